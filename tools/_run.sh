timeout 600 python -m pytest tests/test_bf16_kernels_gpu.py -x -q -k "wgrad" 2>&1 | tail -2
timeout 120 python tools/time_big.py bf16 2>&1 | grep wgrad
timeout 200 python tools/step_time.py --batch 128 --dtype bf16 2>&1 | tail -1
timeout 200 python tools/step_time.py --batch 128 --dtype bf16 --model ssd_vgg 2>&1 | tail -1
