timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_bf16_kernels_gpu.py -x -q -k "conv" > gpurun_out/t_conv.log 2>&1; tail -2 gpurun_out/t_conv.log
timeout 120 python tools/time_big.py bf16 2>&1 | head -2
timeout 120 python tools/time_big.py fp32 2>&1 | head -2
timeout 200 python tools/step_time.py --batch 64 2>&1 | tail -1
