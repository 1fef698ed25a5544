"""A/B of the bf16 engine with and without bf16 storage of the trunk activations (ZSG_B16_ACT): python tools/ab_b16act.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zsg_b200
from zsg_b200 import mdl, dat_loader
from zsg_b200.trainer import FusedStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for act in ("1", "0"):
    os.environ["ZSG_B16_ACT"] = act
    cfg = {"do_norm": False, "use_same_atb": True, "mdl_to_use": "retina", "resize_img": [300, 300], "use_multi": True,
           "use_focal": True, "use_softmax": False, "alpha": 0.25, "gamma": 2, "emb_dim": 300, "matching_threshold": 0.6,
           "use_bidirectional": True, "lstm_dim": 128, "lamb_reg": 1, "acc_iou_threshold": 0.5, "use_lang": True,
           "use_img": True, "device": "cuda:0", "zsg_dtype": "bf16", "zsg_quiet": True}
    torch.manual_seed(0)
    net = mdl.get_default_net(9, cfg); net.train()
    fs = FusedStep(net, [0.5, 1, 2], 4 * np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)]), cfg)
    bs = []
    for i in range(2):
        b = {k: v.cuda() for k, v in dat_loader.synthetic_batch(B, i).items()}; b["qlens_cpu"] = b["qlens"].cpu(); bs.append(b)
    for i in range(4): out = fs.step(bs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): out = fs.step(bs[i % 2])
    e1.record(); torch.cuda.synchronize()
    eng = net.engine_for(B, 20)
    print(f"B16_ACT={act} B={B}: {e0.elapsed_time(e1) / 10:.2f} ms/step  loss {float(out['loss']):.4f}  engine buffers {eng.nbytes / 2**30:.1f} GiB")
    del net, fs, eng; torch.cuda.empty_cache()
