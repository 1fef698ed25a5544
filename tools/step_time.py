"""Device time of the fused training step (diagnostics): python tools/step_time.py [--model M] [--batch B] [--steps N] [--dtype fp32|bf16]
[--no-lstm-overlap] [--no-wgrad-overlap]"""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zsg_b200
from zsg_b200 import mdl, dat_loader
from zsg_b200.trainer import FusedStep

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="retina")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
ap.add_argument("--no-lstm-overlap", action="store_true")
ap.add_argument("--no-wgrad-overlap", action="store_true")
a = ap.parse_args()
cfg = {"do_norm": False, "use_same_atb": True, "mdl_to_use": a.model, "resize_img": [300, 300], "use_multi": True,
       "use_focal": True, "use_softmax": False, "alpha": 0.25, "gamma": 2, "emb_dim": 300, "matching_threshold": 0.6,
       "use_bidirectional": True, "lstm_dim": 128, "lamb_reg": 1, "acc_iou_threshold": 0.5, "use_lang": True,
       "use_img": True, "device": "cuda:0", "zsg_dtype": a.dtype, "zsg_quiet": True}
torch.manual_seed(0)
net = mdl.get_default_net(9, cfg); net.train()
fs = FusedStep(net, [0.5, 1, 2], 4 * np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)]), cfg)
B = a.batch
batches = []
for i in range(2):
    b = {k: v.cuda() for k, v in dat_loader.synthetic_batch(B, i).items()}
    b["qlens_cpu"] = b["qlens"].cpu()
    batches.append(b)
eng = net.engine_for(B, 20)
eng.overlap_lstm = eng.overlap_lstm and not a.no_lstm_overlap
eng.overlap_wgrad = eng.overlap_wgrad and not a.no_wgrad_overlap
for i in range(4):
    out = fs.step(batches[i % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    out = fs.step(batches[i % 2])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"{a.model} B={B} lstm_overlap={eng.overlap_lstm} wgrad_overlap={eng.overlap_wgrad}: {ms:.2f} ms/step  {B / ms * 1e3:.1f} pairs/s  "
      f"loss {float(out['loss']):.4f}  engine buffers {eng.nbytes / 2**30:.1f} GiB  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
