"""One training step inside cudaProfilerStart/Stop, for `ncu --profile-from-start off` (launch lists and full captures of
exactly one step; eager launches, ZSG_GRAPHS=0 is set here).  Usage: python tools/profile_step.py [fp32|bf16] [batch] [model]"""
import os
import sys

os.environ["ZSG_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import zsg_b200  # noqa: F401
from zsg_b200 import dat_loader, mdl
from zsg_b200.trainer import FusedStep

dtype = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
model = sys.argv[3] if len(sys.argv) > 3 else "retina"
cfg = bench.net_cfg(model, dtype, 0)
torch.manual_seed(0)
net = mdl.get_default_net(9, cfg)
net.train()
fs = FusedStep(net, [0.5, 1, 2], 4 * np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)]), cfg)
batch = {k: v.cuda() for k, v in dat_loader.synthetic_batch(B, 1).items()}
batch["qlens_cpu"] = batch["qlens"].cpu()
for _ in range(3):
    fs.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
fs.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step:", dtype, B, model)
