"""Where does the end-to-end step lose time against the device-resident one?  Times the module-API loop (bench.py e2e) in
variants: device-resident batch (no H2D), prefetcher, with / without the per-step .item() read-back."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import zsg_b200  # noqa: F401
from zsg_b200 import dat_loader, evaluator, loss, mdl, optim
from zsg_b200.trainer import FusedStep

dtype = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = bench.net_cfg("retina", dtype, 0)
ratios, scales = [0.5, 1, 2], 4 * np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)])
torch.manual_seed(0)
net = mdl.get_default_net(9, cfg)
net.train()
fs = FusedStep(net, ratios, scales, cfg)
crit, evalr = loss.get_default_loss(ratios, scales, cfg), evaluator.get_default_eval(ratios, scales, cfg)
opt = optim.FusedAdam(net.parameters(), lr=1e-4, betas=(0.9, 0.99), net=net)
host = [dat_loader.synthetic_batch(B, seed=i, pin=True) for i in range(4)]
dev = [{k: v.cuda() for k, v in h.items()} for h in host]
for d, h in zip(dev, host):
    d["qlens_cpu"] = h["qlens"]
N = 12


def step(batch, read=True):
    opt.zero_grad()
    out = net(batch)
    ls = crit(out, batch)
    ls["loss"].mean().backward()
    opt.step()
    met = evalr(out, batch)
    if read:
        return float(ls["loss"].item()), float(met["Acc"].item())


def timed(name, it, read=True):
    for b in it(3):
        step(b, read)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b in it(N):
        step(b, read)
    torch.cuda.synchronize()
    print(f"{name:50s} {(time.perf_counter() - t0) / N * 1e3:7.2f} ms/step")


def resident(n):
    for i in range(n):
        yield dev[i % 4]


def prefetched(n):
    return dat_loader.DevicePrefetcher((host[i % 4] for i in range(n)), "cuda")


def prefetched_staged(n):
    return dat_loader.DevicePrefetcher((host[i % 4] for i in range(n)), "cuda", lstm_state=True)


for i in range(3):
    fs.step(dev[i % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(N):
    fs.step(dev[i % 4])
torch.cuda.synchronize()
print(f"{'FusedStep, resident batch':50s} {(time.perf_counter() - t0) / N * 1e3:7.2f} ms/step")
timed("module API, resident batch, no read-back", resident, read=False)
timed("module API, resident batch, .item() per step", resident, read=True)
timed("module API, prefetcher, no read-back", prefetched, read=False)
timed("module API, prefetcher, .item() per step", prefetched, read=True)
timed("module API, prefetcher + staged LSTM state, .item()", prefetched_staged, read=True)
# host time of one step's Python when nothing has to wait: launch everything, measure until the calls return
torch.cuda.synchronize()
t0 = time.perf_counter()
step(dev[0], read=False)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host time to issue one module-API step: {(t1 - t0) * 1e3:.2f} ms")
