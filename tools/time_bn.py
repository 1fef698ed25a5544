"""Device time of the fp32 BatchNorm passes alone at the layer1 / layer3 sizes (diagnostics): python tools/time_bn.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for rows, c in ((360000, 256), (360000, 64), (23104, 1024)):
    x, r, dy = (torch.randn(rows, c, device="cuda") for _ in range(3))
    y, ylo, dx, dxlo, dz = (torch.empty(rows, c, device="cuda") for _ in range(5))
    sc, sh = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    mean, invstd, gamma = x.mean(0), 1 / x.std(0), torch.rand(c, device="cuda") + 0.5
    sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    dg, db = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    n = rows * c * 4 / 1e9          # GB per fp32 tensor
    t = timeit(lambda: ops.bn_apply(x, sc, sh, y, rows, c, True, r=r, y_lo=ylo)); print(f"[{rows}x{c}] tail      {t*1e3:7.1f} us {4*n/(t*1e-3):6.0f} GB/s")
    t = timeit(lambda: ops.split_act(x, ylo, rows, c, scale=sc, shift=sh, relu=True, z=y)); print(f"[{rows}x{c}] split_act {t*1e3:7.1f} us {3*n/(t*1e-3):6.0f} GB/s")
    t = timeit(lambda: ops.bn_bwd_reduce(dy, x, mean, invstd, sums, rows, c, mask_mode=1, scale=sc, shift=sh)); print(f"[{rows}x{c}] reduce m1 {t*1e3:7.1f} us {2*n/(t*1e-3):6.0f} GB/s")
    t = timeit(lambda: ops.bn_bwd_reduce(dy, x, mean, invstd, sums, rows, c, mask_mode=2, act_out=y, dz_out=dz)); print(f"[{rows}x{c}] reduce m2 {t*1e3:7.1f} us {4*n/(t*1e-3):6.0f} GB/s")
    t = timeit(lambda: ops.bn_bwd_apply(dy, x, mean, invstd, gamma, sums, dx, dg, db, rows, c, mask_mode=1, scale=sc, shift=sh, dx_lo=dxlo)); print(f"[{rows}x{c}] apply m1  {t*1e3:7.1f} us {4*n/(t*1e-3):6.0f} GB/s")
