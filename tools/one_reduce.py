"""One BatchNorm-backward reduce launch at the layer1 block-output size (for ncu): rows = 360000, C = 256, mask mode 2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops
rows, C = 64 * 75 * 75, 256
x, dy, act = (torch.randn(rows, C, device="cuda") for _ in range(3))
dz = torch.empty_like(x)
mean, invstd, scale, shift = (torch.rand(C, device="cuda") + 0.5 for _ in range(4))
sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
for mm in (2, 1, 2, 1):
    ops.bn_bwd_reduce(dy, x, mean, invstd, sums, rows, C, mask_mode=mm, scale=scale, shift=shift, act_out=act if mm == 2 else None,
                      dz_out=dz if mm == 2 else None)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mm in (2, 1):
    a.record()
    for _ in range(5):
        ops.bn_bwd_reduce(dy, x, mean, invstd, sums, rows, C, mask_mode=mm, scale=scale, shift=shift,
                          act_out=act if mm == 2 else None, dz_out=dz if mm == 2 else None)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    byts = rows * C * 4 * (4 if mm == 2 else 2)
    print(f"mask_mode {mm}: {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s")
