"""One store-heavy 1x1 conv launch of the cp.async kernel (for ncu): 64->256 at 75x75, M=360000."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry
B, cin, H, cout, k = 64, 64, 75, 256, 1
x = torch.randn(B, H, H, cin, device="cuda")
w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
hi, lo = torch.empty_like(w), torch.empty_like(w)
ops.split_tf32(w, hi, lo, w.numel())
rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, 0).cuda()
y = torch.empty(B, H, H, cout, device="cuda")
x_lo = torch.empty_like(x)
ops.split_act(x, x_lo, B * H * H, cin)
op = ops.ConvOp(x, hi, y, rows, B * H * H, cin, cout, k, k, w_lo=lo, x_lo=x_lo)
for _ in range(3): op()
torch.cuda.synchronize()
