"""One big conv launch (for ncu): 3x3 256->256, M=123904, optional BN+ReLU prologue."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry
B, cin, H, cout, k = 64, 256, 44, 256, 3
pro = len(sys.argv) > 1 and sys.argv[1] == "pro"
x = torch.randn(B, H, H, cin, device="cuda")
w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
hi, lo = torch.empty_like(w), torch.empty_like(w)
ops.split_tf32(w, hi, lo, w.numel())
rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
y = torch.empty(B, H, H, cout, device="cuda")
sc = torch.rand(cin, device="cuda") + 0.5 if pro else None
sh = torch.randn(cin, device="cuda") if pro else None
if "async" in sys.argv:
    x_lo = torch.empty_like(x)
    ops.split_act(x, x_lo, B * H * H, cin)
    op = ops.ConvOp(x, hi, y, rows, B * H * H, cin, cout, k, k, w_lo=lo, x_lo=x_lo)
    dy = torch.randn(B, H, H, cout, device="cuda"); dy_lo = torch.empty_like(dy)
    ops.split_act(dy, dy_lo, B * H * H, cout)
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    wop = ops.WgradOp(x, dy, dw, rows, B * H * H, cin, cout, k, k, x_lo=x_lo, dy_lo=dy_lo)
    for _ in range(3): op(); wop()
else:
    op = ops.ConvOp(x, hi, y, rows, B * H * H, cin, cout, k, k, w_lo=lo, in_scale=sc, in_shift=sh, in_relu=pro)
    for _ in range(3): op()
torch.cuda.synchronize()
