"""One big conv launch pair (for ncu --set full): 3x3 256->256 forward + its weight gradient, operand-image path.
python tools/one_conv.py [fp32|bf16] [B]      (M = B * 44 * 44)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry
dt = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (128 if dt == "bf16" else 64)
cin, H, cout, k = 256, 44, 256, 3
M = B * H * H
x = torch.randn(B, H, H, cin, device="cuda")
w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
dy = torch.randn(B, H, H, cout, device="cuda")
bias = torch.randn(cout, device="cuda")
rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
y = torch.empty(B, H, H, cout, device="cuda")
dw = torch.zeros(cout, k, k, cin, device="cuda")
if dt == "bf16":
    xi, wi, dyi, wa = x.bfloat16(), w.bfloat16(), dy.bfloat16(), w
else:
    xi, wa, wi, dyi = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w), torch.empty_like(dy)
    ops.split_act(x, xi, M, cin)
    ops.split_tf32(w, wa, wi, w.numel())
    ops.split_act(dy, dyi, M, cout)
op = ops.ConvOp(x, wa, y, rows, M, cin, cout, k, k, w_lo=wi, x_lo=xi, bias=bias, out_relu=True, y_pitch=cout)   # a head layer
wop = ops.WgradOp(x, dy, dw, rows, M, cin, cout, k, k, x_lo=xi, dy_lo=dyi, dy_pitch=cout)
for _ in range(3):
    op(); wop()
torch.cuda.synchronize()
print("done", dt, B, "flop per launch", 2.0 * M * cout * k * k * cin)
