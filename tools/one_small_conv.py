"""One short-K launch for ncu: 1x1 64->256 over M = B*75*75 rows, bf16 storage + BatchNorm statistics, TMA-fed A tiles.
python tools/one_small_conv.py [bf16|fp32] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry
dt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (128 if dt == "bf16" else 64)
cin, cout, H = 64, 256, 75
M = B * H * H
x = torch.randn(M, cin, device="cuda"); w = torch.randn(cout, cin, device="cuda") * 0.05
if dt == "bf16":
    xi, wi, wa = x.bfloat16(), w.bfloat16(), w
    y = torch.empty(M, cout, device="cuda", dtype=torch.bfloat16)
else:
    xi, wa, wi = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
    ops.split_act(x, xi, M, cin); ops.split_tf32(w, wa, wi, w.numel())
    y = torch.empty(M, cout, device="cuda")
st = torch.zeros((M + 127) // 128 * 4 * 2 * cout, device="cuda")
rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, 0).cuda()
op = ops.ConvOp(x, wa, y, rows, M, cin, cout, 1, 1, w_lo=wi, x_lo=xi, stats=st, x_plain=True, y_pitch=cout)
for _ in range(4): op()
torch.cuda.synchronize()
print("done")
