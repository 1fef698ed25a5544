import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry
torch.manual_seed(0)
B, cin, H, cout, k = 2, 256, 8, 128, 3
if len(sys.argv) > 1: B, cin, H, cout, k = [int(v) for v in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
x = torch.randn(B, H, H, cin, device="cuda")
w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
hi, lo = torch.empty_like(w), torch.empty_like(w)
ops.split_tf32(w, hi, lo, w.numel())
rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
y = torch.empty(B, H, H, cout, device="cuda")
for _ in range(reps):
    ops.ConvOp(x, hi, y, rows, B * H * H, cin, cout, k, k, w_lo=lo)()
    torch.cuda.synchronize()
print("ok", cin, float(y.abs().mean()))
