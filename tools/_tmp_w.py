import sys, os
sys.path.insert(0, "/root/repo")
import torch
import zsg_b200
from zsg_b200 import ops, geometry
def run(B, cin, H, cout, k):
    torch.manual_seed(0)
    x = torch.randn(B, H, H, cin, device="cuda"); dy = torch.randn(B, H, H, cout, device="cuda")
    rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
    M = B * H * H
    x_lo, dy_lo = torch.empty_like(x), torch.empty_like(dy)
    ops.split_act(x, x_lo, M, cin); ops.split_act(dy, dy_lo, M, cout)
    dw1 = torch.zeros(cout, k, k, cin, device="cuda"); dw2 = torch.zeros_like(dw1)
    ops.WgradOp(x, dy, dw1, rows, M, cin, cout, k, k)()
    torch.cuda.synchronize()
    ops.WgradOp(x, dy, dw2, rows, M, cin, cout, k, k, x_lo=x_lo, dy_lo=dy_lo)()
    torch.cuda.synchronize()
    print(B, cin, H, cout, k, "max rel diff", float((dw1 - dw2).abs().max() / dw1.abs().max()), flush=True)
import subprocess
if len(sys.argv) > 1:
    run(*[int(v) for v in sys.argv[1:6]])
else:
    for a in ["2 512 10 256 3", "2 264 10 256 3", "2 520 10 256 3", "2 520 10 256 1", "2 136 10 256 3", "2 520 10 128 3"]:
        r = subprocess.run([sys.executable, __file__] + a.split(), capture_output=True, text=True)
        out = [l for l in (r.stdout + r.stderr).split("\n") if "max rel" in l or "timed out" in l]
        print(a, "->", out[:2], flush=True)
