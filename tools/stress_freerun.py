"""Stress the conv kernels with the issuers free-running (impl = 16) against the token-ordered result (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
shapes = [(4, 256, 8, 256, 3), (16, 256, 19, 256, 3), (16, 256, 19, 128, 3), (8, 1024, 19, 256, 1), (1, 520, 10, 256, 3),
          (64, 64, 38, 64, 3), (5, 300, 4, 512, 1), (64, 256, 44, 256, 3)]
for (B, cin, H, cout, k) in shapes:
    x = torch.randn(B, H, H, cin, device="cuda")
    w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    ops.split_tf32(w, hi, lo, w.numel())
    x_lo = torch.empty_like(x)
    ops.split_act(x, x_lo, B * H * H, cin)
    rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
    M = B * H * H
    ref = torch.empty(B, H, H, cout, device="cuda")
    ops.ConvOp(x, hi, ref, rows, M, cin, cout, k, k, w_lo=lo)()
    torch.cuda.synchronize()
    bad = 0
    for rep in range(30):
        for kw in (dict(), dict(x_lo=x_lo)):
            y = torch.full_like(ref, float("nan"))
            ops.ConvOp(x, hi, y, rows, M, cin, cout, k, k, w_lo=lo, impl=impl, **kw)()
            torch.cuda.synchronize()
            bad += int(not torch.equal(y, ref))
    print((B, cin, H, cout, k), "mismatches:", bad, "of 60", flush=True)
print("done")
