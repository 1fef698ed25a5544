"""Run-to-run determinism of the conv kernel (diagnostics).  ZSG_LIB=<path> selects another build of the library."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import _lib
if os.environ.get("ZSG_LIB"):
    _lib.LIB_PATH = os.environ["ZSG_LIB"]
    _lib.SIGNATURES.pop("zsg_debug_set_conv_trace", None)
from zsg_b200 import ops, geometry
torch.manual_seed(0)
for (B, cin, H, cout, k) in ((4, 64, 12, 128, 3), (16, 256, 19, 256, 3), (8, 1024, 19, 256, 1)):
    x = torch.randn(B, H, H, cin, device="cuda")
    w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    ops.split_tf32(w, hi, lo, w.numel())
    rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
    M = B * H * H
    ys = []
    for i in range(4):
        y = torch.empty(B, H, H, cout, device="cuda")
        ops.ConvOp(x, hi, y, rows, M, cin, cout, k, k, w_lo=lo)()
        ys.append(y)
    torch.cuda.synchronize()
    print((B, cin, H, cout, k), "identical to run 0:", [bool(torch.equal(ys[0], y)) for y in ys[1:]],
          "max diff", max(float((ys[0] - y).abs().max()) for y in ys[1:]))
