"""Timing ablations of conv_tc_kernel (diagnostics; results are garbage when a part is switched off).
impl = 8 + bits: 1 = no producers (MMA warp does not wait for `full`), 2 = drain warps skip tcgen05.ld, 4 = MMA warp
does not wait for acc_empty."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry

def run(B, cin, H, cout, k, label):
    x = torch.randn(B, H, H, cin, device="cuda")
    w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    ops.split_tf32(w, hi, lo, w.numel())
    rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
    y = torch.empty(B, H, H, cout, device="cuda")
    M = B * H * H
    fl = 2.0 * M * cout * k * k * cin
    names = {0: "normal", 9: "no producers", 11: "no producers, no drain", 10: "no drain"}
    for impl in names:
        op = ops.ConvOp(x, hi, y, rows, M, cin, cout, k, k, impl=impl, w_lo=lo)
        for _ in range(2): op()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): op()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{label:28s} {names[impl]:30s} {ms:7.3f} ms  {fl/ms/1e9:7.1f} TF/s", flush=True)

run(64, 256, 44, 256, 3, "3x3 256->256 M=123904")
run(64, 64, 75, 256, 1, "1x1 64->256 M=360000")
run(64, 1024, 19, 256, 1, "1x1 1024->256 M=23104")
