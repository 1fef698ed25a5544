"""Timing ablations of the cp.async conv kernel on store-heavy 1x1 shapes (diagnostics; outputs are garbage).
impl = 8 + bits: 2 = drain warps skip tcgen05.ld, 16 = epilogue skips the global stores."""
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
    x_lo = torch.empty_like(x)
    ops.split_act(x, x_lo, M, cin)
    byts = 4.0 * M * (cout + 2 * cin)
    names = {0: "normal", 24: "no stores", 10: "no tmem drain", 26: "no drain, no stores"}
    for impl in names:
        op = ops.ConvOp(x, hi, y, rows, M, cin, cout, k, k, impl=impl, w_lo=lo, x_lo=x_lo)
        for _ in range(2): op()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): op()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{label:28s} {names[impl]:24s} {ms:7.3f} ms  {byts/ms/1e6:7.0f} GB/s", flush=True)
    # memset-like reference: how fast can this GPU write y at all
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): y.fill_(1.0)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{label:28s} {'torch fill_ of y':24s} {ms:7.3f} ms  {4.0*M*cout/ms/1e6:7.0f} GB/s", flush=True)

if len(sys.argv) > 1:
    run(64, 64, 75, 64, 1, "1x1 64->64 M=360000")
    run(64, 64, 75, 64, 3, "3x3 64->64 M=360000")
else:
    run(64, 64, 75, 256, 1, "1x1 64->256 M=360000")
    run(64, 128, 38, 512, 1, "1x1 128->512 M=92416")
    run(64, 256, 75, 64, 1, "1x1 256->64 M=360000")
