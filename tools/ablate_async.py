"""Timing ablations of conv_tc_async_kernel (diagnostics; results are garbage when a part is switched off).
impl = 8 + bits: 1 = no producers, 2 = no TMEM drain, 16 = no stores, 32 = no input (A) loads, 64 = no weight (B) loads.
python tools/ablate_async.py [bf16|fp32] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry

dt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128


def run(B, cin, H, cout, k, label):
    x = torch.randn(B, H, H, cin, device="cuda")
    w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
    if dt == "bf16":
        xi, wi, wa = x.bfloat16(), w.bfloat16(), w
    else:
        xi, wa, wi = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
        ops.split_act(x, xi, B * H * H, cin)
        ops.split_tf32(w, wa, wi, w.numel())
    rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
    y = torch.empty(B, H, H, cout, device="cuda")
    M = B * H * H
    fl = 2.0 * M * cout * k * k * cin
    names = {0: "normal", 8 + 32: "no A loads", 8 + 64: "no B loads", 8 + 96: "no A, no B loads", 8 + 16: "no stores",
             8 + 2: "no drain", 8 + 1: "no producers"}
    for impl, name in names.items():
        op = ops.ConvOp(x, wa, y, rows, M, cin, cout, k, k, impl=impl, w_lo=wi, x_lo=xi)
        for _ in range(2): op()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): op()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{dt} {label:28s} {name:22s} {ms:7.3f} ms  {fl/ms/1e9:7.1f} TF/s", flush=True)


run(B, 256, 44, 256, 3, f"3x3 256->256 M={B*44*44}")
run(B, 1024, 19, 256, 1, f"1x1 1024->256 M={B*361}")
run(B, 256, 19, 1024, 1, f"1x1 256->1024 M={B*361}")
