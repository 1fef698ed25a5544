"""Timing ablations of the short-K (epilogue-bound) 1x1 convs of layer1 (diagnostics; garbage results when a part is off).
impl = 8 + bits: 2 = no TMEM drain, 16 = no output stores, 128 = no BatchNorm statistics, 1 = no producers.
python tools/ablate_epilogue.py [bf16|fp32] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry

dt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128


def run(cin, cout, H, label, residual=False):
    M = B * H * H
    x = torch.randn(M, cin, device="cuda")
    w = torch.randn(cout, cin, device="cuda") * 0.05
    if dt == "bf16":
        xi, wi, wa = x.bfloat16(), w.bfloat16(), w
        y = torch.empty(M, cout, device="cuda", dtype=torch.float32 if residual else torch.bfloat16)
        res = torch.randn(M, cout, device="cuda").bfloat16() if residual else None
    else:
        xi, wa, wi = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
        ops.split_act(x, xi, M, cin)
        ops.split_tf32(w, wa, wi, w.numel())
        y = torch.empty(M, cout, device="cuda")
        res = torch.randn(M, cout, device="cuda") if residual else None
    rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, 0).cuda()
    st = None if residual else torch.zeros((M + 127) // 128 * 4 * 2 * cout, device="cuda")
    byt = M * cin * xi.element_size() * (1 if dt == "bf16" else 2) + M * cout * y.element_size() + (res.numel() * res.element_size() if residual else 0)
    names = {0: "normal", 8 + 128: "no stats", 8 + 16: "no stores", 8 + 144: "no stats, no stores", 8 + 2: "no drain", 8 + 146: "no stats/stores/drain", 8 + 1: "no producers"}
    for impl, name in names.items():
        op = ops.ConvOp(x, wa, y, rows, M, cin, cout, 1, 1, impl=impl, w_lo=wi, x_lo=xi, stats=st, residual=res, x_plain=True, y_pitch=cout)
        for _ in range(2): op()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): op()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        tiles = ((M + 127) // 128) * ((cout + 127) // 128 if cout > 64 else 1)
        print(f"{dt} {label:30s} {name:24s} {ms:7.3f} ms  {byt/ms/1e6:7.0f} GB/s  {ms*1e-3*1.75e9/(tiles/148):7.0f} cyc/tile", flush=True)


run(64, 256, 75, f"1x1 64->256 M={B*5625} +stats")
run(256, 64, 75, f"1x1 256->64 M={B*5625} +stats")
run(64, 256, 75, f"1x1 64->256 M={B*5625} +residual", residual=True)
