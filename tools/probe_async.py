"""cp.async conv path (pre-split input) vs the register path: equality and speed (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry

def run(B, cin, H, cout, k, pro=False, stride=1, reps=5):
    torch.manual_seed(0)
    x = torch.randn(B, H, H, cin, device="cuda")
    w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    ops.split_tf32(w, hi, lo, w.numel())
    Ho = (H + 2 * (k // 2) - k) // stride + 1
    rows = geometry.conv_rows(B, H, H, cin, Ho, Ho, cout, stride, k // 2).cuda()
    M = B * Ho * Ho
    sc = torch.rand(cin, device="cuda") + 0.5 if pro else None
    sh = torch.randn(cin, device="cuda") if pro else None
    y1 = torch.empty(B, Ho, Ho, cout, device="cuda"); y2 = torch.empty_like(y1)
    op1 = ops.ConvOp(x, hi, y1, rows, M, cin, cout, k, k, w_lo=lo, in_scale=sc, in_shift=sh, in_relu=pro)
    z = torch.empty_like(x) if pro else None
    x_lo = torch.empty_like(x)
    def split():
        ops.split_act(x, x_lo, B * H * H, cin, scale=sc, shift=sh, relu=pro, z=z)
    op2 = ops.ConvOp(z if pro else x, hi, y2, rows, M, cin, cout, k, k, w_lo=lo, x_lo=x_lo)
    split(); op1(); op2()
    torch.cuda.synchronize()
    d = float((y1 - y2).abs().max())
    fl = 2.0 * M * cout * k * k * cin
    def timeit(f):
        for _ in range(2): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): f()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    t1, t2, ts = timeit(op1), timeit(op2), timeit(split)
    print(f"B={B} cin={cin} H={H} cout={cout} k={k} s={stride} pro={int(pro)}: max|diff|={d:.2e} (y rms {float(y1.pow(2).mean().sqrt()):.3f})  "
          f"register {t1:.3f} ms {fl/t1/1e9:6.1f} TF/s | async {t2:.3f} ms {fl/t2/1e9:6.1f} TF/s | split {ts:.3f} ms "
          f"({(x.numel() * (12 if pro else 8)) / ts / 1e6:.0f} GB/s)", flush=True)

run(2, 64, 12, 128, 3)
run(2, 64, 12, 128, 3, pro=True)
run(3, 128, 20, 128, 3, stride=2)
run(64, 256, 44, 256, 3)
run(64, 256, 44, 256, 3, pro=True)
run(64, 64, 75, 256, 1)
run(64, 64, 75, 64, 3, pro=True)
run(64, 1024, 19, 256, 1)
run(64, 256, 19, 1024, 1, pro=True)
run(64, 512, 10, 512, 3, pro=True)
run(64, 128, 38, 128, 3, pro=True)
