"""cp.async wgrad path (pre-split operands) vs the register path: agreement and speed (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry

def run(B, cin, H, cout, k, stride=1, pitch=None, reps=5):
    torch.manual_seed(0)
    x = torch.randn(B, H, H, cin, device="cuda")
    Ho = (H + 2 * (k // 2) - k) // stride + 1
    pitch = pitch or cout
    dy = torch.randn(B, Ho, Ho, pitch, device="cuda")
    if pitch != cout: dy[..., cout:] = 0
    rows = geometry.conv_rows(B, H, H, cin, Ho, Ho, pitch, stride, k // 2).cuda()
    M = B * Ho * Ho
    x_lo, dy_lo = torch.empty_like(x), torch.empty_like(dy)
    ops.split_act(x, x_lo, B * H * H, cin)
    ops.split_act(dy, dy_lo, M, pitch)
    dw1 = torch.zeros(cout, k, k, cin, device="cuda"); dw2 = torch.zeros_like(dw1)
    op1 = ops.WgradOp(x, dy, dw1, rows, M, cin, cout, k, k)
    op2 = ops.WgradOp(x, dy, dw2, rows, M, cin, cout, k, k, x_lo=x_lo, dy_lo=dy_lo)
    dw3 = torch.zeros_like(dw1)
    op3 = ops.WgradOp(x, dy, dw3, rows, M, cin, cout, k, k, x_lo=x_lo, dy_lo=dy_lo, dy_pitch=pitch)
    op1(); op2(); op3()
    torch.cuda.synchronize()
    d12 = float((dw1 - dw2).abs().max() / dw1.abs().max()); d13 = float((dw1 - dw3).abs().max() / dw1.abs().max())
    ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).double(), (cout, cin, k, k), dy[..., :cout].permute(0, 3, 1, 2).double(),
                                      stride=stride, padding=k // 2).permute(0, 2, 3, 1) if M * cout * cin * k * k < 3e12 else None
    e = lambda a: float((a.double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()) if ref is not None else float("nan")
    fl = 2.0 * M * cout * k * k * cin
    def timeit(f, dw):
        for _ in range(2): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): f()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    t1, t2, t3 = timeit(op1, dw1), timeit(op2, dw2), timeit(op3, dw3)
    print(f"B={B} cin={cin} H={H} cout={cout} k={k} s={stride}: max rel diff vs register: cp.async {d12:.1e} tma-dy {d13:.1e} | "
          f"register {t1:.3f} ms {fl/t1/1e9:6.1f} TF/s | cp.async {t2:.3f} ms {fl/t2/1e9:6.1f} | tma-dy {t3:.3f} ms {fl/t3/1e9:6.1f} TF/s", flush=True)

run(2, 64, 19, 256, 3)
run(3, 128, 20, 128, 3, stride=2)
run(2, 4, 61, 64, 7, stride=2)
run(2, 256, 5, 45, 3, pitch=48)
run(64, 256, 44, 256, 3)
run(64, 64, 75, 64, 3)
run(64, 64, 75, 256, 1)
run(64, 1024, 19, 256, 1)
run(64, 256, 19, 1024, 1)
run(64, 512, 10, 512, 3)
run(64, 128, 38, 128, 3)
run(64, 520, 44, 256, 3)
