"""Stand-alone timing of the implicit-GEMM kernels on the layer shapes of the bs=64 step, fp32 (3xTF32) vs bf16
operand images: forward and weight gradient, CUDA events over `reps` back-to-back launches after warm-up (inputs of one
launch exceed L2 for the large-M shapes; the small ones are L2-resident in the real step too)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zsg_b200  # noqa: F401
from zsg_b200 import geometry as geo, ops

# (B, cin, H, cout, k, stride, pad)
SHAPES = [(64, 256, 38, 256, 3, 1, 1), (64, 64, 75, 256, 1, 1, 0), (64, 256, 75, 64, 1, 1, 0), (64, 64, 75, 64, 3, 1, 1),
          (64, 128, 38, 128, 3, 1, 1), (64, 128, 38, 512, 1, 1, 0), (64, 256, 19, 256, 3, 1, 1), (64, 256, 19, 1024, 1, 1, 0),
          (64, 1024, 19, 256, 1, 1, 0), (64, 512, 10, 512, 3, 1, 1), (64, 512, 10, 2048, 1, 1, 0), (64, 2048, 10, 256, 3, 2, 1)]
reps = 10


def timed(op):
    for _ in range(3):
        op()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        op()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print(f"{'shape':42s} {'fwd f32':>9s} {'fwd bf16':>9s} {'wg f32':>9s} {'wg bf16':>9s}   (TFLOP/s algorithmic; ms bf16 fwd / wgrad)")
for B, cin, H, cout, k, stride, pad in SHAPES:
    Ho = (H + 2 * pad - k) // stride + 1
    m = B * Ho * Ho
    x = torch.randn(B * H * H, cin, device="cuda")
    w = torch.randn(cout, k, k, cin, device="cuda") / (cin * k * k) ** 0.5
    dy = torch.randn(m, cout, device="cuda")
    y = torch.empty(m, cout, device="cuda")
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    rows = geo.conv_rows(B, H, H, cin, Ho, Ho, cout, stride, pad).cuda()
    x_lo, w_hi, w_lo, dy_lo = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w), torch.empty_like(dy)
    ops.split_act(x, x_lo, x.shape[0], cin)
    ops.split_act(dy, dy_lo, m, cout)
    ops.split_tf32(w, w_hi, w_lo, w.numel())
    xb, wb, dyb = (torch.empty(t.shape, dtype=torch.bfloat16, device="cuda") for t in (x, w, dy))
    ops.cast_bf16(x, xb, x.numel()); ops.cast_bf16(w, wb, w.numel()); ops.cast_bf16(dy, dyb, dy.numel())
    fl = 2.0 * m * cout * k * k * cin
    f32 = timed(ops.ConvOp(x, w_hi, y, rows, m, cin, cout, k, k, w_lo=w_lo, x_lo=x_lo))
    b16 = timed(ops.ConvOp(x, w, y, rows, m, cin, cout, k, k, w_lo=wb, x_lo=xb))
    wf32 = timed(ops.WgradOp(x, dy, dw, rows, m, cin, cout, k, k, x_lo=x_lo, dy_lo=dy_lo, dy_pitch=cout))
    wb16 = timed(ops.WgradOp(x, dy, dw, rows, m, cin, cout, k, k, x_lo=xb, dy_lo=dyb, dy_pitch=cout))
    tf = lambda ms: fl / ms / 1e9
    print(f"{str((B, cin, H, cout, k, stride)):42s} {tf(f32):9.1f} {tf(b16):9.1f} {tf(wf32):9.1f} {tf(wb16):9.1f}   {b16:.3f} / {wb16:.3f}")
