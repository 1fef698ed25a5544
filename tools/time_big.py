"""Device time of the large 3x3 256->256 layer, forward + weight gradient (bf16 / fp32 operand-image paths): python tools/time_big.py [bf16|fp32] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, geometry
dt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (128 if dt == "bf16" else 64)
for cin, H, cout, k in ((256, 44, 256, 3), (256, 19, 1024, 1), (1024, 19, 256, 1), (512, 10, 512, 3)):
    M = B * H * H
    x = torch.randn(B, H, H, cin, device="cuda"); w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
    dy = torch.randn(B, H, H, cout, device="cuda"); bias = torch.randn(cout, device="cuda")
    rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
    y = torch.empty(B, H, H, cout, device="cuda"); dw = torch.zeros(cout, k, k, cin, device="cuda")
    if dt == "bf16":
        xi, wi, dyi, wa = x.bfloat16(), w.bfloat16(), dy.bfloat16(), w
    else:
        xi, wa, wi, dyi = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w), torch.empty_like(dy)
        ops.split_act(x, xi, M, cin); ops.split_tf32(w, wa, wi, w.numel()); ops.split_act(dy, dyi, M, cout)
    op = ops.ConvOp(x, wa, y, rows, M, cin, cout, k, k, w_lo=wi, x_lo=xi, bias=bias, out_relu=True, y_pitch=cout, x_plain=(k == 1))
    wop = ops.WgradOp(x, dy, dw, rows, M, cin, cout, k, k, x_lo=xi, dy_lo=dyi, dy_pitch=cout)
    fl = 2.0 * M * cout * k * k * cin
    for name, f in (("fwd", op), ("wgrad", wop)):
        for _ in range(3): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): f()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"{dt} {k}x{k} {cin}->{cout} M={M} {name:5s} {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s", flush=True)
