import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F, time
import zsg_b200
from zsg_b200 import ops, geometry
torch.backends.cudnn.allow_tf32 = False
B, cin, H, W, cout, k = 2, 256, 12, 12, 256, 3
g = torch.Generator().manual_seed(3)
x = torch.randn(B, cin, H, W, generator=g).cuda()
w = (torch.randn(cout, cin, k, k, generator=g) / 48).cuda().requires_grad_(True)
y = F.conv2d(x, w, None, padding=1)
dy = torch.randn(y.shape, generator=g).cuda()
y.backward(dy)
ref = w.grad.permute(0, 2, 3, 1).contiguous()
rows = geometry.conv_rows(B, H, W, cin, H, W, cout, 1, 1).cuda()
xn, dyn = x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous()
for impl in (1, 0, 7):
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    t0 = time.time()
    try:
        ops.WgradOp(xn, dyn, dw, rows, B * H * W, cin, cout, k, k, impl=impl)()
        torch.cuda.synchronize()
    except Exception as e:
        print("impl", impl, "EXC", repr(e)[:300]); break
    d = (dw - ref)
    print(f"impl {impl}: {time.time()-t0:.2f}s |dw| {float(dw.norm()):.4e} |ref| {float(ref.norm()):.4e} relerr {float(d.norm()/ref.norm()):.3e} nan {int(torch.isnan(dw).sum())} zeros {int((dw==0).sum())}/{dw.numel()}")
    if impl != 1:
        # which output rows/cols are right?
        e = d.view(cout, -1).abs()
        print("   per-n err (first 8):", [f"{float(v):.2e}" for v in e.max(1)[0][:8]], " per-j err (first 8):", [f"{float(v):.2e}" for v in e.max(0)[0][:8]])
