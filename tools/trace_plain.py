"""Timeline of the TMA-fed (x_plain) short-K tiles (diagnostics; needs python zsgnet-pytorch_b200/build.py --trace):
1x1 64->256, bf16 storage + statistics, one K block per tile.  CTA 0 stamps: issuer 8 before full-wait, 9 after, 10 after
token, 11 after issue + commit; drain warp 8: 12 before the acc_full wait, 13 after the TMEM drain, 14 after the stores."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zsg_b200
from zsg_b200 import ops, geometry, _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libzsg_b200_trace.so")
B, cin, cout, H = 128, 64, 256, 75
M = B * H * H
x = torch.randn(M, cin, device="cuda"); w = torch.randn(cout, cin, device="cuda") * 0.05
xi, wi = x.bfloat16(), w.bfloat16()
y = torch.empty(M, cout, device="cuda", dtype=torch.bfloat16)
st = torch.zeros((M + 127) // 128 * 4 * 2 * cout, device="cuda")
rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, 0).cuda()
op = ops.ConvOp(x, w, y, rows, M, cin, cout, 1, 1, w_lo=wi, x_lo=xi, stats=st, x_plain=True, y_pitch=cout)
for _ in range(2): op()
torch.cuda.synchronize()
nblk = 120
buf = torch.zeros(nblk * 16, dtype=torch.int32, device="cuda")
_lib.call("zsg_debug_set_conv_trace", buf.data_ptr(), nblk)
op(); torch.cuda.synchronize()
_lib.call("zsg_debug_set_conv_trace", None, 0)
t = buf.cpu().numpy().astype(np.int64).reshape(nblk, 16) & 0xFFFFFFFF
t0 = t[40, 8]
t = (t - t0) & 0xFFFFFFFF
print("tile | I: top fullOK tokOK issued | D: wait-top drained stored | fullwait issue | drainwait+ld epilogue | issued->drained  period")
prev = None
for g in range(40, 72):
    r = t[g]
    print(f"{g:4d} | {r[8]:7d} {r[9]:7d} {r[10]:7d} {r[11]:7d} | {r[12]:7d} {r[13]:7d} {r[14]:7d} | {r[9]-r[8]:6d} {r[11]-r[10]:5d} | {r[13]-r[12]:6d} {r[14]-r[13]:6d} | {r[13]-r[11]:6d} {(r[14]-prev) if prev is not None else 0:6d}")
    prev = r[14]
