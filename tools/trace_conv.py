"""Pipeline timeline of conv_tc_kernel (diagnostics): CTA 0 stamps clock64 at the hand-overs of its first K blocks.
events: 0 producer before empty-wait, 1 after empty-wait, 2 after the STS loop, 3 after arrive(full);
        4 producer tile top, 5 first group barrier passed, 6 row table in smem (cp.async kernel only);
        8 issuer before full-wait, 9 after full-wait, 10 after token-wait, 11 after issue+commit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zsg_b200
from zsg_b200 import ops, geometry, _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libzsg_b200_trace.so")   # python zsgnet-pytorch_b200/build.py --trace

def run(B, cin, H, cout, k, label, nblk=120, pro=False, use_async=False, impl=0, bf16=False):
    x = torch.randn(B, H, H, cin, device="cuda")
    w = torch.randn(cout, k, k, cin, device="cuda") * 0.05
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    ops.split_tf32(w, hi, lo, w.numel())
    rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
    y = torch.empty(B, H, H, cout, device="cuda")
    M = B * H * H
    sc = torch.rand(cin, device="cuda") + 0.5 if pro else None
    sh = torch.randn(cin, device="cuda") if pro else None
    if bf16:
        op = ops.ConvOp(x, w, y, rows, M, cin, cout, k, k, w_lo=w.bfloat16(), x_lo=x.bfloat16(), impl=impl)
    elif use_async:
        x_lo = torch.empty_like(x)
        ops.split_act(x, x_lo, M, cin)
        op = ops.ConvOp(x, hi, y, rows, M, cin, cout, k, k, w_lo=lo, x_lo=x_lo, impl=impl)
    else:
        op = ops.ConvOp(x, hi, y, rows, M, cin, cout, k, k, w_lo=lo, in_scale=sc, in_shift=sh, in_relu=pro)
    for _ in range(2): op()
    torch.cuda.synchronize()
    buf = torch.zeros(nblk * 16, dtype=torch.int32, device="cuda")
    _lib.call("zsg_debug_set_conv_trace", buf.data_ptr(), nblk)
    op()
    torch.cuda.synchronize()
    _lib.call("zsg_debug_set_conv_trace", None, 0)
    t = buf.cpu().numpy().astype(np.int64).reshape(nblk, 16) & 0xFFFFFFFF
    t0 = t[0, 0]
    t = (t - t0) & 0xFFFFFFFF
    print(f"--- {label}: cycles relative to the first producer stamp; gk | P:top emptyOK stsDone arrived | I:top fullOK tokOK issued | issue-to-issue")
    prev = None
    for g in range(40, min(nblk, 76)):
        r = t[g]
        d = (r[11] - prev) if prev is not None else 0
        prev = r[11]
        print(f"{g:4d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} | {r[8]:7d} {r[9]:7d} {r[10]:7d} {r[11]:7d} | {d:5d}   emptywait {r[1]-r[0]:5d} sts {r[2]-r[1]:5d}  fullwait {r[9]-r[8]:5d} tok {r[10]-r[9]:4d} issue {r[11]-r[10]:4d}  full->arrive lag {r[9]-r[3]:5d}")
    if t[40:64, 4].any():
        print("producer per-tile prologue (cp.async kernel): gk | tile top | first barrier, row table in smem, tap offsets | "
              "previous arrive of this group -> this top")
        for g in range(42, min(nblk, 64)):
            r = t[g]
            if r[4] == 0: continue
            print(f"{g:4d} | {r[4]:8d} | bar1 {r[5]-r[4]:6d}  rows {r[6]-r[5]:6d}  retap {r[0]-r[6]:6d} | {r[4]-t[g-2][3]:6d}")
    print("drain warp 8: tile-first-gk | before acc_full wait, after TMEM drain, after epilogue stores | drain  epilogue  tile period")
    prevd = None
    for g in range(40, min(nblk, 76)):
        r = t[g]
        if r[12] == 0 and r[13] == 0: continue
        print(f"{g:4d} | {r[12]:8d} {r[13]:8d} {r[14]:8d} | {r[13]-r[12]:6d} {r[14]-r[13]:6d} {(r[12]-prevd) if prevd is not None else 0:6d}")
        prevd = r[12]
    per = (t[100, 11] - t[20, 11]) / 80.0
    print(f"average issue period over K blocks 20..100: {per:.0f} cycles (floor 768)")

if "bf16" in sys.argv:
    run(128, 256, 44, 256, 3, "bf16 3x3 256->256", bf16=True, impl=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
elif "small" in sys.argv:
    run(64, 64, 75, 256, 1, "1x1 64->256 M=360000 cp.async path", use_async=True, impl=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
elif "async" in sys.argv:
    run(64, 256, 44, 256, 3, "3x3 256->256 cp.async path", use_async=True)
else:
    run(64, 256, 44, 256, 3, "3x3 256->256 no prologue")
    run(64, 256, 44, 256, 3, "3x3 256->256 BN+ReLU prologue", pro=True)
