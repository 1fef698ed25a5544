"""Per-launch table of the implicit-GEMM kernels in one training step (B=64): dims, time, algorithmic TF/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, collections
import zsg_b200
from zsg_b200 import mdl, ops, dat_loader
import numpy as np
from zsg_b200.trainer import FusedStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
DT = sys.argv[2] if len(sys.argv) > 2 else "fp32"
cfg = {"do_norm": False, "use_same_atb": True, "mdl_to_use": "retina", "resize_img": [300, 300], "use_multi": True,
       "use_focal": True, "use_softmax": False, "alpha": 0.25, "gamma": 2, "emb_dim": 300, "matching_threshold": 0.6,
       "use_bidirectional": True, "lstm_dim": 128, "lamb_reg": 1, "acc_iou_threshold": 0.5, "use_lang": True,
       "use_img": True, "device": "cuda:0", "zsg_dtype": DT, "zsg_quiet": True}
torch.manual_seed(0)
net = mdl.get_default_net(9, cfg); net.train()
fs = FusedStep(net, [0.5, 1, 2], 4 * np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)]), cfg)
batch = {k: v.cuda() for k, v in dat_loader.synthetic_batch(B, 1).items()}
batch["qlens_cpu"] = batch["qlens"].cpu()
for _ in range(2):
    fs.step(batch)
torch.cuda.synchronize()
eng = net.engine_for(B, 20)
eng.overlap_wgrad = False          # weight gradients in line, so that per-launch intervals do not overlap
prof = ops.LaunchProfiler(); ops.PROFILER = prof
fs.step(batch)
torch.cuda.synchronize()
ops.PROFILER = None
eng = net.engine_for(B, 20)
allops = [it[1] for it in eng.fwd if it[0] == "op"] + [o for o in eng.bwd if isinstance(o, (ops.ConvOp, ops.WgradOp))]
assert len(allops) == len(prof.records), (len(allops), len(prof.records))
rows = []
for op, (kern, fl, a, b) in zip(allops, prof.records):
    p = op.p
    ms = a.elapsed_time(b)
    kind = "wgrad" if kern.startswith("wgrad") else ("dgrad" if op in eng.bwd else "fwd")
    rows.append((ms, kind, p.m, p.cin, p.cout, p.r, getattr(p, "in_div", 1), fl / ms / 1e9))
tot = sum(r[0] for r in rows)
print(f"total GEMM time {tot:.1f} ms over {len(rows)} launches")
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for ms, kind, m, cin, cout, r, div, tf in rows:
    key = (kind, m, cin, cout, r, div)
    agg[key][0] += ms; agg[key][1] += tf * ms; agg[key][2] += 1
print(f"{'ms':>8s} {'n':>3s} {'TF/s':>7s}  kind   M       cin  cout  k div")
for key, (ms, tfms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
    kind, m, cin, cout, r, div = key
    print(f"{ms:8.2f} {n:3d} {tfms/ms:7.1f}  {kind:6s} {m:7d} {cin:5d} {cout:5d} {r:2d} {div}")
for kind in ("fwd", "dgrad", "wgrad"):
    ms = sum(r[0] for r in rows if r[1] == kind); fl = sum(r[7] * r[0] for r in rows if r[1] == kind)
    print(f"{kind}: {ms:.1f} ms, {fl/ms:.1f} TF/s")
