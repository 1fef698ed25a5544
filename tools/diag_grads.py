"""Per-parameter gradient error of zsg_b200 vs the CPU fp32 oracle, next to torch CUDA fp32 vs the same oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import mdl, loss, evaluator
from oracle import synth, zsg_oracle as zo

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B, seed = 4, 31
cfg = synth.default_cfg(); cfg["device"] = "cuda"
ratios, scales = synth.ratios_scales(cfg)
net = mdl.get_default_net(9, cfg)
crit = loss.get_default_loss(ratios, scales, cfg)
net.load_state_dict(synth.make_state_dict(0), strict=True)
net.train()
cb = synth.make_batch(B, seed=seed, var_len=True)
torch.manual_seed(seed)
batch = {k: v.cuda() for k, v in cb.items()}
out = net(batch)
ls = crit(out, batch)
ls["loss"].mean().backward()
torch.cuda.synchronize()
sd = synth.make_state_dict(0)
ols, _, og, _, _ = zo.train_step(sd, cb, seed=seed, do_adam=False)
sdg = {k: v.cuda() for k, v in synth.make_state_dict(0).items()}
gls, _, gg, _, _ = zo.train_step(sdg, batch, seed=seed, do_adam=False)
print("loss mine/cpu/cuda", ls["loss"].item(), ols["loss"].item(), gls["loss"].item())
rows = []
for k, g in og.items():
    if g is None:
        continue
    r = g.double()
    mine = net.get_parameter(k).grad.cpu().double()
    cu = gg[k].cpu().double()
    n = r.norm().clamp_min(1e-30)
    rows.append((float((mine - r).norm() / n), float((cu - r).norm() / n), k, float(n)))
rows.sort(reverse=True)
print(f"{'mine vs cpu':>12s} {'cuda vs cpu':>12s}  param (|g|)")
for a, b, k, n in rows[:25]:
    print(f"{a:12.3e} {b:12.3e}  {k} ({n:.3e})")
import statistics
print("median mine", statistics.median(r[0] for r in rows), "median cuda", statistics.median(r[1] for r in rows))
