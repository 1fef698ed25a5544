"""Stage-by-stage comparison of the SSD-VGG trunk with the CPU oracle (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zsg_b200
from zsg_b200 import mdl, loss, evaluator
from oracle import synth, zsg_oracle as zo

cfg = synth.default_cfg("ssd_vgg"); cfg["device"] = "cuda"
ratios, scales = synth.ratios_scales(cfg)
net = mdl.get_default_net(9, cfg)
crit = loss.get_default_loss(ratios, scales, cfg)
B, seed = 2, 31
net.load_state_dict(synth.make_state_dict(0, "ssd_vgg"), strict=True)
net.train()
batch = synth.make_batch(B, seed=seed)
if os.environ.get("VGG_NO_OVERLAP"):
    net.engine_for(B, 20).overlap_wgrad = False
torch.manual_seed(seed)
out = net({k: v.cuda() for k, v in batch.items()})
ls = crit(out, {k: v.cuda() for k, v in batch.items()})
ls["loss"].mean().backward()
torch.cuda.synchronize()
sd = synth.make_state_dict(0, "ssd_vgg")
ols, omet, ograds, oout, _ = zo.train_step(sd, batch, seed=seed, do_adam=False)
torch.manual_seed(seed)
inter = zo.zsgnet_forward(sd, batch, return_inter=True)["_inter"]
eng = net.engine_for(B, 20)
off = eng.dbg["lvl_off"]
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
for i, f in enumerate(inter["feats"]):
    mine = eng.dbg["feat"][off[i]:off[i + 1]].view(B, f.shape[2], f.shape[3], 256).permute(0, 3, 1, 2)
    print("level", i, rel(mine, f))
print("att", rel(out["att_out"], oout["att_out"]), "bbx", rel(out["bbx_out"], oout["bbx_out"]))
print("loss", ls["loss"].item(), ols["loss"].item())
cgrads = ograds
print("parameter, |mine - cpu| / |cpu|, |torch-cuda - cpu| / |cpu|")
for k, g in ograds.items():
    if g is None: continue
    print(f"{k:40s} {rel(net.get_parameter(k).grad, g):.2e}  {rel(cgrads[k], g):.2e}")
