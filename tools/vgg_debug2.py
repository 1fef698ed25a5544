"""Activation-gradient comparison of the SSD-VGG trunk with torch autograd on the CPU (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import zsg_b200
from zsg_b200 import mdl, loss, spec
from oracle import synth, zsg_oracle as zo

cfg = synth.default_cfg("ssd_vgg"); cfg["device"] = "cuda"
ratios, scales = synth.ratios_scales(cfg)
net = mdl.get_default_net(9, cfg)
crit = loss.get_default_loss(ratios, scales, cfg)
B, seed = 2, 31
net.load_state_dict(synth.make_state_dict(0, "ssd_vgg"), strict=True)
net.train()
batch = synth.make_batch(B, seed=seed)
torch.manual_seed(seed)
cb = {k: v.cuda() for k, v in batch.items()}
out = net(cb)
ls = crit(out, cb)
ls["loss"].mean().backward()
torch.cuda.synchronize()
eng = net.engine_for(B, 20)

# reference with retained activation gradients
sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in synth.make_state_dict(0, "ssd_vgg").items()}
e = "backbone.encoder."
acts = {}
x = batch["img"]
for i, L in enumerate(synth.vgg_layers()):
    if L[0] == "conv":
        x = F.conv2d(x, sd[f"{e}vgg.{i}.weight"], sd[f"{e}vgg.{i}.bias"], padding=L[4], dilation=L[5])
        x.retain_grad(); acts[f"pre{i}"] = x
    elif L[0] == "relu":
        x = F.relu(x)
    else:
        x = F.max_pool2d(x, L[1], L[2], L[3], ceil_mode=L[4])
        x.retain_grad(); acts[f"pool{i}"] = x
    if i == 22:
        s0 = x / x.norm(dim=1, keepdim=True)
sources = [s0, x]
for i, (_, _, _, stride, pad) in enumerate(synth.VGG_EXTRAS):
    x = F.conv2d(x, sd[f"{e}extras.{i}.weight"], sd[f"{e}extras.{i}.bias"], stride=stride, padding=pad)
    x.retain_grad(); acts[f"epre{i}"] = x
    x = F.relu(x)
    if i % 2 == 1:
        sources.append(x)
feats = [F.conv2d(sources[j], sd[f"{e}fproj{j + 1}.weight"], sd[f"{e}fproj{j + 1}.bias"]) for j in range(3)] + sources[3:]
for j, f in enumerate(feats):
    f.retain_grad(); acts[f"feat{j}"] = f
torch.manual_seed(seed)
h0, c0 = zo.draw_h0c0(B)
lang = zo.lstm_query_batched(sd, batch["qvec"], batch["qlens"], h0, c0)
att, bbx = zo.fuse_and_head(sd, feats, lang)
ols = zo.zsg_loss(att, bbx, batch["annot"], zo.default_anchors())
ols["loss"].mean().backward()

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)), float((a - b).abs().max()), float(b.abs().max())
def nhwc_of(t, ref):
    Bq, C, H, W = ref.shape
    return t.view(Bq, H, W, C).permute(0, 3, 1, 2)
dfeat = eng.dbg["feat"]  # placeholder to keep name
segs = eng.dbg["segments"]
print("gradient buffers vs autograd (rel l2, max abs err, max abs ref)")
for seg in segs:
    for kind, r in seg:
        if kind == "conv":
            i = int(r["name"].split(".")[-1])
            ref = acts[f"pre{i}"].grad
            print(f"pre{i:<3d}", "%.2e %.2e %.2e" % rel(nhwc_of(r["g"], ref), ref))
        elif kind == "pool":
            idx = [k for k in acts if k.startswith("pool")]
pi = [k for k, L in enumerate(synth.vgg_layers()) if L[0] == "pool"]
pools = [r for seg in segs for kind, r in seg if kind == "pool"]
for k, r in zip(pi, pools):
    ref = acts[f"pool{k}"].grad
    Bq, h, w, c, kk, st_, pad, ho, wo = r["geom"]
    print(f"pool{k} dy vs autograd", "%.2e %.2e %.2e" % rel(nhwc_of(r["dy"], ref), ref))
    # the kernel against torch on the engine's OWN buffers
    a = r["mask"].view(Bq, h, w, c).permute(0, 3, 1, 2).cpu().clone().requires_grad_(True)
    y = F.max_pool2d(a, kk, st_, pad, ceil_mode=(ho != (h + 2 * pad - kk) // st_ + 1))
    y.backward(r["dy"].view(Bq, ho, wo, c).permute(0, 3, 1, 2).cpu())
    want = a.grad * (a > 0)
    print(f"pool{k} kernel dx vs torch on the same inputs", "%.2e %.2e %.2e" % rel(nhwc_of(r["dx"], want), want))
for i, r in enumerate(eng.dbg["ext"]):
    ref = acts[f"epre{i}"].grad
    print(f"epre{i}", "%.2e %.2e %.2e" % rel(nhwc_of(r["g"], ref), ref), "(levels 3-5 hold the masked gradient)" if i in (3, 5, 7) else "")

# argmax agreement of pool5 between the reference activations and the engine's
r = pools[-1]
Bq, h, w, c, kk, st_, pad, ho, wo = r["geom"]
mine = r["mask"].view(Bq, h, w, c).permute(0, 3, 1, 2).cpu()
ref = F.relu(acts["pre28"]).detach()
print("act28 rel err", "%.2e %.2e %.2e" % rel(mine, ref))
_, i1 = F.max_pool2d(mine, kk, st_, pad, return_indices=True)
v2, i2 = F.max_pool2d(ref, kk, st_, pad, return_indices=True)
diff = (i1 != i2)
print("pool5 windows", i1.numel(), "argmax differs in", int(diff.sum()), " of which max value > 0:", int((diff & (v2 > 0)).sum()))
# gap between the two candidates where they differ
flat = ref.flatten(2)
a = flat.gather(2, i1.flatten(2)); b = flat.gather(2, i2.flatten(2))
d = (b - a).flatten()[diff.flatten()]
print("reference value at my argmax vs at its own: max gap", float(d.abs().max()) if d.numel() else 0.0, " exact ties:", int((d == 0).sum()))
print("distinct positive values in ref act28:", int(torch.unique(ref[ref > 0]).numel()), "of", int((ref > 0).sum()))
