"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Recipe follows SURVEY.md section 8(c): stub yacs/spacy/fastprogress/fire, no-download
resnet50, cfg.device='cpu', CPU anchors.  Weights are the seeded ones of oracle/synth.py,
loaded into the reference modules with load_state_dict so that nothing big is committed.
"""
import os
import sys
import types
import json
from functools import partial

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
REF = "/root/reference"


def import_reference():
    class CfgNode(dict):
        def __init__(self, init_dict=None, key_list=None, new_allowed=False):
            super().__init__()
            for k, v in (init_dict or {}).items():
                self[k] = CfgNode(v) if isinstance(v, dict) else v

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

        def freeze(self):
            pass

    yacs = types.ModuleType("yacs")
    yc = types.ModuleType("yacs.config")
    yc.CfgNode = CfgNode
    yacs.config = yc
    spacy = types.ModuleType("spacy")
    spacy.load = lambda *a, **k: None
    fp = types.ModuleType("fastprogress")
    fpp = types.ModuleType("fastprogress.fastprogress")
    fpp.master_bar = lambda x, *a, **k: x
    fpp.progress_bar = lambda x, *a, **k: x
    fp.fastprogress = fpp
    fp.master_bar, fp.progress_bar = fpp.master_bar, fpp.progress_bar
    fire = types.ModuleType("fire")
    fire.Fire = lambda *a, **k: None
    sys.modules.update({"yacs": yacs, "yacs.config": yc, "spacy": spacy, "fastprogress": fp,
                        "fastprogress.fastprogress": fpp, "fire": fire})
    import torchvision.models as tvm
    orig = tvm.resnet50
    tvm.resnet50 = lambda *a, **k: orig(weights=None)
    os.chdir(REF)
    sys.path.insert(0, os.path.join(REF, "code"))
    import anchors, loss, evaluator, mdl          # noqa: E401
    from extended_config import cfg
    cfg.device = "cpu"
    os.chdir(REPO)
    return anchors, loss, evaluator, mdl, cfg


def vgg_cases(synth, anchors, loss, evaluator, mdl, cfg, out_dir):
    """Config 5's trunk: ZSGNet over SSDBackBone(build_ssd('train')) (mdl.py:413-418, minus the torch.load of
    ./weights/vgg16_reducedfc.pth, which is absent: SURVEY.md section 8(c) step 7), seeded weights of
    synth.make_state_dict(0, 'ssd_vgg')."""
    import ssd_vgg
    ratios, scales = synth.ratios_scales()
    cpu = torch.device("cpu")
    cfg.mdl_to_use = "ssd_vgg"
    encoder = ssd_vgg.build_ssd("train", cfg=cfg)
    net = mdl.ZSGNet(mdl.SSDBackBone(encoder, cfg), 9, cfg=cfg)
    crit = loss.get_default_loss(ratios, scales, cfg)
    crit.get_anchors = partial(anchors.create_anchors, ratios=ratios, scales=scales, flatten=True, device=cpu)
    ev = evaluator.get_default_eval(ratios, scales, cfg)
    ev.get_anchors = partial(anchors.create_anchors, ratios=ratios, scales=scales, flatten=True, device=cpu)
    full = {}
    for name, B, seed, var_len in (("vgg2", 2, 31, False), ("vgg3v", 3, 32, True)):
        print("load_state_dict:", net.load_state_dict(synth.make_state_dict(0, "ssd_vgg"), strict=True))
        net.train()
        net.zero_grad()
        crit.anchs = None
        ev.anchs = None
        batch = synth.make_batch(B, seed=seed, var_len=var_len)
        torch.manual_seed(seed)
        out = net(batch)
        ls = crit(out, batch)
        ls["loss"].mean().backward()
        with torch.no_grad():
            met = ev(out, batch)
        grads = {k: p.grad for k, p in net.named_parameters()}
        gnorm = {k: (float(g.double().norm()) if g is not None else None) for k, g in grads.items()}
        e = "backbone.encoder."
        arrs = dict(
            att_stride=out["att_out"].detach().squeeze(-1)[:, ::53].numpy(),
            bbx_stride=out["bbx_out"].detach()[:, ::53].numpy(),
            best_ids=torch.sigmoid(out["att_out"].detach()).squeeze(-1).max(1)[1].numpy(),
            pred_boxes=met["pred_boxes"].numpy(), pred_scores=met["pred_scores"].numpy())
        for k in (e + "vgg.0.bias", e + "vgg.21.bias", e + "vgg.31.bias", e + "fproj1.bias", e + "extras.7.bias",
                  "att_reg_box.5.bias", "lstm.bias_ih_l0"):
            arrs["g:" + k] = grads[k].numpy()
        for k in (e + "vgg.0.weight", e + "vgg.14.weight", e + "vgg.21.weight", e + "vgg.28.weight", e + "vgg.31.weight",
                  e + "vgg.33.weight", e + "fproj1.weight", e + "fproj2.weight", e + "fproj3.weight",
                  e + "extras.1.weight", e + "extras.5.weight", e + "extras.6.weight", "att_reg_box.0.0.weight"):
            arrs["gs:" + k] = grads[k].flatten()[::101].numpy()
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), **arrs)
        full[name] = dict(B=B, seed=seed, var_len=var_len, loss=ls["loss"].item(), cls_ls=ls["cls_ls"].item(),
                          box_ls=ls["box_ls"].item(), Acc=met["Acc"].item(), MaxPos=met["MaxPos"].item(),
                          gnorm=gnorm)
        print(name, full[name]["loss"], full[name]["cls_ls"], full[name]["box_ls"])
    return full


def main():
    from oracle import synth
    anchors, loss, evaluator, mdl, cfg = import_reference()
    if "vgg" in sys.argv[1:]:                      # add the SSD-VGG cases to an existing meta.json
        torch.set_num_threads(8)
        out_dir = os.path.join(REPO, "tests", "golden")
        with open(os.path.join(out_dir, "meta.json")) as f:
            meta = json.load(f)
        meta["vgg_cases"] = vgg_cases(synth, anchors, loss, evaluator, mdl, cfg, out_dir)
        with open(os.path.join(out_dir, "meta.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print("done (vgg)")
        return
    torch.set_num_threads(8)
    ratios, scales = synth.ratios_scales()
    cpu = torch.device("cpu")

    net = mdl.get_default_net(num_anchors=9, cfg=cfg)
    sd = synth.make_state_dict(0)
    missing = net.load_state_dict(sd, strict=True)
    print("load_state_dict:", missing)
    crit = loss.get_default_loss(ratios, scales, cfg)
    crit.get_anchors = partial(anchors.create_anchors, ratios=ratios, scales=scales, flatten=True, device=cpu)
    ev = evaluator.get_default_eval(ratios, scales, cfg)
    ev.get_anchors = partial(anchors.create_anchors, ratios=ratios, scales=scales, flatten=True, device=cpu)

    out_dir = os.path.join(REPO, "tests", "golden")
    meta = {}

    # ---- 1. anchors table --------------------------------------------------------------
    sizes = torch.tensor([[s, s] for s in synth.LEVEL_SIZES])
    anchs = anchors.create_anchors(sizes, ratios=ratios, scales=scales, flatten=True, device=cpu)
    assert anchs.dtype == torch.float64 and anchs.shape == (synth.NUM_ANCHORS, 4)
    np.savez_compressed(os.path.join(out_dir, "anchors.npz"), anchs=anchs.numpy())

    # ---- 2. loss / evaluator on random head outputs, incl. adversarial boxes -----------
    cases = {}
    for name, B, seed, adv in (("rand8", 8, 11, False), ("adv8", 8, 12, True), ("rand3", 3, 13, False)):
        g = torch.Generator().manual_seed(seed)
        batch = synth.make_batch(B, seed=seed, adversarial=adv)
        att = (torch.randn(B, synth.NUM_ANCHORS, 1, generator=g) * 1.5 - 3.0).requires_grad_(True)
        bbx = (torch.randn(B, synth.NUM_ANCHORS, 4, generator=g) * 0.7).requires_grad_(True)
        out = {"att_out": att, "bbx_out": bbx, "feat_sizes": sizes, "num_f_out": torch.tensor([6])}
        crit.anchs = None
        ev.anchs = None
        ls = crit(out, batch)
        ls["loss"].mean().backward()
        with torch.no_grad():
            met = ev(out, batch)
            iou = anchors.IoU_values(batch["annot"], anchs)
            top1 = iou.max(1)[1]
            pos = (iou > cfg["matching_threshold"])
            pos[torch.arange(B), top1] = True
        pos_idx = pos.nonzero()
        cases[name] = dict(
            B=B, seed=seed, adv=adv,
            loss=ls["loss"].item(), cls_ls=ls["cls_ls"].item(), box_ls=ls["box_ls"].item(),
            Acc=met["Acc"].item(), MaxPos=met["MaxPos"].item())
        np.savez_compressed(
            os.path.join(out_dir, f"loss_{name}.npz"),
            annot=batch["annot"].numpy(), top1=top1.numpy(), pos_idx=pos_idx.numpy().astype(np.int32),
            best=met["pred_scores"].numpy(), pred_boxes=met["pred_boxes"].numpy(),
            best_ids=torch.sigmoid(att.detach()).squeeze(-1).max(1)[1].numpy(),
            datt_idx=pos_idx.numpy().astype(np.int32),
            datt_pos=att.grad.squeeze(-1)[pos].numpy(), dbbx_pos=bbx.grad[pos].numpy(),
            datt_stride=att.grad.squeeze(-1)[:, ::97].numpy(),
            dbbx_abs_sum=bbx.grad.abs().sum().item(), datt_abs_sum=att.grad.abs().sum().item(),
            iou_top=iou[torch.arange(B), top1].numpy())
    meta["loss_cases"] = cases

    # ---- 3. full network: forward, loss, backward, metric -------------------------------
    full = {}
    for name, B, seed, var_len in (("net2", 2, 21, False), ("net3v", 3, 22, True)):
        net.load_state_dict(synth.make_state_dict(0), strict=True)
        net.train()
        net.zero_grad()
        crit.anchs = None
        ev.anchs = None
        batch = synth.make_batch(B, seed=seed, var_len=var_len)
        torch.manual_seed(seed)
        out = net(batch)
        ls = crit(out, batch)
        ls["loss"].mean().backward()
        with torch.no_grad():
            met = ev(out, batch)
        grads = {k: p.grad for k, p in net.named_parameters()}
        gnorm = {k: (float(g.double().norm()) if g is not None else None) for k, g in grads.items()}
        sd_after = net.state_dict()
        arrs = dict(
            att_stride=out["att_out"].detach().squeeze(-1)[:, ::53].numpy(),
            bbx_stride=out["bbx_out"].detach()[:, ::53].numpy(),
            att_absmean=out["att_out"].detach().abs().mean().item(),
            best_ids=torch.sigmoid(out["att_out"].detach()).squeeze(-1).max(1)[1].numpy(),
            pred_boxes=met["pred_boxes"].numpy(), pred_scores=met["pred_scores"].numpy(),
            bn1_rm=sd_after["backbone.encoder.bn1.running_mean"].numpy(),
            bn1_rv=sd_after["backbone.encoder.bn1.running_var"].numpy(),
            l4_rm=sd_after["backbone.encoder.layer4.2.bn3.running_mean"].numpy(),
            l4_rv=sd_after["backbone.encoder.layer4.2.bn3.running_var"].numpy(),
        )
        for k in ("att_reg_box.5.bias", "att_reg_box.0.0.bias", "lstm.bias_ih_l0", "lstm.bias_hh_l0_reverse",
                  "backbone.fpn.P3_2.bias", "backbone.encoder.bn1.weight", "backbone.encoder.layer3.2.bn2.bias"):
            arrs["g:" + k] = grads[k].numpy()
        for k in ("att_reg_box.5.weight", "att_reg_box.0.0.weight", "backbone.encoder.conv1.weight",
                  "backbone.encoder.layer1.0.conv1.weight", "backbone.encoder.layer4.2.conv3.weight",
                  "backbone.fpn.P6.weight", "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.weight_ih_l0_reverse"):
            arrs["gs:" + k] = grads[k].flatten()[::101].numpy()
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), **arrs)
        full[name] = dict(B=B, seed=seed, var_len=var_len, loss=ls["loss"].item(), cls_ls=ls["cls_ls"].item(),
                          box_ls=ls["box_ls"].item(), Acc=met["Acc"].item(), MaxPos=met["MaxPos"].item(),
                          gnorm=gnorm)
        print(name, full[name]["loss"], full[name]["cls_ls"], full[name]["box_ls"])
    meta["net_cases"] = full

    # ---- 4. nearest-upsample index tables (fpn_resnet.py:161,166) -----------------------
    up = {}
    for (i, o) in ((10, 19), (19, 38)):
        src = torch.arange(i, dtype=torch.float32).view(1, 1, i, 1).expand(1, 1, i, i).contiguous()
        idx = torch.nn.functional.interpolate(src, size=(o, o))[0, 0, :, 0].long().tolist()
        up[f"{i}->{o}"] = idx
    meta["upsample_idx"] = up
    meta["vgg_cases"] = vgg_cases(synth, anchors, loss, evaluator, mdl, cfg, out_dir)
    meta["torch"] = torch.__version__
    with open(os.path.join(out_dir, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("done")


if __name__ == "__main__":
    main()
