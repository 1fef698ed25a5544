"""Import and drive the UNMODIFIED reference (TheShadow29/zsgnet-pytorch) on CPU.

TEST / MEASUREMENT INFRASTRUCTURE (part of oracle/): used by tests/golden/make_golden.py (golden vectors), tests that run
the reference's own Learner over the drop-in, and bench.py's `--impl reference` / `cpu_baseline` legs.  The product
package never imports this module.

The reference's sources are NOT part of this repository: they are read either from /root/reference (build container) or
from oracle/_ref/ (a git-ignored copy made by oracle/build_ref.py so that the reference arm can run on the GPU box,
where /root/reference does not exist).  Recipe: SURVEY.md section 8(c) -- stub the four absent packages
(yacs, spacy, fastprogress, fire), no-download resnet50, cfg.device = 'cpu', CPU anchors (anchors.py:66 defaults to
device='cuda')."""
import os
import sys
import types
from functools import partial

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CANDIDATES = ("/root/reference", os.path.join(HERE, "_ref"))


def find_reference():
    """Root of a reference tree (has code/mdl.py and configs/cfg.json), or None."""
    for root in CANDIDATES:
        if os.path.exists(os.path.join(root, "code", "mdl.py")) and os.path.exists(os.path.join(root, "configs", "cfg.json")):
            return root
    return None


class CfgNode(dict):
    """Stand-in for yacs.config.CfgNode (extended_config.py:1-12): a dict with attribute access."""

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


_cached = None


def import_reference(root=None):
    """-> dict(anchors, loss, evaluator, mdl, utils, cfg, root): the reference's own modules, imported once."""
    global _cached
    if _cached is not None:
        return _cached
    root = root or find_reference()
    if root is None:
        raise FileNotFoundError("no reference tree: neither /root/reference nor oracle/_ref (python oracle/build_ref.py)")
    yacs, yc = types.ModuleType("yacs"), types.ModuleType("yacs.config")
    yc.CfgNode = CfgNode
    yacs.config = yc
    spacy = types.ModuleType("spacy")
    spacy.load = lambda *a, **k: None
    fp, fpp = types.ModuleType("fastprogress"), types.ModuleType("fastprogress.fastprogress")
    fpp.master_bar = lambda x, *a, **k: x
    fpp.progress_bar = lambda x, *a, **k: x
    fp.fastprogress = fpp
    fp.master_bar, fp.progress_bar = fpp.master_bar, fpp.progress_bar
    fire = types.ModuleType("fire")
    fire.Fire = lambda *a, **k: None
    for name, mod in (("yacs", yacs), ("yacs.config", yc), ("spacy", spacy), ("fastprogress", fp),
                      ("fastprogress.fastprogress", fpp), ("fire", fire)):
        sys.modules.setdefault(name, mod)
    import torchvision.models as tvm
    if not getattr(tvm.resnet50, "_zsg_no_download", False):
        orig = tvm.resnet50
        tvm.resnet50 = lambda *a, **k: orig(weights=None)          # mdl.py:411 asks for ImageNet weights: no network here
        tvm.resnet50._zsg_no_download = True
    cwd = os.getcwd()
    os.chdir(root)                                                  # extended_config.py reads ./configs/*.json
    sys.path.insert(0, os.path.join(root, "code"))
    try:
        import anchors, evaluator, loss, mdl, utils                 # noqa: E401  (the reference's modules)
        from extended_config import cfg
    finally:
        os.chdir(cwd)
    cfg.device = "cpu"
    _cached = dict(anchors=anchors, loss=loss, evaluator=evaluator, mdl=mdl, utils=utils, cfg=cfg, root=root)
    return _cached


def cpu_anchor_fix(obj, ref, ratios, scales):
    """loss.py:37-39 / evaluator.py:42-44 bind create_anchors with its default device='cuda'."""
    obj.get_anchors = partial(ref["anchors"].create_anchors, ratios=ratios, scales=scales, flatten=True,
                              device=torch.device("cpu"))
    return obj


def ratios_scales(cfg):
    """main_dist.py:24-31: the config stores them as strings."""
    import numpy as np
    ratios = eval(cfg["ratios"], {}) if not isinstance(cfg["ratios"], list) else cfg["ratios"]
    scales = cfg["scale_factor"] * np.array(eval(cfg["scales"], {}) if not isinstance(cfg["scales"], list) else cfg["scales"])
    return ratios, scales


def build_reference_step(model="retina", lr=1e-4, seed=0):
    """The reference's training step (utils.py:405-414) on CPU over its own modules: returns step(batch) -> (loss, Acc)."""
    ref = import_reference()
    cfg = ref["cfg"]
    cfg.mdl_to_use = model
    ratios, scales = ratios_scales(cfg)
    torch.manual_seed(seed)
    if model == "ssd_vgg":                                           # mdl.py:413-418 minus the absent ./weights/vgg16_reducedfc.pth
        import ssd_vgg
        net = ref["mdl"].ZSGNet(ref["mdl"].SSDBackBone(ssd_vgg.build_ssd("train", cfg=cfg), cfg), 9, cfg=cfg)
    else:
        net = ref["mdl"].get_default_net(num_anchors=9, cfg=cfg)
    crit = cpu_anchor_fix(ref["loss"].get_default_loss(ratios, scales, cfg), ref, ratios, scales)
    evalr = cpu_anchor_fix(ref["evaluator"].get_default_eval(ratios, scales, cfg), ref, ratios, scales)
    opt = torch.optim.Adam(net.parameters(), lr=lr, betas=(0.9, 0.99))                                       # main_dist.py:50
    net.train()

    def step(batch):
        opt.zero_grad()
        out = net(batch)
        ls = crit(out, batch)
        loss = ls[crit.loss_keys[0]].mean()
        loss.backward()
        opt.step()
        met = evalr(out, batch)
        return float(loss.item()), float(met["Acc"].item())
    return step
