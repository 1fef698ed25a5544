"""Dump the registration order of the UNMODIFIED reference's parameters (named_parameters()) for both trunks into
tests/golden/param_order.json.  torch.optim.Adam's state_dict indexes parameters by that order (utils.py:479-497 saves
optimizer.state_dict()), so the drop-in module must register its parameters in the same order.
Run in the build container only: python tests/golden/make_param_order.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    anchors, loss, evaluator, mdl, cfg = import_reference()
    out = {}
    cfg.mdl_to_use = "retina"
    net = mdl.get_default_net(num_anchors=9, cfg=cfg)
    out["retina"] = [[n, list(p.shape)] for n, p in net.named_parameters()]
    out["retina_buffers"] = [n for n, _ in net.named_buffers()]
    import ssd_vgg
    cfg.mdl_to_use = "ssd_vgg"
    enc = ssd_vgg.build_ssd("train", cfg=cfg)
    net = mdl.ZSGNet(mdl.SSDBackBone(enc, cfg), 9, cfg=cfg)
    out["ssd_vgg"] = [[n, list(p.shape)] for n, p in net.named_parameters()]
    with open(os.path.join(HERE, "param_order.json"), "w") as f:
        json.dump(out, f)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
