"""Extra golden cases generated from the UNMODIFIED reference (run in the build container only):
  loss_zero_area.npz  -- ZSGLoss on a zero-height ground-truth box: the NaN guard of loss.py:128-133.
Usage: python tests/golden/make_golden_extra.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import ref_harness, synth  # noqa: E402


def main():
    ref = ref_harness.import_reference()
    cfg = ref["cfg"]
    ratios, scales = ref_harness.ratios_scales(cfg)
    crit = ref_harness.cpu_anchor_fix(ref["loss"].get_default_loss(ratios, scales, cfg), ref, ratios, scales)
    B, A = 2, synth.NUM_ANCHORS
    batch = synth.make_batch(B, seed=4)
    batch["annot"][0] = torch.tensor([0.1, 0.1, 0.1, 0.4])
    g = torch.Generator().manual_seed(4)
    att = (torch.randn(B, A, 1, generator=g) - 3.0).requires_grad_(True)
    bbx = (torch.randn(B, A, 4, generator=g) * 0.3).requires_grad_(True)
    sizes = torch.tensor([[s, s] for s in synth.LEVEL_SIZES])
    out = {"att_out": att, "bbx_out": bbx, "feat_sizes": sizes, "num_f_out": torch.tensor([6])}
    ls = crit(out, batch)
    ls["loss"].mean().backward()
    anchs = ref["anchors"].create_anchors(sizes, ratios=ratios, scales=scales, flatten=True, device=torch.device("cpu"))
    iou = ref["anchors"].IoU_values(batch["annot"], anchs)
    np.savez_compressed(os.path.join(HERE, "loss_zero_area.npz"),
                        losses=np.array([ls["loss"].item(), ls["cls_ls"].item(), ls["box_ls"].item()]),
                        top1=iou.max(1)[1].numpy(),
                        datt_abs_sum=np.array(0.0 if att.grad is None else att.grad.abs().sum().item()),
                        dbbx_abs_sum=np.array(0.0 if bbx.grad is None else bbx.grad.abs().sum().item()))
    print({k: v.item() for k, v in ls.items()}, att.grad is None)


if __name__ == "__main__":
    main()
