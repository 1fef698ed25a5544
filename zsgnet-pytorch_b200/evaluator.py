"""Drop-in for the reference's code/evaluator.py: `get_default_eval(ratios, scales, cfg)` -> module
with `.met_keys == ['Acc', 'MaxPos']`; forward returns Acc, MaxPos, idxs, pred_boxes, pred_scores
(evaluator.py:48-105).  One fused CUDA pass: argmax of the scores, IoU argmax, decode of just the two
selected boxes per sample (the reference decodes all 17460), IoU against the ground truth."""
from typing import Dict

import torch
from torch import nn

from . import ops, spec
from .anchors import create_anchors
from .loss import _packed_base


class Evaluator(nn.Module):
    def __init__(self, ratios, scales, cfg):
        super().__init__()
        self.cfg = cfg
        self.ratios, self.scales = ratios, scales
        self.met_keys = ["Acc", "MaxPos"]
        self.anchs = None
        self.acc_iou_threshold = cfg["acc_iou_threshold"]

    @torch.no_grad()
    def forward(self, out: Dict[str, torch.Tensor], inp: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        att, bbx = out["att_out"].detach(), out["bbx_out"].detach()
        dev = att.device
        if self.anchs is None:
            sizes = [(int(h), int(w)) for h, w in out["feat_sizes"][:len(spec.LEVEL_SIZES)].cpu().tolist()]
            self.anchs = create_anchors(sizes, self.ratios, self.scales, flatten=True, device=dev)
        B, A = att.shape[0], att.shape[1]
        packed = _packed_base(att, bbx)
        if not packed:
            att, bbx = att.contiguous(), bbx.contiguous()
        sa, sr = (5, 5) if packed else (1, 4)
        best = torch.empty(B, dtype=torch.int64, device=dev)
        scores = torch.empty(B, device=dev)
        boxes = torch.empty(B, 4, dtype=torch.float64, device=dev)
        metrics = torch.empty(2 + 2 * B, device=dev)
        ops.evaluate(att, sa, bbx, sr, inp["annot"].contiguous().float(), self.anchs,
                     inp["img_size"].contiguous().float(), B, A, float(self.acc_iou_threshold), best, scores, boxes,
                     metrics)
        return {"Acc": metrics[0], "MaxPos": metrics[1], "idxs": inp["idxs"], "pred_boxes": boxes,
                "pred_scores": scores, "best_ids": best}


def get_default_eval(ratios, scales, cfg):
    return Evaluator(ratios, scales, cfg)
