"""Checkpoints in the reference's format (Learner.save_model_dict / load_model_dict, utils.py:440-497):
a torch.save'd dict with model_state_dict, optimizer_state_dict, scheduler_state_dict, num_it, num_epoch, cfgtxt,
best_met.  Files written by the reference (incl. the published ones, README.md:87-88) load here and files written
here load there: parameter names and shapes are the reference's (spec.py), tensors are saved contiguous OIHW, and
FusedAdam.state_dict() has torch.optim.Adam's layout.  A 'module.' prefix (the reference saves the DDP wrapper's
state_dict, main_dist.py:37-40) is accepted on load and can be written with ddp_prefix=True."""
import json

import torch


def model_state_dict(net, ddp_prefix=False):
    pre = "module." if ddp_prefix else ""
    return {pre + k: v.detach().contiguous().clone() for k, v in net.state_dict().items()}


def save_model_dict(path, net, optimizer=None, scheduler=None, num_it=0, num_epoch=0, best_met=0.0, cfg=None,
                    ddp_prefix=False):
    ckpt = {"model_state_dict": model_state_dict(net, ddp_prefix), "num_it": num_it, "num_epoch": num_epoch,
            "cfgtxt": json.dumps(dict(cfg) if cfg is not None else {}, default=str), "best_met": best_met}
    if optimizer is not None:
        ckpt["optimizer_state_dict"] = optimizer.state_dict()
    if scheduler is not None:
        ckpt["scheduler_state_dict"] = scheduler.state_dict()
    with open(path, "wb") as f:
        torch.save(ckpt, f)
    return ckpt


def load_model_dict(path, net, optimizer=None, scheduler=None, strict=True):
    """Returns the bookkeeping fields {num_it, num_epoch, best_met} present in the file (utils.py:466-474)."""
    with open(path, "rb") as f:
        ckpt = torch.load(f, map_location="cpu", weights_only=False)
    sd = ckpt["model_state_dict"]
    if sd and all(k.startswith("module.") for k in sd):
        sd = {k[len("module."):]: v for k, v in sd.items()}
    net.load_state_dict(sd, strict=strict)
    if optimizer is not None and "optimizer_state_dict" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    if scheduler is not None and "scheduler_state_dict" in ckpt:
        scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    return {k: ckpt[k] for k in ("num_it", "num_epoch", "best_met") if k in ckpt}
