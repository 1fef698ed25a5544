"""Run one training step op by op with a synchronize after each GEMM launch to find a failing launch (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import zsg_b200
from zsg_b200 import mdl, ops, dat_loader
from zsg_b200.trainer import FusedStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = {"do_norm": False, "use_same_atb": True, "mdl_to_use": "retina", "resize_img": [300, 300], "use_multi": True,
       "use_focal": True, "use_softmax": False, "alpha": 0.25, "gamma": 2, "emb_dim": 300, "matching_threshold": 0.6,
       "use_bidirectional": True, "lstm_dim": 128, "lamb_reg": 1, "acc_iou_threshold": 0.5, "use_lang": True,
       "use_img": True, "device": "cuda:0"}
torch.manual_seed(0)
net = mdl.get_default_net(9, cfg); net.train()
fs = FusedStep(net, [0.5, 1, 2], 4 * np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)]), cfg)
batch = {k: v.cuda() for k, v in dat_loader.synthetic_batch(B, 1).items()}
batch["qlens_cpu"] = batch["qlens"].cpu()
orig_conv, orig_wgrad = ops.ConvOp.__call__, ops.WgradOp.__call__
def wrap(orig, kind):
    def f(self):
        orig(self)
        try:
            torch.cuda.synchronize()
        except Exception as e:
            p = self.p
            print("FAILED", kind, "m", p.m, "cin", p.cin, "cout", p.cout, "r", p.r, "split/div", getattr(p, "split_k", getattr(p, "in_div", 0)),
                  "x_lo", bool(p.x_lo), flush=True)
            raise
    return f
ops.ConvOp.__call__ = wrap(orig_conv, "conv")
ops.WgradOp.__call__ = wrap(orig_wgrad, "wgrad")
fs.step(batch)
torch.cuda.synchronize()
print("step ok")
