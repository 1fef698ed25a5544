"""Device time of the loss pass alone (zsg_loss_grad on packed [B, A, 5] buffers), 20 calls in one CUDA graph.
python tools/time_loss.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import ops, spec, anchors as anc, _lib
import numpy as np

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A = spec.NUM_ANCHORS
dev = "cuda"
g = torch.Generator().manual_seed(0)
out = torch.randn(B, A, 5, generator=g).to(dev)
dout = torch.empty_like(out)
c = torch.rand(B, 2, generator=g) * 1.0 - 0.5
hw = torch.rand(B, 2, generator=g) * 0.5 + 0.2
annot = torch.cat([c - hw / 2, c + hw / 2], 1).float().to(dev)
anchs = anc.create_anchors([(s, s) for s in spec.LEVEL_SIZES], [0.5, 1, 2], 4 * np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)]), flatten=True, device=dev)
pos = torch.zeros(B, A, dtype=torch.uint8, device=dev)
top1 = torch.zeros(B, dtype=torch.int64, device=dev)
ws = torch.zeros(_lib.load().zsg_match_loss_workspace_bytes(B), dtype=torch.uint8, device=dev)
losses = torch.zeros(3, dtype=torch.float64, device=dev)
flat, dflat = out.view(-1), dout.view(-1)
ops.match(annot, anchs, B, A, 0.6, True, top1, pos, ws)


def loss_pass():
    ops.loss_grad(flat[4:], 5, out, 5, annot, anchs, pos, B, A, 0.25, 2.0, 1.0, losses, dflat[4:], 5, dout, 5, ws)


loss_pass(); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr, capture_error_mode="thread_local"):
    for _ in range(20):
        loss_pass()
gr.replay(); torch.cuda.synchronize()
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record(); gr.replay(); a1.record(); torch.cuda.synchronize()
us = a0.elapsed_time(a1) / 20 * 1e3
print(f"B={B} TPB={os.environ.get('ZSG_LOSS_TPB', 'auto')}: {us:.1f} us per call, {B * A * 40 / us / 1e3:.0f} GB/s, positives {int(pos.sum())}, loss {losses.tolist()}")
