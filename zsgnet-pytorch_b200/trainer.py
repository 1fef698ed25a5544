"""Fused training step: the five calls of the reference's hot loop (utils.py:405-414: model, loss,
backward, optimiser, metric) issued back to back on device buffers, without the autograd glue.

The module API (mdl.ZSGNet + loss.ZSGLoss + evaluator.Evaluator under a stock training loop) runs the
same kernels; this class only removes per-step tensor allocation and Python autograd bookkeeping."""
import torch

from . import ops, spec
from .anchors import create_anchors
from .ddp import GradReducer
from .mdl import check_qlens
from .optim import FusedAdam


class FusedStep:
    def __init__(self, net, ratios, scales, cfg, lr=1e-4, reducer=None, broadcast_buffers=True):
        self.net, self.cfg, self.broadcast_buffers = net, cfg, broadcast_buffers
        dev = net.store.device
        sizes = [(s, s) for s in spec.LEVEL_SIZES]
        self.anchs = create_anchors(sizes, ratios, scales, flatten=True, device=dev)
        self.reducer = reducer if reducer is not None else GradReducer(net.store)
        self.opt = FusedAdam(net.parameters(), lr=lr, betas=(0.9, 0.99), net=net, reducer=self.reducer)
        self._per_b = {}
        # the anchor match needs only the annotations: it runs on this stream under the forward pass
        self._match_stream = torch.cuda.Stream(device=dev)

    def _bufs(self, B):
        if B not in self._per_b:
            dev, A = self.net.store.device, spec.NUM_ANCHORS
            self._per_b[B] = dict(
                losses=torch.empty(3, dtype=torch.float64, device=dev), top1=torch.empty(B, dtype=torch.int64, device=dev),
                pos=torch.empty(B, A, dtype=torch.uint8, device=dev), ws=ops.match_loss_workspace(B, dev),
                best=torch.empty(B, dtype=torch.int64, device=dev), scores=torch.empty(B, device=dev),
                boxes=torch.empty(B, 4, dtype=torch.float64, device=dev), metrics=torch.empty(2 + 2 * B, device=dev))
        return self._per_b[B]

    @torch.no_grad()
    def step(self, batch, h0=None, c0=None, do_opt=True, do_eval=True):
        """batch: dict of DEVICE tensors with the reference's keys.  Returns device tensors (no sync)."""
        net, cfg = self.net, self.cfg
        img, qvec, qlens = batch["img"], batch["qvec"], batch["qlens"]
        B, A = img.shape[0], spec.NUM_ANCHORS
        qlens_cpu = batch["qlens_cpu"] if "qlens_cpu" in batch else qlens.cpu()
        max_qlen = check_qlens(qlens_cpu, qvec.shape[1])
        if h0 is None:
            h0, c0 = torch.randn(2, B, 128), torch.randn(2, B, 128)      # mdl.py:279-294
        _, perm = qlens_cpu.sort(0, descending=True)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(B)
        eng = net.engine_for(B, max(max_qlen, 1))
        eng.set_inputs(img, qvec[:, :max_qlen], qlens_cpu, inv, h0, c0)
        if self.broadcast_buffers:
            self.reducer.broadcast_buffers(net)                # DDP broadcast_buffers=True (main_dist.py:39); no-op on one rank
        b = self._bufs(B)
        main = torch.cuda.current_stream()
        self._match_stream.wait_stream(main)                  # previous step's loss pass has read pos / the workspace
        with torch.cuda.stream(self._match_stream):
            ops.match(batch["annot"], self.anchs, B, A, float(cfg["matching_threshold"]), bool(cfg["use_multi"]), b["top1"],
                      b["pos"], b["ws"])
        out = eng.forward(training=True)
        net._bn_n.add_(1)
        flat, dflat = out.view(-1), eng.d_out.view(-1)
        main.wait_stream(self._match_stream)
        ops.loss_grad(flat[4:], 5, out, 5, batch["annot"], self.anchs, b["pos"], B, A, float(cfg["alpha"]), float(cfg["gamma"]),
                      float(cfg["lamb_reg"]), b["losses"], dflat[4:], 5, eng.d_out, 5, b["ws"])
        eng.backward(None, on_bucket=self.reducer.on_bucket if self.reducer.world > 1 else None)
        if do_opt:
            self.opt.step()
        else:
            self.reducer.finish()
        if do_eval:
            ops.evaluate(flat[4:], 5, out, 5, batch["annot"], self.anchs, batch["img_size"], B, A,
                         float(cfg["acc_iou_threshold"]), b["best"], b["scores"], b["boxes"], b["metrics"])
        return {"loss": b["losses"][0], "cls_ls": b["losses"][1], "box_ls": b["losses"][2], "Acc": b["metrics"][0],
                "MaxPos": b["metrics"][1]}
