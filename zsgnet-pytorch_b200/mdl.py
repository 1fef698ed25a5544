"""Drop-in for the reference's code/mdl.py: `get_default_net(num_anchors, cfg) -> nn.Module` whose
`forward(dict) -> dict` has the reference's keys, shapes and dtypes (mdl.py:338-403), the
reference's state_dict key names and train()/eval() BatchNorm semantics —
with every FLOP executed by libzsg_b200.so (hand-written sm_100a kernels).

There is no PyTorch or CPU fallback: constructing the net without a CUDA device, or without the
built library, raises."""
import math
from typing import Any, Dict

import torch
import torch.nn as nn

from . import _lib, ops, spec
from .engine import Engine, ParamStore


class _Node(nn.Module):
    """Name-space node so that parameters get the reference's dotted state_dict keys."""

    def forward(self, *a, **k):
        raise RuntimeError("zsg_b200: sub-modules are containers; call the ZSGNet module itself")


def _attach(root, dotted, tensor, is_param):
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, _Node())
        node = node._modules[p]
    if is_param:
        node.register_parameter(parts[-1], tensor)
    else:
        node.register_buffer(parts[-1], tensor)


class _ZSGNetFn(torch.autograd.Function):
    """One autograd node for the whole network: forward = engine.forward, backward = engine.backward."""

    @staticmethod
    def forward(ctx, net, eng, img, qvec, lens_cpu, inv_perm, h0, c0, *params):
        eng.set_inputs(img, qvec, lens_cpu, inv_perm, h0, c0)
        out = eng.forward(training=net.training)
        ctx.net, ctx.eng = net, eng
        eng.pending_backward = net.training
        return out.clone()                      # the engine's buffer is reused by the next step

    @staticmethod
    def backward(ctx, d_out):
        net, eng = ctx.net, ctx.eng
        if not eng.pending_backward:
            raise RuntimeError("zsg_b200: backward needs a training-mode forward on the same engine "
                               "(one backward per forward; retain_graph is not supported)")
        eng.pending_backward = False
        eng.backward(d_out.contiguous(), on_bucket=net._on_bucket)
        grads = tuple(net.store.grad_view(n) for n in net.param_names)
        return (None,) * 8 + grads


class ZSGNet(nn.Module):
    """Image encoder (cfg['mdl_to_use']: 'retina' = ResNet-50 + FPN, mdl.py:138-159; 'ssd_vgg' = SSD-VGG16,
    mdl.py:162-168 + ssd_vgg.py), bi-LSTM query encoder, language/grid tiling fusion and the shared six-level
    convolutional head of ZSGNet (mdl.py:171-403), on B200."""

    def __init__(self, backbone=None, n_anchors=9, final_bias=0.0, cfg=None, device=None):
        super().__init__()
        if backbone is not None:
            raise NotImplementedError("zsg_b200.ZSGNet builds its own trunk; pass backbone=None")
        if not torch.cuda.is_available():
            raise _lib.ZsgError("zsg_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _lib.load()
        cfg = cfg if cfg is not None else {}
        get = lambda k, d: (cfg[k] if k in cfg else d)
        unsupported = []
        self.model = get("mdl_to_use", "retina")
        if self.model not in spec.MODELS:
            unsupported.append("mdl_to_use=%r (built: 'retina' = ResNet-50+FPN, 'ssd_vgg' = SSD-VGG16)" % self.model)
        if list(get("resize_img", [300, 300])) != [300, 300]:
            unsupported.append("resize_img != [300, 300]")
        if get("do_norm", False):
            unsupported.append("do_norm=true")
        if not (get("use_lang", True) and get("use_img", True)):
            unsupported.append("language-/image-blind ablations")
        if not get("use_same_atb", True):
            unsupported.append("use_same_atb=false")
        if not get("use_bidirectional", True) or get("lstm_dim", 128) != 128 or get("emb_dim", 300) != 300:
            unsupported.append("LSTM other than bidirectional 300->128")
        if n_anchors != spec.ANCHORS_PER_CELL:
            unsupported.append("n_anchors != 9")
        if unsupported:
            raise NotImplementedError("zsg_b200 hot path does not cover: " + "; ".join(unsupported))
        self.cfg = cfg
        self.n_anchors = n_anchors
        dev = torch.device(device if device is not None else "cuda")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.store = ParamStore(dev, self.model)
        self.param_names = [n for n, _, _ in spec.trainable_specs(self.model)]
        for n in self.param_names + [u[0] for u in spec.unused_specs(self.model)]:
            _attach(self, n, nn.Parameter(self.store.view(n)), True)
        # BatchNorm buffers: float stats in one tensor, counters in another (views keep the reference names)
        bspecs = spec.buffer_specs(self.model)
        nf = sum(s[0] for _, s in bspecs if len(s))
        self._bn_f = torch.zeros(nf, device=dev)
        self._bn_n = torch.zeros(sum(1 for _, s in bspecs if not len(s)), dtype=torch.long, device=dev)
        self.bn_buffers, fo, no = {}, 0, 0
        for name, shp in bspecs:
            if len(shp):
                t = self._bn_f[fo:fo + shp[0]]
                fo += shp[0]
                if name.endswith("running_var"):
                    t.fill_(1.0)
            else:
                t = self._bn_n[no]
                no += 1
            self.bn_buffers[name] = t
            _attach(self, name, t, False)
        self._engines = {}
        self._on_bucket = None
        self.reset_parameters()

    # ------------------------------------------------------------------ init / state
    @torch.no_grad()
    def reset_parameters(self):
        """PyTorch-default initialisers (the reference applies none of its own, mdl.py:20-41,227-228;
        torchvision's resnet50 uses kaiming_normal(fan_out) for convs; ssd_vgg.py's convs keep nn.Conv2d's default).
        The ImageNet weights the reference downloads (mdl.py:411) or reads from ./weights/vgg16_reducedfc.pth
        (mdl.py:415-416) are loaded with load_state_dict when available."""
        st = self.store
        for name, shape, kind in spec.trainable_specs(self.model) + spec.unused_specs(self.model):
            v = st.view(name)
            if kind == "conv":
                if name.startswith("backbone.encoder.") and self.model == "retina":
                    nn.init.kaiming_normal_(v, mode="fan_out", nonlinearity="relu")
                else:
                    nn.init.kaiming_uniform_(v, a=math.sqrt(5))
            elif kind == "lin":
                nn.init.kaiming_uniform_(v, a=math.sqrt(5))
            elif kind == "bias":
                wname = name[:-4] + "weight"
                fan_in = int(torch.tensor(st.shapes[wname][1:]).prod())
                nn.init.uniform_(v, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))
            elif kind == "final_bias":
                v.zero_()
                v[torch.arange(4, shape[0], 5)] = -4.0               # mdl.py:214-215
            elif kind == "bn_w":
                v.fill_(1.0)
            elif kind == "bn_b":
                v.zero_()
            elif kind == "lstm":
                nn.init.uniform_(v, -1 / math.sqrt(128.0), 1 / math.sqrt(128.0))

    def _apply(self, fn, *a, **k):
        """Parameters are views of the arena: moving / casting the module would silently detach them."""
        probe = fn(torch.zeros(1, device=self.store.device))
        if probe.device != self.store.device or probe.dtype != torch.float32:
            raise RuntimeError("zsg_b200.ZSGNet lives on its CUDA device in float32; .to()/.half()/.cpu() are not "
                               "supported (build it with device=...)")
        return self

    # ------------------------------------------------------------------ forward
    def engine_for(self, B, T):
        key = (B, T)
        if key not in self._engines:
            if len(self._engines) >= 2:                           # bound memory: keep two shapes (train / last batch)
                self._engines.pop(next(iter(self._engines)))
                torch.cuda.empty_cache()
            bufs = {k: v for k, v in self.bn_buffers.items()}
            self._engines[key] = Engine(self.store, bufs, B, T, self.store.device)
        return self._engines[key]

    def forward(self, inp: Dict[str, Any]):
        img, qvec, qlens = inp["img"], inp["qvec"], inp["qlens"]
        dev = self.store.device
        if img.device != dev:
            raise RuntimeError(f"zsg_b200: inputs must be on {dev} (got {img.device}); no CPU path exists")
        B = img.shape[0]
        # mdl.py:357: the reference synchronises here too (qlens.max().item())
        qlens_cpu = qlens.detach().cpu()
        max_qlen = int(qlens_cpu.max().item())
        qvec = qvec[:, :max_qlen, :]
        # mdl.py:279-294, 307: h0 then c0 from the global CPU RNG, consumed in sorted-row order (309-319)
        h0 = torch.randn(2, B, 128)
        c0 = torch.randn(2, B, 128)
        _, perm = qlens_cpu.sort(0, descending=True)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(B)
        eng = self.engine_for(B, max(max_qlen, 1))
        params = [self.get_parameter(n) for n in self.param_names]
        out = _ZSGNetFn.apply(self, eng, img.contiguous().float(), qvec.float(), qlens_cpu, inv, h0, c0, *params)
        if self.training:
            self._bn_n.add_(1)                                    # num_batches_tracked
        feat_sizes = torch.tensor([[s, s] for s in spec.LEVEL_SIZES], device=dev)
        return {"att_out": out[..., 4:], "bbx_out": out[..., :4], "feat_sizes": feat_sizes,
                "num_f_out": torch.tensor([len(spec.LEVEL_SIZES)], device=dev)}


def get_default_net(num_anchors=1, cfg=None):
    """Same signature as mdl.py:406-422."""
    dev = None
    if cfg is not None and "device" in cfg and str(cfg["device"]).startswith("cuda"):
        dev = cfg["device"]
    return ZSGNet(None, num_anchors, cfg=cfg, device=dev)
