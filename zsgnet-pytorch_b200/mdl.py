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


def check_qlens(qlens_cpu, t_avail):
    """Host-side validation of the query lengths (the reference fails in pack_padded_sequence for both cases, mdl.py:315-316;
    here an unchecked 0 or too long length would read outside the query buffer).  Returns max_qlen."""
    lo, hi = int(qlens_cpu.min().item()), int(qlens_cpu.max().item())
    if lo < 1:
        raise ValueError(f"zsg_b200: qlens must be >= 1 (got {lo}): an empty query cannot be encoded")
    if hi > t_avail:
        raise ValueError(f"zsg_b200: qlens up to {hi} but qvec holds only {t_avail} tokens per query")
    return hi


def draw_lstm_state(qlens_cpu):
    """mdl.py:279-294: h0 then c0 ~ torch.randn(2, B, 128) from the global CPU RNG, assigned to the rows of the
    length-sorted batch (mdl.py:307-319).  Returns (h0, c0, inverse permutation)."""
    B = qlens_cpu.shape[0]
    h0 = torch.randn(2, B, 128)
    c0 = torch.randn(2, B, 128)
    _, perm = qlens_cpu.sort(0, descending=True)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(B)
    return h0, c0, inv


class _ZSGNetFn(torch.autograd.Function):
    """One autograd node for the whole network: forward = engine.forward, backward = engine.backward."""

    @staticmethod
    def forward(ctx, net, eng, img, qvec, lens_cpu, inv_perm, h0, c0, *params):
        eng.set_inputs(img, qvec, lens_cpu, inv_perm, h0, c0, staged=net._staged)
        net._staged = None
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
        if net.direct_grads:
            # the gradients are already where the optimiser reads them: param.grad IS the arena view (no AccumulateGrad
            # clone of 161 strided tensors per step).  Overwrites instead of accumulating: zero_grad() every step, as
            # the reference loop does (utils.py:410).
            for p, g in zip(net._param_list, net._grad_views):
                p.grad = g
            return (None,) * 9
        return (None,) * 8 + tuple(net._grad_views)


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
        self.do_norm = bool(get("do_norm", False))          # mdl.py:118-130, built (zsg_l2norm_fwd / bwd in front of the head)
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
        # arithmetic of the dense contractions: "fp32" (3xTF32, the reference's fp32 results; BASELINE configs[1]) or "bf16"
        # (bf16 tensor-core convs with fp32 accumulation; configs[2..4]).  cfg key `zsg_dtype`, or set_compute_dtype().
        self.compute_dtype = "fp32"
        self.set_compute_dtype(get("zsg_dtype", "fp32"))
        dev = torch.device(device if device is not None else "cuda")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.store = ParamStore(dev, self.model)
        self.param_names = [n for n, _, _ in spec.trainable_specs(self.model)]
        for n in spec.reference_param_order(self.model):        # the reference's registration order (optimizer state_dict)
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
        self._staged = None
        self._param_list = [self.get_parameter(n) for n in self.param_names]
        self._grad_views = [self.store.grad_view(n) for n in self.param_names]
        # cfg key `zsg_direct_grads` (or the attribute): hand gradients to the optimiser by pointing param.grad at the
        # gradient arena instead of returning 161 tensors through autograd.  Off by default: torch's
        # DistributedDataParallel needs the AccumulateGrad hooks of the ordinary path.
        self.direct_grads = bool(get("zsg_direct_grads", False))
        self._anchor = torch.zeros((), device=dev, requires_grad=True)
        self._feat_sizes = torch.tensor([[s, s] for s in spec.LEVEL_SIZES], device=dev)
        self._num_f_out = torch.tensor([len(spec.LEVEL_SIZES)], device=dev)
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

    def set_compute_dtype(self, dtype):
        if dtype not in ("fp32", "bf16"):
            raise ValueError(f"zsg_b200: compute dtype {dtype!r} (choose 'fp32' or 'bf16')")
        self.compute_dtype = dtype
        return self

    # ------------------------------------------------------------------ forward
    T_BUCKET = 50            # dat_loader pads / truncates queries to 50 tokens (dat_loader.py:74, 'phrase_len')

    def engine_for(self, B, T=None):
        """The static launch program for batch size B.  Query length is NOT part of the shape: the LSTM buffers are sized
        for T_BUCKET tokens (the loader's maximum; longer batches round up to the next multiple) and the recurrences run
        to each sample's own length, so batches of any max_qlen share one engine and nothing is rebuilt per step."""
        Ta = max(1, -(-max(int(T or 1), 1) // self.T_BUCKET)) * self.T_BUCKET
        key = (B, Ta, self.compute_dtype)
        eng = self._engines.pop(key, None)
        if eng is None:
            while len(self._engines) >= 2:                        # bound memory: two programs (train batch / ragged last batch)
                self._engines.pop(next(iter(self._engines)))      # least recently used first
                torch.cuda.empty_cache()
            bufs = {k: v for k, v in self.bn_buffers.items()}
            eng = Engine(self.store, bufs, B, Ta, self.store.device, dtype=self.compute_dtype, do_norm=self.do_norm)
        self._engines[key] = eng                                  # most recently used last
        return eng

    def forward(self, inp: Dict[str, Any]):
        img, qvec, qlens = inp["img"], inp["qvec"], inp["qlens"]
        dev = self.store.device
        if img.device != dev:
            raise RuntimeError(f"zsg_b200: inputs must be on {dev} (got {img.device}); no CPU path exists")
        B = img.shape[0]
        # mdl.py:357: the reference synchronises here (qlens.max().item()); a host copy that came with the batch
        # (dat_loader.DevicePrefetcher adds `qlens_cpu`) avoids the read-back
        qlens_cpu = inp["qlens_cpu"] if "qlens_cpu" in inp else qlens.detach().cpu()
        max_qlen = check_qlens(qlens_cpu, qvec.shape[1])
        qvec = qvec[:, :max_qlen, :]
        # mdl.py:279-294, 307: h0 then c0 from the global CPU RNG, consumed in sorted-row order (309-319).  A batch that
        # comes from dat_loader.DevicePrefetcher(lstm_state=True) carries both draws (made when the batch was fetched, in
        # batch order) and the int32 lengths on the device already.
        self._staged = None
        h0 = c0 = inv = None
        if "_zsg_h0c0" in inp and "_zsg_lens" in inp:
            self._staged = (inp["_zsg_lens"], inp["_zsg_h0c0"])
        else:
            h0, c0, inv = draw_lstm_state(qlens_cpu)
        eng = self.engine_for(B, max(max_qlen, 1))
        params = [self._anchor] if self.direct_grads else self._param_list
        out = _ZSGNetFn.apply(self, eng, img.contiguous().float(), qvec.float(), qlens_cpu, inv, h0, c0, *params)
        if self.training:
            self._bn_n.add_(1)                                    # num_batches_tracked
        # constants of the output dict (mdl.py:391-395 builds them per forward with two small host-to-device copies; here
        # they are made once: a per-step copy on this stream would queue behind the prefetch of the next batch's image on
        # the DMA engine and hold up everything launched after it)
        return {"att_out": out[..., 4:], "bbx_out": out[..., :4], "feat_sizes": self._feat_sizes.clone(),
                "num_f_out": self._num_f_out.clone()}


def load_pretrained_encoder(net, path):
    """Pretrained trunk weights into backbone.encoder: a torchvision resnet50 state_dict for 'retina' (the reference builds
    tvm.resnet50(True), mdl.py:411) or vgg16_reducedfc.pth for 'ssd_vgg' (mdl.py:415-416: ssd_net.vgg.load_state_dict)."""
    sd = torch.load(path, map_location="cpu", weights_only=True)
    prefix = "backbone.encoder.vgg." if net.model == "ssd_vgg" else "backbone.encoder."
    own = net.state_dict()
    hit = {prefix + k: v for k, v in sd.items() if prefix + k in own and tuple(own[prefix + k].shape) == tuple(v.shape)}
    if not hit:
        raise ValueError(f"zsg_b200: {path} holds no tensor that fits {prefix}*")
    net.load_state_dict(hit, strict=False)
    return sorted(hit)


def get_default_net(num_anchors=1, cfg=None):
    """Same signature as mdl.py:406-422.  The reference starts from pretrained trunks (ImageNet resnet50 downloaded by
    torchvision, mdl.py:411; ./weights/vgg16_reducedfc.pth, mdl.py:415-416).  There is no network here: cfg['zsg_pretrained']
    (a state_dict file) or, for ssd_vgg, the reference's own path is loaded when present; otherwise the trunk keeps its
    random initialisation and a warning says so (accuracy of a from-scratch run is not the reference's)."""
    import os
    import warnings
    dev = None
    if cfg is not None and "device" in cfg and str(cfg["device"]).startswith("cuda"):
        dev = cfg["device"]
    net = ZSGNet(None, num_anchors, cfg=cfg, device=dev)
    path = cfg["zsg_pretrained"] if (cfg is not None and "zsg_pretrained" in cfg) else None
    if path is None and net.model == "ssd_vgg" and os.path.exists("./weights/vgg16_reducedfc.pth"):
        path = "./weights/vgg16_reducedfc.pth"
    if path:
        load_pretrained_encoder(net, path)
    elif not (cfg is not None and "zsg_quiet" in cfg and cfg["zsg_quiet"]):
        warnings.warn("zsg_b200.get_default_net: no pretrained trunk weights (cfg['zsg_pretrained'] not set; the reference "
                      "starts from ImageNet / vgg16_reducedfc weights): the encoder is randomly initialised", stacklevel=2)
    return net
