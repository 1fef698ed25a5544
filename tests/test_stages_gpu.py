"""Stage-wise parity: one stage of the launch program at a time -- stem, a bottleneck block, the FPN,
the fusion + shared head -- forward AND backward, each FED WITH THE ORACLE'S TENSORS at the bench's per-sample shapes
(300x300, B = 4), so that the fp32 drift of the 50-layer train-mode-BatchNorm network cannot compound and each stage is held
to the 1e-4 bar of BASELINE.json on its own (2e-5 where the stage has no BatchNorm).

The engine is one static program; a stage is run by overwriting its input buffers (and their GEMM operand images) with the
oracle's activations and launching only the stage's slice of the program (Engine.fwd_marks / bwd_marks).

ReLU decisions.  Where a pre-activation lies within fp32 rounding noise of zero, two correct forward passes disagree on
relu'(x); with the random +-1-sized output gradients used here, a handful of such flips among the 5.8 M activations of a
layer1 block already moves a bias gradient by 1e-3 (measured: forward 1e-6, backward 1e-6 in blocks without a flip, 2e-3 in
blocks with one).  The backward comparison therefore runs the oracle with the ENGINE's ReLU masks (oracle relu_masks): same
function on both sides, every kernel of the backward still checked end to end.  The number of disagreeing decisions is
printed.

dtype 'bf16' runs the same stages on the bf16 operand path against the oracle in conv_mode('bf16') (same arithmetic, see
oracle/zsg_oracle.py): what is left is summation order plus values that sit on a bf16 rounding boundary and round the other
way (one bf16 ulp = 0.4 %, on a ~1e-4 fraction of a layer's elements), hence 2e-3 instead of 1e-4."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
B = 4


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().view(-1, t.shape[1])


def rows_to_nchw(buf, h, w):
    """engine buffer [B*h*w, C] (NHWC rows) -> CPU NCHW"""
    return buf.view(B, h, w, -1).permute(0, 3, 1, 2).contiguous().cpu()


def bn_relu_mask(x, bn, h, w):
    """the engine's decision relu'(x * scale + shift) of a BatchNorm + ReLU on load (sign of the fused multiply-add)"""
    return rows_to_nchw((x.double() * bn.scale.double() + bn.shift.double()) > 0, h, w)


def err(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class Stage:
    def __init__(self, dtype):
        import zsg_b200  # noqa: F401
        from zsg_b200 import mdl, ops
        from oracle import synth, zsg_oracle as zo
        self.ops, self.synth, self.zo, self.dtype = ops, synth, zo, dtype
        cfg = synth.default_cfg()
        cfg["device"], cfg["zsg_dtype"] = "cuda", dtype
        self.net = mdl.get_default_net(num_anchors=9, cfg=cfg)
        self.sd = synth.make_state_dict(0)
        self.net.load_state_dict(self.sd, strict=True)
        self.net.train()
        self.batch = synth.make_batch(B, seed=7)
        torch.manual_seed(7)
        self.net({k: v.cuda() for k, v in self.batch.items()})         # builds the engine, weight images, row tables
        self.eng = self.net.engine_for(B, 20)
        self.eng.use_graphs = False
        # the oracle's activations at every stage boundary (same arithmetic as the engine under test)
        torch.manual_seed(7)
        with zo.conv_mode(dtype), torch.no_grad():
            bn = zo.BNState(dict(self.sd), True)
            x = zo.stem(bn.sd, self.batch["img"], bn)
            self.acts = {"x0": x}
            for li, (nblk, width, stride) in enumerate(synth.RESNET_LAYERS, start=1):
                for b in range(nblk):
                    self.acts[f"layer{li}.{b}"] = x                      # INPUT of the block
                    x = zo.bottleneck(bn.sd, x, f"backbone.encoder.layer{li}.{b}.", stride if b == 0 else 1, bn)
                self.acts[f"c{li + 1}"] = x
            torch.manual_seed(7)
            out = zo.zsgnet_forward(dict(self.sd), self.batch, training=True, return_inter=True)
            self.inter = out["_inter"]
        # bf16: one bfloat16 ulp is 2^-8 = 3.9e-3 relative.  With bf16 storage of the trunk (block outputs and conv outputs are
        # bfloat16 tensors) values that sit on a rounding boundary round the other way on either side, and the engine's
        # BatchNorm statistics come from the fp32 accumulators while the oracle's come from the stored (rounded) tensor:
        # measured 2.2e-3 .. 3.0e-3 rms per block forward and up to 4.1e-3 on a weight gradient (layer1.0.conv1), i.e. about
        # one ulp
        self.tol_bn = 1e-4 if dtype == "fp32" else 4e-3
        self.tol = 2e-5 if dtype == "fp32" else 2e-3
        self.tol_g = 1e-4 if dtype == "fp32" else 6e-3

    def sdg(self, keys):
        """leaf copies of the oracle weights that take part in a stage"""
        sd = dict(self.sd)
        for k in keys:
            sd[k] = sd[k].clone().requires_grad_(True)
        return sd

    def put(self, buf, t, image_rows_c=None):
        """oracle tensor (NCHW) -> engine buffer (NHWC rows); regenerates the GEMM operand image of the buffer if it has one"""
        buf.view(-1)[: t.numel()].copy_(nhwc(t).cuda().view(-1))
        for key, (z, img) in self.eng._operand_cache.items():
            if key[0] == buf.data_ptr() and key[3] == id(None) and not key[4] and img is not buf:   # bf16 storage: own image
                self.ops.split_act(buf, img, key[1], key[2])

    def fwd(self, label):
        eng = self.eng
        eng._f64_pool[: eng._f64_used].zero_()
        eng._bn_tickets.zero_()
        lo, hi = eng.fwd_marks[label]
        for item in eng.fwd[lo:hi]:
            item[1]()
        torch.cuda.synchronize()

    def bwd(self, label, fill):
        eng = self.eng
        eng._backward_prologue()                                         # zero gradient arena, transposed weight images
        fill()
        eng._run_bwd_ops(*eng.bwd_marks[label])
        torch.cuda.synchronize()

    def grad(self, name):
        return self.net.store.grad_view(name)


@pytest.fixture(scope="module", params=["fp32", "bf16"])
def st(request):
    assert torch.cuda.is_available()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return Stage(request.param)


def test_stem_stage(st):
    zo, eng = st.zo, st.eng
    e = "backbone.encoder."
    keys = [e + "conv1.weight", e + "bn1.weight", e + "bn1.bias"]
    sd = st.sdg(keys)
    st.ops.nchw_to_nhwc4(st.batch["img"].cuda(), eng._img4)
    st.fwd("stem")
    with zo.conv_mode(st.dtype), zo.relu_masks([bn_relu_mask(eng.dbg["c1"], eng.bns[0], 150, 150)]):
        y = zo.stem(sd, st.batch["img"], zo.BNState(sd, True))
        G = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
        (y * G).sum().backward()
    assert err(eng.dbg["x0"], nhwc(y)) < st.tol_bn
    st.bwd("stem", lambda: st.put(eng.dbg["g_x0"], G))
    for k in keys:
        assert err(st.grad(k), sd[k].grad) < st.tol_g, k


@pytest.mark.parametrize("label", ["layer1.0", "layer1.2", "layer2.0", "layer3.1", "layer4.0"])
def test_bottleneck_stage(st, label):
    """1x1 -> 3x3 (stride) -> 1x1 with three (four) train-mode BatchNorms, shortcut, ReLU: forward, data gradient, every
    parameter gradient.  layer1.0 has the projection shortcut at stride 1, layer2.0 / layer4.0 at stride 2 (parity-class data
    gradients), layer1.2 / layer3.1 the identity shortcut."""
    zo, eng = st.zo, st.eng
    blk = next(b for b in eng.dbg["blocks"] if b["label"] == label)
    p = f"backbone.encoder.{label}."
    keys = [k for k in st.sd if k.startswith(p) and "running" not in k and "num_batches" not in k]
    sd = st.sdg(keys)
    x = st.acts[label].clone().requires_grad_(True)
    stride = 2 if (label.endswith(".0") and not label.startswith("layer1")) else 1
    st.put(blk["inp"], x)
    st.fwd(label)
    h, ho = blk["h"], blk["ho"]
    masks = [bn_relu_mask(blk["r1"], blk["bnA"], h, h), bn_relu_mask(blk["r2"], blk["bnB"], ho, ho),
             rows_to_nchw(blk["out"] > 0, ho, ho)]
    with zo.conv_mode(st.dtype), zo.relu_masks(masks):
        y = zo.bottleneck(sd, x, p, stride, zo.BNState(sd, True))
        G = torch.randn(y.shape, generator=torch.Generator().manual_seed(2))
        (y * G).sum().backward()
    print(f"{st.dtype} {label}: forward {err(blk['out'], nhwc(y)):.2e}; {int(((y > 0) != masks[2]).sum())} of {y.numel()} final "
          "ReLU decisions differ between engine and oracle")
    assert err(blk["out"], nhwc(y)) < st.tol_bn

    def fill():
        st.put(blk["g_out"], G)
        if blk["acc_in"]:
            blk["g_in"].zero_()
    st.bwd(label, fill)
    assert err(blk["g_in"], nhwc(x.grad)) < st.tol_g
    worst = {k: err(st.grad(k), sd[k].grad) for k in keys}
    assert max(worst.values()) < st.tol_g, worst


def test_fpn_stage(st):
    zo, eng = st.zo, st.eng
    keys = [k for k in st.sd if k.startswith("backbone.fpn.")]
    sd = st.sdg(keys)
    cs = [st.inter[k].clone().requires_grad_(True) for k in ("c3", "c4", "c5")]
    for name, c in zip(("c3", "c4", "c5"), cs):
        st.put(eng.dbg[name], c)
    st.fwd("fpn")
    with zo.conv_mode(st.dtype), zo.relu_masks([rows_to_nchw(eng.dbg["fl"][3] > 0, 5, 5)]):      # relu(p6) in front of P7_2
        feats = zo.fpn(sd, *cs)
        gen = torch.Generator().manual_seed(3)
        Gs = [torch.randn(f.shape, generator=gen) for f in feats]
        sum((f * g).sum() for f, g in zip(feats, Gs)).backward()
    for i, f in enumerate(feats):
        assert err(eng.dbg["fl"][i], nhwc(f)) < st.tol, i
    st.bwd("fpn", lambda: [st.put(eng.dbg["dfl"][i], g) for i, g in enumerate(Gs)])
    for name, c in zip(("g_c3", "g_c4", "g_c5"), cs):
        assert err(eng.dbg[name], nhwc(c.grad)) < st.tol_g, name
    worst = {k: err(st.grad(k), sd[k].grad) for k in keys}
    assert max(worst.values()) < st.tol_g, worst


def test_fusion_head_stage(st):
    """concat_we + create_grid tiling (mdl.py:69-104), six shared convs over six levels, [B, A, 5] packing (mdl.py:246-254):
    forward, gradient w.r.t. the six feature maps and the language vector, every head parameter gradient."""
    zo, eng = st.zo, st.eng
    keys = [k for k in st.sd if k.startswith("att_reg_box.")]
    sd = st.sdg(keys)
    feats = [f.clone().requires_grad_(True) for f in st.inter["feats"]]
    lang = st.inter["lang"].clone().requires_grad_(True)
    for i, f in reversed(list(enumerate(feats))):          # level 0 last: its put() regenerates the operand image of the whole
        st.put(eng.dbg["fl"][i], f)                        # level-major matrix the split first head conv reads
    eng.lang.copy_(lang.detach().cuda())
    st.fwd("head")
    lvl_off, sizes = eng.dbg["lvl_off"], st.synth.LEVEL_SIZES
    masks = [rows_to_nchw(eng.dbg["hs"][i][lvl_off[lv]:lvl_off[lv + 1]] > 0, s, s)          # oracle order: level-major, 5 ReLUs each
             for lv, s in enumerate(sizes) for i in range(5)]
    with zo.conv_mode(st.dtype), zo.relu_masks(masks):
        att, bbx = zo.fuse_and_head(sd, feats, lang)
        packed = torch.cat([bbx, att], dim=2)
        G = torch.randn(packed.shape, generator=torch.Generator().manual_seed(4))
        (packed * G).sum().backward()
    assert err(eng.out, packed) < st.tol
    st.bwd("head", lambda: eng.d_out.copy_(G.cuda()))
    for i, f in enumerate(feats):
        assert err(eng.dbg["dfl"][i], nhwc(f.grad)) < st.tol_g, i
    assert err(eng.dbg["dlang"], lang.grad) < st.tol_g
    worst = {k: err(st.grad(k), sd[k].grad) for k in keys}
    assert max(worst.values()) < st.tol_g, worst
