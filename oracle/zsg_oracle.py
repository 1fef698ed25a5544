"""CPU oracle: a restatement of the reference ZSGNet hot path (forward, loss, metric).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file; it
is used by tests/, by bench.py's `cpu_baseline` / `--impl reference` legs and by
`__graft_entry__.smoke()` as the checker, never as the thing measured or shipped.

Parity status: PINNED against outputs of the real reference executed in the
build container (tests/golden/make_golden.py imports /root/reference unmodified,
loads the same seeded weights and dumps tests/golden/*.npz; tests/test_oracle_golden.py
checks this file against those dumps).  The reference itself ships no tests or
golden vectors.

Third-party arithmetic restated here because its source is not under /root/reference:
torchvision.models.resnet50 (pinned torchvision=0.2.2, conda_env_zsg.yml:150; the
container has 0.26.0, same v1.5 bottleneck with the stride on the 3x3), called at
mdl.py:411 and consumed at mdl.py:149-156.

Every function cites the reference lines it follows.  Written functionally over a
flat state_dict with the reference's key names; autograd supplies the backward.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from . import synth

# --------------------------------------------------------------------------------------
# convolution arithmetic
# --------------------------------------------------------------------------------------
# "fp32": F.conv2d as the reference executes it (BASELINE configs 1-2).
# "bf16": what BASELINE configs 3-5 call "bf16 tensor-core convs", restated as the arithmetic the product computes: both
#   operands of every contraction rounded to bfloat16 (round-to-nearest-even), products exact, accumulation and output
#   in fp32 -- forward (x, w), data gradient (dy, w) and weight gradient (x, dy) alike; bias gradients from the unrounded
#   dy.  The LSTM (not a convolution) stays fp32, as in the product.
#   ResNet-50 trunk ("bf16 storage", what torch.autocast(bfloat16) also does to these tensors): every conv output (the
#   tensor a BatchNorm reads), the stem's pooled output and every block output is rounded to bfloat16 where it is stored
#   (store_act); gradients pass those points unrounded.  The block-internal gradients the engine stores are bfloat16 as
#   well (store_grad): the gradient behind the block's final ReLU (= the shortcut gradient) and the gradients of the two
#   inner BatchNorm+ReLU images; the gradient stream between blocks stays fp32.
#   The reference has no bf16 mode of its own; the closest thing it offers, torch.autocast(bfloat16) around the same
#   modules, additionally rounds every conv OUTPUT to bf16 (tests/test_bf16_network_gpu.py measures both against fp32).
CONV_MODE = ["fp32"]


class conv_mode:
    def __init__(self, mode):
        assert mode in ("fp32", "bf16"), mode
        self.mode = mode

    def __enter__(self):
        self.prev, CONV_MODE[0] = CONV_MODE[0], self.mode

    def __exit__(self, *a):
        CONV_MODE[0] = self.prev


def _rb(t):
    return t.bfloat16().to(t.dtype)


class _StoreBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _rb(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _StoreGradBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _rb(g)


def store_grad(x):
    """Identity whose GRADIENT is kept in bfloat16 by the bf16 engine; identity in fp32 mode."""
    return _StoreGradBf16.apply(x) if CONV_MODE[0] == "bf16" else x


def store_act(x):
    """A trunk activation as the bf16 engine keeps it in HBM (bfloat16 only); identity in fp32 mode."""
    return _StoreBf16.apply(x) if CONV_MODE[0] == "bf16" else x


class _Bf16OperandConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, stride, padding, dilation):
        xr, wr = _rb(x), _rb(w)
        ctx.save_for_backward(xr, wr)
        ctx.conf = (stride, padding, dilation, bias is not None, x.shape, w.shape)
        return F.conv2d(xr, wr, bias, stride, padding, dilation)

    @staticmethod
    def backward(ctx, dy):
        xr, wr = ctx.saved_tensors
        stride, padding, dilation, has_bias, xs, ws = ctx.conf
        dyr = _rb(dy)
        gx = torch.nn.grad.conv2d_input(xs, wr, dyr, stride, padding, dilation) if ctx.needs_input_grad[0] else None
        gw = torch.nn.grad.conv2d_weight(xr, ws, dyr, stride, padding, dilation) if ctx.needs_input_grad[1] else None
        gb = dy.sum((0, 2, 3)) if has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None, None


# ReLU with an externally supplied backward mask (stage-wise GPU tests only): where a pre-activation is within fp32
# rounding noise of zero, two correct implementations may disagree on relu'(x); feeding the oracle the decisions of the
# implementation under test compares the two backward passes on the SAME function instead of on neighbouring ones.
_RELU_MASKS = [None]


class relu_masks:
    """with relu_masks([m0, m1, ...]): the i-th relu() call of the block uses m_i (bool, shape of its input) in backward."""

    def __init__(self, masks):
        self.masks = list(masks)

    def __enter__(self):
        self.prev, _RELU_MASKS[0] = _RELU_MASKS[0], self.masks
        return self

    def __exit__(self, *a):
        assert not self.masks or a[0] is not None, f"{len(self.masks)} relu masks were not consumed"
        _RELU_MASKS[0] = self.prev


class _MaskedRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return x.clamp_min(0)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.saved_tensors[0].to(g.dtype), None


def relu(x):
    if _RELU_MASKS[0] is None:
        return F.relu(x)
    m = _RELU_MASKS[0].pop(0)
    assert m.shape == x.shape, (m.shape, x.shape)
    return _MaskedRelu.apply(x, m)


def conv2d(x, w, bias=None, stride=1, padding=0, dilation=1):
    if CONV_MODE[0] == "bf16":                              # every convolution (the LSTM is not one and stays fp32)
        return _Bf16OperandConv.apply(x, w, bias, stride, padding, dilation)
    return F.conv2d(x, w, bias, stride=stride, padding=padding, dilation=dilation)


# --------------------------------------------------------------------------------------
# anchors.py
# --------------------------------------------------------------------------------------


def cell_centres(n):
    """anchors.py:55-56,59-60: linspace(-1+1/n, 1-1/n, n) in float32, [0.] when n == 1."""
    if n > 1:
        return torch.linspace(-1 + 1 / n, 1 - 1 / n, n)
    return torch.tensor([0.0])


def make_grid(H, W):
    """anchors.py:47-63 with flatten=False: [H,W,2], channel 0 = row centre (y),
    channel 1 = column centre (x); float32."""
    ys = cell_centres(H).view(H, 1).expand(H, W)
    xs = cell_centres(W).view(1, W).expand(H, W)
    return torch.stack([ys, xs], dim=2).contiguous()


def make_anchors(sizes, ratios, scales):
    """anchors.py:66-87 + cthw2tlbr (11-15).  Returns [A,4] float64 (y1,x1,y2,x2).

    aspects are float64 (numpy scales, main_dist.py:29); the per-level factor
    [2/h, 2/w] is rounded to float32 FIRST (torch.tensor of python floats) and
    then promoted; cell centres are float32 promoted to float64 by the cat."""
    aspects = torch.tensor([[[s * np.sqrt(r), s * np.sqrt(1 / r)] for s in scales]
                            for r in ratios], dtype=torch.float64).view(-1, 2)
    out = []
    for (h, w) in sizes:
        h, w = int(h), int(w)
        factor = torch.tensor([2 / h, 2 / w], dtype=torch.float32).double()
        hw = aspects * factor                                   # [9,2] f64
        ctr = make_grid(h, w).view(-1, 2).double()              # [h*w,2]
        n, a = ctr.shape[0], hw.shape[0]
        cthw = torch.cat([ctr.unsqueeze(1).expand(n, a, 2), hw.unsqueeze(0).expand(n, a, 2)], dim=2)
        out.append(cthw.reshape(-1, 4))
    cthw = torch.cat(out, dim=0)
    return torch.cat([cthw[:, :2] - cthw[:, 2:] / 2, cthw[:, :2] + cthw[:, 2:] / 2], dim=1)


def iou_gt_vs_anchors(annot, anchs):
    """anchors.py:106-116 called as IoU_values(annot, anchs) (loss.py:76, anchors.py:161).

    annot [B,4] float32, anchs [A,4] float64 -> [B,A] float64.  Op order kept:
    corners max/min in f64; GT area = (y2-y1)*(x2-x1) computed in FLOAT32 and then
    promoted; anchor area in f64; union = (gt + anc) - inter; iou = inter/(union+1e-8)."""
    a64 = annot.double()
    tl = torch.max(a64[:, None, :2], anchs[None, :, :2])
    br = torch.min(a64[:, None, 2:], anchs[None, :, 2:])
    sz = torch.clamp(br - tl, min=0)
    inter = sz[..., 0] * sz[..., 1]
    gt_hw = annot[:, 2:] - annot[:, :2]                          # float32
    gt_area = (gt_hw[:, 0] * gt_hw[:, 1]).double()               # f32 product, then promoted
    an_hw = anchs[:, 2:] - anchs[:, :2]
    an_area = an_hw[:, 0] * an_hw[:, 1]
    union = gt_area[:, None] + an_area[None, :] - inter
    return inter / (union + 1e-8)


def iou_pairwise_diag(boxes, annot):
    """evaluator.py:115: diag(IoU_values(best_boxes, annot)); boxes f64 [B,4], annot f32 [B,4].
    Here `boxes` play the role of anchors.py's first argument (areas in their own dtype)."""
    a64 = annot.double()
    tl = torch.max(boxes[:, :2], a64[:, :2])
    br = torch.min(boxes[:, 2:], a64[:, 2:])
    sz = torch.clamp(br - tl, min=0)
    inter = sz[:, 0] * sz[:, 1]
    b_hw = boxes[:, 2:] - boxes[:, :2]
    b_area = b_hw[:, 0] * b_hw[:, 1]
    g_hw = annot[:, 2:] - annot[:, :2]
    g_area = (g_hw[:, 0] * g_hw[:, 1]).double()
    union = b_area + g_area - inter
    return inter / (union + 1e-8)


def gt_reg_targets(anchs, annot):
    """anchors.py:168-179.  GT centre/size in float32 (tlbr2cthw on the f32 boxes),
    anchors in f64; t_c = (c_gt - c_a)/(hw_a+1e-8); t_hw = log(hw_gt/(hw_a+1e-8))."""
    g_c = (annot[:, :2] + annot[:, 2:]) / 2                      # f32
    g_hw = annot[:, 2:] - annot[:, :2]                           # f32
    a_c = (anchs[:, :2] + anchs[:, 2:]) / 2
    a_hw = anchs[:, 2:] - anchs[:, :2]
    trc = (g_c.double()[:, None, :] - a_c[None]) / (a_hw[None] + 1e-8)
    thw = torch.log(g_hw.double()[:, None, :] / (a_hw[None] + 1e-8))
    return torch.cat([trc, thw], dim=2)


def decode_boxes(anchs, reg):
    """anchors.py:182-197: c = hw_a*t_c + c_a ; hw = exp(t_hw)*hw_a ; back to tlbr.
    The exp is taken in FLOAT32 (anchors.py:193-194: b2 stays f32) and then promoted."""
    a_c = (anchs[:, :2] + anchs[:, 2:]) / 2
    a_hw = anchs[:, 2:] - anchs[:, :2]
    c = a_hw * reg[..., :2] + a_c
    hw = torch.exp(reg[..., 2:]) * a_hw
    return torch.cat([c - hw / 2, c + hw / 2], dim=-1)


# --------------------------------------------------------------------------------------
# loss.py
# --------------------------------------------------------------------------------------


def match(annot, anchs, thr=0.6):
    """loss.py:73-87 with use_multi: pos = (iou > thr) | onehot(argmax_a iou); the argmax
    takes the FIRST maximal index (CPU torch.max)."""
    iou = iou_gt_vs_anchors(annot, anchs)
    top1 = iou.max(1)[1]
    pos = iou > thr
    pos[torch.arange(annot.shape[0], device=annot.device), top1] = True
    return pos, top1, iou


def zsg_loss(att_out, bbx_out, annot, anchs, cfg=None):
    """loss.py:43-143 (focal + smooth-L1; alpha weights the NEGATIVES, loss.py:115-116)."""
    cfg = cfg or synth.default_cfg()
    pos, top1, _ = match(annot, anchs, cfg["matching_threshold"])
    if not cfg["use_multi"]:
        pos = torch.zeros_like(pos)
        pos[torch.arange(annot.shape[0], device=annot.device), top1] = True
    tgt = gt_reg_targets(anchs, annot)
    box_l = F.smooth_l1_loss(bbx_out.double(), tgt, reduction="none")       # f32 vs f64 -> f64
    posf = pos.float()
    box_rows = (box_l.sum(dim=2) * posf).sum(dim=1) / pos.sum(dim=-1).float()
    box_loss = box_rows.mean()
    x = att_out.squeeze(-1)
    p = torch.sigmoid(x)
    w = posf * (1 - p) + (1 - posf) * p
    al = (1 - posf) * cfg["alpha"] + posf * (1 - cfg["alpha"])
    w = (w.pow(cfg["gamma"]) * al).detach()
    cls = F.binary_cross_entropy_with_logits(x, posf, weight=w, reduction="none")
    cls_loss = cls.sum() / pos.sum()
    if torch.isnan(box_loss) or torch.isnan(cls_loss):                       # loss.py:128-133: constants without gradient
        box_loss = (box_loss.new_ones(box_loss.shape) * 0.01).requires_grad_(True)
        cls_loss = cls_loss.new_ones(cls_loss.shape).requires_grad_(True)
    loss = cfg["lamb_reg"] * box_loss + cls_loss
    return {"loss": loss, "cls_ls": cls_loss, "box_ls": box_loss, "pos": pos, "top1": top1}


# --------------------------------------------------------------------------------------
# evaluator.py
# --------------------------------------------------------------------------------------


def evaluate(att_out, bbx_out, annot, img_size, anchs, cfg=None):
    """evaluator.py:48-117.  Acc = mean(IoU(decoded box of argmax sigmoid(att), GT) >= 0.5);
    MaxPos = same with the IoU-argmax anchor; pred_boxes in pixel x1y1x2y2 (f64)."""
    cfg = cfg or synth.default_cfg()
    B = annot.shape[0]
    score, best = torch.sigmoid(att_out).squeeze(-1).max(1)
    top1 = iou_gt_vs_anchors(annot, anchs).max(1)[1]
    boxes = decode_boxes(anchs, bbx_out)                                     # [B,A,4] f64
    ar = torch.arange(B, device=annot.device)
    thr = cfg["acc_iou_threshold"]
    maxpos = (iou_pairwise_diag(boxes[ar, top1], annot) >= thr).float().mean()
    pb = boxes[ar, best]
    acc = (iou_pairwise_diag(pb, annot) >= thr).float().mean()
    px = (pb + 1) / 2
    px = torch.cat([img_size * px[:, :2], img_size * px[:, 2:]], dim=1)       # y1x1y2x2 pixels
    px = px[:, [1, 0, 3, 2]]                                                 # -> x1y1x2y2
    return {"Acc": acc, "MaxPos": maxpos, "idxs_best": best, "pred_boxes": px,
            "pred_scores": score, "top1": top1}


# --------------------------------------------------------------------------------------
# mdl.py / fpn_resnet.py / torchvision resnet50
# --------------------------------------------------------------------------------------


class BNState:
    """Collects running-stat updates so callers can compare them too."""

    def __init__(self, sd, training=True):
        self.sd, self.training = sd, training

    def __call__(self, x, prefix):
        sd = self.sd
        rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
        if self.training:
            sd[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
        return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"],
                            training=self.training, momentum=0.1, eps=1e-5)


def stem(sd, img, bn):
    """mdl.py:149-152: conv7x7/2 -> BN -> ReLU -> maxpool3x3/2."""
    e = "backbone.encoder."
    x = store_act(conv2d(img, sd[e + "conv1.weight"], None, stride=2, padding=3))
    x = relu(bn(x, e + "bn1"))
    return store_act(F.max_pool2d(x, 3, 2, 1))


def bottleneck(sd, x, p, s, bn):
    """torchvision Bottleneck (v1.5: the stride sits on the 3x3): 1x1 -> BN, ReLU -> 3x3(stride s) -> BN, ReLU -> 1x1 (x4) -> BN
    -> + identity (or 1x1/s conv + BN when the block has a `downsample`) -> ReLU.  p = 'backbone.encoder.layerL.B.'."""
    idt = x
    y = store_grad(relu(bn(store_act(conv2d(x, sd[p + "conv1.weight"])), p + "bn1")))
    y = store_grad(relu(bn(store_act(conv2d(y, sd[p + "conv2.weight"], None, stride=s, padding=1)), p + "bn2")))
    y = bn(store_act(conv2d(y, sd[p + "conv3.weight"])), p + "bn3")
    if p + "downsample.0.weight" in sd:
        idt = bn(store_act(conv2d(x, sd[p + "downsample.0.weight"], None, stride=s)), p + "downsample.1")
    return store_act(relu(store_grad(y + idt)))


def resnet50_c3c4c5(sd, img, bn):
    """mdl.py:148-156 over torchvision's ResNet-50 v1.5: stem -> 4 stages of bottlenecks.  Returns C3, C4, C5."""
    e = "backbone.encoder."
    x = stem(sd, img, bn)
    feats = []
    for li, (nblk, width, stride) in enumerate(synth.RESNET_LAYERS, start=1):
        for b in range(nblk):
            x = bottleneck(sd, x, f"{e}layer{li}.{b}.", stride if b == 0 else 1, bn)
        feats.append(x)
    return feats[1], feats[2], feats[3]


def fpn(sd, c3, c4, c5):
    """fpn_resnet.py:154-178 for a 300x300 input: [p3,p4,p5,p6,p7,p8]."""
    f = "backbone.fpn."

    def cv(name, x, stride=1, pad=0):
        return conv2d(x, sd[f + name + ".weight"], sd[f + name + ".bias"], stride=stride, padding=pad)

    p51 = cv("P5_1", c5)
    p5 = cv("P5_2", p51, pad=1)
    p41 = cv("P4_1", c4) + F.interpolate(p51, size=c4.shape[2:])
    p4 = cv("P4_2", p41, pad=1)
    p31 = cv("P3_1", c3) + F.interpolate(p41, size=c3.shape[2:])
    p3 = cv("P3_2", p31, pad=1)
    p6 = cv("P6", c5, stride=2, pad=1)
    p7 = cv("P7_2", relu(p6), stride=2, pad=1)
    p8 = F.adaptive_avg_pool2d(p7, 1)
    return [p3, p4, p5, p6, p7, p8]


def ssd_vgg_feats(sd, img):
    """SSDBackBone.encode_feats (mdl.py:162-168) = SSD.forward (ssd_vgg.py:54-102) for a 300x300 input:
    VGG-16 up to relu(conv4_3) (vgg[0:23]) -> x / ||x||_2 over channels (ssd_vgg.py:80; no eps, no learned
    scale) = source 0; the rest of VGG with pool5 (3x3/1), the dilated conv6 and conv7 = source 1; eight extra
    convs, ReLU after each, every second one a source (92-95); fproj1..3 (1x1 to 256) on the first three
    sources, the last three are used as they are (97-98).  Sizes 38, 19, 10, 5, 3, 1."""
    e = "backbone.encoder."

    def run(x, lo, hi):
        for i, L in enumerate(synth.vgg_layers()):
            if not lo <= i < hi:
                continue
            if L[0] == "conv":
                x = conv2d(x, sd[f"{e}vgg.{i}.weight"], sd[f"{e}vgg.{i}.bias"], padding=L[4], dilation=L[5])
            elif L[0] == "relu":
                x = relu(x)
            else:
                x = F.max_pool2d(x, L[1], L[2], L[3], ceil_mode=L[4])
        return x
    x = run(img, 0, 23)
    sources = [x / x.norm(dim=1, keepdim=True)]
    x = run(x, 23, len(synth.vgg_layers()))
    sources.append(x)
    for i, (_, _, _, stride, pad) in enumerate(synth.VGG_EXTRAS):
        x = relu(conv2d(x, sd[f"{e}extras.{i}.weight"], sd[f"{e}extras.{i}.bias"], stride=stride, padding=pad))
        if i % 2 == 1:
            sources.append(x)
    proj = [conv2d(sources[j], sd[f"{e}fproj{j + 1}.weight"], sd[f"{e}fproj{j + 1}.bias"]) for j in range(3)]
    return proj + sources[3:]


def lstm_query(sd, qvec, qlens, h0, c0):
    """mdl.py:296-336.  One-layer bi-LSTM (gate order i,f,g,o; two bias vectors).  The
    returned vector for sample b is lstm_out[len_b-1, b, :]: the forward direction after
    len_b tokens and the backward direction at position len_b-1, i.e. after ONE step from
    its initial state on token len_b-1 (mdl.py:326-328).  h0,c0: [2,B,128] (mdl.py:279-294).
    Sort/pack/unsort (309-316, 330-331) only permutes rows, and the initial states are
    consumed in SORTED order: sorted row j uses h0[:, j]."""
    B, T, E = qvec.shape
    H = 128
    lens = qlens.long()
    _, perm = qlens.sort(0, descending=True)                     # mdl.py:309
    out = qvec.new_zeros(B, 2 * H)
    for j in range(B):
        b = int(perm[j])
        L = int(lens[b])
        h, c = h0[0, j], c0[0, j]
        for t in range(L):
            g = (sd["lstm.weight_ih_l0"] @ qvec[b, t] + sd["lstm.bias_ih_l0"]
                 + sd["lstm.weight_hh_l0"] @ h + sd["lstm.bias_hh_l0"])
            i, f, gg, o = g[:H].sigmoid(), g[H:2 * H].sigmoid(), g[2 * H:3 * H].tanh(), g[3 * H:].sigmoid()
            c = f * c + i * gg
            h = o * c.tanh()
        hr, cr = h0[1, j], c0[1, j]
        g = (sd["lstm.weight_ih_l0_reverse"] @ qvec[b, L - 1] + sd["lstm.bias_ih_l0_reverse"]
             + sd["lstm.weight_hh_l0_reverse"] @ hr + sd["lstm.bias_hh_l0_reverse"])
        i, f, gg, o = g[:H].sigmoid(), g[H:2 * H].sigmoid(), g[2 * H:3 * H].tanh(), g[3 * H:].sigmoid()
        cr = f * cr + i * gg
        hr = o * cr.tanh()
        out[b] = torch.cat([h, hr])
    return out


def lstm_query_batched(sd, qvec, qlens, h0, c0):
    """Same arithmetic as lstm_query, vectorised over the batch (used by the timed CPU
    baseline so that it is not dominated by python loops)."""
    B, T, E = qvec.shape
    H = 128
    lens = qlens.long()
    _, perm = qlens.sort(0, descending=True)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(B, device=perm.device)
    h, c = h0[0][inv], c0[0][inv]                                 # sample b uses sorted row inv[b]
    gx = qvec @ sd["lstm.weight_ih_l0"].t() + sd["lstm.bias_ih_l0"] + sd["lstm.bias_hh_l0"]
    for t in range(int(lens.max())):
        g = gx[:, t] + h @ sd["lstm.weight_hh_l0"].t()
        i, f, gg, o = g[:, :H].sigmoid(), g[:, H:2 * H].sigmoid(), g[:, 2 * H:3 * H].tanh(), g[:, 3 * H:].sigmoid()
        c2 = f * c + i * gg
        h2 = o * c2.tanh()
        live = (t < lens).unsqueeze(1)
        c, h = torch.where(live, c2, c), torch.where(live, h2, h)
    xl = qvec[torch.arange(B, device=lens.device), lens - 1]
    hr, cr = h0[1][inv], c0[1][inv]
    g = (xl @ sd["lstm.weight_ih_l0_reverse"].t() + sd["lstm.bias_ih_l0_reverse"]
         + hr @ sd["lstm.weight_hh_l0_reverse"].t() + sd["lstm.bias_hh_l0_reverse"])
    i, f, gg, o = g[:, :H].sigmoid(), g[:, H:2 * H].sigmoid(), g[:, 2 * H:3 * H].tanh(), g[:, 3 * H:].sigmoid()
    cr = f * cr + i * gg
    hr = o * cr.tanh()
    return torch.cat([h, hr], dim=1)


def fuse_and_head(sd, feats, lang):
    """mdl.py:69-104 (channel order feat | lang | grid_y | grid_x), 235-244 (6 convs,
    ReLU after the first five), 246-254 + 379-382 (NHWC, view B x HW*9 x 5, cat levels)."""
    outs = []
    B = lang.shape[0]
    for x in feats:
        H, W = x.shape[2], x.shape[3]
        grid = make_grid(H, W).to(x.device).permute(2, 0, 1).unsqueeze(0).expand(B, 2, H, W)
        we = lang.view(B, -1, 1, 1).expand(B, lang.shape[1], H, W)
        y = torch.cat([x, we, grid], dim=1)
        for i in range(5):
            y = relu(conv2d(y, sd[f"att_reg_box.{i}.0.weight"], sd[f"att_reg_box.{i}.0.bias"], padding=1))
        y = conv2d(y, sd["att_reg_box.5.weight"], sd["att_reg_box.5.bias"], padding=1)
        outs.append(y.permute(0, 2, 3, 1).reshape(B, -1, 5))
    out = torch.cat(outs, dim=1)
    return out[..., 4:5], out[..., :4]


def draw_h0c0(B):
    """mdl.py:279-294: two CPU draws from the global RNG per forward, h0 then c0."""
    h0 = torch.randn(2, B, 128)
    c0 = torch.randn(2, B, 128)
    return h0, c0


def zsgnet_forward(sd, batch, training=True, h0c0=None, batched_lstm=True, return_inter=False, do_norm=False):
    """mdl.py:338-403.  The trunk is chosen by the keys present: SSD-VGG (mdl.py:413-418) or ResNet-50+FPN.
    do_norm: cfg do_norm (mdl.py:118-130): every feature pixel and the language vector are L2-normalised before the fusion."""
    img, qvec, qlens = batch["img"], batch["qvec"], batch["qlens"]
    max_qlen = int(qlens.max().item())
    qvec = qvec[:, :max_qlen].contiguous()
    h0, c0 = h0c0 if h0c0 is not None else draw_h0c0(img.shape[0])
    h0, c0 = h0.to(img.device), c0.to(img.device)          # drawn on the CPU, then moved (mdl.py:291-292)
    lang = (lstm_query_batched if batched_lstm else lstm_query)(sd, qvec, qlens, h0, c0)
    c3 = c4 = c5 = None
    if "backbone.encoder.vgg.0.weight" in sd:
        feats = ssd_vgg_feats(sd, img)
    else:
        bn = BNState(sd, training)
        c3, c4, c5 = resnet50_c3c4c5(sd, img, bn)
        feats = fpn(sd, c3, c4, c5)
    lang_raw, feats_raw = lang, feats
    if do_norm:                                              # mdl.py:118-123 and 128-130
        feats = [f / f.norm(dim=1).unsqueeze(1).expand(*f.shape) for f in feats]
        lang = lang / lang.norm(dim=1).unsqueeze(1).expand(*lang.shape)
    att, bbx = fuse_and_head(sd, feats, lang)
    out = {"att_out": att, "bbx_out": bbx,
           "feat_sizes": torch.tensor([[f.shape[2], f.shape[3]] for f in feats]),
           "num_f_out": torch.tensor([len(feats)])}
    if return_inter:
        out["_inter"] = {"lang": lang_raw, "c3": c3, "c4": c4, "c5": c5, "feats": feats_raw}
    return out


_ANCH_CACHE = {}


def default_anchors():
    if "a" not in _ANCH_CACHE:
        ratios, scales = synth.ratios_scales()
        _ANCH_CACHE["a"] = make_anchors([(s, s) for s in synth.LEVEL_SIZES], ratios, scales)
    return _ANCH_CACHE["a"]


def trainable_keys(sd):
    return [k for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k]


def train_step(sd, batch, opt_state=None, lr=1e-4, seed=None, do_adam=True, do_norm=False):
    """One iteration of utils.py:405-414: forward, loss, backward, Adam(betas 0.9,0.99), metric.
    `sd` is updated in place.  Returns losses, metric and the gradients."""
    if seed is not None:
        torch.manual_seed(seed)
    keys = trainable_keys(sd)
    for k in keys:
        sd[k] = sd[k].detach().requires_grad_(True)
    out = zsgnet_forward(sd, batch, training=True, do_norm=do_norm)
    anchs = default_anchors().to(batch["img"].device)
    ls = zsg_loss(out["att_out"], out["bbx_out"], batch["annot"], anchs)
    ls["loss"].mean().backward()
    grads = {k: sd[k].grad for k in keys}
    with torch.no_grad():
        if do_adam:
            if opt_state is None:
                opt_state = {}
            adam_update(sd, grads, opt_state, lr)
        met = evaluate(out["att_out"].detach(), out["bbx_out"].detach(), batch["annot"],
                       batch["img_size"], anchs)
    for k in keys:
        sd[k] = sd[k].detach()
    return ls, met, grads, out, opt_state


def adam_update(sd, grads, state, lr, b1=0.9, b2=0.99, eps=1e-8):
    """torch.optim.Adam as configured at main_dist.py:50 (no weight decay, no amsgrad)."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    for k, g in grads.items():
        if g is None:
            continue
        m = state.setdefault("m." + k, torch.zeros_like(g))
        v = state.setdefault("v." + k, torch.zeros_like(g))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
        sd[k].data.addcdiv_(m, denom, value=-lr / (1 - b1 ** t))
