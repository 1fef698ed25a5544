"""GPU parity tests of the individual C-ABI entry points (run on the B200 box with -m gpu).

Checkers: the CPU oracle (oracle/zsg_oracle.py) for matching / loss / metric / LSTM, and plain
fp32 PyTorch ops (TF32 disabled) for the convolution, BatchNorm and pooling kernels.
Tolerances: integer/index outputs bit-exact; fp32 outputs 1e-4 relative (BASELINE.json)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_npz

gpu = pytest.mark.gpu
pytestmark = gpu

RTOL = 1e-4


@pytest.fixture(scope="module")
def zsg():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import zsg_b200
    from zsg_b200 import ops, geometry, _lib
    assert _lib.load().zsg_device_supported() == 1
    return ops, geometry


def dev(t):
    return t.cuda().contiguous()


def rel_err(a, b):
    a, b = a.detach().cpu().double().flatten(), b.detach().cpu().double().flatten()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------------------- loss / metric
def run_loss(ops, att, bbx, annot, anchs, packed=False):
    B, A = att.shape[0], att.shape[1]
    d = att.device
    if packed:
        buf = torch.cat([bbx, att], dim=2).contiguous()          # [B,A,5]: box 0..3, logit 4
        att_p, reg_p, sa, sr = buf.view(-1)[4:], buf, 5, 5
        dbuf = torch.empty_like(buf)
        datt_p, dreg_p = dbuf.view(-1)[4:], dbuf
    else:
        att_p, reg_p, sa, sr = att.contiguous(), bbx.contiguous(), 1, 4
        datt, dreg = torch.empty(B, A, device=d), torch.empty(B, A, 4, device=d)
        datt_p, dreg_p = datt, dreg
    losses = torch.empty(3, dtype=torch.float64, device=d)
    top1 = torch.empty(B, dtype=torch.int64, device=d)
    pos = torch.empty(B, A, dtype=torch.uint8, device=d)
    ws = ops.match_loss_workspace(B, d)
    ops.match_loss(att_p, sa, reg_p, sr, annot, anchs, B, A, 0.6, 0.25, 2.0, 1.0, True, losses, datt_p, sa, dreg_p, sr,
                   top1, pos, ws)
    torch.cuda.synchronize()
    if packed:
        datt, dreg = dbuf[..., 4], dbuf[..., :4]
    return losses.cpu(), datt.cpu(), dreg.cpu(), top1.cpu(), pos.cpu().bool()


@pytest.mark.parametrize("name", ["rand8", "adv8", "rand3"])
@pytest.mark.parametrize("packed", [False, True])
def test_match_loss_vs_oracle_and_golden(zsg, golden_meta, name, packed):
    ops, _ = zsg
    from oracle import synth, zsg_oracle as zo
    c = golden_meta["loss_cases"][name]
    z = load_npz("loss_" + name)
    B, seed = c["B"], c["seed"]
    g = torch.Generator().manual_seed(seed)
    batch = synth.make_batch(B, seed=seed, adversarial=c["adv"])
    att = (torch.randn(B, synth.NUM_ANCHORS, 1, generator=g) * 1.5 - 3.0)
    bbx = (torch.randn(B, synth.NUM_ANCHORS, 4, generator=g) * 0.7)
    anchs = zo.default_anchors()
    losses, datt, dreg, top1, pos = run_loss(ops, dev(att), dev(bbx), dev(batch["annot"]), dev(anchs), packed)
    # bit-exact index work, against the golden dump of the real reference
    assert np.array_equal(top1.numpy(), z["top1"])
    assert np.array_equal(pos.nonzero().numpy().astype(np.int32), z["pos_idx"])
    for i, k in enumerate(("loss", "cls_ls", "box_ls")):
        assert losses[i].item() == pytest.approx(c[k], rel=RTOL)
    np.testing.assert_allclose(datt[pos].numpy(), z["datt_pos"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(dreg[pos].numpy(), z["dbbx_pos"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(datt[:, ::97].numpy(), z["datt_stride"], rtol=RTOL, atol=1e-10)
    assert float(dreg[~pos].abs().sum()) == 0.0


def test_match_loss_large_batch_vs_oracle(zsg):
    ops, _ = zsg
    from oracle import synth, zsg_oracle as zo
    B = 64
    batch = synth.make_batch(B, seed=77, adversarial=True)
    g = torch.Generator().manual_seed(5)
    att = (torch.randn(B, synth.NUM_ANCHORS, 1, generator=g) * 2 - 3.0).requires_grad_(True)
    bbx = (torch.randn(B, synth.NUM_ANCHORS, 4, generator=g)).requires_grad_(True)
    anchs = zo.default_anchors()
    ref = zo.zsg_loss(att, bbx, batch["annot"], anchs)
    ref["loss"].backward()
    losses, datt, dreg, top1, pos = run_loss(ops, dev(att.detach()), dev(bbx.detach()), dev(batch["annot"]), dev(anchs))
    assert torch.equal(top1, ref["top1"]) and torch.equal(pos, ref["pos"])
    for i, k in enumerate(("loss", "cls_ls", "box_ls")):
        assert losses[i].item() == pytest.approx(ref[k].item(), rel=RTOL)
    np.testing.assert_allclose(datt.numpy(), att.grad.squeeze(-1).numpy(), rtol=RTOL, atol=1e-10)
    np.testing.assert_allclose(dreg.numpy(), bbx.grad.numpy(), rtol=RTOL, atol=1e-10)


def test_match_loss_rejects_bad_arguments(zsg):
    ops, _ = zsg
    from zsg_b200._lib import ZsgError
    t = torch.zeros(8, device="cuda")
    with pytest.raises(ZsgError):
        ops.match_loss(t, 1, t, 4, t, t.double(), 0, 0, 0.6, 0.25, 2.0, 1.0, True, t.double(), t, 1, t, 4,
                       t.long(), t.byte(), t.double())


@pytest.mark.parametrize("name", ["rand8", "adv8"])
def test_eval_vs_oracle_and_golden(zsg, golden_meta, name):
    ops, _ = zsg
    from oracle import synth, zsg_oracle as zo
    c = golden_meta["loss_cases"][name]
    z = load_npz("loss_" + name)
    B, seed = c["B"], c["seed"]
    g = torch.Generator().manual_seed(seed)
    batch = synth.make_batch(B, seed=seed, adversarial=c["adv"])
    att = (torch.randn(B, synth.NUM_ANCHORS, 1, generator=g) * 1.5 - 3.0)
    bbx = (torch.randn(B, synth.NUM_ANCHORS, 4, generator=g) * 0.7)
    anchs = zo.default_anchors()
    A = synth.NUM_ANCHORS
    best = torch.empty(B, dtype=torch.int64, device="cuda")
    scores = torch.empty(B, device="cuda")
    boxes = torch.empty(B, 4, dtype=torch.float64, device="cuda")
    metrics = torch.empty(2 + 2 * B, device="cuda")
    ops.evaluate(dev(att), 1, dev(bbx), 4, dev(batch["annot"]), dev(anchs), dev(batch["img_size"]), B, A, 0.5, best,
                 scores, boxes, metrics)
    torch.cuda.synchronize()
    assert np.array_equal(best.cpu().numpy(), z["best_ids"])
    assert metrics[0].item() == c["Acc"] and metrics[1].item() == c["MaxPos"]
    np.testing.assert_allclose(boxes.cpu().numpy(), z["pred_boxes"], rtol=1e-6)
    np.testing.assert_allclose(scores.cpu().numpy(), z["best"], rtol=1e-6)


# ------------------------------------------------------------------------------- convolution
def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def khwc(w):
    return w.permute(0, 2, 3, 1).contiguous()       # [cout][r][s][cin]


CONV_CASES = [
    # B, cin, H, W, cout, k, stride, pad
    (2, 64, 19, 19, 64, 1, 1, 0),
    (2, 64, 19, 19, 256, 3, 1, 1),
    (3, 128, 20, 18, 128, 3, 2, 1),
    (2, 256, 10, 10, 512, 1, 2, 0),
    (2, 4, 61, 61, 64, 7, 2, 3),
    (2, 256, 5, 5, 45, 3, 1, 1),
    (5, 300, 4, 1, 512, 1, 1, 0),                   # LSTM input projection shape (K tail: 300 = 9*32+12)
    (1, 520, 10, 10, 256, 3, 1, 1),                 # head conv0 with padded fusion channels
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("impl", [0, 1])
def test_conv_fwd_vs_torch(zsg, case, impl):
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    ref = F.conv2d(x, w, bias, stride=stride, padding=pad)
    Ho, Wo = ref.shape[2], ref.shape[3]
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    y = torch.full((B, Ho, Wo, cout), float("nan"), device="cuda")
    op = ops.ConvOp(nhwc(x), khwc(w), y, rows, B * Ho * Wo, cin, cout, k, k, bias=bias, impl=impl)
    op()
    torch.cuda.synchronize()
    assert rel_err(y, nhwc(ref)) < 2e-5


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_tma_weights(zsg, case):
    """Weights pre-split into TF32 hi/lo images and fetched by TMA (the path the engine uses)."""
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(hash(case) % 1000 + 1)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    ref = F.conv2d(x, w, bias, stride=stride, padding=pad)
    Ho, Wo = ref.shape[2], ref.shape[3]
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    wk = khwc(w)
    hi, lo = torch.empty_like(wk), torch.empty_like(wk)
    ops.split_tf32(wk, hi, lo, wk.numel())
    torch.cuda.synchronize()
    assert torch.equal(hi + lo, wk) and int((hi.view(torch.int32) & 0x1FFF).abs().sum()) == 0
    for impl in (0, 1):
        y = torch.full((B, Ho, Wo, cout), float("nan"), device="cuda")
        ops.ConvOp(nhwc(x), hi, y, rows, B * Ho * Wo, cin, cout, k, k, bias=bias, impl=impl, w_lo=lo)()
        torch.cuda.synchronize()
        assert rel_err(y, nhwc(ref)) < 2e-5, impl


@pytest.mark.parametrize("impl", [0, 1])
def test_conv_prologue_epilogue(zsg, impl):
    """BatchNorm affine + ReLU on load (zero padding stays zero), bias, residual, ReLU on store."""
    ops, geo = zsg
    B, cin, H, W, cout = 2, 64, 13, 11, 128
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / 24).cuda()
    sc, sh = (torch.rand(cin, generator=g) + 0.5).cuda(), torch.randn(cin, generator=g).cuda() * 0.3
    res = torch.randn(B, cout, H, W, generator=g).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    a = F.relu(x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    ref = F.relu(F.conv2d(a, w, bias, padding=1) + res)
    rows = geo.conv_rows(B, H, W, cin, H, W, cout, 1, 1).cuda()
    y = torch.empty(B, H, W, cout, device="cuda")
    ops.ConvOp(nhwc(x), khwc(w), y, rows, B * H * W, cin, cout, 3, 3, in_scale=sc, in_shift=sh, in_relu=True, bias=bias,
               out_relu=True, residual=nhwc(res), impl=impl)()
    torch.cuda.synchronize()
    assert rel_err(y, nhwc(ref)) < 2e-5


DGRAD_CASES = [(2, 64, 19, 19, 256, 3, 1, 1), (3, 128, 20, 18, 128, 3, 2, 1), (2, 256, 10, 10, 512, 1, 2, 0),
               (2, 2048, 10, 10, 256, 3, 2, 1), (2, 256, 5, 5, 45, 3, 1, 1), (2, 64, 8, 8, 64, 1, 1, 0)]


@pytest.mark.parametrize("case", DGRAD_CASES)
@pytest.mark.parametrize("impl", [0, 1])
def test_conv_dgrad_vs_torch(zsg, case, impl):
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, cin, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    y = F.conv2d(x, w, None, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    wk = khwc(w)
    wt = torch.empty(cin, k, k, cout, device="cuda")
    ops.weight_transpose_flip(wk, wt, cout, k, k, cin)
    dyn = nhwc(dy)
    cp = (cout + 3) // 4 * 4
    if cp != cout:                                   # the 45-channel head output: pad to 48 (engine does the same)
        wtp = torch.empty(cin, k, k, cp, device="cuda")
        ops.pad_channels(wt, wtp, cin * k * k, cout, cp)
        dyp = torch.empty(B, Ho, Wo, cp, device="cuda")
        ops.pad_channels(dyn, dyp, B * Ho * Wo, cout, cp)
        wt, dyn = wtp, dyp
    rows = geo.dgrad_rows(B, H, W, cin, Ho, Wo, cp, k, stride, pad).cuda()
    dx = torch.empty(B, H, W, cin, device="cuda")
    ops.ConvOp(dyn, wt, dx, rows, B * H * W, cp, cin, k, k, in_div=stride, impl=impl)()
    torch.cuda.synchronize()
    assert rel_err(dx, nhwc(x.grad)) < 2e-5


WGRAD_CASES = [(2, 64, 19, 19, 256, 3, 1, 1), (3, 128, 20, 18, 128, 3, 2, 1), (2, 4, 61, 61, 64, 7, 2, 3),
               (2, 256, 5, 5, 45, 3, 1, 1), (5, 300, 4, 1, 512, 1, 1, 0), (4, 1024, 19, 19, 256, 1, 1, 0)]


@pytest.mark.parametrize("case", WGRAD_CASES)
@pytest.mark.parametrize("impl", [0, 1])
def test_conv_wgrad_vs_torch(zsg, case, impl):
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(13)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda().requires_grad_(True)
    sc, sh = (torch.rand(cin, generator=g) + 0.5).cuda(), torch.randn(cin, generator=g).cuda() * 0.3
    a = F.relu(x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    y = F.conv2d(a, w, None, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    ops.WgradOp(nhwc(x), nhwc(dy), dw, rows, B * Ho * Wo, cin, cout, k, k, in_scale=sc, in_shift=sh, in_relu=True,
                impl=impl)()
    torch.cuda.synchronize()
    assert rel_err(dw, khwc(w.grad)) < 3e-5


def test_wgrad_mn_major_descriptor_probe(zsg):
    """Reports which LBO/SBO reading of the MN-major smem descriptor is right (impl 0 = product, 7 = swapped)."""
    ops, geo = zsg
    B, cin, H, W, cout, k = 2, 256, 12, 12, 256, 3
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / 48).cuda().requires_grad_(True)
    y = F.conv2d(x, w, None, padding=1)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(dy)
    rows = geo.conv_rows(B, H, W, cin, H, W, cout, 1, 1).cuda()
    errs = {}
    for impl in (0, 7):
        dw = torch.zeros(cout, k, k, cin, device="cuda")
        ops.WgradOp(nhwc(x), nhwc(dy), dw, rows, B * H * W, cin, cout, k, k, impl=impl)()
        torch.cuda.synchronize()
        errs[impl] = rel_err(dw, khwc(w.grad))
    print("MN-major descriptor probe: rel err product layout %.3e, swapped LBO/SBO %.3e" % (errs[0], errs[7]))
    assert errs[0] < 3e-5, errs


def test_conv_multilevel_shared_weights(zsg):
    """The head applies one weight set to six levels in a single launch (mdl.py:379-380)."""
    ops, geo = zsg
    B, c, cout = 2, 64, 45
    sizes = (7, 4, 2, 1)
    g = torch.Generator().manual_seed(17)
    w = (torch.randn(cout, c, 3, 3, generator=g) / 24).cuda()
    xs = [torch.randn(B, c, s, s, generator=g).cuda() for s in sizes]
    cells = sum(s * s for s in sizes)
    xin = torch.cat([nhwc(x).reshape(-1) for x in xs])
    tabs, in_off, a_off = [], 0, 0
    A = cells * 9
    for s in sizes:
        # output scattered like permute_correctly + cat: [B, A, 5] with anchor-major packing
        t = geo.conv_rows(B, s, s, c, s, s, cout, 1, 1, in_off=in_off)
        tabs.append((t, s, a_off))
        in_off += B * s * s * c
        a_off += s * s * 9
    import numpy as np
    fixed = []
    for t, s, ao in tabs:
        arr = t.numpy().view(np.dtype([("base", "<i4"), ("y0", "<i2"), ("x0", "<i2"), ("hin", "<i2"), ("win", "<i2"),
                                        ("out", "<i4")])).reshape(-1).copy()
        idx = np.arange(arr.shape[0])
        b, cell = idx // (s * s), idx % (s * s)
        arr["out"] = (b * A + ao + cell * 9) * 5
        fixed.append(torch.from_numpy(arr.view(np.uint8).reshape(-1, 16)))
    rows = torch.cat(fixed).cuda()
    out = torch.empty(B, A, 5, device="cuda")
    ops.ConvOp(xin, khwc(w), out, rows, B * cells, c, cout, 3, 3)()
    torch.cuda.synchronize()
    ref = torch.cat([F.conv2d(x, w, None, padding=1).permute(0, 2, 3, 1).reshape(B, -1, 5) for x in xs], dim=1)
    assert rel_err(out, ref) < 2e-5


# ------------------------------------------------------------------------------- BatchNorm etc.
def test_batchnorm_train_fwd_bwd(zsg):
    ops, _ = zsg
    B, C, H, W = 4, 256, 19, 17
    g = torch.Generator().manual_seed(19)
    x = (torch.randn(B, C, H, W, generator=g) * 2 + 0.7).cuda().requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).cuda().requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.2).cuda().requires_grad_(True)
    rm, rv = torch.zeros(C).cuda(), torch.ones(C).cuda()
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y_ref = F.relu(F.batch_norm(x, rm_ref, rv_ref, gamma, beta, training=True, momentum=0.1, eps=1e-5))
    dy = torch.randn(y_ref.shape, generator=g).cuda()
    y_ref.backward(dy)
    rows = B * H * W
    xn = nhwc(x.detach())
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    mean, invstd, scale, shift = (torch.empty(C, device="cuda") for _ in range(4))
    ops.bn_stats(xn, sums, rows, C)
    ops.bn_finalize(sums, rows, C, gamma.detach(), beta.detach(), 1e-5, 0.1, rm, rv, mean, invstd, scale, shift)
    y = torch.empty_like(xn)
    ops.bn_apply(xn, scale, shift, y, rows, C, relu=True)
    torch.cuda.synchronize()
    assert rel_err(y, nhwc(y_ref)) < 1e-5
    assert rel_err(rm, rm_ref) < 1e-5 and rel_err(rv, rv_ref) < 1e-5
    bsum = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    dyn = nhwc(dy)
    ops.bn_bwd_reduce(dyn, xn, mean, invstd, bsum, rows, C, mask_mode=1, scale=scale, shift=shift)
    dx, dgam, dbet = torch.empty_like(xn), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.bn_bwd_apply(dyn, xn, mean, invstd, gamma.detach(), bsum, dx, dgam, dbet, rows, C, mask_mode=1, scale=scale,
                     shift=shift)
    torch.cuda.synchronize()
    assert rel_err(dx, nhwc(x.grad)) < 1e-4
    assert rel_err(dgam, gamma.grad) < 1e-4 and rel_err(dbet, beta.grad) < 1e-4


def test_bottleneck_tail_and_mask_mode2(zsg):
    ops, _ = zsg
    rows, C = 1000, 512
    g = torch.Generator().manual_seed(23)
    x3, idt = torch.randn(rows, C, generator=g).cuda(), torch.randn(rows, C, generator=g).cuda()
    sc, sh = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    sc2, sh2 = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    y = torch.empty_like(x3)
    ops.bn_apply(x3, sc, sh, y, rows, C, relu=True, r=idt)
    torch.cuda.synchronize()
    assert rel_err(y, F.relu(x3 * sc + sh + idt)) < 1e-6
    ops.bn_apply(x3, sc, sh, y, rows, C, relu=True, r=idt, rscale=sc2, rshift=sh2)
    torch.cuda.synchronize()
    ref = F.relu(x3 * sc + sh + idt * sc2 + sh2)
    assert rel_err(y, ref) < 1e-6
    dy = torch.randn(rows, C, generator=g).cuda()
    mean, invstd = torch.randn(C, generator=g).cuda(), (torch.rand(C, generator=g) + 0.5).cuda()
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    dz = torch.empty_like(dy)
    ops.bn_bwd_reduce(dy, x3, mean, invstd, sums, rows, C, mask_mode=2, act_out=ref, dz_out=dz)
    torch.cuda.synchronize()
    dz_ref = dy * (ref > 0)
    assert torch.equal(dz, dz_ref)
    assert rel_err(sums[:C], dz_ref.double().sum(0)) < 1e-6
    assert rel_err(sums[C:], (dz_ref * (x3 - mean) * invstd).double().sum(0)) < 1e-5


def test_maxpool_upsample_avgpool(zsg, golden_meta):
    ops, _ = zsg
    g = torch.Generator().manual_seed(29)
    B, C, H, W = 2, 64, 30, 30
    x = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    sc, sh = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda() * 0.5
    a = F.relu(x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    a.retain_grad()
    y_ref = F.max_pool2d(a, 3, 2, 1)
    dy = torch.randn(y_ref.shape, generator=g).cuda()
    y_ref.backward(dy)
    Ho, Wo = y_ref.shape[2:]
    y = torch.empty(B, Ho, Wo, C, device="cuda")
    xn = nhwc(x.detach())
    arg = torch.empty(B, Ho, Wo, C, dtype=torch.uint8, device="cuda")
    ops.maxpool_bn_relu_fwd(xn, sc, sh, y, arg, B, H, W, C, Ho, Wo)
    da = torch.empty(B, H, W, C, device="cuda")
    ops.maxpool_bn_relu_bwd(arg, nhwc(dy), da, B, H, W, C, Ho, Wo)
    torch.cuda.synchronize()
    assert rel_err(y, nhwc(y_ref)) < 1e-6          # fmaf vs mul+add in the affine: not bit-equal
    assert rel_err(da, nhwc(a.grad)) < 1e-6
    # nearest upsample + add, with the reference's index tables
    for key, idx in golden_meta["upsample_idx"].items():
        hi, ho = (int(v) for v in key.split("->"))
        src = torch.randn(B, 256, hi, hi, generator=g).cuda().requires_grad_(True)
        dst = torch.randn(B, 256, ho, ho, generator=g).cuda()
        ref = dst + F.interpolate(src, size=(ho, ho))
        dref = torch.randn(ref.shape, generator=g).cuda()
        ref.backward(dref)
        it = torch.tensor(idx, dtype=torch.int32).cuda()
        d = nhwc(dst)
        ops.upsample_add(d, nhwc(src.detach()), it, it, B, ho, ho, hi, hi, 256)
        dsrc = torch.zeros(B, hi, hi, 256, device="cuda")
        ops.upsample_add_bwd(nhwc(dref), dsrc, it, it, B, ho, ho, hi, hi, 256)
        torch.cuda.synchronize()
        assert rel_err(d, nhwc(ref)) < 1e-6 and rel_err(dsrc, nhwc(src.grad)) < 1e-5
    p7 = torch.randn(B, 256, 3, 3, generator=g).cuda()
    out = torch.empty(B, 256, device="cuda")
    ops.avgpool_fwd(nhwc(p7), out, B, 9, 256)
    torch.cuda.synchronize()
    assert rel_err(out, F.adaptive_avg_pool2d(p7, 1).flatten(1)) < 1e-6


def test_fuse_unfuse_colsum_pad_adam(zsg):
    ops, _ = zsg
    from oracle import zsg_oracle as zo
    g = torch.Generator().manual_seed(31)
    B, sizes = 3, (5, 3, 1)
    cells = [s * s for s in sizes]
    tot = sum(cells)
    feats = [torch.randn(B, s * s, 256, generator=g).cuda() for s in sizes]
    lang = torch.randn(B, 256, generator=g).cuda()
    grid = torch.cat([zo.make_grid(s, s).view(-1, 2) for s in sizes]).cuda()
    fused = torch.empty(B * tot, 520, device="cuda")
    ops.fuse_lang_grid(torch.cat([f.reshape(-1) for f in feats]), lang, grid, fused, B, tot, cells, 256, 256, 520)
    torch.cuda.synchronize()
    off, coff = 0, 0
    for f, n in zip(feats, cells):
        blk = fused[off:off + B * n].view(B, n, 520)
        assert torch.equal(blk[..., :256], f)
        assert torch.equal(blk[..., 256:512], lang[:, None, :].expand(B, n, 256))
        assert torch.equal(blk[..., 512:514], grid[coff:coff + n][None].expand(B, n, 2))
        assert float(blk[..., 514:].abs().sum()) == 0
        off += B * n
        coff += n
    dfused = torch.randn(B * tot, 520, generator=g).cuda()
    dfeat, dlang = torch.empty(B * tot, 256, device="cuda"), torch.empty(B, 256, device="cuda")
    ops.unfuse_lang_grid(dfused, dfeat, dlang, B, tot, cells, 256, 256, 520)
    torch.cuda.synchronize()
    assert torch.equal(dfeat, dfused[:, :256])
    ref, off = torch.zeros(B, 256, device="cuda"), 0
    for n in cells:
        ref += dfused[off:off + B * n].view(B, n, 520)[..., 256:512].sum(1)
        off += B * n
    assert rel_err(dlang, ref) < 1e-5
    x = torch.randn(777, 45, generator=g).cuda()
    out = torch.empty(45, device="cuda")
    ops.colsum(x, out, 777, 45)
    rt = torch.from_numpy(np.zeros(777 * 4, dtype=np.int32)).view(777, 4)
    rt[:, 3] = torch.arange(777, dtype=torch.int32).flip(0) * 45
    gat = torch.empty(777, 48, device="cuda")
    ops.gather_rows(x, rt.cuda(), gat, 777, 45, 48)
    torch.cuda.synchronize()
    assert rel_err(out, x.sum(0)) < 1e-5
    assert torch.equal(gat[:, :45], x.flip(0)) and float(gat[:, 45:].abs().sum()) == 0
    src = torch.randn(64 * 49, 3, generator=g).cuda()
    dst = torch.empty(64 * 49, 4, device="cuda")
    ops.pad_channels(src, dst, 64 * 49, 3, 4)
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :3], src) and float(dst[:, 3].abs().sum()) == 0
    img = torch.rand(2, 3, 9, 7, generator=g).cuda()
    o4 = torch.empty(2, 9, 7, 4, device="cuda")
    ops.nchw_to_nhwc4(img, o4)
    torch.cuda.synchronize()
    assert torch.equal(o4[..., :3], img.permute(0, 2, 3, 1))
    # Adam, two steps, against torch.optim.Adam(betas=(0.9, 0.99)) (main_dist.py:50)
    p = torch.randn(10001, generator=g).cuda()
    p_ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3, betas=(0.9, 0.99))
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in (1, 2):
        gr = torch.randn(10001, generator=g).cuda()
        p_ref.grad = gr.clone()
        opt.step()
        ops.adam(p, gr, m, v, p.numel(), 1e-3, 0.9, 0.99, 1e-8, step)
    torch.cuda.synchronize()
    assert rel_err(p, p_ref.detach()) < 1e-5


# ------------------------------------------------------------------------------- LSTM
def test_lstm_recurrences_vs_oracle(zsg):
    ops, _ = zsg
    from oracle import synth, zsg_oracle as zo
    sd = synth.make_state_dict(0)
    B, T, E, H = 5, 20, 300, 128
    batch = synth.make_batch(B, seed=41, var_len=True)
    torch.manual_seed(41)
    h0, c0 = zo.draw_h0c0(B)
    keys = [k for k in sd if k.startswith("lstm.")]
    for k in keys:
        sd[k] = sd[k].clone().requires_grad_(True)
    qv = batch["qvec"].clone().requires_grad_(True)
    ref = zo.lstm_query(sd, qv, batch["qlens"], h0, c0)
    dl = torch.randn(B, 2 * H, generator=torch.Generator().manual_seed(1))
    ref.backward(dl)
    lens = batch["qlens"].int()
    _, perm = batch["qlens"].sort(0, descending=True)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(B)
    d = "cuda"
    gx = (batch["qvec"] @ sd["lstm.weight_ih_l0"].detach().t()).to(d).contiguous()     # the tensor-core part, tested elsewhere
    gates, cs, hp = torch.empty(B, T, 4 * H, device=d), torch.empty(B, T, H, device=d), torch.empty(B, T, H, device=d)
    lang = torch.empty(B, 2 * H, device=d)
    W = {k: sd[k].detach().to(d).contiguous() for k in keys}
    h0f, c0f, h0r, c0r = (dev(t) for t in (h0[0][inv], c0[0][inv], h0[1][inv], c0[1][inv]))
    ops.lstm_fwd_dir(gx, W["lstm.weight_hh_l0"].t().contiguous(), W["lstm.bias_ih_l0"], W["lstm.bias_hh_l0"], h0f, c0f,
                     lens.to(d), B, T, gates, cs, hp, lang)
    xlast, rg = torch.empty(B, E, device=d), torch.empty(B, 4 * H, device=d)
    ops.lstm_rev_step(dev(batch["qvec"]), W["lstm.weight_ih_l0_reverse"], W["lstm.weight_hh_l0_reverse"],
                      W["lstm.bias_ih_l0_reverse"], W["lstm.bias_hh_l0_reverse"], h0r, c0r, lens.to(d), B, T, E, xlast,
                      rg, lang)
    torch.cuda.synchronize()
    assert rel_err(lang, ref.detach()) < 1e-5
    dg = torch.empty(B, T, 4 * H, device=d)
    ops.lstm_bwd_dir(dev(dl), W["lstm.weight_hh_l0"], gates, cs, c0f, lens.to(d), B, T, dg)
    dgr = torch.empty(B, 4 * H, device=d)
    ops.lstm_rev_step_bwd(dev(dl), rg, c0r, B, dgr)
    torch.cuda.synchronize()
    dg2 = dg.view(B * T, 4 * H)
    assert rel_err(dg2.t() @ dev(batch["qvec"]).view(B * T, E), sd["lstm.weight_ih_l0"].grad) < 1e-4
    assert rel_err(dg2.t() @ hp.view(B * T, H), sd["lstm.weight_hh_l0"].grad) < 1e-4
    assert rel_err(dg2.sum(0), sd["lstm.bias_ih_l0"].grad) < 1e-4
    assert rel_err(dgr.t() @ xlast, sd["lstm.weight_ih_l0_reverse"].grad) < 1e-4
    assert rel_err(dgr.t() @ h0r, sd["lstm.weight_hh_l0_reverse"].grad) < 1e-4
    assert rel_err(dgr.sum(0), sd["lstm.bias_hh_l0_reverse"].grad) < 1e-4


# ------------------------------------------------------------------ operand images (cp.async / TMA GEMM paths)
def test_split_act_images(zsg):
    """z = relu(x*scale+shift) with the same fmaf as the in-kernel prologue; lo = z - trunc_tf32(z), exactly."""
    ops, _ = zsg
    g = torch.Generator().manual_seed(5)
    rows, c = 1237, 72
    x = torch.randn(rows, c, generator=g).cuda()
    sc, sh = (torch.rand(c, generator=g) + 0.5).cuda(), torch.randn(c, generator=g).cuda()
    z, lo = torch.empty_like(x), torch.empty_like(x)
    ops.split_act(x, lo, rows, c, scale=sc, shift=sh, relu=True, z=z)
    lo0 = torch.empty_like(x)
    ops.split_act(x, lo0, rows, c)
    torch.cuda.synchronize()
    zr = torch.relu(torch.addcmul(sh, x, sc))                       # single-rounding fma on the GPU as well
    assert float((z - zr).abs().max()) <= 1.2e-7 * float(zr.abs().max())
    hi = (z.view(torch.int32) & ~0x1FFF).view(torch.float32)
    assert torch.equal(hi + lo, z) and torch.equal(lo, z - hi)
    hi0 = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
    assert torch.equal(lo0, x - hi0)


@pytest.mark.parametrize("case", CONV_CASES + [(2, 128, 9, 9, 128, 3, 2, 1)])
@pytest.mark.parametrize("pro", [False, True])
def test_conv_async_equals_register_path(zsg, case, pro):
    """The cp.async kernel (tensor + remainder image) must reproduce the register-path kernel bit for bit and meet
    the same tolerance against torch."""
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(hash(case) % 1000 + 7)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    sc, sh = (torch.rand(cin, generator=g) + 0.5).cuda(), torch.randn(cin, generator=g).cuda() * 0.3
    a = F.relu(x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)) if pro else x
    ref = F.conv2d(a, w, bias, stride=stride, padding=pad)
    Ho, Wo = ref.shape[2], ref.shape[3]
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    wk = khwc(w)
    hi, lo = torch.empty_like(wk), torch.empty_like(wk)
    ops.split_tf32(wk, hi, lo, wk.numel())
    xn = nhwc(x)
    y1 = torch.full((B, Ho, Wo, cout), float("nan"), device="cuda")
    y2 = torch.full_like(y1, float("nan"))
    ops.ConvOp(xn, hi, y1, rows, B * Ho * Wo, cin, cout, k, k, bias=bias, w_lo=lo, in_scale=sc if pro else None,
               in_shift=sh if pro else None, in_relu=pro)()
    z, x_lo = (torch.empty_like(xn) if pro else None), torch.empty_like(xn)
    ops.split_act(xn, x_lo, B * H * W, cin, scale=sc if pro else None, shift=sh if pro else None, relu=pro, z=z)
    ops.ConvOp(z if pro else xn, hi, y2, rows, B * Ho * Wo, cin, cout, k, k, bias=bias, w_lo=lo, x_lo=x_lo)()
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    assert rel_err(y2, nhwc(ref)) < 2e-5


@pytest.mark.parametrize("case", DGRAD_CASES)
def test_conv_dgrad_async_vs_torch(zsg, case):
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(17)
    x = torch.randn(B, cin, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    y = F.conv2d(x, w, None, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    wt = torch.empty(cin, k, k, cout, device="cuda")
    ops.weight_transpose_flip(khwc(w), wt, cout, k, k, cin)
    dyn = nhwc(dy)
    cp = (cout + 3) // 4 * 4
    if cp != cout:
        wtp = torch.empty(cin, k, k, cp, device="cuda")
        ops.pad_channels(wt, wtp, cin * k * k, cout, cp)
        dyp = torch.empty(B, Ho, Wo, cp, device="cuda")
        ops.pad_channels(dyn, dyp, B * Ho * Wo, cout, cp)
        wt, dyn = wtp, dyp
    hi, lo = torch.empty_like(wt), torch.empty_like(wt)
    ops.split_tf32(wt, hi, lo, wt.numel())
    dy_lo = torch.empty_like(dyn)
    ops.split_act(dyn, dy_lo, B * Ho * Wo, cp)
    rows = geo.dgrad_rows(B, H, W, cin, Ho, Wo, cp, k, stride, pad).cuda()
    res = torch.randn(B, H, W, cin, generator=g).cuda()
    mask = torch.randn(B, H, W, cin, generator=g).cuda()
    dx = torch.empty(B, H, W, cin, device="cuda")
    ops.ConvOp(dyn, hi, dx, rows, B * H * W, cp, cin, k, k, in_div=stride, w_lo=lo, x_lo=dy_lo, out_mask=mask, residual=res)()
    torch.cuda.synchronize()
    want = nhwc(x.grad) * (mask > 0) + res
    assert rel_err(dx, want) < 2e-5


@pytest.mark.parametrize("case", WGRAD_CASES + [(2, 520, 10, 10, 256, 3, 1, 1), (2, 264, 7, 9, 128, 1, 1, 0),
                                               (2, 512, 19, 19, 512, 3, 1, 1), (2, 1024, 19, 19, 256, 1, 1, 0)])
@pytest.mark.parametrize("tma_dy", [False, True])
def test_conv_wgrad_async_vs_torch(zsg, case, tma_dy):
    """cp.async weight-gradient kernel (x, dy with remainder images; dy optionally by TMA) against torch, including
    channel counts whose 128-wide tiles straddle filter taps and the ragged 45-channel head output (pitch 48)."""
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(19)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda().requires_grad_(True)
    y = F.conv2d(x, w, None, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    pitch = (cout + 3) // 4 * 4
    dyn = nhwc(dy)
    if pitch != cout:
        dyp = torch.empty(B, Ho, Wo, pitch, device="cuda")
        ops.pad_channels(dyn, dyp, B * Ho * Wo, cout, pitch)
        dyn = dyp
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, pitch, stride, pad).cuda()
    xn = nhwc(x)
    x_lo, dy_lo = torch.empty_like(xn), torch.empty_like(dyn)
    ops.split_act(xn, x_lo, B * H * W, cin)
    ops.split_act(dyn, dy_lo, B * Ho * Wo, pitch)
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    ops.WgradOp(xn, dyn, dw, rows, B * Ho * Wo, cin, cout, k, k, x_lo=x_lo, dy_lo=dy_lo, dy_pitch=pitch if tma_dy else 0)()
    torch.cuda.synchronize()
    assert rel_err(dw, khwc(w.grad)) < 3e-5


def test_bn_kernels_write_operand_images(zsg):
    """bn_apply / bn_bwd_apply optionally emit the TF32 remainder image of their output (same values as zsg_split_act)."""
    ops, _ = zsg
    g = torch.Generator().manual_seed(23)
    rows, C = 777, 64
    x, r, dy = (torch.randn(rows, C, generator=g).cuda() for _ in range(3))
    sc, sh = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    y, y_lo, ref_lo = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    ops.bn_apply(x, sc, sh, y, rows, C, relu=True, r=r, y_lo=y_lo)
    ops.split_act(y, ref_lo, rows, C)
    torch.cuda.synchronize()
    assert torch.equal(y_lo, ref_lo)
    mean, invstd, gamma = torch.randn(C, generator=g).cuda(), (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    sums = torch.randn(2 * C, generator=g).double().cuda()
    dx, dx_lo, dg, db = torch.empty_like(x), torch.empty_like(x), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.bn_bwd_apply(dy, x, mean, invstd, gamma, sums, dx, dg, db, rows, C, dx_lo=dx_lo)
    ops.split_act(dx, ref_lo, rows, C)
    torch.cuda.synchronize()
    assert torch.equal(dx_lo, ref_lo)


# ------------------------------------------------------------------------------- SSD-VGG trunk glue (a-8)
@pytest.mark.parametrize("impl", [0, 1])
def test_dilated_conv_fwd_dgrad_wgrad(zsg, impl):
    """ssd_vgg.py:129: conv6 = 3x3, dilation 6, padding 6 (19x19 stays 19x19); forward, data and weight gradient
    against torch; impl 0 = the cp.async tcgen05 path, 1 = SIMT check kernels."""
    ops, geo = zsg
    B, cin, H, cout, k, pad, dil = 2, 64, 19, 128, 3, 6, 6
    g = torch.Generator().manual_seed(41)
    x = torch.randn(B, cin, H, H, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda().requires_grad_(True)
    bias = torch.randn(cout, generator=g).cuda()
    y = F.conv2d(x, w, bias, padding=pad, dilation=dil)
    assert y.shape[2] == H
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(dy)
    M = B * H * H
    xn, dyn, wk = nhwc(x.detach()), nhwc(dy), khwc(w.detach())
    rows = geo.conv_rows(B, H, H, cin, H, H, cout, 1, pad).cuda()
    hi, lo = torch.empty_like(wk), torch.empty_like(wk)
    ops.split_tf32(wk, hi, lo, wk.numel())
    x_lo, dy_lo = torch.empty_like(xn), torch.empty_like(dyn)
    ops.split_act(xn, x_lo, M, cin)
    ops.split_act(dyn, dy_lo, M, cout)
    yk = torch.empty(B, H, H, cout, device="cuda")
    if impl == 0:
        ops.ConvOp(xn, hi, yk, rows, M, cin, cout, k, k, bias=bias, w_lo=lo, x_lo=x_lo, dil=dil)()
    else:
        ops.ConvOp(xn, wk, yk, rows, M, cin, cout, k, k, bias=bias, impl=1, dil=dil)()
    wt = torch.empty(cin, k, k, cout, device="cuda")
    ops.weight_transpose_flip(wk, wt, cout, k, k, cin)
    thi, tlo = torch.empty_like(wt), torch.empty_like(wt)
    ops.split_tf32(wt, thi, tlo, wt.numel())
    drows = geo.dgrad_rows(B, H, H, cin, H, H, cout, k, 1, pad, dil=dil).cuda()
    dx = torch.empty(B, H, H, cin, device="cuda")
    if impl == 0:
        ops.ConvOp(dyn, thi, dx, drows, M, cout, cin, k, k, w_lo=tlo, x_lo=dy_lo, dil=dil)()
    else:
        ops.ConvOp(dyn, wt, dx, drows, M, cout, cin, k, k, impl=1, dil=dil)()
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    if impl == 0:
        ops.WgradOp(xn, dyn, dw, rows, M, cin, cout, k, k, x_lo=x_lo, dy_lo=dy_lo, dy_pitch=cout, dil=dil)()
    else:
        ops.WgradOp(xn, dyn, dw, rows, M, cin, cout, k, k, impl=1, dil=dil)()
    torch.cuda.synchronize()
    assert rel_err(yk, nhwc(y)) < 2e-5
    assert rel_err(dx, nhwc(x.grad)) < 2e-5
    assert rel_err(dw, khwc(w.grad)) < 3e-5


@pytest.mark.parametrize("case", [(2, 8, 10, 10, 2, 2, 0, False), (2, 12, 7, 9, 2, 2, 0, True), (2, 8, 75, 75, 2, 2, 0, True),
                                  (3, 16, 19, 19, 3, 1, 1, False), (1, 4, 5, 6, 3, 2, 1, False)])
def test_generic_maxpool_fwd_bwd(zsg, case):
    """ssd_vgg.py:115-118,127: MaxPool2d(2,2), MaxPool2d(2,2,ceil_mode=True), MaxPool2d(3,1,1) on ReLU outputs (exact
    zeros make ties: the first maximum in scan order must win, like ATen), backward with the ReLU mask folded in."""
    ops, _ = zsg
    B, C, H, W, k, stride, pad, ceil = case
    g = torch.Generator().manual_seed(43)
    pre = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    a = F.relu(pre)
    y = F.max_pool2d(a, k, stride, pad, ceil_mode=ceil)
    Ho, Wo = y.shape[2], y.shape[3]
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(dy)
    an = nhwc(a.detach())
    yk = torch.empty(B, Ho, Wo, C, device="cuda")
    arg = torch.empty(B * Ho * Wo * C, dtype=torch.uint8, device="cuda")
    ops.maxpool_fwd(an, yk, arg, B, H, W, C, k, stride, pad, Ho, Wo)
    dx = torch.full((B, H, W, C), float("nan"), device="cuda")
    ops.maxpool_bwd(arg, nhwc(dy), dx, B, H, W, C, k, stride, pad, Ho, Wo, mask=an)
    torch.cuda.synchronize()
    assert torch.equal(yk, nhwc(y))
    assert torch.equal(dx, nhwc(pre.grad))


def test_l2norm_fwd_bwd(zsg):
    """ssd_vgg.py:80: s = x / x.norm(dim=1, keepdim=True) on relu(conv4_3) and its autograd (ReLU mask folded in,
    accumulated onto the gradient that arrives through the other consumer of x)."""
    ops, _ = zsg
    B, C, H = 2, 512, 9
    g = torch.Generator().manual_seed(45)
    pre = torch.randn(B, C, H, H, generator=g).cuda().requires_grad_(True)
    x = F.relu(pre)
    s = x / x.norm(dim=1, keepdim=True)
    ds = torch.randn(s.shape, generator=g).cuda()
    s.backward(ds)
    xn = nhwc(x.detach())
    rows = B * H * H
    y, norm = torch.empty_like(xn), torch.empty(rows, device="cuda")
    ops.l2norm_fwd(xn, y, norm, rows, C)
    other = torch.randn(B, H, H, C, generator=g).cuda()
    dx = other.clone()
    ops.l2norm_bwd(nhwc(ds), xn, norm, dx, rows, C, accumulate=True, mask_relu=True)
    dx2 = torch.empty_like(dx)
    ops.l2norm_bwd(nhwc(ds), xn, norm, dx2, rows, C)
    torch.cuda.synchronize()
    assert rel_err(y, nhwc(s)) < 1e-6
    assert rel_err(dx - other, nhwc(pre.grad)) < 2e-5
    assert rel_err(dx2 * (xn > 0), nhwc(pre.grad)) < 2e-5


@pytest.mark.parametrize("case", [(2, 64, 19, 19, 256, 3, 1, 1), (3, 128, 20, 18, 64, 1, 1, 0), (2, 4, 61, 61, 64, 7, 2, 3),
                                  (5, 256, 7, 9, 520, 1, 1, 0)])
def test_conv_epilogue_batchnorm_statistics(zsg, case):
    """zsg_conv_params.stats: per-32-row-group column sums / sums of squares written by the conv epilogue, reduced by
    zsg_bn_stats_partials, must equal zsg_bn_stats over the stored output (and torch's mean / biased variance)."""
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(51)
    x = torch.randn(B, cin, H, W, generator=g).cuda() + 0.3
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    ref = F.conv2d(x, w, None, stride=stride, padding=pad)
    Ho, Wo = ref.shape[2], ref.shape[3]
    M = B * Ho * Wo
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    wk, xn = khwc(w), nhwc(x)
    hi, lo, x_lo = torch.empty_like(wk), torch.empty_like(wk), torch.empty_like(xn)
    ops.split_tf32(wk, hi, lo, wk.numel())
    ops.split_act(xn, x_lo, B * H * W, cin)
    parts = (M + 127) // 128 * 4
    part = torch.full((parts, 2, cout), float("nan"), device="cuda")
    y = torch.empty(B, Ho, Wo, cout, device="cuda")
    ops.ConvOp(xn, hi, y, rows, M, cin, cout, k, k, w_lo=lo, x_lo=x_lo, stats=part)()
    sums = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    ops.bn_stats_partials(part, parts, cout, sums)
    direct = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    pow2 = (cout // 4) & (cout // 4 - 1) == 0                  # zsg_bn_stats wants C/4 to be a power of two
    if pow2:
        ops.bn_stats(y, direct, M, cout)
    # fused reduce + finalize (ticket counter) against the two-launch path, twice to check the counter resets
    gamma, beta = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
    tickets = torch.zeros(64, dtype=torch.int32, device="cuda")
    outs = []
    for _ in range(2):
        s2_ = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
        rm, rv = torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda")
        o = [torch.empty(cout, device="cuda") for _ in range(4)]
        ops.bn_finalize_partials(part, parts, M, cout, gamma, beta, 1e-5, 0.1, rm, rv, *o, s2_, tickets)
        outs.append(o + [rm, rv])
    s3_ = sums.clone()
    rm0, rv0 = torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda")
    o0 = [torch.empty(cout, device="cuda") for _ in range(4)]
    ops.bn_finalize(s3_, M, cout, gamma, beta, 1e-5, 0.1, rm0, rv0, *o0)
    torch.cuda.synchronize()
    assert int(tickets.abs().sum()) == 0
    for o in outs:
        for a, b in zip(o, o0 + [rm0, rv0]):
            np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert not torch.isnan(part).any()
    assert rel_err(y, nhwc(ref)) < 2e-5
    # sums of ~M terms of either sign: absolute tolerance = fp32 rounding of the 32-row partials (values are O(1))
    if pow2:
        np.testing.assert_allclose(sums.cpu().numpy(), direct.cpu().numpy(), rtol=2e-6, atol=1e-7 * M)
    y2 = y.view(M, cout).double()
    np.testing.assert_allclose((sums[:cout] / M).cpu().numpy(), y2.mean(0).cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose((sums[cout:] / M).cpu().numpy(), (y2 * y2).mean(0).cpu().numpy(), rtol=1e-5, atol=1e-7)


def test_batched_weight_transpose_flip(zsg):
    """One launch for many conv weights must equal the per-tensor kernel (which is checked against torch above)."""
    ops, _ = zsg
    g = torch.Generator().manual_seed(61)
    shapes = [(64, 1, 64), (256, 3, 128), (45 + 3, 3, 256), (8, 7, 4), (512, 1, 2048)]     # cout, k, cin
    src = torch.randn(sum(co * k * k * ci for co, k, ci in shapes) + 64, generator=g).cuda()
    dst = torch.full_like(src, float("nan"))
    want = torch.full_like(src, float("nan"))
    entries, off = [], 0
    for co, k, ci in shapes:
        n = co * k * k * ci
        entries.append((off, off, co, k, k, ci))
        ops.weight_transpose_flip(src[off:off + n], want[off:off + n], co, k, k, ci)
        off += n
    tab, total = ops.wtf_table(entries, "cuda")
    assert total == off
    ops.weight_transpose_flip_batched(src, dst, tab, len(entries), total)
    torch.cuda.synchronize()
    assert torch.equal(dst[:off], want[:off]) and torch.isnan(dst[off:]).all()
    # the tiled variant (every cin / cout a multiple of 32, as in the engine's table)
    shapes = [(64, 1, 64), (256, 3, 128), (64, 3, 64), (2048, 1, 512), (256, 3, 256)]
    src = torch.randn(sum(co * k * k * ci for co, k, ci in shapes), generator=g).cuda()
    dst, want = torch.full_like(src, float("nan")), torch.full_like(src, float("nan"))
    entries, off = [], 0
    for co, k, ci in shapes:
        n = co * k * k * ci
        entries.append((off, off, co, k, k, ci))
        ops.weight_transpose_flip(src[off:off + n], want[off:off + n], co, k, k, ci)
        off += n
    assert ops.wtf_table_is_tiled(entries)
    tab, total = ops.wtf_table(entries, "cuda")
    ops.weight_transpose_flip_batched(src, dst, tab, len(entries), total, tiled=True)
    torch.cuda.synchronize()
    assert torch.equal(dst, want)


# ------------------------------------------------------------------------------- NaN handling, anchors (ADVICE)
def test_match_loss_nan_guard_and_stateless_workspace(zsg):
    """loss.py:128-133: a NaN box or class loss is replaced by the constants 0.01 / 1.0, which carry no gradient.  A diverged
    network (NaN score) must not poison the step's gradients, and no call may depend on what an earlier one (or nobody) left
    in the workspace."""
    ops, _ = zsg
    from oracle import synth, zsg_oracle as zo
    B, A = 4, synth.NUM_ANCHORS
    g = torch.Generator().manual_seed(3)
    batch = synth.make_batch(B, seed=3)
    att = (torch.randn(B, A, 1, generator=g) * 1.5 - 3.0)
    bbx = torch.randn(B, A, 4, generator=g) * 0.7
    anchs = zo.default_anchors()
    for packed in (False, True):
        bad = att.clone()
        bad[1, 777, 0] = float("nan")
        d = att.device
        ws = ops.match_loss_workspace(B, "cuda")
        ws.view(torch.int64).fill_(-1)                                     # garbage: the workspace carries no state into a call
        losses = torch.empty(3, dtype=torch.float64, device="cuda")
        top1, pos = torch.empty(B, dtype=torch.int64, device="cuda"), torch.empty(B, A, dtype=torch.uint8, device="cuda")

        def call(a):
            if packed:
                buf = torch.cat([bbx, a], dim=2).cuda().contiguous()
                dbuf = torch.full_like(buf, 7.0)
                ops.match_loss(buf.view(-1)[4:], 5, buf, 5, dev(batch["annot"]), dev(anchs), B, A, 0.6, 0.25, 2.0, 1.0, True,
                               losses, dbuf.view(-1)[4:], 5, dbuf, 5, top1, pos, ws)
                torch.cuda.synchronize()
                return dbuf[..., 4].cpu(), dbuf[..., :4].cpu()
            datt, dreg = torch.full((B, A), 7.0, device="cuda"), torch.full((B, A, 4), 7.0, device="cuda")
            ops.match_loss(dev(a), 1, dev(bbx), 4, dev(batch["annot"]), dev(anchs), B, A, 0.6, 0.25, 2.0, 1.0, True,
                           losses, datt, 1, dreg, 4, top1, pos, ws)
            torch.cuda.synchronize()
            return datt.cpu(), dreg.cpu()
        datt, dreg = call(bad)
        assert losses.cpu().tolist() == [1.0 * 0.01 + 1.0, 1.0, 0.01]
        assert float(datt.abs().sum()) == 0.0 and float(dreg.abs().sum()) == 0.0
        # the same buffers right afterwards with finite scores: the normal result (nothing stuck from the NaN call)
        ref = zo.zsg_loss(att.clone().requires_grad_(True), bbx.clone().requires_grad_(True), batch["annot"], anchs)
        datt, dreg = call(att)
        assert losses[0].item() == pytest.approx(ref["loss"].item(), rel=RTOL)
        assert torch.equal(top1.cpu(), ref["top1"]) and float(datt.abs().sum()) > 0


def test_nan_rows_select_a_valid_anchor_like_torch_max(zsg):
    """torch.max treats NaN as the maximum and returns the first such index (loss.py:77, evaluator.py:74).  A NaN annotation
    (whole IoU row NaN) or all-NaN scores must give index 0, never an out-of-range index (the round-1 kernels
    kept INT_MAX and read the anchor table out of bounds)."""
    ops, _ = zsg
    from oracle import synth, zsg_oracle as zo
    B, A = 3, synth.NUM_ANCHORS
    batch = synth.make_batch(B, seed=9)
    annot = batch["annot"].clone()
    annot[1, 2] = float("nan")
    g = torch.Generator().manual_seed(9)
    att = torch.randn(B, A, 1, generator=g)
    att[2] = float("nan")
    bbx = torch.randn(B, A, 4, generator=g) * 0.3
    anchs = zo.default_anchors()
    losses, datt, dreg, top1, pos = run_loss(ops, dev(att), dev(bbx), dev(annot), dev(anchs))
    iou = zo.iou_gt_vs_anchors(annot, anchs)
    assert torch.isnan(iou[1]).all()
    want_top1 = iou.max(1)[1]                                            # CPU torch.max: first NaN = index 0
    assert want_top1[1].item() == 0 and torch.equal(top1, want_top1)
    assert pos[1].sum().item() == 1 and pos[1, 0]                        # nothing exceeds the threshold, only the top-1
    assert losses.tolist() == [1.01, 1.0, 0.01]                          # NaN loss -> guard constants
    best = torch.empty(B, dtype=torch.int64, device="cuda")
    scores, boxes = torch.empty(B, device="cuda"), torch.empty(B, 4, dtype=torch.float64, device="cuda")
    metrics = torch.empty(2 + 2 * B, device="cuda")
    ops.evaluate(dev(att), 1, dev(bbx), 4, dev(annot), dev(anchs), dev(batch["img_size"]), B, A, 0.5, best, scores, boxes,
                 metrics)
    torch.cuda.synchronize()
    want_best = torch.sigmoid(att).squeeze(-1).max(1)[1]
    assert want_best[2].item() == 0 and torch.equal(best.cpu(), want_best)
    assert torch.isnan(scores[2]).item() and 0.0 <= metrics[0].item() <= 1.0


def test_product_anchor_table_equals_reference_golden(zsg):
    """anchors.create_anchors of the PRODUCT package (what the loss / evaluator kernels read) bit-equal to the table the real
    reference produced (tests/golden/anchors.npz): fp64, with the fp32-rounded 2/h factor of anchors.py:66-87."""
    from zsg_b200.anchors import create_anchors
    from oracle import synth
    ratios, scales = synth.ratios_scales()
    a = create_anchors([(s, s) for s in synth.LEVEL_SIZES], ratios, scales, flatten=True, device="cuda")
    z = load_npz("anchors")
    assert a.dtype == torch.float64 and tuple(a.shape) == (synth.NUM_ANCHORS, 4)
    assert np.array_equal(a.cpu().numpy(), z["anchs"])


def test_zero_area_box_takes_the_nan_guard_like_reference(zsg):
    """a-13: a zero-area ground-truth box has log(0) regression targets.  The reference multiplies the box loss of every anchor
    by the positive mask (loss.py:92), so inf * 0 = NaN reaches the guard of loss.py:128-133: constants, no gradient.  (The
    golden dump tests/golden/loss_zero_area.npz holds what the real reference returned.)"""
    ops, _ = zsg
    from oracle import synth, zsg_oracle as zo
    B, A = 2, synth.NUM_ANCHORS
    batch = synth.make_batch(B, seed=4)
    annot = batch["annot"].clone()
    annot[0] = torch.tensor([0.1, 0.1, 0.1, 0.4])                         # zero height
    g = torch.Generator().manual_seed(4)
    att, bbx = torch.randn(B, A, 1, generator=g) - 3.0, torch.randn(B, A, 4, generator=g) * 0.3
    anchs = zo.default_anchors()
    losses, datt, dreg, top1, pos = run_loss(ops, dev(att), dev(bbx), dev(annot), dev(anchs))
    ref = zo.zsg_loss(att, bbx, annot, anchs)
    z = load_npz("loss_zero_area")
    assert [ref["loss"].item(), ref["cls_ls"].item(), ref["box_ls"].item()] == [1.01, 1.0, 0.01] == z["losses"].tolist()
    assert losses.tolist() == [1.01, 1.0, 0.01]
    assert torch.equal(top1, ref["top1"]) and torch.equal(pos, ref["pos"]) and np.array_equal(top1.numpy(), z["top1"])
    assert float(datt.abs().sum()) == 0.0 and float(dreg.abs().sum()) == 0.0


def test_loss_scalars_are_bit_reproducible_and_split_api_equals_composite(zsg):
    """Partial sums are added in a fixed order (no floating-point atomics): the three loss scalars and both gradients are
    identical bit for bit from run to run; zsg_match followed by zsg_loss_grad is zsg_match_loss."""
    ops, _ = zsg
    from oracle import synth, zsg_oracle as zo
    B, A = 16, synth.NUM_ANCHORS
    g = torch.Generator().manual_seed(8)
    batch = synth.make_batch(B, seed=8, adversarial=True)
    att = dev(torch.randn(B, A, 1, generator=g) * 1.5 - 3.0)
    bbx = dev(torch.randn(B, A, 4, generator=g) * 0.7)
    annot, anchs = dev(batch["annot"]), dev(zo.default_anchors())
    runs = [run_loss(ops, att, bbx, annot, anchs, packed) for packed in (False, False, True)]
    for r in runs[1:]:
        assert torch.equal(runs[0][0], r[0]) and torch.equal(runs[0][1], r[1]) and torch.equal(runs[0][2], r[2])
        assert torch.equal(runs[0][3], r[3]) and torch.equal(runs[0][4], r[4])
    ws = ops.match_loss_workspace(B, "cuda")
    losses = torch.empty(3, dtype=torch.float64, device="cuda")
    top1, pos = torch.empty(B, dtype=torch.int64, device="cuda"), torch.empty(B, A, dtype=torch.uint8, device="cuda")
    datt, dreg = torch.empty(B, A, device="cuda"), torch.empty(B, A, 4, device="cuda")
    ops.match(annot, anchs, B, A, 0.6, True, top1, pos, ws)
    ops.loss_grad(att, 1, bbx, 4, annot, anchs, pos, B, A, 0.25, 2.0, 1.0, losses, datt, 1, dreg, 4, ws)
    torch.cuda.synchronize()
    assert torch.equal(losses.cpu(), runs[0][0]) and torch.equal(datt.cpu(), runs[0][1]) and torch.equal(top1.cpu(), runs[0][3])


@pytest.mark.parametrize("case", [(3, 64, 13, 11, 256), (2, 256, 20, 19, 64), (5, 300, 1, 4, 512), (2, 2048, 10, 10, 256),
                                  (1, 64, 75, 75, 64), (2, 128, 16, 16, 40)])
@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("epi", ["plain", "stats", "full"])
def test_conv_plain_matrix_tma_equals_gather_path(zsg, case, bf16, epi):
    """x_plain (zsg_conv_params): 1x1 stride-1 convs fetch their A tiles by TMA from the plain [m, cin] matrix.  Same smem
    image, same MMAs, same epilogue as the cp.async gather path => bit-identical results (ragged last M tile, K that is
    not a multiple of the K block, every epilogue option)."""
    ops, geo = zsg
    B, cin, H, W, cout = case
    if bf16 and cin % 8:
        pytest.skip("bf16 images need cin % 8 == 0")
    g = torch.Generator().manual_seed(cin * 7 + cout)
    m = B * H * W
    x = torch.randn(m, cin, generator=g).cuda()
    w = (torch.randn(cout, cin, generator=g) / cin ** 0.5).cuda()
    rows = geo.conv_rows(B, H, W, cin, H, W, cout, 1, 0).cuda()
    kw = {}
    if epi == "full":
        kw = dict(bias=torch.randn(cout, generator=g).cuda(), out_mask=torch.randn(m, cout, generator=g).cuda(),
                  residual=torch.randn(m, cout, generator=g).cuda(), accumulate=True)
    if bf16:
        xi, wi = x.to(torch.bfloat16), w.to(torch.bfloat16)
        wa = w
    else:
        xi, wa, wi = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
        ops.split_act(x, xi, m, cin)
        ops.split_tf32(w, wa, wi, w.numel())
    ys, sts = [], []
    for plain in (False, True):
        y = torch.full((m, cout), 0.25, device="cuda")
        st = torch.zeros((m + 127) // 128 * 4 * 2 * cout, device="cuda") if epi == "stats" else None
        ops.ConvOp(x, wa, y, rows, m, cin, cout, 1, 1, w_lo=wi, x_lo=xi, stats=st, x_plain=plain, **kw)()
        ys.append(y)
        sts.append(st)
    torch.cuda.synchronize()
    assert torch.equal(ys[0], ys[1])
    if epi == "stats":
        assert torch.equal(sts[0], sts[1])
    if epi != "full":
        ref = (x.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t()) if bf16 else x @ w.t()
        assert rel_err(ys[1], ref) < (2e-5 if not bf16 else 1e-5)


@pytest.mark.parametrize("case", [(2, 64, 20, 19, 256, 1), (1, 256, 30, 30, 64, 1), (2, 128, 9, 9, 136, 3), (3, 72, 7, 5, 24, 3),
                                  (1, 256, 38, 38, 256, 3)])
@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("opts", ["bias_relu", "mask", "residual_acc", "residual_bf16", "full", "row_add"])
def test_conv_fragment_epilogue_equals_slab_epilogue(zsg, case, bf16, opts):
    """EPI_FRAGX (conv_tc.cu epilogue_fragx: bias / row_add / ReLU mask / residual / residual_bf16 / accumulate / ReLU from the
    fragment registers, chosen for plain [m, y_pitch] outputs with cout % 8 == 0) against the slab epilogue (y_pitch = 0):
    same accumulators, same order of operations => bit-identical outputs; ragged last M tile, cout not a multiple of the tile."""
    ops, geo = zsg
    B, cin, H, W, cout, k = case
    g = torch.Generator().manual_seed(cin * 5 + cout + k)
    m = B * H * W
    x = torch.randn(B, H, W, cin, generator=g).cuda()
    w = (torch.randn(cout, k, k, cin, generator=g) / (cin * k * k) ** 0.5).cuda()
    rows = geo.conv_rows(B, H, W, cin, H, W, cout, 1, k // 2).cuda()
    kw = {}
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    if opts == "bias_relu":
        kw = dict(bias=r(cout), out_relu=True)
    elif opts == "mask":
        kw = dict(out_mask=r(m, cout))
    elif opts == "residual_acc":
        kw = dict(residual=r(m, cout), accumulate=True)
    elif opts == "residual_bf16":
        kw = dict(residual=r(m, cout).to(torch.bfloat16), accumulate=True)
    elif opts == "full":
        kw = dict(bias=r(cout), out_mask=r(m, cout), residual=r(m, cout), accumulate=True, out_relu=True)
    elif opts == "row_add":
        tab = r(37, cout)
        idx = (torch.randint(0, 37, (m, 2), generator=g, dtype=torch.int32) * cout).cuda()
        kw = dict(bias=r(cout), out_relu=True, row_add=tab, row_add_idx=idx)
    if bf16:
        xi, wi, wa = x.to(torch.bfloat16), w.to(torch.bfloat16), w
    else:
        xi, wa, wi = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
        ops.split_act(x, xi, m, cin)
        ops.split_tf32(w, wa, wi, w.numel())
    ys = []
    for pitch in (0, cout):
        y = torch.full((m, cout), 0.25, device="cuda")
        ops.ConvOp(x, wa, y, rows, m, cin, cout, k, k, w_lo=wi, x_lo=xi, y_pitch=pitch, **kw)()
        ys.append(y)
    torch.cuda.synchronize()
    assert torch.equal(ys[0], ys[1])
    if opts == "bias_relu":
        xr, wr = (x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()) if bf16 else (x, w)
        ref = torch.relu(torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr.permute(0, 3, 1, 2), kw["bias"], padding=k // 2))
        assert rel_err(ys[1], ref.permute(0, 2, 3, 1).reshape(m, cout)) < 2e-5


@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("case", [(2, 256, 19, 19, 256, 3), (3, 48, 9, 11, 256, 3), (2, 64, 20, 20, 72, 1)])
def test_conv_epilogue_writes_the_output_operand_image(zsg, case, bf16):
    """zsg_conv_params.y_lo / y_img_bf16: the fragment epilogue stores the GEMM operand image of its OUTPUT (TF32 remainders
    / bfloat16 copy) next to y -- bit-identical to zsg_split_act / zsg_cast_bf16 over the finished tensor, y itself
    unchanged; 128- and 256-column tiles."""
    ops, geo = zsg
    B, cin, H, W, cout, k = case
    g = torch.Generator().manual_seed(cin + cout)
    m = B * H * W
    x = torch.randn(B, H, W, cin, generator=g).cuda()
    w = (torch.randn(cout, k, k, cin, generator=g) / (cin * k * k) ** 0.5).cuda()
    rows = geo.conv_rows(B, H, W, cin, H, W, cout, 1, k // 2).cuda()
    kw = dict(bias=torch.randn(cout, generator=g).cuda(), out_relu=True, out_mask=torch.randn(m, cout, generator=g).cuda())
    if bf16:
        xi, wi, wa = x.to(torch.bfloat16), w.to(torch.bfloat16), w
    else:
        xi, wa, wi = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
        ops.split_act(x, xi, m, cin)
        ops.split_tf32(w, wa, wi, w.numel())
    for impl in ((2, 3) if bf16 and cout % 256 == 0 else (0,)):
        y0, y1 = torch.zeros(m, cout, device="cuda"), torch.zeros(m, cout, device="cuda")
        img = torch.zeros(m, cout, device="cuda", dtype=torch.bfloat16 if bf16 else torch.float32)
        ops.ConvOp(x, wa, y0, rows, m, cin, cout, k, k, w_lo=wi, x_lo=xi, y_pitch=cout, impl=impl, **kw)()
        ops.ConvOp(x, wa, y1, rows, m, cin, cout, k, k, w_lo=wi, x_lo=xi, y_pitch=cout, impl=impl, y_img=img, **kw)()
        want = torch.zeros_like(img)
        if bf16:
            ops.cast_bf16(y0, want, y0.numel())
        else:
            ops.split_act(y0, want, m, cout)
        torch.cuda.synchronize()
        assert torch.equal(y0, y1)
        assert torch.equal(img.view(torch.int16 if bf16 else torch.int32), want.view(torch.int16 if bf16 else torch.int32))
    with pytest.raises(Exception):                           # not a plain [m, y_pitch] output: refused, never silently skipped
        ops.ConvOp(x, wa, y1, rows, m, cin, cout, k, k, w_lo=wi, x_lo=xi, y_pitch=0, y_img=img, **kw)()


@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16_store", "bf16_store_wide"])
@pytest.mark.parametrize("case", [(2, 256, 19, 19, 64, 1), (3, 64, 21, 17, 64, 3), (2, 512, 10, 10, 256, 1), (2, 128, 20, 20, 128, 3)])
def test_conv_epilogue_batchnorm_backward_sums(zsg, case, mode):
    """zsg_conv_params.bnb_*: the data gradient that writes dy of a BatchNorm+ReLU also leaves sum dz / sum dz * x per 32-row
    group (dz = dy where x * scale + shift > 0); zsg_bn_stats_partials + zsg_bn_bwd_center_sums turn them into the sums
    zsg_bn_bwd_reduce (mask_mode 1) produces from dy and x -- compared with that kernel and with torch; y is unchanged."""
    ops, geo = zsg
    B, cin, H, W, cout, k = case
    if mode == "bf16_store_wide" and cout % 256:
        pytest.skip("256-column tiles need cout % 256 == 0")
    bf16, store = mode != "fp32", mode.startswith("bf16_store")
    g = torch.Generator().manual_seed(cin + 3 * cout + k)
    m = B * H * W
    x = torch.randn(B, H, W, cin, generator=g).cuda()
    w = (torch.randn(cout, k, k, cin, generator=g) / (cin * k * k) ** 0.5).cuda()
    rows = geo.conv_rows(B, H, W, cin, H, W, cout, 1, k // 2).cuda()
    bx = (torch.randn(m, cout, generator=g) * 1.5 + 0.3).cuda()          # the BatchNorm's input at the output positions
    mean, var = bx.mean(0), bx.var(0, unbiased=False)
    invstd = (var + 1e-5).rsqrt()
    gamma, beta = (torch.rand(cout, generator=g) + 0.5).cuda(), (torch.randn(cout, generator=g) * 0.2).cuda()
    scale = gamma * invstd
    shift = beta - mean * scale
    if store:
        bx = bx.to(torch.bfloat16)
    if bf16:
        xi, wi, wa = x.to(torch.bfloat16), w.to(torch.bfloat16), w
    else:
        xi, wa, wi = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
        ops.split_act(x, xi, m, cin)
        ops.split_tf32(w, wa, wi, w.numel())
    parts = (m + 127) // 128 * 4
    ydt = torch.bfloat16 if store else torch.float32
    impl = 3 if mode == "bf16_store_wide" else (2 if bf16 else 0)
    y0, y1 = torch.zeros(m, cout, device="cuda", dtype=ydt), torch.zeros(m, cout, device="cuda", dtype=ydt)
    partials = torch.zeros(parts, 2, cout, device="cuda")
    ops.ConvOp(x, wa, y0, rows, m, cin, cout, k, k, w_lo=wi, x_lo=xi, y_pitch=cout, impl=impl)()
    ops.ConvOp(x, wa, y1, rows, m, cin, cout, k, k, w_lo=wi, x_lo=xi, y_pitch=cout, impl=impl,
               bnb=(bx, scale, shift, partials))()
    sums = torch.zeros(2 * cout, device="cuda", dtype=torch.float64)
    ops.bn_stats_partials(partials, parts, cout, sums)
    ops.bn_bwd_center_sums(sums, mean, invstd, cout)
    want = torch.zeros(2 * cout, device="cuda", dtype=torch.float64)
    ops.bn_bwd_reduce(y0, bx, mean, invstd, want, m, cout, mask_mode=1, scale=scale, shift=shift)
    torch.cuda.synchronize()
    assert torch.equal(y0.view(torch.int16 if store else torch.int32), y1.view(torch.int16 if store else torch.int32))
    dz = torch.where(bx.float() * scale + shift > 0, y0.float(), torch.zeros_like(y0, dtype=torch.float32)).double()
    ref = torch.cat([dz.sum(0), (dz * ((bx.double() - mean.double()) * invstd.double())).sum(0)])
    tol = 2e-5 * float(ref.abs().max())
    assert float((sums - want).abs().max()) < tol and float((sums - ref).abs().max()) < tol


def test_split_first_head_conv_equals_the_materialised_convolution(zsg):
    """a-6: conv(W, [feat | lang tiled | grid]) (mdl.py:69-104, 235-244) = conv(W_f, feat) + L[b, border class] + G[cell].
    Forward through zsg_conv_fwd with row_add against F.conv2d over the concatenated tensor, level by level; backward sums
    (per-tap column sums, dW_g) against autograd of the same."""
    ops, geo = zsg
    from zsg_b200 import spec
    from zsg_b200.anchors import cell_grid
    B, N, FC = 3, 256, 514
    sizes = spec.LEVEL_SIZES
    g = torch.Generator().manual_seed(11)
    feats = [torch.randn(B, 256, s, s, generator=g) for s in sizes]
    lang = torch.randn(B, 256, generator=g)
    W = (torch.randn(N, FC, 3, 3, generator=g) / (FC * 9) ** 0.5).requires_grad_(True)
    bias = torch.randn(N, generator=g)
    grids = [cell_grid(s, s) for s in sizes]                                   # [s, s, 2]
    outs = []
    for f, gr in zip(feats, grids):
        s = f.shape[-1]
        fused = torch.cat([f, lang.view(B, 256, 1, 1).expand(B, 256, s, s), gr.permute(2, 0, 1).unsqueeze(0).expand(B, 2, s, s)], 1)
        outs.append(F.relu(F.conv2d(fused, W, bias, padding=1)))
    G_ = [torch.randn(o.shape, generator=g) for o in outs]
    lang_r = lang.clone().requires_grad_(True)
    outs_r = []
    for f, gr in zip(feats, grids):
        s = f.shape[-1]
        fused = torch.cat([f, lang_r.view(B, 256, 1, 1).expand(B, 256, s, s), gr.permute(2, 0, 1).unsqueeze(0).expand(B, 2, s, s)], 1)
        outs_r.append(F.conv2d(fused, W, bias, padding=1))
    sum((o * gg).sum() for o, gg in zip(outs_r, G_)).backward()

    # ---- device side: level-major rows
    tot = spec.TOTAL_CELLS
    lvl_off = np.concatenate([[0], np.cumsum([B * s * s for s in sizes])]).tolist()
    M = B * tot
    rows_level_major = lambda ts: torch.cat([nhwc(t).reshape(-1, t.shape[1]) for t in ts]).cuda().contiguous()
    feat = rows_level_major(feats)
    grid = torch.cat([gr.reshape(-1, 2) for gr in grids])
    tabs = geo.head0_tables(B, sizes, grid)
    Wk = W.detach().permute(0, 2, 3, 1).contiguous().cuda().view(-1)           # [n][t][514]
    wf, wl, wg = torch.empty(N * 9 * 256, device="cuda"), torch.empty(2304 * 256, device="cuda"), torch.empty(2304 * 2, device="cuda")
    ops.copy_cols(Wk, FC, wf, 256, 2304, 256)
    ops.copy_cols(Wk[256:], FC, wl, 256, 2304, 256)
    ops.copy_cols(Wk[512:], FC, wg, 2, 2304, 2)
    V = (lang.cuda() @ wl.view(2304, 256).t()).contiguous()                    # the GEMM itself is tested elsewhere
    radd = torch.empty(B * 16 * N + tot * N, device="cuda")
    ops.head0_lang_grid_terms(V, wg, tabs["gridpatch"].cuda(), radd[:B * 16 * N], radd[B * 16 * N:], B, tot, N)
    tabs_f = []
    for li, s in enumerate(sizes):
        tabs_f.append(geo.conv_rows(B, s, s, 256, s, s, N, 1, 1, in_off=lvl_off[li] * 256, out_off=lvl_off[li] * N))
    rows = torch.cat(tabs_f).contiguous().cuda()
    hi, lo, flo = torch.empty_like(wf), torch.empty_like(wf), torch.empty_like(feat)
    ops.split_tf32(wf, hi, lo, wf.numel())
    ops.split_act(feat, flo, M, 256)
    y = torch.empty(M, N, device="cuda")
    ops.ConvOp(feat, hi, y, rows, M, 256, N, 3, 3, bias=bias.cuda(), out_relu=True, w_lo=lo, x_lo=flo, y_pitch=N, row_add=radd,
               row_add_idx=tabs["row_add_idx"].cuda())()
    torch.cuda.synchronize()
    assert rel_err(y, rows_level_major(outs)) < 2e-5
    # ---- backward sums
    dh = rows_level_major(G_)
    scr = torch.empty(B * 8 * 34 * N, device="cuda")
    St = torch.empty(B, 2304, device="cuda")
    g0 = torch.full((2304 * FC,), 7.0, device="cuda")
    ops.head0_backward_sums(dh, tabs["cell_base"].cuda(), tabs["cell_stride"].cuda(), tabs["cell_cls"].cuda(),
                            tabs["gridpatch"].cuda(), B, tot, N, scr, St, g0[512:], FC)
    torch.cuda.synchronize()
    gW = W.grad.permute(0, 2, 3, 1).reshape(2304, FC)                          # [(n,t)][514]
    assert rel_err(g0.view(2304, FC)[:, 512:], gW[:, 512:]) < 2e-5
    assert bool((g0.view(2304, FC)[:, :512] == 7.0).all())                      # only the two grid columns are written
    assert rel_err(St.t() @ lang.cuda(), gW[:, 256:512]) < 2e-5                # dW_l = St^T x lang
    assert rel_err(St @ wl.view(2304, 256), lang_r.grad) < 2e-5                # d lang = St x W_l
