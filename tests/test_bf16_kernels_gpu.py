"""GPU parity tests of the bf16 operand path (BASELINE configs 3-5: "bf16 tensor-core convs").

The bf16 kernels multiply the bf16 images of both operands exactly (a bf16 x bf16 product is exact in fp32) and
accumulate in fp32, so the checker is a plain fp32 PyTorch convolution over the bf16-ROUNDED operands: what is
left is accumulation order, held to 2e-5 relative (tolerance of the fp32 path: 2e-5)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture(scope="module")
def zsg():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import zsg_b200  # noqa: F401
    from zsg_b200 import ops, geometry, _lib
    assert _lib.load().zsg_device_supported() == 1
    return ops, geometry


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def khwc(w):
    return w.permute(0, 2, 3, 1).contiguous()


def rel_err(a, b):
    a, b = a.detach().cpu().double().flatten(), b.detach().cpu().double().flatten()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rb(t):
    """bf16 round trip (round-to-nearest-even), kept in fp32."""
    return t.bfloat16().float()


def image(ops, t):
    """bf16 image of a contiguous fp32 tensor through zsg_cast_bf16."""
    out = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
    ops.cast_bf16(t, out, t.numel())
    return out


# B, cin, H, W, cout, k, stride, pad
CONV_CASES = [
    (2, 64, 19, 19, 64, 1, 1, 0),
    (2, 64, 19, 19, 256, 3, 1, 1),
    (3, 128, 20, 18, 128, 3, 2, 1),
    (2, 256, 10, 10, 512, 1, 2, 0),
    (2, 8, 61, 61, 64, 7, 2, 3),                    # stem-like with 8 padded channels: K = 392 = 6 K blocks + tail
    (2, 256, 5, 5, 45, 3, 1, 1),                    # ragged cout (head output)
    (1, 520, 10, 10, 256, 3, 1, 1),                 # head conv0 with padded fusion channels (K block spans two taps)
    (4, 64, 75, 75, 256, 1, 1, 0),                  # 176 x 2 tiles on 148 CTAs: several tiles per CTA, 1 K block each
    (2, 256, 38, 38, 256, 3, 1, 1),                 # 36 K blocks: several promotion chunks, ring wraps
    (2, 2048, 10, 10, 256, 3, 2, 1),                # K = 18432 (P6)
]


def test_cast_bf16_matches_torch(zsg):
    ops, _ = zsg
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(1000, 64, generator=g) * 3).cuda()
    x[0, :4] = torch.tensor([0.0, -0.0, 1e-40, 3.3895e38])
    sc, sh = (torch.rand(64, generator=g) + 0.5).cuda(), torch.randn(64, generator=g).cuda()
    out = torch.empty(1000, 64, dtype=torch.bfloat16, device="cuda")
    ops.split_act(x, out, 1000, 64)
    assert torch.equal(out.view(torch.int16), x.bfloat16().view(torch.int16))
    ops.split_act(x, out, 1000, 64, scale=sc, shift=sh, relu=True)
    want = F.relu((x.double() * sc.double() + sh.double()).float()).bfloat16()   # fma(x, scale, shift) like the kernel
    assert torch.equal(out.view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_bf16_vs_torch(zsg, case):
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    ref = F.relu(F.conv2d(rb(x), rb(w), bias, stride=stride, padding=pad))
    Ho, Wo = ref.shape[2], ref.shape[3]
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    xn, wk = nhwc(x), khwc(w)
    y = torch.full((B, Ho, Wo, cout), float("nan"), device="cuda")
    ops.ConvOp(xn, wk, y, rows, B * Ho * Wo, cin, cout, k, k, bias=bias, out_relu=True, x_lo=image(ops, xn),
               w_lo=image(ops, wk))()
    torch.cuda.synchronize()
    assert rel_err(y, nhwc(ref)) < TOL


def test_conv_bf16_statistics_and_run_to_run_bits(zsg):
    """BatchNorm statistics out of the bf16 kernel's epilogue; output identical bit for bit from run to run."""
    ops, geo = zsg
    B, cin, H, W, cout = 3, 128, 20, 18, 64
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / 34).cuda()
    ref = nhwc(F.conv2d(rb(x), rb(w), None, padding=1))
    rows = geo.conv_rows(B, H, W, cin, H, W, cout, 1, 1).cuda()
    m = B * H * W
    parts = (m + 127) // 128 * 4
    xn, wk = nhwc(x), khwc(w)
    xb, wb = image(ops, xn), image(ops, wk)
    outs = []
    for _ in range(2):
        y = torch.full((m, cout), float("nan"), device="cuda")
        stats = torch.zeros(parts, 2, cout, device="cuda")
        ops.ConvOp(xn, wk, y, rows, m, cin, cout, 3, 3, x_lo=xb, w_lo=wb, stats=stats)()
        torch.cuda.synchronize()
        outs.append((y, stats))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    y, stats = outs[0]
    assert rel_err(y, ref.view(m, cout)) < TOL
    np.testing.assert_allclose(stats[:, 0].sum(0).cpu().numpy(), y.sum(0).cpu().numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(stats[:, 1].sum(0).cpu().numpy(), (y * y).sum(0).cpu().numpy(), rtol=1e-4)


# B, cin, H, W, cout, k, stride, pad
WIDE_CASES = [
    (2, 256, 38, 38, 256, 3, 1, 1),                 # 36 K blocks, 23 tiles
    (4, 64, 75, 75, 256, 1, 1, 0),                  # 176 tiles on 148 CTAs: 1 K block each, both accumulators, x_plain
    (2, 256, 19, 19, 1024, 1, 1, 0),                # four column tiles
    (2, 2048, 10, 10, 512, 3, 2, 1),                # K = 18432: whole-K accumulation in TMEM
    (3, 128, 21, 17, 512, 3, 1, 1),                 # ragged last M tile
]


@pytest.mark.parametrize("case", WIDE_CASES)
@pytest.mark.parametrize("epi", ["plain_stats", "b16_stats", "bias_relu", "mask_residual_acc", "residual_bf16"])
def test_conv_bf16_wide_tiles(zsg, case, epi):
    """256-column tiles of the bf16 path (conv_tc.cu mma_loop_wide / conv_epilogue_wide; zsg_conv_params.impl = 3 forces
    them, 2 forces 128 columns): against the fp32 convolution over the bf16-rounded operands, against the 128-column kernel
    (same products, other accumulation order), BatchNorm statistics, every fragment-epilogue option, run-to-run bits."""
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(sum(case) + len(epi))
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    ref = nhwc(F.conv2d(rb(x), rb(w), None, stride=stride, padding=pad))
    Ho, Wo = ref.shape[1], ref.shape[2]
    m = B * Ho * Wo
    ref = ref.reshape(m, cout)
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    xn, wk = nhwc(x), khwc(w)
    xb, wb = image(ops, xn), image(ops, wk)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    kw, y0 = {}, float("nan")
    if epi == "bias_relu":
        kw = dict(bias=r(cout), out_relu=True)
        ref = F.relu(ref + kw["bias"])
    elif epi == "mask_residual_acc":
        kw = dict(out_mask=r(m, cout), residual=r(m, cout), accumulate=True)
        y0 = 0.25
        ref = torch.where(kw["out_mask"] > 0, ref, torch.zeros_like(ref)) + kw["residual"] + y0
    elif epi == "residual_bf16":
        kw = dict(residual=r(m, cout).bfloat16())
        ref = ref + kw["residual"].float()
    parts = (m + 127) // 128 * 4
    outs = []
    for impl in (2, 3, 3):
        if epi == "b16_stats":
            y = torch.zeros(m, cout, dtype=torch.bfloat16, device="cuda")
        else:
            y = torch.full((m, cout), y0, device="cuda")
        st = torch.zeros(parts, 2, cout, device="cuda") if epi.endswith("stats") else None
        ops.ConvOp(xn, wk, y, rows, m, cin, cout, k, k, x_lo=xb, w_lo=wb, stats=st, y_pitch=cout, impl=impl,
                   x_plain=(k == 1 and stride == 1), **kw)()
        torch.cuda.synchronize()
        outs.append((y, st))
    assert torch.equal(outs[1][0], outs[2][0])                                   # run to run
    if epi == "b16_stats":
        d = (outs[1][0].float() - rb(ref)).abs().max() / ref.abs().max()
        assert float(d) < 2.0 ** -7                                               # one bfloat16 ulp at the largest value
        # vs the 128-column kernel: fp32 sums differing in the last bits round to the neighbouring bfloat16 now and then
        a16, b16 = outs[0][0].float(), outs[1][0].float()
        assert (a16 != b16).float().mean() < 2e-2 and bool(((a16 - b16).abs() <= 2.0 ** -7 * torch.maximum(a16.abs(), b16.abs()) + 1e-4).all())
    else:
        # whole-K accumulation in TMEM: the tensor core adds into the fp32 accumulator with truncation, a bias that grows
        # with K (the 128-column kernel promotes every 8 K blocks into registers); K = 18432 is not a layer the product
        # heuristic gives to this kernel
        tol = TOL if cin * k * k <= 8192 else 5e-5
        assert rel_err(outs[1][0], ref) < tol
        assert rel_err(outs[1][0], outs[0][0]) < tol
    if st is not None:
        assert torch.equal(outs[1][1], outs[2][1])
        np.testing.assert_allclose(outs[1][1][:, 0].sum(0).cpu().numpy(), ref.sum(0).cpu().numpy(), rtol=1e-4, atol=2e-3)
        np.testing.assert_allclose(outs[1][1][:, 1].sum(0).cpu().numpy(), (ref * ref).sum(0).cpu().numpy(), rtol=1e-4)
        np.testing.assert_allclose(outs[1][1].cpu().numpy(), outs[0][1].cpu().numpy(), rtol=1e-3, atol=1e-3)


DGRAD_CASES = [(2, 64, 19, 19, 256, 3, 1, 1), (3, 128, 20, 18, 128, 3, 2, 1), (2, 256, 10, 10, 512, 1, 2, 0),
               (2, 2048, 10, 10, 256, 3, 2, 1), (2, 256, 5, 5, 45, 3, 1, 1), (2, 64, 8, 8, 64, 1, 1, 0)]


@pytest.mark.parametrize("case", DGRAD_CASES)
def test_conv_dgrad_bf16_vs_torch(zsg, case):
    """Data gradient = the forward kernel over flipped-transposed weights, output masked / accumulated in the epilogue."""
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, cin, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    y = F.conv2d(x, rb(w), None, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(rb(dy))
    Ho, Wo = y.shape[2], y.shape[3]
    wk = khwc(w)
    wt = torch.empty(cin, k, k, cout, device="cuda")
    ops.weight_transpose_flip(wk, wt, cout, k, k, cin)
    dyn = nhwc(dy)
    cp = (cout + 7) // 8 * 8
    if cp != cout:                                   # the 45-channel head output: pad to 48 (engine does the same)
        wtp = torch.empty(cin, k, k, cp, device="cuda")
        ops.pad_channels(wt, wtp, cin * k * k, cout, cp)
        dyp = torch.empty(B, Ho, Wo, cp, device="cuda")
        ops.pad_channels(dyn, dyp, B * Ho * Wo, cout, cp)
        wt, dyn = wtp, dyp
    rows = geo.dgrad_rows(B, H, W, cin, Ho, Wo, cp, k, stride, pad).cuda()
    mask = torch.randn(B, H, W, cin, generator=g).cuda()
    prev = torch.randn(B, H, W, cin, generator=g).cuda()
    dx = prev.clone()
    ops.ConvOp(dyn, wt, dx, rows, B * H * W, cp, cin, k, k, in_div=stride, x_lo=image(ops, dyn), w_lo=image(ops, wt),
               out_mask=mask, accumulate=True)()
    torch.cuda.synchronize()
    want = prev + nhwc(x.grad) * (mask > 0)
    assert rel_err(dx, want) < TOL


def test_conv_dgrad_bf16_stride2_parity_classes(zsg):
    """3x3 / stride-2 data gradient as four dense 1- or 2-tap convs over dy (engine.conv_dgrad), bf16 operands."""
    ops, geo = zsg
    B, cin, H, W, cout = 2, 64, 19, 19, 128
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, cin, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / 24).cuda()
    y = F.conv2d(x, rb(w), None, stride=2, padding=1)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(rb(dy))
    Ho, Wo = y.shape[2], y.shape[3]
    wt = torch.empty(cin, 3, 3, cout, device="cuda")
    ops.weight_transpose_flip(khwc(w), wt, cout, 3, 3, cin)
    dyn = nhwc(dy)
    dyb = image(ops, dyn)
    dx = torch.full((B, H, W, cin), float("nan"), device="cuda")
    for ey in (0, 1):
        for ex in (0, 1):
            sr, sc = (slice(1, 2), slice(0, 3, 2))[ey], (slice(1, 2), slice(0, 3, 2))[ex]
            wc = wt[:, sr, sc, :].contiguous()
            rows = geo.dgrad_rows_s2_class(B, H, W, cin, Ho, Wo, cout, ey, ex).cuda()
            ops.ConvOp(dyn, wc, dx, rows, rows.shape[0], cout, cin, 1 + ey, 1 + ex, x_lo=dyb, w_lo=image(ops, wc))()
    torch.cuda.synchronize()
    assert rel_err(dx, nhwc(x.grad)) < TOL


WGRAD_CASES = [(2, 64, 19, 19, 256, 3, 1, 1), (3, 128, 20, 18, 128, 3, 2, 1), (2, 8, 61, 61, 64, 7, 2, 3),
               (2, 256, 5, 5, 45, 3, 1, 1), (4, 1024, 19, 19, 256, 1, 1, 0), (2, 520, 10, 10, 256, 3, 1, 1),
               (2, 264, 7, 9, 128, 1, 1, 0), (8, 64, 75, 75, 64, 1, 1, 0), (1, 256, 3, 3, 256, 3, 1, 1)]


@pytest.mark.parametrize("case", WGRAD_CASES)
@pytest.mark.parametrize("split_k", [0, 1])
def test_conv_wgrad_bf16_vs_torch(zsg, case, split_k):
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(13)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda().requires_grad_(True)
    y = F.conv2d(rb(x), w, None, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(rb(dy))
    Ho, Wo = y.shape[2], y.shape[3]
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    xn, dyn = nhwc(x), nhwc(dy)
    pitch = (cout + 7) // 8 * 8
    if pitch != cout:
        dyp = torch.empty(B, Ho, Wo, pitch, device="cuda")
        ops.pad_channels(dyn, dyp, B * Ho * Wo, cout, pitch)
        dyn = dyp
        rows = geo.conv_rows(B, H, W, cin, Ho, Wo, pitch, stride, pad).cuda()
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    ops.WgradOp(xn, dyn, dw, rows, B * Ho * Wo, cin, cout, k, k, x_lo=image(ops, xn), dy_lo=image(ops, dyn),
                dy_pitch=pitch, split_k=split_k)()
    torch.cuda.synchronize()
    assert rel_err(dw, khwc(w.grad)) < 3e-5


WIDE_WGRAD_CASES = [(2, 64, 19, 19, 256, 3, 1, 1), (4, 1024, 19, 19, 256, 1, 1, 0), (2, 520, 10, 10, 256, 3, 1, 1),
                    (1, 256, 3, 3, 256, 3, 1, 1), (6, 256, 44, 44, 256, 3, 1, 1), (3, 128, 40, 36, 512, 3, 2, 1),
                    (70, 64, 10, 10, 256, 1, 1, 0)]


@pytest.mark.parametrize("case", WIDE_WGRAD_CASES)
def test_conv_wgrad_bf16_wide_tiles(zsg, case):
    """wgrad_bf16_wide_kernel (256-column tiles, persistent over units of <= 64 K blocks; zsg_wgrad_params.impl = 3 forces
    it): against autograd over the bf16-rounded operands and against the 128-column kernel; several units per CTA, a unit
    of one K block, ragged last K block, row tiles past Kt."""
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(17)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda().requires_grad_(True)
    y = F.conv2d(rb(x), w, None, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(rb(dy))
    Ho, Wo = y.shape[2], y.shape[3]
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    xn, dyn = nhwc(x), nhwc(dy)
    xb, dyb = image(ops, xn), image(ops, dyn)
    outs = []
    for impl in (2, 3):
        dw = torch.zeros(cout, k, k, cin, device="cuda")
        ops.WgradOp(xn, dyn, dw, rows, B * Ho * Wo, cin, cout, k, k, x_lo=xb, dy_lo=dyb, dy_pitch=cout, impl=impl)()
        torch.cuda.synchronize()
        outs.append(dw)
    assert rel_err(outs[1], khwc(w.grad)) < 3e-5
    assert rel_err(outs[1], outs[0]) < 3e-5


def test_bn_kernels_write_bf16_images(zsg):
    """bn_apply / bn_bwd_apply write the bf16 image of their output next to the fp32 tensor."""
    ops, _ = zsg
    rows, c = 777, 64
    g = torch.Generator().manual_seed(2)
    x, r = torch.randn(rows, c, generator=g).cuda(), torch.randn(rows, c, generator=g).cuda()
    sc, sh = (torch.rand(c, generator=g) + 0.5).cuda(), torch.randn(c, generator=g).cuda()
    y, yb = torch.empty(rows, c, device="cuda"), torch.empty(rows, c, dtype=torch.bfloat16, device="cuda")
    ops.bn_apply(x, sc, sh, y, rows, c, True, r=r, y_lo=yb)
    y2 = torch.empty_like(y)
    ops.bn_apply(x, sc, sh, y2, rows, c, True, r=r)
    torch.cuda.synchronize()
    assert torch.equal(y, y2) and torch.equal(yb.view(torch.int16), y.bfloat16().view(torch.int16))
    dy = torch.randn(rows, c, generator=g).cuda()
    mean, invstd = x.mean(0), 1.0 / (x.var(0, unbiased=False) + 1e-5).sqrt()
    gamma = (torch.rand(c, generator=g) + 0.5).cuda()
    sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    ops.bn_bwd_reduce(dy, x, mean, invstd, sums, rows, c)
    dx, dxb = torch.empty(rows, c, device="cuda"), torch.empty(rows, c, dtype=torch.bfloat16, device="cuda")
    dgam, dbet = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    ops.bn_bwd_apply(dy, x, mean, invstd, gamma, sums, dx, dgam, dbet, rows, c, dx_lo=dxb)
    dx2 = torch.empty_like(dx)
    ops.bn_bwd_apply(dy, x, mean, invstd, gamma, sums, dx2, dgam, dbet, rows, c)
    torch.cuda.synchronize()
    # the image is the rounding of the fp32 value the SAME launch wrote; the fp32-image launch runs the 8-channel kernel
    # (centred form with per-channel coefficients in registers): same formula, last-ulp association differences
    assert torch.equal(dxb.view(torch.int16), dx.bfloat16().view(torch.int16))
    assert float((dx - dx2).abs().max()) <= 2e-6 * float(dx.abs().max())


# ------------------------------------------------------------------------------------------------------------------
# bf16 storage: trunk activations of the bf16 engine are bfloat16 tensors (their own GEMM operand images)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(2, 64, 19, 19, 64, 1, 1, 0), (3, 128, 20, 18, 128, 3, 2, 1), (4, 64, 75, 75, 256, 1, 1, 0)])
def test_conv_bf16_output_storage(zsg, case):
    """zsg_conv_params.y_bf16: the same accumulators stored as bfloat16 (round to nearest even) instead of fp32; the
    BatchNorm statistics still come from the fp32 accumulators (identical to the fp32-output launch)."""
    ops, geo = zsg
    B, cin, H, W, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(31)
    x = nhwc(torch.randn(B, cin, H, W, generator=g)).cuda()
    w = khwc(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    m = B * Ho * Wo
    rows = geo.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad).cuda()
    xi, wi = image(ops, x), image(ops, w)
    nst = (m + 127) // 128 * 4 * 2 * cout
    y32, st32 = torch.empty(m, cout, device="cuda"), torch.zeros(nst, device="cuda")
    y16, st16 = torch.full((m, cout), 7.0, dtype=torch.bfloat16, device="cuda"), torch.zeros(nst, device="cuda")
    plain = k == 1 and stride == 1
    ops.ConvOp(x, w, y32, rows, m, cin, cout, k, k, w_lo=wi, x_lo=xi, stats=st32, x_plain=plain)()
    ops.ConvOp(x, w, y16, rows, m, cin, cout, k, k, w_lo=wi, x_lo=xi, stats=st16, x_plain=plain)()
    torch.cuda.synchronize()
    assert torch.equal(y16, y32.bfloat16()) and torch.equal(st16, st32)


@pytest.mark.parametrize("rows,c", [(1000, 64), (333, 256), (77, 2048), (129, 4)])
def test_bf16_storage_batchnorm_passes(zsg, rows, c):
    """zsg_act_b16 / zsg_bn_apply_b16 / zsg_bn_bwd_reduce_b16 / zsg_bn_bwd_apply_b16 against the fp32 kernels run on the same
    (bfloat16-representable) values: bit-equal images, sums to fp32 rounding.  c = 4 takes the 4-channel fallback."""
    ops, _ = zsg
    g = torch.Generator().manual_seed(rows + c)
    R = lambda *s: torch.randn(*s, generator=g).cuda()
    xb, rb_ = R(rows, c).bfloat16(), R(rows, c).bfloat16()
    x, r = xb.float(), rb_.float()
    sc, sh, rsc, rsh = R(c).abs() + 0.5, R(c) * 0.3, R(c).abs() + 0.5, R(c) * 0.3
    # BatchNorm + ReLU image
    z16, z32 = torch.empty_like(xb), torch.empty_like(xb)
    ops.act_b16(xb, z16, rows, c, scale=sc, shift=sh, relu=True)
    ops.split_act(x, z32, rows, c, scale=sc, shift=sh, relu=True)
    assert torch.equal(z16, z32)
    # bottleneck tail, both shortcut kinds
    for kw16, kw32 in ((dict(r=rb_), dict(r=r)), (dict(r=rb_, rscale=rsc, rshift=rsh), dict(r=r, rscale=rsc, rshift=rsh))):
        y16, y32, i32 = torch.empty_like(xb), torch.empty_like(x), torch.empty_like(xb)
        ops.bn_apply(xb, sc, sh, y16, rows, c, True, **kw16)
        ops.bn_apply(x, sc, sh, y32, rows, c, True, y_lo=i32, **kw32)
        assert torch.equal(y16, i32)
    # backward: reduce + apply, every mask mode, fp32 and bfloat16 gradients
    mean, invstd, gamma = x.mean(0), 1.0 / (x.var(0, unbiased=False) + 1e-5).sqrt(), R(c).abs() + 0.5
    act16 = y16
    for mode in (0, 1, 2):
        for gdt in (torch.float32, torch.bfloat16):
            dy = R(rows, c).to(gdt)
            s16, s32 = torch.zeros(2 * c, dtype=torch.float64, device="cuda"), torch.zeros(2 * c, dtype=torch.float64, device="cuda")
            for zdt in ((torch.float32, torch.bfloat16) if mode == 2 else (None,)):
                s16.zero_()
                dz16 = torch.empty(rows, c, dtype=zdt, device="cuda") if zdt is not None else None
                ops.bn_bwd_reduce(dy, xb, mean, invstd, s16, rows, c, mask_mode=mode, scale=sc, shift=sh,
                                  act_out=act16 if mode == 2 else None, dz_out=dz16)
                dz32 = torch.empty(rows, c, device="cuda") if mode == 2 else None
                s32.zero_()
                ops.bn_bwd_reduce(dy.float(), x, mean, invstd, s32, rows, c, mask_mode=mode, scale=sc, shift=sh,
                                  act_out=act16.float() if mode == 2 else None, dz_out=dz32)
                torch.cuda.synchronize()
                scale_ = s32.abs().max().item() + 1.0
                assert float((s16 - s32).abs().max()) < 2e-5 * scale_ * max(1.0, float(mean.abs().max())), (mode, gdt, zdt)
                if mode == 2:
                    assert torch.equal(dz16.float(), dz32.to(zdt).float())
            dx16, dx32, dxi = torch.empty_like(xb), torch.empty_like(x), torch.empty_like(xb)
            dga, dbe, dga2, dbe2 = (torch.empty(c, device="cuda") for _ in range(4))
            ops.bn_bwd_apply(dy, xb, mean, invstd, gamma, s32, None, dga, dbe, rows, c, mask_mode=mode, scale=sc, shift=sh,
                             act_out=act16 if mode == 2 else None, dx_lo=dx16)
            ops.bn_bwd_apply(dy.float(), x, mean, invstd, gamma, s32, dx32, dga2, dbe2, rows, c, mask_mode=mode, scale=sc,
                             shift=sh, act_out=act16.float() if mode == 2 else None, dx_lo=dxi)
            torch.cuda.synchronize()
            assert torch.equal(dga, dga2) and torch.equal(dbe, dbe2)
            # same formula, different association (A*dz + B*x + C): a few fp32 ulps before the bf16 rounding
            d = (dx16.float() - dx32).abs()
            assert float(d.max()) <= 2 ** -7 * float(dx32.abs().max()), (mode, gdt)
            assert float((d > 2 ** -8 * dx32.abs().clamp_min(1e-3)).float().mean()) < 1e-3


def test_bf16_storage_stem_pool(zsg):
    ops, _ = zsg
    g = torch.Generator().manual_seed(5)
    B, H, W, C = 2, 30, 26, 64
    xb = torch.randn(B, H, W, C, generator=g).cuda().bfloat16()
    sc, sh = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda() * 0.2
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    y16, a16 = torch.empty(B, Ho, Wo, C, dtype=torch.bfloat16, device="cuda"), torch.empty(B * Ho * Wo * C, dtype=torch.uint8, device="cuda")
    y32, a32 = torch.empty(B, Ho, Wo, C, device="cuda"), torch.empty_like(a16)
    ops.maxpool_bn_relu_fwd(xb, sc, sh, y16, a16, B, H, W, C, Ho, Wo)
    ops.maxpool_bn_relu_fwd(xb.float(), sc, sh, y32, a32, B, H, W, C, Ho, Wo)
    torch.cuda.synchronize()
    assert torch.equal(y16, y32.bfloat16()) and torch.equal(a16, a32)
