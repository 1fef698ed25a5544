"""CPU checks of the row tables (geometry.py) that turn forward convs, data gradients, the parity-split stride-2
data gradients and the even-pixel 1x1/s2 scatter into the one gather the CUDA kernels implement.  The gather is
emulated in numpy exactly as include/zsg_b200.h defines it and compared with torch's own conv / conv gradients."""
import numpy as np
import torch
import torch.nn.functional as F

import zsg_b200  # noqa: F401
from zsg_b200 import geometry


def emulate(x, w, table, cin, cout, r, s, y, in_div=1, accumulate=False):
    """y[row.out + n] (+)= sum_{tr,ts,c} x[row.base + ((y0+tr)/div*win + (x0+ts)/div)*cin + c] * w[n][tr][ts][c]."""
    rows = table.numpy().view(geometry.ROW_DTYPE).reshape(-1)
    for e in rows:
        acc = np.zeros(cout)
        for tr in range(r):
            for ts in range(s):
                yy, xx = int(e["y0"]) + tr, int(e["x0"]) + ts
                if in_div == 2:
                    if (yy | xx) & 1:
                        continue
                    yy, xx = yy >> 1, xx >> 1
                if not (0 <= yy < e["hin"] and 0 <= xx < e["win"]):
                    continue
                o = int(e["base"]) + (yy * int(e["win"]) + xx) * cin
                acc += w[:, tr, ts, :] @ x[o:o + cin]
        o = int(e["out"])
        y[o:o + cout] = acc + (y[o:o + cout] if accumulate else 0)
    return y


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def test_forward_rows_match_conv2d():
    g = torch.Generator().manual_seed(0)
    B, cin, H, W, cout, k, stride, pad = 2, 4, 7, 6, 3, 3, 2, 1
    x, w = torch.randn(B, cin, H, W, generator=g).double(), torch.randn(cout, cin, k, k, generator=g).double()
    ref = F.conv2d(x, w, stride=stride, padding=pad)
    Ho, Wo = ref.shape[2:]
    t = geometry.conv_rows(B, H, W, cin, Ho, Wo, cout, stride, pad)
    y = emulate(nhwc(x).numpy().ravel(), w.permute(0, 2, 3, 1).numpy(), t, cin, cout, k, k, np.zeros(B * Ho * Wo * cout))
    assert np.allclose(y, nhwc(ref).numpy().ravel(), atol=1e-12)


def _dgrad_setup(k, stride, pad, H=7, W=6, B=2, cin=4, cout=3, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cin, H, W, generator=g).double().requires_grad_(True)
    w = torch.randn(cout, cin, k, k, generator=g).double()
    y = F.conv2d(x, w, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g).double()
    y.backward(dy)
    wt = w.permute(1, 2, 3, 0).flip(1, 2).contiguous().numpy()           # [cin][R][S][cout], taps flipped (weight_transpose_flip)
    return B, cin, H, W, cout, y.shape[2], y.shape[3], nhwc(dy).numpy().ravel(), wt, nhwc(x.grad).numpy().ravel()


def test_zero_stuffed_dgrad_rows():
    for k, stride, pad in ((3, 1, 1), (3, 2, 1), (1, 2, 0)):
        B, cin, H, W, cout, Ho, Wo, dy, wt, want = _dgrad_setup(k, stride, pad)
        t = geometry.dgrad_rows(B, H, W, cin, Ho, Wo, cout, k, stride, pad)
        dx = emulate(dy, wt, t, cout, cin, k, k, np.zeros(B * H * W * cin), in_div=stride)
        assert np.allclose(dx, want, atol=1e-12), (k, stride, pad)


def test_parity_class_dgrad_equals_zero_stuffed():
    """3x3 / stride 2 / pad 1: four dense launches (1 or 2 taps per axis, sliced flipped weights) cover every input pixel."""
    for H, W in ((7, 6), (8, 8), (5, 9)):
        B, cin, Hh, Ww, cout, Ho, Wo, dy, wt, want = _dgrad_setup(3, 2, 1, H=H, W=W)
        dx = np.full(B * H * W * cin, np.nan)
        for ey in (0, 1):
            for ex in (0, 1):
                sr, sc = (slice(1, 2), slice(0, 3, 2))[ey], (slice(1, 2), slice(0, 3, 2))[ex]
                wc = np.ascontiguousarray(wt[:, sr, sc, :])
                t = geometry.dgrad_rows_s2_class(B, H, W, cin, Ho, Wo, cout, ey, ex)
                emulate(dy, wc, t, cout, cin, 1 + ey, 1 + ex, dx)
        assert not np.isnan(dx).any()
        assert np.allclose(dx, want, atol=1e-12), (H, W)


def test_even_pixel_scatter_for_1x1_stride2():
    B, cin, H, W, cout, Ho, Wo, dy, wt, want = _dgrad_setup(1, 2, 0, H=7, W=8)
    g = np.random.default_rng(0).standard_normal(B * H * W * cin)            # dx already holds another branch's gradient
    dx = g.copy()
    t = geometry.dgrad_rows_1x1_s2(B, H, W, cin, Ho, Wo, cout)
    emulate(dy, wt, t, cout, cin, 1, 1, dx, accumulate=True)
    assert np.allclose(dx, g + want, atol=1e-12)


def test_ssd_vgg_layer_geometry_matches_torch_modules():
    """spec.vgg_layers() / VGG_EXTRAS against torch's own shape arithmetic for ssd_vgg.py:111-154 (300x300 input):
    pool ceil mode, the dilated conv6, the unpadded 3x3 extras; the six sources must be 38, 19, 10, 5, 3, 1."""
    import torch
    import torch.nn as nn
    from zsg_b200 import spec
    x = torch.zeros(1, 3, 300, 300)
    sizes = []
    for i, L in enumerate(spec.vgg_layers()):
        if L[0] == "conv":
            x = nn.Conv2d(L[1], L[2], L[3], padding=L[4], dilation=L[5])(x)
        elif L[0] == "pool":
            x = nn.MaxPool2d(L[1], L[2], L[3], ceil_mode=L[4])(x)
        if i == 22:
            sizes.append(x.shape[-1])
    sizes.append(x.shape[-1])
    assert x.shape[1] == 1024
    for i, (ci, co, k, s, p) in enumerate(spec.VGG_EXTRAS):
        assert x.shape[1] == ci
        x = nn.Conv2d(ci, co, k, stride=s, padding=p)(x)
        if i % 2 == 1:
            sizes.append(x.shape[-1])
    assert tuple(sizes) == spec.LEVEL_SIZES
    names = [n for n, _, _ in spec.trainable_specs("ssd_vgg")]
    assert len(names) == len(set(names)) and names[0] == "backbone.encoder.vgg.0.weight"
    assert not set(names) & {n for n, _, _ in spec.unused_specs("ssd_vgg")}


def test_dilated_dgrad_rows_mirror_forward_taps():
    """Data-gradient table of a dilated conv: flipped tap r' of input pixel y reads dY at y + pad - (R-1-r')*dil, i.e. the
    output pixels whose forward tap (R-1-r') touched y."""
    import numpy as np
    from zsg_b200 import geometry
    B, H, C, R, pad, dil = 1, 19, 8, 3, 6, 6
    t = geometry.dgrad_rows(B, H, H, C, H, H, C, R, 1, pad, dil=dil).numpy().view(geometry.ROW_DTYPE).reshape(-1)
    y, x = 7, 11
    e = t[y * H + x]
    for r in range(R):
        p = e["y0"] + r * dil                       # dY row read by flipped tap r
        fwd_tap = R - 1 - r
        assert p * 1 - pad + fwd_tap * dil == y     # forward: output p reads input p - pad + tap*dil
    assert e["x0"] == x + pad - (R - 1) * dil and e["out"] == (y * H + x) * C
