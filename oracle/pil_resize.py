"""CPU oracle of the image resize on the data path (dat_loader.py:121: `img.resize((300, 300))`, then 136:
`pil2tensor(img).float().div_(255)`): a numpy restatement of Pillow's ImagingResample for 8-bit RGB.

TEST INFRASTRUCTURE ONLY (part of oracle/).  Third-party algorithm: Pillow (pinned pillow=6.1.0 / pillow-simd 5.3.0 in
conda_env_zsg.yml:108,229; the container has 12.2.0).  `Image.resize`'s DEFAULT filter changed in Pillow 7.0: NEAREST
under the pinned version, BICUBIC in the container -- both are restated here (src/libImaging/Resample.c:
precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc; Geometry.c: ImagingScaleAffine) and
pinned against PIL itself by tests/test_gpu_data_cpu.py.  The table builders are shared with the product
(zsg_b200/gpu_data.py builds the same tables; the test checks the two agree), the pixel loops below exist only here."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size, support=2.0, filt=bicubic_filter):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for box = (0, in_size): bounds [out,2] int32, kk [out,ksize] int32."""
    scale = filterscale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic(arr, out_w, out_h):
    """arr [h, w, 3] uint8 -> [out_h, out_w, 3] uint8 exactly like PIL.Image.resize((out_w, out_h), BICUBIC)."""
    h, w, _ = arr.shape
    need_h, need_v = out_w != w, out_h != h
    bh, kh = precompute_coeffs(w, out_w)
    bv, kv = precompute_coeffs(h, out_h)
    y_first = int(bv[0, 0])
    y_last = int(bv[-1, 0] + bv[-1, 1])
    img = arr
    if need_h:
        rows = arr[y_first:y_last].astype(np.int64)
        tmp = np.empty((y_last - y_first, out_w, 3), np.uint8)
        for xx in range(out_w):
            xmin, xmax = bh[xx]
            acc = (rows[:, xmin:xmin + xmax, :] * kh[xx, :xmax, None].astype(np.int64)).sum(1) + (1 << (PRECISION_BITS - 1))
            tmp[:, xx, :] = _clip8(acc)
        img = tmp
        bv = bv.copy()
        bv[:, 0] -= y_first
    if need_v:
        src = img.astype(np.int64)
        out = np.empty((out_h, img.shape[1], 3), np.uint8)
        for yy in range(out_h):
            ymin, ymax = bv[yy]
            acc = (src[ymin:ymin + ymax] * kv[yy, :ymax, None, None].astype(np.int64)).sum(0) + (1 << (PRECISION_BITS - 1))
            out[yy] = _clip8(acc)
        img = out
    return img.copy() if img is arr else img


def nearest_tables(in_w, in_h, out_w, out_h):
    """Geometry.c ImagingScaleAffine index tables (the running coordinate is accumulated in double, as there)."""
    def tab(n_in, n_out):
        a = float(np.float32(n_in) - np.float32(0.0)) / n_out
        o = 0.0 + a * 0.5
        t = np.empty(n_out, np.int32)
        for i in range(n_out):
            t[i] = -1 if o < 0.0 else int(o)
            o += a
        return t
    return tab(in_w, out_w), tab(in_h, out_h)


def resize_nearest(arr, out_w, out_h):
    """exactly like PIL.Image.resize((out_w, out_h), NEAREST) (the default before Pillow 7.0)"""
    h, w, _ = arr.shape
    xt, yt = nearest_tables(w, h, out_w, out_h)
    assert (xt >= 0).all() and (xt < w).all() and (yt >= 0).all() and (yt < h).all()
    return arr[yt][:, xt]


def to_unit_float(arr):
    """pil2tensor(img, np.float_).float().div_(255) (dat_loader.py:136, utils.py:521-529): CHW float32 in [0, 1]."""
    return (arr.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)).astype(np.float32)
