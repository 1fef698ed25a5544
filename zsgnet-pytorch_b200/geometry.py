"""Row tables (zsg_row_t) that drive the implicit-GEMM kernels.  Built once per shape with torch
ops on the host and uploaded; see include/zsg_b200.h for the entry layout."""
import numpy as np
import torch


ROW_DTYPE = np.dtype([("base", "<i4"), ("y0", "<i2"), ("x0", "<i2"), ("hin", "<i2"), ("win", "<i2"), ("out", "<i4")])


def _pack(base, y0, x0, hin, win, out):
    n = base.shape[0]
    arr = np.zeros(n, dtype=ROW_DTYPE)
    arr["base"], arr["y0"], arr["x0"], arr["hin"], arr["win"], arr["out"] = base, y0, x0, hin, win, out
    assert arr.itemsize == 16
    return torch.from_numpy(arr.view(np.uint8).reshape(n, 16).copy())


def conv_rows(B, hin, win, cin, hout, wout, cout, stride, pad, in_off=0, out_off=0, out_row_stride=None):
    """Forward conv: one row per output pixel (b, p, q), b-major."""
    out_row_stride = cout if out_row_stride is None else out_row_stride
    b, p, q = np.meshgrid(np.arange(B), np.arange(hout), np.arange(wout), indexing="ij")
    b, p, q = b.ravel(), p.ravel(), q.ravel()
    base = in_off + b * (hin * win * cin)
    out = out_off + ((b * hout + p) * wout + q) * out_row_stride
    assert base.max() < 2 ** 31 and out.max() < 2 ** 31
    n = b.shape[0]
    return _pack(base, p * stride - pad, q * stride - pad, np.full(n, hin), np.full(n, win), out)


def dgrad_rows(B, hin, win, cin, hout, wout, cout, R, stride, pad, dy_off=0, dx_off=0):
    """Data gradient of the conv above: one row per INPUT pixel (b, y, x); the gathered tensor is dY
    (plane hout x wout, cout channels).  Used with flipped-transposed weights, in_div = stride."""
    b, y, x = np.meshgrid(np.arange(B), np.arange(hin), np.arange(win), indexing="ij")
    b, y, x = b.ravel(), y.ravel(), x.ravel()
    base = dy_off + b * (hout * wout * cout)
    out = dx_off + ((b * hin + y) * win + x) * cin
    n = b.shape[0]
    off = pad - (R - 1)
    return _pack(base, y + off, x + off, np.full(n, hout), np.full(n, wout), out)


def concat_rows(tables):
    return torch.cat(tables, dim=0).contiguous()
