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


def dgrad_rows(B, hin, win, cin, hout, wout, cout, R, stride, pad, dy_off=0, dx_off=0, dil=1):
    """Data gradient of the conv above: one row per INPUT pixel (b, y, x); the gathered tensor is dY
    (plane hout x wout, cout channels).  Used with flipped-transposed weights, in_div = stride, and the
    forward conv's tap spacing `dil` (flipped tap r' reads dY at y + pad - (R-1)*dil + r'*dil)."""
    b, y, x = np.meshgrid(np.arange(B), np.arange(hin), np.arange(win), indexing="ij")
    b, y, x = b.ravel(), y.ravel(), x.ravel()
    base = dy_off + b * (hout * wout * cout)
    out = dx_off + ((b * hin + y) * win + x) * cin
    n = b.shape[0]
    off = pad - (R - 1) * dil
    return _pack(base, y + off, x + off, np.full(n, hout), np.full(n, wout), out)


def dgrad_rows_s2_class(B, hin, win, cin, hout, wout, cout, ey, ex, dy_off=0, dx_off=0):
    """Data gradient of a 3x3 / stride 2 / pad 1 conv, restricted to the input pixels (y, x) = (2i + ey, 2j + ex).
    For one parity class only the filter rows r with (y + 1 - r) even contribute (r = 1 for ey = 0; r = 2, 0 for
    ey = 1, reading dY rows i, i + 1), so the class is a small stride-1 conv over dY with 1 or 2 taps per axis and
    no zero-stuffed taps at all (in_div = 1).  Rows are (b, i, j); `out` scatters to the class's pixels of dX."""
    ni, nj = (hin - ey + 1) // 2, (win - ex + 1) // 2
    b, i, j = np.meshgrid(np.arange(B), np.arange(ni), np.arange(nj), indexing="ij")
    b, i, j = b.ravel(), i.ravel(), j.ravel()
    base = dy_off + b * (hout * wout * cout)
    out = dx_off + ((b * hin + 2 * i + ey) * win + 2 * j + ex) * cin
    n = b.shape[0]
    return _pack(base, i, j, np.full(n, hout), np.full(n, wout), out)


def dgrad_rows_1x1_s2(B, hin, win, cin, hout, wout, cout, dy_off=0, dx_off=0):
    """Data gradient of a 1x1 / stride 2 conv: only the even input pixels receive anything, one row per dY pixel
    (b, p, q) scattering to dX pixel (2p, 2q); the caller accumulates into a dX that is already written."""
    b, p, q = np.meshgrid(np.arange(B), np.arange(hout), np.arange(wout), indexing="ij")
    b, p, q = b.ravel(), p.ravel(), q.ravel()
    base = dy_off + b * (hout * wout * cout)
    out = dx_off + ((b * hin + 2 * p) * win + 2 * q) * cin
    n = b.shape[0]
    return _pack(base, p, q, np.full(n, hout), np.full(n, wout), out)


def concat_rows(tables):
    return torch.cat(tables, dim=0).contiguous()


def head0_tables(B, level_sizes, grid_yx):
    """Host tables of the split first head conv (engine.py, csrc/elementwise.cu): per cell (all levels, level-major) the border
    class, the grid values under the nine taps (0 outside the level), the row of (sample 0, cell) and the row stride between
    samples in the level-major [sum_l B * cells_l, C] matrices; per output row the two offsets zsg_conv_params.row_add adds
    (language class sums L [B][16][256], then the grid term G [cells][256] in the same buffer).  grid_yx: [cells, 2] as
    mdl.py:77-104 tiles it (anchors.cell_grid per level)."""
    grid_yx = np.asarray(grid_yx, dtype=np.float32).reshape(-1, 2)
    total = sum(s * s for s in level_sizes)
    gp = np.zeros((total, 9, 2), dtype=np.float32)
    cls = np.zeros(total, dtype=np.int32)
    base = np.zeros(total, dtype=np.int32)
    stride = np.zeros(total, dtype=np.int32)
    ridx = []
    cb, row_off = 0, 0
    for S in level_sizes:
        ys, xs = np.divmod(np.arange(S * S), S)
        ry = (ys == 0).astype(np.int32) | ((ys == S - 1).astype(np.int32) << 1)
        rx = (xs == 0).astype(np.int32) | ((xs == S - 1).astype(np.int32) << 1)
        cls[cb:cb + S * S] = ry * 4 + rx
        for t in range(9):
            ny, nx = ys + t // 3 - 1, xs + t % 3 - 1
            ok = (ny >= 0) & (ny < S) & (nx >= 0) & (nx < S)
            nb = cb + np.clip(ny, 0, S - 1) * S + np.clip(nx, 0, S - 1)
            gp[cb:cb + S * S, t, :] = np.where(ok[:, None], grid_yx[nb], 0.0)
        base[cb:cb + S * S] = row_off + np.arange(S * S)
        stride[cb:cb + S * S] = S * S
        b = np.repeat(np.arange(B), S * S)
        local = np.tile(np.arange(S * S), B)
        ridx.append(np.stack([(b * 16 + cls[cb + local]) * 256, B * 16 * 256 + (cb + local) * 256], axis=1))
        cb += S * S
        row_off += B * S * S
    return dict(gridpatch=torch.from_numpy(gp.reshape(total, 18)), cell_cls=torch.from_numpy(cls),
                cell_base=torch.from_numpy(base), cell_stride=torch.from_numpy(stride),
                row_add_idx=torch.from_numpy(np.concatenate(ridx).astype(np.int32)).contiguous())
