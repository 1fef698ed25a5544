"""Device-side data path (the reference's dat_loader.py:98-146).

The reference resizes every image with Pillow on a DataLoader worker, converts it to a float tensor and looks the phrase's
word vectors up, per sample, on the CPU.  At ~1200 pairs/s per GPU that is several cores per GPU for the resize alone.
Here a worker only decodes the JPEG and tokenises; the batch is then assembled on the GPU in three launches
(csrc/data.cu): Pillow's two-pass fixed-point resize bit for bit, /255 into NCHW float32, and the embedding gather.

Host side of the resize: the coefficient tables of Pillow's ImagingResample (Resample.c precompute_coeffs /
normalize_coeffs_8bpc; Geometry.c ImagingScaleAffine for NEAREST), built per distinct source size and cached.
`resample`: "bicubic" = Image.resize's default since Pillow 7.0 (what the reference computes in this container), "nearest" =
its default under the reference's pinned pillow 6.1.0 (conda_env_zsg.yml:108)."""
import math

import numpy as np
import torch

from . import ops
from ._lib import call, ptr, stream

PRECISION_BITS = 32 - 8 - 2
DESC = np.dtype([("src_off", "<i8"), ("tmp_off", "<i8"), ("h", "<i4"), ("w", "<i4"), ("y_first", "<i4"), ("n_rows", "<i4"),
                 ("hk_off", "<i4"), ("vk_off", "<i4"), ("hksize", "<i4"), ("vksize", "<i4")])
assert DESC.itemsize == 48


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1, np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


def bicubic_table(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc (box = the whole axis), vectorised over the output axis with
    the C code's operation order (sequential weight sum).  Returns bounds [out, 2] and coefficients [out, ksize], int32."""
    scale = filterscale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = 0.0 + (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # (int) truncation of a non-negative or clamped value
    xmin = np.where(center - support + 0.5 < 0, 0, xmin)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    w = _bicubic((x + xmin[:, None] - center[:, None] + 0.5) * ss)
    w = np.where(x < xmax[:, None], w, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for j in range(ksize):                                                     # the C loop's order: pairwise np.sum would round differently
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww[:, None] != 0.0, ww[:, None], 1.0), w)
    kk = np.where(w < 0, (-0.5 + w * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + w * (1 << PRECISION_BITS)).astype(np.int64))
    kk = np.where(x < xmax[:, None], kk, 0).astype(np.int32)
    return np.stack([xmin, xmax], axis=1).astype(np.int32), kk


def nearest_table(n_in, n_out):
    """Geometry.c ImagingScaleAffine: source index per output position (coordinate accumulated in double, as there)."""
    a = float(np.float32(n_in) - np.float32(0.0)) / n_out
    o = 0.0 + a * 0.5
    t = np.empty(n_out, np.int32)
    for i in range(n_out):
        t[i] = -1 if o < 0.0 else int(o)
        o += a
    assert t.min() >= 0 and t.max() < n_in
    return t


class ResizePlan:
    """Tables and row window for one source size."""

    def __init__(self, h, w, out_h, out_w, resample):
        self.h, self.w = h, w
        if resample == "nearest":
            self.htab, self.vtab = nearest_table(w, out_w), nearest_table(h, out_h)
            self.hksize = self.vksize = 0
            self.y_first, self.n_rows = 0, 0
            return
        bh, kh = bicubic_table(w, out_w)
        bv, kv = bicubic_table(h, out_h)
        self.y_first = int(bv[0, 0])
        self.n_rows = int(bv[-1, 0] + bv[-1, 1]) - self.y_first
        bv = bv.copy()
        bv[:, 0] -= self.y_first
        self.hksize, self.vksize = kh.shape[1], kv.shape[1]
        self.htab = np.concatenate([bh, kh], axis=1).reshape(-1)
        self.vtab = np.concatenate([bv, kv], axis=1).reshape(-1)


class GpuImageStage:
    """list of decoded images (uint8 [h, w, 3] numpy arrays or tensors, any sizes) -> float32 [B, 3, out_h, out_w] on the
    device, bit-identical to `pil2tensor(img.resize((out_w, out_h))).float().div_(255)` of the reference."""

    def __init__(self, device, out_hw=(300, 300), resample="bicubic"):
        assert resample in ("bicubic", "nearest"), resample
        self.device, self.out_h, self.out_w, self.resample = torch.device(device), out_hw[0], out_hw[1], resample
        self._plans = {}
        self._tables = None            # device copy of all tables seen so far
        self._tab_off = {}             # (h, w) -> (hk_off, vk_off)
        self._tab_host = []
        self._tab_len = 0

    def plan(self, h, w):
        key = (h, w)
        if key not in self._plans:
            p = ResizePlan(h, w, self.out_h, self.out_w, self.resample)
            self._plans[key] = p
            self._tab_off[key] = (self._tab_len, self._tab_len + p.htab.size)
            self._tab_host += [p.htab, p.vtab]
            self._tab_len += p.htab.size + p.vtab.size
            self._tables = None        # re-upload lazily (new source sizes are rare after the first epochs)
        return self._plans[key]

    def __call__(self, images, out=None):
        n = len(images)
        arrs = [np.ascontiguousarray(im.numpy() if isinstance(im, torch.Tensor) else im) for im in images]
        desc = np.zeros(n, dtype=DESC)
        src_off = tmp_off = max_rows = 0
        for i, a in enumerate(arrs):
            assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3, (a.dtype, a.shape)
            h, w = a.shape[:2]
            p = self.plan(h, w)
            hk, vk = self._tab_off[(h, w)]
            desc[i] = (src_off, tmp_off, h, w, p.y_first, p.n_rows, hk, vk, p.hksize, p.vksize)
            src_off += (a.size + 15) // 16 * 16
            tmp_off += (p.n_rows * self.out_w * 3 + 15) // 16 * 16
            max_rows = max(max_rows, p.n_rows)
        host = torch.empty(src_off, dtype=torch.uint8).pin_memory()
        hv = host.numpy()
        for a, d in zip(arrs, desc):
            hv[d["src_off"]:d["src_off"] + a.size] = a.reshape(-1)
        if self._tables is None:
            self._tables = torch.from_numpy(np.concatenate(self._tab_host)).to(self.device)
        src = host.to(self.device, non_blocking=True)
        ddesc = torch.from_numpy(desc.view(np.uint8).reshape(-1).copy()).to(self.device, non_blocking=True)
        ws = torch.empty(max(tmp_off, 16), dtype=torch.uint8, device=self.device)
        if out is None:
            out = torch.empty(n, 3, self.out_h, self.out_w, dtype=torch.float32, device=self.device)
        call("zsg_resize_rgb8", ptr(src), ptr(ddesc), ptr(self._tables), n, self.out_h, self.out_w, max_rows,
             int(self.resample == "nearest"), ptr(ws), ptr(out), stream())
        for t in (src, ddesc, ws):
            t.record_stream(torch.cuda.current_stream(self.device))
        return out


def embed_gather(tokens, table, out=None):
    """tokens int32 [..] (device; -1 = padding -> zeros), table float32 [V, D] (device) -> float32 [.., D]."""
    assert tokens.dtype == torch.int32 and table.dtype == torch.float32 and table.is_contiguous()
    tokens = tokens.contiguous()
    if out is None:
        out = torch.empty(*tokens.shape, table.shape[1], dtype=torch.float32, device=table.device)
    call("zsg_embed_gather", ptr(tokens), ptr(table), ptr(out), tokens.numel(), table.shape[1], stream())
    return out


class GpuBatchStage:
    """Raw batch (dat_loader.raw_collater: decoded images of any size + token ids or vectors) -> the reference's batch dict
    on the device (dat_loader.py:136-144, 187-196): img B x 3 x H x W float32 in [0, 1], qvec B x T' x 300, qlens, annot,
    orig_annot, img_size, idxs, all float32."""

    def __init__(self, device, out_hw=(300, 300), resample="bicubic", embed_table=None):
        self.device = torch.device(device)
        self.images = GpuImageStage(device, out_hw, resample)
        self.table = embed_table.to(self.device).contiguous() if embed_table is not None else None

    def __call__(self, raw):
        out = {k: v.to(self.device, non_blocking=True).float() for k, v in raw.items()
               if k in ("idxs", "qlens", "annot", "orig_annot", "img_size", "qvec")}
        out["img"] = self.images(raw["img_raw"])
        max_qlen = int(raw["qlens"].max().item())
        if "tokens" in raw:
            assert self.table is not None, "token ids need embed_table"
            tok = raw["tokens"][:, :max_qlen].to(torch.int32).to(self.device, non_blocking=True)
            out["qvec"] = embed_gather(tok, self.table)
        else:
            out["qvec"] = out["qvec"][:, :max_qlen]
        out["qlens_cpu"] = raw["qlens"].float()
        return out
