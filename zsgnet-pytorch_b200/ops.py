"""Tensor-level wrappers over the C ABI (include/zsg_b200.h).  torch supplies device memory and
streams only; every computation below is a kernel of libzsg_b200.so."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ConvParams, WgradParams, call, ptr, stream

IMPL_TC, IMPL_SIMT = 0, 1
PROFILER = None          # set by bench.py: per-launch CUDA-event timing of the implicit-GEMM kernels


class LaunchProfiler:
    """CUDA events around every implicit-GEMM launch on the launching stream (bench.py roofline)."""

    def __init__(self):
        self.records = []                # (kernel, flops, start_event, end_event)

    def timed(self, op, entry):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        call(entry, op.ref, stream())
        b.record()
        self.records.append((op.kernel, op.flops, a, b))

    def summary(self):
        out = {}
        for k, f, a, b in self.records:
            d = out.setdefault(k, dict(launches=0, flops=0.0, ms=0.0))
            d["launches"] += 1
            d["flops"] += f
            d["ms"] += a.elapsed_time(b)
        return out


def _is_bf16(t):
    return t is not None and t.dtype == torch.bfloat16


class ConvOp:
    """One implicit-GEMM launch description with its parameter block kept alive (pointers are
    fixed because the engine's buffers are static)."""

    def __init__(self, x, w, y, rows, m, cin, cout, r, s, in_div=1, in_scale=None, in_shift=None, in_relu=False,
                 bias=None, out_relu=False, out_mask=None, residual=None, accumulate=False, impl=IMPL_TC, w_lo=None,
                 x_lo=None, dil=1, stats=None, x_plain=False, y_pitch=0, row_add=None, row_add_idx=None, y_img=None, bnb=None):
        # bnb = (x, scale, shift, partials): BatchNorm backward sums as a by-product of this data gradient (zsg_conv_params.bnb_*)
        bnb = bnb if bnb is not None else (None, None, None, None)
        self.keep = (x, w, w_lo, y, rows, in_scale, in_shift, bias, out_mask, residual, x_lo, stats, row_add, row_add_idx, y_img, bnb)
        yb = None
        if _is_bf16(y):                                      # bf16 storage: the output tensor itself is bfloat16
            y, yb = None, y
        rb = None
        if _is_bf16(residual):
            residual, rb = None, residual
        # operand images: float32 x_lo / w_lo = TF32 remainders (3xTF32 path); bfloat16 x_lo / w_lo = the bf16 copies of the
        # operands (bf16 path of BASELINE configs 3-5; goes into the x_bf16 / w_bf16 fields of the parameter block)
        self.bf16 = _is_bf16(x_lo)
        assert self.bf16 == _is_bf16(w_lo), "conv operand images must both be bf16 or both fp32"
        xb, wb = (x_lo, w_lo) if self.bf16 else (None, None)
        if self.bf16:
            x_lo = w_lo = None
        self.p = ConvParams(ptr(x), ptr(w), ptr(w_lo), ptr(y), ptr(rows), m, cin, cout, r, s, in_div, ptr(in_scale),
                            ptr(in_shift), int(in_relu), ptr(bias), int(out_relu), ptr(out_mask), ptr(residual), int(accumulate),
                            impl, ptr(x_lo), dil, ptr(stats), ptr(xb), ptr(wb), int(bool(x_plain)), ptr(yb), ptr(rb), int(y_pitch),
                            ptr(row_add), ptr(row_add_idx),
                            # y_img: the operand image of the output, written by the epilogue (fp32 = TF32 remainders, bf16 = copy)
                            ptr(None if _is_bf16(y_img) else y_img), ptr(y_img if _is_bf16(y_img) else None),
                            ptr(bnb[0]), ptr(bnb[1]), ptr(bnb[2]), ptr(bnb[3]))
        if y_pitch and rows.is_cuda:                         # the claim is checked once, when the launch is described
            out = rows.view(torch.int32).view(-1, 4)[:m, 3].to(torch.int64)
            assert bool((out == torch.arange(m, device=rows.device) * y_pitch).all()), "y_pitch does not describe this row table"
        self.ref = C.byref(self.p)
        self.flops = 2.0 * m * cout * r * s * cin / (in_div * in_div)     # algorithmic (valid taps only)
        self.kernel = ("conv_bf16_kernel" if self.bf16 else "conv_tc_async_kernel" if x_lo is not None else
                       "conv_tc_kernel" if w_lo is not None else "conv_tc_generic_kernel")

    def __call__(self):
        if PROFILER is not None:
            return PROFILER.timed(self, "zsg_conv_fwd")
        call("zsg_conv_fwd", self.ref, stream())


class WgradOp:
    def __init__(self, x, dy, dw, rows, m, cin, cout, r, s, in_scale=None, in_shift=None, in_relu=False, split_k=0,
                 impl=IMPL_TC, x_lo=None, dy_lo=None, dy_pitch=0, dil=1):
        self.keep = (x, dy, dw, rows, in_scale, in_shift, x_lo, dy_lo)
        self.dw = dw
        self.bf16 = _is_bf16(x_lo)                           # bf16 images (see ConvOp)
        assert self.bf16 == _is_bf16(dy_lo), "wgrad operand images must both be bf16 or both fp32"
        xb, dyb = (x_lo, dy_lo) if self.bf16 else (None, None)
        if self.bf16:
            x_lo = dy_lo = None
        self.p = WgradParams(ptr(x), ptr(dy), ptr(dw), ptr(rows), m, cin, cout, r, s, ptr(in_scale), ptr(in_shift),
                             int(in_relu), split_k, impl, ptr(x_lo), ptr(dy_lo), dy_pitch, dil, ptr(xb), ptr(dyb))
        self.ref = C.byref(self.p)
        self.flops = 2.0 * m * cout * r * s * cin
        self.kernel = "wgrad_bf16_kernel" if self.bf16 else "wgrad_tc_async_kernel" if x_lo is not None else "wgrad_tc_kernel"

    def __call__(self):
        if PROFILER is not None:
            return PROFILER.timed(self, "zsg_conv_wgrad")
        call("zsg_conv_wgrad", self.ref, stream())


def copy_cols(src, src_ld, dst, dst_ld, rows, c):
    call("zsg_copy_cols", ptr(src), src_ld, ptr(dst), dst_ld, rows, c, stream())


def head0_lang_grid_terms(v, wg, gridpatch, lang_cls, grid_term, b, total_cells, n):
    call("zsg_head0_lang_grid_terms", ptr(v), ptr(wg), ptr(gridpatch), ptr(lang_cls), ptr(grid_term), b, total_cells, n, stream())


def head0_backward_sums(dh, cell_base, cell_stride, cell_cls, gridpatch, b, total_cells, n, scratch, tap_sums, dwg, dwg_ld):
    call("zsg_head0_backward_sums", ptr(dh), ptr(cell_base), ptr(cell_stride), ptr(cell_cls), ptr(gridpatch), b, total_cells, n,
         ptr(scratch), scratch.numel(), ptr(tap_sums), ptr(dwg), dwg_ld, stream())


def weight_transpose_flip(w, wt, cout, r, s, cin):
    call("zsg_weight_transpose_flip", ptr(w), ptr(wt), cout, r, s, cin, stream())


WTF_DESC = np.dtype([("src", "<i8"), ("dst", "<i8"), ("begin", "<i8"), ("cout", "<i4"), ("r", "<i4"), ("s", "<i4"), ("cin", "<i4")])


def wtf_table(entries, device):
    """Device table for weight_transpose_flip_batched from [(src_off, dst_off, cout, r, s, cin)] (element offsets)."""
    arr = np.zeros(len(entries), dtype=WTF_DESC)
    begin = 0
    for i, (so, do, cout, r, s, cin) in enumerate(entries):
        arr[i] = (so, do, begin, cout, r, s, cin)
        begin += cout * r * s * cin
    assert arr.itemsize == 40
    return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(device), begin


def weight_transpose_flip_batched(src_base, dst_base, table, n, total, tiled=False):
    """tiled: every cin / cout of the table is a multiple of 32 (wtf_table_is_tiled)."""
    call("zsg_weight_transpose_flip_batched32" if tiled else "zsg_weight_transpose_flip_batched", ptr(src_base), ptr(dst_base),
         ptr(table), n, total, stream())


def wtf_table_is_tiled(entries):
    return all(cout % 32 == 0 and cin % 32 == 0 for _, _, cout, _, _, cin in entries)


def split_act(x, lo, rows, c, scale=None, shift=None, relu=False, z=None):
    """Operand image of relu?(x * scale + shift).  float32 `lo`: the TF32 remainder (and z = the activated tensor when a
    prologue is given); bfloat16 `lo`: the bf16 copy (z is not needed by the bf16 GEMMs and is not written)."""
    if _is_bf16(lo):
        call("zsg_cast_bf16", ptr(x), ptr(scale), ptr(shift), int(relu), ptr(lo), rows, c, stream())
        return
    call("zsg_split_act", ptr(x), ptr(scale), ptr(shift), int(relu), ptr(z), ptr(lo), rows, c, stream())


def cast_bf16(x, out, n):
    """out[i] = bf16(x[i]) over a flat array (weights); n % 4 == 0."""
    call("zsg_cast_bf16", ptr(x), None, None, 0, ptr(out), n // 4, 4, stream())


def split_tf32(w, hi, lo, n):
    call("zsg_split_tf32", ptr(w), ptr(hi), ptr(lo), n, stream())


def pad_channels(src, dst, n, csrc, cdst):
    call("zsg_pad_channels", ptr(src), ptr(dst), n, csrc, cdst, stream())


def nchw_to_nhwc4(img, out):
    """out float32 [B,H,W,4] (fp32 path) or bfloat16 [B,H,W,8] (bf16 path: padded image = operand image)."""
    b, _, h, w = img.shape
    call("zsg_nchw_to_nhwc8_bf16" if _is_bf16(out) else "zsg_nchw_to_nhwc4", ptr(img), ptr(out), b, h, w, stream())


def colsum(x, out, rows, c, accumulate=False, ld=None):
    call("zsg_colsum", ptr(x), ptr(out), rows, c, c if ld is None else ld, int(accumulate), stream())


def gather_rows(src, rows, dst, m, csrc, cdst):
    call("zsg_gather_rows", ptr(src), ptr(rows), ptr(dst), m, csrc, cdst, stream())


def bn_stats(x, sums, rows, c):
    call("zsg_bn_stats", ptr(x), ptr(sums), rows, c, stream())


def bn_stats_partials(partials, parts, c, sums):
    call("zsg_bn_stats_partials", ptr(partials), parts, c, ptr(sums), stream())


def bn_bwd_center_sums(sums, mean, invstd, c):
    call("zsg_bn_bwd_center_sums", ptr(sums), ptr(mean), ptr(invstd), c, stream())


def bn_finalize_partials(partials, parts, rows, c, gamma, beta, eps, momentum, rm, rv, mean, invstd, scale, shift, sums,
                         tickets):
    call("zsg_bn_finalize_partials", ptr(partials), parts, rows, c, ptr(gamma), ptr(beta), eps, momentum, ptr(rm), ptr(rv),
         ptr(mean), ptr(invstd), ptr(scale), ptr(shift), ptr(sums), ptr(tickets), stream())


def bn_finalize(sums, rows, c, gamma, beta, eps, momentum, rm, rv, mean, invstd, scale, shift):
    call("zsg_bn_finalize", ptr(sums), rows, c, ptr(gamma), ptr(beta), eps, momentum, ptr(rm), ptr(rv), ptr(mean),
         ptr(invstd), ptr(scale), ptr(shift), stream())


def bn_eval_affine(rm, rv, gamma, beta, eps, c, scale, shift):
    call("zsg_bn_eval_affine", ptr(rm), ptr(rv), ptr(gamma), ptr(beta), eps, c, ptr(scale), ptr(shift), stream())


def bn_apply(x, scale, shift, y, rows, c, relu, r=None, rscale=None, rshift=None, y_lo=None):
    if _is_bf16(x):                                          # bf16 storage: bfloat16 in, bfloat16 out (y is the only copy)
        assert _is_bf16(y) and y_lo is None and (r is None or _is_bf16(r))
        call("zsg_bn_apply_b16", ptr(x), ptr(scale), ptr(shift), ptr(r), ptr(rscale), ptr(rshift), int(relu), ptr(y), rows, c,
             stream())
        return
    call("zsg_bn_apply_bf16" if _is_bf16(y_lo) else "zsg_bn_apply", ptr(x), ptr(scale), ptr(shift), ptr(r), ptr(rscale),
         ptr(rshift), int(relu), ptr(y), ptr(y_lo), rows, c, stream())


def act_b16(x, out, rows, c, scale=None, shift=None, relu=False):
    """bf16 storage: out = bf16(relu?(x * scale + shift)) over a bfloat16 tensor."""
    assert _is_bf16(x) and _is_bf16(out)
    call("zsg_act_b16", ptr(x), ptr(scale), ptr(shift), int(relu), ptr(out), rows, c, stream())


def bn_bwd_reduce(dy, x, mean, invstd, sums, rows, c, mask_mode=0, scale=None, shift=None, act_out=None, dz_out=None):
    if _is_bf16(x):                                          # bf16 storage
        assert act_out is None or _is_bf16(act_out)
        call("zsg_bn_bwd_reduce_b16", ptr(dy), int(_is_bf16(dy)), ptr(x), ptr(mean), ptr(invstd), ptr(scale), ptr(shift),
             ptr(act_out), mask_mode, ptr(dz_out), int(_is_bf16(dz_out)), ptr(sums), rows, c, stream())
        return
    _bn_bwd_reduce_f32(dy, x, mean, invstd, sums, rows, c, mask_mode, scale, shift, act_out, dz_out)


def _bn_bwd_reduce_f32(dy, x, mean, invstd, sums, rows, c, mask_mode=0, scale=None, shift=None, act_out=None, dz_out=None):
    call("zsg_bn_bwd_reduce", ptr(dy), ptr(x), ptr(mean), ptr(invstd), ptr(scale), ptr(shift), ptr(act_out), mask_mode,
         ptr(dz_out), ptr(sums), rows, c, stream())


def bn_bwd_apply(dy, x, mean, invstd, gamma, sums, dx, dgamma, dbeta, rows, c, mask_mode=0, scale=None, shift=None,
                 act_out=None, dx_lo=None):
    if _is_bf16(x):                                          # bf16 storage: only the bfloat16 dx is written
        assert dx is None and _is_bf16(dx_lo) and (act_out is None or _is_bf16(act_out))
        call("zsg_bn_bwd_apply_b16", ptr(dy), int(_is_bf16(dy)), ptr(x), ptr(mean), ptr(invstd), ptr(gamma), ptr(scale),
             ptr(shift), ptr(act_out), mask_mode, ptr(sums), ptr(dx_lo), ptr(dgamma), ptr(dbeta), rows, c, stream())
        return
    call("zsg_bn_bwd_apply_bf16" if _is_bf16(dx_lo) else "zsg_bn_bwd_apply", ptr(dy), ptr(x), ptr(mean), ptr(invstd),
         ptr(gamma), ptr(scale), ptr(shift), ptr(act_out), mask_mode, ptr(sums), ptr(dx), ptr(dx_lo), ptr(dgamma), ptr(dbeta),
         rows, c, stream())


def maxpool_bn_relu_fwd(x, scale, shift, y, argmax, b, h, w, c, ho, wo):
    if _is_bf16(x):
        assert _is_bf16(y)
        call("zsg_maxpool_bn_relu_fwd_b16", ptr(x), ptr(scale), ptr(shift), ptr(y), ptr(argmax), b, h, w, c, ho, wo, stream())
        return
    call("zsg_maxpool_bn_relu_fwd", ptr(x), ptr(scale), ptr(shift), ptr(y), ptr(argmax), b, h, w, c, ho, wo, stream())


def maxpool_bn_relu_bwd(argmax, dy, da, b, h, w, c, ho, wo):
    call("zsg_maxpool_bn_relu_bwd", ptr(argmax), ptr(dy), ptr(da), b, h, w, c, ho, wo, stream())


def upsample_add(dst, src, iy, ix, b, ho, wo, hi, wi, c):
    call("zsg_upsample_add", ptr(dst), ptr(src), ptr(iy), ptr(ix), b, ho, wo, hi, wi, c, stream())


def upsample_add_bwd(ddst, dsrc, iy, ix, b, ho, wo, hi, wi, c):
    call("zsg_upsample_add_bwd", ptr(ddst), ptr(dsrc), ptr(iy), ptr(ix), b, ho, wo, hi, wi, c, stream())


def avgpool_fwd(x, y, b, hw, c):
    call("zsg_avgpool_fwd", ptr(x), ptr(y), b, hw, c, stream())


def avgpool_bwd(dy, dx, b, hw, c):
    call("zsg_avgpool_bwd", ptr(dy), ptr(dx), b, hw, c, stream())


def maxpool_fwd(x, y, argmax, b, h, w, c, k, stride, pad, ho, wo):
    call("zsg_maxpool_fwd", ptr(x), ptr(y), ptr(argmax), b, h, w, c, k, stride, pad, ho, wo, stream())


def maxpool_bwd(argmax, dy, dx, b, h, w, c, k, stride, pad, ho, wo, mask=None):
    call("zsg_maxpool_bwd", ptr(argmax), ptr(dy), ptr(mask), ptr(dx), b, h, w, c, k, stride, pad, ho, wo, stream())


def l2norm_fwd(x, y, norm, rows, c):
    call("zsg_l2norm_fwd", ptr(x), ptr(y), ptr(norm), rows, c, stream())


def l2norm_bwd(dy, x, norm, dx, rows, c, accumulate=False, mask_relu=False):
    call("zsg_l2norm_bwd", ptr(dy), ptr(x), ptr(norm), ptr(dx), rows, c, int(accumulate), int(mask_relu), stream())


def relu_bwd(dy, x, dx, n, accumulate=False):
    call("zsg_relu_bwd", ptr(dy), ptr(x), ptr(dx), n, int(accumulate), stream())


def axpy(x, y, a, n):
    call("zsg_axpy", ptr(x), ptr(y), a, n, stream())


def scale_dev(x, scale):
    """x: [..., width] view whose rows are `stride` elements apart (last dim contiguous); scale: device double."""
    width = x.shape[-1]
    stride = x.stride(-2) if x.dim() > 1 else width
    rows = x.numel() // width
    scale = scale.reshape(1).contiguous()
    call("zsg_scale_dev", ptr(x), rows, width, stride, ptr(scale), stream())


def _lvl(cells):
    return (C.c_int32 * len(cells))(*cells)


def fuse_lang_grid(feat, lang, grid_yx, fused, b, total_cells, cells, cfeat, clang, cpad):
    call("zsg_fuse_lang_grid", ptr(feat), ptr(lang), ptr(grid_yx), ptr(fused), b, total_cells, _lvl(cells), len(cells),
         cfeat, clang, cpad, stream())


def unfuse_lang_grid(dfused, dfeat, dlang, b, total_cells, cells, cfeat, clang, cpad):
    call("zsg_unfuse_lang_grid", ptr(dfused), ptr(dfeat), ptr(dlang), b, total_cells, _lvl(cells), len(cells), cfeat,
         clang, cpad, stream())


def lstm_fwd_dir(gx, whh_t, b_ih, b_hh, h0, c0, lens, b, t, gates, cs, hprev, lang):
    call("zsg_lstm_fwd_dir", ptr(gx), ptr(whh_t), ptr(b_ih), ptr(b_hh), ptr(h0), ptr(c0), ptr(lens), b, t, ptr(gates),
         ptr(cs), ptr(hprev), ptr(lang), stream())


def lstm_rev_step(qvec, wih, whh, b_ih, b_hh, h0, c0, lens, b, t, e, xlast, gates, lang):
    call("zsg_lstm_rev_step", ptr(qvec), ptr(wih), ptr(whh), ptr(b_ih), ptr(b_hh), ptr(h0), ptr(c0), ptr(lens), b, t, e,
         ptr(xlast), ptr(gates), ptr(lang), stream())


def lstm_bwd_dir(dlang, whh, gates, cs, c0, lens, b, t, dgates):
    call("zsg_lstm_bwd_dir", ptr(dlang), ptr(whh), ptr(gates), ptr(cs), ptr(c0), ptr(lens), b, t, ptr(dgates), stream())


def lstm_rev_step_bwd(dlang, gates, c0, b, dgates):
    call("zsg_lstm_rev_step_bwd", ptr(dlang), ptr(gates), ptr(c0), b, ptr(dgates), stream())


def match(annot, anchors, b, a, thr, use_multi, top1, pos, ws):
    call("zsg_match", ptr(annot), ptr(anchors), b, a, thr, int(use_multi), ptr(top1), ptr(pos), ptr(ws), ws.numel() * 8, stream())


def loss_grad(att, att_stride, reg, reg_stride, annot, anchors, pos, b, a, alpha, gamma, lamb, losses, d_att, d_att_stride,
              d_reg, d_reg_stride, ws):
    call("zsg_loss_grad", ptr(att), att_stride, ptr(reg), reg_stride, ptr(annot), ptr(anchors), ptr(pos), b, a, alpha, gamma,
         lamb, ptr(losses), ptr(d_att), d_att_stride, ptr(d_reg), d_reg_stride, ptr(ws), ws.numel() * 8, stream())


def match_loss_workspace(b, device):
    """Scratch that carries the positives' counts from the match to the loss pass (include/zsg_b200.h)."""
    n = _lib.load().zsg_match_loss_workspace_bytes(b)
    return torch.zeros((n + 7) // 8, dtype=torch.float64, device=device)


def eval_workspace(b, device):
    n = _lib.load().zsg_eval_workspace_bytes(b)
    return torch.zeros((n + 7) // 8, dtype=torch.float64, device=device)


def match_loss(att, att_stride, reg, reg_stride, annot, anchors, b, a, thr, alpha, gamma, lamb, use_multi, losses,
               d_att, d_att_stride, d_reg, d_reg_stride, top1, pos, ws):
    call("zsg_match_loss", ptr(att), att_stride, ptr(reg), reg_stride, ptr(annot), ptr(anchors), b, a, thr, alpha,
         gamma, lamb, int(use_multi), ptr(losses), ptr(d_att), d_att_stride, ptr(d_reg), d_reg_stride, ptr(top1),
         ptr(pos), ptr(ws), ws.numel() * 8, stream())


_EVAL_WS = {}


def evaluate(att, att_stride, reg, reg_stride, annot, anchors, img_size, b, a, thr, best_ids, scores, boxes, metrics,
             ws=None):
    if ws is None:                                       # one cached scratch per (device, batch): calls on a stream serialise
        key = (att.device, b)
        if key not in _EVAL_WS:
            _EVAL_WS[key] = eval_workspace(b, att.device)
        ws = _EVAL_WS[key]
    call("zsg_eval", ptr(att), att_stride, ptr(reg), reg_stride, ptr(annot), ptr(anchors), ptr(img_size), b, a, thr,
         ptr(best_ids), ptr(scores), ptr(boxes), ptr(metrics), ptr(ws), ws.numel() * 8, stream())


def adam(p, g, m, v, n, lr, b1, b2, eps, step, grad_scale=1.0):
    call("zsg_adam", ptr(p), ptr(g), ptr(m), ptr(v), n, lr, b1, b2, eps, step, grad_scale, stream())
