// Anchor matching + focal / smooth-L1 loss + gradient, and the Acc@0.5 evaluator.
//
// Replaces loss.py:73-135 (≈25 ATen launches, a 1.22 GB torch.eye and 3 host syncs) and
// evaluator.py:74-99 by three launches each.  IoU follows anchors.py:90-116 operation by
// operation in float64 with explicit round-to-nearest intrinsics so that nvcc cannot contract
// a*b+c into an FMA: the positive mask and both argmaxes are bit-exact with the CPU reference.
//
// HBM roofline: 40 B per (sample, anchor): read att 4 + reg 16, write d_att 4 + d_reg 16
//  + 1 B mask; the 559 KB float64 anchor table stays in L2.
#include <math.h>
#include "common.cuh"

namespace zsg {

struct Box4 { double y1, x1, y2, x2; };

__device__ __forceinline__ Box4 load_anchor(const double* __restrict__ anchors, int a) {
  const double2* p = reinterpret_cast<const double2*>(anchors + 4 * (size_t)a);
  double2 u = __ldg(p), v = __ldg(p + 1);
  return {u.x, u.y, v.x, v.y};
}

// anchors.py:106-116 as IoU_values(annot, anchs): g* float32 GT, an float64 anchor.
__device__ __forceinline__ double iou_gt_anchor(float g0, float g1, float g2, float g3, const Box4& an) {
  double tl0 = fmax((double)g0, an.y1), tl1 = fmax((double)g1, an.x1);
  double br0 = fmin((double)g2, an.y2), br1 = fmin((double)g3, an.x2);
  double s0 = fmax(__dsub_rn(br0, tl0), 0.0), s1 = fmax(__dsub_rn(br1, tl1), 0.0);
  double inter = __dmul_rn(s0, s1);
  float gh = __fsub_rn(g2, g0), gw = __fsub_rn(g3, g1);
  double gt_area = (double)__fmul_rn(gh, gw);                       // float32 product, then promoted
  double an_area = __dmul_rn(__dsub_rn(an.y2, an.y1), __dsub_rn(an.x2, an.x1));
  double uni = __dsub_rn(__dadd_rn(gt_area, an_area), inter);
  return __ddiv_rn(inter, __dadd_rn(uni, 1e-8));
}

// evaluator.py:115 as IoU_values(best_boxes, annot): box float64 first, GT float32 second.
__device__ __forceinline__ double iou_box_gt(const Box4& bx, float g0, float g1, float g2, float g3) {
  double tl0 = fmax(bx.y1, (double)g0), tl1 = fmax(bx.x1, (double)g1);
  double br0 = fmin(bx.y2, (double)g2), br1 = fmin(bx.x2, (double)g3);
  double s0 = fmax(__dsub_rn(br0, tl0), 0.0), s1 = fmax(__dsub_rn(br1, tl1), 0.0);
  double inter = __dmul_rn(s0, s1);
  double b_area = __dmul_rn(__dsub_rn(bx.y2, bx.y1), __dsub_rn(bx.x2, bx.x1));
  float gh = __fsub_rn(g2, g0), gw = __fsub_rn(g3, g1);
  double gt_area = (double)__fmul_rn(gh, gw);
  double uni = __dsub_rn(__dadd_rn(b_area, gt_area), inter);
  return __ddiv_rn(inter, __dadd_rn(uni, 1e-8));
}

struct ArgMaxD { double v; int i; };
__device__ __forceinline__ ArgMaxD better(ArgMaxD a, ArgMaxD b) {   // larger value, then lower index
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMaxD block_argmax(ArgMaxD x, ArgMaxD* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMaxD y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sm[w] = x;
  __syncthreads();
  if (w == 0) {
    x = (l < nw) ? sm[l] : ArgMaxD{-1.0, 0x7fffffff};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ArgMaxD y;
      y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
      y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
      x = better(x, y);
    }
    if (l == 0) sm[0] = x;
  }
  __syncthreads();
  ArgMaxD r = sm[0];
  __syncthreads();
  return r;
}

struct LossWs {                      // device workspace header (doubles first for alignment)
  double cls_sum;
  unsigned long long npos_total;
  int nan_flag;
  int pad;
  // followed by double box_row[B]; int npos_row[B]
};

// ---- pass 1: one CTA per sample: IoU row, first-index argmax, positives ---------------------
__global__ void __launch_bounds__(1024) match_rows_kernel(const float* __restrict__ annot,
                                                          const double* __restrict__ anchors, int A,
                                                          double thr, int use_multi, uint8_t* __restrict__ pos,
                                                          int64_t* __restrict__ top1, LossWs* ws,
                                                          int* __restrict__ npos_row) {
  __shared__ ArgMaxD sm[32];
  __shared__ int cnt_sm[32];
  const int b = blockIdx.x;
  const float g0 = annot[4 * b], g1 = annot[4 * b + 1], g2 = annot[4 * b + 2], g3 = annot[4 * b + 3];
  ArgMaxD best{-1.0, 0x7fffffff};
  int cnt = 0;
  uint8_t* prow = pos + (size_t)b * A;
  for (int a = threadIdx.x; a < A; a += blockDim.x) {
    double v = iou_gt_anchor(g0, g1, g2, g3, load_anchor(anchors, a));
    best = better(best, ArgMaxD{v, a});
    int p = (use_multi && v > thr) ? 1 : 0;
    cnt += p;
    prow[a] = (uint8_t)p;
  }
  best = block_argmax(best, sm);
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) cnt_sm[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += cnt_sm[w];
    if (!(use_multi && best.v > thr)) total += 1;                   // the top-1 anchor is always positive
    prow[best.i] = 1;
    top1[b] = best.i;
    npos_row[b] = total;
    atomicAdd(&ws->npos_total, (unsigned long long)total);
  }
}

// ---- pass 2: per (sample, anchor) loss terms and gradients ----------------------------------
__global__ void __launch_bounds__(256) loss_grad_kernel(
    const float* __restrict__ att, int64_t att_stride, const float* __restrict__ reg, int64_t reg_stride,
    const float* __restrict__ annot, const double* __restrict__ anchors, const uint8_t* __restrict__ pos, int B, int A,
    float alpha, float gamma, double lamb_reg, float* __restrict__ d_att, int64_t d_att_stride,
    float* __restrict__ d_reg, int64_t d_reg_stride, LossWs* ws, double* __restrict__ box_row,
    const int* __restrict__ npos_row) {
  const int b = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  double cls_l = 0.0, box_l = 0.0;
  if (a < A) {
    const size_t e = (size_t)b * A + a;
    const float x = att[e * att_stride];
    const bool t = pos[e] != 0;
    // loss.py:103-125 in float32, like the reference
    const float p = 1.0f / (1.0f + expf(-x));
    float w = t ? (1.0f - p) : p;
    w = (gamma == 2.0f) ? w * w : powf(w, gamma);
    w *= t ? (1.0f - alpha) : alpha;                                  // alpha weights the negatives
    const float tf = t ? 1.0f : 0.0f;
    const float bce = fmaxf(x, 0.0f) - x * tf + log1pf(expf(-fabsf(x)));
    cls_l = (double)(w * bce);
    const float inv_np = 1.0f / (float)ws->npos_total;
    d_att[e * d_att_stride] = w * (p - tf) * inv_np;
    float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t) {                                                          // few per row: divergence is cheap
      const float g0 = annot[4 * b], g1 = annot[4 * b + 1], g2 = annot[4 * b + 2], g3 = annot[4 * b + 3];
      const Box4 an = load_anchor(anchors, a);
      // anchors.py:168-179: GT centre/size in float32, anchors in float64
      const float gc0 = __fdiv_rn(__fadd_rn(g0, g2), 2.0f), gc1 = __fdiv_rn(__fadd_rn(g1, g3), 2.0f);
      const float gh = __fsub_rn(g2, g0), gw = __fsub_rn(g3, g1);
      const double ac0 = __ddiv_rn(__dadd_rn(an.y1, an.y2), 2.0), ac1 = __ddiv_rn(__dadd_rn(an.x1, an.x2), 2.0);
      const double ah = __dadd_rn(__dsub_rn(an.y2, an.y1), 1e-8), aw = __dadd_rn(__dsub_rn(an.x2, an.x1), 1e-8);
      double tg[4];
      tg[0] = ((double)gc0 - ac0) / ah;
      tg[1] = ((double)gc1 - ac1) / aw;
      tg[2] = log((double)gh / ah);
      tg[3] = log((double)gw / aw);
      const float* r = reg + e * reg_stride;
      const double gscale = lamb_reg / ((double)B * (double)npos_row[b]);
      float dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double d = (double)r[k] - tg[k];
        const double ad = fabs(d);
        box_l += (ad < 1.0) ? 0.5 * d * d : ad - 0.5;                // SmoothL1, beta = 1 (loss.py:41,91)
        dv[k] = (float)(gscale * fmin(fmax(d, -1.0), 1.0));
      }
      dr = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
    float* dp = d_reg + e * d_reg_stride;
    if (d_reg_stride == 4) {
      *reinterpret_cast<float4*>(dp) = dr;
    } else {
      dp[0] = dr.x; dp[1] = dr.y; dp[2] = dr.z; dp[3] = dr.w;
    }
  }
  // block reduction -> one atomic per block and quantity
  __shared__ double red[2][8];
  cls_l = warp_sum(cls_l);
  box_l = warp_sum(box_l);
  const int w_ = threadIdx.x >> 5, l_ = threadIdx.x & 31;
  if (l_ == 0) { red[0][w_] = cls_l; red[1][w_] = box_l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0.0, bx = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { c += red[0][i]; bx += red[1][i]; }
    atomicAdd(&ws->cls_sum, c);
    if (bx != 0.0) atomicAdd(&box_row[b], bx);
  }
}

__global__ void loss_finalize_kernel(LossWs* ws, const double* __restrict__ box_row, const int* __restrict__ npos_row,
                                     int B, double lamb_reg, double* __restrict__ losses) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double box = 0.0;
  for (int b = 0; b < B; ++b) box += box_row[b] / (double)(float)npos_row[b];
  box /= (double)B;
  float cls = (float)ws->cls_sum / (float)ws->npos_total;           // f32 / count, like loss.py:125
  int bad = (box != box) || (cls != cls);
  if (bad) { box = 0.01; cls = 1.0f; }                               // loss.py:128-133
  ws->nan_flag = bad;
  losses[0] = lamb_reg * box + (double)cls;
  losses[1] = (double)cls;
  losses[2] = box;
}

// loss.py:128-133: a NaN step carries no gradient.  Returns immediately in the normal case.
__global__ void zero_if_nan_kernel(const LossWs* ws, float* d_att, int64_t sa, float* d_reg, int64_t sr, size_t n) {
  if (!ws->nan_flag) return;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    d_att[e * sa] = 0.f;
    for (int k = 0; k < 4; ++k) d_reg[e * sr + k] = 0.f;
  }
}

// ---- evaluator ------------------------------------------------------------------------------
struct ArgMaxF { float v; int i; };
__device__ __forceinline__ ArgMaxF betterf(ArgMaxF a, ArgMaxF b) {
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

__device__ __forceinline__ Box4 decode_box(const Box4& an, const float* r) {   // anchors.py:182-197
  const double ac0 = __ddiv_rn(__dadd_rn(an.y1, an.y2), 2.0), ac1 = __ddiv_rn(__dadd_rn(an.x1, an.x2), 2.0);
  const double ah = __dsub_rn(an.y2, an.y1), aw = __dsub_rn(an.x2, an.x1);
  const double c0 = __dadd_rn(__dmul_rn(ah, (double)r[0]), ac0), c1 = __dadd_rn(__dmul_rn(aw, (double)r[1]), ac1);
  const double h = __dmul_rn((double)expf(r[2]), ah), w = __dmul_rn((double)expf(r[3]), aw);   // exp in float32
  const double hh = __ddiv_rn(h, 2.0), hw = __ddiv_rn(w, 2.0);
  return {__dsub_rn(c0, hh), __dsub_rn(c1, hw), __dadd_rn(c0, hh), __dadd_rn(c1, hw)};
}

__global__ void __launch_bounds__(1024) eval_rows_kernel(const float* __restrict__ att, int64_t att_stride,
                                                         const float* __restrict__ reg, int64_t reg_stride,
                                                         const float* __restrict__ annot,
                                                         const double* __restrict__ anchors,
                                                         const float* __restrict__ img_size, int A, double thr,
                                                         int64_t* __restrict__ best_ids, float* __restrict__ scores,
                                                         double* __restrict__ pred_boxes, float* __restrict__ flags) {
  __shared__ ArgMaxD smd[32];
  __shared__ ArgMaxF smf[32];
  const int b = blockIdx.x;
  const float g0 = annot[4 * b], g1 = annot[4 * b + 1], g2 = annot[4 * b + 2], g3 = annot[4 * b + 3];
  ArgMaxD bi{-1.0, 0x7fffffff};
  ArgMaxF bs{-1.0f, 0x7fffffff};
  for (int a = threadIdx.x; a < A; a += blockDim.x) {
    bi = better(bi, ArgMaxD{iou_gt_anchor(g0, g1, g2, g3, load_anchor(anchors, a)), a});
    const float x = att[((size_t)b * A + a) * att_stride];
    bs = betterf(bs, ArgMaxF{1.0f / (1.0f + expf(-x)), a});          // evaluator.py:74-75
  }
  bi = block_argmax(bi, smd);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMaxF y;
    y.v = __shfl_xor_sync(0xffffffffu, bs.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, bs.i, o);
    bs = betterf(bs, y);
  }
  if ((threadIdx.x & 31) == 0) smf[threadIdx.x >> 5] = bs;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) bs = betterf(bs, smf[w]);
    const Box4 bx = decode_box(load_anchor(anchors, bs.i), reg + ((size_t)b * A + bs.i) * reg_stride);
    const Box4 mx = decode_box(load_anchor(anchors, bi.i), reg + ((size_t)b * A + bi.i) * reg_stride);
    flags[b] = iou_box_gt(bx, g0, g1, g2, g3) >= thr ? 1.f : 0.f;
    flags[gridDim.x + b] = iou_box_gt(mx, g0, g1, g2, g3) >= thr ? 1.f : 0.f;
    best_ids[b] = bs.i;
    scores[b] = bs.v;
    // evaluator.py:96-97: ((box+1)/2) * (h,w), then y1x1y2x2 -> x1y1x2y2
    const double h = (double)img_size[2 * b], w = (double)img_size[2 * b + 1];
    const double py1 = __dmul_rn(h, __ddiv_rn(__dadd_rn(bx.y1, 1.0), 2.0)), px1 = __dmul_rn(w, __ddiv_rn(__dadd_rn(bx.x1, 1.0), 2.0));
    const double py2 = __dmul_rn(h, __ddiv_rn(__dadd_rn(bx.y2, 1.0), 2.0)), px2 = __dmul_rn(w, __ddiv_rn(__dadd_rn(bx.x2, 1.0), 2.0));
    double* o = pred_boxes + 4 * b;
    o[0] = px1; o[1] = py1; o[2] = px2; o[3] = py2;
  }
}

__global__ void eval_finalize_kernel(const float* __restrict__ flags, int B, float* __restrict__ metrics) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float a = 0.f, m = 0.f;
  for (int b = 0; b < B; ++b) { a += flags[b]; m += flags[B + b]; }
  metrics[0] = a / (float)B;
  metrics[1] = m / (float)B;
}

}  // namespace zsg

using namespace zsg;

extern "C" size_t zsg_match_loss_workspace_bytes(int b) {
  return sizeof(LossWs) + (size_t)b * (sizeof(double) + sizeof(int)) + 16;
}

extern "C" int zsg_match_loss(const float* att, int64_t att_stride, const float* reg, int64_t reg_stride,
                              const float* annot, const double* anchors, int b, int a, double match_thr, float alpha,
                              float gamma, double lamb_reg, int use_multi, double* losses, float* d_att,
                              int64_t d_att_stride, float* d_reg, int64_t d_reg_stride, int64_t* top1, uint8_t* pos,
                              void* workspace, size_t ws_bytes, zsg_stream_t stream) {
  ZSG_REQUIRE(att && reg && annot && anchors && losses && d_att && d_reg && top1 && pos && workspace,
              "zsg_match_loss: null pointer");
  ZSG_REQUIRE(b > 0 && a > 0, "zsg_match_loss: empty batch (b=%d a=%d)", b, a);
  ZSG_REQUIRE(ws_bytes >= zsg_match_loss_workspace_bytes(b), "zsg_match_loss: workspace too small");
  ZSG_REQUIRE(((uintptr_t)anchors & 15) == 0, "zsg_match_loss: anchors must be 16-byte aligned");
  ZSG_REQUIRE(d_reg_stride != 4 || ((uintptr_t)d_reg & 15) == 0, "zsg_match_loss: d_reg must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  LossWs* ws = reinterpret_cast<LossWs*>(workspace);
  double* box_row = reinterpret_cast<double*>(ws + 1);
  int* npos_row = reinterpret_cast<int*>(box_row + b);
  cudaMemsetAsync(workspace, 0, zsg_match_loss_workspace_bytes(b), st);
  match_rows_kernel<<<b, 1024, 0, st>>>(annot, anchors, a, match_thr, use_multi, pos, top1, ws, npos_row);
  dim3 grid((a + 255) / 256, b);
  loss_grad_kernel<<<grid, 256, 0, st>>>(att, att_stride, reg, reg_stride, annot, anchors, pos, b, a, alpha, gamma,
                                         lamb_reg, d_att, d_att_stride, d_reg, d_reg_stride, ws, box_row, npos_row);
  loss_finalize_kernel<<<1, 32, 0, st>>>(ws, box_row, npos_row, b, lamb_reg, losses);
  zero_if_nan_kernel<<<num_sms(), 256, 0, st>>>(ws, d_att, d_att_stride, d_reg, d_reg_stride, (size_t)b * a);
  return check_launch("zsg_match_loss");
}

extern "C" int zsg_eval(const float* att, int64_t att_stride, const float* reg, int64_t reg_stride,
                        const float* annot, const double* anchors, const float* img_size, int b, int a,
                        double iou_thr, int64_t* best_ids, float* pred_scores, double* pred_boxes, float* metrics,
                        zsg_stream_t stream) {
  ZSG_REQUIRE(att && reg && annot && anchors && img_size && best_ids && pred_scores && pred_boxes && metrics,
              "zsg_eval: null pointer");
  ZSG_REQUIRE(b > 0 && a > 0, "zsg_eval: empty batch");
  cudaStream_t st = as_stream(stream);
  // flags live behind the two metrics: metrics must hold 2 + 2*b floats
  float* flags = metrics + 2;
  eval_rows_kernel<<<b, 1024, 0, st>>>(att, att_stride, reg, reg_stride, annot, anchors, img_size, a, iou_thr,
                                       best_ids, pred_scores, pred_boxes, flags);
  eval_finalize_kernel<<<1, 32, 0, st>>>(flags, b, metrics);
  return check_launch("zsg_eval");
}
