// Anchor matching + focal / smooth-L1 loss + gradient, and the Acc@0.5 evaluator.
//
// Replaces loss.py:73-135 (≈25 ATen launches, a 1.22 GB torch.eye and 3 host syncs) by two launches and
// evaluator.py:74-99 by one (round 1: memset + four, and two; the match pass ran one CTA per sample).  IoU follows anchors.py:90-116 operation by
// operation in float64 with explicit round-to-nearest intrinsics so that nvcc cannot contract
// a*b+c into an FMA: the positive mask and both argmaxes are bit-exact with the CPU reference.
//
// HBM roofline: 40 B per (sample, anchor): read att 4 + reg 16, write d_att 4 + d_reg 16
//  + 1 B mask; the 559 KB float64 anchor table stays in L2.
#include <math.h>
#include <stdlib.h>
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

// torch.max semantics for the row argmax (anchors.py / loss.py:77, evaluator.py:74): the largest value wins, NaN counts as
// the largest of all, ties (and several NaNs) go to the LOWEST index.  A diverged network (all-NaN scores) or a NaN
// annotation therefore still yields a valid index, like the reference, instead of an out-of-range one.
struct ArgMaxD { double v; int i; };
__device__ __forceinline__ ArgMaxD better(ArgMaxD a, ArgMaxD b) {
  const bool an = a.v != a.v, bn = b.v != b.v;
  const bool gt = bn ? !an : (!an && b.v > a.v);
  const bool eq = (an && bn) || b.v == a.v;
  return (gt || (eq && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMaxD warp_argmax(ArgMaxD x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMaxD y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  return x;
}
constexpr double ARG_NONE = -1.0;       // below every IoU / score; paired with index INT_MAX, replaced by the first candidate
__device__ __forceinline__ ArgMaxD block_argmax(ArgMaxD x, ArgMaxD* sm) {
  x = warp_argmax(x);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sm[w] = x;
  __syncthreads();
  if (w == 0) {
    x = warp_argmax((l < nw) ? sm[l] : ArgMaxD{ARG_NONE, 0x7fffffff});
    if (l == 0) sm[0] = x;
  }
  __syncthreads();
  ArgMaxD r = sm[0];
  __syncthreads();
  return r;
}

// Device workspace (no state between calls: zsg_match clears the counters it uses).  Partial sums of the loss are written
// to per-tile slots and added up in a fixed order by a one-block finalize kernel: no floating-point atomics, so the three
// loss scalars are bit-reproducible from run to run.
constexpr int MATCH_MAX_CHUNKS = 32;
constexpr int LOSS_MAX_TILES = 128;          // tiles of 256 anchors per row: A <= 32768
struct LossWs {
  unsigned long long npos_total;
  unsigned int rows_done;
  int nan_flag;
  // followed by the arrays of ws_*_off() below
};
struct MatchPart { double v; int i; int cnt; };

__host__ __device__ inline size_t ws_al(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline size_t ws_rowcls_off() { return ws_al(sizeof(LossWs)); }                              // double row_cls[B]
__host__ __device__ inline size_t ws_box_off(int B) { return ws_rowcls_off() + (size_t)B * sizeof(double); }     // double row_box[B]
__host__ __device__ inline size_t ws_npos_off(int B) { return ws_box_off(B) + (size_t)B * sizeof(double); }      // int npos_row[B]
__host__ __device__ inline size_t ws_ticket_off(int B) { return ws_npos_off(B) + (size_t)B * sizeof(int); }      // unsigned match_ticket[B]
__host__ __device__ inline size_t ws_lticket_off(int B) { return ws_ticket_off(B) + (size_t)B * sizeof(unsigned); }   // unsigned loss_ticket[B]
__host__ __device__ inline size_t ws_clean_end(int B) { return ws_lticket_off(B) + (size_t)B * sizeof(unsigned); }    // [0, here) is left zero
__host__ __device__ inline size_t ws_part_off(int B) { return ws_al(ws_clean_end(B)); }                          // MatchPart part[B][MATCH_MAX_CHUNKS]
__host__ __device__ inline size_t ws_tile_off(int B) { return ws_part_off(B) + (size_t)B * MATCH_MAX_CHUNKS * sizeof(MatchPart); }   // double2 tile[B][LOSS_MAX_TILES]
__host__ __device__ inline size_t ws_total(int B) { return ws_tile_off(B) + (size_t)B * LOSS_MAX_TILES * sizeof(double2); }

// IoU of the match pass: exactly iou_gt_anchor, but the division is skipped where the boxes do not intersect
// (inter == 0 gives 0 / (positive) == +0.0 in the reference too; most of the 17460 anchors of a row are such).
__device__ __forceinline__ double iou_match(float g0, float g1, float g2, float g3, const Box4& an, bool row_nan) {
  const double tl0 = fmax((double)g0, an.y1), tl1 = fmax((double)g1, an.x1);
  const double br0 = fmin((double)g2, an.y2), br1 = fmin((double)g3, an.x2);
  const double s0 = fmax(__dsub_rn(br0, tl0), 0.0), s1 = fmax(__dsub_rn(br1, tl1), 0.0);
  const double inter = __dmul_rn(s0, s1);
  double v = 0.0;
  if (inter != 0.0) {
    const float gh = __fsub_rn(g2, g0), gw = __fsub_rn(g3, g1);
    const double gt_area = (double)__fmul_rn(gh, gw);                // float32 product, then promoted
    const double an_area = __dmul_rn(__dsub_rn(an.y2, an.y1), __dsub_rn(an.x2, an.x1));
    const double uni = __dsub_rn(__dadd_rn(gt_area, an_area), inter);
    v = __ddiv_rn(inter, __dadd_rn(uni, 1e-8));
  }
  // torch.max / torch.min propagate NaN (CUDA's fmax / fmin drop it): a NaN annotation makes the whole IoU row NaN
  return row_nan ? __longlong_as_double(0x7ff8000000000000LL) : v;
}

// ---- pass 1: grid (chunks, B): IoU of a chunk of one row, positives, partial first-index argmax; the last chunk of a
// row to finish (ticket) reduces the partials: top-1 anchor, positives per row and in total.  With one CTA per sample
// (round 1) 64 CTAs ran on 148 SMs and the pass took as long as the loss pass itself.
__global__ void __launch_bounds__(256) match_rows_kernel(const float* __restrict__ annot,
                                                         const double* __restrict__ anchors, int A, int per_chunk,
                                                         double thr, int use_multi, uint8_t* __restrict__ pos,
                                                         int64_t* __restrict__ top1, uint8_t* __restrict__ wsb, int B) {
  __shared__ ArgMaxD sm[32];
  __shared__ int cnt_sm[8];
  __shared__ bool last;
  LossWs* ws = reinterpret_cast<LossWs*>(wsb);
  int* npos_row = reinterpret_cast<int*>(wsb + ws_npos_off(B));
  unsigned* row_ticket = reinterpret_cast<unsigned*>(wsb + ws_ticket_off(B));
  MatchPart* part = reinterpret_cast<MatchPart*>(wsb + ws_part_off(B));
  const int b = blockIdx.y, ch = blockIdx.x, nch = gridDim.x;
  const float g0 = annot[4 * b], g1 = annot[4 * b + 1], g2 = annot[4 * b + 2], g3 = annot[4 * b + 3];
  const bool row_nan = (g0 != g0) || (g1 != g1) || (g2 != g2) || (g3 != g3);
  ArgMaxD best{ARG_NONE, 0x7fffffff};
  int cnt = 0;
  uint8_t* prow = pos + (size_t)b * A;
  const int a_end = min(A, (ch + 1) * per_chunk);
  for (int a = ch * per_chunk + threadIdx.x; a < a_end; a += blockDim.x) {
    const double v = iou_match(g0, g1, g2, g3, load_anchor(anchors, a), row_nan);
    best = better(best, ArgMaxD{v, a});
    const int p = (use_multi && v > thr) ? 1 : 0;
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
    part[b * MATCH_MAX_CHUNKS + ch] = MatchPart{best.v, best.i, total};
    __threadfence();                                                // partial and this chunk's pos bytes before the ticket
    last = atomicAdd(&row_ticket[b], 1u) == (unsigned)nch - 1;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < 32) {
    __threadfence();
    ArgMaxD x{ARG_NONE, 0x7fffffff};
    int c = 0;
    if ((int)threadIdx.x < nch) {
      const volatile MatchPart* q = part + b * MATCH_MAX_CHUNKS + threadIdx.x;
      x = ArgMaxD{q->v, q->i};
      c = q->cnt;
    }
    x = warp_argmax(x);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (threadIdx.x == 0) {
      const int idx = min(max(x.i, 0), A - 1);
      if (!(use_multi && x.v > thr)) c += 1;                        // the top-1 anchor is always positive
      prow[idx] = 1;
      top1[b] = idx;
      npos_row[b] = c;
      atomicAdd(&ws->npos_total, (unsigned long long)c);
      row_ticket[b] = 0;                                            // leave the workspace clean
    }
  }
}

// ---- pass 2: per (sample, anchor) loss terms and gradients; the last block (ticket) turns the sums into the three
// loss scalars, applies the reference's NaN guard and clears the workspace.  PACKED: att / reg and d_att / d_reg are the
// [..., 4] / [..., :4] views of [B, A, 5] buffers (what the head writes, mdl.py:381-382): the block's 256 anchors are
// 5120 contiguous bytes, moved with 128-bit accesses through shared memory instead of five stride-5 scalar loads and
// stores per thread.
struct LossTerms { float d_att; float4 d_reg; double cls_l, box_l; };

// box term of a POSITIVE anchor (a few per row of 17460): fp64 targets, smooth-L1 and its gradient.  (Measured out of
// line: 48 instead of 76 registers, 5 instead of 3 blocks per SM -- and 21-24 us instead of 18 us per pass at B = 64.)
__device__ __forceinline__ void box_terms(const float* __restrict__ annot, const double* __restrict__ anchors, int b, int a, int B,
                                       float r0, float r1, float r2, float r3, double lamb_reg, int npos_b, float4* d_reg,
                                       double* box_l) {
  const float r[4] = {r0, r1, r2, r3};
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
  const double gscale = lamb_reg / ((double)B * (double)npos_b);
  float dv[4];
  double bl = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double d = (double)r[k] - tg[k];
    const double ad = fabs(d);
    bl += (ad < 1.0) ? 0.5 * d * d : ad - 0.5;                      // SmoothL1, beta = 1 (loss.py:41,91)
    dv[k] = (float)(gscale * fmin(fmax(d, -1.0), 1.0));
  }
  *d_reg = make_float4(dv[0], dv[1], dv[2], dv[3]);
  *box_l = bl;
}

__device__ __forceinline__ LossTerms loss_terms(float x, const float r[4], bool t, int b, int a, int B,
                                                const float* __restrict__ annot, const double* __restrict__ anchors,
                                                float alpha, float gamma, double lamb_reg, float inv_np, int npos_b) {
  LossTerms o;
  // loss.py:103-125 in float32, like the reference
  const float p = 1.0f / (1.0f + expf(-x));
  float w = t ? (1.0f - p) : p;
  w = (gamma == 2.0f) ? w * w : powf(w, gamma);
  w *= t ? (1.0f - alpha) : alpha;                                  // alpha weights the negatives
  const float tf = t ? 1.0f : 0.0f;
  const float bce = fmaxf(x, 0.0f) - x * tf + log1pf(expf(-fabsf(x)));
  o.cls_l = (double)(w * bce);
  o.d_att = w * (p - tf) * inv_np;
  o.d_reg = make_float4(0.f, 0.f, 0.f, 0.f);
  o.box_l = 0.0;
  if (t)                                                            // few per row: divergence is cheap
    box_terms(annot, anchors, b, a, B, r[0], r[1], r[2], r[3], lamb_reg, npos_b, &o.d_reg, &o.box_l);
  return o;
}

// block-wide sum of a double (all threads call; result valid in thread 0); `red` holds 8 doubles
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();                                                  // previous use of `red` is over
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// grid (tiles of 256 anchors, B).  Every block leaves its two partial sums in its own slot and is done: no atomics, no
// fence, nothing serial at the end of a block (with ~30 waves of blocks per SM, a fence + ticket per block was most of the
// kernel's duration; same-address atomicAdds on the sums, round 1, retire one every ~2.5-5 ns: 22 us for 4416 blocks).
// loss_finalize_kernel adds the slots up.
template <bool PACKED>
__global__ void __launch_bounds__(256) loss_grad_kernel(
    const float* __restrict__ att, int64_t att_stride, const float* __restrict__ reg, int64_t reg_stride,
    const float* __restrict__ annot, const double* __restrict__ anchors, const uint8_t* __restrict__ pos, int B, int A,
    float alpha, float gamma, double lamb_reg, float* __restrict__ d_att, int64_t d_att_stride,
    float* __restrict__ d_reg, int64_t d_reg_stride, uint8_t* __restrict__ wsb) {
  __shared__ __align__(16) float tile[PACKED ? 256 * 5 : 4];
  __shared__ double red[8];
  const LossWs* ws = reinterpret_cast<const LossWs*>(wsb);
  const int* npos_row = reinterpret_cast<const int*>(wsb + ws_npos_off(B));
  double2* tiles = reinterpret_cast<double2*>(wsb + ws_tile_off(B));
  const int b = blockIdx.y;
  const int a0 = blockIdx.x * 256;
  const int a = a0 + threadIdx.x;
  const int na = min(256, A - a0);                                  // anchors of this block
  const float inv_np = 1.0f / (float)ws->npos_total;
  double cls_l = 0.0, box_l = 0.0;
  if (PACKED) {
    // reg points at element [0, 0, 0] of the packed buffer, d_reg likewise; rows start 16-byte aligned (A * 5 % 4 == 0)
    const float* src = reg + ((size_t)b * A + a0) * 5;
    float* dst = d_reg + ((size_t)b * A + a0) * 5;
    const int nf = na * 5;                                          // floats of this block; a0 * 5 % 4 == 0
    for (int i = threadIdx.x * 4; i < nf; i += 256 * 4) {
      if (i + 3 < nf) *reinterpret_cast<float4*>(tile + i) = __ldg(reinterpret_cast<const float4*>(src + i));
      else for (int k = i; k < nf; ++k) tile[k] = __ldg(src + k);
    }
    __syncthreads();
    LossTerms o;
    if (a < A) {
      const float r[4] = {tile[threadIdx.x * 5], tile[threadIdx.x * 5 + 1], tile[threadIdx.x * 5 + 2], tile[threadIdx.x * 5 + 3]};
      const float x = tile[threadIdx.x * 5 + 4];
      o = loss_terms(x, r, pos[(size_t)b * A + a] != 0, b, a, B, annot, anchors, alpha, gamma, lamb_reg, inv_np, npos_row[b]);
      cls_l = o.cls_l;
      box_l = o.box_l;
    }
    __syncthreads();
    if (a < A) {
      float* q = tile + threadIdx.x * 5;
      q[0] = o.d_reg.x; q[1] = o.d_reg.y; q[2] = o.d_reg.z; q[3] = o.d_reg.w; q[4] = o.d_att;
    }
    __syncthreads();
    for (int i = threadIdx.x * 4; i < nf; i += 256 * 4) {
      if (i + 3 < nf) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(tile + i);
      else for (int k = i; k < nf; ++k) dst[k] = tile[k];
    }
  } else if (a < A) {
    const size_t e = (size_t)b * A + a;
    const float* rp = reg + e * reg_stride;
    const float r[4] = {rp[0], rp[1], rp[2], rp[3]};
    const LossTerms o = loss_terms(att[e * att_stride], r, pos[e] != 0, b, a, B, annot, anchors, alpha, gamma, lamb_reg,
                                   inv_np, npos_row[b]);
    cls_l = o.cls_l;
    box_l = o.box_l;
    d_att[e * d_att_stride] = o.d_att;
    float* dp = d_reg + e * d_reg_stride;
    if (d_reg_stride == 4) {
      *reinterpret_cast<float4*>(dp) = o.d_reg;
    } else {
      dp[0] = o.d_reg.x; dp[1] = o.d_reg.y; dp[2] = o.d_reg.z; dp[3] = o.d_reg.w;
    }
  }
  const double c = block_sum(cls_l, red), bx = block_sum(box_l, red);
  if (threadIdx.x == 0) tiles[(size_t)b * LOSS_MAX_TILES + blockIdx.x] = make_double2(c, bx);
}

// The product path (packed [B, A, 5] buffers, what the head writes): a block walks `tpb` consecutive tiles of 256 anchors of
// one row and leaves ONE pair of partial sums.  Nothing in the tile loop is block-wide: a warp owns 32 anchors = 160
// contiguous floats, moves them with 128-bit accesses through its own 640-byte slab of shared memory (__syncwarp only) and
// has the next tile's loads in flight while it computes.  With one tile per block (above) the pass was a chain of ~10
// block-wide phases for 256 anchors -- 4416 blocks of ~7 us lifetime at B = 64 -- and sat at a quarter of the HBM roof.
__global__ void __launch_bounds__(256) loss_grad_packed_kernel(
    const float* __restrict__ reg, const float* __restrict__ annot, const double* __restrict__ anchors,
    const uint8_t* __restrict__ pos, int B, int A, int tpb, float alpha, float gamma, double lamb_reg,
    float* __restrict__ d_reg, uint8_t* __restrict__ wsb) {
  __shared__ __align__(16) float slab[8][160];
  __shared__ double red[8];
  const LossWs* ws = reinterpret_cast<const LossWs*>(wsb);
  const int* npos_row = reinterpret_cast<const int*>(wsb + ws_npos_off(B));
  double2* tiles = reinterpret_cast<double2*>(wsb + ws_tile_off(B));
  const int b = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_np = 1.0f / (float)ws->npos_total;
  const int npos_b = npos_row[b];
  float* my = slab[w];
  const int t0 = blockIdx.x * tpb;
  double cls_l = 0.0, box_l = 0.0;
  // the warp's 32 anchors of tile t: floats [0, nf) of src / dst below; nf % 4 == 0 because A % 4 == 0
  auto first = [&](int t) { return t * 256 + w * 32; };
  float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
  uint8_t cp = 0;
  {
    const int a0 = first(t0), nf = max(0, min(32, A - a0)) * 5;
    const float* src = reg + ((size_t)b * A + a0) * 5;
    if (lane * 4 < nf) c0 = __ldg(reinterpret_cast<const float4*>(src) + lane);
    if (128 + lane * 4 < nf) c1 = __ldg(reinterpret_cast<const float4*>(src) + 32 + lane);
    if (a0 + lane < A) cp = pos[(size_t)b * A + a0 + lane];
  }
  for (int t = t0; t < t0 + tpb; ++t) {
    const int a0 = first(t);
    if (a0 >= A) break;                                             // warp-uniform
    const int nf = min(32, A - a0) * 5;
    // next tile's loads first: they are in flight while this one is computed
    float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f), n1 = n0;
    uint8_t np_ = 0;
    if (t + 1 < t0 + tpb) {
      const int a1 = first(t + 1), nf1 = max(0, min(32, A - a1)) * 5;
      const float* src = reg + ((size_t)b * A + a1) * 5;
      if (lane * 4 < nf1) n0 = __ldg(reinterpret_cast<const float4*>(src) + lane);
      if (128 + lane * 4 < nf1) n1 = __ldg(reinterpret_cast<const float4*>(src) + 32 + lane);
      if (a1 + lane < A) np_ = pos[(size_t)b * A + a1 + lane];
    }
    if (lane * 4 < nf) *reinterpret_cast<float4*>(my + lane * 4) = c0;
    if (128 + lane * 4 < nf) *reinterpret_cast<float4*>(my + 128 + lane * 4) = c1;
    __syncwarp();
    const bool live = a0 + lane < A;
    LossTerms o;
    if (live) {
      const float r[4] = {my[lane * 5], my[lane * 5 + 1], my[lane * 5 + 2], my[lane * 5 + 3]};
      o = loss_terms(my[lane * 5 + 4], r, cp != 0, b, a0 + lane, B, annot, anchors, alpha, gamma, lamb_reg, inv_np, npos_b);
      cls_l += o.cls_l;
      box_l += o.box_l;
    }
    __syncwarp();
    if (live) {
      float* q = my + lane * 5;
      q[0] = o.d_reg.x; q[1] = o.d_reg.y; q[2] = o.d_reg.z; q[3] = o.d_reg.w; q[4] = o.d_att;
    }
    __syncwarp();
    float* dst = d_reg + ((size_t)b * A + a0) * 5;
    if (lane * 4 < nf) *reinterpret_cast<float4*>(dst + lane * 4) = *reinterpret_cast<const float4*>(my + lane * 4);
    if (128 + lane * 4 < nf) *reinterpret_cast<float4*>(dst + 128 + lane * 4) = *reinterpret_cast<const float4*>(my + 128 + lane * 4);
    __syncwarp();
    c0 = n0; c1 = n1; cp = np_;
  }
  const double c = block_sum(cls_l, red), bx = block_sum(box_l, red);
  if (threadIdx.x == 0) tiles[(size_t)b * LOSS_MAX_TILES + blockIdx.x] = make_double2(c, bx);
}

// one block: rows in parallel (each adds its tiles in tile order), then the rows in row order (loss.py:91-143)
__global__ void __launch_bounds__(256) loss_finalize_kernel(uint8_t* __restrict__ wsb, int B, int A, int ntiles, double lamb_reg,
                                                            double* __restrict__ losses, float* __restrict__ d_att,
                                                            int64_t d_att_stride, float* __restrict__ d_reg,
                                                            int64_t d_reg_stride) {
  LossWs* ws = reinterpret_cast<LossWs*>(wsb);
  double* row_cls = reinterpret_cast<double*>(wsb + ws_rowcls_off());
  double* row_box = reinterpret_cast<double*>(wsb + ws_box_off(B));
  const int* npos_row = reinterpret_cast<const int*>(wsb + ws_npos_off(B));
  const double2* tiles = reinterpret_cast<const double2*>(wsb + ws_tile_off(B));
  __shared__ int bad_s;
  constexpr int SROWS = 1024;                                       // rows kept in shared memory for the serial row-order sum
  __shared__ double s_cls[SROWS], s_box[SROWS];
  // LPR lanes per row (a power of two, 256 / B clamped to 1..32): lane `sub` adds the slots sub, sub + LPR, ... in order, a
  // butterfly adds the lanes -- a fixed order, so the scalars stay bit-reproducible -- instead of one thread walking all of a
  // row's slots through dependent L2 round trips (that chain was ~3 us of the 8 us this one-block kernel takes at B = 64)
  int lpr = 1;
  while (lpr < 32 && lpr * 2 * B <= (int)blockDim.x) lpr *= 2;
  const int rows_per_pass = blockDim.x / lpr;
  const int sub = threadIdx.x % lpr;
  for (int b0 = 0; b0 < B; b0 += rows_per_pass) {                     // (uniform trip count: the shuffles below are warp-wide)
    const int b = b0 + threadIdx.x / lpr;
    double rc = 0.0, rb = 0.0;
    if (b < B) {
      const double2* t = tiles + (size_t)b * LOSS_MAX_TILES;
      for (int i = sub; i < ntiles; i += lpr) { rc += t[i].x; rb += t[i].y; }
    }
    for (int off = lpr >> 1; off > 0; off >>= 1) {
      rc += __shfl_xor_sync(0xffffffffu, rc, off);
      rb += __shfl_xor_sync(0xffffffffu, rb, off);
    }
    if (b < B && sub == 0) {
      rb = rb / (double)(float)npos_row[b];                         // loss.py:93: row sum / positives of the row (float count)
      row_cls[b] = rc;
      row_box[b] = rb;
      if (b < SROWS) { s_cls[b] = rc; s_box[b] = rb; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double box = 0.0, clsd = 0.0;
    for (int i = 0; i < B; ++i) {
      box += i < SROWS ? s_box[i] : row_box[i];
      clsd += i < SROWS ? s_cls[i] : row_cls[i];
    }
    box /= (double)B;
    float cls = (float)clsd / (float)ws->npos_total;                // f32 / count, like loss.py:125
    // loss.py:128-133.  The reference multiplies the per-anchor box loss of ALL anchors by the mask (loss.py:92): a zero-area
    // or inverted box (log(0) / log(<0) targets) gives inf * 0 = NaN there, while only positives are evaluated here (inf):
    // an infinite box loss is the same condition.
    const int bad = (box != box) || (cls != cls) || isinf(box);
    if (bad) { box = 0.01; cls = 1.0f; }
    losses[0] = lamb_reg * box + (double)cls;
    losses[1] = (double)cls;
    losses[2] = box;
    ws->nan_flag = bad;
    bad_s = bad;
  }
  __syncthreads();
  if (bad_s) {
    // a NaN step carries no gradient (the constants above have none); rare, hence one block
    const size_t n = (size_t)B * A;
    for (size_t e = threadIdx.x; e < n; e += blockDim.x) {
      d_att[e * d_att_stride] = 0.f;
      for (int k = 0; k < 4; ++k) d_reg[e * d_reg_stride + k] = 0.f;
    }
  }
}

// ---- evaluator ------------------------------------------------------------------------------
struct ArgMaxF { float v; int i; };
__device__ __forceinline__ ArgMaxF betterf(ArgMaxF a, ArgMaxF b) {   // torch.max semantics, see better()
  const bool an = a.v != a.v, bn = b.v != b.v;
  const bool gt = bn ? !an : (!an && b.v > a.v);
  const bool eq = (an && bn) || b.v == a.v;
  return (gt || (eq && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMaxF warp_argmaxf(ArgMaxF x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMaxF y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = betterf(x, y);
  }
  return x;
}

__device__ __forceinline__ Box4 decode_box(const Box4& an, const float* r) {   // anchors.py:182-197
  const double ac0 = __ddiv_rn(__dadd_rn(an.y1, an.y2), 2.0), ac1 = __ddiv_rn(__dadd_rn(an.x1, an.x2), 2.0);
  const double ah = __dsub_rn(an.y2, an.y1), aw = __dsub_rn(an.x2, an.x1);
  const double c0 = __dadd_rn(__dmul_rn(ah, (double)r[0]), ac0), c1 = __dadd_rn(__dmul_rn(aw, (double)r[1]), ac1);
  const double h = __dmul_rn((double)expf(r[2]), ah), w = __dmul_rn((double)expf(r[3]), aw);   // exp in float32
  const double hh = __ddiv_rn(h, 2.0), hw = __ddiv_rn(w, 2.0);
  return {__dsub_rn(c0, hh), __dsub_rn(c1, hw), __dadd_rn(c0, hh), __dadd_rn(c1, hw)};
}

// grid (chunks, B) like the match pass: partial argmaxes of the IoU row and of the scores; the last chunk of a row
// decodes the two selected boxes; the last row to finish (second ticket) averages the flags into Acc / MaxPos.
// scratch (behind the flags in `metrics`, see zsg_eval): EvalPart[B][chunks], unsigned row_ticket[B], unsigned done.
struct EvalPart { double iv; int ii; float sv; int si; int pad; };
constexpr int EVAL_CHUNKS = 8;

__global__ void __launch_bounds__(256) eval_rows_kernel(const float* __restrict__ att, int64_t att_stride,
                                                        const float* __restrict__ reg, int64_t reg_stride,
                                                        const float* __restrict__ annot,
                                                        const double* __restrict__ anchors,
                                                        const float* __restrict__ img_size, int A, int per_chunk,
                                                        double thr, int64_t* __restrict__ best_ids,
                                                        float* __restrict__ scores, double* __restrict__ pred_boxes,
                                                        float* __restrict__ metrics, EvalPart* __restrict__ part,
                                                        unsigned* __restrict__ tickets) {
  __shared__ ArgMaxD smd[32];
  __shared__ ArgMaxF smf[8];
  __shared__ int last;
  const int b = blockIdx.y, ch = blockIdx.x, nch = gridDim.x, B = gridDim.y;
  float* flags = metrics + 2;
  const float g0 = annot[4 * b], g1 = annot[4 * b + 1], g2 = annot[4 * b + 2], g3 = annot[4 * b + 3];
  const bool row_nan = (g0 != g0) || (g1 != g1) || (g2 != g2) || (g3 != g3);
  ArgMaxD bi{ARG_NONE, 0x7fffffff};
  ArgMaxF bs{-1.0f, 0x7fffffff};
  const int a_end = min(A, (ch + 1) * per_chunk);
  for (int a = ch * per_chunk + threadIdx.x; a < a_end; a += blockDim.x) {
    bi = better(bi, ArgMaxD{iou_match(g0, g1, g2, g3, load_anchor(anchors, a), row_nan), a});
    const float x = att[((size_t)b * A + a) * att_stride];
    bs = betterf(bs, ArgMaxF{1.0f / (1.0f + expf(-x)), a});          // evaluator.py:74-75
  }
  bi = block_argmax(bi, smd);
  bs = warp_argmaxf(bs);
  if ((threadIdx.x & 31) == 0) smf[threadIdx.x >> 5] = bs;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) bs = betterf(bs, smf[w]);
    part[b * EVAL_CHUNKS + ch] = EvalPart{bi.v, bi.i, bs.v, bs.i, 0};
    __threadfence();
    last = atomicAdd(&tickets[b], 1u) == (unsigned)nch - 1;
    if (last) {
      __threadfence();
      const volatile EvalPart* q = part + b * EVAL_CHUNKS;
      bi = ArgMaxD{ARG_NONE, 0x7fffffff};
      bs = ArgMaxF{-1.0f, 0x7fffffff};
      for (int c = 0; c < nch; ++c) {
        bi = better(bi, ArgMaxD{q[c].iv, q[c].ii});
        bs = betterf(bs, ArgMaxF{q[c].sv, q[c].si});
      }
      const int si = min(max(bs.i, 0), A - 1), ii = min(max(bi.i, 0), A - 1);
      const Box4 bx = decode_box(load_anchor(anchors, si), reg + ((size_t)b * A + si) * reg_stride);
      const Box4 mx = decode_box(load_anchor(anchors, ii), reg + ((size_t)b * A + ii) * reg_stride);
      flags[b] = iou_box_gt(bx, g0, g1, g2, g3) >= thr ? 1.f : 0.f;
      flags[B + b] = iou_box_gt(mx, g0, g1, g2, g3) >= thr ? 1.f : 0.f;
      best_ids[b] = si;
      scores[b] = bs.v;
      // evaluator.py:96-97: ((box+1)/2) * (h,w), then y1x1y2x2 -> x1y1x2y2
      const double h = (double)img_size[2 * b], w = (double)img_size[2 * b + 1];
      const double py1 = __dmul_rn(h, __ddiv_rn(__dadd_rn(bx.y1, 1.0), 2.0)), px1 = __dmul_rn(w, __ddiv_rn(__dadd_rn(bx.x1, 1.0), 2.0));
      const double py2 = __dmul_rn(h, __ddiv_rn(__dadd_rn(bx.y2, 1.0), 2.0)), px2 = __dmul_rn(w, __ddiv_rn(__dadd_rn(bx.x2, 1.0), 2.0));
      double* o = pred_boxes + 4 * b;
      o[0] = px1; o[1] = py1; o[2] = px2; o[3] = py2;
      tickets[b] = 0;
      __threadfence();
      if (atomicAdd(&tickets[B], 1u) == (unsigned)B - 1) {           // last row: the two means (evaluator.py:108-117)
        __threadfence();
        float acc = 0.f, mp = 0.f;
        for (int i = 0; i < B; ++i) { acc += __ldcg(flags + i); mp += __ldcg(flags + B + i); }
        metrics[0] = acc / (float)B;
        metrics[1] = mp / (float)B;
        tickets[B] = 0;
      }
    }
  }
}

}  // namespace zsg

using namespace zsg;

extern "C" size_t zsg_match_loss_workspace_bytes(int b) { return ws_total(b < 1 ? 1 : b); }

extern "C" size_t zsg_eval_workspace_bytes(int b) {
  return (size_t)(b < 1 ? 1 : b) * EVAL_CHUNKS * sizeof(EvalPart) + ((size_t)b + 1) * sizeof(unsigned) + 16;
}

static int check_ws(const char* who, void* workspace, size_t ws_bytes, int b, int a) {
  ZSG_REQUIRE(workspace, "%s: null workspace", who);
  ZSG_REQUIRE(b > 0 && a > 0 && b <= 65535, "%s: bad batch (b=%d a=%d)", who, b, a);
  ZSG_REQUIRE(ws_bytes >= zsg_match_loss_workspace_bytes(b), "%s: workspace too small", who);
  ZSG_REQUIRE(((uintptr_t)workspace & 15) == 0, "%s: workspace must be 16-byte aligned", who);
  ZSG_REQUIRE((a + 255) / 256 <= LOSS_MAX_TILES, "%s: a=%d exceeds %d anchors per row", who, a, LOSS_MAX_TILES * 256);
  return ZSG_OK;
}

extern "C" int zsg_match(const float* annot, const double* anchors, int b, int a, double match_thr, int use_multi,
                         int64_t* top1, uint8_t* pos, void* workspace, size_t ws_bytes, zsg_stream_t stream) {
  ZSG_REQUIRE(annot && anchors && top1 && pos, "zsg_match: null pointer");
  if (int rc = check_ws("zsg_match", workspace, ws_bytes, b, a)) return rc;
  ZSG_REQUIRE(((uintptr_t)anchors & 15) == 0, "zsg_match: anchors must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
  cudaMemsetAsync(wsb, 0, ws_clean_end(b), st);                     // counters and tickets of this call
  // chunks per row: enough CTAs for ~4 per SM, at most MATCH_MAX_CHUNKS, at least 256 anchors each
  int nch = (num_sms() * 4 + b - 1) / b;
  if (nch > MATCH_MAX_CHUNKS) nch = MATCH_MAX_CHUNKS;
  if (nch > (a + 255) / 256) nch = (a + 255) / 256;
  if (nch < 1) nch = 1;
  const int per_chunk = (a + nch - 1) / nch;
  nch = (a + per_chunk - 1) / per_chunk;
  match_rows_kernel<<<dim3(nch, b), 256, 0, st>>>(annot, anchors, a, per_chunk, match_thr, use_multi, pos, top1, wsb, b);
  return check_launch("zsg_match");
}

extern "C" int zsg_loss_grad(const float* att, int64_t att_stride, const float* reg, int64_t reg_stride, const float* annot,
                             const double* anchors, const uint8_t* pos, int b, int a, float alpha, float gamma,
                             double lamb_reg, double* losses, float* d_att, int64_t d_att_stride, float* d_reg,
                             int64_t d_reg_stride, void* workspace, size_t ws_bytes, zsg_stream_t stream) {
  ZSG_REQUIRE(att && reg && annot && anchors && pos && losses && d_att && d_reg, "zsg_loss_grad: null pointer");
  if (int rc = check_ws("zsg_loss_grad", workspace, ws_bytes, b, a)) return rc;
  ZSG_REQUIRE(((uintptr_t)anchors & 15) == 0, "zsg_loss_grad: anchors must be 16-byte aligned");
  ZSG_REQUIRE(d_reg_stride != 4 || ((uintptr_t)d_reg & 15) == 0, "zsg_loss_grad: d_reg must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
  int ntiles = (a + 255) / 256;
  dim3 grid(ntiles, b);
  const bool packed = att_stride == 5 && reg_stride == 5 && d_att_stride == 5 && d_reg_stride == 5 && att == reg + 4 &&
                      d_att == d_reg + 4 && ((((uintptr_t)reg | (uintptr_t)d_reg) & 15) == 0) && (a % 4 == 0);
  if (packed) {
    // tiles per block (measured sweep, tools/time_loss.py: 1 -> 29.9 us, 2 -> 23.8, 4 -> 18.0 at B = 64): about one wave of
    // 8 blocks per SM
    int tpb = (int)(((int64_t)b * ntiles + (int64_t)num_sms() * 8 - 1) / ((int64_t)num_sms() * 8));
    tpb = tpb < 1 ? 1 : (tpb > 16 ? 16 : tpb);
    if (const char* e = getenv("ZSG_LOSS_TPB")) tpb = atoi(e) > 0 ? atoi(e) : tpb;      // diagnostics
    ntiles = (ntiles + tpb - 1) / tpb;                      // slots per row the finalize kernel adds up
    loss_grad_packed_kernel<<<dim3(ntiles, b), 256, 0, st>>>(reg, annot, anchors, pos, b, a, tpb, alpha, gamma, lamb_reg, d_reg,
                                                              wsb);
  } else
    loss_grad_kernel<false><<<grid, 256, 0, st>>>(att, att_stride, reg, reg_stride, annot, anchors, pos, b, a, alpha, gamma,
                                                  lamb_reg, d_att, d_att_stride, d_reg, d_reg_stride, wsb);
  loss_finalize_kernel<<<1, 256, 0, st>>>(wsb, b, a, ntiles, lamb_reg, losses, d_att, d_att_stride, d_reg, d_reg_stride);
  return check_launch("zsg_loss_grad");
}

extern "C" int zsg_match_loss(const float* att, int64_t att_stride, const float* reg, int64_t reg_stride,
                              const float* annot, const double* anchors, int b, int a, double match_thr, float alpha,
                              float gamma, double lamb_reg, int use_multi, double* losses, float* d_att,
                              int64_t d_att_stride, float* d_reg, int64_t d_reg_stride, int64_t* top1, uint8_t* pos,
                              void* workspace, size_t ws_bytes, zsg_stream_t stream) {
  if (int rc = zsg_match(annot, anchors, b, a, match_thr, use_multi, top1, pos, workspace, ws_bytes, stream)) return rc;
  return zsg_loss_grad(att, att_stride, reg, reg_stride, annot, anchors, pos, b, a, alpha, gamma, lamb_reg, losses, d_att,
                       d_att_stride, d_reg, d_reg_stride, workspace, ws_bytes, stream);
}

extern "C" int zsg_eval(const float* att, int64_t att_stride, const float* reg, int64_t reg_stride,
                        const float* annot, const double* anchors, const float* img_size, int b, int a,
                        double iou_thr, int64_t* best_ids, float* pred_scores, double* pred_boxes, float* metrics,
                        void* workspace, size_t ws_bytes, zsg_stream_t stream) {
  ZSG_REQUIRE(att && reg && annot && anchors && img_size && best_ids && pred_scores && pred_boxes && metrics && workspace,
              "zsg_eval: null pointer");
  ZSG_REQUIRE(b > 0 && a > 0 && b <= 65535, "zsg_eval: bad batch");
  ZSG_REQUIRE(ws_bytes >= zsg_eval_workspace_bytes(b) && ((uintptr_t)workspace & 15) == 0, "zsg_eval: workspace too small or misaligned");
  cudaStream_t st = as_stream(stream);
  int nch = (num_sms() * 2 + b - 1) / b;
  if (nch > EVAL_CHUNKS) nch = EVAL_CHUNKS;
  if (nch > (a + 255) / 256) nch = (a + 255) / 256;
  if (nch < 1) nch = 1;
  const int per_chunk = (a + nch - 1) / nch;
  nch = (a + per_chunk - 1) / per_chunk;
  EvalPart* part = reinterpret_cast<EvalPart*>(workspace);
  unsigned* tickets = reinterpret_cast<unsigned*>(part + (size_t)b * EVAL_CHUNKS);
  // flags live behind the two metrics: metrics must hold 2 + 2*b floats
  eval_rows_kernel<<<dim3(nch, b), 256, 0, st>>>(att, att_stride, reg, reg_stride, annot, anchors, img_size, a, per_chunk,
                                                 iou_thr, best_ids, pred_scores, pred_boxes, metrics, part, tickets);
  return check_launch("zsg_eval");
}
