// HBM-bound glue of the ZSGNet hot path: BatchNorm (train) statistics / apply / backward,
// stem max-pool, FPN nearest upsample + add, global average pool, language/grid tiling,
// bias-gradient column sums, layout helpers and Adam.  All tensors NHWC float32, channel
// counts multiples of 4 so that every access is a 16-byte vector, channels fastest => coalesced.
#include <cuda_bf16.h>
#include <math.h>
#include "common.cuh"

namespace zsg {

static inline int grid_for(int64_t work_items, int threads, int waves = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 fma4(float4 x, float4 s, float4 b) {
  return make_float4(fmaf(x.x, s.x, b.x), fmaf(x.y, s.y, b.y), fmaf(x.z, s.z, b.z), fmaf(x.w, s.w, b.w));
}
__device__ __forceinline__ float4 relu4(float4 v) {
  return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}
__device__ __forceinline__ float4 tf32_lo4(float4 v) {          // remainder after TF32 truncation (operand image)
  float4 l;
  l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  return l;
}

// bf16 operand image (the bf16 GEMM path of configs 3-5): round-to-nearest-even, 4 values = one 8-byte store
__device__ __forceinline__ void st4_bf16(uint16_t* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a);
  u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
// Activation tensors are float32, or -- trunk activations of the bf16 engine (conv outputs, block outputs: "bf16 storage")
// -- bfloat16.  Kernels take them as const void* plus a flag word; e = element index (a multiple of 4).
enum { TF_ACT = 1,    // x / r / act_out (forward activations) are bfloat16
       TF_DY = 2,     // dy is bfloat16
       TF_DZ = 4,     // dz_out is bfloat16
       TF_Y = 8 };    // y (the kernel's forward output) is bfloat16
__device__ __forceinline__ float4 ldx4(const void* p, bool b16, size_t e) {
  if (b16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p) + e);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xFFFF0000u));
  }
  return ld4(reinterpret_cast<const float*>(p) + e);
}
__device__ __forceinline__ void stx4(void* p, bool b16, size_t e, float4 v) {
  if (b16) st4_bf16(reinterpret_cast<uint16_t*>(p) + e, v);
  else st4(reinterpret_cast<float*>(p) + e, v);
}
// operand image of a value the kernel has in registers: TF32 remainder (float) or bf16 copy, whichever the caller passed
__device__ __forceinline__ void st_image(float* lo, uint16_t* b16, int64_t i4, float4 v) {
  if (lo) st4(lo + i4, tf32_lo4(v));
  if (b16) st4_bf16(b16 + i4, v);
}

// ------------------------------------------------------------------------------------------
// Per-channel reductions over rows.  Block = 256 threads = (256/cols) row lanes x cols float4
// columns (cols = min(C/4, 256), power of two); grid.y walks column groups when C/4 > 256.
// Partial sums: fp32 over 8-row chunks, then double; one double atomic per (block, channel).
// ------------------------------------------------------------------------------------------
template <int MODE>   // 0: stats (sum x, sum x^2); 1: bn backward (sum dz, sum dz*xhat)
__global__ void __launch_bounds__(256) channel_reduce_kernel(
    const void* __restrict__ x, const void* __restrict__ dy, const void* __restrict__ act_out,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ scale,
    const float* __restrict__ shift, int mask_mode, void* __restrict__ dz_out, double* __restrict__ sums,
    int64_t rows, int C, int cols, int tf) {
  const bool xb = tf & TF_ACT, gb = tf & TF_DY, zb = tf & TF_DZ;
  const int lane_c = threadIdx.x % cols;
  const int lane_r = threadIdx.x / cols;
  const int rpb = 256 / cols;
  const int c = (blockIdx.y * cols + lane_c) * 4;
  double a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  float4 mu = make_float4(0, 0, 0, 0), is = mu, sc = mu, sh = mu;
  if (MODE == 1) {
    mu = ld4(mean + c);
    is = ld4(invstd + c);
    if (mask_mode == 1) { sc = ld4(scale + c); sh = ld4(shift + c); }
  }
  const int64_t stride = (int64_t)gridDim.x * rpb;
  int64_t r = (int64_t)blockIdx.x * rpb + lane_r;
  constexpr int U = MODE == 0 ? 8 : 4;                   // rows in flight per thread: the stats pass has one load stream
                                                         // (3.7 TB/s with 4 rows in flight), the backward pass two or three
  while (r < rows) {
    float4 v[U], g[U], a[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * stride;
      ok[u] = rr < rows;
      const size_t off = (size_t)(ok[u] ? rr : r) * C + c;
      v[u] = ldx4(x, xb, off);
      if (MODE == 1) {
        g[u] = ldx4(dy, gb, off);
        if (mask_mode == 2) a[u] = ldx4(act_out, xb, off);
      }
    }
    float f1[4] = {0, 0, 0, 0}, f2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      if (MODE == 0) {
        f1[0] += v[u].x; f1[1] += v[u].y; f1[2] += v[u].z; f1[3] += v[u].w;
        f2[0] = fmaf(v[u].x, v[u].x, f2[0]); f2[1] = fmaf(v[u].y, v[u].y, f2[1]);
        f2[2] = fmaf(v[u].z, v[u].z, f2[2]); f2[3] = fmaf(v[u].w, v[u].w, f2[3]);
      } else {
        float4 gg = g[u];
        if (mask_mode == 1) {
          const float4 t = fma4(v[u], sc, sh);
          gg.x = t.x > 0.f ? gg.x : 0.f; gg.y = t.y > 0.f ? gg.y : 0.f;
          gg.z = t.z > 0.f ? gg.z : 0.f; gg.w = t.w > 0.f ? gg.w : 0.f;
        } else if (mask_mode == 2) {
          gg.x = a[u].x > 0.f ? gg.x : 0.f; gg.y = a[u].y > 0.f ? gg.y : 0.f;
          gg.z = a[u].z > 0.f ? gg.z : 0.f; gg.w = a[u].w > 0.f ? gg.w : 0.f;
          if (dz_out) stx4(dz_out, zb, (size_t)(r + u * stride) * C + c, gg);
        }
        f1[0] += gg.x; f1[1] += gg.y; f1[2] += gg.z; f1[3] += gg.w;
        f2[0] = fmaf(gg.x, (v[u].x - mu.x) * is.x, f2[0]); f2[1] = fmaf(gg.y, (v[u].y - mu.y) * is.y, f2[1]);
        f2[2] = fmaf(gg.z, (v[u].z - mu.z) * is.z, f2[2]); f2[3] = fmaf(gg.w, (v[u].w - mu.w) * is.w, f2[3]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { a1[k] += (double)f1[k]; a2[k] += (double)f2[k]; }
    r += U * stride;
  }
  __shared__ double sm[2][256][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { sm[0][threadIdx.x][k] = a1[k]; sm[1][threadIdx.x][k] = a2[k]; }
  __syncthreads();
  if (lane_r == 0) {
    for (int j = 1; j < rpb; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) { a1[k] += sm[0][j * cols + lane_c][k]; a2[k] += sm[1][j * cols + lane_c][k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&sums[c + k], a1[k]);
      atomicAdd(&sums[C + c + k], a2[k]);
    }
  }
}

static int launch_channel_reduce(int mode, const void* x, const void* dy, const void* act_out, const float* mean,
                                 const float* invstd, const float* scale, const float* shift, int mask_mode,
                                 void* dz_out, double* sums, int64_t rows, int C, cudaStream_t st, int tf = 0) {
  ZSG_REQUIRE(C % 4 == 0 && C >= 4, "channel reduce: C=%d must be a multiple of 4", C);
  int c4 = C / 4;
  int cols = c4 >= 256 ? 256 : c4;
  ZSG_REQUIRE((cols & (cols - 1)) == 0 && c4 % cols == 0, "channel reduce: C/4=%d must be a power of two", c4);
  int rpb = 256 / cols;
  int gy = c4 / cols;
  int64_t want = (rows + (int64_t)rpb * 8 - 1) / ((int64_t)rpb * 8);
  int64_t cap = (int64_t)num_sms() * 8 / gy;
  int gx = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  dim3 grid(gx, gy);
  if (mode == 0)
    channel_reduce_kernel<0><<<grid, 256, 0, st>>>(x, dy, act_out, mean, invstd, scale, shift, mask_mode, dz_out,
                                                   sums, rows, C, cols, tf);
  else
    channel_reduce_kernel<1><<<grid, 256, 0, st>>>(x, dy, act_out, mean, invstd, scale, shift, mask_mode, dz_out,
                                                   sums, rows, C, cols, tf);
  return check_launch("channel_reduce");
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, int64_t rows, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                   float* running_var, float* mean, float* invstd, float* scale, float* shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double n = (double)rows;
  double m = sums[c] / n;
  double var = sums[C + c] / n - m * m;
  if (var < 0.0) var = 0.0;
  double is = 1.0 / sqrt(var + (double)eps);
  mean[c] = (float)m;
  invstd[c] = (float)is;
  float sc = gamma[c] * (float)is;
  scale[c] = sc;
  shift[c] = beta[c] - (float)m * sc;
  if (running_mean) {
    double unb = rows > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void bn_eval_affine_kernel(const float* rm, const float* rv, const float* gamma, const float* beta,
                                      float eps, int C, float* scale, float* shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float is = 1.0f / sqrtf(rv[c] + eps);
  float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const void* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, const void* __restrict__ r,
                                                       const float* __restrict__ rscale,
                                                       const float* __restrict__ rshift, int relu,
                                                       float* __restrict__ y, float* __restrict__ y_lo,
                                                       uint16_t* __restrict__ y_b16, int64_t n4, int c4, int tf) {
  const bool xb = tf & TF_ACT;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    float4 v = fma4(ldx4(x, xb, i * 4), ld4(scale + c), ld4(shift + c));
    if (r) {
      float4 q = ldx4(r, xb, i * 4);
      if (rscale) q = fma4(q, ld4(rscale + c), ld4(rshift + c));
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    if (relu) v = relu4(v);
    if (y) st4(y + i * 4, v);                              // bf16 storage: the bfloat16 tensor (y_b16) is the only copy
    st_image(y_lo, y_b16, i * 4, v);
  }
}

// operand preparation for the cp.async GEMM paths: z = relu?(x * scale + shift), lo = z - trunc_tf32(z)
__global__ void __launch_bounds__(256) split_act_kernel(const void* __restrict__ x, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int relu, float* __restrict__ z,
                                                        float* __restrict__ lo, uint16_t* __restrict__ b16, int64_t n4,
                                                        int c4, int tf) {
  const bool xb = tf & TF_ACT;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = ldx4(x, xb, i * 4);
    if (scale) {
      const int c = (int)(i % c4) * 4;
      v = fma4(v, ld4(scale + c), ld4(shift + c));
    }
    if (relu) v = relu4(v);
    if (z) st4(z + i * 4, v);
    st_image(lo, b16, i * 4, v);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(
    const void* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ scale,
    const float* __restrict__ shift, const void* __restrict__ act_out, int mask_mode,
    const double* __restrict__ sums, float* __restrict__ dx, float* __restrict__ dx_lo, uint16_t* __restrict__ dx_b16,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int C, int tf) {
  const bool xb = tf & TF_ACT, gb = tf & TF_DY;
  const int c4 = C / 4;
  const int64_t n4 = rows * c4;
  const float inv_n = 1.0f / (float)rows;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = (float)sums[c];
      if (dgamma) dgamma[c] = (float)sums[C + c];
    }
  }
  // four float4 per stream in flight per thread (one at a time this pass ran at 3.8 TB/s)
  constexpr int U = 4;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n4; i0 += U * nthreads) {
    float4 g[U], v[U], a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * nthreads;
      if (i < n4) {
        g[u] = ldx4(dy, gb, i * 4);
        v[u] = ldx4(x, xb, i * 4);
        if (mask_mode == 2) a[u] = ldx4(act_out, xb, i * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * nthreads;
      if (i >= n4) break;
      const int c = (int)(i % c4) * 4;
      float4 gg = g[u];
      if (mask_mode == 1) {
        const float4 t = fma4(v[u], ld4(scale + c), ld4(shift + c));
        gg.x = t.x > 0.f ? gg.x : 0.f; gg.y = t.y > 0.f ? gg.y : 0.f;
        gg.z = t.z > 0.f ? gg.z : 0.f; gg.w = t.w > 0.f ? gg.w : 0.f;
      } else if (mask_mode == 2) {
        gg.x = a[u].x > 0.f ? gg.x : 0.f; gg.y = a[u].y > 0.f ? gg.y : 0.f;
        gg.z = a[u].z > 0.f ? gg.z : 0.f; gg.w = a[u].w > 0.f ? gg.w : 0.f;
      }
      const float4 mu = ld4(mean + c), is = ld4(invstd + c), ga = ld4(gamma + c);
      const float s1[4] = {(float)sums[c] * inv_n, (float)sums[c + 1] * inv_n, (float)sums[c + 2] * inv_n,
                           (float)sums[c + 3] * inv_n};
      const float s2[4] = {(float)sums[C + c] * inv_n, (float)sums[C + c + 1] * inv_n, (float)sums[C + c + 2] * inv_n,
                           (float)sums[C + c + 3] * inv_n};
      float4 o;
      o.x = ga.x * is.x * (gg.x - s1[0] - (v[u].x - mu.x) * is.x * s2[0]);
      o.y = ga.y * is.y * (gg.y - s1[1] - (v[u].y - mu.y) * is.y * s2[1]);
      o.z = ga.z * is.z * (gg.z - s1[2] - (v[u].z - mu.z) * is.z * s2[2]);
      o.w = ga.w * is.w * (gg.w - s1[3] - (v[u].w - mu.w) * is.w * s2[3]);
      if (dx) st4(dx + i * 4, o);                         // bf16 path: only the image is consumed (by the two GEMMs that follow)
      st_image(dx_lo, dx_b16, i * 4, o);
    }
  }
}

// Column sums of the [parts][2][C] partial statistics a conv epilogue wrote (zsg_conv_params.stats) into the double
// sums of bn_finalize.  Block = 8 part-lanes x 32 channels; grid.y strides over the parts; one double atomic per
// (block, channel, moment), like channel_reduce_kernel.
__global__ void __launch_bounds__(256) bn_partials_kernel(const float* __restrict__ partials, int64_t parts, int C,
                                                          double* __restrict__ sums, int64_t rows,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, float momentum, float* running_mean,
                                                          float* running_var, float* mean, float* invstd, float* scale,
                                                          float* shift, int* __restrict__ tickets) {
  __shared__ double sh[2][8][32];
  __shared__ int last;
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double a = 0.0, b = 0.0;
  if (c < C) {
    // four parts (eight loads) in flight per thread: with one the pass was a chain of ~10 dependent L2 round trips
    const int64_t step = (int64_t)gridDim.y * 8;
    int64_t p = (int64_t)blockIdx.y * 8 + pl;
    for (; p + 3 * step < parts; p += 4 * step) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[2 * u] = partials[((p + u * step) * 2) * C + c];
        v[2 * u + 1] = partials[((p + u * step) * 2 + 1) * C + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a += (double)v[2 * u]; b += (double)v[2 * u + 1]; }
    }
    for (; p < parts; p += step) {
      a += (double)partials[(p * 2) * C + c];
      b += (double)partials[(p * 2 + 1) * C + c];
    }
  }
  sh[0][pl][cl] = a;
  sh[1][pl][cl] = b;
  __syncthreads();
  if (pl < 2 && c < C) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[pl][i][cl];
    atomicAdd(sums + pl * C + c, t);
  }
  if (!tickets) return;
  // fused finalize: the last block of this channel group to arrive (ticket counter, as in the threadfence-reduction
  // pattern) turns the completed sums into mean / invstd / scale / shift and the running statistics.
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&tickets[blockIdx.x], 1) == (int)gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0;           // ready for the next BatchNorm on this stream
  if (threadIdx.x < 32 && c < C) {
    const double n = (double)rows;
    const double m = __ldcg(sums + c) / n;
    double var = __ldcg(sums + C + c) / n - m * m;
    if (var < 0.0) var = 0.0;
    const double is = 1.0 / sqrt(var + (double)eps);
    mean[c] = (float)m;
    invstd[c] = (float)is;
    const float sc = gamma[c] * (float)is;
    scale[c] = sc;
    shift[c] = beta[c] - (float)m * sc;
    if (running_mean) {
      const double unb = rows > 1 ? var * n / (n - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 storage kernels: EIGHT channels per thread, so that a bfloat16 tensor is read / written with 16-byte accesses
// like the fp32 kernels above (with 4 channels = 8 bytes per access these passes are bound by the number of memory
// requests in flight, not by bytes: halving the bytes alone changed nothing, measured).  x / r / act_out / y are
// bfloat16; dy and dz_out are float32 or bfloat16 (template flags).
// ---------------------------------------------------------------------------------------------------------------
template <bool B16>
__device__ __forceinline__ void ldraw8(const void* p, size_t e, uint4& a, uint4& b) {
  if (B16) {
    a = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + e);
  } else {
    a = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p) + e);
    b = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p) + e + 4);
  }
}
template <bool B16>
__device__ __forceinline__ void unpack8(const uint4& a, const uint4& b, float (&f)[8]) {
  if (B16) {
    f[0] = __uint_as_float(a.x << 16); f[1] = __uint_as_float(a.x & 0xFFFF0000u);
    f[2] = __uint_as_float(a.y << 16); f[3] = __uint_as_float(a.y & 0xFFFF0000u);
    f[4] = __uint_as_float(a.z << 16); f[5] = __uint_as_float(a.z & 0xFFFF0000u);
    f[6] = __uint_as_float(a.w << 16); f[7] = __uint_as_float(a.w & 0xFFFF0000u);
  } else {
    f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
    f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <bool B16>
__device__ __forceinline__ void st8(void* p, size_t e, const float (&f)[8]) {
  if (B16) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p) + e) = u;
  } else {
    st4(reinterpret_cast<float*>(p) + e, make_float4(f[0], f[1], f[2], f[3]));
    st4(reinterpret_cast<float*>(p) + e + 4, make_float4(f[4], f[5], f[6], f[7]));
  }
}
__device__ __forceinline__ void ldc8(const float* p, int c, float (&f)[8]) {       // per-channel parameters (L1 / L2 resident)
  const float4 a = ld4(p + c), b = ld4(p + c + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Thread layout of the kernels below: a thread owns the SAME 8 channels for the whole launch (column lane_c of `cols`
// = min(C / 8, 256) columns per block, grid.y walks column groups) and strides over rows, so every per-channel
// coefficient is loaded once into registers and the row loop is loads + a few FMAs + one 16-byte store.
struct Cols8 {
  int c, lane_r, rpb;
  __device__ __forceinline__ Cols8(int cols) {
    c = (blockIdx.y * cols + threadIdx.x % cols) * 8;
    lane_r = threadIdx.x / cols;
    rpb = 256 / cols;
  }
};
static inline bool cols8_geometry(int C, int64_t rows, int rows_per_iter, int& cols, dim3& grid, int waves = 8) {
  if (C % 8) return false;
  const int c8 = C / 8;
  cols = c8 >= 256 ? 256 : c8;
  if ((cols & (cols - 1)) != 0 || c8 % cols != 0) return false;
  const int rpb = 256 / cols, gy = c8 / cols;
  const int64_t want = (rows + (int64_t)rpb * rows_per_iter - 1) / ((int64_t)rpb * rows_per_iter);
  int64_t cap = (int64_t)num_sms() * waves / gy;
  if (cap < 1) cap = 1;
  grid = dim3((unsigned)(want < 1 ? 1 : (want > cap ? cap : want)), gy);
  return true;
}

// XB = storage of the forward activations (x, r, act_out and the kernel's own output): true = bfloat16 ("bf16 storage",
// the tensor is its own GEMM operand image), false = float32 (fp32 engine: the output is the fp32 tensor `y` and / or its
// TF32 remainder image `y_lo`).
template <bool XB>
__device__ __forceinline__ void st_act8(void* y, float* y_lo, size_t e, const float (&v)[8]) {
  if (XB) {
    st8<true>(y, e, v);
  } else {
    if (y) st8<false>(y, e, v);
    if (y_lo) {
      st4(y_lo + e, tf32_lo4(make_float4(v[0], v[1], v[2], v[3])));
      st4(y_lo + e + 4, tf32_lo4(make_float4(v[4], v[5], v[6], v[7])));
    }
  }
}

// y = relu?(x * scale + shift [+ r * rscale + rshift | + r]): BatchNorm+ReLU image of a conv output (r = NULL) and the
// bottleneck tail
template <bool XB>
__global__ void __launch_bounds__(256) bn_apply8_kernel(const void* __restrict__ x, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const void* __restrict__ r,
                                                        const float* __restrict__ rscale, const float* __restrict__ rshift,
                                                        int relu, void* __restrict__ y, float* __restrict__ y_lo,
                                                        int64_t rows, int C, int cols) {
  const Cols8 t(cols);
  float sc[8], sh[8], rs[8], rh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = 1.f; sh[k] = 0.f; rs[k] = 1.f; rh[k] = 0.f; }
  if (scale) { ldc8(scale, t.c, sc); ldc8(shift, t.c, sh); }
  if (rscale) { ldc8(rscale, t.c, rs); ldc8(rshift, t.c, rh); }
#ifndef ZSG_APPLY_U
#define ZSG_APPLY_U 2      // measured (tools/time_bn.py): 4 rows in flight help the 369 MB tensors (6.3 vs 6.0 TB/s), 2 the smaller ones
#endif
  constexpr int U = XB ? 4 : ZSG_APPLY_U;
  const int64_t stride = (int64_t)gridDim.x * t.rpb;
  for (int64_t r0 = (int64_t)blockIdx.x * t.rpb + t.lane_r; r0 < rows; r0 += U * stride) {
    uint4 xa[U], xb2[U], ra[U], rb2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = r0 + u * stride;
      if (row < rows) {
        ldraw8<XB>(x, (size_t)row * C + t.c, xa[u], xb2[u]);
        if (r) ldraw8<XB>(r, (size_t)row * C + t.c, ra[u], rb2[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = r0 + u * stride;
      if (row >= rows) break;
      float v[8], q[8];
      unpack8<XB>(xa[u], xb2[u], v);
      if (scale) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
      }
      if (r) {
        unpack8<XB>(ra[u], rb2[u], q);
        if (rscale) {
#pragma unroll
          for (int k = 0; k < 8; ++k) q[k] = fmaf(q[k], rs[k], rh[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += q[k];
      }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
      }
      st_act8<XB>(y, y_lo, (size_t)row * C + t.c, v);
    }
  }
}

// BatchNorm backward reduce: sums[0:C] = sum dz, sums[C:2C] = sum dz * xhat with dz = dy * relu-mask.
// fp32 tensors: xhat = (x - mean) * invstd per element, as the 4-channel kernel.  bf16 storage: accumulates sum dz * x and
// centres once at the end, sum dz * xhat = invstd * (sum dz * x - mean * sum dz), in fp64.
template <bool XB, bool GB, bool ZB>
__global__ void __launch_bounds__(256) bn_bwd_reduce8_kernel(
    const void* __restrict__ x, const void* __restrict__ dy, const void* __restrict__ act_out,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ scale,
    const float* __restrict__ shift, int mask_mode, void* __restrict__ dz_out, double* __restrict__ sums, int64_t rows,
    int C, int cols) {
  const Cols8 t(cols);
  const int c = t.c;
  double a1[8], a2[8];
  float sc[8], sh[8], mu[8], is[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a1[k] = 0.0; a2[k] = 0.0; sc[k] = 0.f; sh[k] = 0.f; mu[k] = 0.f; is[k] = 1.f; }
  if (mask_mode == 1) { ldc8(scale, c, sc); ldc8(shift, c, sh); }
  if (!XB) { ldc8(mean, c, mu); ldc8(invstd, c, is); }
  const int64_t stride = (int64_t)gridDim.x * t.rpb;
  constexpr int U = (XB && GB) ? 4 : 2;
  for (int64_t r = (int64_t)blockIdx.x * t.rpb + t.lane_r; r < rows; r += U * stride) {
    uint4 xa[U], xb2[U], ga[U], gb2[U], aa[U], ab2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * stride;
      if (rr < rows) {
        const size_t off = (size_t)rr * C + c;
        ldraw8<XB>(x, off, xa[u], xb2[u]);
        ldraw8<GB>(dy, off, ga[u], gb2[u]);
        if (mask_mode == 2) ldraw8<XB>(act_out, off, aa[u], ab2[u]);
      }
    }
    float f1[8], f2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { f1[k] = 0.f; f2[k] = 0.f; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u * stride >= rows) break;
      float v[8], g[8];
      unpack8<XB>(xa[u], xb2[u], v);
      unpack8<GB>(ga[u], gb2[u], g);
      if (mask_mode == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = fmaf(v[k], sc[k], sh[k]) > 0.f ? g[k] : 0.f;
      } else if (mask_mode == 2) {
        float a[8];
        unpack8<XB>(aa[u], ab2[u], a);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = a[k] > 0.f ? g[k] : 0.f;
        if (dz_out) st8<ZB>(dz_out, (size_t)(r + u * stride) * C + c, g);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f1[k] += g[k];
        f2[k] = XB ? fmaf(g[k], v[k], f2[k]) : fmaf(g[k], (v[k] - mu[k]) * is[k], f2[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { a1[k] += (double)f1[k]; a2[k] += (double)f2[k]; }
  }
  __shared__ double sm[2][256][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sm[0][threadIdx.x][k] = a1[k]; sm[1][threadIdx.x][k] = a2[k]; }
  __syncthreads();
  if (t.lane_r == 0) {
    const int lane_c = threadIdx.x % cols;
    for (int j = 1; j < t.rpb; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) { a1[k] += sm[0][j * cols + lane_c][k]; a2[k] += sm[1][j * cols + lane_c][k]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&sums[c + k], a1[k]);
      atomicAdd(&sums[C + c + k], XB ? ((double)invstd[c + k]) * (a2[k] - (double)mean[c + k] * a1[k]) : a2[k]);
    }
  }
}

// dx = gamma * invstd * (dz - s1/n - xhat * s2/n) = A * (dz - s1/n) + Bc * (x - mean), per-channel A, Bc in registers
// (bf16 storage: folded further into A * dz + Bc * x + Cc)
template <bool XB, bool GB>
__global__ void __launch_bounds__(256) bn_bwd_apply8_kernel(
    const void* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ scale,
    const float* __restrict__ shift, const void* __restrict__ act_out, int mask_mode,
    const double* __restrict__ sums, void* __restrict__ dx, float* __restrict__ dx_lo, float* __restrict__ dgamma,
    float* __restrict__ dbeta, int64_t rows, int C, int cols) {
  const Cols8 t(cols);
  const int c = t.c;
  if (blockIdx.x == 0 && t.lane_r == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (dbeta) dbeta[c + k] = (float)sums[c + k];
      if (dgamma) dgamma[c + k] = (float)sums[C + c + k];
    }
  }
  float A[8], Bc[8], Cc[8], mu[8], sc[8], sh[8];
  {
    const float inv_n = 1.0f / (float)rows;
    float is[8], gam[8];
    ldc8(mean, c, mu); ldc8(invstd, c, is); ldc8(gamma, c, gam);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float s1 = (float)sums[c + k] * inv_n, s2 = (float)sums[C + c + k] * inv_n;
      A[k] = gam[k] * is[k];
      Bc[k] = -A[k] * is[k] * s2;
      Cc[k] = XB ? -A[k] * s1 - Bc[k] * mu[k] : s1;          // fp32: Cc holds s1, the centred form is kept
      sc[k] = 0.f; sh[k] = 0.f;
    }
    if (mask_mode == 1) { ldc8(scale, c, sc); ldc8(shift, c, sh); }
  }
  constexpr int U = (XB && GB) ? 4 : 2;
  const int64_t stride = (int64_t)gridDim.x * t.rpb;
  for (int64_t r = (int64_t)blockIdx.x * t.rpb + t.lane_r; r < rows; r += U * stride) {
    uint4 xa[U], xb2[U], ga[U], gb2[U], aa[U], ab2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * stride;
      if (rr < rows) {
        const size_t off = (size_t)rr * C + c;
        ldraw8<GB>(dy, off, ga[u], gb2[u]);
        ldraw8<XB>(x, off, xa[u], xb2[u]);
        if (mask_mode == 2) ldraw8<XB>(act_out, off, aa[u], ab2[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * stride;
      if (rr >= rows) break;
      float v[8], g[8], o[8];
      unpack8<XB>(xa[u], xb2[u], v);
      unpack8<GB>(ga[u], gb2[u], g);
      if (mask_mode == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = fmaf(v[k], sc[k], sh[k]) > 0.f ? g[k] : 0.f;
      } else if (mask_mode == 2) {
        float a[8];
        unpack8<XB>(aa[u], ab2[u], a);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = a[k] > 0.f ? g[k] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        o[k] = XB ? fmaf(A[k], g[k], fmaf(Bc[k], v[k], Cc[k])) : fmaf(Bc[k], v[k] - mu[k], A[k] * (g[k] - Cc[k]));
      st_act8<XB>(dx, dx_lo, (size_t)rr * C + c, o);
    }
  }
}

// ------------------------------------- stem max-pool ----------------------------------------
// forward also records, per output element, which tap of the 3x3 window won (first maximum in row-major scan
// order over the valid taps, like ATen's max_pool2d): code = dy*3 + dx.  Backward is then a pure gather.
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const void* __restrict__ x, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, void* __restrict__ y,
                                                          uint8_t* __restrict__ argmax, int B, int H, int W, int C,
                                                          int Ho, int Wo, int tf) {
  const bool xb = tf & TF_ACT, yb = tf & TF_Y;
  const int c4 = C / 4;
  const int64_t n = (int64_t)B * Ho * Wo * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    int64_t t = i / c4;
    const int q = (int)(t % Wo); t /= Wo;
    const int p = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const float4 sc = ld4(scale + c), sh = ld4(shift + c);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 code = make_uchar4(0, 0, 0, 0);
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = 2 * p - 1 + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = 2 * q - 1 + dx;
        if (xx < 0 || xx >= W) continue;
        const float4 v = relu4(fma4(ldx4(x, xb, (((size_t)b * H + yy) * W + xx) * C + c), sc, sh));
        const unsigned char k = (unsigned char)(dy * 3 + dx);
        if (v.x > m.x) { m.x = v.x; code.x = k; }
        if (v.y > m.y) { m.y = v.y; code.y = k; }
        if (v.z > m.z) { m.z = v.z; code.z = k; }
        if (v.w > m.w) { m.w = v.w; code.w = k; }
      }
    }
    stx4(y, yb, i * 4, m);
    if (argmax) *reinterpret_cast<uchar4*>(argmax + i * 4) = code;
  }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const uint8_t* __restrict__ argmax,
                                                          const float* __restrict__ dy, float* __restrict__ da, int B,
                                                          int H, int W, int C, int Ho, int Wo) {
  const int c4 = C / 4;
  const int64_t n = (int64_t)B * H * W * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    int64_t t = i / c4;
    const int xx = (int)(t % W); t /= W;
    const int yy = (int)(t % H);
    const int b = (int)(t / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int p0 = yy / 2, p1 = (yy + 1) / 2, q0 = xx / 2, q1 = (xx + 1) / 2;     // windows containing (yy, xx)
    for (int p = p0; p <= p1; ++p) {
      if (p >= Ho) continue;
      for (int q = q0; q <= q1; ++q) {
        if (q >= Wo) continue;
        const unsigned char k = (unsigned char)((yy - (2 * p - 1)) * 3 + (xx - (2 * q - 1)));
        const size_t o = (((size_t)b * Ho + p) * Wo + q) * C + c;
        const uchar4 code = *reinterpret_cast<const uchar4*>(argmax + o);
        const float4 g = ld4(dy + o);
        if (code.x == k) acc.x += g.x;
        if (code.y == k) acc.y += g.y;
        if (code.z == k) acc.z += g.z;
        if (code.w == k) acc.w += g.w;
      }
    }
    st4(da + i * 4, acc);
  }
}

// ----------------------------- SSD-VGG trunk glue (ssd_vgg.py) -------------------------------
// Generic max-pool (k x k window, stride, symmetric padding, output size given by the caller so that floor and
// ceil mode are the same kernel).  The code of the winning tap (first maximum in row-major scan order over the
// taps inside the image, as ATen's max_pool2d) is recorded per element; the backward is a gather over the
// windows that contain an input pixel, optionally masked by the ReLU that feeds the pool.
__global__ void __launch_bounds__(256) pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       uint8_t* __restrict__ argmax, int B, int H, int W, int C, int K,
                                                       int S, int P, int Ho, int Wo) {
  const int c4 = C / 4;
  const int64_t n = (int64_t)B * Ho * Wo * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    int64_t t = i / c4;
    const int q = (int)(t % Wo); t /= Wo;
    const int p = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 code = make_uchar4(255, 255, 255, 255);
    for (int dy = 0; dy < K; ++dy) {
      const int yy = p * S - P + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = 0; dx < K; ++dx) {
        const int xx = q * S - P + dx;
        if (xx < 0 || xx >= W) continue;
        const float4 v = ld4(x + (((size_t)b * H + yy) * W + xx) * C + c);
        const unsigned char k = (unsigned char)(dy * K + dx);
        if (v.x > m.x || code.x == 255) { m.x = v.x; code.x = k; }
        if (v.y > m.y || code.y == 255) { m.y = v.y; code.y = k; }
        if (v.z > m.z || code.z == 255) { m.z = v.z; code.z = k; }
        if (v.w > m.w || code.w == 255) { m.w = v.w; code.w = k; }
      }
    }
    st4(y + i * 4, m);
    if (argmax) *reinterpret_cast<uchar4*>(argmax + i * 4) = code;
  }
}

__global__ void __launch_bounds__(256) pool_bwd_kernel(const uint8_t* __restrict__ argmax, const float* __restrict__ dy,
                                                       const float* __restrict__ mask, float* __restrict__ dx, int B,
                                                       int H, int W, int C, int K, int S, int P, int Ho, int Wo) {
  const int c4 = C / 4;
  const int64_t n = (int64_t)B * H * W * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    int64_t t = i / c4;
    const int xx = (int)(t % W); t /= W;
    const int yy = (int)(t % H);
    const int b = (int)(t / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // windows p with p*S - P <= yy <= p*S - P + K - 1
    int p0 = yy + P - K + 1; p0 = p0 <= 0 ? 0 : (p0 + S - 1) / S;
    int q0 = xx + P - K + 1; q0 = q0 <= 0 ? 0 : (q0 + S - 1) / S;
    const int p1 = min((yy + P) / S, Ho - 1), q1 = min((xx + P) / S, Wo - 1);
    for (int p = p0; p <= p1; ++p)
      for (int q = q0; q <= q1; ++q) {
        const unsigned char k = (unsigned char)((yy - (p * S - P)) * K + (xx - (q * S - P)));
        const size_t o = (((size_t)b * Ho + p) * Wo + q) * C + c;
        const uchar4 code = *reinterpret_cast<const uchar4*>(argmax + o);
        const float4 g = ld4(dy + o);
        if (code.x == k) acc.x += g.x;
        if (code.y == k) acc.y += g.y;
        if (code.z == k) acc.z += g.z;
        if (code.w == k) acc.w += g.w;
      }
    if (mask) {
      const float4 v = ld4(mask + i * 4);
      if (!(v.x > 0.f)) acc.x = 0.f;
      if (!(v.y > 0.f)) acc.y = 0.f;
      if (!(v.z > 0.f)) acc.z = 0.f;
      if (!(v.w > 0.f)) acc.w = 0.f;
    }
    st4(dx + i * 4, acc);
  }
}

// s = x / ||x||_2 over channels (ssd_vgg.py:80): one warp per pixel row, the row stays in registers between the
// reduction and the division (C <= 1024).  The norm is kept for the backward.
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                         float* __restrict__ norm, int64_t rows, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarp) {
    float4 v[8];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j * 32 + lane) * 4;
      v[j] = c < C ? ld4(x + r * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    const float nrm = sqrtf(warp_sum(ss));
    if (lane == 0) norm[r] = nrm;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) st4(y + r * C + c, make_float4(v[j].x / nrm, v[j].y / nrm, v[j].z / nrm, v[j].w / nrm));
    }
  }
}

// dx (+)= [x > 0] * (dy / n - x * sum_c(dy * x) / n^3): the autograd of x / x.norm(dim=1, keepdim=True), with the
// ReLU that produced x (vgg[22]) folded in when mask_relu is set.
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         const float* __restrict__ norm, float* __restrict__ dx,
                                                         int64_t rows, int C, int accumulate, int mask_relu) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarp) {
    float4 v[8], g[8];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j * 32 + lane) * 4;
      v[j] = g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C) { v[j] = ld4(x + r * C + c); g[j] = ld4(dy + r * C + c); }
      dot += v[j].x * g[j].x + v[j].y * g[j].y + v[j].z * g[j].z + v[j].w * g[j].w;
    }
    dot = warp_sum(dot);
    const float n = norm[r];
    const float k = dot / (n * n * n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c >= C) continue;
      float4 o = make_float4(g[j].x / n - v[j].x * k, g[j].y / n - v[j].y * k, g[j].z / n - v[j].z * k,
                             g[j].w / n - v[j].w * k);
      if (mask_relu) {
        if (!(v[j].x > 0.f)) o.x = 0.f;
        if (!(v[j].y > 0.f)) o.y = 0.f;
        if (!(v[j].z > 0.f)) o.z = 0.f;
        if (!(v[j].w > 0.f)) o.w = 0.f;
      }
      if (accumulate) { const float4 a = ld4(dx + r * C + c); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
      st4(dx + r * C + c, o);
    }
  }
}

// --------------------------------- FPN nearest upsample + add --------------------------------
__global__ void __launch_bounds__(256) upsample_add_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                           const int32_t* __restrict__ iy, const int32_t* __restrict__ ix,
                                                           int B, int Ho, int Wo, int Hi, int Wi, int C) {
  const int c4 = C / 4;
  const int64_t n = (int64_t)B * Ho * Wo * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    int64_t t = i / c4;
    const int x = (int)(t % Wo); t /= Wo;
    const int y = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float4 d = ld4(dst + i * 4);
    const float4 s = ld4(src + (((size_t)b * Hi + iy[y]) * Wi + ix[x]) * C + c);
    d.x += s.x; d.y += s.y; d.z += s.z; d.w += s.w;
    st4(dst + i * 4, d);
  }
}

__global__ void __launch_bounds__(256) upsample_add_bwd_kernel(const float* __restrict__ ddst, float* __restrict__ dsrc,
                                                               const int32_t* __restrict__ iy,
                                                               const int32_t* __restrict__ ix, int B, int Ho, int Wo,
                                                               int Hi, int Wi, int C) {
  const int c4 = C / 4;
  const int64_t n = (int64_t)B * Hi * Wi * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    int64_t t = i / c4;
    const int xs = (int)(t % Wi); t /= Wi;
    const int ys = (int)(t % Hi);
    const int b = (int)(t / Hi);
    float4 acc = ld4(dsrc + i * 4);
    for (int y = 0; y < Ho; ++y) {
      if (iy[y] != ys) continue;
      for (int x = 0; x < Wo; ++x) {
        if (ix[x] != xs) continue;
        const float4 g = ld4(ddst + (((size_t)b * Ho + y) * Wo + x) * C + c);
        acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
      }
    }
    st4(dsrc + i * 4, acc);
  }
}

__global__ void avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int HW, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  int b = i / C, c = i % C;
  float s = 0.f;
  for (int k = 0; k < HW; ++k) s += x[((size_t)b * HW + k) * C + c];
  y[i] = s / (float)HW;
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int HW, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW * C) return;
  int c = i % C, b = i / (HW * C);
  dx[i] += dy[b * C + c] / (float)HW;
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                       float* __restrict__ dx, int64_t n4, int accumulate) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 g = ld4(dy + i * 4), v = ld4(x + i * 4);
    g.x = v.x > 0.f ? g.x : 0.f; g.y = v.y > 0.f ? g.y : 0.f;
    g.z = v.z > 0.f ? g.z : 0.f; g.w = v.w > 0.f ? g.w : 0.f;
    if (accumulate) { float4 o = ld4(dx + i * 4); g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w; }
    st4(dx + i * 4, g);
  }
}

__global__ void __launch_bounds__(256) axpy_kernel(const float* __restrict__ x, float* __restrict__ y, float a,
                                                   int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaf(a, x[i], y[i]);
}

__global__ void __launch_bounds__(256) scale_dev_kernel(float* __restrict__ x, int64_t rows, int width, int64_t stride,
                                                        const double* __restrict__ scale) {
  const float s = (float)*scale;
  const int64_t n = rows * width;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[(i / width) * stride + (i % width)] *= s;
}

// ------------------------------ language / grid tiling fusion --------------------------------
struct Levels { int cells[8]; int n; };

__device__ __forceinline__ void locate_row(const Levels& lv, int B, int64_t row, int& lvl, int& b, int& cell,
                                           int& cell_base) {
  int64_t off = 0;
  int cb = 0;
  lvl = 0;
  for (int l = 0; l < lv.n; ++l) {
    const int64_t cnt = (int64_t)B * lv.cells[l];
    if (row < off + cnt) { lvl = l; break; }
    off += cnt;
    cb += lv.cells[l];
  }
  const int64_t rr = row - off;
  b = (int)(rr / lv.cells[lvl]);
  cell = (int)(rr % lv.cells[lvl]);
  cell_base = cb;
}

__global__ void __launch_bounds__(256) fuse_kernel(const float* __restrict__ feat, const float* __restrict__ lang,
                                                   const float* __restrict__ grid_yx, float* __restrict__ fused, int B,
                                                   int total_cells, Levels lv, int cfeat, int clang, int cpad) {
  const int p4 = cpad / 4;
  const int64_t n = (int64_t)B * total_cells * p4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % p4) * 4;
    const int64_t row = i / p4;
    float4 v;
    if (c < cfeat) {
      v = ld4(feat + row * cfeat + c);
    } else {
      int lvl, b, cell, cb;
      locate_row(lv, B, row, lvl, b, cell, cb);
      if (c < cfeat + clang) {
        v = ld4(lang + (size_t)b * clang + (c - cfeat));
      } else if (c == cfeat + clang) {
        v = make_float4(grid_yx[2 * (cb + cell)], grid_yx[2 * (cb + cell) + 1], 0.f, 0.f);
      } else {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    st4(fused + i * 4, v);
  }
}

__global__ void __launch_bounds__(256) unfuse_feat_kernel(const float* __restrict__ dfused, float* __restrict__ dfeat,
                                                          int64_t rows, int cfeat, int cpad) {
  const int f4 = cfeat / 4;
  const int64_t n = rows * f4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % f4) * 4;
    const int64_t row = i / f4;
    st4(dfeat + i * 4, ld4(dfused + row * cpad + c));
  }
}

// one block per (sample, level, chunk of 64 cells): dlang[b][j] += sum over the chunk's cells (four rows in flight per
// thread; one block per whole level walked 1444 cells with a single load in flight: 181 us for 139 MB)
constexpr int UNFUSE_CHUNK = 64;
__global__ void __launch_bounds__(256) unfuse_lang_kernel(const float* __restrict__ dfused, float* __restrict__ dlang,
                                                          int B, Levels lv, int cfeat, int clang, int cpad) {
  const int b = blockIdx.x, l = blockIdx.y;
  const int c0 = blockIdx.z * UNFUSE_CHUNK;
  const int ncell = lv.cells[l];
  if (c0 >= ncell) return;
  const int c1 = min(c0 + UNFUSE_CHUNK, ncell);
  int64_t off = 0;
  for (int k = 0; k < l; ++k) off += (int64_t)B * lv.cells[k];
  const float* base = dfused + (off + (int64_t)b * ncell) * cpad + cfeat;
  for (int j = threadIdx.x; j < clang; j += blockDim.x) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int cell = c0;
    for (; cell + 3 < c1; cell += 4) {
      s0 += base[(size_t)cell * cpad + j];
      s1 += base[(size_t)(cell + 1) * cpad + j];
      s2 += base[(size_t)(cell + 2) * cpad + j];
      s3 += base[(size_t)(cell + 3) * cpad + j];
    }
    for (; cell < c1; ++cell) s0 += base[(size_t)cell * cpad + j];
    atomicAdd(&dlang[(size_t)b * clang + j], (s0 + s1) + (s2 + s3));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// First head conv WITHOUT the concatenated tensor (mdl.py:69-104 builds [feat | lang tiled | grid] and convolves it).
// conv(W, [feat|lang|grid]) = conv(W_f, feat) + L[b, class(cell)] + G[cell]:
//   V[b, (n,t)]   = sum_c W_l[n,t,c] * lang[b,c]                       a [B,256] x [256,2304] GEMM (zsg_conv_fwd)
//   L[b, cls, n]  = sum over the taps t that fall inside the level at a cell of border class cls of V[b,(n,t)]
//   G[cell, n]    = sum_t sum_g W_g[n,t,g] * gridpatch[cell, t, g]     batch independent
// and the conv over feat adds L + G per output row in its epilogue (zsg_conv_params.row_add).  Border class of a cell
// (y, x) of an S x S level: ry = (y == 0) | (y == S-1) << 1, rx likewise, cls = ry * 4 + rx; tap (r, s) is valid iff
// !(r == 0 && ry & 1) && !(r == 2 && ry & 2) and the same in x.  gridpatch[cell][t][g] = grid value g of the neighbour
// cell under tap t, 0 outside the level (host-built table, like the row tables).
// Backward: S_t[b,(n,t)] = sum over the cells where t is valid of dh0[b,cell,n] gives d lang = S_t x W_l (GEMM) and
// dW_l = S_t^T x lang (weight-gradient GEMM); dW_g[n,t,g] = sum_{b,cell} dh0[b,cell,n] * gridpatch[cell,t,g].
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tap_valid(int cls, int t) {
  const int r = t / 3, q = t % 3, ry = cls >> 2, rx = cls & 3;
  return !(r == 0 && (ry & 1)) && !(r == 2 && (ry & 2)) && !(q == 0 && (rx & 1)) && !(q == 2 && (rx & 2));
}

// dst[row][0:c] = src[row][0:c] with separate row pitches (channel slices of [.., 514] weights and their gradients)
__global__ void __launch_bounds__(256) copy_cols_kernel(const float* __restrict__ src, int64_t src_ld, float* __restrict__ dst,
                                                        int64_t dst_ld, int64_t rows, int c) {
  const int64_t n = rows * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c;
    const int k = (int)(i % c);
    dst[r * dst_ld + k] = src[r * src_ld + k];
  }
}

// L[b][cls][n] from V[b][n * 9 + t]; grid (B, 16), 256 threads = n
__global__ void __launch_bounds__(256) lang_class_kernel(const float* __restrict__ V, float* __restrict__ L, int N) {
  const int b = blockIdx.x, cls = blockIdx.y;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* v = V + ((size_t)b * N + n) * 9;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t)
      if (tap_valid(cls, t)) acc += v[t];
    L[((size_t)b * 16 + cls) * N + n] = acc;
  }
}

// G[cell][n] = sum_{t,g} Wg[(n * 9 + t) * 2 + g] * gp[cell][t * 2 + g]; one block per cell
__global__ void __launch_bounds__(256) grid_term_kernel(const float* __restrict__ Wg, const float* __restrict__ gp,
                                                        float* __restrict__ G, int N) {
  const int cell = blockIdx.x;
  float p[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) p[k] = gp[(size_t)cell * 18 + k];
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* w = Wg + (size_t)n * 18;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 18; ++k) acc = fmaf(w[k], p[k], acc);
    G[(size_t)cell * N + n] = acc;
  }
}

// Backward reduce over the rows of dh0 (level-major [sum_l B * cells_l, N], N = 256 = blockDim.x): block (b, chunk) walks
// the cells chunk, chunk + nch, ... of sample b; per border class the column sums (shared memory, a thread only ever
// touches its own column) and per (tap, grid channel) the sums weighted with the grid patch (registers).
//   Sp[b][chunk][16][N], Wp[b][chunk][N][18]   partial sums, added up in fixed order by head0_finish_kernel
__global__ void __launch_bounds__(256) head0_reduce_kernel(const float* __restrict__ dh, const int32_t* __restrict__ cell_base,
                                                           const int32_t* __restrict__ cell_stride,
                                                           const int32_t* __restrict__ cell_cls, const float* __restrict__ gp,
                                                           int total_cells, int N, float* __restrict__ Sp,
                                                           float* __restrict__ Wp) {
  __shared__ float s[16][256];
  const int b = blockIdx.x, ch = blockIdx.y, nch = gridDim.y, n = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 16; ++k) s[k][n] = 0.f;
  float wg[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) wg[k] = 0.f;
  constexpr int U = 4;
  for (int c0 = ch; c0 < total_cells; c0 += nch * U) {
    float v[U];
    int cl[U], ce[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ce[u] = c0 + u * nch;
      v[u] = 0.f;
      cl[u] = 0;
      if (ce[u] < total_cells) {
        cl[u] = cell_cls[ce[u]];
        v[u] = dh[((size_t)cell_base[ce[u]] + (size_t)b * cell_stride[ce[u]]) * N + n];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ce[u] >= total_cells) break;
      s[cl[u]][n] += v[u];
      const float2* g = reinterpret_cast<const float2*>(gp + (size_t)ce[u] * 18);     // 72-byte rows: 8-byte aligned
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float2 gk = __ldg(g + k);
        wg[2 * k] = fmaf(v[u], gk.x, wg[2 * k]);
        wg[2 * k + 1] = fmaf(v[u], gk.y, wg[2 * k + 1]);
      }
    }
  }
  float* sp = Sp + ((size_t)b * nch + ch) * 16 * N;
#pragma unroll
  for (int k = 0; k < 16; ++k) sp[(size_t)k * N + n] = s[k][n];
  float* wp = Wp + (((size_t)b * nch + ch) * N + n) * 18;
#pragma unroll
  for (int k = 0; k < 18; ++k) wp[k] = wg[k];
}

// St[b][n * 9 + t] = sum_chunk sum_{cls: t valid} Sp ; dWg (row pitch ld, starting at column 0 of dwg) = sum_b sum_chunk Wp
__global__ void __launch_bounds__(256) head0_finish_kernel(const float* __restrict__ Sp, const float* __restrict__ Wp, int B,
                                                           int nch, int N, float* __restrict__ St, float* __restrict__ dwg,
                                                           int64_t dwg_ld) {
  const int n = threadIdx.x;
  if ((int)blockIdx.x < B) {
    const int b = blockIdx.x;
    float cls_sum[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a = 0.f;
      for (int c = 0; c < nch; ++c) a += Sp[(((size_t)b * nch + c) * 16 + k) * N + n];
      cls_sum[k] = a;
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (tap_valid(k, t)) a += cls_sum[k];
      St[((size_t)b * N + n) * 9 + t] = a;
    }
  } else {
    // blocks B .. B + 17: one (tap, grid channel) each
    const int k = blockIdx.x - B;
    float a = 0.f;
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < nch; ++c) a += Wp[(((size_t)b * nch + c) * N + n) * 18 + k];
    dwg[((size_t)n * 9 + k / 2) * dwg_ld + (k & 1)] = a;
  }
}

// ------------------------------------ layout helpers -----------------------------------------
__global__ void __launch_bounds__(256) weight_transpose_flip_kernel(const float* __restrict__ w, float* __restrict__ wt,
                                                                    int Cout, int R, int S, int Cin) {
  const int64_t n = (int64_t)Cout * R * S * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes wt [cin][r][s][cout]
    const int k = (int)(i % Cout);
    int64_t t = i / Cout;
    const int s = (int)(t % S); t /= S;
    const int r = (int)(t % R);
    const int c = (int)(t / R);
    wt[i] = w[(((size_t)k * R + (R - 1 - r)) * S + (S - 1 - s)) * Cin + c];
  }
}

__global__ void __launch_bounds__(256) weight_transpose_flip_batched_kernel(const float* __restrict__ src,
                                                                            float* __restrict__ dst,
                                                                            const zsg_wtf_desc* __restrict__ descs, int n,
                                                                            int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;                                // last entry with begin <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].begin <= i) lo = mid; else hi = mid - 1;
    }
    const zsg_wtf_desc d = descs[lo];
    const int64_t j = i - d.begin;                         // index into wt [cin][r][s][cout]
    const int k = (int)(j % d.cout);
    int64_t t = j / d.cout;
    const int s_ = (int)(t % d.s); t /= d.s;
    const int r_ = (int)(t % d.r);
    const int c = (int)(t / d.r);
    dst[d.dst + j] = src[d.src + (((int64_t)k * d.r + (d.r - 1 - r_)) * d.s + (d.s - 1 - s_)) * d.cin + c];
  }
}

// The same through 32 x 32 shared-memory tiles (every cin / cout of the batched copies is a multiple of 32): reads run along
// cin, writes along cout, both coalesced.  The element-wise kernel above reads 32 different lines per warp load: 0.88 TB/s
// for 280 MB = 342 us per backward; this one is bound by the copy itself.
__global__ void __launch_bounds__(256) weight_transpose_flip_tiled_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                          const zsg_wtf_desc* __restrict__ descs, int n,
                                                                          int64_t total_tiles) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    int lo = 0, hi = n - 1;                                // last entry with begin <= t * 1024
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].begin <= t * 1024) lo = mid; else hi = mid - 1;
    }
    const zsg_wtf_desc d = descs[lo];
    int64_t local = t - d.begin / 1024;
    const int nct = d.cin / 32, nkt = d.cout / 32;
    const int ct = (int)(local % nct); local /= nct;
    const int kt = (int)(local % nkt);
    const int tap = (int)(local / nkt);
    const int r_ = tap / d.s, s_ = tap % d.s;
    const int rr = d.r - 1 - r_, ss = d.s - 1 - s_;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = kt * 32 + ty + 8 * i, c = ct * 32 + tx;
      tile[ty + 8 * i][tx] = src[d.src + (((int64_t)k * d.r + rr) * d.s + ss) * d.cin + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = ct * 32 + ty + 8 * i, k = kt * 32 + tx;
      dst[d.dst + (((int64_t)c * d.r + r_) * d.s + s_) * d.cout + k] = tile[tx][ty + 8 * i];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi,
                                                         float* __restrict__ lo, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = w[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

__global__ void __launch_bounds__(256) pad_channels_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           int64_t n, int cs, int cd) {
  const int64_t tot = n * cd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cd);
    const int64_t r = i / cd;
    dst[i] = c < cs ? src[r * cs + c] : 0.f;
  }
}

__global__ void __launch_bounds__(256) nchw_to_nhwc4_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                            int B, int H, int W) {
  const int64_t n = (int64_t)B * H * W;
  const int64_t plane = (int64_t)H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / plane, p = i % plane;
    const float* s = img + b * 3 * plane + p;
    st4(out + i * 4, make_float4(s[0], s[plane], s[2 * plane], 0.f));
  }
}

// the same for the bf16 operand path: 8 bf16 channels per pixel (r, g, b, 0 x 5) = one 16-byte store, which is at once the
// padded NHWC image and its GEMM operand image (no fp32 copy, no cast pass)
__global__ void __launch_bounds__(256) nchw_to_nhwc8_bf16_kernel(const float* __restrict__ img, uint16_t* __restrict__ out,
                                                                 int B, int H, int W) {
  const int64_t n = (int64_t)B * H * W;
  const int64_t plane = (int64_t)H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / plane, p = i % plane;
    const float* s = img + b * 3 * plane + p;
    const __nv_bfloat162 rg = __floats2bfloat162_rn(s[0], s[plane]), b0 = __floats2bfloat162_rn(s[2 * plane], 0.f);
    uint4 u = make_uint4(*reinterpret_cast<const uint32_t*>(&rg), *reinterpret_cast<const uint32_t*>(&b0), 0u, 0u);
    *reinterpret_cast<uint4*>(out + i * 8) = u;
  }
}

// out[c] (+)= sum_rows x[row][c]; block = 32 x 8, each block owns 32 channels and a row slab
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows,
                                                     int C, int ld) {
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int64_t r = (int64_t)blockIdx.y * 8 + ry; r < rows; r += (int64_t)gridDim.y * 8) s += x[(size_t)r * ld + c];
  sm[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < C) {
    for (int j = 1; j < 8; ++j) s += sm[j][threadIdx.x & 31];
    atomicAdd(&out[c], s);
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const zsg_row_t* __restrict__ rows,
                                                          float* __restrict__ dst, int64_t m, int cs, int cd) {
  const int64_t tot = m * cd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cd);
    const int64_t r = i / cd;
    dst[i] = c < cs ? src[(int64_t)rows[r].out + c] : 0.f;
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                   float b1, float b2, float eps, float bc1, float bc2_sqrt,
                                                   float gscale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

}  // namespace zsg

using namespace zsg;

extern "C" int zsg_bn_stats(const float* x, double* sums, int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(x && sums && rows > 0, "zsg_bn_stats: bad arguments");
  return launch_channel_reduce(0, x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, sums, rows, c,
                               as_stream(stream));
}

extern "C" int zsg_bn_stats_partials(const float* partials, int64_t parts, int c, double* sums, zsg_stream_t stream) {
  ZSG_REQUIRE(partials && sums && parts > 0 && c > 0, "zsg_bn_stats_partials: bad arguments");
  const int gx = (c + 31) / 32;
  int64_t gy = (parts + 63) / 64;                         // >= 8 parts per part-lane
  const int64_t cap = (int64_t)num_sms() * 4 / gx + 1;
  if (gy > cap) gy = cap;
  bn_partials_kernel<<<dim3(gx, (unsigned)gy), 256, 0, as_stream(stream)>>>(partials, parts, c, sums, 0, nullptr, nullptr, 0.f,
                                                                            0.f, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                            nullptr, nullptr);
  return check_launch("zsg_bn_stats_partials");
}

__global__ void bn_bwd_center_sums_kernel(double* __restrict__ sums, const float* __restrict__ mean,
                                          const float* __restrict__ invstd, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) sums[C + c] = (double)invstd[c] * (sums[C + c] - (double)mean[c] * sums[c]);
}

extern "C" int zsg_bn_bwd_center_sums(double* sums, const float* mean, const float* invstd, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(sums && mean && invstd && c > 0, "zsg_bn_bwd_center_sums: bad arguments");
  bn_bwd_center_sums_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(sums, mean, invstd, c);
  return check_launch("zsg_bn_bwd_center_sums");
}

extern "C" int zsg_bn_finalize_partials(const float* partials, int64_t parts, int64_t rows, int c, const float* gamma,
                                        const float* beta, float eps, float momentum, float* running_mean,
                                        float* running_var, float* mean, float* invstd, float* scale, float* shift,
                                        double* sums, int* tickets, zsg_stream_t stream) {
  ZSG_REQUIRE(partials && sums && tickets && gamma && beta && mean && invstd && scale && shift && parts > 0 && rows > 0 &&
                  c > 0 && c <= 32 * 64,
              "zsg_bn_finalize_partials: bad arguments (c <= 2048, tickets = 64 zeroed ints)");
  const int gx = (c + 31) / 32;
  int64_t gy = (parts + 63) / 64;
  const int64_t cap = (int64_t)num_sms() * 4 / gx + 1;
  if (gy > cap) gy = cap;
  bn_partials_kernel<<<dim3(gx, (unsigned)gy), 256, 0, as_stream(stream)>>>(partials, parts, c, sums, rows, gamma, beta, eps,
                                                                            momentum, running_mean, running_var, mean, invstd,
                                                                            scale, shift, tickets);
  return check_launch("zsg_bn_finalize_partials");
}

extern "C" int zsg_bn_finalize(const double* sums, int64_t rows, int c, const float* gamma, const float* beta,
                               float eps, float momentum, float* running_mean, float* running_var, float* mean,
                               float* invstd, float* scale, float* shift, zsg_stream_t stream) {
  ZSG_REQUIRE(sums && gamma && beta && mean && invstd && scale && shift && rows > 0, "zsg_bn_finalize: bad arguments");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(sums, rows, c, gamma, beta, eps, momentum,
                                                                      running_mean, running_var, mean, invstd, scale,
                                                                      shift);
  return check_launch("zsg_bn_finalize");
}

extern "C" int zsg_bn_eval_affine(const float* running_mean, const float* running_var, const float* gamma,
                                  const float* beta, float eps, int c, float* scale, float* shift,
                                  zsg_stream_t stream) {
  ZSG_REQUIRE(running_mean && running_var && gamma && beta && scale && shift, "zsg_bn_eval_affine: null pointer");
  bn_eval_affine_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(running_mean, running_var, gamma, beta, eps, c,
                                                                         scale, shift);
  return check_launch("zsg_bn_eval_affine");
}

extern "C" int zsg_bn_apply(const float* x, const float* scale, const float* shift, const float* r,
                            const float* rscale, const float* rshift, int relu, float* y, float* y_lo, int64_t rows,
                            int c, zsg_stream_t stream) {
  ZSG_REQUIRE(x && scale && shift && y && c % 4 == 0, "zsg_bn_apply: bad arguments");
  {
    int cols;
    dim3 grid;
    if ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)r | (uintptr_t)y_lo) & 15) == 0 && cols8_geometry(c, rows, 4, cols, grid)) {
      bn_apply8_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(x, scale, shift, r, rscale, rshift, relu, y, y_lo, rows, c,
                                                                   cols);
      return check_launch("zsg_bn_apply");
    }
  }
  int64_t n4 = rows * (c / 4);
  bn_apply_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(x, scale, shift, r, rscale, rshift, relu, y, y_lo,
                                                                     nullptr, n4, c / 4, 0);
  return check_launch("zsg_bn_apply");
}

extern "C" int zsg_bn_apply_bf16(const float* x, const float* scale, const float* shift, const float* r,
                                 const float* rscale, const float* rshift, int relu, float* y, uint16_t* y_bf16,
                                 int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(x && scale && shift && y && y_bf16 && c % 4 == 0, "zsg_bn_apply_bf16: bad arguments");
  int64_t n4 = rows * (c / 4);
  bn_apply_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(x, scale, shift, r, rscale, rshift, relu, y, nullptr,
                                                                     y_bf16, n4, c / 4, 0);
  return check_launch("zsg_bn_apply_bf16");
}

extern "C" int zsg_cast_bf16(const float* x, const float* scale, const float* shift, int relu, uint16_t* out,
                             int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(x && out && c > 0 && c % 4 == 0, "zsg_cast_bf16: bad arguments");
  ZSG_REQUIRE(!scale == !shift, "zsg_cast_bf16: scale and shift go together");
  ZSG_REQUIRE((((uintptr_t)x & 15) | ((uintptr_t)out & 7)) == 0, "zsg_cast_bf16: x must be 16-byte and out 8-byte aligned");
  if (rows <= 0) return ZSG_OK;
  const int64_t n4 = rows * (c / 4);
  split_act_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(x, scale, shift, relu, nullptr, nullptr, out, n4, c / 4, 0);
  return check_launch("zsg_cast_bf16");
}

extern "C" int zsg_split_act(const float* x, const float* scale, const float* shift, int relu, float* z, float* lo,
                             int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(x && lo && c > 0 && c % 4 == 0, "zsg_split_act: bad arguments");
  ZSG_REQUIRE(!scale == !shift, "zsg_split_act: scale and shift go together");
  ZSG_REQUIRE(z || (!scale && !relu), "zsg_split_act: a prologue needs an output tensor z");
  if (rows <= 0) return ZSG_OK;
  {
    int cols;
    dim3 grid;
    if ((((uintptr_t)x | (uintptr_t)z | (uintptr_t)lo) & 15) == 0 && cols8_geometry(c, rows, 4, cols, grid)) {
      bn_apply8_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(x, scale, shift, nullptr, nullptr, nullptr, relu, z, lo, rows, c,
                                                                   cols);
      return check_launch("zsg_split_act");
    }
  }
  const int64_t n4 = rows * (c / 4);
  split_act_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(x, scale, shift, relu, z, lo, nullptr, n4, c / 4, 0);
  return check_launch("zsg_split_act");
}

extern "C" int zsg_bn_bwd_reduce(const float* dy, const float* x, const float* mean, const float* invstd,
                                 const float* scale, const float* shift, const float* act_out, int mask_mode,
                                 float* dz_out, double* sums, int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(dy && x && mean && invstd && sums, "zsg_bn_bwd_reduce: null pointer");
  ZSG_REQUIRE(mask_mode != 1 || (scale && shift), "zsg_bn_bwd_reduce: mask_mode 1 needs scale/shift");
  ZSG_REQUIRE(mask_mode != 2 || act_out, "zsg_bn_bwd_reduce: mask_mode 2 needs act_out");
  {
    int cols;
    dim3 grid;
    if ((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)act_out | (uintptr_t)dz_out) & 15) == 0 &&
        cols8_geometry(c, rows, 4, cols, grid, 2)) {
      bn_bwd_reduce8_kernel<false, false, false><<<grid, 256, 0, as_stream(stream)>>>(x, dy, act_out, mean, invstd, scale, shift,
                                                                                      mask_mode, dz_out, sums, rows, c, cols);
      return check_launch("zsg_bn_bwd_reduce");
    }
  }
  return launch_channel_reduce(1, x, dy, act_out, mean, invstd, scale, shift, mask_mode, dz_out, sums, rows, c,
                               as_stream(stream));
}

extern "C" int zsg_bn_bwd_apply(const float* dy, const float* x, const float* mean, const float* invstd,
                                const float* gamma, const float* scale, const float* shift, const float* act_out,
                                int mask_mode, const double* sums, float* dx, float* dx_lo, float* dgamma,
                                float* dbeta, int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(dy && x && mean && invstd && gamma && sums && dx && c % 4 == 0, "zsg_bn_bwd_apply: bad arguments");
  {
    int cols;
    dim3 grid;
    if ((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)act_out | (uintptr_t)dx | (uintptr_t)dx_lo) & 15) == 0 &&
        cols8_geometry(c, rows, 4, cols, grid)) {
      bn_bwd_apply8_kernel<false, false><<<grid, 256, 0, as_stream(stream)>>>(dy, x, mean, invstd, gamma, scale, shift, act_out,
                                                                              mask_mode, sums, dx, dx_lo, dgamma, dbeta, rows, c,
                                                                              cols);
      return check_launch("zsg_bn_bwd_apply");
    }
  }
  bn_bwd_apply_kernel<<<grid_for(rows * (c / 4), 256), 256, 0, as_stream(stream)>>>(
      dy, x, mean, invstd, gamma, scale, shift, act_out, mask_mode, sums, dx, dx_lo, nullptr, dgamma, dbeta, rows, c, 0);
  return check_launch("zsg_bn_bwd_apply");
}

extern "C" int zsg_bn_bwd_apply_bf16(const float* dy, const float* x, const float* mean, const float* invstd,
                                     const float* gamma, const float* scale, const float* shift, const float* act_out,
                                     int mask_mode, const double* sums, float* dx, uint16_t* dx_bf16, float* dgamma,
                                     float* dbeta, int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(dy && x && mean && invstd && gamma && sums && dx_bf16 && c % 4 == 0, "zsg_bn_bwd_apply_bf16: bad arguments");
  bn_bwd_apply_kernel<<<grid_for(rows * (c / 4), 256), 256, 0, as_stream(stream)>>>(
      dy, x, mean, invstd, gamma, scale, shift, act_out, mask_mode, sums, dx, nullptr, dx_bf16, dgamma, dbeta, rows, c, 0);
  return check_launch("zsg_bn_bwd_apply_bf16");
}

extern "C" int zsg_maxpool_bn_relu_fwd(const float* x, const float* scale, const float* shift, float* y,
                                       uint8_t* argmax, int b, int h, int w, int c, int ho, int wo,
                                       zsg_stream_t stream) {
  ZSG_REQUIRE(x && scale && shift && y && c % 4 == 0, "zsg_maxpool_bn_relu_fwd: bad arguments");
  maxpool_fwd_kernel<<<grid_for((int64_t)b * ho * wo * (c / 4), 256), 256, 0, as_stream(stream)>>>(
      x, scale, shift, y, argmax, b, h, w, c, ho, wo, 0);
  return check_launch("zsg_maxpool_bn_relu_fwd");
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 storage (trunk of the bf16 engine): conv outputs, BatchNorm+ReLU images and block outputs live in HBM as
// bfloat16 only.  Same kernels, bfloat16 loads / stores; statistics, affine and reductions stay fp32 / fp64.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int zsg_act_b16(const uint16_t* x, const float* scale, const float* shift, int relu, uint16_t* out,
                           int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(x && out && c > 0 && c % 4 == 0, "zsg_act_b16: bad arguments");
  ZSG_REQUIRE(!scale == !shift, "zsg_act_b16: scale and shift go together");
  ZSG_REQUIRE((((uintptr_t)x | (uintptr_t)out) & 7) == 0, "zsg_act_b16: x and out must be 8-byte aligned");
  if (rows <= 0) return ZSG_OK;
  int cols;
  dim3 grid;
  if ((((uintptr_t)x | (uintptr_t)out) & 15) == 0 && cols8_geometry(c, rows, 8, cols, grid)) {
    bn_apply8_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(x, scale, shift, nullptr, nullptr, nullptr, relu, out, nullptr,
                                                                rows, c, cols);
    return check_launch("zsg_act_b16");
  }
  const int64_t n4 = rows * (c / 4);
  split_act_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(x, scale, shift, relu, nullptr, nullptr, out, n4, c / 4,
                                                                      TF_ACT);
  return check_launch("zsg_act_b16");
}

extern "C" int zsg_bn_apply_b16(const uint16_t* x, const float* scale, const float* shift, const uint16_t* r,
                                const float* rscale, const float* rshift, int relu, uint16_t* y, int64_t rows, int c,
                                zsg_stream_t stream) {
  ZSG_REQUIRE(x && scale && shift && y && c % 4 == 0, "zsg_bn_apply_b16: bad arguments");
  int cols;
  dim3 grid;
  if ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)r) & 15) == 0 && cols8_geometry(c, rows, 8, cols, grid)) {
    bn_apply8_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(x, scale, shift, r, rscale, rshift, relu, y, nullptr, rows, c,
                                                                cols);
    return check_launch("zsg_bn_apply_b16");
  }
  int64_t n4 = rows * (c / 4);
  bn_apply_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(x, scale, shift, r, rscale, rshift, relu, nullptr,
                                                                     nullptr, y, n4, c / 4, TF_ACT);
  return check_launch("zsg_bn_apply_b16");
}

extern "C" int zsg_bn_bwd_reduce_b16(const void* dy, int dy_is_b16, const uint16_t* x, const float* mean,
                                     const float* invstd, const float* scale, const float* shift, const uint16_t* act_out,
                                     int mask_mode, void* dz_out, int dz_is_b16, double* sums, int64_t rows, int c,
                                     zsg_stream_t stream) {
  ZSG_REQUIRE(dy && x && mean && invstd && sums, "zsg_bn_bwd_reduce_b16: null pointer");
  ZSG_REQUIRE(mask_mode != 1 || (scale && shift), "zsg_bn_bwd_reduce_b16: mask_mode 1 needs scale/shift");
  ZSG_REQUIRE(mask_mode != 2 || act_out, "zsg_bn_bwd_reduce_b16: mask_mode 2 needs act_out");
  int cols;
  dim3 grid;
  if ((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)act_out | (uintptr_t)dz_out) & 15) == 0 &&
      cols8_geometry(c, rows, 8, cols, grid, 2)) {
    cudaStream_t st = as_stream(stream);
#define ZSG_RED8(GB, ZB)                                                                                                        \
  bn_bwd_reduce8_kernel<true, GB, ZB><<<grid, 256, 0, st>>>(x, dy, act_out, mean, invstd, scale, shift, mask_mode, dz_out, sums, \
                                                            rows, c, cols)
    if (dy_is_b16) { if (dz_is_b16) ZSG_RED8(true, true); else ZSG_RED8(true, false); }
    else { if (dz_is_b16) ZSG_RED8(false, true); else ZSG_RED8(false, false); }
#undef ZSG_RED8
    return check_launch("zsg_bn_bwd_reduce_b16");
  }
  return launch_channel_reduce(1, x, dy, act_out, mean, invstd, scale, shift, mask_mode, dz_out, sums, rows, c,
                               as_stream(stream), TF_ACT | (dy_is_b16 ? TF_DY : 0) | (dz_is_b16 ? TF_DZ : 0));
}

extern "C" int zsg_bn_bwd_apply_b16(const void* dy, int dy_is_b16, const uint16_t* x, const float* mean,
                                    const float* invstd, const float* gamma, const float* scale, const float* shift,
                                    const uint16_t* act_out, int mask_mode, const double* sums, uint16_t* dx_bf16,
                                    float* dgamma, float* dbeta, int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(dy && x && mean && invstd && gamma && sums && dx_bf16 && c % 4 == 0, "zsg_bn_bwd_apply_b16: bad arguments");
  int cols;
  dim3 grid;
  if ((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)act_out | (uintptr_t)dx_bf16) & 15) == 0 &&
      cols8_geometry(c, rows, 8, cols, grid)) {
    if (dy_is_b16)
      bn_bwd_apply8_kernel<true, true><<<grid, 256, 0, as_stream(stream)>>>(dy, x, mean, invstd, gamma, scale, shift, act_out,
                                                                             mask_mode, sums, dx_bf16, nullptr, dgamma, dbeta, rows,
                                                                             c, cols);
    else
      bn_bwd_apply8_kernel<true, false><<<grid, 256, 0, as_stream(stream)>>>(dy, x, mean, invstd, gamma, scale, shift, act_out,
                                                                              mask_mode, sums, dx_bf16, nullptr, dgamma, dbeta, rows,
                                                                              c, cols);
    return check_launch("zsg_bn_bwd_apply_b16");
  }
  bn_bwd_apply_kernel<<<grid_for(rows * (c / 4), 256), 256, 0, as_stream(stream)>>>(
      dy, x, mean, invstd, gamma, scale, shift, act_out, mask_mode, sums, nullptr, nullptr, dx_bf16, dgamma, dbeta, rows, c,
      TF_ACT | (dy_is_b16 ? TF_DY : 0));
  return check_launch("zsg_bn_bwd_apply_b16");
}

extern "C" int zsg_maxpool_bn_relu_fwd_b16(const uint16_t* x, const float* scale, const float* shift, uint16_t* y,
                                           uint8_t* argmax, int b, int h, int w, int c, int ho, int wo,
                                           zsg_stream_t stream) {
  ZSG_REQUIRE(x && scale && shift && y && c % 4 == 0, "zsg_maxpool_bn_relu_fwd_b16: bad arguments");
  maxpool_fwd_kernel<<<grid_for((int64_t)b * ho * wo * (c / 4), 256), 256, 0, as_stream(stream)>>>(
      x, scale, shift, y, argmax, b, h, w, c, ho, wo, TF_ACT | TF_Y);
  return check_launch("zsg_maxpool_bn_relu_fwd_b16");
}

extern "C" int zsg_maxpool_bn_relu_bwd(const uint8_t* argmax, const float* dy, float* da, int b, int h, int w, int c,
                                       int ho, int wo, zsg_stream_t stream) {
  ZSG_REQUIRE(argmax && dy && da && c % 4 == 0, "zsg_maxpool_bn_relu_bwd: bad arguments");
  maxpool_bwd_kernel<<<grid_for((int64_t)b * h * w * (c / 4), 256, 16), 256, 0, as_stream(stream)>>>(argmax, dy, da, b,
                                                                                                      h, w, c, ho, wo);
  return check_launch("zsg_maxpool_bn_relu_bwd");
}

extern "C" int zsg_upsample_add(float* dst, const float* src, const int32_t* idx_y, const int32_t* idx_x, int b,
                                int ho, int wo, int hi, int wi, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(dst && src && idx_y && idx_x && c % 4 == 0, "zsg_upsample_add: bad arguments");
  upsample_add_kernel<<<grid_for((int64_t)b * ho * wo * (c / 4), 256), 256, 0, as_stream(stream)>>>(
      dst, src, idx_y, idx_x, b, ho, wo, hi, wi, c);
  return check_launch("zsg_upsample_add");
}

extern "C" int zsg_upsample_add_bwd(const float* ddst, float* dsrc, const int32_t* idx_y, const int32_t* idx_x, int b,
                                    int ho, int wo, int hi, int wi, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(ddst && dsrc && idx_y && idx_x && c % 4 == 0, "zsg_upsample_add_bwd: bad arguments");
  upsample_add_bwd_kernel<<<grid_for((int64_t)b * hi * wi * (c / 4), 256), 256, 0, as_stream(stream)>>>(
      ddst, dsrc, idx_y, idx_x, b, ho, wo, hi, wi, c);
  return check_launch("zsg_upsample_add_bwd");
}

extern "C" int zsg_avgpool_fwd(const float* x, float* y, int b, int hw, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(x && y, "zsg_avgpool_fwd: null pointer");
  avgpool_fwd_kernel<<<(b * c + 255) / 256, 256, 0, as_stream(stream)>>>(x, y, b, hw, c);
  return check_launch("zsg_avgpool_fwd");
}
extern "C" int zsg_avgpool_bwd(const float* dy, float* dx, int b, int hw, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(dy && dx, "zsg_avgpool_bwd: null pointer");
  avgpool_bwd_kernel<<<(b * hw * c + 255) / 256, 256, 0, as_stream(stream)>>>(dy, dx, b, hw, c);
  return check_launch("zsg_avgpool_bwd");
}

extern "C" int zsg_maxpool_fwd(const float* x, float* y, uint8_t* argmax, int b, int h, int w, int c, int k, int stride,
                               int pad, int ho, int wo, zsg_stream_t stream) {
  ZSG_REQUIRE(x && y && c % 4 == 0 && k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad < k, "zsg_maxpool_fwd: bad arguments");
  ZSG_REQUIRE(ho >= 1 && wo >= 1 && (ho - 1) * stride - pad < h && (wo - 1) * stride - pad < w,
              "zsg_maxpool_fwd: a window starts outside the image");
  pool_fwd_kernel<<<grid_for((int64_t)b * ho * wo * (c / 4), 256), 256, 0, as_stream(stream)>>>(x, y, argmax, b, h, w, c, k,
                                                                                                 stride, pad, ho, wo);
  return check_launch("zsg_maxpool_fwd");
}

extern "C" int zsg_maxpool_bwd(const uint8_t* argmax, const float* dy, const float* mask, float* dx, int b, int h, int w,
                               int c, int k, int stride, int pad, int ho, int wo, zsg_stream_t stream) {
  ZSG_REQUIRE(argmax && dy && dx && c % 4 == 0 && k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad < k,
              "zsg_maxpool_bwd: bad arguments");
  pool_bwd_kernel<<<grid_for((int64_t)b * h * w * (c / 4), 256, 16), 256, 0, as_stream(stream)>>>(argmax, dy, mask, dx, b, h,
                                                                                                   w, c, k, stride, pad, ho, wo);
  return check_launch("zsg_maxpool_bwd");
}

extern "C" int zsg_l2norm_fwd(const float* x, float* y, float* norm, int64_t rows, int c, zsg_stream_t stream) {
  ZSG_REQUIRE(x && y && norm && rows > 0 && c > 0 && c % 4 == 0 && c <= 1024, "zsg_l2norm_fwd: bad arguments (c <= 1024, c % 4 == 0)");
  l2norm_fwd_kernel<<<grid_for(rows * 32, 256), 256, 0, as_stream(stream)>>>(x, y, norm, rows, c);
  return check_launch("zsg_l2norm_fwd");
}

extern "C" int zsg_l2norm_bwd(const float* dy, const float* x, const float* norm, float* dx, int64_t rows, int c,
                              int accumulate, int mask_relu, zsg_stream_t stream) {
  ZSG_REQUIRE(dy && x && norm && dx && rows > 0 && c > 0 && c % 4 == 0 && c <= 1024, "zsg_l2norm_bwd: bad arguments");
  l2norm_bwd_kernel<<<grid_for(rows * 32, 256), 256, 0, as_stream(stream)>>>(dy, x, norm, dx, rows, c, accumulate, mask_relu);
  return check_launch("zsg_l2norm_bwd");
}

extern "C" int zsg_relu_bwd(const float* dy, const float* x, float* dx, int64_t n, int accumulate,
                            zsg_stream_t stream) {
  ZSG_REQUIRE(dy && x && dx && n % 4 == 0, "zsg_relu_bwd: bad arguments");
  relu_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(dy, x, dx, n / 4, accumulate);
  return check_launch("zsg_relu_bwd");
}

extern "C" int zsg_axpy(const float* x, float* y, float a, int64_t n, zsg_stream_t stream) {
  ZSG_REQUIRE(x && y, "zsg_axpy: null pointer");
  axpy_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, y, a, n);
  return check_launch("zsg_axpy");
}

extern "C" int zsg_scale_dev(float* x, int64_t rows, int width, int64_t stride, const double* scale,
                             zsg_stream_t stream) {
  ZSG_REQUIRE(x && scale && rows > 0 && width > 0 && stride >= width, "zsg_scale_dev: bad arguments");
  scale_dev_kernel<<<grid_for(rows * width, 256), 256, 0, as_stream(stream)>>>(x, rows, width, stride, scale);
  return check_launch("zsg_scale_dev");
}

static int make_levels(const int32_t* lvl_cells, int nlvl, Levels& lv) {
  ZSG_REQUIRE(lvl_cells && nlvl >= 1 && nlvl <= 8, "levels: nlvl=%d out of range", nlvl);
  lv.n = nlvl;
  for (int i = 0; i < 8; ++i) lv.cells[i] = i < nlvl ? lvl_cells[i] : 0;
  return ZSG_OK;
}

/* lvl_cells is a HOST array (six ints); everything else is device memory. */
extern "C" int zsg_fuse_lang_grid(const float* feat, const float* lang, const float* grid_yx, float* fused, int b,
                                  int total_cells, const int32_t* lvl_cells, int nlvl, int cfeat, int clang, int cpad,
                                  zsg_stream_t stream) {
  ZSG_REQUIRE(feat && lang && grid_yx && fused, "zsg_fuse_lang_grid: null pointer");
  ZSG_REQUIRE(cfeat % 4 == 0 && clang % 4 == 0 && cpad % 4 == 0 && cpad >= cfeat + clang + 2,
              "zsg_fuse_lang_grid: channel counts must be multiples of 4");
  Levels lv;
  if (int rc = make_levels(lvl_cells, nlvl, lv)) return rc;
  fuse_kernel<<<grid_for((int64_t)b * total_cells * (cpad / 4), 256), 256, 0, as_stream(stream)>>>(
      feat, lang, grid_yx, fused, b, total_cells, lv, cfeat, clang, cpad);
  return check_launch("zsg_fuse_lang_grid");
}

extern "C" int zsg_unfuse_lang_grid(const float* dfused, float* dfeat, float* dlang, int b, int total_cells,
                                    const int32_t* lvl_cells, int nlvl, int cfeat, int clang, int cpad,
                                    zsg_stream_t stream) {
  ZSG_REQUIRE(dfused && dfeat && dlang, "zsg_unfuse_lang_grid: null pointer");
  Levels lv;
  if (int rc = make_levels(lvl_cells, nlvl, lv)) return rc;
  cudaStream_t st = as_stream(stream);
  unfuse_feat_kernel<<<grid_for((int64_t)b * total_cells * (cfeat / 4), 256), 256, 0, st>>>(
      dfused, dfeat, (int64_t)b * total_cells, cfeat, cpad);
  cudaMemsetAsync(dlang, 0, (size_t)b * clang * sizeof(float), st);
  int max_cells = 1;
  for (int i = 0; i < nlvl; ++i) max_cells = lv.cells[i] > max_cells ? lv.cells[i] : max_cells;
  unfuse_lang_kernel<<<dim3(b, nlvl, (max_cells + UNFUSE_CHUNK - 1) / UNFUSE_CHUNK), 256, 0, st>>>(dfused, dlang, b, lv, cfeat,
                                                                                                  clang, cpad);
  return check_launch("zsg_unfuse_lang_grid");
}

extern "C" int zsg_copy_cols(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t rows, int c,
                             zsg_stream_t stream) {
  ZSG_REQUIRE(src && dst && rows > 0 && c > 0 && src_ld >= c && dst_ld >= c, "zsg_copy_cols: bad arguments");
  copy_cols_kernel<<<grid_for(rows * c, 256), 256, 0, as_stream(stream)>>>(src, src_ld, dst, dst_ld, rows, c);
  return check_launch("zsg_copy_cols");
}

extern "C" int zsg_head0_lang_grid_terms(const float* v, const float* wg, const float* gridpatch, float* lang_cls,
                                         float* grid_term, int b, int total_cells, int n, zsg_stream_t stream) {
  ZSG_REQUIRE(v && wg && gridpatch && lang_cls && grid_term && b > 0 && total_cells > 0 && n > 0,
              "zsg_head0_lang_grid_terms: bad arguments");
  cudaStream_t st = as_stream(stream);
  lang_class_kernel<<<dim3(b, 16), 256, 0, st>>>(v, lang_cls, n);
  grid_term_kernel<<<total_cells, 256, 0, st>>>(wg, gridpatch, grid_term, n);
  return check_launch("zsg_head0_lang_grid_terms");
}

extern "C" int zsg_head0_backward_sums(const float* dh, const int32_t* cell_base, const int32_t* cell_stride,
                                       const int32_t* cell_cls, const float* gridpatch, int b, int total_cells, int n,
                                       float* scratch, size_t scratch_floats, float* tap_sums, float* dwg, int64_t dwg_ld,
                                       zsg_stream_t stream) {
  ZSG_REQUIRE(dh && cell_base && cell_stride && cell_cls && gridpatch && scratch && tap_sums && dwg,
              "zsg_head0_backward_sums: null pointer");
  ZSG_REQUIRE(n == 256, "zsg_head0_backward_sums: n=%d (the head width of the path is 256)", n);
  const int nch = 8;
  ZSG_REQUIRE(scratch_floats >= (size_t)b * nch * (16 + 18) * n && ((uintptr_t)gridpatch & 7) == 0,
              "zsg_head0_backward_sums: scratch too small or gridpatch not 8-byte aligned");
  float* Sp = scratch;
  float* Wp = scratch + (size_t)b * nch * 16 * n;
  cudaStream_t st = as_stream(stream);
  head0_reduce_kernel<<<dim3(b, nch), 256, 0, st>>>(dh, cell_base, cell_stride, cell_cls, gridpatch, total_cells, n, Sp, Wp);
  head0_finish_kernel<<<b + 18, 256, 0, st>>>(Sp, Wp, b, nch, n, tap_sums, dwg, dwg_ld);
  return check_launch("zsg_head0_backward_sums");
}

extern "C" int zsg_weight_transpose_flip(const float* w, float* wt, int cout, int rs_r, int rs_s, int cin,
                                         zsg_stream_t stream) {
  ZSG_REQUIRE(w && wt, "zsg_weight_transpose_flip: null pointer");
  weight_transpose_flip_kernel<<<grid_for((int64_t)cout * rs_r * rs_s * cin, 256), 256, 0, as_stream(stream)>>>(
      w, wt, cout, rs_r, rs_s, cin);
  return check_launch("zsg_weight_transpose_flip");
}

extern "C" int zsg_weight_transpose_flip_batched(const float* src_base, float* dst_base, const zsg_wtf_desc* descs, int n,
                                                 int64_t total, zsg_stream_t stream) {
  ZSG_REQUIRE(src_base && dst_base && descs && n > 0 && total > 0, "zsg_weight_transpose_flip_batched: bad arguments");
  weight_transpose_flip_batched_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(src_base, dst_base, descs, n, total);
  return check_launch("zsg_weight_transpose_flip_batched");
}

extern "C" int zsg_weight_transpose_flip_batched32(const float* src_base, float* dst_base, const zsg_wtf_desc* descs, int n,
                                                   int64_t total, zsg_stream_t stream) {
  ZSG_REQUIRE(src_base && dst_base && descs && n > 0 && total > 0 && total % 1024 == 0,
              "zsg_weight_transpose_flip_batched32: bad arguments (every cin and cout must be a multiple of 32)");
  const int64_t tiles = total / 1024;
  const int64_t cap = (int64_t)num_sms() * 16;
  weight_transpose_flip_tiled_kernel<<<(unsigned)(tiles < cap ? tiles : cap), 256, 0, as_stream(stream)>>>(src_base, dst_base,
                                                                                                          descs, n, tiles);
  return check_launch("zsg_weight_transpose_flip_batched32");
}

extern "C" int zsg_split_tf32(const float* w, float* hi, float* lo, int64_t n, zsg_stream_t stream) {
  ZSG_REQUIRE(w && hi && lo && n > 0, "zsg_split_tf32: bad arguments");
  split_tf32_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(w, hi, lo, n);
  return check_launch("zsg_split_tf32");
}

extern "C" int zsg_pad_channels(const float* src, float* dst, int64_t n, int csrc, int cdst, zsg_stream_t stream) {
  ZSG_REQUIRE(src && dst && csrc > 0 && cdst > 0, "zsg_pad_channels: bad arguments");
  pad_channels_kernel<<<grid_for(n * cdst, 256), 256, 0, as_stream(stream)>>>(src, dst, n, csrc, cdst);
  return check_launch("zsg_pad_channels");
}

extern "C" int zsg_nchw_to_nhwc4(const float* img, float* out, int b, int h, int w, zsg_stream_t stream) {
  ZSG_REQUIRE(img && out, "zsg_nchw_to_nhwc4: null pointer");
  nchw_to_nhwc4_kernel<<<grid_for((int64_t)b * h * w, 256), 256, 0, as_stream(stream)>>>(img, out, b, h, w);
  return check_launch("zsg_nchw_to_nhwc4");
}

extern "C" int zsg_nchw_to_nhwc8_bf16(const float* img, uint16_t* out, int b, int h, int w, zsg_stream_t stream) {
  ZSG_REQUIRE(img && out && b > 0 && h > 0 && w > 0 && ((uintptr_t)out & 15) == 0, "zsg_nchw_to_nhwc8_bf16: bad arguments");
  nchw_to_nhwc8_bf16_kernel<<<grid_for((int64_t)b * h * w, 256), 256, 0, as_stream(stream)>>>(img, out, b, h, w);
  return check_launch("zsg_nchw_to_nhwc8_bf16");
}

extern "C" int zsg_gather_rows(const float* src, const zsg_row_t* rows, float* dst, int64_t m, int csrc, int cdst,
                               zsg_stream_t stream) {
  ZSG_REQUIRE(src && rows && dst && m > 0 && csrc > 0 && cdst > 0, "zsg_gather_rows: bad arguments");
  gather_rows_kernel<<<grid_for(m * cdst, 256), 256, 0, as_stream(stream)>>>(src, rows, dst, m, csrc, cdst);
  return check_launch("zsg_gather_rows");
}

extern "C" int zsg_colsum(const float* x, float* out, int64_t rows, int c, int ld, int accumulate, zsg_stream_t stream) {
  ZSG_REQUIRE(x && out && rows > 0 && c > 0 && ld >= c, "zsg_colsum: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (!accumulate) cudaMemsetAsync(out, 0, (size_t)c * sizeof(float), st);
  int gx = (c + 31) / 32;
  int64_t gy = (rows + 255) / 256;
  int64_t cap = (int64_t)num_sms() * 8 / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  colsum_kernel<<<dim3(gx, (unsigned)gy), 256, 0, st>>>(x, out, rows, c, ld);
  return check_launch("zsg_colsum");
}

extern "C" int zsg_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                        float eps, int step, float grad_scale, zsg_stream_t stream) {
  ZSG_REQUIRE(p && g && m && v && step >= 1, "zsg_adam: bad arguments");
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2,
                                                                grad_scale);
  return check_launch("zsg_adam");
}
