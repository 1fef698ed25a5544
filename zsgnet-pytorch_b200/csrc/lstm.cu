// bi-LSTM query encoder recurrences (mdl.py:296-336), forward and BPTT.
//
// What the reference consumes is lstm_out[len-1]: the forward direction after len tokens and
// the reverse direction at position len-1, i.e. after ONE step from its random initial state.
// The remaining reverse-direction steps never reach the loss, so they are not computed: the
// results (values and gradients) are identical and 19/20 of the reverse recurrence disappears.
//
// The dense parts (x*W_ih^T and every weight gradient) run on the tensor cores through
// zsg_conv_fwd / zsg_conv_wgrad (a 1x1 "conv" over B*T rows); what is left here is the
// strictly sequential part: h(t-1)*W_hh^T + gate math, latency-bound, one CTA per sample
// with W_hh streamed from L2 (256 KB, shared by all CTAs).
#include <math.h>
#include "common.cuh"

namespace zsg {

constexpr int H = 128;
constexpr int G = 4 * H;   // gate rows: i, f, g, o (PyTorch order)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// one CTA (512 threads) per sample; thread j owns gate row j
__global__ void __launch_bounds__(G) lstm_fwd_dir_kernel(const float* __restrict__ gx, const float* __restrict__ whh_t,
                                                         const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                         const float* __restrict__ h0, const float* __restrict__ c0,
                                                         const int32_t* __restrict__ lens, int T,
                                                         float* __restrict__ gates, float* __restrict__ cs,
                                                         float* __restrict__ hprev, float* __restrict__ lang) {
  __shared__ float h_s[H];
  __shared__ float g_s[G];
  const int b = blockIdx.x, j = threadIdx.x;
  const int L = lens[b];
  float c = 0.f;
  if (j < H) { h_s[j] = h0[(size_t)b * H + j]; c = c0[(size_t)b * H + j]; }
  const float bias = b_ih[j] + b_hh[j];
  __syncthreads();
  for (int t = 0; t < L; ++t) {
    float acc = gx[((size_t)b * T + t) * G + j] + bias;
#pragma unroll 8
    for (int k = 0; k < H; ++k) acc = fmaf(__ldg(whh_t + (size_t)k * G + j), h_s[k], acc);
    const float a = (j >= 2 * H && j < 3 * H) ? tanhf(acc) : sigmoidf_(acc);
    g_s[j] = a;
    gates[((size_t)b * T + t) * G + j] = a;
    if (j < H) hprev[((size_t)b * T + t) * H + j] = h_s[j];
    __syncthreads();
    if (j < H) {
      c = g_s[H + j] * c + g_s[j] * g_s[2 * H + j];
      cs[((size_t)b * T + t) * H + j] = c;
      h_s[j] = g_s[3 * H + j] * tanhf(c);
    }
    __syncthreads();
  }
  if (j < H) lang[(size_t)b * 2 * H + j] = h_s[j];
}

// reverse direction, single step on token len-1.  One CTA per sample, warp per gate row group.
__global__ void __launch_bounds__(G) lstm_rev_step_kernel(const float* __restrict__ qvec, const float* __restrict__ wih,
                                                          const float* __restrict__ whh, const float* __restrict__ b_ih,
                                                          const float* __restrict__ b_hh, const float* __restrict__ h0,
                                                          const float* __restrict__ c0, const int32_t* __restrict__ lens,
                                                          int T, int E, float* __restrict__ xlast,
                                                          float* __restrict__ gates, float* __restrict__ lang) {
  extern __shared__ float sm[];           // x[E] | h[H] | g[G]
  float* x_s = sm;
  float* h_s = sm + E;
  float* g_s = h_s + H;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int L = lens[b];
  const float* xr = qvec + ((size_t)b * T + (L - 1)) * E;
  for (int k = tid; k < E; k += blockDim.x) { float v = xr[k]; x_s[k] = v; xlast[(size_t)b * E + k] = v; }
  if (tid < H) h_s[tid] = h0[(size_t)b * H + tid];
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  for (int row = warp; row < G; row += nwarp) {
    float acc = 0.f;
    for (int k = lane; k < E; k += 32) acc = fmaf(wih[(size_t)row * E + k], x_s[k], acc);
    for (int k = lane; k < H; k += 32) acc = fmaf(whh[(size_t)row * H + k], h_s[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += b_ih[row] + b_hh[row];
      const float a = (row >= 2 * H && row < 3 * H) ? tanhf(acc) : sigmoidf_(acc);
      g_s[row] = a;
      gates[(size_t)b * G + row] = a;
    }
  }
  __syncthreads();
  if (tid < H) {
    const float c = g_s[H + tid] * c0[(size_t)b * H + tid] + g_s[tid] * g_s[2 * H + tid];
    lang[(size_t)b * 2 * H + H + tid] = g_s[3 * H + tid] * tanhf(c);
  }
}

// BPTT, forward direction.  Thread j<H owns hidden unit j for the cell math; all 512 threads
// share the dh(t-1) = W_hh^T * dG(t) product (4 partial sums per unit, reduced through smem).
__global__ void __launch_bounds__(G) lstm_bwd_dir_kernel(const float* __restrict__ dlang, const float* __restrict__ whh,
                                                         const float* __restrict__ gates, const float* __restrict__ cs,
                                                         const float* __restrict__ c0, const int32_t* __restrict__ lens,
                                                         int T, float* __restrict__ dgates) {
  __shared__ float dg_s[G];
  __shared__ float part[4][H];
  const int b = blockIdx.x, j = threadIdx.x;
  const int L = lens[b];
  float dh = 0.f, dc = 0.f;
  if (j < H) dh = dlang[(size_t)b * 2 * H + j];
  for (int t = T - 1; t >= L; --t) dgates[((size_t)b * T + t) * G + j] = 0.f;
  for (int t = L - 1; t >= 0; --t) {
    const float* gt = gates + ((size_t)b * T + t) * G;
    if (j < H) {
      const float i = gt[j], f = gt[H + j], g = gt[2 * H + j], o = gt[3 * H + j];
      const float c = cs[((size_t)b * T + t) * H + j];
      const float cp = t > 0 ? cs[((size_t)b * T + t - 1) * H + j] : c0[(size_t)b * H + j];
      const float th = tanhf(c);
      const float d_o = dh * th;
      dc += dh * o * (1.f - th * th);
      dg_s[j] = dc * g * i * (1.f - i);
      dg_s[H + j] = dc * cp * f * (1.f - f);
      dg_s[2 * H + j] = dc * i * (1.f - g * g);
      dg_s[3 * H + j] = d_o * o * (1.f - o);
      dc = dc * f;
    }
    __syncthreads();
    dgates[((size_t)b * T + t) * G + j] = dg_s[j];
    {
      const int k = j & (H - 1), q = j >> 7;             // unit k, row quarter q
      float acc = 0.f;
#pragma unroll 8
      for (int r = q * H; r < (q + 1) * H; ++r) acc = fmaf(__ldg(whh + (size_t)r * H + k), dg_s[r], acc);
      part[q][k] = acc;
    }
    __syncthreads();
    if (j < H) dh = part[0][j] + part[1][j] + part[2][j] + part[3][j];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(H) lstm_rev_step_bwd_kernel(const float* __restrict__ dlang,
                                                              const float* __restrict__ gates,
                                                              const float* __restrict__ c0, float* __restrict__ dgates) {
  const int b = blockIdx.x, j = threadIdx.x;
  const float* gt = gates + (size_t)b * G;
  const float i = gt[j], f = gt[H + j], g = gt[2 * H + j], o = gt[3 * H + j];
  const float cp = c0[(size_t)b * H + j];
  const float c = f * cp + i * g;
  const float th = tanhf(c);
  const float dh = dlang[(size_t)b * 2 * H + H + j];
  const float dc = dh * o * (1.f - th * th);
  float* dg = dgates + (size_t)b * G;
  dg[j] = dc * g * i * (1.f - i);
  dg[H + j] = dc * cp * f * (1.f - f);
  dg[2 * H + j] = dc * i * (1.f - g * g);
  dg[3 * H + j] = dh * th * o * (1.f - o);
}

}  // namespace zsg

using namespace zsg;

extern "C" int zsg_lstm_fwd_dir(const float* gx, const float* whh_t, const float* b_ih, const float* b_hh,
                                const float* h0, const float* c0, const int32_t* lens, int b, int t, float* gates,
                                float* cs, float* hprev, float* lang, zsg_stream_t stream) {
  ZSG_REQUIRE(gx && whh_t && b_ih && b_hh && h0 && c0 && lens && gates && cs && hprev && lang,
              "zsg_lstm_fwd_dir: null pointer");
  ZSG_REQUIRE(b > 0 && t > 0, "zsg_lstm_fwd_dir: empty batch");
  lstm_fwd_dir_kernel<<<b, G, 0, as_stream(stream)>>>(gx, whh_t, b_ih, b_hh, h0, c0, lens, t, gates, cs, hprev, lang);
  return check_launch("zsg_lstm_fwd_dir");
}

extern "C" int zsg_lstm_rev_step(const float* qvec, const float* wih, const float* whh, const float* b_ih,
                                 const float* b_hh, const float* h0, const float* c0, const int32_t* lens, int b, int t,
                                 int e, float* xlast, float* gates, float* lang, zsg_stream_t stream) {
  ZSG_REQUIRE(qvec && wih && whh && b_ih && b_hh && h0 && c0 && lens && xlast && gates && lang,
              "zsg_lstm_rev_step: null pointer");
  size_t smem = (size_t)(e + H + G) * sizeof(float);
  lstm_rev_step_kernel<<<b, G, smem, as_stream(stream)>>>(qvec, wih, whh, b_ih, b_hh, h0, c0, lens, t, e, xlast, gates,
                                                          lang);
  return check_launch("zsg_lstm_rev_step");
}

extern "C" int zsg_lstm_bwd_dir(const float* dlang, const float* whh, const float* gates, const float* cs,
                                const float* c0, const int32_t* lens, int b, int t, float* dgates,
                                zsg_stream_t stream) {
  ZSG_REQUIRE(dlang && whh && gates && cs && c0 && lens && dgates, "zsg_lstm_bwd_dir: null pointer");
  lstm_bwd_dir_kernel<<<b, G, 0, as_stream(stream)>>>(dlang, whh, gates, cs, c0, lens, t, dgates);
  return check_launch("zsg_lstm_bwd_dir");
}

extern "C" int zsg_lstm_rev_step_bwd(const float* dlang, const float* gates, const float* c0, int b, float* dgates,
                                     zsg_stream_t stream) {
  ZSG_REQUIRE(dlang && gates && c0 && dgates, "zsg_lstm_rev_step_bwd: null pointer");
  lstm_rev_step_bwd_kernel<<<b, H, 0, as_stream(stream)>>>(dlang, gates, c0, dgates);
  return check_launch("zsg_lstm_rev_step_bwd");
}
