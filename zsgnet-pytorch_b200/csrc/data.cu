// Data-path kernels (dat_loader.py:98-146): what the reference does per sample on a CPU worker after
// the JPEG is decoded -- `img.resize((300, 300))` (Pillow ImagingResample, 8-bit fixed point), `pil2tensor(...).float()
// .div_(255)` and the word-vector lookup -- for a whole batch in three launches.  Byte / integer work, HBM-bound:
// results are bit-identical to Pillow's (tests/test_gpu_data_gpu.py), the coefficient / index tables are built on the
// host exactly like Resample.c precompute_coeffs / Geometry.c ImagingScaleAffine (zsg_b200/gpu_data.py).
#include "common.cuh"

namespace zsg {

constexpr int PRECISION_BITS = 32 - 8 - 2;               // Resample.c

__device__ __forceinline__ uint8_t clip8(int v) {         // Resample.c clip8(): clip8_lookups[v >> PRECISION_BITS]
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass over the source rows the vertical pass needs: tmp[row][xx][c], row = source row - y_first
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, const zsg_resize_desc* __restrict__ descs,
                                                       const int32_t* __restrict__ tables, int out_w,
                                                       uint8_t* __restrict__ tmp) {
  const zsg_resize_desc d = descs[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d.n_rows * out_w) return;
  const int row = idx / out_w, xx = idx - row * out_w;
  const uint8_t* line = src + d.src_off + (int64_t)(d.y_first + row) * d.w * 3;
  uint8_t* o = tmp + d.tmp_off + (int64_t)idx * 3;
  if (d.w == out_w) {                                     // Pillow skips the pass: the rows are used as they are
    o[0] = line[xx * 3]; o[1] = line[xx * 3 + 1]; o[2] = line[xx * 3 + 2];
    return;
  }
  const int32_t* t = tables + d.hk_off + (int64_t)xx * (2 + d.hksize);
  const int xmin = t[0], xmax = t[1];
  const int32_t* k = t + 2;
  const uint8_t* p = line + xmin * 3;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < xmax; ++x) {
    const int kx = __ldg(k + x);
    s0 += (int)p[3 * x] * kx;
    s1 += (int)p[3 * x + 1] * kx;
    s2 += (int)p[3 * x + 2] * kx;
  }
  o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// vertical pass + pil2tensor + /255: out[img][c][yy][xx] float32
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ tmp, const zsg_resize_desc* __restrict__ descs,
                                                       const int32_t* __restrict__ tables, int out_h, int out_w,
                                                       float* __restrict__ out) {
  const zsg_resize_desc d = descs[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= out_h * out_w) return;
  const int yy = idx / out_w, xx = idx - yy * out_w;
  const uint8_t* base = tmp + d.tmp_off + (int64_t)xx * 3;
  uint8_t v0, v1, v2;
  if (d.h == out_h) {                                     // no vertical pass (y_first == 0, n_rows == h)
    const uint8_t* p = base + (int64_t)yy * out_w * 3;
    v0 = p[0]; v1 = p[1]; v2 = p[2];
  } else {
    const int32_t* t = tables + d.vk_off + (int64_t)yy * (2 + d.vksize);
    const int ymin = t[0], ymax = t[1];                   // already relative to y_first
    const int32_t* k = t + 2;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < ymax; ++y) {
      const uint8_t* p = base + (int64_t)(ymin + y) * out_w * 3;
      const int ky = __ldg(k + y);
      s0 += (int)p[0] * ky;
      s1 += (int)p[1] * ky;
      s2 += (int)p[2] * ky;
    }
    v0 = clip8(s0); v1 = clip8(s1); v2 = clip8(s2);
  }
  const int64_t plane = (int64_t)out_h * out_w;
  float* o = out + (int64_t)blockIdx.y * 3 * plane + idx;
  o[0] = __fdiv_rn((float)v0, 255.0f);
  o[plane] = __fdiv_rn((float)v1, 255.0f);
  o[2 * plane] = __fdiv_rn((float)v2, 255.0f);
}

// NEAREST (Image.resize's default before Pillow 7.0, i.e. under the reference's pinned pillow 6.1): index tables
// xtab[out_w] at hk_off, ytab[out_h] at vk_off
__global__ void __launch_bounds__(256) resize_nearest_kernel(const uint8_t* __restrict__ src,
                                                             const zsg_resize_desc* __restrict__ descs,
                                                             const int32_t* __restrict__ tables, int out_h, int out_w,
                                                             float* __restrict__ out) {
  const zsg_resize_desc d = descs[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= out_h * out_w) return;
  const int yy = idx / out_w, xx = idx - yy * out_w;
  const int ys = __ldg(tables + d.vk_off + yy), xs = __ldg(tables + d.hk_off + xx);
  const uint8_t* p = src + d.src_off + ((int64_t)ys * d.w + xs) * 3;
  const int64_t plane = (int64_t)out_h * out_w;
  float* o = out + (int64_t)blockIdx.y * 3 * plane + idx;
  o[0] = __fdiv_rn((float)p[0], 255.0f);
  o[plane] = __fdiv_rn((float)p[1], 255.0f);
  o[2 * plane] = __fdiv_rn((float)p[2], 255.0f);
}

// out[i][0:dim] = table[tokens[i]][0:dim] (zeros for tokens[i] < 0): the per-token `q.vector` lookup of dat_loader.py:115
__global__ void __launch_bounds__(256) embed_gather_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ table,
                                                           float* __restrict__ out, int64_t n, int dim4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * dim4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / dim4;
    const int c = (int)(i - row * dim4);
    const int tok = __ldg(tokens + row);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tok >= 0) v = __ldg(reinterpret_cast<const float4*>(table) + (int64_t)tok * dim4 + c);
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

}  // namespace zsg

using namespace zsg;

extern "C" int zsg_resize_rgb8(const uint8_t* src, const zsg_resize_desc* descs, const int32_t* tables, int n_images,
                               int out_h, int out_w, int max_rows, int nearest, uint8_t* workspace, float* out,
                               zsg_stream_t stream) {
  ZSG_REQUIRE(src && descs && tables && out, "zsg_resize_rgb8: null pointer");
  ZSG_REQUIRE(n_images > 0 && n_images <= 65535 && out_h > 0 && out_w > 0, "zsg_resize_rgb8: bad sizes");
  ZSG_REQUIRE(nearest || (workspace && max_rows > 0), "zsg_resize_rgb8: the two-pass filter needs a workspace");
  cudaStream_t st = as_stream(stream);
  const int out_px = out_h * out_w;
  if (nearest) {
    resize_nearest_kernel<<<dim3((out_px + 255) / 256, n_images), 256, 0, st>>>(src, descs, tables, out_h, out_w, out);
    return check_launch("zsg_resize_rgb8(nearest)");
  }
  resize_h_kernel<<<dim3((max_rows * out_w + 255) / 256, n_images), 256, 0, st>>>(src, descs, tables, out_w, workspace);
  resize_v_kernel<<<dim3((out_px + 255) / 256, n_images), 256, 0, st>>>(workspace, descs, tables, out_h, out_w, out);
  return check_launch("zsg_resize_rgb8");
}

extern "C" int zsg_embed_gather(const int32_t* tokens, const float* table, float* out, int64_t n_tokens, int dim,
                                zsg_stream_t stream) {
  ZSG_REQUIRE(tokens && table && out && n_tokens > 0, "zsg_embed_gather: bad arguments");
  ZSG_REQUIRE(dim > 0 && dim % 4 == 0 && ((((uintptr_t)table | (uintptr_t)out) & 15) == 0),
              "zsg_embed_gather: dim must be a multiple of 4 and the tensors 16-byte aligned");
  const int64_t work = n_tokens * (dim / 4);
  int blocks = (int)((work + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  embed_gather_kernel<<<blocks, 256, 0, as_stream(stream)>>>(tokens, table, out, n_tokens, dim / 4);
  return check_launch("zsg_embed_gather");
}
