// Microbenchmark (diagnostics, not product): which (TMEM lane, column) lands in which (thread, register) of
// tcgen05.ld.16x256b.x8 -- the accumulator-fragment layout the register epilogue of conv_tc.cu relies on.
// TMEM is filled with value = lane * 1000 + column through tcgen05.st.32x32b (thread = lane), then read back with
// 16x256b.x8 at lane offsets 0 and 16.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_layout tmem_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  // warp w may touch lanes [32 * (w % 4), +32): 4 warps fill all 128 lanes, 64 columns
  const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) v[j] = (uint32_t)((warp * 32 + lane) * 1000 + c0 + j);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                   "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int h = 0; h < 2; ++h) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr + ((uint32_t)(h * 16) << 16))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) out[((warp * 2 + h) * 32 + lane) * 32 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64u) : "memory");
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 4 * 2 * 32 * 32 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[4 * 2 * 32 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  // check the hypothesis: reg 4*u + 2*hh + c of thread t (load h) = lane 32*w + 16*h + 8*hh + t/4, column 8*u + 2*(t%4) + c
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int hd = 0; hd < 2; ++hd) for (int t = 0; t < 32; ++t) for (int i = 0; i < 32; ++i) {
    const int u = i / 4, hh = (i / 2) & 1, c = i & 1;
    const uint32_t want = (uint32_t)((32 * w + 16 * hd + 8 * hh + t / 4) * 1000 + 8 * u + 2 * (t % 4) + c);
    if (h[((w * 2 + hd) * 32 + t) * 32 + i] != want) ++bad;
  }
  printf("hypothesis mismatches: %d\n", bad);
  for (int t = 0; t < 6; ++t) {
    printf("warp 0, load 0, thread %d:", t);
    for (int i = 0; i < 12; ++i) printf(" %u", h[t * 32 + i]);
    printf("\n");
  }
  return 0;
}
