// Microbenchmark (diagnostics): two issuing threads (in two warps) alternate 12-MMA groups with an mbarrier token
// handoff, each doing ~W dummy dependent instructions of "prologue" before waiting for its token.
// Compares with a single issuer doing the same prologue.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) {
  long long t0 = clock64();
  while (!mbar_try(bar, ph)) if (clock64() - t0 > 2000000000LL) { printf("timeout\n"); __trap(); }
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t busy(uint32_t x, int w) {      // w dependent integer instructions
  for (int i = 0; i < w; ++i) x = x * 1664525u + 1013904223u;
  return x;
}

__device__ __forceinline__ void group12(uint32_t tmem, uint32_t base16, uint64_t desc_hi, uint32_t idesc, uint32_t first) {
  const uint32_t a_hi = base16, a_lo = base16 + 1024, b_hi = base16 + 2048, b_lo = b_hi + 1024;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t dah = desc_hi | (a_hi + 2 * k), dal = desc_hi | (a_lo + 2 * k), dbh = desc_hi | (b_hi + 2 * k), dbl = desc_hi | (b_lo + 2 * k);
    mma(tmem, dal, dbh, idesc, k ? 1u : first);
    mma(tmem, dah, dbl, idesc, 1u);
    mma(tmem, dah, dbh, idesc, 1u);
  }
}

// issuers: 1 = single thread, 2 = ping-pong between warp 0 and warp 1
__global__ void __launch_bounds__(128, 1) bench(int iters, int issuers, int work, long long* out, float* result) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  __shared__ uint64_t done[2], tok[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;   // 192 KB of ones
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(smem_u32(&done[0]), 1); mbar_init(smem_u32(&done[1]), 1);
      mbar_init(smem_u32(&tok[0]), 1); mbar_init(smem_u32(&tok[1]), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
  const uint32_t stage16 = (64 * 1024) >> 4;
  long long t0 = clock64();
  uint32_t sink = 0;
  const int nissue = issuers == 3 ? 2 : issuers;
  if (warp < nissue) {
    if (elect_one()) {
      // thread `warp` issues groups g = warp, warp + issuers, ...; group g uses smem stage g % 3; accumulator zeroed at g == 0 only
      for (int g = warp; g < iters; g += nissue) {
        sink = busy(sink + g, work);                                  // "prologue": barrier checks, descriptors ...
        const uint32_t base16 = (smem_u32(sm) >> 4) + (g % 3) * stage16 + (sink & 0);
        if (issuers == 2 && g > 0) mbar_wait(smem_u32(&tok[warp]), ((g - 1) >> 1) & 1);      // my turn?
        if (issuers == 3) group12(tmem + 128 * warp, base16, desc_hi, idesc, g >= 2 ? 1u : 0u);
        else group12(tmem, base16, desc_hi, idesc, g ? 1u : 0u);
        if (issuers == 2) mbar_arrive(smem_u32(&tok[warp ^ 1]));                              // hand over
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done[warp])) : "memory");
      mbar_wait(smem_u32(&done[warp]), 0);
    }
    __syncwarp();
  }
  long long t1 = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = sink; }
  // read back accumulator element (lane 0.., column 0) : expected iters * 12 * 8 (all ones) if nothing was lost / reordered
  if (warp == 0) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tmem) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t r2;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r2) : "r"(tmem + 128) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0 && lane == 5) result[0] = __uint_as_float(r) + (issuers == 3 ? __uint_as_float(r2) : 0.f);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

void run(const char* name, int issuers, int work, int iters = 4000) {
  long long* d; float* r;
  cudaMalloc(&d, 64); cudaMalloc(&r, 16);
  cudaMemset(d, 0, 64);
  const int smem = 192 * 1024 + 2048;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) bench<<<148, 128, smem>>>(iters, issuers, work, d, r);
  cudaError_t e = cudaDeviceSynchronize();
  long long hh[2]; float hr = 0;
  cudaMemcpy(hh, d, 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(&hr, r, 4, cudaMemcpyDeviceToHost);
  printf("%-22s issuers=%d work=%3d : %7.1f cyc/MMA (%5.1f%% of the 64-cycle floor)  acc[5][0]=%.0f expected %.0f  (%s)\n", name, issuers, work,
         (double)hh[0] / (iters * 12.0), 6400.0 * iters * 12.0 / (double)hh[0], hr, iters * 12.0 * 8.0, cudaGetErrorString(e));
  cudaFree(d); cudaFree(r);
}

int main() {
  for (int work : {0, 20, 50, 100}) {
    run("single issuer", 1, work);
    run("ping-pong", 2, work);
    run("two free-running", 3, work);
  }
  for (int rep = 0; rep < 20; ++rep) run("two free-running long", 3, 37 + rep, 200000);
  return 0;
}
