// Microbenchmark (diagnostics, not product): issue rate of tcgen05.mma kind::tf32 / kind::f16 in SS and TS mode,
// K-major and MN-major operands, with and without concurrent shared-memory store traffic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu ; run: ./mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) {
  long long t0 = clock64();
  while (!mbar_try(bar, ph)) if (clock64() - t0 > 2000000000LL) { printf("timeout\n"); __trap(); }
}

// mode: 0 = SS K-major tf32, 1 = TS tf32 (A in TMEM), 2 = SS MN-major tf32 (BASE32B swizzle), 3 = SS K-major f16 (K=16), 4 = TS f16
template <int MODE>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint32_t a_t, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (MODE == 0 || MODE == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (MODE == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (MODE == 3)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int MODE, int N>
__global__ void __launch_bounds__(288, 1) bench(int iters, int sts_warps, int sts_per_iter, long long* out, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t bars2[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;   // 192 KB
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars2[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  constexpr bool F16 = MODE >= 3;
  constexpr bool MN = MODE == 2;
  // idesc: D fp32 (1<<4); A,B fmt: tf32 = 2, f16 = 0 (bf16 = 1); N>>3 at 17; M>>4 at 24; MN-major bits 15,16
  const uint32_t idesc = (1u << 4) | ((F16 ? 1u : 2u) << 7) | ((F16 ? 1u : 2u) << 10) | (MN ? (3u << 15) : 0u) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t desc_hi = MN ? ((256ull << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61))
                              : ((1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61));
  const uint32_t kstep = MN ? 64u : 2u;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    // 3 stages of (A 16K + A' 16K + B N*128 + B' N*128)
    const uint32_t stage_bytes = 2 * 16384 + 2 * N * 128;
    const uint32_t nst = (N == 128) ? 3 : 2;
    __syncwarp();
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t base = (smem_u32(sm) + (it % nst) * stage_bytes) >> 4;
      const uint32_t a_hi = base, a_lo = base + 1024, b_hi = base + 2048, b_lo = b_hi + (N * 128 >> 4);
      const uint32_t ta = tmem + 256 + (it % 3) * 64;       // TS: A hi at cols +0..31, lo at +32..63
      const uint32_t dacc = (variant & 4) ? tmem + (((it >> 2) & 1) * 128) : tmem;     // variant 4: two accumulators, 4 iterations each
      const uint32_t first = (variant & 4) ? ((it & 3) ? 1u : 0u) : (it ? 1u : 0u);
      if (variant & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      bool leader = lane == 0;
      if (variant & 2) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        leader = pred != 0;
      }
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dah = desc_hi | ((a_hi + kstep * k) & 0x3FFF), dal = desc_hi | ((a_lo + kstep * k) & 0x3FFF);
          const uint64_t dbh = desc_hi | ((b_hi + kstep * k) & 0x3FFF), dbl = desc_hi | ((b_lo + kstep * k) & 0x3FFF);
          const uint32_t kc = F16 ? 4 * k : 8 * k;
          mma<MODE>(dacc, dal, ta + 32 + kc, dbh, idesc, k ? 1u : first);
          mma<MODE>(dacc, dah, ta + kc, dbl, idesc, 1u);
          mma<MODE>(dacc, dah, ta + kc, dbh, idesc, 1u);
        }
        if (variant & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars2[it & 3])) : "memory");
        if ((variant & 8) && (it & 3) == 3) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars2[4 + ((it >> 2) & 1)])) : "memory");
      }
      __syncwarp();
    }
    if (lane == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[3] = (long long)(g1 - g0); }
  } else if (warp <= sts_warps) {
    // concurrent store traffic into a region the MMAs do not read (last 32 KB of the 192 KB)
    float4* dst = reinterpret_cast<float4*>(sm + 160 * 1024) + (warp - 1) * 32 + lane;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    long long tstart = clock64();
    long long nst = 0;
    // unpaced: sts_per_iter > 0 => store as fast as possible for a fixed count, report own duration
    const int total = 20000;
    for (int it = 0; it < total; it += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(dst + j * 256)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
      if (sts_per_iter < 100) { long long t = clock64(); while (clock64() - t < sts_per_iter) {} }
    }
    nst = total;
    long long tend = clock64();
    if (lane == 0 && blockIdx.x == 0 && warp == 1) out[2] = tend - tstart;
    if (lane == 0 && blockIdx.x == 0 && warp == 1) out[1] = nst * 512 * sts_warps;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE, int N>
void run(const char* name, int sts_warps, int sts_per_iter, int iters = 6000, int variant = 0) {
  long long* d;
  cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  const int smem = 192 * 1024 + 2048;
  cudaFuncSetAttribute(bench<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) bench<MODE, N><<<148, 288, smem>>>(iters, sts_warps, sts_per_iter, d, variant);
  cudaError_t e = cudaDeviceSynchronize();
  long long hh[4] = {0, 0, 0, 0};
  cudaMemcpy(hh, d, 32, cudaMemcpyDeviceToHost);
  long long h = hh[0];
  const double cyc = (double)h / (iters * 12.0);
  const double kper = (MODE >= 3) ? 16 : 8;
  printf("%-34s N=%3d sts_warps=%d x%2d : %7.1f cyc/MMA  -> %6.0f FLOP/cyc/SM  (%s)\n", name, N, sts_warps, sts_per_iter, cyc,
         2.0 * 128 * N * kper / cyc, cudaGetErrorString(e));
  printf("      %lld cycles in %lld ns -> %.0f MHz effective SM clock\n", hh[0], hh[3], 1e3 * (double)hh[0] / (double)hh[3]);
  if (sts_warps) printf("      concurrent STS: %.0f B per MMA-time (%.1f B/cyc)\n", (double)hh[1] / (double)hh[2] * 64.0, (double)hh[1] / (double)hh[2]);
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1) {                                           // N = 128 vs 256, kind::f16 and kind::tf32, with concurrent store traffic
    run<3, 128>("f16 SS", 0, 0);
    run<3, 256>("f16 SS", 0, 0);
    run<3, 128>("f16 SS + STS", 4, 0);
    run<3, 256>("f16 SS + STS", 4, 0);
    run<3, 128>("f16 SS + paced STS", 4, 40);
    run<3, 256>("f16 SS + paced STS", 4, 40);
    run<0, 128>("tf32 SS", 0, 0);
    run<0, 256>("tf32 SS", 0, 0);
    run<0, 128>("tf32 SS + STS", 4, 0);
    run<0, 256>("tf32 SS + STS", 4, 0);
    run<0, 128>("tf32 SS + paced STS", 4, 40);
    run<0, 256>("tf32 SS + paced STS", 4, 40);
    run<2, 128>("tf32 SS MN-major (weight gradient)", 0, 0);
    run<2, 128>("tf32 SS MN-major + STS", 4, 0);
    return 0;
  }
  run<0, 128>("plain", 0, 0, 6000, 0);
  run<0, 128>("+commit per 12", 0, 0, 6000, 1);
  run<0, 128>("+fence+elect", 0, 0, 6000, 2);
  run<0, 128>("+commit+fence+elect", 0, 0, 6000, 3);
  run<0, 128>("+2 accumulators (first=0 every 48)", 0, 0, 6000, 4);
  run<0, 128>("+2 acc + commit", 0, 0, 6000, 5);
  run<0, 128>("all (commit, fence, elect, 2 acc, chunk commit)", 0, 0, 6000, 15);
  return 0;
}
