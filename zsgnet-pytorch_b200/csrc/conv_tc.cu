// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05, accumulators in TMEM).
//
// One kernel family serves every dense contraction of the hot path: ResNet-50 / FPN / SSD-VGG / head
// convolutions forward, their data gradients (the same kernel over flipped-transposed weights),
// the LSTM input projection (a 1x1 conv over B*T rows) and, in the wgrad kernel, every weight
// gradient.  Geometry comes from a 16-byte row table (zsg_row_t) plus a tap spacing, so strides, padding,
// dilation, the stride-2 data gradient and the six-level shared head are data, not code.  The epilogue can
// emit the BatchNorm statistics of its output (zsg_conv_params.stats) so that y is not read again for them.
//
// Arithmetic: fp32 in HBM; D += Ah*Bh + Ah*Bl + Al*Bh is issued as three kind::tf32 MMAs (3xTF32): fp32-accurate
// products, fp32 accumulation in TMEM, promoted to registers every few K blocks.  This is what lets the fp32
// configuration meet the reference's 1e-4 tolerance while still running on the tensor pipe.  The tensor core
// reads a raw fp32 word as its TF32 truncation, so a tensor is its own high part; the low part is a second
// image (lo = v - trunc(v)).
//
// Kernels in this file (all share the tiling, the barrier protocol, the two MMA-issuer warps and the drain):
//   conv_tc_async_kernel<BN>          product path: input + remainder image by cp.async, weight images by TMA
//   wgrad_tc_async_kernel<BN,TMA_DY>  product path: x images by cp.async, dy images by TMA (or cp.async)
//   conv_tc_kernel<BN,PRO>            register path: gather -> BatchNorm affine / ReLU -> split in registers -> smem
//   conv_tc_generic_kernel<BN>        register path with weights gathered as well (no pre-split weight images)
//   wgrad_tc_kernel<BN>               register path of the weight gradient
//   conv_simt_kernel / wgrad_simt_kernel   plain-FMA check kernels (tests only)
// Stages are handed over with mbarriers: full[s] (producer arrivals / cp.async completions / TMA bytes),
// empty[s] and acc_full[a] (tcgen05.commit), acc_empty[a] (256 drain arrivals), token[w] (issuer hand-over).
// Warp roles and the chunked-promotion scheme are described above setup_pipeline().
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "common.cuh"

namespace zsg {

constexpr int TM = 128;             // tile rows = TMEM lanes
constexpr int KB = 32;              // fp32 K elements per stage = one 128-byte swizzle row
constexpr int NPROD = 128;          // producer / epilogue threads
constexpr int A_TILE_BYTES = TM * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU.  Waiting warps back off with
// nanosleep: 18 warps share 4 schedulers and up to 10 of them wait at any time; spinning on try_wait took the issue
// slots of the warps that had work (the epilogue of a tile ran at 10 cycles per instruction).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
      printf("zsg conv: mbarrier wait timed out (block %d,%d,%d thread %d, wait site %d, parity %u)\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// K-major, 128-byte swizzle, 8-row groups 1024 B apart (SBO); LBO unused for swizzled K-major.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n, bool mn_major = false) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load WITHOUT the wait: several can be in flight, one tmem_ld_wait() covers them.  With eight warps draining
// at once a load-and-wait costs ~230 cycles (B300_MICROARCH: LDTM 8-warp effective latency), so the four x16 loads of a
// 64-column accumulator half, each waited for, were ~1.2 k of the ~5 k cycles a 1-K-block tile takes (tools/trace_plain.py).
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
        "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- TMA (cp.async.bulk.tensor) for the weight operand --------------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// TF32-exact high part and fp32 residual (both representable: hi + lo == v exactly).
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = v - hi;
}
__device__ __forceinline__ void store_split(uint8_t* tile_hi, uint8_t* tile_lo, int row, int chunk, float4 v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x);
  split_tf32(v.y, h.y, l.y);
  split_tf32(v.z, h.z, l.z);
  split_tf32(v.w, h.w, l.w);
  const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
  *reinterpret_cast<float4*>(tile_hi + off) = h;
  *reinterpret_cast<float4*>(tile_lo + off) = l;
}

// --------------------------------------------------------------------------------------------
// CTA layout (18 warps):
//   warps 0-3, 4-7   two producer groups; group g fills the stages of K-blocks kb == g (mod 2), so
//                    twice as many gather loads are in flight and neither group waits on the other
//   warps 8-15       drain + epilogue: quadrant (warp & 3) = TMEM lanes, (warp - 8) >> 2 = column half
//   warps 16, 17     MMA issue, one elected thread each, taking turns K block by K block (16 also owns TMEM)
//
// Chunked promotion: the tensor core adds every MMA into the fp32 TMEM accumulator with truncation,
// which biases long reductions toward zero (measured -1.4e-5 relative at K = 2304).  So TMEM only
// ever accumulates CHUNK_KB K-blocks; the drain warps pull each finished chunk out of TMEM and add
// it into fp32 registers with round-to-nearest while the MMAs of the next chunk run into the
// second TMEM accumulator.  The bias no longer grows with K, and the epilogue already holds the
// result in registers.
// --------------------------------------------------------------------------------------------
constexpr int NGROUP = 2;            // producer groups
constexpr int CHUNK_KB = 8;          // K-blocks per chunk: each of the two issuers accumulates its (up to) 4 in its own TMEM
                                     // accumulator before the drain warps promote both into fp32 registers
#ifndef ZSG_CHUNK_KB_BF16
#define ZSG_CHUNK_KB_BF16 8
#endif
template <bool BF16> struct ChunkKB { static constexpr int V = BF16 ? ZSG_CHUNK_KB_BF16 : CHUNK_KB; };
constexpr int NDRAIN = 256;          // drain / epilogue threads
constexpr int DRAIN_WARP0 = 8;
constexpr int MMA_WARP = 16;           // issuer warps: 16 and 17 (TMEM is allocated / freed by 16)
constexpr int NTHREADS2 = 18 * 32;

// BF16 = the bf16 operand path (BASELINE configs 3-5): ONE image per operand (the bf16 copy), 64 elements per
// 128-byte swizzle row, one kind::f16 MMA per product instead of three kind::tf32 ones.  A stage is then half the
// bytes for twice the K extent, so the ring is deeper.
template <int BN, bool BF16 = false>
struct Smem {
  static constexpr int BN_ = BN;
  static constexpr bool BF16_ = BF16;
  static constexpr int NIMG = BF16 ? 1 : 2;                 // operand images per tile (bf16 copy | TF32 hi + lo)
  static constexpr int KBE = BF16 ? 64 : 32;                // K elements per stage (one 128-byte swizzle row)
  static constexpr int ES = BF16 ? 2 : 4;                   // operand element size in bytes
  static constexpr int B_TILE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = NIMG * (A_TILE_BYTES + B_TILE_BYTES);
  static constexpr int STAGES = BF16 ? (BN > 128 ? 4 : 6) : (BN >= 128 ? 3 : 4);
  static constexpr int TILES_BYTES = STAGES * STAGE_BYTES;
  static constexpr int ROWS_OFF = TILES_BYTES;              // TM row entries (fwd) / 2 groups x 2 x 32|64 (wgrad)
  static constexpr int BAR_OFF = ROWS_OFF + NGROUP * TM * 16;
  static constexpr int EPI_OFF = BAR_OFF + 256;             // epilogue staging: 8 warps x 32 rows x 20 floats
  static constexpr int EPI_WARP_BYTES = 32 * 20 * 4;
  static constexpr int TOTAL = EPI_OFF + 8 * EPI_WARP_BYTES + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = BN > 128 ? 512 : 4 * BN;   // (chunk parity) x (issuer) accumulators; BN = 256: two, by tile parity
};

constexpr int MAX_STAGES = 8;
constexpr int TMEM_SLOT_OFF = 192;     // byte offset of the TMEM base-address word inside the barrier block
struct PipeBars {                      // mbarrier addresses are computed, never indexed from memory
  uint32_t bar0;
  __device__ __forceinline__ uint32_t full(int s) const { return bar0 + 8 * s; }
  __device__ __forceinline__ uint32_t empty(int s) const { return bar0 + 64 + 8 * s; }
  __device__ __forceinline__ uint32_t acc_full(int a) const { return bar0 + 128 + 8 * a; }
  __device__ __forceinline__ uint32_t acc_empty(int a) const { return bar0 + 160 + 8 * a; }
  __device__ __forceinline__ uint32_t tmem_slot() const { return bar0 + TMEM_SLOT_OFF; }
  __device__ __forceinline__ uint32_t token(uint32_t w) const { return bar0 + 200 + 8 * w; }
};

// Diagnostics (tools/trace_conv.py): when a trace buffer is registered, CTA 0 records clock timestamps of the
// pipeline hand-overs of its first K blocks: trace[gk * 16 + event].
__device__ unsigned int* g_trace = nullptr;
__device__ int g_trace_blocks = 0;
__device__ __forceinline__ void trace(int gk, int ev) {
#ifdef ZSG_TRACE                       // build.py --trace; the stamps cost ~30 cycles each even when no buffer is registered
  if (g_trace != nullptr && blockIdx.x == 0 && gk < g_trace_blocks) g_trace[gk * 16 + ev] = (unsigned int)clock64();
#endif
}

template <int BN, bool BF16 = false>
__device__ __forceinline__ PipeBars setup_pipeline(uint8_t* sm, int warp, int lane, int full_count = NPROD) {
  using S = Smem<BN, BF16>;
  static_assert(S::STAGES <= MAX_STAGES, "barrier block holds MAX_STAGES full/empty pairs");
  PipeBars pb;
  pb.bar0 = smem_u32(sm + S::BAR_OFF);
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < S::STAGES; ++i) { mbar_init(pb.full(i), full_count); mbar_init(pb.empty(i), 1); }
      for (int i = 0; i < 4; ++i) { mbar_init(pb.acc_full(i), 1); mbar_init(pb.acc_empty(i), NDRAIN); }
      for (int i = 0; i < 2; ++i) mbar_init(pb.token(i), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(pb.tmem_slot(), S::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return pb;
}

// MMA issue (3xTF32: lo*hi + hi*lo + hi*hi), one TMEM accumulator per chunk.
//
// The tensor pipe only queues about two MMAs (measured: every UTCHMMA of a K block stalls equally on issue, and
// the pipe idled 42 % of the time while this warp spent ~700 cycles per K block outside the MMA issue stall), so
// everything this warp executes between the last MMA of one K block and the first of the next is exposed.  Hence:
// ring position and phases are running counters (no division), the four descriptors of a stage differ from one
// base word by constants, and the twelve MMAs of a K block are issued from ONE asm block whose only other
// instructions are the 32-bit adds that advance the descriptor start addresses.
//
// MN_MAJOR (weight-gradient kernel): both operands are stored as they lie in memory, [pixel][channel].  For 32-bit
// operands the only MN-major layout the tensor core accepts is SWIZZLE_128B_BASE32B (layout type 1): atoms of
// 4 pixel rows x 128 B (32 channels) in which the 32-byte granule g of row r is stored at granule g ^ r (byte-address
// bits [5,7) ^= bits [7,9)).  Tile = [channel atom][pixel group of 4][4][128 B]: LBO = 4096 B between channel atoms, SBO = 512 B
// between pixel groups; one MMA (K = 8) consumes two pixel groups (1024 B).
// Two issuing warps.  Measured on B200 (tools/micro/mma_pingpong.cu): the tensor pipe queues only about two MMAs, so
// with ONE issuing thread every instruction between the last MMA of a K block and the first of the next (barrier
// waits, fences, descriptor set-up: ~50 SASS instructions) is exposed -- a single issuer with that much work per
// 12 MMAs reaches 71 % of the 64-cycle/MMA floor, two issuers reach 100 %.  So warps 16 and 17 each elect one
// thread; thread w owns the K blocks with (global index) % 2 == w and accumulates them into ITS OWN TMEM
// accumulator (chunk parity x issuer = 4 accumulators), so the two never touch the same accumulator or stage and
// results do not depend on how the pipe interleaves the tail of one K block with the head of the next (with a
// shared accumulator that made the fp32 sums differ from run to run).  They still take turns through a token
// (mbarrier) so that only one thread issues at a time: issuing from both concurrently hung the pipe now and then.
// Issuer w owns the K blocks g with (g / G) % 2 == w.  G = 1: strict alternation.  Measured (tools/time_big.py, bf16, where a K
// block is only 4 MMAs = 256 cycles): G = 1, 2 and 4 give the same 670 / 530 TFLOP/s on the 3x3 256->256 forward / weight
// gradient -- the hand-over between the two issuing threads is not what limits the bf16 kernels.
#ifndef ZSG_ISSUE_GROUP_BF16
#define ZSG_ISSUE_GROUP_BF16 1
#endif
#ifndef ZSG_ISSUE_GROUP_FP32
#define ZSG_ISSUE_GROUP_FP32 1
#endif
template <bool BF16> struct IssueGroup { static constexpr uint32_t G = BF16 ? ZSG_ISSUE_GROUP_BF16 : ZSG_ISSUE_GROUP_FP32; };
template <bool BF16>
__device__ __forceinline__ int issuer_count(int g0, int n, int w) {     // K blocks g0 .. g0 + n - 1 owned by issuer w
  constexpr int G = (int)IssueGroup<BF16>::G;
  int c = 0;
  for (int k = 0; k < n; ++k) c += (((g0 + k) / G) & 1) == w;
  return c;
}
struct Issuer {
  uint32_t w;                // 0 / 1
  uint32_t stage;            // ring position of K block g
  uint32_t phase;            // parity of full[stage] for K block g
  uint32_t tokens = 0;       // tokens consumed so far (parity of my token barrier)
  uint32_t g = 0;            // global K-block index over all tiles of this CTA
  uint32_t chunk = 0;        // global chunk index (selects the accumulator pair and its parity)
};

template <int BN, bool MN_MAJOR>
__device__ __forceinline__ void issue_kblock(uint32_t tmem_d, uint32_t a_hi_lo, uint32_t desc_hi, uint32_t acc_first) {
  constexpr uint32_t idesc = umma_idesc_tf32(BN, MN_MAJOR);
  constexpr uint32_t A16 = A_TILE_BYTES >> 4, B16 = (BN * 128) >> 4, KSTEP = MN_MAJOR ? 64u : 2u;
  asm volatile(
      "{\n\t"
      ".reg .pred p, pt;\n\t"
      ".reg .b32 a0, a1, b0, b1;\n\t"
      ".reg .b64 dah, dal, dbh, dbl;\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b32 a0, %1;\n\t"
      "add.u32 a1, a0, %5;\n\t"
      "add.u32 b0, a1, %5;\n\t"
      "add.u32 b1, b0, %6;\n\t"
      "mov.b64 dah, {a0, %2};\n\tmov.b64 dal, {a1, %2};\n\tmov.b64 dbh, {b0, %2};\n\tmov.b64 dbl, {b1, %2};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dal, dbh, %4, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbl, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbh, %4, pt;\n\t"
      "add.u32 a0, a0, %7;\n\tadd.u32 a1, a1, %7;\n\tadd.u32 b0, b0, %7;\n\tadd.u32 b1, b1, %7;\n\t"
      "mov.b64 dah, {a0, %2};\n\tmov.b64 dal, {a1, %2};\n\tmov.b64 dbh, {b0, %2};\n\tmov.b64 dbl, {b1, %2};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dal, dbh, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbl, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbh, %4, pt;\n\t"
      "add.u32 a0, a0, %7;\n\tadd.u32 a1, a1, %7;\n\tadd.u32 b0, b0, %7;\n\tadd.u32 b1, b1, %7;\n\t"
      "mov.b64 dah, {a0, %2};\n\tmov.b64 dal, {a1, %2};\n\tmov.b64 dbh, {b0, %2};\n\tmov.b64 dbl, {b1, %2};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dal, dbh, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbl, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbh, %4, pt;\n\t"
      "add.u32 a0, a0, %7;\n\tadd.u32 a1, a1, %7;\n\tadd.u32 b0, b0, %7;\n\tadd.u32 b1, b1, %7;\n\t"
      "mov.b64 dah, {a0, %2};\n\tmov.b64 dal, {a1, %2};\n\tmov.b64 dbh, {b0, %2};\n\tmov.b64 dbl, {b1, %2};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dal, dbh, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbl, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbh, %4, pt;\n\t"
      "}"
      ::"r"(tmem_d), "r"(a_hi_lo), "r"(desc_hi), "r"(acc_first), "n"(idesc), "n"(A16), "n"(B16), "n"(KSTEP)
      : "memory");
}

// kind::f16 with bf16 operands, fp32 accumulate, M = 128 (instruction-descriptor fields as in kind::tf32: bits 4-5 D format
// 1 = f32, 7-9 / 10-12 A / B format 1 = bf16, 15 / 16 A / B MN-major, 17-22 N >> 3, 24-28 M >> 4).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n, bool mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TM >> 4) << 24);
}
// bf16 K block = 64 K elements = four MMAs of K = 16.  K-major tiles (forward / data gradient): an MMA consumes 32 bytes
// of every 128-byte swizzle row, start address += 2 (x 16 B).  MN-major tiles (weight gradient; [channel atom of 64]
// [pixel][128 B], plain 128-byte swizzle = what TMA SWIZZLE_128B and the cp.async producers write): an MMA consumes 16
// pixels = 2048 B, start address += 128.
template <int BN, bool MN_MAJOR>
__device__ __forceinline__ void issue_kblock_bf16(uint32_t tmem_d, uint32_t a_lo32, uint32_t desc_hi, uint32_t acc_first) {
  constexpr uint32_t idesc = umma_idesc_bf16(BN, MN_MAJOR);
  constexpr uint32_t A16 = A_TILE_BYTES >> 4, KSTEP = MN_MAJOR ? 128u : 2u;
  asm volatile(
      "{\n\t"
      ".reg .pred p, pt;\n\t"
      ".reg .b32 a0, b0;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b32 a0, %1;\n\t"
      "add.u32 b0, a0, %5;\n\t"
      "mov.b64 da, {a0, %2};\n\tmov.b64 db, {b0, %2};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "add.u32 a0, a0, %6;\n\tadd.u32 b0, b0, %6;\n\t"
      "mov.b64 da, {a0, %2};\n\tmov.b64 db, {b0, %2};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.u32 a0, a0, %6;\n\tadd.u32 b0, b0, %6;\n\t"
      "mov.b64 da, {a0, %2};\n\tmov.b64 db, {b0, %2};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.u32 a0, a0, %6;\n\tadd.u32 b0, b0, %6;\n\t"
      "mov.b64 da, {a0, %2};\n\tmov.b64 db, {b0, %2};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "}"
      ::"r"(tmem_d), "r"(a_lo32), "r"(desc_hi), "r"(acc_first), "n"(idesc), "n"(A16), "n"(KSTEP)
      : "memory");
}

template <int BN, bool BF16 = false>
__device__ __forceinline__ Issuer issuer_init(uint32_t w) {
  Issuer is;
  is.w = w;
  is.stage = 0;
  is.phase = 0;
  return is;
}

// Executed by the elected thread of issuer warp `is.w` for the K blocks of one tile (both issuers walk all K blocks
// to keep the chunk bookkeeping, each acts on its own).  A chunk never spans two tiles.
template <int BN, bool MN_MAJOR = false, bool BF16 = false>
__device__ __forceinline__ void mma_loop(uint8_t* sm, const PipeBars& pb, uint32_t tmem_base, int nkb, Issuer& is,
                                         bool swap_lbo_sbo = false, int ablate = 0) {
  using S = Smem<BN, BF16>;
  // MN-major strides (x 16 B).  tf32: atoms of 4 pixels x 128 B, 4096 B between channel atoms (SWIZZLE_128B_BASE32B).
  // bf16: atoms of 8 pixels x 128 B (64 channels) = 1024 B (SBO), 64 pixels x 128 B = 8192 B between channel atoms (LBO),
  // plain SWIZZLE_128B (layout type 2), as in the canonical ((8,n),(8,k)):((1,LBO),(8,SBO)) uint128 layout.
  const uint32_t lbo = MN_MAJOR ? (BF16 ? 512u : (swap_lbo_sbo ? 32u : 256u)) : 1u;
  const uint32_t sbo = MN_MAJOR ? (BF16 ? 64u : (swap_lbo_sbo ? 256u : 32u)) : 64u;
  const uint32_t desc_hi = sbo | (1u << 14) | (((MN_MAJOR && !BF16) ? 1u : 2u) << 29);   // bits 32..63 of the descriptor
  const uint32_t lo0 = ((smem_u32(sm) >> 4) & 0x3FFFu) | (lbo << 16);                // a_hi tile of stage 0
  uint32_t mine = 0;                                        // my K blocks so far in the current chunk
  for (int kb = 0; kb < nkb; ++kb, ++is.g) {
    const uint32_t acc = (is.chunk & 1u) * 2u + is.w;
    // Opening a chunk: my accumulator of this parity must have been drained (two chunks ago).  BOTH issuers wait, whether or
    // not they own a K block of the chunk: each commits acc_full for its accumulator at the end of every chunk, and an
    // issuer with nothing to do in a run of chunks (single-K-block tiles: it owns every other tile, always with the same
    // accumulator) would otherwise complete acc_full phases faster than the drain warps consume them -- their parity wait
    // then misses a phase and hangs (seen on the bf16 path, where K = 64 is ONE K block per tile).
    if (kb % ChunkKB<BF16>::V == 0) mbar_wait(pb.acc_empty(acc), ((is.chunk >> 1) & 1u) ^ 1u, 100 + is.g);
    constexpr uint32_t G = IssueGroup<BF16>::G;
    if (((is.g / G) & 1u) == is.w) {
      trace(is.g, 8);
      if (!(ablate & 1)) mbar_wait(pb.full(is.stage), is.phase, 1000 + is.g);     // ablate bit 0 (diagnostics): producers are off
      trace(is.g, 9);
      fence_proxy_async();                                  // cp.async-filled tiles (generic proxy) -> tensor core (async proxy)
      if (is.g > 0 && is.g % G == 0 && !(ablate & 8)) {    // first K block of my group: my turn
#ifdef ZSG_TOKEN_SPIN
        while (!mbar_try_wait(pb.token(is.w), is.tokens & 1u)) {}
#else
        mbar_wait(pb.token(is.w), is.tokens & 1u, 4000 + is.g);
#endif
        ++is.tokens;
      }
      trace(is.g, 10);
      tc_fence_after();
      if (BF16)
        issue_kblock_bf16<BN, MN_MAJOR>(tmem_base + acc * BN, lo0 + is.stage * (uint32_t)(S::STAGE_BYTES >> 4), desc_hi,
                                        mine == 0 ? 0u : 1u);
      else
        issue_kblock<BN, MN_MAJOR>(tmem_base + acc * BN, lo0 + is.stage * (uint32_t)(S::STAGE_BYTES >> 4), desc_hi,
                                   mine == 0 ? 0u : 1u);
      if (is.g % G == G - 1 && !(ablate & 8)) {             // last K block of my group: the other issuer may go
        tc_fence_before();
        mbar_arrive(pb.token(is.w ^ 1u));
      }
      umma_commit(pb.empty(is.stage));                      // frees the stage once the MMAs above have read it
      trace(is.g, 11);
      ++mine;
    }
    if (++is.stage == (uint32_t)S::STAGES) { is.stage = 0; is.phase ^= 1u; }
    if (kb % ChunkKB<BF16>::V == ChunkKB<BF16>::V - 1 || kb == nkb - 1) {   // chunk closed: my part of it (possibly empty) is complete
      umma_commit(pb.acc_full(acc));
      ++is.chunk;
      mine = 0;
    }
  }
}

// drain warps: promote every finished TMEM chunk into fp32 registers (round-to-nearest adds), issuer 0's part first.
// gkb0 = global index of the tile's first K block (decides which issuer owns which K block of the chunk).
template <int BN, bool BF16 = false>
__device__ __forceinline__ void drain_loop(const PipeBars& pb, uint32_t tmem_base, int nkb, int gkb0, int quadrant, int half,
                                           float (&acc)[BN / 2], int& gchunk, int ablate = 0) {
  // The first accumulator of a tile that holds anything is loaded straight into `acc` (all of its x16 loads in flight, one
  // wait); later ones go through 16 temporaries, one load at a time.
  bool fresh = true;
  const int nchunks = (nkb + ChunkKB<BF16>::V - 1) / ChunkKB<BF16>::V;
  for (int cc = 0; cc < nchunks; ++cc, ++gchunk) {
    const int c = gchunk;
    const int k0 = cc * ChunkKB<BF16>::V;
    const int n = (nkb - k0 < ChunkKB<BF16>::V) ? nkb - k0 : ChunkKB<BF16>::V;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int a = (c & 1) * 2 + w;
      const int count = issuer_count<BF16>(gkb0 + k0, n, w);           // K blocks issuer w put into its accumulator
      mbar_wait(pb.acc_full(a), (c >> 1) & 1, 2000 + c * 2 + w);
      tc_fence_after();
      if (count > 0 && !(ablate & 2)) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quadrant * 32) << 16) + (uint32_t)a * BN + half * (BN / 2);
        if (fresh) {
#pragma unroll
          for (int cb = 0; cb < BN / 32; ++cb) tmem_ld16_nowait(taddr + cb * 16, acc + cb * 16);
          tmem_ld_wait();
          fresh = false;
        } else {
#pragma unroll
          for (int cb = 0; cb < BN / 32; ++cb) {            // 16 temporaries: 32 spilled in the generic-epilogue kernels
            uint32_t r[16];
            tmem_ld16(taddr + cb * 16, r);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[cb * 16 + j] += __uint_as_float(r[j]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(pb.acc_empty(a));
    }
  }
  if (fresh) {                                              // nothing was accumulated (diagnostic ablations only)
#pragma unroll
    for (int i = 0; i < BN / 2; ++i) acc[i] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Register ("fragment") epilogue of the plain / bf16-storage kernels.
//
// tcgen05.ld.16x256b hands a warp its accumulator in the layout of an MMA C fragment (verified on B200,
// tools/micro/tmem_layout.cu): register 4u + 2hh + c of lane t of the load at TMEM-lane offset 16h holds
//     row 16h + 8hh + t/4,   column 8u + 2(t%4) + c        (of the warp's 32 rows x CW columns)
// so the four lanes of a quad cover 8 consecutive columns (32 bytes) of one row and a warp instruction covers 8 rows with
// full sectors: the tile goes registers -> global with NO transposition through shared memory.  The slab epilogue
// (32x32b loads, row per lane) spent ~1200 instructions per warp and tile on STS / syncwarp / LDS / address work and ran
// at about one instruction per 4 cycles per warp: 5.2 k cycles per 128x128 tile against 0.25 k cycles of MMA on the
// 1-K-block tiles (tools/ablate_epilogue.py, tools/trace_plain.py).  Here: ~450.
// BatchNorm statistics: per column the lane's four rows are added in registers, then a halving butterfly over the 8 lanes
// of a column group (xor 16, 8, 4: each step sends half of the values) leaves every lane with one column pair's sum.
// ---------------------------------------------------------------------------------------------------------------
template <int U>
__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, float* r);
template <>
__device__ __forceinline__ void tmem_ld_frag<8>(uint32_t taddr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
        "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]),
        "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]),
        "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_frag<4>(uint32_t taddr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
        "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
      : "r"(taddr)
      : "memory");
}

// drain_loop in the fragment layout: acc[h * 4U + 4u + 2hh + c]
template <int BN, bool BF16 = false>
__device__ __forceinline__ void drain_loop_frag(const PipeBars& pb, uint32_t tmem_base, int nkb, int gkb0, int quadrant, int half,
                                                float (&acc)[BN / 2], int& gchunk, int ablate = 0) {
  constexpr int U = BN / 16;                                // 8-column units of the warp's BN / 2 columns
  bool fresh = true;
  const int nchunks = (nkb + ChunkKB<BF16>::V - 1) / ChunkKB<BF16>::V;
  for (int cc = 0; cc < nchunks; ++cc, ++gchunk) {
    const int c = gchunk;
    const int k0 = cc * ChunkKB<BF16>::V;
    const int n = (nkb - k0 < ChunkKB<BF16>::V) ? nkb - k0 : ChunkKB<BF16>::V;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int a = (c & 1) * 2 + w;
      const int count = issuer_count<BF16>(gkb0 + k0, n, w);
      mbar_wait(pb.acc_full(a), (c >> 1) & 1, 2000 + c * 2 + w);
      tc_fence_after();
      if (count > 0 && !(ablate & 2)) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quadrant * 32) << 16) + (uint32_t)a * BN + half * (BN / 2);
        if (fresh) {                                        // both lane halves in flight, one wait
          tmem_ld_frag<U>(taddr, acc);
          tmem_ld_frag<U>(taddr + (16u << 16), acc + 4 * U);
          tmem_ld_wait();
          fresh = false;
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float r[4 * U];
            tmem_ld_frag<U>(taddr + ((uint32_t)(16 * h) << 16), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4 * U; ++j) acc[h * 4 * U + j] += r[j];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(pb.acc_empty(a));
    }
  }
  if (fresh) {
#pragma unroll
    for (int i = 0; i < BN / 2; ++i) acc[i] = 0.f;
  }
}

// EPI_PLAIN / EPI_B16 epilogue of one warp's 32 rows x BN/2 columns, straight from the fragment registers
template <int BN, bool B16>
__device__ __forceinline__ void epilogue_frag(const zsg_conv_params& p, const float (&acc)[BN / 2], int m0, int n0, int quadrant,
                                              int half, int lane, int ablate) {
  constexpr int U = BN / 16;
  const int q = lane & 3, g = lane >> 2;                    // column pair inside a unit, row inside an 8-row group
  const int nbase = n0 + half * (BN / 2) + 2 * q;           // column of (u = 0, c = 0)
  // ---- BatchNorm statistics of the warp's 32 rows (rows past p.m hold zeros: their A rows were zero-filled)
  //   forward (p.stats):        s1 = sum y,  s2 = sum y^2
  //   backward (p.bnb_partials): this launch is the data gradient that produces dy of a BatchNorm+ReLU whose INPUT x lies at
  //                             the output positions (p.bnb_x): s1 = sum dz, s2 = sum dz * x with dz = dy where
  //                             x * scale + shift > 0 -- the reduce pass of that BatchNorm's backward (zsg_bn_bwd_reduce,
  //                             mask_mode 1) as a by-product; y itself is stored unmasked, as without it
  float* const stp = p.stats ? p.stats : p.bnb_partials;
  if (stp && !(ablate & 128)) {
    float s1[2 * U], s2[2 * U];
    if (p.stats) {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float a0 = acc[4 * u + c], a1 = acc[4 * u + 2 + c], a2 = acc[4 * U + 4 * u + c], a3 = acc[4 * U + 4 * u + 2 + c];
          s1[2 * u + c] = (a0 + a1) + (a2 + a3);
          s2[2 * u + c] = fmaf(a0, a0, a1 * a1) + fmaf(a2, a2, a3 * a3);
        }
    } else {
      // unit by unit: the unit's scale / shift pair once, its four rows' x pairs in flight together
      int64_t roff[4];
      bool rok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {                         // i = 2h + hh
        const int r = m0 + quadrant * 32 + 16 * (i >> 1) + 8 * (i & 1) + g;
        rok[i] = r < p.m;                                   // (rows past m: zero accumulators, and nothing to read)
        roff[i] = (int64_t)r * p.y_pitch + nbase;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        s1[2 * u] = s1[2 * u + 1] = s2[2 * u] = s2[2 * u + 1] = 0.f;
        if (nbase + 8 * u >= p.cout) continue;
        const float2 sc = __ldg(reinterpret_cast<const float2*>(p.bnb_scale + nbase + 8 * u));
        const float2 sh = __ldg(reinterpret_cast<const float2*>(p.bnb_shift + nbase + 8 * u));
        float2 xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xv[i] = make_float2(0.f, 0.f);
          if (rok[i]) {
            if (B16) {
              const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(p.bnb_x) + roff[i] + 8 * u);
              xv[i] = make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
            } else {
              xv[i] = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(p.bnb_x) + roff[i] + 8 * u);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v0 = acc[(i >> 1) * 4 * U + 4 * u + 2 * (i & 1)], v1 = acc[(i >> 1) * 4 * U + 4 * u + 2 * (i & 1) + 1];
          if (B16) {                                        // the stored gradient is the bfloat16 rounding: sum what is stored
            v0 = __bfloat162float(__float2bfloat16_rn(v0));
            v1 = __bfloat162float(__float2bfloat16_rn(v1));
          }
          v0 = fmaf(xv[i].x, sc.x, sh.x) > 0.f ? v0 : 0.f;
          v1 = fmaf(xv[i].y, sc.y, sh.y) > 0.f ? v1 : 0.f;
          s1[2 * u] += v0;
          s1[2 * u + 1] += v1;
          s2[2 * u] = fmaf(v0, xv[i].x, s2[2 * u]);
          s2[2 * u + 1] = fmaf(v1, xv[i].y, s2[2 * u + 1]);
        }
      }
    }
    // halving butterfly over the 8 lanes of a column group: lane bit 4 keeps the upper / lower half of the values, ...
    int keep = 2 * U;                                       // number of values still held; they sit in s*[0 .. keep)
#pragma unroll
    for (int bit = 16; bit >= 4; bit >>= 1) {
      if (keep >= 2) {
        const int hk = keep / 2;
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int j = 0; j < 2 * U; ++j) {
          if (j < hk) {
            const float send1 = up ? s1[j] : s1[j + hk], send2 = up ? s2[j] : s2[j + hk];
            const float r1 = __shfl_xor_sync(0xffffffffu, send1, bit), r2 = __shfl_xor_sync(0xffffffffu, send2, bit);
            s1[j] = (up ? s1[j + hk] : s1[j]) + r1;
            s2[j] = (up ? s2[j + hk] : s2[j]) + r2;
          }
        }
        keep = hk;
      } else {                                              // one value left: plain butterfly
        s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], bit);
        s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], bit);
      }
    }
    // values held now: k = kbase .. kbase + keep - 1 of the original 2U (k = 2u + c), kbase from the lane bits that chose halves
    float* st1 = stp + ((int64_t)((m0 / TM) * 4 + quadrant) * 2) * p.cout;
    float* st2 = st1 + p.cout;
    if (U == 8) {                                           // keep == 2: unit u = 4*b4 + 2*b3 + b2, both columns of the pair
      const int u = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      const int col = nbase + 8 * u;
      if (col < p.cout) {
        *reinterpret_cast<float2*>(st1 + col) = make_float2(s1[0], s1[1]);
        *reinterpret_cast<float2*>(st2 + col) = make_float2(s2[0], s2[1]);
      }
    } else {                                                // U == 4: keep == 1: k = 4*b4 + 2*b3 + b2 -> u = k / 2, c = k & 1
      const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      const int col = nbase + 8 * (k >> 1) + (k & 1);
      if (col < p.cout) { st1[col] = s1[0]; st2[col] = s2[0]; }
    }
  }
  if (ablate & 16) return;
  // ---- stores: four rows per lane (16h + 8hh + g), a column pair per unit
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int r = m0 + quadrant * 32 + 16 * h + 8 * hh + g;
      if (r >= p.m) continue;
      const int64_t off = (p.y_pitch > 0 ? (int64_t)r * p.y_pitch : (int64_t)__ldg(&p.rows[r].out)) + nbase;
      const bool al = (off & 1) == 0;                       // row offsets of the path are multiples of the channel count
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (nbase + 8 * u >= p.cout) break;
        const float v0 = acc[h * 4 * U + 4 * u + 2 * hh], v1 = acc[h * 4 * U + 4 * u + 2 * hh + 1];
        if (B16) {
          if (al) {
            const __nv_bfloat162 b = __floats2bfloat162_rn(v0, v1);
            *reinterpret_cast<__nv_bfloat162*>(p.y_bf16 + off + 8 * u) = b;
          } else {
            reinterpret_cast<__nv_bfloat16*>(p.y_bf16)[off + 8 * u] = __float2bfloat16_rn(v0);
            reinterpret_cast<__nv_bfloat16*>(p.y_bf16)[off + 8 * u + 1] = __float2bfloat16_rn(v1);
          }
        } else {
          if (al) *reinterpret_cast<float2*>(p.y + off + 8 * u) = make_float2(v0, v1);
          else { p.y[off + 8 * u] = v0; p.y[off + 8 * u + 1] = v1; }
        }
      }
    }
}

// EPI_FRAGX: the options of the generic epilogue (bias, row_add, ReLU mask, residual / residual_bf16, accumulate, ReLU; same
// order of operations, bit-identical results) straight from the fragment registers, for plain [m, y_pitch] outputs with cout % 8 == 0.
// The slab epilogue consumed its read operands two 16-byte loads at a time per lane (8 KB in flight per SM): the short-K data
// gradients that add into the gradient stream between blocks (y += ..., or + residual) ran at half of what HBM allows
// (10.2 k cycles per 128x128 tile against 5.3 k for its 128 KB).  Here a lane owns two rows x U column pairs per pass and
// issues all 2U loads of an operand before using any: 16 x 8 B x 32 lanes x 8 warps = 32 KB in flight.
template <int BN>
__device__ __forceinline__ void epilogue_fragx(const zsg_conv_params& p, float (&acc)[BN / 2], int m0, int n0, int quadrant,
                                               int half, int lane, int ablate) {
  constexpr int U = BN / 16;
  const int q = lane & 3, g = lane >> 2;
  const int nbase = n0 + half * (BN / 2) + 2 * q;
  if (ablate & 16) return;
#define VX(hh, u) acc[h * 4 * U + 4 * (u) + 2 * (hh)]
#define VY(hh, u) acc[h * 4 * U + 4 * (u) + 2 * (hh) + 1]
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int64_t off[2];                                         // y_pitch > 0 and even (pick_epilogue): every access is 8-byte aligned
    bool ok[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int r = m0 + quadrant * 32 + 16 * h + 8 * hh + g;
      ok[hh] = r < p.m;
      off[hh] = (int64_t)r * p.y_pitch + nbase;
    }
    if (p.bias) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (nbase + 8 * u >= p.cout) break;
        const float2 b = __ldg(reinterpret_cast<const float2*>(p.bias + nbase + 8 * u));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) { VX(hh, u) += b.x; VY(hh, u) += b.y; }
      }
    }
    if (p.row_add) {                                        // language + grid terms of the first head conv (L2-resident tables)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (!ok[hh]) continue;
        const int2 ra = __ldg(reinterpret_cast<const int2*>(p.row_add_idx) + (m0 + quadrant * 32 + 16 * h + 8 * hh + g));
        float2 t0[U], t1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (nbase + 8 * u >= p.cout) break;
          t0[u] = __ldg(reinterpret_cast<const float2*>(p.row_add + ra.x + nbase + 8 * u));
          t1[u] = __ldg(reinterpret_cast<const float2*>(p.row_add + ra.y + nbase + 8 * u));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (nbase + 8 * u >= p.cout) break;
          VX(hh, u) += t0[u].x + t1[u].x;
          VY(hh, u) += t0[u].y + t1[u].y;
        }
      }
    }
    if (p.out_mask) {
      float2 t[2][U];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          t[hh][u] = make_float2(1.f, 1.f);
          if (ok[hh] && nbase + 8 * u < p.cout) t[hh][u] = __ldg(reinterpret_cast<const float2*>(p.out_mask + off[hh] + 8 * u));
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!(t[hh][u].x > 0.f)) VX(hh, u) = 0.f;
          if (!(t[hh][u].y > 0.f)) VY(hh, u) = 0.f;
        }
    }
    if (p.residual) {
      float2 t[2][U];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          t[hh][u] = make_float2(0.f, 0.f);
          if (ok[hh] && nbase + 8 * u < p.cout) t[hh][u] = *reinterpret_cast<const float2*>(p.residual + off[hh] + 8 * u);
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int u = 0; u < U; ++u) { VX(hh, u) += t[hh][u].x; VY(hh, u) += t[hh][u].y; }
    }
    if (p.residual_bf16) {                                  // bf16 storage: the shortcut gradient is a bfloat16 tensor
      uint32_t t[2][U];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          t[hh][u] = 0u;
          if (ok[hh] && nbase + 8 * u < p.cout) t[hh][u] = *reinterpret_cast<const uint32_t*>(p.residual_bf16 + off[hh] + 8 * u);
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          VX(hh, u) += __uint_as_float(t[hh][u] << 16);
          VY(hh, u) += __uint_as_float(t[hh][u] & 0xFFFF0000u);
        }
    }
    if (p.accumulate) {
      float2 t[2][U];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          t[hh][u] = make_float2(0.f, 0.f);
          if (ok[hh] && nbase + 8 * u < p.cout) t[hh][u] = *reinterpret_cast<const float2*>(p.y + off[hh] + 8 * u);
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int u = 0; u < U; ++u) { VX(hh, u) += t[hh][u].x; VY(hh, u) += t[hh][u].y; }
    }
    if (p.out_relu) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int u = 0; u < U; ++u) { VX(hh, u) = fmaxf(VX(hh, u), 0.f); VY(hh, u) = fmaxf(VY(hh, u), 0.f); }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      if (!ok[hh]) continue;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (nbase + 8 * u >= p.cout) break;
        *reinterpret_cast<float2*>(p.y + off[hh] + 8 * u) = make_float2(VX(hh, u), VY(hh, u));
      }
      if (p.y_lo) {                                         // operand image of the output: TF32 remainders (as zsg_split_act)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (nbase + 8 * u >= p.cout) break;
          float h0, l0, h1, l1;
          split_tf32(VX(hh, u), h0, l0);
          split_tf32(VY(hh, u), h1, l1);
          *reinterpret_cast<float2*>(p.y_lo + off[hh] + 8 * u) = make_float2(l0, l1);
        }
      }
      if (p.y_img_bf16) {                                   // ... or the bfloat16 copy (as zsg_cast_bf16)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (nbase + 8 * u >= p.cout) break;
          *reinterpret_cast<__nv_bfloat162*>(p.y_img_bf16 + off[hh] + 8 * u) = __floats2bfloat162_rn(VX(hh, u), VY(hh, u));
        }
      }
    }
  }
#undef VX
#undef VY
}

// ragged rows (channel count or row offset not a multiple of 4: the [B, A, 5] head output): scalar, out of line.
// (Arguments by value: a reference to the kernel parameter block forces a local-memory copy of it, and with ~6 KB of
// L1 left every read of that copy is an L2 round trip.)
__device__ __noinline__ void epilogue_store_ragged(float* y, const float* out_mask, const float* residual, int accumulate,
                                                   int out_relu, int cout, int off_r, int n, float4 v4) {
  const float v[4] = {v4.x, v4.y, v4.z, v4.w};
  float* yrow = y + (int64_t)off_r + n;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    if (n + q >= cout) break;
    float u = v[q];
    if (out_mask && !(out_mask[(int64_t)off_r + n + q] > 0.f)) u = 0.f;
    if (residual) u += residual[(int64_t)off_r + n + q];
    if (accumulate) u += yrow[q];
    if (out_relu) u = fmaxf(u, 0.f);
    yrow[q] = u;
  }
}

// drain + epilogue warps of the forward / data-gradient kernels.
//
// EPI selects the epilogue at COMPILE time.  The slab loop is unrolled BN/32 times, so every option costs its code four
// times over, and the three warp roles of this kernel already fill the instruction cache: with the bf16-storage and
// bf16-residual options added as run-time branches the kernel grew from 3.3 k to 4.4 k SASS instructions and the big 3x3
// layers, which use neither, lost 13 % (measured).  Each product kernel now carries one of
//   EPI_PLAIN   y = acc (fp32), optional BatchNorm statistics            -- every conv that feeds a BatchNorm, plain data gradients
//   EPI_B16     y_bf16 = bf16(acc), optional statistics                  -- the same under bf16 storage
//   EPI_GENERIC bias / ReLU / mask / residual(_bf16) / accumulate / ragged cout (slab epilogue)
//   EPI_FRAGX   the same options for cout % 8 == 0, from the fragment registers (epilogue_fragx)
// and the register-path kernels keep EPI_ANY (PLAIN or GENERIC decided at run time, as before).
enum { EPI_PLAIN = 0, EPI_B16 = 1, EPI_GENERIC = 2, EPI_ANY = 3, EPI_FRAGX = 4 };

__host__ __device__ inline bool epilogue_is_plain(const zsg_conv_params& p) {
  return !p.bias && !p.out_mask && !p.residual && !p.residual_bf16 && !p.accumulate && !p.out_relu && !p.row_add &&
         (p.cout & 3) == 0;
}

// ---------------------------------------------------------------------------------------------------------------
// BN = 256 tiles of the bf16 path ("wide" kernel).  Measured on B200 (tools/micro/mma_bench.cu, tools/trace_conv.py): a
// 128x128x16 kind::f16 MMA issues every 75.6 cycles (85 % of the pipe), a 128x256x16 one every 128.0 (100 %); and in the
// 128-column kernel the PRODUCERS bound the big bf16 layers -- a producer group needs ~1.8 k cycles per K block (950 per K
// block over the two groups) against 300 cycles of MMA.  A 256-column tile does twice the MMA work per gathered A tile.
// Pipeline: ONE issuing thread (a K block is 4 x 128 cycles, the ~50 instructions between K blocks hide behind the two MMAs
// the pipe queues), the whole K extent accumulates in TMEM (no chunked promotion: its truncation bias, 1e-5 at K = 2304, is
// far below a bfloat16 ulp), two 256-column accumulators taken by tile parity, and the drain warps run the fragment
// epilogues of the 128-column kernel over their 128 columns in two passes of 64 straight out of TMEM.
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
__device__ __forceinline__ void mma_loop_wide(uint8_t* sm, const PipeBars& pb, uint32_t tmem_base, int nkb, int total_tiles,
                                              int ablate) {
  using S = Smem<BN, true>;
  const uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);   // K-major, SBO = 1024 B, 128-byte swizzle
  const uint32_t lo0 = ((smem_u32(sm) >> 4) & 0x3FFFu) | (1u << 16);
  uint32_t stage = 0, phase = 0, nt = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++nt) {
    const uint32_t a = nt & 1u;
    mbar_wait(pb.acc_empty(a), ((nt >> 1) & 1u) ^ 1u, 100 + nt);      // drained two tiles ago
    tc_fence_after();
    for (int kb = 0; kb < nkb; ++kb) {
      if (!(ablate & 1)) mbar_wait(pb.full(stage), phase, 1000 + kb);
      fence_proxy_async();
      tc_fence_after();
      issue_kblock_bf16<BN, false>(tmem_base + a * BN, lo0 + stage * (uint32_t)(S::STAGE_BYTES >> 4), desc_hi, kb == 0 ? 0u : 1u);
      umma_commit(pb.empty(stage));
      if (++stage == (uint32_t)S::STAGES) { stage = 0; phase ^= 1u; }
    }
    umma_commit(pb.acc_full(a));
  }
}

template <int BN, int EPI>
__device__ __forceinline__ void conv_epilogue_wide(const zsg_conv_params& p, const PipeBars& pb, uint32_t tmem_base, int warp,
                                                   int lane, int tiles_n, int total_tiles, int ablate) {
  static_assert(BN == 256 && (EPI == EPI_PLAIN || EPI == EPI_B16 || EPI == EPI_FRAGX), "wide kernel: fragment epilogues only");
  const int dw = warp - DRAIN_WARP0;
  const int quadrant = dw & 3, half = dw >> 2;              // TMEM lanes, 128-column half of the tile
  const int row = quadrant * 32 + lane;
  uint32_t nt = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++nt) {
    const int n0 = (tile % tiles_n) * BN + half * 128;
    const int m0 = (tile / tiles_n) * TM;
    if (EPI == EPI_FRAGX && (p.residual || p.residual_bf16 || p.accumulate || p.out_mask)) {
      const int ntile = tile + gridDim.x;                   // L2 prefetch of the next tile's read operands (see conv_epilogue)
      if (ntile < total_tiles) {
        const int nn0 = (ntile % tiles_n) * BN + half * 128, nm = (ntile / tiles_n) * TM + row;
        if (nm < p.m && nn0 < p.cout) {
          const int64_t e = (int64_t)nm * p.y_pitch + nn0;
          const int cols = (p.cout - nn0 < 128) ? p.cout - nn0 : 128;
          for (int c = 0; c < cols; c += 32) {
            if (p.residual_bf16 && (c & 63) == 0) prefetch_l2(p.residual_bf16 + e + c);
            if (p.residual) prefetch_l2(p.residual + e + c);
            if (p.out_mask) prefetch_l2(p.out_mask + e + c);
            if (p.accumulate) prefetch_l2(p.y + e + c);
          }
        }
      }
    }
    if (EPI != EPI_FRAGX && p.bnb_partials) {                // BatchNorm input rows of the next tile (backward sums)
      const int ntile = tile + gridDim.x;
      if (ntile < total_tiles) {
        const int nn0 = (ntile % tiles_n) * BN + half * 128, nm = (ntile / tiles_n) * TM + row;
        if (nm < p.m && nn0 < p.cout) {
          const int64_t e = (int64_t)nm * p.y_pitch + nn0;
          if (EPI == EPI_B16) {
            for (int c = 0; c < 128; c += 64) prefetch_l2(reinterpret_cast<const uint16_t*>(p.bnb_x) + e + c);
          } else {
            for (int c = 0; c < 128; c += 32) prefetch_l2(reinterpret_cast<const float*>(p.bnb_x) + e + c);
          }
        }
      }
    }
    const uint32_t a = nt & 1u;
    mbar_wait(pb.acc_full(a), (nt >> 1) & 1u, 2000 + nt);
    tc_fence_after();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float acc[64];
      const uint32_t taddr = tmem_base + ((uint32_t)(quadrant * 32) << 16) + a * BN + (uint32_t)(half * 128 + j * 64);
      if (!(ablate & 2)) {
        tmem_ld_frag<8>(taddr, acc);
        tmem_ld_frag<8>(taddr + (16u << 16), acc + 32);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      }
      if (j == 1) {                                         // every TMEM read of this tile is done: the accumulator may be reused
        tc_fence_before();
        mbar_arrive(pb.acc_empty(a));
      }
      if (EPI == EPI_FRAGX) epilogue_fragx<128>(p, acc, m0, n0, quadrant, j, lane, ablate);
      else epilogue_frag<128, EPI == EPI_B16>(p, acc, m0, n0, quadrant, j, lane, ablate);
    }
  }
}

template <int BN, bool BF16 = false, int EPI = EPI_ANY>
__device__ __forceinline__ void conv_epilogue(const zsg_conv_params& p, uint8_t* sm, const PipeBars& pb, uint32_t tmem_base,
                                              int warp, int lane, int nkb, int tiles_n, int total_tiles, int ablate = 0) {
  using S = Smem<BN, BF16>;
  constexpr bool HAS_PLAIN = EPI == EPI_PLAIN || EPI == EPI_ANY;
  constexpr bool HAS_B16 = EPI == EPI_B16;
  constexpr bool HAS_GENERIC = EPI == EPI_GENERIC || EPI == EPI_ANY;
  constexpr bool HAS_STATS = EPI != EPI_GENERIC;
  // ------------------------------ drain + epilogue ------------------------------
  const int dw = warp - DRAIN_WARP0;
  const int quadrant = dw & 3, half = dw >> 2;
  const int row = quadrant * 32 + lane;
  int gchunk = 0, gkb0 = 0;
  if (EPI == EPI_PLAIN || EPI == EPI_B16 || EPI == EPI_FRAGX) {   // register epilogue: no shared-memory transposition
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n0 = (tile % tiles_n) * BN;
      const int m0 = (tile / tiles_n) * TM;
      if (EPI == EPI_FRAGX && p.y_pitch > 0 && (p.residual || p.residual_bf16 || p.accumulate || p.out_mask)) {
        // the epilogue's read operands of the NEXT tile are asked for now (one to four 128-byte lines per thread and
        // operand), so that they are L2 hits when the loads are issued
        const int nt = tile + gridDim.x;
        if (nt < total_tiles) {
          const int nn0 = (nt % tiles_n) * BN + half * (BN / 2), nm = (nt / tiles_n) * TM + row;
          if (nm < p.m && nn0 < p.cout) {
            const int64_t e = (int64_t)nm * p.y_pitch + nn0;
            const int cols = (p.cout - nn0 < BN / 2) ? p.cout - nn0 : BN / 2;
            if (p.residual_bf16) prefetch_l2(p.residual_bf16 + e);
            for (int c = 0; c < cols; c += 32) {
              if (p.residual) prefetch_l2(p.residual + e + c);
              if (p.out_mask) prefetch_l2(p.out_mask + e + c);
              if (p.accumulate) prefetch_l2(p.y + e + c);
            }
          }
        }
      }
      if (EPI != EPI_FRAGX && p.bnb_partials) {              // the BatchNorm input rows of the NEXT tile (backward sums): into L2 now
        const int nt = tile + gridDim.x;
        if (nt < total_tiles) {
          const int nn0 = (nt % tiles_n) * BN + half * (BN / 2), nm = (nt / tiles_n) * TM + row;
          if (nm < p.m && nn0 < p.cout) {
            const int64_t e = (int64_t)nm * p.y_pitch + nn0;
            if (EPI == EPI_B16) {
              for (int c = 0; c < BN / 2; c += 64) prefetch_l2(reinterpret_cast<const uint16_t*>(p.bnb_x) + e + c);
            } else {
              for (int c = 0; c < BN / 2; c += 32) prefetch_l2(reinterpret_cast<const float*>(p.bnb_x) + e + c);
            }
          }
        }
      }
      float acc[BN / 2];
      if (dw == 0 && lane == 0) trace(gkb0, 12);
      drain_loop_frag<BN, BF16>(pb, tmem_base, nkb, gkb0, quadrant, half, acc, gchunk, ablate);
      if (dw == 0 && lane == 0) trace(gkb0, 13);
      if (EPI == EPI_FRAGX) epilogue_fragx<BN>(p, acc, m0, n0, quadrant, half, lane, ablate);
      else epilogue_frag<BN, EPI == EPI_B16>(p, acc, m0, n0, quadrant, half, lane, ablate);
      if (dw == 0 && lane == 0) trace(gkb0, 14);
      gkb0 += nkb;
    }
    return;
  }
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n0 = (tile % tiles_n) * BN;
    const int m0 = (tile / tiles_n) * TM;
    int out_off = 0;
    const bool row_ok = m0 + row < p.m;
    // y_pitch: the output is a plain [m, y_pitch] matrix -- no table read (on 1-2-K-block tiles the drain has nothing to
    // wait for and the ~700-cycle load was exposed once per tile)
    if (row_ok) out_off = p.y_pitch > 0 ? (m0 + row) * p.y_pitch : __ldg(&p.rows[m0 + row].out);
    if (HAS_GENERIC && p.y_pitch > 0 && (p.residual || p.residual_bf16 || p.accumulate || p.out_mask)) {
      // The epilogue's read operands (shortcut gradient, ReLU mask, the tensor it accumulates into) are streamed once from
      // HBM and consumed two 16-byte loads at a time: on the short-K data gradients that was ~9 k of 11 k cycles per tile
      // (tools/ablate_epilogue.py).  Each thread asks L2 for its row of the NEXT tile (one to four 128-byte lines) now;
      // by the time the loads are issued they are L2 hits.
      const int nt = tile + gridDim.x;
      if (nt < total_tiles) {
        const int nn0 = (nt % tiles_n) * BN + half * (BN / 2), nm = (nt / tiles_n) * TM + row;
        if (nm < p.m && nn0 < p.cout) {
          const int64_t e = (int64_t)nm * p.y_pitch + nn0;
          const int cols = (p.cout - nn0 < BN / 2) ? p.cout - nn0 : BN / 2;
          if (p.residual_bf16) prefetch_l2(p.residual_bf16 + e);
          for (int c = 0; c < cols; c += 32) {
            if (p.residual) prefetch_l2(p.residual + e + c);
            if (p.out_mask) prefetch_l2(p.out_mask + e + c);
            if (p.accumulate) prefetch_l2(p.y + e + c);
          }
        }
      }
    }
    int ra0 = 0, ra1 = 0;                                  // row_add: offsets of this row's two addend rows
    if (HAS_GENERIC && p.row_add && row_ok) {
      const int2 ra = __ldg(reinterpret_cast<const int2*>(p.row_add_idx) + m0 + row);
      ra0 = ra.x;
      ra1 = ra.y;
    }
    float acc[BN / 2];
    if (dw == 0 && lane == 0) trace(gkb0, 12);
    drain_loop<BN, BF16>(pb, tmem_base, nkb, gkb0, quadrant, half, acc, gchunk, ablate);
    if (dw == 0 && lane == 0) trace(gkb0, 13);
    const int gkb_tile = gkb0;
    gkb0 += nkb;
    // Epilogue through a per-warp smem slab: a thread owns one row of the accumulator, but global memory wants
    // lanes along channels.  16 columns at a time are transposed through smem (row stride 20 floats keeps the
    // 128-bit accesses conflict-free); then 4 lanes cover one row's 64 B and a warp instruction touches 8 rows
    // with full 32-byte sectors, for the store and for the residual / mask / accumulate reads alike.
    float* stg = reinterpret_cast<float*>(sm + S::EPI_OFF + dw * S::EPI_WARP_BYTES);
    const int rsel = lane >> 2, c4 = (lane & 3) * 4;
    // Plain stores: the row offsets of the rows this lane stores are fetched once per tile and the per-slab work is four
    // LDS.128 + four STG.128, against ~100 instructions per pass of the generic path.
    const bool plain = EPI == EPI_PLAIN || (EPI == EPI_ANY && epilogue_is_plain(p) && !(ablate & 16));
    int64_t po[4] = {0, 0, 0, 0};
    bool pk[4] = {false, false, false, false};
    bool use_plain = false;                                // vector stores: every row offset is 16-byte aligned
    if (HAS_PLAIN && plain) {
      bool al = true;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int o = __shfl_sync(0xffffffffu, out_off, i * 8 + rsel);
        pk[i] = __shfl_sync(0xffffffffu, (int)row_ok, i * 8 + rsel) != 0;
        po[i] = (int64_t)o + n0 + half * (BN / 2) + c4;
        al = al && (o & 3) == 0;
      }
      use_plain = __all_sync(0xffffffffu, al);
    }
    // bf16 storage (y_bf16): lane = (8-column half, row) -- two passes of 16 rows, 8 fp32 columns -> ONE 16-byte store,
    // the two halves of a row's 32 bytes in the same instruction (full sectors); slab reads stay conflict-free
    const int hb = lane >> 4, rb = lane & 15;
    int64_t pob[2] = {0, 0};
    bool pkb[2] = {false, false};
    if (HAS_B16) {
      bool al = true;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int o = __shfl_sync(0xffffffffu, out_off, rb + 16 * i);
        pkb[i] = __shfl_sync(0xffffffffu, (int)row_ok, rb + 16 * i) != 0;
        pob[i] = (int64_t)o + n0 + half * (BN / 2) + hb * 8;
        al = al && (o & 7) == 0;
      }
      use_plain = __all_sync(0xffffffffu, al);
    }
#pragma unroll
    for (int slab = 0; slab < BN / 32; ++slab) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(stg + lane * 20 + j) =
            make_float4(acc[slab * 16 + j], acc[slab * 16 + j + 1], acc[slab * 16 + j + 2], acc[slab * 16 + j + 3]);
      __syncwarp();
      if (HAS_STATS && p.stats && !(ablate & 128)) {
        // BatchNorm statistics of this warp's 32 rows x 16 columns while they sit in the slab: lane = (row half,
        // column); the two halves walk rows 20 (or 4) apart, which puts them on different banks (row stride 20 floats).
        // Rows past p.m hold zeros (their A rows were zero-filled), so they add nothing.
        const int col = lane & 15, hsel = lane >> 4;
        float s1 = 0.f, s2 = 0.f, t1 = 0.f, t2 = 0.f;        // two chains each: the 16-deep dependent add was exposed
#pragma unroll
        for (int r = 0; r < 16; r += 2) {
          const int ra = hsel ? 16 + ((r + 4) & 15) : r, rb2 = hsel ? 16 + ((r + 5) & 15) : r + 1;
          const float v = stg[ra * 20 + col], u = stg[rb2 * 20 + col];
          s1 += v;
          s2 = fmaf(v, v, s2);
          t1 += u;
          t2 = fmaf(u, u, t2);
        }
        s1 += t1;
        s2 += t2;
        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
        const int ncol = n0 + half * (BN / 2) + slab * 16 + col;
        if (ncol < p.cout)
          p.stats[((int64_t)((m0 / TM) * 4 + quadrant) * 2 + hsel) * p.cout + ncol] = hsel ? s2 : s1;
      }
      const int n = n0 + half * (BN / 2) + slab * 16 + c4;
      if (HAS_B16) {
        if (use_plain) {
          if (n0 + half * (BN / 2) + slab * 16 + hb * 8 < p.cout) {      // cout % 8 == 0: the 8 columns are inside
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float* src = stg + (rb + 16 * i) * 20 + hb * 8;
              const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
              const __nv_bfloat162 b0 = __floats2bfloat162_rn(v0.x, v0.y), b1 = __floats2bfloat162_rn(v0.z, v0.w);
              const __nv_bfloat162 b2 = __floats2bfloat162_rn(v1.x, v1.y), b3 = __floats2bfloat162_rn(v1.z, v1.w);
              uint4 u;
              u.x = *reinterpret_cast<const uint32_t*>(&b0);
              u.y = *reinterpret_cast<const uint32_t*>(&b1);
              u.z = *reinterpret_cast<const uint32_t*>(&b2);
              u.w = *reinterpret_cast<const uint32_t*>(&b3);
              if (pkb[i] && !(ablate & 16)) *reinterpret_cast<uint4*>(p.y_bf16 + pob[i] + slab * 16) = u;
            }
          }
          continue;
        }
      } else if (HAS_PLAIN && use_plain) {
        if (n < p.cout) {                                   // cout % 4 == 0: the whole float4 is inside
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(stg + (i * 8 + rsel) * 20 + c4);
            if (pk[i] && !(ablate & 16)) *reinterpret_cast<float4*>(p.y + po[i] + slab * 16) = v;
          }
        }
        continue;
      }
      if (!HAS_GENERIC) {
        // EPI_PLAIN / EPI_B16 over a row table whose offsets are not 16-byte aligned (no table of the path is): scalar stores,
        // not unrolled -- correct, compact, never on the hot path
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int r_ = i * 8 + rsel;
          const int o = __shfl_sync(0xffffffffu, out_off, r_);
          const bool ok = __shfl_sync(0xffffffffu, (int)row_ok, r_) != 0;
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            if (!ok || n + q >= p.cout) continue;
            const float v = stg[r_ * 20 + c4 + q];
            if (HAS_B16) reinterpret_cast<__nv_bfloat16*>(p.y_bf16)[(int64_t)o + n + q] = __float2bfloat16_rn(v);
            else p.y[(int64_t)o + n + q] = v;
          }
        }
        continue;
      }
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias && n < p.cout) {
        if (n + 3 < p.cout) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        else { bias4.x = __ldg(p.bias + n); if (n + 1 < p.cout) bias4.y = __ldg(p.bias + n + 1); if (n + 2 < p.cout) bias4.z = __ldg(p.bias + n + 2); }
      }
      // Two rows per pass, one operand array at a time (mask, residual, accumulate read) through ONE pair of
      // temporaries: the two loads of a pair are in flight together, which halves the exposed load latencies of the
      // short-K data gradients (row by row they cost ~800 cycles each).  Loop, not unrolled: inlined four times with
      // the ragged path the epilogue was > 100 KB of code and ran out of the instruction cache.
#pragma unroll 1
      for (int ih = 0; ih < 2; ++ih) {
        const int r0 = (ih * 2) * 8 + rsel, r1 = r0 + 8;
        const int o0 = __shfl_sync(0xffffffffu, out_off, r0), o1 = __shfl_sync(0xffffffffu, out_off, r1);
        const bool k0 = __shfl_sync(0xffffffffu, (int)row_ok, r0) != 0 && n < p.cout;
        const bool k1 = __shfl_sync(0xffffffffu, (int)row_ok, r1) != 0 && n < p.cout;
        float4 v0 = *reinterpret_cast<const float4*>(stg + r0 * 20 + c4);
        float4 v1 = *reinterpret_cast<const float4*>(stg + r1 * 20 + c4);
        v0.x += bias4.x; v0.y += bias4.y; v0.z += bias4.z; v0.w += bias4.w;
        v1.x += bias4.x; v1.y += bias4.y; v1.z += bias4.z; v1.w += bias4.w;
        const bool vec = ((p.cout & 3) == 0) && (((o0 | o1) & 3) == 0) && (n + 3 < p.cout);
        int i00 = 0, i01 = 0, i10 = 0, i11 = 0;             // row_add offsets: shuffled by ALL lanes (`vec` may diverge on a ragged cout)
        if (p.row_add) {
          i00 = __shfl_sync(0xffffffffu, ra0, r0); i01 = __shfl_sync(0xffffffffu, ra1, r0);
          i10 = __shfl_sync(0xffffffffu, ra0, r1); i11 = __shfl_sync(0xffffffffu, ra1, r1);
        }
        if (vec) {
          const int64_t a0 = (int64_t)o0 + n, a1 = (int64_t)o1 + n;
          float4 q0, q1;
          if (p.row_add) {                                  // language + grid terms of the first head conv (L2-resident tables)
            if (k0) {
              q0 = __ldg(reinterpret_cast<const float4*>(p.row_add + i00 + n));
              q1 = __ldg(reinterpret_cast<const float4*>(p.row_add + i01 + n));
              v0.x += q0.x + q1.x; v0.y += q0.y + q1.y; v0.z += q0.z + q1.z; v0.w += q0.w + q1.w;
            }
            if (k1) {
              q0 = __ldg(reinterpret_cast<const float4*>(p.row_add + i10 + n));
              q1 = __ldg(reinterpret_cast<const float4*>(p.row_add + i11 + n));
              v1.x += q0.x + q1.x; v1.y += q0.y + q1.y; v1.z += q0.z + q1.z; v1.w += q0.w + q1.w;
            }
          }
          if (p.out_mask) {
            if (k0) q0 = __ldg(reinterpret_cast<const float4*>(p.out_mask + a0));
            if (k1) q1 = __ldg(reinterpret_cast<const float4*>(p.out_mask + a1));
            if (k0) { if (!(q0.x > 0.f)) v0.x = 0.f; if (!(q0.y > 0.f)) v0.y = 0.f; if (!(q0.z > 0.f)) v0.z = 0.f; if (!(q0.w > 0.f)) v0.w = 0.f; }
            if (k1) { if (!(q1.x > 0.f)) v1.x = 0.f; if (!(q1.y > 0.f)) v1.y = 0.f; if (!(q1.z > 0.f)) v1.z = 0.f; if (!(q1.w > 0.f)) v1.w = 0.f; }
          }
          if (p.residual) {
            if (k0) q0 = *reinterpret_cast<const float4*>(p.residual + a0);
            if (k1) q1 = *reinterpret_cast<const float4*>(p.residual + a1);
            if (k0) { v0.x += q0.x; v0.y += q0.y; v0.z += q0.z; v0.w += q0.w; }
            if (k1) { v1.x += q1.x; v1.y += q1.y; v1.z += q1.z; v1.w += q1.w; }
          }
          if (EPI == EPI_GENERIC && p.residual_bf16) {      // bf16 storage: the shortcut gradient is a bfloat16 tensor
            uint2 u0 = make_uint2(0u, 0u), u1 = make_uint2(0u, 0u);
            if (k0) u0 = *reinterpret_cast<const uint2*>(p.residual_bf16 + a0);
            if (k1) u1 = *reinterpret_cast<const uint2*>(p.residual_bf16 + a1);
            v0.x += __uint_as_float(u0.x << 16); v0.y += __uint_as_float(u0.x & 0xFFFF0000u);
            v0.z += __uint_as_float(u0.y << 16); v0.w += __uint_as_float(u0.y & 0xFFFF0000u);
            v1.x += __uint_as_float(u1.x << 16); v1.y += __uint_as_float(u1.x & 0xFFFF0000u);
            v1.z += __uint_as_float(u1.y << 16); v1.w += __uint_as_float(u1.y & 0xFFFF0000u);
          }
          if (p.accumulate) {
            if (k0) q0 = *reinterpret_cast<const float4*>(p.y + a0);
            if (k1) q1 = *reinterpret_cast<const float4*>(p.y + a1);
            if (k0) { v0.x += q0.x; v0.y += q0.y; v0.z += q0.z; v0.w += q0.w; }
            if (k1) { v1.x += q1.x; v1.y += q1.y; v1.z += q1.z; v1.w += q1.w; }
          }
          if (p.out_relu) {
            v0.x = fmaxf(v0.x, 0.f); v0.y = fmaxf(v0.y, 0.f); v0.z = fmaxf(v0.z, 0.f); v0.w = fmaxf(v0.w, 0.f);
            v1.x = fmaxf(v1.x, 0.f); v1.y = fmaxf(v1.y, 0.f); v1.z = fmaxf(v1.z, 0.f); v1.w = fmaxf(v1.w, 0.f);
          }
          if (!(ablate & 16)) {
            if (k0) *reinterpret_cast<float4*>(p.y + a0) = v0;
            if (k1) *reinterpret_cast<float4*>(p.y + a1) = v1;
          }
        } else {
          if ((k0 || k1) && (p.row_add || (EPI == EPI_GENERIC && p.residual_bf16))) __trap();   // checked on the host: cout % 4 == 0; only unaligned row offsets get here
          if (k0) epilogue_store_ragged(p.y, p.out_mask, p.residual, p.accumulate, p.out_relu, p.cout, o0, n, v0);
          if (k1) epilogue_store_ragged(p.y, p.out_mask, p.residual, p.accumulate, p.out_relu, p.cout, o1, n, v1);
        }
      }
    }
    if (dw == 0 && lane == 0) trace(gkb_tile, 14);
  }
}

// ============================================================================================
// forward / data-gradient kernel, generic variant: weights gathered through registers (p.w_lo == NULL) and every
// option decided at run time.  The product path (pre-split weights, TMA) is conv_tc_kernel below.
// ============================================================================================
template <int BN, bool TMA_W>
__global__ void __launch_bounds__(NTHREADS2, 1) conv_tc_generic_kernel(const zsg_conv_params p,
                                                               const __grid_constant__ CUtensorMap tm_hi,
                                                               const __grid_constant__ CUtensorMap tm_lo) {
  // Persistent: one CTA per SM walks the output tiles (n fastest, so the CTAs that share im2col rows run at the
  // same time and hit L2).  TMEM and the mbarriers are set up once; the smem ring, the two TMEM accumulators and
  // all barrier phases run on counters that continue across tiles, so the epilogue of tile i overlaps the gathers
  // and MMAs of tile i+1.
  using S = Smem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.r * p.s * p.cin;
  const int nkb = (K + KB - 1) / KB;
  const int tiles_n = (p.cout + BN - 1) / BN;
  const int total_tiles = tiles_n * ((p.m + TM - 1) / TM);

  PipeBars pb = setup_pipeline<BN>(sm, warp, lane, TMA_W ? NPROD + 1 : NPROD);   // contains __syncthreads
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + S::BAR_OFF + TMEM_SLOT_OFF);

  if (warp >= MMA_WARP) {
    if (elect_one()) {
      Issuer is = issuer_init<BN>(warp - MMA_WARP);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) mma_loop<BN>(sm, pb, tmem_base, nkb, is);
    }
    __syncwarp();
  } else if (warp < DRAIN_WARP0) {
    // ------------------------------ producers ------------------------------
    // Thread (rsub, chunk) owns the 16-byte chunk `chunk` of rows rsub, rsub+16, ... of every K block of its
    // group.  Its filter tap changes only every cin/32 K blocks, so the bounds check and the pixel offset of
    // its 8 rows are cached per tap; the swizzled smem offset is a compile-time function of (it, rsub, chunk).
    const int group = warp >> 2;
    const int t = tid & 127;
    const int chunk = t & 7;
    const int rsub = t >> 3;
    const int soff = rsub * 128 + ((chunk ^ (rsub & 7)) << 4);      // + it * 2048 (16 rows x 128 B)
    const int ntap = p.r * p.s;
    int4* rows_g = reinterpret_cast<int4*>(sm + S::ROWS_OFF) + group * TM;      // this group's copy of the row table
    int gkb0 = 0;                                                   // global index of the tile's first K block
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, gkb0 += nkb) {
      const int n0 = (tile % tiles_n) * BN;
      const int m0 = (tile / tiles_n) * TM;
      {
        int4 e = make_int4(0, 0, 0, 0);                   // hin = win = 0 => every tap out of bounds
        if (m0 + t < p.m) e = __ldg(reinterpret_cast<const int4*>(p.rows) + m0 + t);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");   // previous tile's reads are done
        rows_g[t] = e;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
      }
      const int kb_first = (group - gkb0) & 1;            // K blocks with (gkb0 + kb) % NGROUP == group
      int c = chunk * 4 + kb_first * KB, tap = 0, tr = 0, ts = 0;
      while (c >= p.cin) { c -= p.cin; ++tap; if (++ts == p.s) { ts = 0; ++tr; } }
      int cached_tap = -1;
      int off[8];                                         // element offset of the tap's pixel, -1 = padding
      for (int kb = kb_first; kb < nkb; kb += NGROUP) {
        const int gk = gkb0 + kb;
        const int s = gk % S::STAGES;
        const bool kvalid = tap < ntap;
        if (tap != cached_tap) {
          cached_tap = tap;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int4 e = rows_g[it * 16 + rsub];
            int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
            const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
            bool ok = kvalid;
            if (p.in_div == 2) { ok = ok && (((yy | xx) & 1) == 0); yy >>= 1; xx >>= 1; }
            ok = ok && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
            off[it] = ok ? e.x + (yy * win + xx) * p.cin : -1;
          }
        }
        // issue the gather loads before waiting for the stage: they only need registers
        // (p.impl >= 2 are timing ablations used by tools/ablate_conv.py: 2 = no gather loads, 3 = also no smem
        //  stores, 4 = also no proxy fence, 5 = everything but the proxy fence; results are garbage then)
        float4 va[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          va[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (off[it] >= 0 && (p.impl < 2 || p.impl == 5))
            va[it] = __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)off[it] + c));
        }
        float4 vb[TMA_W ? 1 : BN / 16];
        if (!TMA_W) {
          const int kk = kb * KB + chunk * 4;
#pragma unroll
          for (int it = 0; it < BN / 16; ++it) {
            const int n = n0 + it * 16 + rsub;
            vb[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kvalid && n < p.cout) vb[it] = __ldg(reinterpret_cast<const float4*>(p.w + (int64_t)n * K + kk));
          }
        }
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.in_scale && kvalid) {
          sc = __ldg(reinterpret_cast<const float4*>(p.in_scale + c));
          sh = __ldg(reinterpret_cast<const float4*>(p.in_shift + c));
        }
        mbar_wait(pb.empty(s), ((gk / S::STAGES) & 1) ^ 1);
        uint8_t* a_hi = sm + s * S::STAGE_BYTES;
        uint8_t* a_lo = a_hi + A_TILE_BYTES;
        uint8_t* b_hi = a_lo + A_TILE_BYTES;
        uint8_t* b_lo = b_hi + S::B_TILE_BYTES;
        if (TMA_W && t == 0 && p.impl != 6) {             // weights: two TMA tiles, no register pass
          mbar_arrive_expect_tx(pb.full(s), 2 * S::B_TILE_BYTES);
          tma_load_2d(smem_u32(b_hi), &tm_hi, kb * KB, n0, pb.full(s));
          tma_load_2d(smem_u32(b_lo), &tm_lo, kb * KB, n0, pb.full(s));
        }
        if (TMA_W && t == 0 && p.impl == 6) mbar_arrive(pb.full(s));   // ablation 6: no weight TMA
        if (p.impl != 3 && p.impl != 4)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          float4 v = va[it];
          if (off[it] >= 0) {
            if (p.in_scale) {
              v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
              v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
            }
            if (p.in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          }
          float4 h, l;
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
          *reinterpret_cast<float4*>(a_hi + soff + it * 2048) = h;
          *reinterpret_cast<float4*>(a_lo + soff + it * 2048) = l;
        }
        if (!TMA_W) {
#pragma unroll
          for (int it = 0; it < BN / 16; ++it) store_split(b_hi, b_lo, it * 16 + rsub, chunk, vb[it]);
        }
        if (p.impl != 4 && p.impl != 5) fence_proxy_async();
        mbar_arrive(pb.full(s));
        // advance this thread's (tap, channel) by NGROUP K blocks
        c += NGROUP * KB;
        while (c >= p.cin) { c -= p.cin; ++tap; if (++ts == p.s) { ts = 0; ++tr; } }
      }
    }
  } else {
    conv_epilogue<BN>(p, sm, pb, tmem_base, warp, lane, nkb, tiles_n, total_tiles);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ============================================================================================
// forward / data-gradient kernel, product variant.
//
// Same tiling, barriers, MMA and drain code as above; the producers are written for instruction count, because
// ncu showed them issue-bound (about 440 SASS instructions per thread per K block, the MMA warp idle a fifth of
// the time waiting for `full`):  the prologue (PRO bit 1 = BatchNorm affine, bit 0 = ReLU) is a template
// parameter; gather loads, the affine and the zero-padding rule are predicated instructions instead of
// branches; one IMAD.WIDE forms each address; the affine and the hi/lo subtraction use packed f32x2 math;
// weights always arrive by TMA; one elected lane per producer warp arrives on `full`.
// ============================================================================================
__device__ __forceinline__ float4 ldg128_pred(const float* ptr, int off) {     // zeros when off < 0
  float4 v;
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %5, 0;\n\t"
      "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(ptr + off), "r"(off));
  return v;
}
// v = v * sc + sh where off >= 0 (padding rows stay exactly zero)
__device__ __forceinline__ void affine4_pred(float4& v, const float4& sc, const float4& sh, int off) {
  asm("{\n\t.reg .pred q;\n\t.reg .b64 a, b, c;\n\tsetp.ge.s32 q, %12, 0;\n\t"
      "mov.b64 a, {%0, %1};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%8, %9};\n\t"
      "@q fma.rn.f32x2 a, a, b, c;\n\tmov.b64 {%0, %1}, a;\n\t"
      "mov.b64 a, {%2, %3};\n\tmov.b64 b, {%6, %7};\n\tmov.b64 c, {%10, %11};\n\t"
      "@q fma.rn.f32x2 a, a, b, c;\n\tmov.b64 {%2, %3}, a;\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "f"(sc.x), "f"(sc.y), "f"(sc.z), "f"(sc.w), "f"(sh.x), "f"(sh.y), "f"(sh.z), "f"(sh.w), "r"(off));
}
__device__ __forceinline__ void affine4(float4& v, const float4& sc, const float4& sh) {
  asm("{\n\t.reg .b64 a, b, c;\n\t"
      "mov.b64 a, {%0, %1};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%8, %9};\n\t"
      "fma.rn.f32x2 a, a, b, c;\n\tmov.b64 {%0, %1}, a;\n\t"
      "mov.b64 a, {%2, %3};\n\tmov.b64 b, {%6, %7};\n\tmov.b64 c, {%10, %11};\n\t"
      "fma.rn.f32x2 a, a, b, c;\n\tmov.b64 {%2, %3}, a;\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "f"(sc.x), "f"(sc.y), "f"(sc.z), "f"(sc.w), "f"(sh.x), "f"(sh.y), "f"(sh.z), "f"(sh.w));
}
__device__ __forceinline__ void sts128(uint32_t dst, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void split4(const float4& v, float4& h, float4& l) {
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  asm("{\n\t.reg .b64 a, b;\n\t"
      "mov.b64 a, {%4, %5};\n\tmov.b64 b, {%8, %9};\n\tsub.rn.f32x2 a, a, b;\n\tmov.b64 {%0, %1}, a;\n\t"
      "mov.b64 a, {%6, %7};\n\tmov.b64 b, {%10, %11};\n\tsub.rn.f32x2 a, a, b;\n\tmov.b64 {%2, %3}, a;\n\t}"
      : "=f"(l.x), "=f"(l.y), "=f"(l.z), "=f"(l.w)
      : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w));
}

constexpr int FULL_COUNT_TMA = NPROD / 32 + 1;

// Register re-balancing (cp.async kernels): the producers only form addresses, the drain warps hold 64 accumulators
// per thread plus the epilogue's state.  The carve-out leaves ~6 KB of L1, so a spilled register costs an L2 round
// trip (two spilled values in the epilogue loop cost 7.5 k cycles per tile).  Launch-time count: 96 -- registers are
// allocated per 4 warps, so the 18 warps pay for 20 and 20 x 32 x 96 = 61440 is the most that fits (a build with
// __maxnreg__(112) fails to launch: "too many resources").  Producers give up 96 -> 56, drain warps take exactly what
// was freed, 96 -> 136: setmaxnreg.inc draws only from what the CTA's own warps released, a larger request never returns.
__device__ __forceinline__ void regs_release_producer() { asm volatile("setmaxnreg.dec.sync.aligned.u32 56;"); }
__device__ __forceinline__ void regs_take_drain() { asm volatile("setmaxnreg.inc.sync.aligned.u32 136;"); }
     // one elected arrive per producer warp + the expect_tx arrive

template <int BN, int PRO>
__global__ void __launch_bounds__(NTHREADS2, 1) conv_tc_kernel(const zsg_conv_params p,
                                                               const __grid_constant__ CUtensorMap tm_hi,
                                                               const __grid_constant__ CUtensorMap tm_lo) {
  using S = Smem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.r * p.s * p.cin;
  const int nkb = (K + KB - 1) / KB;
  const int tiles_n = (p.cout + BN - 1) / BN;
  const int total_tiles = tiles_n * ((p.m + TM - 1) / TM);

  const int ablate = p.impl >= 8 ? p.impl - 8 : 0;     // diagnostics (tools/ablate_conv.py): 1 = no producers, 2 = no TMEM drain, 8 = issuers without token
  PipeBars pb = setup_pipeline<BN>(sm, warp, lane, FULL_COUNT_TMA);   // contains __syncthreads
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + S::BAR_OFF + TMEM_SLOT_OFF);

  if (warp >= MMA_WARP) {
    if (elect_one()) {
      Issuer is = issuer_init<BN>(warp - MMA_WARP);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) mma_loop<BN>(sm, pb, tmem_base, nkb, is, false, ablate);
    }
    __syncwarp();
  } else if (warp < DRAIN_WARP0 && !(ablate & 1)) {
    // ------------------------------ producers ------------------------------
    const int group = warp >> 2;
    const int t = tid & 127;
    const int chunk = t & 7;
    const int rsub = t >> 3;
    const uint32_t soff = rsub * 128 + ((chunk ^ (rsub & 7)) << 4);      // + it * 2048 (16 rows x 128 B)
    const int ntap = p.r * p.s;
    const int cin = p.cin;
    int4* rows_g = reinterpret_cast<int4*>(sm + S::ROWS_OFF) + group * TM;      // this group's copy of the row table
    const uint32_t tiles0 = smem_u32(sm);
    int gkb0 = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, gkb0 += nkb) {
      const int n0 = (tile % tiles_n) * BN;
      const int m0 = (tile / tiles_n) * TM;
      {
        int4 e = make_int4(0, 0, 0, 0);                   // hin = win = 0 => every tap out of bounds
        if (m0 + t < p.m) e = __ldg(reinterpret_cast<const int4*>(p.rows) + m0 + t);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");   // previous tile's reads are done
        rows_g[t] = e;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
      }
      const int kb_first = (group - gkb0) & 1;            // K blocks with (gkb0 + kb) % NGROUP == group
      // (c, tap, tr, ts, off[], mask) always describe the block whose gather loads are issued NEXT.  The loads of
      // block i + NGROUP are issued row by row while block i is converted, into the registers block i just freed:
      // a whole iteration ahead, so their latency never sits between `empty` and `full`.
      int c = chunk * 4 + kb_first * KB, tap = 0, tr = 0, ts = 0;
      while (c >= cin) { c -= cin; ++tap; if (++ts == p.s) { ts = 0; ++tr; } }
      int cached_tap = -1;
      int off[8];                                         // element offset of the tap's pixel, -1 = padding
      uint32_t mask = 0;                                  // bit it = row it is a real pixel at the cached tap
      auto retap = [&]() {
        cached_tap = tap;
        const bool kvalid = tap < ntap;
        mask = 0;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int4 e = rows_g[it * 16 + rsub];
          int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
          const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
          bool ok = kvalid;
          if (p.in_div == 2) { ok = ok && (((yy | xx) & 1) == 0); yy >>= 1; xx >>= 1; }
          ok = ok && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
          off[it] = ok ? e.x + (yy * win + xx) * cin : -1;
          mask |= ok ? (1u << it) : 0u;
        }
      };
      float4 va[8];
      if (kb_first < nkb) {                               // prologue: loads of this group's first block of the tile
        retap();
        const float* xb = p.x + c;
        asm("" : "+l"(xb));
#pragma unroll
        for (int it = 0; it < 8; ++it) va[it] = ldg128_pred(xb, off[it]);
      }
      for (int kb = kb_first; kb < nkb; kb += NGROUP) {
        const int gk = gkb0 + kb;
        const int s = gk % S::STAGES;
        const uint32_t cur_mask = mask;                   // validity of the rows held in va[]
        float4 sc, sh;
        if (PRO & 2) {
          sc = __ldg(reinterpret_cast<const float4*>(p.in_scale + c));
          sh = __ldg(reinterpret_cast<const float4*>(p.in_shift + c));
        }
        // advance (tap, channel) to this group's next block and prepare its addresses
        const bool has_next = kb + NGROUP < nkb;
        c += NGROUP * KB;
        while (c >= cin) { c -= cin; ++tap; if (++ts == p.s) { ts = 0; ++tr; } }
        if (has_next && tap != cached_tap) retap();
        const float* xb = p.x + c;
        asm("" : "+l"(xb));                                // keep xb a 64-bit base: each address is one IMAD.WIDE
        if (t == 0) trace(gk, 0);
        mbar_wait(pb.empty(s), ((gk / S::STAGES) & 1) ^ 1, 3000 + gk);
        if (t == 0) trace(gk, 1);
        const uint32_t a_hi = tiles0 + s * S::STAGE_BYTES;
        if ((t >> 5) == 0) {                               // weights: two TMA tiles, no register pass
          if (elect_one()) {
            mbar_arrive_expect_tx(pb.full(s), 2 * S::B_TILE_BYTES);
            tma_load_2d(a_hi + 2 * A_TILE_BYTES, &tm_hi, kb * KB, n0, pb.full(s));
            tma_load_2d(a_hi + 2 * A_TILE_BYTES + S::B_TILE_BYTES, &tm_lo, kb * KB, n0, pb.full(s));
          }
          __syncwarp();
        }
        const bool cur_full = cur_mask == 0xFFu, next_full = mask == 0xFFu;
        if (cur_full && next_full && has_next) {           // interior: no per-row predicates anywhere
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            float4 v = va[it];
            va[it] = __ldg(reinterpret_cast<const float4*>(xb + off[it]));
            if (PRO & 2) affine4(v, sc, sh);
            if (PRO & 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            float4 h, l;
            split4(v, h, l);
            sts128(a_hi + soff + it * 2048, h);
            sts128(a_hi + soff + it * 2048 + A_TILE_BYTES, l);
          }
        } else {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            float4 v = va[it];
            if (has_next) va[it] = ldg128_pred(xb, off[it]);
            if (PRO & 2) affine4_pred(v, sc, sh, (cur_mask >> it) & 1u ? 0 : -1);
            if (PRO & 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            float4 h, l;
            split4(v, h, l);
            sts128(a_hi + soff + it * 2048, h);
            sts128(a_hi + soff + it * 2048 + A_TILE_BYTES, l);
          }
        }
        if (t == 0) trace(gk, 2);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(pb.full(s));
        if (t == 0) trace(gk, 3);
      }
    }
  } else if (warp >= DRAIN_WARP0) {
    conv_epilogue<BN>(p, sm, pb, tmem_base, warp, lane, nkb, tiles_n, total_tiles, ablate);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ============================================================================================
// forward / data-gradient kernel, pre-split input operand (p.x_lo != NULL).
//
// The input tensor comes with its TF32 remainders (zsg_split_act): the tensor core reads a raw fp32 word as its
// TF32 truncation, so `x` itself is the high-part tile and `x_lo` the low-part tile, and both go global -> shared
// with cp.async (16 B per request, zero-filled for padding taps) -- no register pass, no arithmetic.  Measured
// before this existed: the register-path producers (gather, affine, split, 16 STS.128 per thread and K block) took
// 1000-2700 cycles per K block and group against a 768-cycle MMA floor.
// ============================================================================================
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// BF16 (p.x_bf16 / p.w_bf16): the same kernel over the bf16 images -- one image per operand, 64 channels per K block,
// 8 cp.async per producer thread and K block, one TMA weight tile, four kind::f16 MMAs (issue_kblock_bf16).
template <int BN, bool BF16, int EPI>
__global__ void __launch_bounds__(NTHREADS2, 1) conv_tc_async_kernel(const zsg_conv_params p,
                                                                     const __grid_constant__ CUtensorMap tm_hi,
                                                                     const __grid_constant__ CUtensorMap tm_lo,
                                                                     const __grid_constant__ CUtensorMap tm_a,
                                                                     const __grid_constant__ CUtensorMap tm_a_lo) {
  using S = Smem<BN, BF16>;
  constexpr int KBE = S::KBE;              // K elements per stage
  constexpr int CE = 16 / S::ES;           // elements per 16-byte chunk
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.r * p.s * p.cin;
  const int nkb = (K + KBE - 1) / KBE;
  const int tiles_n = (p.cout + BN - 1) / BN;
  const int total_tiles = tiles_n * ((p.m + TM - 1) / TM);
  const int ablate = p.impl >= 8 ? p.impl - 8 : 0;
  // x_plain (1x1 stride-1 convs and their data gradients: A is a plain [m, cin] matrix): the A tiles come by TMA like the
  // weights, `full` then counts only the expect_tx arrive.  Otherwise: 128 cp.async completions + the expect_tx arrive.
  const bool plain_a = p.x_plain != 0;
  PipeBars pb = setup_pipeline<BN, BF16>(sm, warp, lane, plain_a ? 1 : NPROD + 1);
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + S::BAR_OFF + TMEM_SLOT_OFF);

  static_assert(BN <= 128 || (BF16 && EPI != EPI_GENERIC), "256-column tiles: bf16 path, fragment epilogues");
  if (warp >= MMA_WARP) {
    if (BN > 128) {                                         // wide kernel: one issuing thread (warp 17 idles)
      if (warp == MMA_WARP && elect_one()) mma_loop_wide<BN>(sm, pb, tmem_base, nkb, total_tiles, ablate);
    } else if (elect_one()) {
      Issuer is = issuer_init<BN, BF16>(warp - MMA_WARP);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x)
        mma_loop<BN, false, BF16>(sm, pb, tmem_base, nkb, is, false, ablate);
    }
    __syncwarp();
  } else if (warp < DRAIN_WARP0) {
    // ------------------------------ producers ------------------------------
    regs_release_producer();
    if (plain_a) {
      // One thread feeds the whole ring: per K block one expect_tx and 2 (bf16) or 4 (hi + lo images) TMA tiles.  The
      // per-tile chain of the gather path (row entries -> group barriers -> tap offsets -> 8-16 cp.async per thread) took
      // ~7 k cycles per tile on the 1-2-K-block tiles of the 64-channel layers against 0.25-1.5 k cycles of MMA time.
      if (warp == 0 && !(ablate & 1)) {
        if (elect_one()) {
          const uint32_t tiles0 = smem_u32(sm);
          uint32_t stage = 0, phase = 0;
          for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n0 = (tile % tiles_n) * BN;
            const int m0 = (tile / tiles_n) * TM;
            for (int kb = 0; kb < nkb; ++kb) {
              mbar_wait(pb.empty(stage), phase ^ 1u, 3000 + kb);
              const uint32_t a_hi = tiles0 + stage * S::STAGE_BYTES;
              mbar_arrive_expect_tx(pb.full(stage), S::NIMG * (A_TILE_BYTES + S::B_TILE_BYTES));
              tma_load_2d(a_hi, &tm_a, kb * KBE, m0, pb.full(stage));
              if (!BF16) tma_load_2d(a_hi + A_TILE_BYTES, &tm_a_lo, kb * KBE, m0, pb.full(stage));
              tma_load_2d(a_hi + S::NIMG * A_TILE_BYTES, &tm_hi, kb * KBE, n0, pb.full(stage));
              if (!BF16) tma_load_2d(a_hi + 2 * A_TILE_BYTES + S::B_TILE_BYTES, &tm_lo, kb * KBE, n0, pb.full(stage));
              if (++stage == (uint32_t)S::STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
        __syncwarp();
      }
    } else if (!(ablate & 1)) {
    const int group = warp >> 2;
    const int t = tid & 127;
    const int chunk = t & 7;
    const int rsub = t >> 3;
    const uint32_t soff = rsub * 128 + ((chunk ^ (rsub & 7)) << 4);      // + it * 2048 (16 rows x 128 B)
    const int ntap = p.r * p.s;
    const int cin = p.cin;
    int4* rows_g = reinterpret_cast<int4*>(sm + S::ROWS_OFF) + group * TM;      // this group's copy of the row table
    const uint32_t tiles0 = smem_u32(sm);
    const uint8_t* xbase = BF16 ? reinterpret_cast<const uint8_t*>(p.x_bf16) : reinterpret_cast<const uint8_t*>(p.x);
    const uint8_t* xlbase = reinterpret_cast<const uint8_t*>(p.x_lo);
    int gkb0 = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, gkb0 += nkb) {
      const int n0 = (tile % tiles_n) * BN;
      const int m0 = (tile / tiles_n) * TM;
      if (t == 0) trace(gkb0 + ((group - gkb0) & 1), 4);  // tile top (stamps exist in the build.py --trace build only)
      {
        int4 e = make_int4(0, 0, 0, 0);                   // hin = win = 0 => every tap out of bounds
        if (m0 + t < p.m) e = __ldg(reinterpret_cast<const int4*>(p.rows) + m0 + t);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");   // previous tile's reads are done
        if (t == 0) trace(gkb0 + ((group - gkb0) & 1), 5);  // first group barrier passed
        rows_g[t] = e;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
      }
      if (t == 0) trace(gkb0 + ((group - gkb0) & 1), 6);  // row table of the tile is in shared memory
      const int kb_first = (group - gkb0) & 1;            // K blocks with (gkb0 + kb) % NGROUP == group
      int c = chunk * CE + kb_first * KBE, tap = 0, tr = 0, ts = 0;
      while (c >= cin) { c -= cin; ++tap; if (++ts == p.s) { ts = 0; ++tr; } }
      int cached_tap = -1;
      int off[8];                                         // element offset of the tap's pixel, -1 = padding
      for (int kb = kb_first; kb < nkb; kb += NGROUP) {
        const int gk = gkb0 + kb;
        const int s = gk % S::STAGES;
        if (tap != cached_tap) {
          cached_tap = tap;
          const bool kvalid = tap < ntap;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int4 e = rows_g[it * 16 + rsub];
            int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
            const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
            bool ok = kvalid;
            if (p.in_div == 2) { ok = ok && (((yy | xx) & 1) == 0); yy >>= 1; xx >>= 1; }
            ok = ok && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
            off[it] = ok ? e.x + (yy * win + xx) * cin : -1;
          }
        }
        const uint8_t* xh = xbase + (int64_t)c * S::ES;
        const uint8_t* xl = BF16 ? xh : xlbase + (int64_t)c * S::ES;
        asm("" : "+l"(xh));
        asm("" : "+l"(xl));
        if (t == 0) trace(gk, 0);
        mbar_wait(pb.empty(s), ((gk / S::STAGES) & 1) ^ 1, 3000 + gk);
        if (t == 0) trace(gk, 1);
        const uint32_t a_hi = tiles0 + s * S::STAGE_BYTES;
        if ((t >> 5) == 0) {                               // weights: one TMA tile per image
          if (elect_one()) {
            if (ablate & 64) {                              // diagnostics: no weight loads
              mbar_arrive(pb.full(s));
            } else {
              mbar_arrive_expect_tx(pb.full(s), S::NIMG * S::B_TILE_BYTES);
              tma_load_2d(a_hi + S::NIMG * A_TILE_BYTES, &tm_hi, kb * KBE, n0, pb.full(s));
              if (!BF16) tma_load_2d(a_hi + 2 * A_TILE_BYTES + S::B_TILE_BYTES, &tm_lo, kb * KBE, n0, pb.full(s));
            }
          }
          __syncwarp();
        }
        // One max + one multiply-add per address: a padding tap (off = -1) reads nothing (0 source bytes = 16 bytes of zeros)
        // and is pointed at the chunk of pixel 0, a valid address.  The 64-bit select / add chain this replaces and a uniform
        // branch inside the loop were 13 issue slots per copy; the producers are bound by issue slots (tools/trace_conv.py:
        // ~95 cycles per copy).
        if (!(ablate & 32)) {                               // diagnostics bit 5: no input loads
          const int64_t lo_delta = xl - xh;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int o = off[it];
            const uint32_t nbytes = o >= 0 ? 16u : 0u;
            const uint32_t dst = a_hi + soff + it * 2048;
            const uint8_t* src = xh + (int64_t)max(o, 0) * S::ES;
            cp_async16(dst, src, nbytes);
            if (!BF16) cp_async16(dst + A_TILE_BYTES, src + lo_delta, nbytes);
          }
        }
        // arrive on `full` when this thread's copies have landed; the thread itself moves on to its next K block.
        // (The copies are generic-proxy writes: the issuer runs fence.proxy.async after its wait on `full`.)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(pb.full(s)) : "memory");
        if (t == 0) trace(gk, 3);
        c += NGROUP * KBE;
        while (c >= cin) { c -= cin; ++tap; if (++ts == p.s) { ts = 0; ++tr; } }
      }
    }
    }
  } else if (warp >= DRAIN_WARP0) {
    regs_take_drain();
    if constexpr (BN > 128) conv_epilogue_wide<BN, EPI>(p, pb, tmem_base, warp, lane, tiles_n, total_tiles, ablate);
    else conv_epilogue<BN, BF16, EPI>(p, sm, pb, tmem_base, warp, lane, nkb, tiles_n, total_tiles, ablate);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ============================================================================================
// weight-gradient kernel:  D[j = (tap,c)][n] = sum_pix X_gathered[pix][j] * dY[pix][n]
// Both operands are read as they lie in memory (pixel rows, channels contiguous) with 128-bit coalesced loads
// and stored MN-major; no transposition anywhere.
// ============================================================================================
template <int BN>
__global__ void __launch_bounds__(NTHREADS2, 1) wgrad_tc_kernel(const zsg_wgrad_params p, int kb_per_split) {
  using S = Smem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN;
  const int j0 = blockIdx.y * TM;
  const int Kt = p.r * p.s * p.cin;                         // rows of D
  const int nkb_total = (p.m + KB - 1) / KB;
  const int kb_begin = blockIdx.z * kb_per_split;
  int kb_end = kb_begin + kb_per_split;
  if (kb_end > nkb_total) kb_end = nkb_total;
  const int nkb = kb_end - kb_begin;                        // >= 1 by construction of the grid

  PipeBars pb = setup_pipeline<BN>(sm, warp, lane);
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + S::BAR_OFF + TMEM_SLOT_OFF);

  if (warp >= MMA_WARP) {
    if (elect_one()) {
      Issuer is = issuer_init<BN>(warp - MMA_WARP);
      mma_loop<BN, true>(sm, pb, tmem_base, nkb, is, p.impl == 7);
    }
    __syncwarp();
  } else if (warp < DRAIN_WARP0) {
    const int group = warp >> 2;
    const int t = tid & 127;
    const int mc = t & 31;                                  // 16-byte chunk (4 channels) of the 128-channel tile row
    const int ps = t >> 5;                                  // pixel sub-index 0..3
    const int atom_off = (mc >> 3) * 4096;                  // channel atom
    const int cj = mc & 7;
    // A side: this thread's 4 channels j..j+3 of D's row index = (tap, c)
    const int j = j0 + mc * 4;
    const bool jvalid = j < Kt;
    int tap = 0, c = 0, tr = 0, ts = 0;
    if (jvalid) { tap = j / p.cin; c = j - tap * p.cin; tr = tap / p.s; ts = tap - tr * p.s; }
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.in_scale && jvalid) {
      sc = __ldg(reinterpret_cast<const float4*>(p.in_scale + c));
      sh = __ldg(reinterpret_cast<const float4*>(p.in_shift + c));
    }
    // B side: 4 output channels n..n+3
    const int n = n0 + mc * 4;
    const bool bthread = mc * 4 < BN;
    const int4* rows = reinterpret_cast<const int4*>(p.rows);
    int4* ent = reinterpret_cast<int4*>(sm + S::ROWS_OFF) + group * 64;       // [2][32] entries per group
    int it = 0;
    for (int i = group; i < nkb; i += NGROUP, ++it) {
      const int s = i % S::STAGES;
      const int pix0 = (kb_begin + i) * KB;
      int4* eb = ent + (it & 1) * 32;
      if (t < 32) {                                          // stage this K block's 32 row entries
        int4 e = make_int4(0, 0, 0, 0);                      // hin = 0 => masked (also past the last pixel)
        if (pix0 + t < p.m) e = __ldg(rows + pix0 + t);
        eb[t] = e;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
      uint8_t* a_hi = sm + s * S::STAGE_BYTES;
      uint8_t* a_lo = a_hi + A_TILE_BYTES;
      uint8_t* b_hi = a_lo + A_TILE_BYTES;
      uint8_t* b_lo = b_hi + S::B_TILE_BYTES;
      bool waited = false;
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {                       // 2 x (4 pixels of A + 4 pixels of B) in flight
        float4 xa[4], yb[4];
        bool oka[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int pixel = (hp * 4 + q) * 4 + ps;           // 0..31 within the K block
          const int4 e = eb[pixel];
          const int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
          const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
          oka[q] = jvalid && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
          xa[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          yb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (oka[q]) xa[q] = __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)e.x + (int64_t)(yy * win + xx) * p.cin + c));
          if (bthread && hin > 0 && n < p.cout) {
            const float* src = p.dy + (int64_t)e.w + n;
            if (((e.w & 3) == 0) && n + 3 < p.cout) {
              yb[q] = __ldg(reinterpret_cast<const float4*>(src));
            } else {                                         // ragged channel count (45): scalar, bounded
              yb[q].x = __ldg(src);
              if (n + 1 < p.cout) yb[q].y = __ldg(src + 1);
              if (n + 2 < p.cout) yb[q].z = __ldg(src + 2);
              if (n + 3 < p.cout) yb[q].w = __ldg(src + 3);
            }
          }
        }
        if (!waited) { mbar_wait(pb.empty(s), ((i / S::STAGES) & 1) ^ 1); waited = true; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int pixel = (hp * 4 + q) * 4 + ps;
          const int r4 = pixel & 3;
          const int off = atom_off + (pixel >> 2) * 512 + r4 * 128 + ((((cj >> 1) ^ r4) << 1) | (cj & 1)) * 16;
          float4 v = xa[q];
          if (oka[q]) {
            if (p.in_scale) {
              v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
              v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
            }
            if (p.in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          }
          float4 h, l;
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
          *reinterpret_cast<float4*>(a_hi + off) = h;
          *reinterpret_cast<float4*>(a_lo + off) = l;
          if (bthread) {
            const float4 w = yb[q];
            split_tf32(w.x, h.x, l.x); split_tf32(w.y, h.y, l.y); split_tf32(w.z, h.z, l.z); split_tf32(w.w, h.w, l.w);
            *reinterpret_cast<float4*>(b_hi + off) = h;
            *reinterpret_cast<float4*>(b_lo + off) = l;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(pb.full(s));
    }
  } else {
    // drain + epilogue: lanes own consecutive j => coalesced reductions into dw[n][j]
    const int dw = warp - DRAIN_WARP0;
    const int quadrant = dw & 3, half = dw >> 2;
    float acc[BN / 2];
    int gchunk = 0;
    drain_loop<BN>(pb, tmem_base, nkb, 0, quadrant, half, acc, gchunk);
    const int j = j0 + quadrant * 32 + lane;
    if (j < Kt) {
#pragma unroll
      for (int q = 0; q < BN / 2; ++q) {
        const int nn = n0 + half * (BN / 2) + q;
        if (nn < p.cout) atomicAdd(p.dw + (int64_t)nn * Kt + j, acc[q]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ============================================================================================
// weight-gradient kernel, pre-split operands (p.x_lo and p.dy_lo given): same tiling, layouts and epilogue as
// wgrad_tc_kernel, but both operands and their TF32 remainders go global -> shared with cp.async (no register
// pass: the register-path producers ran ~850 SASS instructions per thread and K block), completion is signalled
// straight to `full` by cp.async.mbarrier.arrive, and the row entries of the next K block are fetched while the
// current one is in flight.
// ============================================================================================
// TMA_DY: dy / dy_lo are plain [m, dy_pitch] matrices (always true for the rows of a forward table): the B tile is
// fetched by TMA (four 32-pixel x 32-channel boxes per image, SWIZZLE_128B_ATOM_32B = the MN-major operand layout),
// which halves the cp.async instruction stream -- ncu showed the LSU, not the tensor pipe, pacing this kernel.
template <int BN, bool TMA_DY>
__global__ void __launch_bounds__(NTHREADS2, 1) wgrad_tc_async_kernel(const zsg_wgrad_params p, int kb_per_split,
                                                                      const __grid_constant__ CUtensorMap tm_dy,
                                                                      const __grid_constant__ CUtensorMap tm_dy_lo) {
  using S = Smem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN;
  const int j0 = blockIdx.y * TM;
  const int Kt = p.r * p.s * p.cin;                         // rows of D
  const int nkb_total = (p.m + KB - 1) / KB;
  const int kb_begin = blockIdx.z * kb_per_split;
  int kb_end = kb_begin + kb_per_split;
  if (kb_end > nkb_total) kb_end = nkb_total;
  const int nkb = kb_end - kb_begin;                        // >= 1 by construction of the grid

  PipeBars pb = setup_pipeline<BN>(sm, warp, lane, TMA_DY ? NPROD + 1 : NPROD);
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + S::BAR_OFF + TMEM_SLOT_OFF);

  if (warp >= MMA_WARP) {
    if (elect_one()) {
      Issuer is = issuer_init<BN>(warp - MMA_WARP);
      mma_loop<BN, true>(sm, pb, tmem_base, nkb, is);
    }
    __syncwarp();
  } else if (warp < DRAIN_WARP0) {
    regs_release_producer();
    const int group = warp >> 2;
    const int t = tid & 127;
    const int mc = t & 31;                                  // 16-byte chunk (4 channels) of the 128-channel tile row
    const int ps = t >> 5;                                  // pixel sub-index 0..3
    const uint32_t atom_off = (mc >> 3) * 4096;             // channel atom
    const int cj = mc & 7;
    // A side: this thread's 4 channels j..j+3 of D's row index = (tap, c)
    const int j = j0 + mc * 4;
    const bool jvalid = j < Kt;
    int tap = 0, c = 0, tr = 0, ts = 0;
    if (jvalid) { tap = j / p.cin; c = j - tap * p.cin; tr = tap / p.s; ts = tap - tr * p.s; }
    // B side: 4 output channels n..n+3 (ragged channel counts: only the valid bytes are copied, the rest is zero)
    const int n = n0 + mc * 4;
    const bool bthread = mc * 4 < BN && n < p.cout;
    const uint32_t bbytes = bthread ? (p.cout - n >= 4 ? 16u : 4u * (uint32_t)(p.cout - n)) : 0u;
    const float* xh = p.x + c;
    const float* xl = p.x_lo + c;
    const float* yh = p.dy + (bthread ? n : 0);
    const float* yl = p.dy_lo + (bthread ? n : 0);
    const int4* rows = reinterpret_cast<const int4*>(p.rows);
    int4* ent = reinterpret_cast<int4*>(sm + S::ROWS_OFF) + group * 64;       // [2][32] entries per group
    const uint32_t tiles0 = smem_u32(sm);
    // every group of 8 lanes (32 channels) lies inside one tap and is valid or not as a whole (cin % 32 == 0): each lane then
    // decodes ONE pixel per K block for its group's tap and the group exchanges the offsets (the per-lane decode of all 8
    // pixels was most of the producers' issue slots; with whole-warp uniformity only, the 64-channel layers fell back to it)
    const int tap0 = __shfl_sync(0xffffffffu, tap, lane & 24);      // (outside the &&: every lane must take part)
    const int jv0 = __shfl_sync(0xffffffffu, (int)jvalid, lane & 24);
    const bool warp_uniform = __all_sync(0xffffffffu, tap == tap0 && (int)jvalid == jv0) != 0;
    // prologue: entries of this group's first K block
    if (group < nkb) {
      if (t < 32) {
        int4 e = make_int4(0, 0, 0, 0);                      // hin = 0 => masked (also past the last pixel)
        const int pix = (kb_begin + group) * KB + t;
        if (pix < p.m) e = __ldg(rows + pix);
        ent[t] = e;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
    }
    int it = 0;
    for (int i = group; i < nkb; i += NGROUP, ++it) {
      const int s = i % S::STAGES;
      const int4* eb = ent + (it & 1) * 32;
      int4 e_next = make_int4(0, 0, 0, 0);                   // entries of my next K block: in flight during this one
      const bool has_next = i + NGROUP < nkb;
      if (has_next && t < 32) {
        const int pix = (kb_begin + i + NGROUP) * KB + t;
        if (pix < p.m) e_next = __ldg(rows + pix);
      }
      mbar_wait(pb.empty(s), ((i / S::STAGES) & 1) ^ 1, 5000 + i);
      const uint32_t a_hi = tiles0 + s * S::STAGE_BYTES;
      const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES;
      if (TMA_DY && (t >> 5) == 0) {
        if (elect_one()) {
          const int pix0 = (kb_begin + i) * KB;
          mbar_arrive_expect_tx(pb.full(s), 2 * S::B_TILE_BYTES);
#pragma unroll
          for (int atom = 0; atom < BN / 32; ++atom) {
            tma_load_2d(b_hi + atom * 4096, &tm_dy, n0 + atom * 32, pix0, pb.full(s));
            tma_load_2d(b_hi + S::B_TILE_BYTES + atom * 4096, &tm_dy_lo, n0 + atom * 32, pix0, pb.full(s));
          }
        }
        __syncwarp();
      }
      if (warp_uniform) {
        // lane (group, q) decodes pixel q * 4 + ps once for its group's tap, the group exchanges the offsets.  (The register
        // prefetch of the entry that the bf16 kernels use instead of this shared-memory staging made this kernel spill at its 56
        // producer registers and gained nothing: 213 TFLOP/s either way.)
        const int4 e = eb[(lane & 7) * 4 + ps];
        const int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
        const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
        const bool oka = jvalid && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
        const int my_ao = oka ? e.x + (yy * win + xx) * p.cin : -1;
        const int my_bo = hin > 0 ? e.w : -1;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int ao = __shfl_sync(0xffffffffu, my_ao, (lane & 24) + q), bo = __shfl_sync(0xffffffffu, my_bo, (lane & 24) + q);
          const int pixel = q * 4 + ps;
          const int r4 = ps;                                 // pixel & 3
          const uint32_t off = atom_off + q * 512 + r4 * 128 + ((((cj >> 1) ^ r4) << 1) | (cj & 1)) * 16;
          (void)pixel;
          const float* src = xh + max(ao, 0);               // padding: nothing is read (0 source bytes), any valid address
          cp_async16(a_hi + off, src, ao < 0 ? 0u : 16u);
          cp_async16(a_hi + A_TILE_BYTES + off, src + (xl - xh), ao < 0 ? 0u : 16u);
          if (!TMA_DY && mc * 4 < BN) {
            cp_async16(b_hi + off, yh + (bo < 0 ? 0 : bo), bo < 0 ? 0u : bbytes);
            cp_async16(b_hi + S::B_TILE_BYTES + off, yl + (bo < 0 ? 0 : bo), bo < 0 ? 0u : bbytes);
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int pixel = q * 4 + ps;                      // 0..31 within the K block
          const int4 e = eb[pixel];
          const int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
          const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
          const bool oka = jvalid && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
          const int r4 = pixel & 3;
          const uint32_t off = atom_off + (pixel >> 2) * 512 + r4 * 128 + ((((cj >> 1) ^ r4) << 1) | (cj & 1)) * 16;
          const int64_t ao = oka ? (int64_t)e.x + (int64_t)(yy * win + xx) * p.cin : 0;
          cp_async16(a_hi + off, xh + ao, oka ? 16u : 0u);
          cp_async16(a_hi + A_TILE_BYTES + off, xl + ao, oka ? 16u : 0u);
          if (!TMA_DY && mc * 4 < BN) {
            const bool okb = hin > 0;
            const int64_t bo = okb ? (int64_t)e.w : 0;
            cp_async16(b_hi + off, yh + bo, okb ? bbytes : 0u);
            cp_async16(b_hi + S::B_TILE_BYTES + off, yl + bo, okb ? bbytes : 0u);
          }
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(pb.full(s)) : "memory");
      if (has_next) {
        if (t < 32) ent[((it + 1) & 1) * 32 + t] = e_next;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
      }
    }
  } else {
    // drain + epilogue: lanes own consecutive j => coalesced reductions into dw[n][j]
    regs_take_drain();
    const int dw = warp - DRAIN_WARP0;
    const int quadrant = dw & 3, half = dw >> 2;
    float acc[BN / 2];
    int gchunk = 0;
    drain_loop<BN>(pb, tmem_base, nkb, 0, quadrant, half, acc, gchunk);
    const int j = j0 + quadrant * 32 + lane;
    if (j < Kt) {
#pragma unroll
      for (int q = 0; q < BN / 2; ++q) {
        const int nn = n0 + half * (BN / 2) + q;
        if (nn < p.cout) atomicAdd(p.dw + (int64_t)nn * Kt + j, acc[q]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ============================================================================================
// weight-gradient kernel, bf16 operand images (p.x_bf16, p.dy_bf16; BASELINE configs 3-5).  Same roles, barriers and
// drain as wgrad_tc_async_kernel; a K block is 64 pixels = four kind::f16 MMAs of K = 16.  Both operands are MN-major
// (channels contiguous per pixel, as they lie in memory): tile = [atom of 64 channels][64 pixels][128 B] with the plain
// 128-byte swizzle (16-byte chunk index ^= pixel & 7).  dy arrives by TMA (CU_TENSOR_MAP_SWIZZLE_128B writes exactly
// this layout, one 64 x 64 box per atom), the gathered x by cp.async (8 x 16 B per producer thread and K block).
// ============================================================================================
template <int BN>
__global__ void __launch_bounds__(NTHREADS2, 1) wgrad_bf16_kernel(const zsg_wgrad_params p, int kb_per_split,
                                                                  const __grid_constant__ CUtensorMap tm_dy) {
  using S = Smem<BN, true>;
  constexpr int KBP = 64;                                   // pixels per K block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN;
  const int j0 = blockIdx.y * TM;
  const int Kt = p.r * p.s * p.cin;                         // rows of D
  const int nkb_total = (p.m + KBP - 1) / KBP;
  const int kb_begin = blockIdx.z * kb_per_split;
  int kb_end = kb_begin + kb_per_split;
  if (kb_end > nkb_total) kb_end = nkb_total;
  const int nkb = kb_end - kb_begin;                        // >= 1 by construction of the grid

  PipeBars pb = setup_pipeline<BN, true>(sm, warp, lane, NPROD + 1);
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + S::BAR_OFF + TMEM_SLOT_OFF);

  if (warp >= MMA_WARP) {
    if (elect_one()) {
      Issuer is = issuer_init<BN, true>(warp - MMA_WARP);
      mma_loop<BN, true, true>(sm, pb, tmem_base, nkb, is);
    }
    __syncwarp();
  } else if (warp < DRAIN_WARP0) {
    regs_release_producer();
    const int group = warp >> 2;
    const int t = tid & 127;
    const int mc = t & 15;                                  // 16-byte chunk (8 channels) of the 128-channel tile row
    const int ps = t >> 4;                                  // pixel sub-index 0..7 (= pixel & 7: the swizzle key)
    const uint32_t toff = (mc >> 3) * 8192 + ps * 128 + (((mc & 7) ^ ps) << 4);   // + q * 1024 (8 pixels x 128 B)
    // A side: this thread's 8 channels j..j+7 of D's row index = (tap, c)
    const int j = j0 + mc * 8;
    const bool jvalid = j < Kt;
    int tap = 0, c = 0, tr = 0, ts = 0;
    if (jvalid) { tap = j / p.cin; c = j - tap * p.cin; tr = tap / p.s; ts = tap - tr * p.s; }
    const uint16_t* xb = p.x_bf16 + c;
    // every group of 8 lanes (64 channels) inside one tap, valid or not as a whole (cin % 64 == 0): see the fp32 kernel
    const int tap0 = __shfl_sync(0xffffffffu, tap, lane & 24);      // (outside the &&: every lane must take part)
    const int jv0 = __shfl_sync(0xffffffffu, (int)jvalid, lane & 24);
    const bool warp_uniform = __all_sync(0xffffffffu, tap == tap0 && (int)jvalid == jv0) != 0;
    const int4* rows = reinterpret_cast<const int4*>(p.rows);
    int4* ent = reinterpret_cast<int4*>(sm + S::ROWS_OFF) + group * 128;      // [2][64] entries per group
    const uint32_t tiles0 = smem_u32(sm);
    // prologue: entries of this group's first K block (staged through shared memory; the register prefetch of the wide kernel
    // makes the 64-column instance of this one spill at its 56 producer registers)
    if (group < nkb) {
      if (t < KBP) {
        int4 e = make_int4(0, 0, 0, 0);                      // hin = 0 => masked (also past the last pixel)
        const int pix = (kb_begin + group) * KBP + t;
        if (pix < p.m) e = __ldg(rows + pix);
        ent[t] = e;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
    }
    int it = 0;
    for (int i = group; i < nkb; i += NGROUP, ++it) {
      const int s = i % S::STAGES;
      const int4* eb = ent + (it & 1) * KBP;
      int4 e_next = make_int4(0, 0, 0, 0);                   // entries of my next K block: in flight during this one
      const bool has_next = i + NGROUP < nkb;
      if (has_next && t < KBP) {
        const int pix = (kb_begin + i + NGROUP) * KBP + t;
        if (pix < p.m) e_next = __ldg(rows + pix);
      }
      mbar_wait(pb.empty(s), ((i / S::STAGES) & 1) ^ 1, 5000 + i);
      const uint32_t a_tile = tiles0 + s * S::STAGE_BYTES;
      const uint32_t b_tile = a_tile + A_TILE_BYTES;
      if ((t >> 5) == 0) {
        if (elect_one()) {
          const int pix0 = (kb_begin + i) * KBP;
          mbar_arrive_expect_tx(pb.full(s), S::B_TILE_BYTES);
#pragma unroll
          for (int atom = 0; atom < BN / 64; ++atom) tma_load_2d(b_tile + atom * 8192, &tm_dy, n0 + atom * 64, pix0, pb.full(s));
        }
        __syncwarp();
      }
      if (warp_uniform) {
        // each lane decodes ONE pixel (row q = lane & 7 of its own column ps) for its 8-lane group's tap and the group
        // exchanges the offsets -- eight decodes per lane were most of the producers' issue slots
        const int4 e = eb[(lane & 7) * 8 + ps];          // my group's tap, my pixel column ps, pixel row q = lane & 7
        const int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
        const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
        const bool oka = jvalid && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
        const int my_ao = oka ? e.x + (yy * win + xx) * p.cin : -1;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int ao = __shfl_sync(0xffffffffu, my_ao, (lane & 24) + q);
          cp_async16(a_tile + toff + q * 1024, xb + max(ao, 0), ao < 0 ? 0u : 16u);   // padding: 0 bytes from a valid address
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int4 e = eb[q * 8 + ps];                     // pixel q * 8 + ps of the K block
          const int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
          const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
          const bool oka = jvalid && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
          const int64_t ao = oka ? (int64_t)e.x + (int64_t)(yy * win + xx) * p.cin : 0;
          cp_async16(a_tile + toff + q * 1024, xb + ao, oka ? 16u : 0u);
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(pb.full(s)) : "memory");
      if (has_next) {
        if (t < KBP) ent[((it + 1) & 1) * KBP + t] = e_next;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
      }
    }
  } else {
    // drain + epilogue: lanes own consecutive j => coalesced reductions into dw[n][j]
    regs_take_drain();
    const int dw = warp - DRAIN_WARP0;
    const int quadrant = dw & 3, half = dw >> 2;
    float acc[BN / 2];
    int gchunk = 0;
    drain_loop<BN, true>(pb, tmem_base, nkb, 0, quadrant, half, acc, gchunk);
    const int j = j0 + quadrant * 32 + lane;
    if (j < Kt) {
#pragma unroll
      for (int q = 0; q < BN / 2; ++q) {
        const int nn = n0 + half * (BN / 2) + q;
        if (nn < p.cout) atomicAdd(p.dw + (int64_t)nn * Kt + j, acc[q]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ============================================================================================
// bf16 weight gradient over 256-column tiles ("wide", see mma_loop_wide): the gathered x tile -- the side that costs
// producer time -- feeds twice the MMA work.  Persistent over work units (column tile, row tile, pixel split); a unit is at
// most WG_UNIT_KB K blocks, accumulated in TMEM without promotion (256 MMAs: the truncation bias stays at a few 1e-6), one
// issuing thread, two accumulators by unit parity, and the drain warps reduce a finished unit into dw (fp32 red.add, lanes
// along j) while the MMAs of the next unit run.
// ============================================================================================
constexpr int WG_UNIT_KB = 64;
__global__ void __launch_bounds__(NTHREADS2, 1) wgrad_bf16_wide_kernel(const zsg_wgrad_params p, int kb_per_split, int splits,
                                                                       const __grid_constant__ CUtensorMap tm_dy) {
  constexpr int BN = 256;
  using S = Smem<BN, true>;
  constexpr int KBP = 64;                                   // pixels per K block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kt = p.r * p.s * p.cin;                         // rows of D
  const int tiles_n = (p.cout + BN - 1) / BN, tiles_j = (Kt + TM - 1) / TM;
  const int nkb_total = (p.m + KBP - 1) / KBP;
  // unit u: row tile fastest, pixel split slowest -- the CTAs running at the same time then read the SAME pixels of x and dy
  // (all (tap, c) row tiles of a few splits): with the split fastest every pixel range came back from DRAM once per row tile
  // (ncu: 2.05 GB read per launch of the 3x3 256->256 layer at a 39 % L2 hit rate, DRAM 68 % busy; algorithmic 254 MB)
  const int units = tiles_n * tiles_j * splits;
  PipeBars pb = setup_pipeline<BN, true>(sm, warp, lane, NPROD + 1);
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + S::BAR_OFF + TMEM_SLOT_OFF);
#define WG_UNIT(u)                                                                         \
  const int tj = (u) % tiles_j, tn = ((u) / tiles_j) % tiles_n, sp = (u) / (tiles_j * tiles_n); \
  const int kb_begin = sp * kb_per_split;                                                  \
  const int nkb = (kb_begin + kb_per_split > nkb_total ? nkb_total : kb_begin + kb_per_split) - kb_begin; \
  const int n0 = tn * BN, j0 = tj * TM;

  if (warp >= MMA_WARP) {
    if (warp == MMA_WARP && elect_one()) {
      const uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);                  // MN-major: SBO = 1024 B, 128-byte swizzle
      const uint32_t lo0 = ((smem_u32(sm) >> 4) & 0x3FFFu) | (512u << 16);     // LBO = 8192 B between channel atoms
      uint32_t stage = 0, phase = 0, nu = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++nu) {
        WG_UNIT(u)
        (void)n0; (void)j0;
        const uint32_t a = nu & 1u;
        mbar_wait(pb.acc_empty(a), ((nu >> 1) & 1u) ^ 1u, 100 + nu);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(pb.full(stage), phase, 1000 + kb);
          fence_proxy_async();
          tc_fence_after();
          issue_kblock_bf16<BN, true>(tmem_base + a * BN, lo0 + stage * (uint32_t)(S::STAGE_BYTES >> 4), desc_hi, kb == 0 ? 0u : 1u);
          umma_commit(pb.empty(stage));
          if (++stage == (uint32_t)S::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(pb.acc_full(a));
      }
    }
    __syncwarp();
  } else if (warp < DRAIN_WARP0) {
    regs_release_producer();
    const int group = warp >> 2;
    const int t = tid & 127;
    const int mc = t & 15;                                  // 16-byte chunk (8 channels) of the 128-channel tile row
    const int ps = t >> 4;                                  // pixel sub-index 0..7 (= pixel & 7: the swizzle key)
    const uint32_t toff = (mc >> 3) * 8192 + ps * 128 + (((mc & 7) ^ ps) << 4);   // + q * 1024 (8 pixels x 128 B)
    const int4* rows = reinterpret_cast<const int4*>(p.rows);
    int4* ent = reinterpret_cast<int4*>(sm + S::ROWS_OFF) + group * 128;      // [2][64] entries per group
    const uint32_t tiles0 = smem_u32(sm);
    int g0 = 0;                                             // K blocks of this CTA before the current unit (ring position)
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      WG_UNIT(u)
      const int j = j0 + mc * 8;                            // this thread's 8 channels j..j+7 of D's row index = (tap, c)
      const bool jvalid = j < Kt;
      int tap = 0, c = 0, tr = 0, ts = 0;
      if (jvalid) { tap = j / p.cin; c = j - tap * p.cin; tr = tap / p.s; ts = tap - tr * p.s; }
      const uint16_t* xb = p.x_bf16 + c;
      const int tap0 = __shfl_sync(0xffffffffu, tap, lane & 24);     // groups of 8 lanes (64 channels): see the fp32 kernel
      const int jv0 = __shfl_sync(0xffffffffu, (int)jvalid, lane & 24);
      const bool warp_uniform = __all_sync(0xffffffffu, tap == tap0 && (int)jvalid == jv0) != 0;
      const int first = (group - g0) & 1;                   // K blocks with (g0 + i) % NGROUP == group
      // row entries: register prefetch on the shared-decode path, shared-memory staging otherwise (see wgrad_bf16_kernel)
      const int my_pl = (lane & 7) * 8 + ps;
      int4 e_cur = make_int4(0, 0, 0, 0);                    // hin = 0 => masked (also past the last pixel)
      if (warp_uniform) {
        if (first < nkb) {
          const int pix = (kb_begin + first) * KBP + my_pl;
          if (pix < p.m) e_cur = __ldg(rows + pix);
        }
      } else {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");   // the previous unit's entries have been read
        if (first < nkb && t < KBP) {
          int4 e = make_int4(0, 0, 0, 0);
          const int pix = (kb_begin + first) * KBP + t;
          if (pix < p.m) e = __ldg(rows + pix);
          ent[t] = e;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
      }
      int it = 0;
      for (int i = first; i < nkb; i += NGROUP, ++it) {
        const int g = g0 + i;
        const int s = g % S::STAGES;
        const int4* eb = ent + (it & 1) * KBP;
        int4 e_next = make_int4(0, 0, 0, 0);                 // entries of my next K block: in flight during this one
        const bool has_next = i + NGROUP < nkb;
        if (has_next && (warp_uniform || t < KBP)) {
          const int pix = (kb_begin + i + NGROUP) * KBP + (warp_uniform ? my_pl : t);
          if (pix < p.m) e_next = __ldg(rows + pix);
        }
        mbar_wait(pb.empty(s), ((g / S::STAGES) & 1) ^ 1, 5000 + i);
        const uint32_t a_tile = tiles0 + s * S::STAGE_BYTES;
        const uint32_t b_tile = a_tile + A_TILE_BYTES;
        if ((t >> 5) == 0) {
          if (elect_one()) {
            const int pix0 = (kb_begin + i) * KBP;
            mbar_arrive_expect_tx(pb.full(s), S::B_TILE_BYTES);
#pragma unroll
            for (int atom = 0; atom < BN / 64; ++atom) tma_load_2d(b_tile + atom * 8192, &tm_dy, n0 + atom * 64, pix0, pb.full(s));
          }
          __syncwarp();
        }
        if (warp_uniform) {
          // each lane decodes ONE pixel (row q = lane & 7 of its own column ps) for its 8-lane group's tap and the group
          // exchanges the offsets -- eight decodes per lane were most of the producers' issue slots
          const int4 e = e_cur;                            // my group's tap, my pixel column ps, pixel row q = lane & 7
          const int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
          const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
          const bool oka = jvalid && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
          const int my_ao = oka ? e.x + (yy * win + xx) * p.cin : -1;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int ao = __shfl_sync(0xffffffffu, my_ao, (lane & 24) + q);
            cp_async16(a_tile + toff + q * 1024, xb + max(ao, 0), ao < 0 ? 0u : 16u);   // padding: 0 bytes from a valid address
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int4 e = eb[q * 8 + ps];                     // pixel q * 8 + ps of the K block
            const int yy = (int)(short)(e.y & 0xFFFF) + tr * p.dil, xx = (e.y >> 16) + ts * p.dil;
            const int hin = (int)(short)(e.z & 0xFFFF), win = e.z >> 16;
            const bool oka = jvalid && (unsigned)yy < (unsigned)hin && (unsigned)xx < (unsigned)win;
            const int64_t ao = oka ? (int64_t)e.x + (int64_t)(yy * win + xx) * p.cin : 0;
            cp_async16(a_tile + toff + q * 1024, xb + ao, oka ? 16u : 0u);
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(pb.full(s)) : "memory");
        if (warp_uniform) {
          e_cur = e_next;
        } else if (has_next) {
          if (t < KBP) ent[((it + 1) & 1) * KBP + t] = e_next;
          asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(NPROD) : "memory");
        }
      }
      g0 += nkb;
    }
  } else {
    // drain + epilogue: lanes own consecutive j => coalesced reductions into dw[n][j]
    regs_take_drain();
    const int dw = warp - DRAIN_WARP0;
    const int quadrant = dw & 3, half = dw >> 2;
    uint32_t nu = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++nu) {
      WG_UNIT(u)
      (void)nkb;
      const uint32_t a = nu & 1u;
      mbar_wait(pb.acc_full(a), (nu >> 1) & 1u, 2000 + nu);
      tc_fence_after();
      const int j = j0 + quadrant * 32 + lane;
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        float acc[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(quadrant * 32) << 16) + a * BN + (uint32_t)(half * 128 + jp * 64);
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) tmem_ld16_nowait(taddr + cb * 16, acc + cb * 16);
        tmem_ld_wait();
        if (jp == 1) {
          tc_fence_before();
          mbar_arrive(pb.acc_empty(a));
        }
        if (j < Kt) {
#pragma unroll
          for (int q = 0; q < 64; ++q) {
            const int nn = n0 + half * 128 + jp * 64 + q;
            if (nn < p.cout) atomicAdd(p.dw + (int64_t)nn * Kt + j, acc[q]);
          }
        }
      }
    }
  }
#undef WG_UNIT
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ============================================================================================
// SIMT check kernels (tests only): the same gather semantics in plain fp32 FMAs, one thread
// per output element.  Independent of every tcgen05 / smem-layout assumption above.
// ============================================================================================
__global__ void conv_simt_kernel(const zsg_conv_params p) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)p.m * p.cout) return;
  const int n = (int)(idx % p.cout);
  const int m = (int)(idx / p.cout);
  const zsg_row_t e = p.rows[m];
  const int K = p.r * p.s * p.cin;
  float acc = 0.f;
  for (int tr = 0; tr < p.r; ++tr)
    for (int ts = 0; ts < p.s; ++ts) {
      int yy = e.y0 + tr * p.dil, xx = e.x0 + ts * p.dil;
      if (p.in_div == 2) {
        if ((yy | xx) & 1) continue;
        yy >>= 1;
        xx >>= 1;
      }
      if ((unsigned)yy >= (unsigned)e.hin || (unsigned)xx >= (unsigned)e.win) continue;
      const float* xp = p.x + (int64_t)e.base + (int64_t)(yy * e.win + xx) * p.cin;
      const float* wp = p.w + (int64_t)n * K + (tr * p.s + ts) * p.cin;
      for (int c = 0; c < p.cin; ++c) {
        float v = xp[c];
        if (p.in_scale) v = fmaf(v, p.in_scale[c], p.in_shift[c]);
        if (p.in_relu) v = fmaxf(v, 0.f);
        acc = fmaf(v, p.w_lo ? wp[c] + p.w_lo[(int64_t)n * K + (tr * p.s + ts) * p.cin + c] : wp[c], acc);
      }
    }
  if (p.bias) acc += p.bias[n];
  if (p.out_mask && !(p.out_mask[(int64_t)e.out + n] > 0.f)) acc = 0.f;
  if (p.residual) acc += p.residual[(int64_t)e.out + n];
  if (p.accumulate) acc += p.y[(int64_t)e.out + n];
  if (p.out_relu) acc = fmaxf(acc, 0.f);
  p.y[(int64_t)e.out + n] = acc;
}

__global__ void wgrad_simt_kernel(const zsg_wgrad_params p) {
  const int Kt = p.r * p.s * p.cin;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)p.cout * Kt) return;
  const int j = (int)(idx % Kt);
  const int n = (int)(idx / Kt);
  const int tap = j / p.cin, c = j - tap * p.cin, tr = tap / p.s, ts = tap - tr * p.s;
  float acc = 0.f;
  for (int m = 0; m < p.m; ++m) {
    const zsg_row_t e = p.rows[m];
    const int yy = e.y0 + tr * p.dil, xx = e.x0 + ts * p.dil;
    if ((unsigned)yy >= (unsigned)e.hin || (unsigned)xx >= (unsigned)e.win) continue;
    float v = p.x[(int64_t)e.base + (int64_t)(yy * e.win + xx) * p.cin + c];
    if (p.in_scale) v = fmaf(v, p.in_scale[c], p.in_shift[c]);
    if (p.in_relu) v = fmaxf(v, 0.f);
    acc = fmaf(v, p.dy[(int64_t)e.out + n], acc);
  }
  p.dw[idx] += acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// weights [cout][K] fp32, K contiguous: box = 32 K elements (128 B, swizzle 128B) x BN rows; OOB reads give 0
static int make_weight_map(CUtensorMap* map, const float* w, int cout, int K, int bn) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return ZSG_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for cout=%d K=%d", (int)r, cout, K); return ZSG_ECUDA; }
  return ZSG_OK;
}

template <int BN>
static int launch_conv_generic(const zsg_conv_params& p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_generic_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Smem<BN>::TOTAL);
    if (e != cudaSuccess) { set_error("conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ZSG_ECUDA; }
    attr_done = true;
  }
  CUtensorMap tm_hi, tm_lo;
  memset(&tm_hi, 0, sizeof(tm_hi));
  memset(&tm_lo, 0, sizeof(tm_lo));
  const int total_tiles = ((p.cout + BN - 1) / BN) * ((p.m + TM - 1) / TM);
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();      // persistent: one CTA per SM
  conv_tc_generic_kernel<BN, false><<<grid, NTHREADS2, Smem<BN>::TOTAL, st>>>(p, tm_hi, tm_lo);
  return check_launch("zsg_conv_fwd");
}

template <int BN, int PRO>
static int launch_conv(const zsg_conv_params& p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, PRO>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::TOTAL);
    if (e != cudaSuccess) { set_error("conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ZSG_ECUDA; }
    attr_done = true;
  }
  CUtensorMap tm_hi, tm_lo;
  const int K = p.r * p.s * p.cin;
  if (int rc = make_weight_map(&tm_hi, p.w, p.cout, K, BN)) return rc;
  if (int rc = make_weight_map(&tm_lo, p.w_lo, p.cout, K, BN)) return rc;
  const int total_tiles = ((p.cout + BN - 1) / BN) * ((p.m + TM - 1) / TM);
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();      // persistent: one CTA per SM
  conv_tc_kernel<BN, PRO><<<grid, NTHREADS2, Smem<BN>::TOTAL, st>>>(p, tm_hi, tm_lo);
  return check_launch("zsg_conv_fwd");
}

static inline int pick_epilogue(const zsg_conv_params& p) {
  static const bool fragx = [] { const char* e = getenv("ZSG_EPI_FRAGX"); return !(e && e[0] == '0'); }();   // 0: slab epilogue (A/B runs)
  if (p.y_bf16) return EPI_B16;
  if (epilogue_is_plain(p) && !p.y_lo && !p.y_img_bf16) return EPI_PLAIN;
  const uintptr_t ptrs = (uintptr_t)p.y | (uintptr_t)p.bias | (uintptr_t)p.out_mask | (uintptr_t)p.residual | (uintptr_t)p.row_add |
                         (uintptr_t)p.y_lo;
  return (fragx && p.cout % 8 == 0 && p.y_pitch > 0 && p.y_pitch % 2 == 0 && (ptrs & 7) == 0 && (((uintptr_t)p.residual_bf16 | (uintptr_t)p.y_img_bf16) & 3) == 0)
             ? EPI_FRAGX : EPI_GENERIC;
}

template <int BN, int EPI>
static int launch_conv_async_epi(const zsg_conv_params& p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_async_kernel<BN, false, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::TOTAL);
    if (e != cudaSuccess) { set_error("conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ZSG_ECUDA; }
    attr_done = true;
  }
  CUtensorMap tm_hi, tm_lo, tm_a, tm_a_lo;
  const int K = p.r * p.s * p.cin;
  if (int rc = make_weight_map(&tm_hi, p.w, p.cout, K, BN)) return rc;
  if (int rc = make_weight_map(&tm_lo, p.w_lo, p.cout, K, BN)) return rc;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_a_lo, 0, sizeof(tm_a_lo));
  if (p.x_plain) {                                          // A = x as a plain [m, cin] matrix: box of 32 channels x 128 rows
    if (int rc = make_weight_map(&tm_a, p.x, p.m, p.cin, TM)) return rc;
    if (int rc = make_weight_map(&tm_a_lo, p.x_lo, p.m, p.cin, TM)) return rc;
  }
  const int total_tiles = ((p.cout + BN - 1) / BN) * ((p.m + TM - 1) / TM);
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();      // persistent: one CTA per SM
  conv_tc_async_kernel<BN, false, EPI><<<grid, NTHREADS2, Smem<BN>::TOTAL, st>>>(p, tm_hi, tm_lo, tm_a, tm_a_lo);
  return check_launch("zsg_conv_fwd");
}
template <int BN>
static int launch_conv_async(const zsg_conv_params& p, cudaStream_t st) {
  switch (pick_epilogue(p)) {
    case EPI_PLAIN: return launch_conv_async_epi<BN, EPI_PLAIN>(p, st);
    case EPI_B16: return launch_conv_async_epi<BN, EPI_B16>(p, st);
    case EPI_FRAGX: return launch_conv_async_epi<BN, EPI_FRAGX>(p, st);
    default: return launch_conv_async_epi<BN, EPI_GENERIC>(p, st);
  }
}

// bf16 matrices for TMA: [rows][inner] bf16, inner contiguous; box = 64 inner elements (128 B, swizzle 128B) x box_rows.
// Weights [cout][K] (K-major B operand of the forward kernel) and dy [m][pitch] (MN-major B operand of the weight
// gradient) use the same encoding; out-of-range reads give 0.
static int make_bf16_map(CUtensorMap* map, const uint16_t* base, int rows, int inner, int box_rows, const char* what) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return ZSG_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)inner * sizeof(uint16_t)};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for bf16 %s rows=%d inner=%d", (int)r, what, rows, inner); return ZSG_ECUDA; }
  return ZSG_OK;
}

template <int BN, int EPI>
static int launch_conv_bf16_epi(const zsg_conv_params& p, cudaStream_t st) {
  using S = Smem<BN, true>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_async_kernel<BN, true, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) { set_error("conv(bf16): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ZSG_ECUDA; }
    attr_done = true;
  }
  CUtensorMap tm_w, tm_unused;
  memset(&tm_unused, 0, sizeof(tm_unused));
  const int K = p.r * p.s * p.cin;
  if (int rc = make_bf16_map(&tm_w, p.w_bf16, p.cout, K, BN, "weights")) return rc;
  CUtensorMap tm_a;
  memset(&tm_a, 0, sizeof(tm_a));
  if (p.x_plain)                                            // A = x_bf16 as a plain [m, cin] matrix: 64 channels x 128 rows
    if (int rc = make_bf16_map(&tm_a, p.x_bf16, p.m, p.cin, TM, "x")) return rc;
  const int total_tiles = ((p.cout + BN - 1) / BN) * ((p.m + TM - 1) / TM);
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();      // persistent: one CTA per SM
  conv_tc_async_kernel<BN, true, EPI><<<grid, NTHREADS2, S::TOTAL, st>>>(p, tm_w, tm_unused, tm_a, tm_unused);
  return check_launch("zsg_conv_fwd(bf16)");
}
// 256-column tiles when the layer has them and they pay: twice the MMA work per gathered A tile and the full issue rate, but
// half as many tiles -- the partial last wave of the persistent grid weighs more.  wide_gain = measured speed ratio of the two
// kernels on full waves (tools/time_big.py).
static bool use_wide_bf16(const zsg_conv_params& p, int epi) {
  static const int mode = [] { const char* e = getenv("ZSG_WIDE_BF16"); return e ? atoi(e) : 1; }();   // 0 off, 1 heuristic, 2 whenever legal
  if (epi == EPI_GENERIC || p.cout % 256 != 0 || p.impl == 2) return false;      // impl 2 / 3 (tests): 128- / 256-column tiles
  if (p.impl == 3) return true;
  if (mode == 0) return false;
  if (mode == 2) return true;
  const double tm = (p.m + TM - 1) / TM, sms = num_sms();
  const double t128 = tm * (p.cout / 128), t256 = tm * (p.cout / 256);
  const double eff128 = t128 / (ceil(t128 / sms) * sms), eff256 = t256 / (ceil(t256 / sms) * sms);
  const double wide_gain = (p.r * p.s > 1) ? 1.4 : 1.1;    // gather-fed 3x3 layers 1.3-1.5x, TMA-fed 1x1 layers 1.0-1.1x (epilogue-bound)
  return eff256 * wide_gain > eff128;
}

template <int BN>
static int launch_conv_bf16(const zsg_conv_params& p, cudaStream_t st) {
  if constexpr (BN == 128) {
    const int epi = pick_epilogue(p);
    if (use_wide_bf16(p, epi)) {
      switch (epi) {
        case EPI_PLAIN: return launch_conv_bf16_epi<256, EPI_PLAIN>(p, st);
        case EPI_B16: return launch_conv_bf16_epi<256, EPI_B16>(p, st);
        default: return launch_conv_bf16_epi<256, EPI_FRAGX>(p, st);
      }
    }
  }
  switch (pick_epilogue(p)) {
    case EPI_PLAIN: return launch_conv_bf16_epi<BN, EPI_PLAIN>(p, st);
    case EPI_B16: return launch_conv_bf16_epi<BN, EPI_B16>(p, st);
    case EPI_FRAGX: return launch_conv_bf16_epi<BN, EPI_FRAGX>(p, st);
    default: return launch_conv_bf16_epi<BN, EPI_GENERIC>(p, st);
  }
}

template <int BN>
static int launch_conv_pro(const zsg_conv_params& p, cudaStream_t st) {
  switch ((p.in_scale ? 2 : 0) | (p.in_relu ? 1 : 0)) {
    case 0: return launch_conv<BN, 0>(p, st);
    case 1: return launch_conv<BN, 1>(p, st);
    case 2: return launch_conv<BN, 2>(p, st);
    default: return launch_conv<BN, 3>(p, st);
  }
}

// dy [m][pitch] fp32: box = 32 channels (128 B) x 32 pixels, 32-byte-atom 128 B swizzle (MN-major tf32 operand layout)
static int make_dy_map(CUtensorMap* map, const float* dy, int m, int pitch) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return ZSG_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)m};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
  cuuint32_t box[2] = {32, (cuuint32_t)KB};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dy), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for dy m=%d pitch=%d", (int)r, m, pitch); return ZSG_ECUDA; }
  return ZSG_OK;
}

// split-K factor shared by the weight-gradient launchers: minimise (waves of CTAs) x (K blocks per CTA + fixed per-CTA
// cost), so that the grid fills whole waves of the SMs while every CTA keeps at least 8 K blocks
static int choose_split_k(int tiles, int nkb, int fixed) {
  const int sms = num_sms();
  const int max_split = nkb >= 16 ? nkb / 8 : 1;
  long best_cost = -1;
  int split = 1;
  for (int sk = 1; sk <= max_split && sk <= 256; ++sk) {
    const int per_cta = (nkb + sk - 1) / sk;
    const int ctas = tiles * ((nkb + per_cta - 1) / per_cta);
    const long waves = (ctas + sms - 1) / sms;
    const long cost = waves * (per_cta + fixed);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; split = sk; }
  }
  return split;
}

template <int BN, int MODE>          // MODE 0 = register path, 1 = cp.async operands, 2 = cp.async x + TMA dy
static int launch_wgrad(const zsg_wgrad_params& p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = MODE == 0 ? cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::TOTAL)
                  : MODE == 1 ? cudaFuncSetAttribute(wgrad_tc_async_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::TOTAL)
                              : cudaFuncSetAttribute(wgrad_tc_async_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::TOTAL);
    if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ZSG_ECUDA; }
    attr_done = true;
  }
  CUtensorMap tm_dy, tm_dy_lo;
  memset(&tm_dy, 0, sizeof(tm_dy));
  memset(&tm_dy_lo, 0, sizeof(tm_dy_lo));
  if (MODE == 2) {
    if (int rc = make_dy_map(&tm_dy, p.dy, p.m, p.dy_pitch)) return rc;
    if (int rc = make_dy_map(&tm_dy_lo, p.dy_lo, p.m, p.dy_pitch)) return rc;
  }
  const int Kt = p.r * p.s * p.cin;
  const int tiles = ((p.cout + BN - 1) / BN) * ((Kt + TM - 1) / TM);
  const int nkb = (p.m + KB - 1) / KB;
  int split = p.split_k > 0 ? p.split_k : choose_split_k(tiles, nkb, 24);   // fixed = pipeline fill + atomic epilogue, in K-block times
  if (split > nkb) split = nkb;
  const int per = (nkb + split - 1) / split;
  split = (nkb + per - 1) / per;                            // no empty splits
  dim3 grid((p.cout + BN - 1) / BN, (Kt + TM - 1) / TM, split);
  if (MODE == 0) wgrad_tc_kernel<BN><<<grid, NTHREADS2, Smem<BN>::TOTAL, st>>>(p, per);
  else if (MODE == 1) wgrad_tc_async_kernel<BN, false><<<grid, NTHREADS2, Smem<BN>::TOTAL, st>>>(p, per, tm_dy, tm_dy_lo);
  else wgrad_tc_async_kernel<BN, true><<<grid, NTHREADS2, Smem<BN>::TOTAL, st>>>(p, per, tm_dy, tm_dy_lo);
  return check_launch("zsg_conv_wgrad");
}

template <int BN>
static int launch_wgrad_bf16(const zsg_wgrad_params& p, cudaStream_t st) {
  using S = Smem<BN, true>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) { set_error("wgrad(bf16): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ZSG_ECUDA; }
    attr_done = true;
  }
  CUtensorMap tm_dy;
  if (int rc = make_bf16_map(&tm_dy, p.dy_bf16, p.m, p.dy_pitch, 64, "dy")) return rc;
  const int Kt = p.r * p.s * p.cin;
  const int tiles = ((p.cout + BN - 1) / BN) * ((Kt + TM - 1) / TM);
  const int nkb = (p.m + 63) / 64;
  int split = p.split_k > 0 ? p.split_k : choose_split_k(tiles, nkb, 40);   // a bf16 K block is a third of the MMA time of a
                                                                            // tf32 one: the fixed cost weighs more
  if (split > nkb) split = nkb;
  const int per = (nkb + split - 1) / split;
  split = (nkb + per - 1) / per;                            // no empty splits
  dim3 grid((p.cout + BN - 1) / BN, (Kt + TM - 1) / TM, split);
  wgrad_bf16_kernel<BN><<<grid, NTHREADS2, S::TOTAL, st>>>(p, per, tm_dy);
  return check_launch("zsg_conv_wgrad(bf16)");
}

// 256-column tiles for the bf16 weight gradient (wgrad_bf16_wide_kernel) when the layer has them: units of at most
// WG_UNIT_KB K blocks, the unit length chosen so that the persistent grid ends on a full wave.
static int launch_wgrad_bf16_wide(const zsg_wgrad_params& p, cudaStream_t st) {
  using S = Smem<256, true>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_bf16_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) { set_error("wgrad(bf16 wide): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ZSG_ECUDA; }
    attr_done = true;
  }
  CUtensorMap tm_dy;
  if (int rc = make_bf16_map(&tm_dy, p.dy_bf16, p.m, p.dy_pitch, 64, "dy")) return rc;
  const int Kt = p.r * p.s * p.cin;
  const int tiles = (p.cout / 256) * ((Kt + TM - 1) / TM);
  const int nkb = (p.m + 63) / 64;
  const int sms = num_sms();
  int best_per = WG_UNIT_KB;
  double best_cost = -1.0;
  for (int per = WG_UNIT_KB; per >= 24 && per >= 1; --per) {
    const int splits = (nkb + per - 1) / per;
    const double units = (double)tiles * splits, waves = ceil(units / sms);
    const double cost = waves * (per + 6);                  // + per-unit overhead in K-block times
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_per = per; }
  }
  int per = nkb < best_per ? nkb : best_per;
  const int splits = (nkb + per - 1) / per;
  const int units = tiles * splits;
  const int grid = units < sms ? units : sms;
  wgrad_bf16_wide_kernel<<<grid, NTHREADS2, S::TOTAL, st>>>(p, per, splits, tm_dy);
  return check_launch("zsg_conv_wgrad(bf16 wide)");
}
static bool use_wide_wgrad_bf16(const zsg_wgrad_params& p) {
  static const int mode = [] { const char* e = getenv("ZSG_WIDE_WGRAD_BF16"); return e ? atoi(e) : 1; }();   // 0 off, 1 heuristic, 2 whenever legal
  if (p.cout % 256 != 0 || p.impl == 2 || p.split_k > 0 || p.dy_pitch <= 0) return false;
  if (p.impl == 3 || mode == 2) return true;
  if (mode == 0) return false;
  const int Kt = p.r * p.s * p.cin;
  const double units = (double)(p.cout / 256) * ((Kt + TM - 1) / TM) * (((p.m + 63) / 64 + WG_UNIT_KB - 1) / WG_UNIT_KB);
  return units >= 32.0;                                     // measured (tools/step_time.py): wide wherever legal beats a two-wave threshold by 1 ms / step
}

}  // namespace zsg

using namespace zsg;

extern "C" int zsg_debug_set_conv_trace(unsigned int* buf, int nblocks) {
  cudaError_t e = cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_trace_blocks, &nblocks, sizeof(nblocks));
  if (e != cudaSuccess) { set_error("zsg_debug_set_conv_trace: %s", cudaGetErrorString(e)); return ZSG_ECUDA; }
  return ZSG_OK;
}

extern "C" int zsg_conv_fwd(const zsg_conv_params* pp, zsg_stream_t stream) {
  ZSG_REQUIRE(pp, "zsg_conv_fwd: null params");
  zsg_conv_params p = *pp;
  ZSG_REQUIRE(p.dil >= 0 && p.dil <= 64, "zsg_conv_fwd: dil=%d out of range", p.dil);
  if (p.dil == 0) p.dil = 1;
  ZSG_REQUIRE(((p.x && p.w) || (p.x_bf16 && p.w_bf16)) && (p.y || p.y_bf16) && p.rows, "zsg_conv_fwd: null pointer");
  ZSG_REQUIRE(!p.residual_bf16 || (!p.residual && p.cout % 4 == 0 && p.impl != 1 && ((uintptr_t)p.residual_bf16 & 7) == 0),
              "zsg_conv_fwd: residual_bf16 excludes residual, needs cout %% 4 == 0 (and row offsets %% 4 == 0) and the tcgen05 path");
  ZSG_REQUIRE(!p.y_bf16 || (!p.bias && !p.out_relu && !p.out_mask && !p.residual && !p.residual_bf16 && !p.accumulate && p.cout % 8 == 0 &&
                            p.impl != 1 && (p.x_lo || p.x_bf16) && ((uintptr_t)p.y_bf16 & 15) == 0),
              "zsg_conv_fwd: y_bf16 needs a plain output (no bias / ReLU / mask / residual / accumulate), cout %% 8 == 0, an "
              "operand-image input and the tcgen05 path");
  ZSG_REQUIRE(p.m > 0 && p.cout > 0 && p.r > 0 && p.s > 0, "zsg_conv_fwd: empty problem");
  ZSG_REQUIRE(p.cin > 0 && p.cin % 4 == 0, "zsg_conv_fwd: cin=%d must be a multiple of 4", p.cin);
  ZSG_REQUIRE(p.in_div == 1 || p.in_div == 2, "zsg_conv_fwd: in_div must be 1 or 2");
  ZSG_REQUIRE((((uintptr_t)p.x | (uintptr_t)p.w) & 15) == 0, "zsg_conv_fwd: x and w must be 16-byte aligned");
  ZSG_REQUIRE(p.impl != 1 || (p.x && p.w), "zsg_conv_fwd: the SIMT check kernel reads the fp32 operands");
  ZSG_REQUIRE(!p.in_scale || p.in_shift, "zsg_conv_fwd: in_scale without in_shift");
  ZSG_REQUIRE(!p.x_plain || (p.r == 1 && p.s == 1 && p.in_div == 1), "zsg_conv_fwd: x_plain needs r = s = 1 and in_div = 1");
  ZSG_REQUIRE(p.y_pitch == 0 || p.y_pitch >= p.cout, "zsg_conv_fwd: y_pitch=%d must be 0 or >= cout", p.y_pitch);
  ZSG_REQUIRE(!p.row_add == !p.row_add_idx, "zsg_conv_fwd: row_add and row_add_idx go together");
  ZSG_REQUIRE((!p.y_lo && !p.y_img_bf16) ||
                  (!(p.y_lo && p.y_img_bf16) && !p.y_bf16 && !p.stats && p.impl != 1 && (p.x_lo || p.x_bf16) && p.cout % 8 == 0 &&
                   p.y_pitch > 0 && p.y_pitch % 2 == 0 && ((uintptr_t)p.y & 7) == 0 && ((uintptr_t)p.y_lo & 7) == 0 &&
                   ((uintptr_t)p.y_img_bf16 & 3) == 0 && pick_epilogue(p) == EPI_FRAGX),
              "zsg_conv_fwd: y_lo / y_img_bf16 (one of them) need a plain [m, y_pitch] fp32 output with cout %% 8 == 0 and an even "
              "y_pitch on the operand-image tcgen05 paths (the fragment epilogue writes them)");
  ZSG_REQUIRE(!p.row_add || (p.cout % 4 == 0 && (p.x_lo || p.x_bf16) && p.impl != 1 && !p.y_bf16 && !p.stats &&
                             (((uintptr_t)p.row_add & 15) | ((uintptr_t)p.row_add_idx & 7)) == 0),
              "zsg_conv_fwd: row_add needs cout %% 4 == 0, an operand-image input, the tcgen05 path and aligned tables");
  ZSG_REQUIRE(!p.bnb_partials == !p.bnb_x && !p.bnb_partials == !p.bnb_scale && !p.bnb_partials == !p.bnb_shift,
              "zsg_conv_fwd: bnb_x / bnb_scale / bnb_shift / bnb_partials go together");
  ZSG_REQUIRE(!p.bnb_partials || (!p.stats && !p.y_lo && !p.y_img_bf16 && epilogue_is_plain(p) && p.impl != 1 && (p.x_lo || p.x_bf16) &&
                                  p.cout % 8 == 0 && p.y_pitch > 0 && p.y_pitch % 2 == 0 &&
                                  (((uintptr_t)p.bnb_x | (uintptr_t)p.bnb_scale | (uintptr_t)p.bnb_shift | (uintptr_t)p.bnb_partials) & 7) == 0),
              "zsg_conv_fwd: bnb_* (BatchNorm backward sums from a data gradient's epilogue) need a plain [m, y_pitch] output "
              "(fp32 y with fp32 bnb_x, or y_bf16 with bfloat16 bnb_x), cout %% 8 == 0, the operand-image tcgen05 paths");
  ZSG_REQUIRE(!p.stats || (!p.bias && !p.out_relu && !p.out_mask && !p.residual && !p.residual_bf16 && !p.accumulate && p.impl != 1),
              "zsg_conv_fwd: stats needs a plain output (no bias / ReLU / mask / residual / accumulate) on the tcgen05 path");
  cudaStream_t st = as_stream(stream);
  if (p.impl == 1) {
    const int64_t n = (int64_t)p.m * p.cout;
    conv_simt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p);
    return check_launch("zsg_conv_fwd(simt)");
  }
  if (!zsg_device_supported()) { set_error("zsg_conv_fwd: tcgen05 path needs an sm_100 device"); return ZSG_EARCH; }
  if (p.x_bf16 || p.w_bf16) {
    ZSG_REQUIRE(p.x_bf16 && p.w_bf16, "zsg_conv_fwd: x_bf16 and w_bf16 go together");
    ZSG_REQUIRE(p.cin % 8 == 0, "zsg_conv_fwd: the bf16 path needs cin=%d to be a multiple of 8", p.cin);
    ZSG_REQUIRE(!p.in_scale && !p.in_relu, "zsg_conv_fwd: bf16 images exclude the on-load affine / ReLU (apply them in zsg_cast_bf16)");
    ZSG_REQUIRE((((uintptr_t)p.x_bf16 | (uintptr_t)p.w_bf16) & 15) == 0, "zsg_conv_fwd: x_bf16 and w_bf16 must be 16-byte aligned");
    return p.cout <= 64 ? launch_conv_bf16<64>(p, st) : launch_conv_bf16<128>(p, st);
  }
  if (p.x_lo) {
    ZSG_REQUIRE(p.w_lo, "zsg_conv_fwd: x_lo needs pre-split weights (w_lo)");
    ZSG_REQUIRE(!p.in_scale && !p.in_relu, "zsg_conv_fwd: x_lo excludes the on-load affine / ReLU (apply them in zsg_split_act)");
    ZSG_REQUIRE((((uintptr_t)p.x_lo | (uintptr_t)p.w_lo) & 15) == 0, "zsg_conv_fwd: x_lo and w_lo must be 16-byte aligned");
    return p.cout <= 64 ? launch_conv_async<64>(p, st) : launch_conv_async<128>(p, st);
  }
  if (p.w_lo) {
    ZSG_REQUIRE(((uintptr_t)p.w_lo & 15) == 0, "zsg_conv_fwd: w_lo must be 16-byte aligned");
    return p.cout <= 64 ? launch_conv_pro<64>(p, st) : launch_conv_pro<128>(p, st);
  }
  return p.cout <= 64 ? launch_conv_generic<64>(p, st) : launch_conv_generic<128>(p, st);
}

extern "C" int zsg_conv_wgrad(const zsg_wgrad_params* pp, zsg_stream_t stream) {
  ZSG_REQUIRE(pp, "zsg_conv_wgrad: null params");
  zsg_wgrad_params p = *pp;
  ZSG_REQUIRE(p.dil >= 0 && p.dil <= 64, "zsg_conv_wgrad: dil=%d out of range", p.dil);
  if (p.dil == 0) p.dil = 1;
  ZSG_REQUIRE(((p.x && p.dy) || (p.x_bf16 && p.dy_bf16)) && p.dw && p.rows, "zsg_conv_wgrad: null pointer");
  ZSG_REQUIRE(p.impl != 1 || (p.x && p.dy), "zsg_conv_wgrad: the SIMT check kernel reads the fp32 operands");
  ZSG_REQUIRE(p.m > 0 && p.cout > 0 && p.cin > 0 && p.r > 0 && p.s > 0, "zsg_conv_wgrad: empty problem");
  ZSG_REQUIRE(p.impl == 1 || p.cin % 4 == 0, "zsg_conv_wgrad: cin=%d must be a multiple of 4", p.cin);
  ZSG_REQUIRE(!p.in_scale || p.in_shift, "zsg_conv_wgrad: in_scale without in_shift");
  cudaStream_t st = as_stream(stream);
  if (p.impl == 1) {
    const int64_t n = (int64_t)p.cout * p.r * p.s * p.cin;
    wgrad_simt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p);
    return check_launch("zsg_conv_wgrad(simt)");
  }
  if (!zsg_device_supported()) { set_error("zsg_conv_wgrad: tcgen05 path needs an sm_100 device"); return ZSG_EARCH; }
  if (p.x_bf16 || p.dy_bf16) {
    ZSG_REQUIRE(p.x_bf16 && p.dy_bf16, "zsg_conv_wgrad: x_bf16 and dy_bf16 go together");
    ZSG_REQUIRE(p.cin % 8 == 0, "zsg_conv_wgrad: the bf16 path needs cin=%d to be a multiple of 8", p.cin);
    ZSG_REQUIRE(!p.in_scale && !p.in_relu, "zsg_conv_wgrad: bf16 images exclude the on-load affine / ReLU");
    ZSG_REQUIRE(p.dy_pitch >= p.cout && p.dy_pitch % 8 == 0,
                "zsg_conv_wgrad: the bf16 path needs dy as a plain [m, dy_pitch] matrix, dy_pitch >= cout and a multiple of 8");
    ZSG_REQUIRE((((uintptr_t)p.x_bf16 | (uintptr_t)p.dy_bf16) & 15) == 0, "zsg_conv_wgrad: bf16 operands must be 16-byte aligned");
    if (use_wide_wgrad_bf16(p)) return launch_wgrad_bf16_wide(p, st);
    return p.cout <= 64 ? launch_wgrad_bf16<64>(p, st) : launch_wgrad_bf16<128>(p, st);
  }
  if (p.x_lo || p.dy_lo) {
    ZSG_REQUIRE(p.x_lo && p.dy_lo, "zsg_conv_wgrad: x_lo and dy_lo go together");
    ZSG_REQUIRE(!p.in_scale && !p.in_relu, "zsg_conv_wgrad: pre-split operands exclude the on-load affine / ReLU");
    ZSG_REQUIRE(p.cin % 4 == 0, "zsg_conv_wgrad: cin must be a multiple of 4");
    ZSG_REQUIRE((((uintptr_t)p.x | (uintptr_t)p.x_lo | (uintptr_t)p.dy | (uintptr_t)p.dy_lo) & 15) == 0,
                "zsg_conv_wgrad: operands must be 16-byte aligned");
    if (p.dy_pitch > 0) {
      ZSG_REQUIRE(p.dy_pitch >= p.cout && p.dy_pitch % 4 == 0, "zsg_conv_wgrad: dy_pitch must be >= cout and a multiple of 4");
      return p.cout <= 64 ? launch_wgrad<64, 2>(p, st) : launch_wgrad<128, 2>(p, st);
    }
    return p.cout <= 64 ? launch_wgrad<64, 1>(p, st) : launch_wgrad<128, 1>(p, st);
  }
  return p.cout <= 64 ? launch_wgrad<64, 0>(p, st) : launch_wgrad<128, 0>(p, st);
}
