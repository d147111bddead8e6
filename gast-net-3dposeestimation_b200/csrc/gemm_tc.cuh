// tcgen05 (5th-gen tensor core) GEMM core of the lifting path: error-compensated TF32.
//
//   D[128 rows, 128 cols] (fp32, TMEM) += A_hi.B_hi  (kind::tf32)  +  A_lo.B_hi + A_hi.B_lo  (kind::f16, bf16)
//
// where x_hi = tf32(x), x_lo = x - x_hi.  The two correction products are 2^-11 of the result, so
// they do not need TF32 operands: with bf16 operands (2^-9 relative operand error) they are good
// to 2^-20 of the result, the same order as the dropped A_lo.B_lo term (2^-22) -- the result still
// matches an fp32 FFMA GEMM to ~1e-6 (the parity bar is 1e-4 abs; a single TF32 pass would not,
// SURVEY.md §7).  bf16 MMAs run at twice the TF32 rate, and both corrections together are ONE
// K=64 bf16 product  [A_lo | A_hi] . [B_hi | B_lo]^T  per 32-wide K chunk, so a chunk costs
// 4 (tf32) + 4 (bf16) tensor-pipe slots instead of the 12 of 3xTF32 (round 1).  Same tiles,
// A-gather and fused epilogues as gemm_ffma.cuh, so both cores are interchangeable per launch.
//
// Persistent, warp-specialised CTA (one per SM), launched as 2-CTA clusters that work as a CTA PAIR (tcgen05 cta_group::2,
// template parameter CG = 2): adjacent M tiles, same N tile; the rank-0 CTA issues M = 256 MMAs for both, each CTA loads
// only its half of the B tile (see the comment at the kernel).  512 threads, registers re-balanced with setmaxnreg
// (136 / 168 / 40 / 168):
//   warps 0-3   A converters : raw A rows (TMA or cp.async ring in smem) -> hi (tf32) | packed bf16 [lo | hi] ->
//                              tcgen05.st into the A ring in TENSOR memory (TS-form MMA)
//   warp  8     B producer   : TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) of this CTA's 64 rows of the pre-split weights,
//                              bytes counted on the rank-0 CTA's barrier (CG = 1: whole tile, multicast across the cluster)
//   warp  9     MMA issuer   : one elected thread (of the rank-0 CTA) issues tcgen05.mma kind::tf32 / kind::f16, accumulators in TMEM
//   warp  10    A producer   : TMA (cp.async.bulk.tensor.3d) of the raw A tile when the frame map is affine
//   warps 4-7   accumulate + epilogue of accumulator columns 0-63
//   warps 12-15 accumulate + epilogue of accumulator columns 64-127
//               tcgen05.ld of every flush group, added into fp32 REGISTERS (round-to-nearest), then smem
//               staging for the joint mixing -> bias/BN shift/ReLU/residual | SemCH neighbour mix | attention mix
//
// Two-level accumulation.  Measured on B200 (tools/tc_probe.py, profiles/r01_tc_numerics.md): the
// tensor core aligns and TRUNCATES its addends, so a long accumulation chain in TMEM is biased
// towards zero by ~1 ulp per MMA (K=1536: -1e-5 relative; whole model: 6e-5 abs, MPJPE biased).
// Therefore a tile is accumulated in TMEM over at most TC_FLUSH chunks at a time (into a ring of 2
// accumulator buffers) and the group sums are added in registers with RN.  The correction products
// A_lo.B_hi + A_hi.B_lo (2^-11 of the result) go into the SAME accumulator as A_hi.B_hi (round 1 kept them in
// a third TMEM buffer over the whole K): that frees 128 tensor-memory columns, which buy a 4-stage instead of a
// 2-stage A ring -- the hand-over latency converter -> MMA -> converter of a 2-stage ring, not the tensor pipe,
// bounded the chunk rate at ~800 cycles (profiles/r02_tc_attribution_bf16corr.md) -- and it removes the
// per-tile wait for the correction buffer to drain.  Every MMA truncates the accumulator once whatever the size
// of its addend, so a group now sees 8 truncations per chunk instead of 4; the pre-compensation through W_lo
// (tc_split_kernel) counts them accordingly.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>
#include <vector>
#include "gast_common.cuh"

namespace gast {

struct TcWeights {
  float* hi = nullptr;   // [N][K] tf32-rounded weights
  float* lo = nullptr;   // the same bytes as [N][2K] bf16: per 32-wide K chunk [hi | lo] (correction operand)
  CUtensorMap map_hi, map_lo;
  int N = 0, K = 0;
  int prec = 0;          // 0: `lo` holds the bf16 [hi | lo] rows; 1 (3xTF32, training): `lo` holds fp32 W - tf32(W), [N][K]
  bool ready = false;
};

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;
#ifndef GAST_TC_BSTAGES
#define GAST_TC_BSTAGES 3
#endif
#ifndef GAST_TC_RSTAGES
#define GAST_TC_RSTAGES 4
#endif
constexpr int TC_BSTAGES = GAST_TC_BSTAGES;   // B (weights) ring in shared memory, filled by TMA
constexpr int TC_RSTAGES = GAST_TC_RSTAGES;   // raw A ring in shared memory, filled by cp.async (no register staging, no MSHR cap)
// ... and when it is filled by TMA (16 KB per slot): the bytes in flight per SM bound the chunk rate -- 4 slots = 64 KB
// against ~3500 cycles of load latency under load = 18 B/clk = one 16 KB chunk per ~880 cycles, the measured period of
// the CTA-pair kernel -- so the shared memory the CTA pair frees (half a B tile per CTA) goes into this ring
#ifndef GAST_TC_RSTAGES_TMA
#define GAST_TC_RSTAGES_TMA 4
#endif
constexpr int TC_RSTAGES_TMA = GAST_TC_RSTAGES_TMA;
constexpr int TC_RAW_BYTES = (TC_RSTAGES * 128 * 36 * 4 > TC_RSTAGES_TMA * 16384) ? TC_RSTAGES * 128 * 36 * 4 : TC_RSTAGES_TMA * 16384;
static_assert(TC_RSTAGES <= 8 && TC_RSTAGES_TMA <= 8, "8 raw-slot barriers");
#ifndef GAST_TC_ASTAGES
#define GAST_TC_ASTAGES 4
#endif
constexpr int TC_ASTAGES = GAST_TC_ASTAGES;   // A (activations) ring in TENSOR MEMORY, filled by tcgen05.st (<= 4: 64 columns each)
static_assert(TC_ASTAGES >= 2 && TC_ASTAGES <= 4, "the A ring has 256 tensor-memory columns");
constexpr int TC_FLUSH = 4;     // K chunks accumulated in TMEM before the sum is flushed to registers
#ifndef GAST_TC_CLUSTER
#define GAST_TC_CLUSTER 2
#endif
constexpr int TC_CLUSTER = GAST_TC_CLUSTER;   // CTAs per cluster: same N tile, adjacent M tiles, B multicast by TMA
constexpr int TC_THREADS = 512;   // 4 warpgroups: A converters | epilogue (cols 0-63) | TMA, MMA, 1 idle | epilogue (cols 64-127)
// registers per thread after setmaxnreg: A converters / epilogue groups / TMA+MMA warps (sum x 128 threads <= 64K)
#ifndef GAST_TC_REG_A
#define GAST_TC_REG_A 136
#define GAST_TC_REG_E 168
#define GAST_TC_REG_M 40
#endif
constexpr int TC_REG_A = GAST_TC_REG_A, TC_REG_E = GAST_TC_REG_E, TC_REG_M = GAST_TC_REG_M;
static_assert(TC_REG_A + 2 * TC_REG_E + TC_REG_M <= 512, "register file over-subscribed");

// A converters (TMA-fed path): 1 = software-pipelined -- the raw rows of chunk c+1 are requested from shared memory
// before the hand-over of chunk c (wait for the tensor-memory stage, tcgen05.st, arrive), so the shared-memory
// latency and the wait for the TMA data overlap the hand-over latencies; 0 = one chunk at a time (round 1)
#ifndef GAST_TC_CONV_PIPE
#define GAST_TC_CONV_PIPE 1
#endif
constexpr int TC_CONV_PIPE = GAST_TC_CONV_PIPE;
#ifndef GAST_TC_CONV_ST_EARLY
#define GAST_TC_CONV_ST_EARLY 1
#endif
#ifndef GAST_TC_TRUNC_SPLIT
#define GAST_TC_TRUNC_SPLIT 0
#endif

// MMA issue: 1 = two issuing warps (9 and 11) take alternate flush groups (each owns one accumulator buffer and its
// own set of operand "full" barriers), 0 = warp 9 issues everything
#ifndef GAST_TC_DUAL_ISSUE
#define GAST_TC_DUAL_ISSUE 0
#endif
constexpr int TC_DUAL_ISSUE = GAST_TC_DUAL_ISSUE;
constexpr int TC_NISSUE = TC_DUAL_ISSUE ? 2 : 1;
#ifndef GAST_TC_CG_DEFAULT
#define GAST_TC_CG_DEFAULT 2
#endif
#ifndef GAST_TC_REMOTE_RELEASE_CLUSTER
#define GAST_TC_REMOTE_RELEASE_CLUSTER 0
#endif
#ifndef GAST_TC_CG2_DEEP
#define GAST_TC_CG2_DEEP 1
#endif
constexpr int TC_CG_DEFAULT = GAST_TC_CG_DEFAULT;   // 2 = CTA-pair MMAs (tcgen05 cta_group::2) unless GAST_TC_CG says otherwise


constexpr int TC_EN = 64;         // accumulator columns owned by one epilogue warpgroup
constexpr int TC_STAGE_BYTES = 2 * 16384;            // B_hi, B_lo : 128 rows x 128 B each
constexpr int TC_SLD = 68;                            // SemCH coefficient slab row stride (floats): 64 channels + 4, conflict-free 16B rows
constexpr int TC_MAX_NNZ = 64;    // SemCH coefficient slab rows
constexpr int TC_MAXDEG = 6;      // SemCH: neighbours per joint kept in a packed register (17j: <= 5)
constexpr int TC_JMAX = 20;
constexpr int TC_XLD = 36;                            // raw A / 32-column staging row stride (floats): conflict-free row-per-thread access
// Epilogue scratch, laid out per epilogue kind:
//   PLAIN : 8 warp-private 32 x TC_XLD patches (two groups x 4 warps)
//   SEMCH : 2 x [128][TC_XLD] H1 staging (one per group) | coefficient slab [TC_MAX_NNZ][TC_SLD] (64 channels)
//   GLOBAL: 2 x [128][TC_XLD] g staging (one per group) | a/b tile of group 1 (group 0 uses TC_OFF_AB)
constexpr int TC_EPI_BYTES = 2 * 128 * TC_XLD * 4 + TC_MAX_NNZ * TC_SLD * 4;
constexpr int TC_OFF_STAGING = TC_BSTAGES * TC_STAGE_BYTES;
constexpr int TC_OFF_AB = TC_OFF_STAGING + TC_EPI_BYTES;
constexpr int TC_OFF_XPOSE = (TC_OFF_AB + 128 * 8 * 4 + 1023) / 1024 * 1024;  // raw A ring (1024-aligned: TMA SWIZZLE_128B)
static_assert(TC_XLD == 36, "TC_RAW_BYTES assumes the 36-float row stride of the cp.async layout");
constexpr int TC_OFF_BAR = TC_OFF_XPOSE + TC_RAW_BYTES;
static_assert(TC_OFF_STAGING % 1024 == 0 && TC_OFF_XPOSE % 1024 == 0, "swizzled regions must be 1024-byte aligned");
constexpr int TC_SMEM_BYTES = TC_OFF_BAR + 512 + 1024;
static_assert(TC_SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory per CTA");   // 26 mbarriers + tmem ptr, + alignment slack

// ----------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// non-blocking probe (one try_wait), so that several barriers can be polled with overlapping latency
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}

// multicast form: the box lands at the same smem offset in every CTA of `mask`, and each of those
// CTAs' mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- cta_group::2 (CTA pair) forms.  All tcgen05 instructions of one kernel use the same cta_group, so these are
// selected by the kernel's CG template parameter.
// shared::cluster address of `addr` (a shared::cta address of this CTA's window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier of any CTA of the cluster (shared::cluster address from mapa_u32)
// (default semantics = release at CTA scope, like the local arrives: the hand-over it signals is a tensor-memory
//  write / read ordered by tcgen05.fence, not generic-proxy memory.  The .release.cluster form measured 2x slower
//  kernels: every converter warp paid a cluster-scope fence per chunk, profiles/r02_tc_attribution.md #16.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
#if GAST_TC_REMOTE_RELEASE_CLUSTER
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
// TMA load whose bytes are counted on an mbarrier of ANOTHER CTA of the pair (`bar` is a shared::cluster address):
// both CTAs load their half of the B tile into their own shared memory and signal the MMA-issuing CTA's barrier
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs, 128 rows each] (+)= A[tmem of each CTA] . B[64 rows from each CTA's smem]^T : M = 256, N = 128
__device__ __forceinline__ void umma_tf32_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// completion of this thread's earlier MMAs -> the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::tf32, single CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// same with the A operand read from tensor memory (TS form): halves the shared-memory
// traffic of the MMA, which at M=N=128 otherwise saturates the 128 B/clk smem port by itself
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// D[tmem] (+)= A[tmem, bf16 pairs packed per 32-bit column] . B[smem desc, bf16]^T, kind::f16, K = 16 per instruction
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// two fp32 -> one 32-bit register of two bf16 (round to nearest even): `lo_k` in bits 0-15 (the lower k index)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_k, float hi_k) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_k), "f"(lo_k));
  return d;
}
// two fp32 -> packed fp16 pair (lower k in the low half), round to nearest, SATURATING at +-65504 (no inf operand), and
// the pair back as two fp32 (exact)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo_k, float hi_k) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_k), "f"(lo_k));
  return d;
}
__device__ __forceinline__ void unpack_f16x2(uint32_t d, float& lo_k, float& hi_k) {
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo_k), "=f"(hi_k) : "r"(d));
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// arrive on the same barrier offset in every CTA of `mask` (frees a multicast-filled smem stage)
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (quadrant base + i).
// Asynchronous: the registers are valid only after tmem_wait_ld(), which names them as in/out
// operands so that the compiler cannot move their uses above the wait.
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}
// registers -> tensor memory: thread i of the warp writes lane (quadrant base + i), 32 columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  tmem_ld32_async(taddr, r);
  tmem_wait_ld(r);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// finite inputs only: add half an ulp of the 10-bit mantissa and clear the 13 low bits
__device__ __forceinline__ float tf32_rn_fast(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 writes zeros
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}

// SemCH neighbour mix with packed fp32x2 FMAs (fma.rn.f32x2): o[0..3] += cf * hv element by element, each half rounded
// like the scalar FFMA (bit-identical), half the FMA issue slots of the epilogue warps
#ifndef GAST_TC_SEMCH_FFMA2
#define GAST_TC_SEMCH_FFMA2 1
#endif
__device__ __forceinline__ void semch_fma2(const float4 cf, const float4 hv, float* o) {
  const float2 lo = __ffma2_rn(make_float2(cf.x, cf.y), make_float2(hv.x, hv.y), make_float2(o[0], o[1]));
  const float2 hi = __ffma2_rn(make_float2(cf.z, cf.w), make_float2(hv.z, hv.w), make_float2(o[2], o[3]));
  o[0] = lo.x; o[1] = lo.y; o[2] = hi.x; o[3] = hi.y;
}

// named barrier of one epilogue warpgroup (128 threads): id 1 = columns 0-63, id 2 = columns 64-127
__device__ __forceinline__ void epi_bar_sync(uint32_t id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row atoms of 1024 B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);       // start address, 16 B units
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

// Tensor-memory map (512 columns x 128 lanes x 32 bit):
//   [0,256)   accumulator ring (2 x 128): A_hi.B_hi + A_lo.B_hi + A_hi.B_lo of one flush group (<= TC_FLUSH chunks)
//   [256,512) A operand ring (4 stages x {hi: 32 tf32 columns | 16 columns of bf16 pairs of lo | 16 of hi}):
//             row m in lane m, k along the columns
constexpr uint32_t TC_NMAIN = 2;
constexpr uint32_t TC_A_COL = 256;

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);
// cta_group::2: M = 256 over the CTA pair (128 rows in each CTA's tensor memory)
constexpr uint32_t TC_IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                               ((uint32_t)((2 * TC_BM) >> 4) << 24);
constexpr uint32_t TC_IDESC2_BF = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                  ((uint32_t)((2 * TC_BM) >> 4) << 24);
// PREC = 2 (CTA pair): every operand fp16 (kind::f16 format code 0).  (Mixed formats -- a bf16 remainder against an fp16
// hi operand -- raise "illegal instruction": A and B of one kind::f16 MMA must have the same type, GPU session Z4.)
constexpr uint32_t TC_IDESC2_HH = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
// weights are split as 2^TC_F16_WSHIFT * W (their remainder then sits in fp16's normal range: |W| ~ 0.03 has a remainder of
// ~1e-5, below fp16's smallest normal 6.1e-5); the epilogue multiplies the group sums by 2^-TC_F16_WSHIFT (exact)
constexpr int TC_F16_WSHIFT = 8;
constexpr float TC_F16_WSCALE = 256.f, TC_F16_WUNSCALE = 1.f / 256.f;
// the same with A=B=bf16 (kind::f16 format code 1), for the correction products
constexpr uint32_t TC_IDESC_BF = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                 ((uint32_t)(TC_BM >> 4) << 24);

// issuer of the flush group a chunk belongs to: groups are numbered over all tiles of a CTA
struct TcGroupOf {
  int ct = 0; uint32_t grp = 0;
  __device__ __forceinline__ int role() const { return (int)(grp % TC_NISSUE); }
  __device__ __forceinline__ void next(int nchunks_) {
    ++ct;
    if (ct == nchunks_) { ct = 0; ++grp; }
    else if ((ct % TC_FLUSH) == 0) ++grp;
  }
};

// ----------------------------------------------------------------------------------------
// kernel
// ----------------------------------------------------------------------------------------
// DBG != 0 builds timing-experiment variants (tools/tc_probe.py --perf): 1 = no TMEM->register
// flush, 2 = no global A loads, 3 = main MMA only, 4 = no tcgen05.st of A, 5 = 1+2+4,
// 6 = full kernel + clock64() attribution of every role's waits (written to p.dbg[blockIdx.x*32 + i]),
// 7 = correction (bf16) MMAs only, 8 = 5 + no B loads (MMAs on stale shared memory), 9 = 8 + A converters reduced to
// their barrier hand-shakes (no shared-memory reads, no split): the floor of the MMA issue loop and the barrier protocol.
// PREC: 0 = one TF32 product + ONE bf16 correction product per chunk (inference: 2e-6 rms per GEMM); 1 = 3xTF32 --
// both correction products A_lo.B_hi and A_hi.B_lo as kind::tf32 MMAs on fp32 `lo` operands (12 MMAs per chunk, 3e-7 rms:
// as good as an fp32 FFMA GEMM).  The training path uses PREC = 1: its parity bar is the reference's own fp32 gradient
// noise (tests/test_gpu_train.py), which the bf16 corrections would exceed.  Same rings, barriers and tensor-memory map:
// the 32 "pair" columns of an A stage hold 32 fp32 lo values instead of 16 + 16 bf16 pairs, the second 16 KB of a B
// stage the fp32 lo tile instead of the bf16 [hi | lo] rows.
// CG: 1 = every CTA issues its own M = 128 MMAs and the two CTAs of a cluster multicast the B tile to each other;
// 2 = CTA PAIR (tcgen05 cta_group::2): the rank-0 CTA issues M = 256 MMAs for both -- each CTA converts its own 128 rows
// of A into its own tensor memory and accumulates its own 128 rows, but loads only ITS HALF of the B tile (64 of the
// 128 weight rows) into its shared memory; the tensor cores of the pair read both halves.  Per CTA and chunk that is
// 16 KB instead of 32 KB of TMA writes and 16 KB instead of 32 KB of operand reads: at M = N = 128 the shared-memory
// port (TMA writes of A and B + MMA reads of B + the converters' reads of A = 96 KB per ~1000-cycle chunk at
// 128 B/clk), not the issuing thread or the converters, is what no single-CTA variant could get past
// (profiles/r02_tc_attribution.md, experiments 10-13).  Protocol differences: the operand "full" barriers and the
// accumulator "empty" barriers that the issuer waits on live in the rank-0 CTA and collect arrivals from both CTAs
// (remote mbarrier arrives; the peer's B TMA counts its bytes on the leader's barrier), every commit is multicast to
// both CTAs.
template <int EPI, int DBG = 0, int PREC = 0, int CG = 1>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmP p, const __grid_constant__ CUtensorMap map_hi,
               const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap amap0,
               const __grid_constant__ CUtensorMap amap1, const __grid_constant__ CUtensorMap amap2,
               int n_tiles_n, int total_tiles) {
  // total_tiles counts (M-tile group, N tile) work items: a group is TC_CLUSTER adjacent M tiles
  // No integer round-trip on this pointer: the compiler must keep knowing it is SHARED memory,
  // otherwise every staging / transposition access becomes a generic LD/ST (ncu: 45 % of the
  // A-producer samples were long-scoreboard stalls on generic loads from the patch).
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  unsigned char* smem = tc_smem_raw;
  uint32_t sbase_ = smem_u32(smem);
  // opaque to the optimiser: otherwise it re-derives the shared-window base from SR_CgaCtaId (a slow special-register
  // read) in front of every mbarrier operation of the latency-bound MMA issue loop
  asm volatile("" : "+r"(sbase_));
  const uint32_t sbase = sbase_;
  if ((sbase & 1023u) != 0) __trap();   // SWIZZLE_128B atoms need 1024-byte alignment
  float* staging = reinterpret_cast<float*>(smem + TC_OFF_STAGING);
  float* ab_s = reinterpret_cast<float*>(smem + TC_OFF_AB);
  const uint32_t bar0 = sbase + TC_OFF_BAR;
  // mbarriers (8 B each):  b_full[8 stages][2 issuers] @0  b_empty[8] @128  a_full[4][2] @192  a_empty[4] @256
  //                        main_full[2] @288  main_empty[2] @304 ; tmem ptr @320 ; raw_full[8] @328  raw_empty[8] @392
  // The operand "full" barriers exist once per issuing warp: the issuer of a chunk's flush group is the only
  // waiter of that barrier instance, so it sees EVERY phase of it (an mbarrier parity wait cannot tell a phase
  // from the one two later, which an issuer that skips the other issuer's chunks would otherwise run into).
  // CTA pair: a B stage holds this CTA's 64 weight rows only (8 KB hi + 8 KB pk), so the same 96 KB give a ring of
  // twice the depth -- the operand round trip now includes a remote complete_tx and a multicast commit
  constexpr int BST = (CG == 2 && GAST_TC_CG2_DEEP) ? 2 * TC_BSTAGES : TC_BSTAGES;        // B stages
  constexpr int BSB = (CG == 2 && GAST_TC_CG2_DEEP) ? TC_STAGE_BYTES / 2 : TC_STAGE_BYTES; // bytes per B stage
  constexpr int BPK = (CG == 2 && GAST_TC_CG2_DEEP) ? 8192 : 16384;                        // offset of the pk / lo half
  constexpr uint32_t BB_FULL = 0, BB_EMPTY = 128, BA_FULL = 192, BA_EMPTY = 256, BM_FULL = 288, BM_EMPTY = 304,
                     B_TMEMPTR = 320, BR_FULL = 328, BR_EMPTY = 392;
  static_assert(BST <= 8, "barrier area: 8 B stages");
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + TC_OFF_BAR + B_TMEMPTR);
  constexpr uint32_t NMAIN = TC_NMAIN;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int J = p.J;
  // Cluster of TC_CLUSTER CTAs: work item q = (group of adjacent M tiles, N tile); this CTA takes
  // M tile TC_CLUSTER*(q / n_tiles_n) + rank.  An M tile beyond the last one is a dummy (all rows
  // invalid) that still takes part in the B multicast protocol, so all CTAs of a cluster run the
  // same number of chunks.
  const uint32_t crank = (TC_CLUSTER > 1) ? cluster_ctarank() : 0u;
  const int cid = blockIdx.x / TC_CLUSTER, ncl = gridDim.x / TC_CLUSTER;
  auto tile_f0 = [&](int q) { return (TC_CLUSTER * (q / n_tiles_n) + (int)crank) * p.fpt; };

  static_assert(CG == 1 || (TC_CLUSTER == 2 && !TC_DUAL_ISSUE && (DBG == 0 || DBG == 4 || DBG == 5 || DBG == 8 || DBG == 9)), "the CTA-pair form needs 2-CTA clusters");
  if (tid == 0) {
    for (int s = 0; s < BST; ++s) {
      for (int q = 0; q < 2; ++q) mbar_init(bar0 + BB_FULL + 8 * (2 * s + q), 1);     // expect_tx arrive + TMA bytes
      mbar_init(bar0 + BB_EMPTY + 8 * s, CG == 2 ? 1 : TC_CLUSTER);   // tcgen05.commit of every issuing CTA of the cluster
    }
    for (int s = 0; s < TC_ASTAGES; ++s) {
      for (int q = 0; q < 2; ++q) mbar_init(bar0 + BA_FULL + 8 * (2 * s + q), 4 * CG);     // 4 A-producer warps (of each CTA of the pair)
      mbar_init(bar0 + BA_EMPTY + 8 * s, 1);    // tcgen05.commit
    }
    for (int b = 0; b < (int)NMAIN; ++b) {
      mbar_init(bar0 + BM_FULL + 8 * b, 1);     // tcgen05.commit
      mbar_init(bar0 + BM_EMPTY + 8 * b, 8 * CG);    // 8 accumulate/epilogue warps (two column groups) (of each CTA of the pair)
    }
    for (int r = 0; r < 8; ++r) {
      mbar_init(bar0 + BR_FULL + 8 * r, 1);     // raw A slot: expect_tx arrive + TMA bytes
      mbar_init(bar0 + BR_EMPTY + 8 * r, 4);    // 4 A-producer warps have read it
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) { if (CG == 2) tmem_alloc2(sbase + TC_OFF_BAR + B_TMEMPTR, 512); else tmem_alloc(sbase + TC_OFF_BAR + B_TMEMPTR, 512); }
  tc_fence_before();
  __syncthreads();
  if (TC_CLUSTER > 1) cluster_sync_all();   // peer barriers initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  // CTA pair: the barriers the MMA issuer waits on are the rank-0 CTA's (shared::cluster addresses for remote arrives)
  const uint32_t lead_bar0 = (CG == 2) ? mapa_u32(bar0, 0) : bar0;

  int nchunks = 0;
  for (int s = 0; s < p.nseg; ++s) nchunks += p.seg[s].K / TC_BK;

  if (warp < 4) {
    if (TC_REG_A >= 128) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REG_A)); else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REG_A));
    // ================================================================= A producers
    // Warp w owns tile rows 32w..32w+31 (== its TMEM lane quadrant).  Global loads are COALESCED:
    // instruction i of a chunk covers rows 32w + 4i + (lane>>3), 8 lanes x 16 B = one 128-byte
    // line per row along the channel axis (a row-per-thread mapping costs 32 L1 tag lookups per
    // instruction and made the A path the bottleneck, profiles/r01_tc_attribution.md).  At hand-over
    // the warp transposes the chunk through a private 32x36-float smem patch so that thread t
    // holds row t, splits hi/lo and writes both into the A ring in tensor memory (tcgen05.st).
    const int c16 = lane & 7;              // 16-byte chunk of the row
    const int rsub = lane >> 3;            // row within the 4-row group of one instruction
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float* xpose = reinterpret_cast<float*>(smem + TC_OFF_XPOSE) + warp * (32 * TC_XLD);
    long long roff[8];
    // (frame-in-tile, joint) of this thread's 8 rows, packed one byte each: fr*32 + jj
    unsigned long long frjj = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = warp * 32 + 4 * i + rsub;
      const int fr_ = row / J, jj_ = row - fr_ * J;
      frjj |= (unsigned long long)((fr_ << 5) | jj_) << (8 * i);
    }
    int c_tile = -1, c_seg = -1;
    // row offsets of the current (tile, segment): ONE division per call, the frames of a tile
    // are consecutive (the per-row divisions of the first version cost ~1000 cycles per call)
    auto ensure = [&](int tile, int sg) {
      if (tile == c_tile && sg == c_seg) return;
      c_tile = tile; c_seg = sg;
      const int f0 = tile_f0(tile);
      const int nf = min(p.fpt, p.F - f0);          // <= 0 for a dummy tile
      const RowMap mp = p.seg[sg].map;
      const long long ld = p.seg[sg].ld;
      const int b0 = f0 / mp.T_out, t0 = f0 - b0 * mp.T_out;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int code = (int)((frjj >> (8 * i)) & 0xff);
        const int fr_ = code >> 5, jj_ = code & 31;
        if (fr_ < nf) {
          int b = b0, t = t0 + fr_;
          while (t >= mp.T_out) { t -= mp.T_out; ++b; }
          long long fin = (long long)b * mp.T_in + (long long)t * mp.t_mul + mp.t_off;
          roff[i] = (fin * J + jj_) * ld;
        } else {
          roff[i] = -1;
        }
      }
    };
    long long my_chunks = 0;
    for (int q = cid; q < total_tiles; q += ncl) my_chunks += nchunks;
    int stage = 0, slot = 0;
    uint32_t phase = 0, rphase = 0;
    long long tA_wait = 0, tA_st = 0, tA_tot = clock64(), tA_n = 0, tA_ld = 0, tA_x = 0, tA_is = 0;
    // Raw A reaches shared memory either by TMA (affine frame maps: warp 10 issues one
    // cp.async.bulk.tensor.3d per chunk, 128-byte-swizzled rows) or by a cp.async gather issued
    // here (general frame maps: dilated stages on long sequences, dense ablation).  Either way it
    // never passes through registers on its way in: the LSU path caps the bytes in flight per SM
    // (profiles/r01_tc_attribution.md), TMA does not.
    const uint32_t raw0 = sbase + TC_OFF_XPOSE + (uint32_t)(warp * 32 * TC_XLD * 4);
    auto issue = [&](int sg, int k0, int slot_) {
      const ASeg& sgm = p.seg[sg];
      const int tap = k0 / sgm.Kc;
      const float* src = sgm.base + tap * sgm.tap_stride + (k0 - tap * sgm.Kc) + c16 * 4;
      const uint32_t dst = raw0 + (uint32_t)(slot_ * 128 * TC_XLD * 4) + (uint32_t)(rsub * TC_XLD * 4 + c16 * 16);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = roff[i] >= 0 && DBG != 2 && DBG != 5 && DBG != 8 && DBG != 9;
        cp_async16(dst + (uint32_t)(4 * i * TC_XLD * 4), ok ? (const void*)(src + roff[i]) : (const void*)sgm.base,
                   ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    int tile = cid, sg = 0, k0 = 0;
    bool have = tile < total_tiles;
    auto advance = [&]() {
      k0 += TC_BK;
      if (k0 >= p.seg[sg].K) { k0 = 0; ++sg; if (sg >= p.nseg) { sg = 0; tile += ncl; } }
      have = tile < total_tiles;
    };
    const bool atma = p.a_tma != 0;
    if (!atma) {
      for (int r = 0; r < TC_RSTAGES; ++r) {   // prologue: fill the ring
        if (have) { ensure(tile, sg); issue(sg, k0, r); advance(); }
        else cp_async_commit();               // keep group accounting uniform
      }
    }
    const int my_row = warp * 32 + lane;
    const bool row_in_box = my_row < p.fpt * J;
    // raw rows of one chunk -> 8 float4 of this thread's row (swizzle-aware, conflict-free)
    // (rows of the tile that the TMA box does not cover -- 119..127 for J = 17 -- are never written by the TMA:
    //  their owners zero them once in every raw slot instead of masking 32 values per chunk)
    auto load_raw = [&](int slot_, float4 (&xr)[8]) {
      const float* rowp = reinterpret_cast<const float*>(smem + TC_OFF_XPOSE + slot_ * 16384 + my_row * 128);
      const int sw = my_row & 7;              // TMA SWIZZLE_128B: 16-byte chunk index ^= row % 8
#pragma unroll
      for (int i = 0; i < 8; ++i) xr[i] = *reinterpret_cast<const float4*>(rowp + ((i ^ sw) << 2));
    };
    auto split_row = [&](const float4 (&xr)[8], uint32_t* hi, uint32_t* lo) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 x = xr[i];
#if GAST_TC_TRUNC_SPLIT
        if (PREC == 0) {
          // split by TRUNCATION: the raw value is the tf32 operand (the tensor core reads the upper 19 bits), the
          // correction is x - trunc(x): one integer op per element less than rounding.  Alone it measured no faster
          // (the single MMA issuer paces the chunk rate then, experiment 11); with two issuers the converters are
          // the limiter (experiment 10), so the two are meant to be used together.  Costs a bit of accuracy: the
          // remainder is one-signed and up to 2^-10 |x| (per-GEMM error 3.2e-6 instead of 2.1e-6 rms).
          const float t0 = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u), t1 = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          const float t2 = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u), t3 = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          hi[4 * i + 0] = __float_as_uint(x.x); hi[4 * i + 1] = __float_as_uint(x.y);
          hi[4 * i + 2] = __float_as_uint(x.z); hi[4 * i + 3] = __float_as_uint(x.w);
          lo[2 * i + 0] = pack_bf16x2(x.x - t0, x.y - t1);
          lo[2 * i + 1] = pack_bf16x2(x.z - t2, x.w - t3);
          lo[16 + 2 * i + 0] = pack_bf16x2(x.x, x.y);
          lo[16 + 2 * i + 1] = pack_bf16x2(x.z, x.w);
          continue;
        }
#endif
        // (splitting by TRUNCATION -- the raw value as the tf32 operand, x - trunc(x) as the correction -- saves an
        //  integer op per element but measured no faster and doubles the error of the bf16 correction:
        //  profiles/r02_tc_attribution.md, experiment 11)
        if (PREC == 2) {
          // fp16 hi (11 significant bits, like tf32) | fp16 remainder: 32 pair columns [h16(k = 0..31) | lo(k = 0..31)]
          const uint32_t p01 = pack_f16x2_sat(x.x, x.y), p23 = pack_f16x2_sat(x.z, x.w);
          float g0, g1, g2, g3;
          unpack_f16x2(p01, g0, g1); unpack_f16x2(p23, g2, g3);
          lo[2 * i + 0] = p01; lo[2 * i + 1] = p23;
          lo[16 + 2 * i + 0] = pack_f16x2_sat(x.x - g0, x.y - g1);
          lo[16 + 2 * i + 1] = pack_f16x2_sat(x.z - g2, x.w - g3);
          continue;
        }
        const float h0 = tf32_rn_fast(x.x), h1 = tf32_rn_fast(x.y), h2 = tf32_rn_fast(x.z), h3 = tf32_rn_fast(x.w);
        hi[4 * i + 0] = __float_as_uint(h0); hi[4 * i + 1] = __float_as_uint(h1);
        hi[4 * i + 2] = __float_as_uint(h2); hi[4 * i + 3] = __float_as_uint(h3);
        if (PREC == 1) {                       // 3xTF32: the remainder as a tf32 operand of its own (rounded, not truncated)
          lo[4 * i + 0] = __float_as_uint(tf32_rn_fast(x.x - h0)); lo[4 * i + 1] = __float_as_uint(tf32_rn_fast(x.y - h1));
          lo[4 * i + 2] = __float_as_uint(tf32_rn_fast(x.z - h2)); lo[4 * i + 3] = __float_as_uint(tf32_rn_fast(x.w - h3));
        } else {
          lo[2 * i + 0] = pack_bf16x2(x.x - h0, x.y - h1);
          lo[2 * i + 1] = pack_bf16x2(x.z - h2, x.w - h3);
          lo[16 + 2 * i + 0] = pack_bf16x2(h0, h1);
          lo[16 + 2 * i + 1] = pack_bf16x2(h2, h3);
        }
      }
    };
    TcGroupOf gof;                            // which issuer waits for the chunk being converted
    if (atma && TC_CONV_PIPE && DBG == 0) {
      float4 xr[8];
      if (!row_in_box) {
#pragma unroll
        for (int sl = 0; sl < TC_RSTAGES_TMA; ++sl) {
          float4* rz = reinterpret_cast<float4*>(smem + TC_OFF_XPOSE + sl * 16384 + my_row * 128);
#pragma unroll
          for (int i = 0; i < 8; ++i) rz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (my_chunks > 0) {
        mbar_wait(bar0 + BR_FULL + 8 * slot, rphase);
        load_raw(slot, xr);
      }
      for (long long c = 0; c < my_chunks; ++c) {
        uint32_t hi[32], lo[32];
        split_row(xr, hi, lo);                 // consumes every raw value: the shared-memory reads are complete
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + BR_EMPTY + 8 * slot);   // raw slot free for the TMA producer
        if (++slot == TC_RSTAGES_TMA) { slot = 0; rphase ^= 1; }
        const bool more = c + 1 < my_chunks;
        bool got_next = false;
#if !GAST_TC_CONV_ST_EARLY
        if (more) {                            // next chunk's rows on their way while this one is handed over
          mbar_wait(bar0 + BR_FULL + 8 * slot, rphase);
          load_raw(slot, xr);
          got_next = true;
        }
#endif
        mbar_wait(bar0 + BA_EMPTY + 8 * stage, phase ^ 1);
        tc_fence_after();
        const uint32_t ta = tmem_base + lane_off + TC_A_COL + stage * 64;
        if (PREC != 2) tmem_st32(ta, hi);      // (PREC 2: no fp32 operand -- the stage is its 32 pair columns)
        tmem_st32(ta + 32, lo);
#if GAST_TC_CONV_ST_EARLY
        // the tensor-memory stores are in flight: request the next chunk's rows in their shadow (only if they have
        // landed -- a late TMA must not hold back the hand-over of a chunk that is already converted)
        if (more && __all_sync(0xffffffffu, mbar_try(bar0 + BR_FULL + 8 * slot, rphase))) {
          load_raw(slot, xr);
          got_next = true;
        }
#endif
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(lead_bar0 + BA_FULL + 8 * (2 * stage + gof.role()));
          else mbar_arrive(bar0 + BA_FULL + 8 * (2 * stage + gof.role()));
        }
        if (more && !got_next) {
          mbar_wait(bar0 + BR_FULL + 8 * slot, rphase);
          load_raw(slot, xr);
        }
        gof.next(nchunks);
        if (++stage == TC_ASTAGES) { stage = 0; phase ^= 1; }
      }
    } else
    for (long long c = 0; c < my_chunks; ++c) {
      long long tq0 = 0;
      if (DBG == 6) tq0 = clock64();
      if (atma) {
        mbar_wait(bar0 + BR_FULL + 8 * slot, rphase);
      } else {
        cp_async_wait<TC_RSTAGES - 1>();      // the oldest group (this chunk) has landed
        __syncwarp();
      }
      if (DBG == 6) { long long tq1 = clock64(); tA_ld += tq1 - tq0; tq0 = tq1; }
      uint32_t hi[32], lo[32];
      // hi = RN to tf32 (11 significant bits); lo = x - hi exactly.  The correction operand is the
      // K = 64 bf16 row [lo(k = 0..31) | hi(k = 0..31)], two values per 32-bit column (lower k in the
      // low half): lo[0..15] = pairs of x_lo, lo[16..31] = pairs of x_hi, both rounded to nearest.
      const float* rowp = atma
          ? reinterpret_cast<const float*>(smem + TC_OFF_XPOSE + slot * 16384 + my_row * 128)
          : reinterpret_cast<const float*>(smem + TC_OFF_XPOSE) + (size_t)slot * 128 * TC_XLD + (size_t)my_row * TC_XLD;
      const int sw = atma ? (my_row & 7) : 0;   // TMA SWIZZLE_128B: 16-byte chunk index ^= row % 8
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (DBG != 9) x = *reinterpret_cast<const float4*>(rowp + ((i ^ sw) << 2));
        if (atma && !row_in_box) x = make_float4(0.f, 0.f, 0.f, 0.f);   // rows the box does not cover
        if (PREC == 2) {
          const uint32_t p01 = pack_f16x2_sat(x.x, x.y), p23 = pack_f16x2_sat(x.z, x.w);
          float g0, g1, g2, g3;
          unpack_f16x2(p01, g0, g1); unpack_f16x2(p23, g2, g3);
          lo[2 * i + 0] = p01; lo[2 * i + 1] = p23;
          lo[16 + 2 * i + 0] = pack_f16x2_sat(x.x - g0, x.y - g1);
          lo[16 + 2 * i + 1] = pack_f16x2_sat(x.z - g2, x.w - g3);
          continue;
        }
        const float h0 = tf32_rn_fast(x.x), h1 = tf32_rn_fast(x.y), h2 = tf32_rn_fast(x.z), h3 = tf32_rn_fast(x.w);
        hi[4 * i + 0] = __float_as_uint(h0); hi[4 * i + 1] = __float_as_uint(h1);
        hi[4 * i + 2] = __float_as_uint(h2); hi[4 * i + 3] = __float_as_uint(h3);
        if (PREC == 1) {
          lo[4 * i + 0] = __float_as_uint(tf32_rn_fast(x.x - h0)); lo[4 * i + 1] = __float_as_uint(tf32_rn_fast(x.y - h1));
          lo[4 * i + 2] = __float_as_uint(tf32_rn_fast(x.z - h2)); lo[4 * i + 3] = __float_as_uint(tf32_rn_fast(x.w - h3));
        } else {
          lo[2 * i + 0] = pack_bf16x2(x.x - h0, x.y - h1);
          lo[2 * i + 1] = pack_bf16x2(x.z - h2, x.w - h3);
          lo[16 + 2 * i + 0] = pack_bf16x2(h0, h1);
          lo[16 + 2 * i + 1] = pack_bf16x2(h2, h3);
        }
      }
      __syncwarp();                         // slot free
      if (DBG == 6) { long long tq1 = clock64(); tA_x += tq1 - tq0; tq0 = tq1; }
      if (atma) {
        if (lane == 0) mbar_arrive(bar0 + BR_EMPTY + 8 * slot);
      } else {
        if (have) { ensure(tile, sg); issue(sg, k0, slot); advance(); }
        else cp_async_commit();
      }
      if (DBG == 6) tA_is += clock64() - tq0;
      long long t0 = 0;
      if (DBG == 6) t0 = clock64();
      mbar_wait(bar0 + BA_EMPTY + 8 * stage, phase ^ 1);
      if (DBG == 6) { long long t1 = clock64(); tA_wait += t1 - t0; t0 = t1; ++tA_n; }
      tc_fence_after();
      const uint32_t ta = tmem_base + lane_off + TC_A_COL + stage * 64;
      if (DBG != 4 && DBG != 5 && DBG != 8 && DBG != 9) {
        if (PREC != 2) tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(lead_bar0 + BA_FULL + 8 * (2 * stage + gof.role()));
        else mbar_arrive(bar0 + BA_FULL + 8 * (2 * stage + gof.role()));
      }
      gof.next(nchunks);
      if (DBG == 6) tA_st += clock64() - t0;
      if (++stage == TC_ASTAGES) { stage = 0; phase ^= 1; }
      if (++slot == (atma ? TC_RSTAGES_TMA : TC_RSTAGES)) { slot = 0; rphase ^= 1; }
    }
    if (!atma) cp_async_wait<0>();
    if (DBG == 6 && tid == 0 && p.dbg) {
      unsigned long long* d = p.dbg + (size_t)blockIdx.x * 32;
      d[0] = (unsigned long long)tA_n; d[1] = (unsigned long long)tA_wait; d[2] = (unsigned long long)tA_st;
      d[3] = (unsigned long long)(clock64() - tA_tot);
      d[4] = (unsigned long long)tA_ld; d[5] = (unsigned long long)tA_x; d[6] = (unsigned long long)tA_is;
    }
  } else if (warp >= 8 && warp < 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REG_M));
    if (warp == 8) {
    // ================================================================= B producer (TMA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long tB_wait = 0;
      TcGroupOf gof;
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        const int n0 = (tile % n_tiles_n) * TC_BN;
        for (int c = 0; c < nchunks; ++c, gof.next(nchunks)) {
          long long t0 = 0;
          if (DBG == 6) t0 = clock64();
          mbar_wait(bar0 + BB_EMPTY + 8 * stage, phase ^ 1);     // every CTA of the cluster consumed it
          if (DBG == 6) tB_wait += clock64() - t0;
          const uint32_t full = bar0 + BB_FULL + 8 * (2 * stage + gof.role());
          if (DBG == 8 || DBG == 9) {                            // experiment: MMAs on stale smem, no B traffic
            if (CG == 1 || crank == 0) mbar_arrive(full);
            if (++stage == BST) { stage = 0; phase ^= 1; }
            continue;
          }
          if (CG == 2) {
            // CTA pair: this CTA's 64 weight rows (hi and pk) into ITS shared memory; the bytes of both CTAs are counted
            // on the rank-0 CTA's barrier, which its issuer waits on
            if (crank == 0) mbar_arrive_expect_tx(full, (PREC == 2 ? 1 : 2) * 16384);
            const uint32_t dst = sbase + stage * BSB;
            const uint32_t lfull = lead_bar0 + BB_FULL + 8 * (2 * stage + gof.role());
            if (PREC != 2)   // (PREC 2: the [h16 | lo] rows are the only B operand: 8 KB per CTA and chunk)
            tma_load_2d_cg2(dst, &map_hi, lfull, c * TC_BK, n0 + (TC_BN / 2) * (int)crank);
            tma_load_2d_cg2(dst + BPK, &map_lo, lfull, (PREC == 1 ? 1 : 2) * c * TC_BK, n0 + (TC_BN / 2) * (int)crank);
            if (++stage == BST) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(full, 2 * 16384);                // own share + the peers' shares
          if (TC_CLUSTER > 1) {
            // this CTA fetches rows [R*rank, R*rank+R) of the 128-row B tile (hi and lo) and
            // multicasts them into every CTA of the cluster: each weight byte crosses the L2->SM
            // fabric once per cluster instead of once per CTA
            constexpr int R = TC_BN / TC_CLUSTER;
            const uint32_t dst = sbase + stage * TC_STAGE_BYTES + crank * (R * 128);
            const uint16_t mask = (uint16_t)((1u << TC_CLUSTER) - 1);
            tma_load_2d_mc(dst, &map_hi, full, c * TC_BK, n0 + R * (int)crank, mask);
            tma_load_2d_mc(dst + 16384, &map_lo, full, (PREC == 1 ? 1 : 2) * c * TC_BK, n0 + R * (int)crank, mask);   // bf16 [hi | lo] row of this chunk (PREC 1: fp32 lo tile)
          } else {
            const uint32_t dst = sbase + stage * TC_STAGE_BYTES;
            tma_load_2d(dst, &map_hi, full, c * TC_BK, n0);
            tma_load_2d(dst + 16384, &map_lo, full, (PREC == 1 ? 1 : 2) * c * TC_BK, n0);
          }
          if (++stage == TC_BSTAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (DBG == 6 && p.dbg) p.dbg[(size_t)blockIdx.x * 32 + 8] = (unsigned long long)tB_wait;
    }
    __syncwarp();
    } else if (warp == 10) {
    // ================================================================= A producer by TMA (affine frame maps)
    if (lane == 0 && p.a_tma) {
      int slot = 0;
      uint32_t rphase = 0;
      const uint32_t bytes = (uint32_t)(p.fpt * J * 128);
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        const int f0 = tile_f0(tile);
        for (int sg = 0; sg < p.nseg; ++sg) {
          const int Kc = p.seg[sg].Kc;
          for (int k0 = 0; k0 < p.seg[sg].K; k0 += TC_BK) {
            const int tap = k0 / Kc;
            const int mi = p.a_map0[sg] + tap;
            const CUtensorMap* mp = (mi == 0) ? &amap0 : (mi == 1) ? &amap1 : &amap2;
            mbar_wait(bar0 + BR_EMPTY + 8 * slot, rphase ^ 1);
            const uint32_t full = bar0 + BR_FULL + 8 * slot;
            if (DBG == 8 || DBG == 9) {                      // experiment: no A traffic
              mbar_arrive(full);
              if (++slot == TC_RSTAGES_TMA) { slot = 0; rphase ^= 1; }
              continue;
            }
            mbar_arrive_expect_tx(full, bytes);
            // box {32 channels, J joints, fpt frames} -> fpt*J dense 128-byte rows, swizzled;
            // frames past the end of the tensor are zero-filled (ragged last tile, dummy tiles)
            tma_load_3d(sbase + TC_OFF_XPOSE + slot * 16384, mp, full, k0 - tap * Kc, 0, f0);
            if (++slot == TC_RSTAGES_TMA) { slot = 0; rphase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
    } else if (TC_DUAL_ISSUE && (warp == 9 || warp == 11)) {
    // ================================================================= MMA issuers (two)
    // The 8 MMAs of a chunk execute in ~500 cycles, but one thread needs ~620 cycles per chunk for the barrier
    // checks, descriptors, commits and probes around them ("MMAs + barrier hand-shakes only" ran at 0.68 ms where
    // the tensor pipe needs 0.41 ms, profiles/r02_tc_attribution.md): the issuing thread paces the kernel.  So
    // warp 9 issues the even flush groups and warp 11 the odd ones.  A flush group has its own accumulator buffer
    // (group g -> buffer g % 2), so issuer `role` owns buffer `role` and no ordering between the two threads' MMAs
    // is needed; operand stages are consumed in ring order, each released by the commit of the thread that used
    // it, and each issuer waits on its OWN instance of the operand "full" barriers (see the barrier table).
    {
      const int role = (warp == 11) ? 1 : 0;
      int bs = 0, as = 0;                  // ring positions of the next chunk in global order (both issuers' chunks)
      uint32_t aph = 0, bph = 0;           // bit s: parity this issuer waits for next on its barrier of stage s
      uint32_t gcount = 0;                 // flush groups so far (both issuers')
      long long tM[5] = {0, 0, 0, 0, 0};
      bool pre_a = false, pre_b = false;
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        for (int c0 = 0; c0 < nchunks; c0 += TC_FLUSH, ++gcount) {
          const int glen = min(TC_FLUSH, nchunks - c0);
          if ((int)(gcount % TC_NISSUE) != role) {        // the other issuer's group: step over its chunks
            as = (as + glen) % TC_ASTAGES;
            bs = (bs + glen) % TC_BSTAGES;
            continue;
          }
          const uint32_t mb = gcount % NMAIN;
          mbar_wait(bar0 + BM_EMPTY + 8 * mb, ((gcount / NMAIN) & 1) ^ 1);   // the epilogue has drained this buffer
          const uint32_t d_main = tmem_base + mb * TC_BN;
          for (int cg = 0; cg < glen; ++cg) {
            long long t0 = 0, t2 = 0, t3 = 0;
            if (DBG == 6) t0 = clock64();
            if (!pre_a) mbar_wait(bar0 + BA_FULL + 8 * (2 * as + role), (aph >> as) & 1u);
            if (DBG == 6) t2 = clock64();
            if (!pre_b) mbar_wait(bar0 + BB_FULL + 8 * (2 * bs + role), (bph >> bs) & 1u);
            if (DBG == 6) t3 = clock64();
            aph ^= 1u << as; bph ^= 1u << bs;
            tc_fence_after();
            const uint32_t a_hi = tmem_base + TC_A_COL + as * 64, a_pk = a_hi + 32;
            const uint32_t sb = sbase + bs * TC_STAGE_BYTES;
            const uint64_t b_hi = make_smem_desc(sb), b_pk = make_smem_desc(sb + 16384);
            const bool last_of_group = (cg == glen - 1);
            const int as_n = (as + 1 == TC_ASTAGES) ? 0 : as + 1, bs_n = (bs + 1 == TC_BSTAGES) ? 0 : bs + 1;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < TC_BK / 8; ++k) {
                const uint64_t adv = (uint64_t)(k * 2);      // 8 fp32 = 32 B = 2 x 16 B
                if (DBG != 7) umma_tf32_ts(d_main, a_hi + 8 * k, b_hi + adv, TC_IDESC, (cg | k) ? 1u : 0u);
                if (PREC == 1) {
                  umma_tf32_ts(d_main, a_pk + 8 * k, b_hi + adv, TC_IDESC, 1u);
                  umma_tf32_ts(d_main, a_hi + 8 * k, b_pk + adv, TC_IDESC, 1u);
                } else
                if (DBG != 3) umma_bf16_ts(d_main, a_pk + 8 * k, b_pk + adv, TC_IDESC_BF, (DBG == 7 && !(cg | k)) ? 0u : 1u);
              }
              if (TC_CLUSTER > 1) umma_commit_mc(bar0 + BB_EMPTY + 8 * bs, (uint16_t)((1u << TC_CLUSTER) - 1));
              else umma_commit(bar0 + BB_EMPTY + 8 * bs);     // frees the B smem stage when the MMAs retire
              umma_commit(bar0 + BA_EMPTY + 8 * as);          // frees the A tmem stage
              if (last_of_group) umma_commit(bar0 + BM_FULL + 8 * mb);   // group sum ready
            }
            __syncwarp();
            if (!last_of_group) {          // (after a group this issuer's next chunk is a whole group away)
              pre_a = mbar_try(bar0 + BA_FULL + 8 * (2 * as_n + role), (aph >> as_n) & 1u);
              pre_b = mbar_try(bar0 + BB_FULL + 8 * (2 * bs_n + role), (bph >> bs_n) & 1u);
            } else {
              pre_a = pre_b = false;
            }
            as = as_n; bs = bs_n;
            if (DBG == 6) {
              const long long t4 = clock64();
              tM[0] += 1; tM[2] += t2 - t0; tM[3] += t3 - t2; tM[4] += t4 - t3;
            }
          }
        }
      }
      if (DBG == 6 && p.dbg && lane == 0 && role == 0)
        for (int i = 0; i < 5; ++i) p.dbg[(size_t)blockIdx.x * 32 + 16 + i] = (unsigned long long)tM[i];
    }
    } else if (warp == 9 && (CG == 1 || crank == 0)) {
    // ================================================================= MMA issuer (CTA pair: of the rank-0 CTA only)
    // The whole warp runs this loop converged so that descriptors and barrier addresses live in
    // uniform registers; one elected lane issues.  (Issuing from inside `if (lane == 0)` made
    // every tcgen05.mma pay a ~55-cycle vector->uniform waterfall: 824 cycles per chunk for a
    // 768-cycle MMA budget, profiles/r01_tc_attribution.md.)
    {
      int bs = 0, as = 0;
      uint32_t bphase = 0, aphase = 0;
      uint32_t mcount = 0;                 // main buffers handed out so far
      long long tM[5] = {0, 0, 0, 0, 0};
      // The readiness probes of the NEXT chunk (~70 cycles each) are launched right after the MMAs of the
      // current one have been queued and only consumed at the top of the next iteration.
      bool pre_a = false, pre_b = false, pre_m = false;
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t mb = mcount % NMAIN;
          const int cg = c % TC_FLUSH;                    // position in the flush group
          long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
          if (DBG == 6) t0 = clock64();
          if (cg == 0 && !pre_m) mbar_wait(bar0 + BM_EMPTY + 8 * mb, ((mcount / NMAIN) & 1) ^ 1);
          pre_m = false;
          if (DBG == 6) t1 = clock64();
          if (!pre_a) mbar_wait(bar0 + BA_FULL + 16 * as, aphase);
          if (DBG == 6) t2 = clock64();
          if (!pre_b) mbar_wait(bar0 + BB_FULL + 16 * bs, bphase);
          if (DBG == 6) t3 = clock64();
          tc_fence_after();
          const uint32_t d_main = tmem_base + mb * TC_BN;
          const uint32_t a_hi = tmem_base + TC_A_COL + as * 64, a_pk = a_hi + 32;
          const uint32_t sb = sbase + bs * BSB;
          const uint64_t b_hi = make_smem_desc(sb), b_pk = make_smem_desc(sb + BPK);
          const bool last_of_group = (cg == TC_FLUSH - 1 || c == nchunks - 1);
          // ring positions of the next chunk (same rings across tile boundaries)
          const int as_n = (as + 1 == TC_ASTAGES) ? 0 : as + 1, bs_n = (bs + 1 == BST) ? 0 : bs + 1;
          const uint32_t aph_n = (as + 1 == TC_ASTAGES) ? (aphase ^ 1) : aphase;
          const uint32_t bph_n = (bs + 1 == BST) ? (bphase ^ 1) : bphase;
          // ONE elected block per chunk: every elect.sync + reconvergence costs ~50 cycles of this thread, and the
          // thread's instruction stream, not the tensor pipe, paces the chunk rate (8 MMAs execute in ~500 cycles,
          // the loop body took ~700 with one elected block per k-step: profiles/r02_tc_attribution.md).  The probes
          // of the next chunk's barriers follow in the shadow of the MMAs just queued.
          if (elect_one()) {
            if (PREC == 2) {
              // 6 MMAs of K = 16 on 16-bit operands (tensor-pipe time of 3 bf16 passes instead of 4): per 16-wide k-step
              //   A_h16.B_h16 + A_lo.B_h16 + A_h16.B_lo ; pair columns [h16 (16) | lo (16)], B row [h16 (64 B) | lo (64 B)]
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                umma_bf16_ts2(d_main, a_pk + 8 * k, b_pk + (uint64_t)(2 * k), TC_IDESC2_HH, (cg | k) ? 1u : 0u);
                umma_bf16_ts2(d_main, a_pk + 16 + 8 * k, b_pk + (uint64_t)(2 * k), TC_IDESC2_HH, 1u);
                umma_bf16_ts2(d_main, a_pk + 8 * k, b_pk + (uint64_t)(2 * (k + 2)), TC_IDESC2_HH, 1u);
              }
            } else
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);      // 8 fp32 = 32 B = 2 x 16 B
              if (CG == 2) {
                umma_tf32_ts2(d_main, a_hi + 8 * k, b_hi + adv, TC_IDESC2, (cg | k) ? 1u : 0u);
                if (PREC == 1) {
                  umma_tf32_ts2(d_main, a_pk + 8 * k, b_hi + adv, TC_IDESC2, 1u);
                  umma_tf32_ts2(d_main, a_hi + 8 * k, b_pk + adv, TC_IDESC2, 1u);
                } else {
                  umma_bf16_ts2(d_main, a_pk + 8 * k, b_pk + adv, TC_IDESC2_BF, 1u);
                }
                continue;
              }
              if (DBG != 7) umma_tf32_ts(d_main, a_hi + 8 * k, b_hi + adv, TC_IDESC, (cg | k) ? 1u : 0u);
              if (PREC == 1) {                  // 3xTF32: A_lo.B_hi and A_hi.B_lo as tf32 products of their own
                umma_tf32_ts(d_main, a_pk + 8 * k, b_hi + adv, TC_IDESC, 1u);
                umma_tf32_ts(d_main, a_hi + 8 * k, b_pk + adv, TC_IDESC, 1u);
              } else
              // correction: bf16 k-step k of the K=64 row [A_lo | A_hi] . [B_hi | B_lo]^T (8 columns of bf16
              // pairs in tensor memory, 32 bytes of the swizzled B row, like a tf32 k-step).  (Issuing the four
              // tf32 MMAs first and the four bf16 ones after them measured 5 % slower than interleaving them.)
              if (DBG != 3) umma_bf16_ts(d_main, a_pk + 8 * k, b_pk + adv, TC_IDESC_BF, (DBG == 7 && !(cg | k)) ? 0u : 1u);
            }
            if (CG == 2) {                                  // every release goes to both CTAs of the pair
              umma_commit_mc2(bar0 + BB_EMPTY + 8 * bs, (uint16_t)3);
              umma_commit_mc2(bar0 + BA_EMPTY + 8 * as, (uint16_t)3);
              if (last_of_group) umma_commit_mc2(bar0 + BM_FULL + 8 * mb, (uint16_t)3);
            } else {
            if (TC_CLUSTER > 1) umma_commit_mc(bar0 + BB_EMPTY + 8 * bs, (uint16_t)((1u << TC_CLUSTER) - 1));
            else umma_commit(bar0 + BB_EMPTY + 8 * bs);     // frees the B smem stage when the MMAs retire
            umma_commit(bar0 + BA_EMPTY + 8 * as);          // frees the A tmem stage
            if (last_of_group) umma_commit(bar0 + BM_FULL + 8 * mb);   // group sum ready
            }
          }
          __syncwarp();
          pre_a = mbar_try(bar0 + BA_FULL + 16 * as_n, aph_n);
          pre_b = mbar_try(bar0 + BB_FULL + 16 * bs_n, bph_n);
          if (last_of_group) {
            ++mcount;
            // the next chunk opens a flush group: probe its accumulator buffer as well
            pre_m = mbar_try(bar0 + BM_EMPTY + 8 * (mcount % NMAIN), ((mcount / NMAIN) & 1) ^ 1);
          }
          bs = bs_n; bphase = bph_n;
          as = as_n; aphase = aph_n;
          if (DBG == 6) {
            const long long t4 = clock64();
            tM[0] += 1; tM[1] += t1 - t0; tM[2] += t2 - t1; tM[3] += t3 - t2; tM[4] += t4 - t3;
          }
        }
      }
      if (DBG == 6 && p.dbg && lane == 0)
        for (int i = 0; i < 5; ++i) p.dbg[(size_t)blockIdx.x * 32 + 16 + i] = (unsigned long long)tM[i];
    }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REG_E));
    // ================================================================= epilogue warps 4..7 and 12..15
    // Two warpgroups split the tile by COLUMNS: group eh owns accumulator columns [64 eh, 64 eh + 64).
    // (One group reading all 128 columns needed ~8400 cycles per tile for its tcgen05.ld + epilogue,
    // 2.7x the MMA time of a K=128 tile: profiles/r01_tc_attribution.md.)  Warp w of either group
    // reads TMEM lane quadrant w % 4; each group has its own staging area and named barrier.
    const int eh = (warp >= 12) ? 1 : 0;
    const int ew = warp & 3;                 // TMEM lane quadrant
    const int et = ew * 32 + lane;           // 0..127 inside the group
    const int r = et;                        // tile row of this thread
    const int fr = r / J, ji = r - fr * J;
    const int fb = fr * J;                   // first row of this thread's frame
    const uint32_t ebar = 1u + (uint32_t)eh; // named barrier of this group
    uint32_t mcount = 0;
    long long tE_n = 0, tE_wait = 0, tE_tot = clock64(), tE_tiles = 0, tE_cw = 0, tE_cl = 0, tE_fl = 0;
    const uint32_t lane_off = ((uint32_t)(ew * 32) << 16) + (uint32_t)(eh * TC_EN);
    float* scratch = staging;                // TC_EPI_BYTES, laid out per epilogue kind below
    // SemCH tables of this thread (built once per kernel), for both masks:
    //   sA[m]: bits 0-3 degree of this thread's joint (<= TC_MAXDEG, checked by tc_supported), bit 4 "neighbour 0 is the
    //          joint itself", then 5 bits per neighbour joint.  The joint itself is moved to the FRONT of its list: its
    //          term comes from registers (X.W0 of the same row), so no staged row is read for it.
    //   sB[m]: 6 bits per neighbour k: the row of the coefficient slab that holds coef(ji, k).  The slab is k-major --
    //          neighbour 0 of every joint, then neighbour 1 of every joint that has one, ... -- so the 8 consecutive
    //          joints of a quarter warp read 8 consecutive rows: conflict-free.  (In the CSR order of the parameter the
    //          rows of a quarter warp are deg(j) apart and collided: 2.1 wavefronts per ideal one,
    //          profiles/r02_w_lines_semch1.txt; the self rows were another 40 % of the staged-row reads.)
    //   lz   : byte (4 m + u): the CSR nonzero that belongs into slab row (et + 128 u) / 8, 0xff = none -- this thread's
    //          (up to 4) float4 slots of the per-tile coefficient load
    unsigned long long sA0 = 0, sA1 = 0, sB0 = 0, sB1 = 0, lz = ~0ull;
    if (EPI == EPI_SEMCH) {
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (!p.coef[m]) continue;
        const NbrTable& nb = p.nbr[m];
        unsigned long long a = 0, b = 0;
        int base = 0;                                   // first slab row of neighbour index k
        for (int k = 0; k < TC_MAXDEG; ++k) {
          int rank = 0;                                 // joints before j that have a neighbour k
          for (int j = 0; j < J; ++j) {
            const int z0 = nb.row_ptr[j], dg = nb.row_ptr[j + 1] - z0;
            if (k >= dg) continue;
            int sp = -1;                                // position of the joint itself in its CSR row
            for (int z = 0; z < dg; ++z) if (nb.col[z0 + z] == j) sp = z;
            const int idx = (sp < 0) ? k : (k == 0 ? sp : (k - 1 < sp ? k - 1 : k));
            const int z = z0 + idx, row = base + rank;
            if (j == ji) {
              if (k == 0) a |= (unsigned long long)dg | ((unsigned long long)(sp >= 0 ? 1 : 0) << 4);
              a |= (unsigned long long)nb.col[z] << (5 + 5 * k);
              b |= (unsigned long long)row << (6 * k);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (row == ((et + 128 * u) >> 3)) lz = (lz & ~(0xffull << (8 * (4 * m + u)))) | ((unsigned long long)z << (8 * (4 * m + u)));
            ++rank;
          }
          base += rank;
        }
        if (m == 0) { sA0 = a; sB0 = b; } else { sA1 = a; sB1 = b; }
      }
    }
    for (int tile = cid; tile < total_tiles; tile += ncl) {
      const int tn = tile % n_tiles_n;
      const int f0 = tile_f0(tile);
      const int nf = max(0, min(p.fpt, p.F - f0));
      const int vrows = nf * J;
      const int n0 = tn * TC_BN + eh * TC_EN;          // first output column of this group
      const bool valid = r < vrows;
      const long long orow = (long long)(f0 + fr) * J + ji;

      if (EPI == EPI_SEMCH) {
        // this group's 32 channels of the coefficient slab -> smem (its previous readers are this
        // group's own warps, past the barrier that ends the loop body)
        const int mask_ = tn / p.tiles_per_mask;
        const int c0_ = (tn - mask_ * p.tiles_per_mask) * 64 + eh * 32;
        const int nnz_ = p.nbr[mask_].row_ptr[J];
        float* coef_w = scratch + 2 * 128 * TC_XLD + eh * 32;
        // (all loads first, then the stores: nnz <= TC_MAX_NNZ = 64 rows x 8 float4 = at most 4 per thread;
        //  load->store pairs one after the other exposed one global latency each, ~2000 cycles per tile)
        float4 cfv[TC_MAX_NNZ * 8 / 128];
        static_assert(TC_MAX_NNZ * 8 / 128 == 4, "lz holds 4 slots per mask");
#pragma unroll
        for (int u = 0; u < TC_MAX_NNZ * 8 / 128; ++u) {
          const int i = et + 128 * u, g = (i & 7) * 4;
          const int z = (int)((lz >> (8 * (4 * mask_ + u))) & 0xff);       // CSR nonzero of slab row i / 8
          cfv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (z != 0xff && c0_ + g < p.C) cfv[u] = ldg4(p.coef[mask_] + (long long)z * p.C + c0_ + g);
        }
#pragma unroll
        for (int u = 0; u < TC_MAX_NNZ * 8 / 128; ++u) {
          const int i = et + 128 * u, g = (i & 7) * 4;
          if (i < nnz_ * 8) *reinterpret_cast<float4*>(coef_w + (i >> 3) * TC_SLD + g) = cfv[u];
        }
      }
      float* ab_g = eh ? (scratch + 2 * 128 * TC_XLD) : ab_s;   // per-group copy of the a/b tile
      if (EPI == EPI_GLOBAL) {
        const int H2_ = 2 * p.heads;                  // <= 8 (tc_supported): loads first, then the stores
        float abv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = et + 128 * u;
          abv[u] = (u < H2_ && i / H2_ < vrows) ? __ldg(p.ab + ((long long)f0 * J) * H2_ + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (u < H2_) ab_g[et + 128 * u] = abv[u];
      }

      // ---- level-2 accumulation: group sums (TMEM) -> fp32 registers, round-to-nearest adds.
      // The first group is loaded straight into the accumulator registers; later groups and the
      // correction accumulator go through 2x32 temporaries.  (Issuing the correction loads together
      // with the first group when K is a single flush group made ptxas spill ~175 registers.)
      float acc[TC_EN];
      const int ngroups = (nchunks + TC_FLUSH - 1) / TC_FLUSH;
      const bool noflush = (DBG == 1 || DBG == 5 || DBG == 8 || DBG == 9);
      {
        const uint32_t mb = mcount % NMAIN;
        long long t0 = 0;
        if (DBG == 6) t0 = clock64();
        mbar_wait(bar0 + BM_FULL + 8 * mb, (mcount / NMAIN) & 1);
        if (DBG == 6) { tE_n += 1; tE_wait += clock64() - t0; t0 = clock64(); }
        tc_fence_after();
        const uint32_t taddr = tmem_base + mb * TC_BN + lane_off;
        if (noflush) {
#pragma unroll
          for (int i = 0; i < TC_EN; ++i) acc[i] = 0.f;
        } else {
          uint32_t* au = reinterpret_cast<uint32_t*>(acc);
          tmem_ld32_async(taddr, au);
          tmem_ld32_async(taddr + 32, au + 32);
          tmem_wait_ld(au);
          tmem_wait_ld(au + 32);
          if (PREC == 2) {                     // weights were split as 2^8 W
#pragma unroll
            for (int i = 0; i < TC_EN; ++i) acc[i] *= TC_F16_WUNSCALE;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(lead_bar0 + BM_EMPTY + 8 * mb);
          else mbar_arrive(bar0 + BM_EMPTY + 8 * mb);
        }
        ++mcount;
        if (DBG == 6) tE_fl += clock64() - t0;
      }
      for (int c = 1; c < ngroups; ++c) {
        const uint32_t mb = mcount % NMAIN;
        long long t0 = 0;
        if (DBG == 6) t0 = clock64();
        mbar_wait(bar0 + BM_FULL + 8 * mb, (mcount / NMAIN) & 1);
        if (DBG == 6) { tE_n += 1; tE_wait += clock64() - t0; }
        tc_fence_after();
        const uint32_t taddr = tmem_base + mb * TC_BN + lane_off;
        if (!noflush) {
          uint32_t va[32], vb[32];
          tmem_ld32_async(taddr, va);
          tmem_ld32_async(taddr + 32, vb);
          tmem_wait_ld(va);
          tmem_wait_ld(vb);
          // (packed add.rn.f32x2 here made ptxas spill in the PLAIN kernel: the pair alignment fights the tcgen05.ld targets)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (PREC == 2) {                   // exact power-of-two scale folded into the round-to-nearest add
              acc[i] = fmaf(__uint_as_float(va[i]), TC_F16_WUNSCALE, acc[i]);
              acc[32 + i] = fmaf(__uint_as_float(vb[i]), TC_F16_WUNSCALE, acc[32 + i]);
            } else {
            acc[i] += __uint_as_float(va[i]);
            acc[32 + i] += __uint_as_float(vb[i]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(lead_bar0 + BM_EMPTY + 8 * mb);
          else mbar_arrive(bar0 + BM_EMPTY + 8 * mb);
        }
        ++mcount;
      }
      if (DBG == 6) ++tE_tiles;

      if (EPI == EPI_PLAIN) {
        // Row-per-thread registers -> warp-private smem patch -> COALESCED 128-bit stores (8 lanes
        // cover one 128-byte row segment).  Storing straight from the row-per-thread layout issues
        // 32 half-sector writes per instruction: clock64 attribution showed ~7500 of the ~8000
        // epilogue cycles of a tile in those stores (profiles/r01_tc_attribution.md).
        float* patch = scratch + (eh * 4 + ew) * (32 * TC_XLD);
        const int rsub = lane >> 3, ch = lane & 7;
        // Residual rows of the 8 tile rows this lane stores (stage 1x1 convs, gast_net.py:174): ONE 32-bit
        // division per tile -- the frames of a tile are consecutive.  (map_frame()'s 64-bit divisions per
        // stored float4 made the residual GEMMs 2.5x slower than their MMA time.)
        unsigned rres[8];                        // residual row offsets in 16-byte units (res_ld % 4 == 0)
        if (p.res) {
          const int b0 = f0 / p.res_map.T_out, t0 = f0 - b0 * p.res_map.T_out;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = ew * 32 + 4 * i + rsub;
            const int frr = rr / J, jr = rr - frr * J;
            int b = b0, t = t0 + frr;
            while (t >= p.res_map.T_out) { t -= p.res_map.T_out; ++b; }
            rres[i] = (unsigned)(((((long long)b * p.res_map.T_in + (long long)t * p.res_map.t_mul + p.res_map.t_off) * J + jr) *
                                  (long long)p.res_ld) >> 2);
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int nq = n0 + q * 32;
          if (nq < p.N) {                       // warp-uniform
            const int n = nq + ch * 4;
            // residual values of this lane's 8 stores: all 8 loads in flight BEFORE the patch round trip
            // (one exposed HBM latency per 32 columns instead of one per stored float4)
            float4 rv[8];
            if (p.res) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = ew * 32 + 4 * i + rsub;
                rv[i] = (rr < vrows && n < p.N) ? ldg4(p.res + 4ull * rres[i] + n) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int nn = nq + g * 4;
              float4 o = make_float4(acc[q * 32 + g * 4], acc[q * 32 + g * 4 + 1], acc[q * 32 + g * 4 + 2],
                                     acc[q * 32 + g * 4 + 3]);
              if (p.bias && nn < p.N) { float4 bb = ldg4(p.bias + nn); o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w; }
              if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              *reinterpret_cast<float4*>(patch + lane * TC_XLD + g * 4) = o;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = ew * 32 + 4 * i + rsub;          // tile row stored by this lane
              if (rr < vrows && n < p.N) {
                float4 o = *reinterpret_cast<const float4*>(patch + (4 * i + rsub) * TC_XLD + ch * 4);
                if (p.res) { o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w; }
                // rows of a tile are consecutive in the output: row index = f0*J + rr
                *reinterpret_cast<float4*>(p.out + ((long long)f0 * J + rr) * p.ld_out + n) = o;
              }
            }
            __syncwarp();
          }
        }
      } else if (EPI == EPI_SEMCH) {
        // Weight rows are interleaved per 32 channels (tc_split_kernel, semch order): this group's
        // 64 columns are  acc[0..31] = X.W0 (self term), acc[32..63] = X.W1 (neighbour term)  of
        // channels c0 .. c0+31.
        const int mask = tn / p.tiles_per_mask;
        const int c0 = (tn - mask * p.tiles_per_mask) * 64 + eh * 32;
        float* Hs = scratch + eh * (128 * TC_XLD);               // [128 rows][32 ch], stride TC_XLD
        const float* coef_r = scratch + 2 * 128 * TC_XLD + eh * 32;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(Hs + r * TC_XLD + g * 4) =
              make_float4(acc[32 + g * 4], acc[32 + g * 4 + 1], acc[32 + g * 4 + 2], acc[32 + g * 4 + 3]);
        const float* h0 = acc;
        epi_bar_sync(ebar);
        float ov[32];
        {
          // this row's neighbour list comes packed in a register (built once per kernel): the loop is
          // fully unrolled and predicated, so the shared-memory loads of all neighbours can be in flight
          // together (a runtime z loop with an indexed constant load per step exposed ~2 latencies per neighbour)
          const unsigned long long ta_ = mask ? sA1 : sA0, tb_ = mask ? sB1 : sB0;
          const int cnt = valid ? (int)(ta_ & 0xf) : 0;
          const bool self0 = ((ta_ >> 4) & 1) != 0;
#pragma unroll
          for (int gq = 0; gq < 2; ++gq) {           // 16 channels at a time
            float o[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
              const int c = c0 + gq * 16 + g * 4;
              if (p.shift && c < p.C) sh = ldg4(p.shift + mask * p.C + c);
              o[g * 4] = sh.x; o[g * 4 + 1] = sh.y; o[g * 4 + 2] = sh.z; o[g * 4 + 3] = sh.w;
            }
#pragma unroll
            for (int k = 0; k < TC_MAXDEG; ++k) {
              if (k < cnt) {
                const int jn = (int)((ta_ >> (5 + 5 * k)) & 31);
                const float* crow = coef_r + (int)((tb_ >> (6 * k)) & 63) * TC_SLD + gq * 16;
                if (k == 0 && self0) {                 // the joint itself: X.W0 of this row, from registers
#pragma unroll
                  for (int g = 0; g < 4; ++g) {
                    const float4 cf = *reinterpret_cast<const float4*>(crow + g * 4);
#if GAST_TC_SEMCH_FFMA2
                    semch_fma2(cf, make_float4(h0[gq * 16 + g * 4], h0[gq * 16 + g * 4 + 1], h0[gq * 16 + g * 4 + 2],
                                               h0[gq * 16 + g * 4 + 3]), o + g * 4);
#else
                    o[g * 4] = fmaf(cf.x, h0[gq * 16 + g * 4], o[g * 4]); o[g * 4 + 1] = fmaf(cf.y, h0[gq * 16 + g * 4 + 1], o[g * 4 + 1]);
                    o[g * 4 + 2] = fmaf(cf.z, h0[gq * 16 + g * 4 + 2], o[g * 4 + 2]); o[g * 4 + 3] = fmaf(cf.w, h0[gq * 16 + g * 4 + 3], o[g * 4 + 3]);
#endif
                  }
                } else {
                  const float* hrow = Hs + (fb + jn) * TC_XLD + gq * 16;
#pragma unroll
                  for (int g = 0; g < 4; ++g) {
                    const float4 cf = *reinterpret_cast<const float4*>(crow + g * 4);
                    const float4 hv = *reinterpret_cast<const float4*>(hrow + g * 4);
#if GAST_TC_SEMCH_FFMA2
                    semch_fma2(cf, hv, o + g * 4);
#else
                    o[g * 4] = fmaf(cf.x, hv.x, o[g * 4]); o[g * 4 + 1] = fmaf(cf.y, hv.y, o[g * 4 + 1]);
                    o[g * 4 + 2] = fmaf(cf.z, hv.z, o[g * 4 + 2]); o[g * 4 + 3] = fmaf(cf.w, hv.w, o[g * 4 + 3]);
#endif
                  }
                }
              }
            }
#pragma unroll
            for (int g = 0; g < 16; ++g) ov[gq * 16 + g] = p.relu ? fmaxf(o[g], 0.f) : o[g];
          }
        }
        epi_bar_sync(ebar);  // every neighbour read of the staged H1 tile is done: reuse it for the output
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(Hs + r * TC_XLD + g * 4) =
              make_float4(ov[g * 4], ov[g * 4 + 1], ov[g * 4 + 2], ov[g * 4 + 3]);
        __syncwarp();
        {
          // coalesced: 8 lanes cover the 128-byte row segment of one row, 4 rows per instruction
          const int rs = lane >> 3, chq = lane & 7;
          const int c = c0 + chq * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = ew * 32 + 4 * i + rs;
            if (rr < vrows && c < p.C)
              *reinterpret_cast<float4*>(p.out + ((long long)f0 * J + rr) * p.ld_out + mask * p.C + c) =
                  *reinterpret_cast<const float4*>(Hs + rr * TC_XLD + chq * 4);
          }
        }
        epi_bar_sync(ebar);  // staging / coefficient slab free for the next tile
      } else {               // EPI_GLOBAL
        const int H2 = 2 * p.heads;
        float* Gs = scratch + eh * (128 * TC_XLD);               // [128 rows][32 cols], stride TC_XLD
        float att[TC_JMAX];                                      // attention row of (this joint, head att_h)
        int att_h = -1;
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const int nb0 = n0 + ps * 32;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int n = nb0 + g * 4;
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bg && n < p.N) bb = ldg4(p.bg + n);
            *reinterpret_cast<float4*>(Gs + r * TC_XLD + g * 4) =
                make_float4(acc[ps * 32 + g * 4] + bb.x, acc[ps * 32 + g * 4 + 1] + bb.y,
                            acc[ps * 32 + g * 4 + 2] + bb.z, acc[ps * 32 + g * 4 + 3] + bb.w);
          }
          epi_bar_sync(ebar);
          // (a register/patch-staged, coalesced-store variant of this mix measured 70 % slower than the
          //  direct row-per-thread stores below: profiles/r01_tc_attribution.md)
          if (valid && nb0 < p.N) {
            const int nend = min(nb0 + 32, p.N);
            const int h_lo = nb0 / p.Cg, h_hi = (nend - 1) / p.Cg;
            for (int h = h_lo; h <= h_hi; ++h) {
              // attention row of joint ji: softmax_j(LeakyReLU_0.2(a_i + b_j)) + C_k[i,j]; kept across the
              // two 32-column passes when they belong to the same head (Cg >= 64)
              if (h != att_h) {
                att_h = h;
                const float a = ab_g[r * H2 + 2 * h];
                float mx = -3.4e38f;
#pragma unroll
                for (int j = 0; j < TC_JMAX; ++j) {
                  if (j < J) {
                    float s = a + ab_g[(fb + j) * H2 + 2 * h + 1];
                    s = (s >= 0.f) ? s : 0.2f * s;
                    att[j] = s;
                    mx = fmaxf(mx, s);
                  }
                }
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < TC_JMAX; ++j)
                  if (j < J) { att[j] = expf(att[j] - mx); sum += att[j]; }
                const float inv = 1.f / sum;
                const float* ck = p.ck + ((long long)h * J + ji) * J;
#pragma unroll
                for (int j = 0; j < TC_JMAX; ++j)
                  if (j < J) att[j] = att[j] * inv + __ldg(ck + j);
              }
              const int cbeg = max(h * p.Cg, nb0), cend = min((h + 1) * p.Cg, nend);
              // y[i, n] = sum_j att[j] g[j, n]: 4 float4 columns per iteration = 16 independent FMA chains
              // (measured neutral against one float4 at a time; an explicit one-step-ahead load of g was 4 % slower)
              int n = cbeg;
              for (; n + 16 <= cend; n += 16) {
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0, o2 = o0, o3 = o0;
                const float* scol = Gs + fb * TC_XLD + (n - nb0);
#pragma unroll
                for (int j = 0; j < TC_JMAX; ++j) {
                  if (j < J) {
                    const float a_ = att[j];
                    const float4 g0 = *reinterpret_cast<const float4*>(scol + j * TC_XLD);
                    const float4 g1 = *reinterpret_cast<const float4*>(scol + j * TC_XLD + 4);
                    const float4 g2 = *reinterpret_cast<const float4*>(scol + j * TC_XLD + 8);
                    const float4 g3 = *reinterpret_cast<const float4*>(scol + j * TC_XLD + 12);
                    o0.x = fmaf(a_, g0.x, o0.x); o0.y = fmaf(a_, g0.y, o0.y); o0.z = fmaf(a_, g0.z, o0.z); o0.w = fmaf(a_, g0.w, o0.w);
                    o1.x = fmaf(a_, g1.x, o1.x); o1.y = fmaf(a_, g1.y, o1.y); o1.z = fmaf(a_, g1.z, o1.z); o1.w = fmaf(a_, g1.w, o1.w);
                    o2.x = fmaf(a_, g2.x, o2.x); o2.y = fmaf(a_, g2.y, o2.y); o2.z = fmaf(a_, g2.z, o2.z); o2.w = fmaf(a_, g2.w, o2.w);
                    o3.x = fmaf(a_, g3.x, o3.x); o3.y = fmaf(a_, g3.y, o3.y); o3.z = fmaf(a_, g3.z, o3.z); o3.w = fmaf(a_, g3.w, o3.w);
                  }
                }
                float* op = p.out + orow * p.ld_out + n;
                *reinterpret_cast<float4*>(op) = o0;
                *reinterpret_cast<float4*>(op + 4) = o1;
                *reinterpret_cast<float4*>(op + 8) = o2;
                *reinterpret_cast<float4*>(op + 12) = o3;
              }
              for (; n < cend; n += 4) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* scol = Gs + fb * TC_XLD + (n - nb0);
#pragma unroll
                for (int j = 0; j < TC_JMAX; ++j) {
                  if (j < J) {
                    float4 g4 = *reinterpret_cast<const float4*>(scol + j * TC_XLD);
                    o.x = fmaf(att[j], g4.x, o.x); o.y = fmaf(att[j], g4.y, o.y);
                    o.z = fmaf(att[j], g4.z, o.z); o.w = fmaf(att[j], g4.w, o.w);
                  }
                }
                *reinterpret_cast<float4*>(p.out + orow * p.ld_out + n) = o;
              }
            }
          }
          epi_bar_sync(ebar);  // staging (and, after the last pass, the a/b tile) free for reuse
        }
      }
    }
    if (DBG == 6 && et == 0 && eh == 0 && p.dbg) {
      p.dbg[(size_t)blockIdx.x * 32 + 27] = (unsigned long long)tE_tiles;
      p.dbg[(size_t)blockIdx.x * 32 + 28] = (unsigned long long)tE_cw;
      p.dbg[(size_t)blockIdx.x * 32 + 29] = (unsigned long long)tE_cl;
      p.dbg[(size_t)blockIdx.x * 32 + 30] = (unsigned long long)tE_fl;
      p.dbg[(size_t)blockIdx.x * 32 + 24] = (unsigned long long)tE_n;
      p.dbg[(size_t)blockIdx.x * 32 + 25] = (unsigned long long)tE_wait;
      p.dbg[(size_t)blockIdx.x * 32 + 26] = (unsigned long long)(clock64() - tE_tot);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (TC_CLUSTER > 1) cluster_sync_all();   // peers may multicast into / arrive on this CTA until they finish
  if (warp == 9) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
  (void)0;
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
// Residual truncation bias of the in-TMEM accumulation of one flush group (measured -8.7e-8 relative
// per 4 MMAs, tools/tc_probe.py): the partial sum after every MMA is truncated by ~0.5 ulp towards
// zero, i.e. the product of MMA s is under-counted by TC_TRUNC_C * (MMAs left in its group).
// It is added back through the correction accumulator by folding it into W_lo (it is <= 2^-20 of W,
// within W_lo's own rounding budget).
constexpr float TC_TRUNC_C = 3.5e-8f;

// semch != 0: W is the SemCH packing of gast_api.cu ([W0 of 64 channels | W1 of the same 64] per
// 128-row tile); the copies interleave it per 32 channels ([W0 32 | W1 32 | W0 next 32 | W1 next 32]) so
// that each epilogue warpgroup (64 accumulator columns) holds self and neighbour terms of the same channels.
// `pk` is the bf16 operand of the correction product: row n, chunk c (32 k's) holds
// [hi(k = 32c .. 32c+31) | lo(k = 32c .. 32c+31)] as 64 bf16 = one 128-byte swizzle row of the B stage.
__global__ void tc_split_kernel(const float* __restrict__ w, float* __restrict__ hi, unsigned short* __restrict__ pk,
                                long long n, int K, int semch, int prec) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long src = i;
  const long long row = i / K;
  if (semch) {
    const int within = (int)(row & 127), half = within >> 6, sub = within & 63;
    const int srow = (sub >> 5) * 64 + half * 32 + (sub & 31);
    src = ((row & ~127LL) + srow) * K + (i - row * K);
  }
  float x = w[src];
  float h = tf32_rna(x);
  const int k = (int)(i - row * K);
  const int nchunks = K / TC_BK;
  const int c = k / TC_BK, g = c / TC_FLUSH;
  const int glen = min(TC_FLUSH, nchunks - g * TC_FLUSH);            // chunks in this flush group
  // MMA index of this product inside its flush group: a chunk issues 8 MMAs into the group's accumulator, in the
  // order main(k-step 0), correction(0), main(1), correction(1), ...; every one of them truncates the accumulator
  // (prec 1, 3xTF32: 12 MMAs per chunk in the order main(k), A_lo.B_hi(k), A_hi.B_lo(k))
  const int per_k = prec == 1 ? 3 : 2;
  if (prec == 2) {
    // fp16 hi | fp16 remainder of 2^8 W, one 128-byte row per chunk as in prec 0; 6 MMAs per chunk in the order
    // (A_h16.B_h16, A_lo.B_h16, A_h16.B_lo) per 16-wide k-step
    const float xs = x * TC_F16_WSCALE;            // (tc_prepare_weights checked max|W| * 2^8 < 65504)
    const __half hh = __float2half_rn(xs);
    const float hf = __half2float(hh);
    const int s2 = (c - g * TC_FLUSH) * 6 + 3 * ((k % TC_BK) / 16);
    const float left = (float)(glen * 6 - s2);
    const float lo2 = (xs - hf) + TC_TRUNC_C * left * hf;
    unsigned short* prow2 = pk + row * 2 * K + (long long)c * 2 * TC_BK + (k - c * TC_BK);
    prow2[0] = __half_as_ushort(hh);
    prow2[TC_BK] = __half_as_ushort(__float2half_rn(lo2));
    hi[i] = hf;
    return;
  }
  const int s = (c - g * TC_FLUSH) * (per_k * TC_BK / 8) + per_k * ((k % TC_BK) / 8);
  const float steps_left = (float)(glen * (per_k * TC_BK / 8) - s);      // truncations this product still sees
  hi[i] = h;
  const float lo = (x - h) + TC_TRUNC_C * steps_left * h;
  if (prec == 1) {                       // fp32 lo matrix [N][K], the tf32 operand of the A_hi.B_lo product
    reinterpret_cast<float*>(pk)[i] = tf32_rna(lo);
    return;
  }
  unsigned short* prow = pk + row * 2 * K + (long long)c * 2 * TC_BK + (k - c * TC_BK);
  prow[0] = __bfloat16_as_ushort(__float2bfloat16_rn(h));
  prow[TC_BK] = __bfloat16_as_ushort(__float2bfloat16_rn(lo));
}

typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tc_encode_fn tc_get_encode() {
  static tc_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tc_encode_fn>(p);
  }
  return fn;
}

// W: [N][K] fp32 K-major (device).  Allocates hi/lo once, splits, encodes the TMA maps.
// Returns 0 on success, a cudaError_t / -1 otherwise.
__global__ void tc_absmax_kernel(const float* __restrict__ w, long long n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float a = fabsf(w[i]);
    m = (a > m || a != a) ? (a != a ? __int_as_float(0x7f800000) : a) : m;     // NaN counts as "too large"
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));             // non-negative floats order like their bits
}

inline int tc_prepare_weights(TcWeights& t, const float* W, int N, int K, cudaStream_t st,
                              std::vector<void*>* owned, int semch = 0, int prec = 0) {
  t.ready = false;
  if (K % TC_BK != 0 || N % 4 != 0) return 0;            // shape not taken by this core (FFMA runs it)
  if (prec == 2 && K < 256) prec = 0;   // K = 128 layers are bound by HBM / their epilogue: nothing to gain, keep fp32's range
  if (prec == 2) {
    // the fp16 form needs 2^8 |W| inside fp16's range: a GEMM whose (BatchNorm-folded) weights exceed it keeps tf32 + bf16
    // (the check reads max|W| back: when the caller is capturing a CUDA graph it cannot, and the GEMM keeps tf32 + bf16)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) prec = 0;
  }
  if (prec == 2) {
    static unsigned* d_max_dev[64] = {};                  // one word per device
    int dev = 0;
    cudaGetDevice(&dev);
    unsigned*& d_max = d_max_dev[dev & 63];
    if (!d_max && cudaMalloc(&d_max, sizeof(unsigned)) != cudaSuccess) return -1;
    unsigned hmax = 0;
    cudaMemsetAsync(d_max, 0, sizeof(unsigned), st);
    tc_absmax_kernel<<<64, 256, 0, st>>>(W, (long long)N * K, d_max);
    if (cudaMemcpyAsync(&hmax, d_max, sizeof(unsigned), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) return -1;
    float fmax_;
    memcpy(&fmax_, &hmax, sizeof(float));
    if (!(fmax_ * TC_F16_WSCALE < 60000.f)) prec = 0;
  }
  tc_encode_fn enc = tc_get_encode();
  if (!enc) return -1;
  if (!t.hi || t.N != N || t.K != K || t.prec != prec) {
    void* a = nullptr; void* b = nullptr;
    cudaError_t e = cudaMalloc(&a, sizeof(float) * (size_t)N * K);
    if (e != cudaSuccess) return (int)e;
    e = cudaMalloc(&b, sizeof(float) * (size_t)N * K);
    if (e != cudaSuccess) return (int)e;
    owned->push_back(a); owned->push_back(b);
    t.hi = (float*)a; t.lo = (float*)b; t.N = N; t.K = K; t.prec = prec;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)(TC_BN / TC_CLUSTER)};   // each CTA fetches its share
    cuuint32_t estr[2] = {1, 1};
    CUresult r1 = enc(&t.map_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, t.hi, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // the bf16 [hi | lo] rows: [N][2K] bf16 (the same bytes per row as the fp32 matrix), 64 elements = 128 B per box row
    cuuint64_t dims_b[2] = {(cuuint64_t)2 * K, (cuuint64_t)N};
    cuuint32_t box_b[2] = {(cuuint32_t)(2 * TC_BK), (cuuint32_t)(TC_BN / TC_CLUSTER)};
    CUresult r2 = prec == 1
        ? enc(&t.map_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, t.lo, dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        : enc(&t.map_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, t.lo, dims_b, strides, box_b, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return -1;
  }
  long long n = (long long)N * K;
  if (semch && N % 128) return -1;
  tc_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, t.hi, reinterpret_cast<unsigned short*>(t.lo), n, K, semch, prec);
  t.ready = true;
  return 0;
}

inline bool tc_supported(const GemmP& p, int epi, const TcWeights& t) {
  if (!t.ready || t.K != p.ldw) return false;
  if (p.J > TC_JMAX || p.N % 4) return false;
  if (p.res) {   // the epilogue keeps residual row offsets as 32-bit counts of 16-byte units
    const long long max_row = ((long long)(p.F / p.res_map.T_out) + 1) * p.res_map.T_in * p.J;
    if (p.res_ld % 4 || max_row * p.res_ld * 4 >= (1LL << 34)) return false;
  }
  for (int s = 0; s < p.nseg; ++s) {
    const ASeg& sg = p.seg[s];
    if (sg.K % TC_BK || sg.Kc % TC_BK || sg.ld % 4 || sg.tap_stride % 4) return false;
    if (reinterpret_cast<uintptr_t>(sg.base) & 15) return false;
  }
  if (epi == EPI_SEMCH) {
    if (p.C % 4) return false;
    for (int m = 0; m < 2; ++m) {
      if (!p.coef[m]) continue;
      if (p.nbr[m].row_ptr[p.J] > TC_MAX_NNZ) return false;
      for (int j = 0; j < p.J; ++j)
        if (p.nbr[m].row_ptr[j + 1] - p.nbr[m].row_ptr[j] > TC_MAXDEG) return false;
    }
  }
  if (epi == EPI_GLOBAL && (p.Cg % 4 || p.heads > 4)) return false;
  return true;
}

// Tensor maps of the A operand: one per (segment, temporal tap), possible when the segment's frame
// map is affine in the output frame index f:  in_frame = t_mul * f + t_off  (T_in == t_mul * T_out).
// dims {channels, joints, frames}, box {32, J, fpt}, SWIZZLE_128B, out-of-range frames read as 0.
struct TcAMaps {
  CUtensorMap m[3];
  bool ok = false;
};

inline void tc_build_amaps(GemmP& p, TcAMaps& am) {
  am.ok = false;
  p.a_tma = 0;
  tc_encode_fn enc = tc_get_encode();
  if (!enc) return;
  int nmaps = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const ASeg& sg = p.seg[s];
    const int taps = sg.K / sg.Kc;
    if ((long long)sg.map.T_in != (long long)sg.map.t_mul * sg.map.T_out) return;   // not affine
    if (nmaps + taps > 3) return;
    p.a_map0[s] = nmaps;
    for (int tp = 0; tp < taps; ++tp) {
      const float* base = sg.base + (long long)sg.map.t_off * p.J * sg.ld + (long long)tp * sg.tap_stride;
      cuuint64_t dims[3] = {(cuuint64_t)sg.Kc, (cuuint64_t)p.J, (cuuint64_t)p.F};
      cuuint64_t strides[2] = {(cuuint64_t)sg.ld * sizeof(float),
                               (cuuint64_t)sg.map.t_mul * p.J * sg.ld * sizeof(float)};
      cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)p.J, (cuuint32_t)p.fpt};
      cuuint32_t estr[3] = {1, 1, 1};
      if (strides[0] % 16 || strides[1] % 16 || (reinterpret_cast<uintptr_t>(base) & 15)) return;
      CUresult r = enc(&am.m[nmaps], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return;
      ++nmaps;
    }
  }
  for (int i = nmaps; i < 3; ++i) am.m[i] = am.m[0];
  am.ok = nmaps > 0;
  p.a_tma = am.ok ? 1 : 0;
}

template <int EPI, int DBG, int PREC = 0, int CG = 1>
inline int tc_launch_one(int grid, cudaStream_t st, const GemmP& p_in, const TcWeights& t, int nt, int items) {
  GemmP p = p_in;
  TcAMaps am;
  tc_build_amaps(p, am);
  if (!am.ok) { am.m[0] = t.map_hi; am.m[1] = t.map_hi; am.m[2] = t.map_hi; }
  static bool attr_set[64];                 // the shared-memory opt-in is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute((const void*)gemm_tc_kernel<EPI, DBG, PREC, CG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TC_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI, DBG, PREC, CG>, p, t.map_hi, t.map_lo, am.m[0], am.m[1], am.m[2], nt, items);
}

inline int tc_launch(int sm_count, cudaStream_t st, int epi, const GemmP& p, const TcWeights& t, int dbg = 0) {
  const int mt = (p.F + p.fpt - 1) / p.fpt;
  const int nt = (p.N + TC_BN - 1) / TC_BN;
  const long long items = (long long)((mt + TC_CLUSTER - 1) / TC_CLUSTER) * nt;
  if (items > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  const int max_cl = sm_count / TC_CLUSTER;
  const int grid = TC_CLUSTER * (int)(items < max_cl ? items : max_cl);
  // CTA-pair form (cta_group::2): GAST_TC_CG=1 selects the single-CTA MMAs with B multicast (A/B runs)
  static const int cg = (TC_CLUSTER == 2 && !TC_DUAL_ISSUE) ? (getenv("GAST_TC_CG") ? atoi(getenv("GAST_TC_CG")) : TC_CG_DEFAULT) : 1;
  if (t.prec == 1) {                       // 3xTF32 (training): plain epilogue only
    if (epi != EPI_PLAIN || dbg) return (int)cudaErrorInvalidValue;
#if GAST_TC_CLUSTER == 2 && !GAST_TC_DUAL_ISSUE
    if (cg == 2) return tc_launch_one<EPI_PLAIN, 0, 1, 2>(grid, st, p, t, nt, (int)items);
#endif
    return tc_launch_one<EPI_PLAIN, 0, 1>(grid, st, p, t, nt, (int)items);
  }
#if GAST_TC_CLUSTER == 2 && !GAST_TC_DUAL_ISSUE
  if (t.prec == 2) {                       // fp16 hi + bf16 remainder (GAST_TC_F16=1): CTA pair only
    if (cg != 2 || dbg) return (int)cudaErrorInvalidValue;
    if (epi == EPI_PLAIN) return tc_launch_one<EPI_PLAIN, 0, 2, 2>(grid, st, p, t, nt, (int)items);
    if (epi == EPI_SEMCH) return tc_launch_one<EPI_SEMCH, 0, 2, 2>(grid, st, p, t, nt, (int)items);
    return tc_launch_one<EPI_GLOBAL, 0, 2, 2>(grid, st, p, t, nt, (int)items);
  }
  if (cg == 2 && !dbg) {
    if (epi == EPI_PLAIN) return tc_launch_one<EPI_PLAIN, 0, 0, 2>(grid, st, p, t, nt, (int)items);
    if (epi == EPI_SEMCH) return tc_launch_one<EPI_SEMCH, 0, 0, 2>(grid, st, p, t, nt, (int)items);
    return tc_launch_one<EPI_GLOBAL, 0, 0, 2>(grid, st, p, t, nt, (int)items);
  }
#endif
#if GAST_TC_CLUSTER == 2 && !GAST_TC_DUAL_ISSUE
  if (cg == 2 && dbg == 4) return tc_launch_one<EPI_PLAIN, 4, 0, 2>(grid, st, p, t, nt, (int)items);
  if (cg == 2 && dbg == 5) return tc_launch_one<EPI_PLAIN, 5, 0, 2>(grid, st, p, t, nt, (int)items);
  if (cg == 2 && dbg == 8) return tc_launch_one<EPI_PLAIN, 8, 0, 2>(grid, st, p, t, nt, (int)items);
  if (cg == 2 && dbg == 9) return tc_launch_one<EPI_PLAIN, 9, 0, 2>(grid, st, p, t, nt, (int)items);
#endif
  if (dbg) {
    if (epi != EPI_PLAIN) return (int)cudaErrorInvalidValue;
    switch (dbg) {
      case 1: return tc_launch_one<EPI_PLAIN, 1>(grid, st, p, t, nt, (int)items);
      case 2: return tc_launch_one<EPI_PLAIN, 2>(grid, st, p, t, nt, (int)items);
      case 3: return tc_launch_one<EPI_PLAIN, 3>(grid, st, p, t, nt, (int)items);
      case 4: return tc_launch_one<EPI_PLAIN, 4>(grid, st, p, t, nt, (int)items);
      case 5: return tc_launch_one<EPI_PLAIN, 5>(grid, st, p, t, nt, (int)items);
      case 6: return tc_launch_one<EPI_PLAIN, 6>(grid, st, p, t, nt, (int)items);
      case 7: return tc_launch_one<EPI_PLAIN, 7>(grid, st, p, t, nt, (int)items);
      case 8: return tc_launch_one<EPI_PLAIN, 8>(grid, st, p, t, nt, (int)items);
      case 9: return tc_launch_one<EPI_PLAIN, 9>(grid, st, p, t, nt, (int)items);
      default: return (int)cudaErrorInvalidValue;
    }
  }
  if (epi == EPI_PLAIN) return tc_launch_one<EPI_PLAIN, 0>(grid, st, p, t, nt, (int)items);
  if (epi == EPI_SEMCH) return tc_launch_one<EPI_SEMCH, 0>(grid, st, p, t, nt, (int)items);
  return tc_launch_one<EPI_GLOBAL, 0>(grid, st, p, t, nt, (int)items);
}

}  // namespace gast
