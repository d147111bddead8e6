// HBM-bound kernels of the lifting path in their streaming form (round 2): the expand stage staged through shared
// memory and the attention mix of MultiGlobalGraph as a persistent kernel fed by bulk copies (cp.async.bulk +
// mbarrier).  Both replace latency-bound "load, then compute, then store" kernels of kernels_misc.cuh, which stay as
// the fallback for shapes these do not cover (GAST_EXPAND_STAGED=0 / GAST_MIX_STREAM=0 select them explicitly).
#pragma once
#include "gast_common.cuh"
#include "kernels_misc.cuh"
#include "gemm_tc.cuh"      // mbarrier / shared-address PTX wrappers

namespace gast {

// ---------------------------------------------------------------------------------------
// expand stage (gast_net.py:163-164), staged: a block owns EXS_ROWS consecutive output rows.  Their taps*Fin input
// values are gathered ONCE into shared memory (expand_kernel has every one of the C/4 threads of a row request them
// again through L1: 48 loads in flight per thread, 126 registers, 16 warps per SM, 1.7 TB/s of output); a thread then
// keeps the folded weights of its 4 channels in registers and walks the block's rows: 2 broadcast shared-memory reads,
// 4*KF FMAs and one coalesced 16-byte store per row, ~50 registers, full occupancy.  Same FMA order as expand_kernel
// (bias first, then k ascending), so the two are bit-identical.
// ---------------------------------------------------------------------------------------
constexpr int EXS_ROWS = 128;
constexpr int EXS_THREADS = 256;
template <int KFT>
__global__ void __launch_bounds__(EXS_THREADS)
expand_rows_kernel(const float* __restrict__ x, const float* __restrict__ We, const float* __restrict__ be,
                   float* __restrict__ out, long long rows, int J, int T, int T0, int stride, int taps, int Fin, int C) {
  constexpr int KP = (KFT + 3) & ~3;                      // padded row of staged inputs (float4 reads)
  __shared__ __align__(16) float xs[EXS_ROWS * KP];
  const int tid = threadIdx.x;
  const int KF = taps * Fin;
  const long long r0 = (long long)blockIdx.x * EXS_ROWS;
  const int nrows = (int)((rows - r0 < EXS_ROWS) ? (rows - r0) : EXS_ROWS);
  for (int e = tid; e < EXS_ROWS * KP; e += EXS_THREADS) {
    const int rr = e / KP, k = e - rr * KP;
    float v = 0.f;
    if (rr < nrows && k < KF) {
      const int row = (int)(r0 + rr);                     // the host checks rows < 2^31
      const int f = row / J, j = row - f * J;
      const int b = f / T0, t = f - b * T0;
      const int kk = k / Fin, i = k - kk * Fin;
      v = __ldg(x + (((long long)b * T + (long long)t * stride + kk) * J + j) * Fin + i);
    }
    xs[e] = v;
  }
  const int cq = C >> 2;                                  // channel groups (threads) per row
  const int rpp = EXS_THREADS / cq;                       // rows per pass of the block
  const int cg = tid % cq, rsub = tid / cq;
  const int c = cg * 4;
  float w[4][KFT];
#pragma unroll
  for (int k = 0; k < KFT; ++k) {
    float4 wk = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < KF) wk = ldg4(We + (long long)k * C + c);
    w[0][k] = wk.x; w[1][k] = wk.y; w[2][k] = wk.z; w[3][k] = wk.w;
  }
  const float4 b4 = ldg4(be + c);
  __syncthreads();
  if (rsub >= rpp) return;                                // (256 % cq != 0: the surplus threads have no row)
  float* op = out + (r0 + rsub) * C + c;
  const long long ostep = (long long)rpp * C;
#pragma unroll 4
  for (int rr = rsub; rr < nrows; rr += rpp, op += ostep) {
    float xv[KP];
#pragma unroll
    for (int q = 0; q < KP / 4; ++q) {
      const float4 t4 = *reinterpret_cast<const float4*>(xs + rr * KP + q * 4);
      xv[q * 4] = t4.x; xv[q * 4 + 1] = t4.y; xv[q * 4 + 2] = t4.z; xv[q * 4 + 3] = t4.w;
    }
    float v0 = b4.x, v1 = b4.y, v2 = b4.z, v3 = b4.w;
#pragma unroll
    for (int k = 0; k < KFT; ++k) {
      v0 = fmaf(w[0][k], xv[k], v0); v1 = fmaf(w[1][k], xv[k], v1);
      v2 = fmaf(w[2][k], xv[k], v2); v3 = fmaf(w[3][k], xv[k], v3);
    }
    *reinterpret_cast<float4*>(op) = make_float4(fmaxf(v0, 0.f), fmaxf(v1, 0.f), fmaxf(v2, 0.f), fmaxf(v3, 0.f));
  }
}

// true when expand_rows_kernel covers the shape (else the caller keeps expand_kernel)
static inline bool expand_rows_ok(int C) { return C % 4 == 0 && C / 4 >= 1 && C / 4 <= EXS_THREADS; }

// ---------------------------------------------------------------------------------------
// Attention mix of MultiGlobalGraph (global_attention.py:74-80), streaming form.  Same arithmetic and the same work
// split as global_mix_kernel (one thread = one frame x 4 channels, its J float4 of g in registers, the attention rows
// of the block's frames in shared memory), but PERSISTENT and fed by bulk copies: the g rows and the a/b values of a
// group of `fpb` frames are one contiguous slab each (ldg == heads*Cg), fetched by cp.async.bulk into a ring of STAGES
// shared-memory stages and signalled on an mbarrier.  The slab of group n+1 (and n+2) is in flight while group n is
// computed, and a stage is re-armed as soon as its g values sit in registers, so neither the a/b loads before the
// softmax nor the g loads after the block barrier are exposed (global_mix_kernel: 43 % of the HBM rate, 16 % of its
// stall samples on those loads and 11 % on the barrier between its two phases, profiles/r02_w_lines_global_mix.txt).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

constexpr int MIXS_THREADS = 128;
constexpr int MIXS_HDR = 64;      // bytes reserved for the stage barriers

// shared memory: [barriers | attention rows fpb*heads*J*MIX_JP | STAGES x a/b slab | STAGES x g slab]
static inline size_t mix_stream_smem(int fpb, int J, int heads, int Ng, int stages) {
  const size_t abf = ((size_t)fpb * J * 2 * heads + 3) & ~(size_t)3;
  return MIXS_HDR + sizeof(float) * ((size_t)fpb * heads * J * MIX_JP + stages * (abf + (size_t)fpb * J * Ng));
}

template <int STAGES>
__global__ void __launch_bounds__(MIXS_THREADS)
global_mix_stream_kernel(const float* __restrict__ G, const float* __restrict__ ab, const float* __restrict__ ck,
                         float* __restrict__ Y, long long F, int J, int heads, int Cg, int fpb) {
  extern __shared__ __align__(128) unsigned char mixs_smem[];
  constexpr int NT = MIXS_THREADS;
  const int tid = threadIdx.x;
  const int H2 = 2 * heads, Ng = heads * Cg;
  const int abf = (fpb * J * H2 + 3) & ~3;                 // floats per a/b stage
  const int gf = fpb * J * Ng;                             // floats per g stage
  float* att_s = reinterpret_cast<float*>(mixs_smem + MIXS_HDR);
  float* ab_s = att_s + fpb * heads * J * MIX_JP;
  float* g_s = ab_s + STAGES * abf;
  const uint32_t bar0 = smem_u32(mixs_smem);
  const long long ngroups = (F + fpb - 1) / fpb;

  auto issue = [&](long long grp, int s) {                 // one thread: arm the stage and start both copies
    const long long f0 = grp * fpb;
    const int nf = (int)((F - f0 < fpb) ? (F - f0) : fpb);
    const uint32_t gb = (uint32_t)nf * J * Ng * 4u, abb = (uint32_t)nf * J * H2 * 4u;
    mbar_arrive_expect_tx(bar0 + 8 * s, gb + abb);
    bulk_g2s(smem_u32(g_s + (size_t)s * gf), G + f0 * J * (long long)Ng, gb, bar0 + 8 * s);
    bulk_g2s(smem_u32(ab_s + (size_t)s * abf), ab + f0 * J * (long long)H2, abb, bar0 + 8 * s);
  };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_mbar_init();
    for (int s = 0; s < STAGES; ++s) {
      const long long grp = blockIdx.x + (long long)s * gridDim.x;
      if (grp < ngroups) issue(grp, s);
    }
  }
  __syncthreads();

  const int GV = Ng / 4;                                   // threads per frame (host: fpb * GV == NT)
  const int fs = tid / GV, c = (tid - fs * GV) * 4;
  const int h = c / Cg;
  const int nrow = fpb * heads * J;
  int it = 0;
  for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, ++it) {
    const int s = it % STAGES;
    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
    const long long f0 = grp * fpb;
    const int nf = (int)((F - f0 < fpb) ? (F - f0) : fpb);
    mbar_wait(bar0 + 8 * s, ph);
    // ---- attention rows of this group's frames (a/b from the staged slab)
    const float* abq = ab_s + (size_t)s * abf;
    for (int e = tid; e < nrow; e += NT) {
      const int i = e % J, hh = (e / J) % heads, fr = e / (J * heads);
      float* dst = att_s + (size_t)e * MIX_JP;
      if (fr < nf) {
        const float* abf_ = abq + (fr * J) * H2;
        const float a = abf_[i * H2 + 2 * hh];
        float v[MIX_JMAX];
        float mx = -3.4e38f;
#pragma unroll
        for (int j = 0; j < MIX_JMAX; ++j)
          if (j < J) {
            float sc = a + abf_[j * H2 + 2 * hh + 1];
            sc = (sc >= 0.f) ? sc : 0.2f * sc;
            v[j] = sc;
            mx = fmaxf(mx, sc);
          }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < MIX_JMAX; ++j)
          if (j < J) { v[j] = expf(v[j] - mx); sum += v[j]; }
        const float inv = 1.f / sum;
        const float* ckr = ck + ((long long)hh * J + i) * J;
#pragma unroll
        for (int j = 0; j < MIX_JMAX; ++j) dst[j] = (j < J) ? v[j] * inv + __ldg(ckr + j) : 0.f;
      }
    }
    // ---- this thread's g values: shared memory -> registers
    float4 g[MIX_JMAX];
    const float* gq = g_s + (size_t)s * gf + (size_t)(fs * J) * Ng + c;
    if (fs < nf) {
#pragma unroll
      for (int j = 0; j < MIX_JMAX; ++j)
        if (j < J) g[j] = *reinterpret_cast<const float4*>(gq + (size_t)j * Ng);
    }
    __syncthreads();                                       // attention rows complete; stage s fully read
    if (tid == 0) {
      const long long nxt = grp + (long long)STAGES * gridDim.x;
      if (nxt < ngroups) issue(nxt, s);
    }
    // ---- mix
    if (fs < nf) {
      const float* arow = att_s + (size_t)((fs * heads + h) * J) * MIX_JP;
      float* yp = Y + ((f0 + fs) * J) * (long long)Ng + c;
      for (int i = 0; i < J; ++i) {
        float a[MIX_JMAX];
#pragma unroll
        for (int q = 0; q < MIX_JMAX / 4; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(arow + i * MIX_JP + q * 4);
          a[q * 4] = t.x; a[q * 4 + 1] = t.y; a[q * 4 + 2] = t.z; a[q * 4 + 3] = t.w;
        }
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < MIX_JMAX; ++j)
          if (j < J) MixVec<float4>::fma(a[j], g[j], o);
        *reinterpret_cast<float4*>(yp + (long long)i * Ng) = o;
      }
    }
    __syncthreads();                                       // the attention rows are rewritten by the next group
  }
}

// ---------------------------------------------------------------------------------------
// Collapsed theta/phi row dots (global_attention.py:60-72; rowdot8_kernel of kernels_misc.cuh), streaming form.
// rowdot8_kernel keeps 2 rows (1 KB at C = 128) in flight per warp and ~30 warps per SM: 30 KB against ~3000 cycles of
// load latency under load = 3 TB/s, 35 % of its issue slots.  Here every WARP runs its own bulk-copy pipeline: a ring of
// RDS_STAGES chunks of consecutive rows (2-4 KB, ldx == K) in shared memory, each signalled on the warp's own mbarrier;
// lane 0 re-arms a stage as soon as the warp has consumed it.  The bytes in flight (2 chunks per warp, 32 warps per SM)
// no longer depend on registers, and there is no block-wide barrier in the loop.  Same FMA order and the same halving
// butterfly as rowdot8_kernel: bit-identical results.
// ---------------------------------------------------------------------------------------
constexpr int RDS_THREADS = 256;
constexpr int RDS_WARPS = RDS_THREADS / 32;
constexpr int RDS_STAGES = 3;
constexpr int RDS_HDR = 256;      // RDS_WARPS x RDS_STAGES barriers

static inline int rowdot_chunk_bytes(int K) { return (2 * K * 4 > 2048) ? 2 * K * 4 : 2048; }
static inline size_t rowdot_stream_smem(int K, int nst) {
  return RDS_HDR + sizeof(float) * 8 * (size_t)K + (size_t)RDS_WARPS * nst * rowdot_chunk_bytes(K);
}

__global__ void __launch_bounds__(RDS_THREADS, 4)
rowdot8_stream_kernel(const float* __restrict__ X, const float* __restrict__ U, const float* __restrict__ cab,
                      float* __restrict__ ab, long long rows, int K, int chunk_bytes, int nst) {
  extern __shared__ __align__(128) unsigned char rds_smem[];
  float* us = reinterpret_cast<float*>(rds_smem + RDS_HDR);             // [8][K]
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int cf = chunk_bytes / 4;                                       // floats per chunk
  float* xs = us + 8 * K + (size_t)wrp * nst * cf;                      // this warp's ring of nst (<= RDS_STAGES) chunks
  const uint32_t bar0 = smem_u32(rds_smem) + 8 * RDS_STAGES * wrp;
  const int rpc = cf / K;                                               // rows per chunk
  const long long nchunks = (rows + rpc - 1) / rpc;
  const long long gw = (long long)blockIdx.x * RDS_WARPS + wrp, GW = (long long)gridDim.x * RDS_WARPS;

  auto issue = [&](long long ci, int s) {
    const long long r0 = ci * rpc;
    const int nr = (int)((rows - r0 < rpc) ? (rows - r0) : rpc);
    const uint32_t bytes = (uint32_t)nr * K * 4u;
    mbar_arrive_expect_tx(bar0 + 8 * s, bytes);
    bulk_g2s(smem_u32(xs + (size_t)s * cf), X + r0 * K, bytes, bar0 + 8 * s);
  };
  if (lane == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_mbar_init();
    for (int s = 0; s < nst; ++s) {
      const long long ci = gw + (long long)s * GW;
      if (ci < nchunks) issue(ci, s);
    }
  }
  for (int i = tid * 4; i < 8 * K; i += RDS_THREADS * 4) *reinterpret_cast<float4*>(us + i) = ldg4(U + i);
  __syncthreads();

  const int sel = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);   // q owned after the butterfly
  const float cq = __ldg(cab + sel);
  const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
  int it = 0;
  for (long long ci = gw; ci < nchunks; ci += GW, ++it) {
    const int s = it % nst;
    const uint32_t ph = (uint32_t)(it / nst) & 1u;
    const long long r0 = ci * rpc;
    const int nr = (int)((rows - r0 < rpc) ? (rows - r0) : rpc);
    const float* slab = xs + (size_t)s * cf;
    mbar_wait(bar0 + 8 * s, ph);
    for (int ra = 0; ra < nr; ra += 2) {
      const int rb = ra + 1;
      const bool has1 = rb < nr;
      const float* xr0 = slab + (size_t)ra * K;
      const float* xr1 = slab + (size_t)(has1 ? rb : ra) * K;
      float va[8], vb[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { va[q] = 0.f; vb[q] = 0.f; }
      for (int k = lane * 4; k < K; k += 128) {
        const float4 xa = *reinterpret_cast<const float4*>(xr0 + k);
        const float4 xb = *reinterpret_cast<const float4*>(xr1 + k);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 u = *reinterpret_cast<const float4*>(us + q * K + k);
          va[q] = fmaf(xa.x, u.x, fmaf(xa.y, u.y, fmaf(xa.z, u.z, fmaf(xa.w, u.w, va[q]))));
          vb[q] = fmaf(xb.x, u.x, fmaf(xb.y, u.y, fmaf(xb.z, u.z, fmaf(xb.w, u.w, vb[q]))));
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float* v = r ? vb : va;
        float w4[4], w2[2], w1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float keep = hi16 ? v[i + 4] : v[i], send = hi16 ? v[i] : v[i + 4];
          w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float keep = hi8 ? w4[i + 2] : w4[i], send = hi8 ? w4[i] : w4[i + 2];
          w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
          const float keep = hi4 ? w2[1] : w2[0], send = hi4 ? w2[0] : w2[1];
          w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
        w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
        if ((lane & 3) == 0 && (r == 0 || has1)) ab[(r0 + (r ? rb : ra)) * 8 + sel] = w1 + cq;
      }
    }
    __syncwarp();                                          // every lane is done with stage s
    if (lane == 0) {
      const long long nxt = ci + (long long)nst * GW;
      if (nxt < nchunks) issue(nxt, s);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Attention mix, tiled form: the g slab stays in SHARED memory (bulk-copied, as above) and a thread owns 4 channels x
// IW output joints -- per source joint j one 16-byte read of g (conflict-free), one 16-byte read of the IW attention
// values (stored [j][i], broadcast within a head) and 4*IW FMAs.  No g values in registers: ~50 registers instead of
// 120-165, so 24-36 warps per SM instead of 4-16 hide the shared-memory and FMA latencies that bound both
// global_mix_kernel and global_mix_stream_kernel (measured, profiles/r02_z_*: the register-resident stream kernel is
// 0-35 % SLOWER than global_mix_kernel at 2 stages, equal at 3).  Accumulation order over j is the same, so the result is
// bit-identical to the other two.
// ---------------------------------------------------------------------------------------
constexpr int MIXT_MAXTHREADS = 384;
#ifndef GAST_MIXT_MINB
#define GAST_MIXT_MINB 2
#endif
static inline int mixt_nig(int J, int IW) { return (J + IW - 1) / IW; }
static inline size_t mix_tile_smem(int fpb, int J, int heads, int Ng, int IW, int stages) {
  const size_t abf = ((size_t)fpb * J * 2 * heads + 3) & ~(size_t)3;
  return MIXS_HDR + sizeof(float) * ((size_t)fpb * heads * J * mixt_nig(J, IW) * 4 + stages * (abf + (size_t)fpb * J * Ng));
}

template <int IW, int STAGES>
__global__ void __launch_bounds__(MIXT_MAXTHREADS, GAST_MIXT_MINB)
global_mix_tile_kernel(const float* __restrict__ G, const float* __restrict__ ab, const float* __restrict__ ck,
                       float* __restrict__ Y, long long F, int J, int heads, int Cg, int fpb) {
  extern __shared__ __align__(128) unsigned char mixt_smem[];
  const int NT = blockDim.x, tid = threadIdx.x;
  const int H2 = 2 * heads, Ng = heads * Cg;
  const int NIG = (J + IW - 1) / IW, JP = NIG * 4;         // attention row [j] : NIG groups of 4 slots (IW used)
  const int abf = (fpb * J * H2 + 3) & ~3;
  const int gf = fpb * J * Ng;
  float* att_s = reinterpret_cast<float*>(mixt_smem + MIXS_HDR);     // [fpb][heads][J (j)][JP (i slots)]
  float* ab_s = att_s + fpb * heads * J * JP;
  float* g_s = ab_s + STAGES * abf;
  const uint32_t bar0 = smem_u32(mixt_smem);
  const long long ngroups = (F + fpb - 1) / fpb;

  auto issue = [&](long long grp, int s) {
    const long long f0 = grp * fpb;
    const int nf = (int)((F - f0 < fpb) ? (F - f0) : fpb);
    const uint32_t gb = (uint32_t)nf * J * Ng * 4u, abb = (uint32_t)nf * J * H2 * 4u;
    mbar_arrive_expect_tx(bar0 + 8 * s, gb + abb);
    bulk_g2s(smem_u32(g_s + (size_t)s * gf), G + f0 * J * (long long)Ng, gb, bar0 + 8 * s);
    bulk_g2s(smem_u32(ab_s + (size_t)s * abf), ab + f0 * J * (long long)H2, abb, bar0 + 8 * s);
  };
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_mbar_init();
    for (int s = 0; s < STAGES; ++s) {
      const long long grp = blockIdx.x + (long long)s * gridDim.x;
      if (grp < ngroups) issue(grp, s);
    }
  }
  for (int e = tid; e < fpb * heads * J * JP; e += NT) att_s[e] = 0.f;   // unused slots stay finite
  __syncthreads();

  const int GV = Ng / 4;
  const int items = fpb * NIG * GV;                        // item = (frame slot, joint group, channel quad)
  const int nrow = fpb * heads * J;
  int it = 0;
  for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, ++it) {
    const int s = it % STAGES;
    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
    const long long f0 = grp * fpb;
    const int nf = (int)((F - f0 < fpb) ? (F - f0) : fpb);
    mbar_wait(bar0 + 8 * s, ph);
    // ---- attention rows (softmax over j of row i), stored transposed: att_s[fs][h][j][slot(i)]
    const float* abq = ab_s + (size_t)s * abf;
    for (int e = tid; e < nrow; e += NT) {
      const int i = e % J, hh = (e / J) % heads, fr = e / (J * heads);
      if (fr >= nf) continue;
      const float* abr = abq + (fr * J) * H2;
      const float a = abr[i * H2 + 2 * hh];
      float v[MIX_JMAX];
      float mx = -3.4e38f;
#pragma unroll
      for (int j = 0; j < MIX_JMAX; ++j)
        if (j < J) {
          float sc = a + abr[j * H2 + 2 * hh + 1];
          sc = (sc >= 0.f) ? sc : 0.2f * sc;
          v[j] = sc;
          mx = fmaxf(mx, sc);
        }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < MIX_JMAX; ++j)
        if (j < J) { v[j] = expf(v[j] - mx); sum += v[j]; }
      const float inv = 1.f / sum;
      const float* ckr = ck + ((long long)hh * J + i) * J;
      float* dst = att_s + (size_t)((fr * heads + hh) * J) * JP + (i / IW) * 4 + (i % IW);
#pragma unroll
      for (int j = 0; j < MIX_JMAX; ++j)
        if (j < J) dst[j * JP] = v[j] * inv + __ldg(ckr + j);
    }
    __syncthreads();
    // ---- mix
    const float* gq = g_s + (size_t)s * gf;
    for (int w = tid; w < items; w += NT) {
      const int cq = w % GV, t = w / GV;
      const int ig = t % NIG, fs = t / NIG;
      if (fs >= nf) continue;
      const int c = cq * 4, h = c / Cg;
      const float* gp = gq + (size_t)(fs * J) * Ng + c;
      const float* ap = att_s + (size_t)((fs * heads + h) * J) * JP + ig * 4;
      float4 o[IW];
#pragma unroll
      for (int u = 0; u < IW; ++u) o[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int j = 0; j < J; ++j) {
        const float4 g = *reinterpret_cast<const float4*>(gp + (size_t)j * Ng);
        const float4 a = *reinterpret_cast<const float4*>(ap + j * JP);
        MixVec<float4>::fma(a.x, g, o[0]);
        if (IW > 1) MixVec<float4>::fma(a.y, g, o[1]);
        if (IW > 2) MixVec<float4>::fma(a.z, g, o[2]);
        if (IW > 3) MixVec<float4>::fma(a.w, g, o[3]);
      }
      float* yp = Y + ((f0 + fs) * J + ig * IW) * (long long)Ng + c;
#pragma unroll
      for (int u = 0; u < IW; ++u)
        if (ig * IW + u < J) *reinterpret_cast<float4*>(yp + (long long)u * Ng) = o[u];
    }
    __syncthreads();                                       // slab s and the attention rows are free
    if (tid == 0) {
      const long long nxt = grp + (long long)STAGES * gridDim.x;
      if (nxt < ngroups) issue(nxt, s);
    }
  }
}

// ---------------------------------------------------------------------------------------
// shrink (gast_net.py:60,99) with mpjpe (common/loss.py:5-11) in its epilogue: the warp that produced the 3 coordinates
// of an output joint also takes ||y - target|| of it.  Partial sums are double (like mpjpe_kernel); the last block to
// finish (atomic ticket) adds the block partials in index order, so the loss does not depend on the scheduling, and
// resets the ticket for the next call (CUDA-graph replay safe).  y is written as by shrink_kernel (same FMA order).
// ---------------------------------------------------------------------------------------
constexpr int SHL_MAXBLOCKS = 1024;
__global__ void __launch_bounds__(256)
shrink_mpjpe_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ Ws, float* __restrict__ y,
                    long long rows, int K, const float* __restrict__ target, double* __restrict__ partial,
                    unsigned* __restrict__ ticket, float* __restrict__ loss) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const long long nw = (long long)gridDim.x * 8;
  double acc = 0.0;
  for (long long row = (long long)blockIdx.x * 8 + wrp; row < rows; row += nw) {
    const float* xr = X + row * ldx;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      float4 xv = ldg4(xr + k);
      float4 w0 = ldg4(Ws + k), w1 = ldg4(Ws + K + k), w2 = ldg4(Ws + 2 * K + k);
      a0 = fmaf(xv.x, w0.x, fmaf(xv.y, w0.y, fmaf(xv.z, w0.z, fmaf(xv.w, w0.w, a0))));
      a1 = fmaf(xv.x, w1.x, fmaf(xv.y, w1.y, fmaf(xv.z, w1.z, fmaf(xv.w, w1.w, a1))));
      a2 = fmaf(xv.x, w2.x, fmaf(xv.y, w2.y, fmaf(xv.z, w2.z, fmaf(xv.w, w2.w, a2))));
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) {
      y[row * 3 + 0] = a0; y[row * 3 + 1] = a1; y[row * 3 + 2] = a2;
      const float d0 = a0 - target[row * 3], d1 = a1 - target[row * 3 + 1], d2 = a2 - target[row * 3 + 2];
      acc += (double)sqrtf(fmaf(d2, d2, fmaf(d1, d1, d0 * d0)));       // same expression as mpjpe_kernel (D = 3)
    }
  }
  __shared__ double sh[8];
  __shared__ bool last;
  if (lane == 0) sh[wrp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int i = 0; i < 8; ++i) v += sh[i];
    partial[blockIdx.x] = v;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    double v = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) v += __ldcg(partial + i);
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) { *loss = (float)(v / (double)rows); *ticket = 0u; }
  }
}

}  // namespace gast
