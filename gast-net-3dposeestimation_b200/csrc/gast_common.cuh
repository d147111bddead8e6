// Shared device-side types of libgast_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gast {

constexpr float BN_EPS = 1e-5f;  // nn.BatchNorm2d default used throughout the reference

// Frame remap of a gathered operand: output frame f = b*T_out + t  ->  input frame
// b*T_in + t*t_mul + t_off.  Identity is {1,1,1,0} on flattened frames.
struct RowMap {
  int T_out, T_in, t_mul, t_off;
};

__host__ __device__ inline long long map_frame(const RowMap& m, long long f) {
  long long b = f / m.T_out;
  long long t = f - b * m.T_out;
  return b * m.T_in + t * m.t_mul + m.t_off;
}

// One K-segment of the A operand: rows gathered from `base` (channels-last, row stride ld).
// K may span several temporal taps: k -> (tap = k / Kc, c = k % Kc), the tap adds
// tap * tap_stride floats (dilation * J * ld) to the row address.
struct ASeg {
  const float* base;
  int ld;
  int K;
  int Kc;
  long long tap_stride;
  RowMap map;
};

// Sparse row structure of a LocalGraph mask (J <= 32, nnz <= 160), row-major nonzero order
// == the order of the learnable `e` (local_attention.py:25,41).
struct NbrTable {
  unsigned char row_ptr[36];
  unsigned char col[164];
};

struct GemmP {
  ASeg seg[3];
  int nseg;
  const float* W;   // [N][ldw], K-major
  int ldw;
  int N;
  int F;            // output frames
  int J;
  int fpt;          // frames per 128-row tile
  float* out;
  int ld_out;
  const float* bias;
  int relu;
  const float* res; // optional residual, added after the activation (gast_net.py:174)
  int res_ld;
  RowMap res_map;
  // EPI_SEMCH
  const float* coef[2];   // [nnz][C], softmaxed adjacency x BN scale
  const float* shift;     // [2][C] BN shift (or SemGraphConv bias), may be null
  int C;
  int tiles_per_mask;
  NbrTable nbr[2];
  // EPI_GLOBAL
  const float* ab;  // [rows][2*heads]: a_h(i), b_h(j)
  const float* ck;  // [heads][J][J]
  const float* bg;  // [N]
  int heads;
  int Cg;
  int tc_mode;                 // unused by the product path
  // tcgen05 core, A operand by TMA (set by tc_launch when every segment's frame map is affine):
  int a_tma;                   // 1 = raw A tiles arrive by cp.async.bulk.tensor.3d, 0 = cp.async gather
  int a_map0[3];               // first tensor-map index of each segment (one map per temporal tap)
  unsigned long long* dbg;     // DBG==6 timing variant of the tcgen05 kernel: per-CTA cycle counters
};

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace gast
