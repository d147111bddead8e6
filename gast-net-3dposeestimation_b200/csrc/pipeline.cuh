// The callers and data formats either side of the lifting forward, kept on the
// device: training-batch assembly (ChunkedGenerator), keypoint format conversion and screen
// normalisation in front of the network; camera-to-world, MPJPE / P-MPJPE and the optimiser step
// behind it.  All are HBM-/latency-bound elementwise or tiny-reduction kernels; what they buy is
// that a 200k clips/s forward is not fed and drained by numpy on the host.
//
// Included at the end of gast_api.cu (uses fail / CUDA_OK / cdiv / JointPerm / make_perm).
#pragma once
#include <algorithm>
#include <math.h>

namespace gast {

// ------------------------------------------------------------------------------------------
// N1: ChunkedGenerator.next_epoch (common/generators.py:93-154): one training batch gathered from
// the concatenated sequences.  pairs[b] = {seq, start_3d, end_3d (unused), flip}.
//   batch_2d[b,t] = edge-clamped frame (start_3d - pad - causal_shift + t) of sequence seq, mirrored
//   when flip (feature 0 negated, left/right keypoints swapped); batch_3d likewise without padding;
//   batch_cam[b] = cameras[seq] with entries 2 and 7 negated when flip (:139-144).
// ------------------------------------------------------------------------------------------
__global__ void chunk_gather_kernel(const float* __restrict__ poses, const long long* __restrict__ seq_start,
                                    const int* __restrict__ pairs, float* __restrict__ out, int B, int Tc, int J, int F,
                                    int lead, JointPerm perm) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)B * Tc * J * F;
  if (idx >= n) return;
  const int c = (int)(idx % F);
  const int j = (int)((idx / F) % J);
  const int t = (int)((idx / ((long long)F * J)) % Tc);
  const int b = (int)(idx / ((long long)F * J * Tc));
  const int seq = pairs[4 * b], start = pairs[4 * b + 1], flip = pairs[4 * b + 3];
  const long long s0 = seq_start[seq], len = seq_start[seq + 1] - s0;
  long long fr = (long long)start - lead + t;
  fr = fr < 0 ? 0 : (fr > len - 1 ? len - 1 : fr);              // np.pad(..., 'edge')
  const int js = flip ? perm.p[j] : j;
  float v = poses[((s0 + fr) * J + js) * F + c];
  out[idx] = (flip && c == 0) ? -v : v;
}

__global__ void chunk_cam_kernel(const float* __restrict__ cams, const int* __restrict__ pairs, float* __restrict__ out,
                                 int B, int ncam) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * ncam) return;
  const int b = idx / ncam, k = idx - b * ncam;
  float v = cams[(long long)pairs[4 * b] * ncam + k];
  if (pairs[4 * b + 3] && (k == 2 || k == 7)) v = -v;
  out[idx] = v;
}

// ------------------------------------------------------------------------------------------
// N3: keypoint formats (tools/mpii_coco_h36m.py).  One thread per frame; float32 arithmetic in the
// order numpy evaluates it (sequential sums, true division), so results are bit-identical for
// float32 input.  valid[t] = (sum of the frame's coordinates != 0), the mask np.where is taken of.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float np_pairwise_sum(const float* a, int n) {   // numpy's float32 add.reduce, 8 <= n < 128
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = a[k];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], a[i + k]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}

// coco_h36m (tools/mpii_coco_h36m.py:20-48): kp (17 x 2) of one frame -> h (17 x 2)
__device__ __forceinline__ void coco_to_h36m_frame(const float* kp, int ld, float* h) {
  auto K = [&](int j, int c) { return kp[j * ld + c]; };
  const int h36m_coco_order[13] = {9, 11, 14, 12, 15, 13, 16, 4, 1, 5, 2, 6, 3};
  const int coco_order[13] = {0, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
  float htps[4][2];
  htps[0][0] = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(K(1, 0), K(2, 0)), K(3, 0)), K(4, 0)), 4.f);   // head x :25
  htps[0][1] = __fsub_rn(__fadd_rn(K(1, 1), K(2, 1)), K(0, 1));                                      // head y :26
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float th = __fdiv_rn(__fadd_rn(K(5, c), K(6, c)), 2.f);                                          // thorax :27
    th = __fadd_rn(th, __fdiv_rn(__fsub_rn(K(0, c), th), 3.f));                                      // :28
    htps[1][c] = th;
    htps[2][c] = __fdiv_rn(__fadd_rn(K(11, c), K(12, c)), 2.f);                                      // pelvis :30
    htps[3][c] = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(K(5, c), K(6, c)), K(11, c)), K(12, c)), 4.f);   // spine :31
  }
  const int spple[4] = {10, 8, 0, 7};
#pragma unroll
  for (int i = 0; i < 4; ++i) { h[spple[i] * 2] = htps[i][0]; h[spple[i] * 2 + 1] = htps[i][1]; }
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    h[h36m_coco_order[i] * 2] = K(coco_order[i], 0);
    h[h36m_coco_order[i] * 2 + 1] = K(coco_order[i], 1);
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {                                                                       // neck :36
    const float m = __fdiv_rn(__fadd_rn(K(5, c), K(6, c)), 2.f);
    h[9 * 2 + c] = __fsub_rn(h[9 * 2 + c], __fdiv_rn(__fsub_rn(h[9 * 2 + c], m), 4.f));
  }
  {                                                                                                   // spine x :37
    const float m = __fdiv_rn(__fadd_rn(h[0], h[8 * 2]), 2.f);
    h[7 * 2] = __fadd_rn(h[7 * 2], __fmul_rn(2.f, __fsub_rn(h[7 * 2], m)));
  }
  {                                                                                                   // thorax y :38
    const float m = __fdiv_rn(__fadd_rn(K(1, 1), K(2, 1)), 2.f);
    h[8 * 2 + 1] = __fsub_rn(h[8 * 2 + 1], __fdiv_rn(__fmul_rn(__fsub_rn(m, K(0, 1)), 2.f), 3.f));
  }
}

// mode 0: coco_h36m            kp (T,17,2)  -> out (T,17,2)
// mode 1: mpii_h36m            kp (T,16,2)  -> out (T,17,2)   (:51-59)
// mode 2: coco_h36m_toe_format kp (T,Jin>=22,2) -> out (T,19,2)   (:62-78)
__global__ void kpt_convert_kernel(const float* __restrict__ kp, float* __restrict__ out, int* __restrict__ valid,
                                   int T, int Jin, int mode) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* k = kp + (long long)t * Jin * 2;
  float o[38];
  int Jo = 17;
  if (mode == 0) {
    coco_to_h36m_frame(k, 2, o);
  } else if (mode == 1) {
    const int h36m_mpii_order[16] = {3, 2, 1, 4, 5, 6, 0, 8, 9, 10, 16, 15, 14, 11, 12, 13};
#pragma unroll
    for (int i = 0; i < 34; ++i) o[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[h36m_mpii_order[i] * 2] = k[i * 2]; o[h36m_mpii_order[i] * 2 + 1] = k[i * 2 + 1]; }
#pragma unroll
    for (int c = 0; c < 2; ++c)
      o[7 * 2 + c] = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(k[2 * 2 + c], k[3 * 2 + c]), k[12 * 2 + c]), k[13 * 2 + c]), 4.f);
  } else {
    Jo = 19;
    float h[34];
    coco_to_h36m_frame(k, 2, h);
    const int toe_order[17] = {0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18};
#pragma unroll
    for (int i = 0; i < 38; ++i) o[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 17; ++i) { o[toe_order[i] * 2] = h[i * 2]; o[toe_order[i] * 2 + 1] = h[i * 2 + 1]; }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      o[4 * 2 + c] = __fdiv_rn(__fadd_rn(k[20 * 2 + c], k[21 * 2 + c]), 2.f);
      o[8 * 2 + c] = __fdiv_rn(__fadd_rn(k[17 * 2 + c], k[18 * 2 + c]), 2.f);
    }
  }
  float* op = out + (long long)t * Jo * 2;
  for (int i = 0; i < Jo * 2; ++i) op[i] = o[i];
  if (valid) valid[t] = (np_pairwise_sum(o, Jo * 2) != 0.f) ? 1 : 0;
}

// normalize_screen_coordinates (common/camera.py:8-12): X/w*2 - [1, h/w].  numpy evaluates X/w*2 in
// float32 and the subtraction of the Python-float list in float64; the result is rounded to float32
// here (callers cast it: reconstruction.py:143, main.py:46).   image_coordinates (:15-19) is the inverse.
__global__ void screen_norm_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, float w, double hw,
                                   int inverse) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const double c = (idx & 1) ? hw : 1.0;
  if (!inverse) {
    const float s = __fmul_rn(__fdiv_rn(x[idx], w), 2.f);
    out[idx] = (float)((double)s - c);
  } else {
    out[idx] = (float)(((double)x[idx] + c) * (double)w / 2.0);
  }
}

// camera_to_world (common/camera.py:27-28) with one quaternion for every point: qort(q, v) + t
// (common/quaternion.py:4-18: v + 2 * (q0 * (qv x v) + qv x (qv x v)))
__global__ void cam_to_world_kernel(const float* __restrict__ x, float* __restrict__ out, long long npts, float q0, float q1,
                                    float q2, float q3, float t0, float t1, float t2) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const float vx = x[3 * i], vy = x[3 * i + 1], vz = x[3 * i + 2];
  const float ux = q2 * vz - q3 * vy, uy = q3 * vx - q1 * vz, uz = q1 * vy - q2 * vx;
  const float wx = q2 * uz - q3 * uy, wy = q3 * ux - q1 * uz, wz = q1 * uy - q2 * ux;
  out[3 * i] = vx + 2.f * (q0 * ux + wx) + t0;
  out[3 * i + 1] = vy + 2.f * (q0 * uy + wy) + t1;
  out[3 * i + 2] = vz + 2.f * (q0 * uz + wz) + t2;
}

// ------------------------------------------------------------------------------------------
// N2: mpjpe (common/loss.py:5-11) forward + backward in one pass.  loss = mean_i ||p_i - t_i||;
// d loss / d p_i = (p_i - t_i) / (||.|| n)  (0 where the norm is 0, like torch.norm's backward).
// ------------------------------------------------------------------------------------------
constexpr int MPJPE_BLOCKS = 256;

__global__ void mpjpe_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long n, int D,
                             float gscale, float* __restrict__ dpred, double* __restrict__ partial) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d[4], s = 0.f;
    for (int c = 0; c < D; ++c) { d[c] = pred[i * D + c] - tgt[i * D + c]; s = fmaf(d[c], d[c], s); }
    const float nrm = sqrtf(s);
    acc += (double)nrm;
    if (dpred) {
      const float k = nrm > 0.f ? gscale / nrm : 0.f;
      for (int c = 0; c < D; ++c) dpred[i * D + c] = d[c] * k;
    }
  }
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

__global__ void mpjpe_final_kernel(const double* __restrict__ partial, int nb, long long n, float* __restrict__ loss) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nb; i += 32) v += partial[i];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (threadIdx.x == 0) *loss = (float)(v / (double)n);
}

// ------------------------------------------------------------------------------------------
// N2: p_mpjpe (common/loss.py:14-53): per frame, similarity-Procrustes alignment of the prediction to
// the target (3x3 SVD), then the mean joint distance.  One warp per frame, lanes over joints; the 3x3
// decomposition runs in double (Jacobi on H^T H) on every lane redundantly.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ inline void jacobi_eig3(double a[3][3], double v[3][3]) {   // a symmetric -> a diagonal, v eigenvectors (columns)
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p][q];
        if (fabs(apq) < 1e-300) continue;
        const double th = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double tt = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

__global__ void p_mpjpe_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, float* __restrict__ err,
                               int N, int J) {
  const int frame = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (frame >= N) return;
  const bool on = lane < J;
  double X[3] = {0, 0, 0}, Y[3] = {0, 0, 0};
  if (on)
    for (int c = 0; c < 3; ++c) { X[c] = tgt[((long long)frame * J + lane) * 3 + c]; Y[c] = pred[((long long)frame * J + lane) * 3 + c]; }
  double muX[3], muY[3];
  for (int c = 0; c < 3; ++c) { muX[c] = warp_sum_d(X[c]) / J; muY[c] = warp_sum_d(Y[c]) / J; }
  double X0[3], Y0[3], nx = 0, ny = 0;
  for (int c = 0; c < 3; ++c) {
    X0[c] = on ? X[c] - muX[c] : 0.0; Y0[c] = on ? Y[c] - muY[c] : 0.0;
    nx += X0[c] * X0[c]; ny += Y0[c] * Y0[c];
  }
  const double normX = sqrt(warp_sum_d(nx)), normY = sqrt(warp_sum_d(ny));
  for (int c = 0; c < 3; ++c) { X0[c] /= normX; Y0[c] /= normY; }
  double H[3][3];                                         // H = X0^T Y0   (:33)
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) H[i][j] = warp_sum_d(X0[i] * Y0[j]);
  // SVD  H = U diag(s) V^T  from the eigen-decomposition of H^T H
  double B[3][3], V[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[i][j] = H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j];
  jacobi_eig3(B, V);
  double lam[3] = {B[0][0], B[1][1], B[2][2]};
  int ord[3] = {0, 1, 2};                                 // descending eigenvalues (numpy's singular-value order)
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2 - i; ++j)
    if (lam[ord[j]] < lam[ord[j + 1]]) { int tq = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = tq; }
  double Vs[3][3], U[3][3], s[3];
  for (int k = 0; k < 3; ++k) {
    s[k] = sqrt(fmax(lam[ord[k]], 0.0));
    for (int i = 0; i < 3; ++i) Vs[i][k] = V[i][ord[k]];
  }
  for (int k = 0; k < 2; ++k)
    for (int i = 0; i < 3; ++i)
      U[i][k] = (H[i][0] * Vs[0][k] + H[i][1] * Vs[1][k] + H[i][2] * Vs[2][k]) / fmax(s[k], 1e-300);
  {
    double u2[3] = {H[0][0] * Vs[0][2] + H[0][1] * Vs[1][2] + H[0][2] * Vs[2][2],
                    H[1][0] * Vs[0][2] + H[1][1] * Vs[1][2] + H[1][2] * Vs[2][2],
                    H[2][0] * Vs[0][2] + H[2][1] * Vs[1][2] + H[2][2] * Vs[2][2]};
    const double cr[3] = {U[1][0] * U[2][1] - U[2][0] * U[1][1], U[2][0] * U[0][1] - U[0][0] * U[2][1],
                          U[0][0] * U[1][1] - U[1][0] * U[0][1]};
    if (s[2] > 1e-9 * s[0]) { for (int i = 0; i < 3; ++i) U[i][2] = u2[i] / s[2]; }
    else { for (int i = 0; i < 3; ++i) U[i][2] = cr[i]; }   // rank-deficient H: any unit vector completing the basis
  }
  auto det3 = [](double m[3][3]) {
    return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  };
  if (det3(Vs) * det3(U) < 0.0) {                         // det(R) = -1: flip the last singular direction (:39-43)
    for (int i = 0; i < 3; ++i) Vs[i][2] = -Vs[i][2];
    s[2] = -s[2];
  }
  double R[3][3];                                         // R = V U^T
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + Vs[i][2] * U[j][2];
  const double a = (s[0] + s[1] + s[2]) * normX / normY;  // scale (:45-47)
  double e = 0.0;
  if (on) {
    double d2 = 0.0;
    for (int c = 0; c < 3; ++c) {
      const double yr = Y[0] * R[0][c] + Y[1] * R[1][c] + Y[2] * R[2][c];
      const double mr = muY[0] * R[0][c] + muY[1] * R[1][c] + muY[2] * R[2][c];
      const double al = a * yr + (muX[c] - a * mr);       // a * pred . R + t   (:48-51)
      d2 += (al - X[c]) * (al - X[c]);
    }
    e = sqrt(d2);
  }
  e = warp_sum_d(e);
  if (lane == 0) err[frame] = (float)(e / J);
}

// ------------------------------------------------------------------------------------------
// N2: Adam with amsgrad (trainval.py:78 optim.Adam(..., amsgrad=True); torch/optim/adam.py single-tensor
// update) over every parameter in ONE launch.  table[i] = {param ptr, grad ptr, state offset, count}:
// a chunk of <= ADAM_CHUNK elements of one tensor; the moments live in flat state buffers.
// ------------------------------------------------------------------------------------------
constexpr int ADAM_CHUNK = 4096;

__global__ void adam_multi_kernel(const long long* __restrict__ table, float* __restrict__ exp_avg,
                                  float* __restrict__ exp_avg_sq, float* __restrict__ max_sq, float one_m_b1, float b2,
                                  float one_m_b2, float step_size, float bc2_sqrt, float eps, float wd) {
  const long long* e = table + 4LL * blockIdx.x;
  float* p = reinterpret_cast<float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  const long long so = e[2];
  const int cnt = (int)e[3];
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    float pv = p[i], gv = g[i];
    if (wd != 0.f) gv = fmaf(wd, pv, gv);
    float m = exp_avg[so + i], v = exp_avg_sq[so + i];
    m = m + one_m_b1 * (gv - m);                          // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + one_m_b2 * gv * gv;                      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    exp_avg[so + i] = m;
    exp_avg_sq[so + i] = v;
    float vd = v;
    if (max_sq) { vd = fmaxf(max_sq[so + i], v); max_sq[so + i] = vd; }
    const float denom = sqrtf(vd) / bc2_sqrt + eps;
    p[i] = pv - step_size * (m / denom);
  }
}

}  // namespace gast

extern "C" int gast_chunk_gather(const float* poses_2d, const float* poses_3d, const float* cameras,
                                 const int64_t* seq_start, int32_t n_seq, const int32_t* pairs, int32_t B,
                                 int32_t chunk, int32_t pad, int32_t causal_shift, int32_t J2, int32_t F2, int32_t J3,
                                 int32_t ncam, int32_t n_sym2, const int32_t* kps_left, const int32_t* kps_right,
                                 int32_t n_sym3, const int32_t* joints_left, const int32_t* joints_right,
                                 float* batch_2d, float* batch_3d, float* batch_cam, void* stream) {
  if (!poses_2d || !seq_start || !pairs || !batch_2d || B <= 0 || chunk <= 0 || pad < 0 || n_seq <= 0 || J2 <= 0 || F2 <= 0)
    return fail("gast_chunk_gather: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  JointPerm p2, p3;
  if (make_perm(J2, n_sym2, kps_left, kps_right, &p2)) return 1;
  const int Tc = chunk + 2 * pad;
  const long long n2 = (long long)B * Tc * J2 * F2;
  chunk_gather_kernel<<<cdiv(n2, 256), 256, 0, st>>>(poses_2d, reinterpret_cast<const long long*>(seq_start), pairs,
                                                     batch_2d, B, Tc, J2, F2, pad + causal_shift, p2);
  if (poses_3d && batch_3d) {
    if (make_perm(J3, n_sym3, joints_left, joints_right, &p3)) return 1;
    const long long n3 = (long long)B * chunk * J3 * 3;
    chunk_gather_kernel<<<cdiv(n3, 256), 256, 0, st>>>(poses_3d, reinterpret_cast<const long long*>(seq_start), pairs,
                                                       batch_3d, B, chunk, J3, 3, 0, p3);
  }
  if (cameras && batch_cam && ncam > 0)
    chunk_cam_kernel<<<cdiv((long long)B * ncam, 256), 256, 0, st>>>(cameras, pairs, batch_cam, B, ncam);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int gast_keypoints_convert(const float* kpts, float* out, int32_t* valid, int32_t T, int32_t J_in, int32_t mode,
                                      void* stream) {
  if (!kpts || !out || T <= 0) return fail("gast_keypoints_convert: bad arguments");
  if ((mode == 0 && J_in != 17) || (mode == 1 && J_in != 16) || (mode == 2 && J_in < 22) || mode < 0 || mode > 2)
    return fail("gast_keypoints_convert: mode %d does not take %d joints", mode, J_in);
  kpt_convert_kernel<<<cdiv(T, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(kpts, out, valid, T, J_in, mode);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int gast_normalize_screen(const float* x, float* out, int64_t n_points, float w, float h, int32_t inverse,
                                     void* stream) {
  if (!x || !out || n_points <= 0 || w == 0.f) return fail("gast_normalize_screen: bad arguments");
  const long long n = 2 * n_points;
  screen_norm_kernel<<<cdiv(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, n, w, (double)h / (double)w,
                                                                                      inverse);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int gast_camera_to_world(const float* x, float* out, int64_t n_points, const float* q, const float* t,
                                    void* stream) {
  if (!x || !out || !q || n_points <= 0) return fail("gast_camera_to_world: bad arguments");
  cam_to_world_kernel<<<cdiv(n_points, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, out, n_points, q[0], q[1], q[2], q[3], t ? t[0] : 0.f, t ? t[1] : 0.f, t ? t[2] : 0.f);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" size_t gast_mpjpe_workspace_bytes(void) { return sizeof(double) * MPJPE_BLOCKS; }

extern "C" int gast_mpjpe(const float* pred, const float* target, int64_t n_points, int32_t D, float* loss, float* dpred,
                          float grad_scale, void* workspace, void* stream) {
  if (!pred || !target || !loss || !workspace || n_points <= 0 || D < 1 || D > 4) return fail("gast_mpjpe: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nb = (int)std::min<long long>(MPJPE_BLOCKS, (n_points + 255) / 256);
  mpjpe_kernel<<<nb, 256, 0, st>>>(pred, target, n_points, D, grad_scale / (float)n_points, dpred,
                                  reinterpret_cast<double*>(workspace));
  mpjpe_final_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const double*>(workspace), nb, n_points, loss);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int gast_p_mpjpe(const float* pred, const float* target, int32_t N, int32_t J, float* per_frame, void* stream) {
  if (!pred || !target || !per_frame || N <= 0 || J < 3 || J > 32) return fail("gast_p_mpjpe: bad arguments (3 <= J <= 32)");
  p_mpjpe_kernel<<<cdiv((long long)N * 32, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, target, per_frame, N, J);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int32_t gast_adam_chunk(void) { return ADAM_CHUNK; }

extern "C" int gast_adam_step(const int64_t* table, int32_t n_chunks, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                              double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                              void* stream) {
  if (!table || !exp_avg || !exp_avg_sq || n_chunks <= 0 || step < 1) return fail("gast_adam_step: bad arguments");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  adam_multi_kernel<<<n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(table), exp_avg, exp_avg_sq, max_exp_avg_sq, (float)(1.0 - beta1), (float)beta2,
      (float)(1.0 - beta2), (float)(lr / bc1), (float)sqrt(bc2), (float)eps, (float)weight_decay);
  CUDA_OK(cudaGetLastError());
  return 0;
}
