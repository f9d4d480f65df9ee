// Loss evaluation of the reference's training / validation loops (forward values):
//   MSE, L1 (nn.MSELoss / nn.L1Loss, trainers/trainer.py:103-104), gradient difference loss (trainers/trainer.py:65-83)
//   and the bidirectional patch contrastive loss BiPatchNCE (models/contrastive_loss.py:28-60), combined as in
//   Trainer.criterion (trainers/trainer.py:88-109).  Inputs are the (P, B, E) sequence-first slices pred[-P:] and
//   y_expected[-P:] the reference passes (trainers/trainer.py:145, :224).
// Two kernels + a finalise: one streaming pass computes the MSE / L1 / GDL sums (HBM-bound: 8 bytes in per element,
// 128-bit loads, deterministic two-stage reduction in double); the contrastive loss runs one CTA per (clip, frame)
// with both (h*w) x 4 feature matrices in shared memory and an online log-sum-exp per row, in both directions -
// the reference materialises two (N*T, hw, hw) score tensors and an int64 mask of the same size per call.
#pragma once
#include "common.cuh"

namespace sdvg {

constexpr int kLossBlocks = 592;   // 148 SMs x 4
constexpr int kLossThreads = 256;

struct LossArgs {
  const float* x; const float* y;   // prediction, ground truth: (P, B, E) contiguous
  int P, B, h, w;                   // E = 4 h w
  float alpha;                      // GDL exponent
  float inv_temperature;
  double* partial;                  // [kLossBlocks][4]: sum (x-y)^2, sum |x-y|, GDL sum, unused
  double* nce_partial;              // [P*B][2]: sum over rows of (lse - diag), direction 1 and 2
};

__device__ __forceinline__ float gdl_term(float dx, float dy, float alpha) {
  const float v = fabsf(fabsf(dx) - fabsf(dy));
  if (alpha == 1.0f) return v;
  if (alpha == 2.0f) return v * v;
  return powf(v, alpha);
}

__global__ void __launch_bounds__(kLossThreads) loss_elementwise_kernel(const __grid_constant__ LossArgs a) {
  __shared__ double red[3][kLossThreads / 32];
  pdl_wait();
  pdl_trigger();
  const int hw = a.h * a.w, E = 4 * hw;
  const long long total = static_cast<long long>(a.P) * a.B * E;
  double s_mse = 0.0, s_l1 = 0.0, s_gdl = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(kLossThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kLossThreads) {
    const float xv = __ldg(a.x + i), yv = __ldg(a.y + i);
    const float d = xv - yv;
    float f_mse = d * d, f_l1 = fabsf(d), f_gdl = 0.f;
    const int e = static_cast<int>(i % E);
    const int p = e % hw, r = p / a.w, c = p - r * a.w;
    if (r + 1 < a.h) f_gdl += gdl_term(__ldg(a.x + i + a.w) - xv, __ldg(a.y + i + a.w) - yv, a.alpha);   // vertical
    if (c + 1 < a.w) f_gdl += gdl_term(__ldg(a.x + i + 1) - xv, __ldg(a.y + i + 1) - yv, a.alpha);       // horizontal
    s_mse += f_mse; s_l1 += f_l1; s_gdl += f_gdl;
  }
  double v[3] = {s_mse, s_l1, s_gdl};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int wi = 0; wi < kLossThreads / 32; ++wi) t += red[threadIdx.x][wi];
    a.partial[blockIdx.x * 4 + threadIdx.x] = t;
  }
}

// One CTA per (clip n, frame t).  Row i of direction 1: lse_j(gt_i . pred_j / tau) - gt_i . pred_i / tau;
// direction 2 swaps the roles (scores2 = scores1^T).  Features are (hw, C=4): element (i, c) = v[t][n][c*hw + i].
__global__ void __launch_bounds__(256) loss_nce_kernel(const __grid_constant__ LossArgs a) {
  extern __shared__ float4 nce_smem[];   // [2][hw]: pred features, gt features
  __shared__ double red[2][8];
  pdl_wait();
  pdl_trigger();
  const int hw = a.h * a.w;
  const int n = blockIdx.x % a.B, t = blockIdx.x / a.B;
  const float* xp = a.x + (static_cast<size_t>(t) * a.B + n) * 4 * hw;
  const float* yp = a.y + (static_cast<size_t>(t) * a.B + n) * 4 * hw;
  float4* sp = nce_smem;
  float4* sg = nce_smem + hw;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    sp[i] = make_float4(__ldg(xp + i), __ldg(xp + hw + i), __ldg(xp + 2 * hw + i), __ldg(xp + 3 * hw + i));
    sg[i] = make_float4(__ldg(yp + i), __ldg(yp + hw + i), __ldg(yp + 2 * hw + i), __ldg(yp + 3 * hw + i));
  }
  __syncthreads();
  double acc1 = 0.0, acc2 = 0.0;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const float4 gi = sg[i], pi = sp[i];
    float m1 = -INFINITY, s1 = 0.f, m2 = -INFINITY, s2 = 0.f;
    for (int j = 0; j < hw; ++j) {
      const float4 pj = sp[j], gj = sg[j];
      const float v1 = (gi.x * pj.x + gi.y * pj.y + gi.z * pj.z + gi.w * pj.w) * a.inv_temperature;
      const float v2 = (pi.x * gj.x + pi.y * gj.y + pi.z * gj.z + pi.w * gj.w) * a.inv_temperature;
      if (v1 > m1) { s1 = s1 * __expf(m1 - v1) + 1.f; m1 = v1; } else { s1 += __expf(v1 - m1); }
      if (v2 > m2) { s2 = s2 * __expf(m2 - v2) + 1.f; m2 = v2; } else { s2 += __expf(v2 - m2); }
    }
    const float diag = (gi.x * pi.x + gi.y * pi.y + gi.z * pi.z + gi.w * pi.w) * a.inv_temperature;
    acc1 += static_cast<double>(m1 + logf(s1) - diag);
    acc2 += static_cast<double>(m2 + logf(s2) - diag);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
    acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc1; red[1][threadIdx.x >> 5] = acc2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double tsum = 0.0;
    for (int wi = 0; wi < static_cast<int>(blockDim.x >> 5); ++wi) tsum += red[threadIdx.x][wi];
    a.nce_partial[blockIdx.x * 2 + threadIdx.x] = tsum;
  }
}

struct LossFinalArgs {
  const double* partial; int n_blocks;
  const double* nce_partial; int n_ct;     // n_ct = P*B (0: contrastive term not requested)
  double numel, nce_rows;                  // P*B*E ; P*B*hw
  int use_mse, use_l1, use_gdl, use_nce;
  float lambda_gdl, lambda_nce;
  float* out;                              // [5]: total, mse, l1, gdl, nce
};

__global__ void loss_finalize_kernel(const __grid_constant__ LossFinalArgs a) {
  pdl_wait();
  pdl_trigger();
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  // one warp: fixed strided order + shuffle tree (deterministic), instead of one thread walking ~2000 partials
  const int lane = threadIdx.x;
  double s[3] = {0.0, 0.0, 0.0};
  for (int b = lane; b < a.n_blocks; b += 32)
    for (int k = 0; k < 3; ++k) s[k] += a.partial[b * 4 + k];
  double n1 = 0.0, n2 = 0.0;
  for (int b = lane; b < a.n_ct; b += 32) { n1 += a.nce_partial[b * 2]; n2 += a.nce_partial[b * 2 + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    for (int k = 0; k < 3; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  }
  if (lane != 0) return;
  const double mse = s[0] / a.numel, l1 = s[1] / a.numel, gdl = s[2] / a.numel;
  const double nce = a.n_ct ? 0.5 * (n1 + n2) / a.nce_rows : 0.0;
  const double total = a.use_mse * mse + a.use_l1 * l1 + a.use_gdl * a.lambda_gdl * gdl + a.use_nce * a.lambda_nce * nce;
  a.out[0] = static_cast<float>(total); a.out[1] = static_cast<float>(mse); a.out[2] = static_cast<float>(l1);
  a.out[3] = static_cast<float>(gdl); a.out[4] = static_cast<float>(nce);
}

}  // namespace sdvg
