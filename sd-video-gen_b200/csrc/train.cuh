// Kernels of the training step (trainers/trainer.py:123-165: teacher-forced forward, criterion, loss.backward(),
// Adam) that are not GEMMs.  Every linear layer's backward is two tensor-core GEMMs of the same kernels the
// forward uses (gemm_tc*.cuh):
//     dX[M,K] = dY[M,N] . W[N,K]          A = dY planes,            B = W^T planes [K][N]
//     dW[N,K] = dY^T[N,M] . X[M,K]        A = dY^T planes [N][Mp],  B = X^T planes [K][Mp]
// so what is needed around them is (1) pack_t: fp32 matrix -> operand planes, transposed operand planes and
// column sums (the bias gradient) in one pass, (2) LayerNorm backward, (3) attention backward, (4) the gradient of
// the criterion (MSE / L1 / GDL / BiPatchNCE), (5) Adam.  All HBM/L2-bound streaming kernels with coalesced
// 128-byte rows; nothing here allocates or synchronises.
//
// Gradients flow through 16-bit operand planes, so they are kept multiplied by a power-of-two loss scale S chosen
// on the device from max|dL/dpred| (loss_scale_kernel); every parameter gradient is multiplied by 1/S where it is
// written (GEMM epilogue Epilogue::alpha_dev, column sums here), so the gradient vector handed to the all-reduce
// and to Adam is unscaled fp32.
#pragma once
#include "common.cuh"

namespace sdvg {

constexpr int kTrainMaxS = 16;   // tokens per clip in a training step (the reference uses 6 and 5)

// ------------------------------------------------------------------------------------------------ pack_t
struct PackTArgs {
  const float* src; int ld_src;        // [R][C] fp32
  int R, C;
  int perm_S, perm_B;                  // perm_S > 0: destination row r reads source row (r % S) * B + r / S (sequence-major -> clip-major)
  float mul; const float* mul_dev;     // v = src * mul * (mul_dev ? *mul_dev : 1)
  uint16_t* out_hi; uint16_t* out_lo; int ld16;            // planes [R][C] (nullable)
  uint16_t* t_hi; uint16_t* t_lo; int ld_t; int Rpad;      // transposed planes [C][ld_t]; columns [R, Rpad) are zero-filled (nullable)
  float* colsum; const float* colsum_mul_dev; int colsum_accumulate;   // colsum[c] (+)= (*colsum_mul_dev) * sum_r v[r][c]  (nullable)
  int bf16;
  int rows_per_block;                  // rows one CTA walks through (a multiple of 128)
  uint32_t drop_thr, drop_key; float drop_scale;   // drop_thr != 0: v *= dropout mask of element (r, c) (index r * C + c)
};

__global__ void __launch_bounds__(256) pack_t_kernel(const __grid_constant__ PackTArgs a) {
  __shared__ float tile[32][129];
  __shared__ float red[8][32];
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  const float mul = a.mul * (a.mul_dev ? __ldg(a.mul_dev) : 1.0f);
  const int row_begin = blockIdx.y * a.rows_per_block;
  int row_end = row_begin + a.rows_per_block;
  const int limit = a.t_hi ? a.Rpad : a.R;
  if (row_end > limit) row_end = limit;
  float cs = 0.f;
  for (int r0 = row_begin; r0 < row_end; r0 += 128) {
    // all 16 row loads of this lane are issued before any is used (the kernel is latency-bound: 400 KB per launch)
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = r0 + warp + 8 * i;
      v[i] = 0.f;
      if (r < a.R && c < a.C) {
        const int sr = a.perm_S > 0 ? (r % a.perm_S) * a.perm_B + r / a.perm_S : r;
        v[i] = __ldg(a.src + static_cast<size_t>(sr) * a.ld_src + c);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int rl = warp + 8 * i, r = r0 + rl;
      float x = v[i] * mul;
      if (a.drop_thr)
        x = drop_hash(a.drop_key, static_cast<uint32_t>(r) * static_cast<uint32_t>(a.C) + static_cast<uint32_t>(c)) >= a.drop_thr
                ? x * a.drop_scale : 0.f;
      if (a.out_hi && r < a.R && c < a.C) {
        const uint16_t h = to_plane_hi(x, a.bf16);
        a.out_hi[static_cast<size_t>(r) * a.ld16 + c] = h;
        if (a.out_lo) a.out_lo[static_cast<size_t>(r) * a.ld16 + c] = to_plane_lo(x, h);
      }
      cs += x;
      tile[lane][rl] = x;
    }
    __syncthreads();
    if (a.t_hi) {
      // warp w writes columns c0 + w, w + 8, ...; each lane 4 consecutive rows (8 bytes per plane)
      for (int cl = warp; cl < 32; cl += 8) {
        const int cc = blockIdx.x * 32 + cl;
        const int r = r0 + lane * 4;
        if (cc < a.C && r < a.Rpad) {
          uint16_t h[4], l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float v = tile[cl][lane * 4 + i];
            h[i] = to_plane_hi(v, a.bf16);
            l[i] = to_plane_lo(v, h[i]);
          }
          const size_t o = static_cast<size_t>(cc) * a.ld_t + r;
          *reinterpret_cast<uint2*>(a.t_hi + o) = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
          if (a.t_lo) *reinterpret_cast<uint2*>(a.t_lo + o) = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
        }
      }
    }
    __syncthreads();
  }
  if (a.colsum) {
    red[warp][lane] = cs;
    __syncthreads();
    if (warp == 0 && c < a.C) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][lane];
      t *= a.colsum_mul_dev ? __ldg(a.colsum_mul_dev) : 1.0f;
      a.colsum[c] = a.colsum_accumulate ? a.colsum[c] + t : t;
    }
  }
}

inline cudaError_t launch_pack_t(PackTArgs a, cudaStream_t stream) {
  if (a.t_hi && (a.Rpad % 4 != 0 || a.ld_t % 4 != 0)) return cudaErrorInvalidValue;
  const int limit = a.t_hi ? a.Rpad : a.R;
  // the column sums need every row in one CTA; otherwise split long matrices (weights) over the grid
  a.rows_per_block = a.colsum ? ((limit + 127) / 128) * 128 : 128;
  const dim3 grid(ceil_div(a.C, 32), ceil_div(limit, a.rows_per_block));
  return launch_kernel(pack_t_kernel, grid, dim3(256), 0, stream, a);
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// y = LN(x) = xhat * w + b, xhat = (x - mean) * rstd.  With g = dy * w:
//   dx = rstd * (g - mean_j(g) - xhat * mean_j(g * xhat));   dw = sum_rows dy * xhat;   db = sum_rows dy.
// CTAs [0, rows): one row each (dx).  CTAs [rows, rows + d/32): 32 columns each over all rows (dw, db).
struct LnBwdArgs {
  const float* dy; int ld_dy;
  const float* x; int ld_x;        // saved LayerNorm input
  const float2* stats;             // (mean, rstd) per row, saved by the forward kernel
  const float* w;
  int rows, d;
  float* dx; int ld_dx;
  float* dw; float* db;            // parameter gradients [d]
  const float* param_mul_dev;      // 1 / loss scale
};

__global__ void __launch_bounds__(256) ln_backward_kernel(const __grid_constant__ LnBwdArgs a) {
  __shared__ float red[2][8][32];
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (static_cast<int>(blockIdx.x) < a.rows) {
    const int r = blockIdx.x;
    const float2 st = __ldg(a.stats + r);
    const float* dy = a.dy + static_cast<size_t>(r) * a.ld_dy;
    const float* x = a.x + static_cast<size_t>(r) * a.ld_x;
    float s1 = 0.f, s2 = 0.f;
    for (int j = threadIdx.x; j < a.d; j += 256) {
      const float g = __ldg(dy + j) * __ldg(a.w + j);
      const float xh = (__ldg(x + j) - st.x) * st.y;
      s1 += g; s2 = fmaf(g, xh, s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) { red[0][warp][0] = s1; red[1][warp][0] = s2; }
    __syncthreads();
    s1 = 0.f; s2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s1 += red[0][w][0]; s2 += red[1][w][0]; }
    const float m1 = s1 / static_cast<float>(a.d), m2 = s2 / static_cast<float>(a.d);
    float* dx = a.dx + static_cast<size_t>(r) * a.ld_dx;
    for (int j = threadIdx.x; j < a.d; j += 256) {
      const float g = __ldg(dy + j) * __ldg(a.w + j);
      const float xh = (__ldg(x + j) - st.x) * st.y;
      dx[j] = st.y * (g - m1 - xh * m2);
    }
    return;
  }
  const int c = (blockIdx.x - a.rows) * 32 + lane;
  float sw = 0.f, sb = 0.f;
  if (c < a.d) {
    for (int r = warp; r < a.rows; r += 8) {
      const float2 st = __ldg(a.stats + r);
      const float g = __ldg(a.dy + static_cast<size_t>(r) * a.ld_dy + c);
      const float xh = (__ldg(a.x + static_cast<size_t>(r) * a.ld_x + c) - st.x) * st.y;
      sw = fmaf(g, xh, sw); sb += g;
    }
  }
  red[0][warp][lane] = sw; red[1][warp][lane] = sb;
  __syncthreads();
  if (warp == 0 && c < a.d) {
    float tw = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tw += red[0][w][lane]; tb += red[1][w][lane]; }
    const float m = a.param_mul_dev ? __ldg(a.param_mul_dev) : 1.0f;
    a.dw[c] = tw * m; a.db[c] = tb * m;
  }
}

inline cudaError_t launch_ln_backward(const LnBwdArgs& a, cudaStream_t stream) {
  return launch_kernel(ln_backward_kernel, dim3(a.rows + ceil_div(a.d, 32)), dim3(256), 0, stream, a);
}

// ------------------------------------------------------------------------------------------------ attention backward
// One warp per (clip, head); S <= kTrainMaxS so the score matrices live in shared memory (1 KB each per warp).
//   P = softmax(scale Q K^T + mask);  O = P V
//   dV = P^T dO;  dP = dO V^T;  dS = P o (dP - rowsum(P o dP));  dQ = scale dS K;  dK = scale dS^T Q
struct AttnBwdArgs {
  const float* q; int ldq; const float* k; const float* v; int ldkv;
  const float* dO; int ld_do;
  float* dq; int ld_dq; float* dk; float* dv; int ld_dkv;
  int clips, heads, hd, Sq, Sk, causal;
  float scale;
  int cshift;   // hd == 32 << cshift, or -1
  uint32_t drop_thr, drop_key; float drop_scale;   // dropout on the attention probabilities: index ((b H + h) Sq + i) Sk + j
};

// SMAX bounds both sequence lengths at compile time (6 covers the reference's 5 / 6 token windows) so that the
// score / output loops unroll and their shared-memory loads are issued in batches instead of one per FMA.
template <int SMAX>
__global__ void __launch_bounds__(128) attention_backward_kernel(const __grid_constant__ AttnBwdArgs a) {
  extern __shared__ float attn_bwd_smem[];   // per warp: Q[Sq][hd+1], dO[Sq][hd+1], K[Sk][hd+1], V[Sk][hd+1]
  __shared__ float sP[4][kTrainMaxS][kTrainMaxS];
  __shared__ float sD[4][kTrainMaxS][kTrainMaxS];
  pdl_wait();
  pdl_trigger();
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = blockIdx.x * 4 + wib;
  if (wg >= a.clips * a.heads) return;
  const int b = wg / a.heads, h = wg - b * a.heads;
  const int hd = a.hd, Sq = a.Sq, Sk = a.Sk, ldp = hd + 1;
  const float* q = a.q + static_cast<size_t>(b) * Sq * a.ldq + h * hd;
  const float* k = a.k + static_cast<size_t>(b) * Sk * a.ldkv + h * hd;
  const float* v = a.v + static_cast<size_t>(b) * Sk * a.ldkv + h * hd;
  const float* dO = a.dO + static_cast<size_t>(b) * Sq * a.ld_do + h * hd;
  float* sQ = attn_bwd_smem + static_cast<size_t>(wib) * (2 * Sq + 2 * Sk) * ldp;
  float* sO = sQ + Sq * ldp;
  float* sK = sO + Sq * ldp;
  float* sV = sK + Sk * ldp;
  float (*P)[kTrainMaxS] = sP[wib];
  float (*D)[kTrainMaxS] = sD[wib];
  // stage the four operands (smem rows are laid out Q, dO, K, V).  hd = 32 << cshift: 32-element chunks, eight
  // independent loads in flight per lane (16 per batch), shifts instead of divisions; other head sizes take the plain loop
  {
    const int nrows = 2 * Sq + 2 * Sk;
    auto row_ptr = [&](int r) -> const float* {
      return r < Sq ? q + static_cast<size_t>(r) * a.ldq
           : r < 2 * Sq ? dO + static_cast<size_t>(r - Sq) * a.ld_do
           : r < 2 * Sq + Sk ? k + static_cast<size_t>(r - 2 * Sq) * a.ldkv
                             : v + static_cast<size_t>(r - 2 * Sq - Sk) * a.ldkv;
    };
    if (a.cshift >= 0) {
      const int nchunks = nrows << a.cshift, cmask = (1 << a.cshift) - 1;
      for (int c0 = 0; c0 < nchunks; c0 += 16) {
        float tmp[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int ci = c0 + u;
          tmp[u] = ci < nchunks ? __ldg(row_ptr(ci >> a.cshift) + ((ci & cmask) << 5) + lane) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int ci = c0 + u;
          if (ci < nchunks) sQ[(ci >> a.cshift) * ldp + ((ci & cmask) << 5) + lane] = tmp[u];
        }
      }
    } else {
      for (int r = 0; r < nrows; ++r)
        for (int e = lane; e < hd; e += 32) sQ[r * ldp + e] = __ldg(row_ptr(r) + e);
    }
  }
  __syncwarp();
  // scores and dP: one (query, key) pair per lane (row pitch hd+1: conflict-free)
  for (int p = lane; p < Sq * Sk; p += 32) {
    const int i = p / Sk, j = p - i * Sk;
    float s = 0.f, dp = 0.f;
#pragma unroll 8
    for (int e = 0; e < hd; ++e) {
      s = fmaf(sQ[i * ldp + e], sK[j * ldp + e], s);
      dp = fmaf(sO[i * ldp + e], sV[j * ldp + e], dp);
    }
    const bool masked = a.causal && j > i + (Sk - Sq);
    P[i][j] = masked ? -INFINITY : s * a.scale;
    D[i][j] = dp;
  }
  __syncwarp();
  if (lane < Sq) {
    const int i = lane;
    float mx = -INFINITY;
    for (int j = 0; j < Sk; ++j) mx = fmaxf(mx, P[i][j]);
    float sum = 0.f;
    for (int j = 0; j < Sk; ++j) { const float p = __expf(P[i][j] - mx); P[i][j] = p; sum += p; }
    const float inv = 1.0f / sum;
    float dot = 0.f;
    const uint32_t drow = static_cast<uint32_t>((b * a.heads + h) * Sq + i) * static_cast<uint32_t>(Sk);
    for (int j = 0; j < Sk; ++j) {
      P[i][j] *= inv;
      if (a.drop_thr) D[i][j] = drop_hash(a.drop_key, drow + j) >= a.drop_thr ? D[i][j] * a.drop_scale : 0.f;   // dP = dP_d o mask / (1-p)
      dot = fmaf(P[i][j], D[i][j], dot);
    }
    for (int j = 0; j < Sk; ++j) {
      D[i][j] = P[i][j] * (D[i][j] - dot);   // dS
      if (a.drop_thr) P[i][j] = drop_hash(a.drop_key, drow + j) >= a.drop_thr ? P[i][j] * a.drop_scale : 0.f;   // P_d, for dV
    }
  }
  __syncwarp();
  for (int e = lane; e < hd; e += 32) {
    float kk[SMAX], qq[SMAX], oo[SMAX];
#pragma unroll
    for (int j = 0; j < SMAX; ++j) kk[j] = j < Sk ? sK[j * ldp + e] : 0.f;
#pragma unroll
    for (int i = 0; i < SMAX; ++i) { qq[i] = i < Sq ? sQ[i * ldp + e] : 0.f; oo[i] = i < Sq ? sO[i * ldp + e] : 0.f; }
#pragma unroll
    for (int i = 0; i < SMAX; ++i) {
      if (i < Sq) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < SMAX; ++j) if (j < Sk) acc = fmaf(D[i][j], kk[j], acc);
        a.dq[(static_cast<size_t>(b) * Sq + i) * a.ld_dq + h * hd + e] = acc * a.scale;
      }
    }
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
      if (j < Sk) {
        float ak = 0.f, av = 0.f;
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
          if (i < Sq) { ak = fmaf(D[i][j], qq[i], ak); av = fmaf(P[i][j], oo[i], av); }
        }
        a.dk[(static_cast<size_t>(b) * Sk + j) * a.ld_dkv + h * hd + e] = ak * a.scale;
        a.dv[(static_cast<size_t>(b) * Sk + j) * a.ld_dkv + h * hd + e] = av;
      }
    }
  }
}

inline cudaError_t launch_attention_backward(const AttnBwdArgs& a_in, cudaStream_t stream) {
  AttnBwdArgs a = a_in;
  a.cshift = -1;
  for (int sft = 0; sft < 4; ++sft) if (a.hd == (32 << sft)) a.cshift = sft;
  if (a.Sq > kTrainMaxS || a.Sk > kTrainMaxS) return cudaErrorInvalidValue;
  const size_t smem = 4 * static_cast<size_t>(2 * a.Sq + 2 * a.Sk) * (a.hd + 1) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 40 * 1024) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(attention_backward_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_backward_kernel<kTrainMaxS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return e;
      attr_set[dev & 63] = true;
    }
  }
  const dim3 grid(ceil_div(a.clips * a.heads, 4));
  if (a.Sq <= 6 && a.Sk <= 6) return launch_kernel(attention_backward_kernel<6>, grid, dim3(128), smem, stream, a);
  return launch_kernel(attention_backward_kernel<kTrainMaxS>, grid, dim3(128), smem, stream, a);
}

// ------------------------------------------------------------------------------------------------ attention forward (dropout)
// Training-mode attention with dropout on the probabilities (torch.nn.MultiheadAttention(dropout=p), as constructed
// by nn.Transformer at models/transformer.py:38-44).  Same staging as the backward kernel; used only when p > 0 -
// without dropout the training forward uses the inference attention kernels.
struct AttnTrainArgs {
  const float* q; int ldq; const float* k; const float* v; int ldkv;
  int clips, heads, hd, Sq, Sk, causal;
  float scale;
  float* out32; int ld32;
  uint16_t* out_hi; uint16_t* out_lo; int ld16; int bf16;
  uint32_t drop_thr, drop_key; float drop_scale;
  int cshift;   // hd == 32 << cshift, or -1
};

__global__ void __launch_bounds__(128) attention_train_fwd_kernel(const __grid_constant__ AttnTrainArgs a) {
  extern __shared__ float attn_fwd_smem[];   // per warp: Q[Sq][hd+1], K[Sk][hd+1], V[Sk][hd+1]
  __shared__ float sP[4][kTrainMaxS][kTrainMaxS];
  pdl_wait();
  pdl_trigger();
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = blockIdx.x * 4 + wib;
  if (wg >= a.clips * a.heads) return;
  const int b = wg / a.heads, h = wg - b * a.heads;
  const int hd = a.hd, Sq = a.Sq, Sk = a.Sk, ldp = hd + 1;
  const float* q = a.q + static_cast<size_t>(b) * Sq * a.ldq + h * hd;
  const float* k = a.k + static_cast<size_t>(b) * Sk * a.ldkv + h * hd;
  const float* v = a.v + static_cast<size_t>(b) * Sk * a.ldkv + h * hd;
  float* sQ = attn_fwd_smem + static_cast<size_t>(wib) * (Sq + 2 * Sk) * ldp;
  float* sK = sQ + Sq * ldp;
  float* sV = sK + Sk * ldp;
  float (*P)[kTrainMaxS] = sP[wib];
  {
    // stage Q, K, V (smem rows in that order); hd = 32 << cshift: 16 independent loads in flight per lane
    const int nrows = Sq + 2 * Sk;
    auto row_ptr = [&](int r) -> const float* {
      return r < Sq ? q + static_cast<size_t>(r) * a.ldq
           : r < Sq + Sk ? k + static_cast<size_t>(r - Sq) * a.ldkv : v + static_cast<size_t>(r - Sq - Sk) * a.ldkv;
    };
    if (a.cshift >= 0) {
      const int nchunks = nrows << a.cshift, cmask = (1 << a.cshift) - 1;
      for (int c0 = 0; c0 < nchunks; c0 += 16) {
        float tmp[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int ci = c0 + u;
          tmp[u] = ci < nchunks ? __ldg(row_ptr(ci >> a.cshift) + ((ci & cmask) << 5) + lane) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int ci = c0 + u;
          if (ci < nchunks) sQ[(ci >> a.cshift) * ldp + ((ci & cmask) << 5) + lane] = tmp[u];
        }
      }
    } else {
      for (int r = 0; r < nrows; ++r)
        for (int e = lane; e < hd; e += 32) sQ[r * ldp + e] = __ldg(row_ptr(r) + e);
    }
  }
  __syncwarp();
  for (int p = lane; p < Sq * Sk; p += 32) {
    const int i = p / Sk, j = p - i * Sk;
    float s = 0.f;
#pragma unroll 8
    for (int e = 0; e < hd; ++e) s = fmaf(sQ[i * ldp + e], sK[j * ldp + e], s);
    P[i][j] = (a.causal && j > i + (Sk - Sq)) ? -INFINITY : s * a.scale;
  }
  __syncwarp();
  if (lane < Sq) {
    const int i = lane;
    float mx = -INFINITY;
    for (int j = 0; j < Sk; ++j) mx = fmaxf(mx, P[i][j]);
    float sum = 0.f;
    for (int j = 0; j < Sk; ++j) { const float pv = __expf(P[i][j] - mx); P[i][j] = pv; sum += pv; }
    const float inv = 1.0f / sum;
    const uint32_t drow = static_cast<uint32_t>((b * a.heads + h) * Sq + i) * static_cast<uint32_t>(Sk);
    for (int j = 0; j < Sk; ++j) {
      float pv = P[i][j] * inv;
      if (a.drop_thr) pv = drop_hash(a.drop_key, drow + j) >= a.drop_thr ? pv * a.drop_scale : 0.f;
      P[i][j] = pv;
    }
  }
  __syncwarp();
  for (int e = lane; e < hd; e += 32) {
    for (int i = 0; i < Sq; ++i) {
      float acc = 0.f;
      for (int j = 0; j < Sk; ++j) acc = fmaf(P[i][j], sV[j * ldp + e], acc);
      const size_t row = static_cast<size_t>(b) * Sq + i;
      const int col = h * hd + e;
      if (a.out32) a.out32[row * a.ld32 + col] = acc;
      if (a.out_hi) {
        const uint16_t hi = to_plane_hi(acc, a.bf16);
        a.out_hi[row * a.ld16 + col] = hi;
        if (a.out_lo) a.out_lo[row * a.ld16 + col] = to_plane_lo(acc, hi);
      }
    }
  }
}

inline cudaError_t launch_attention_train_fwd(const AttnTrainArgs& a_in, cudaStream_t stream) {
  AttnTrainArgs a = a_in;
  a.cshift = -1;
  for (int sft = 0; sft < 4; ++sft) if (a.hd == (32 << sft)) a.cshift = sft;
  if (a.Sq > kTrainMaxS || a.Sk > kTrainMaxS) return cudaErrorInvalidValue;
  const size_t smem = 4 * static_cast<size_t>(a.Sq + 2 * a.Sk) * (a.hd + 1) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 40 * 1024) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(attention_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return e;
      attr_set[dev & 63] = true;
    }
  }
  return launch_kernel(attention_train_fwd_kernel, dim3(ceil_div(a.clips * a.heads, 4)), dim3(128), smem, stream, a);
}

// ------------------------------------------------------------------------------------------------ criterion gradient
// d criterion / d pred for the (P, B, E) slices of trainers/trainer.py:145 (sequence-major, E = 4 h w):
//   MSE 2 (x - y) / n;  L1 sign(x - y) / n;  GDL (trainers/trainer.py:65-83): each forward difference a = x[+1] - x
//   contributes alpha |abs(a) - abs(b)|^(alpha-1) sign(abs(a) - abs(b)) sign(a) / n to x[+1] and its negative to x;
//   BiPatchNCE (models/contrastive_loss.py:28-60): direction 1 back-propagates through the diagonal scores only
//   (the off-diagonal ones use pred.detach()), direction 2 through the whole row.
struct LossGradArgs {
  const float* x; const float* y;   // prediction, ground truth (P, B, E)
  int P, B, h, w;
  float c_mse, c_l1, c_gdl, alpha;  // coefficients already divided by n = P B E
  float c_nce, inv_temperature;     // 0.5 * lambda_c / (P B hw)
  float* grad;                      // (P, B, E), overwritten by the elementwise kernel, += by the NCE kernel
  unsigned int* amax_bits;          // max |grad| as float bits (atomicMax on non-negative floats)
};

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
__device__ __forceinline__ float gdl_dterm(float da, float db, float alpha) {
  const float u = fabsf(da) - fabsf(db);
  float m;
  if (alpha == 2.0f) m = 2.0f * fabsf(u);
  else if (alpha == 1.0f) m = 1.0f;
  else m = alpha * powf(fabsf(u), alpha - 1.0f);
  return m * sgn(u) * sgn(da);
}

__global__ void __launch_bounds__(256) loss_grad_elementwise_kernel(const __grid_constant__ LossGradArgs a) {
  pdl_wait();
  pdl_trigger();
  const int hw = a.h * a.w, E = 4 * hw;
  const long long total = static_cast<long long>(a.P) * a.B * E;
  float amax = 0.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const float xv = __ldg(a.x + i), yv = __ldg(a.y + i);
    const float d = xv - yv;
    float g = a.c_mse * 2.0f * d + a.c_l1 * sgn(d);
    if (a.c_gdl != 0.f) {
      const int e = static_cast<int>(i % E);
      const int p = e % hw, r = p / a.w, c = p - r * a.w;
      float t = 0.f;
      if (r + 1 < a.h) t -= gdl_dterm(__ldg(a.x + i + a.w) - xv, __ldg(a.y + i + a.w) - yv, a.alpha);
      if (r > 0) t += gdl_dterm(xv - __ldg(a.x + i - a.w), yv - __ldg(a.y + i - a.w), a.alpha);
      if (c + 1 < a.w) t -= gdl_dterm(__ldg(a.x + i + 1) - xv, __ldg(a.y + i + 1) - yv, a.alpha);
      if (c > 0) t += gdl_dterm(xv - __ldg(a.x + i - 1), yv - __ldg(a.y + i - 1), a.alpha);
      g = fmaf(a.c_gdl, t, g);
    }
    a.grad[i] = g;
    amax = fmaxf(amax, fabsf(g));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(a.amax_bits, __float_as_uint(amax));
}

// One CTA per (clip n, frame t), one thread per patch i (strided).  Feature (i, c) = v[t][n][c*hw + i].
__global__ void __launch_bounds__(256) loss_grad_nce_kernel(const __grid_constant__ LossGradArgs a) {
  extern __shared__ float4 nce_g_smem[];   // [2][hw]: pred, gt
  pdl_wait();
  pdl_trigger();
  const int hw = a.h * a.w;
  const int n = blockIdx.x % a.B, t = blockIdx.x / a.B;
  const size_t base = (static_cast<size_t>(t) * a.B + n) * 4 * hw;
  float4* sp = nce_g_smem;
  float4* sg = nce_g_smem + hw;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    sp[i] = make_float4(__ldg(a.x + base + i), __ldg(a.x + base + hw + i), __ldg(a.x + base + 2 * hw + i), __ldg(a.x + base + 3 * hw + i));
    sg[i] = make_float4(__ldg(a.y + base + i), __ldg(a.y + base + hw + i), __ldg(a.y + base + 2 * hw + i), __ldg(a.y + base + 3 * hw + i));
  }
  __syncthreads();
  const float it = a.inv_temperature;
  float amax = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const float4 pi = sp[i], gi = sg[i];
    // direction 1, row i: scores gt_i . pred_j / tau; only p1[i][i] is needed
    float m1 = -INFINITY, l1 = 0.f;
    // direction 2, row i: scores pred_i . gt_j / tau; softmax-weighted sum of gt_j (online rescaling)
    float m2 = -INFINITY, l2 = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < hw; ++j) {
      const float4 pj = sp[j], gj = sg[j];
      const float s1 = (gi.x * pj.x + gi.y * pj.y + gi.z * pj.z + gi.w * pj.w) * it;
      const float s2 = (pi.x * gj.x + pi.y * gj.y + pi.z * gj.z + pi.w * gj.w) * it;
      if (s1 > m1) { l1 *= __expf(m1 - s1); m1 = s1; }
      l1 += __expf(s1 - m1);
      if (s2 > m2) {
        const float f = __expf(m2 - s2);
        l2 *= f; acc.x *= f; acc.y *= f; acc.z *= f; acc.w *= f; m2 = s2;
      }
      const float e2 = __expf(s2 - m2);
      l2 += e2;
      acc.x = fmaf(e2, gj.x, acc.x); acc.y = fmaf(e2, gj.y, acc.y); acc.z = fmaf(e2, gj.z, acc.z); acc.w = fmaf(e2, gj.w, acc.w);
    }
    const float sii = (gi.x * pi.x + gi.y * pi.y + gi.z * pi.z + gi.w * pi.w) * it;
    const float p1 = __expf(sii - m1) / l1;
    const float inv2 = 1.0f / l2;
    const float k = a.c_nce * it;
    const float g0 = k * ((p1 - 2.0f) * gi.x + acc.x * inv2);
    const float g1 = k * ((p1 - 2.0f) * gi.y + acc.y * inv2);
    const float g2 = k * ((p1 - 2.0f) * gi.z + acc.z * inv2);
    const float g3 = k * ((p1 - 2.0f) * gi.w + acc.w * inv2);
    float* o = a.grad + base + i;
    const float r0 = o[0] + g0, r1 = o[hw] + g1, r2 = o[2 * hw] + g2, r3 = o[3 * hw] + g3;
    o[0] = r0; o[hw] = r1; o[2 * hw] = r2; o[3 * hw] = r3;
    amax = fmaxf(fmaxf(amax, fmaxf(fabsf(r0), fabsf(r1))), fmaxf(fabsf(r2), fabsf(r3)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(a.amax_bits, __float_as_uint(amax));
}

// max |x| of an upstream gradient handed in by the caller (torch.autograd bridge, sdvg_train_backward_from): feeds
// the same power-of-two loss scale as the fused criterion gradient kernels above.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, unsigned int* amax_bits) {
  float amax = 0.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) amax = fmaxf(amax, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.f && amax < INFINITY) atomicMax(amax_bits, __float_as_uint(amax));
}

// scale[0] = S = 2^k with S * amax in [32, 64), scale[1] = 1 / S; resets amax for the next step.
__global__ void loss_scale_kernel(unsigned int* amax_bits, float* scale) {
  pdl_wait();
  pdl_trigger();
  const float amax = __uint_as_float(*amax_bits);
  float s = 1.0f;
  if (amax > 0.f && amax < INFINITY) {
    int ex;
    frexpf(amax, &ex);            // amax = f * 2^ex, f in [0.5, 1)
    int k = 6 - ex;
    k = k > 100 ? 100 : (k < -100 ? -100 : k);
    s = ldexpf(1.0f, k);
  }
  scale[0] = s; scale[1] = 1.0f / s;
  *amax_bits = 0u;
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam(params, lr) as constructed at trainers/trainer.py:365: betas (0.9, 0.999), eps 1e-8, no weight
// decay, no amsgrad.  One flat pass over all parameters: 16 bytes read, 12 written per parameter.
struct AdamArgs {
  float* p; const float* g; float* m; float* v;
  long long n;
  float beta1, beta2, eps, step_size, inv_bc2_sqrt, gmul;   // step_size = lr / (1 - beta1^t); gmul = 1 / world size
};

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamArgs a) {
  pdl_wait();
  pdl_trigger();
  const long long n4 = a.n >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.g) + i);
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float* pp = &p.x; const float* gp = &g4.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g = gp[k] * a.gmul;
      mp[k] = mp[k] + (g - mp[k]) * (1.0f - a.beta1);
      vp[k] = vp[k] * a.beta2 + (1.0f - a.beta2) * g * g;
      const float denom = sqrtf(vp[k]) * a.inv_bc2_sqrt + a.eps;
      pp[k] = pp[k] - a.step_size * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
  }
}

}  // namespace sdvg
