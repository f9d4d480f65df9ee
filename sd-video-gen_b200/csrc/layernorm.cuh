// K3 - LayerNorm over the model width (post-norm layers of torch.nn.Transformer, eps 1e-5, biased variance),
// optionally chained with a second LayerNorm (last encoder layer: norm2 -> encoder.norm; last decoder layer:
// norm3 -> decoder.norm).  The residual add already happened in the producing GEMM's epilogue.
// One warp per token row held entirely in registers (two-pass mean / variance), 128-bit loads and stores;
// emits the fp32 residual stream and/or the 16-bit operand planes of the next GEMM.  HBM/L2-bound:
// 4 d bytes in, (4 + 2..4) d bytes out per row.
#pragma once
#include "common.cuh"

namespace sdvg {

struct LnArgs {
  const float* x; int ldx;
  int rows, d;
  const float* w1; const float* b1;
  const float* w2; const float* b2;  // nullptr: single LayerNorm
  float eps;
  // only rows with (row % rows_per_clip) >= first_token are processed; outputs are compacted to
  // (clip, token - first_token) rows when compact != 0 (last-token pruning of the final decoder layer)
  int rows_per_clip, first_token, compact;
  float* out32; int ld32;
  uint16_t* out_hi; uint16_t* out_lo; int ld16; int bf16;
  float2* stats;  // optional [rows]: (mean, rstd) of the first LayerNorm, indexed like the output rows (deferred residual)
};

template <int NV>
__device__ __forceinline__ float2 ln_inplace(float4 (&v)[NV], int d, int lane, const float* __restrict__ w,
                                             const float* __restrict__ b, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if ((i * 32 + lane) * 4 < d) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if ((i * 32 + lane) * 4 < d) {
      const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / static_cast<float>(d) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
      v[i].x = (v[i].x - mean) * rstd * ww.x + bb.x;
      v[i].y = (v[i].y - mean) * rstd * ww.y + bb.y;
      v[i].z = (v[i].z - mean) * rstd * ww.z + bb.z;
      v[i].w = (v[i].w - mean) * rstd * ww.w + bb.w;
    }
  }
  return make_float2(mean, rstd);
}

template <int NV>
__global__ void __launch_bounds__(128, NV <= 16 ? 4 : 2) layernorm_kernel(const __grid_constant__ LnArgs a) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);  // index among processed rows
  const int keep = a.rows_per_clip - a.first_token;
  const int clips = a.rows / a.rows_per_clip;
  if (r >= clips * keep) return;
  const int clip = r / keep, tok = a.first_token + (r - clip * keep);
  const size_t in_row = static_cast<size_t>(clip) * a.rows_per_clip + tok;
  const size_t out_row = a.compact ? static_cast<size_t>(r) : in_row;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    v[i] = c < a.d ? *reinterpret_cast<const float4*>(a.x + in_row * a.ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float2 st1 = ln_inplace<NV>(v, a.d, lane, a.w1, a.b1, a.eps);
  if (a.stats && lane == 0) a.stats[out_row] = st1;
  if (a.w2) ln_inplace<NV>(v, a.d, lane, a.w2, a.b2, a.eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c >= a.d) continue;
    if (a.out32) *reinterpret_cast<float4*>(a.out32 + out_row * a.ld32 + c) = v[i];
    if (a.out_hi) {
      const float f[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      uint16_t hi[4], lo[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) { hi[t] = to_plane_hi(f[t], a.bf16); lo[t] = to_plane_lo(f[t], hi[t]); }
      *reinterpret_cast<uint2*>(a.out_hi + out_row * a.ld16 + c) =
          make_uint2(hi[0] | (uint32_t(hi[1]) << 16), hi[2] | (uint32_t(hi[3]) << 16));
      if (a.out_lo)
        *reinterpret_cast<uint2*>(a.out_lo + out_row * a.ld16 + c) =
            make_uint2(lo[0] | (uint32_t(lo[1]) << 16), lo[2] | (uint32_t(lo[3]) << 16));
    }
  }
}

// Block-per-row variant for wide rows (d = 512 * NV4): 128 threads share one row, NV4 float4 per thread, all
// loads issued up front, two block reductions (mean, centred variance).  ncu r1b showed the warp-per-row kernel
// at d=2048 holding 64 values per lane in 128 registers: 14 resident warps per SM, ~2100 instructions per row,
// issue-bound at 48 % issue utilisation and ~4 TB/s.  This one keeps 16 values per thread (~40 registers).
__device__ __forceinline__ float block_sum_128(float v, float* red, int lane, int warp) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const float t = (red[0] + red[1]) + (red[2] + red[3]);
  __syncthreads();
  return t;
}

template <int NV4>
__global__ void __launch_bounds__(128) layernorm_block_kernel(const __grid_constant__ LnArgs a) {
  __shared__ float red[4];
  pdl_wait();
  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int keep = a.rows_per_clip - a.first_token;
  const int r = blockIdx.x;
  const int clip = r / keep, tok = a.first_token + (r - clip * keep);
  const size_t in_row = static_cast<size_t>(clip) * a.rows_per_clip + tok;
  const size_t out_row = a.compact ? static_cast<size_t>(r) : in_row;
  const float inv_d = 1.0f / static_cast<float>(a.d);
  float4 v[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) v[i] = *reinterpret_cast<const float4*>(a.x + in_row * a.ldx + (i * 128 + tid) * 4);
  const float* ws[2] = {a.w1, a.w2};
  const float* bs[2] = {a.b1, a.b2};
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && !a.w2) break;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = block_sum_128(s, red, lane, warp) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const float rstd = rsqrtf(block_sum_128(q, red, lane, warp) * inv_d + a.eps);
    if (pass == 0 && a.stats && tid == 0) a.stats[out_row] = make_float2(mean, rstd);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = (i * 128 + tid) * 4;
      const float4 ww = __ldg(reinterpret_cast<const float4*>(ws[pass] + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bs[pass] + c));
      v[i].x = (v[i].x - mean) * rstd * ww.x + bb.x;
      v[i].y = (v[i].y - mean) * rstd * ww.y + bb.y;
      v[i].z = (v[i].z - mean) * rstd * ww.z + bb.z;
      v[i].w = (v[i].w - mean) * rstd * ww.w + bb.w;
    }
  }
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = (i * 128 + tid) * 4;
    if (a.out32) *reinterpret_cast<float4*>(a.out32 + out_row * a.ld32 + c) = v[i];
    if (a.out_hi) {
      uint2 h;
      if (a.bf16) {
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[i].x, v[i].y), p1 = __floats2bfloat162_rn(v[i].z, v[i].w);
        h = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
      } else {
        const __half2 p0 = __floats2half2_rn(v[i].x, v[i].y), p1 = __floats2half2_rn(v[i].z, v[i].w);
        h = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
      }
      *reinterpret_cast<uint2*>(a.out_hi + out_row * a.ld16 + c) = h;
      if (a.out_lo) {
        const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
        const float2 h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
        const __half2 l0 = __floats2half2_rn((v[i].x - h0.x) * kSplitScale, (v[i].y - h0.y) * kSplitScale);
        const __half2 l1 = __floats2half2_rn((v[i].z - h1.x) * kSplitScale, (v[i].w - h1.y) * kSplitScale);
        *reinterpret_cast<uint2*>(a.out_lo + out_row * a.ld16 + c) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
      }
    }
  }
}

// ---- LayerNorm folded into the neighbouring GEMMs (Epilogue::ln_in / stat_out, common.cuh) ---------------------------
// Row statistics from the partial sums the producing GEMM's epilogue warps left behind: one thread per row adds its
// `slots` partials in a fixed order (deterministic), biased variance as E[y^2] - mean^2 (16-bit modes only: the rows are
// post-norm residual sums, |mean| <~ std, so the cancellation costs a few ulps of fp32).
__global__ void __launch_bounds__(256) ln_stats_finalize_kernel(const float2* __restrict__ part, int ld, int slots, int rows, float inv_d,
                                                                float eps, float2* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  // one warp per row: lane i takes slots i, i + 32, ...; fixed-shape butterfly sum (deterministic).  (First version: one
  // thread per row walking its 22 slots - 20 blocks of serialised L2 round trips, 28 us per launch.)
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float2* p = part + static_cast<size_t>(r) * ld;
  float s = 0.f, q = 0.f;
  for (int i = lane; i < slots; i += 32) { const float2 v = __ldcg(p + i); s += v.x; q += v.y; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if (lane == 0) {
    const float mean = s * inv_d;
    const float var = fmaxf(q * inv_d - mean * mean, 0.f);
    stats[r] = make_float2(mean, rsqrtf(var + eps));
  }
}
inline cudaError_t launch_ln_stats_finalize(const float2* part, int ld, int slots, int rows, int d, float eps, float2* stats,
                                            cudaStream_t stream) {
  return launch_kernel(ln_stats_finalize_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, stream, part, ld, slots, rows,
                       1.0f / static_cast<float>(d), eps, stats);
}

// Weight planes of a GEMM that consumes LayerNorm(y) directly from y: W' = W diag(gamma) in the operand format, with
// c[n] = sum_k W'[n][k] (of the ROUNDED plane values, so that the mean term cancels exactly as the tensor cores see it)
// and b'[n] = b[n] + sum_k beta[k] W[n][k].  One block per weight row; runs when the weights change, not per step.
__global__ void __launch_bounds__(128) ln_fold_weights_kernel(const float* __restrict__ w, int K, int ld16, const float* __restrict__ bias,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              uint16_t* __restrict__ hi, float* __restrict__ c, float* __restrict__ bfold,
                                                              int bf16) {
  __shared__ float red[8];
  const int n = blockIdx.x, tid = threadIdx.x;
  const float* wr = w + static_cast<size_t>(n) * K;
  uint16_t* hr = hi + static_cast<size_t>(n) * ld16;
  float cs = 0.f, bs = 0.f;
  for (int k = tid; k < ld16; k += 128) {
    uint16_t h = 0;
    if (k < K) {
      const float wv = wr[k];
      h = to_plane_hi(wv * gamma[k], bf16);
      cs += bf16 ? __bfloat162float(__ushort_as_bfloat16(h)) : __half2float(__ushort_as_half(h));
      bs = fmaf(beta[k], wv, bs);
    }
    hr[k] = h;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { cs += __shfl_xor_sync(0xffffffffu, cs, o); bs += __shfl_xor_sync(0xffffffffu, bs, o); }
  if ((tid & 31) == 0) { red[tid >> 5] = cs; red[4 + (tid >> 5)] = bs; }
  __syncthreads();
  if (tid == 0) {
    c[n] = (red[0] + red[1]) + (red[2] + red[3]);
    bfold[n] = bias[n] + ((red[4] + red[5]) + (red[6] + red[7]));
  }
}

inline cudaError_t launch_layernorm(const LnArgs& a, cudaStream_t stream) {
  if (a.d % 4 != 0 || a.d > 4096) return cudaErrorInvalidValue;
  const int keep = a.rows_per_clip - a.first_token;
  const int nrows = (a.rows / a.rows_per_clip) * keep;
  if (nrows == 0) return cudaSuccess;
  if (a.d % 512 == 0 && a.ldx % 4 == 0) {
    switch (a.d / 512) {
      case 1: return launch_kernel(layernorm_block_kernel<1>, dim3(nrows), dim3(128), 0, stream, a);
      case 2: return launch_kernel(layernorm_block_kernel<2>, dim3(nrows), dim3(128), 0, stream, a);
      case 3: return launch_kernel(layernorm_block_kernel<3>, dim3(nrows), dim3(128), 0, stream, a);
      case 4: return launch_kernel(layernorm_block_kernel<4>, dim3(nrows), dim3(128), 0, stream, a);
      case 6: return launch_kernel(layernorm_block_kernel<6>, dim3(nrows), dim3(128), 0, stream, a);
      case 8: return launch_kernel(layernorm_block_kernel<8>, dim3(nrows), dim3(128), 0, stream, a);
      default: break;
    }
  }
  const int grid = ceil_div(nrows, 4);
  const int nv = ceil_div(a.d, 128);
  if (nv <= 1) return launch_kernel(layernorm_kernel<1>, dim3(grid), dim3(128), 0, stream, a);
  else if (nv <= 2) return launch_kernel(layernorm_kernel<2>, dim3(grid), dim3(128), 0, stream, a);
  else if (nv <= 4) return launch_kernel(layernorm_kernel<4>, dim3(grid), dim3(128), 0, stream, a);
  else if (nv <= 8) return launch_kernel(layernorm_kernel<8>, dim3(grid), dim3(128), 0, stream, a);
  else if (nv <= 16) return launch_kernel(layernorm_kernel<16>, dim3(grid), dim3(128), 0, stream, a);
  else return launch_kernel(layernorm_kernel<32>, dim3(grid), dim3(128), 0, stream, a);
}

}  // namespace sdvg
