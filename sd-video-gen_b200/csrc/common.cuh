// Shared types for the sdvg kernels: the GEMM epilogue description and 16-bit plane helpers.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace sdvg {

// Split-operand scale: x ~= hi + lo * 2^-11 with hi, lo fp16 (22 significant bits).
constexpr float kSplitScale = 2048.0f;
constexpr float kSplitInv = 1.0f / 2048.0f;

// What every GEMM does to its fp32 accumulator before storing (in this order):
//   v = acc + bias[col];  v *= alpha;  v += pe[pe_index[row / rows_per_clip]][col];  relu;  v += residual[row][col]
// and where the result goes: an fp32 matrix, and/or 16-bit operand planes (hi and, for split consumers, lo)
// that the next GEMM reads through TMA.
struct Epilogue {
  const float* bias = nullptr;      // [N]
  float alpha = 1.0f;
  const float* pe = nullptr;        // [64][ld_pe] positional table (models/positional_encoding.py:17-31)
  int ld_pe = 0;
  const int* pe_index = nullptr;    // [clips]; nullptr -> clip index itself
  int rows_per_clip = 1;
  int relu = 0;
  const float* residual = nullptr;  // [M][ld_res]
  int ld_res = 0;
  float* out32 = nullptr;           // fp32 output
  int ld32 = 0;
  uint16_t* out_hi = nullptr;       // 16-bit planes [M][ld16]
  uint16_t* out_lo = nullptr;
  int ld16 = 0;
  int bf16 = 0;                     // hi plane is bf16 instead of fp16 (lo plane is always fp16)
  // row mapping of the fp32 output (planes always use identity):
  //   0 identity;  1 clip-major (b*S+s) -> sequence-major (s*B+b), the reference's (S,B,E) return layout
  //   (models/transformer.py:60-68);  2 only the last token of every clip is written, to row b
  //   (prediction/predict.py:42 keeps pred[:, -1]).
  //   3 clip-strided: token (clip b, position s) -> row b*out_clip_rows + out_row_off + s (token-local caches
  //   of the rollout, indexed by history slot)
  int row_map = 0;
  int clips = 0;
  int out_clip_rows = 0, out_row_off = 0;
  // residual read with the same clip-strided addressing when res_clip_rows > 0
  int res_clip_rows = 0, res_row_off = 0;
  // deferred LayerNorm of the residual: when ln_stats != nullptr, `residual` holds the PRE-norm sums y and the
  // value added is (y - mean[row]) * rstd[row] * ln_w[col] + ln_b[col], i.e. LayerNorm(y) recomputed on the fly
  // from the row statistics the LayerNorm kernel left behind (it then does not have to write its fp32 output)
  const float2* ln_stats = nullptr;
  const float* ln_w = nullptr;
  const float* ln_b = nullptr;
  // LayerNorm folded into the neighbouring GEMMs (large batches, 16-bit modes; Engine::fold_*):
  //  * consumer (ln_in != 0): the A planes hold the PRE-norm sums y and the weight planes are W diag(gamma), so
  //      acc = sum_k y_k gamma_k W_nk  and  LayerNorm(y) W^T + b = rstd[row] * (acc - mean[row] * c[col]) + b'[col]
  //    with ln_stats = (mean, rstd) per row, ln_w = c (row sums of the folded planes), bias = b' = b + W beta;
  //  * producer (stat_out != nullptr): besides its outputs, every epilogue warp leaves the (sum, sum of squares) of
  //    the values it stored, per row and column slot: stat_out[row * stat_ld + slot] - summed in fixed slot order by
  //    ln_stats_finalize_kernel (deterministic), so no kernel ever reads the row again to normalise it.
  int ln_in = 0;
  float2* stat_out = nullptr;
  int stat_ld = 0;
  //  * small batches (stat_in != nullptr, together with ln_stats as the mode flag): no kernel turned the partials into
  //    (mean, rstd) - every epilogue warp sums the stat_in_n slots of its 32 rows itself while its tile's main loop
  //    runs (one launch less per LayerNorm where a launch is all latency; at large batch the 3 us kernel is cheaper
  //    than the strided slot reads of every tile, profiles/README.md).  mean = sum * stat_inv_d,
  //    rstd = rsqrt(max(sumsq * stat_inv_d - mean^2, 0) + stat_eps), as ln_stats_finalize_kernel.
  const float2* stat_in = nullptr;
  int stat_in_ld = 0, stat_in_n = 0;
  float stat_inv_d = 0.f, stat_eps = 0.f;
  // training (backward GEMMs): alpha_dev multiplies like alpha but is read from device memory (1 / loss scale);
  // gate zeroes the value where gate[row][col] <= 0 (ReLU backward against the saved activation), before the residual
  const float* alpha_dev = nullptr;
  const float* gate = nullptr;
  int ld_gate = 0;
  float gate_scale = 1.0f;   // kept values are multiplied by this (1 / (1 - p) when the gate also carries a dropout mask)
  // training-mode dropout (after ReLU, before the residual): element (row, col) is kept iff
  // drop_hash(drop_key, row * drop_cols + col) >= drop_thr, and then multiplied by drop_scale = 1 / (1 - p)
  uint32_t drop_thr = 0;
  uint32_t drop_key = 0;
  int drop_cols = 0;
  float drop_scale = 1.0f;
};

// Counter-based dropout mask (murmur3 finaliser of a per-site key and the element index): the same mask is
// regenerated in the backward pass, and the CPU checker of the test-suite restates it bit for bit for the parity tests.
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t key, uint32_t idx) { return fmix32(idx * 0x9E3779B1u + key); }
__host__ __device__ __forceinline__ uint32_t drop_site_key(uint64_t seed, uint32_t step, uint32_t site) {
  return fmix32(static_cast<uint32_t>(seed) ^ fmix32(site * 0x632BE5ABu + step)) ^ static_cast<uint32_t>(seed >> 32);
}

__device__ __forceinline__ int epi_out_row(const Epilogue& e, int row) {
  if (e.row_map == 0) return row;
  const int b = row / e.rows_per_clip, s = row - b * e.rows_per_clip;
  if (e.row_map == 1) return s * e.clips + b;
  if (e.row_map == 3) return b * e.out_clip_rows + e.out_row_off + s;
  return (s == e.rows_per_clip - 1) ? b : -1;
}

__device__ __forceinline__ int epi_res_row(const Epilogue& e, int row) {
  if (e.res_clip_rows == 0) return row;
  const int b = row / e.rows_per_clip, s = row - b * e.rows_per_clip;
  return b * e.res_clip_rows + e.res_row_off + s;
}

// Row-dependent part resolved once per row: which PE row this token's clip uses (-1: none).
__device__ __forceinline__ int epi_pe_row(const Epilogue& e, int row) {
  if (!e.pe) return -1;
  const int b = row / e.rows_per_clip;
  return e.pe_index ? __ldg(e.pe_index + b) : b;
}

__device__ __forceinline__ float epi_value(const Epilogue& e, float acc, int row, int col, int pe_row) {
  float v = acc;
  if (e.ln_in) {
    const float2 st = __ldg(e.ln_stats + row);
    v = st.y * (v - st.x * __ldg(e.ln_w + col));
  }
  if (e.bias) v += __ldg(e.bias + col);
  v *= e.alpha;
  if (e.alpha_dev) v *= __ldg(e.alpha_dev);
  if (e.gate) v = __ldg(e.gate + static_cast<size_t>(row) * e.ld_gate + col) > 0.0f ? v * e.gate_scale : 0.0f;
  if (pe_row >= 0) v += __ldg(e.pe + static_cast<size_t>(pe_row) * e.ld_pe + col);
  if (e.relu) v = fmaxf(v, 0.0f);
  if (e.drop_thr)
    v = drop_hash(e.drop_key, static_cast<uint32_t>(row) * static_cast<uint32_t>(e.drop_cols) + static_cast<uint32_t>(col)) >= e.drop_thr
            ? v * e.drop_scale : 0.0f;
  if (e.residual) {
    const int rrow = epi_res_row(e, row);
    float r = __ldg(e.residual + static_cast<size_t>(rrow) * e.ld_res + col);
    if (e.ln_stats) {
      const float2 st = __ldg(e.ln_stats + rrow);
      r = (r - st.x) * st.y * __ldg(e.ln_w + col) + __ldg(e.ln_b + col);
    }
    v += r;
  }
  return v;
}

// 16-bit plane conversions ---------------------------------------------------------------------------
__device__ __forceinline__ uint16_t to_plane_hi(float v, int bf16) {
  if (bf16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(v));
}
// lo plane (split mode, fp16 only): lo = fp16((v - hi) * 2^11)
__device__ __forceinline__ uint16_t to_plane_lo(float v, uint16_t hi) {
  const float r = (v - __half2float(__ushort_as_half(hi))) * kSplitScale;
  return __half_as_ushort(__float2half_rn(r));
}

// Store one value to all requested destinations (scalar path).
__device__ __forceinline__ void epi_store(const Epilogue& e, float v, int row, int out_row, int col) {
  if (e.out32 && out_row >= 0) e.out32[static_cast<size_t>(out_row) * e.ld32 + col] = v;
  if (e.out_hi) {
    const uint16_t h = to_plane_hi(v, e.bf16);
    e.out_hi[static_cast<size_t>(row) * e.ld16 + col] = h;
    if (e.out_lo) e.out_lo[static_cast<size_t>(row) * e.ld16 + col] = to_plane_lo(v, h);
  }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The rollout is a chain of ~130 dependent kernels per model pass, many of
// them 10-50 us long, so the kernel-boundary bubble (launch latency + drain + the next kernel's prologue:
// mbarrier init, TMEM allocation, tensor-map prefetch, cluster sync) is a measurable share of the step.  Every
// kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization; it runs its prologue, then
// pdl_wait() blocks until the previous kernel has completed and flushed (so all data hazards are exactly those
// of plain stream order), then pdl_trigger() lets the next kernel's CTAs be scheduled as SM resources free up.
inline bool& pdl_enabled() { static bool on = true; return on; }

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace sdvg
