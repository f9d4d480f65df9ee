// K2 - fused multi-head attention for very short sequences (S <= 32 tokens; the path uses 5, 6 or 10).
// softmax(Q K^T / sqrt(hd) + mask) V per (clip, head), i.e. the scaled_dot_product_attention inside
// torch.nn.MultiheadAttention as called by the layers constructed at models/transformer.py:38-44.
//
// One warp per (clip, head): the head dimension is spread over the lanes (128-bit loads when hd % 128 == 0),
// scores are lane-partial dot products reduced with shuffles, softmax runs in registers in fp32 (the layer-0
// scores are near one-hot because emb*sqrt(d) is un-normalised - SURVEY.md fact 6 - so nothing here is 16-bit),
// K/V rows are re-read from L1.  HBM-bound: bytes = (Sq + 2 Sk) hd 4 in, Sq hd (2..8) out per (clip, head).
// Mask kinds: 0 none, 1 causal (key <= query; models/transformer.py:70-89 without materialising the matrix),
// 2 additive fp32 (Sq x Sk) as passed to forward(..., tgt_mask).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace sdvg {

constexpr int kAttnMaxS = 32;

struct AttnArgs {
  const float* q; int ldq;
  const float* k; const float* v; int ldkv;
  long long q_clip_stride, kv_clip_stride;  // elements between consecutive clips (0: Sq*ldq / Sk*ldkv, i.e. packed rows)
  int clips, heads, hd, Sq, Sk;
  int mask_kind; const float* mask;
  float scale;
  int q_first;        // first query row to compute (Sq-1 when only the last token is needed)
  int out_compact;    // 1: output rows are packed per clip from q_first on, i.e. row = b*(Sq-q_first) + (i-q_first)
  float* out32; int ld32;
  uint16_t* out_hi; uint16_t* out_lo; int ld16; int bf16;
};

// Lane-distributed row fragment: element index of (chunk c, lane, v) is (c*32 + lane)*VEC + v.
template <int VEC, int NCH>
__device__ __forceinline__ void load_frag(float (&f)[NCH * VEC], const float* __restrict__ row, int hd, int lane) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int e0 = (c * 32 + lane) * VEC;
    if constexpr (VEC == 4) {
      const float4 t = e0 < hd ? __ldg(reinterpret_cast<const float4*>(row + e0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      f[c * 4 + 0] = t.x; f[c * 4 + 1] = t.y; f[c * 4 + 2] = t.z; f[c * 4 + 3] = t.w;
    } else if constexpr (VEC == 2) {
      const float2 t = e0 < hd ? __ldg(reinterpret_cast<const float2*>(row + e0)) : make_float2(0.f, 0.f);
      f[c * 2 + 0] = t.x; f[c * 2 + 1] = t.y;
    } else {
      f[c] = e0 < hd ? __ldg(row + e0) : 0.f;
    }
  }
}

template <int VEC, int NCH>
__global__ void __launch_bounds__(128) attention_kernel(const __grid_constant__ AttnArgs a) {
  constexpr int EPL = VEC * NCH;
  pdl_wait();
  pdl_trigger();
  const int warp_global = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp_global >= a.clips * a.heads) return;
  const int b = warp_global / a.heads, h = warp_global - b * a.heads;
  const int hd = a.hd;
  const float* qb = a.q + static_cast<size_t>(b) * a.q_clip_stride + h * hd;
  const float* kb = a.k + static_cast<size_t>(b) * a.kv_clip_stride + h * hd;
  const float* vb = a.v + static_cast<size_t>(b) * a.kv_clip_stride + h * hd;
  constexpr float kLog2e = 1.4426950408889634f;

  for (int i = a.q_first; i < a.Sq; ++i) {
    float ql[EPL];
    load_frag<VEC, NCH>(ql, qb + static_cast<size_t>(i) * a.ldq, hd, lane);
    float sc[kAttnMaxS];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kAttnMaxS; ++j) {
      sc[j] = -INFINITY;
      if (j < a.Sk) {
        float kl[EPL];
        load_frag<VEC, NCH>(kl, kb + static_cast<size_t>(j) * a.ldkv, hd, lane);
        float part = 0.f;
#pragma unroll
        for (int t = 0; t < EPL; ++t) part = fmaf(ql[t], kl[t], part);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        float s = part * a.scale;
        if (a.mask_kind == 1) { if (j > i + (a.Sk - a.Sq)) s = -INFINITY; }
        else if (a.mask_kind == 2) s += __ldg(a.mask + i * a.Sk + j);
        sc[j] = s;
        mx = fmaxf(mx, s);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kAttnMaxS; ++j) {
      if (j < a.Sk) { sc[j] = exp2f((sc[j] - mx) * kLog2e); sum += sc[j]; }
    }
    const float inv = 1.0f / sum;
    float ol[EPL];
#pragma unroll
    for (int t = 0; t < EPL; ++t) ol[t] = 0.f;
#pragma unroll
    for (int j = 0; j < kAttnMaxS; ++j) {
      if (j < a.Sk) {
        const float p = sc[j] * inv;
        float vl[EPL];
        load_frag<VEC, NCH>(vl, vb + static_cast<size_t>(j) * a.ldkv, hd, lane);
#pragma unroll
        for (int t = 0; t < EPL; ++t) ol[t] = fmaf(p, vl[t], ol[t]);
      }
    }
    const size_t row = a.out_compact ? static_cast<size_t>(b) * (a.Sq - a.q_first) + (i - a.q_first) : static_cast<size_t>(b) * a.Sq + i;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int e0 = (c * 32 + lane) * VEC;
      if (e0 >= hd) continue;
      const int col = h * hd + e0;
      if constexpr (VEC == 4) {
        if (a.out32)
          *reinterpret_cast<float4*>(a.out32 + row * a.ld32 + col) =
              make_float4(ol[c * 4], ol[c * 4 + 1], ol[c * 4 + 2], ol[c * 4 + 3]);
        if (a.out_hi) {
          uint16_t hi[4], lo[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) { hi[v] = to_plane_hi(ol[c * 4 + v], a.bf16); lo[v] = to_plane_lo(ol[c * 4 + v], hi[v]); }
          *reinterpret_cast<uint2*>(a.out_hi + row * a.ld16 + col) =
              make_uint2(hi[0] | (uint32_t(hi[1]) << 16), hi[2] | (uint32_t(hi[3]) << 16));
          if (a.out_lo)
            *reinterpret_cast<uint2*>(a.out_lo + row * a.ld16 + col) =
                make_uint2(lo[0] | (uint32_t(lo[1]) << 16), lo[2] | (uint32_t(lo[3]) << 16));
        }
      } else {
        const float val = ol[c];
        if (a.out32) a.out32[row * a.ld32 + col] = val;
        if (a.out_hi) {
          const uint16_t hi = to_plane_hi(val, a.bf16);
          a.out_hi[row * a.ld16 + col] = hi;
          if (a.out_lo) a.out_lo[row * a.ld16 + col] = to_plane_lo(val, hi);
        }
      }
    }
  }
}

// Fast path (hd == 32 * VEC * NCH exactly, Sk <= SMAX): K and V of the (clip, head) are loaded ONCE into
// registers with all 128-bit loads in flight together, then every query row is scored from registers.
// The generic kernel above re-reads K/V per query row through L1 and is latency bound (measured 1.46 TB/s
// algorithmic on B200 at S=5, hd=256); this one issues 2*Sk*NCH independent loads per lane up front.
template <int VEC, int NCH, int SMAX>
__global__ void __launch_bounds__(128) attention_reg_kernel(const __grid_constant__ AttnArgs a) {
  constexpr int EPL = VEC * NCH;
  constexpr float kLog2e = 1.4426950408889634f;
  pdl_wait();
  pdl_trigger();
  const int warp_global = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp_global >= a.clips * a.heads) return;
  const int b = warp_global / a.heads, h = warp_global - b * a.heads;
  const int hd = a.hd;
  const float* qb = a.q + static_cast<size_t>(b) * a.q_clip_stride + h * hd;
  const float* kb = a.k + static_cast<size_t>(b) * a.kv_clip_stride + h * hd;
  const float* vb = a.v + static_cast<size_t>(b) * a.kv_clip_stride + h * hd;
  float kr[SMAX][EPL], vr[SMAX][EPL];
#pragma unroll
  for (int j = 0; j < SMAX; ++j) {
    if (j < a.Sk) {
      load_frag<VEC, NCH>(kr[j], kb + static_cast<size_t>(j) * a.ldkv, hd, lane);
      load_frag<VEC, NCH>(vr[j], vb + static_cast<size_t>(j) * a.ldkv, hd, lane);
    }
  }
  const float scale2 = a.scale * kLog2e;  // softmax in base 2: exp(x) = exp2(x * log2 e)
  for (int i = a.q_first; i < a.Sq; ++i) {
    float ql[EPL];
    load_frag<VEC, NCH>(ql, qb + static_cast<size_t>(i) * a.ldq, hd, lane);
    float sc[SMAX];
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
      float part = 0.f;
      if (j < a.Sk) {
#pragma unroll
        for (int t = 0; t < EPL; ++t) part = fmaf(ql[t], kr[j][t], part);
      }
      sc[j] = part;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < SMAX; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], o);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
      float s = sc[j] * scale2;
      if (j >= a.Sk) s = -INFINITY;
      else if (a.mask_kind == 1) { if (j > i + (a.Sk - a.Sq)) s = -INFINITY; }
      else if (a.mask_kind == 2) s += __ldg(a.mask + i * a.Sk + j) * kLog2e;
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
      sc[j] = (j < a.Sk) ? exp2f(sc[j] - mx) : 0.f;
      sum += sc[j];
    }
    const float inv = 1.0f / sum;
    float ol[EPL];
#pragma unroll
    for (int t = 0; t < EPL; ++t) ol[t] = 0.f;
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
      if (j < a.Sk) {
        const float p = sc[j] * inv;
#pragma unroll
        for (int t = 0; t < EPL; ++t) ol[t] = fmaf(p, vr[j][t], ol[t]);
      }
    }
    const size_t row = a.out_compact ? static_cast<size_t>(b) * (a.Sq - a.q_first) + (i - a.q_first) : static_cast<size_t>(b) * a.Sq + i;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = h * hd + (c * 32 + lane) * VEC;
      if (a.out32) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) a.out32[row * a.ld32 + col + v] = ol[c * VEC + v];
      }
      if (a.out_hi) {
        uint16_t hi[VEC], lo[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { hi[v] = to_plane_hi(ol[c * VEC + v], a.bf16); lo[v] = to_plane_lo(ol[c * VEC + v], hi[v]); }
        if constexpr (VEC == 4) {
          *reinterpret_cast<uint2*>(a.out_hi + row * a.ld16 + col) = make_uint2(hi[0] | (uint32_t(hi[1]) << 16), hi[2] | (uint32_t(hi[3]) << 16));
          if (a.out_lo) *reinterpret_cast<uint2*>(a.out_lo + row * a.ld16 + col) = make_uint2(lo[0] | (uint32_t(lo[1]) << 16), lo[2] | (uint32_t(lo[3]) << 16));
        } else if constexpr (VEC == 2) {
          *reinterpret_cast<uint32_t*>(a.out_hi + row * a.ld16 + col) = hi[0] | (uint32_t(hi[1]) << 16);
          if (a.out_lo) *reinterpret_cast<uint32_t*>(a.out_lo + row * a.ld16 + col) = lo[0] | (uint32_t(lo[1]) << 16);
        } else {
          a.out_hi[row * a.ld16 + col] = hi[0];
          if (a.out_lo) a.out_lo[row * a.ld16 + col] = lo[0];
        }
      }
    }
  }
}

// Exact-shape variant of the fast path: Sq, Sk and the mask kind are template parameters, so every loop is fully
// unrolled without predicates (ncu r1b: the predicated SMAX version executed ~2560 instructions per (clip, head)
// and was issue-bound at 46 % issue utilisation and 2.8 TB/s; the shapes the path actually uses are few:
// 5/6/10 tokens, no mask or causal).
template <int VEC, int NCH, int SQ, int SK, int MASK>
__global__ void __launch_bounds__(128) attention_exact_kernel(const __grid_constant__ AttnArgs a) {
  constexpr int EPL = VEC * NCH;
  constexpr float kLog2e = 1.4426950408889634f;
  pdl_wait();
  pdl_trigger();
  const int warp_global = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp_global >= a.clips * a.heads) return;
  const int b = warp_global / a.heads, h = warp_global - b * a.heads;
  constexpr int hd = 32 * EPL;
  const float* qb = a.q + static_cast<size_t>(b) * a.q_clip_stride + h * hd;
  const float* kb = a.k + static_cast<size_t>(b) * a.kv_clip_stride + h * hd;
  const float* vb = a.v + static_cast<size_t>(b) * a.kv_clip_stride + h * hd;
  float kr[SK][EPL], vr[SK][EPL];
#pragma unroll
  for (int j = 0; j < SK; ++j) load_frag<VEC, NCH>(kr[j], kb + static_cast<size_t>(j) * a.ldkv, hd, lane);
#pragma unroll
  for (int j = 0; j < SK; ++j) load_frag<VEC, NCH>(vr[j], vb + static_cast<size_t>(j) * a.ldkv, hd, lane);
  const float scale2 = a.scale * kLog2e;
  // the query row of iteration i+1 is loaded while row i is scored: a load issued and consumed inside one
  // iteration costs a full L2 round trip per row (measured: 37 us per launch at S=5, all of it latency)
  float qn[EPL];
  load_frag<VEC, NCH>(qn, qb + static_cast<size_t>(a.q_first) * a.ldq, hd, lane);
#pragma unroll 1
  for (int i = a.q_first; i < SQ; ++i) {
    float ql[EPL];
#pragma unroll
    for (int t = 0; t < EPL; ++t) ql[t] = qn[t];
    if (i + 1 < SQ) load_frag<VEC, NCH>(qn, qb + static_cast<size_t>(i + 1) * a.ldq, hd, lane);
    float sc[SK];
#pragma unroll
    for (int j = 0; j < SK; ++j) {
      float part = 0.f;
#pragma unroll
      for (int t = 0; t < EPL; ++t) part = fmaf(ql[t], kr[j][t], part);
      sc[j] = part;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < SK; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], o);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SK; ++j) {
      float s = sc[j] * scale2;
      if (MASK == 1 && j > i + (SK - SQ)) s = -INFINITY;
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < SK; ++j) { sc[j] = exp2f(sc[j] - mx); sum += sc[j]; }
    const float inv = 1.0f / sum;
    float ol[EPL];
#pragma unroll
    for (int t = 0; t < EPL; ++t) ol[t] = 0.f;
#pragma unroll
    for (int j = 0; j < SK; ++j) {
      const float p = sc[j] * inv;
#pragma unroll
      for (int t = 0; t < EPL; ++t) ol[t] = fmaf(p, vr[j][t], ol[t]);
    }
    const size_t row = a.out_compact ? static_cast<size_t>(b) * (SQ - a.q_first) + (i - a.q_first) : static_cast<size_t>(b) * SQ + i;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = h * hd + (c * 32 + lane) * VEC;
      if (a.out32) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) a.out32[row * a.ld32 + col + v] = ol[c * VEC + v];
      }
      if (a.out_hi) {
        uint16_t hi[VEC], lo[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { hi[v] = to_plane_hi(ol[c * VEC + v], a.bf16); lo[v] = to_plane_lo(ol[c * VEC + v], hi[v]); }
        if constexpr (VEC == 4) {
          *reinterpret_cast<uint2*>(a.out_hi + row * a.ld16 + col) = make_uint2(hi[0] | (uint32_t(hi[1]) << 16), hi[2] | (uint32_t(hi[3]) << 16));
          if (a.out_lo) *reinterpret_cast<uint2*>(a.out_lo + row * a.ld16 + col) = make_uint2(lo[0] | (uint32_t(lo[1]) << 16), lo[2] | (uint32_t(lo[3]) << 16));
        } else if constexpr (VEC == 2) {
          *reinterpret_cast<uint32_t*>(a.out_hi + row * a.ld16 + col) = hi[0] | (uint32_t(hi[1]) << 16);
          if (a.out_lo) *reinterpret_cast<uint32_t*>(a.out_lo + row * a.ld16 + col) = lo[0] | (uint32_t(lo[1]) << 16);
        } else {
          a.out_hi[row * a.ld16 + col] = hi[0];
          if (a.out_lo) a.out_lo[row * a.ld16 + col] = lo[0];
        }
      }
    }
  }
}

template <int VEC, int NCH, int SQ, int SK>
inline cudaError_t launch_attention_exact_m(const AttnArgs& a, int grid, cudaStream_t stream) {
  if (a.mask_kind == 0) return launch_kernel(attention_exact_kernel<VEC, NCH, SQ, SK, 0>, dim3(grid), dim3(128), 0, stream, a);
  if (a.mask_kind == 1) return launch_kernel(attention_exact_kernel<VEC, NCH, SQ, SK, 1>, dim3(grid), dim3(128), 0, stream, a);
  return cudaErrorNotSupported;
}
template <int VEC, int NCH>
inline cudaError_t launch_attention_exact(const AttnArgs& a, int grid, cudaStream_t stream) {
  if (a.Sq == 5 && a.Sk == 5) return launch_attention_exact_m<VEC, NCH, 5, 5>(a, grid, stream);
  if (a.Sq == 6 && a.Sk == 6) return launch_attention_exact_m<VEC, NCH, 6, 6>(a, grid, stream);
  if (a.Sq == 10 && a.Sk == 10) return launch_attention_exact_m<VEC, NCH, 10, 10>(a, grid, stream);
  if (a.Sq == 5 && a.Sk == 6) return launch_attention_exact_m<VEC, NCH, 5, 6>(a, grid, stream);
  if (a.Sq == 1 && a.Sk == 5) return launch_attention_exact_m<VEC, NCH, 1, 5>(a, grid, stream);
  if (a.Sq == 1 && a.Sk == 6) return launch_attention_exact_m<VEC, NCH, 1, 6>(a, grid, stream);
  if (a.Sq == 1 && a.Sk == 10) return launch_attention_exact_m<VEC, NCH, 1, 10>(a, grid, stream);
  return cudaErrorNotSupported;
}

// returns cudaErrorNotSupported when Sk is beyond the register-resident variants (caller falls back)
template <int VEC, int NCH>
inline cudaError_t launch_attention_reg(const AttnArgs& a, int grid, cudaStream_t stream) {
  if (a.Sk <= 6) return launch_kernel(attention_reg_kernel<VEC, NCH, 6>, dim3(grid), dim3(128), 0, stream, a);
  if (a.Sk <= 10) return launch_kernel(attention_reg_kernel<VEC, NCH, 10>, dim3(grid), dim3(128), 0, stream, a);
  return cudaErrorNotSupported;
}

// 16-bit-input variant (fp16 / bf16 Q, K, V written by the projection GEMM's epilogue as operand planes): used by
// the 16-bit precision modes for every attention except layer 0, whose un-normalised inputs need fp32 scores
// (SURVEY.md fact 6).  Halves the bytes of the bandwidth-bound attention kernels and of the QKV epilogue stores;
// measured effect on the C1 golden rollout: 8.3e-4 -> 1.05e-3 per-frame max-rel (oracle emulation), bar 5e-3.
// Lane l owns head elements [l*VEC, (l+1)*VEC): one 16-byte load per row at hd = 256.
template <int VEC>
__device__ __forceinline__ void load_frag16(float (&f)[VEC], const uint16_t* __restrict__ row, int lane, int bf16) {
  const uint16_t* p = row + lane * VEC;
  uint32_t w[(VEC + 1) / 2];
  if constexpr (VEC == 8) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
  } else if constexpr (VEC == 4) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    w[0] = t.x; w[1] = t.y;
  } else if constexpr (VEC == 2) {
    w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
  } else {
    w[0] = __ldg(p);
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const uint16_t h = static_cast<uint16_t>(i & 1 ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu));
    f[i] = bf16 ? __bfloat162float(__ushort_as_bfloat16(h)) : __half2float(__ushort_as_half(h));
  }
}

template <int VEC, int SQ, int SK, int MASK>
__global__ void __launch_bounds__(128) attention_exact16_kernel(const __grid_constant__ AttnArgs a) {
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr int hd = 32 * VEC;
  pdl_wait();
  pdl_trigger();
  const int warp_global = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp_global >= a.clips * a.heads) return;
  const int b = warp_global / a.heads, h = warp_global - b * a.heads;
  const uint16_t* qb = reinterpret_cast<const uint16_t*>(a.q) + static_cast<size_t>(b) * a.q_clip_stride + h * hd;
  const uint16_t* kb = reinterpret_cast<const uint16_t*>(a.k) + static_cast<size_t>(b) * a.kv_clip_stride + h * hd;
  const uint16_t* vb = reinterpret_cast<const uint16_t*>(a.v) + static_cast<size_t>(b) * a.kv_clip_stride + h * hd;
  float kr[SK][VEC], vr[SK][VEC];
#pragma unroll
  for (int j = 0; j < SK; ++j) load_frag16<VEC>(kr[j], kb + static_cast<size_t>(j) * a.ldkv, lane, a.bf16);
#pragma unroll
  for (int j = 0; j < SK; ++j) load_frag16<VEC>(vr[j], vb + static_cast<size_t>(j) * a.ldkv, lane, a.bf16);
  const float scale2 = a.scale * kLog2e;
  float qn[VEC];  // next query row in flight while the current one is scored (see attention_exact_kernel)
  load_frag16<VEC>(qn, qb + static_cast<size_t>(a.q_first) * a.ldq, lane, a.bf16);
#pragma unroll 1
  for (int i = a.q_first; i < SQ; ++i) {
    float ql[VEC];
#pragma unroll
    for (int t = 0; t < VEC; ++t) ql[t] = qn[t];
    if (i + 1 < SQ) load_frag16<VEC>(qn, qb + static_cast<size_t>(i + 1) * a.ldq, lane, a.bf16);
    float sc[SK];
#pragma unroll
    for (int j = 0; j < SK; ++j) {
      float part = 0.f;
#pragma unroll
      for (int t = 0; t < VEC; ++t) part = fmaf(ql[t], kr[j][t], part);
      sc[j] = part;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < SK; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], o);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SK; ++j) {
      float sv = sc[j] * scale2;
      if (MASK == 1 && j > i + (SK - SQ)) sv = -INFINITY;
      sc[j] = sv;
      mx = fmaxf(mx, sv);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < SK; ++j) { sc[j] = exp2f(sc[j] - mx); sum += sc[j]; }
    const float inv = 1.0f / sum;
    float ol[VEC];
#pragma unroll
    for (int t = 0; t < VEC; ++t) ol[t] = 0.f;
#pragma unroll
    for (int j = 0; j < SK; ++j) {
      const float pj = sc[j] * inv;
#pragma unroll
      for (int t = 0; t < VEC; ++t) ol[t] = fmaf(pj, vr[j][t], ol[t]);
    }
    const size_t row = a.out_compact ? static_cast<size_t>(b) * (SQ - a.q_first) + (i - a.q_first) : static_cast<size_t>(b) * SQ + i;
    const int col = h * hd + lane * VEC;
    if (a.out32) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) a.out32[row * a.ld32 + col + v] = ol[v];
    }
    if (a.out_hi) {
      uint16_t hi[VEC], lo[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) { hi[v] = to_plane_hi(ol[v], a.bf16); lo[v] = to_plane_lo(ol[v], hi[v]); }
      if constexpr (VEC == 8) {
        *reinterpret_cast<uint4*>(a.out_hi + row * a.ld16 + col) =
            make_uint4(hi[0] | (uint32_t(hi[1]) << 16), hi[2] | (uint32_t(hi[3]) << 16), hi[4] | (uint32_t(hi[5]) << 16), hi[6] | (uint32_t(hi[7]) << 16));
        if (a.out_lo)
          *reinterpret_cast<uint4*>(a.out_lo + row * a.ld16 + col) =
              make_uint4(lo[0] | (uint32_t(lo[1]) << 16), lo[2] | (uint32_t(lo[3]) << 16), lo[4] | (uint32_t(lo[5]) << 16), lo[6] | (uint32_t(lo[7]) << 16));
      } else if constexpr (VEC == 4) {
        *reinterpret_cast<uint2*>(a.out_hi + row * a.ld16 + col) = make_uint2(hi[0] | (uint32_t(hi[1]) << 16), hi[2] | (uint32_t(hi[3]) << 16));
        if (a.out_lo) *reinterpret_cast<uint2*>(a.out_lo + row * a.ld16 + col) = make_uint2(lo[0] | (uint32_t(lo[1]) << 16), lo[2] | (uint32_t(lo[3]) << 16));
      } else if constexpr (VEC == 2) {
        *reinterpret_cast<uint32_t*>(a.out_hi + row * a.ld16 + col) = hi[0] | (uint32_t(hi[1]) << 16);
        if (a.out_lo) *reinterpret_cast<uint32_t*>(a.out_lo + row * a.ld16 + col) = lo[0] | (uint32_t(lo[1]) << 16);
      } else {
        a.out_hi[row * a.ld16 + col] = hi[0];
        if (a.out_lo) a.out_lo[row * a.ld16 + col] = lo[0];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-core variant for 16-bit Q/K/V planes, head dim 256, Sq <= 8, Sk <= 8 (the rollout's 5- and 6-token
// windows and the pruned last layer's single query).  The warp-level kernels above are issue-bound (ncu r1b:
// ~1850 instructions per (clip, head), 54 % issue utilisation, DRAM 27 %); here the two products run on
// mma.sync.m16n8k16 (legacy tensor path - the tiles are 5x5x256, far too small for tcgen05):
//   S = Q K^T : A = Q (rows padded to 16 by aliasing), B = K ([key][hd] rows are the "col" operand as stored)
//   O^T = V^T P^T : A = V^T through ldmatrix.trans, B = P^T - which is exactly the score accumulator fragment
//                   each thread already holds (row q = lane/4, keys 2t, 2t+1), so no shuffles are needed
// Softmax runs on the accumulator fragments in fp32 (row reductions across the 4 lanes of a quad); P is rounded
// to the plane format for the second product.  One warp per (clip, head), operands staged in padded smem by
// cp.async (rows of 512 B + 16 B pad: ldmatrix conflict-free), output staged through smem for 16-byte stores.
inline bool& attention_mma_enabled() { static bool on = true; return on; }  // SDVG_ATTN_MMA=0 disables
constexpr int kMmaHd = 256;
constexpr int kMmaRow = kMmaHd + 8;   // halves per smem row
#ifndef SDVG_ATTN_GROUP
#define SDVG_ATTN_GROUP 4
#endif
constexpr int kAttnMmaGroup = SDVG_ATTN_GROUP;   // K steps / output tiles whose fragments are in flight together

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
template <bool BF16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if constexpr (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Shared memory of one warp: [Sq query rows | Sk key rows | padding up to 8 rows] [Sk value rows] [one zero row].
// Only the rows that exist are staged (the first version reserved 8 + 8 + 9 rows = 13.2 KB per warp, which capped the
// SM at 16 resident warps of a kernel whose time is load latency; 5-token windows need 16 rows = 8.4 KB -> 26 warps).
// ldmatrix rows beyond Sq / Sk alias row 0 of their operand (their scores are masked / their outputs never stored);
// the first 8 rows double as the output staging area once the scores are done.
__host__ __device__ inline int attn_mma_rows(int Sq, int Sk) { return (Sq + Sk > 8 ? Sq + Sk : 8) + Sk + 1; }

template <bool BF16, int MASK>
__global__ void __launch_bounds__(128) attention_mma_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ __align__(16) uint8_t attn_smem[];
  constexpr float kLog2e = 1.4426950408889634f;
  pdl_wait();
  pdl_trigger();
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * 4 + wib;
  if (warp_global >= a.clips * a.heads) return;
  const int b = warp_global / a.heads, h = warp_global - b * a.heads;
  const int Sq = a.Sq, Sk = a.Sk;
  const int qk_rows = Sq + Sk > 8 ? Sq + Sk : 8;
  uint16_t* smw = reinterpret_cast<uint16_t*>(attn_smem) + static_cast<size_t>(wib) * attn_mma_rows(Sq, Sk) * kMmaRow;
  uint16_t* sq = smw;                                   // [Sq] query rows (rows 0..7 of this block: output staging later)
  uint16_t* sk = smw + Sq * kMmaRow;                    // [Sk] key rows
  uint16_t* sv = smw + qk_rows * kMmaRow;               // [Sk] value rows + the zero row (every key >= Sk points there)
  const uint16_t* qb = reinterpret_cast<const uint16_t*>(a.q) + static_cast<size_t>(b) * a.q_clip_stride + h * kMmaHd;
  const uint16_t* kb = reinterpret_cast<const uint16_t*>(a.k) + static_cast<size_t>(b) * a.kv_clip_stride + h * kMmaHd;
  const uint16_t* vb = reinterpret_cast<const uint16_t*>(a.v) + static_cast<size_t>(b) * a.kv_clip_stride + h * kMmaHd;

  // ---- stage Q, K, V rows (512 B each: 32 lanes x 16 B) and the zero row
  for (int r = a.q_first; r < Sq; ++r) cp_async16(ptx::smem_u32(sq + r * kMmaRow + lane * 8), qb + static_cast<size_t>(r) * a.ldq + lane * 8);
  for (int r = 0; r < Sk; ++r) {
    cp_async16(ptx::smem_u32(sk + r * kMmaRow + lane * 8), kb + static_cast<size_t>(r) * a.ldkv + lane * 8);
    cp_async16(ptx::smem_u32(sv + r * kMmaRow + lane * 8), vb + static_cast<size_t>(r) * a.ldkv + lane * 8);
  }
  *reinterpret_cast<uint4*>(sv + Sk * kMmaRow + lane * 8) = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  // ---- S = Q K^T  (16 x 8 tile; query rows outside [q_first, Sq) and key rows >= Sk alias a staged row: masked below)
  const int mid = lane >> 3, mr = lane & 7;           // ldmatrix: matrix id, row inside the matrix
  const int qrow = (mr >= a.q_first && mr < Sq) ? mr : a.q_first, krow = mr < Sk ? mr : 0;
  const uint32_t q_addr = ptx::smem_u32(sq + qrow * kMmaRow + (mid >> 1) * 8);   // A: M0 rows0-7 k0-7 | M1 rows8-15(alias) k0-7 | M2 k8-15 | M3
  const uint32_t k_addr = ptx::smem_u32(sk + krow * kMmaRow + (mid & 1) * 8);    // B (x2, lanes 0-15): n rows, k 0-7 | k 8-15
  // (fragments of four K steps are loaded ahead of the four HMMAs that use them, and the steps alternate between two
  // accumulators: ncu showed every HMMA of the plain loop stalled on its own ldmatrix (short scoreboard 30 %) and on
  // the previous HMMA's accumulator)
  float sc[4] = {0.f, 0.f, 0.f, 0.f}, sc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  constexpr int G = kAttnMmaGroup;
  for (int kk = 0; kk < kMmaHd / 16; kk += G) {
    uint32_t af[G][4], bf[G][2];
#pragma unroll
    for (int u = 0; u < G; ++u) {
      ldmatrix_x4(af[u], q_addr + (kk + u) * 32);
      ldmatrix_x2(bf[u], k_addr + (kk + u) * 32);
    }
#pragma unroll
    for (int u = 0; u < G; ++u) {
      if (u & 1) mma_16816<BF16>(sc2, af[u], bf[u]);
      else mma_16816<BF16>(sc, af[u], bf[u]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) sc[e] += sc2[e];
  // ---- softmax over the keys of row g = lane / 4 (this thread holds keys 2t, 2t+1)
  const int g = lane >> 2, t = lane & 3;
  const float scale2 = a.scale * kLog2e;
  float p0 = sc[0] * scale2, p1 = sc[1] * scale2;
  const int j0 = 2 * t, j1 = 2 * t + 1;
  const bool row_ok = g >= a.q_first && g < Sq;
  bool ok0 = row_ok && j0 < Sk, ok1 = row_ok && j1 < Sk;
  if (MASK == 1) { ok0 = ok0 && j0 <= g + (Sk - Sq); ok1 = ok1 && j1 <= g + (Sk - Sq); }
  p0 = ok0 ? p0 : -INFINITY;
  p1 = ok1 ? p1 : -INFINITY;
  float mx = fmaxf(p0, p1);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  if (mx == -INFINITY) mx = 0.f;                       // rows outside [q_first, Sq): all weights zero, no NaN
  p0 = exp2f(p0 - mx);
  p1 = exp2f(p1 - mx);
  float sum = p0 + p1;
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  p0 *= inv; p1 *= inv;
  uint32_t pb[2];                                      // B fragment of P^T: keys 2t, 2t+1 of query g | keys 8.. = 0
  if constexpr (BF16) { const __nv_bfloat162 pp = __floats2bfloat162_rn(p0, p1); pb[0] = *reinterpret_cast<const uint32_t*>(&pp); }
  else { const __half2 pp = __floats2half2_rn(p0, p1); pb[0] = *reinterpret_cast<const uint32_t*>(&pp); }
  pb[1] = 0u;

  // ---- O^T = V^T P^T : 16 tiles of 16 head elements x 8 queries
  // A through ldmatrix.trans of V ([key][hd]): M0 keys0-7 hd m0..+7 | M1 keys0-7 hd m0+8.. | M2 keys8-15 | M3 keys8-15
  const int vrow = (mid >> 1) ? Sk : (mr < Sk ? mr : Sk);             // keys >= Sk (and all of 8..15) -> zero row
  const uint32_t v_addr = ptx::smem_u32(sv + vrow * kMmaRow + (mid & 1) * 8);
  __syncwarp();                                                       // everyone is done reading the query / key rows
  uint16_t* so = smw;                                                 // 8 staging rows over the query + key rows
#pragma unroll
  for (int m0 = 0; m0 < kMmaHd / 16; m0 += G) {
    // G independent tiles at a time: their ldmatrix and HMMA latencies overlap, the conversions follow
    uint32_t af[G][4];
    float o[G][4];
#pragma unroll
    for (int u = 0; u < G; ++u) ldmatrix_x4_trans(af[u], v_addr + (m0 + u) * 32);
#pragma unroll
    for (int u = 0; u < G; ++u) {
      o[u][0] = 0.f; o[u][1] = 0.f; o[u][2] = 0.f; o[u][3] = 0.f;
      mma_16816<BF16>(o[u], af[u], pb);
    }
#pragma unroll
    for (int u = 0; u < G; ++u) {
      // o[0], o[1] = O[q = 2t, 2t+1][hd = 16 m + g];  o[2], o[3] = same queries, hd = 16 m + g + 8
      const int hd0 = (m0 + u) * 16 + g;
      if constexpr (BF16) {
        so[(2 * t) * kMmaRow + hd0] = __bfloat16_as_ushort(__float2bfloat16_rn(o[u][0]));
        so[(2 * t + 1) * kMmaRow + hd0] = __bfloat16_as_ushort(__float2bfloat16_rn(o[u][1]));
        so[(2 * t) * kMmaRow + hd0 + 8] = __bfloat16_as_ushort(__float2bfloat16_rn(o[u][2]));
        so[(2 * t + 1) * kMmaRow + hd0 + 8] = __bfloat16_as_ushort(__float2bfloat16_rn(o[u][3]));
      } else {
        so[(2 * t) * kMmaRow + hd0] = __half_as_ushort(__float2half_rn(o[u][0]));
        so[(2 * t + 1) * kMmaRow + hd0] = __half_as_ushort(__float2half_rn(o[u][1]));
        so[(2 * t) * kMmaRow + hd0 + 8] = __half_as_ushort(__float2half_rn(o[u][2]));
        so[(2 * t + 1) * kMmaRow + hd0 + 8] = __half_as_ushort(__float2half_rn(o[u][3]));
      }
    }
  }
  __syncwarp();
  // ---- store the valid query rows, 16 bytes per lane
  for (int i = a.q_first; i < Sq; ++i) {
    const size_t row = a.out_compact ? static_cast<size_t>(b) * (Sq - a.q_first) + (i - a.q_first) : static_cast<size_t>(b) * Sq + i;
    const uint4 v4 = *reinterpret_cast<const uint4*>(&so[i * kMmaRow + lane * 8]);
    *reinterpret_cast<uint4*>(a.out_hi + row * a.ld16 + h * kMmaHd + lane * 8) = v4;
  }
}

// Same scheme for 9..16 queries / keys (the window-10 configurations): two key tiles for Q K^T, two query tiles
// for O^T = V^T P^T (the V^T fragments are shared by both), 2 warps per CTA (26 KB of operands per warp).
struct AttnMmaSmem16 {                     // per warp
  uint16_t q[16][kMmaRow];                 // query rows (reused for the output rows)
  uint16_t k[16][kMmaRow];
  uint16_t v[17][kMmaRow];                 // row 16 = zeros: every key row >= Sk points here
};

template <bool BF16, int MASK>
__global__ void __launch_bounds__(64) attention_mma16_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ __align__(16) uint8_t attn_smem[];
  constexpr float kLog2e = 1.4426950408889634f;
  pdl_wait();
  pdl_trigger();
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * 2 + wib;
  if (warp_global >= a.clips * a.heads) return;
  const int b = warp_global / a.heads, h = warp_global - b * a.heads;
  AttnMmaSmem16& sm = reinterpret_cast<AttnMmaSmem16*>(attn_smem)[wib];
  const int Sq = a.Sq, Sk = a.Sk;
  const uint16_t* qb = reinterpret_cast<const uint16_t*>(a.q) + static_cast<size_t>(b) * a.q_clip_stride + h * kMmaHd;
  const uint16_t* kb = reinterpret_cast<const uint16_t*>(a.k) + static_cast<size_t>(b) * a.kv_clip_stride + h * kMmaHd;
  const uint16_t* vb = reinterpret_cast<const uint16_t*>(a.v) + static_cast<size_t>(b) * a.kv_clip_stride + h * kMmaHd;
  for (int r = a.q_first; r < Sq; ++r) cp_async16(ptx::smem_u32(&sm.q[r][lane * 8]), qb + static_cast<size_t>(r) * a.ldq + lane * 8);
  for (int r = 0; r < Sk; ++r) {
    cp_async16(ptx::smem_u32(&sm.k[r][lane * 8]), kb + static_cast<size_t>(r) * a.ldkv + lane * 8);
    cp_async16(ptx::smem_u32(&sm.v[r][lane * 8]), vb + static_cast<size_t>(r) * a.ldkv + lane * 8);
  }
  *reinterpret_cast<uint4*>(&sm.v[16][lane * 8]) = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  // ---- S = Q K^T : 16 queries x 16 keys = two n-tiles
  const int mid = lane >> 3, mr = lane & 7;
  const uint32_t q_addr = ptx::smem_u32(&sm.q[(mid & 1) * 8 + mr][(mid >> 1) * 8]);   // M0 rows0-7 k0-7 | M1 rows8-15 k0-7 | M2, M3: k8-15
  const uint32_t k_addr = ptx::smem_u32(&sm.k[(mid >> 1) * 8 + mr][(mid & 1) * 8]);   // x4: M0 keys0-7 k0-7 | M1 keys0-7 k8-15 | M2 keys8-15 k0-7 | M3 keys8-15 k8-15
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kk = 0; kk < kMmaHd / 16; ++kk) {
    uint32_t af[4], bf[4];
    ldmatrix_x4(af, q_addr + kk * 32);
    ldmatrix_x4(bf, k_addr + kk * 32);
    const uint32_t b0[2] = {bf[0], bf[1]}, b1[2] = {bf[2], bf[3]};
    mma_16816<BF16>(s0, af, b0);
    mma_16816<BF16>(s1, af, b1);
  }
  // ---- softmax: rows g (values s0[0..1], s1[0..1]) and g + 8 (s0[2..3], s1[2..3]); keys 2t, 2t+1, 8+2t, 9+2t
  const int g = lane >> 2, t = lane & 3;
  const float scale2 = a.scale * kLog2e;
  uint32_t pb_lo[2], pb_hi[2];   // B fragments of P^T for queries 0-7 and 8-15
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = g + half * 8;
    float p[4] = {s0[half * 2] * scale2, s0[half * 2 + 1] * scale2, s1[half * 2] * scale2, s1[half * 2 + 1] * scale2};
    const int key[4] = {2 * t, 2 * t + 1, 8 + 2 * t, 9 + 2 * t};
    const bool row_ok = row >= a.q_first && row < Sq;
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bool ok = row_ok && key[i] < Sk;
      if (MASK == 1) ok = ok && key[i] <= row + (Sk - Sq);
      p[i] = ok ? p[i] : -INFINITY;
      mx = fmaxf(mx, p[i]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    if (mx == -INFINITY) mx = 0.f;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { p[i] = exp2f(p[i] - mx); sum += p[i]; }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
    uint32_t w0, w1;
    if constexpr (BF16) {
      const __nv_bfloat162 a0 = __floats2bfloat162_rn(p[0] * inv, p[1] * inv), a1 = __floats2bfloat162_rn(p[2] * inv, p[3] * inv);
      w0 = *reinterpret_cast<const uint32_t*>(&a0); w1 = *reinterpret_cast<const uint32_t*>(&a1);
    } else {
      const __half2 a0 = __floats2half2_rn(p[0] * inv, p[1] * inv), a1 = __floats2half2_rn(p[2] * inv, p[3] * inv);
      w0 = *reinterpret_cast<const uint32_t*>(&a0); w1 = *reinterpret_cast<const uint32_t*>(&a1);
    }
    if (half == 0) { pb_lo[0] = w0; pb_lo[1] = w1; } else { pb_hi[0] = w0; pb_hi[1] = w1; }
  }
  // ---- O^T = V^T P^T
  const int vkey = (mid >> 1) * 8 + mr;                                    // M0/M1: keys 0-7, M2/M3: keys 8-15
  const uint32_t v_addr = ptx::smem_u32(&sm.v[vkey < Sk ? vkey : 16][(mid & 1) * 8]);
  __syncwarp();
  uint16_t* so = &sm.q[0][0];
#pragma unroll
  for (int m = 0; m < kMmaHd / 16; ++m) {
    uint32_t af[4];
    ldmatrix_x4_trans(af, v_addr + m * 32);
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
    mma_16816<BF16>(o0, af, pb_lo);   // queries 0-7
    mma_16816<BF16>(o1, af, pb_hi);   // queries 8-15
    const int hd0 = m * 16 + g;
#pragma unroll
    for (int qt = 0; qt < 2; ++qt) {
      const float* o = qt ? o1 : o0;
      const int q0 = qt * 8 + 2 * t;
      if constexpr (BF16) {
        so[q0 * kMmaRow + hd0] = __bfloat16_as_ushort(__float2bfloat16_rn(o[0]));
        so[(q0 + 1) * kMmaRow + hd0] = __bfloat16_as_ushort(__float2bfloat16_rn(o[1]));
        so[q0 * kMmaRow + hd0 + 8] = __bfloat16_as_ushort(__float2bfloat16_rn(o[2]));
        so[(q0 + 1) * kMmaRow + hd0 + 8] = __bfloat16_as_ushort(__float2bfloat16_rn(o[3]));
      } else {
        so[q0 * kMmaRow + hd0] = __half_as_ushort(__float2half_rn(o[0]));
        so[(q0 + 1) * kMmaRow + hd0] = __half_as_ushort(__float2half_rn(o[1]));
        so[q0 * kMmaRow + hd0 + 8] = __half_as_ushort(__float2half_rn(o[2]));
        so[(q0 + 1) * kMmaRow + hd0 + 8] = __half_as_ushort(__float2half_rn(o[3]));
      }
    }
  }
  __syncwarp();
  for (int i = a.q_first; i < Sq; ++i) {
    const size_t row = a.out_compact ? static_cast<size_t>(b) * (Sq - a.q_first) + (i - a.q_first) : static_cast<size_t>(b) * Sq + i;
    const uint4 v4 = *reinterpret_cast<const uint4*>(&so[i * kMmaRow + lane * 8]);
    *reinterpret_cast<uint4*>(a.out_hi + row * a.ld16 + h * kMmaHd + lane * 8) = v4;
  }
}

// usable when Q/K/V are 16-bit planes, hd == 256, at most 16 queries and keys, plane output without a lo plane
inline bool attention_mma_supported(const AttnArgs& a) {
  return a.hd == kMmaHd && a.Sq <= 16 && a.Sk <= 16 && (a.mask_kind == 0 || a.mask_kind == 1) && a.out_hi && !a.out_lo &&
         !a.out32 && a.ldq % 8 == 0 && a.ldkv % 8 == 0 && a.q_clip_stride % 8 == 0 && a.kv_clip_stride % 8 == 0 && a.ld16 % 8 == 0;
}

inline cudaError_t launch_attention_mma16(const AttnArgs& a, cudaStream_t stream) {
  const int grid = ceil_div(a.clips * a.heads, 2);
  const size_t smem = 2 * sizeof(AttnMmaSmem16);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attention_mma16_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_mma16_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_mma16_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_mma16_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  if (a.bf16) {
    if (a.mask_kind == 1) return launch_kernel(attention_mma16_kernel<true, 1>, dim3(grid), dim3(64), smem, stream, a);
    return launch_kernel(attention_mma16_kernel<true, 0>, dim3(grid), dim3(64), smem, stream, a);
  }
  if (a.mask_kind == 1) return launch_kernel(attention_mma16_kernel<false, 1>, dim3(grid), dim3(64), smem, stream, a);
  return launch_kernel(attention_mma16_kernel<false, 0>, dim3(grid), dim3(64), smem, stream, a);
}

inline cudaError_t launch_attention_mma(const AttnArgs& a, cudaStream_t stream) {
  if (a.Sq > 8 || a.Sk > 8) return launch_attention_mma16(a, stream);
  const int grid = ceil_div(a.clips * a.heads, 4);
  const size_t smem = 4 * static_cast<size_t>(attn_mma_rows(a.Sq, a.Sk)) * kMmaRow * sizeof(uint16_t);
  const size_t smem_max = 4 * static_cast<size_t>(attn_mma_rows(8, 8)) * kMmaRow * sizeof(uint16_t);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    const size_t smem = smem_max;
    cudaError_t e = cudaFuncSetAttribute(attention_mma_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_mma_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_mma_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_mma_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  if (a.bf16) {
    if (a.mask_kind == 1) return launch_kernel(attention_mma_kernel<true, 1>, dim3(grid), dim3(128), smem, stream, a);
    return launch_kernel(attention_mma_kernel<true, 0>, dim3(grid), dim3(128), smem, stream, a);
  }
  if (a.mask_kind == 1) return launch_kernel(attention_mma_kernel<false, 1>, dim3(grid), dim3(128), smem, stream, a);
  return launch_kernel(attention_mma_kernel<false, 0>, dim3(grid), dim3(128), smem, stream, a);
}

// shapes the 16-bit-input kernel is instantiated for (the engine asks before choosing the 16-bit Q/K/V layout)
inline bool attention16_supported(int hd, int Sq, int Sk, int mask_kind) {
  const bool shape = (Sq == 5 && Sk == 5) || (Sq == 6 && Sk == 6) || (Sq == 10 && Sk == 10) || (Sq == 5 && Sk == 6) ||
                     (Sq == 1 && (Sk == 5 || Sk == 6 || Sk == 10));
  return shape && (hd == 256 || hd == 128 || hd == 64 || hd == 32) && (mask_kind == 0 || mask_kind == 1);
}

template <int VEC, int SQ, int SK>
inline cudaError_t launch_attention16_m(const AttnArgs& a, int grid, cudaStream_t stream) {
  if (a.mask_kind == 0) return launch_kernel(attention_exact16_kernel<VEC, SQ, SK, 0>, dim3(grid), dim3(128), 0, stream, a);
  return launch_kernel(attention_exact16_kernel<VEC, SQ, SK, 1>, dim3(grid), dim3(128), 0, stream, a);
}
template <int VEC>
inline cudaError_t launch_attention16_v(const AttnArgs& a, int grid, cudaStream_t stream) {
  if (a.Sq == 5 && a.Sk == 5) return launch_attention16_m<VEC, 5, 5>(a, grid, stream);
  if (a.Sq == 6 && a.Sk == 6) return launch_attention16_m<VEC, 6, 6>(a, grid, stream);
  if (a.Sq == 10 && a.Sk == 10) return launch_attention16_m<VEC, 10, 10>(a, grid, stream);
  if (a.Sq == 1 && a.Sk == 5) return launch_attention16_m<VEC, 1, 5>(a, grid, stream);
  if (a.Sq == 1 && a.Sk == 6) return launch_attention16_m<VEC, 1, 6>(a, grid, stream);
  if (a.Sq == 1 && a.Sk == 10) return launch_attention16_m<VEC, 1, 10>(a, grid, stream);
  return launch_attention16_m<VEC, 5, 6>(a, grid, stream);
}
// Q/K/V given as 16-bit planes (a.q/a.k/a.v point at uint16 data; ldq/ldkv/clip strides in elements).
inline cudaError_t launch_attention16(const AttnArgs& a_in, cudaStream_t stream) {
  AttnArgs a = a_in;
  if (a.q_clip_stride == 0) a.q_clip_stride = static_cast<long long>(a.Sq) * a.ldq;
  if (a.kv_clip_stride == 0) a.kv_clip_stride = static_cast<long long>(a.Sk) * a.ldkv;
  if (!attention16_supported(a.hd, a.Sq, a.Sk, a.mask_kind) || a.ldq % 8 || a.ldkv % 8 || a.q_clip_stride % 8 ||
      a.kv_clip_stride % 8)
    return cudaErrorInvalidValue;
  if (attention_mma_enabled() && attention_mma_supported(a)) return launch_attention_mma(a, stream);
  const int grid = ceil_div(a.clips * a.heads, 4);
  switch (a.hd) {
    case 256: return launch_attention16_v<8>(a, grid, stream);
    case 128: return launch_attention16_v<4>(a, grid, stream);
    case 64: return launch_attention16_v<2>(a, grid, stream);
    default: return launch_attention16_v<1>(a, grid, stream);
  }
}

inline cudaError_t launch_attention(const AttnArgs& a_in, cudaStream_t stream) {
  AttnArgs a = a_in;
  if (a.q_clip_stride == 0) a.q_clip_stride = static_cast<long long>(a.Sq) * a.ldq;
  if (a.kv_clip_stride == 0) a.kv_clip_stride = static_cast<long long>(a.Sk) * a.ldkv;
  if (a.Sk > kAttnMaxS || a.Sq > kAttnMaxS || a.hd > 256) return cudaErrorInvalidValue;
  const int warps = a.clips * a.heads;
  const int grid = ceil_div(warps, 4);
  const bool al4 = (a.ldq % 4 == 0) && (a.ldkv % 4 == 0) && (a.ld32 % 4 == 0) && (a.q_clip_stride % 4 == 0) &&
                   (a.kv_clip_stride % 4 == 0);
  cudaError_t fast = cudaErrorNotSupported;
  if (al4 && a.hd == 256) fast = launch_attention_exact<4, 2>(a, grid, stream);
  else if (al4 && a.hd == 128) fast = launch_attention_exact<4, 1>(a, grid, stream);
  else if (al4 && a.hd == 64) fast = launch_attention_exact<2, 1>(a, grid, stream);
  else if (al4 && a.hd == 32) fast = launch_attention_exact<1, 1>(a, grid, stream);
  if (fast != cudaErrorNotSupported) return fast;
  if (al4 && a.hd == 256) fast = launch_attention_reg<4, 2>(a, grid, stream);
  else if (al4 && a.hd == 128) fast = launch_attention_reg<4, 1>(a, grid, stream);
  else if (al4 && a.hd == 64) fast = launch_attention_reg<2, 1>(a, grid, stream);
  else if (al4 && a.hd == 32) fast = launch_attention_reg<1, 1>(a, grid, stream);
  if (fast != cudaErrorNotSupported) return fast;
  const bool vec = (a.hd % 128 == 0) && al4;
  if (vec && a.hd == 128) return launch_kernel(attention_kernel<4, 1>, dim3(grid), dim3(128), 0, stream, a);
  else if (vec && a.hd == 256) return launch_kernel(attention_kernel<4, 2>, dim3(grid), dim3(128), 0, stream, a);
  else if (a.hd <= 32) return launch_kernel(attention_kernel<1, 1>, dim3(grid), dim3(128), 0, stream, a);
  else if (a.hd <= 64) return launch_kernel(attention_kernel<1, 2>, dim3(grid), dim3(128), 0, stream, a);
  else if (a.hd <= 128) return launch_kernel(attention_kernel<1, 4>, dim3(grid), dim3(128), 0, stream, a);
  else return launch_kernel(attention_kernel<1, 8>, dim3(grid), dim3(128), 0, stream, a);
}

}  // namespace sdvg
