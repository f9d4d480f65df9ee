// Latent window plumbing around the transformer (prediction/predict.py:124-133,188-196; utils/sd_utils.py:31,143,159):
// gather the current window of latent frames out of the device-resident history (ring of slots per clip),
// insert the SOS frame (constant 2.0), apply the 0.18215 latent scale, and emit either fp32 rows or the
// 16-bit operand planes the embedding GEMM reads through TMA.  The same kernel exports predictions
// (history slots -> (B, n_pred, E), with the 1/0.18215 egress scale) and ingests the context.
// Pure streaming: 4 E bytes in, 2..8 E bytes out per token, 128-bit accesses.
#pragma once
#include "common.cuh"

namespace sdvg {

constexpr int kPackMaxTokens = 32;

struct PackArgs {
  const float* src;          // [clips][src_clip_stride] with token t at offset slot[t] * src_slot_stride
  long long src_clip_stride; long long src_slot_stride;
  int clips, tokens, width;  // width = E (multiple of 4)
  int slot[kPackMaxTokens];  // source slot per output token; -1 -> constant row `fill`
  float fill, scale;
  float* out32; long long out_clip_stride; long long out_tok_stride;   // fp32 destination (elements)
  uint16_t* out_hi; uint16_t* out_lo; int ld16; int bf16;             // planes: row = clip * tokens + t
  // weight packing reuses this kernel with tokens = 1 and slot[0] = 0
};

__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackArgs a) {
  pdl_wait();
  pdl_trigger();
  const int w4 = a.width >> 2;
  const long long total = static_cast<long long>(a.clips) * a.tokens * w4;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * 256) {
    const int c4 = static_cast<int>(idx % w4);
    const long long rt = idx / w4;
    const int t = static_cast<int>(rt % a.tokens);
    const long long b = rt / a.tokens;
    float4 v;
    const int s = a.slot[t];
    if (s < 0) v = make_float4(a.fill, a.fill, a.fill, a.fill);
    else {
      v = __ldg(reinterpret_cast<const float4*>(a.src + b * a.src_clip_stride + s * a.src_slot_stride) + c4);
      v.x *= a.scale; v.y *= a.scale; v.z *= a.scale; v.w *= a.scale;
    }
    if (a.out32) *(reinterpret_cast<float4*>(a.out32 + b * a.out_clip_stride + t * a.out_tok_stride) + c4) = v;
    if (a.out_hi) {
      const float f[4] = {v.x, v.y, v.z, v.w};
      uint16_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { hi[i] = to_plane_hi(f[i], a.bf16); lo[i] = to_plane_lo(f[i], hi[i]); }
      const size_t o = static_cast<size_t>(rt) * a.ld16 + c4 * 4;
      *reinterpret_cast<uint2*>(a.out_hi + o) = make_uint2(hi[0] | (uint32_t(hi[1]) << 16), hi[2] | (uint32_t(hi[3]) << 16));
      if (a.out_lo)
        *reinterpret_cast<uint2*>(a.out_lo + o) = make_uint2(lo[0] | (uint32_t(lo[1]) << 16), lo[2] | (uint32_t(lo[3]) << 16));
    }
  }
}

// dst[b][:] += src[b][:]  (or += fill when src == nullptr): the residual prediction of the reference's
// prediction/predict_diff.py:33  (pred[:, -1] = pred[:, -1] + y_input[:, -2]).
struct AddArgs {
  float* dst; long long dst_clip_stride;
  const float* src; long long src_clip_stride;
  float fill;
  int clips, width;
};

__global__ void __launch_bounds__(256) add_rows_kernel(const __grid_constant__ AddArgs a) {
  pdl_wait();
  pdl_trigger();
  const int w4 = a.width >> 2;
  const long long total = static_cast<long long>(a.clips) * w4;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * 256) {
    const int c4 = static_cast<int>(idx % w4);
    const long long b = idx / w4;
    float4* d = reinterpret_cast<float4*>(a.dst + b * a.dst_clip_stride) + c4;
    float4 v = *d;
    const float4 s = a.src ? __ldg(reinterpret_cast<const float4*>(a.src + b * a.src_clip_stride) + c4)
                           : make_float4(a.fill, a.fill, a.fill, a.fill);
    v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    *d = v;
  }
}

inline cudaError_t launch_add_rows(const AddArgs& a, int num_sms, cudaStream_t stream) {
  const long long total = static_cast<long long>(a.clips) * (a.width >> 2);
  if (total == 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  return launch_kernel(add_rows_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, a);
}

inline cudaError_t launch_pack(const PackArgs& a, int num_sms, cudaStream_t stream) {
  if (a.width % 4 != 0 || a.tokens > kPackMaxTokens) return cudaErrorInvalidValue;
  const long long total = static_cast<long long>(a.clips) * a.tokens * (a.width >> 2);
  if (total == 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  return launch_kernel(pack_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, a);
}

}  // namespace sdvg
