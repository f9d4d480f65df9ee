// K1 - tensor-core GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T, then the fused Epilogue (common.cuh).
//
// Replaces every nn.Linear / packed in-proj of the reference path (models/transformer.py:37,45 and the
// torch.nn.Transformer layers built at :38-44): embed, QKV, cross-Q/KV, out-proj, FF1, FF2, out.
//
// Design (one CTA = one SM, persistent over output tiles, 192 threads):
//   warp 0 : TMA producer   - cp.async.bulk.tensor 2-D loads of the A (activation) and W (weight) K-slabs
//                             (64 x 16-bit = 128-byte rows, SWIZZLE_128B) into a multi-stage smem ring
//   warp 1 : MMA issuer     - one lane issues tcgen05.mma kind::f16 (M=128, N=BN, K=16) with fp32
//                             accumulators in TMEM; tcgen05.commit releases smem stages / publishes tiles
//   warps 2-5 : epilogue    - tcgen05.ld (32 lanes x 32 columns per warp), transpose through padded smem so
//                             global traffic is row-contiguous, bias / scale / PE / ReLU / residual, stores of
//                             the fp32 result and/or the 16-bit operand planes of the next GEMM
//   TMEM accumulators are double-buffered: the epilogue of tile i overlaps the main loop of tile i+1.
//
// Both operands are K-major (PyTorch weights are [N][K], so x W^T needs no transpose).
//
// SPLIT = true is the fp32-parity mode: A and W each come as two fp16 planes (x = hi + 2^-11 lo) and three
// products are accumulated - hi*hi into accumulator 0, hi*lo + lo*hi into accumulator 1 (kept scaled by 2^11,
// so nothing underflows) - and combined in the epilogue.  22 significant bits per operand at 3 MMAs.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace sdvg {

constexpr int kTcBM = 128;
constexpr int kTcBK = 64;
constexpr int kTcThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)
constexpr int kTcEpiWarps = 8;
constexpr int kTcEpiStride = 36;  // floats; 32x32 transpose tile, padded to keep 128-bit accesses conflict-free
constexpr int kTcSmemLimit = 232448;  // 227 KB

// AROWS: rows of shared memory a stage keeps for the A operand.  128 = the whole UMMA tile.  48 (small batches, every
// token row fits the 48-row TMA box): the MMA still reads 128 rows from the stage's base - rows 48..127 are whatever
// follows in shared memory (the B tile, the next stages) and only feed accumulator rows >= M, which no epilogue stores -
// and the ring gets twice the stages.  A narrow-tile launch is bound by the bytes one SM keeps in flight (8 stages of
// 6 + 4 KB against ~1 us of HBM latency): C1 launch chain 808.6 us per pass with 8 stages, 799.5 with 9 (same box).
template <int BN, bool SPLIT, int AROWS = 128>
struct TcCfg {
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN");
  static_assert(!(SPLIT && BN == 256), "split mode needs two accumulators per tile: BN <= 128");
  static_assert(AROWS == 128 || (AROWS % 8 == 0 && !SPLIT && BN <= 64), "compact A stages: one-plane narrow tiles only");
  static constexpr int kPlanes = SPLIT ? 2 : 1;
  static constexpr int kABytes = AROWS * kTcBK * 2;
  static constexpr int kBBytes = BN * kTcBK * 2;
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kEpiBytes = kTcEpiWarps * 32 * kTcEpiStride * 4 + kTcEpiWarps * 32 * 8;   // transpose tiles + (mean, rstd) of each warp's 32 rows
  // epilogue warps that take part: two per quadrant split the tile's 32-column chunks (one per quadrant if BN == 32)
  static constexpr int kEpiActive = BN >= 64 ? 8 : 4;
  static constexpr int kBarBytes = 1024;
  static constexpr int kMaxStages = (kTcSmemLimit - 1024 /*align slack*/ - kEpiBytes - kBarBytes) / kStageBytes;
  static constexpr int kStageCap = AROWS < 128 ? 16 : 8;   // small-M launches are bound by the bytes in flight per SM
  static constexpr int kStages = kMaxStages > kStageCap ? kStageCap : kMaxStages;
  static constexpr int kColsPerTile = BN * kPlanes;
  static constexpr int kTmemCols = (2 * kColsPerTile) < 32 ? 32 : (2 * kColsPerTile);
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiBytes + kBarBytes;
  static_assert(kStages >= 2, "pipeline depth");
  static_assert(kTmemCols <= 512, "TMEM");
  // the last stage's 128-row A read stays inside the allocation
  static_assert(kStageBytes + kEpiBytes + kBarBytes >= 128 * kTcBK * 2, "A overrun");
};

__device__ __forceinline__ unsigned long long global_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SDVG_TRACE(slot) do { if (args.trace && blockIdx.x == 0) args.trace[slot] = global_timer(); } while (0)

struct TcGemmArgs {
  int M, N, K;
  int bf16;  // operand format of the hi planes
  int vec4;  // all epilogue pointers / pitches are 16-byte aligned and N % 4 == 0 (set by the launcher)
  int a_box_rows;             // rows of A the TMA loads per K block (128, or 64 / 32 for small-M problems; one-CTA kernel)
  int max_stages;             // 0 = all smem stages; >0 caps the TMA ring depth (pipeline experiments, SDVG_STAGES)
  unsigned long long* trace;  // optional [64] device buffer: CTA 0 records %globaltimer at pipeline events (tools/gemm_trace.py)
  Epilogue epi;
  // split-K for small-M problems (one-CTA kernel, one tile per CTA group): `ksplit` consecutive CTAs share a tile,
  // CTA r accumulates K blocks [r nk / ks, (r+1) nk / ks); r > 0 leave their fp32 partial tiles in ks_ws and bump
  // ks_flags[tile], r == 0 waits, adds them in a fixed order (deterministic) and runs the fused epilogue.
  int ksplit;
  float* ks_ws;             // [tiles][ksplit-1][128][BN]
  unsigned int* ks_flags;   // [tiles], zero between launches (reset by the last reader)
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 pack_hi4(const float4& v, int bf16) {
  if (bf16) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
// lo = fp16((v - fp16(v)) * 2^11), four at a time (fp16 hi planes only)
__device__ __forceinline__ uint2 pack_lo4(const float4& v, const uint2& hi) {
  const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&hi.x));
  const float2 h1 = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
  const __half2 a = __floats2half2_rn((v.x - h0.x) * kSplitScale, (v.y - h0.y) * kSplitScale);
  const __half2 b = __floats2half2_rn((v.z - h1.x) * kSplitScale, (v.w - h1.y) * kSplitScale);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// Epilogue of one accumulator tile for one warp: 32 TMEM lanes (rows row0..row0+31) x BN columns starting at
// global column col_base.  TMEM -> registers -> padded smem transpose -> row-contiguous 128-bit global accesses.
// `release()` is called by lane 0 once every TMEM read of the tile has completed (hands the accumulator back
// to the MMA issuer before the global stores are issued).
// FANCY = false is the lean path of the 70-odd per-layer GEMMs (bias, ReLU, residual, fp32 / plane outputs);
// FANCY = true adds what only the embedding and output projections need (scale, positional rows, row remapping).
// The first version of this routine was one generic loop; ncu showed the four epilogue warps issue-bound on
// per-element address arithmetic, integer division and generic-space smem accesses (~18 us per 128x256 tile,
// longer than the tile's MMA main loop), so everything row- or tile-invariant is hoisted and the staging
// buffer is addressed in the shared window explicitly.
// Bias / residual operands of up to PF 32-column chunks, loaded ahead of use.  The epilogue warps fill this for
// the NEXT tile before they block on its accumulator barrier, so the ~1 us L2 latency of the residual rows is
// hidden behind the tile's MMA main loop (with a one-chunk look-ahead the last tile's epilogue - which nothing
// overlaps - still took 4 us of a 32 us launch; see tools/gemm_trace.py).
template <int PF>
struct EpiPre {
  float4 b4[PF];
  float4 res[PF][8];
  float4 lw[PF], lb[PF];  // deferred-LayerNorm weights of the chunk's columns (Epilogue::ln_stats mode)
  float2* st;             // shared memory, [32]: (mean, rstd) of the warp's 32 rows (16 registers per thread when it was an
                          // array here: the epilogue is at the register limit and spilled once the folded LayerNorm came in)
};

template <bool FANCY>
__device__ __forceinline__ void epi_prefetch_chunk(const TcGemmArgs& args, int row0, int col, int lr, float4& b4,
                                                   float4 (&res)[8], float4& lw, float4& lb) {
  const Epilogue& e = args.epi;
  const int rows_valid = args.M - row0;
  b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < args.N) {
    if (e.bias) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
    if (e.ln_stats) {
      lw = __ldg(reinterpret_cast<const float4*>(e.ln_w + col));
      lb = __ldg(reinterpret_cast<const float4*>(e.ln_b + col));
    }
    if (e.residual) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + lr;
        size_t rrow = static_cast<size_t>(row0 + rr);
        if constexpr (FANCY) rrow = static_cast<size_t>(epi_res_row(e, row0 + rr));
        res[i] = rr < rows_valid ? __ldg(reinterpret_cast<const float4*>(e.residual + rrow * e.ld_res + col))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// Issue the loads of the first min(NCW, PF) chunks of a tile (chunks c0 .. c0+NCW-1 belong to this warp).
template <bool FANCY, int NCW, int PF>
__device__ __forceinline__ void tc_epilogue_prefetch(const TcGemmArgs& args, int row0, int col_base, int lane, int c0,
                                                     EpiPre<PF>& pre) {
  if (!args.vec4) return;
  const int lr = lane >> 3, lc = (lane & 7) * 4;
  if (args.epi.ln_stats) {
    const int row = row0 + lane;
    int rrow = row;
    if constexpr (FANCY) rrow = epi_res_row(args.epi, row);
    float2 stv = make_float2(0.f, 0.f);
    if (row < args.M) {
      if (args.epi.stat_in) {   // small batch: the row's partial sums straight from the producer's epilogue slots
        const float2* p = args.epi.stat_in + static_cast<size_t>(rrow) * args.epi.stat_in_ld;
        const int n = args.epi.stat_in_n;
        float s = 0.f, q = 0.f;
#pragma unroll 8
        for (int i = 0; i < n; ++i) { const float2 v = __ldcg(p + i); s += v.x; q += v.y; }
        const float mean = s * args.epi.stat_inv_d;
        stv = make_float2(mean, rsqrtf(fmaxf(q * args.epi.stat_inv_d - mean * mean, 0.f) + args.epi.stat_eps));
      } else {
        stv = __ldg(args.epi.ln_stats + rrow);
      }
    }
    pre.st[lane] = stv;
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < (NCW < PF ? NCW : PF); ++j)
    epi_prefetch_chunk<FANCY>(args, row0, col_base + (c0 + j) * 32 + lc, lr, pre.b4[j], pre.res[j], pre.lw[j], pre.lb[j]);
}

// split-K: add the peers' partial sums of this lane's row (32 columns of chunk c) to the accumulator registers
__device__ __forceinline__ void ks_add_partials(uint32_t (&r0)[32], const float* part_row, int kparts, int part_stride, int c) {
  for (int p = 0; p < kparts; ++p) {
    const float4* s4 = reinterpret_cast<const float4*>(part_row + static_cast<size_t>(p) * part_stride + c * 32);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 v = __ldcg(s4 + u);   // written by another SM during this launch: bypass L1
      r0[4 * u] = __float_as_uint(__uint_as_float(r0[4 * u]) + v.x);
      r0[4 * u + 1] = __float_as_uint(__uint_as_float(r0[4 * u + 1]) + v.y);
      r0[4 * u + 2] = __float_as_uint(__uint_as_float(r0[4 * u + 2]) + v.z);
      r0[4 * u + 3] = __float_as_uint(__uint_as_float(r0[4 * u + 3]) + v.w);
    }
  }
}

// split-K peer (rank > 0): dump this warp's 32 rows x NCW chunks of the accumulator as fp32 into the workspace
template <int BN, bool SPLIT, int NCW, typename Release>
__device__ __forceinline__ void tc_epilogue_partial(uint32_t tbase, int c0, int lane, float* dst_row, Release release) {
#pragma unroll
  for (int j = 0; j < NCW; ++j) {
    const int c = c0 + j;
    uint32_t r0[32];
    ptx::tmem_ld_32x32b_x32(tbase + c * 32, r0);
    if (SPLIT) {
      uint32_t r1[32];
      ptx::tmem_ld_32x32b_x32(tbase + BN + c * 32, r1);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) r0[i] = __float_as_uint(fmaf(__uint_as_float(r1[i]), kSplitInv, __uint_as_float(r0[i])));
    } else {
      ptx::tmem_ld_wait();
    }
    float4* d4 = reinterpret_cast<float4*>(dst_row + c * 32);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      __stcg(d4 + u, make_float4(__uint_as_float(r0[4 * u]), __uint_as_float(r0[4 * u + 1]), __uint_as_float(r0[4 * u + 2]),
                                 __uint_as_float(r0[4 * u + 3])));
  }
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) release();
  __threadfence();
  __syncwarp();
}

// Epilogue of one accumulator tile for one warp: 32 TMEM lanes (rows row0..row0+31) x NCW chunks of 32 columns
// starting at chunk c0 of the tile whose first global column is col_base.  TMEM -> registers -> padded smem
// transpose -> row-contiguous 128-bit global accesses.  `release()` is called by lane 0 once every TMEM read of
// the tile has completed (hands the accumulator back to the MMA issuer before the last global stores).
// FANCY = false is the lean path of the 70-odd per-layer GEMMs (bias, ReLU, residual, fp32 / plane outputs);
// FANCY = true adds what only a few GEMMs need (scale, positional rows, row remapping, clip-strided residual).
// The first version of this routine was one generic loop; ncu showed the epilogue warps issue-bound on
// per-element address arithmetic, integer division and generic-space smem accesses (~18 us per 128x256 tile,
// longer than the tile's MMA main loop), so everything row- or tile-invariant is hoisted and the staging
// buffer is addressed in the shared window explicitly.
template <int BN, bool SPLIT, bool FANCY, int NCW, int PF, typename Release>
__device__ __forceinline__ void tc_epilogue_tile(const TcGemmArgs& args, float* stg, uint32_t tbase, int row0,
                                                 int col_base, int lane, int c0, EpiPre<PF>& pre, Release release,
                                                 const float* part_row = nullptr, int kparts = 0) {
  const Epilogue& e = args.epi;
  const int M = args.M, N = args.N;
  const uint32_t stg_addr = ptx::smem_u32(stg);
  if (!args.vec4) {
    // unaligned / odd-N fallback: one column per lane, one row per pass (never on the model's hot path)
#pragma unroll 1
    for (int c = c0; c < c0 + NCW; ++c) {
      uint32_t r0[32];
      ptx::tmem_ld_32x32b_x32(tbase + c * 32, r0);
      if (SPLIT) {
        uint32_t r1[32];
        ptx::tmem_ld_32x32b_x32(tbase + BN + c * 32, r1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) r0[i] = __float_as_uint(fmaf(__uint_as_float(r1[i]), kSplitInv, __uint_as_float(r0[i])));
      } else {
        ptx::tmem_ld_wait();
      }
      if (kparts) ks_add_partials(r0, part_row, kparts, kTcBM * BN, c);
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128(stg_addr + (lane * kTcEpiStride + 4 * j) * 4, r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
      if (c == c0 + NCW - 1) { ptx::tc_fence_before(); __syncwarp(); if (lane == 0) release(); } else { __syncwarp(); }
      const int col = col_base + c * 32 + lane;
      if (col < N) {
        for (int r = 0; r < 32; ++r) {
          const int row = row0 + r;
          if (row >= M) break;
          const float v = epi_value(e, stg[r * kTcEpiStride + lane], row, col, epi_pe_row(e, row));
          epi_store(e, v, row, epi_out_row(e, row), col);
        }
      }
      __syncwarp();
    }
    return;
  }

  // transposed read-back: 8 lanes x float4 cover the 32 columns of one row, 4 rows per pass, 8 passes
  const int lr = lane >> 3, lc = (lane & 7) * 4;
  const int rows_valid = M - row0;  // rows of this warp's slab inside the matrix (>= 32: all)
  const float* residual = e.residual;
  float* out32 = e.out32;
  uint16_t* out_hi = e.out_hi;
  uint16_t* out_lo = e.out_lo;
  const int relu = e.relu, bf16 = e.bf16;
  const size_t ld32 = e.ld32, ld16 = e.ld16;
  const uint32_t rd_addr = stg_addr + (lr * kTcEpiStride + lc) * 4;
  const uint32_t wr_addr = stg_addr + lane * kTcEpiStride * 4;
  const bool ln_in = e.ln_in != 0;
  float2* const stat_out = e.stat_out;
  float ssum[8], ssq[8];   // producer of a folded LayerNorm: this lane's share of (sum, sum of squares) of its 8 rows
#pragma unroll
  for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }

#pragma unroll
  for (int j = 0; j < NCW; ++j) {
    const int c = c0 + j;
    constexpr int kNoSlot = 0;
    (void)kNoSlot;
    uint32_t r0[32];
    ptx::tmem_ld_32x32b_x32(tbase + c * 32, r0);
    if (SPLIT) {
      uint32_t r1[32];
      ptx::tmem_ld_32x32b_x32(tbase + BN + c * 32, r1);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) r0[i] = __float_as_uint(fmaf(__uint_as_float(r1[i]), kSplitInv, __uint_as_float(r0[i])));
    } else {
      ptx::tmem_ld_wait();
    }
    if (kparts) ks_add_partials(r0, part_row, kparts, kTcBM * BN, c);
#pragma unroll
    for (int u = 0; u < 8; ++u) sts128(wr_addr + 16 * u, r0[4 * u], r0[4 * u + 1], r0[4 * u + 2], r0[4 * u + 3]);
    if (j == NCW - 1) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) release();
    } else {
      __syncwarp();
    }
    const float4 b4 = pre.b4[j % PF];
    const float4 lw = pre.lw[j % PF], lb = pre.lb[j % PF];
    const bool ln_res = e.ln_stats != nullptr;
    const int col = col_base + c * 32 + lc;
    if (col < N) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + lr;
        float4 v = lds128(rd_addr + i * 4 * kTcEpiStride * 4);
        if (rr < rows_valid) {
          const int row = row0 + rr;
          if (ln_in) {   // LayerNorm folded into this GEMM: lw = row sums of the gamma-scaled weight planes
            const float2 st = pre.st[i * 4 + lr];
            v.x = st.y * (v.x - st.x * lw.x); v.y = st.y * (v.y - st.x * lw.y);
            v.z = st.y * (v.z - st.x * lw.z); v.w = st.y * (v.w - st.x * lw.w);
          }
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          int out_row = row;
          if constexpr (FANCY) {
            const float al = e.alpha_dev ? e.alpha * __ldg(e.alpha_dev) : e.alpha;
            v.x *= al; v.y *= al; v.z *= al; v.w *= al;
            if (e.gate) {
              const float4 gt = __ldg(reinterpret_cast<const float4*>(e.gate + static_cast<size_t>(row) * e.ld_gate + col));
              const float gs = e.gate_scale;
              v.x = gt.x > 0.f ? v.x * gs : 0.f; v.y = gt.y > 0.f ? v.y * gs : 0.f;
              v.z = gt.z > 0.f ? v.z * gs : 0.f; v.w = gt.w > 0.f ? v.w * gs : 0.f;
            }
            if (e.pe || e.row_map) {
              const int b = row / e.rows_per_clip, sidx = row - b * e.rows_per_clip;
              if (e.pe) {
                const int p = e.pe_index ? __ldg(e.pe_index + b) : b;
                const float4 pv = __ldg(reinterpret_cast<const float4*>(e.pe + static_cast<size_t>(p) * e.ld_pe + col));
                v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
              }
              if (e.row_map == 1) out_row = sidx * e.clips + b;
              else if (e.row_map == 2) out_row = (sidx == e.rows_per_clip - 1) ? b : -1;
              else if (e.row_map == 3) out_row = b * e.out_clip_rows + e.out_row_off + sidx;
            }
          }
          if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          if constexpr (FANCY) {
            if (e.drop_thr) {   // training-mode dropout, mask regenerated from (row, col) in the backward pass
              const uint32_t i0 = static_cast<uint32_t>(row) * static_cast<uint32_t>(e.drop_cols) + static_cast<uint32_t>(col);
              const float ds = e.drop_scale;
              v.x = drop_hash(e.drop_key, i0) >= e.drop_thr ? v.x * ds : 0.f;
              v.y = drop_hash(e.drop_key, i0 + 1) >= e.drop_thr ? v.y * ds : 0.f;
              v.z = drop_hash(e.drop_key, i0 + 2) >= e.drop_thr ? v.z * ds : 0.f;
              v.w = drop_hash(e.drop_key, i0 + 3) >= e.drop_thr ? v.w * ds : 0.f;
            }
          }
          if (residual) {
            float4 rv = pre.res[j % PF][i];
            if (ln_res) {  // LayerNorm of the pre-norm sums, recomputed from the row statistics
              const float2 st = pre.st[i * 4 + lr];
              rv.x = (rv.x - st.x) * st.y * lw.x + lb.x; rv.y = (rv.y - st.x) * st.y * lw.y + lb.y;
              rv.z = (rv.z - st.x) * st.y * lw.z + lb.z; rv.w = (rv.w - st.x) * st.y * lw.w + lb.w;
            }
            v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
          }
          if (stat_out) {
            ssum[i] += (v.x + v.y) + (v.z + v.w);
            ssq[i] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
          }
          if (out32 && out_row >= 0) *reinterpret_cast<float4*>(out32 + static_cast<size_t>(out_row) * ld32 + col) = v;
          if (out_hi) {
            const uint2 h = pack_hi4(v, bf16);
            const size_t o = static_cast<size_t>(row) * ld16 + col;
            *reinterpret_cast<uint2*>(out_hi + o) = h;
            if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = pack_lo4(v, h);
          }
        }
      }
    }
    // this slot is free again: start the loads of chunk j + PF of this tile
    if (j + PF < NCW)
      epi_prefetch_chunk<FANCY>(args, row0, col_base + (c + PF) * 32 + lc, lr, pre.b4[j % PF], pre.res[j % PF],
                                pre.lw[j % PF], pre.lb[j % PF]);
    __syncwarp();
  }
  if (stat_out) {
    // the 8 lanes that share a row (lr fixed, lc = 0..28) add up their parts; slot = this warp's position along N
    constexpr int kSlotsPerTile = (BN / 32) / NCW;
    const int slot = (col_base / BN) * kSlotsPerTile + c0 / NCW;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = ssum[i], q = ssq[i];
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      const int rr = i * 4 + lr;
      if ((lane & 7) == 0 && rr < rows_valid) stat_out[static_cast<size_t>(row0 + rr) * e.stat_ld + slot] = make_float2(s, q);
    }
  }
}

// The embedding / output projections are the only GEMMs that scale, add positional rows or remap output rows.
inline bool epilogue_is_fancy(const Epilogue& e) {
  return e.alpha != 1.0f || e.pe != nullptr || e.row_map != 0 || e.res_clip_rows != 0 || e.alpha_dev != nullptr ||
         e.gate != nullptr || e.drop_thr != 0;
}

template <int BN, bool SPLIT, bool FANCY, int AROWS = 128>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ TcGemmArgs args) {
  using Cfg = TcCfg<BN, SPLIT, AROWS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* epi_stage = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]        MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform (see gemm_tc2.cuh)
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { SDVG_TRACE(0); if (args.trace && blockIdx.x == 0) args.trace[40] = clock64(); }
  const int M = args.M, N = args.N, K = args.K;
  const int m_tiles = (M + kTcBM - 1) / kTcBM;
  const int n_tiles = (N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = (K + kTcBK - 1) / kTcBK;
  // split-K: `ks` consecutive CTAs share one tile (ks == 1: persistent over tiles as before)
  const int ks = args.ksplit > 1 ? args.ksplit : 1;
  const int cta = static_cast<int>(blockIdx.x) / ks, kr = static_cast<int>(blockIdx.x) - cta * ks;
  const int ncta = static_cast<int>(gridDim.x) / ks;
  const int kb0 = (num_kb * kr) / ks, kb1 = (num_kb * (kr + 1)) / ks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull_bar[b], 1);
      ptx::mbar_init(&tempty_bar[b], Cfg::kEpiActive);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmA_hi);
    ptx::prefetch_tensormap(&tmB_hi);
    if (SPLIT) {
      ptx::prefetch_tensormap(&tmA_lo);
      ptx::prefetch_tensormap(&tmB_lo);
    }
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only smem / TMEM / kernel parameters: it may overlap the previous kernel's tail
  if (threadIdx.x == 0) SDVG_TRACE(1);
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) SDVG_TRACE(2);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    const uint32_t a_box_bytes = static_cast<uint32_t>((args.a_box_rows > 0 ? args.a_box_rows : kTcBM) * kTcBK * 2);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cta; t < total_tiles; t += ncta) {
      const int n_blk = t / m_tiles, m_blk = t - n_blk * m_tiles;
      // (filling the run of free stages in one go, like the MMA warp consumes the full ones, measured SLOWER here:
      // C1 launch chain 801 vs 774 us per pass)
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sp = stage_base + stage * Cfg::kStageBytes;
        uint8_t* sb = sp + Cfg::kPlanes * Cfg::kABytes;
        if (ptx::elect_one()) {
          // rows of the smem A tile beyond a_box_rows keep stale bits: they only feed accumulator rows >= M,
          // which the epilogue never stores
          ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kPlanes * (a_box_bytes + Cfg::kBBytes));
          ptx::tma_load_2d(sp, &tmA_hi, &full_bar[stage], kb * kTcBK, m_blk * kTcBM);
          if (SPLIT) ptx::tma_load_2d(sp + Cfg::kABytes, &tmA_lo, &full_bar[stage], kb * kTcBK, m_blk * kTcBM);
          ptx::tma_load_2d(sb, &tmB_hi, &full_bar[stage], kb * kTcBK, n_blk * BN);
          if (SPLIT) ptx::tma_load_2d(sb + Cfg::kBBytes, &tmB_lo, &full_bar[stage], kb * kTcBK, n_blk * BN);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    const uint32_t idesc = ptx::make_idesc_f16(kTcBM, BN, args.bf16 != 0);
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t buf_phase = 0;
    for (int t = cta; t < total_tiles; t += ncta) {
      ptx::mbar_wait(&tempty_bar[buf], buf_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d0 = tmem_base + buf * Cfg::kColsPerTile;
      const uint32_t d1 = d0 + BN;
      if constexpr (AROWS < 128) {   // (at 49..128 rows the batched loop measured 3-4 % SLOWER than one blocking wait per stage)
        // Stages are consumed in batches: lane l tests the barrier of stage (stage + l), the batch is the run of stages
        // that are already full (at least one: then the warp blocks on it), and ONE elected thread issues every MMA and
        // commit of the batch back to back.  With narrow tiles an MMA executes in 48 cycles (tools/mma_rate.cu) while one
        // wait + fence + election + descriptor set-up per K block cost ~315: the K loop of a 40 x 2048 x 2048 GEMM was 5.4 us
        // of its 8.2 us launch (tools/gemm_trace.py).
        constexpr int kBatch = Cfg::kStages < 4 ? Cfg::kStages : 4;
        const uint32_t ring_lo = ptx::kmajor_sw128_desc_lo(ptx::smem_u32(stage_base));
        int kb = kb0;
        while (kb < kb1) {
          int want = kb1 - kb < kBatch ? kb1 - kb : kBatch;
          bool ready = false;
          if (lane < want) {
            int s2 = stage + lane; uint32_t ph = phase;
            if (s2 >= Cfg::kStages) { s2 -= Cfg::kStages; ph ^= 1; }
            ready = ptx::mbar_test_wait(&full_bar[s2], ph);
          }
          const uint32_t mask = __ballot_sync(0xffffffffu, ready);
          int nst = __ffs(~mask) - 1;                 // leading run of full stages
          if (nst > want) nst = want;
          if (nst == 0) { ptx::mbar_wait(&full_bar[stage], phase); nst = 1; }
          ptx::tc_fence_after();
          if (kb == kb0 && t == cta && lane == 0) SDVG_TRACE(8);
          if (ptx::elect_one()) {
            int st = stage;
            for (int j = 0; j < nst; ++j) {
              const uint32_t a_hi = ring_lo + static_cast<uint32_t>(st) * (Cfg::kStageBytes >> 4);
              const uint32_t b_hi = a_hi + ((Cfg::kPlanes * Cfg::kABytes) >> 4);
              const uint32_t a_lo = a_hi + (Cfg::kABytes >> 4);
              const uint32_t b_lo = b_hi + (Cfg::kBBytes >> 4);
  #pragma unroll
              for (int k = 0; k < kTcBK / 16; ++k) {
                const uint32_t acc = (kb + j != kb0 || k != 0) ? 1u : 0u;
                const uint32_t adv = static_cast<uint32_t>(k * 2);  // 16 elements * 2 B = 32 B = 2 << 4
                ptx::umma_f16_lo(d0, a_hi + adv, b_hi + adv, idesc, acc);
                if (SPLIT) {
                  ptx::umma_f16_lo(d1, a_hi + adv, b_lo + adv, idesc, acc);
                  ptx::umma_f16_lo(d1, a_lo + adv, b_hi + adv, idesc, 1u);
                }
              }
              ptx::umma_commit(&empty_bar[st]);                              // smem stage reusable once these MMAs retire
              if (kb + j == kb1 - 1) {
                ptx::umma_commit(&tfull_bar[buf]);                           // accumulator(s) of this tile complete
                if (t == cta) SDVG_TRACE(16);
              }
              if (++st == Cfg::kStages) st = 0;
            }
          }
          __syncwarp();
          stage += nst;
          if (stage >= Cfg::kStages) { stage -= Cfg::kStages; phase ^= 1; }
          kb += nst;
        }
      } else {   // full-height stages / wide tiles: one blocking wait per stage is cheaper (C2 step: 49.37 vs 49.51 ms)
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kPlanes * Cfg::kABytes;
          const uint64_t a_hi = ptx::make_kmajor_sw128_desc(sa);
          const uint64_t b_hi = ptx::make_kmajor_sw128_desc(sb);
          const uint64_t a_lo = ptx::make_kmajor_sw128_desc(sa + Cfg::kABytes);
          const uint64_t b_lo = ptx::make_kmajor_sw128_desc(sb + Cfg::kBBytes);
          if (ptx::elect_one()) {
  #pragma unroll
            for (int k = 0; k < kTcBK / 16; ++k) {
              const uint32_t acc = (kb != kb0 || k != 0) ? 1u : 0u;
              const uint64_t adv = static_cast<uint64_t>(k * 2);  // 16 elements * 2 B = 32 B = 2 << 4
              ptx::umma_f16(d0, a_hi + adv, b_hi + adv, idesc, acc);
              if (SPLIT) {
                ptx::umma_f16(d1, a_hi + adv, b_lo + adv, idesc, acc);
                ptx::umma_f16(d1, a_lo + adv, b_hi + adv, idesc, 1u);
              }
            }
            ptx::umma_commit(&empty_bar[stage]);                         // smem stage reusable once these MMAs retire
            if (kb == kb1 - 1) ptx::umma_commit(&tfull_bar[buf]);        // accumulator(s) of this tile complete
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (++buf == 2) { buf = 0; buf_phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (2 per 32-lane TMEM quadrant)
    const int q = warp & 3;               // TMEM lane quadrant this warp is allowed to read
    const int half = (warp - 2) >> 2;     // which half of the tile's column chunks
    constexpr int kChunks = BN / 32;
    constexpr int NCW = Cfg::kEpiActive == 8 ? kChunks / 2 : kChunks;   // chunks per participating warp
    constexpr int PF = NCW < 2 ? NCW : 2;
    const bool active = Cfg::kEpiActive == 8 || half == 0;
    const int c0 = Cfg::kEpiActive == 8 ? half * NCW : 0;
    float* stg = epi_stage + (warp - 2) * 32 * kTcEpiStride;
    int buf = 0;
    uint32_t buf_phase = 0;
    for (int t = cta; t < total_tiles && active; t += ncta) {
      const int n_blk = t / m_tiles, m_blk = t - n_blk * m_tiles;
      const int row0 = m_blk * kTcBM + q * 32;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * Cfg::kColsPerTile;
      uint64_t* done_bar = &tempty_bar[buf];
      if (kr != 0) {
        // split-K peer: fp32 partial tile -> workspace, then publish
        ptx::mbar_wait(&tfull_bar[buf], buf_phase);
        ptx::tc_fence_after();
        float* dst = args.ks_ws + ((static_cast<size_t>(t) * (ks - 1) + (kr - 1)) * kTcBM + q * 32 + lane) * BN;
        tc_epilogue_partial<BN, SPLIT, NCW>(tbase, c0, lane, dst, [done_bar]() { ptx::mbar_arrive(done_bar); });
        if (lane == 0) atomicAdd(args.ks_flags + t, 1u);
        if (++buf == 2) { buf = 0; buf_phase ^= 1; }
        continue;
      }
      EpiPre<PF> pre;
      pre.st = reinterpret_cast<float2*>(epi_stage + kTcEpiWarps * 32 * kTcEpiStride) + (warp - 2) * 32;
      tc_epilogue_prefetch<FANCY, NCW, PF>(args, row0, n_blk * BN, lane, c0, pre);
      ptx::mbar_wait(&tfull_bar[buf], buf_phase);
      ptx::tc_fence_after();
      if (t == cta && warp == 2 && lane == 0) SDVG_TRACE(24);
      const float* part_row = nullptr;
      const unsigned int expected = static_cast<unsigned int>(Cfg::kEpiActive * (ks - 1));
      if (ks > 1) {
        // every peer warp has published its rows once the counter reaches kEpiActive * (ks - 1); all CTAs of the
        // launch are co-resident (grid <= SM count, one CTA per SM), so the peers always make progress
        if (lane == 0) {
          const volatile unsigned int* f = args.ks_flags + t;
          const long long t0 = clock64();
          while (*f < expected) {
            __nanosleep(40);
            if (clock64() - t0 > 4000000000LL) { printf("sdvg gemm: split-K wait timed out (tile %d)\n", t); __trap(); }
          }
        }
        __syncwarp();
        __threadfence();
        part_row = args.ks_ws + (static_cast<size_t>(t) * (ks - 1) * kTcBM + q * 32 + lane) * BN;
      }
      tc_epilogue_tile<BN, SPLIT, FANCY, NCW, PF>(args, stg, tbase, row0, n_blk * BN, lane, c0, pre,
                                                  [done_bar]() { ptx::mbar_arrive(done_bar); }, part_row, ks - 1);
      if (ks > 1 && lane == 0) {
        // the last reader of this tile's partials re-arms the flag for the next launch
        if (atomicAdd(args.ks_flags + t, 1u) == expected + Cfg::kEpiActive - 1) args.ks_flags[t] = 0u;
      }
      if (t == cta && warp == 2 && lane == 0) SDVG_TRACE(32);
      if (++buf == 2) { buf = 0; buf_phase ^= 1; }
    }
  }

  __syncthreads();
  if (threadIdx.x == 0) SDVG_TRACE(4);
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
  if (threadIdx.x == 0) { SDVG_TRACE(5); if (args.trace && blockIdx.x == 0) args.trace[41] = clock64(); }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D map over a row-major 16-bit matrix [rows][cols] (cols contiguous, row pitch ld elements):
// box = 64 columns (128 B, one swizzle span) x box_rows rows.
inline bool make_tmap_2d(CUtensorMap* out, const void* base, int rows, int cols, int ld, int box_rows, bool bf16) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kTcBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline bool epilogue_vec4_ok(const Epilogue& e, int N) {
  auto al = [](const void* p, int bytes) { return (reinterpret_cast<uintptr_t>(p) % bytes) == 0; };
  if (N % 4 != 0) return false;
  if (e.bias && !al(e.bias, 16)) return false;
  if (e.pe && (!al(e.pe, 16) || e.ld_pe % 4 != 0)) return false;
  if (e.residual && (!al(e.residual, 16) || e.ld_res % 4 != 0)) return false;
  if (e.ln_stats && (!al(e.ln_stats, 8) || !al(e.ln_w, 16) || !al(e.ln_b, 16))) return false;
  if (e.gate && (!al(e.gate, 16) || e.ld_gate % 4 != 0)) return false;
  if (e.out32 && (!al(e.out32, 16) || e.ld32 % 4 != 0)) return false;
  if (e.out_hi && (!al(e.out_hi, 8) || e.ld16 % 4 != 0)) return false;
  if (e.out_lo && !al(e.out_lo, 8)) return false;
  return true;
}

constexpr int kTcCompactRows = 48;   // compact A stages (TcCfg AROWS) when every token row fits the 48-row TMA box

template <int BN, bool SPLIT, bool FANCY, int AROWS = 128>
inline cudaError_t launch_gemm_tc_f(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                                    const CUtensorMap& b_lo, const TcGemmArgs& args, int num_sms,
                                    cudaStream_t stream) {
  if constexpr (AROWS == 128 && !SPLIT && BN <= 64) {
    if (args.M <= kTcCompactRows && args.a_box_rows > 0 && args.a_box_rows <= kTcCompactRows && args.ksplit <= 1)
      return launch_gemm_tc_f<BN, SPLIT, FANCY, kTcCompactRows>(a_hi, a_lo, b_hi, b_lo, args, num_sms, stream);
  }
  using Cfg = TcCfg<BN, SPLIT, AROWS>;
  static bool attr_set[64] = {};  // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, SPLIT, FANCY, AROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  const int tiles = ceil_div(args.M, kTcBM) * ceil_div(args.N, BN);
  int grid = tiles < num_sms ? tiles : num_sms;
  if (args.ksplit > 1) {
    // split-K needs every CTA of the launch resident at once (rank 0 spins on its peers)
    if (tiles * args.ksplit > num_sms || !args.ks_ws || !args.ks_flags || args.ksplit > ceil_div(args.K, kTcBK))
      return cudaErrorInvalidConfiguration;
    grid = tiles * args.ksplit;
  }
  TcGemmArgs a2 = args;
  a2.vec4 = epilogue_vec4_ok(args.epi, args.N) ? 1 : 0;
  return launch_kernel(gemm_tc_kernel<BN, SPLIT, FANCY, AROWS>, dim3(grid), dim3(kTcThreads), Cfg::kSmemBytes, stream, a_hi, a_lo, b_hi, b_lo, a2);
}

template <int BN, bool SPLIT>
inline cudaError_t launch_gemm_tc_t(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                                    const CUtensorMap& b_lo, const TcGemmArgs& args, int num_sms,
                                    cudaStream_t stream) {
  if (epilogue_is_fancy(args.epi)) return launch_gemm_tc_f<BN, SPLIT, true>(a_hi, a_lo, b_hi, b_lo, args, num_sms, stream);
  return launch_gemm_tc_f<BN, SPLIT, false>(a_hi, a_lo, b_hi, b_lo, args, num_sms, stream);
}

}  // namespace sdvg
