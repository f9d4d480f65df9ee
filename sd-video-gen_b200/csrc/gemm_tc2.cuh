// K1b - CTA-pair tensor-core GEMM (tcgen05 cta_group::2):  C[M,N] = A[M,K] * W[N,K]^T + Epilogue.
//
// Why a second kernel: the one-CTA kernel (gemm_tc.cuh) streams 48 KB of operands per 128x256x64 MMA block
// through L2->SMEM; ncu shows its tensor pipe ~44 % active with neither L2 nor DRAM saturated - the TMA ring
// cannot keep enough bytes in flight.  Pairing two SMs on a 256 x BN tile halves the W traffic of each CTA
// (each loads 128 rows of A and BN/2 rows of W; the pair MMA reads both halves), so a stage is 28-32 KB, the
// ring is 6-7 stages deep and operand bytes per FLOP drop 1.5x.
//
//   cluster = 2 CTAs (one TPC), leader = cluster rank 0, persistent over 256 x BN tiles
//   warp 0 (both CTAs): TMA producer - own A rows / own half of W, completion bytes land on the LEADER's full
//                       barrier (count 2: one arrive.expect_tx per CTA)
//   warp 1 (leader)   : issues tcgen05.mma.cta_group::2 (M=256, N=BN, K=16); tcgen05.commit ... multicast
//                       releases the smem stage / publishes the accumulator in BOTH CTAs
//   warps 2-9 (both)  : epilogue of the CTA's own 128 accumulator rows, two warps per TMEM lane quadrant (code
//                       shared with gemm_tc.cuh); they hand the TMEM buffer back by arriving on the leader's
//                       barrier (count 16)
// BN in {64, 128, 192, 256}: 192 exists because at M=5120 (1024 clips x 5 tokens) an N=2048 GEMM is 160 pair
// tiles of 256x256 for 74 pairs (3 rounds, 72 % filled) but 220 tiles of 256x192 (3 rounds, 99 % filled, each
// 25 % cheaper).
#pragma once
#include "gemm_tc.cuh"

namespace sdvg {

constexpr int kTc2BM = 256;  // rows per pair tile (128 per CTA)

template <int BN, bool SPLIT>
struct Tc2Cfg {
  static_assert(BN == 64 || BN == 128 || BN == 192 || BN == 256, "BN");
  static_assert(!(SPLIT && BN > 128), "split mode needs two accumulators per tile: BN <= 128");
  static constexpr int kPlanes = SPLIT ? 2 : 1;
  static constexpr int kABytes = kTcBM * kTcBK * 2;       // this CTA's 128 rows of A
  static constexpr int kBRows = BN / 2;                    // this CTA's half of the W tile
  static constexpr int kBBytes = kBRows * kTcBK * 2;
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kEpiBytes = kTcEpiWarps * 32 * kTcEpiStride * 4 + kTcEpiWarps * 32 * 8;   // transpose tiles + (mean, rstd) of each warp's 32 rows
  static constexpr int kBarBytes = 1024;
  static constexpr int kMaxStages = (kTcSmemLimit - 1024 - kEpiBytes - kBarBytes) / kStageBytes;
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kColsPerTile = BN * kPlanes;
  static constexpr int kTmemCols = 2 * kColsPerTile <= 32 ? 32 : 2 * kColsPerTile <= 64 ? 64 : 2 * kColsPerTile <= 128 ? 128
                                 : 2 * kColsPerTile <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiBytes + kBarBytes;
  static_assert(kBBytes % 1024 == 0, "swizzle atom alignment");
  static_assert(2 * kColsPerTile <= 512, "TMEM");
};

template <int BN, bool SPLIT, bool FANCY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ TcGemmArgs args) {
  using Cfg = Tc2Cfg<BN, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* epi_stage = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;                      // [kStages] used in the leader: both producers -> MMA
  uint64_t* empty_bar = bars + Cfg::kStages;      // [kStages] per CTA: MMA (multicast commit) -> own producer
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;  // [2] per CTA: MMA (multicast commit) -> own epilogue
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] used in the leader: both epilogues -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  // warp index / cluster rank as provably warp-uniform values: the producer and MMA roles run with the whole
  // warp converged and elect one lane per instruction, so that TMA / tcgen05 operands live in uniform registers.
  // (With `if (lane == 0)` around the role loops ptxas wrapped every UTMALDG / UTCHMMA / UTCBAR in an
  // ELECT + R2UR.BROADCAST + BRA.U.ANY loop: ~540 cycles per K block regardless of tile width.)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0);
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int M = args.M, N = args.N, K = args.K;
  const int m_tiles = (M + kTc2BM - 1) / kTc2BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = (K + kTcBK - 1) / kTcBK;
  const int nstages = (args.max_stages > 0 && args.max_stages < Cfg::kStages) ? args.max_stages : Cfg::kStages;
  if (threadIdx.x == 0) { SDVG_TRACE(0); if (args.trace && blockIdx.x == 0) args.trace[40] = clock64(); }

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 2);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull_bar[b], 1);
      ptx::mbar_init(&tempty_bar[b], 2 * kTcEpiWarps);
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
    ptx::tmem_alloc_cg2(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish_cg2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // peer barriers are initialised before anyone signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) SDVG_TRACE(1);
  pdl_wait();  // prologue above overlaps the previous kernel's tail (see common.cuh)
  pdl_trigger();
  if (threadIdx.x == 0) SDVG_TRACE(2);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < total_tiles; t += num_pairs) {
      const int n_blk = t / m_tiles, m_blk = t - n_blk * m_tiles;
      const int row_a = m_blk * kTc2BM + static_cast<int>(rank) * kTcBM;
      const int row_b = n_blk * BN + static_cast<int>(rank) * Cfg::kBRows;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sp = stage_base + stage * Cfg::kStageBytes;
        uint8_t* sb = sp + Cfg::kPlanes * Cfg::kABytes;
        const uint32_t lead_full = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx_cluster(lead_full, Cfg::kStageBytes);
          ptx::tma_load_2d_cg2(sp, &tmA_hi, lead_full, kb * kTcBK, row_a);
          if (SPLIT) ptx::tma_load_2d_cg2(sp + Cfg::kABytes, &tmA_lo, lead_full, kb * kTcBK, row_a);
          ptx::tma_load_2d_cg2(sb, &tmB_hi, lead_full, kb * kTcBK, row_b);
          if (SPLIT) ptx::tma_load_2d_cg2(sb + Cfg::kBBytes, &tmB_lo, lead_full, kb * kTcBK, row_b);
        }
        __syncwarp();
        if (t == pair && kb == 0 && lane == 0) SDVG_TRACE(3);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; whole warp, one elected lane issues)
    if (rank == 0) {
      const uint32_t idesc = ptx::make_idesc_f16(kTc2BM, BN, args.bf16 != 0);
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t buf_phase = 0;
      int tile_no = 0;
      for (int t = pair; t < total_tiles; t += num_pairs, ++tile_no) {
        ptx::mbar_wait(&tempty_bar[buf], buf_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + buf * Cfg::kColsPerTile;
        const uint32_t d1 = d0 + BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (kb == 0 && tile_no < 8 && lane == 0) SDVG_TRACE(8 + tile_no);
          const uint32_t sa = ptx::smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kPlanes * Cfg::kABytes;
          const uint64_t a_hi = ptx::make_kmajor_sw128_desc(sa);
          const uint64_t b_hi = ptx::make_kmajor_sw128_desc(sb);
          const uint64_t a_lo = ptx::make_kmajor_sw128_desc(sa + Cfg::kABytes);
          const uint64_t b_lo = ptx::make_kmajor_sw128_desc(sb + Cfg::kBBytes);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kTcBK / 16; ++k) {
              const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
              const uint64_t adv = static_cast<uint64_t>(k * 2);
              ptx::umma_f16_cg2(d0, a_hi + adv, b_hi + adv, idesc, acc);
              if (SPLIT) {
                ptx::umma_f16_cg2(d1, a_hi + adv, b_lo + adv, idesc, acc);
                ptx::umma_f16_cg2(d1, a_lo + adv, b_hi + adv, idesc, 1u);
              }
            }
            ptx::umma_commit_cg2_mc(&empty_bar[stage], 0x3);
            if (kb == num_kb - 1) ptx::umma_commit_cg2_mc(&tfull_bar[buf], 0x3);
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (tile_no < 8 && lane == 0) SDVG_TRACE(16 + tile_no);
        if (++buf == 2) { buf = 0; buf_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (own 128 rows, 2 per quadrant)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int NCW = BN / 64;              // 32-column chunks per warp
    constexpr int PF = NCW < 2 ? NCW : 2;     // chunks whose bias / residual operands are loaded ahead (3 spills registers)
    const int c0 = half * NCW;
    float* stg = epi_stage + (warp - 2) * 32 * kTcEpiStride;
    int buf = 0;
    uint32_t buf_phase = 0;
    int tile_no = 0;
    for (int t = pair; t < total_tiles; t += num_pairs, ++tile_no) {
      const int n_blk = t / m_tiles, m_blk = t - n_blk * m_tiles;
      const int row0 = m_blk * kTc2BM + static_cast<int>(rank) * kTcBM + q * 32;
      EpiPre<PF> pre;
      pre.st = reinterpret_cast<float2*>(epi_stage + kTcEpiWarps * 32 * kTcEpiStride) + (warp - 2) * 32;
      tc_epilogue_prefetch<FANCY, NCW, PF>(args, row0, n_blk * BN, lane, c0, pre);   // in flight during the main loop
      ptx::mbar_wait(&tfull_bar[buf], buf_phase);
      ptx::tc_fence_after();
      if (warp == 2 && lane == 0 && tile_no < 8) SDVG_TRACE(24 + tile_no);
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * Cfg::kColsPerTile;
      const uint32_t lead_tempty = ptx::mapa_u32(ptx::smem_u32(&tempty_bar[buf]), 0);
      tc_epilogue_tile<BN, SPLIT, FANCY, NCW, PF>(args, stg, tbase, row0, n_blk * BN, lane, c0, pre,
                                                  [lead_tempty]() { ptx::mbar_arrive_cluster(lead_tempty); });
      if (warp == 2 && lane == 0 && tile_no < 8) SDVG_TRACE(32 + tile_no);
      if (++buf == 2) { buf = 0; buf_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SDVG_TRACE(4);
  ptx::cluster_sync_all();  // nobody leaves while the peer may still signal or read this CTA
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols);
  }
  if (threadIdx.x == 0) { SDVG_TRACE(5); if (args.trace && blockIdx.x == 0) args.trace[41] = clock64(); }
}

template <int BN, bool SPLIT, bool FANCY>
inline cudaError_t launch_gemm_tc2_f(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                                     const CUtensorMap& b_lo, const TcGemmArgs& args, int num_sms,
                                     cudaStream_t stream) {
  using Cfg = Tc2Cfg<BN, SPLIT>;
  static bool attr_set[64] = {};  // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<BN, SPLIT, FANCY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  const int tiles = ceil_div(args.M, kTc2BM) * ceil_div(args.N, BN);
  const int pairs = num_sms / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  TcGemmArgs a2 = args;
  a2.vec4 = epilogue_vec4_ok(args.epi, args.N) ? 1 : 0;
  return launch_kernel(gemm_tc2_kernel<BN, SPLIT, FANCY>, dim3(grid), dim3(kTcThreads), Cfg::kSmemBytes, stream, a_hi, a_lo, b_hi, b_lo, a2);
}

template <int BN, bool SPLIT>
inline cudaError_t launch_gemm_tc2_t(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                                     const CUtensorMap& b_lo, const TcGemmArgs& args, int num_sms,
                                     cudaStream_t stream) {
  if (epilogue_is_fancy(args.epi)) return launch_gemm_tc2_f<BN, SPLIT, true>(a_hi, a_lo, b_hi, b_lo, args, num_sms, stream);
  return launch_gemm_tc2_f<BN, SPLIT, false>(a_hi, a_lo, b_hi, b_lo, args, num_sms, stream);
}

}  // namespace sdvg
