// K4 - persistent small-batch kernel: one launch runs a whole "program" of the model's ops (GEMM, LayerNorm,
// attention, window plumbing) for M = clips x tokens <= 128 rows - a full forward pass, or every pass of a rollout
// (prediction/predict.py:143-197, whose real regime is batch 1).  BASELINE's north star calls it the "persistent
// rollout kernel"; it exists because at small batch the pass is pure weight streaming (0.875 GB of 16-bit weights per
// pass for the d2048 4e/8d model = 0.134 ms at the measured HBM rate) while the per-kernel path is ~118 dependent
// launches of 5-9 us of fixed cost each (0.86 ms per pass, profiles/README.md round 1).
//
// Decomposition of one GEMM  C[M,N] = X[M,K] W[N,K]^T  over the machine (33 clusters x 4 CTAs, one CTA per SM - what
// cudaOccupancyMaxActiveClusters gives on B200: 8-CTA clusters only fit 15 times, which made a 16-tile GEMM two rounds):
//   * the weight matrix is cut along N into tiles of T <= 128 rows, T chosen per GEMM so that the tiles fill the clusters
//     in whole rounds (N = 2048: 32 tiles of 64 rows, one round), and along K into 4 slabs, one per CTA of the cluster.
//     Every weight byte is read from HBM exactly once, by exactly one SM.
//   * a CTA needs only ITS K slab of the activations (M x K/4, from L2).  With an N-only split every CTA would pull
//     all of X for every GEMM: 2.5x the weight bytes through L2 -> SM at M = 40, more than the HBM stream itself.
//   * tcgen05.mma runs "swapped": the weight tile is the M = 128 operand (lanes beyond T compute garbage that is never
//     read), the token rows are the N operand (N = M rounded up to 16), so the accumulator D[n][m] in TMEM has one
//     weight row per lane and one token per column, and the MMA cost scales with the token count.
//   * the 4 K-partials of a tile are reduced through distributed shared memory: each CTA stages its partial tile in
//     shared memory, one bulk copy (cp.async.bulk shared::cta -> shared::cluster, completion bytes on the owner's
//     mbarrier - no fences, no polling) per owner hands every CTA of the cluster the 16 token rows it owns; the owner
//     sums the 4 partials in fixed order (deterministic) and applies the fused Epilogue (common.cuh), coalesced along
//     the model width.
// Warp roles per CTA: warp 0 streams the CTA's weight blocks through a TMA ring and runs AHEAD of the dependency
// chain (weights depend on nothing), so HBM keeps streaming while the other roles wait at a barrier; warp 2 issues the
// MMAs; warps 4-7 are the workers: they copy the CTA's activation slab from L2 into the swizzled operand layout
// (plain 16-byte loads: a third of the latency of a TMA round trip for a 5 KB box), run the reduction + epilogue, and
// execute the row ops (LayerNorm: one row per CTA, attention: one warp per (clip, head, query)).
// Ops are separated by a grid barrier (one atomic counter in L2, relaxed polls + one acquire fence); every wait is
// bounded and traps instead of hanging the GPU.
#pragma once
#include <cstring>

#include "attention.cuh"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "layernorm.cuh"
#include "pack.cuh"
#include "ptx.cuh"

namespace sdvg {

constexpr int kPkThreads = 256;
constexpr int kPkCluster = 4;
constexpr int kPkTileN = 128;                          // UMMA M: weight rows per MMA tile (T <= 128 of them valid)
constexpr int kPkWSlotBytes = kPkTileN * kTcBK * 2;    // one slot of the weight ring: 16 KB = floor(128 / T) K blocks of T rows
constexpr int kPkMaxRows = 128;                        // tokens per program (UMMA N)
constexpr int kPkRpo = 16;                             // token rows per owner CTA and exchange group
constexpr int kPkGroup = kPkCluster * kPkRpo;          // tokens per exchange group (64)
constexpr int kPkRedBytes = kPkCluster * kPkTileN * kPkRpo * 4;   // partial tiles received from the 4 sources (32 KB)
constexpr long long kPkSpinLimit = 4000000000LL;
constexpr int kPkChunk = 4;                            // weight-slot groups the MMA warp awaits and issues as one batch

enum PkType : int { PK_GEMM = 1, PK_LN = 2, PK_ATTN = 3, PK_PACK = 4, PK_ADD = 5 };

struct PkGemm {
  int M, N, K, MT;              // token rows, weight rows, reduction length, M rounded up to 16
  int split, bf16;
  int w_hi, w_lo;               // tensor-map table indices of the weight planes (box = T rows x 64 columns)
  int T, n_tiles;               // rows per weight tile, number of tiles
  int lda, pad;                 // row pitch (elements) of the activation planes
  const uint16_t* a_hi;         // activation operand planes [rows][lda] (rows >= MT)
  const uint16_t* a_lo;
  Epilogue epi;
};

struct alignas(16) PkOp {
  int type;
  int in16;                     // attention: Q/K/V are 16-bit planes
  int pad[2];
  union U {
    PkGemm g;
    LnArgs ln;
    AttnArgs at;
    PackArgs pk;
    AddArgs ad;
    U() { std::memset(static_cast<void*>(this), 0, sizeof(*this)); }
  } u;
};

static_assert(sizeof(PkOp) <= 1024 && sizeof(PkOp) % 4 == 0, "the workers keep a copy of the current op in 1 KB of shared memory");

struct PkParams {
  const PkOp* ops;
  int n_ops;
  const CUtensorMap* maps;
  unsigned int* sync;           // [0]: arrivals (every CTA once per op), reset to 0 by the last CTA out
  int w_slots, a_stages, a_stage_bytes;
  unsigned long long* trace;    // optional [n_ops][32] %globaltimer stamps of CTA `trace_cta` (SDVG_PK_TRACE)
  int trace_cta;
};

#define PK_TRACE(ev) do { if (P.trace && static_cast<int>(blockIdx.x) == P.trace_cta) P.trace[static_cast<size_t>(oi) * 32 + (ev)] = global_timer(); } while (0)

// Tile height of a GEMM with N weight rows on `clusters` clusters: whole rounds, tiles as even as possible.
inline int pk_tile_rows(int N, int clusters) {
  const int rounds = (N + clusters * kPkTileN - 1) / (clusters * kPkTileN);
  int t = (N + rounds * clusters - 1) / (rounds * clusters);
  t = (t + 7) / 8 * 8;
  return t > kPkTileN ? kPkTileN : (t < 8 ? 8 : t);
}

namespace pkx {

__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_relaxed_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t ncluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
// warp-uniform copies of values read from memory: TMA / tcgen05 operands must sit in uniform registers, otherwise
// ptxas wraps every UTMALDG / UTCHMMA in an ELECT + R2UR waterfall (~600 cycles per K block, measured)
__device__ __forceinline__ int uni(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ const void* uni_ptr(const void* p) {
  const unsigned long long v = reinterpret_cast<unsigned long long>(p);
  const unsigned lo = __shfl_sync(0xffffffffu, static_cast<unsigned>(v), 0);
  const unsigned hi = __shfl_sync(0xffffffffu, static_cast<unsigned>(v >> 32), 0);
  return reinterpret_cast<const void*>((static_cast<unsigned long long>(hi) << 32) | lo);
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_local, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_local), "r"(bytes), "r"(bar_cluster)
               : "memory");
}

// tcgen05.mma with the descriptors given as (low word, constant high word): the issuing warp is alone on its scheduler,
// so the MMA block loop is bound by its instruction count - 64-bit descriptor arithmetic on the uniform datapath made
// it ~130 instructions (670 cycles) per 64-column block.  Only the 14-bit start-address field changes.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // stride 1024 B, version 1, SWIZZLE_128B (bits 32..63)
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}
// the four K = 16 steps of one 64-column block (operand start addresses advance by 32 bytes = 2 descriptor units)
// One accumulator per term: tools/mma_rate.cu measures 48 cycles per 128 x (<= 64) x 16 MMA whether the MMAs of a tile
// accumulate into one, two or four TMEM regions (the first version alternated even / odd K steps between two
// accumulators and paid two more TMEM loads per staged block for it).
template <bool SPLIT>
__device__ __forceinline__ void mma_block(uint32_t d0, uint32_t d1, uint32_t mtp, uint32_t w_hi, uint32_t w_lo, uint32_t a_hi, uint32_t a_lo,
                                          uint32_t idesc, bool first) {
#pragma unroll
  for (int k = 0; k < kTcBK / 16; ++k) {
    const uint32_t acc = (!first || k >= 1) ? 1u : 0u;
    umma_lo(d0, w_hi + 2 * k, a_hi + 2 * k, idesc, acc);
    if (SPLIT) {   // split modes use fp16 hi planes, lo planes are always fp16: one instruction descriptor
      umma_lo(d1, w_hi + 2 * k, a_lo + 2 * k, idesc, acc);
      umma_lo(d1, w_lo + 2 * k, a_hi + 2 * k, idesc, 1u);
    }
  }
}

// Grid barrier wait: every CTA adds 1 to the counter when it has finished an op, so "all CTAs are done with every op
// before `op`" is counter >= op * gridDim.x (no CTA can be more than one op ahead of the slowest one).
// The poll is a RELAXED load (ld.acquire.gpu compiles to LDG + CCTL.IVALL: an L1 invalidation per poll); one acquire
// fence after the poll succeeds.
__device__ __forceinline__ void grid_wait(const unsigned* ctr, unsigned target) {
  if (ld_relaxed_gpu(ctr) < target) {
    const long long t0 = clock64();
    while (ld_relaxed_gpu(ctr) < target) {
      if (clock64() - t0 > kPkSpinLimit) {
        printf("sdvg persistent: grid barrier timed out (block %d thread %d target %u have %u)\n", blockIdx.x, threadIdx.x, target,
               ld_relaxed_gpu(ctr));
        __trap();
      }
    }
  }
  fence_acq_rel_gpu();
}

// K blocks [kb0, kb1) of CTA `rank` of a cluster
__device__ __forceinline__ void k_range(int nk, int rank, int& kb0, int& kb1) {
  kb0 = (nk * rank) / kPkCluster;
  kb1 = (nk * (rank + 1)) / kPkCluster;
}

// Activation slab loader.  K block `kb` of the activation plane (MT rows x 64 elements) goes into an operand stage in
// the 128-byte swizzled K-major layout tcgen05.mma reads (what TMA's SWIZZLE_128B would have produced): row r, 16-byte
// chunk c lands at r * 128 + ((c ^ (r & 7)) * 16).  Asynchronous copies (LDGSTS, L2 -> shared memory without a register
// round trip): the blocks of a whole K slab are in flight together, and each thread's arrival on the stage's mbarrier
// fires when its copies have landed (cp.async.mbarrier.arrive.noinc).  A thread's (row, chunk) pieces are the same for
// every block of an op, so their addresses are computed once per op (first version: one exposed L2 round trip per
// block, 8 us per GEMM; second: 30 instructions of address arithmetic per piece, 0.35 us per block).
struct APieces {
  const uint16_t* src[8];   // global address of the piece in K block 0 (hi plane)
  uint32_t dst[8];          // swizzled offset inside a stage
  int n;
};
__device__ __forceinline__ void a_pieces_init(APieces& ap, const uint16_t* plane, int lda, int MT, int wt) {
  const int n16 = MT * 8;
  ap.n = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = wt + i * 128;
    const int r = idx >> 3, c = idx & 7;
    ap.src[i] = plane + static_cast<size_t>(r) * lda + c * 8;
    ap.dst[i] = r * 128 + ((c ^ (r & 7)) << 4);
    if (idx < n16) ap.n = i + 1;
  }
}
template <int NP>
__device__ __forceinline__ void load_a_block(const APieces& ap, long long plane_off, int kb, uint32_t stage_addr) {
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    if (i < ap.n)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage_addr + ap.dst[i]), "l"(ap.src[i] + plane_off + kb * kTcBK) : "memory");
  }
}
// every block of a K slab into consecutive ring stages starting at `s` (the ring holds the whole slab)
template <int NP>
__device__ __forceinline__ void issue_slab(const APieces& ap, long long lo_off, bool split, int kb0, int kb1, int s, int a_stages,
                                           uint32_t a_ring_addr, uint32_t a_stage_bytes) {
  // (unrolled: one LDGSTS holds its address registers until the load/store unit has read them - with one K block per
  // loop iteration every iteration waited ~190 cycles for the previous block's copies to release theirs: 0.8 us per slab)
#pragma unroll 4
  for (int kb = kb0; kb < kb1; ++kb) {
    load_a_block<NP>(ap, 0, kb, a_ring_addr + static_cast<uint32_t>(s) * a_stage_bytes);
    if (++s == a_stages) s = 0;
    if (split) {
      load_a_block<NP>(ap, lo_off, kb, a_ring_addr + static_cast<uint32_t>(s) * a_stage_bytes);
      if (++s == a_stages) s = 0;
    }
  }
}
// (separate from the copies: ARRIVES.LDGSTSBAR holds the warp until its outstanding copies have landed - signalling
// after every block serialised the slab into one L2 round trip per block)
__device__ __forceinline__ void mbar_arrive_count(uint32_t bar_addr, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(count) : "memory");
}
__device__ __forceinline__ void a_block_arrive(uint32_t bar_addr) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_addr) : "memory");
}

// Stage this CTA's partial accumulator for one group of 64 tokens (one weight row per thread, token columns in TMEM):
// stage[r][n][16] holds tokens [tok0 + 16 r, +16) of weight row n - the block owner r of the cluster expects from this
// source, so one bulk DSMEM copy per owner moves it.  (The first version pushed 16-byte st.async packets straight from
// registers: 1280 remote transactions per tile, each updating the owner's mbarrier; the owners' own shared-memory
// reads queued behind them for microseconds.)
__device__ __forceinline__ void stage_partial(uint32_t tacc, uint32_t mtp, bool split, int wt, int tok0, int owners, uint32_t stage_local) {
  uint32_t v[kPkCluster][16];
#pragma unroll
  for (int r = 0; r < kPkCluster; ++r)
    if (r < owners) tmem_ld_x16(tacc + tok0 + r * kPkRpo, v[r]);
  ptx::tmem_ld_wait();
  if (split) {
#pragma unroll
    for (int r = 0; r < kPkCluster; ++r) {
      if (r < owners) {
        uint32_t w[16];
        tmem_ld_x16(tacc + 2 * mtp + tok0 + r * kPkRpo, w);    // cross terms, kept scaled by 2^11
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[r][j] = __float_as_uint(fmaf(__uint_as_float(w[j]), kSplitInv, __uint_as_float(v[r][j])));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kPkCluster; ++r) {
    if (r < owners) {
      const uint32_t dst = stage_local + ((r * kPkTileN + wt) * kPkRpo) * 4;
#pragma unroll
      for (int j = 0; j < 16; j += 4) sts128(dst + j * 4, v[r][j], v[r][j + 1], v[r][j + 2], v[r][j + 3]);
    }
  }
}

// Owner side.  Thread mapping of the owner's 16 token rows x T weight rows (= output columns): consecutive threads take
// consecutive columns (coalesced global access); narrow tiles put 2 or 4 threads on a column, each with 8 or 4 rows.
// RPT = rows per thread.
template <int RPT>
struct EpiOperands {
  float bias, lw, lb;
  float res[RPT];
  float2 st[RPT];
};

// Operands of the fused epilogue, loaded BEFORE the owner blocks on the partial tiles of its peers so that their L2
// latency hides behind that wait.  Lean path only (identity row mapping); the rare paths load in place.  Absent
// operands get neutral values so that the lean epilogue below is one branch-free expression per element.
template <int RPT>
__device__ __forceinline__ void epi_prefetch(const Epilogue& e, int M, int N, int row0, int n, bool lean, EpiOperands<RPT>& o) {
  o.bias = 0.f; o.lw = 1.f; o.lb = 0.f;
#pragma unroll
  for (int j = 0; j < RPT; ++j) { o.res[j] = 0.f; o.st[j] = make_float2(0.f, 1.f); }
  if (n >= N) return;
  if (e.bias) o.bias = __ldg(e.bias + n);
  const float2* stats = e.ln_stats;
  if (stats) { o.lw = __ldg(e.ln_w + n); o.lb = __ldg(e.ln_b + n); }
  const float* res = e.residual;
  if (res && lean) {
    const int ld = e.ld_res;
    const float* p = res + static_cast<size_t>(row0) * ld + n;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      if (row0 + j < M) {
        o.res[j] = __ldcg(p + static_cast<size_t>(j) * ld);
        if (stats) o.st[j] = __ldcg(stats + row0 + j);
      }
    }
  }
}

// Sum the partials of the cluster's CTAs in rank order, then the fused epilogue (same arithmetic and order as
// epi_value, common.cuh).  Operands that other CTAs wrote during this launch (residual stream, row statistics, saved
// activations) are read with ld.global.cg - L1 is not coherent across SMs.  The workers are one warp per scheduler, so
// this code is bound by its instruction count: every Epilogue field is copied to a register first (the struct sits in
// shared memory behind a generic pointer - with the output stores in between the compiler otherwise reloads a dozen
// fields per row), the lean path has no branches, pointers advance by the row pitch.
template <int RPT>
__device__ __forceinline__ void reduce_epilogue(const Epilogue& ein, int M, int N, const float* red, int col, int jrow0, int row0,
                                                int n, uint32_t src_mask, bool lean, const EpiOperands<RPT>& o) {
  float acc[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) acc[j] = 0.f;
#pragma unroll
  for (int s = 0; s < kPkCluster; ++s) {
    if (!((src_mask >> s) & 1u)) continue;
    const float4* p = reinterpret_cast<const float4*>(red + (s * kPkTileN + col) * kPkRpo + jrow0);
#pragma unroll
    for (int i = 0; i < RPT / 4; ++i) {
      const float4 t = p[i];
      acc[4 * i] += t.x; acc[4 * i + 1] += t.y; acc[4 * i + 2] += t.z; acc[4 * i + 3] += t.w;
    }
  }
  if (n >= N) return;
  float* const out32 = ein.out32; uint16_t* const out_hi = ein.out_hi; uint16_t* const out_lo = ein.out_lo;
  const size_t ld32 = ein.ld32, ld16 = ein.ld16;
  const bool relu = ein.relu != 0, bf16 = ein.bf16 != 0;
  if (lean) {
    const float alpha = ein.alpha;
    const int nvalid = M - row0;
    float* p32 = out32 + static_cast<size_t>(row0) * ld32 + n;
    uint16_t* ph = out_hi + static_cast<size_t>(row0) * ld16 + n;
    uint16_t* pl = out_lo + static_cast<size_t>(row0) * ld16 + n;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      float v = (acc[j] + o.bias) * alpha;
      v = relu ? fmaxf(v, 0.0f) : v;
      v += (o.res[j] - o.st[j].x) * o.st[j].y * o.lw + o.lb;      // neutral operands when there is no residual / LayerNorm
      const uint16_t hf = __half_as_ushort(__float2half_rn(v));
      const uint16_t hb = __bfloat16_as_ushort(__float2bfloat16_rn(v));
      const uint16_t h = bf16 ? hb : hf;
      const uint16_t l = __half_as_ushort(__float2half_rn((v - __half2float(__ushort_as_half(hf))) * kSplitScale));
      const bool ok = j < nvalid;
      if (ok && out32) *p32 = v;
      if (ok && out_hi) *ph = h;
      if (ok && out_lo) *pl = l;
      p32 += ld32; ph += ld16; pl += ld16;
    }
    return;
  }
  // everything else (embedding: scale + positional rows; output projection / caches: row remapping; training: gates,
  // dropout, device-side scale): a few GEMMs per pass
  const Epilogue e = ein;
  const float alpha = e.alpha_dev ? e.alpha * __ldcg(e.alpha_dev) : e.alpha;
#pragma unroll 1
  for (int j = 0; j < RPT; ++j) {
    const int row = row0 + j;
    if (row >= M) break;
    float v = (acc[j] + o.bias) * alpha;
    if (e.gate) v = __ldcg(e.gate + static_cast<size_t>(row) * e.ld_gate + n) > 0.0f ? v * e.gate_scale : 0.0f;
    const int pe_row = epi_pe_row(e, row);
    if (pe_row >= 0) v += __ldg(e.pe + static_cast<size_t>(pe_row) * e.ld_pe + n);
    if (relu) v = fmaxf(v, 0.0f);
    if (e.drop_thr)
      v = drop_hash(e.drop_key, static_cast<uint32_t>(row) * static_cast<uint32_t>(e.drop_cols) + static_cast<uint32_t>(n)) >= e.drop_thr
              ? v * e.drop_scale : 0.0f;
    if (e.residual) {
      const int rrow = epi_res_row(e, row);
      float r = __ldcg(e.residual + static_cast<size_t>(rrow) * e.ld_res + n);
      if (e.ln_stats) {
        const float2 st = __ldcg(e.ln_stats + rrow);
        r = (r - st.x) * st.y * o.lw + o.lb;
      }
      v += r;
    }
    const int out_row = epi_out_row(e, row);
    if (out32 && out_row >= 0) out32[static_cast<size_t>(out_row) * ld32 + n] = v;
    if (out_hi) {
      const uint16_t h = to_plane_hi(v, bf16);
      out_hi[static_cast<size_t>(row) * ld16 + n] = h;
      if (out_lo) out_lo[static_cast<size_t>(row) * ld16 + n] = to_plane_lo(v, h);
    }
  }
}

// owner's work for one tile and token group, RPT rows per thread
template <int RPT>
__device__ __forceinline__ void owner_tile(const PkGemm& g, const float* red, int wt, int row_base, int n0, int tile_rows, uint32_t src_mask,
                                           uint64_t* red_full, uint32_t& full_phase, bool lean) {
  constexpr int TPC = kPkRpo / RPT;            // threads per column
  constexpr int COLS = 128 / TPC;
  const int col = wt % COLS, part = wt / COLS;
  const int jrow0 = part * RPT;
  const int row0 = row_base + jrow0;
  const int n = col < tile_rows ? n0 + col : g.N;   // threads beyond the tile's rows idle (n >= N)
  EpiOperands<RPT> eo;
  epi_prefetch<RPT>(g.epi, g.M, g.N, row0, n, lean, eo);
  ptx::mbar_wait(red_full, full_phase);
  full_phase ^= 1;
  reduce_epilogue<RPT>(g.epi, g.M, g.N, red, col, jrow0, row0, n, src_mask, lean, eo);
}

// ------------------------------------------------------------------------------------------------ row ops
__device__ __forceinline__ float worker_sum(float v, float* red, int lane, int w4) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[w4] = v;
  worker_bar();
  const float t = (red[0] + red[1]) + (red[2] + red[3]);
  worker_bar();
  return t;
}

__device__ __forceinline__ void store_planes4(const float4& v, uint16_t* hi_p, uint16_t* lo_p, int bf16) {
  const uint2 h = pack_hi4(v, bf16);
  *reinterpret_cast<uint2*>(hi_p) = h;
  if (lo_p) *reinterpret_cast<uint2*>(lo_p) = pack_lo4(v, h);
}

// LayerNorm of one row by the CTA's 128 worker threads (same arithmetic as layernorm_block_kernel).
template <int NV4>
__device__ __forceinline__ void ln_row(const LnArgs& a, size_t in_row, size_t out_row, int wt, float* red) {
  const int lane = wt & 31, w4 = wt >> 5;
  const float inv_d = 1.0f / static_cast<float>(a.d);
  float4 v[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = (i * 128 + wt) * 4;
    v[i] = c < a.d ? __ldcg(reinterpret_cast<const float4*>(a.x + in_row * a.ldx + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float* ws[2] = {a.w1, a.w2};
  const float* bs[2] = {a.b1, a.b2};
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && !a.w2) break;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = worker_sum(s, red, lane, w4) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      if ((i * 128 + wt) * 4 < a.d) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rstd = rsqrtf(worker_sum(q, red, lane, w4) * inv_d + a.eps);
    if (pass == 0 && a.stats && wt == 0) a.stats[out_row] = make_float2(mean, rstd);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = (i * 128 + wt) * 4;
      if (c < a.d) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(ws[pass] + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bs[pass] + c));
        v[i].x = (v[i].x - mean) * rstd * ww.x + bb.x;
        v[i].y = (v[i].y - mean) * rstd * ww.y + bb.y;
        v[i].z = (v[i].z - mean) * rstd * ww.z + bb.z;
        v[i].w = (v[i].w - mean) * rstd * ww.w + bb.w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = (i * 128 + wt) * 4;
    if (c >= a.d) continue;
    if (a.out32) *reinterpret_cast<float4*>(a.out32 + out_row * a.ld32 + c) = v[i];
    if (a.out_hi) store_planes4(v[i], a.out_hi + out_row * a.ld16 + c, a.out_lo ? a.out_lo + out_row * a.ld16 + c : nullptr, a.bf16);
  }
}

__device__ __forceinline__ void run_ln(const LnArgs a, int wt, float* red) {
  const int keep = a.rows_per_clip - a.first_token;
  const int nrows = (a.rows / a.rows_per_clip) * keep;
  const int nv4 = (a.d + 511) / 512;
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int clip = r / keep, tok = a.first_token + (r - clip * keep);
    const size_t in_row = static_cast<size_t>(clip) * a.rows_per_clip + tok;
    const size_t out_row = a.compact ? static_cast<size_t>(r) : in_row;
    switch (nv4) {
      case 1: ln_row<1>(a, in_row, out_row, wt, red); break;
      case 2: ln_row<2>(a, in_row, out_row, wt, red); break;
      case 3: ln_row<3>(a, in_row, out_row, wt, red); break;
      case 4: ln_row<4>(a, in_row, out_row, wt, red); break;
      case 5: case 6: ln_row<6>(a, in_row, out_row, wt, red); break;
      default: ln_row<8>(a, in_row, out_row, wt, red); break;
    }
  }
}

// Attention of one (clip, head, query row) by one warp: online softmax over the keys, K/V rows of the next key in
// flight while the current one is reduced.  Lane l holds elements (c*32 + l)*VEC .. +VEC of the head.
template <int VEC, int NCH, bool IN16>
struct AttnFrag {
  static constexpr int EPL = VEC * NCH;
  static __device__ __forceinline__ void load(float (&f)[EPL], const void* row, int hd, int lane, int bf16) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int e0 = (c * 32 + lane) * VEC;
      if constexpr (IN16) {
        const uint16_t* p = static_cast<const uint16_t*>(row) + e0;
        uint16_t h[VEC];
        if (e0 < hd) {
          if constexpr (VEC == 8) {
            const uint4 t = __ldcg(reinterpret_cast<const uint4*>(p));
            h[0] = t.x & 0xffff; h[1] = t.x >> 16; h[2] = t.y & 0xffff; h[3] = t.y >> 16;
            h[4] = t.z & 0xffff; h[5] = t.z >> 16; h[6] = t.w & 0xffff; h[7] = t.w >> 16;
          } else if constexpr (VEC == 4) {
            const uint2 t = __ldcg(reinterpret_cast<const uint2*>(p));
            h[0] = t.x & 0xffff; h[1] = t.x >> 16; h[2] = t.y & 0xffff; h[3] = t.y >> 16;
          } else if constexpr (VEC == 2) {
            const uint32_t t = __ldcg(reinterpret_cast<const uint32_t*>(p));
            h[0] = t & 0xffff; h[1] = t >> 16;
          } else {
            h[0] = __ldcg(p);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            f[c * VEC + v] = bf16 ? __bfloat162float(__ushort_as_bfloat16(h[v])) : __half2float(__ushort_as_half(h[v]));
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) f[c * VEC + v] = 0.f;
        }
      } else {
        const float* p = static_cast<const float*>(row) + e0;
        if constexpr (VEC == 4) {
          const float4 t = e0 < hd ? __ldcg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
          f[c * 4] = t.x; f[c * 4 + 1] = t.y; f[c * 4 + 2] = t.z; f[c * 4 + 3] = t.w;
        } else if constexpr (VEC == 2) {
          const float2 t = e0 < hd ? __ldcg(reinterpret_cast<const float2*>(p)) : make_float2(0.f, 0.f);
          f[c * 2] = t.x; f[c * 2 + 1] = t.y;
        } else {
          f[c] = e0 < hd ? __ldcg(p) : 0.f;
        }
      }
    }
  }
};

template <int VEC, int NCH, bool IN16>
__device__ __forceinline__ void attn_task(const AttnArgs& a, int b, int h, int i, int lane) {
  using F = AttnFrag<VEC, NCH, IN16>;
  constexpr int EPL = F::EPL;
  constexpr int ES = IN16 ? 2 : 4;
  constexpr float kLog2e = 1.4426950408889634f;
  const int hd = a.hd;
  const char* qb = reinterpret_cast<const char*>(a.q) + (static_cast<size_t>(b) * a.q_clip_stride + static_cast<size_t>(i) * a.ldq + h * hd) * ES;
  const char* kb = reinterpret_cast<const char*>(a.k) + (static_cast<size_t>(b) * a.kv_clip_stride + h * hd) * ES;
  const char* vb = reinterpret_cast<const char*>(a.v) + (static_cast<size_t>(b) * a.kv_clip_stride + h * hd) * ES;
  const size_t kv_pitch = static_cast<size_t>(a.ldkv) * ES;
  // keys visible to this query row
  int nk = a.Sk;
  if (a.mask_kind == 1) { const int lim = i + (a.Sk - a.Sq) + 1; nk = lim < nk ? lim : nk; }
  float ql[EPL], kc[EPL], vc[EPL], kn[EPL], vn[EPL], o[EPL];
  F::load(ql, qb, hd, lane, a.bf16);
#pragma unroll
  for (int t = 0; t < EPL; ++t) { o[t] = 0.f; kn[t] = 0.f; vn[t] = 0.f; }
  if (nk > 0) { F::load(kn, kb, hd, lane, a.bf16); F::load(vn, vb, hd, lane, a.bf16); }
  float mx = -INFINITY, sum = 0.f;
  for (int j = 0; j < nk; ++j) {
#pragma unroll
    for (int t = 0; t < EPL; ++t) { kc[t] = kn[t]; vc[t] = vn[t]; }
    if (j + 1 < nk) {
      F::load(kn, kb + (j + 1) * kv_pitch, hd, lane, a.bf16);
      F::load(vn, vb + (j + 1) * kv_pitch, hd, lane, a.bf16);
    }
    float part = 0.f;
#pragma unroll
    for (int t = 0; t < EPL; ++t) part = fmaf(ql[t], kc[t], part);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    float s = part * a.scale;
    if (a.mask_kind == 2) s += __ldg(a.mask + i * a.Sk + j);
    if (s == -INFINITY) continue;
    const float mnew = fmaxf(mx, s);
    const float corr = exp2f((mx - mnew) * kLog2e);   // 0 on the first visible key (mx = -inf)
    const float p = exp2f((s - mnew) * kLog2e);
    sum = sum * corr + p;
#pragma unroll
    for (int t = 0; t < EPL; ++t) o[t] = fmaf(p, vc[t], o[t] * corr);
    mx = mnew;
  }
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  const size_t row = a.out_compact ? static_cast<size_t>(b) * (a.Sq - a.q_first) + (i - a.q_first) : static_cast<size_t>(b) * a.Sq + i;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int e0 = (c * 32 + lane) * VEC;
    if (e0 >= hd) continue;
    const int col = h * hd + e0;
    if constexpr (VEC % 4 == 0) {
#pragma unroll
      for (int v = 0; v < VEC; v += 4) {
        const float4 f4 = make_float4(o[c * VEC + v] * inv, o[c * VEC + v + 1] * inv, o[c * VEC + v + 2] * inv, o[c * VEC + v + 3] * inv);
        if (a.out32) *reinterpret_cast<float4*>(a.out32 + row * a.ld32 + col + v) = f4;
        if (a.out_hi) store_planes4(f4, a.out_hi + row * a.ld16 + col + v, a.out_lo ? a.out_lo + row * a.ld16 + col + v : nullptr, a.bf16);
      }
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float val = o[c * VEC + v] * inv;
        if (a.out32) a.out32[row * a.ld32 + col + v] = val;
        if (a.out_hi) {
          const uint16_t hi = to_plane_hi(val, a.bf16);
          a.out_hi[row * a.ld16 + col + v] = hi;
          if (a.out_lo) a.out_lo[row * a.ld16 + col + v] = to_plane_lo(val, hi);
        }
      }
    }
  }
}

template <bool IN16>
__device__ __forceinline__ void run_attn(const AttnArgs a, int wt) {
  const int lane = wt & 31;
  const int gw = blockIdx.x * 4 + (wt >> 5), nw = gridDim.x * 4;
  const int nq = a.Sq - a.q_first;
  const int tasks = a.clips * a.heads * nq;
  // vector width of the lane fragment: the widest one the head size and every pitch allow
  constexpr int ES = IN16 ? 2 : 4;
  const long long al = (a.ldq | a.ldkv | a.q_clip_stride | a.kv_clip_stride | a.hd) * ES |
                       (reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v));
  // (vector paths also store 4 outputs at a time: 8-byte plane stores, 16-byte fp32 stores)
  const bool out_ok = a.ld16 % 4 == 0 && a.ld32 % 4 == 0 && reinterpret_cast<uintptr_t>(a.out_hi) % 8 == 0 &&
                      reinterpret_cast<uintptr_t>(a.out_lo) % 8 == 0 && reinterpret_cast<uintptr_t>(a.out32) % 16 == 0;
  const bool v16 = (al % 16) == 0 && out_ok, v8 = (al % 8) == 0 && out_ok;
  for (int t = gw; t < tasks; t += nw) {
    const int b = t / (a.heads * nq), r = t - b * a.heads * nq;
    const int h = r / nq, i = a.q_first + (r - h * nq);
    if constexpr (IN16) {
      if (v16 && a.hd == 256) attn_task<8, 1, true>(a, b, h, i, lane);
      else if (v8 && a.hd <= 128 && a.hd % 4 == 0) attn_task<4, 1, true>(a, b, h, i, lane);
      else if (v8 && a.hd % 4 == 0) attn_task<4, 2, true>(a, b, h, i, lane);
      else attn_task<1, 8, true>(a, b, h, i, lane);
    } else {
      if (v16 && a.hd <= 128 && a.hd % 4 == 0) attn_task<4, 1, false>(a, b, h, i, lane);
      else if (v16 && a.hd % 4 == 0) attn_task<4, 2, false>(a, b, h, i, lane);
      else attn_task<1, 8, false>(a, b, h, i, lane);
    }
  }
}

__device__ __forceinline__ void run_pack(const PackArgs a, int wt) {
  const int w4 = a.width >> 2;
  const long long total = static_cast<long long>(a.clips) * a.tokens * w4;
  for (long long idx = blockIdx.x * 128LL + wt; idx < total; idx += static_cast<long long>(gridDim.x) * 128) {
    const int c4 = static_cast<int>(idx % w4);
    const long long rt = idx / w4;
    const int t = static_cast<int>(rt % a.tokens);
    const long long b = rt / a.tokens;
    float4 v;
    const int s = a.slot[t];
    if (s < 0) v = make_float4(a.fill, a.fill, a.fill, a.fill);
    else {
      v = __ldcg(reinterpret_cast<const float4*>(a.src + b * a.src_clip_stride + s * a.src_slot_stride) + c4);
      v.x *= a.scale; v.y *= a.scale; v.z *= a.scale; v.w *= a.scale;
    }
    if (a.out32) *(reinterpret_cast<float4*>(a.out32 + b * a.out_clip_stride + t * a.out_tok_stride) + c4) = v;
    if (a.out_hi) {
      const size_t o = static_cast<size_t>(rt) * a.ld16 + c4 * 4;
      store_planes4(v, a.out_hi + o, a.out_lo ? a.out_lo + o : nullptr, a.bf16);
    }
  }
}

__device__ __forceinline__ void run_add(const AddArgs a, int wt) {
  const int w4 = a.width >> 2;
  const long long total = static_cast<long long>(a.clips) * w4;
  for (long long idx = blockIdx.x * 128LL + wt; idx < total; idx += static_cast<long long>(gridDim.x) * 128) {
    const int c4 = static_cast<int>(idx % w4);
    const long long b = idx / w4;
    float4* d = reinterpret_cast<float4*>(a.dst + b * a.dst_clip_stride) + c4;
    float4 v = __ldcg(d);
    const float4 s = a.src ? __ldcg(reinterpret_cast<const float4*>(a.src + b * a.src_clip_stride) + c4)
                           : make_float4(a.fill, a.fill, a.fill, a.fill);
    v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    *d = v;
  }
}

}  // namespace pkx

__global__ void __cluster_dims__(kPkCluster, 1, 1) __launch_bounds__(kPkThreads, 1)
persistent_kernel(const __grid_constant__ PkParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_ring = smem;
  uint8_t* a_ring = w_ring + static_cast<size_t>(P.w_slots) * kPkWSlotBytes;
  float* red = reinterpret_cast<float*>(a_ring + static_cast<size_t>(P.a_stages) * P.a_stage_bytes);   // [4 sources][128][16]
  float* stage = red + kPkCluster * kPkTileN * kPkRpo;                                                  // [4 owners][128][16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red) + 2 * kPkRedBytes);
  uint64_t* w_full = bars;                       // [w_slots]   TMA -> MMA
  uint64_t* w_empty = w_full + P.w_slots;        // [w_slots]   MMA -> TMA
  uint64_t* a_full = w_empty + P.w_slots;        // [a_stages]  workers (128 arrivals) -> MMA
  uint64_t* a_empty = a_full + P.a_stages;       // [a_stages]  MMA -> workers
  uint64_t* t_full = a_empty + P.a_stages;       // [2] MMA -> workers
  uint64_t* t_empty = t_full + 2;                // [2] workers -> MMA
  uint64_t* red_full = t_empty + 2;              // partial tiles of every source CTA have landed (tx bytes)
  uint64_t* red_empty = red_full + 1;            // every owner of the cluster has consumed the previous tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(red_empty + 1);
  float* ln_red = reinterpret_cast<float*>(tmem_slot + 2);   // [4]
  PkOp* op_s = reinterpret_cast<PkOp*>(reinterpret_cast<uint8_t*>(bars) + 1024);   // the workers' copy of the current op

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0);
  const int cluster = static_cast<int>(__shfl_sync(0xffffffffu, pkx::cluster_id_x(), 0));
  const int n_clusters = static_cast<int>(__shfl_sync(0xffffffffu, pkx::ncluster_id_x(), 0));
  const unsigned n_cta = gridDim.x;
  unsigned int* ctr = P.sync;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.w_slots; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < P.a_stages; ++s) { ptx::mbar_init(&a_full[s], 128); ptx::mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&t_full[b], 1); ptx::mbar_init(&t_empty[b], 4); }
    ptx::mbar_init(red_full, 1);
    ptx::mbar_init(red_empty, kPkCluster);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();
  ptx::cluster_sync_all();   // peer barriers exist before anyone signals them

  if (warp == 0) {
    // ------------------------------------------------------------------ weight stream (runs ahead of the barriers)
    int slot = 0;
    uint32_t phase = 0;
    for (int oi = 0; oi < P.n_ops; ++oi) {
      const PkOp* op = P.ops + oi;
      if (pkx::uni(op->type) != PK_GEMM) continue;
      const PkGemm& g = op->u.g;
      const int K = pkx::uni(g.K), T = pkx::uni(g.T), n_tiles = pkx::uni(g.n_tiles), split = pkx::uni(g.split);
      const int nk = (K + kTcBK - 1) / kTcBK;
      int kb0, kb1;
      pkx::k_range(nk, static_cast<int>(rank), kb0, kb1);
      if (kb0 == kb1) continue;
      const CUtensorMap* mh = static_cast<const CUtensorMap*>(pkx::uni_ptr(P.maps + pkx::uni(g.w_hi)));
      const CUtensorMap* ml = static_cast<const CUtensorMap*>(pkx::uni_ptr(P.maps + pkx::uni(g.w_lo)));
      const int cps = kPkTileN / T;                      // K blocks per ring slot
      const uint32_t blk_bytes = static_cast<uint32_t>(T) * kTcBK * 2;
      const int planes = split ? 2 : 1;
      if (ptx::elect_one()) { ptx::prefetch_tensormap(mh); if (split) ptx::prefetch_tensormap(ml); }
      for (int t = cluster; t < n_tiles; t += n_clusters) {
        for (int kb = kb0; kb < kb1; kb += cps) {
          const int nb = kb1 - kb < cps ? kb1 - kb : cps;
          for (int p = 0; p < planes; ++p) {
            ptx::mbar_wait(&w_empty[slot], phase ^ 1);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&w_full[slot], blk_bytes * nb);
              uint8_t* dst = w_ring + static_cast<size_t>(slot) * kPkWSlotBytes;
              for (int i = 0; i < nb; ++i)
                ptx::tma_load_2d(dst + i * blk_bytes, p ? ml : mh, &w_full[slot], (kb + i) * kTcBK, t * T);
            }
            __syncwarp();
            if (++slot == P.w_slots) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer
    int ws = 0, as = 0, buf = 0;
    uint32_t wphase = 0, aphase = 0, tph = 0;   // tph: bit b = phase of accumulator buffer b
    for (int oi = 0; oi < P.n_ops; ++oi) {
      const PkOp* op = P.ops + oi;
      if (pkx::uni(op->type) != PK_GEMM) continue;
      const PkGemm& g = op->u.g;
      const int K = pkx::uni(g.K), T = pkx::uni(g.T), n_tiles = pkx::uni(g.n_tiles), MT = pkx::uni(g.MT);
      const bool split = pkx::uni(g.split) != 0;
      const bool bf = pkx::uni(g.bf16) != 0;
      const int nk = (K + kTcBK - 1) / kTcBK;
      int kb0, kb1;
      pkx::k_range(nk, static_cast<int>(rank), kb0, kb1);
      if (kb0 == kb1) continue;
      const uint32_t idesc = ptx::make_idesc_f16(kPkTileN, MT, bf);
      const int cps = kPkTileN / T;
      const uint32_t blk_bytes = static_cast<uint32_t>(T) * kTcBK * 2;
      const int a_need = (kb1 - kb0) * (split ? 2 : 1);     // activation stages one tile consumes
      const bool a_all = a_need <= P.a_stages;
      // TMEM: [hi*hi even | hi*hi odd | cross even | cross odd], mtp columns each; two such buffers when MT <= 64
      const uint32_t mtp = MT <= 64 ? 64u : 128u;
      const int nbuf = MT <= 64 ? 2 : 1;
      if (buf >= nbuf) buf = 0;
      const uint32_t w_ring_lo = pkx::desc_lo(ptx::smem_u32(w_ring)), a_ring_lo = pkx::desc_lo(ptx::smem_u32(a_ring));
      const uint32_t a_step = static_cast<uint32_t>(P.a_stage_bytes) >> 4, blk_step = blk_bytes >> 4;
      for (int t = cluster; t < n_tiles; t += n_clusters) {
        ptx::mbar_wait(&t_empty[buf], ((tph >> buf) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + buf * (4 * mtp);
        const uint32_t d1 = d0 + 2 * mtp;
        // the K slab of the activations arrives in one L2 round trip: when the ring holds all of it, wait for its LAST
        // stage only (a thread's arrival for a stage fires after all its earlier copies), then ONE generic->async proxy
        // fence instead of one per block
        if (a_all) {
          int s = as + a_need - 1; uint32_t ph = aphase;
          if (s >= P.a_stages) { s -= P.a_stages; ph ^= 1; }
          ptx::mbar_wait(&a_full[s], ph);
          ptx::fence_proxy_async_smem();
          ptx::tc_fence_after();
        }
        if (a_all) {
          // The whole K slab of the activations is in the ring: the weight slots of up to kPkChunk slot groups are awaited
          // by different lanes at once (a try_wait on a completed barrier is ~90 cycles), then ONE elected thread issues
          // every MMA and commit of the chunk back to back.  tools/mma_rate.cu: a 128 x (<= 64) x 16 MMA issues every 48
          // cycles whatever its width or accumulator, a commit costs 45; the per-block waits, fences and elections of the
          // first version made a 32-MMA tile 4500 cycles instead of ~2100.
          const int planes = split ? 2 : 1;
          int kb = kb0;
          bool first_chunk = true;
          while (kb < kb1) {
            const int groups_left = (kb1 - kb + cps - 1) / cps;
            int ng = groups_left < kPkChunk ? groups_left : kPkChunk;
            if (ng * planes > P.w_slots) ng = P.w_slots / planes;
            const int nsl = ng * planes;
            if (lane < nsl) {
              int j = ws + lane; uint32_t ph = wphase;
              if (j >= P.w_slots) { j -= P.w_slots; ph ^= 1; }
              ptx::mbar_wait(&w_full[j], ph);
            }
            __syncwarp();
            ptx::tc_fence_after();
            if (lane == 0 && t == cluster && first_chunk) PK_TRACE(16);
            const int nblk = kb1 - kb < ng * cps ? kb1 - kb : ng * cps;
            if (ptx::elect_one()) {
              int lws = ws, las = as, lkb = kb;
              for (int gi = 0; gi < ng; ++gi) {
                const int nb = kb1 - lkb < cps ? kb1 - lkb : cps;
                const int lws1 = (lws + 1 == P.w_slots) ? 0 : lws + 1;
                const uint32_t w_hi = w_ring_lo + static_cast<uint32_t>(lws) * (kPkWSlotBytes >> 4);
                const uint32_t w_lo = w_ring_lo + static_cast<uint32_t>(lws1) * (kPkWSlotBytes >> 4);
                for (int i = 0; i < nb; ++i) {
                  const int las1 = (las + 1 == P.a_stages) ? 0 : las + 1;
                  const uint32_t a_hi = a_ring_lo + static_cast<uint32_t>(las) * a_step;
                  const uint32_t a_lo = a_ring_lo + static_cast<uint32_t>(las1) * a_step;
                  const bool first = (lkb + i == kb0), last = (lkb + i == kb1 - 1);
                  if (split) pkx::mma_block<true>(d0, d1, mtp, w_hi + i * blk_step, w_lo + i * blk_step, a_hi, a_lo, idesc, first);
                  else pkx::mma_block<false>(d0, d1, mtp, w_hi + i * blk_step, w_lo, a_hi, a_lo, idesc, first);
                  ptx::umma_commit(&a_empty[las]);
                  if (split) ptx::umma_commit(&a_empty[las1]);
                  if (i == nb - 1) { ptx::umma_commit(&w_empty[lws]); if (split) ptx::umma_commit(&w_empty[lws1]); }
                  if (last) ptx::umma_commit(&t_full[buf]);
                  las = split ? ((las1 + 1 == P.a_stages) ? 0 : las1 + 1) : las1;
                }
                lws = split ? ((lws1 + 1 == P.w_slots) ? 0 : lws1 + 1) : lws1;
                lkb += nb;
              }
            }
            __syncwarp();
            if (lane == 0 && t == cluster && first_chunk) PK_TRACE(17);
            first_chunk = false;
            as += nblk * planes;
            if (as >= P.a_stages) { as -= P.a_stages; aphase ^= 1; }
            ws += nsl;
            if (ws >= P.w_slots) { ws -= P.w_slots; wphase ^= 1; }
            kb += nblk;
          }
        } else {
          for (int kb = kb0; kb < kb1; kb += cps) {
            const int nb = kb1 - kb < cps ? kb1 - kb : cps;
            const int ws1 = (ws + 1 == P.w_slots) ? 0 : ws + 1;
            ptx::mbar_wait(&w_full[ws], wphase);
            if (split) ptx::mbar_wait(&w_full[ws1], ws1 == 0 ? wphase ^ 1 : wphase);
            ptx::tc_fence_after();
            const uint32_t w_hi = w_ring_lo + static_cast<uint32_t>(ws) * (kPkWSlotBytes >> 4);
            const uint32_t w_lo = w_ring_lo + static_cast<uint32_t>(ws1) * (kPkWSlotBytes >> 4);
            for (int i = 0; i < nb; ++i) {
              const int as1 = (as + 1 == P.a_stages) ? 0 : as + 1;
              if (!a_all) {
                ptx::mbar_wait(&a_full[as], aphase);
                if (split) ptx::mbar_wait(&a_full[as1], as1 == 0 ? aphase ^ 1 : aphase);
                ptx::fence_proxy_async_smem();   // the activation stage was written with generic-proxy copies
                ptx::tc_fence_after();
              }
              if (lane == 0 && t == cluster && kb + i - kb0 < 4) PK_TRACE(16 + kb + i - kb0);
              const uint32_t a_hi = a_ring_lo + static_cast<uint32_t>(as) * a_step;
              const uint32_t a_lo = a_ring_lo + static_cast<uint32_t>(as1) * a_step;
              const bool first = (kb + i == kb0), last = (kb + i == kb1 - 1);
              if (ptx::elect_one()) {
                if (split) pkx::mma_block<true>(d0, d1, mtp, w_hi + i * blk_step, w_lo + i * blk_step, a_hi, a_lo, idesc, first);
                else pkx::mma_block<false>(d0, d1, mtp, w_hi + i * blk_step, w_lo, a_hi, a_lo, idesc, first);
                ptx::umma_commit(&a_empty[as]);
                if (split) ptx::umma_commit(&a_empty[as1]);
                if (i == nb - 1) { ptx::umma_commit(&w_empty[ws]); if (split) ptx::umma_commit(&w_empty[ws1]); }
                if (last) ptx::umma_commit(&t_full[buf]);
              }
              __syncwarp();
              if (split) { if (++as == P.a_stages) { as = 0; aphase ^= 1; } }
              if (++as == P.a_stages) { as = 0; aphase ^= 1; }
            }
            if (split) { if (++ws == P.w_slots) { ws = 0; wphase ^= 1; } }
            if (++ws == P.w_slots) { ws = 0; wphase ^= 1; }
          }
        }
        tph ^= 1u << buf;
        if (++buf >= nbuf) buf = 0;
      }
      if (lane == 0) PK_TRACE(4);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ workers: activation loads, reduction + epilogue, row ops
    const int wt = static_cast<int>(threadIdx.x) - 128;   // 0..127 = TMEM lane = row of the weight tile
    const int q = warp & 3;
    int buf = 0, as = 0;
    uint32_t tph = 0, aphase = 0;
    uint32_t full_phase = 0, empty_it = 0;
    const uint32_t a_ring_addr = ptx::smem_u32(a_ring), a_full_addr = ptx::smem_u32(a_full);
    const uint32_t red_local = ptx::smem_u32(red);
    const uint32_t stage_local = ptx::smem_u32(stage);
    const uint32_t red_full_local = ptx::smem_u32(red_full);
    const uint32_t red_empty_local = ptx::smem_u32(red_empty);
    for (int oi = 0; oi < P.n_ops; ++oi) {
      // the op descriptor goes to shared memory once
      {
        const uint32_t* gsrc = reinterpret_cast<const uint32_t*>(P.ops + oi);
        uint32_t* sdst = reinterpret_cast<uint32_t*>(op_s);
        for (int i = wt; i < static_cast<int>(sizeof(PkOp) / 4); i += 128) sdst[i] = __ldg(gsrc + i);
      }
      const PkOp* op = op_s;
      pkx::worker_bar();
      // everything this op reads from global memory was written before the previous op's barrier; the op's own set-up
      // (descriptor decode, address tables) runs before the wait, while the slowest CTA is still arriving
      auto wait_previous_op = [&]() {
        if (wt == 0) { PK_TRACE(5); if (P.trace && static_cast<int>(blockIdx.x) == P.trace_cta) P.trace[static_cast<size_t>(oi) * 32 + 14] = clock64(); pkx::grid_wait(ctr, static_cast<unsigned>(oi) * n_cta); PK_TRACE(6); }
        pkx::worker_bar();
      };
      const int type = op->type;
      if (type != PK_GEMM) wait_previous_op();
      if (type == PK_GEMM) {
        const PkGemm& g = op->u.g;
        const int M = g.M, N = g.N, T = g.T, MT = g.MT, n_tiles = g.n_tiles, lda = g.lda;
        const bool split = g.split != 0;
        const uint16_t* a_hi = g.a_hi;
        const uint16_t* a_lo = g.a_lo;
        const int nk = (g.K + kTcBK - 1) / kTcBK;
        int kb0, kb1;
        pkx::k_range(nk, static_cast<int>(rank), kb0, kb1);
        const bool have_k = kb0 < kb1;
        uint32_t src_mask = 0;
        for (int s = 0; s < kPkCluster; ++s) {
          int a0, a1;
          pkx::k_range(nk, s, a0, a1);
          if (a0 < a1) src_mask |= 1u << s;
        }
        const int a_need = (kb1 - kb0) * (split ? 2 : 1);
        const bool a_all = a_need <= P.a_stages;
        const uint32_t mtp = MT <= 64 ? 64u : 128u;
        const int nbuf = MT <= 64 ? 2 : 1;
        if (have_k && buf >= nbuf) buf = 0;     // (the MMA warp does the same, and only for ops it takes part in)
        pkx::APieces ap;
        pkx::a_pieces_init(ap, a_hi, lda, M, wt);   // rows M..MT-1 of a stage keep whatever they held: token columns nobody reads
        const long long lo_off = a_lo - a_hi;   // element offset of the lo plane (same layout)
        const Epilogue& e = g.epi;
        const bool lean = e.row_map == 0 && e.res_clip_rows == 0 && !e.pe && !e.gate && !e.drop_thr && !e.alpha_dev;
        const int groups = (M + kPkGroup - 1) / kPkGroup;
        const int rpt_sel = T <= 32 ? 4 : (T <= 64 ? 8 : 16);
        wait_previous_op();
        for (int t = cluster; t < n_tiles; t += n_clusters) {
          if (have_k) {
            // this CTA's K slab of the activations -> swizzled operand stages, every block in flight at once.  Stages are
            // released in order, so when the slab fits the ring only its last stage's release has to be observed.
            if (wt == 0 && t == cluster) PK_TRACE(21);
            if (a_all) {
              int sl = as + a_need - 1; uint32_t ph = aphase;
              if (sl >= P.a_stages) { sl -= P.a_stages; ph ^= 1; }
              ptx::mbar_wait(&a_empty[sl], ph ^ 1);
            }
            if (wt == 0 && t == cluster) PK_TRACE(22);
            if (a_all) {
              if (ap.n <= 1) pkx::issue_slab<1>(ap, lo_off, split, kb0, kb1, as, P.a_stages, a_ring_addr, P.a_stage_bytes);
              else if (ap.n <= 3) pkx::issue_slab<3>(ap, lo_off, split, kb0, kb1, as, P.a_stages, a_ring_addr, P.a_stage_bytes);
              else pkx::issue_slab<8>(ap, lo_off, split, kb0, kb1, as, P.a_stages, a_ring_addr, P.a_stage_bytes);
              // each thread waits for its own copies, then ONE lane per warp arrives for its 32 threads on every stage
              // of the slab.  (cp.async.mbarrier.arrive.noinc per thread and stage cost ~100 ns apiece: 0.8 us of the
              // 1.15 us slab phase at 8 stages, 1.6 us at 16 - whatever the number of bytes copied.)
              if (wt == 0 && t == cluster) PK_TRACE(23);
              asm volatile("cp.async.wait_all;" ::: "memory");
              __syncwarp();
              if (wt == 0 && t == cluster) PK_TRACE(24);
              for (int i = 0; i < a_need; ++i) {
                if (lane == 0) pkx::mbar_arrive_count(a_full_addr + as * 8, 32u);
                if (++as == P.a_stages) { as = 0; aphase ^= 1; }
              }
            } else {
              for (int kb = kb0; kb < kb1; ++kb) {
                for (int p = 0; p < (split ? 2 : 1); ++p) {
                  ptx::mbar_wait(&a_empty[as], aphase ^ 1);
                  pkx::load_a_block<8>(ap, p ? lo_off : 0, kb, a_ring_addr + static_cast<uint32_t>(as) * P.a_stage_bytes);
                  pkx::a_block_arrive(a_full_addr + as * 8);
                  if (++as == P.a_stages) { as = 0; aphase ^= 1; }
                }
              }
            }
            if (wt == 0 && t == cluster) PK_TRACE(20);
            ptx::mbar_wait(&t_full[buf], (tph >> buf) & 1u);
            ptx::tc_fence_after();
            if (wt == 0 && t == cluster) PK_TRACE(7);
          }
          const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * (4 * mtp);
          const int tile_rows = N - t * T < T ? N - t * T : T;
          for (int tg = 0; tg < groups; ++tg) {
            const int tok0 = tg * kPkGroup;
            const int rows_g = M - tok0 < kPkGroup ? M - tok0 : kPkGroup;
            const int owners = (rows_g + kPkRpo - 1) / kPkRpo;       // CTAs of the cluster that own token rows of this group
            const bool owner = static_cast<int>(rank) < owners;
            if (owner && wt == 0)
              ptx::mbar_arrive_expect_tx(red_full, static_cast<uint32_t>(__popc(src_mask)) * T * kPkRpo * 4);
            if (have_k) {
              // all owners are done with the previous partials (so this CTA's previous bulk copies have landed too)
              ptx::mbar_wait(red_empty, (empty_it & 1u) ^ 1u);
              pkx::stage_partial(tacc, mtp, split, wt, tok0, owners, stage_local);
              if (tg == groups - 1) {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&t_empty[buf]);
              }
              ptx::fence_proxy_async_smem();   // staged with generic stores, read by the bulk-copy engine
              pkx::worker_bar();
              if (wt < owners)
                pkx::bulk_copy_to_peer(ptx::mapa_u32(red_local + rank * (kPkTileN * kPkRpo * 4), static_cast<uint32_t>(wt)),
                                       stage_local + wt * (kPkTileN * kPkRpo * 4), static_cast<uint32_t>(T) * kPkRpo * 4,
                                       ptx::mapa_u32(red_full_local, static_cast<uint32_t>(wt)));
              if (wt == 0 && t == cluster && tg == 0) PK_TRACE(8);
            }
            if (owner) {
              const int row_base = tok0 + static_cast<int>(rank) * kPkRpo;
              if (rpt_sel == 16) pkx::owner_tile<16>(g, red, wt, row_base, t * T, tile_rows, src_mask, red_full, full_phase, lean);
              else if (rpt_sel == 8) pkx::owner_tile<8>(g, red, wt, row_base, t * T, tile_rows, src_mask, red_full, full_phase, lean);
              else pkx::owner_tile<4>(g, red, wt, row_base, t * T, tile_rows, src_mask, red_full, full_phase, lean);
              if (wt == 0 && t == cluster && tg == 0) PK_TRACE(9);
            }
            pkx::worker_bar();   // all reads of the partial buffer are done
            if (wt < kPkCluster) ptx::mbar_arrive_cluster(ptx::mapa_u32(red_empty_local, static_cast<uint32_t>(wt)));
            ++empty_it;
          }
          if (have_k) { tph ^= 1u << buf; if (++buf >= nbuf) buf = 0; }
        }
      } else if (type == PK_LN) {
        pkx::run_ln(op->u.ln, wt, ln_red);
      } else if (type == PK_ATTN) {
        if (op->in16) pkx::run_attn<true>(op->u.at, wt);
        else pkx::run_attn<false>(op->u.at, wt);
      } else if (type == PK_PACK) {
        pkx::run_pack(op->u.pk, wt);
      } else if (type == PK_ADD) {
        pkx::run_add(op->u.ad, wt);
      }
      // publish: bar.sync orders the 128 workers' writes before thread 0's cumulative gpu-scope release (like grid.sync())
      if (wt == 0) PK_TRACE(11);
      pkx::worker_bar();
      if (wt == 0) {
        pkx::fence_acq_rel_gpu();
        PK_TRACE(12);
        if (P.trace && static_cast<int>(blockIdx.x) == P.trace_cta) P.trace[static_cast<size_t>(oi) * 32 + 15] = clock64();
        pkx::red_relaxed_gpu_add(ctr, 1u);
        PK_TRACE(13);
      }
    }
    // last one out re-arms the counter for the next launch (nobody polls it any more once all have arrived)
    if (blockIdx.x == 0 && wt == 0) {
      pkx::grid_wait(ctr, static_cast<unsigned>(P.n_ops) * n_cta);
      *reinterpret_cast<volatile unsigned*>(ctr) = 0u;
      __threadfence();
    }
  }

  __syncthreads();
  ptx::cluster_sync_all();   // nobody leaves while a peer may still copy into this CTA
  if (warp == 2) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// shared memory plan of a launch whose widest GEMM has MT token rows
struct PkSmemPlan { int w_slots, a_stages, a_stage_bytes, total; };
inline PkSmemPlan pk_smem_plan(int mt_max) {
  PkSmemPlan p{};
  if (mt_max < 16) mt_max = 16;
  p.a_stage_bytes = mt_max * kTcBK * 2;
  int a_bytes = 48 * 1024;                           // one tile's K slab of the activations at MT = 48 (8 blocks)
  p.a_stages = a_bytes / p.a_stage_bytes;
  if (p.a_stages < 4) p.a_stages = 4;                // split precision consumes two stages per MMA block
  if (p.a_stages > 16) p.a_stages = 16;
  const int fixed = 1024 /*align*/ + p.a_stages * p.a_stage_bytes + 2 * kPkRedBytes + 2048 /*barriers + op copy*/;
  p.w_slots = (kTcSmemLimit - fixed) / kPkWSlotBytes;
  if (p.w_slots > 12) p.w_slots = 12;
  p.total = fixed + p.w_slots * kPkWSlotBytes;
  return p;
}

}  // namespace sdvg
