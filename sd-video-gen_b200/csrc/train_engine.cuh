// Host side of the training step (trainers/trainer.py:123-165): teacher-forced forward with saved activations,
// criterion, backward, Adam.  Uses the Engine's weights, tensor-core GEMM dispatch, LayerNorm and attention
// kernels; adds the transposed weight planes, the flat gradient / moment vectors (same offsets as the Engine's
// parameter arena, so one NCCL all-reduce covers all gradients) and the backward-only kernels of train.cuh.
//
// Per linear layer y = x W^T + b the backward pass is
//     pack_t(dy)  -> dy planes [M][N], dy^T planes [N][Mp], db = colsum(dy) / S
//     pack_t(x)   -> x^T planes [K][Mp]
//     dW = dy^T x      tensor-core GEMM, epilogue multiplies by 1/S, writes the fp32 gradient in place
//     dx = dy W        tensor-core GEMM against the W^T planes, epilogue adds the residual-path gradient
// S is the device-side power-of-two loss scale (train.cuh).  Dropout (model.train(), DROPOUT_P of the config) is
// applied at nn.Transformer's sites with the library's counter-based masks (common.cuh: drop_hash) - in the forward
// epilogues / the training attention kernel, and regenerated in the backward pass; torch's Philox stream cannot be
// reproduced by another implementation, so parity is tested under the same masks (the test-suite's CPU checker restates the hash) and for p = 0.
#pragma once
#include "engine.cuh"
#include "loss.cuh"
#include "train.cuh"

namespace sdvg {

struct TrainLoss {
  int P;                     // frames_to_predict: the loss covers the last P target positions (trainers/trainer.py:145)
  int use_mse, use_l1, use_gdl;
  float lambda_gdl, alpha;
  int use_nce;
  float temperature, lambda_nce;
};

// dropout mask of one site: element kept iff drop_hash(key, index) >= thr, then multiplied by scale (thr == 0: off)
struct Drop { uint32_t thr = 0, key = 0; float scale = 1.f; };

class Trainer {
 public:
  Engine& g;
  explicit Trainer(Engine& e) : g(e) {}

  struct TLinear {
    Linear fwd;            // forward view: N x K, planes [N][K]
    Linear t;              // transposed view: N' = K, K' = N, planes W^T [K][N]
    float* gw = nullptr;   // fp32 gradient [N][K] inside `grads`
    float* gb = nullptr;
  };
  struct TAttn { TLinear qkv, q, kv, out; };
  struct LnGrad { float* w = nullptr; float* b = nullptr; };
  struct TEnc { TAttn sa; TLinear ff1, ff2; LnGrad n1, n2; };
  struct TDec { TAttn sa, ca; TLinear ff1, ff2; LnGrad n1, n2, n3; };
  struct EncSave { float *qkv, *a, *y1, *x1, *h, *y2; float2 *st1, *st2; };
  struct DecSave { float *qkv, *a, *y1, *x1, *qc, *kvc, *ac, *y2, *x2, *h, *y3; float2 *st1, *st2, *st3; };
  struct WT { uint16_t* hi = nullptr; uint16_t* lo = nullptr; int ld = 0; };

  bool ready = false;
  int rows = 0;            // row capacity of every activation buffer (= Engine::max_rows, a multiple of 128)
  float* grads = nullptr;  // flat gradient vector, same layout as Engine::arena
  float* adam_m = nullptr;
  float* adam_v = nullptr;
  long long adam_t = 0;
  float* scale = nullptr;  // [0] = loss scale S, [1] = 1 / S
  unsigned int* amax = nullptr;
  float* loss_out = nullptr;  // [5]
  size_t decoder_offset = 0;  // first element of the decoder-side bucket (decoder layers, decoder.norm, out)

  std::vector<WT> wt;      // per weight slot
  TLinear t_emb, t_out;
  std::vector<TEnc> tenc;
  std::vector<TDec> tdec;
  LnGrad g_encnorm, g_decnorm;

  std::vector<float*> xe, xd;   // layer inputs: xe[l] input of encoder layer l, xe[Le] input of encoder.norm
  std::vector<EncSave> se;
  std::vector<DecSave> sd;
  float *mem32 = nullptr, *fin32 = nullptr, *pred = nullptr, *dpred = nullptr;
  float2 *st_enc = nullptr, *st_dec = nullptr;

  uint16_t *dA_hi = nullptr, *dA_lo = nullptr;
  std::map<int, ActBuf> dA;     // dY operand planes [rows][width], one view (and tensor maps) per width
  ActBuf dAT;                   // dY^T planes [Cmax][rows]  (A operand of the weight-gradient GEMMs)
  Linear XT, XTmem;             // X^T planes [Cmax][rows]   (B operand of the weight-gradient GEMMs)
  float *gA = nullptr, *gB = nullptr, *gWide = nullptr, *gAttn = nullptr, *gQc = nullptr, *gKvc = nullptr, *gMem = nullptr,
        *gEmbT = nullptr;

  // The weight-gradient branch of every linear layer (pack_xt + dW GEMM) runs on a side stream next to the
  // activation-gradient chain (dX GEMM, LayerNorm / attention backward): both are small launches that fill a
  // fraction of the SMs.  ev_dy: dY planes ready (main -> side); ev_side: side finished with dAT / XT (side -> main).
  cudaStream_t side = nullptr;
  cudaEvent_t ev_dy = nullptr, ev_side = nullptr, ev_fwd = nullptr;
  bool side_pending = false;
  bool use_side = true;     // SDVG_TRAIN_STREAMS=0 keeps everything on the caller's stream

  ~Trainer() {
    if (side) cudaStreamDestroy(side);
    if (ev_dy) cudaEventDestroy(ev_dy);
    if (ev_side) cudaEventDestroy(ev_side);
    if (ev_fwd) cudaEventDestroy(ev_fwd);
  }

  // Gradient-ready notifications for a data-parallel caller: while the backward pass is being enqueued, `ready_cb`
  // is called with consecutive ranges [offset, offset + count) of the flat gradient vector whose values are final
  // once the work enqueued so far on the stream has run (the caller records an event there and starts the
  // all-reduce of that range on its communication stream).  Ranges arrive from the end of the vector to its start
  // and tile it exactly; a range covers `layers_per_bucket` layers.
  typedef void (*ReadyFn)(void* user, long long offset, long long count);
  ReadyFn ready_cb = nullptr;
  void* ready_user = nullptr;
  int layers_per_bucket = 3;
  size_t ready_hi = 0;       // everything in [ready_hi, arena_count) has been announced

  cudaError_t announce(size_t lo, cudaStream_t st) {
    if (!ready_cb || lo >= ready_hi) return cudaSuccess;
    SDVG_CK(weight_branch_join(st));     // the weight gradients of these layers come from the side stream
    ready_cb(ready_user, static_cast<long long>(lo), static_cast<long long>(ready_hi - lo));
    ready_hi = lo;
    return cudaSuccess;
  }
  size_t offset_of(const float* p) const { return static_cast<size_t>(p - g.arena); }

  // Training-mode dropout (DROPOUT_P of the reference's configs; nn.Transformer applies it to the embedding + PE sum,
  // the attention probabilities, every sub-layer output before its residual add and the FFN hidden activation).
  // Masks come from a counter-based hash of (seed, step, site, element) - common.cuh drop_hash - so the backward pass
  // regenerates them instead of storing them; torch's own Philox stream is not reproduced (include/sdvg.h).
  float drop_p = 0.f;
  uint64_t drop_seed = 0;
  uint32_t drop_step = 0;
  enum Site { SA_P = 0, SA_OUT = 1, CA_P = 2, CA_OUT = 3, FF_H = 4, FF_OUT = 5 };
  Drop drop_at(uint32_t site) const {
    Drop dr;
    if (drop_p <= 0.f) return dr;
    double t = static_cast<double>(drop_p) * 4294967296.0;
    dr.thr = t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
    if (dr.thr == 0) dr.thr = 1;
    dr.key = drop_site_key(drop_seed, drop_step, site);
    dr.scale = 1.0f / (1.0f - drop_p);
    return dr;
  }
  static uint32_t enc_site(int l, int which) { return 1000u + static_cast<uint32_t>(l) * 10u + which; }
  static uint32_t dec_site(int l, int which) { return 2000u + static_cast<uint32_t>(l) * 10u + which; }
  static void set_drop(Epilogue& e, const Drop& dr, int cols) {
    e.drop_thr = dr.thr; e.drop_key = dr.key; e.drop_scale = dr.scale; e.drop_cols = cols;
  }

  // ------------------------------------------------------------------ set-up
  int slot_containing(const float* p) const {
    for (size_t i = 0; i < g.slots.size(); ++i)
      if (p >= g.slots[i].dev && p < g.slots[i].dev + g.slots[i].count) return static_cast<int>(i);
    return -1;
  }
  float* grad_of(const float* param) const { return grads + (param - g.arena); }

  bool make_tlinear(TLinear& tl, const Linear& L) {
    tl.fwd = L;
    const int si = slot_containing(L.w32);
    if (si < 0) return false;
    const WeightSlot& s = g.slots[si];
    const int K = static_cast<int>(s.shape[1]);
    const int row0 = static_cast<int>((L.w32 - s.dev) / K);
    tl.gw = grad_of(L.w32);
    tl.gb = grad_of(L.bias);
    tl.t.N = L.K; tl.t.K = L.N; tl.t.w32 = nullptr; tl.t.bias = nullptr; tl.t.split = L.split;
    tl.t.p.rows = L.K; tl.t.p.cols = L.N; tl.t.p.ld = wt[si].ld;
    tl.t.p.hi = wt[si].hi + row0;
    tl.t.p.lo = wt[si].lo ? wt[si].lo + row0 : nullptr;
    return g.map_planes(tl.t.p, true);
  }
  bool make_tattn(TAttn& t, const AttnWeights& a) {
    return make_tlinear(t.qkv, a.qkv) && make_tlinear(t.q, a.q) && make_tlinear(t.kv, a.kv) && make_tlinear(t.out, a.out);
  }
  LnGrad ln_grad(const LNParam& p) const { return LnGrad{grad_of(p.w), grad_of(p.b)}; }

  template <typename T>
  bool alloc(T** out, size_t count) { return g.dalloc(out, count) == cudaSuccess; }

  int init() {
    if (ready) return SDVG_OK;
    if (!g.tc()) return g.fail(SDVG_ERR_UNSUPPORTED, "training needs a tensor-core precision mode (fp32, mixed, fp16 or bf16)");
    if (g.cfg.max_tokens > kTrainMaxS) return g.fail(SDVG_ERR_UNSUPPORTED, "training supports at most %d tokens per clip", kTrainMaxS);
    const int d = g.cfg.dim_model, E = g.cfg.latent_dim, ff = g.cfg.dim_feedforward;
    const int Le = static_cast<int>(g.enc.size()), Ld = static_cast<int>(g.dec.size());
    rows = g.max_rows;
    const size_t n = g.arena_count;
    bool ok = alloc(&grads, n) && alloc(&adam_m, n) && alloc(&adam_v, n) && alloc(&scale, 2) && alloc(&amax, 1) && alloc(&loss_out, 8);
    if (!ok) return g.fail(SDVG_ERR_CUDA, "out of device memory (training state)");
    const float one[2] = {1.0f, 1.0f};
    cudaMemcpy(scale, one, sizeof one, cudaMemcpyHostToDevice);
    // transposed planes of every GEMM weight
    wt.resize(g.slots.size());
    for (size_t i = 0; i < g.slots.size(); ++i) {
      const WeightSlot& s = g.slots[i];
      if (!s.is_matrix) continue;
      const int N = static_cast<int>(s.shape[0]), K = static_cast<int>(s.shape[1]);
      wt[i].ld = round_up(N, kTcBK);
      ok = ok && alloc(&wt[i].hi, static_cast<size_t>(K) * wt[i].ld);
      if (s.need_lo) ok = ok && alloc(&wt[i].lo, static_cast<size_t>(K) * wt[i].ld);
    }
    if (!ok) return g.fail(SDVG_ERR_CUDA, "out of device memory (transposed weights)");
    ok = make_tlinear(t_emb, g.embedding) && make_tlinear(t_out, g.out_proj);
    tenc.resize(Le); tdec.resize(Ld);
    for (int l = 0; l < Le && ok; ++l) {
      ok = make_tattn(tenc[l].sa, g.enc[l].sa) && make_tlinear(tenc[l].ff1, g.enc[l].ff1) && make_tlinear(tenc[l].ff2, g.enc[l].ff2);
      tenc[l].n1 = ln_grad(g.enc[l].n1); tenc[l].n2 = ln_grad(g.enc[l].n2);
    }
    for (int l = 0; l < Ld && ok; ++l) {
      ok = make_tattn(tdec[l].sa, g.dec[l].sa) && make_tattn(tdec[l].ca, g.dec[l].ca) && make_tlinear(tdec[l].ff1, g.dec[l].ff1) &&
           make_tlinear(tdec[l].ff2, g.dec[l].ff2);
      tdec[l].n1 = ln_grad(g.dec[l].n1); tdec[l].n2 = ln_grad(g.dec[l].n2); tdec[l].n3 = ln_grad(g.dec[l].n3);
    }
    if (!ok) return g.fail(SDVG_ERR_CUDA, "tensor map creation failed (transposed weights)");
    g_encnorm = ln_grad(g.enc_norm); g_decnorm = ln_grad(g.dec_norm);
    decoder_offset = Ld > 0 ? static_cast<size_t>(g.dec[0].sa.qkv.w32 - g.arena) : static_cast<size_t>(g.dec_norm.w - g.arena);

    // saved activations
    const size_t R = static_cast<size_t>(rows);
    xe.resize(Le + 1); xd.resize(Ld + 1); se.resize(Le); sd.resize(Ld);
    for (auto& p : xe) ok = ok && alloc(&p, R * d);
    for (auto& p : xd) ok = ok && alloc(&p, R * d);
    for (auto& s : se)
      ok = ok && alloc(&s.qkv, R * 3 * d) && alloc(&s.a, R * d) && alloc(&s.y1, R * d) && alloc(&s.x1, R * d) && alloc(&s.h, R * ff) &&
           alloc(&s.y2, R * d) && alloc(&s.st1, R) && alloc(&s.st2, R);
    for (auto& s : sd)
      ok = ok && alloc(&s.qkv, R * 3 * d) && alloc(&s.a, R * d) && alloc(&s.y1, R * d) && alloc(&s.x1, R * d) && alloc(&s.qc, R * d) &&
           alloc(&s.kvc, R * 2 * d) && alloc(&s.ac, R * d) && alloc(&s.y2, R * d) && alloc(&s.x2, R * d) && alloc(&s.h, R * ff) &&
           alloc(&s.y3, R * d) && alloc(&s.st1, R) && alloc(&s.st2, R) && alloc(&s.st3, R);
    ok = ok && alloc(&mem32, R * d) && alloc(&fin32, R * d) && alloc(&pred, R * E) && alloc(&dpred, R * E) && alloc(&st_enc, R) &&
         alloc(&st_dec, R);
    // backward work buffers
    const int wide = std::max(std::max(3 * d, ff), E);
    const bool lo = g.split_first();
    ok = ok && alloc(&dA_hi, R * round_up(wide, kTcBK));
    if (lo) ok = ok && alloc(&dA_lo, R * round_up(wide, kTcBK));
    ok = ok && alloc(&gA, R * d) && alloc(&gB, R * d) && alloc(&gWide, R * wide) && alloc(&gAttn, R * d) && alloc(&gQc, R * d) &&
         alloc(&gKvc, R * 2 * d) && alloc(&gMem, R * d) && alloc(&gEmbT, R * d);
    if (!ok) return g.fail(SDVG_ERR_CUDA, "out of device memory (training workspace)");
    for (int w : {d, 2 * d, 3 * d, ff, E}) {
      if (dA.count(w)) continue;
      ActBuf v;
      v.p.rows = rows; v.p.cols = w; v.p.ld = round_up(w, kTcBK); v.p.hi = dA_hi; v.p.lo = dA_lo;
      if (!g.map_planes(v.p, false)) return g.fail(SDVG_ERR_CUDA, "tensor map creation failed (gradient planes)");
      dA[w] = v;
    }
    const int cmax = wide;
    dAT.p.rows = cmax; dAT.p.cols = rows; dAT.p.ld = rows;
    XT.p.rows = cmax; XT.p.cols = rows; XT.p.ld = rows;
    XTmem.p.rows = d; XTmem.p.cols = rows; XTmem.p.ld = rows;
    ok = alloc(&dAT.p.hi, static_cast<size_t>(cmax) * rows) && alloc(&XT.p.hi, static_cast<size_t>(cmax) * rows) &&
         alloc(&XTmem.p.hi, static_cast<size_t>(d) * rows);
    if (lo) ok = ok && alloc(&dAT.p.lo, static_cast<size_t>(cmax) * rows) && alloc(&XT.p.lo, static_cast<size_t>(cmax) * rows) &&
                 alloc(&XTmem.p.lo, static_cast<size_t>(d) * rows);
    if (!ok) return g.fail(SDVG_ERR_CUDA, "out of device memory (transposed operands)");
    if (!g.map_planes(dAT.p, false) || !g.map_planes(XT.p, true) || !g.map_planes(XTmem.p, true))
      return g.fail(SDVG_ERR_CUDA, "tensor map creation failed (transposed operands)");
    XT.split = lo; XTmem.split = lo;
    if (const char* v = std::getenv("SDVG_TRAIN_STREAMS")) use_side = std::atoi(v) != 0;
    if (use_side) {
      if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_dy, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_side, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_fwd, cudaEventDisableTiming) != cudaSuccess)
        return g.fail(SDVG_ERR_CUDA, "stream / event creation failed");
    }
    ready = true;
    wt_stale = true;
    return SDVG_OK;
  }

  // ------------------------------------------------------------------ weight planes
  bool wt_stale = true;
  // (Re)build the operand planes of every GEMM weight from the fp32 parameters: W planes [N][K] and W^T planes [K][N].
  cudaError_t repack_weights(bool forward_planes, cudaStream_t st, size_t lo = 0, size_t hi = ~static_cast<size_t>(0)) {
    for (size_t i = 0; i < g.slots.size(); ++i) {
      const WeightSlot& s = g.slots[i];
      if (!s.is_matrix) continue;
      const size_t off = static_cast<size_t>(s.dev - g.arena);
      if (off < lo || off >= hi) continue;     // only the matrices of the parameter range [lo, hi)
      PackTArgs a{};
      a.src = s.dev; a.ld_src = static_cast<int>(s.shape[1]); a.R = static_cast<int>(s.shape[0]); a.C = static_cast<int>(s.shape[1]);
      a.mul = 1.0f;
      if (forward_planes) { a.out_hi = s.hi; a.out_lo = s.lo; a.ld16 = s.ld16; }
      a.t_hi = wt[i].hi; a.t_lo = wt[i].lo; a.ld_t = wt[i].ld; a.Rpad = wt[i].ld;
      a.bf16 = g.bf16();
      Engine::Scope sc(&g, KC_PACK, 0.0, static_cast<double>(s.count) * (forward_planes ? 12.0 : 8.0), st);
      SDVG_CK(launch_pack_t(a, st));
    }
    if (lo == 0) {   // a full rebuild, or the last range of a ranged one (ranges run from the end of the arena to its start)
      wt_stale = false;
      if (forward_planes) {   // the rollout path's derived operands follow the weights: stacked cross-attention K|V, folded LayerNorms
        SDVG_CK(g.restack_cross(st));
        SDVG_CK(g.refold(st));
      }
    }
    return cudaSuccess;
  }

  // ------------------------------------------------------------------ small wrappers
  ActBuf act(float* f32, int ld, const Planes& p) const { ActBuf a; a.f32 = f32; a.ld32 = ld; a.p = p; return a; }

  cudaError_t attention_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, int B, int Sq, int Sk, int causal,
                            float* out32, cudaStream_t st, const Drop& dr = Drop{}) {
    if (dr.thr) {
      AttnTrainArgs t{};
      t.q = q; t.ldq = ldq; t.k = k; t.v = v; t.ldkv = ldkv;
      t.clips = B; t.heads = g.cfg.num_heads; t.hd = g.cfg.dim_model / g.cfg.num_heads; t.Sq = Sq; t.Sk = Sk; t.causal = causal;
      t.scale = 1.0f / sqrtf(static_cast<float>(t.hd));
      t.out32 = out32; t.ld32 = g.cfg.dim_model;
      t.out_hi = g.attn.p.hi; t.out_lo = g.attn.p.lo; t.ld16 = g.attn.p.ld; t.bf16 = g.bf16();
      t.drop_thr = dr.thr; t.drop_key = dr.key; t.drop_scale = dr.scale;
      Engine::Scope sc(&g, KC_ATTN, 0.0, 4.0 * B * g.cfg.dim_model * (2.0 * Sq + 2.0 * Sk), st);
      return launch_attention_train_fwd(t, st);
    }
    AttnArgs a{};
    a.q = q; a.ldq = ldq; a.k = k; a.v = v; a.ldkv = ldkv;
    a.clips = B; a.heads = g.cfg.num_heads; a.hd = g.cfg.dim_model / g.cfg.num_heads; a.Sq = Sq; a.Sk = Sk;
    a.mask_kind = causal ? 1 : 0;
    a.scale = 1.0f / sqrtf(static_cast<float>(a.hd));
    a.out32 = out32; a.ld32 = g.cfg.dim_model;
    a.out_hi = g.attn.p.hi; a.out_lo = g.attn.p.lo; a.ld16 = g.attn.p.ld; a.bf16 = g.bf16();
    Engine::Scope sc(&g, KC_ATTN, 0.0, 4.0 * B * g.cfg.dim_model * (2.0 * Sq + 2.0 * Sk), st);
    return launch_attention(a, st);
  }

  cudaError_t ln_fwd(const float* y, int M, int S, const LNParam& n, float* out32, const Planes& planes, float2* stats, cudaStream_t st) {
    ActBuf in = act(const_cast<float*>(y), g.cfg.dim_model, Planes{});
    ActBuf dst = act(out32, g.cfg.dim_model, planes);
    return g.layernorm(in, M, n, nullptr, dst, true, S, 0, st, stats);
  }

  cudaError_t ln_bwd(const float* dy, const float* y, const float2* stats, const LNParam& n, const LnGrad& gp, int M, float* dx,
                     cudaStream_t st) {
    LnBwdArgs a{};
    const int d = g.cfg.dim_model;
    a.dy = dy; a.ld_dy = d; a.x = y; a.ld_x = d; a.stats = stats; a.w = n.w; a.rows = M; a.d = d;
    a.dx = dx; a.ld_dx = d; a.dw = gp.w; a.db = gp.b; a.param_mul_dev = scale + 1;
    Engine::Scope sc(&g, KC_LN, 0.0, 4.0 * M * d * 5.0, st);
    return launch_ln_backward(a, st);
  }

  cudaError_t attention_bwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* dO, float* dq, int ld_dq,
                            float* dk, float* dv, int ld_dkv, int B, int Sq, int Sk, int causal, cudaStream_t st,
                            const Drop& dr = Drop{}) {
    AttnBwdArgs a{};
    a.drop_thr = dr.thr; a.drop_key = dr.key; a.drop_scale = dr.scale;
    a.q = q; a.ldq = ldq; a.k = k; a.v = v; a.ldkv = ldkv; a.dO = dO; a.ld_do = g.cfg.dim_model;
    a.dq = dq; a.ld_dq = ld_dq; a.dk = dk; a.dv = dv; a.ld_dkv = ld_dkv;
    a.clips = B; a.heads = g.cfg.num_heads; a.hd = g.cfg.dim_model / g.cfg.num_heads; a.Sq = Sq; a.Sk = Sk; a.causal = causal;
    a.scale = 1.0f / sqrtf(static_cast<float>(a.hd));
    Engine::Scope sc(&g, KC_ATTN, 0.0, 4.0 * B * g.cfg.dim_model * (3.0 * Sq + 4.0 * Sk), st);
    return launch_attention_backward(a, st);
  }

  // dY fp32 [M][N] -> operand planes, transposed operand planes, bias gradient
  cudaError_t pack_dy(const float* dy, int ld, int M, int N, float* gb, bool accumulate_b, cudaStream_t st, float mul = 1.0f,
                      const float* mul_dev = nullptr, int perm_S = 0, int perm_B = 0, bool want_planes = true,
                      const Drop& dr = Drop{}) {
    PackTArgs a{};
    a.drop_thr = dr.thr; a.drop_key = dr.key; a.drop_scale = dr.scale;
    a.src = dy; a.ld_src = ld; a.R = M; a.C = N; a.perm_S = perm_S; a.perm_B = perm_B; a.mul = mul; a.mul_dev = mul_dev;
    if (want_planes) { const ActBuf& v = dA.at(N); a.out_hi = v.p.hi; a.out_lo = v.p.lo; a.ld16 = v.p.ld; }
    a.t_hi = dAT.p.hi; a.t_lo = dAT.p.lo; a.ld_t = dAT.p.ld; a.Rpad = round_up(M, kTcBK);
    a.colsum = gb; a.colsum_mul_dev = scale + 1; a.colsum_accumulate = accumulate_b ? 1 : 0;
    a.bf16 = g.bf16();
    Engine::Scope sc(&g, KC_PACK, 0.0, static_cast<double>(M) * N * 12.0, st);
    return launch_pack_t(a, st);
  }
  // X fp32 [M][K] -> transposed operand planes
  cudaError_t pack_xt(const float* x, int ld, int M, int K, Linear& dst, cudaStream_t st) {
    PackTArgs a{};
    a.src = x; a.ld_src = ld; a.R = M; a.C = K; a.mul = 1.0f;
    a.t_hi = dst.p.hi; a.t_lo = dst.p.lo; a.ld_t = dst.p.ld; a.Rpad = round_up(M, kTcBK);
    a.bf16 = g.bf16();
    Engine::Scope sc(&g, KC_PACK, 0.0, static_cast<double>(M) * K * 8.0, st);
    return launch_pack_t(a, st);
  }
  // dW[N][K] (+)= (1/S) dY^T X, from the planes left by pack_dy / pack_xt
  cudaError_t grad_w(int N, int K, int M, float* gw, bool accumulate, Linear& xt, cudaStream_t st) {
    Linear L = xt;
    L.N = K; L.K = round_up(M, kTcBK); L.bias = nullptr;
    Epilogue e;
    e.out32 = gw; e.ld32 = K; e.alpha_dev = scale + 1;
    if (accumulate) { e.residual = gw; e.ld_res = K; }
    return g.gemm(dAT, L, N, e, st);
  }
  // dX[M][K] = dY W (+ residual), from the planes left by pack_dy
  cudaError_t grad_x(const TLinear& tl, int M, float* dx, const float* residual, const float* gate, int ld_gate, cudaStream_t st,
                     float gate_scale = 1.0f) {
    Epilogue e;
    e.out32 = dx; e.ld32 = tl.t.N;
    if (residual) { e.residual = residual; e.ld_res = tl.t.N; }
    if (gate) { e.gate = gate; e.ld_gate = ld_gate; e.gate_scale = gate_scale; }
    return g.gemm(dA.at(tl.t.K), tl.t, M, e, st);
  }
  // full backward of one linear layer whose input x (fp32, saved) has M rows
  // out_drop: dropout that the forward pass applied to this layer's OUTPUT (dy is masked the same way on the way in);
  // gate / gate_scale: ReLU (+ dropout) that the forward pass applied to this layer's INPUT (masks dx)
  cudaError_t linear_bwd(const TLinear& tl, const float* dy, int ld_dy, const float* x, int M, float* dx, const float* residual,
                         cudaStream_t st, const float* gate = nullptr, int ld_gate = 0, Linear* xt_ready = nullptr,
                         const Drop& out_drop = Drop{}, float gate_scale = 1.0f) {
    SDVG_CK(weight_branch_begin(st));
    SDVG_CK(pack_dy(dy, ld_dy, M, tl.fwd.N, tl.gb, false, st, 1.0f, nullptr, 0, 0, true, out_drop));
    SDVG_CK(weight_branch(tl.fwd.N, tl.fwd.K, M, tl.gw, false, xt_ready ? nullptr : x, tl.fwd.K, xt_ready ? *xt_ready : XT, st));
    if (dx) SDVG_CK(grad_x(tl, M, dx, residual, gate, ld_gate, st, gate_scale));
    return cudaSuccess;
  }
  // before dAT is overwritten: the side stream must be done with the previous layer's dW GEMM
  cudaError_t weight_branch_begin(cudaStream_t st) {
    if (side && side_pending) { SDVG_CK(cudaStreamWaitEvent(st, ev_side, 0)); side_pending = false; }
    return cudaSuccess;
  }
  // pack_xt (unless the transposed operand is already there) + dW GEMM, on the side stream when enabled
  cudaError_t weight_branch(int N, int K, int M, float* gw, bool accumulate, const float* x, int ld_x, Linear& xt, cudaStream_t st) {
    cudaStream_t ws = side ? side : st;
    if (side) { SDVG_CK(cudaEventRecord(ev_dy, st)); }
    if (x) SDVG_CK(pack_xt(x, ld_x, M, K, xt, ws));
    if (side) SDVG_CK(cudaStreamWaitEvent(side, ev_dy, 0));
    SDVG_CK(grad_w(N, K, M, gw, accumulate, xt, ws));
    if (side) { SDVG_CK(cudaEventRecord(ev_side, side)); side_pending = true; }
    return cudaSuccess;
  }
  // the caller's stream waits for everything the side stream still has in flight
  cudaError_t weight_branch_join(cudaStream_t st) {
    if (side && side_pending) { SDVG_CK(cudaStreamWaitEvent(st, ev_side, 0)); side_pending = false; }
    return cudaSuccess;
  }
  // the side stream may not start before the forward pass (saved activations, loss scale) is complete
  cudaError_t weight_branch_fork(cudaStream_t st) {
    if (!side) return cudaSuccess;
    SDVG_CK(cudaEventRecord(ev_fwd, st));
    return cudaStreamWaitEvent(side, ev_fwd, 0);
  }

  // ------------------------------------------------------------------ forward with saved activations
  int Bc = 0, Ssc = 0, Stc = 0;        // shapes of the saved forward
  const float* src_c = nullptr; const float* tgt_c = nullptr;

  cudaError_t forward(const float* src, const float* tgt, int B, int Ss, int St, const int* pe_index, cudaStream_t st) {
    const int d = g.cfg.dim_model, E = g.cfg.latent_dim;
    const int Ms = B * Ss, Mt = B * St;
    const int Le = static_cast<int>(g.enc.size()), Ld = static_cast<int>(g.dec.size());
    const float sqrt_d = sqrtf(static_cast<float>(d));
    Bc = B; Ssc = Ss; Stc = St; src_c = src; tgt_c = tgt;
    ++drop_step;
    SDVG_CK(g.ingest(src, static_cast<long long>(Ss) * E, E, nullptr, B, Ss, 1.0f, g.lat_s, st));
    SDVG_CK(g.ingest(tgt, static_cast<long long>(St) * E, E, nullptr, B, St, 1.0f, g.lat_t, st));
    auto embed = [&](const ActBuf& lat, int S, float* out32, const ActBuf& planes, uint32_t site) -> cudaError_t {
      Epilogue e;  // models/transformer.py:53-56, dropout of positional_encoding.py:35
      e.alpha = sqrt_d; e.pe = g.pe_table; e.ld_pe = d; e.pe_index = pe_index; e.rows_per_clip = S;
      set_drop(e, drop_at(site), d);
      e.out32 = out32; e.ld32 = d; e.out_hi = planes.p.hi; e.out_lo = planes.p.lo; e.ld16 = planes.p.ld;
      return g.gemm(lat, g.embedding, B * S, e, st);
    };
    auto plain = [&](float* out32, int ld) { Epilogue e; e.out32 = out32; e.ld32 = ld; return e; };
    // ---- encoder
    SDVG_CK(embed(g.lat_s, Ss, xe[0], g.emb_s, 0));
    const ActBuf* x = &g.emb_s;
    for (int l = 0; l < Le; ++l) {
      const EncLayer& L = g.enc[l]; EncSave& s = se[l];
      SDVG_CK(g.gemm(*x, L.sa.qkv, Ms, plain(s.qkv, 3 * d), st));
      SDVG_CK(attention_fwd(s.qkv, 3 * d, s.qkv + d, s.qkv + 2 * d, 3 * d, B, Ss, Ss, 0, s.a, st, drop_at(enc_site(l, SA_P))));
      Epilogue eo = plain(s.y1, d); eo.residual = xe[l]; eo.ld_res = d;
      set_drop(eo, drop_at(enc_site(l, SA_OUT)), d);
      SDVG_CK(g.gemm(g.attn, L.sa.out, Ms, eo, st));
      SDVG_CK(ln_fwd(s.y1, Ms, Ss, L.n1, s.x1, g.xs.p, s.st1, st));
      Epilogue e1 = plain(s.h, g.cfg.dim_feedforward); e1.relu = 1;
      e1.out_hi = g.ffh.p.hi; e1.out_lo = g.ffh.p.lo; e1.ld16 = g.ffh.p.ld;
      set_drop(e1, drop_at(enc_site(l, FF_H)), g.cfg.dim_feedforward);
      SDVG_CK(g.gemm(g.xs, L.ff1, Ms, e1, st));
      Epilogue e2 = plain(s.y2, d); e2.residual = s.x1; e2.ld_res = d;
      set_drop(e2, drop_at(enc_site(l, FF_OUT)), d);
      SDVG_CK(g.gemm(g.ffh, L.ff2, Ms, e2, st));
      SDVG_CK(ln_fwd(s.y2, Ms, Ss, L.n2, xe[l + 1], g.xs.p, s.st2, st));
      x = &g.xs;
    }
    SDVG_CK(ln_fwd(xe[Le], Ms, Ss, g.enc_norm, mem32, g.mem.p, st_enc, st));
    // ---- decoder
    SDVG_CK(embed(g.lat_t, St, xd[0], g.emb_t, 1));
    const ActBuf* y = &g.emb_t;
    for (int l = 0; l < Ld; ++l) {
      const DecLayer& L = g.dec[l]; DecSave& s = sd[l];
      SDVG_CK(g.gemm(*y, L.sa.qkv, Mt, plain(s.qkv, 3 * d), st));
      SDVG_CK(attention_fwd(s.qkv, 3 * d, s.qkv + d, s.qkv + 2 * d, 3 * d, B, St, St, 1, s.a, st, drop_at(dec_site(l, SA_P))));
      Epilogue eo = plain(s.y1, d); eo.residual = xd[l]; eo.ld_res = d;
      set_drop(eo, drop_at(dec_site(l, SA_OUT)), d);
      SDVG_CK(g.gemm(g.attn, L.sa.out, Mt, eo, st));
      SDVG_CK(ln_fwd(s.y1, Mt, St, L.n1, s.x1, g.xt.p, s.st1, st));
      SDVG_CK(g.gemm(g.xt, L.ca.q, Mt, plain(s.qc, d), st));
      SDVG_CK(g.gemm(g.mem, L.ca.kv, Ms, plain(s.kvc, 2 * d), st));
      SDVG_CK(attention_fwd(s.qc, d, s.kvc, s.kvc + d, 2 * d, B, St, Ss, 0, s.ac, st, drop_at(dec_site(l, CA_P))));
      Epilogue eo2 = plain(s.y2, d); eo2.residual = s.x1; eo2.ld_res = d;
      set_drop(eo2, drop_at(dec_site(l, CA_OUT)), d);
      SDVG_CK(g.gemm(g.attn, L.ca.out, Mt, eo2, st));
      SDVG_CK(ln_fwd(s.y2, Mt, St, L.n2, s.x2, g.xt.p, s.st2, st));
      Epilogue e1 = plain(s.h, g.cfg.dim_feedforward); e1.relu = 1;
      e1.out_hi = g.ffh.p.hi; e1.out_lo = g.ffh.p.lo; e1.ld16 = g.ffh.p.ld;
      set_drop(e1, drop_at(dec_site(l, FF_H)), g.cfg.dim_feedforward);
      SDVG_CK(g.gemm(g.xt, L.ff1, Mt, e1, st));
      Epilogue e2 = plain(s.y3, d); e2.residual = s.x2; e2.ld_res = d;
      set_drop(e2, drop_at(dec_site(l, FF_OUT)), d);
      SDVG_CK(g.gemm(g.ffh, L.ff2, Mt, e2, st));
      SDVG_CK(ln_fwd(s.y3, Mt, St, L.n3, xd[l + 1], g.xt.p, s.st3, st));
      y = &g.xt;
    }
    SDVG_CK(ln_fwd(xd[Ld], Mt, St, g.dec_norm, fin32, g.fin.p, st_dec, st));
    Epilogue oe = plain(pred, E);   // (S_tgt, B, E) like the reference's return value (models/transformer.py:60-68)
    oe.row_map = 1; oe.rows_per_clip = St; oe.clips = B;
    return g.gemm(g.fin, g.out_proj, Mt, oe, st);
  }

  // ------------------------------------------------------------------ criterion: values + d/dpred
  int loss_and_grad(const float* expected, int B, int St, const TrainLoss& lc, cudaStream_t st) {
    const int E = g.cfg.latent_dim;
    int hs = 1;
    while (4 * hs * hs < E) ++hs;
    if (4 * hs * hs != E) return g.fail(SDVG_ERR_INVALID, "latent width %d is not 4*h*h", E);
    const int P = lc.P;
    if (P <= 0 || P > St) return g.fail(SDVG_ERR_INVALID, "frames_to_predict %d outside (0, %d]", P, St);
    const size_t off = static_cast<size_t>(St - P) * B * E;
    int rc = sdvg_criterion(g.cfg.device, pred + off, expected + off, P, B, hs, hs, lc.use_mse, lc.use_l1, lc.use_gdl, lc.lambda_gdl,
                            lc.alpha, lc.use_nce, lc.temperature, lc.lambda_nce, loss_out, st);
    g.launches += lc.use_nce ? 3 : 2;
    if (rc != SDVG_OK) return g.fail(rc, "criterion failed: %s", sdvg_last_error(nullptr));
    cudaError_t e = cudaSuccess;
    if (P < St) e = cudaMemsetAsync(dpred, 0, off * sizeof(float), st);
    LossGradArgs a{};
    a.x = pred + off; a.y = expected + off; a.P = P; a.B = B; a.h = hs; a.w = hs;
    const double numel = static_cast<double>(P) * B * E;
    a.c_mse = lc.use_mse ? static_cast<float>(1.0 / numel) : 0.f;
    a.c_l1 = lc.use_l1 ? static_cast<float>(1.0 / numel) : 0.f;
    a.c_gdl = lc.use_gdl ? static_cast<float>(static_cast<double>(lc.lambda_gdl) / numel) : 0.f;
    a.alpha = lc.alpha;
    a.c_nce = lc.use_nce ? static_cast<float>(0.5 * lc.lambda_nce / (static_cast<double>(P) * B * hs * hs)) : 0.f;
    a.inv_temperature = 1.0f / lc.temperature;
    a.grad = dpred + off; a.amax_bits = amax;
    const long long total = static_cast<long long>(P) * B * E;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > g.num_sms * 8) blocks = g.num_sms * 8;
    if (e == cudaSuccess) {
      Engine::Scope sc(&g, KC_PACK, 0.0, 12.0 * total, st);
      e = launch_kernel(loss_grad_elementwise_kernel, dim3(blocks), dim3(256), 0, st, a);
    }
    if (e == cudaSuccess && lc.use_nce) {
      const int hw = hs * hs;
      const size_t smem = static_cast<size_t>(hw) * 2 * sizeof(float4);
      if (smem > 48 * 1024) {
        static bool attr_set[64] = {};
        if (!attr_set[g.cfg.device & 63]) {
          e = cudaFuncSetAttribute(loss_grad_nce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          attr_set[g.cfg.device & 63] = e == cudaSuccess;
        }
      }
      const int threads = hw >= 256 ? 256 : (hw >= 128 ? 128 : 64);
      if (e == cudaSuccess) {
        Engine::Scope sc(&g, KC_PACK, 0.0, 16.0 * total, st);
        e = launch_kernel(loss_grad_nce_kernel, dim3(P * B), dim3(threads), smem, st, a);
      }
    }
    if (e == cudaSuccess) {
      Engine::Scope sc(&g, KC_PACK, 0.0, 16.0, st);
      e = launch_kernel(loss_scale_kernel, dim3(1), dim3(1), 0, st, amax, scale);
    }
    if (e != cudaSuccess) return g.fail_cuda(e, "criterion gradient");
    return SDVG_OK;
  }

  // ------------------------------------------------------------------ backward
  // part 1: output projection, decoder.norm, decoder layers (gradients of the decoder-side bucket are final);
  // part 2: target embedding, encoder.norm, encoder layers, source embedding.
  cudaError_t backward_decoder(cudaStream_t st) {
    const int d = g.cfg.dim_model, E = g.cfg.latent_dim, ff = g.cfg.dim_feedforward;
    const int B = Bc, Ss = Ssc, St = Stc, Ms = B * Ss, Mt = B * St;
    const int Ld = static_cast<int>(g.dec.size());
    // out projection: dpred is (S_tgt, B, E) and unscaled -> clip-major rows, multiplied by the loss scale
    ready_hi = g.arena_count;
    SDVG_CK(weight_branch_fork(st));
    SDVG_CK(weight_branch_begin(st));
    SDVG_CK(pack_dy(dpred, E, Mt, E, t_out.gb, false, st, 1.0f, scale, St, B));
    SDVG_CK(weight_branch(E, d, Mt, t_out.gw, false, fin32, d, XT, st));
    SDVG_CK(grad_x(t_out, Mt, gA, nullptr, nullptr, 0, st));
    SDVG_CK(ln_bwd(gA, xd[Ld], st_dec, g.dec_norm, g_decnorm, Mt, gB, st));
    float* gin = gB;    // gradient w.r.t. the current layer's output
    float* gtmp = gA;
    bool mem_started = false;
    for (int l = Ld - 1; l >= 0; --l) {
      const DecLayer& L = g.dec[l]; const TDec& T = tdec[l]; const DecSave& s = sd[l];
      // x_out = LN3(y3), y3 = x2 + FFN(x2)
      SDVG_CK(ln_bwd(gin, s.y3, s.st3, L.n3, T.n3, Mt, gtmp, st));                       // gtmp = dy3
      const float ds = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
      SDVG_CK(linear_bwd(T.ff2, gtmp, d, s.h, Mt, gWide, nullptr, st, s.h, ff, nullptr, drop_at(dec_site(l, FF_OUT)), ds));   // gWide = dh (ReLU / dropout gated)
      SDVG_CK(linear_bwd(T.ff1, gWide, ff, s.x2, Mt, gin, gtmp, st));                    // gin = dx2 = dh W1 + dy3
      // x2 = LN2(y2), y2 = x1 + CA(x1, mem)
      SDVG_CK(ln_bwd(gin, s.y2, s.st2, L.n2, T.n2, Mt, gtmp, st));                       // gtmp = dy2
      SDVG_CK(linear_bwd(T.ca.out, gtmp, d, s.ac, Mt, gAttn, nullptr, st, nullptr, 0, nullptr, drop_at(dec_site(l, CA_OUT))));   // gAttn = d(attention output)
      SDVG_CK(attention_bwd(s.qc, d, s.kvc, s.kvc + d, 2 * d, gAttn, gQc, d, gKvc, gKvc + d, 2 * d, B, St, Ss, 0, st,
                            drop_at(dec_site(l, CA_P))));
      SDVG_CK(linear_bwd(T.ca.q, gQc, d, s.x1, Mt, gin, gtmp, st));                      // gin = dx1 = dq Wq + dy2
      if (l == Ld - 1) SDVG_CK(pack_xt(mem32, d, Ms, d, XTmem, st));   // main stream: ordered before every later ev_dy
      SDVG_CK(linear_bwd(T.ca.kv, gKvc, 2 * d, nullptr, Ms, gMem, mem_started ? gMem : nullptr, st, nullptr, 0, &XTmem));
      mem_started = true;
      // x1 = LN1(y1), y1 = x + SA(x)
      SDVG_CK(ln_bwd(gin, s.y1, s.st1, L.n1, T.n1, Mt, gtmp, st));                       // gtmp = dy1
      SDVG_CK(linear_bwd(T.sa.out, gtmp, d, s.a, Mt, gAttn, nullptr, st, nullptr, 0, nullptr, drop_at(dec_site(l, SA_OUT))));
      SDVG_CK(attention_bwd(s.qkv, 3 * d, s.qkv + d, s.qkv + 2 * d, 3 * d, gAttn, gWide, 3 * d, gWide + d, gWide + 2 * d, 3 * d, B, St,
                            St, 1, st, drop_at(dec_site(l, SA_P))));
      SDVG_CK(linear_bwd(T.sa.qkv, gWide, 3 * d, xd[l], Mt, l == 0 ? gEmbT : gin, gtmp, st));   // dx = dqkv Wqkv + dy1
      if (l == 0 || (Ld - l) % layers_per_bucket == 0) SDVG_CK(announce(offset_of(L.sa.qkv.w32), st));
    }
    SDVG_CK(weight_branch_join(st));
    SDVG_CK(announce(decoder_offset, st));
    if (Ld == 0) {
      SDVG_CK(cudaMemcpyAsync(gEmbT, gB, static_cast<size_t>(Mt) * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
      SDVG_CK(cudaMemsetAsync(gMem, 0, static_cast<size_t>(Ms) * d * sizeof(float), st));
    }
    return cudaSuccess;
  }

  cudaError_t backward_encoder(cudaStream_t st) {
    const int d = g.cfg.dim_model, E = g.cfg.latent_dim, ff = g.cfg.dim_feedforward;
    const int B = Bc, Ss = Ssc, St = Stc, Ms = B * Ss, Mt = B * St;
    const int Le = static_cast<int>(g.enc.size());
    const float sqrt_d = sqrtf(static_cast<float>(d));
    // target embedding: emb = (x W^T + b) sqrt(d) + PE
    SDVG_CK(weight_branch_fork(st));
    SDVG_CK(weight_branch_begin(st));
    SDVG_CK(pack_dy(gEmbT, d, Mt, d, t_emb.gb, false, st, sqrt_d, nullptr, 0, 0, false, drop_at(1)));
    SDVG_CK(weight_branch(d, E, Mt, t_emb.gw, false, tgt_c, E, XT, st));
    // encoder.norm
    SDVG_CK(ln_bwd(gMem, xe[Le], st_enc, g.enc_norm, g_encnorm, Ms, gB, st));
    float* gin = gB;
    float* gtmp = gA;
    for (int l = Le - 1; l >= 0; --l) {
      const EncLayer& L = g.enc[l]; const TEnc& T = tenc[l]; const EncSave& s = se[l];
      SDVG_CK(ln_bwd(gin, s.y2, s.st2, L.n2, T.n2, Ms, gtmp, st));                       // dy2
      const float ds = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
      SDVG_CK(linear_bwd(T.ff2, gtmp, d, s.h, Ms, gWide, nullptr, st, s.h, ff, nullptr, drop_at(enc_site(l, FF_OUT)), ds));
      SDVG_CK(linear_bwd(T.ff1, gWide, ff, s.x1, Ms, gin, gtmp, st));                    // dx1
      SDVG_CK(ln_bwd(gin, s.y1, s.st1, L.n1, T.n1, Ms, gtmp, st));                       // dy1
      SDVG_CK(linear_bwd(T.sa.out, gtmp, d, s.a, Ms, gAttn, nullptr, st, nullptr, 0, nullptr, drop_at(enc_site(l, SA_OUT))));
      SDVG_CK(attention_bwd(s.qkv, 3 * d, s.qkv + d, s.qkv + 2 * d, 3 * d, gAttn, gWide, 3 * d, gWide + d, gWide + 2 * d, 3 * d, B, Ss,
                            Ss, 0, st, drop_at(enc_site(l, SA_P))));
      SDVG_CK(linear_bwd(T.sa.qkv, gWide, 3 * d, xe[l], Ms, gin, gtmp, st));             // dx
      if (l == 0 || (Le - l) % layers_per_bucket == 0) SDVG_CK(announce(offset_of(L.sa.qkv.w32), st));
    }
    // source embedding (same weights as the target embedding: accumulate)
    SDVG_CK(weight_branch_begin(st));
    SDVG_CK(pack_dy(gin, d, Ms, d, t_emb.gb, true, st, sqrt_d, nullptr, 0, 0, false, drop_at(0)));
    SDVG_CK(weight_branch(d, E, Ms, t_emb.gw, true, src_c, E, XT, st));
    SDVG_CK(weight_branch_join(st));
    return announce(0, st);
  }

  // forward + criterion + backward.  part 0: everything; 1: up to and including the decoder backward; 2: the rest.
  int forward_backward(const float* src, const float* tgt, const float* expected, int B, int Ss, int St, const TrainLoss& lc,
                       const int* pe_index, float* losses, int part, cudaStream_t st) {
    int rc = init();
    if (rc != SDVG_OK) return rc;
    if (part < 0 || part > 2) return g.fail(SDVG_ERR_INVALID, "part must be 0, 1 or 2");
    if (part != 2) {
      if (!src || !tgt || !expected) return g.fail(SDVG_ERR_INVALID, "null tensor");
      if (B <= 0 || Ss <= 0 || St <= 0 || B > g.cfg.max_clips || Ss > g.cfg.max_tokens || St > g.cfg.max_tokens)
        return g.fail(SDVG_ERR_INVALID, "B=%d S_src=%d S_tgt=%d exceed the handle's limits (%d clips, %d tokens)", B, Ss, St,
                      g.cfg.max_clips, g.cfg.max_tokens);
      if (!pe_index && B > 64) return g.fail(SDVG_ERR_BATCH, "B=%d > 64 without pe_index (models/positional_encoding.py:35)", B);
      if (lc.use_mse && lc.use_l1) return g.fail(SDVG_ERR_INVALID, "use_mse and use_l1 together is an invalid loss combination");
      if ((rc = g.check_ready(st)) != SDVG_OK) return rc;
      cudaError_t e = cudaSuccess;
      if (wt_stale) e = repack_weights(false, st);
      if (e == cudaSuccess) e = forward(src, tgt, B, Ss, St, pe_index ? pe_index : g.pe_mod64, st);
      if (e != cudaSuccess) return g.fail_cuda(e, "training forward");
      if ((rc = loss_and_grad(expected, B, St, lc, st)) != SDVG_OK) return rc;
      if (losses) {
        e = cudaMemcpyAsync(losses, loss_out, 5 * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return g.fail_cuda(e, "loss copy");
      }
      if ((e = backward_decoder(st)) != cudaSuccess) return g.fail_cuda(e, "decoder backward");
    }
    if (part != 1) {
      if (Bc == 0) return g.fail(SDVG_ERR_STATE, "part 2 of the backward pass without part 1");
      cudaError_t e = backward_encoder(st);
      if (e != cudaSuccess) return g.fail_cuda(e, "encoder backward");
    }
    return SDVG_OK;
  }

  // ------------------------------------------------------------------ torch.autograd bridge
  // The reference's loop body is `pred = model(...)`, `loss = loss_fn(...)`, `loss.backward()`, `opt.step()`
  // (trainers/trainer.py:141-165) with torch's own criterion and optimiser.  These two calls are the halves of
  // forward_backward() around a loss that lives in the caller: the training-mode forward pass (activations saved,
  // dropout at nn.Transformer's sites) and the backward pass from the upstream gradient dL/dpred.
  int forward_only(const float* src, const float* tgt, int B, int Ss, int St, const int* pe_index, float* pred_out, cudaStream_t st) {
    int rc = init();
    if (rc != SDVG_OK) return rc;
    if (!src || !tgt || !pred_out) return g.fail(SDVG_ERR_INVALID, "null tensor");
    if (B <= 0 || Ss <= 0 || St <= 0 || B > g.cfg.max_clips || Ss > g.cfg.max_tokens || St > g.cfg.max_tokens)
      return g.fail(SDVG_ERR_INVALID, "B=%d S_src=%d S_tgt=%d exceed the handle's limits (%d clips, %d tokens)", B, Ss, St,
                    g.cfg.max_clips, g.cfg.max_tokens);
    if (!pe_index && B > 64) return g.fail(SDVG_ERR_BATCH, "B=%d > 64 without pe_index (models/positional_encoding.py:35)", B);
    if ((rc = g.check_ready(st)) != SDVG_OK) return rc;
    cudaError_t e = cudaSuccess;
    if (wt_stale) e = repack_weights(false, st);
    if (e == cudaSuccess) e = forward(src, tgt, B, Ss, St, pe_index ? pe_index : g.pe_mod64, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(pred_out, pred, static_cast<size_t>(St) * B * g.cfg.latent_dim * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return g.fail_cuda(e, "training forward");
    return SDVG_OK;
  }

  int backward_from(const float* dpred_in, cudaStream_t st) {
    if (!ready || Bc == 0) return g.fail(SDVG_ERR_STATE, "backward without a saved training forward pass");
    if (!dpred_in) return g.fail(SDVG_ERR_INVALID, "null gradient");
    const long long n = static_cast<long long>(Stc) * Bc * g.cfg.latent_dim;
    cudaError_t e = cudaMemcpyAsync(dpred, dpred_in, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) {
      Engine::Scope sc(&g, KC_PACK, 0.0, 4.0 * n, st);
      long long blocks = (n + 255) / 256;
      if (blocks > g.num_sms * 4) blocks = g.num_sms * 4;
      e = launch_kernel(absmax_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, static_cast<const float*>(dpred), n, amax);
    }
    if (e == cudaSuccess) {
      Engine::Scope sc(&g, KC_PACK, 0.0, 16.0, st);
      e = launch_kernel(loss_scale_kernel, dim3(1), dim3(1), 0, st, amax, scale);
    }
    if (e == cudaSuccess) e = backward_decoder(st);
    if (e == cudaSuccess) e = backward_encoder(st);
    if (e != cudaSuccess) return g.fail_cuda(e, "backward from upstream gradient");
    return SDVG_OK;
  }

  // torch.optim.Adam.step() on the flat vectors, then the operand planes of every weight are rebuilt.
  // Ranged form: parameters [offset, offset + count) only (offset a multiple of 64, as announced by the gradient-ready
  // callback), so that the update of the layers whose gradients are final runs next to the rest of the backward pass;
  // `begin_step` advances the step counter of the bias correction (first range of a step).
  int adam_step(float lr, float beta1, float beta2, float eps, float grad_mul, cudaStream_t st, long long offset = 0,
                long long count = -1, bool begin_step = true) {
    if (!ready) return g.fail(SDVG_ERR_STATE, "no gradients: call the backward pass first");
    if (count < 0) count = static_cast<long long>(g.arena_count) - offset;
    if (offset < 0 || offset % 4 != 0 || count % 4 != 0 || offset + count > static_cast<long long>(g.arena_count))
      return g.fail(SDVG_ERR_INVALID, "bad parameter range [%lld, +%lld)", offset, count);
    if (begin_step) ++adam_t;
    if (adam_t == 0) return g.fail(SDVG_ERR_STATE, "ranged Adam step without begin_step");
    AdamArgs a{};
    a.p = g.arena + offset; a.g = grads + offset; a.m = adam_m + offset; a.v = adam_v + offset; a.n = count;
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.gmul = grad_mul;
    const double bc1 = 1.0 - std::pow(static_cast<double>(beta1), static_cast<double>(adam_t));
    const double bc2 = 1.0 - std::pow(static_cast<double>(beta2), static_cast<double>(adam_t));
    a.step_size = static_cast<float>(static_cast<double>(lr) / bc1);
    a.inv_bc2_sqrt = static_cast<float>(1.0 / std::sqrt(bc2));
    cudaError_t e;
    {
      Engine::Scope sc(&g, KC_PACK, 0.0, 28.0 * a.n, st);
      long long blocks = (count / 4 + 255) / 256;
      if (blocks > g.num_sms * 8) blocks = g.num_sms * 8;
      e = launch_kernel(adam_kernel, dim3(static_cast<unsigned>(blocks < 1 ? 1 : blocks)), dim3(256), 0, st, a);
    }
    if (e == cudaSuccess) e = repack_weights(true, st, static_cast<size_t>(offset), static_cast<size_t>(offset + count));
    if (e != cudaSuccess) return g.fail_cuda(e, "Adam step");
    return SDVG_OK;
  }
};

}  // namespace sdvg
