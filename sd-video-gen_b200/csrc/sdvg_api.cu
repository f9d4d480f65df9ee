// C ABI of libsdvg.so (declared in include/sdvg.h).  Thin, exception-free shims over sdvg::Engine.
#include <cstdlib>
#include <new>

#include "engine.cuh"
#include "loss.cuh"
#include "train_engine.cuh"

using sdvg::Engine;

struct sdvg_handle {
  Engine eng;
  sdvg::Trainer* trainer = nullptr;   // created by the first training call
  ~sdvg_handle() { delete trainer; }
};

static thread_local std::string g_create_error;

// SDVG_PDL=0 in the environment disables programmatic dependent launch (A/B measurements); default on.
static void read_env_options() {
  const char* v = std::getenv("SDVG_PDL");
  if (v) sdvg::pdl_enabled() = std::atoi(v) != 0;
  if (const char* m = std::getenv("SDVG_ATTN_MMA")) sdvg::attention_mma_enabled() = std::atoi(m) != 0;
}

extern "C" {

int sdvg_version(void) { return SDVG_VERSION; }

const char* sdvg_last_error(const sdvg_handle* h) { return h ? h->eng.err.c_str() : g_create_error.c_str(); }

int sdvg_create(const sdvg_config* cfg, sdvg_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return SDVG_ERR_INVALID; }
  *out = nullptr;
  read_env_options();
  sdvg_handle* h = new (std::nothrow) sdvg_handle();
  if (!h) { g_create_error = "out of host memory"; return SDVG_ERR_INVALID; }
  int r;
  try {
    r = h->eng.init(*cfg);
  } catch (const std::exception& ex) {
    h->eng.err = std::string("exception: ") + ex.what();
    r = SDVG_ERR_INVALID;
  }
  if (r != SDVG_OK) {
    g_create_error = h->eng.err;
    delete h;
    return r;
  }
  *out = h;
  return SDVG_OK;
}

void sdvg_destroy(sdvg_handle* h) {
  if (!h) return;
  cudaSetDevice(h->eng.cfg.device);
  cudaDeviceSynchronize();
  delete h;
}

int sdvg_workspace_bytes(const sdvg_handle* h, size_t* bytes) {
  if (!h || !bytes) return SDVG_ERR_INVALID;
  *bytes = h->eng.bytes_owned;
  return SDVG_OK;
}

int sdvg_set_weight(sdvg_handle* h, const char* key, const void* data, const int64_t* shape, int32_t ndim) {
  if (!h) return SDVG_ERR_INVALID;
  if (!data || !shape) return h->eng.fail(SDVG_ERR_INVALID, "null argument");
  cudaSetDevice(h->eng.cfg.device);
  if (h->trainer) h->trainer->wt_stale = true;
  return h->eng.set_weight(key, data, shape, ndim);
}

int sdvg_num_weights(const sdvg_handle* h) { return h ? static_cast<int>(h->eng.slots.size()) : SDVG_ERR_INVALID; }

const char* sdvg_weight_key(const sdvg_handle* h, int32_t i) {
  if (!h || i < 0 || i >= static_cast<int>(h->eng.slots.size())) return nullptr;
  return h->eng.slots[i].key.c_str();
}

int sdvg_finalize_weights(sdvg_handle* h, void* stream) {
  if (!h) return SDVG_ERR_INVALID;
  cudaSetDevice(h->eng.cfg.device);
  return h->eng.finalize(static_cast<cudaStream_t>(stream));
}

int sdvg_forward(sdvg_handle* h, const float* src, const float* tgt, int32_t B, int32_t S_src, int32_t S_tgt,
                 int32_t mask_kind, const float* mask, const int32_t* pe_index, float* out, void* stream) {
  if (!h) return SDVG_ERR_INVALID;
  cudaSetDevice(h->eng.cfg.device);
  return h->eng.forward(src, tgt, B, S_src, S_tgt, mask_kind, mask, pe_index, out, static_cast<cudaStream_t>(stream));
}

int sdvg_rollout(sdvg_handle* h, const float* ctx, int32_t B, int32_t C, int32_t n_pred, int32_t window,
                 int32_t flags, const float* teacher, const int32_t* pe_index, float scale_in, float scale_out,
                 float* out, void* stream) {
  if (!h) return SDVG_ERR_INVALID;
  cudaSetDevice(h->eng.cfg.device);
  return h->eng.rollout(ctx, B, C, n_pred, window, flags, teacher, pe_index, scale_in, scale_out, out,
                        static_cast<cudaStream_t>(stream));
}

int sdvg_timing_enable(sdvg_handle* h, int32_t on) {
  if (!h) return SDVG_ERR_INVALID;
  h->eng.timing = on != 0;
  return SDVG_OK;
}

int sdvg_timing_read(sdvg_handle* h, double* ms, int64_t* launches, double* flops, double* bytes) {
  if (!h) return SDVG_ERR_INVALID;
  cudaSetDevice(h->eng.cfg.device);
  return h->eng.timing_read(ms, launches, flops, bytes);
}

int64_t sdvg_launch_count(const sdvg_handle* h) { return h ? h->eng.launches : 0; }

// ---------------------------------------------------------------------------------------------------------
// training step
static sdvg::Trainer* trainer_of(sdvg_handle* h) {
  if (!h->trainer) h->trainer = new (std::nothrow) sdvg::Trainer(h->eng);
  return h->trainer;
}

int sdvg_train_backward(sdvg_handle* h, const float* src, const float* tgt, const float* expected, int32_t B, int32_t S_src,
                        int32_t S_tgt, const sdvg_loss_config* loss, const int32_t* pe_index, float* losses, int32_t part,
                        void* stream) {
  if (!h) return SDVG_ERR_INVALID;
  if (!loss) return h->eng.fail(SDVG_ERR_INVALID, "null loss configuration");
  cudaSetDevice(h->eng.cfg.device);
  sdvg::Trainer* t = trainer_of(h);
  if (!t) return h->eng.fail(SDVG_ERR_INVALID, "out of host memory");
  sdvg::TrainLoss lc{loss->frames_to_predict, loss->use_mse != 0, loss->use_l1 != 0, loss->use_gdl != 0, loss->lambda_gdl, loss->alpha,
                     loss->use_contrastive != 0, loss->temperature, loss->lambda_contrastive};
  try {
    return t->forward_backward(src, tgt, expected, B, S_src, S_tgt, lc, pe_index, losses, part, static_cast<cudaStream_t>(stream));
  } catch (const std::exception& ex) {
    return h->eng.fail(SDVG_ERR_INVALID, "exception: %s", ex.what());
  }
}

int sdvg_train_forward(sdvg_handle* h, const float* src, const float* tgt, int32_t B, int32_t S_src, int32_t S_tgt,
                       const int32_t* pe_index, float* pred, void* stream) {
  if (!h) return SDVG_ERR_INVALID;
  cudaSetDevice(h->eng.cfg.device);
  sdvg::Trainer* t = trainer_of(h);
  if (!t) return h->eng.fail(SDVG_ERR_INVALID, "out of host memory");
  try {
    return t->forward_only(src, tgt, B, S_src, S_tgt, pe_index, pred, static_cast<cudaStream_t>(stream));
  } catch (const std::exception& ex) {
    return h->eng.fail(SDVG_ERR_INVALID, "exception: %s", ex.what());
  }
}

int sdvg_train_backward_from(sdvg_handle* h, const float* dpred, void* stream) {
  if (!h) return SDVG_ERR_INVALID;
  cudaSetDevice(h->eng.cfg.device);
  if (!h->trainer) return h->eng.fail(SDVG_ERR_STATE, "backward without a saved training forward pass");
  try {
    return h->trainer->backward_from(dpred, static_cast<cudaStream_t>(stream));
  } catch (const std::exception& ex) {
    return h->eng.fail(SDVG_ERR_INVALID, "exception: %s", ex.what());
  }
}

int sdvg_train_gradients(sdvg_handle* h, float** grads, int64_t* count, int64_t* decoder_offset) {
  if (!h) return SDVG_ERR_INVALID;
  cudaSetDevice(h->eng.cfg.device);
  sdvg::Trainer* t = trainer_of(h);
  if (!t) return h->eng.fail(SDVG_ERR_INVALID, "out of host memory");
  int rc = t->init();
  if (rc != SDVG_OK) return rc;
  if (grads) *grads = t->grads;
  if (count) *count = static_cast<int64_t>(h->eng.arena_count);
  if (decoder_offset) *decoder_offset = static_cast<int64_t>(t->decoder_offset);
  return SDVG_OK;
}

int sdvg_train_set_dropout(sdvg_handle* h, float p, uint64_t seed) {
  if (!h) return SDVG_ERR_INVALID;
  if (!(p >= 0.f) || p >= 1.f) return h->eng.fail(SDVG_ERR_INVALID, "dropout probability must be in [0, 1)");
  sdvg::Trainer* t = trainer_of(h);
  if (!t) return h->eng.fail(SDVG_ERR_INVALID, "out of host memory");
  t->drop_p = p; t->drop_seed = seed;
  return SDVG_OK;
}

int sdvg_train_set_ready_callback(sdvg_handle* h, sdvg_grad_ready_fn fn, void* user, int32_t layers_per_bucket) {
  if (!h) return SDVG_ERR_INVALID;
  sdvg::Trainer* t = trainer_of(h);
  if (!t) return h->eng.fail(SDVG_ERR_INVALID, "out of host memory");
  t->ready_cb = fn; t->ready_user = user;
  t->layers_per_bucket = layers_per_bucket > 0 ? layers_per_bucket : 3;
  return SDVG_OK;
}

int sdvg_param_range(const sdvg_handle* h, const char* key, int64_t* offset, int64_t* count) {
  if (!h) return SDVG_ERR_INVALID;
  auto it = h->eng.slot_of.find(key ? key : "");
  if (it == h->eng.slot_of.end()) return SDVG_ERR_INVALID;
  const sdvg::WeightSlot& s = h->eng.slots[it->second];
  if (offset) *offset = static_cast<int64_t>(s.dev - h->eng.arena);
  if (count) *count = static_cast<int64_t>(s.count);
  return SDVG_OK;
}

int sdvg_train_prediction(sdvg_handle* h, const float** pred) {
  if (!h || !pred) return SDVG_ERR_INVALID;
  if (!h->trainer || !h->trainer->ready || h->trainer->Bc == 0) return h->eng.fail(SDVG_ERR_STATE, "no training forward pass has run");
  *pred = h->trainer->pred;
  return SDVG_OK;
}

int sdvg_train_adam_step(sdvg_handle* h, float lr, float beta1, float beta2, float eps, float grad_mul, void* stream) {
  if (!h) return SDVG_ERR_INVALID;
  cudaSetDevice(h->eng.cfg.device);
  if (!h->trainer) return h->eng.fail(SDVG_ERR_STATE, "no gradients: call sdvg_train_backward first");
  return h->trainer->adam_step(lr, beta1, beta2, eps, grad_mul, static_cast<cudaStream_t>(stream));
}

int sdvg_train_adam_step_range(sdvg_handle* h, float lr, float beta1, float beta2, float eps, float grad_mul, int64_t offset,
                               int64_t count, int32_t begin_step, void* stream) {
  if (!h) return SDVG_ERR_INVALID;
  cudaSetDevice(h->eng.cfg.device);
  if (!h->trainer) return h->eng.fail(SDVG_ERR_STATE, "no gradients: call sdvg_train_backward first");
  return h->trainer->adam_step(lr, beta1, beta2, eps, grad_mul, static_cast<cudaStream_t>(stream), offset, count, begin_step != 0);
}

int sdvg_get_weight(sdvg_handle* h, const char* key, float* out, void* stream) {
  if (!h) return SDVG_ERR_INVALID;
  if (!out) return h->eng.fail(SDVG_ERR_INVALID, "null argument");
  auto it = h->eng.slot_of.find(key ? key : "");
  if (it == h->eng.slot_of.end()) return h->eng.fail(SDVG_ERR_INVALID, "unknown state_dict key '%s'", key ? key : "(null)");
  cudaSetDevice(h->eng.cfg.device);
  const sdvg::WeightSlot& s = h->eng.slots[it->second];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(out, s.dev, s.count * sizeof(float), cudaMemcpyDefault, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return h->eng.fail_cuda(e, "weight read-back");
  return SDVG_OK;
}

// ---------------------------------------------------------------------------------------------------------
int sdvg_gemm(int32_t device, int32_t precision, const float* A, const float* W, const float* bias, int32_t relu,
              float* C, int32_t M, int32_t N, int32_t K, int32_t block_n, int32_t iters, float* ms, void* stream) {
  using namespace sdvg;
  if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || K % 8 != 0) { g_create_error = "sdvg_gemm: bad argument"; return SDVG_ERR_INVALID; }
  if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "sdvg_gemm: no such CUDA device"; return SDVG_ERR_CUDA; }
  read_env_options();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (iters < 1) iters = 1;
  Epilogue e;
  e.bias = bias; e.relu = relu; e.out32 = C; e.ld32 = N;
  cudaEvent_t ev0, ev1;
  cudaEventCreate(&ev0); cudaEventCreate(&ev1);
  cudaError_t err = cudaSuccess;
  int rc = SDVG_OK;
  std::vector<void*> tmp;
  auto cleanup = [&]() { for (void* p : tmp) cudaFree(p); cudaEventDestroy(ev0); cudaEventDestroy(ev1); };
  if (precision == SDVG_FP32_SIMT) {
    cudaEventRecord(ev0, st);
    for (int i = 0; i < iters && err == cudaSuccess; ++i) err = launch_gemm_simt(A, K, W, K, M, N, K, e, st);
    cudaEventRecord(ev1, st);
  } else {
    const bool split = precision == SDVG_FP32, bf = precision == SDVG_BF16;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10 || !get_encode_fn()) { cleanup(); g_create_error = "sdvg_gemm: needs an sm_100 device"; return SDVG_ERR_CUDA; }
    const int Kp = round_up(K, kTcBK), Mp = round_up(M, kTc2BM);
    uint16_t *a_hi = nullptr, *a_lo = nullptr, *w_hi = nullptr, *w_lo = nullptr;
    auto alloc16 = [&](uint16_t** p, size_t n) {
      if (cudaMalloc(reinterpret_cast<void**>(p), n * 2) != cudaSuccess) return false;
      tmp.push_back(*p);
      return cudaMemsetAsync(*p, 0, n * 2, st) == cudaSuccess;
    };
    bool ok = alloc16(&a_hi, size_t(Mp) * Kp) && alloc16(&w_hi, size_t(N) * Kp);
    if (ok && split) ok = alloc16(&a_lo, size_t(Mp) * Kp) && alloc16(&w_lo, size_t(N) * Kp);
    if (!ok) { cleanup(); g_create_error = "sdvg_gemm: out of device memory"; return SDVG_ERR_CUDA; }
    auto pack16 = [&](const float* src, int rows, uint16_t* hi, uint16_t* lo) {
      PackArgs a{};
      a.src = src; a.src_clip_stride = K; a.clips = rows; a.tokens = 1; a.width = K; a.slot[0] = 0; a.scale = 1.f;
      a.out_hi = hi; a.out_lo = lo; a.ld16 = Kp; a.bf16 = bf;
      return launch_pack(a, prop.multiProcessorCount, st);
    };
    err = pack16(A, M, a_hi, a_lo);
    if (err == cudaSuccess) err = pack16(W, N, w_hi, w_lo);
    Engine tmp_eng;
    tmp_eng.num_sms = prop.multiProcessorCount;
    tmp_eng.cfg.precision = precision;
    if (block_n == 9999) {
      // the persistent small-batch kernel (persistent.cuh) running a one-op program: M <= 128
      ActBuf Ab;
      Ab.p.hi = a_hi; Ab.p.lo = a_lo; Ab.p.rows = Mp; Ab.p.cols = K; Ab.p.ld = Kp;
      Linear L;
      L.N = N; L.K = K; L.bias = bias; L.split = split;
      L.p.hi = w_hi; L.p.lo = w_lo; L.p.rows = N; L.p.cols = K; L.p.ld = Kp;
      tmp_eng.use_pk = true; tmp_eng.pk_mode = 1;   // this entry point asks for the persistent kernel explicitly
      if (err != cudaSuccess || !tmp_eng.pk_eligible(M)) { cleanup(); g_create_error = "sdvg_gemm: persistent kernel unavailable (M > 128 or no cluster launch)"; return SDVG_ERR_UNSUPPORTED; }
      tmp_eng.pk_begin();
      int repeat = 1;   // SDVG_PK_REPEAT=n: the same op n times in one program (per-op cost in steady state)
      if (const char* rv = std::getenv("SDVG_PK_REPEAT")) repeat = std::atoi(rv) > 0 ? std::atoi(rv) : 1;
      for (int r = 0; r < repeat; ++r) tmp_eng.gemm(Ab, L, M, e, st);
      Engine::PkProgram* pr = tmp_eng.pk_end({9999}, st);
      if (!pr) { cleanup(); g_create_error = "sdvg_gemm: persistent program rejected"; return SDVG_ERR_UNSUPPORTED; }
      cudaEventRecord(ev0, st);
      for (int i = 0; i < iters && err == cudaSuccess; ++i) err = tmp_eng.pk_launch(*pr, st);
      cudaEventRecord(ev1, st);
      if (err == cudaSuccess) err = cudaStreamSynchronize(st);
      if (err != cudaSuccess) { g_create_error = std::string("sdvg_gemm (persistent): ") + cudaGetErrorString(err); rc = SDVG_ERR_CUDA; }
      else if (ms) { float t = 0.f; cudaEventElapsedTime(&t, ev0, ev1); *ms = t / iters; }
      cleanup();
      return rc;
    }
    // block_n >= 1000 (tests): split-K factor block_n / 1000 with tile width block_n % 1000 (one-CTA kernel, M <= 128)
    int force_ks = 1;
    if (block_n >= 1000) { force_ks = block_n / 1000; block_n %= 1000; }
    TilePlan plan = block_n == 0 ? (std::getenv("SDVG_KSPLIT") && std::atoi(std::getenv("SDVG_KSPLIT")) && M <= kTcBM && K > 256
                                     ? (tmp_eng.use_ksplit = true, tmp_eng.choose_small_m(M, N, K, split)) : tmp_eng.choose_plan(M, N, split, K))
                                 : TilePlan{block_n < 0, block_n < 0 ? -block_n : block_n, force_ks};
    if (plan.ks > 1) {
      void *w = nullptr, *f = nullptr;
      const size_t wbytes = static_cast<size_t>(prop.multiProcessorCount) * kTcBM * 128 * sizeof(float);
      if (cudaMalloc(&w, wbytes) != cudaSuccess || cudaMalloc(&f, 4096) != cudaSuccess) {
        if (w) cudaFree(w);
        cleanup(); g_create_error = "sdvg_gemm: out of device memory"; return SDVG_ERR_CUDA;
      }
      tmp.push_back(w); tmp.push_back(f);
      cudaMemsetAsync(f, 0, 4096, st);
      tmp_eng.ks_ws = static_cast<float*>(w); tmp_eng.ks_flags = static_cast<unsigned int*>(f);
    }
    const int bn = plan.bn;
    const bool ok1 = !plan.pair && (bn == 32 || bn == 64 || bn == 128 || bn == 256) && !(split && bn == 256) && !(bn > N && bn > 32);
    const bool ok2 = plan.pair && (bn == 64 || bn == 128 || bn == 192 || bn == 256) && !(split && bn > 128) && bn / 2 <= N;
    if (!ok1 && !ok2) { cleanup(); g_create_error = "sdvg_gemm: bad block_n"; return SDVG_ERR_INVALID; }
    Planes pa, pb;
    const int box = plan.pair ? bn / 2 : (bn > N ? N : bn);
    const int bi = box_index(plan.pair ? bn / 2 : bn);
    ok = make_tmap_2d(&pb.tm_hi[bi], w_hi, N, Kp, Kp, box, bf);
    if (ok && split) ok = make_tmap_2d(&pb.tm_lo[bi], w_lo, N, Kp, Kp, box, false);
    for (int s = 0; s < kNumBoxes && ok; ++s) {
      ok = make_tmap_2d(&pa.tm_hi[s], a_hi, Mp, Kp, Kp, kActBoxRows[s], bf);
      if (ok && split) ok = make_tmap_2d(&pa.tm_lo[s], a_lo, Mp, Kp, Kp, kActBoxRows[s], false);
    }
    if (!ok) { cleanup(); g_create_error = "sdvg_gemm: cuTensorMapEncodeTiled failed"; return SDVG_ERR_CUDA; }
    TcGemmArgs args{M, N, K, bf ? 1 : 0, 0, kTcBM, 0, nullptr, e};
    if (const char* sv = std::getenv("SDVG_STAGES")) args.max_stages = std::atoi(sv);
    // SDVG_TRACE_BUF=<device pointer to 64 uint64> : pipeline timestamps of CTA 0 (tools/gemm_trace.py)
    if (const char* tb = std::getenv("SDVG_TRACE_BUF")) args.trace = reinterpret_cast<unsigned long long*>(std::strtoull(tb, nullptr, 0));
    cudaEventRecord(ev0, st);
    for (int i = 0; i < iters && err == cudaSuccess; ++i) err = tmp_eng.gemm_tc_dispatch(pa, pb, split, plan, args, st);
    cudaEventRecord(ev1, st);
  }
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) {
    g_create_error = std::string("sdvg_gemm: ") + cudaGetErrorString(err);
    rc = SDVG_ERR_CUDA;
  } else if (ms) {
    float t = 0.f;
    cudaEventElapsedTime(&t, ev0, ev1);
    *ms = t / iters;
  }
  cleanup();
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
int sdvg_criterion(int32_t device, const float* pred, const float* target, int32_t P, int32_t B, int32_t h, int32_t w,
                   int32_t use_mse, int32_t use_l1, int32_t use_gdl, float lambda_gdl, float alpha,
                   int32_t use_contrastive, float temperature, float lambda_contrastive, float* out, void* stream) {
  using namespace sdvg;
  if (!pred || !target || !out || P <= 0 || B <= 0 || h <= 0 || w <= 0 || temperature <= 0.f) {
    g_create_error = "sdvg_criterion: bad argument"; return SDVG_ERR_INVALID;
  }
  if (use_mse && use_l1) {  // trainers/trainer.py:107-109: "Invalid loss function combination"
    g_create_error = "sdvg_criterion: use_mse and use_l1 together is an invalid loss combination in the reference";
    return SDVG_ERR_INVALID;
  }
  if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "sdvg_criterion: no such CUDA device"; return SDVG_ERR_CUDA; }
  read_env_options();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // per-device scratch for the two-stage reductions (grown on demand, never freed: a few hundred KB)
  static double* scratch[64] = {};
  static size_t scratch_elems[64] = {};
  const size_t need = static_cast<size_t>(kLossBlocks) * 4 + static_cast<size_t>(P) * B * 2;
  const int di = device & 63;
  if (scratch_elems[di] < need) {
    if (scratch[di]) { cudaStreamSynchronize(st); cudaFree(scratch[di]); }
    if (cudaMalloc(reinterpret_cast<void**>(&scratch[di]), need * sizeof(double)) != cudaSuccess) {
      scratch[di] = nullptr; scratch_elems[di] = 0;
      g_create_error = "sdvg_criterion: out of device memory"; return SDVG_ERR_CUDA;
    }
    scratch_elems[di] = need;
  }
  const int hw = h * w;
  if (use_contrastive && static_cast<size_t>(hw) * 32 > 200 * 1024) {
    g_create_error = "sdvg_criterion: feature map too large for the contrastive kernel"; return SDVG_ERR_UNSUPPORTED;
  }
  LossArgs la{};
  la.x = pred; la.y = target; la.P = P; la.B = B; la.h = h; la.w = w; la.alpha = alpha;
  la.inv_temperature = 1.0f / temperature;
  la.partial = scratch[di]; la.nce_partial = scratch[di] + static_cast<size_t>(kLossBlocks) * 4;
  const long long total = static_cast<long long>(P) * B * 4 * hw;
  int blocks = static_cast<int>((total + kLossThreads - 1) / kLossThreads);
  if (blocks > kLossBlocks) blocks = kLossBlocks;
  cudaError_t err = launch_kernel(loss_elementwise_kernel, dim3(blocks), dim3(kLossThreads), 0, st, la);
  if (err == cudaSuccess && use_contrastive) {
    const size_t smem = static_cast<size_t>(hw) * 2 * sizeof(float4);
    static bool attr_set[64] = {};
    if (smem > 48 * 1024 && !attr_set[di]) {
      err = cudaFuncSetAttribute(loss_nce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set[di] = err == cudaSuccess;
    }
    const int threads = hw >= 256 ? 256 : (hw >= 128 ? 128 : 64);
    if (err == cudaSuccess) err = launch_kernel(loss_nce_kernel, dim3(P * B), dim3(threads), smem, st, la);
  }
  if (err == cudaSuccess) {
    LossFinalArgs fa{};
    fa.partial = la.partial; fa.n_blocks = blocks; fa.nce_partial = la.nce_partial; fa.n_ct = use_contrastive ? P * B : 0;
    fa.numel = static_cast<double>(total); fa.nce_rows = static_cast<double>(P) * B * hw;
    fa.use_mse = use_mse != 0; fa.use_l1 = use_l1 != 0; fa.use_gdl = use_gdl != 0; fa.use_nce = use_contrastive != 0;
    fa.lambda_gdl = lambda_gdl; fa.lambda_nce = lambda_contrastive; fa.out = out;
    err = launch_kernel(loss_finalize_kernel, dim3(1), dim3(32), 0, st, fa);
  }
  if (err != cudaSuccess) { g_create_error = std::string("sdvg_criterion: ") + cudaGetErrorString(err); return SDVG_ERR_CUDA; }
  return SDVG_OK;
}

}  // extern "C"
