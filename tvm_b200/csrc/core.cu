// Library-wide host state: thread-local error string, launch counter, SM count, split-KV workspace.
#include "common.cuh"

#include <mutex>

namespace tvmb200 {

std::atomic<int64_t> g_launch_count{0};

std::string& last_error_ref() {
  static thread_local std::string err;
  return err;
}

int set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return 1;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------------------------------------------
// Context: everything the reference bakes into a compiled kernel set (rope_scaling, rotary_dim / theta / scale of
// fused_rope, layer_sliding_window_size) plus the device scratch the launches need (split-KV workspace, the work
// counter of the persistent prefill kernel, the block counter of the peer gather).  One DEFAULT context serves the
// plain C ABI and the `__tvm_ffi_*` module symbols; every host cache owns its own; tvmb200_context_create makes more
// (bound tvm-ffi closures, ffi_api.cc).  Scratch is keyed by (device, stream): two streams never share partials or
// counters, so launches of one context on different streams -- and of different contexts -- may run concurrently.
// ---------------------------------------------------------------------------------------------------------------
Context::~Context() {
  for (auto& kv : scratch_) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(static_cast<int>(kv.first.first));
    if (kv.second.ws) cudaFree(kv.second.ws);
    if (kv.second.counters) cudaFree(kv.second.counters);
    cudaSetDevice(prev);
  }
}

Context* default_context() {
  static Context* ctx = new Context();  // never destroyed: the CUDA runtime may be gone at process exit
  return ctx;
}

static thread_local Context* t_ctx = nullptr;
Context* current_context() { return t_ctx ? t_ctx : default_context(); }
Context* enter_context(Context* c) {
  Context* prev = t_ctx;
  t_ctx = c;
  return prev;
}
ContextScope::ContextScope(Context* c) : prev_(t_ctx) { t_ctx = c; }
ContextScope::~ContextScope() { t_ctx = prev_; }

static int scratch_of(Context* c, cudaStream_t st, Context::Scratch** out) {
  int dev = 0;
  TVMB200_CUDA(cudaGetDevice(&dev));
  TVMB200_CHECK(dev >= 0 && dev < 64, "device id %d out of range", dev);
  *out = &c->scratch_[{dev, reinterpret_cast<uintptr_t>(st)}];
  return 0;
}

int get_workspace(int64_t bytes, cudaStream_t st, void** out) {
  Context* c = current_context();
  std::lock_guard<std::mutex> lk(c->mu_);
  Context::Scratch* w = nullptr;
  if (int rc = scratch_of(c, st, &w)) return rc;
  if (w->ws_bytes < bytes) {
    // growing is synchronous (cudaFree / cudaMalloc): never inside a reference callback once the cache has reserved
    // its worst case (tvmb200_reserve_workspace / the host cache's constructor), never inside CUDA-graph capture
    const int64_t want = bytes < (int64_t(32) << 20) ? (int64_t(32) << 20) : bytes + bytes / 4;
    if (w->ws) {
      TVMB200_CUDA(cudaDeviceSynchronize());
      TVMB200_CUDA(cudaFree(w->ws));
      w->ws = nullptr;
      w->ws_bytes = 0;
    }
    TVMB200_CUDA(cudaMalloc(&w->ws, static_cast<size_t>(want)));
    w->ws_bytes = want;
  }
  if (out) *out = w->ws;
  return 0;
}

// two zero-initialised int32 counters per (context, device, stream): [0] the persistent prefill kernel's work queue,
// [64] the peer gather's block ticket (256 bytes apart from each other's cache line)
int get_counters(cudaStream_t st, int32_t** out) {
  Context* c = current_context();
  std::lock_guard<std::mutex> lk(c->mu_);
  Context::Scratch* w = nullptr;
  if (int rc = scratch_of(c, st, &w)) return rc;
  if (!w->counters) {
    TVMB200_CUDA(cudaMalloc(&w->counters, 512));
    TVMB200_CUDA(cudaMemset(w->counters, 0, 512));
  }
  *out = w->counters;
  return 0;
}

int32_t layer_sliding_window_size() { return current_context()->layer_sws.load(); }

RopeScaling rope_scaling() {
  Context* c = current_context();
  std::lock_guard<std::mutex> lk(c->mu_);
  return c->rs;
}
RopeVariant rope_variant() {
  Context* c = current_context();
  std::lock_guard<std::mutex> lk(c->mu_);
  return c->rv;
}
int check_no_rope_variant(const char* who) {
  const int kind = rope_variant().kind;
  if (kind == 0) return 0;
  return set_error("%s: rope scaling kind %d (gptj / llama4 / yarn) is only implemented by split_rotary (rope mode "
                   "\"normal\"); rotations inside the attention kernels need the default or llama3 frequencies", who, kind);
}

int context_set_rope_scaling(Context* c, int32_t kind, float factor, float low_freq_factor, float high_freq_factor,
                             float original_max_position_embeddings) {
  TVMB200_CHECK(kind == TVMB200_ROPE_SCALING_NONE || kind == TVMB200_ROPE_SCALING_LLAMA3 || kind == TVMB200_ROPE_SCALING_GPTJ ||
                    kind == TVMB200_ROPE_SCALING_LLAMA4,
                "set_rope_scaling: kind %d unsupported (0 = none, 1 = llama3, 2 = gptj, 3 = llama4; yarn has its own setter, "
                "longrope is not implemented)", kind);
  RopeScaling rs = {0, 1.0f, 0.0f, 0.0f};
  RopeVariant rv = {0, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (kind == TVMB200_ROPE_SCALING_LLAMA3 || kind == TVMB200_ROPE_SCALING_LLAMA4) {
    const bool equal = high_freq_factor == low_freq_factor;
    TVMB200_CHECK(factor > 0.f && original_max_position_embeddings > 0.f && (kind == TVMB200_ROPE_SCALING_LLAMA4 || !equal) &&
                      (!equal || low_freq_factor != 0.f),
                  "set_rope_scaling: llama3 / llama4 need factor > 0, original_max_position_embeddings > 0 and (llama3) "
                  "high_freq_factor != low_freq_factor");
    // both share the smooth-interpolation constants; llama3 rotates pair-wise (RopeScaling), llama4 element-wise
    const RopeVariant v = make_rope_variant(3, factor, low_freq_factor, high_freq_factor, original_max_position_embeddings);
    if (kind == TVMB200_ROPE_SCALING_LLAMA3)
      rs = {1, v.p0, v.p1, v.p2};
    else
      rv = v;
  } else if (kind == TVMB200_ROPE_SCALING_GPTJ) {
    rv = make_rope_variant(2, 0.f, 0.f, 0.f, 0.f);
  }
  std::lock_guard<std::mutex> lk(c->mu_);
  c->rs = rs;
  c->rv = rv;
  return 0;
}

int context_set_rope_scaling_yarn(Context* c, float factor, float original_max_position_embeddings, float beta_fast,
                                  float beta_slow, float inv_theta_log_scale) {
  TVMB200_CHECK(factor > 0.f && original_max_position_embeddings > 0.f && beta_fast > 0.f && beta_slow > 0.f,
                "set_rope_scaling_yarn: factor, original_max_position_embeddings, beta_fast and beta_slow must be positive");
  std::lock_guard<std::mutex> lk(c->mu_);
  c->rs = {0, 1.0f, 0.0f, 0.0f};
  c->rv = {5, factor, original_max_position_embeddings, beta_fast, beta_slow, inv_theta_log_scale};
  return 0;
}

void context_copy_settings(Context* dst, Context* src) {
  if (dst == src) return;
  RopeScaling rs;
  RopeVariant rv;
  {
    std::lock_guard<std::mutex> lk(src->mu_);
    rs = src->rs;
    rv = src->rv;
  }
  std::lock_guard<std::mutex> lk(dst->mu_);
  dst->rs = rs;
  dst->rv = rv;
  dst->layer_sws.store(src->layer_sws.load());
  dst->rope_theta = src->rope_theta;
  dst->rope_scale = src->rope_scale;
  dst->rotary_dim = src->rotary_dim;
}

}  // namespace tvmb200

extern "C" const char* tvmb200_last_error(void) { return tvmb200::last_error_ref().c_str(); }
extern "C" const char* tvmb200_version(void) { return "tvm_b200 0.1 (sm_100a)"; }
extern "C" int64_t tvmb200_launch_count(void) { return tvmb200::g_launch_count.load(); }
extern "C" void tvmb200_set_layer_sliding_window_size(int32_t size) {
  tvmb200::current_context()->layer_sws.store(size);
}
extern "C" int tvmb200_set_rope_scaling(int32_t kind, float factor, float low_freq_factor, float high_freq_factor,
                                        float original_max_position_embeddings) {
  return tvmb200::context_set_rope_scaling(tvmb200::current_context(), kind, factor, low_freq_factor, high_freq_factor,
                                           original_max_position_embeddings);
}
extern "C" int tvmb200_set_rope_scaling_yarn(float factor, float original_max_position_embeddings, float beta_fast,
                                             float beta_slow, float inv_theta_log_scale) {
  return tvmb200::context_set_rope_scaling_yarn(tvmb200::current_context(), factor, original_max_position_embeddings,
                                                beta_fast, beta_slow, inv_theta_log_scale);
}
extern "C" int32_t tvmb200_get_rope_scaling_kind(void) {
  tvmb200::Context* c = tvmb200::current_context();
  std::lock_guard<std::mutex> lk(c->mu_);
  return c->rv.kind != 0 ? c->rv.kind : c->rs.kind;
}
extern "C" int tvmb200_reserve_workspace(int device_id, int64_t bytes) {
  return tvmb200_reserve_workspace_stream(device_id, bytes, nullptr);
}
extern "C" int tvmb200_reserve_workspace_stream(int device_id, int64_t bytes, tvmb200_stream_t stream) {
  TVMB200_CHECK(device_id >= 0 && device_id < 64, "device id %d out of range", device_id);
  int prev = 0;
  TVMB200_CUDA(cudaGetDevice(&prev));
  TVMB200_CUDA(cudaSetDevice(device_id));
  int rc = tvmb200::get_workspace(bytes, static_cast<cudaStream_t>(stream), nullptr);
  cudaSetDevice(prev);
  return rc;
}

// ---- contexts (include/tvm_b200.h) ----------------------------------------------------------------------------
extern "C" int tvmb200_context_create(tvmb200_context_t* out) {
  TVMB200_CHECK(out != nullptr, "tvmb200_context_create: null argument");
  tvmb200::Context* c = new tvmb200::Context();
  tvmb200::context_copy_settings(c, tvmb200::current_context());
  *out = reinterpret_cast<tvmb200_context_t>(c);
  return 0;
}
extern "C" void tvmb200_context_retain(tvmb200_context_t c) {
  if (c) reinterpret_cast<tvmb200::Context*>(c)->refs.fetch_add(1);
}
extern "C" void tvmb200_context_release(tvmb200_context_t c) {
  tvmb200::Context* ctx = reinterpret_cast<tvmb200::Context*>(c);
  if (ctx && ctx->refs.fetch_sub(1) == 1) delete ctx;
}
extern "C" tvmb200_context_t tvmb200_context_enter(tvmb200_context_t c) {
  tvmb200::Context* prev = tvmb200::enter_context(reinterpret_cast<tvmb200::Context*>(c));
  return reinterpret_cast<tvmb200_context_t>(prev);
}
