// Library-wide host state: thread-local error string, launch counter, SM count, split-KV workspace.
#include "common.cuh"

#include <mutex>

namespace tvmb200 {

std::atomic<int64_t> g_launch_count{0};
static std::atomic<int32_t> g_layer_sliding_window_size{1024};

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

struct Workspace {
  void* ptr = nullptr;
  int64_t bytes = 0;
};
static Workspace g_ws[64];
static std::mutex g_ws_mu;

static int ensure_workspace(int dev, int64_t bytes, void** out) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  Workspace& w = g_ws[dev];
  if (w.bytes < bytes) {
    // growing is synchronous (cudaFree/cudaMalloc); callers that capture CUDA graphs pre-size the
    // workspace with tvmb200_reserve_workspace.
    int64_t want = bytes < (int64_t(8) << 20) ? (int64_t(8) << 20) : bytes + bytes / 4;
    if (w.ptr) {
      TVMB200_CUDA(cudaDeviceSynchronize());
      TVMB200_CUDA(cudaFree(w.ptr));
      w.ptr = nullptr;
      w.bytes = 0;
    }
    TVMB200_CUDA(cudaMalloc(&w.ptr, static_cast<size_t>(want)));
    w.bytes = want;
  }
  if (out) *out = w.ptr;
  return 0;
}

int get_workspace(int64_t bytes, void** out) {
  int dev = 0;
  TVMB200_CUDA(cudaGetDevice(&dev));
  TVMB200_CHECK(dev >= 0 && dev < 64, "device id %d out of range", dev);
  return ensure_workspace(dev, bytes, out);
}

int32_t layer_sliding_window_size() { return g_layer_sliding_window_size.load(); }

static std::mutex g_rope_mu;
static RopeScaling g_rope_scaling = {0, 1.0f, 0.0f, 0.0f};
RopeScaling rope_scaling() {
  std::lock_guard<std::mutex> lk(g_rope_mu);
  return g_rope_scaling;
}

static RopeVariant g_rope_variant = {0, 0.f, 0.f, 0.f, 0.f, 0.f};
RopeVariant rope_variant() {
  std::lock_guard<std::mutex> lk(g_rope_mu);
  return g_rope_variant;
}
int check_no_rope_variant(const char* who) {
  const int kind = rope_variant().kind;
  if (kind == 0) return 0;
  return set_error("%s: rope scaling kind %d (gptj / llama4 / yarn) is only implemented by split_rotary (rope mode "
                   "\"normal\"); rotations inside the attention kernels need the default or llama3 frequencies", who, kind);
}

}  // namespace tvmb200

extern "C" const char* tvmb200_last_error(void) { return tvmb200::last_error_ref().c_str(); }
extern "C" const char* tvmb200_version(void) { return "tvm_b200 0.1 (sm_100a)"; }
extern "C" int64_t tvmb200_launch_count(void) { return tvmb200::g_launch_count.load(); }
extern "C" void tvmb200_set_layer_sliding_window_size(int32_t size) {
  tvmb200::g_layer_sliding_window_size.store(size);
}
extern "C" int tvmb200_set_rope_scaling(int32_t kind, float factor, float low_freq_factor, float high_freq_factor,
                                        float original_max_position_embeddings) {
  TVMB200_CHECK(kind == TVMB200_ROPE_SCALING_NONE || kind == TVMB200_ROPE_SCALING_LLAMA3 || kind == TVMB200_ROPE_SCALING_GPTJ ||
                    kind == TVMB200_ROPE_SCALING_LLAMA4,
                "set_rope_scaling: kind %d unsupported (0 = none, 1 = llama3, 2 = gptj, 3 = llama4; yarn has its own setter, "
                "longrope is not implemented)", kind);
  tvmb200::RopeScaling rs = {0, 1.0f, 0.0f, 0.0f};
  tvmb200::RopeVariant rv = {0, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (kind == TVMB200_ROPE_SCALING_LLAMA3 || kind == TVMB200_ROPE_SCALING_LLAMA4) {
    const bool equal = high_freq_factor == low_freq_factor;
    TVMB200_CHECK(factor > 0.f && original_max_position_embeddings > 0.f && (kind == TVMB200_ROPE_SCALING_LLAMA4 || !equal) &&
                      (!equal || low_freq_factor != 0.f),
                  "set_rope_scaling: llama3 / llama4 need factor > 0, original_max_position_embeddings > 0 and (llama3) "
                  "high_freq_factor != low_freq_factor");
    // both share the smooth-interpolation constants; llama3 rotates pair-wise (RopeScaling), llama4 element-wise
    const tvmb200::RopeVariant v = tvmb200::make_rope_variant(3, factor, low_freq_factor, high_freq_factor,
                                                              original_max_position_embeddings);
    if (kind == TVMB200_ROPE_SCALING_LLAMA3)
      rs = {1, v.p0, v.p1, v.p2};
    else
      rv = v;
  } else if (kind == TVMB200_ROPE_SCALING_GPTJ) {
    rv = tvmb200::make_rope_variant(2, 0.f, 0.f, 0.f, 0.f);
  }
  std::lock_guard<std::mutex> lk(tvmb200::g_rope_mu);
  tvmb200::g_rope_scaling = rs;
  tvmb200::g_rope_variant = rv;
  return 0;
}
extern "C" int tvmb200_set_rope_scaling_yarn(float factor, float original_max_position_embeddings, float beta_fast,
                                             float beta_slow, float inv_theta_log_scale) {
  TVMB200_CHECK(factor > 0.f && original_max_position_embeddings > 0.f && beta_fast > 0.f && beta_slow > 0.f,
                "set_rope_scaling_yarn: factor, original_max_position_embeddings, beta_fast and beta_slow must be positive");
  std::lock_guard<std::mutex> lk(tvmb200::g_rope_mu);
  tvmb200::g_rope_scaling = {0, 1.0f, 0.0f, 0.0f};
  tvmb200::g_rope_variant = {5, factor, original_max_position_embeddings, beta_fast, beta_slow, inv_theta_log_scale};
  return 0;
}
extern "C" int32_t tvmb200_get_rope_scaling_kind(void) {
  std::lock_guard<std::mutex> lk(tvmb200::g_rope_mu);
  return tvmb200::g_rope_variant.kind != 0 ? tvmb200::g_rope_variant.kind : tvmb200::g_rope_scaling.kind;
}
extern "C" int tvmb200_reserve_workspace(int device_id, int64_t bytes) {
  TVMB200_CHECK(device_id >= 0 && device_id < 64, "device id %d out of range", device_id);
  int prev = 0;
  TVMB200_CUDA(cudaGetDevice(&prev));
  TVMB200_CUDA(cudaSetDevice(device_id));
  int rc = tvmb200::ensure_workspace(device_id, bytes, nullptr);
  cudaSetDevice(prev);
  return rc;
}
