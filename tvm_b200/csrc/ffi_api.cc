// tvm-ffi packed-function layer: every callback of the reference's PagedKVCache is exported as a
// `__tvm_ffi_<name>` symbol with the TVMFFISafeCallType ABI and the reference's positional argument
// order (src/runtime/vm/attn_backend.h:234-243, 388-396, 507-515, 618-627, 665-673 and
// paged_kv_cache.cc:1360-1373, 728, 759, 1718, 2292), then forwarded to the plain C ABI of
// include/tvm_b200.h.  Only the tvm-ffi *C* API is used (no tvm::ffi C++ templates), so the same
// .so loads under pip tvm_ffi 0.1.9 and the reference's vendored 0.1.14 (call-ABI type indices
// are identical).  libtvm_ffi symbols are resolved lazily with dlsym so the library also loads
// in a process that only uses the plain C ABI.
//
// Stream: kernels launch on TVMFFIEnvGetStream(kDLCUDA, device_id), the stream the reference cache
// switches with DeviceAPI::SetStream (paged_kv_cache.cc:724,755,2352).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <dlpack/dlpack.h>
#include <tvm/ffi/c_api.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <iterator>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/tvm_b200.h"
#include "../../include/tvm_b200_cache.h"
#include "common.cuh"

namespace {

typedef void* (*PFN_GetStream)(int32_t, int32_t);
typedef void (*PFN_SetRaised)(const char*, const char*);

void* ffi_sym(const char* name) {
  void* s = dlsym(RTLD_DEFAULT, name);
  if (s) return s;
  void* h = dlopen("libtvm_ffi.so", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libtvm_ffi.so", RTLD_NOW | RTLD_GLOBAL);
  return h ? dlsym(h, name) : nullptr;
}

void* env_stream(int device_id) {
  static PFN_GetStream fn = reinterpret_cast<PFN_GetStream>(ffi_sym("TVMFFIEnvGetStream"));
  return fn ? fn(kDLCUDA, device_id) : nullptr;
}

int raise(const char* kind, const std::string& msg) {
  static PFN_SetRaised fn = reinterpret_cast<PFN_SetRaised>(ffi_sym("TVMFFIErrorSetRaisedFromCStr"));
  if (fn) fn(kind, msg.c_str());
  return -1;
}

std::string fmt(const char* f, ...) {
  char buf[768];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return buf;
}

struct Err {
  std::string kind, msg;
};

struct Tensor {
  void* data = nullptr;
  const DLTensor* t = nullptr;
  int ndim() const { return t->ndim; }
  int64_t shape(int i) const { return t->shape[i]; }
};

// ---- argument decoding (throws Err) -----------------------------------------------------------------
Tensor arg_tensor(const TVMFFIAny* args, int i, const char* fn, const char* name) {
  const TVMFFIAny& a = args[i];
  const DLTensor* t = nullptr;
  if (a.type_index == kTVMFFIDLTensorPtr) {
    t = static_cast<const DLTensor*>(a.v_ptr);
  } else if (a.type_index == kTVMFFITensor) {
    t = TVMFFITensorGetDLTensorPtr(a.v_obj);
  } else {
    throw Err{"TypeError", fmt("%s: argument %d (%s) must be a Tensor, got type index %d", fn, i, name, a.type_index)};
  }
  if (t->strides != nullptr) {
    int64_t expect = 1;
    for (int d = t->ndim - 1; d >= 0; --d) {
      if (t->shape[d] != 1 && t->strides[d] != expect)
        throw Err{"ValueError", fmt("%s: argument %d (%s) must be compact row-major", fn, i, name)};
      expect *= t->shape[d];
    }
  }
  Tensor r;
  r.t = t;
  r.data = static_cast<char*>(t->data) + t->byte_offset;
  return r;
}

int64_t arg_int(const TVMFFIAny* args, int i, const char* fn, const char* name) {
  const TVMFFIAny& a = args[i];
  if (a.type_index == kTVMFFIInt || a.type_index == kTVMFFIBool) return a.v_int64;
  throw Err{"TypeError", fmt("%s: argument %d (%s) must be an int, got type index %d", fn, i, name, a.type_index)};
}

// the TIR binder also accepts an int for a float parameter (tvm_ffi_binder.cc:488-534)
double arg_float(const TVMFFIAny* args, int i, const char* fn, const char* name) {
  const TVMFFIAny& a = args[i];
  if (a.type_index == kTVMFFIFloat) return a.v_float64;
  if (a.type_index == kTVMFFIInt || a.type_index == kTVMFFIBool) return static_cast<double>(a.v_int64);
  throw Err{"TypeError", fmt("%s: argument %d (%s) must be a float, got type index %d", fn, i, name, a.type_index)};
}

std::string arg_str(const TVMFFIAny* args, int i, const char* fn, const char* name) {
  const TVMFFIAny& a = args[i];
  if (a.type_index == kTVMFFIRawStr) return a.v_c_str;
  if (a.type_index == kTVMFFISmallStr) return std::string(a.v_bytes, a.small_str_len);
  if (a.type_index == kTVMFFIStr) {
    const TVMFFIByteArray* b = TVMFFIBytesGetByteArrayPtr(a.v_obj);
    return std::string(b->data, b->size);
  }
  throw Err{"TypeError", fmt("%s: argument %d (%s) must be a string, got type index %d", fn, i, name, a.type_index)};
}

void expect_nargs(int got, int want, const char* fn) {
  if (got != want) throw Err{"TypeError", fmt("%s expects %d arguments, got %d", fn, want, got)};
}

int kv_dtype(const Tensor& x, const char* fn, const char* name) {
  const DLDataType d = x.t->dtype;
  if (d.lanes == 1 && d.bits == 16 && d.code == kDLFloat) return TVMB200_F16;
  if (d.lanes == 1 && d.bits == 16 && d.code == kDLBfloat) return TVMB200_BF16;
  throw Err{"ValueError", fmt("%s: %s has dtype (code %d, bits %d); the sm_100a kernels support float16 and bfloat16", fn, name, d.code, d.bits)};
}
void expect_dtype(const Tensor& x, int code, int bits, const char* fn, const char* name, const char* what) {
  const DLDataType d = x.t->dtype;
  if (!(d.lanes == 1 && d.bits == bits && d.code == code))
    throw Err{"ValueError", fmt("%s: %s must be %s", fn, name, what)};
}
void expect_i32(const Tensor& x, const char* fn, const char* name) { expect_dtype(x, kDLInt, 32, fn, name, "int32"); }
void expect_f32(const Tensor& x, const char* fn, const char* name) { expect_dtype(x, kDLFloat, 32, fn, name, "float32"); }
void expect_ndim(const Tensor& x, int nd, const char* fn, const char* name) {
  if (x.ndim() != nd) throw Err{"ValueError", fmt("%s: %s.ndim is expected to equal %d, got %d", fn, name, nd, x.ndim())};
}
void expect_shape(bool ok, const char* fn, const char* what) {
  if (!ok) throw Err{"ValueError", fmt("%s: shape mismatch: %s", fn, what)};
}
void expect_same_dtype(const Tensor& a, const Tensor& b, const char* fn, const char* what) {
  if (a.t->dtype.code != b.t->dtype.code || a.t->dtype.bits != b.t->dtype.bits)
    throw Err{"ValueError", fmt("%s: dtype mismatch: %s", fn, what)};
}

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  int dev;
  explicit DeviceGuard(const Tensor& x, const char* fn) {
    const DLDevice d = x.t->device;
    if (d.device_type != kDLCUDA && d.device_type != kDLCUDAManaged)
      throw Err{"ValueError", fmt("%s: tensors must live on a CUDA device (device_type %d); there is no CPU fallback", fn, d.device_type)};
    dev = d.device_id;
    if (cudaGetDevice(&prev) != cudaSuccess)
      throw Err{"RuntimeError", fmt("%s: no usable CUDA device: %s", fn, cudaGetErrorString(cudaGetLastError()))};
    if (prev != dev) {
      cudaSetDevice(dev);
      switched = true;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
void expect_same_device(const Tensor& a, const Tensor& b, const char* fn, const char* name) {
  if (a.t->device.device_type != b.t->device.device_type || a.t->device.device_id != b.t->device.device_id)
    throw Err{"ValueError", fmt("%s: %s is on a different device", fn, name)};
}

void check_rc(int rc) {
  if (rc != 0) throw Err{"RuntimeError", tvmb200_last_error()};
}

struct PagesInfo {
  int64_t num_pages;
  int hkv, page_size, d, dtype;
};
PagesInfo pages_info(const Tensor& pages, const char* fn) {
  expect_ndim(pages, 5, fn, "pages");
  expect_shape(pages.shape(1) == 2, fn, "pages.shape[1] must be 2");
  PagesInfo pi;
  pi.num_pages = pages.shape(0);
  pi.hkv = static_cast<int>(pages.shape(2));
  pi.page_size = static_cast<int>(pages.shape(3));
  pi.d = static_cast<int>(pages.shape(4));
  pi.dtype = kv_dtype(pages, fn, "pages");
  return pi;
}

#define TVMB200_FFI_BEGIN() try {
#define TVMB200_FFI_END()                                   \
    if (result) { result->type_index = kTVMFFINone; result->zero_padding = 0; result->v_int64 = 0; } \
    return 0;                                               \
  } catch (const Err& e) {                                  \
    return raise(e.kind.c_str(), e.msg);                    \
  } catch (const std::exception& e) {                       \
    return raise("RuntimeError", e.what());                 \
  }

// ---- the callbacks --------------------------------------------------------------------------------

// (pages, k_data, v_data, position_map)
int impl_transpose_append(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "tir_kv_cache_transpose_append";
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 4, fn);
  Tensor pages = arg_tensor(args, 0, fn, "pages"), k = arg_tensor(args, 1, fn, "k_data"),
         v = arg_tensor(args, 2, fn, "v_data"), pm = arg_tensor(args, 3, fn, "position_map");
  PagesInfo pi = pages_info(pages, fn);
  expect_ndim(k, 3, fn, "k_data");
  expect_ndim(v, 3, fn, "v_data");
  expect_ndim(pm, 1, fn, "position_map");
  expect_i32(pm, fn, "position_map");
  expect_same_dtype(pages, k, fn, "k_data vs pages");
  expect_same_dtype(pages, v, fn, "v_data vs pages");
  const int64_t nt = k.shape(0);
  expect_shape(k.shape(1) == pi.hkv && k.shape(2) == pi.d, fn, "k_data must be [ntoken, num_kv_heads, head_dim]");
  expect_shape(v.shape(0) == nt && v.shape(1) == pi.hkv && v.shape(2) == pi.d, fn, "v_data must match k_data");
  expect_shape(pm.shape(0) == nt, fn, "position_map must be [ntoken]");
  DeviceGuard g(pages, fn);
  expect_same_device(pages, k, fn, "k_data");
  expect_same_device(pages, v, fn, "v_data");
  expect_same_device(pages, pm, fn, "position_map");
  check_rc(tvmb200_transpose_append(pages.data, k.data, v.data, static_cast<const int32_t*>(pm.data), nt,
                                    pi.num_pages, pi.hkv, pi.page_size, pi.d, pi.dtype, env_stream(g.dev)));
  TVMB200_FFI_END();
}

// (pages, position_map, k_data, v_data, layer_id)
int impl_debug_get_kv(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "tir_kv_cache_debug_get_kv";
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 5, fn);
  Tensor pages = arg_tensor(args, 0, fn, "pages"), pm = arg_tensor(args, 1, fn, "position_map"),
         k = arg_tensor(args, 2, fn, "k_data"), v = arg_tensor(args, 3, fn, "v_data");
  const int64_t layer_id = arg_int(args, 4, fn, "layer_id");
  PagesInfo pi = pages_info(pages, fn);
  expect_ndim(k, 4, fn, "k_data");
  expect_ndim(v, 4, fn, "v_data");
  expect_ndim(pm, 1, fn, "position_map");
  expect_i32(pm, fn, "position_map");
  expect_same_dtype(pages, k, fn, "k_data vs pages");
  expect_same_dtype(pages, v, fn, "v_data vs pages");
  const int64_t layers = k.shape(0), seqlen = k.shape(1);
  expect_shape(k.shape(2) == pi.hkv && k.shape(3) == pi.d, fn, "k_data must be [layers, seqlen, num_kv_heads, head_dim]");
  expect_shape(v.shape(0) == layers && v.shape(1) == seqlen && v.shape(2) == pi.hkv && v.shape(3) == pi.d, fn, "v_data must match k_data");
  expect_shape(pm.shape(0) == seqlen, fn, "position_map must be [seqlen]");
  DeviceGuard g(pages, fn);
  check_rc(tvmb200_debug_get_kv(pages.data, static_cast<const int32_t*>(pm.data), k.data, v.data, layer_id, layers,
                                seqlen, pi.num_pages, pi.hkv, pi.page_size, pi.d, pi.dtype, env_stream(g.dev)));
  TVMB200_FFI_END();
}

// (pages, src_page_id, tgt_page_id, copy_length)
int impl_copy_single_page(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "copy_single_page";
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 4, fn);
  Tensor pages = arg_tensor(args, 0, fn, "pages");
  const int64_t src = arg_int(args, 1, fn, "src_page_id"), tgt = arg_int(args, 2, fn, "tgt_page_id"),
                len = arg_int(args, 3, fn, "copy_length");
  PagesInfo pi = pages_info(pages, fn);
  DeviceGuard g(pages, fn);
  check_rc(tvmb200_copy_single_page(pages.data, src, tgt, len, pi.num_pages, pi.hkv, pi.page_size, pi.d, pi.dtype,
                                    env_stream(g.dev)));
  TVMB200_FFI_END();
}

// (pages, copy_length_indptr, copy_src_dst_pos, batch_size)
int impl_compact_kv_copy(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "compact_kv_copy";
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 4, fn);
  Tensor pages = arg_tensor(args, 0, fn, "pages"), ip = arg_tensor(args, 1, fn, "copy_length_indptr"),
         sd = arg_tensor(args, 2, fn, "copy_src_dst_pos");
  const int64_t batch = arg_int(args, 3, fn, "batch_size");
  PagesInfo pi = pages_info(pages, fn);
  expect_ndim(ip, 1, fn, "copy_length_indptr");
  expect_ndim(sd, 2, fn, "copy_src_dst_pos");
  expect_i32(ip, fn, "copy_length_indptr");
  expect_i32(sd, fn, "copy_src_dst_pos");
  expect_shape(ip.shape(0) == batch + 1, fn, "copy_length_indptr must be [batch_size + 1]");
  expect_shape(sd.shape(0) == 2, fn, "copy_src_dst_pos must be [2, total_copy_length]");
  DeviceGuard g(pages, fn);
  check_rc(tvmb200_compact_kv_copy(pages.data, static_cast<const int32_t*>(ip.data),
                                   static_cast<const int32_t*>(sd.data), static_cast<int32_t>(batch),
                                   static_cast<int32_t>(sd.shape(1)), pi.num_pages, pi.hkv, pi.page_size, pi.d,
                                   pi.dtype, env_stream(g.dev)));
  TVMB200_FFI_END();
}

// (qkv, position_map, q, k, v, apply_rope)     rope theta/scale/rotary_dim are context state (see set_rope_params)
int impl_fused_rope(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "fused_rope";
  TVMB200_FFI_BEGIN();
  // 6 arguments = the reference signature (theta/scale/rotary_dim from module state);
  // 9 arguments = explicit (..., apply_rope, rope_theta, rope_scale, rotary_dim) used by our own host
  if (n != 6 && n != 9) throw Err{"TypeError", fmt("%s expects 6 (or 9) arguments, got %d", fn, n)};
  tvmb200::Context* ctx = tvmb200::current_context();
  float theta = ctx->rope_theta, scale = ctx->rope_scale;
  int rotary_dim = ctx->rotary_dim;
  if (n == 9) {
    theta = static_cast<float>(arg_float(args, 6, fn, "rope_theta"));
    scale = static_cast<float>(arg_float(args, 7, fn, "rope_scale"));
    rotary_dim = static_cast<int>(arg_int(args, 8, fn, "rotary_dim"));
  }
  Tensor qkv = arg_tensor(args, 0, fn, "qkv"), pm = arg_tensor(args, 1, fn, "position_map"),
         q = arg_tensor(args, 2, fn, "q"), k = arg_tensor(args, 3, fn, "k"), v = arg_tensor(args, 4, fn, "v");
  // the longrope flavour of the reference PrimFunc takes `ext_factors` (a tensor) here (position_embedding.py:567-667)
  if (args[5].type_index == kTVMFFITensor || args[5].type_index == kTVMFFIDLTensorPtr)
    throw Err{"ValueError", fmt("%s: rope_ext_factors (longrope, fused_rope_longrope_scaling) is not supported", fn)};
  const int64_t apply_rope = arg_int(args, 5, fn, "apply_rope");
  expect_ndim(qkv, 3, fn, "qkv");
  expect_ndim(q, 3, fn, "q");
  expect_ndim(k, 3, fn, "k");
  expect_ndim(v, 3, fn, "v");
  expect_ndim(pm, 1, fn, "position_map");
  expect_i32(pm, fn, "position_map");
  const int dtype = kv_dtype(qkv, fn, "qkv");
  expect_same_dtype(qkv, q, fn, "q vs qkv");
  expect_same_dtype(qkv, k, fn, "k vs qkv");
  expect_same_dtype(qkv, v, fn, "v vs qkv");
  const int64_t nt = qkv.shape(0);
  const int hq = static_cast<int>(q.shape(1)), hkv = static_cast<int>(k.shape(1)), d = static_cast<int>(qkv.shape(2));
  expect_shape(qkv.shape(1) == hq + 2 * hkv, fn, "qkv.shape[1] must equal num_q_heads + 2*num_kv_heads");
  expect_shape(q.shape(0) == nt && q.shape(2) == d, fn, "q must be [seq_len, num_q_heads, head_dim]");
  expect_shape(k.shape(0) == nt && k.shape(2) == d, fn, "k must be [seq_len, num_kv_heads, head_dim]");
  expect_shape(v.shape(0) == nt && v.shape(1) == hkv && v.shape(2) == d, fn, "v must be [seq_len, num_kv_heads, head_dim]");
  expect_shape(pm.shape(0) == nt, fn, "position_map must be [seq_len]");
  DeviceGuard g(qkv, fn);
  check_rc(tvmb200_split_rotary(qkv.data, static_cast<const int32_t*>(pm.data), q.data, k.data, v.data, nt, hq, hkv,
                                d, rotary_dim, apply_rope, scale, theta, dtype, env_stream(g.dev)));
  TVMB200_FFI_END();
}

// (theta, scale[, rotary_dim]) -- the reference bakes these into the fused_rope PrimFunc when it is
// built (position_embedding.py:444-452); a loaded .so needs them as state of the (current / bound) context.
int impl_set_rope_params(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "set_rope_params";
  TVMB200_FFI_BEGIN();
  if (n != 2 && n != 3) throw Err{"TypeError", "set_rope_params expects (theta, scale[, rotary_dim])"};
  tvmb200::Context* ctx = tvmb200::current_context();
  ctx->rope_theta = static_cast<float>(arg_float(args, 0, fn, "theta"));
  ctx->rope_scale = static_cast<float>(arg_float(args, 1, fn, "scale"));
  ctx->rotary_dim = n == 3 ? static_cast<int>(arg_int(args, 2, fn, "rotary_dim")) : 0;
  TVMB200_FFI_END();
}

// (kind, factor, low_freq_factor, high_freq_factor, original_max_position_embeddings): the rope_scaling dict the
// reference bakes into its PrimFuncs (position_embedding.py:257-299); kind 0 = default, 1 = llama3, 2 = gptj, 3 = llama4
int impl_set_rope_scaling(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "set_rope_scaling";
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 5, fn);
  if (tvmb200_set_rope_scaling(static_cast<int32_t>(arg_int(args, 0, fn, "kind")), static_cast<float>(arg_float(args, 1, fn, "factor")),
                               static_cast<float>(arg_float(args, 2, fn, "low_freq_factor")),
                               static_cast<float>(arg_float(args, 3, fn, "high_freq_factor")),
                               static_cast<float>(arg_float(args, 4, fn, "original_max_position_embeddings"))) != 0)
    throw Err{"ValueError", tvmb200_last_error()};
  TVMB200_FFI_END();
}

// (factor, original_max_position_embeddings, beta_fast, beta_slow[, inv_theta_log_scale]): rope_scaling = yarn
int impl_set_rope_scaling_yarn(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "set_rope_scaling_yarn";
  TVMB200_FFI_BEGIN();
  if (n != 4 && n != 5) throw Err{"TypeError", "set_rope_scaling_yarn expects (factor, original_max_position_embeddings, beta_fast, beta_slow[, inv_theta_log_scale])"};
  if (tvmb200_set_rope_scaling_yarn(static_cast<float>(arg_float(args, 0, fn, "factor")),
                                    static_cast<float>(arg_float(args, 1, fn, "original_max_position_embeddings")),
                                    static_cast<float>(arg_float(args, 2, fn, "beta_fast")),
                                    static_cast<float>(arg_float(args, 3, fn, "beta_slow")),
                                    n == 5 ? static_cast<float>(arg_float(args, 4, fn, "inv_theta_log_scale")) : 0.f) != 0)
    throw Err{"ValueError", tvmb200_last_error()};
  TVMB200_FFI_END();
}

// (layer_sliding_window_size)
int impl_set_layer_sliding_window_size(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "set_layer_sliding_window_size";
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 1, fn);
  tvmb200_set_layer_sliding_window_size(static_cast<int32_t>(arg_int(args, 0, fn, "size")));
  TVMB200_FFI_END();
}

// (v, s, v_other, s_other)
int impl_merge_state_inplace(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "merge_state_inplace";
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 4, fn);
  Tensor v = arg_tensor(args, 0, fn, "v"), s = arg_tensor(args, 1, fn, "s"), vo = arg_tensor(args, 2, fn, "v_other"),
         so = arg_tensor(args, 3, fn, "s_other");
  expect_ndim(v, 3, fn, "v");
  expect_ndim(s, 2, fn, "s");
  expect_ndim(vo, 3, fn, "v_other");
  expect_ndim(so, 2, fn, "s_other");
  const int dtype = kv_dtype(v, fn, "v");
  expect_same_dtype(v, vo, fn, "v_other vs v");
  expect_f32(s, fn, "s");
  expect_f32(so, fn, "s_other");
  const int64_t N = v.shape(0);
  const int H = static_cast<int>(v.shape(1)), D = static_cast<int>(v.shape(2));
  expect_shape(vo.shape(0) == N && vo.shape(1) == H && vo.shape(2) == D, fn, "v_other must match v");
  expect_shape(s.shape(0) == N && s.shape(1) == H && so.shape(0) == N && so.shape(1) == H, fn, "s, s_other must be [N, H]");
  DeviceGuard g(v, fn);
  check_rc(tvmb200_merge_state_inplace(v.data, static_cast<float*>(s.data), vo.data, static_cast<const float*>(so.data),
                                       N, H, D, dtype, env_stream(g.dev)));
  TVMB200_FFI_END();
}

struct PagedArgs {
  Tensor pages, page_indptr, page_values, length_info, k_rope_pos_offset;
  PagesInfo pi;
  int batch, nnz, sliding;
};
PagedArgs paged_args(const TVMFFIAny* args, int i0, const char* fn) {
  PagedArgs a;
  a.pages = arg_tensor(args, i0, fn, "pages");
  a.page_indptr = arg_tensor(args, i0 + 1, fn, "page_indptr");
  a.page_values = arg_tensor(args, i0 + 2, fn, "page_values");
  a.length_info = arg_tensor(args, i0 + 3, fn, "length_info");
  a.k_rope_pos_offset = arg_tensor(args, i0 + 4, fn, "k_rope_pos_offset");
  a.pi = pages_info(a.pages, fn);
  expect_ndim(a.page_indptr, 1, fn, "page_indptr");
  expect_ndim(a.page_values, 1, fn, "page_values");
  expect_ndim(a.k_rope_pos_offset, 1, fn, "k_rope_pos_offset");
  expect_i32(a.page_indptr, fn, "page_indptr");
  expect_i32(a.page_values, fn, "page_values");
  expect_i32(a.length_info, fn, "length_info");
  expect_i32(a.k_rope_pos_offset, fn, "k_rope_pos_offset");
  a.batch = static_cast<int>(a.page_indptr.shape(0)) - 1;
  a.nnz = static_cast<int>(a.page_values.shape(0));
  // [B] for a cache built without sliding-window support, [3,B] otherwise (paged_kv_cache.cc:2443-2452)
  if (a.length_info.ndim() == 1) {
    a.sliding = 0;
    expect_shape(a.length_info.shape(0) == a.batch, fn, "length_info must be [batch_size]");
  } else if (a.length_info.ndim() == 2) {
    a.sliding = 1;
    expect_shape(a.length_info.shape(0) == 3 && a.length_info.shape(1) == a.batch, fn, "length_info must be [3, batch_size]");
  } else {
    throw Err{"ValueError", fmt("%s: length_info.ndim is expected to equal 1 or 2", fn)};
  }
  expect_shape(a.k_rope_pos_offset.shape(0) == a.batch, fn, "k_rope_pos_offset must be [batch_size]");
  return a;
}

// (q, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position, output, lse,
//  rotary_mode, rope_scale, rope_theta, sm_scale)
int impl_decode(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "batch_decode_paged_kv";
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 13, fn);
  Tensor q = arg_tensor(args, 0, fn, "Q");
  PagedArgs pa = paged_args(args, 1, fn);
  Tensor qpos = arg_tensor(args, 6, fn, "q_rope_position"), out = arg_tensor(args, 7, fn, "output"),
         lse = arg_tensor(args, 8, fn, "lse");
  const int rotary_mode = static_cast<int>(arg_int(args, 9, fn, "rotary_mode"));
  const float rope_scale = static_cast<float>(arg_float(args, 10, fn, "rope_scale"));
  const float rope_theta = static_cast<float>(arg_float(args, 11, fn, "rope_theta"));
  const float sm_scale = static_cast<float>(arg_float(args, 12, fn, "sm_scale"));
  expect_ndim(q, 3, fn, "Q");
  expect_ndim(out, 3, fn, "output");
  expect_ndim(lse, 2, fn, "lse");
  expect_ndim(qpos, 1, fn, "q_rope_position");
  expect_i32(qpos, fn, "q_rope_position");
  expect_f32(lse, fn, "lse");
  expect_same_dtype(pa.pages, q, fn, "Q vs pages");
  expect_same_dtype(pa.pages, out, fn, "output vs pages");
  const int B = static_cast<int>(q.shape(0)), hq = static_cast<int>(q.shape(1));
  expect_shape(B == pa.batch, fn, "page_indptr must be [batch_size + 1]");
  expect_shape(q.shape(2) == pa.pi.d, fn, "Q.shape[2] must equal head_dim of pages");
  expect_shape(out.shape(0) == B && out.shape(1) == hq && out.shape(2) == pa.pi.d, fn, "output must match Q");
  expect_shape(lse.shape(0) == B && lse.shape(1) == hq, fn, "lse must be [batch_size, num_qo_heads]");
  expect_shape(qpos.shape(0) == B, fn, "q_rope_position must be [batch_size]");
  DeviceGuard g(q, fn);
  expect_same_device(q, pa.pages, fn, "pages");
  expect_same_device(q, out, fn, "output");
  check_rc(tvmb200_attention_decode(
      q.data, pa.pages.data, static_cast<const int32_t*>(pa.page_indptr.data),
      static_cast<const int32_t*>(pa.page_values.data), static_cast<const int32_t*>(pa.length_info.data),
      static_cast<const int32_t*>(pa.k_rope_pos_offset.data), static_cast<const int32_t*>(qpos.data), out.data,
      static_cast<float*>(lse.data), B, pa.nnz, pa.pi.num_pages, hq, pa.pi.hkv, pa.pi.page_size, pa.pi.d,
      pa.sliding, rotary_mode, rope_scale, rope_theta, sm_scale, pa.pi.dtype, env_stream(g.dev)));
  TVMB200_FFI_END();
}

// (q, q_indptr, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position,
//  output, lse, causal, rotary_mode, rope_scale, rope_theta, sm_scale)
// tree flavour: (..., output, lse, rotary_mode, rope_scale, rope_theta, sm_scale, tree_indptr, tree_order)
int prefill_paged_common(const TVMFFIAny* args, int32_t n, TVMFFIAny* result, bool tree, const char* fn) {
  TVMB200_FFI_BEGIN();
  expect_nargs(n, tree ? 16 : 15, fn);
  Tensor q = arg_tensor(args, 0, fn, "q"), qi = arg_tensor(args, 1, fn, "q_indptr");
  PagedArgs pa = paged_args(args, 2, fn);
  Tensor qpos = arg_tensor(args, 7, fn, "q_rope_position"), out = arg_tensor(args, 8, fn, "output"),
         lse = arg_tensor(args, 9, fn, "lse");
  int causal = 0, i = 10;
  if (!tree) causal = static_cast<int>(arg_int(args, i++, fn, "causal"));
  const int rotary_mode = static_cast<int>(arg_int(args, i++, fn, "rotary_mode"));
  const float rope_scale = static_cast<float>(arg_float(args, i++, fn, "rope_scale"));
  const float rope_theta = static_cast<float>(arg_float(args, i++, fn, "rope_theta"));
  const float sm_scale = static_cast<float>(arg_float(args, i++, fn, "sm_scale"));
  expect_ndim(q, 3, fn, "q");
  expect_ndim(qi, 1, fn, "q_indptr");
  expect_ndim(out, 3, fn, "output");
  expect_ndim(lse, 2, fn, "lse");
  expect_ndim(qpos, 1, fn, "q_rope_position");
  expect_i32(qi, fn, "q_indptr");
  expect_i32(qpos, fn, "q_rope_position");
  expect_f32(lse, fn, "lse");
  expect_same_dtype(pa.pages, q, fn, "q vs pages");
  expect_same_dtype(pa.pages, out, fn, "output vs pages");
  const int total = static_cast<int>(q.shape(0)), hq = static_cast<int>(q.shape(1));
  expect_shape(qi.shape(0) == pa.batch + 1, fn, "q_indptr and page_indptr must both be [batch_size + 1]");
  expect_shape(q.shape(2) == pa.pi.d, fn, "q.shape[2] must equal head_dim of pages");
  expect_shape(out.shape(0) == total && out.shape(1) == hq && out.shape(2) == pa.pi.d, fn, "output must match q");
  expect_shape(lse.shape(0) == total && lse.shape(1) == hq, fn, "lse must be [total_len, num_qo_heads]");
  expect_shape(qpos.shape(0) == total, fn, "q_rope_position must be [total_len]");
  DeviceGuard g(q, fn);
  expect_same_device(q, pa.pages, fn, "pages");
  if (tree) {
    Tensor ti = arg_tensor(args, i++, fn, "tree_order_indptr"), to = arg_tensor(args, i++, fn, "tree_order");
    expect_i32(ti, fn, "tree_order_indptr");
    expect_i32(to, fn, "tree_order");
    expect_ndim(ti, 1, fn, "tree_order_indptr");
    expect_ndim(to, 2, fn, "tree_order");
    expect_shape(ti.shape(0) == pa.batch + 1 && to.shape(1) == 2, fn, "tree_order_indptr [batch_size+1], tree_order [tree_size, 2]");
    check_rc(tvmb200_attention_prefill_tree_paged(
        q.data, static_cast<const int32_t*>(qi.data), pa.pages.data, static_cast<const int32_t*>(pa.page_indptr.data),
        static_cast<const int32_t*>(pa.page_values.data), static_cast<const int32_t*>(pa.length_info.data),
        static_cast<const int32_t*>(pa.k_rope_pos_offset.data), static_cast<const int32_t*>(qpos.data), out.data,
        static_cast<float*>(lse.data), pa.batch, total, pa.nnz, pa.pi.num_pages, hq, pa.pi.hkv, pa.pi.page_size,
        pa.pi.d, rotary_mode, rope_scale, rope_theta, sm_scale, static_cast<const int32_t*>(ti.data),
        static_cast<const int32_t*>(to.data), pa.pi.dtype, env_stream(g.dev)));
  } else {
    check_rc(tvmb200_attention_prefill_paged(
        q.data, static_cast<const int32_t*>(qi.data), pa.pages.data, static_cast<const int32_t*>(pa.page_indptr.data),
        static_cast<const int32_t*>(pa.page_values.data), static_cast<const int32_t*>(pa.length_info.data),
        static_cast<const int32_t*>(pa.k_rope_pos_offset.data), static_cast<const int32_t*>(qpos.data), out.data,
        static_cast<float*>(lse.data), pa.batch, total, pa.nnz, pa.pi.num_pages, hq, pa.pi.hkv, pa.pi.page_size,
        pa.pi.d, pa.sliding, tvmb200::layer_sliding_window_size(), causal, rotary_mode, rope_scale, rope_theta,
        sm_scale, pa.pi.dtype, env_stream(g.dev)));
  }
  TVMB200_FFI_END();
}
int impl_prefill_paged(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  return prefill_paged_common(args, n, result, false, "batch_prefill_paged_kv");
}
int impl_tree_paged(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  return prefill_paged_common(args, n, result, true, "tree_attn_paged_kv");
}

// ragged: (q, q_indptr, k, v, kv_indptr, q_rope_position, k_rope_pos_offset, output, lse,
//          causal, rotary_mode, rope_scale, rope_theta, sm_scale)
// tree:   (q, q_indptr, k, v, kv_indptr, q_rope_position, mn_indptr, mask, output, lse,
//          rotary_mode, rope_scale, rope_theta, sm_scale)
int prefill_ragged_common(const TVMFFIAny* args, int32_t n, TVMFFIAny* result, bool tree, const char* fn) {
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 14, fn);
  Tensor q = arg_tensor(args, 0, fn, "q"), qi = arg_tensor(args, 1, fn, "q_indptr"), k = arg_tensor(args, 2, fn, "k"),
         v = arg_tensor(args, 3, fn, "v"), ki = arg_tensor(args, 4, fn, "kv_indptr"),
         qpos = arg_tensor(args, 5, fn, "q_rope_position");
  int i = 6;
  Tensor a6 = arg_tensor(args, i++, fn, tree ? "mn_indptr" : "k_rope_pos_offset");
  Tensor mask;
  if (tree) mask = arg_tensor(args, i++, fn, "mask");
  Tensor out = arg_tensor(args, i++, fn, "output"), lse = arg_tensor(args, i++, fn, "lse");
  int causal = 0;
  if (!tree) causal = static_cast<int>(arg_int(args, i++, fn, "causal"));
  const int rotary_mode = static_cast<int>(arg_int(args, i++, fn, "rotary_mode"));
  const float rope_scale = static_cast<float>(arg_float(args, i++, fn, "rope_scale"));
  const float rope_theta = static_cast<float>(arg_float(args, i++, fn, "rope_theta"));
  const float sm_scale = static_cast<float>(arg_float(args, i++, fn, "sm_scale"));
  expect_ndim(q, 3, fn, "q");
  expect_ndim(k, 3, fn, "k");
  expect_ndim(v, 3, fn, "v");
  expect_ndim(qi, 1, fn, "q_indptr");
  expect_ndim(ki, 1, fn, "kv_indptr");
  expect_ndim(qpos, 1, fn, "q_rope_position");
  expect_ndim(out, 3, fn, "output");
  expect_ndim(lse, 2, fn, "lse");
  expect_i32(qi, fn, "q_indptr");
  expect_i32(ki, fn, "kv_indptr");
  expect_i32(qpos, fn, "q_rope_position");
  expect_i32(a6, fn, tree ? "mn_indptr" : "k_rope_pos_offset");
  expect_f32(lse, fn, "lse");
  const int dtype = kv_dtype(q, fn, "q");
  expect_same_dtype(q, k, fn, "k vs q");
  expect_same_dtype(q, v, fn, "v vs q");
  expect_same_dtype(q, out, fn, "output vs q");
  const int total = static_cast<int>(q.shape(0)), hq = static_cast<int>(q.shape(1)), d = static_cast<int>(q.shape(2));
  const int total_kv = static_cast<int>(k.shape(0)), hkv = static_cast<int>(k.shape(1));
  const int batch = static_cast<int>(qi.shape(0)) - 1;
  expect_shape(ki.shape(0) == batch + 1, fn, "q_indptr and kv_indptr must both be [batch_size + 1]");
  expect_shape(k.shape(2) == d, fn, "k.shape[2] must equal q.shape[2]");
  expect_shape(v.shape(0) == total_kv && v.shape(1) == hkv && v.shape(2) == d, fn,
               "v must match k (the sm_100a path needs d_v == d_qk)");
  expect_shape(out.shape(0) == total && out.shape(1) == hq && out.shape(2) == d, fn, "output must match q");
  expect_shape(lse.shape(0) == total && lse.shape(1) == hq, fn, "lse must be [total_len, num_qo_heads]");
  expect_shape(qpos.shape(0) == total, fn, "q_rope_position must be [total_len]");
  DeviceGuard g(q, fn);
  expect_same_device(q, k, fn, "k");
  expect_same_device(q, v, fn, "v");
  if (tree) {
    expect_i32(mask, fn, "mask");
    expect_ndim(a6, 1, fn, "mn_indptr");
    expect_shape(a6.shape(0) == batch + 1, fn, "mn_indptr must be [batch_size + 1]");
    check_rc(tvmb200_attention_prefill_tree_ragged(
        q.data, static_cast<const int32_t*>(qi.data), k.data, v.data, static_cast<const int32_t*>(ki.data),
        static_cast<const int32_t*>(qpos.data), static_cast<const int32_t*>(a6.data),
        static_cast<const int32_t*>(mask.data), out.data, static_cast<float*>(lse.data), batch, total, total_kv, hq,
        hkv, d, rotary_mode, rope_scale, rope_theta, sm_scale, dtype, env_stream(g.dev)));
  } else {
    expect_ndim(a6, 1, fn, "k_rope_pos_offset");
    expect_shape(a6.shape(0) == batch, fn, "k_rope_pos_offset must be [batch_size]");
    check_rc(tvmb200_attention_prefill_ragged(
        q.data, static_cast<const int32_t*>(qi.data), k.data, v.data, static_cast<const int32_t*>(ki.data),
        static_cast<const int32_t*>(qpos.data), static_cast<const int32_t*>(a6.data), out.data,
        static_cast<float*>(lse.data), batch, total, total_kv, hq, hkv, d, causal, rotary_mode, rope_scale,
        rope_theta, sm_scale, dtype, env_stream(g.dev)));
  }
  TVMB200_FFI_END();
}
int impl_prefill_ragged(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  return prefill_ragged_common(args, n, result, false, "batch_prefill_ragged_kv");
}
int impl_tree_ragged(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  return prefill_ragged_common(args, n, result, true, "batch_tree_attn");
}

int impl_launch_count(const TVMFFIAny*, int32_t, TVMFFIAny* result) {
  result->type_index = kTVMFFIInt;
  result->zero_padding = 0;
  result->v_int64 = tvmb200_launch_count();
  return 0;
}


// =====================================================================================================
// the host cache under the Relax VM's own global names.  A compiled model calls
// `vm.builtin.paged_attention_kv_cache_create` and the `vm.builtin.kv_state_*` / `attention_kv_cache_*` functions BY
// NAME (src/runtime/vm/kv_state.cc:33-116, paged_kv_cache.cc:2535-2639; callers python/tvm/relax/frontend/nn/llm/
// kv_cache.py:124-351); register_vm_builtins() re-registers those names onto tvm_b200's cache (kv_cache_host.cc), so the
// unmodified model runs on the sm_100a kernels.  The cache travels through the VM registers as a ref-counted ffi object
// (destroyed with its last reference); the callback arguments of the constructor (13..27) are accepted and ignored.
// The MLA entries are registered too and raise.
// =====================================================================================================
// The cache travels through the VM registers as a tvm-ffi Function object whose resource handle is the cache and
// whose deleter destroys it -- a ref-counted ffi Object, like the reference's PagedAttentionKVCacheObj, made with
// nothing but the tvm-ffi C API.  Live handles are kept in a table (object -> cache), so any other Function passed
// where a cache is expected is refused.  A raw opaque pointer (the C ABI's tvmb200_cache_t) is accepted as well.
std::mutex g_handles_mu;
std::unordered_map<void*, tvmb200_cache_t> g_handles;

int cache_handle_call(void* self, const TVMFFIAny*, int32_t, TVMFFIAny* result) {
  result->type_index = kTVMFFIOpaquePtr;  // calling the handle returns the raw tvmb200_cache_t
  result->zero_padding = 0;
  result->v_ptr = self;
  return 0;
}
void cache_handle_delete(void* self) {
  {
    std::lock_guard<std::mutex> lk(g_handles_mu);
    for (auto it = g_handles.begin(); it != g_handles.end();)
      it = it->second == self ? g_handles.erase(it) : std::next(it);
  }
  tvmb200_cache_destroy(static_cast<tvmb200_cache_t>(self));
}

tvmb200_cache_t arg_cache(const TVMFFIAny* args, int i, const char* fn) {
  if (args[i].type_index == kTVMFFIOpaquePtr && args[i].v_ptr != nullptr) return static_cast<tvmb200_cache_t>(args[i].v_ptr);
  if (args[i].type_index == kTVMFFIFunction) {
    std::lock_guard<std::mutex> lk(g_handles_mu);
    auto it = g_handles.find(args[i].v_obj);
    if (it != g_handles.end()) return it->second;
  }
  throw Err{"TypeError", fmt("%s: argument %d must be the cache returned by vm.builtin.paged_attention_kv_cache_create "
                             "of tvm_b200 (got type index %d)", fn, i, args[i].type_index)};
}

struct ShapeView {
  const int64_t* data = nullptr;
  int64_t size = 0;
  int64_t operator[](int64_t i) const { return data[i]; }
};

ShapeView arg_shape(const TVMFFIAny* args, int i, const char* fn, const char* name) {
  if (args[i].type_index != kTVMFFIShape)
    throw Err{"TypeError", fmt("%s: argument %d (%s) must be a Shape, got type index %d", fn, i, name, args[i].type_index)};
  const TVMFFIShapeCell* c = TVMFFIShapeGetCellPtr(args[i].v_obj);
  return ShapeView{c->data, static_cast<int64_t>(c->size)};
}

void cache_rc(int rc) {
  if (rc != 0) throw Err{"RuntimeError", tvmb200_last_error()};
}

void ret_int(TVMFFIAny* result, int64_t v) {
  result->type_index = kTVMFFIInt;
  result->zero_padding = 0;
  result->v_int64 = v;
}

#define TVMB200_VM_BEGIN() try {
#define TVMB200_VM_END()                       \
    return 0;                                  \
  } catch (const Err& e) {                     \
    return raise(e.kind.c_str(), e.msg);       \
  } catch (const std::exception& e) {          \
    return raise("RuntimeError", e.what());    \
  }
#define TVMB200_VM_NONE() do { result->type_index = kTVMFFINone; result->zero_padding = 0; result->v_int64 = 0; } while (0)

int vm_create(void*, const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.paged_attention_kv_cache_create";
  TVMB200_VM_BEGIN();
  if (n != 28 && n != 29) throw Err{"TypeError", fmt("%s expects 28 or 29 arguments, got %d", fn, n)};
  const ShapeView cfg = arg_shape(args, 0, fn, "cache_config"), li = arg_shape(args, 1, fn, "layer_indptr");
  if (cfg.size != 5 && cfg.size != 6) throw Err{"ValueError", fmt("%s: cache_config must have 5 or 6 entries", fn)};
  // the reference picks layer_indptr[group .. group + 1] of the calling Disco worker's pipeline group
  // (paged_kv_cache.cc:2541-2551); there is no Disco worker here, so only a single stage slice can be meant
  if (li.size != 2)
    throw Err{"ValueError", fmt("%s: layer_indptr has %ld entries; tvm_b200 serves one pipeline stage per cache -- pass "
                                "this worker's [begin, end) pair", fn, (long)li.size)};
  const int64_t hq = arg_int(args, 2, fn, "num_qo_heads"), hkv = arg_int(args, 3, fn, "num_kv_heads");
  const int64_t d_qk = arg_int(args, 4, fn, "qk_head_dim"), d_v = arg_int(args, 5, fn, "v_head_dim");
  if (d_qk != d_v) throw Err{"ValueError", fmt("%s: qk_head_dim %ld != v_head_dim %ld (MLA is outside this hot path)", fn, (long)d_qk, (long)d_v)};
  const ShapeView kinds = arg_shape(args, 6, fn, "attn_kinds");
  // enable_kv_transfer: the reference puts the pages on the NVSHMEM symmetric heap (paged_kv_cache.cc:376-405); here the
  // receivers' pools are registered afterwards with the `kv_cache_set_remote_pages` packed function (peer-mapped pointers)
  const bool enable_kv_transfer = arg_int(args, 7, fn, "enable_kv_transfer") != 0;
  if (args[11].type_index != kTVMFFINone)
    throw Err{"ValueError", fmt("%s: rope_ext_factors (longrope) is not supported", fn)};
  const Tensor init = arg_tensor(args, 12, fn, "init");
  if (init.t->device.device_type != kDLCUDA) throw Err{"ValueError", fmt("%s: the cache lives on a CUDA device; there is no CPU fallback", fn)};
  std::vector<int32_t> kinds32(kinds.size);
  for (int64_t i = 0; i < kinds.size; ++i) kinds32[i] = static_cast<int32_t>(kinds[i]);
  tvmb200_cache_config c;
  std::memset(&c, 0, sizeof(c));
  c.reserved_num_seqs = cfg[0];
  c.total_token_capacity = cfg[1];
  c.prefill_chunk_size = cfg[2];
  c.page_size = cfg[3];
  c.support_sliding_window = static_cast<int32_t>(cfg[4]);
  c.layer_sliding_window_size = cfg.size == 6 ? cfg[5] : 0;
  c.layer_id_begin_offset = li[0];  // worker group 0 (single pipeline stage)
  c.num_layers = li[1] - li[0];
  c.num_qo_heads = hq;
  c.num_kv_heads = hkv;
  c.head_dim = d_qk;
  c.attn_kinds = kinds32.empty() ? nullptr : kinds32.data();
  c.rope_mode = static_cast<int32_t>(arg_int(args, 8, fn, "rope_mode"));
  c.rotary_scale = arg_float(args, 9, fn, "rotary_scale");
  c.rotary_theta = arg_float(args, 10, fn, "rotary_theta");
  c.dtype = kv_dtype(init, fn, "init");
  c.device_id = init.t->device.device_id;
  // arguments 13..27 (the compiled callbacks) are not used: the cache drives tvm_b200's own kernels.  What the
  // reference compiled INTO them -- rope_scaling, rotary_dim -- is taken from the calling thread's current kernel-set
  // context (tvmb200_set_rope_scaling / the `set_rope_scaling` packed function) at this moment and stays with this
  // cache, so two models with different scalings can live in one process.
  typedef int (*PFN_Create)(void*, TVMFFISafeCallType, void (*)(void*), TVMFFIObjectHandle*);
  static PFN_Create create = reinterpret_cast<PFN_Create>(ffi_sym("TVMFFIFunctionCreate"));
  if (!create) throw Err{"RuntimeError", "libtvm_ffi.so (TVMFFIFunctionCreate) is not loaded"};
  tvmb200_cache_t cache = nullptr;
  cache_rc(tvmb200_cache_create(&c, &cache));
  if (enable_kv_transfer && tvmb200_cache_enable_kv_transfer(cache, 0, 64, static_cast<int32_t>(hkv)) != 0) {
    tvmb200_cache_destroy(cache);
    throw Err{"RuntimeError", tvmb200_last_error()};
  }
  TVMFFIObjectHandle h = nullptr;
  if (create(cache, cache_handle_call, cache_handle_delete, &h) != 0) {
    tvmb200_cache_destroy(cache);
    return -1;
  }
  {
    std::lock_guard<std::mutex> lk(g_handles_mu);
    g_handles[h] = cache;
  }
  result->type_index = kTVMFFIFunction;
  result->zero_padding = 0;
  result->v_obj = static_cast<TVMFFIObject*>(h);
  TVMB200_VM_END();
}

int vm_clear(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.kv_state_clear";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 1, fn);
  cache_rc(tvmb200_cache_clear(arg_cache(a, 0, fn)));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_add_sequence(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.kv_state_add_sequence";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 2, fn);
  cache_rc(tvmb200_cache_add_sequence(arg_cache(a, 0, fn), arg_int(a, 1, fn, "seq_id")));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_remove_sequence(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.kv_state_remove_sequence";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 2, fn);
  cache_rc(tvmb200_cache_remove_sequence(arg_cache(a, 0, fn), arg_int(a, 1, fn, "seq_id")));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_fork_sequence(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.kv_state_fork_sequence";
  TVMB200_VM_BEGIN();
  if (n != 3 && n != 4) throw Err{"TypeError", fmt("%s expects (cache, parent, child[, fork_pos])", fn)};
  cache_rc(tvmb200_cache_fork_sequence(arg_cache(a, 0, fn), arg_int(a, 1, fn, "parent_seq_id"),
                                       arg_int(a, 2, fn, "child_seq_id"), n == 4 ? arg_int(a, 3, fn, "fork_pos") : -1));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_popn(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.kv_state_popn";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 3, fn);
  cache_rc(tvmb200_cache_popn(arg_cache(a, 0, fn), arg_int(a, 1, fn, "seq_id"), static_cast<int32_t>(arg_int(a, 2, fn, "n"))));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_begin_forward(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.kv_state_begin_forward";
  TVMB200_VM_BEGIN();
  if (n != 3 && n != 4) throw Err{"TypeError", "KVState BeginForward only accepts 3 or 4 arguments"};
  const ShapeView ids = arg_shape(a, 1, fn, "seq_ids"), lens = arg_shape(a, 2, fn, "append_lengths");
  if (ids.size != lens.size) throw Err{"ValueError", fmt("%s: seq_ids and append_lengths differ in length", fn)};
  ShapeView tree;
  if (n == 4 && a[3].type_index != kTVMFFINone) tree = arg_shape(a, 3, fn, "token_tree_parent_ptr");
  cache_rc(tvmb200_cache_begin_forward(arg_cache(a, 0, fn), ids.data, lens.data, static_cast<int32_t>(ids.size),
                                       tree.data, static_cast<int32_t>(tree.size)));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_end_forward(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.kv_state_end_forward";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 1, fn);
  cache_rc(tvmb200_cache_end_forward(arg_cache(a, 0, fn)));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_enable_sliding_window(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_enable_sliding_window_for_seq";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 4, fn);
  cache_rc(tvmb200_cache_enable_sliding_window_for_seq(arg_cache(a, 0, fn), arg_int(a, 1, fn, "seq_id"),
                                                       static_cast<int32_t>(arg_int(a, 2, fn, "sliding_window_size")),
                                                       static_cast<int32_t>(arg_int(a, 3, fn, "attn_sink_size"))));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_commit_tree_nodes(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_commit_accepted_token_tree_nodes";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 3, fn);
  const ShapeView ids = arg_shape(a, 1, fn, "seq_ids"), leaves = arg_shape(a, 2, fn, "leaf_indices");
  if (ids.size != leaves.size) throw Err{"ValueError", fmt("%s: seq_ids and leaf_indices differ in length", fn)};
  cache_rc(tvmb200_cache_commit_accepted_token_tree_nodes(arg_cache(a, 0, fn), ids.data, leaves.data, static_cast<int32_t>(ids.size)));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_empty(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_empty";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 1, fn);
  int32_t v = 0;
  cache_rc(tvmb200_cache_empty(arg_cache(a, 0, fn), &v));
  result->type_index = kTVMFFIBool;
  result->zero_padding = 0;
  result->v_int64 = v != 0;
  TVMB200_VM_END();
}
int vm_num_available_pages(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_get_num_available_pages";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 1, fn);
  int32_t v = 0;
  cache_rc(tvmb200_cache_get_num_available_pages(arg_cache(a, 0, fn), &v));
  ret_int(result, v);
  TVMB200_VM_END();
}
int vm_total_sequence_length(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_get_total_sequence_length";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 1, fn);
  int32_t v = 0;
  cache_rc(tvmb200_cache_get_total_sequence_length(arg_cache(a, 0, fn), &v));
  ret_int(result, v);
  TVMB200_VM_END();
}

// a Tensor view of q_rope_position_map (valid until the next begin_forward), as GetQueryPositions returns
struct PosView {
  DLManagedTensor m;
  int64_t shape[1];
};
void pos_view_deleter(DLManagedTensor* m) { delete reinterpret_cast<PosView*>(m->manager_ctx); }

int vm_query_positions(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_get_query_positions";
  typedef int (*PFN_FromDLPack)(DLManagedTensor*, int32_t, int32_t, TVMFFIObjectHandle*);
  static PFN_FromDLPack from_dlpack = reinterpret_cast<PFN_FromDLPack>(ffi_sym("TVMFFITensorFromDLPack"));
  TVMB200_VM_BEGIN();
  expect_nargs(n, 1, fn);
  if (!from_dlpack) throw Err{"RuntimeError", "libtvm_ffi.so (TVMFFITensorFromDLPack) is not loaded"};
  tvmb200_cache_t c = arg_cache(a, 0, fn);
  int dev = 0;
  cudaGetDevice(&dev);
  const int32_t* ptr = nullptr;
  int64_t len = 0;
  cache_rc(tvmb200_cache_get_query_positions(c, &ptr, &len, env_stream(dev)));
  PosView* v = new PosView();
  v->shape[0] = len;
  v->m.dl_tensor.data = const_cast<int32_t*>(ptr);
  v->m.dl_tensor.device = DLDevice{kDLCUDA, dev};
  v->m.dl_tensor.ndim = 1;
  v->m.dl_tensor.dtype = DLDataType{kDLInt, 32, 1};
  v->m.dl_tensor.shape = v->shape;
  v->m.dl_tensor.strides = nullptr;
  v->m.dl_tensor.byte_offset = 0;
  v->m.manager_ctx = v;
  v->m.deleter = pos_view_deleter;
  TVMFFIObjectHandle h = nullptr;
  if (from_dlpack(&v->m, 0, 0, &h) != 0) return -1;  // error already raised by tvm-ffi
  result->type_index = kTVMFFITensor;
  result->zero_padding = 0;
  result->v_obj = static_cast<TVMFFIObject*>(h);
  TVMB200_VM_END();
}
int vm_debug_get_kv(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_debug_get_kv";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 6, fn);
  const Tensor k = arg_tensor(a, 4, fn, "k_data"), v = arg_tensor(a, 5, fn, "v_data");
  cache_rc(tvmb200_cache_debug_get_kv(arg_cache(a, 0, fn), arg_int(a, 1, fn, "seq_id"), arg_int(a, 2, fn, "start_pos"),
                                      arg_int(a, 3, fn, "end_pos"), k.data, v.data, env_stream(k.t->device.device_id)));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
struct CacheShape {
  int64_t hq, hkv, d, dtype;
};
CacheShape cache_shape(tvmb200_cache_t c) {
  int64_t v[6];
  cache_rc(tvmb200_cache_shape(c, v));
  return CacheShape{v[0], v[1], v[2], v[3]};
}
void expect_cache_dtype(const Tensor& t, const CacheShape& cs, const char* fn, const char* name) {
  if (kv_dtype(t, fn, name) != cs.dtype)
    throw Err{"ValueError", fmt("%s: %s has a different dtype than the cache's pages (%s)", fn, name,
                                cs.dtype == TVMB200_F16 ? "float16" : "bfloat16")};
}
// q / o style tensor [n, heads, head_dim] of the cache's dtype
void expect_cache_heads(const Tensor& t, const CacheShape& cs, int64_t heads, const char* fn, const char* name) {
  expect_cache_dtype(t, cs, fn, name);
  if (t.shape(1) != heads || t.shape(2) != cs.d)
    throw Err{"ValueError", fmt("%s: %s must be [n, %ld, %ld] for this cache, got [%ld, %ld, %ld]", fn, name, (long)heads,
                                (long)cs.d, (long)t.shape(0), (long)t.shape(1), (long)t.shape(2))};
}
int vm_attention_with_fused_qkv(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_attention_with_fused_qkv";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 5, fn);
  const Tensor qkv = arg_tensor(a, 3, fn, "qkv_data"), o = arg_tensor(a, 4, fn, "o_data");
  if (qkv.t->device.device_type != kDLCUDA || o.t->device.device_type != kDLCUDA)
    throw Err{"ValueError", fmt("%s: qkv_data / o_data must be CUDA tensors (there is no CPU fallback)", fn)};
  if (qkv.ndim() != 3 || o.ndim() != 3) throw Err{"ValueError", fmt("%s: qkv_data and o_data must be 3-D", fn)};
  tvmb200_cache_t c = arg_cache(a, 0, fn);
  // the reference's checks (paged_kv_cache.cc:1303-1335): dtype of the pages, head counts, head_dim, o vs qkv
  const CacheShape cs = cache_shape(c);
  expect_cache_dtype(qkv, cs, fn, "qkv_data");
  expect_cache_dtype(o, cs, fn, "o_data");
  if (qkv.shape(1) != cs.hq + 2 * cs.hkv || o.shape(1) != cs.hq || qkv.shape(2) != cs.d || o.shape(2) != cs.d ||
      o.shape(0) != qkv.shape(0))
    throw Err{"ValueError", fmt("%s: qkv_data must be [n, %ld, %ld] and o_data [n, %ld, %ld] for this cache, got [%ld, %ld, %ld] "
                                "and [%ld, %ld, %ld]", fn, (long)(cs.hq + 2 * cs.hkv), (long)cs.d, (long)cs.hq, (long)cs.d,
                                (long)qkv.shape(0), (long)qkv.shape(1), (long)qkv.shape(2), (long)o.shape(0), (long)o.shape(1),
                                (long)o.shape(2))};
  cache_rc(tvmb200_cache_attention_with_fused_qkv(c, arg_int(a, 1, fn, "layer_id"), arg_float(a, 2, fn, "sm_scale"),
                                                  qkv.data, o.data, qkv.shape(0), env_stream(qkv.t->device.device_id)));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
// shared checks of the q / k / v / o / lse tensors of the split attention entries (kv_state.cc:84-115)
void expect_cuda_3d(const Tensor& t, const char* fn, const char* name) {
  if (t.t->device.device_type != kDLCUDA) throw Err{"ValueError", fmt("%s: %s must be a CUDA tensor (there is no CPU fallback)", fn, name)};
  if (t.ndim() != 3) throw Err{"ValueError", fmt("%s: %s must be 3-D", fn, name)};
  (void)kv_dtype(t, fn, name);
}
void expect_lse(const Tensor& lse, const Tensor& q, const char* fn, const char* name) {
  if (lse.t->device.device_type != kDLCUDA) throw Err{"ValueError", fmt("%s: %s must be a CUDA tensor", fn, name)};
  expect_f32(lse, fn, name);
  if (lse.ndim() != 2 || lse.shape(0) != q.shape(0) || lse.shape(1) != q.shape(1))
    throw Err{"ValueError", fmt("%s: %s must be [total_len, num_qo_heads] float32", fn, name)};
}
int vm_self_attention(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_self_attention";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 8, fn);
  const Tensor q = arg_tensor(a, 3, fn, "q_data"), k = arg_tensor(a, 4, fn, "k_data"), v = arg_tensor(a, 5, fn, "v_data");
  const Tensor o = arg_tensor(a, 6, fn, "o_data"), lse = arg_tensor(a, 7, fn, "lse_data");
  expect_cuda_3d(q, fn, "q_data");
  expect_cuda_3d(k, fn, "k_data");
  expect_cuda_3d(v, fn, "v_data");
  expect_cuda_3d(o, fn, "o_data");
  expect_lse(lse, q, fn, "lse_data");
  if (k.shape(0) != q.shape(0) || v.shape(0) != q.shape(0) || o.shape(0) != q.shape(0))
    throw Err{"ValueError", fmt("%s: q / k / v / o must have the same number of rows", fn)};
  tvmb200_cache_t c = arg_cache(a, 0, fn);
  const CacheShape cs = cache_shape(c);
  expect_cache_heads(q, cs, cs.hq, fn, "q_data");
  expect_cache_heads(k, cs, cs.hkv, fn, "k_data");
  expect_cache_heads(v, cs, cs.hkv, fn, "v_data");
  expect_cache_heads(o, cs, cs.hq, fn, "o_data");
  cache_rc(tvmb200_cache_self_attention(c, arg_int(a, 1, fn, "layer_id"), arg_float(a, 2, fn, "sm_scale"),
                                        q.data, k.data, v.data, o.data, static_cast<float*>(lse.data), q.shape(0),
                                        env_stream(q.t->device.device_id)));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_cross_attention(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_cross_attention";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 6, fn);
  const Tensor q = arg_tensor(a, 3, fn, "q_data"), o = arg_tensor(a, 4, fn, "o_data"), lse = arg_tensor(a, 5, fn, "lse_data");
  expect_cuda_3d(q, fn, "q_data");
  expect_cuda_3d(o, fn, "o_data");
  expect_lse(lse, q, fn, "lse_data");
  if (o.shape(0) != q.shape(0)) throw Err{"ValueError", fmt("%s: q and o must have the same number of rows", fn)};
  tvmb200_cache_t c = arg_cache(a, 0, fn);
  const CacheShape cs = cache_shape(c);
  expect_cache_heads(q, cs, cs.hq, fn, "q_data");
  expect_cache_heads(o, cs, cs.hq, fn, "o_data");
  cache_rc(tvmb200_cache_cross_attention(c, arg_int(a, 1, fn, "layer_id"), arg_float(a, 2, fn, "sm_scale"),
                                         q.data, o.data, static_cast<float*>(lse.data), q.shape(0),
                                         env_stream(q.t->device.device_id)));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
int vm_attention_with_shared_kv(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_attention_with_shared_kv";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 7, fn);
  const Tensor q = arg_tensor(a, 3, fn, "q_data"), k = arg_tensor(a, 4, fn, "current_k_data");
  const Tensor v = arg_tensor(a, 5, fn, "current_v_data"), o = arg_tensor(a, 6, fn, "o_data");
  expect_cuda_3d(q, fn, "q_data");
  expect_cuda_3d(k, fn, "current_k_data");
  expect_cuda_3d(v, fn, "current_v_data");
  expect_cuda_3d(o, fn, "o_data");
  if (k.shape(0) != q.shape(0) || v.shape(0) != q.shape(0) || o.shape(0) != q.shape(0))
    throw Err{"ValueError", fmt("%s: q / current_k / current_v / o must have the same number of rows", fn)};
  tvmb200_cache_t c = arg_cache(a, 0, fn);
  const CacheShape cs = cache_shape(c);
  expect_cache_heads(q, cs, cs.hq, fn, "q_data");
  expect_cache_heads(k, cs, cs.hkv, fn, "current_k_data");
  expect_cache_heads(v, cs, cs.hkv, fn, "current_v_data");
  expect_cache_heads(o, cs, cs.hq, fn, "o_data");
  cache_rc(tvmb200_cache_attention_with_shared_kv(c, arg_int(a, 1, fn, "source_layer_id"),
                                                  arg_float(a, 2, fn, "sm_scale"), q.data, k.data, v.data, o.data, q.shape(0),
                                                  env_stream(q.t->device.device_id)));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
// returns Array<Tensor>{o_self_attn, lse_self_attn}, built with tvm-ffi's own "ffi.Array" constructor
int vm_merge_attn_output_inplace(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.attention_kv_cache_merge_attn_output_inplace";
  typedef int (*PFN_GetGlobal)(const TVMFFIByteArray*, TVMFFIObjectHandle*);
  typedef int (*PFN_Call)(TVMFFIObjectHandle, TVMFFIAny*, int32_t, TVMFFIAny*);
  typedef int (*PFN_DecRef)(TVMFFIObjectHandle);
  static PFN_GetGlobal get_global = reinterpret_cast<PFN_GetGlobal>(ffi_sym("TVMFFIFunctionGetGlobal"));
  static PFN_Call call = reinterpret_cast<PFN_Call>(ffi_sym("TVMFFIFunctionCall"));
  static PFN_DecRef dec_ref = reinterpret_cast<PFN_DecRef>(ffi_sym("TVMFFIObjectDecRef"));
  TVMB200_VM_BEGIN();
  expect_nargs(n, 5, fn);
  if (!get_global || !call || !dec_ref) throw Err{"RuntimeError", "libtvm_ffi.so is not loaded"};
  const Tensor os = arg_tensor(a, 1, fn, "o_self_attn"), ls = arg_tensor(a, 2, fn, "lse_self_attn");
  const Tensor oc = arg_tensor(a, 3, fn, "o_cross_attn"), lc = arg_tensor(a, 4, fn, "lse_cross_attn");
  expect_cuda_3d(os, fn, "o_self_attn");
  expect_cuda_3d(oc, fn, "o_cross_attn");
  expect_lse(ls, os, fn, "lse_self_attn");
  expect_lse(lc, os, fn, "lse_cross_attn");
  expect_same_dtype(os, oc, fn, "o_cross_attn vs o_self_attn");
  for (int i = 0; i < 3; ++i)
    if (oc.shape(i) != os.shape(i)) throw Err{"ValueError", fmt("%s: o_cross_attn must have the shape of o_self_attn", fn)};
  cache_rc(tvmb200_cache_merge_attn_output_inplace(arg_cache(a, 0, fn), os.data, static_cast<float*>(ls.data), oc.data,
                                                   static_cast<const float*>(lc.data), os.shape(0), os.shape(1), os.shape(2),
                                                   env_stream(os.t->device.device_id)));
  static const char kArray[] = "ffi.Array";
  const TVMFFIByteArray name{kArray, sizeof(kArray) - 1};
  TVMFFIObjectHandle ctor = nullptr;
  if (get_global(&name, &ctor) != 0) return -1;
  if (ctor == nullptr) throw Err{"RuntimeError", "tvm-ffi global function ffi.Array is missing"};
  TVMFFIAny pair[2] = {a[1], a[2]};
  const int rc = call(ctor, pair, 2, result);
  dec_ref(ctor);
  if (rc != 0) return -1;
  TVMB200_VM_END();
}
// ffi.Shape(*ints) through tvm-ffi's own constructor (no object layout assumptions)
int return_shape(const std::vector<int64_t>& v, TVMFFIAny* result) {
  typedef int (*PFN_GetGlobal)(const TVMFFIByteArray*, TVMFFIObjectHandle*);
  typedef int (*PFN_Call)(TVMFFIObjectHandle, TVMFFIAny*, int32_t, TVMFFIAny*);
  typedef int (*PFN_DecRef)(TVMFFIObjectHandle);
  static PFN_GetGlobal get_global = reinterpret_cast<PFN_GetGlobal>(ffi_sym("TVMFFIFunctionGetGlobal"));
  static PFN_Call call = reinterpret_cast<PFN_Call>(ffi_sym("TVMFFIFunctionCall"));
  static PFN_DecRef dec_ref = reinterpret_cast<PFN_DecRef>(ffi_sym("TVMFFIObjectDecRef"));
  if (!get_global || !call || !dec_ref) throw Err{"RuntimeError", "libtvm_ffi.so is not loaded"};
  static const char kShape[] = "ffi.Shape";
  const TVMFFIByteArray name{kShape, sizeof(kShape) - 1};
  TVMFFIObjectHandle ctor = nullptr;
  if (get_global(&name, &ctor) != 0) return -1;
  if (ctor == nullptr) throw Err{"RuntimeError", "tvm-ffi global function ffi.Shape is missing"};
  std::vector<TVMFFIAny> args(v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    args[i].type_index = kTVMFFIInt;
    args[i].zero_padding = 0;
    args[i].v_int64 = v[i];
  }
  const int rc = call(ctor, args.data(), static_cast<int32_t>(args.size()), result);
  dec_ref(ctor);
  return rc;
}
// kv_cache_disagg_prepare_recv(cache, seq_id, append_length) -> Shape  (kv_state.cc, paged_kv_cache.cc:1220-1248)
int vm_disagg_prepare_recv(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.kv_cache_disagg_prepare_recv";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 3, fn);
  const int64_t len = arg_int(a, 2, fn, "append_length");
  if (len <= 0) throw Err{"ValueError", fmt("%s: append_length %ld", fn, (long)len)};
  std::vector<int64_t> out(static_cast<size_t>(2 * len + 1));
  int64_t used = 0;
  cache_rc(tvmb200_cache_disagg_prepare_recv(arg_cache(a, 0, fn), arg_int(a, 1, fn, "seq_id"), len, out.data(),
                                             static_cast<int64_t>(out.size()), &used));
  out.resize(static_cast<size_t>(used));
  if (return_shape(out, result) != 0) return -1;
  TVMB200_VM_END();
}
// kv_cache_disagg_mark_send(cache, seq_id, begin, compressed_remote_position_map, recver_pe_offset)  (:1250-1301)
int vm_disagg_mark_send(void*, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "vm.builtin.kv_cache_disagg_mark_send";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 5, fn);
  const ShapeView m = arg_shape(a, 3, fn, "compressed_remote_position_map");
  cache_rc(tvmb200_cache_disagg_mark_send(arg_cache(a, 0, fn), arg_int(a, 1, fn, "seq_id"), arg_int(a, 2, fn, "begin"), m.data,
                                          m.size, static_cast<int32_t>(arg_int(a, 4, fn, "recver_pe_offset"))));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
// kv_cache_enable_kv_transfer(cache, local_tp_rank, num_pe, remote_num_kv_heads): the geometry NVSHMEM / Disco would supply
int impl_kv_cache_enable_kv_transfer(const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "kv_cache_enable_kv_transfer";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 4, fn);
  cache_rc(tvmb200_cache_enable_kv_transfer(arg_cache(a, 0, fn), static_cast<int32_t>(arg_int(a, 1, fn, "local_tp_rank")),
                                            static_cast<int32_t>(arg_int(a, 2, fn, "num_pe")),
                                            static_cast<int32_t>(arg_int(a, 3, fn, "remote_num_kv_heads"))));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
// kv_cache_set_remote_pages(cache, pe, local_layer, pages): `pages` = the receiver's pool of that layer as a (peer-mapped)
// CUDA tensor, an opaque pointer or an integer address
int impl_kv_cache_set_remote_pages(const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "kv_cache_set_remote_pages";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 4, fn);
  void* ptr = nullptr;
  if (a[3].type_index == kTVMFFIInt) ptr = reinterpret_cast<void*>(static_cast<uintptr_t>(a[3].v_int64));
  else if (a[3].type_index == kTVMFFIOpaquePtr) ptr = a[3].v_ptr;
  else ptr = arg_tensor(a, 3, fn, "pages").data;
  cache_rc(tvmb200_cache_set_remote_pages(arg_cache(a, 0, fn), static_cast<int32_t>(arg_int(a, 1, fn, "pe")),
                                          arg_int(a, 2, fn, "local_layer"), ptr));
  TVMB200_VM_NONE();
  TVMB200_VM_END();
}
// kv_cache_pages(cache, local_layer) -> opaque pointer of this cache's pool of the layer (what a peer registers)
int impl_kv_cache_pages(const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "kv_cache_pages";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 2, fn);
  void* ptr = nullptr;
  int64_t np = 0;
  cache_rc(tvmb200_cache_pages(arg_cache(a, 0, fn), arg_int(a, 1, fn, "local_layer"), &ptr, &np));
  result->type_index = kTVMFFIOpaquePtr;
  result->zero_padding = 0;
  result->v_ptr = ptr;
  TVMB200_VM_END();
}
int vm_unsupported(void*, const TVMFFIAny*, int32_t, TVMFFIAny*) {
  return raise("RuntimeError", "tvm_b200: this vm.builtin KV-cache entry (MLA) is "
                               "outside the PagedKVCache MHA hot path and is not implemented");
}

struct VmBuiltin {
  const char* name;
  TVMFFISafeCallType fn;
};
const VmBuiltin kVmBuiltins[] = {
    {"vm.builtin.paged_attention_kv_cache_create", vm_create},
    {"vm.builtin.kv_state_clear", vm_clear},
    {"vm.builtin.kv_state_add_sequence", vm_add_sequence},
    {"vm.builtin.kv_state_remove_sequence", vm_remove_sequence},
    {"vm.builtin.kv_state_fork_sequence", vm_fork_sequence},
    {"vm.builtin.kv_state_popn", vm_popn},
    {"vm.builtin.kv_state_begin_forward", vm_begin_forward},
    {"vm.builtin.kv_state_end_forward", vm_end_forward},
    {"vm.builtin.attention_kv_cache_enable_sliding_window_for_seq", vm_enable_sliding_window},
    {"vm.builtin.attention_kv_cache_commit_accepted_token_tree_nodes", vm_commit_tree_nodes},
    {"vm.builtin.attention_kv_cache_empty", vm_empty},
    {"vm.builtin.attention_kv_cache_get_num_available_pages", vm_num_available_pages},
    {"vm.builtin.attention_kv_cache_get_total_sequence_length", vm_total_sequence_length},
    {"vm.builtin.attention_kv_cache_get_query_positions", vm_query_positions},
    {"vm.builtin.attention_kv_cache_debug_get_kv", vm_debug_get_kv},
    {"vm.builtin.attention_kv_cache_attention_with_fused_qkv", vm_attention_with_fused_qkv},
    {"vm.builtin.kv_cache_disagg_prepare_recv", vm_disagg_prepare_recv},
    {"vm.builtin.kv_cache_disagg_mark_send", vm_disagg_mark_send},
    {"vm.builtin.attention_kv_cache_debug_get_kv_mla", vm_unsupported},
    {"vm.builtin.attention_kv_cache_self_attention", vm_self_attention},
    {"vm.builtin.attention_kv_cache_cross_attention", vm_cross_attention},
    {"vm.builtin.attention_kv_cache_attention_with_shared_kv", vm_attention_with_shared_kv},
    {"vm.builtin.attention_kv_cache_merge_attn_output_inplace", vm_merge_attn_output_inplace},
    {"vm.builtin.attention_kv_cache_append_mla_kv", vm_unsupported},
};

// (allow_override) -> number of names registered
int impl_register_vm_builtins(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "register_vm_builtins";
  TVMB200_VM_BEGIN();
  if (n > 1) throw Err{"TypeError", "register_vm_builtins expects ([allow_override = 1])"};
  const int allow = n == 1 ? static_cast<int>(arg_int(args, 0, fn, "allow_override")) : 1;
  const int rc = tvmb200_register_vm_builtins(allow);
  if (rc < 0) return -1;
  ret_int(result, rc);
  TVMB200_VM_END();
}

}  // namespace

extern "C" int tvmb200_register_vm_builtins(int allow_override) {
  typedef int (*PFN_Create)(void*, TVMFFISafeCallType, void (*)(void*), TVMFFIObjectHandle*);
  typedef int (*PFN_SetGlobal)(const TVMFFIByteArray*, TVMFFIObjectHandle, int);
  typedef int (*PFN_DecRef)(TVMFFIObjectHandle);
  static PFN_Create create = reinterpret_cast<PFN_Create>(ffi_sym("TVMFFIFunctionCreate"));
  static PFN_SetGlobal set_global = reinterpret_cast<PFN_SetGlobal>(ffi_sym("TVMFFIFunctionSetGlobal"));
  static PFN_DecRef dec_ref = reinterpret_cast<PFN_DecRef>(ffi_sym("TVMFFIObjectDecRef"));
  if (!create || !set_global || !dec_ref) {
    raise("RuntimeError", "tvmb200_register_vm_builtins: libtvm_ffi.so is not loaded in this process");
    return -1;
  }
  int count = 0;
  for (const VmBuiltin& b : kVmBuiltins) {
    TVMFFIObjectHandle f = nullptr;
    if (create(nullptr, b.fn, nullptr, &f) != 0) return -1;
    const TVMFFIByteArray name{b.name, std::strlen(b.name)};
    const int rc = set_global(&name, f, allow_override);
    dec_ref(f);
    if (rc != 0) return -1;
    ++count;
  }
  return count;
}

// role names used by the reference cache constructor (paged_kv_cache.cc:2573-2603), the global_symbols of the reference
// PrimFuncs they replace, and the state the reference bakes in at TIR build time
#define TVMB200_CALLBACKS(X)                                                  \
  X(f_transpose_append, impl_transpose_append)                                \
  X(f_attention_decode, impl_decode)                                          \
  X(f_attention_decode_sliding_window, impl_decode)                           \
  X(f_attention_prefill, impl_prefill_paged)                                  \
  X(f_attention_prefill_sliding_window, impl_prefill_paged)                   \
  X(f_attention_prefill_ragged, impl_prefill_ragged)                          \
  X(f_attention_prefill_with_tree_mask, impl_tree_ragged)                     \
  X(f_attention_prefill_with_tree_mask_paged_kv, impl_tree_paged)             \
  X(f_merge_inplace, impl_merge_state_inplace)                                \
  X(f_split_rotary, impl_fused_rope)                                          \
  X(f_copy_single_page, impl_copy_single_page)                                \
  X(f_debug_get_kv, impl_debug_get_kv)                                        \
  X(f_compact_copy, impl_compact_kv_copy)                                     \
  X(tir_kv_cache_transpose_append, impl_transpose_append)                     \
  X(batch_decode_paged_kv, impl_decode)                                       \
  X(batch_decode_paged_kv_sliding_window, impl_decode)                        \
  X(batch_prefill_paged_kv, impl_prefill_paged)                               \
  X(batch_prefill_paged_kv_sliding_window, impl_prefill_paged)                \
  X(batch_prefill_ragged_kv, impl_prefill_ragged)                             \
  X(batch_tree_attn, impl_tree_ragged)                                        \
  X(tree_attn_paged_kv, impl_tree_paged)                                      \
  X(merge_state_inplace, impl_merge_state_inplace)                            \
  X(fused_rope, impl_fused_rope)                                              \
  X(copy_single_page, impl_copy_single_page)                                  \
  X(tir_kv_cache_debug_get_kv, impl_debug_get_kv)                             \
  X(compact_kv_copy, impl_compact_kv_copy)                                    \
  X(set_rope_params, impl_set_rope_params)                                    \
  X(set_layer_sliding_window_size, impl_set_layer_sliding_window_size)        \
  X(set_rope_scaling, impl_set_rope_scaling)                                  \
  X(set_rope_scaling_yarn, impl_set_rope_scaling_yarn)                        \
  X(launch_count, impl_launch_count)                                          \
  X(kv_cache_enable_kv_transfer, impl_kv_cache_enable_kv_transfer)            \
  X(kv_cache_set_remote_pages, impl_kv_cache_set_remote_pages)                \
  X(kv_cache_pages, impl_kv_cache_pages)                                      \
  X(register_vm_builtins, impl_register_vm_builtins)

namespace {

// ---- kernel-set contexts for the packed path --------------------------------------------------------------------
// The module symbols above run in the DEFAULT context.  The reference compiles one kernel set per model (rope_scaling,
// rotary_dim, layer window baked in); the equivalent here is a context plus closures bound to it:
//   ctx = context_create()                                   (opaque pointer; context_release(ctx) when done)
//   bind_context(ctx, "set_rope_scaling")(1, 8.0, 1.0, 4.0, 8192)
//   f_decode = bind_context(ctx, "f_attention_decode")       (an ffi Function that holds a reference on ctx)
// Two models with different rope scalings, or two streams, then never share settings or scratch.
typedef int (*ImplFn)(const TVMFFIAny*, int32_t, TVMFFIAny*);
struct Bound {
  tvmb200_context_t ctx;
  ImplFn impl;
};
int bound_call(void* self, const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  Bound* b = static_cast<Bound*>(self);
  tvmb200::ContextScope scope(reinterpret_cast<tvmb200::Context*>(b->ctx));
  return b->impl(args, n, result);
}
void bound_delete(void* self) {
  Bound* b = static_cast<Bound*>(self);
  tvmb200_context_release(b->ctx);
  delete b;
}
ImplFn find_impl(const std::string& name) {
#define X(sym, impl) \
  if (name == #sym) return impl;
  TVMB200_CALLBACKS(X)
#undef X
  return nullptr;
}
tvmb200_context_t arg_context(const TVMFFIAny* args, int i, const char* fn) {
  if (args[i].type_index != kTVMFFIOpaquePtr || args[i].v_ptr == nullptr)
    throw Err{"TypeError", fmt("%s: argument %d must be a context made by context_create", fn, i)};
  return static_cast<tvmb200_context_t>(args[i].v_ptr);
}
int impl_context_create(const TVMFFIAny*, int32_t n, TVMFFIAny* result) {
  static const char* fn = "context_create";
  TVMB200_VM_BEGIN();
  expect_nargs(n, 0, fn);
  tvmb200_context_t c = nullptr;
  check_rc(tvmb200_context_create(&c));
  result->type_index = kTVMFFIOpaquePtr;
  result->zero_padding = 0;
  result->v_ptr = c;
  TVMB200_VM_END();
}
int impl_context_release(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "context_release";
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 1, fn);
  tvmb200_context_release(arg_context(args, 0, fn));
  TVMB200_FFI_END();
}
int impl_bind_context(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "bind_context";
  typedef int (*PFN_Create)(void*, TVMFFISafeCallType, void (*)(void*), TVMFFIObjectHandle*);
  static PFN_Create create = reinterpret_cast<PFN_Create>(ffi_sym("TVMFFIFunctionCreate"));
  TVMB200_VM_BEGIN();
  expect_nargs(n, 2, fn);
  if (!create) throw Err{"RuntimeError", "libtvm_ffi.so (TVMFFIFunctionCreate) is not loaded"};
  tvmb200_context_t ctx = arg_context(args, 0, fn);
  const std::string name = arg_str(args, 1, fn, "name");
  ImplFn impl = find_impl(name);
  if (!impl) throw Err{"ValueError", fmt("%s: tvm_b200 exports no packed function named \"%s\"", fn, name.c_str())};
  Bound* b = new Bound{ctx, impl};
  tvmb200_context_retain(ctx);
  TVMFFIObjectHandle h = nullptr;
  if (create(b, bound_call, bound_delete, &h) != 0) {
    bound_delete(b);
    return -1;
  }
  result->type_index = kTVMFFIFunction;
  result->zero_padding = 0;
  result->v_obj = static_cast<TVMFFIObject*>(h);
  TVMB200_VM_END();
}

// ---- nvshmem.KVTransfer / nvshmem.KVTransferPageToPage without NVSHMEM ---------------------------------------------------
//   f = bind_kv_transfer(Shape(pe_0_ptr, pe_1_ptr, ...), local_tp_rank, page_to_page)
// returns a Function with the REFERENCE's signature (src/runtime/extra/contrib/nvshmem/kv_transfer.cu:139-257, :259-325):
//   f(remote_pages, k, v, remote_position_map, remote_tp_group_pe_offset, transfer_stream)                     (page_to_page 0)
//   f(remote_pages, local_pages, remote_position_map, local_position_map, remote_tp_group_pe_offset, stream)   (page_to_page 1)
// `remote_pages` is, as in the reference, the caller's OWN pool of the layer: only its geometry is read (shape[2] = the
// receivers' kv heads per rank, shape[3] = page size), and its data pointer locates the layer inside the pool, so that one
// table of pool BASE pointers per PE serves every layer (NVSHMEM's symmetric addressing: same offset on every PE).  A
// maintainer registers the two results under `nvshmem.KVTransfer` / `nvshmem.KVTransferPageToPage`.
struct BoundTransfer {
  std::vector<void*> pe_base;   // base address of every PE's page pool (peer-mapped)
  void* local_base;             // base address of this rank's own pool (pe_base[own pe]); layer offset = pages.data - local_base
  int32_t local_tp_rank;
  bool page_to_page;
};
int bound_transfer_call(void* self, const TVMFFIAny* a, int32_t n, TVMFFIAny* result) {
  static const char* fn = "nvshmem.KVTransfer";
  BoundTransfer* b = static_cast<BoundTransfer*>(self);
  TVMB200_FFI_BEGIN();
  expect_nargs(n, 6, fn);
  const Tensor pages = arg_tensor(a, 0, fn, "remote_pages");
  if (pages.ndim() != 5) throw Err{"ValueError", fmt("%s: remote_pages must be [num_pages, 2, kv_heads, page_size, head_dim]", fn)};
  const int dtype = kv_dtype(pages, fn, "remote_pages");
  const int64_t layer_off = static_cast<char*>(pages.data) - static_cast<char*>(b->local_base);
  std::vector<void*> table(b->pe_base.size());
  for (size_t i = 0; i < table.size(); ++i) table[i] = b->pe_base[i] ? static_cast<char*>(b->pe_base[i]) + layer_off : nullptr;
  void* stream = nullptr;
  if (a[5].type_index == kTVMFFIOpaquePtr) stream = a[5].v_ptr;
  else if (a[5].type_index == kTVMFFINone) stream = env_stream(pages.t->device.device_id);
  else stream = reinterpret_cast<void*>(static_cast<uintptr_t>(arg_int(a, 5, fn, "transfer_stream")));
  const int32_t remote_h = static_cast<int32_t>(pages.shape(2)), page = static_cast<int32_t>(pages.shape(3)),
                d = static_cast<int32_t>(pages.shape(4));
  if (!b->page_to_page) {
    const Tensor k = arg_tensor(a, 1, fn, "k"), v = arg_tensor(a, 2, fn, "v"), pos = arg_tensor(a, 3, fn, "remote_position_map"),
                 pe = arg_tensor(a, 4, fn, "remote_tp_group_pe_offset");
    if (k.ndim() != 3 || v.ndim() != 3 || k.shape(2) != d) throw Err{"ValueError", fmt("%s: k / v must be [ntokens, kv_heads, head_dim]", fn)};
    check_rc(tvmb200_kv_transfer(table.data(), k.data, v.data, static_cast<const int32_t*>(pos.data),
                                 static_cast<const int32_t*>(pe.data), pos.shape(0), static_cast<int32_t>(k.shape(1)), remote_h, page,
                                 d, b->local_tp_rank, static_cast<int32_t>(table.size()), dtype, stream));
  } else {
    const Tensor lp = arg_tensor(a, 1, fn, "local_pages"), rpos = arg_tensor(a, 2, fn, "remote_position_map"),
                 lpos = arg_tensor(a, 3, fn, "local_position_map"), pe = arg_tensor(a, 4, fn, "remote_tp_group_pe_offset");
    if (lp.ndim() != 5 || lp.shape(4) != d) throw Err{"ValueError", fmt("%s: local_pages must be a 5-d page pool of the same head_dim", fn)};
    check_rc(tvmb200_kv_transfer_page_to_page(table.data(), lp.data, static_cast<const int32_t*>(rpos.data),
                                              static_cast<const int32_t*>(lpos.data), static_cast<const int32_t*>(pe.data),
                                              rpos.shape(0), static_cast<int32_t>(lp.shape(2)), remote_h, page, d, b->local_tp_rank,
                                              static_cast<int32_t>(table.size()), dtype, stream));
  }
  TVMB200_FFI_END();
}
void bound_transfer_delete(void* self) { delete static_cast<BoundTransfer*>(self); }
int impl_bind_kv_transfer(const TVMFFIAny* args, int32_t n, TVMFFIAny* result) {
  static const char* fn = "bind_kv_transfer";
  typedef int (*PFN_Create)(void*, TVMFFISafeCallType, void (*)(void*), TVMFFIObjectHandle*);
  static PFN_Create create = reinterpret_cast<PFN_Create>(ffi_sym("TVMFFIFunctionCreate"));
  TVMB200_VM_BEGIN();
  expect_nargs(n, 4, fn);
  if (!create) throw Err{"RuntimeError", "libtvm_ffi.so (TVMFFIFunctionCreate) is not loaded"};
  const ShapeView ptrs = arg_shape(args, 0, fn, "pe_pool_base_pointers");
  if (ptrs.size < 1 || ptrs.size > 64) throw Err{"ValueError", fmt("%s: 1..64 processing elements", fn)};
  const int64_t own = arg_int(args, 1, fn, "own_pe");
  if (own < 0 || own >= ptrs.size) throw Err{"ValueError", fmt("%s: own_pe %ld out of range", fn, (long)own)};
  BoundTransfer* b = new BoundTransfer;
  for (int64_t i = 0; i < ptrs.size; ++i) b->pe_base.push_back(reinterpret_cast<void*>(static_cast<uintptr_t>(ptrs[i])));
  b->local_base = b->pe_base[static_cast<size_t>(own)];
  b->local_tp_rank = static_cast<int32_t>(arg_int(args, 2, fn, "local_tp_rank"));
  b->page_to_page = arg_int(args, 3, fn, "page_to_page") != 0;
  TVMFFIObjectHandle h = nullptr;
  if (create(b, bound_transfer_call, bound_transfer_delete, &h) != 0) {
    delete b;
    return -1;
  }
  result->type_index = kTVMFFIFunction;
  result->zero_padding = 0;
  result->v_obj = static_cast<TVMFFIObject*>(h);
  TVMB200_VM_END();
}

}  // namespace

#define TVMB200_EXPORT(name, impl)                                                                      \
  extern "C" __attribute__((visibility("default"))) int __tvm_ffi_##name(void* self, const TVMFFIAny* args, \
                                                                         int32_t num_args, TVMFFIAny* result) { \
    (void)self;                                                                                         \
    return impl(args, num_args, result);                                                                \
  }

TVMB200_CALLBACKS(TVMB200_EXPORT)
TVMB200_EXPORT(context_create, impl_context_create)
TVMB200_EXPORT(context_release, impl_context_release)
TVMB200_EXPORT(bind_context, impl_bind_context)
TVMB200_EXPORT(bind_kv_transfer, impl_bind_kv_transfer)
