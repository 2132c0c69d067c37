// Shared device/host helpers for the B200 (sm_100a) PagedKVCache kernel set.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/tvm_b200.h"

namespace tvmb200 {

// ---------------------------------------------------------------------------------------------
// host side: error reporting + launch accounting
// ---------------------------------------------------------------------------------------------
std::string& last_error_ref();
int set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;

#define TVMB200_CHECK(cond, ...)                      \
  do {                                                \
    if (!(cond)) return ::tvmb200::set_error(__VA_ARGS__); \
  } while (0)

#define TVMB200_CUDA(expr)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ::tvmb200::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__),   \
                                  __FILE__, __LINE__, #expr);                               \
  } while (0)

// checks the launch, bumps the launch counter
#define TVMB200_LAUNCH_OK()                      \
  do {                                           \
    TVMB200_CUDA(cudaGetLastError());            \
    ::tvmb200::g_launch_count.fetch_add(1);      \
  } while (0)

int num_sms();  // SM count of the current device (cached)

// scratch of the current context for launches on (current device, stream): split-KV partials (O / LSE / chunk
// offsets) and two zeroed int32 counters ([0] prefill work queue, [64] peer-gather ticket); see Context below
int get_workspace(int64_t bytes, cudaStream_t st, void** out);
int get_counters(cudaStream_t st, int32_t** out);
int32_t layer_sliding_window_size();

// TMA tensor maps (decode.cu): 2-D [rows][cols] with box [box_rows][64], cached by (base, shape); 3-D uncached
int get_tmap_2d_cached(CUtensorMap* out, const void* base, int dtype, uint64_t rows, uint64_t cols,
                       uint32_t box_rows);
int make_tmap_3d(CUtensorMap* out, const void* base, int dtype, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b1, uint32_t b2);

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNegInit = -5e4f;  // the reference's running-max sentinel (_decode_kernels.py:134)

template <typename T>
struct DT;
template <>
struct DT<__half> {
  using T2 = __half2;
  static __device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
  static __device__ __forceinline__ __half from_f(float x) { return __float2half_rn(x); }
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ __half neg(__half x) { return __hneg(x); }
};
template <>
struct DT<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static __device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float x) { return __float2bfloat16_rn(x); }
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    // bf16 -> f32 is a 16-bit shift
    float2 r;
    r.x = __uint_as_float(u << 16);
    r.y = __uint_as_float(u & 0xffff0000u);
    return r;
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ __nv_bfloat16 neg(__nv_bfloat16 x) { return __hneg(x); }
};

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_na_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- mbarrier / TMA (cp.async.bulk.tensor) ---------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// ---- programmatic dependent launch (griddepcontrol): no-ops unless the launch carries the attribute --------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// non-suspending poll (try_wait may park the thread for a HW time slice; pollers that watch several barriers use this)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tmap, int c0,
                                            int c1, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst_smem),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- legacy tensor path (mma.sync), used by the HBM-bound decode kernel and the generic prefill ----
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                            uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                                  uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
template <typename T>
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                          uint32_t a3, uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&c)[4], uint32_t a0, uint32_t a1,
                                                  uint32_t a2, uint32_t a3, uint32_t b0,
                                                  uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], uint32_t a0, uint32_t a1,
                                                         uint32_t a2, uint32_t a3, uint32_t b0,
                                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src, bool pred) {
  int sz = pred ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// RoPE frequency scaling: the reference bakes a `rope_scaling` dict into its PrimFuncs at build time
// (switch_rope_freq_func, position_embedding.py:257-299); here it is process state set by tvmb200_set_rope_scaling and
// copied into every launch.  kind 0 = rope_freq_default, 1 = rope_freq_llama3 (position_embedding.py:130-160).
struct RopeScaling {
  int kind;
  float inv_factor, alpha, beta;  // llama3: 1/factor, orig_max_pos / (2 pi (high - low)), low / (high - low)
};
RopeScaling rope_scaling();  // setting of the calling thread's current context (core.cu)

// The denominator the rotation angle is divided by: freq = pos * scale / rope_denominator(d).
//   default: theta^((2d mod rd)/rd)                                  (position_embedding.py:63)
//   llama3:  1 / (orig * ((1 - smooth) / factor + smooth)), orig = theta^-((2d mod rd)/rd),
//            smooth = clamp(alpha * orig - beta, 0, 1)                (position_embedding.py:143-153)
__device__ __forceinline__ float rope_denominator(int d, int rotary_dim, float theta, const RopeScaling& rs) {
  const float den = powf(theta, static_cast<float>((d * 2) % rotary_dim) / static_cast<float>(rotary_dim));
  if (rs.kind == 0) return den;
  const float orig = 1.0f / den;
  const float smooth = fmaxf(0.0f, fminf(1.0f, rs.alpha * orig - rs.beta));
  return 1.0f / ((1.0f - smooth) * orig * rs.inv_factor + smooth * orig);
}

// Element-wise RoPE variants of f_split_rotary that do not fit the "one angle per rotated pair" form above: kind 2 =
// rope_freq_gptj (interleaved pairs, position_embedding.py:70-76, :509-514), 3 = rope_freq_llama4 (:79-127), 5 =
// rope_freq_yarn (:223-254; the ramp runs over the element index, so the two halves of a pair get different angles).
// Only tvmb200_split_rotary[_append] implements them (split_rotary_variant_kernel); kind 0 = none active.  While one is
// active every in-kernel rotation (rotary_mode = 1 of the attention entries, the fused decode step) is rejected.
struct RopeVariant {
  int kind;
  float p0, p1, p2, p3;
  // gptj:   -
  // llama4: p0 = 1/factor (smooth) or factor (equal factors), p1 = alpha or the wavelength threshold, p2 = beta,
  //         p3 != 0: high_freq_factor == low_freq_factor (threshold branch)
  // yarn (set): p0 = factor, p1 = original_max_position_embeddings, p2 = beta_fast, p3 = beta_slow; at launch p1 / p2
  //         become (low, high - low) of the correction range for the launch's rotary_dim and theta
  float inv_theta_log_scale;  // yarn; 0 = 1 / (2 ln theta) taken at launch (kv_cache.py:355-366)
};
RopeVariant rope_variant();  // setting of the calling thread's current context (core.cu)

int check_no_rope_variant(const char* who);  // error (non-zero) when a variant is active (core.cu)

// host: the stored form of kind 2 / 3 from the numbers of the reference's rope_scaling dict (llama3-style arguments)
inline RopeVariant make_rope_variant(int kind, float factor, float low_freq_factor, float high_freq_factor,
                                     float original_max_position_embeddings) {
  RopeVariant rv = {kind, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (kind != 3) return rv;
  if (high_freq_factor == low_freq_factor) {  // threshold branch (position_embedding.py:99-109)
    rv.p0 = factor;
    rv.p1 = original_max_position_embeddings / low_freq_factor;
    rv.p3 = 1.f;
    return rv;
  }
  const double inv_diff = 1.0 / (static_cast<double>(high_freq_factor) - static_cast<double>(low_freq_factor));
  rv.p0 = static_cast<float>(1.0 / factor);
  rv.p1 = static_cast<float>(original_max_position_embeddings / (2.0 * 3.14159265358979323846) * inv_diff);
  rv.p2 = static_cast<float>(low_freq_factor * inv_diff);
  return rv;
}

// the rotation angle of element d in [0, rotary_dim) at scaled position s, float32 like the reference
__device__ __forceinline__ float rope_variant_angle(float s, int d, int rd, float theta, const RopeVariant& v) {
  if (v.kind == 2) {
    return s / powf(theta, static_cast<float>((2 * (d / 2)) % rd) / static_cast<float>(rd));
  } else if (v.kind == 3) {
    const float orig = 1.0f / powf(theta, static_cast<float>(2 * (d / 2)) / static_cast<float>(rd));
    if (v.p3 != 0.0f) {
      const float wavelength = 6.283185307179586f / orig;
      return s * (wavelength > v.p1 ? orig / v.p0 : orig);
    }
    const float smooth = fmaxf(0.0f, fminf(1.0f, v.p1 * orig - v.p2));
    return s * ((1.0f - smooth) * orig * v.p0 + smooth * orig);
  } else {
    const float den = powf(theta, static_cast<float>((d * 2) % rd) / static_cast<float>(rd));
    const float extra = 1.0f / den, inter = 1.0f / (v.p0 * den);
    const float ramp = (static_cast<float>(d) - v.p1) / v.p2;
    const float mask = 1.0f - fmaxf(fminf(ramp, 1.0f), 0.0f);
    return s * (inter * (1.0f - mask) + extra * mask);
  }
}

// cos * x + sin * partner with ONE fixed rounding sequence (product, then fused multiply-add), so that every kernel
// that rotates -- split_rotary, the inline-RoPE loads, the fused decode step -- produces the same bits
__device__ __forceinline__ float rope_mix(float c, float x, float s, float partner) {
  return __fmaf_rn(c, x, __fmul_rn(s, partner));
}

// block-wide exclusive scan helper for the device-side work schedulers.  vals in smem [n+1]:
// on entry s[i] (i<n) holds the count of item i; on exit s[i] = sum_{j<i}, s[n] = total.
// All threads of the block must call it.  tmp: smem scratch of >= 33 ints.
__device__ __forceinline__ void block_exclusive_scan(int* s, int n, int* tmp) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (n + nt - 1) / nt;
  const int beg = min(tid * per, n), end = min(beg + per, n);
  int sum = 0;
  for (int i = beg; i < end; ++i) sum += s[i];
  // scan of per-thread sums
  const int lane = tid & 31, warp = tid >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) tmp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int nw = (nt + 31) >> 5;
    int w = lane < nw ? tmp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += y;
    }
    tmp[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) tmp[32] = wi;
  }
  __syncthreads();
  int run = tmp[warp] + incl - sum;
  for (int i = beg; i < end; ++i) {
    int c = s[i];
    s[i] = run;
    run += c;
  }
  if (tid == 0) s[n] = tmp[32];
  __syncthreads();
}

// What the reference compiles into one kernel set + the device scratch its launches need (core.cu).
struct Context {
  struct Scratch {
    void* ws = nullptr;
    int64_t ws_bytes = 0;
    int32_t* counters = nullptr;
  };
  std::mutex mu_;
  RopeScaling rs = {0, 1.0f, 0.0f, 0.0f};
  RopeVariant rv = {0, 0.f, 0.f, 0.f, 0.f, 0.f};
  std::atomic<int32_t> layer_sws{1024};
  // what the 6-argument (reference) form of fused_rope uses (position_embedding.py:444-452 bakes them in)
  float rope_theta = 10000.0f, rope_scale = 1.0f;
  int rotary_dim = 0;
  std::map<std::pair<int, uintptr_t>, Scratch> scratch_;  // (device, stream) -> scratch
  std::atomic<int> refs{1};
  ~Context();
};
Context* default_context();
Context* current_context();                  // the scope entered on this thread, else the default context
Context* enter_context(Context* c);          // returns the previous scope (nullptr = default); not RAII
struct ContextScope {
  explicit ContextScope(Context* c);
  ~ContextScope();
  Context* prev_;
};
int context_set_rope_scaling(Context* c, int32_t kind, float factor, float low_freq_factor, float high_freq_factor,
                             float original_max_position_embeddings);
int context_set_rope_scaling_yarn(Context* c, float factor, float original_max_position_embeddings, float beta_fast,
                                  float beta_slow, float inv_theta_log_scale);
void context_copy_settings(Context* dst, Context* src);

}  // namespace tvmb200
