// Page-cache data movement kernels: append, debug-get, page fork copy, tree-commit compaction,
// fused QKV split + RoPE, and the in-place LSE merge.  All HBM-bound: 128-bit accesses, streaming
// cache hints, bit-exact copies.
//
// Reference (TIR) counterparts:
//   transpose_append   python/tvm/relax/frontend/nn/llm/_page_kernels.py:40-74
//   debug_get_kv       _page_kernels.py:106-136
//   copy_single_page   _page_kernels.py:169-189
//   compact_kv_copy    _page_kernels.py:235-263
//   split_rotary       python/tvm/relax/frontend/nn/llm/position_embedding.py:444-565
//   merge_state_inplace python/tvm/relax/frontend/nn/llm/_decode_kernels.py:414-526
#include "common.cuh"

namespace tvmb200 {

// one thread = one 16-byte vector of K and the matching vector of V
__global__ void __launch_bounds__(256)
transpose_append_kernel(uint4* __restrict__ pages, const uint4* __restrict__ k,
                        const uint4* __restrict__ v, const int32_t* __restrict__ position_map,
                        int64_t total_vecs, int num_kv_heads, int page_size, int row_vecs) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total_vecs) return;
  const int per_token = num_kv_heads * row_vecs;
  const int64_t t = idx / per_token;
  const int r = static_cast<int>(idx - t * per_token);
  const int h = r / row_vecs;
  const int c = r - h * row_vecs;
  const int pos = __ldg(position_map + t);
  if (pos < 0) return;  // -1 = "do not append" (reference: position_map != -1)
  const int64_t page = pos / page_size;
  const int slot = pos - static_cast<int>(page) * page_size;
  const int64_t dst_k = (((page * 2 + 0) * num_kv_heads + h) * page_size + slot) * row_vecs + c;
  const int64_t dst_v = dst_k + static_cast<int64_t>(num_kv_heads) * page_size * row_vecs;
  const uint4 kv = ldg_nc_v4(k + idx);
  const uint4 vv = ldg_nc_v4(v + idx);
  pages[dst_k] = kv;
  pages[dst_v] = vv;
}

__global__ void __launch_bounds__(256)
debug_get_kv_kernel(const uint4* __restrict__ pages, const int32_t* __restrict__ position_map,
                    uint4* __restrict__ k_out, uint4* __restrict__ v_out, int64_t total_vecs,
                    int num_kv_heads, int page_size, int row_vecs) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total_vecs) return;
  const int per_token = num_kv_heads * row_vecs;
  const int64_t t = idx / per_token;
  const int r = static_cast<int>(idx - t * per_token);
  const int h = r / row_vecs;
  const int c = r - h * row_vecs;
  const int pos = __ldg(position_map + t);
  const int64_t page = pos / page_size;
  const int slot = pos - static_cast<int>(page) * page_size;
  const int64_t src_k = (((page * 2 + 0) * num_kv_heads + h) * page_size + slot) * row_vecs + c;
  const int64_t src_v = src_k + static_cast<int64_t>(num_kv_heads) * page_size * row_vecs;
  k_out[idx] = pages[src_k];
  v_out[idx] = pages[src_v];
}

// pages[tgt, kv, h, 0:copy_length, :] = pages[src, kv, h, 0:copy_length, :]
__global__ void __launch_bounds__(256)
copy_single_page_kernel(uint4* __restrict__ pages, int64_t src_page, int64_t tgt_page,
                        int copy_length, int num_kv_heads, int page_size, int row_vecs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_head = copy_length * row_vecs;
  const int total = 2 * num_kv_heads * per_head;
  if (idx >= total) return;
  const int kvh = idx / per_head;  // (kv, head) flattened
  const int r = idx - kvh * per_head;
  const int64_t in_page = static_cast<int64_t>(kvh) * page_size * row_vecs + r;
  const int64_t page_vecs = static_cast<int64_t>(2) * num_kv_heads * page_size * row_vecs;
  pages[tgt_page * page_vecs + in_page] = pages[src_page * page_vecs + in_page];
}

// thread = (sequence b, kv head h, 16-byte column c); walks that sequence's copy list IN ORDER
// (a destination may be a later source: the serial order of the reference must be kept).
__global__ void __launch_bounds__(256)
compact_kv_copy_kernel(uint4* __restrict__ pages, const int32_t* __restrict__ indptr,
                       const int32_t* __restrict__ src_dst, int batch_size, int total_copy_length,
                       int num_kv_heads, int page_size, int row_vecs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_seq = num_kv_heads * row_vecs;
  if (idx >= batch_size * per_seq) return;
  const int b = idx / per_seq;
  const int r = idx - b * per_seq;
  const int h = r / row_vecs;
  const int c = r - h * row_vecs;
  const int beg = indptr[b], end = indptr[b + 1];
  const int64_t v_off = static_cast<int64_t>(num_kv_heads) * page_size * row_vecs;
  for (int i = beg; i < end; ++i) {
    const int sp = src_dst[i];
    const int dp = src_dst[total_copy_length + i];
    const int64_t spage = sp / page_size, dpage = dp / page_size;
    const int sslot = sp - static_cast<int>(spage) * page_size;
    const int dslot = dp - static_cast<int>(dpage) * page_size;
    const int64_t s = (((spage * 2) * num_kv_heads + h) * page_size + sslot) * row_vecs + c;
    const int64_t d = (((dpage * 2) * num_kv_heads + h) * page_size + dslot) * row_vecs + c;
    const uint4 kk = pages[s];
    const uint4 vv = pages[s + v_off];
    pages[d] = kk;
    pages[d + v_off] = vv;
  }
}

// ---------------------------------------------------------------------------------------------
// fused QKV split + RoPE.  One CTA per token: the D/2 (cos, sin) pairs of the token's position are
// computed ONCE into shared memory (the reference recomputes powf/cosf/sinf per element), then
// every (head, 8-element vector pair) is rotated with 128-bit loads/stores.
// ---------------------------------------------------------------------------------------------
// APPEND = true additionally scatters the (rotated) k and v rows of token t into the paged cache at slot
// append_pos[t] (>= 0), i.e. f_split_rotary followed by f_transpose_append in one launch; the values written to
// the pages are the very registers written to k / v, so the result is bit-identical to the two-kernel sequence.
template <typename T, bool APPEND>
__global__ void __launch_bounds__(256)
split_rotary_kernel(const uint4* __restrict__ qkv, const int32_t* __restrict__ position_map,
                    uint4* __restrict__ q, uint4* __restrict__ k, uint4* __restrict__ v,
                    int num_qo_heads, int num_kv_heads, int head_dim, int rotary_dim,
                    int apply_rope, float rope_scale, float rope_theta, uint4* __restrict__ pages,
                    const int32_t* __restrict__ append_pos, int page_size, const RopeScaling rs) {
  extern __shared__ float2 cs[];  // [rotary_dim/2] (cos, sin)
  pdl_launch_dependents();  // a dependent launched programmatically (the decode kernel) may start its prologue now
  const int64_t t = blockIdx.x;
  const int row_vecs = head_dim / 8;
  const int half_vecs = rotary_dim / 16;  // vectors in one rotary half
  const int fused_heads = num_qo_heads + 2 * num_kv_heads;
  if (apply_rope > 0) {
    const float pos = static_cast<float>(position_map[t]) * rope_scale;
    for (int d = threadIdx.x; d < rotary_dim / 2; d += blockDim.x) {
      const float freq = pos / rope_denominator(d, rotary_dim, rope_theta, rs);
      float s, c;
      sincosf(freq, &s, &c);
      cs[d] = make_float2(c, s);
    }
    __syncthreads();
  }
  const uint4* src = qkv + t * fused_heads * row_vecs;
  // work item = (head, vector index j in [0, row_vecs)); rotated heads pair j with j +- half_vecs
  for (int w = threadIdx.x; w < fused_heads * row_vecs; w += blockDim.x) {
    const int h = w / row_vecs;
    const int j = w - h * row_vecs;
    uint4* dst;
    uint4* dst_page = nullptr;
    if (h < num_qo_heads) {
      dst = q + (t * num_qo_heads + h) * row_vecs + j;
    } else {
      const int is_v = h >= num_qo_heads + num_kv_heads;
      const int hh = h - num_qo_heads - is_v * num_kv_heads;
      dst = (is_v ? v : k) + (t * num_kv_heads + hh) * row_vecs + j;
      if (APPEND) {
        const int32_t slot = append_pos[t];
        if (slot >= 0) {
          const int64_t pg = slot / page_size, off = slot - pg * page_size;
          dst_page = pages + (((pg * 2 + is_v) * num_kv_heads + hh) * page_size + off) * row_vecs + j;
        }
      }
    }
    uint4 x = ldg_nc_v4(src + w);
    const bool rot = apply_rope > 0 && h < num_qo_heads + num_kv_heads && j < 2 * half_vecs;
    if (rot) {
      const bool lower = j < half_vecs;
      const uint4 p = ldg_nc_v4(src + (lower ? w + half_vecs : w - half_vecs));
      const int d0 = (lower ? j : j - half_vecs) * 8;  // frequency index of element 0
      const T* xe = reinterpret_cast<const T*>(&x);
      const T* pe = reinterpret_cast<const T*>(&p);
      uint4 o;
      T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 f = cs[d0 + e];
        // reference: cos*x + sin*(d < rd/2 ? -x[d+rd/2] : x[d-rd/2]), negation done in dtype
        const float partner = DT<T>::to_f(lower ? DT<T>::neg(pe[e]) : pe[e]);
        oe[e] = DT<T>::from_f(rope_mix(f.x, DT<T>::to_f(xe[e]), f.y, partner));
      }
      x = o;
    }
    *dst = x;
    if (APPEND && dst_page != nullptr) *dst_page = x;
  }
}

// Large token counts (prefill chunks): ONE WARP per token, tokens grid-strided over a grid sized to the machine.  The
// per-CTA kernel above spends most of a CTA's life in its serial prologue (64 powf + sincosf on two warps, a barrier,
// then three vectors per thread): 3.6 TB/s on 32768 tokens.  Here the frequency denominators are computed once per
// CTA, a warp computes its token's (cos, sin) table into its own shared-memory slice (no CTA barrier in the loop), and
// every lane rotates whole (lower, upper) vector pairs, so each input vector is loaded exactly once.  The arithmetic
// per element is the very same sequence as above (pos / den -> sincosf -> rope_mix): outputs are bit-identical.
template <typename T, bool APPEND>
__global__ void __launch_bounds__(256, 4)
split_rotary_warp_kernel(const uint4* __restrict__ qkv, const int32_t* __restrict__ position_map,
                         uint4* __restrict__ q, uint4* __restrict__ k, uint4* __restrict__ v, int64_t ntoken,
                         int num_qo_heads, int num_kv_heads, int head_dim, int rotary_dim,
                         int apply_rope, float rope_scale, float rope_theta, uint4* __restrict__ pages,
                         const int32_t* __restrict__ append_pos, int page_size, const RopeScaling rs) {
  extern __shared__ float sm_rot[];  // den[rotary_dim/2] | per warp: float2 cs[rotary_dim/2]
  const int nfreq = rotary_dim / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* den = sm_rot;
  float2* cs = reinterpret_cast<float2*>(sm_rot + nfreq) + warp * nfreq;
  pdl_launch_dependents();
  if (apply_rope > 0)
    for (int d = threadIdx.x; d < nfreq; d += blockDim.x) den[d] = rope_denominator(d, rotary_dim, rope_theta, rs);
  __syncthreads();
  const int row_vecs = head_dim / 8;
  const int half_vecs = apply_rope > 0 ? rotary_dim / 16 : 0;  // vectors in one rotary half (0: plain split)
  const int rot_heads = num_qo_heads + num_kv_heads;
  const int fused_heads = rot_heads + num_kv_heads;
  const int tail_vecs = row_vecs - 2 * half_vecs;              // unrotated vectors of a q / k head
  const int warps_per_cta = blockDim.x >> 5;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * warps_per_cta + warp; t < ntoken;
       t += static_cast<int64_t>(gridDim.x) * warps_per_cta) {
    if (apply_rope > 0) {
      const float pos = static_cast<float>(position_map[t]) * rope_scale;
      for (int d = lane; d < nfreq; d += 32) {
        float sn, c;
        sincosf(pos / den[d], &sn, &c);
        cs[d] = make_float2(c, sn);
      }
      __syncwarp();
    }
    const uint4* src = qkv + t * fused_heads * row_vecs;
    uint4* page_k = nullptr;  // row of (page, K, head 0, slot) -- head hh adds hh * page_size * row_vecs
    int64_t v_off = 0;
    if (APPEND) {
      const int32_t slot = append_pos[t];
      if (slot >= 0) {
        const int64_t pg = slot / page_size, off = slot - pg * page_size;
        page_k = pages + ((pg * 2 * num_kv_heads) * page_size + off) * row_vecs;
        v_off = static_cast<int64_t>(num_kv_heads) * page_size * row_vecs;
      }
    }
    const int64_t head_stride = static_cast<int64_t>(page_size) * row_vecs;
    // rotated (lower, upper) vector pairs of the q and k heads
#pragma unroll 2
    for (int idx = lane; idx < rot_heads * half_vecs; idx += 32) {
      const int h = idx / half_vecs, j = idx - h * half_vecs;
      const uint4 lo = ldg_nc_v4(src + h * row_vecs + j);
      const uint4 hi = ldg_nc_v4(src + h * row_vecs + j + half_vecs);
      const T* le = reinterpret_cast<const T*>(&lo);
      const T* he = reinterpret_cast<const T*>(&hi);
      uint4 olo, ohi;
      T* ol = reinterpret_cast<T*>(&olo);
      T* oh = reinterpret_cast<T*>(&ohi);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 f = cs[j * 8 + e];
        // reference: cos*x + sin*(d < rd/2 ? -x[d+rd/2] : x[d-rd/2]), negation done in dtype
        ol[e] = DT<T>::from_f(rope_mix(f.x, DT<T>::to_f(le[e]), f.y, DT<T>::to_f(DT<T>::neg(he[e]))));
        oh[e] = DT<T>::from_f(rope_mix(f.x, DT<T>::to_f(he[e]), f.y, DT<T>::to_f(le[e])));
      }
      if (h < num_qo_heads) {
        uint4* dst = q + (t * num_qo_heads + h) * row_vecs + j;
        dst[0] = olo;
        dst[half_vecs] = ohi;
      } else {
        const int hh = h - num_qo_heads;
        uint4* dst = k + (t * num_kv_heads + hh) * row_vecs + j;
        dst[0] = olo;
        dst[half_vecs] = ohi;
        if (APPEND && page_k != nullptr) {
          uint4* pd = page_k + hh * head_stride + j;
          pd[0] = olo;
          pd[half_vecs] = ohi;
        }
      }
    }
    // unrotated vectors of the q / k heads (rotary_dim < head_dim, or no rotation at all)
    for (int idx = lane; idx < rot_heads * tail_vecs; idx += 32) {
      const int h = idx / tail_vecs, j = 2 * half_vecs + idx - h * tail_vecs;
      const uint4 x = ldg_nc_v4(src + h * row_vecs + j);
      if (h < num_qo_heads) {
        q[(t * num_qo_heads + h) * row_vecs + j] = x;
      } else {
        const int hh = h - num_qo_heads;
        k[(t * num_kv_heads + hh) * row_vecs + j] = x;
        if (APPEND && page_k != nullptr) page_k[hh * head_stride + j] = x;
      }
    }
    // v heads: a copy
#pragma unroll 2
    for (int idx = lane; idx < num_kv_heads * row_vecs; idx += 32) {
      const int hh = idx / row_vecs, j = idx - hh * row_vecs;
      const uint4 x = ldg_nc_v4(src + (rot_heads + hh) * row_vecs + j);
      v[(t * num_kv_heads + hh) * row_vecs + j] = x;
      if (APPEND && page_k != nullptr) page_k[v_off + hh * head_stride + j] = x;
    }
    __syncwarp();  // the warp's (cos, sin) slice is rewritten for its next token
  }
}

constexpr int64_t kWarpPerTokenMin = 1024;  // below this the one-CTA-per-token kernel has the shorter critical path

template <typename T, bool APPEND>
static void launch_split_rotary_warp(const void* qkv, const int32_t* position_map, void* q, void* k, void* v, int64_t ntoken,
                                     int num_qo_heads, int num_kv_heads, int head_dim, int rotary_dim, int apply,
                                     float rope_scale, float rope_theta, void* pages, const int32_t* append_pos,
                                     int page_size, cudaStream_t st) {
  const int warps = 8;
  const size_t smem = static_cast<size_t>(rotary_dim / 2) * sizeof(float) * (1 + 2 * warps);
  const int64_t want = (ntoken + warps - 1) / warps, cap = static_cast<int64_t>(num_sms()) * 4;  // one resident wave
  split_rotary_warp_kernel<T, APPEND><<<static_cast<unsigned>(want < cap ? want : cap), warps * 32, smem, st>>>(
      static_cast<const uint4*>(qkv), position_map, static_cast<uint4*>(q), static_cast<uint4*>(k), static_cast<uint4*>(v),
      ntoken, num_qo_heads, num_kv_heads, head_dim, rotary_dim, apply, rope_scale, rope_theta, static_cast<uint4*>(pages),
      append_pos, page_size, rope_scaling());
}

// f_split_rotary for the element-wise RoPE variants (gptj / llama4 / yarn, common.cuh RopeVariant): same CTA-per-token
// layout, but the table holds one (cos, sin) per ELEMENT of the rotary range -- the two halves of a pair may rotate by
// different angles (llama4, yarn) -- and gptj pairs neighbouring elements (2i, 2i+1) instead of the two halves.
template <typename T>
__global__ void __launch_bounds__(256)
split_rotary_variant_kernel(const uint4* __restrict__ qkv, const int32_t* __restrict__ position_map,
                            uint4* __restrict__ q, uint4* __restrict__ k, uint4* __restrict__ v, int num_qo_heads,
                            int num_kv_heads, int head_dim, int rotary_dim, float rope_scale, float rope_theta,
                            const RopeVariant rv) {
  extern __shared__ float2 cs[];  // [rotary_dim] (cos, sin)
  const int64_t t = blockIdx.x;
  const int row_vecs = head_dim / 8;
  const int half_vecs = rotary_dim / 16;
  const int fused_heads = num_qo_heads + 2 * num_kv_heads;
  const float pos = static_cast<float>(position_map[t]) * rope_scale;
  for (int d = threadIdx.x; d < rotary_dim; d += blockDim.x) {
    float sn, c;
    sincosf(rope_variant_angle(pos, d, rotary_dim, rope_theta, rv), &sn, &c);
    cs[d] = make_float2(c, sn);
  }
  __syncthreads();
  const uint4* src = qkv + t * fused_heads * row_vecs;
  const bool interleaved = rv.kind == 2;
  for (int w = threadIdx.x; w < fused_heads * row_vecs; w += blockDim.x) {
    const int h = w / row_vecs;
    const int j = w - h * row_vecs;
    uint4* dst;
    if (h < num_qo_heads) {
      dst = q + (t * num_qo_heads + h) * row_vecs + j;
    } else {
      const int is_v = h >= num_qo_heads + num_kv_heads;
      const int hh = h - num_qo_heads - is_v * num_kv_heads;
      dst = (is_v ? v : k) + (t * num_kv_heads + hh) * row_vecs + j;
    }
    uint4 x = ldg_nc_v4(src + w);
    if (h < num_qo_heads + num_kv_heads && j < 2 * half_vecs) {
      const bool lower = j < half_vecs;
      const T* xe = reinterpret_cast<const T*>(&x);
      uint4 p = x;
      if (!interleaved) p = ldg_nc_v4(src + (lower ? w + half_vecs : w - half_vecs));
      const T* pe = reinterpret_cast<const T*>(&p);
      uint4 o;
      T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 f = cs[j * 8 + e];
        float partner;
        if (interleaved)  // d even: -x[d + 1], d odd: x[d - 1]; negation done in dtype like the reference
          partner = DT<T>::to_f((e & 1) ? xe[e - 1] : DT<T>::neg(xe[e + 1]));
        else
          partner = DT<T>::to_f(lower ? DT<T>::neg(pe[e]) : pe[e]);
        oe[e] = DT<T>::from_f(rope_mix(f.x, DT<T>::to_f(xe[e]), f.y, partner));
      }
      x = o;
    }
    *dst = x;
  }
}

// launch-time form of the process-wide variant: the yarn correction range depends on rotary_dim and theta
static RopeVariant variant_for_launch(RopeVariant rv, int rotary_dim, float rope_theta) {
  if (rv.kind == 5) {
    const double itls = rv.inv_theta_log_scale > 0.f ? static_cast<double>(rv.inv_theta_log_scale)
                                                     : 1.0 / (2.0 * std::log(static_cast<double>(rope_theta)));
    const double two_pi = 2.0 * 3.14159265358979323846;
    double low = rotary_dim * std::log(rv.p1 / (rv.p2 * two_pi)) * itls;   // yarn_find_correction_range, :192-220
    double high = rotary_dim * std::log(rv.p1 / (rv.p3 * two_pi)) * itls;
    low = std::max(low, 0.0);
    high = std::min(high, static_cast<double>(rotary_dim - 1));
    if (low == high) high += 0.001;
    rv.p1 = static_cast<float>(low);
    rv.p2 = static_cast<float>(high - low);
  }
  return rv;
}

static int launch_split_rotary_variant(const void* qkv, const int32_t* position_map, void* q, void* k, void* v,
                                       int64_t ntoken, int32_t num_qo_heads, int32_t num_kv_heads, int32_t head_dim,
                                       int32_t rotary_dim, float rope_scale, float rope_theta, int dtype, cudaStream_t st) {
  const RopeVariant rv = variant_for_launch(rope_variant(), rotary_dim, rope_theta);
  const size_t smem = static_cast<size_t>(rotary_dim) * sizeof(float2);
  if (dtype == TVMB200_F16) {
    split_rotary_variant_kernel<__half><<<static_cast<unsigned>(ntoken), 256, smem, st>>>(
        static_cast<const uint4*>(qkv), position_map, static_cast<uint4*>(q), static_cast<uint4*>(k),
        static_cast<uint4*>(v), num_qo_heads, num_kv_heads, head_dim, rotary_dim, rope_scale, rope_theta, rv);
  } else {
    split_rotary_variant_kernel<__nv_bfloat16><<<static_cast<unsigned>(ntoken), 256, smem, st>>>(
        static_cast<const uint4*>(qkv), position_map, static_cast<uint4*>(q), static_cast<uint4*>(k),
        static_cast<uint4*>(v), num_qo_heads, num_kv_heads, head_dim, rotary_dim, rope_scale, rope_theta, rv);
  }
  TVMB200_LAUNCH_OK();
  return 0;
}

// (V,S) <- merge((V,S),(V',S')).  A block owns whole rows (row = n*H+h): thread = (row, 8-element
// vector); S[row] is rewritten by the row's vector-0 thread after a block barrier so every thread of
// the row has read the old value first.
template <typename T>
__global__ void __launch_bounds__(256)
merge_state_inplace_kernel(uint4* __restrict__ v, float* __restrict__ s,
                           const uint4* __restrict__ v_other, const float* __restrict__ s_other,
                           int64_t rows, int row_vecs, int rows_per_block) {
  const int r_in = threadIdx.x / row_vecs;
  const int c = threadIdx.x - r_in * row_vecs;
  const int64_t row = blockIdx.x * static_cast<int64_t>(rows_per_block) + r_in;
  const bool active = r_in < rows_per_block && row < rows;
  float s_max = 0.f, a = 0.f, b = 0.f;
  if (active) {
    const float s_val = s[row], so_val = s_other[row];
    s_max = fmaxf(s_val, so_val);
    a = exp2f(s_val - s_max);
    b = exp2f(so_val - s_max);
    const float scale = a / (a + b);
    const float other_scale = b / (a + b);
    const int64_t idx = row * row_vecs + c;
    uint4 x = v[idx];
    const uint4 y = ldg_nc_v4(v_other + idx);
    T* xe = reinterpret_cast<T*>(&x);
    const T* ye = reinterpret_cast<const T*>(&y);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      xe[e] = DT<T>::from_f(DT<T>::to_f(xe[e]) * scale + DT<T>::to_f(ye[e]) * other_scale);
    v[idx] = x;
  }
  __syncthreads();
  if (active && c == 0) s[row] = log2f(a + b) + s_max;
}

}  // namespace tvmb200

using namespace tvmb200;

static inline int elem_vec_check(int head_dim) { return head_dim % 8 == 0; }

extern "C" int tvmb200_transpose_append(void* pages, const void* k, const void* v,
                                        const int32_t* position_map, int64_t ntoken,
                                        int64_t num_pages, int32_t num_kv_heads, int32_t page_size,
                                        int32_t head_dim, int dtype, tvmb200_stream_t stream) {
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "transpose_append: unsupported dtype %d", dtype);
  TVMB200_CHECK(elem_vec_check(head_dim), "transpose_append: head_dim %d must be a multiple of 8", head_dim);
  TVMB200_CHECK(ntoken >= 0 && num_pages >= 0 && num_kv_heads > 0 && page_size > 0, "transpose_append: bad shape");
  if (ntoken == 0) return 0;
  const int row_vecs = head_dim / 8;
  const int64_t total = ntoken * num_kv_heads * row_vecs;
  const int64_t blocks = (total + 255) / 256;
  transpose_append_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint4*>(pages), static_cast<const uint4*>(k), static_cast<const uint4*>(v),
      position_map, total, num_kv_heads, page_size, row_vecs);
  TVMB200_LAUNCH_OK();
  return 0;
}

extern "C" int tvmb200_debug_get_kv(const void* pages, const int32_t* position_map, void* k_out,
                                    void* v_out, int64_t layer_id, int64_t num_layers,
                                    int64_t seqlen, int64_t num_pages, int32_t num_kv_heads,
                                    int32_t page_size, int32_t head_dim, int dtype,
                                    tvmb200_stream_t stream) {
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "debug_get_kv: unsupported dtype %d", dtype);
  TVMB200_CHECK(elem_vec_check(head_dim), "debug_get_kv: head_dim %d must be a multiple of 8", head_dim);
  TVMB200_CHECK(layer_id >= 0 && layer_id < num_layers, "debug_get_kv: layer_id %ld out of range [0,%ld)", (long)layer_id, (long)num_layers);
  if (seqlen == 0) return 0;
  const int row_vecs = head_dim / 8;
  const int64_t total = seqlen * num_kv_heads * row_vecs;
  const int64_t blocks = (total + 255) / 256;
  uint4* ko = static_cast<uint4*>(k_out) + layer_id * total;
  uint4* vo = static_cast<uint4*>(v_out) + layer_id * total;
  debug_get_kv_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(pages), position_map, ko, vo, total, num_kv_heads, page_size, row_vecs);
  TVMB200_LAUNCH_OK();
  return 0;
}

extern "C" int tvmb200_copy_single_page(void* pages, int64_t src_page_id, int64_t tgt_page_id,
                                        int64_t copy_length, int64_t num_pages,
                                        int32_t num_kv_heads, int32_t page_size, int32_t head_dim,
                                        int dtype, tvmb200_stream_t stream) {
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "copy_single_page: unsupported dtype %d", dtype);
  TVMB200_CHECK(elem_vec_check(head_dim), "copy_single_page: head_dim %d must be a multiple of 8", head_dim);
  TVMB200_CHECK(src_page_id >= 0 && src_page_id < num_pages && tgt_page_id >= 0 && tgt_page_id < num_pages,
                "copy_single_page: page id out of range (src %ld, tgt %ld, num_pages %ld)", (long)src_page_id, (long)tgt_page_id, (long)num_pages);
  TVMB200_CHECK(copy_length >= 0 && copy_length <= page_size, "copy_single_page: copy_length %ld exceeds page_size %d", (long)copy_length, page_size);
  if (copy_length == 0) return 0;
  const int row_vecs = head_dim / 8;
  const int total = 2 * num_kv_heads * static_cast<int>(copy_length) * row_vecs;
  copy_single_page_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint4*>(pages), src_page_id, tgt_page_id, static_cast<int>(copy_length),
      num_kv_heads, page_size, row_vecs);
  TVMB200_LAUNCH_OK();
  return 0;
}

extern "C" int tvmb200_compact_kv_copy(void* pages, const int32_t* copy_length_indptr,
                                       const int32_t* copy_src_dst_pos, int32_t batch_size,
                                       int32_t total_copy_length, int64_t num_pages,
                                       int32_t num_kv_heads, int32_t page_size, int32_t head_dim,
                                       int dtype, tvmb200_stream_t stream) {
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "compact_kv_copy: unsupported dtype %d", dtype);
  TVMB200_CHECK(elem_vec_check(head_dim), "compact_kv_copy: head_dim %d must be a multiple of 8", head_dim);
  if (batch_size <= 0 || total_copy_length <= 0) return 0;
  const int row_vecs = head_dim / 8;
  const int total = batch_size * num_kv_heads * row_vecs;
  compact_kv_copy_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint4*>(pages), copy_length_indptr, copy_src_dst_pos, batch_size,
      total_copy_length, num_kv_heads, page_size, row_vecs);
  TVMB200_LAUNCH_OK();
  return 0;
}

extern "C" int tvmb200_split_rotary(const void* qkv, const int32_t* position_map, void* q, void* k,
                                    void* v, int64_t ntoken, int32_t num_qo_heads,
                                    int32_t num_kv_heads, int32_t head_dim, int32_t rotary_dim,
                                    int64_t apply_rope, float rope_scale, float rope_theta,
                                    int dtype, tvmb200_stream_t stream) {
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "split_rotary: unsupported dtype %d", dtype);
  if (rotary_dim <= 0) rotary_dim = head_dim;
  TVMB200_CHECK(head_dim % 8 == 0 && rotary_dim % 16 == 0 && rotary_dim <= head_dim,
                "split_rotary: head_dim %d / rotary_dim %d unsupported (need D %% 8 == 0, rd %% 16 == 0)", head_dim, rotary_dim);
  if (ntoken == 0) return 0;
  const size_t smem = static_cast<size_t>(rotary_dim / 2) * sizeof(float2);
  const int apply = apply_rope > 0 ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (apply && rope_variant().kind != 0)
    return launch_split_rotary_variant(qkv, position_map, q, k, v, ntoken, num_qo_heads, num_kv_heads, head_dim, rotary_dim,
                                       rope_scale, rope_theta, dtype, st);
  if (ntoken >= kWarpPerTokenMin) {
    if (dtype == TVMB200_F16)
      launch_split_rotary_warp<__half, false>(qkv, position_map, q, k, v, ntoken, num_qo_heads, num_kv_heads, head_dim,
                                              rotary_dim, apply, rope_scale, rope_theta, nullptr, nullptr, 16, st);
    else
      launch_split_rotary_warp<__nv_bfloat16, false>(qkv, position_map, q, k, v, ntoken, num_qo_heads, num_kv_heads, head_dim,
                                                     rotary_dim, apply, rope_scale, rope_theta, nullptr, nullptr, 16, st);
    TVMB200_LAUNCH_OK();
    return 0;
  }
  if (dtype == TVMB200_F16) {
    split_rotary_kernel<__half, false><<<static_cast<unsigned>(ntoken), 256, smem, st>>>(
        static_cast<const uint4*>(qkv), position_map, static_cast<uint4*>(q), static_cast<uint4*>(k),
        static_cast<uint4*>(v), num_qo_heads, num_kv_heads, head_dim, rotary_dim, apply, rope_scale, rope_theta,
        nullptr, nullptr, 16, rope_scaling());
  } else {
    split_rotary_kernel<__nv_bfloat16, false><<<static_cast<unsigned>(ntoken), 256, smem, st>>>(
        static_cast<const uint4*>(qkv), position_map, static_cast<uint4*>(q), static_cast<uint4*>(k),
        static_cast<uint4*>(v), num_qo_heads, num_kv_heads, head_dim, rotary_dim, apply, rope_scale, rope_theta,
        nullptr, nullptr, 16, rope_scaling());
  }
  TVMB200_LAUNCH_OK();
  return 0;
}

extern "C" int tvmb200_split_rotary_append(const void* qkv, const int32_t* q_rope_position_map,
                                           const int32_t* append_position_map, void* q, void* k, void* v,
                                           void* pages, int64_t ntoken, int64_t num_pages, int32_t num_qo_heads,
                                           int32_t num_kv_heads, int32_t page_size, int32_t head_dim,
                                           int32_t rotary_dim, int64_t apply_rope, float rope_scale,
                                           float rope_theta, int dtype, tvmb200_stream_t stream) {
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "split_rotary_append: unsupported dtype %d", dtype);
  if (rotary_dim <= 0) rotary_dim = head_dim;
  TVMB200_CHECK(head_dim % 8 == 0 && rotary_dim % 16 == 0 && rotary_dim <= head_dim,
                "split_rotary_append: head_dim %d / rotary_dim %d unsupported (need D %% 8 == 0, rd %% 16 == 0)", head_dim, rotary_dim);
  TVMB200_CHECK(page_size > 0 && num_pages >= 0, "split_rotary_append: bad page geometry (%d slots, %ld pages)", page_size, (long)num_pages);
  if (ntoken == 0) return 0;
  const size_t smem = static_cast<size_t>(rotary_dim / 2) * sizeof(float2);
  const int apply = apply_rope > 0 ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (apply && rope_variant().kind != 0) {  // gptj / llama4 / yarn: the variant kernel, then the plain append
    if (int rc = launch_split_rotary_variant(qkv, q_rope_position_map, q, k, v, ntoken, num_qo_heads, num_kv_heads, head_dim,
                                             rotary_dim, rope_scale, rope_theta, dtype, st))
      return rc;
    return tvmb200_transpose_append(pages, k, v, append_position_map, ntoken, num_pages, num_kv_heads, page_size, head_dim,
                                    dtype, stream);
  }
  if (ntoken >= kWarpPerTokenMin) {
    if (dtype == TVMB200_F16)
      launch_split_rotary_warp<__half, true>(qkv, q_rope_position_map, q, k, v, ntoken, num_qo_heads, num_kv_heads, head_dim,
                                             rotary_dim, apply, rope_scale, rope_theta, pages, append_position_map, page_size, st);
    else
      launch_split_rotary_warp<__nv_bfloat16, true>(qkv, q_rope_position_map, q, k, v, ntoken, num_qo_heads, num_kv_heads,
                                                    head_dim, rotary_dim, apply, rope_scale, rope_theta, pages,
                                                    append_position_map, page_size, st);
    TVMB200_LAUNCH_OK();
    return 0;
  }
  if (dtype == TVMB200_F16) {
    split_rotary_kernel<__half, true><<<static_cast<unsigned>(ntoken), 256, smem, st>>>(
        static_cast<const uint4*>(qkv), q_rope_position_map, static_cast<uint4*>(q), static_cast<uint4*>(k),
        static_cast<uint4*>(v), num_qo_heads, num_kv_heads, head_dim, rotary_dim, apply, rope_scale, rope_theta,
        static_cast<uint4*>(pages), append_position_map, page_size, rope_scaling());
  } else {
    split_rotary_kernel<__nv_bfloat16, true><<<static_cast<unsigned>(ntoken), 256, smem, st>>>(
        static_cast<const uint4*>(qkv), q_rope_position_map, static_cast<uint4*>(q), static_cast<uint4*>(k),
        static_cast<uint4*>(v), num_qo_heads, num_kv_heads, head_dim, rotary_dim, apply, rope_scale, rope_theta,
        static_cast<uint4*>(pages), append_position_map, page_size, rope_scaling());
  }
  TVMB200_LAUNCH_OK();
  return 0;
}

extern "C" int tvmb200_merge_state_inplace(void* v, float* s, const void* v_other,
                                           const float* s_other, int64_t n, int32_t num_heads,
                                           int32_t head_dim, int dtype, tvmb200_stream_t stream) {
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "merge_state_inplace: unsupported dtype %d", dtype);
  TVMB200_CHECK(elem_vec_check(head_dim), "merge_state_inplace: head_dim %d must be a multiple of 8", head_dim);
  if (n == 0) return 0;
  const int row_vecs = head_dim / 8;
  TVMB200_CHECK(row_vecs <= 256, "merge_state_inplace: head_dim %d too large", head_dim);
  const int64_t rows = n * num_heads;
  const int rows_per_block = 256 / row_vecs;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block);
  if (dtype == TVMB200_F16) {
    merge_state_inplace_kernel<__half><<<blocks, 256, 0, st>>>(
        static_cast<uint4*>(v), s, static_cast<const uint4*>(v_other), s_other, rows, row_vecs, rows_per_block);
  } else {
    merge_state_inplace_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        static_cast<uint4*>(v), s, static_cast<const uint4*>(v_other), s_other, rows, row_vecs, rows_per_block);
  }
  TVMB200_LAUNCH_OK();
  return 0;
}
