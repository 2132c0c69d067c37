// Gather / rotate pre-pass in front of the tcgen05 prefill kernel (prefill_tc05.cu).
//
// The tensor-core kernel wants K tiles it can hand to the MMA as they land: 128 consecutive KV POSITIONS, already
// rotated.  Two flavours of the reference's prefill callbacks do not offer that:
//   * rotary_mode = 1 (RoPEMode::kInline: the cache holds un-rotated K; _kernel_common.py:115-127 rotates q at
//     q_rope_position[row] and every K row at k_rope_pos_offset[b] + its position while loading) -- 128 x 64 sincosf
//     per K tile cannot hide behind ~500 clk of tensor work;
//   * the `_sliding_window` flavours ([3, B] length_info: position -> slot is pos < sink ? pos : pos - sink + offset,
//     _kernel_common.py:147-170), where 128 consecutive positions are not page-aligned rows any more.  Every cache
//     built with sliding-window support runs RoPE inline (paged_kv_cache.cc:343-344), so the two come together.
// This pass streams the sequence's K (and, for paged sources, V) once through HBM into position-ordered ragged
// scratch, rotating K on the way, and rotates q into scratch; the attention then runs on the ragged tcgen05 kernel
// with rotary_mode 0.  It costs one extra read + write of the visible KV (HBM-bound, a few per cent of the
// contraction it unlocks) instead of the mma.sync fallback (about a quarter of the tcgen05 throughput).
// One warp per row; the per-row (cos, sin) table is shared by all heads of the row (as split_rotary_warp_kernel).
#include "prefill.cuh"

namespace tvmb200 {

namespace {

constexpr int kD = 128;
constexpr int kWarps = 8;

// kv_len_b of every sequence of a paged batch (_kernel_common.py:155-159) -> exclusive scan (one CTA)
__global__ void __launch_bounds__(1024)
prepass_offsets_kernel(const int32_t* __restrict__ page_indptr, const int32_t* __restrict__ length_info, int batch,
                       int sliding, int32_t* __restrict__ kv_indptr_out) {
  extern __shared__ int s_len[];  // [batch + 1] + 40
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    const int np = page_indptr[b + 1] - page_indptr[b];
    int len = 0;
    if (np > 0) {
      len = (np - 1) * 16 + length_info[b];
      if (sliding) len += length_info[2 * batch + b] - length_info[batch + b];  // - window offset + sink
    }
    s_len[b] = len;
  }
  __syncthreads();
  block_exclusive_scan(s_len, batch, s_len + batch + 1);
  for (int b = threadIdx.x; b <= batch; b += blockDim.x) kv_indptr_out[b] = s_len[b];
}

template <typename T, bool PAGED>
__global__ void __launch_bounds__(kWarps * 32, 4)
prepass_rows_kernel(const PrepassParams a) {
  __shared__ float den[kD / 2];
  __shared__ float2 cs_all[kWarps][kD / 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* cs = cs_all[warp];
  if (a.rotary)
    for (int d = threadIdx.x; d < kD / 2; d += blockDim.x) den[d] = rope_denominator(d, kD, a.rope_theta, a.rs);
  __syncthreads();
  constexpr int RV = kD / 8, HV = kD / 16;  // 16-byte vectors per head row / per rotary half
  const int32_t* kv_indptr = PAGED ? a.kv_indptr_out : a.kv_indptr;
  const int64_t total_kv = kv_indptr[a.batch];
  const int64_t n_q = a.rotary ? a.n_q : 0;
  // rows [0, n_q): q; [n_q, n_q + total_kv): K (and V); then up to 128 rows of zeros behind the last V row, so that
  // the last KV tile of the last sequence multiplies P = 0 with zeros, not with whatever the scratch held
  const int64_t room = a.kv_rows_bound - total_kv;
  const int64_t pad = PAGED ? (room < 128 ? room : 128) : 0;
  const int64_t rows = n_q + total_kv + pad;
  auto table = [&](float pos) {
    for (int d = lane; d < kD / 2; d += 32) {
      float sn, c;
      sincosf(pos / den[d], &sn, &c);
      cs[d] = make_float2(c, sn);
    }
    __syncwarp();
  };
  // rotate `heads` head rows of D elements (head h at src + h * src_stride vectors) into consecutive rows of dst: one
  // (lower, upper) pair of 16-byte vectors per lane step
  auto rotate_rows = [&](const uint4* __restrict__ src, int64_t src_stride, uint4* __restrict__ dst, int heads) {
#pragma unroll 2
    for (int idx = lane; idx < heads * HV; idx += 32) {
      const int h = idx / HV, j = idx - h * HV;
      const uint4 lo = ldg_nc_v4(src + h * src_stride + j), hi = ldg_nc_v4(src + h * src_stride + j + HV);
      const T* le = reinterpret_cast<const T*>(&lo);
      const T* he = reinterpret_cast<const T*>(&hi);
      uint4 olo, ohi;
      T* ol = reinterpret_cast<T*>(&olo);
      T* oh = reinterpret_cast<T*>(&ohi);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 f = cs[j * 8 + e];
        ol[e] = DT<T>::from_f(rope_mix(f.x, DT<T>::to_f(le[e]), f.y, DT<T>::to_f(DT<T>::neg(he[e]))));
        oh[e] = DT<T>::from_f(rope_mix(f.x, DT<T>::to_f(he[e]), f.y, DT<T>::to_f(le[e])));
      }
      dst[h * RV + j] = olo;
      dst[h * RV + j + HV] = ohi;
    }
  };
  auto copy_rows = [&](const uint4* __restrict__ src, int64_t src_stride, uint4* __restrict__ dst, int heads) {
#pragma unroll 2
    for (int idx = lane; idx < heads * RV; idx += 32) {
      const int h = idx / RV, j = idx - h * RV;
      dst[idx] = ldg_nc_v4(src + h * src_stride + j);
    }
  };
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kWarps + warp; r < rows; r += static_cast<int64_t>(gridDim.x) * kWarps) {
    if (r < n_q) {
      table(static_cast<float>(a.q_rope_position[r]) * a.rope_scale);
      rotate_rows(static_cast<const uint4*>(a.q) + r * a.hq * RV, RV, static_cast<uint4*>(a.q_out) + r * a.hq * RV, a.hq);
      __syncwarp();
      continue;
    }
    const int64_t kr = r - n_q;
    uint4* k_dst = static_cast<uint4*>(a.k_out) + kr * a.hkv * RV;
    if (kr >= total_kv) {  // zero padding behind the last V row (paged sources only)
      uint4* v_dst = static_cast<uint4*>(a.v_out) + kr * a.hkv * RV;
      for (int idx = lane; idx < a.hkv * RV; idx += 32) v_dst[idx] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    int lo = 0, hi = a.batch;  // largest b with kv_indptr[b] <= kr
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(kv_indptr + mid) <= kr) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int i = static_cast<int>(kr - __ldg(kv_indptr + b));  // position within the sequence's visible KV
    if (a.rotary) {
      const int pos = a.tree_k_rope ? a.q_rope_position[kr] : a.k_rope_pos_offset[b] + i;
      table(static_cast<float>(pos) * a.rope_scale);
    }
    if (PAGED) {
      int slot = i;
      if (a.sliding) {
        const int sink = a.length_info[2 * a.batch + b], off = a.length_info[a.batch + b];
        slot = i < sink ? i : i - sink + off;
      }
      const int64_t pid = __ldg(a.page_values + a.page_indptr[b] + (slot >> 4));
      const uint4* pg = static_cast<const uint4*>(a.pages);
      const int64_t head_stride = 16 * RV;  // one (page, K|V, head) block
      uint4* v_dst = static_cast<uint4*>(a.v_out) + kr * a.hkv * RV;
      const uint4* k_src = pg + (pid * 2 + 0) * a.hkv * head_stride + (slot & 15) * RV;
      const uint4* v_src = pg + (pid * 2 + 1) * a.hkv * head_stride + (slot & 15) * RV;
      if (a.rotary) rotate_rows(k_src, head_stride, k_dst, a.hkv); else copy_rows(k_src, head_stride, k_dst, a.hkv);
      copy_rows(v_src, head_stride, v_dst, a.hkv);
    } else {
      rotate_rows(static_cast<const uint4*>(a.k) + kr * a.hkv * RV, RV, k_dst, a.hkv);  // ragged sources: rotary only
    }
    __syncwarp();  // the warp's (cos, sin) slice is rewritten for its next row
  }
}

}  // namespace

int64_t prepass_scratch_bytes(bool paged, bool rotary, int64_t n_q, int hq, int hkv, int64_t kv_rows_bound, int batch,
                              int64_t off[4]) {
  auto al = [](int64_t x) { return (x + 255) / 256 * 256; };
  int64_t at = 0;
  off[0] = at; at += al((static_cast<int64_t>(batch) + 1) * 4);                    // kv_indptr
  off[1] = at; at += rotary ? al(n_q * hq * kD * 2) : 0;                           // q
  off[2] = at; at += al(kv_rows_bound * hkv * kD * 2);                             // k
  off[3] = at; at += paged ? al(kv_rows_bound * hkv * kD * 2) : 0;                 // v
  return at;
}

int launch_prefill_prepass(const PrepassParams& a, bool paged, int dtype, cudaStream_t st) {
  if (paged) {
    const size_t smem = (static_cast<size_t>(a.batch) + 1 + 40) * sizeof(int);
    prepass_offsets_kernel<<<1, 1024, smem, st>>>(a.page_indptr, a.length_info, a.batch, a.sliding, a.kv_indptr_out);
    TVMB200_LAUNCH_OK();
  }
  const int64_t rows = (a.rotary ? a.n_q : 0) + a.kv_rows_bound;
  const int64_t want = (rows + kWarps - 1) / kWarps, cap = static_cast<int64_t>(num_sms()) * 4;
  const unsigned grid = static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
  if (dtype == TVMB200_F16) {
    if (paged) prepass_rows_kernel<__half, true><<<grid, kWarps * 32, 0, st>>>(a);
    else prepass_rows_kernel<__half, false><<<grid, kWarps * 32, 0, st>>>(a);
  } else {
    if (paged) prepass_rows_kernel<__nv_bfloat16, true><<<grid, kWarps * 32, 0, st>>>(a);
    else prepass_rows_kernel<__nv_bfloat16, false><<<grid, kWarps * 32, 0, st>>>(a);
  }
  TVMB200_LAUNCH_OK();
  return 0;
}

}  // namespace tvmb200
