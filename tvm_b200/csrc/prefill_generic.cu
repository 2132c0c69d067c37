// General batch prefill attention: ragged Q against paged KV (f_attention_prefill), ragged KV
// (f_attention_prefill_ragged) and the two token-tree variants, with causal / layer-sliding-window /
// tree masks, per-sequence sliding window + attention sinks, GQA and optional inline RoPE.
//
// Reference: python/tvm/relax/frontend/nn/llm/_prefill_kernels.py:54-391 (paged), :677-923 (ragged),
//            tree_attn.py:48-65 (tree mask), :68-603 (tree ragged), :606-1259 (tree paged),
//            _kernel_common.py:115-170 (inline rope, masks, kv length helpers).
//
// This is the shape-generic path (any ragged batch, D in {64,128}, every mask and rope flag): a
// FlashAttention-2 style kernel on the legacy tensor path (mma.sync m16n8k16, fp32 accumulate).  The
// D=128 / rotary_mode=0 hot configurations are served by the tcgen05/TMEM kernel in
// prefill_tc05.cu; this kernel is what every other configuration runs on.
//
//  * work item = (64-row Q tile, kv head).  Rows are GQA-folded exactly like the reference
//    (row = token * group + head_in_group, _prefill_kernels.py:318-324) so a K/V tile is shared
//    by the whole group.  Items are enumerated on the device from q_indptr by every CTA (block scan).
//  * K/V tiles of 64 tokens are gathered with 16-byte cp.async (page-table lookup per row for the
//    paged variants) into XOR-swizzled shared memory, double buffered.
//  * online softmax in the base-2 domain with the reference's -5e4 sentinel; LSE = m + log2(d).
#include "prefill.cuh"

#include <type_traits>

namespace tvmb200 {

constexpr int BM = 64;  // Q rows per CTA (4 warps x 16)
constexpr int BN = 64;  // KV tokens per tile

// physical byte offset of logical 16-byte chunk c of row r in a [rows][D] 16-bit tile
template <int D>
__device__ __forceinline__ uint32_t swz(int r, int c) {
  return static_cast<uint32_t>(r) * (D * 2) + static_cast<uint32_t>((c & ~7) | ((c ^ r) & 7)) * 16;
}

template <typename T, int D>
__device__ __forceinline__ void rope_rows_inplace(uint8_t* tile, int rows, const float* s_denom,
                                                  const int* s_pos, float rope_scale) {
  // item = (row, chunk pair): chunk j of the lower half pairs with chunk j + D/16
  constexpr int HC = D / 16;
  for (int it = threadIdx.x; it < rows * HC; it += blockDim.x) {
    const int r = it / HC, j = it - r * HC;
    uint4* plo = reinterpret_cast<uint4*>(tile + swz<D>(r, j));
    uint4* phi = reinterpret_cast<uint4*>(tile + swz<D>(r, j + HC));
    uint4 lo = *plo, hi = *phi;
    T* le = reinterpret_cast<T*>(&lo);
    T* he = reinterpret_cast<T*>(&hi);
    const float pos = static_cast<float>(s_pos[r]) * rope_scale;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float freq = pos / s_denom[j * 8 + e];
      float sn, cs;
      sincosf(freq, &sn, &cs);
      const float xl = DT<T>::to_f(le[e]), xh = DT<T>::to_f(he[e]);
      // _kernel_common.py:115-127: cos*x + sin*(d < rd/2 ? -x[d+rd/2] : x[d-rd/2])
      const float nl = cs * xl + sn * DT<T>::to_f(DT<T>::neg(he[e]));
      const float nh = cs * xh + sn * xl;
      le[e] = DT<T>::from_f(nl);
      he[e] = DT<T>::from_f(nh);
    }
    *plo = lo;
    *phi = hi;
  }
}

template <typename T, int D, bool PAGED>
__global__ void __launch_bounds__(128, 2)
prefill_generic_kernel(const PrefillParams p) {
  constexpr int KS = D / 16;      // k-steps of QK^T
  constexpr int NT_O = D / 8;     // n-tiles of O
  constexpr int CH = D / 8;       // 16-byte chunks per row
  constexpr int TILE_BYTES = BN * D * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;                         // BM x D
  uint8_t* sK = sQ + BM * D * 2;              // 2 stages
  uint8_t* sV = sK + 2 * TILE_BYTES;          // 2 stages
  int* s_tiles = reinterpret_cast<int*>(sV + 2 * TILE_BYTES);  // [B+1] scan
  int* s_tmp = s_tiles + p.batch + 1;                         // 40
  float* s_denom = reinterpret_cast<float*>(s_tmp + 40);      // [D/2] rope denominators
  int* s_pos = reinterpret_cast<int*>(s_denom + D / 2);       // [64] rope positions of the tile rows

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int B = p.batch, g = p.group;

  // ---- enumerate Q tiles -----------------------------------------------------------------------
  for (int b = tid; b < B; b += blockDim.x) {
    const int rows = (p.q_indptr[b + 1] - p.q_indptr[b]) * g;
    s_tiles[b] = (rows + BM - 1) / BM;
  }
  __syncthreads();
  block_exclusive_scan(s_tiles, B, s_tmp);
  const int n_items = s_tiles[B] * p.num_kv_heads;
  const int item = blockIdx.x;
  if (item >= n_items) return;
  // heavy (late) tiles first helps the causal tail; keep simple: reverse order
  const int ritem = n_items - 1 - item;
  const int tg = ritem / p.num_kv_heads;
  const int h = ritem - tg * p.num_kv_heads;
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_tiles[mid] <= tg) lo = mid; else hi = mid;
  }
  const int b = lo;
  const int tile = tg - s_tiles[b];
  const int q_beg = p.q_indptr[b];
  const int qo_len = p.q_indptr[b + 1] - q_beg;
  const int row0 = tile * BM;  // first folded row of this tile

  // ---- KV extent -------------------------------------------------------------------------------
  int kv_len, kv_beg = 0, pg_beg = 0, sw_off = 0, sink = 0;
  if (PAGED) {
    pg_beg = p.page_indptr[b];
    const int npg = p.page_indptr[b + 1] - pg_beg;
    if (npg == 0) {
      kv_len = 0;
    } else if (p.sliding) {
      sw_off = p.length_info[B + b];
      sink = p.length_info[2 * B + b];
      kv_len = (npg - 1) * 16 + p.length_info[b] - sw_off + sink;
    } else {
      kv_len = (npg - 1) * 16 + p.length_info[b];
    }
  } else {
    kv_beg = p.kv_indptr[b];
    kv_len = p.kv_indptr[b + 1] - kv_beg;
  }
  int tree_beg = 0, tree_len = 0;
  if (p.mask_mode == kMaskTree) {
    tree_beg = p.tree_indptr[b];
    tree_len = p.tree_indptr[b + 1] - tree_beg;
  }
  const int tree_start = kv_len - tree_len;
  // last query token of the tile bounds the causal extent
  const int qi_last = min(qo_len - 1, (row0 + BM - 1) / g);
  int kv_end = kv_len;
  if (p.mask_mode == kMaskCausal) kv_end = max(0, min(kv_len, kv_len - qo_len + qi_last + 1));
  const int n_kv_tiles = (kv_end + BN - 1) / BN;

  if (p.rotary_mode == 1) {
    for (int d = tid; d < D / 2; d += blockDim.x) s_denom[d] = rope_denominator(d, D, p.rope_theta, p.rs);
  }

  // ---- loaders -----------------------------------------------------------------------------------
  const int ld_c = tid & (CH - 1);        // chunk handled by this thread
  const int ld_r0 = tid / CH;             // first row; rows step by 128/CH
  constexpr int LD_RSTEP = 128 / CH;
  const T* qbase = static_cast<const T*>(p.q);
  auto load_q = [&]() {
#pragma unroll
    for (int r = ld_r0; r < BM; r += LD_RSTEP) {
      const int R = row0 + r;
      const int qi = R / g;
      const bool ok = qi < qo_len;
      const T* src = qbase + ((static_cast<int64_t>(q_beg + (ok ? qi : 0)) * p.num_qo_heads) + h * g + (R - qi * g)) * D + ld_c * 8;
      cp_async_16(smem_u32(sQ + swz<D>(r, ld_c)), src, ok);
    }
  };
  auto load_kv = [&](int t, int stage) {
    uint8_t* dk = sK + stage * TILE_BYTES;
    uint8_t* dv = sV + stage * TILE_BYTES;
#pragma unroll
    for (int r = ld_r0; r < BN; r += LD_RSTEP) {
      const int j = t * BN + r;
      const bool ok = j < kv_end;
      const T *srck, *srcv;
      if (PAGED) {
        const int slot = (j < sink) ? j : j - sink + sw_off;  // _get_seq_offset
        const int page = ok ? __ldg(p.page_values + pg_beg + (slot >> 4)) : 0;
        const int64_t rowk = ((static_cast<int64_t>(page) * 2) * p.num_kv_heads + h) * 16 + (slot & 15);
        srck = static_cast<const T*>(p.pages) + rowk * D + ld_c * 8;
        srcv = srck + static_cast<int64_t>(p.num_kv_heads) * 16 * D;
      } else {
        const int64_t row = (static_cast<int64_t>(kv_beg + (ok ? j : 0)) * p.num_kv_heads + h) * D + ld_c * 8;
        srck = static_cast<const T*>(p.k) + row;
        srcv = static_cast<const T*>(p.v) + row;
      }
      cp_async_16(smem_u32(dk + swz<D>(r, ld_c)), srck, ok);
      cp_async_16(smem_u32(dv + swz<D>(r, ld_c)), srcv, ok);
    }
  };

  load_q();
  if (n_kv_tiles > 0) load_kv(0, 0);
  cp_async_commit();

  // ---- per-thread row bookkeeping: rows (warp*16 + lane/4) and +8 -------------------------------
  const int r_a = warp * 16 + (lane >> 2);
  const int qi_a = (row0 + r_a) / g, qi_b = (row0 + r_a + 8) / g;
  int child_a0 = 0, child_a1 = 0, child_b0 = 0, child_b1 = 0;  // tree (order, end) of the rows
  (void)child_a1; (void)child_b1;
  if (p.mask_mode == kMaskTree) {
    const int ca = qi_a + tree_len - qo_len, cb = qi_b + tree_len - qo_len;
    if (qi_a < qo_len && ca >= 0) child_a0 = p.tree_order[(tree_beg + ca) * 2];
    if (qi_b < qo_len && cb >= 0) child_b0 = p.tree_order[(tree_beg + cb) * 2];
  }

  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_a = kNegInit, m_b = kNegInit, d_a = 0.f, d_b = 0.f;
  uint32_t qf[KS][4];

  auto visible = [&](int qi, int child_order, int j) -> bool {
    if (j >= kv_len) return false;
    switch (p.mask_mode) {
      case kMaskCausal:
        return j < kv_len - qo_len + qi + 1;
      case kMaskLayerSliding: {
        const int visible_past = max(p.layer_sws - qi - 1, 0);
        return j >= max(kv_len - visible_past, 0);
      }
      case kMaskTree: {
        if (j < tree_start) return true;
        const int2 par = *reinterpret_cast<const int2*>(p.tree_order + (tree_beg + (j - tree_start)) * 2);
        return child_order >= par.x && child_order < par.y;
      }
      default:
        return true;
    }
  };

  for (int t = 0; t < n_kv_tiles; ++t) {
    const int stage = t & 1;
    if (t + 1 < n_kv_tiles) load_kv(t + 1, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    if (p.rotary_mode == 1) {
      if (t == 0) {
        // rotate Q rows in place
        for (int r = tid; r < BM; r += blockDim.x) {
          const int qi = (row0 + r) / g;
          s_pos[r] = qi < qo_len ? p.q_rope_position[q_beg + qi] : 0;
        }
        __syncthreads();
        rope_rows_inplace<T, D>(sQ, BM, s_denom, s_pos, p.rope_scale);
        __syncthreads();
      }
      for (int r = tid; r < BN; r += blockDim.x) {
        const int j = t * BN + r;
        int pos = 0;
        if (j < kv_end) {
          if (!PAGED && p.tree_k_rope) pos = p.q_rope_position[kv_beg + j];
          else pos = p.k_rope_pos_offset[b] + j;
        }
        s_pos[r] = pos;
      }
      __syncthreads();
      rope_rows_inplace<T, D>(sK + stage * TILE_BYTES, BN, s_denom, s_pos, p.rope_scale);
      __syncthreads();
    }

    if (t == 0) {
      // Q fragments (A operand): rows warp*16.., 16x16 per k-step
      const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldmatrix_x4(smem_u32(sQ + swz<D>(row, ks * 2 + (lane >> 4))), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }

    const uint32_t kb = smem_u32(sK + stage * TILE_BYTES);
    const uint32_t vb = smem_u32(sV + stage * TILE_BYTES);

    // ---- S = Q K^T (16 x 64 per warp) -------------------------------------------------------------
    float s[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int jp = 0; jp < BN / 16; ++jp) {
        // x4: (ntile 2jp, k lo), (ntile 2jp, k hi), (ntile 2jp+1, k lo), (ntile 2jp+1, k hi)
        const int mi = lane >> 3;
        const int row = (jp * 2 + (mi >> 1)) * 8 + (lane & 7);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(kb + swz<D>(row, ks * 2 + (mi & 1)), b0, b1, b2, b3);
        mma_16816<T>(s[jp * 2], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_16816<T>(s[jp * 2 + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }

    // ---- mask + scale ----------------------------------------------------------------------------
    const int j0 = t * BN;
    bool need_mask = (j0 + BN > kv_len) || p.mask_mode == kMaskLayerSliding || p.mask_mode == kMaskTree;
    if (p.mask_mode == kMaskCausal) {
      const int qi_first = row0 / g;
      need_mask = need_mask || (j0 + BN > kv_len - qo_len + qi_first + 1);
    }
    float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) {
      const int jc = j0 + i * 8 + (lane & 3) * 2;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float va = s[i][e] * p.scale_log2, vb2 = s[i][2 + e] * p.scale_log2;
        if (need_mask) {
          if (!visible(qi_a, child_a0, jc + e)) va = -INFINITY;
          if (!visible(qi_b, child_b0, jc + e)) vb2 = -INFINITY;
        }
        s[i][e] = va;
        s[i][2 + e] = vb2;
        mx_a = fmaxf(mx_a, va);
        mx_b = fmaxf(mx_b, vb2);
      }
    }
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
    const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
    const float f_a = fast_exp2(m_a - mn_a), f_b = fast_exp2(m_b - mn_b);
    m_a = mn_a;
    m_b = mn_b;
    d_a *= f_a;
    d_b *= f_b;
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
      o[i][0] *= f_a; o[i][1] *= f_a; o[i][2] *= f_b; o[i][3] *= f_b;
    }
    // bf16 keeps only 8 mantissa bits of P; the reference keeps P in fp32.  Split P = hi + lo (two bf16
    // parts, ~16 bits) so that rows dominated by a few keys stay inside the 2e-3 parity bar.
    constexpr bool kSplitP = std::is_same<T, __nv_bfloat16>::value;
    uint32_t pf[BN / 16][4];
    uint32_t pl[kSplitP ? BN / 16 : 1][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) {
      const float p0 = fast_exp2(s[i][0] - m_a), p1 = fast_exp2(s[i][1] - m_a);
      const float p2 = fast_exp2(s[i][2] - m_b), p3 = fast_exp2(s[i][3] - m_b);
      d_a += p0 + p1;
      d_b += p2 + p3;
      const uint32_t h01 = DT<T>::pack(p0, p1), h23 = DT<T>::pack(p2, p3);
      pf[i >> 1][(i & 1) * 2 + 0] = h01;
      pf[i >> 1][(i & 1) * 2 + 1] = h23;
      if (kSplitP) {
        const float2 f01 = DT<T>::to_f2(h01), f23 = DT<T>::to_f2(h23);
        pl[i >> 1][(i & 1) * 2 + 0] = DT<T>::pack(p0 - f01.x, p1 - f01.y);
        pl[i >> 1][(i & 1) * 2 + 1] = DT<T>::pack(p2 - f23.x, p3 - f23.y);
      }
    }

    // ---- O += P V ----------------------------------------------------------------------------------
#pragma unroll
    for (int kk = 0; kk < BN / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < NT_O / 2; ++np) {
        // x4.trans: (tok lo, d chunk 2np), (tok hi, 2np), (tok lo, 2np+1), (tok hi, 2np+1)
        const int mi = lane >> 3;
        const int row = kk * 16 + (mi & 1) * 8 + (lane & 7);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(vb + swz<D>(row, np * 2 + (mi >> 1)), b0, b1, b2, b3);
        mma_16816<T>(o[np * 2], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b0, b1);
        mma_16816<T>(o[np * 2 + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b2, b3);
        if (kSplitP) {
          mma_16816<T>(o[np * 2], pl[kk][0], pl[kk][1], pl[kk][2], pl[kk][3], b0, b1);
          mma_16816<T>(o[np * 2 + 1], pl[kk][0], pl[kk][1], pl[kk][2], pl[kk][3], b2, b3);
        }
      }
    }
    __syncthreads();  // everyone is done with this stage before it is refilled
  }
  cp_async_wait<0>();

  // ---- epilogue --------------------------------------------------------------------------------------
  d_a += __shfl_xor_sync(0xffffffffu, d_a, 1);
  d_a += __shfl_xor_sync(0xffffffffu, d_a, 2);
  d_b += __shfl_xor_sync(0xffffffffu, d_b, 1);
  d_b += __shfl_xor_sync(0xffffffffu, d_b, 2);
  const float inv_a = d_a > 0.f ? 1.f / d_a : 0.f, inv_b = d_b > 0.f ? 1.f / d_b : 0.f;
  T* obase = static_cast<T*>(p.output);
  if (qi_a < qo_len) {
    const int hq = h * g + (row0 + r_a) - qi_a * g;
    T* dst = obase + (static_cast<int64_t>(q_beg + qi_a) * p.num_qo_heads + hq) * D + (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < NT_O; ++i)
      *reinterpret_cast<uint32_t*>(dst + i * 8) = DT<T>::pack(o[i][0] * inv_a, o[i][1] * inv_a);
    if ((lane & 3) == 0)
      p.lse[static_cast<int64_t>(q_beg + qi_a) * p.num_qo_heads + hq] = d_a > 0.f ? m_a + log2f(d_a) : kNegInit;
  }
  if (qi_b < qo_len) {
    const int hq = h * g + (row0 + r_a + 8) - qi_b * g;
    T* dst = obase + (static_cast<int64_t>(q_beg + qi_b) * p.num_qo_heads + hq) * D + (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < NT_O; ++i)
      *reinterpret_cast<uint32_t*>(dst + i * 8) = DT<T>::pack(o[i][2] * inv_b, o[i][3] * inv_b);
    if ((lane & 3) == 0)
      p.lse[static_cast<int64_t>(q_beg + qi_b) * p.num_qo_heads + hq] = d_b > 0.f ? m_b + log2f(d_b) : kNegInit;
  }
}

template <typename T, int D, bool PAGED>
static int launch_generic(const PrefillParams& p, int total_q_len, cudaStream_t st) {
  const int64_t max_tiles = (static_cast<int64_t>(total_q_len) * p.group + BM - 1) / BM + p.batch;
  const int64_t grid = max_tiles * p.num_kv_heads;
  const size_t smem = static_cast<size_t>(BM) * D * 2 + 4ull * BN * D * 2 +
                      (static_cast<size_t>(p.batch) + 1 + 40) * 4 + (D / 2) * 4 + 64 * 4;
  auto kern = prefill_generic_kernel<T, D, PAGED>;
  TVMB200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<static_cast<unsigned>(grid), 128, smem, st>>>(p);
  TVMB200_LAUNCH_OK();
  return 0;
}

int launch_prefill_generic(const PrefillParams& p, bool paged, int total_q_len, int head_dim, int dtype,
                           cudaStream_t st) {
  if (dtype == TVMB200_F16) {
    if (head_dim == 128) return paged ? launch_generic<__half, 128, true>(p, total_q_len, st) : launch_generic<__half, 128, false>(p, total_q_len, st);
    return paged ? launch_generic<__half, 64, true>(p, total_q_len, st) : launch_generic<__half, 64, false>(p, total_q_len, st);
  }
  if (head_dim == 128) return paged ? launch_generic<__nv_bfloat16, 128, true>(p, total_q_len, st) : launch_generic<__nv_bfloat16, 128, false>(p, total_q_len, st);
  return paged ? launch_generic<__nv_bfloat16, 64, true>(p, total_q_len, st) : launch_generic<__nv_bfloat16, 64, false>(p, total_q_len, st);
}

}  // namespace tvmb200
