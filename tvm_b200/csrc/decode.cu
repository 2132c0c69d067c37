// Batch decode attention over the paged KV cache (f_attention_decode).
//
// Reference: python/tvm/relax/frontend/nn/llm/_decode_kernels.py:49-178 (CPU semantics),
//            :181-411 (reference GPU schedule: one 512-thread CTA per (seq, kv-head), 8-byte loads,
//            no split-KV).
//
// B200 design (HBM-bound; every KV byte is read from DRAM exactly once):
//  * split-KV by a BALANCED CONTIGUOUS PARTITION: the (sequence, kv head, page) triples, in that order, form one
//    line of nnz_pages * Hkv page-heads; CTA k of a persistent grid (2 CTAs per SM) owns the k-th `quota` of it, so
//    every CTA streams the same number of bytes whatever the mix of sequence lengths, and a (sequence, head)
//    segment is cut only where a CTA boundary falls inside it (2-4 work items per CTA instead of the 8-9 of the
//    first version's fixed-size chunks: the per-item prologue / epilogue is what kept that one at 0.87-0.95 of the
//    copy peak depending on the shape).  The walk is done on the DEVICE from page_indptr (the callback only gets
//    device arrays): one binary search per CTA, then consecutive segments.
//  * inside an item each WARP owns an independent TMA pipeline: lane 0 issues
//    cp.async.bulk.tensor (128B-swizzled boxes of 16 slots x 64 elements; one page/head K block is
//    4 KiB contiguous in HBM) into the warp's private ring of NSTAGE stages and waits on the warp's
//    own mbarriers -- there is no CTA-wide barrier in the main loop.
//  * math runs on the legacy tensor path with the KV tokens as the MMA M dimension:
//    S^T[16 tok x 8 q] = K[16 x D] . Q^T, online softmax over the token axis with warp shuffles,
//    O^T[D x 8 q] += V^T[D x 16] . P^T, so the whole GQA group (<= 8 query heads) shares every K/V
//    byte and the FMA/cvt pressure of a SIMT kernel (which would make this kernel issue-bound on
//    B200) disappears.  For bf16, P is split into hi+lo bf16 parts so the PV product
//    keeps ~16 bits of P (the reference keeps P in fp32).
//  * the warps of an item merge (m, d, O) through shared memory; an item that covers its whole (sequence, head)
//    segment writes O/LSE directly, the others write fp32 partials into the CTA's two slots, which
//    decode_merge_kernel reduces (base-2 LSE merge, same arithmetic as f_merge_inplace).
#include <cstdlib>

#include "common.cuh"

#include <mutex>
#include <type_traits>
#include <unordered_map>

namespace tvmb200 {

constexpr int kMaxBatchSmem = 8192;  // page_indptr scan lives in shared memory (sized per launch)

struct DecodeParams {
  const void* q;               // [B, Hq, D]
  const int32_t* page_indptr;  // [B+1]
  const int32_t* page_values;  // [nnz]
  const int32_t* length_info;  // [B] or [3,B]
  const int32_t* k_rope_pos_offset;
  const int32_t* q_rope_position;
  void* output;        // [B, Hq, D]
  float* lse;          // [B, Hq]
  float* part_o;       // [2 * grid, group, D] fp32 (normalised partial outputs), slot = 2 * cta + (continues ? 1 : 0)
  float* part_lse;     // [2 * grid, group]
  int batch;
  int num_qo_heads;
  int num_kv_heads;    // VIRTUAL kv heads = kv_heads_real * vsplit: what the partition, the items and the merge see
  int kv_heads_real;   // heads of the page pool / the fused qkv tensor
  int vsplit;          // GQA groups above 8 query heads per kv head are cut into `vsplit` virtual heads of 8 (the MMA holds 8
                       // query heads per KV pass); virtual head v reads the pages of kv head v / vsplit -- a second pass
                       // over the same pages, normally out of L2 (its item sits next to the first on the line)
  int group;  // Hq / Hkv
  // Balanced split-KV plan: the (sequence, kv head, page) triples in that order form one line of `total` page-heads
  // (sequence b, head h starts at page_indptr[b] * Hkv + h * np_b); CTA k owns [k * quota, (k + 1) * quota) -- every
  // CTA streams the same number of bytes, and a (sequence, head) segment is cut only at CTA boundaries.  A piece that
  // does not cover its whole segment writes an fp32 partial to one of the CTA's two slots: 2k when the segment began
  // in an earlier CTA and ends here, 2k + 1 when it continues into the next CTA (a CTA has at most one of each).
  int64_t total;       // nnz_pages * num_kv_heads
  int quota;           // page-heads per CTA
  // fused split_rotary + transpose_append + decode (FUSED instantiation): q / new k / new v are read from the fused
  // qkv tensor, rotated in the kernel, and the new token is written to its page slot by the item that owns that page
  const void* qkv;              // [B, Hq + 2 Hkv, D]
  const int32_t* append_slot;   // [B] slot id (page * 16 + offset) of the new token = the sequence's last slot, or -1
  void* pages;                  // [P, 2, Hkv, 16, D]
  int fused_apply_rope;
  int sliding;  // length_info is [3,B]
  int rotary_mode;
  float rope_scale;
  float rope_theta;
  RopeScaling rs;
  float scale_log2;  // sm_scale * log2(e)
};

// shared-memory plan per CTA (dynamic):
//   [0, NW*NSTAGE*STAGE_BYTES)   K/V stages, 1024-aligned, per warp
//   then: mbarriers (NW*NSTAGE * 8 B), page_indptr copy (batch+1 ints), 40 spare ints
template <int D>
struct DecodeCfg {
  static constexpr int kPage = 16;
  static constexpr int kBlockBytes = kPage * D * 2;        // one K (or V) block of a page/head
  static constexpr int kStageBytes = 2 * kBlockBytes;      // K + V
  static constexpr int kHalves = D / 64;                   // 128-byte column groups per row
  static constexpr int kHalfBytes = kPage * 128;           // 2 KiB per TMA box
};

template <typename T, int D, int NW, int NSTAGE, bool ROPE, bool FUSED>
__global__ void __launch_bounds__(NW * 32)
decode_kernel(const __grid_constant__ CUtensorMap tmap, const DecodeParams p) {
  using Cfg = DecodeCfg<D>;
  constexpr int KS = D / 16;  // k-steps of QK^T == m-tiles of O^T
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16-byte aligned: align by hand
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t stages_base = smem_base + warp * NSTAGE * Cfg::kStageBytes;
  const uint32_t bars_base = smem_base + NW * NSTAGE * Cfg::kStageBytes;
  int* s_chunk_off = reinterpret_cast<int*>(smem_gen + NW * NSTAGE * Cfg::kStageBytes + NW * NSTAGE * 8);
  int* s_scan_tmp = s_chunk_off + (p.batch + 1);
  // inline-RoPE scratch (rotary_mode == 1 only): D/2 frequency denominators + the rotated Q group
  float* s_denom = reinterpret_cast<float*>(
      (reinterpret_cast<uintptr_t>(s_scan_tmp + 40) + 15) & ~static_cast<uintptr_t>(15));
  T* s_q = reinterpret_cast<T*>(s_denom + D / 2);
  auto bar_addr = [&](int w, int s) -> uint32_t { return bars_base + (w * NSTAGE + s) * 8; };

  // ---- setup: barriers, work enumeration -----------------------------------------------------
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < NW * NSTAGE; ++i) mbar_init(bars_base + i * 8, 1);
    mbar_fence_init();
  }
  // Programmatic dependent launch: when launched with the stream-serialization attribute this grid may start while
  // the previous kernel of the stream (the fused rotary + append) drains; nothing it produced -- or that it still
  // reads -- is touched above this line.  Without the attribute the wait returns at once.
  pdl_wait();
  const int B = p.batch;
  int* s_indptr = s_chunk_off;  // page_indptr staged in shared memory (the item walk below reads it repeatedly)
  for (int b = threadIdx.x; b <= B; b += blockDim.x) s_indptr[b] = p.page_indptr[b];
  if (ROPE || FUSED)
    for (int d = threadIdx.x; d < D / 2; d += blockDim.x) s_denom[d] = rope_denominator(d, D, p.rope_theta, p.rs);
  __syncthreads();
  if (blockIdx.x == 0) {
    // sequences without pages belong to no CTA's range: the empty result (O = 0, lse = the -5e4 sentinel)
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      if (s_indptr[b + 1] != s_indptr[b]) continue;
      for (int i = 0; i < p.num_qo_heads * D; ++i)
        static_cast<T*>(p.output)[static_cast<int64_t>(b) * p.num_qo_heads * D + i] = DT<T>::from_f(0.f);
      for (int i = 0; i < p.num_qo_heads; ++i) p.lse[static_cast<int64_t>(b) * p.num_qo_heads + i] = kNegInit;
    }
  }
  // (the line's true length comes from page_indptr; the host's nnz_pages only sized the grid)
  const int64_t total = static_cast<int64_t>(s_indptr[B]) * p.num_kv_heads < p.total
                            ? static_cast<int64_t>(s_indptr[B]) * p.num_kv_heads : p.total;
  const int64_t lin_beg = static_cast<int64_t>(blockIdx.x) * p.quota;
  const int64_t lin_end = lin_beg + p.quota < total ? lin_beg + p.quota : total;

  // per-warp pipeline bookkeeping: number of loads issued / consumed so far (monotonic across items,
  // so mbarrier phase parity is (count / NSTAGE) & 1)
  uint32_t n_consumed = 0, n_issued = 0;

  const int g = p.group;
  const int qrow = lane >> 2;           // query head within the group held by this lane (B-frag n)
  const int qc0 = (lane & 3) * 2;       // S^T / O^T column pair (query heads qc0, qc0+1)

  // first segment of this CTA's range: the last sequence b with page_indptr[b] * Hkv <= lin_beg (it has pages: a
  // sequence without pages shares its start with its successor)
  int b = 0;
  {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (static_cast<int64_t>(s_indptr[mid]) * p.num_kv_heads <= lin_beg) lo = mid; else hi = mid;
    }
    b = lo;
  }
  int h = 0, pg0 = 0;
  if (lin_beg < lin_end) {
    const int np = s_indptr[b + 1] - s_indptr[b];
    const int off = static_cast<int>(lin_beg - static_cast<int64_t>(s_indptr[b]) * p.num_kv_heads);
    h = off / np;
    pg0 = off - h * np;
  }
  for (int64_t lin = lin_beg; lin < lin_end;) {
    const int pg_beg_seq = s_indptr[b];
    const int n_pages_seq = s_indptr[b + 1] - pg_beg_seq;
    const int64_t room = lin_end - lin;
    const int pg1 = n_pages_seq - pg0 <= room ? n_pages_seq : pg0 + static_cast<int>(room);
    const bool whole = pg0 == 0 && pg1 == n_pages_seq;   // the segment lies inside this CTA's range: final result
    const int slot = 2 * blockIdx.x + (pg1 < n_pages_seq ? 1 : 0);

    // sequence length bookkeeping (_kernel_common.py:155-170)
    int last_page_len, sw_off = 0, sink = 0;
    if (p.sliding) {
      last_page_len = p.length_info[b];
      sw_off = p.length_info[B + b];
      sink = p.length_info[2 * B + b];
    } else {
      last_page_len = p.length_info[b];
    }
    const int total_slots = n_pages_seq > 0 ? (n_pages_seq - 1) * Cfg::kPage + last_page_len : 0;
    // valid slot s (slot = index into the sequence's page list * 16 + offset):
    //   s < sink   or   sw_off <= s < total_slots          (non-sliding: sink = sw_off = 0)
    // (kv_len = total_slots - sw_off + sink; position -> slot is pos<sink ? pos : pos-sink+sw_off)

    // ---- Q^T fragments (B operand, [k = d][n = q head]) -----------------------------------------
    uint32_t qf[KS][2];
    if (ROPE || FUSED) {
      // rotate the group's Q rows at q_rope_position[b] into shared memory (_kernel_common.py:115-127); FUSED reads
      // them from the fused qkv tensor like f_split_rotary (position_embedding.py:444-565) and skips the rotation
      // when the cache's RoPE mode is "none"
      const float qpos = static_cast<float>(p.q_rope_position[b]) * p.rope_scale;
      const T* qg = FUSED ? static_cast<const T*>(p.qkv) +
                                (static_cast<int64_t>(b) * (p.num_qo_heads + 2 * p.kv_heads_real) + h * g) * D
                          : static_cast<const T*>(p.q) + (static_cast<int64_t>(b) * p.num_qo_heads + h * g) * D;
      const bool rotate = !FUSED || p.fused_apply_rope;
      for (int it = threadIdx.x; it < g * (D / 2); it += blockDim.x) {
        const int qh = it / (D / 2), d = it - qh * (D / 2);
        const T xl = qg[qh * D + d], xh = qg[qh * D + d + D / 2];
        if (rotate) {
          float sn, cs;
          sincosf(qpos / s_denom[d], &sn, &cs);
          s_q[qh * D + d] = DT<T>::from_f(rope_mix(cs, DT<T>::to_f(xl), sn, DT<T>::to_f(DT<T>::neg(xh))));
          s_q[qh * D + d + D / 2] = DT<T>::from_f(rope_mix(cs, DT<T>::to_f(xh), sn, DT<T>::to_f(xl)));
        } else {
          s_q[qh * D + d] = xl;
          s_q[qh * D + d + D / 2] = xh;
        }
      }
      __syncthreads();
      const bool qv = qrow < g;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int d0 = ks * 16 + (lane & 3) * 2;
        qf[ks][0] = qv ? *reinterpret_cast<const uint32_t*>(s_q + qrow * D + d0) : 0u;
        qf[ks][1] = qv ? *reinterpret_cast<const uint32_t*>(s_q + qrow * D + d0 + 8) : 0u;
      }
    } else {
      const T* qp = static_cast<const T*>(p.q) + (static_cast<int64_t>(b) * p.num_qo_heads + h * g + qrow) * D;
      const bool qv = qrow < g;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int d0 = ks * 16 + (lane & 3) * 2;
        qf[ks][0] = qv ? *reinterpret_cast<const uint32_t*>(qp + d0) : 0u;
        qf[ks][1] = qv ? *reinterpret_cast<const uint32_t*>(qp + d0 + 8) : 0u;
      }
    }

    float o[KS][4];
#pragma unroll
    for (int i = 0; i < KS; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = kNegInit, m1 = kNegInit;  // running max of columns qc0, qc0+1 (warp-uniform per column)
    float d0s = 0.f, d1s = 0.f;          // per-lane partial denominators

    // pages of this warp: pg0 + warp, pg0 + warp + NW, ...
    const int my_n = (pg1 - pg0 - warp + NW - 1) / NW > 0 ? (pg1 - pg0 - warp + NW - 1) / NW : 0;
    const int32_t* my_pages = p.page_values + pg_beg_seq + pg0 + warp;

    int ids_cache = 0;  // page ids of this warp's pages [32*batch_k, 32*batch_k+32), one per lane
    int ids_batch = -1;
    auto page_id_of = [&](int j) -> int {  // warp-uniform j
      const int bk = j >> 5;
      if (bk != ids_batch) {
        const int jj = bk * 32 + lane;
        ids_cache = jj < my_n ? __ldg(my_pages + static_cast<int64_t>(jj) * NW) : 0;
        ids_batch = bk;
      }
      return __shfl_sync(0xffffffffu, ids_cache, j & 31);
    };
    auto issue = [&](int j) {  // all lanes call (shuffle inside); lane 0 issues the TMA
      const int pid = page_id_of(j);
      const uint32_t st = n_issued % NSTAGE;
      if (lane == 0) {
        const uint32_t bar = bar_addr(warp, st);
        const uint32_t dst = stages_base + st * Cfg::kStageBytes;
        mbar_expect_tx(bar, Cfg::kStageBytes);
        const int row_k = ((pid * 2 + 0) * p.kv_heads_real + h / p.vsplit) * Cfg::kPage;
        const int row_v = ((pid * 2 + 1) * p.kv_heads_real + h / p.vsplit) * Cfg::kPage;
#pragma unroll
        for (int hf = 0; hf < Cfg::kHalves; ++hf) {
          tma_load_2d(dst + hf * Cfg::kHalfBytes, &tmap, hf * 64, row_k, bar, kEvictFirst);
          tma_load_2d(dst + Cfg::kBlockBytes + hf * Cfg::kHalfBytes, &tmap, hf * 64, row_v, bar, kEvictFirst);
        }
      }
      ++n_issued;
    };

    int issued_here = 0;
    for (; issued_here < my_n && issued_here < NSTAGE; ++issued_here) issue(issued_here);

    for (int j = 0; j < my_n; ++j) {
      const uint32_t st = n_consumed % NSTAGE;
      mbar_wait(bar_addr(warp, st), (n_consumed / NSTAGE) & 1);
      const uint32_t kb = stages_base + st * Cfg::kStageBytes;
      const uint32_t vb = kb + Cfg::kBlockBytes;

      if (FUSED) {
        // f_transpose_append for this (sequence, kv head): the new token is the sequence's last slot.  The warp that
        // holds the last page rotates the new k, drops k and v into their row of the staged page (so this step's
        // attention sees them) and writes them to the page in HBM (so the next steps do).  D = 128: 4 elements / lane.
        if (pg0 + warp + j * NW == n_pages_seq - 1 && p.append_slot[b] >= 0) {
          const int r = last_page_len - 1;
          const int pid = page_id_of(j);
          // the documented precondition: the append slot IS the sequence's last slot (BeginForward's layout for a
          // decode step).  Anything else would silently attend to a stale row: fail loudly instead.
          if (p.append_slot[b] != pid * Cfg::kPage + r) __trap();
          const T* kn = static_cast<const T*>(p.qkv) +
                        (static_cast<int64_t>(b) * (p.num_qo_heads + 2 * p.kv_heads_real) + p.num_qo_heads + h / p.vsplit) * D;
          const T* vn = kn + static_cast<int64_t>(p.kv_heads_real) * D;
          const int e0 = lane * 4;
          const bool lower = e0 < D / 2;
          uint2 kx = *reinterpret_cast<const uint2*>(kn + e0);
          const uint2 kp = *reinterpret_cast<const uint2*>(kn + (lower ? e0 + D / 2 : e0 - D / 2));
          const uint2 vx = *reinterpret_cast<const uint2*>(vn + e0);
          if (p.fused_apply_rope) {
            const float pos = static_cast<float>(p.q_rope_position[b]) * p.rope_scale;
            T* xe = reinterpret_cast<T*>(&kx);
            const T* pe = reinterpret_cast<const T*>(&kp);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float sn, cs;
              sincosf(pos / s_denom[(e0 + e) & (D / 2 - 1)], &sn, &cs);
              const float partner = DT<T>::to_f(lower ? DT<T>::neg(pe[e]) : pe[e]);
              xe[e] = DT<T>::from_f(rope_mix(cs, DT<T>::to_f(xe[e]), sn, partner));
            }
          }
          const int c = lane >> 1;  // 16-byte chunk of the row
          const uint32_t off = (c >> 3) * Cfg::kHalfBytes + r * 128 + (((c & 7) ^ (r & 7)) << 4) + (lane & 1) * 8;
          asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(kb + off), "r"(kx.x), "r"(kx.y) : "memory");
          asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(vb + off), "r"(vx.x), "r"(vx.y) : "memory");
          T* pg = static_cast<T*>(p.pages);
          const int64_t row_k = ((static_cast<int64_t>(pid) * 2 + 0) * p.kv_heads_real + h / p.vsplit) * Cfg::kPage + r;
          const int64_t row_v = ((static_cast<int64_t>(pid) * 2 + 1) * p.kv_heads_real + h / p.vsplit) * Cfg::kPage + r;
          *reinterpret_cast<uint2*>(pg + row_k * D + e0) = kx;
          *reinterpret_cast<uint2*>(pg + row_v * D + e0) = vx;
          fence_proxy_async();  // generic-proxy writes before the next TMA refill of this stage
          __syncwarp();
        }
      }

      // slot validity of this page
      const int slot0 = (pg0 + warp + j * NW) * Cfg::kPage;
      uint32_t vmask;  // bit t = slot0 + t is a live KV entry
      {
        const int hi_end = min(max(total_slots - slot0, 0), 16);
        const int lo_beg = min(max(sw_off - slot0, 0), 16);
        const int sink_end = min(max(sink - slot0, 0), 16);
        const uint32_t window = ((1u << hi_end) - 1u) & ~((1u << lo_beg) - 1u);
        const uint32_t sinkm = ((1u << min(sink_end, hi_end)) - 1u);
        vmask = window | sinkm;
      }
      if (vmask != 0xffffu) {
        // zero the dead V rows so that 0 * garbage cannot produce NaN (rows are 128-byte swizzle
        // units, so zeroing a whole row is layout independent)
        for (int r = 0; r < 16; ++r) {
          if (!((vmask >> r) & 1u)) {
#pragma unroll
            for (int hf = 0; hf < Cfg::kHalves; ++hf) {
              uint32_t a = vb + hf * Cfg::kHalfBytes + r * 128 + lane * 4;
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(0u) : "memory");
            }
          }
        }
        fence_proxy_async();  // generic-proxy writes before the next TMA refill of this stage
        __syncwarp();
      }

      if (ROPE) {
        // rotate the 16 K rows of this page in place at position k_rope_pos_offset[b] + kv row index
        // (kv row index = position in the visible KV, i.e. slot mapped back through sink / window)
        const int kofs = p.k_rope_pos_offset[b];
        constexpr int HC = D / 16;  // 16-byte chunk pairs per row
        for (int it = lane; it < 16 * HC; it += 32) {
          const int r = it / HC, j = it - r * HC;
          if (!((vmask >> r) & 1u)) continue;
          const int slot = slot0 + r;
          const int row_idx = slot < sink ? slot : slot - sw_off + sink;
          const float pos = static_cast<float>(kofs + row_idx) * p.rope_scale;
          const int cl = j, ch = j + HC;
          const uint32_t al = kb + (cl >> 3) * Cfg::kHalfBytes + r * 128 + (((cl & 7) ^ (r & 7)) << 4);
          const uint32_t ah = kb + (ch >> 3) * Cfg::kHalfBytes + r * 128 + (((ch & 7) ^ (r & 7)) << 4);
          uint4 lo, hi;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(al));
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "r"(ah));
          T* le = reinterpret_cast<T*>(&lo);
          T* he = reinterpret_cast<T*>(&hi);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float sn, cs;
            sincosf(pos / s_denom[j * 8 + e], &sn, &cs);
            const float xl = DT<T>::to_f(le[e]), xh = DT<T>::to_f(he[e]);
            le[e] = DT<T>::from_f(rope_mix(cs, xl, sn, DT<T>::to_f(DT<T>::neg(he[e]))));
            he[e] = DT<T>::from_f(rope_mix(cs, xh, sn, xl));
          }
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(al), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ah), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
        }
        fence_proxy_async();  // generic-proxy writes before the next TMA refill of this stage
        __syncwarp();
      }

      // ---- S^T = K . Q^T ------------------------------------------------------------------------
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int c = ks * 2 + (lane >> 4);
          const uint32_t addr = kb + (c >> 3) * Cfg::kHalfBytes + row * 128 + (((c & 7) ^ (row & 7)) << 4);
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(addr, a0, a1, a2, a3);
          mma_16816<T>(s, a0, a1, a2, a3, qf[ks][0], qf[ks][1]);
        }
      }
      // s[0],s[1]: token (lane>>2), heads qc0,qc0+1; s[2],s[3]: token (lane>>2)+8
      {
        const int t0 = lane >> 2;
        const bool v0 = (vmask >> t0) & 1u, v1 = (vmask >> (t0 + 8)) & 1u;
        s[0] = v0 ? s[0] * p.scale_log2 : -INFINITY;
        s[1] = v0 ? s[1] * p.scale_log2 : -INFINITY;
        s[2] = v1 ? s[2] * p.scale_log2 : -INFINITY;
        s[3] = v1 ? s[3] * p.scale_log2 : -INFINITY;
      }
      float mx0 = fmaxf(s[0], s[2]), mx1 = fmaxf(s[1], s[3]);
#pragma unroll
      for (int o_ = 4; o_ < 32; o_ <<= 1) {
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o_));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o_));
      }
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      if (__any_sync(0xffffffffu, (mn0 > m0) || (mn1 > m1))) {
        const float f0 = fast_exp2(m0 - mn0), f1 = fast_exp2(m1 - mn1);
        d0s *= f0;
        d1s *= f1;
#pragma unroll
        for (int i = 0; i < KS; ++i) {
          o[i][0] *= f0; o[i][1] *= f1; o[i][2] *= f0; o[i][3] *= f1;
        }
        m0 = mn0;
        m1 = mn1;
      }
      const float p0 = fast_exp2(s[0] - m0), p1 = fast_exp2(s[1] - m1);
      const float p2 = fast_exp2(s[2] - m0), p3 = fast_exp2(s[3] - m1);
      d0s += p0 + p2;
      d1s += p1 + p3;

      // ---- P^T as B operand ([k = token][n = head]) via 8x8 transposes --------------------------
      uint32_t pb0 = movmatrix_trans(DT<T>::pack(p0, p1));
      uint32_t pb1 = movmatrix_trans(DT<T>::pack(p2, p3));
      uint32_t pl0 = 0, pl1 = 0;
      constexpr bool kSplitP = std::is_same<T, __nv_bfloat16>::value;
      if (kSplitP) {
        // residuals of the bf16 rounding of P
        const float2 h01 = DT<T>::to_f2(DT<T>::pack(p0, p1));
        const float2 h23 = DT<T>::to_f2(DT<T>::pack(p2, p3));
        pl0 = movmatrix_trans(DT<T>::pack(p0 - h01.x, p1 - h01.y));
        pl1 = movmatrix_trans(DT<T>::pack(p2 - h23.x, p3 - h23.y));
      }

      // ---- O^T += V^T . P^T ---------------------------------------------------------------------
      {
        const int trow = (lane & 7) + (lane >> 4) * 8;
#pragma unroll
        for (int mt = 0; mt < KS; ++mt) {
          const int c = mt * 2 + ((lane >> 3) & 1);
          const uint32_t addr = vb + (c >> 3) * Cfg::kHalfBytes + trow * 128 + (((c & 7) ^ (trow & 7)) << 4);
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4_trans(addr, a0, a1, a2, a3);
          mma_16816<T>(o[mt], a0, a1, a2, a3, pb0, pb1);
          if (kSplitP) mma_16816<T>(o[mt], a0, a1, a2, a3, pl0, pl1);
        }
      }
      ++n_consumed;
      __syncwarp();
      if (issued_here < my_n) {
        issue(issued_here);
        ++issued_here;
      }
    }

    // ---- merge the NW warps of this item through shared memory --------------------------------
    // every warp writes into its OWN stage region (its pipeline is drained), layout:
    //   float O[8 heads][132] (pad 132 -> conflict free), float m[8], float d[8]
    constexpr int OS = D + 4;
    static_assert((8 * OS + 16) * 4 <= NSTAGE * Cfg::kStageBytes, "merge scratch must fit in the warp's stages");
    // reduce per-lane partial denominators over the 8 lanes that share a column
#pragma unroll
    for (int o_ = 4; o_ < 32; o_ <<= 1) {
      d0s += __shfl_xor_sync(0xffffffffu, d0s, o_);
      d1s += __shfl_xor_sync(0xffffffffu, d1s, o_);
    }
    float* my_scr = reinterpret_cast<float*>(smem_gen + warp * NSTAGE * Cfg::kStageBytes);
#pragma unroll
    for (int mt = 0; mt < KS; ++mt) {
      const int dd = mt * 16 + (lane >> 2);
      my_scr[qc0 * OS + dd] = o[mt][0];
      my_scr[(qc0 + 1) * OS + dd] = o[mt][1];
      my_scr[qc0 * OS + dd + 8] = o[mt][2];
      my_scr[(qc0 + 1) * OS + dd + 8] = o[mt][3];
    }
    if (lane < 4) {
      my_scr[8 * OS + qc0] = m0;
      my_scr[8 * OS + qc0 + 1] = m1;
      my_scr[8 * OS + 8 + qc0] = d0s;
      my_scr[8 * OS + 8 + qc0 + 1] = d1s;
    }
    // generic-proxy writes to stage memory must be ordered before the async-proxy (TMA) refill
    fence_proxy_async();
    __syncthreads();
    {
      // thread -> (head qh, dims): NW*32 threads cover g*D outputs
      for (int idx = threadIdx.x; idx < g * D; idx += NW * 32) {
        const int qh = idx / D, dd = idx - qh * D;
        float mm = kNegInit;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const float* scr = reinterpret_cast<const float*>(smem_gen + w * NSTAGE * Cfg::kStageBytes);
          mm = fmaxf(mm, scr[8 * OS + qh]);
        }
        float acc = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const float* scr = reinterpret_cast<const float*>(smem_gen + w * NSTAGE * Cfg::kStageBytes);
          const float f = fast_exp2(scr[8 * OS + qh] - mm);
          acc += scr[qh * OS + dd] * f;
          den += scr[8 * OS + 8 + qh] * f;
        }
        const int hq = h * g + qh;
        const bool empty = den == 0.f;
        const float outv = empty ? 0.f : acc / den;
        const float lsev = empty ? kNegInit : mm + log2f(den);
        if (whole) {
          static_cast<T*>(p.output)[(static_cast<int64_t>(b) * p.num_qo_heads + hq) * D + dd] = DT<T>::from_f(outv);
          if (dd == 0) p.lse[static_cast<int64_t>(b) * p.num_qo_heads + hq] = lsev;
        } else {
          p.part_o[(static_cast<int64_t>(slot) * g + qh) * D + dd] = outv;
          if (dd == 0) p.part_lse[static_cast<int64_t>(slot) * g + qh] = lsev;
        }
      }
    }
    __syncthreads();  // scratch (= stage memory) is reused by the next item's TMA
    // next segment of the line: the next head of this sequence, then the next sequence that has pages
    lin += pg1 - pg0;
    pg0 = 0;
    if (++h == p.num_kv_heads) {
      h = 0;
      do { ++b; } while (b < B && s_indptr[b + 1] == s_indptr[b]);
    }
  }
  // (triggering at the top of the kernel instead, and again at the top of the merge kernel, was measured: no gain)
  pdl_launch_dependents();  // the merge kernel's blocks may be scheduled; they still wait for this grid's completion
}

// Head-sharded multi-GPU decode (north_star: KV-head groups across one 8 x B200 box; reference: Disco tensor
// parallelism + ncclAllGather of the per-head outputs, src/runtime/extra/disco/nccl/nccl.cc:136-144).  Instead of a
// collective behind the kernel, the merge kernel itself stores this rank's heads into EVERY rank's gathered
// [n, total_heads, D] buffer through NVLink peer pointers, and the last block to finish raises this rank's flag in
// every peer's flag array (release at system scope); a consumer waits until all `n` flags carry the step's epoch.
struct PeerGather {
  void* out[8];        // gathered output buffer of rank i (peer-mapped device pointer)
  uint32_t* flags[8];  // flag array of rank i: flags[i][r] = last epoch rank r has completely written into rank i
  int n;               // number of ranks (0 = no gather)
  int rank;
  int head_offset;     // first gathered head of this rank
  int total_heads;
  uint32_t epoch;
  int32_t* done;       // block counter (zero between launches)
};

// The pieces of segment (b, h): CTAs k0 .. k1 of the partition; piece c lives in slot 2 (k0 + c) + (c < nc - 1).
struct SegPieces {
  int k0, nc;
  __device__ __forceinline__ int64_t slot(int c) const { return 2 * static_cast<int64_t>(k0 + c) + (c < nc - 1 ? 1 : 0); }
};
__device__ __forceinline__ SegPieces seg_pieces(const int32_t* __restrict__ page_indptr, int b, int h, int num_kv_heads,
                                                int quota) {
  const int p0 = page_indptr[b], np = page_indptr[b + 1] - p0;
  SegPieces sp;
  if (np == 0) {
    sp.k0 = 0;
    sp.nc = 0;
    return sp;
  }
  const int64_t s0 = static_cast<int64_t>(p0) * num_kv_heads + static_cast<int64_t>(h) * np;
  sp.k0 = static_cast<int>(s0 / quota);
  sp.nc = static_cast<int>((s0 + np - 1) / quota) - sp.k0 + 1;
  return sp;
}

// reduce the partial (O, LSE) of the (sequence, head) segments that were cut by a CTA boundary.  grid = (B, Hq), D threads.
template <typename T, int D>
__global__ void __launch_bounds__(D)
decode_merge_kernel(const float* __restrict__ part_o, const float* __restrict__ part_lse,
                    const int32_t* __restrict__ page_indptr, T* __restrict__ output,
                    float* __restrict__ lse, int num_qo_heads, int num_kv_heads, int quota) {
  const int b = blockIdx.x, hq = blockIdx.y, dd = threadIdx.x;
  const int g = num_qo_heads / num_kv_heads, h = hq / g, qh = hq - h * g;
  // page_indptr is an input of the step (not produced by decode_kernel): the piece list can be worked out before
  // the programmatic dependency resolves
  const SegPieces sp = seg_pieces(page_indptr, b, h, num_kv_heads, quota);
  // launched as a programmatic dependent of decode_kernel: partials are complete past this point.  EVERY block waits,
  // also the ones with nothing to merge: the completion of this grid must imply the completion of decode_kernel, or the
  // next step's kernel (a programmatic dependent of THIS grid) could overtake it.
  pdl_wait();
  if (sp.nc <= 1) return;  // empty, or written directly by the CTA that held the whole segment
  const int nc = sp.nc;
  // The kernel is pure latency (a few KiB per block): keep the loads independent -- the piece LSEs go to shared
  // memory in one parallel sweep, the partial outputs are read four at a time -- instead of two serial chains.
  __shared__ float s_lse[D];
  const bool staged = nc <= D;
  if (staged) {
    if (dd < nc) s_lse[dd] = part_lse[sp.slot(dd) * g + qh];
    __syncthreads();
  }
  auto lse_of = [&](int c) { return staged ? s_lse[c] : part_lse[sp.slot(c) * g + qh]; };
  auto po = [&](int c) { return part_o[(sp.slot(c) * g + qh) * D + dd]; };
  float mm = kNegInit;
  for (int c = 0; c < nc; ++c) mm = fmaxf(mm, lse_of(c));
  float acc = 0.f, den = 0.f;
  int c = 0;
  for (; c + 4 <= nc; c += 4) {
    const float v0 = po(c + 0), v1 = po(c + 1), v2 = po(c + 2), v3 = po(c + 3);
    const float w0 = exp2f(lse_of(c + 0) - mm), w1 = exp2f(lse_of(c + 1) - mm);
    const float w2 = exp2f(lse_of(c + 2) - mm), w3 = exp2f(lse_of(c + 3) - mm);
    acc += w0 * v0; den += w0;
    acc += w1 * v1; den += w1;
    acc += w2 * v2; den += w2;
    acc += w3 * v3; den += w3;
  }
  for (; c < nc; ++c) {
    const float w = exp2f(lse_of(c) - mm);
    acc += w * po(c);
    den += w;
  }
  const T outv = DT<T>::from_f(acc / den);
  output[(static_cast<int64_t>(b) * num_qo_heads + hq) * D + dd] = outv;
  if (dd == 0) lse[static_cast<int64_t>(b) * num_qo_heads + hq] = mm + log2f(den);
}

// Peer-gather flavour of the merge (pg.n > 0): a block owns whole sequences -- all local heads of sequence b are one
// contiguous run of Hq_local * D elements in every rank's gathered [n, total_heads, D] buffer -- and every thread slot
// produces 8 outputs, so each peer receives 16-byte vector stores (a warp writes 512 contiguous bytes per peer: full
// NVLink write packets; the first version issued 2-byte stores from a (B, Hq) grid).  Completion is one system-scope
// fence + one ticket per BLOCK (a few hundred per launch instead of B * Hq); the last block raises this rank's epoch in
// every peer's flag array and then waits, in the same kernel, until every peer's epoch has arrived here: when the
// launch completes, the gathered buffer of this step is complete on this rank and no separate wait launch is needed.
// (No cycle: every rank raises its flags before it waits, and waits only for flags.)
template <typename T, int D>
__global__ void __launch_bounds__(256)
decode_merge_gather_kernel(const float* __restrict__ part_o, const float* __restrict__ part_lse,
                           const int32_t* __restrict__ page_indptr, T* __restrict__ output, float* __restrict__ lse,
                           int num_qo_heads, int num_kv_heads, int quota, int batch, const PeerGather pg) {
  constexpr int VPH = D / 8;  // 16-byte vectors per head
  pdl_wait();  // launched as a programmatic dependent of decode_kernel: partials and direct outputs are complete
  const int nslots = num_qo_heads * VPH;
  const int g = num_qo_heads / num_kv_heads;
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    for (int slot = threadIdx.x; slot < nslots; slot += blockDim.x) {
      const int hq = slot / VPH, d0 = (slot - hq * VPH) * 8;
      const int h = hq / g, qh = hq - h * g;
      const SegPieces sp = seg_pieces(page_indptr, b, h, num_kv_heads, quota);
      const int nc = sp.nc;
      union { uint4 u; T h[8]; } pk;
      T* out_row = output + (static_cast<int64_t>(b) * num_qo_heads + hq) * D + d0;
      if (nc <= 1) {
        // final already (written by the CTA that held the whole segment, or the empty result): forward it
        pk.u = *reinterpret_cast<const uint4*>(out_row);
      } else {
        auto lse_of = [&](int c) { return part_lse[sp.slot(c) * g + qh]; };
        auto po = [&](int c) { return part_o + (sp.slot(c) * g + qh) * D + d0; };
        float mm = kNegInit;
        for (int c = 0; c < nc; ++c) mm = fmaxf(mm, lse_of(c));
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float den = 0.f;
        // (same accumulation order as decode_merge_kernel: the gathered result is bit-identical to the plain path)
        for (int c = 0; c < nc; ++c) {
          const float4 a0 = *reinterpret_cast<const float4*>(po(c)), a1 = *reinterpret_cast<const float4*>(po(c) + 4);
          const float w0 = exp2f(lse_of(c) - mm);
          acc[0] += w0 * a0.x; acc[1] += w0 * a0.y; acc[2] += w0 * a0.z; acc[3] += w0 * a0.w;
          acc[4] += w0 * a1.x; acc[5] += w0 * a1.y; acc[6] += w0 * a1.z; acc[7] += w0 * a1.w;
          den += w0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) pk.h[i] = DT<T>::from_f(acc[i] / den);
        *reinterpret_cast<uint4*>(out_row) = pk.u;
        if (d0 == 0) lse[static_cast<int64_t>(b) * num_qo_heads + hq] = mm + log2f(den);
      }
      const int64_t at = (static_cast<int64_t>(b) * pg.total_heads + pg.head_offset + hq) * D + d0;
#pragma unroll 1
      for (int i = 0; i < pg.n; ++i) *reinterpret_cast<uint4*>(static_cast<T*>(pg.out[i]) + at) = pk.u;
    }
  }
  // completion: the block's peer stores are ordered (barrier, then ONE system-scope fence by thread 0 -- fences are
  // cumulative over what the barrier made visible to it) before its ticket.  In the last block, thread i raises this
  // rank's epoch in peer i's flag array and then waits for peer i's epoch here: the n release stores (each waits for
  // the acknowledgement of everything before it) and the n polls run side by side instead of one after the other --
  // issued by one thread they cost one NVLink round trip per PEER, 30 us per step at 8 ranks.
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const int last = atomicAdd(pg.done, 1) == static_cast<int>(gridDim.x) - 1;
    if (last) *pg.done = 0;
    s_last = last;
  }
  __syncthreads();
  if (s_last && threadIdx.x < pg.n) {
    const int i = threadIdx.x;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pg.flags[i] + pg.rank), "r"(pg.epoch) : "memory");
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(pg.flags[pg.rank] + i) : "memory");
    } while (static_cast<int32_t>(v - pg.epoch) < 0);
  }
}

// wait until every rank's flag in this rank's flag array carries `epoch` (one thread per rank)
__global__ void wait_peer_flags_kernel(const uint32_t* __restrict__ flags, int n, uint32_t epoch) {
  const int i = threadIdx.x;
  if (i < n) {
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
    } while (static_cast<int32_t>(v - epoch) < 0);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map cache + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(sym);
  });
  return fn;
}

// 2-D view of a 16-bit row-major matrix [rows, cols], box = [box_rows, 64 elements], 128B swizzle
int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, uint64_t rows, uint64_t cols,
                 uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled fn = get_encode_fn();
  TVMB200_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  TVMB200_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA needs a 16-byte aligned base pointer");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype == TVMB200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TVMB200_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// 3-D view [d2][d1][d0] of a 16-bit tensor with byte strides (stride1, stride2); box = [b2][b1][64], 128B swizzle
int make_tmap_3d(CUtensorMap* out, const void* base, int dtype, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b1, uint32_t b2) {
  PFN_encodeTiled fn = get_encode_fn();
  TVMB200_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  TVMB200_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA needs a 16-byte aligned base pointer");
  TVMB200_CHECK(stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, "TMA needs 16-byte aligned strides");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {64, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dtype == TVMB200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TVMB200_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", static_cast<int>(r));
  return 0;
}

struct TmapKey {
  const void* base;
  uint64_t rows, cols;
  uint32_t box_rows;
  int dtype;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && box_rows == o.box_rows && dtype == o.dtype;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    return std::hash<const void*>()(k.base) ^ (k.rows * 1315423911u) ^ (k.cols << 7) ^ (k.box_rows << 3) ^ k.dtype;
  }
};

int get_tmap_2d_cached(CUtensorMap* out, const void* base, int dtype, uint64_t rows, uint64_t cols,
                       uint32_t box_rows) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{base, rows, cols, box_rows, dtype};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  int rc = make_tmap_2d(out, base, dtype, rows, cols, box_rows, 64);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return 0;
}

template <typename T, int D, int NW, int NSTAGE, bool ROPE, bool FUSED>
static int launch_decode_impl(const CUtensorMap& tmap, const DecodeParams& p, int grid, bool need_merge,
                         cudaStream_t st, const PeerGather& pg) {
  using Cfg = DecodeCfg<D>;
  const size_t smem = 1024 + static_cast<size_t>(NW) * NSTAGE * Cfg::kStageBytes + NW * NSTAGE * 8 +
                      (static_cast<size_t>(p.batch) + 1 + 40) * sizeof(int) + 16 + (D / 2) * sizeof(float) +
                      8 * D * sizeof(T);
  auto kern = decode_kernel<T, D, NW, NSTAGE, ROPE, FUSED>;
  TVMB200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchAttribute pdl[1];
  pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = pdl;
  cfg.numAttrs = 1;
  TVMB200_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, p));
  TVMB200_LAUNCH_OK();
  if (pg.n > 0) {
    const int nslots = p.num_qo_heads * (D / 8);
    const int max_blocks = 2 * num_sms();
    cfg.gridDim = dim3(p.batch < max_blocks ? p.batch : max_blocks);
    cfg.blockDim = dim3(nslots <= 64 ? 64 : nslots <= 128 ? 128 : 256);
    cfg.dynamicSmemBytes = 0;
    TVMB200_CUDA(cudaLaunchKernelEx(&cfg, decode_merge_gather_kernel<T, D>, static_cast<const float*>(p.part_o),
                                    static_cast<const float*>(p.part_lse), p.page_indptr, static_cast<T*>(p.output), p.lse,
                                    p.num_qo_heads, p.num_kv_heads, p.quota, p.batch, pg));
    TVMB200_LAUNCH_OK();
  } else if (need_merge) {
    cfg.gridDim = dim3(p.batch, p.num_qo_heads);
    cfg.blockDim = dim3(D);
    cfg.dynamicSmemBytes = 0;
    TVMB200_CUDA(cudaLaunchKernelEx(&cfg, decode_merge_kernel<T, D>, static_cast<const float*>(p.part_o),
                                    static_cast<const float*>(p.part_lse), p.page_indptr, static_cast<T*>(p.output), p.lse,
                                    p.num_qo_heads, p.num_kv_heads, p.quota));
    TVMB200_LAUNCH_OK();
  }
  return 0;
}

template <typename T, int D, int NW, int NSTAGE>
static int launch_decode(const CUtensorMap& tmap, const DecodeParams& p, int grid, bool need_merge,
                         cudaStream_t st, const PeerGather& pg) {
  // inline RoPE is a separate instantiation so the default (pre-rotated K) path keeps its register budget
  if (p.qkv != nullptr) {
    if constexpr (D == 128) return launch_decode_impl<T, D, NW, NSTAGE, false, true>(tmap, p, grid, need_merge, st, pg);
    return set_error("attention_decode_fused_qkv: head_dim %d unsupported (128)", D);
  }
  return p.rotary_mode == 1 ? launch_decode_impl<T, D, NW, NSTAGE, true, false>(tmap, p, grid, need_merge, st, pg)
                            : launch_decode_impl<T, D, NW, NSTAGE, false, false>(tmap, p, grid, need_merge, st, pg);
}

}  // namespace tvmb200

using namespace tvmb200;

static int decode_entry(const void* q, const void* pages, const int32_t* page_indptr,
                        const int32_t* page_values, const int32_t* length_info,
                        const int32_t* k_rope_pos_offset, const int32_t* q_rope_position, void* output, float* lse,
                        int32_t batch_size, int32_t nnz_pages, int64_t num_pages, int32_t num_qo_heads,
                        int32_t num_kv_heads, int32_t page_size, int32_t head_dim, int sliding_window,
                        int rotary_mode, float rope_scale, float rope_theta, float sm_scale, int dtype,
                        tvmb200_stream_t stream, const PeerGather& pg, const void* fused_qkv = nullptr,
                        const int32_t* append_slot = nullptr, int fused_apply_rope = 0) {
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "attention_decode: unsupported dtype %d", dtype);
  TVMB200_CHECK(page_size == 16, "attention_decode: page_size %d unsupported (the B200 path is built for 16-slot pages)", page_size);
  TVMB200_CHECK(head_dim == 128 || head_dim == 64, "attention_decode: head_dim %d unsupported (64 or 128)", head_dim);
  TVMB200_CHECK(num_kv_heads > 0 && num_qo_heads % num_kv_heads == 0, "attention_decode: num_qo_heads %d not a multiple of num_kv_heads %d", num_qo_heads, num_kv_heads);
  const int group_real = num_qo_heads / num_kv_heads;
  TVMB200_CHECK(group_real <= 8 || group_real % 8 == 0, "attention_decode: GQA group size %d unsupported (up to 8, or a multiple of 8)", group_real);
  const int vsplit = group_real > 8 ? group_real / 8 : 1;
  const int group = group_real / vsplit;
  const int32_t kv_heads_real = num_kv_heads;
  num_kv_heads *= vsplit;  // virtual heads from here on (the tensor map below keeps the real count)
  TVMB200_CHECK(batch_size <= kMaxBatchSmem, "attention_decode: batch %d exceeds %d", batch_size, kMaxBatchSmem);
  TVMB200_CHECK(rotary_mode == 0 || rotary_mode == 1, "attention_decode: rotary_mode %d (0 or 1)", rotary_mode);
  if (batch_size <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  // split-KV plan from host-known sizes only (nnz_pages, batch, heads, SM count)
  constexpr int NW = 4;
  const int sms = num_sms();
  const int ctas_per_sm = 2;
  const int grid_max = sms * ctas_per_sm;
  // balanced contiguous partition of the nnz_pages * Hkv page-heads (see DecodeParams): one quota per resident CTA,
  // never below 8 page-heads (tiny batches: fewer CTAs rather than one-page pieces), a multiple of NW so that a full
  // piece splits evenly over the warps
  const int64_t total = static_cast<int64_t>(nnz_pages) * num_kv_heads;
  int64_t quota = (total + grid_max - 1) / grid_max;
  if (quota < 8) quota = 8;
  quota = (quota + NW - 1) / NW * NW;
  TVMB200_CHECK(quota < (1ll << 30), "attention_decode: %lld page-heads are too many", static_cast<long long>(total));
  int grid = static_cast<int>((total + quota - 1) / quota);
  if (grid < 1) grid = 1;  // no pages at all: one CTA still writes the empty results
  const bool need_merge = grid > 1 || pg.n > 0;

  // workspace: part_lse [2 * grid, group] | part_o [2 * grid, group, D]  (two partial slots per CTA)
  const int64_t lse_bytes = (static_cast<int64_t>(2) * grid * group * 4 + 255) / 256 * 256;
  const int64_t o_bytes = static_cast<int64_t>(2) * grid * group * head_dim * 4;
  void* ws = nullptr;
  if (int rc = get_workspace(lse_bytes + o_bytes, st, &ws)) return rc;

  DecodeParams p;
  p.q = q;
  p.page_indptr = page_indptr;
  p.page_values = page_values;
  p.length_info = length_info;
  p.k_rope_pos_offset = k_rope_pos_offset;
  p.q_rope_position = q_rope_position;
  p.output = output;
  p.lse = lse;
  p.part_lse = static_cast<float*>(ws);
  p.part_o = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + lse_bytes);
  p.batch = batch_size;
  p.num_qo_heads = num_qo_heads;
  p.num_kv_heads = num_kv_heads;
  p.kv_heads_real = kv_heads_real;
  p.vsplit = vsplit;
  p.group = group;
  p.total = total;
  p.quota = static_cast<int>(quota);
  p.qkv = fused_qkv;
  p.append_slot = append_slot;
  p.pages = const_cast<void*>(pages);
  p.fused_apply_rope = fused_apply_rope;
  p.sliding = sliding_window ? 1 : 0;
  p.rotary_mode = rotary_mode;
  p.rope_scale = rope_scale;
  p.rope_theta = rope_theta;
  p.rs = rope_scaling();
  p.scale_log2 = sm_scale * kLog2e;
  if (rotary_mode == 1 || fused_apply_rope)
    if (int rc = check_no_rope_variant("attention_decode (in-kernel RoPE)")) return rc;

  CUtensorMap tmap;
  const uint64_t rows = static_cast<uint64_t>(num_pages) * 2 * kv_heads_real * page_size;
  if (int rc = get_tmap_2d_cached(&tmap, pages, dtype, rows, head_dim, 16)) return rc;

  if (dtype == TVMB200_F16) {
    if (head_dim == 128) return launch_decode<__half, 128, NW, 3>(tmap, p, grid, need_merge, st, pg);
    return launch_decode<__half, 64, NW, 6>(tmap, p, grid, need_merge, st, pg);
  } else {
    if (head_dim == 128) return launch_decode<__nv_bfloat16, 128, NW, 3>(tmap, p, grid, need_merge, st, pg);
    return launch_decode<__nv_bfloat16, 64, NW, 6>(tmap, p, grid, need_merge, st, pg);
  }
}

extern "C" int tvmb200_attention_decode(const void* q, const void* pages, const int32_t* page_indptr,
                                        const int32_t* page_values, const int32_t* length_info,
                                        const int32_t* k_rope_pos_offset,
                                        const int32_t* q_rope_position, void* output, float* lse,
                                        int32_t batch_size, int32_t nnz_pages, int64_t num_pages,
                                        int32_t num_qo_heads, int32_t num_kv_heads,
                                        int32_t page_size, int32_t head_dim, int sliding_window,
                                        int rotary_mode, float rope_scale, float rope_theta,
                                        float sm_scale, int dtype, tvmb200_stream_t stream) {
  PeerGather pg = {};
  return decode_entry(q, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position, output, lse,
                      batch_size, nnz_pages, num_pages, num_qo_heads, num_kv_heads, page_size, head_dim, sliding_window,
                      rotary_mode, rope_scale, rope_theta, sm_scale, dtype, stream, pg);
}

extern "C" int tvmb200_attention_decode_fused_qkv(const void* qkv, const int32_t* q_rope_position,
                                                  const int32_t* append_position_map, void* pages,
                                                  const int32_t* page_indptr, const int32_t* page_values,
                                                  const int32_t* length_info, const int32_t* k_rope_pos_offset,
                                                  void* output, float* lse, int32_t batch_size, int32_t nnz_pages,
                                                  int64_t num_pages, int32_t num_qo_heads, int32_t num_kv_heads,
                                                  int32_t page_size, int32_t head_dim, int sliding_window,
                                                  int64_t apply_rope, float rope_scale, float rope_theta, float sm_scale,
                                                  int dtype, tvmb200_stream_t stream) {
  TVMB200_CHECK(head_dim == 128, "attention_decode_fused_qkv: head_dim %d unsupported (128)", head_dim);
  TVMB200_CHECK(qkv != nullptr && append_position_map != nullptr && pages != nullptr, "attention_decode_fused_qkv: null argument");
  TVMB200_CHECK(!sliding_window, "attention_decode_fused_qkv: per-sequence sliding windows append after the attention; use the separate calls");
  PeerGather pg = {};
  // rotary_mode 0: the cached K is already rotated (RoPE mode "normal") or never rotated ("none"); q and the new k
  // are rotated here when apply_rope > 0
  return decode_entry(nullptr, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position, output, lse,
                      batch_size, nnz_pages, num_pages, num_qo_heads, num_kv_heads, page_size, head_dim, sliding_window, 0,
                      rope_scale, rope_theta, sm_scale, dtype, stream, pg, qkv, append_position_map, apply_rope > 0 ? 1 : 0);
}

static int fill_peer_gather(PeerGather* pg, void* const* peer_outputs, uint32_t* const* peer_flags, int32_t world,
                            int32_t rank, uint32_t epoch, int32_t num_qo_heads, tvmb200_stream_t stream) {
  TVMB200_CHECK(world >= 1 && world <= 8 && rank >= 0 && rank < world, "attention_decode_gather: world %d / rank %d (1..8 ranks of one box)", world, rank);
  TVMB200_CHECK(peer_outputs != nullptr && peer_flags != nullptr, "attention_decode_gather: peer pointer arrays are null");
  pg->n = world;
  pg->rank = rank;
  pg->head_offset = rank * num_qo_heads;
  pg->total_heads = world * num_qo_heads;
  pg->epoch = epoch;
  for (int i = 0; i < world; ++i) {
    TVMB200_CHECK(peer_outputs[i] != nullptr && peer_flags[i] != nullptr, "attention_decode_gather: peer %d pointer is null", i);
    pg->out[i] = peer_outputs[i];
    pg->flags[i] = peer_flags[i];
  }
  // block ticket of this (context, device, stream): zero between launches
  int32_t* counters = nullptr;
  if (int rc = get_counters(static_cast<cudaStream_t>(stream), &counters)) return rc;
  pg->done = counters + 64;
  return 0;
}

extern "C" int tvmb200_attention_decode_gather(const void* q, const void* pages, const int32_t* page_indptr,
                                               const int32_t* page_values, const int32_t* length_info,
                                               const int32_t* k_rope_pos_offset, const int32_t* q_rope_position,
                                               void* output, float* lse, int32_t batch_size, int32_t nnz_pages,
                                               int64_t num_pages, int32_t num_qo_heads, int32_t num_kv_heads,
                                               int32_t page_size, int32_t head_dim, int sliding_window, int rotary_mode,
                                               float rope_scale, float rope_theta, float sm_scale, int dtype,
                                               void* const* peer_outputs, uint32_t* const* peer_flags, int32_t world,
                                               int32_t rank, uint32_t epoch, tvmb200_stream_t stream) {
  PeerGather pg = {};
  if (int rc = fill_peer_gather(&pg, peer_outputs, peer_flags, world, rank, epoch, num_qo_heads, stream)) return rc;
  if (batch_size <= 0) return 0;
  return decode_entry(q, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position, output, lse,
                      batch_size, nnz_pages, num_pages, num_qo_heads, num_kv_heads, page_size, head_dim, sliding_window,
                      rotary_mode, rope_scale, rope_theta, sm_scale, dtype, stream, pg);
}

extern "C" int tvmb200_attention_decode_fused_qkv_gather(
    const void* qkv, const int32_t* q_rope_position, const int32_t* append_position_map, void* pages,
    const int32_t* page_indptr, const int32_t* page_values, const int32_t* length_info, const int32_t* k_rope_pos_offset,
    void* output, float* lse, int32_t batch_size, int32_t nnz_pages, int64_t num_pages, int32_t num_qo_heads,
    int32_t num_kv_heads, int32_t page_size, int32_t head_dim, int sliding_window, int64_t apply_rope, float rope_scale,
    float rope_theta, float sm_scale, int dtype, void* const* peer_outputs, uint32_t* const* peer_flags, int32_t world,
    int32_t rank, uint32_t epoch, tvmb200_stream_t stream) {
  TVMB200_CHECK(head_dim == 128, "attention_decode_fused_qkv_gather: head_dim %d unsupported (128)", head_dim);
  TVMB200_CHECK(qkv != nullptr && append_position_map != nullptr && pages != nullptr, "attention_decode_fused_qkv_gather: null argument");
  TVMB200_CHECK(!sliding_window, "attention_decode_fused_qkv_gather: per-sequence sliding windows append after the attention; use the separate calls");
  PeerGather pg = {};
  if (int rc = fill_peer_gather(&pg, peer_outputs, peer_flags, world, rank, epoch, num_qo_heads, stream)) return rc;
  if (batch_size <= 0) return 0;
  return decode_entry(nullptr, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position, output, lse,
                      batch_size, nnz_pages, num_pages, num_qo_heads, num_kv_heads, page_size, head_dim, sliding_window, 0,
                      rope_scale, rope_theta, sm_scale, dtype, stream, pg, qkv, append_position_map, apply_rope > 0 ? 1 : 0);
}

extern "C" int tvmb200_wait_peer_flags(const uint32_t* flags, int32_t world, uint32_t epoch, tvmb200_stream_t stream) {
  TVMB200_CHECK(flags != nullptr && world >= 1 && world <= 8, "wait_peer_flags: bad arguments");
  wait_peer_flags_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags, world, epoch);
  TVMB200_LAUNCH_OK();
  return 0;
}
