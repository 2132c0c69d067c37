// Batch prefill attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Serves f_attention_prefill_ragged (K4) and f_attention_prefill (K3) for the hot configurations:
// head_dim 128, rotary_mode 0 (K pre-rotated: RoPEMode::kNormal / kNone), mask none or causal, no
// per-sequence sliding window, GQA group in {1,2,4,8,16}.  Everything else runs on prefill_generic.cu.
// Reference semantics: python/tvm/relax/frontend/nn/llm/_prefill_kernels.py:217-391 (paged), :795-923
// (ragged); numerics per _kernel_common.py:222-334 (base-2 online softmax, fp32 accumulate, LSE = m + log2 d).
//
// Design (persistent: one CTA per SM pulls work items = 256 GQA-folded query rows of one sequence x one KV head = two
// 128-row UMMA tiles from a device-wide queue, long items first):
//   warp 0      work fetch (atomic counter -> 2-entry item ring in shared memory) + TMA producer: Q tiles (3-D box:
//               64 cols x g heads x 128/g tokens = the reference's row fold row = token*g + head,
//               _prefill_kernels.py:318-324) once per item, K_j / V_j tiles of 128 KV rows through a 4-slot ring
//               (ragged: one 3-D box per 64-col half; paged: eight 16-row page boxes per half, page ids looked up by
//               the producer lanes).  The next item's Q / K / V loads start while the current item finishes.
//   warp 1      MMA issuer (one elected thread).  Per Q tile t and KV tile j: S_t = Q_t K_j^T as eight N = 128 SS
//               MMAs (K-major operands) into the tile's 128-column S region, and O_t += P_t V_j as two groups of
//               four TS MMAs (P read from TMEM, V an MN-major shared-memory operand), one group per 64-column half
//               of P as soon as that half is ready.  N = 128 is deliberate: an SS-mode MMA never takes fewer than
//               ~64 clk (scripts/umma_bench.cu: 63 clk at N = 64, 67 at N = 128, 128 at N = 256), so 64-column QK
//               steps run the tensor pipe at half rate.  P arrives in a fixed order -- (t0, lower half), (t0, upper),
//               (t1, lower), (t1, upper) -- and QK_t(j+1) is issued right behind the upper-half PV of tile t, so one
//               Q tile's PV + QK occupy the tensor pipe while the other tile's warpgroup runs its softmax.
//   warps 2, 3  TMEM allocation (512 columns: S_0, S_1 of 128 columns each, O_0, O_1 of 128; P aliases the first 32
//               columns of the S half it was computed from) and, for bf16 inputs, conversion of every V tile to fp16 in
//               shared memory, so that P can be fp16 (kind::f16 wants one format for A and B).
//   warps 4-7   softmax warpgroup of tile 0, warps 8-11 of tile 1: one thread per row; per KV tile: tcgen05.ld the 128
//               S values, mask (diagonal / tail tiles only), ONE row maximum and LAZY rescale decision (O in TMEM is
//               only rescaled when the max grows by more than 2^8), then per 64-column half exp2 (4 of 16 pairs on
//               the FMA pipe), pack to fp16, tcgen05.st P, arrive.  Each tile stops at its own last visible step
//               under a causal mask.  The same threads normalise and store O / LSE at the end of an item.
// Measured (scripts/softmax_bench.cu, scripts/prefill_trace.py): the kernel is bound by the softmax instruction
// stream, not by the tensor pipe (~55 % busy) nor by the MUFU.
// All hand-offs are mbarriers (TMA complete_tx, tcgen05.commit, thread arrives), reused across items with running
// use counts for the wait parity; no __syncthreads in the loop.
#include "prefill.cuh"
#include "tc05.cuh"

#include <cstdlib>
#include <mutex>
#include <type_traits>

namespace tvmb200 {

namespace {

constexpr int kRows = 128;                 // rows of a UMMA tile (M)
constexpr int kKV = 128;                   // KV rows per shared-memory tile
constexpr int kStep = 64;                  // KV columns per softmax step (N of QK^T, K of PV)
constexpr int kD = 128;                    // head dim
constexpr int kHalfBytes = kRows * 128;    // one 64-column half of a 128-row tile: 16 KiB
constexpr int kTileBytes = 2 * kHalfBytes; // 32 KiB
constexpr int kSlots = 4;                  // K/V ring
constexpr int kThreads = 384;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: P <= 2^8
// P is always fp16 (11 bits; the reference keeps P in fp32): kind::f16 UMMA needs A and B in one format, so for bf16
// inputs the V tile is converted to fp16 in shared memory by two otherwise idle warps (exact for |v| in [2^-14, 65504],
// saturating above).  The first tcgen05 versions kept P in bf16 and added a second PV pass with the rounding residual
// P_lo = P - bf16(P) for steps with dominant weights: slower (956 vs 976 TFLOP/s) and much more code.
// of every 16 column pairs, how many take 2^x on the FMA pipe instead of the MUFU.  Swept on C3 in round 2 (fp16 P, burst /
// sustained TFLOP/s): 0: 981 / 875, 1: 981 / 875, 2: 981-1001 / 874-876, 3: 977 / 873, 4: 974-995 / 865, 6: 982 / 855,
// 8: 965 / 835 -- flat within 1 % up to 4, then worse: the MUFU is not what the softmax warps wait for
#ifdef TVMB200_POLY_PAIRS
template <typename PT>
constexpr int kPolyPairsOf = TVMB200_POLY_PAIRS;
#else
template <typename PT>
constexpr int kPolyPairsOf = 2;
#endif

// NS = K / V tile slots: 4 = one ring K_j, V_j, K_j+1, V_j+1; 5 (VR3 instantiations) = a K ring of 2 + a V ring of 3
template <int NS>
struct SmemLayoutT {
  static constexpr int q = 0;
  static constexpr int kv = q + 2 * kTileBytes;
  static constexpr int bars = kv + NS * kTileBytes;   // 33 mbarriers
  static constexpr int tmem_ptr = bars + 288;
  static constexpr int item = bars + 304;     // int[2]: work-item ring filled by the producer warp
  static constexpr int scan = bars + 320;
};
// S_FULL / P_READY / PV_DONE: one barrier per (tile, S half) = index 2 t + h.  The kernel is persistent, so every
// barrier is used across work items: each role keeps running use counts and waits for parity (count & 1).
enum Bar {
  Q_FULL = 0, KV_FULL = 1, KV_EMPTY = 6, S_FULL = 11, P_READY = 15, PV_DONE = 19, Q_EMPTY = 23, ITEM_FULL = 24,
  ITEM_EMPTY = 26, V_CONV = 28, NUM_BARS = 33   // KV_FULL / KV_EMPTY / V_CONV: one per tile slot (up to 5)
};
// The f-th K / V tile load of a CTA (K tiles even f, V tiles odd f) -> its slot and the parity of its use of that slot.
// VR3: V tiles go round three slots of their own, so a V tile is requested two KV tiles ahead of its PV instead of one:
// the paged bf16 kernel's tile period was set by the loop slot free -> TMA from HBM -> convert -> PV -> slot free.
template <bool VR3>
__device__ __forceinline__ int slot_of(uint32_t f) {
  return VR3 ? ((f & 1) ? 2 + static_cast<int>((f >> 1) % 3u) : static_cast<int>((f >> 1) & 1u)) : static_cast<int>(f & 3u);
}
template <bool VR3>
__device__ __forceinline__ uint32_t phase_of(uint32_t f) {
  return VR3 ? ((f & 1) ? ((f >> 1) / 3u) & 1u : (f >> 2) & 1u) : (f >> 2) & 1u;
}

#ifdef TVMB200_TRACE
// tuning aid: clock64 stamps of one CTA (role 0 = MMA warp, 1 / 2 = softmax warpgroup 0 / 1), [role][step][slot];
// the stamps are those of work item number TVMB200_TRACE (whichever CTA fetched it)
__device__ long long g_trace[3][64][8];
#define TRACE(role, step, slot)                                                             \
  do {                                                                                      \
    if (trace_item == TVMB200_TRACE && (threadIdx.x & 31) == 0 && (step) < 64)              \
      g_trace[role][step][slot] = clock64();                                                \
  } while (0)
#else
#define TRACE(role, step, slot)
#endif

template <typename PT>
__device__ __forceinline__ uint32_t pack_p(float lo, float hi) {
  return DT<PT>::pack(lo, hi);
}

// one work item = 256 GQA-folded query rows of one sequence x one KV head
struct Item {
  int h, q_beg, qo_len, row0, tok0, nqt, kv_len, kv_beg, pg_beg, n_pages, ns0, ns1, n_kv;
  int tree_beg, tree_len;  // kMaskTree: the sequence's rows of tree_order; the mask covers the trailing tree_len columns
  int j0;                  // first 128-row KV tile of this item (kv_splits > 1: the item covers tiles [j0, j0 + n_kv))
  int split;               // which part (kv_splits > 1)
};

template <bool PAGED, bool SPLIT>
__device__ __forceinline__ Item decode_item(const PrefillParams& p, const int* s_tiles, int id, int n_items) {
  Item it;
  const int B = p.batch, g = p.group;
  const int ritem0 = n_items - 1 - id;  // late (= long-KV under a causal mask) tiles first
  const int S = SPLIT ? p.kv_splits : 1;
  const int ritem = SPLIT ? ritem0 / S : ritem0;
  it.split = SPLIT ? ritem0 - ritem * S : 0;
  const int tg = ritem / p.num_kv_heads;
  it.h = ritem - tg * p.num_kv_heads;
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_tiles[mid] <= tg) lo = mid; else hi = mid;
  }
  const int b = lo;
  const int pair = tg - s_tiles[b];
  it.q_beg = p.q_indptr[b];
  it.qo_len = p.q_indptr[b + 1] - it.q_beg;
  it.row0 = pair * 2 * kRows;            // first folded row of the item
  it.tok0 = it.row0 / g;                 // first query token (relative to the sequence)
  const int tok_per_tile = kRows / g;
  it.nqt = (it.qo_len * g - it.row0 > kRows) ? 2 : 1;  // second tile entirely outside the sequence?
  it.kv_beg = 0;
  it.pg_beg = 0;
  it.n_pages = 0;
  if (PAGED) {
    it.pg_beg = p.page_indptr[b];
    it.n_pages = p.page_indptr[b + 1] - it.pg_beg;
    it.kv_len = it.n_pages > 0 ? (it.n_pages - 1) * 16 + p.length_info[b] : 0;
  } else {
    it.kv_beg = p.kv_indptr[b];
    it.kv_len = p.kv_indptr[b + 1] - it.kv_beg;
  }
  it.tree_beg = 0;
  it.tree_len = 0;
  if (p.mask_mode == kMaskTree) {
    it.tree_beg = p.tree_indptr[b];
    it.tree_len = p.tree_indptr[b + 1] - it.tree_beg;
  }
  const bool causal = p.mask_mode == kMaskCausal;
  // visible KV extent of each tile's last valid token bounds that tile's step count
  auto steps_of = [&](int t) {
    const int tok_last = min(it.qo_len, it.tok0 + (t + 1) * tok_per_tile) - 1;
    const int kv_end = causal ? max(0, min(it.kv_len, it.kv_len - it.qo_len + tok_last + 1)) : it.kv_len;
    return (t < it.nqt) ? (kv_end + kStep - 1) / kStep : 0;
  };
  it.ns0 = steps_of(0);
  it.ns1 = steps_of(1);
  it.n_kv = (max(it.ns0, it.ns1) + 1) >> 1;  // 128-row K / V tiles to load
  it.j0 = 0;
  if (SPLIT) {
    const int per = (it.n_kv + S - 1) / S;
    it.j0 = min(it.split * per, it.n_kv);
    const int j1 = min(it.n_kv, it.j0 + per);
    it.ns0 = max(0, min(it.ns0 - 2 * it.j0, 2 * (j1 - it.j0)));
    it.ns1 = max(0, min(it.ns1 - 2 * it.j0, 2 * (j1 - it.j0)));
    it.n_kv = (max(it.ns0, it.ns1) + 1) >> 1;
  }
  return it;
}

}  // namespace

// Persistent: one CTA per SM walks a device-wide work queue (an atomic counter; long items first).  The producer warp
// fetches the next item id and publishes it through a two-entry shared-memory ring, so the next item's Q and first
// K / V tiles are in flight -- and its first QK^T issued -- while the softmax warpgroups still normalise and store the
// previous item's O: TMEM allocation, barrier setup, the Q load from HBM and the O store no longer sit between two
// tiles' tensor work (they cost ~12 % with one CTA per tile).
// XMASK = true adds the two masks that are not a plain "columns [0, limit)": the per-layer sliding window (a LOWER bound
// per row, _kernel_common.py:130-144) and the token-tree mask on the trailing tree_len columns (an ancestor test per
// (row, column), tree_attn.py:48-65).  SPLIT = true cuts every item's KV range into p.kv_splits parts that write fp32
// partials.  Separate instantiations: the causal / mask-free kernel of the hot configuration keeps its registers.
template <typename T, typename PT, bool PAGED, bool XMASK, bool SPLIT, bool VR3 = false>
__global__ void __launch_bounds__(kThreads, 1)
prefill_tc05_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const PrefillParams p, const uint32_t idesc_qk,
                    const uint32_t idesc_pv, int* __restrict__ work_counter) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using SmemLayout = SmemLayoutT<VR3 ? 5 : 4>;
  constexpr int kNumSlots = VR3 ? 5 : 4;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int B = p.batch, g = p.group;
  static_assert(std::is_same<PT, __half>::value, "P is fp16: bf16 inputs convert V in shared memory");
  // bf16 inputs with fp16 P: warps 2 and 3 convert every V tile to fp16 in shared memory (exact for |v| in
  // [2^-14, 65504], saturating above), so P keeps 11 bits without a residual pass and PV is an f16 x f16 MMA
  constexpr bool kConvertV = !std::is_same<T, PT>::value;
  int* s_tiles = reinterpret_cast<int*>(sgen + SmemLayout::scan);
  int* s_tmp = s_tiles + B + 1;
  volatile int* s_item = reinterpret_cast<volatile int*>(sgen + SmemLayout::item);
  auto bar = [&](int i) -> uint32_t { return sbase + SmemLayout::bars + i * 8; };

  // ---- work enumeration: 256-row tile pairs per sequence ---------------------------------------------------------
  for (int b = tid; b < B; b += kThreads) {
    const int rows = (p.q_indptr[b + 1] - p.q_indptr[b]) * g;
    s_tiles[b] = (rows + 2 * kRows - 1) / (2 * kRows);
  }
  __syncthreads();
  block_exclusive_scan(s_tiles, B, s_tmp);
  const int n_items = s_tiles[B] * p.num_kv_heads * (SPLIT ? p.kv_splits : 1);
  const bool causal = p.mask_mode == kMaskCausal;
  const int tok_per_tile = kRows / g;

  // ---- one-time setup ------------------------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    if (!PAGED) tma_prefetch_desc(&tm_v);
    mbar_init(bar(Q_FULL), 1);
    mbar_init(bar(Q_EMPTY), 1);
    for (int i = 0; i < kNumSlots; ++i) {
      mbar_init(bar(KV_FULL + i), 1);
      mbar_init(bar(KV_EMPTY + i), 1);
      mbar_init(bar(V_CONV + i), 2);
    }
    for (int t = 0; t < 2; ++t) {
      for (int h = 0; h < 2; ++h) {
        mbar_init(bar(S_FULL + 2 * t + h), 1);
        mbar_init(bar(P_READY + 2 * t + h), kRows);
        mbar_init(bar(PV_DONE + 2 * t + h), 1);
      }
      mbar_init(bar(ITEM_FULL + t), 1);
      mbar_init(bar(ITEM_EMPTY + t), kConvertV ? 11 : 9);  // the MMA warp + the eight softmax warps (+ two V converters)
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tc05::tmem_alloc(sbase + SmemLayout::tmem_ptr, 512);
    tc05::tmem_relinquish();
  }
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + SmemLayout::tmem_ptr);
  const uint32_t sq = sbase + SmemLayout::q, skv = sbase + SmemLayout::kv;

  if (warp < 4) {
    tc05::setmaxnreg_dec<72>();
    if (warp == 0) {
      // =========================== TMA producer + work fetch ===========================
      uint32_t fill = 0;   // K / V tiles loaded so far (all items): ring slot = fill & 3
      uint32_t n_q = 0;    // Q loads so far
      for (int k = 0;; ++k) {
        const int islot = k & 1;
        int id = 0;
        if (lane == 0) {
          mbar_wait(bar(ITEM_EMPTY + islot), ((k >> 1) & 1) ^ 1);
          id = atomicAdd(work_counter, 1);
          if (id == n_items + static_cast<int>(gridDim.x) - 1) *work_counter = 0;  // the very last fetch of the launch
          s_item[islot] = id;
          mbar_arrive(bar(ITEM_FULL + islot));
        }
        id = __shfl_sync(0xffffffffu, id, 0);
        if (id >= n_items) break;
        const Item it = decode_item<PAGED, SPLIT>(p, s_tiles, id, n_items);
        if (it.n_kv == 0) continue;
        if (lane == 0) {
          mbar_wait(bar(Q_EMPTY), (n_q & 1) ^ 1);  // every QK^T of the previous item has read Q
          mbar_expect_tx(bar(Q_FULL), it.nqt * kTileBytes);
          for (int t = 0; t < it.nqt; ++t)
            for (int hf = 0; hf < 2; ++hf)
              tc05::tma_load_3d(sq + t * kTileBytes + hf * kHalfBytes, &tm_q, hf * 64, it.h * g,
                                it.q_beg + it.tok0 + t * tok_per_tile, bar(Q_FULL), kEvictFirst);
        }
        ++n_q;
        for (int fl = 0; fl < 2 * it.n_kv; ++fl, ++fill) {
          const int j = (SPLIT ? it.j0 : 0) + (fl >> 1), is_v = fl & 1, slot = slot_of<VR3>(fill);  // j: absolute KV tile
          int pid = 0;
          if (PAGED && lane < 16) {
            // the page id does not depend on the ring slot: its load runs under the wait for the slot (the paged kernel's
            // tile period is set by the loop slot free -> page id -> TMA from HBM -> [bf16: convert] -> PV -> slot free)
            const int pi = min(j * 8 + (lane >> 1), it.n_pages - 1);  // tail boxes re-read the last page (rows masked / zeroed)
            pid = __ldg(p.page_values + it.pg_beg + pi);
          }
          mbar_wait(bar(KV_EMPTY + slot), phase_of<VR3>(fill) ^ 1);
          const uint32_t dst = skv + slot * kTileBytes;
          if (!PAGED) {
            if (lane == 0) {
              mbar_expect_tx(bar(KV_FULL + slot), kTileBytes);
              const CUtensorMap* tm = is_v ? &tm_v : &tm_k;
              for (int hf = 0; hf < 2; ++hf)
                tc05::tma_load_3d(dst + hf * kHalfBytes, tm, hf * 64, it.h, it.kv_beg + j * kKV, bar(KV_FULL + slot),
                                  kEvictLast);
            }
          } else {
            if (lane == 0) mbar_expect_tx(bar(KV_FULL + slot), kTileBytes);
            __syncwarp();
            if (lane < 16) {
              const int i = lane >> 1, hf = lane & 1;
              const int row = ((pid * 2 + is_v) * p.num_kv_heads + it.h) * 16;
              tma_load_2d(dst + hf * kHalfBytes + i * 16 * 128, &tm_k, hf * 64, row, bar(KV_FULL + slot), kEvictLast);
            }
          }
        }
      }
    } else if (warp == 1) {
      // =========================== MMA issuer ===========================
      // descriptors: high words are constant per operand kind, low words = (address >> 4) | LBO field
      const uint64_t dkm = tc05::make_smem_desc(0, 16, 1024);          // K-major operands (Q, K)
      const uint64_t dmn = tc05::make_smem_desc(0, kHalfBytes, 1024);  // MN-major operand (V)
      const uint32_t kmaj_hi = static_cast<uint32_t>(dkm >> 32), kmaj_lo = static_cast<uint32_t>(dkm);
      const uint32_t mnmaj_hi = static_cast<uint32_t>(dmn >> 32), mnmaj_lo = static_cast<uint32_t>(dmn);
      const uint32_t q_lo = kmaj_lo + (sq >> 4), k_lo = kmaj_lo + (skv >> 4), v_lo = mnmaj_lo + (skv >> 4);
      // S_t[0:128) = Q_t x (K tile in `kslot`)^T, one N = 128 MMA per 16-wide slice of the head dim
      auto issue_qk = [&](int t, uint32_t kslot) {
        const uint32_t a0 = q_lo + ((t * kTileBytes) >> 4);
        const uint32_t b0 = k_lo + ((kslot * kTileBytes) >> 4);
        const uint32_t d = tmem + t * 128;
#pragma unroll
        for (int s = 0; s < kD / 16; ++s) {
          const uint32_t off = ((s >> 2) * kHalfBytes + (s & 3) * 32) >> 4;
          tc05::mma_ss_w(d, a0 + off, kmaj_hi, b0 + off, kmaj_hi, idesc_qk, s > 0);
        }
      };
      auto issue_pv = [&](int t, uint32_t vslot, int half, bool acc) {
        const uint32_t b0 = v_lo + ((vslot * kTileBytes + half * (kStep * 128)) >> 4);
        const uint32_t p0 = tmem + t * 128 + half * kStep;
        const uint32_t d = tmem + 256 + t * 128;
#pragma unroll
        for (int s = 0; s < kStep / 16; ++s) {
          // A = P_t[:, 16s .. 16s+15] = 8 packed columns; B = V rows 16s.. (MN-major: LBO = next 64-col half)
          tc05::mma_ts_w(d, p0 + s * 8, b0 + ((s * 16 * 128) >> 4), mnmaj_hi, idesc_pv, (acc || s > 0) ? 1u : 0u);
        }
      };
      uint32_t fill = 0, n_q = 0;
      uint32_t n_p[2][2] = {{0, 0}, {0, 0}};  // real (tile, half) steps issued so far = completions of P_READY waited
      for (int k = 0;; ++k) {
        const int islot = k & 1;
        mbar_wait(bar(ITEM_FULL + islot), (k >> 1) & 1);
        const int id = s_item[islot];
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(ITEM_EMPTY + islot));
        if (id >= n_items) break;
        const int trace_item = id;
        (void)trace_item;
        const Item it = decode_item<PAGED, SPLIT>(p, s_tiles, id, n_items);
        const int n_kv = it.n_kv;
        if (n_kv == 0) continue;
        auto ns = [&](int t) { return t ? it.ns1 : it.ns0; };
        mbar_wait(bar(Q_FULL), n_q & 1);
        ++n_q;
        mbar_wait(bar(KV_FULL + slot_of<VR3>(fill)), phase_of<VR3>(fill));  // K_0
        tc05::fence_after_sync();
        if (tc05::elect_one()) {
          for (int t = 0; t < 2; ++t)
            if (ns(t) > 0) {
              issue_qk(t, slot_of<VR3>(fill));
              tc05::commit(bar(S_FULL + 2 * t));
              tc05::commit(bar(S_FULL + 2 * t + 1));
            }
          tc05::commit(bar(KV_EMPTY + slot_of<VR3>(fill)));
          if (n_kv == 1) tc05::commit(bar(Q_EMPTY));  // no further QK^T in this item
        }
        __syncwarp();
        // The softmax warpgroups hand P over in a fixed order: (t0, lower half), (t0, upper half), (t1, lower),
        // (t1, upper), next KV tile.  Behind the upper half of a tile goes its QK for KV tile j+1, i.e. one Q tile's
        // PV + QK occupy the tensor pipe while the other tile's warpgroup runs its softmax.
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t fv = fill + 2 * j + 1, fk1 = fill + 2 * j + 2;
          const int vslot = slot_of<VR3>(fv), k1slot = slot_of<VR3>(fk1);
          const bool more_k = j + 1 < n_kv;
          mbar_wait(bar((kConvertV ? V_CONV : KV_FULL) + vslot), phase_of<VR3>(fv));
          if (kConvertV) tc05::fence_after_sync();
          if (PAGED && !kConvertV && j == n_kv - 1) {
            // last tile: rows past kv_len of the V tile hold whatever is in the page (maybe NaN bit patterns);
            // P is exactly 0 there but 0 * NaN = NaN, so zero them before the tensor core reads them
            const int valid = it.kv_len - ((SPLIT ? it.j0 : 0) + j) * kKV;
            if (valid < kKV) {
              for (int e = lane; e < (kKV - valid) * 16; e += 32) {
                const int r = valid + (e >> 4), c = e & 15;
                const uint32_t a = skv + vslot * kTileBytes + (c >> 3) * kHalfBytes + r * 128 + ((c & 7) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u) : "memory");
              }
              fence_proxy_async();
              __syncwarp();
            }
          }
          if (more_k) mbar_wait(bar(KV_FULL + k1slot), phase_of<VR3>(fk1));
          for (int t = 0; t < 2; ++t) {
            const int nst = ns(t);
            if (2 * j >= nst) continue;
            for (int h = 0; h < 2; ++h) {
              const int s = 2 * j + h;
              if (s >= nst) break;
              mbar_wait(bar(P_READY + 2 * t + h), n_p[t][h] & 1);
              ++n_p[t][h];
              TRACE(0, s, 3 * t + 1);
              tc05::fence_after_sync();
              const bool last_of_tile = h == 1 || s == nst - 1;
              if (tc05::elect_one()) {
                issue_pv(t, vslot, h, s > 0);
                tc05::commit(bar(PV_DONE + 2 * t + h));
                if (last_of_tile && 2 * (j + 1) < nst) {
                  issue_qk(t, k1slot);
                  tc05::commit(bar(S_FULL + 2 * t));
                  tc05::commit(bar(S_FULL + 2 * t + 1));
                }
              }
              __syncwarp();
              TRACE(0, s, 3 * t + 2);
            }
          }
          if (tc05::elect_one()) {
            tc05::commit(bar(KV_EMPTY + vslot));
            if (more_k) tc05::commit(bar(KV_EMPTY + k1slot));
            if (j + 2 == n_kv) tc05::commit(bar(Q_EMPTY));  // the item's last QK^T (of KV tile n_kv-1) is issued
          }
          __syncwarp();
        }
        fill += 2 * n_kv;
      }
    } else if (kConvertV) {
      // =========================== V converters (warps 2 and 3): bf16 -> fp16 in place ===========================
      uint32_t fill = 0;
      for (int k = 0;; ++k) {
        const int islot = k & 1;
        mbar_wait(bar(ITEM_FULL + islot), (k >> 1) & 1);
        const int id = s_item[islot];
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(ITEM_EMPTY + islot));
        if (id >= n_items) break;
        const Item it = decode_item<PAGED, SPLIT>(p, s_tiles, id, n_items);
        for (int j = 0; j < it.n_kv; ++j) {
          const uint32_t fv = fill + 2 * j + 1;
          const int vslot = slot_of<VR3>(fv);
          mbar_wait(bar(KV_FULL + vslot), phase_of<VR3>(fv));
          // warp 2 owns the first 64-column half of the tile (16 KiB), warp 3 the second; the element-wise
          // conversion does not care about the 128B swizzle.  Rows past kv_len of a paged tile (whatever the page
          // holds, maybe NaN bit patterns) become zeros: P is exactly 0 there but 0 * NaN = NaN.
          uint4* half = reinterpret_cast<uint4*>(sgen + SmemLayout::kv + vslot * kTileBytes + (warp - 2) * kHalfBytes);
          const int valid = (PAGED && j == it.n_kv - 1) ? it.kv_len - ((SPLIT ? it.j0 : 0) + j) * kKV : kKV;
          // eight 16-byte vectors per lane in flight: all loads, then the conversions, then all stores (written as a plain
          // load / convert / store loop the in-place update serialises on shared-memory latency -- ncu showed the two
          // converter warps busy for the whole KV-tile period of the paged kernel, the MMA warp waiting on V_CONV)
          constexpr int kBatch = 8;
          for (int e0 = lane; e0 < kHalfBytes / 16; e0 += 32 * kBatch) {
            uint4 v[kBatch];
#pragma unroll
            for (int i = 0; i < kBatch; ++i) v[i] = half[e0 + 32 * i];
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
              if (((e0 + 32 * i) >> 3) < valid) {  // 8 x 16 B per 128-byte row of the half
                v[i].x = tc05::bf16x2_to_f16x2_sat(v[i].x);
                v[i].y = tc05::bf16x2_to_f16x2_sat(v[i].y);
                v[i].z = tc05::bf16x2_to_f16x2_sat(v[i].z);
                v[i].w = tc05::bf16x2_to_f16x2_sat(v[i].w);
              } else {
                v[i] = make_uint4(0u, 0u, 0u, 0u);
              }
            }
#pragma unroll
            for (int i = 0; i < kBatch; ++i) half[e0 + 32 * i] = v[i];
          }
          fence_proxy_async();  // generic-proxy writes before the tensor core (async proxy) reads the tile
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(V_CONV + vslot));
        }
        fill += 2 * it.n_kv;
      }
    }
  } else {
    // =========================== softmax / correction / epilogue ===========================
    tc05::setmaxnreg_inc<216>();
    const int t = (warp - 4) >> 2;          // which Q tile
    const int wq = warp & 3;                // TMEM lane quarter of this warp
    const int r = wq * 32 + lane;           // row within the tile
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t t_s = tmem + lane_addr + t * 128;
    const uint32_t t_o = tmem + lane_addr + 256 + t * 128;
    const float sc = p.scale_log2;
    uint32_t n_s = 0;             // KV tiles (= QK^T results) of my Q tile consumed so far, all items
    uint32_t n_p[2] = {0, 0};     // my steps so far per S half = my arrivals on P_READY(t, half)
    for (int k = 0;; ++k) {
      const int islot = k & 1;
      mbar_wait(bar(ITEM_FULL + islot), (k >> 1) & 1);
      const int id = s_item[islot];
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(ITEM_EMPTY + islot));
      if (id >= n_items) break;
      const int trace_item = id;
      (void)trace_item;
      const Item it = decode_item<PAGED, SPLIT>(p, s_tiles, id, n_items);
      const int h = it.h, q_beg = it.q_beg, qo_len = it.qo_len, kv_len = it.kv_len, nqt = it.nqt;
      const int R = it.row0 + t * kRows + r;     // folded row within the sequence
      const int tok = R / g;
      const bool valid = t < nqt && tok < qo_len;
      const int limit = causal ? max(0, min(kv_len, kv_len - qo_len + tok + 1)) : kv_len;  // visible columns [0, limit)
      // XMASK: columns below `lower` are hidden (layer sliding window: the last max(sws - tok - 1, 0) columns are
      // visible); tree: row = node `tok + tree_len - qo_len` of the tree, whose dfs order must fall into the
      // [order, subtree end) interval of the column's node
      int lower = 0, my_order = 0;
      const int tree_start = kv_len - it.tree_len;
      if (XMASK) {
        if (p.mask_mode == kMaskLayerSliding) lower = max(kv_len - max(p.layer_sws - tok - 1, 0), 0);
        if (p.mask_mode == kMaskTree) {
          const int child = tok + it.tree_len - qo_len;
          if (tok < qo_len && child >= 0) my_order = p.tree_order[2 * (it.tree_beg + child)];
        }
      }
      float m_used = kNegInit, l = 0.f;
      const int my_ns = t ? it.ns1 : it.ns0;
      if (t >= nqt) continue;
      const uint32_t ninf = __float_as_uint(-INFINITY);
      // columns [0, rem) of a 64-column half are visible
      auto mask_half = [&](uint32_t (&x0)[32], uint32_t (&x1)[32], int rem) {
        if (__any_sync(0xffffffffu, rem < kStep)) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (c >= rem) x0[c] = ninf;
            if (c + 32 >= rem) x1[c] = ninf;
          }
        }
      };
      // XMASK: hide columns below `lower` and, inside the tree region, columns whose node is not an ancestor-or-self of
      // the row's node; col0 = sequence column of x0[0]
      auto mask_extra = [&](uint32_t (&x0)[32], uint32_t (&x1)[32], int col0) {
        if (__any_sync(0xffffffffu, lower > col0)) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (col0 + c < lower) x0[c] = ninf;
            if (col0 + 32 + c < lower) x1[c] = ninf;
          }
        }
        if (p.mask_mode == kMaskTree && col0 + kStep > tree_start) {  // (uniform over the CTA)
          const int2* ord = reinterpret_cast<const int2*>(p.tree_order) + it.tree_beg - tree_start;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int ca = col0 + c, cb = col0 + 32 + c;
            if (ca >= tree_start && ca < kv_len) {
              const int2 par = __ldg(ord + ca);
              if (my_order < par.x || my_order >= par.y) x0[c] = ninf;
            }
            if (cb >= tree_start && cb < kv_len) {
              const int2 par = __ldg(ord + cb);
              if (my_order < par.x || my_order >= par.y) x1[c] = ninf;
            }
          }
        }
      };
      // four independent FMNMX3 chains per half
      auto half_max = [&](const uint32_t (&x0)[32], const uint32_t (&x1)[32]) {
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          mx0 = tc05::fmax3(mx0, __uint_as_float(x0[c]), __uint_as_float(x0[c + 1]));
          mx1 = tc05::fmax3(mx1, __uint_as_float(x0[c + 2]), __uint_as_float(x0[c + 3]));
          mx2 = tc05::fmax3(mx2, __uint_as_float(x1[c]), __uint_as_float(x1[c + 1]));
          mx3 = tc05::fmax3(mx3, __uint_as_float(x1[c + 2]), __uint_as_float(x1[c + 3]));
        }
        return fmaxf(tc05::fmax3(mx0, mx1, mx2), mx3);
      };
      // lazy rescale: keep the old reference max while the true max is within 2^8 of it.  `prev_half` >= 0 names the
      // S half whose PV was issued last: O_t must be complete before it is rescaled.
      auto rescale = [&](float mx, int prev_half) {
        const float m_new = fmaxf(m_used, mx * sc);
        const bool grow = m_new - m_used > kRescaleThreshold;
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2(m_used - m_new) : 1.0f;
          if (grow) {
            m_used = m_new;
            l *= alpha;
          }
          if (prev_half >= 0) {
            mbar_wait(bar(PV_DONE + 2 * t + prev_half), (n_p[prev_half] - 1) & 1);
            tc05::fence_after_sync();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              uint32_t o[32];
              tc05::ld32(t_o + cc * 32, o);
              tc05::wait_ld();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
              tc05::st32(t_o + cc * 32, o);
            }
          }
        }
      };
      // exp2, pack, store P of one 64-column half, hand it to the MMA warp.  Scale/subtract and the row sum
      // run as packed FFMA2 / FADD2; kPolyPairs of every 16 pairs take 2^x on the FMA pipe instead of the MUFU unit.
      auto do_half = [&](const uint32_t (&x0)[32], const uint32_t (&x1)[32], int hb, int si) {
        const float mneg = -m_used;
        const float2 sc2 = make_float2(sc, sc), mneg2 = make_float2(mneg, mneg);
        const uint32_t t_sb = t_s + hb * kStep;
        ++n_p[hb];
        float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
        // 2^(s*scale - m) of column pair `pi` of a 32-column chunk (pi is a compile-time index once unrolled)
        auto exp_pair = [&](uint32_t u0, uint32_t u1, int pi) {
          constexpr int kPolyPairs = kPolyPairsOf<PT>;
          const bool poly = ((pi + 1) * kPolyPairs) / 16 != (pi * kPolyPairs) / 16;
          const float2 x = tc05::ffma2(make_float2(__uint_as_float(u0), __uint_as_float(u1)), sc2, mneg2);
          float2 a;
          if (poly) {
            a = tc05::exp2_poly2(x);
          } else {
            a.x = fast_exp2(x.x);
            a.y = fast_exp2(x.y);
          }
          if (pi & 1) sum_b = tc05::fadd2(sum_b, a); else sum_a = tc05::fadd2(sum_a, a);
          return a;
        };
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float2 a = exp_pair(x0[c], x0[c + 1], c >> 1);
          pk[c >> 1] = pack_p<PT>(a.x, a.y);
        }
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float2 a = exp_pair(x1[c], x1[c + 1], c >> 1);
          pk[16 + (c >> 1)] = pack_p<PT>(a.x, a.y);
        }
#ifdef TVMB200_SYNCCHECK
        // compute-sanitizer's synccheck flags an mbarrier whose phase completes twice without a wait in between
        // ("missing wait").  PV_DONE is committed for every PV but only waited for when O is rescaled or stored; the
        // phase a waiter asks for is always its own latest P hand-off, so skipped phases are harmless.  This build
        // observes every phase (the wait returns at once: S_FULL of this tile was committed behind that PV).
        if (n_p[hb] > 1) mbar_wait(bar(PV_DONE + 2 * t + hb), (n_p[hb] - 2) & 1);
#endif
        tc05::st32(t_sb, pk);
        sum_a = tc05::fadd2(sum_a, sum_b);
        l += sum_a.x + sum_a.y;
        if (wq == 0) TRACE(1 + t, si, 5);
        tc05::wait_st();
        tc05::fence_before_sync();
        mbar_arrive(bar(P_READY + 2 * t + hb));
        if (wq == 0) TRACE(1 + t, si, 3);
      };
      // One iteration = one KV tile (128 columns): all S values are loaded and reduced to ONE row maximum /
      // rescale decision, then the two 64-column halves are exponentiated, packed and handed over one after the
      // other, so PV of the lower half runs while the upper half is still in the exp2 phase (-4 % time).
      const int my_tiles = (my_ns + 1) >> 1;
      for (int j = 0; j < my_tiles; ++j) {
        const bool has_b = 2 * j + 1 < my_ns;       // upper half visible to some row of the tile (CTA-uniform)
        if (wq == 0) TRACE(1 + t, 2 * j, 0);
        mbar_wait(bar(S_FULL + 2 * t), n_s & 1);
        mbar_wait(bar(S_FULL + 2 * t + 1), n_s & 1);  // (committed with the first one: every phase is observed)
        ++n_s;
        if (wq == 0) TRACE(1 + t, 2 * j, 1);
        tc05::fence_after_sync();
        uint32_t sa0[32], sa1[32], sb0[32], sb1[32];
        tc05::ld32(t_s + 0, sa0);
        tc05::ld32(t_s + 32, sa1);
        if (has_b) {
          tc05::ld32(t_s + 64, sb0);
          tc05::ld32(t_s + 96, sb1);
        }
        tc05::wait_ld();
        if (wq == 0) TRACE(1 + t, 2 * j, 2);
        const int col_a = ((SPLIT ? it.j0 : 0) + j) * kKV;     // sequence column of the tile's first S column
        const int rem_a = limit - col_a;
        mask_half(sa0, sa1, rem_a);
        if (XMASK) mask_extra(sa0, sa1, col_a);
        const float mx_a = half_max(sa0, sa1);
        float mx_b = -INFINITY;
        if (has_b) {
          mask_half(sb0, sb1, rem_a - kStep);
          if (XMASK) mask_extra(sb0, sb1, col_a + kStep);
          mx_b = half_max(sb0, sb1);
        }
        if (wq == 0) TRACE(1 + t, 2 * j, 4);
        rescale(fmaxf(mx_a, mx_b), j > 0 ? 1 : -1);  // the last PV of KV tile j-1 is its upper half
        do_half(sa0, sa1, 0, 2 * j);
        if (has_b) do_half(sb0, sb1, 1, 2 * j + 1);
      }
      // ---- epilogue: O / l -> global, LSE ----------------------------------------------------------------
      T* orow = nullptr;
      float* prow = nullptr;   // kv_splits > 1: this part's fp32 partial row
      if (valid) {
        const int hq = h * g + (R - tok * g);
        const int64_t at = static_cast<int64_t>(q_beg + tok) * p.num_qo_heads + hq;
        const float lse_v = l > 0.f ? m_used + log2f(l) : kNegInit;
        if (SPLIT) {
          const int64_t pat = static_cast<int64_t>(it.split) * p.total_q * p.num_qo_heads + at;
          prow = p.part_o + pat * kD;
          p.part_lse[pat] = lse_v;
        } else {
          orow = static_cast<T*>(p.output) + at * kD;
          p.lse[at] = lse_v;
        }
      }
      if (my_ns > 0) {
        const int hl = (my_ns - 1) & 1;
        mbar_wait(bar(PV_DONE + 2 * t + hl), (n_p[hl] - 1) & 1);
        tc05::fence_after_sync();
        const float inv = l > 0.f ? 1.0f / l : 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t o[32];
          tc05::ld32(t_o + cc * 32, o);
          tc05::wait_ld();
          if (SPLIT && prow != nullptr) {
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              *reinterpret_cast<float4*>(prow + cc * 32 + c) =
                  make_float4(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv, __uint_as_float(o[c + 2]) * inv,
                              __uint_as_float(o[c + 3]) * inv);
          } else if (valid) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              uint4 v;
              v.x = DT<T>::pack(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv);
              v.y = DT<T>::pack(__uint_as_float(o[c + 2]) * inv, __uint_as_float(o[c + 3]) * inv);
              v.z = DT<T>::pack(__uint_as_float(o[c + 4]) * inv, __uint_as_float(o[c + 5]) * inv);
              v.w = DT<T>::pack(__uint_as_float(o[c + 6]) * inv, __uint_as_float(o[c + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + cc * 32 + c) = v;
            }
          }
        }
        // O_t is free for the next item's first PV: that PV waits for this warpgroup's next P_READY, which these
        // very threads only signal after the loads above
        tc05::fence_before_sync();
      } else if (SPLIT && prow != nullptr) {
#pragma unroll
        for (int c = 0; c < kD; c += 4) *reinterpret_cast<float4*>(prow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (valid) {
        // no visible KV at all: O = 0, LSE = -5e4 (the reference's empty result)
#pragma unroll
        for (int c = 0; c < kD; c += 8) *reinterpret_cast<uint4*>(orow + c) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }

  // ---- teardown ------------------------------------------------------------------------------------------------
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc05::fence_after_sync();
    tc05::tmem_dealloc(tmem, 512);
  }
}

// kv_splits > 1: O = sum_s w_s O_s / sum_s w_s, w_s = 2^(lse_s - max lse), LSE = max + log2(sum w) over the parts of a row
// (the arithmetic of f_merge_inplace); a part that saw no visible column carries the -5e4 sentinel and weighs nothing.
// One warp per (token, head) row, four dims per lane.
template <typename T>
__global__ void __launch_bounds__(256)
prefill_merge_splits_kernel(const float* __restrict__ part_o, const float* __restrict__ part_lse, T* __restrict__ output,
                            float* __restrict__ lse, int64_t rows, int splits) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float mm = kNegInit;
  for (int s = 0; s < splits; ++s) mm = fmaxf(mm, part_lse[s * rows + row]);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float den = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float ls = part_lse[s * rows + row];
    if (ls <= kNegInit) continue;
    const float w = exp2f(ls - mm);
    const float4 v = *reinterpret_cast<const float4*>(part_o + (s * rows + row) * kD + lane * 4);
    acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    den += w;
  }
  const float inv = den > 0.f ? 1.f / den : 0.f;
  uint2 pk;
  pk.x = DT<T>::pack(acc.x * inv, acc.y * inv);
  pk.y = DT<T>::pack(acc.z * inv, acc.w * inv);
  *reinterpret_cast<uint2*>(output + row * kD + lane * 4) = pk;
  if (lane == 0) lse[row] = den > 0.f ? mm + log2f(den) : kNegInit;
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static std::atomic<int> g_prefill_impl{0};   // 0 auto, 1 force generic, 2 force tcgen05 (where eligible)
std::atomic<int64_t> g_kv_split_launches{0};  // tcgen05 launches that cut their items' KV range (prefill_api.cu reports it)

bool tc05_eligible(const PrefillParams& p, bool paged, int total_q_len, int head_dim) {
  const int impl = g_prefill_impl.load();
  if (impl == 1) return false;
  // (inline RoPE and per-sequence sliding windows reach this kernel through the gather / rotate pre-pass of
  // prefill_api.cu, which hands over position-ordered, rotated ragged K / V: rotary_mode 0, no slot remap)
  if (head_dim != kD || p.rotary_mode != 0 || p.sliding) return false;
  const int g = p.group;
  if (!(g == 1 || g == 2 || g == 4 || g == 8 || g == 16)) return false;
  if (p.batch > 2048) return false;
  if (impl == 2) return true;
  // auto: the 256-row tiles only pay off once there is enough work to fill them
  return static_cast<int64_t>(total_q_len) * g >= 2048;
}

// How many parts to cut every item's KV range into.  Known on the host: the number of 256-row items (bounds) and the
// average KV length; wanted: the launch's makespan on `grid` CTAs, ceil(items * S / grid) rounds of ceil(tiles / S) KV
// tiles, as short as possible without making parts shorter than 8 tiles.  S = 1 whenever the items already fill the
// machine several times over (C3) or the contexts are short.
static int choose_kv_splits(int64_t items, int64_t avg_kv_len, int grid) {
  const int64_t tiles = (avg_kv_len + kKV - 1) / kKV;
  if (items <= 0 || tiles < 16 || items >= 4 * static_cast<int64_t>(grid)) return 1;
  int best = 1;
  int64_t best_cost = ((items + grid - 1) / grid) * tiles;
  for (int s = 2; s <= 8 && tiles / s >= 8; ++s) {
    const int64_t cost = ((items * s + grid - 1) / grid) * ((tiles + s - 1) / s) + 2 * s;  // (+ a little per extra part)
    if (cost < best_cost) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

template <typename T, typename PT, bool PAGED>
static int launch_tc05_t(const PrefillParams& p_in, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                         int total_q_len, cudaStream_t st, int64_t avg_kv_len) {
  PrefillParams p = p_in;
  const int64_t max_pairs = (static_cast<int64_t>(total_q_len) * p.group + 2 * kRows - 1) / (2 * kRows) + p.batch;
  // (for the split decision: pairs without the per-sequence rounding slack)
  const int64_t est_items = ((static_cast<int64_t>(total_q_len) * p.group + 2 * kRows - 1) / (2 * kRows)) * p.num_kv_heads;
  p.kv_splits = avg_kv_len > 0 ? choose_kv_splits(est_items > p.batch * p.num_kv_heads ? est_items : p.batch * p.num_kv_heads,
                                                  avg_kv_len, num_sms()) : 1;
  p.total_q = total_q_len;
  {  // partials are capped at 256 MiB of the context's scratch
    const int64_t rows = static_cast<int64_t>(total_q_len) * p.num_qo_heads;
    while (p.kv_splits > 1 && p.kv_splits * rows * (kD + 1) * 4 > (int64_t(256) << 20)) --p.kv_splits;
  }
  if (p.kv_splits > 1) {
    g_kv_split_launches++;
    const int64_t rows = static_cast<int64_t>(total_q_len) * p.num_qo_heads;
    const int64_t lse_bytes = (p.kv_splits * rows * 4 + 255) / 256 * 256;
    void* ws = nullptr;
    if (int rc = get_workspace(lse_bytes + p.kv_splits * rows * kD * 4, st, &ws)) return rc;
    p.part_lse = static_cast<float*>(ws);
    p.part_o = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + lse_bytes);
  }
  const int64_t max_items = max_pairs * p.num_kv_heads * p.kv_splits;
  const int grid = static_cast<int>(max_items < num_sms() ? max_items : num_sms());  // persistent: one CTA per SM
  const size_t tail = (static_cast<size_t>(p.batch) + 1 + 40) * sizeof(int);
  const bool xmask = p.mask_mode == kMaskLayerSliding || p.mask_mode == kMaskTree;
  const bool split = p.kv_splits > 1;
  // paged bf16 (pages come from HBM once, V passes through the converter warps): a V ring of three slots, when the fifth
  // tile still fits next to the batch's scan array (227 KiB of shared memory per CTA).  TVMB200_PREFILL_VR3=0: A/B knob
  constexpr bool kCanVr3 = PAGED && !std::is_same<T, PT>::value;  // (fp16 pages: 946 TFLOP/s on C5 with either ring)
  static const bool vr3_on = [] {
    const char* e = getenv("TVMB200_PREFILL_VR3");
    return !(e && atoi(e) == 0);
  }();
  const bool vr3 = kCanVr3 && vr3_on && 1024 + SmemLayoutT<5>::scan + tail <= 227 * 1024;
  const size_t smem = 1024 + (vr3 ? SmemLayoutT<5>::scan : SmemLayoutT<4>::scan) + tail;
  auto kern = xmask ? (split ? prefill_tc05_kernel<T, PT, PAGED, true, true> : prefill_tc05_kernel<T, PT, PAGED, true, false>)
                    : (split ? prefill_tc05_kernel<T, PT, PAGED, false, true> : prefill_tc05_kernel<T, PT, PAGED, false, false>);
  if constexpr (kCanVr3) {
    if (vr3)
      kern = xmask ? (split ? prefill_tc05_kernel<T, PT, PAGED, true, true, true> : prefill_tc05_kernel<T, PT, PAGED, true, false, true>)
                   : (split ? prefill_tc05_kernel<T, PT, PAGED, false, true, true> : prefill_tc05_kernel<T, PT, PAGED, false, false, true>);
  }
  TVMB200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  constexpr uint32_t fa = std::is_same<T, __half>::value ? 0u : 1u;
  constexpr uint32_t fp = std::is_same<PT, __half>::value ? 0u : 1u;
  const uint32_t idesc_qk = tc05::make_idesc(fa, fa, 0, 0, kRows, kKV);
  // V is converted to the P format in shared memory when the two differ (bf16 inputs, fp16 P)
  const uint32_t idesc_pv = tc05::make_idesc(fp, fp, 0, 1, kRows, kD);
  // work-queue counter of this (context, device, stream)
  int32_t* counter = nullptr;
  if (int rc = get_counters(st, &counter)) return rc;
  // the kernel resets the counter with its last fetch; the memset only matters after an aborted launch
  TVMB200_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
  kern<<<static_cast<unsigned>(grid), kThreads, smem, st>>>(tq, tk, tv, p, idesc_qk, idesc_pv, counter);
  TVMB200_LAUNCH_OK();
  if (p.kv_splits > 1) {
    const int64_t rows = static_cast<int64_t>(total_q_len) * p.num_qo_heads;
    prefill_merge_splits_kernel<T><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(
        p.part_o, p.part_lse, static_cast<T*>(p.output), p.lse, rows, p.kv_splits);
    TVMB200_LAUNCH_OK();
  }
  return 0;
}

// avg_kv_len > 0 allows the KV split (the caller passes 0 when the context's scratch is already in use: pre-pass route)
int launch_prefill_tc05(const PrefillParams& p, bool paged, int total_q_len, int total_kv_len, int64_t num_pages,
                        int dtype, cudaStream_t st, int64_t avg_kv_len) {
  CUtensorMap tq, tk, tv;
  const int g = p.group;
  const uint64_t row_q = static_cast<uint64_t>(p.num_qo_heads) * kD * 2;
  if (int rc = make_tmap_3d(&tq, p.q, dtype, kD, p.num_qo_heads, total_q_len, kD * 2, row_q, g, kRows / g)) return rc;
  if (paged) {
    const uint64_t rows = static_cast<uint64_t>(num_pages) * 2 * p.num_kv_heads * 16;
    if (int rc = get_tmap_2d_cached(&tk, p.pages, dtype, rows, kD, 16)) return rc;
    tv = tk;
  } else {
    const uint64_t row_kv = static_cast<uint64_t>(p.num_kv_heads) * kD * 2;
    if (int rc = make_tmap_3d(&tk, p.k, dtype, kD, p.num_kv_heads, total_kv_len, kD * 2, row_kv, 1, kKV)) return rc;
    if (int rc = make_tmap_3d(&tv, p.v, dtype, kD, p.num_kv_heads, total_kv_len, kD * 2, row_kv, 1, kKV)) return rc;
  }
  if (dtype == TVMB200_F16)
    return paged ? launch_tc05_t<__half, __half, true>(p, tq, tk, tv, total_q_len, st, avg_kv_len)
                 : launch_tc05_t<__half, __half, false>(p, tq, tk, tv, total_q_len, st, avg_kv_len);
  return paged ? launch_tc05_t<__nv_bfloat16, __half, true>(p, tq, tk, tv, total_q_len, st, avg_kv_len)
               : launch_tc05_t<__nv_bfloat16, __half, false>(p, tq, tk, tv, total_q_len, st, avg_kv_len);
}

}  // namespace tvmb200

extern "C" void tvmb200_set_prefill_impl(int impl) { tvmb200::g_prefill_impl.store(impl); }

#ifdef TVMB200_TRACE
extern "C" __attribute__((visibility("default"))) int tvmb200_debug_prefill_trace(long long* out) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, tvmb200::g_trace, sizeof(tvmb200::g_trace)));
}
#endif
