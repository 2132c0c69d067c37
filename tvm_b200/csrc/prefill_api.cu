// C entry points of the four prefill callbacks: argument validation, parameter block, dispatch.
#include "prefill.cuh"

#include <atomic>

using namespace tvmb200;

static int check_common(const char* name, int dtype, int head_dim, int num_qo_heads, int num_kv_heads,
                        int batch_size, int rotary_mode) {
  if (rotary_mode == 1)
    if (int rc = check_no_rope_variant(name)) return rc;
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "%s: unsupported dtype %d", name, dtype);
  TVMB200_CHECK(head_dim == 128 || head_dim == 64, "%s: head_dim %d unsupported (64 or 128)", name, head_dim);
  TVMB200_CHECK(num_kv_heads > 0 && num_qo_heads % num_kv_heads == 0,
                "%s: num_qo_heads %d not a multiple of num_kv_heads %d", name, num_qo_heads, num_kv_heads);
  TVMB200_CHECK(batch_size >= 0 && batch_size <= 8192, "%s: batch %d out of range [0, 8192]", name, batch_size);
  return 0;
}

static void fill_base(PrefillParams& p, const void* q, const int32_t* q_indptr, void* output, float* lse,
                      int batch, int hq, int hkv, int rotary_mode, float rope_scale, float rope_theta,
                      float sm_scale) {
  p = PrefillParams{};
  p.q = q;
  p.q_indptr = q_indptr;
  p.output = output;
  p.lse = lse;
  p.batch = batch;
  p.num_qo_heads = hq;
  p.num_kv_heads = hkv;
  p.group = hq / hkv;
  p.rotary_mode = rotary_mode;
  p.rope_scale = rope_scale;
  p.rope_theta = rope_theta;
  p.rs = rope_scaling();
  p.scale_log2 = sm_scale * kLog2e;
}

// Dispatch of one prefill callback: the tcgen05 kernel directly when the inputs already have the shape it wants; through
// the gather / rotate pre-pass when they need inline RoPE or the sliding-window slot remap and the scratch that takes
// stays under the cap; the mma.sync kernel otherwise (head_dim 64, odd GQA groups, tiny batches).
static std::atomic<int64_t> g_path_counts[3];  // launches by path: generic, tcgen05, tcgen05 behind the pre-pass
static std::atomic<int64_t> g_prepass_cap{int64_t(2) << 30};

static int dispatch_prefill(PrefillParams& p, bool paged, int total_q_len, int total_kv_len, int32_t nnz_pages,
                            int64_t num_pages, int head_dim, int dtype, cudaStream_t st) {
  const bool needs_prepass = p.rotary_mode == 1 || p.sliding;
  if (!needs_prepass) {
    if (tc05_eligible(p, paged, total_q_len, head_dim)) {
      g_path_counts[1]++;
      const int64_t kv_total = paged ? static_cast<int64_t>(nnz_pages) * 16 : total_kv_len;
      return launch_prefill_tc05(p, paged, total_q_len, total_kv_len, num_pages, dtype, st,
                                 p.batch > 0 ? kv_total / p.batch : 0);
    }
    g_path_counts[0]++;
    return launch_prefill_generic(p, paged, total_q_len, head_dim, dtype, st);
  }
  PrefillParams t = p;
  t.rotary_mode = 0;
  t.sliding = 0;
  t.tree_k_rope = 0;
  const bool rotary = p.rotary_mode == 1;
  const int64_t kv_rows = paged ? static_cast<int64_t>(nnz_pages) * 16 : total_kv_len;
  int64_t off[4];
  const int64_t bytes = prepass_scratch_bytes(paged, rotary, total_q_len, p.num_qo_heads, p.num_kv_heads, kv_rows, p.batch, off);
  if (!tc05_eligible(t, false, total_q_len, head_dim) || bytes > g_prepass_cap.load() || kv_rows == 0) {
    g_path_counts[0]++;
    return launch_prefill_generic(p, paged, total_q_len, head_dim, dtype, st);
  }
  void* ws = nullptr;
  if (int rc = get_workspace(bytes, st, &ws)) return rc;
  uint8_t* base = static_cast<uint8_t*>(ws);
  PrepassParams a = {};
  a.q = p.q;
  a.q_rope_position = p.q_rope_position;
  a.pages = p.pages;
  a.page_indptr = p.page_indptr;
  a.page_values = p.page_values;
  a.length_info = p.length_info;
  a.k = p.k;
  a.kv_indptr = p.kv_indptr;
  a.k_rope_pos_offset = p.k_rope_pos_offset;
  a.kv_indptr_out = reinterpret_cast<int32_t*>(base + off[0]);
  a.q_out = base + off[1];
  a.k_out = base + off[2];
  a.v_out = base + off[3];
  a.n_q = total_q_len;
  a.kv_rows_bound = kv_rows;
  a.batch = p.batch;
  a.hq = p.num_qo_heads;
  a.hkv = p.num_kv_heads;
  a.sliding = p.sliding;
  a.rotary = rotary ? 1 : 0;
  a.tree_k_rope = p.tree_k_rope;
  a.rope_scale = p.rope_scale;
  a.rope_theta = p.rope_theta;
  a.rs = p.rs;
  if (int rc = launch_prefill_prepass(a, paged, dtype, st)) return rc;
  if (rotary) t.q = a.q_out;
  t.k = a.k_out;
  t.v = paged ? a.v_out : p.v;
  t.kv_indptr = paged ? a.kv_indptr_out : p.kv_indptr;
  t.pages = nullptr;
  g_path_counts[2]++;
  return launch_prefill_tc05(t, false, total_q_len, static_cast<int>(kv_rows), 0, dtype, st);
}

// [0] mma.sync kernel, [1] tcgen05 kernel, [2] tcgen05 kernel behind the gather / rotate pre-pass, [3] tcgen05 launches
// (of [1]) that cut their items' KV range into parts (tests: no silent fallback on the shapes the tensor-core path covers)
namespace tvmb200 {
extern std::atomic<int64_t> g_kv_split_launches;
}
extern "C" TVMB200_API void tvmb200_debug_prefill_path_counts(int64_t out[4]) {
  for (int i = 0; i < 3; ++i) out[i] = g_path_counts[i].load();
  out[3] = tvmb200::g_kv_split_launches.load();
}
// upper bound of the pre-pass scratch (bytes) above which inline-RoPE / sliding-window prefill stays on the mma.sync kernel
extern "C" TVMB200_API void tvmb200_set_prefill_prepass_cap(int64_t bytes) { g_prepass_cap.store(bytes); }

extern "C" int tvmb200_attention_prefill_paged(
    const void* q, const int32_t* q_indptr, const void* pages, const int32_t* page_indptr,
    const int32_t* page_values, const int32_t* length_info, const int32_t* k_rope_pos_offset,
    const int32_t* q_rope_position, void* output, float* lse, int32_t batch_size, int32_t total_q_len,
    int32_t nnz_pages, int64_t num_pages, int32_t num_qo_heads, int32_t num_kv_heads, int32_t page_size,
    int32_t head_dim, int sliding_window, int32_t layer_sliding_window_size, int causal,
    int rotary_mode, float rope_scale, float rope_theta, float sm_scale, int dtype,
    tvmb200_stream_t stream) {
  if (int rc = check_common("attention_prefill", dtype, head_dim, num_qo_heads, num_kv_heads, batch_size, rotary_mode)) return rc;
  TVMB200_CHECK(page_size == 16, "attention_prefill: page_size %d unsupported (16)", page_size);
  TVMB200_CHECK(rotary_mode == 0 || rotary_mode == 1, "attention_prefill: rotary_mode %d", rotary_mode);
  if (batch_size == 0 || total_q_len == 0) return 0;
  PrefillParams p;
  fill_base(p, q, q_indptr, output, lse, batch_size, num_qo_heads, num_kv_heads, rotary_mode, rope_scale, rope_theta, sm_scale);
  p.pages = pages;
  p.page_indptr = page_indptr;
  p.page_values = page_values;
  p.length_info = length_info;
  p.k_rope_pos_offset = k_rope_pos_offset;
  p.q_rope_position = q_rope_position;
  p.sliding = sliding_window ? 1 : 0;
  // _kernel_common.py:138-144: the layer-sliding mask replaces the causal mask only when
  // causal > 0 and the kernel was built with a sliding window size
  if (causal > 0 && sliding_window && layer_sliding_window_size > 0) {
    p.mask_mode = kMaskLayerSliding;
    p.layer_sws = layer_sliding_window_size;
  } else {
    p.mask_mode = causal > 0 ? kMaskCausal : kMaskNone;
  }
  return dispatch_prefill(p, true, total_q_len, 0, nnz_pages, num_pages, head_dim, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int tvmb200_attention_prefill_ragged(
    const void* q, const int32_t* q_indptr, const void* k, const void* v, const int32_t* kv_indptr,
    const int32_t* q_rope_position, const int32_t* k_rope_pos_offset, void* output, float* lse,
    int32_t batch_size, int32_t total_q_len, int32_t total_kv_len, int32_t num_qo_heads,
    int32_t num_kv_heads, int32_t head_dim, int causal, int rotary_mode, float rope_scale,
    float rope_theta, float sm_scale, int dtype, tvmb200_stream_t stream) {
  if (int rc = check_common("attention_prefill_ragged", dtype, head_dim, num_qo_heads, num_kv_heads, batch_size, rotary_mode)) return rc;
  TVMB200_CHECK(rotary_mode == 0 || rotary_mode == 1, "attention_prefill_ragged: rotary_mode %d", rotary_mode);
  if (batch_size == 0 || total_q_len == 0) return 0;
  PrefillParams p;
  fill_base(p, q, q_indptr, output, lse, batch_size, num_qo_heads, num_kv_heads, rotary_mode, rope_scale, rope_theta, sm_scale);
  p.k = k;
  p.v = v;
  p.kv_indptr = kv_indptr;
  p.k_rope_pos_offset = k_rope_pos_offset;
  p.q_rope_position = q_rope_position;
  p.mask_mode = causal > 0 ? kMaskCausal : kMaskNone;
  return dispatch_prefill(p, false, total_q_len, total_kv_len, 0, 0, head_dim, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int tvmb200_attention_prefill_tree_ragged(
    const void* q, const int32_t* q_indptr, const void* k, const void* v, const int32_t* kv_indptr,
    const int32_t* q_rope_position, const int32_t* mn_indptr, const int32_t* mask, void* output,
    float* lse, int32_t batch_size, int32_t total_q_len, int32_t total_kv_len, int32_t num_qo_heads,
    int32_t num_kv_heads, int32_t head_dim, int rotary_mode, float rope_scale, float rope_theta,
    float sm_scale, int dtype, tvmb200_stream_t stream) {
  if (int rc = check_common("attention_prefill_with_tree_mask", dtype, head_dim, num_qo_heads, num_kv_heads, batch_size, rotary_mode)) return rc;
  TVMB200_CHECK(rotary_mode == 0 || rotary_mode == 1, "attention_prefill_with_tree_mask: rotary_mode %d", rotary_mode);
  if (batch_size == 0 || total_q_len == 0) return 0;
  PrefillParams p;
  fill_base(p, q, q_indptr, output, lse, batch_size, num_qo_heads, num_kv_heads, rotary_mode, rope_scale, rope_theta, sm_scale);
  p.k = k;
  p.v = v;
  p.kv_indptr = kv_indptr;
  p.q_rope_position = q_rope_position;
  p.tree_k_rope = 1;
  p.tree_indptr = mn_indptr;
  p.tree_order = mask;
  p.mask_mode = kMaskTree;
  return dispatch_prefill(p, false, total_q_len, total_kv_len, 0, 0, head_dim, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int tvmb200_attention_prefill_tree_paged(
    const void* q, const int32_t* q_indptr, const void* pages, const int32_t* page_indptr,
    const int32_t* page_values, const int32_t* length_info, const int32_t* k_rope_pos_offset,
    const int32_t* q_rope_position, void* output, float* lse, int32_t batch_size, int32_t total_q_len,
    int32_t nnz_pages, int64_t num_pages, int32_t num_qo_heads, int32_t num_kv_heads, int32_t page_size,
    int32_t head_dim, int rotary_mode, float rope_scale, float rope_theta, float sm_scale,
    const int32_t* tree_order_indptr, const int32_t* tree_order, int dtype, tvmb200_stream_t stream) {
  if (int rc = check_common("attention_prefill_with_tree_mask_paged_kv", dtype, head_dim, num_qo_heads, num_kv_heads, batch_size, rotary_mode)) return rc;
  TVMB200_CHECK(page_size == 16, "attention_prefill_with_tree_mask_paged_kv: page_size %d unsupported (16)", page_size);
  // the reference asserts this too (tree_attn.py:699, 930)
  TVMB200_CHECK(rotary_mode == 0, "Inline rotary mode is not supported in tree attention.");
  if (batch_size == 0 || total_q_len == 0) return 0;
  PrefillParams p;
  fill_base(p, q, q_indptr, output, lse, batch_size, num_qo_heads, num_kv_heads, rotary_mode, rope_scale, rope_theta, sm_scale);
  p.pages = pages;
  p.page_indptr = page_indptr;
  p.page_values = page_values;
  p.length_info = length_info;
  p.k_rope_pos_offset = k_rope_pos_offset;
  p.q_rope_position = q_rope_position;
  p.tree_indptr = tree_order_indptr;
  p.tree_order = tree_order;
  p.mask_mode = kMaskTree;
  return dispatch_prefill(p, true, total_q_len, 0, nnz_pages, num_pages, head_dim, dtype, static_cast<cudaStream_t>(stream));
}
