// C entry points of the four prefill callbacks: argument validation, parameter block, dispatch.
#include "prefill.cuh"

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
  if (tc05_eligible(p, true, total_q_len, head_dim))
    return launch_prefill_tc05(p, true, total_q_len, 0, num_pages, dtype, static_cast<cudaStream_t>(stream));
  return launch_prefill_generic(p, true, total_q_len, head_dim, dtype, static_cast<cudaStream_t>(stream));
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
  if (tc05_eligible(p, false, total_q_len, head_dim))
    return launch_prefill_tc05(p, false, total_q_len, total_kv_len, 0, dtype, static_cast<cudaStream_t>(stream));
  return launch_prefill_generic(p, false, total_q_len, head_dim, dtype, static_cast<cudaStream_t>(stream));
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
  return launch_prefill_generic(p, false, total_q_len, head_dim, dtype, static_cast<cudaStream_t>(stream));
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
  return launch_prefill_generic(p, true, total_q_len, head_dim, dtype, static_cast<cudaStream_t>(stream));
}
