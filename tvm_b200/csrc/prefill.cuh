// Shared parameter block of the prefill kernels (generic mma.sync path and tcgen05 path).
#pragma once
#include "common.cuh"

namespace tvmb200 {

enum MaskMode : int { kMaskNone = 0, kMaskCausal = 1, kMaskLayerSliding = 2, kMaskTree = 3 };

struct PrefillParams {
  const void* q;            // [n, Hq, D]
  const int32_t* q_indptr;  // [B+1]
  // paged KV
  const void* pages;  // [P, 2, Hkv, 16, D]
  const int32_t* page_indptr;
  const int32_t* page_values;
  const int32_t* length_info;  // [B] or [3,B]
  // ragged KV
  const void* k;  // [m, Hkv, D]
  const void* v;
  const int32_t* kv_indptr;
  // rope
  const int32_t* k_rope_pos_offset;  // [B]
  const int32_t* q_rope_position;    // [n]
  // tree
  const int32_t* tree_indptr;  // [B+1]
  const int32_t* tree_order;   // [tree_size, 2]
  void* output;                // [n, Hq, D]
  float* lse;                  // [n, Hq]
  int batch;
  int num_qo_heads;
  int num_kv_heads;
  int group;
  int sliding;  // length_info is [3,B]
  int mask_mode;
  int layer_sws;  // layer sliding window size (kMaskLayerSliding)
  int rotary_mode;
  int tree_k_rope;  // tree ragged: K rope position is q_rope_position[kv row] (tree_attn.py:429)
  float rope_scale;
  float rope_theta;
  RopeScaling rs;
  float scale_log2;
  // tcgen05 kernel only: the KV range of every work item is cut into kv_splits parts (launches with few, long items:
  // a 64-node tree against a 32K context is one 256-row item per (sequence, kv head)); parts write normalised fp32
  // partials [kv_splits, n, Hq, D] / [kv_splits, n, Hq] that prefill_merge_splits_kernel reduces
  int kv_splits;
  int total_q;
  float* part_o;
  float* part_lse;
};

int launch_prefill_generic(const PrefillParams& p, bool paged, int total_q_len, int head_dim, int dtype,
                           cudaStream_t st);

// gather / rotate pre-pass (prefill_prepass.cu): position-ordered, rotated ragged q / K / V for the tcgen05 kernel
struct PrepassParams {
  const void* q;                     // [n_q, hq, 128]  (rotated into q_out when rotary)
  const int32_t* q_rope_position;    // [n_q]
  const void* pages;                 // paged source
  const int32_t* page_indptr;
  const int32_t* page_values;
  const int32_t* length_info;        // [B] or [3, B]
  const void* k;                     // ragged source (rotary only; V is used in place)
  const int32_t* kv_indptr;
  const int32_t* k_rope_pos_offset;  // [B]
  void* q_out;
  void* k_out;                       // [kv_rows_bound, hkv, 128]
  void* v_out;                       // paged sources only
  int32_t* kv_indptr_out;            // paged sources only: [B + 1], written by the pass
  int64_t n_q;
  int64_t kv_rows_bound;             // rows k_out / v_out can hold (paged: nnz_pages * 16; ragged: total_kv_len)
  int batch, hq, hkv;
  int sliding;                       // length_info is [3, B]
  int rotary;                        // rotate q and K
  int tree_k_rope;                   // K row r is rotated at q_rope_position[r] (tree_attn.py:429)
  float rope_scale, rope_theta;
  RopeScaling rs;
};
int64_t prepass_scratch_bytes(bool paged, bool rotary, int64_t n_q, int hq, int hkv, int64_t kv_rows_bound, int batch,
                              int64_t off[4]);
int launch_prefill_prepass(const PrepassParams& a, bool paged, int dtype, cudaStream_t st);

// tcgen05 / TMEM path (prefill_tc05.cu): D = 128, rotary_mode 0, no per-sequence slot remap; masks: none, causal, layer
// sliding window, token tree
bool tc05_eligible(const PrefillParams& p, bool paged, int total_q_len, int head_dim);
int launch_prefill_tc05(const PrefillParams& p, bool paged, int total_q_len, int total_kv_len, int64_t num_pages,
                        int dtype, cudaStream_t st, int64_t avg_kv_len = 0);

}  // namespace tvmb200
