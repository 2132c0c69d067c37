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
};

int launch_prefill_generic(const PrefillParams& p, bool paged, int total_q_len, int head_dim, int dtype,
                           cudaStream_t st);

// tcgen05 / TMEM path (prefill_tc05.cu): D = 128, rotary_mode 0, mask none / causal, no sliding window
bool tc05_eligible(const PrefillParams& p, bool paged, int total_q_len, int head_dim);
int launch_prefill_tc05(const PrefillParams& p, bool paged, int total_q_len, int total_kv_len, int64_t num_pages,
                        int dtype, cudaStream_t st);

}  // namespace tvmb200
