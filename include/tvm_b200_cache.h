/*
 * tvm_b200_cache.h -- C ABI of the host-side paged KV cache that drives the sm_100a kernel set.
 *
 * It is the B200-side counterpart of the reference's C++ object `PagedAttentionKVCacheObj`
 * (/root/reference/src/runtime/vm/paged_kv_cache.cc:75-2526) and of the `vm.builtin.kv_state_*` /
 * `vm.builtin.attention_kv_cache_*` packed functions registered in src/runtime/vm/kv_state.cc:33-116: the same
 * entry points, argument meaning and error behaviour, with plain pointers instead of ffi objects.  Page / block /
 * sequence bookkeeping and every int32 auxiliary array handed to the kernels are bit-identical to the reference
 * (pinned by tests/golden/kvcache_*.npz, captured from the reference itself).
 *
 * Conventions: every function returns 0 on success; on error nothing is launched and the message is in
 * tvmb200_last_error() (include/tvm_b200.h).  Device pointers are raw CUDA device pointers.
 */
#ifndef TVM_B200_CACHE_H_
#define TVM_B200_CACHE_H_

#include <stdint.h>

#include "tvm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tvmb200_cache_s* tvmb200_cache_t;

/* AttnKind (attn_utils.h:66-71) and RoPEMode (attn_utils.h:341-345) values accepted here */
#define TVMB200_ATTN_MHA 0
#define TVMB200_ATTN_MHA_SLIDING 3
#define TVMB200_ROPE_NONE 0
#define TVMB200_ROPE_NORMAL 1
#define TVMB200_ROPE_INLINE 2

/*! \brief Arguments of `vm.builtin.paged_attention_kv_cache_create` (paged_kv_cache.cc:2535-2639) that are not callbacks. */
typedef struct {
  int64_t reserved_num_seqs;         /* cache_config[0] */
  int64_t total_token_capacity;      /* cache_config[1] */
  int64_t prefill_chunk_size;        /* cache_config[2] */
  int64_t page_size;                 /* cache_config[3] (the sm_100a kernels need 16) */
  int32_t support_sliding_window;    /* cache_config[4] */
  int64_t layer_sliding_window_size; /* cache_config[5]; <= 0 means the reference default 1024 */
  int64_t layer_id_begin_offset;     /* layer_indptr[group], pipeline-stage slice */
  int64_t num_layers;                /* layer_indptr[group+1] - layer_indptr[group] */
  int64_t num_qo_heads;
  int64_t num_kv_heads;
  int64_t head_dim;
  const int32_t* attn_kinds;         /* [layer_id_begin_offset + num_layers] or NULL = all MHA */
  int32_t rope_mode;
  double rotary_scale;
  double rotary_theta;
  int32_t dtype;                     /* TVMB200_F16 / TVMB200_BF16 */
  int32_t device_id;                 /* CUDA device; -1 = planning-only cache: bookkeeping + call trace, it owns no
                                        device memory and launches nothing (attention calls only record the plan) */
} tvmb200_cache_config;

TVMB200_API int tvmb200_cache_create(const tvmb200_cache_config* cfg, tvmb200_cache_t* out);
TVMB200_API void tvmb200_cache_destroy(tvmb200_cache_t c);

/* vm.builtin.kv_state_* (kv_state.cc:33-53) */
TVMB200_API int tvmb200_cache_clear(tvmb200_cache_t c);
TVMB200_API int tvmb200_cache_add_sequence(tvmb200_cache_t c, int64_t seq_id);
TVMB200_API int tvmb200_cache_remove_sequence(tvmb200_cache_t c, int64_t seq_id);
TVMB200_API int tvmb200_cache_fork_sequence(tvmb200_cache_t c, int64_t parent_seq_id, int64_t child_seq_id, int64_t fork_pos);
TVMB200_API int tvmb200_cache_popn(tvmb200_cache_t c, int64_t seq_id, int32_t n);
/*! \brief token_tree_parent_ptr may be NULL (no tree); otherwise it has sum(append_lengths) entries.  When the call
 *  fails (unknown sequence, cache full, invalid tree, empty batch) the cache is left exactly as it was before the call
 *  (sequence lengths, pages, free-page stack) and there is no current batch: the attention entries,
 *  commit_accepted_token_tree_nodes and get_query_positions return an error until a begin_forward completes. */
TVMB200_API int tvmb200_cache_begin_forward(tvmb200_cache_t c, const int64_t* seq_ids, const int64_t* append_lengths,
                                            int32_t batch_size, const int64_t* token_tree_parent_ptr, int32_t tree_size);
TVMB200_API int tvmb200_cache_end_forward(tvmb200_cache_t c);

/* vm.builtin.attention_kv_cache_* (kv_state.cc:57-116) */
TVMB200_API int tvmb200_cache_enable_sliding_window_for_seq(tvmb200_cache_t c, int64_t seq_id, int32_t sliding_window_size,
                                                            int32_t attn_sink_size);
TVMB200_API int tvmb200_cache_commit_accepted_token_tree_nodes(tvmb200_cache_t c, const int64_t* seq_ids,
                                                               const int64_t* leaf_indices, int32_t n);
TVMB200_API int tvmb200_cache_empty(tvmb200_cache_t c, int32_t* out);
TVMB200_API int tvmb200_cache_get_num_available_pages(tvmb200_cache_t c, int32_t* out);
TVMB200_API int tvmb200_cache_get_total_sequence_length(tvmb200_cache_t c, int32_t* out);
/*! \brief device pointer + length of q_rope_position_map (AttentionKVCacheObj::GetQueryPositions). */
TVMB200_API int tvmb200_cache_get_query_positions(tvmb200_cache_t c, const int32_t** dev_ptr, int64_t* n,
                                                  tvmb200_stream_t stream);
/*!
 * \brief attention_with_fused_qkv(layer_id, sm_scale, qkv [n, Hq+2Hkv, D], o [n, Hq, D]) on `stream`
 *        (AttentionWithFusedQKV, paged_kv_cache.cc:1303-1402): split+rope -> (append) -> attention over every
 *        block depth with in-place LSE merge -> (append).
 */
TVMB200_API int tvmb200_cache_attention_with_fused_qkv(tvmb200_cache_t c, int64_t layer_id, double sm_scale,
                                                       const void* qkv, void* o, int64_t qkv_rows,
                                                       tvmb200_stream_t stream);
/*!
 * \brief self_attention(layer_id, sm_scale, q [n, Hq, D], k, v [n, Hkv, D], o [n, Hq, D], lse [n, Hq] f32)
 *  (`vm.builtin.attention_kv_cache_self_attention`, kv_state.cc:84-90; SelfAttention, paged_kv_cache.cc:1404-1445 ->
 *  MHASelfAttnInternal :2182-2206): the step's tokens against themselves (causal, or the token-tree mask); nothing is
 *  read from or appended to the pages.  rows = n must equal the batch's total append length.
 */
TVMB200_API int tvmb200_cache_self_attention(tvmb200_cache_t c, int64_t layer_id, double sm_scale, const void* q,
                                             const void* k, const void* v, void* o, float* lse, int64_t rows,
                                             tvmb200_stream_t stream);
/*!
 * \brief cross_attention(layer_id, sm_scale, q, o, lse) (`vm.builtin.attention_kv_cache_cross_attention`,
 *  kv_state.cc:91-96; CrossAttention, paged_kv_cache.cc:1447-1485 -> MHACrossAttnInternal :2216-2299 with
 *  is_first_kernel = true, causal = false): q against the cached KV of every block depth, merged in place.  When no
 *  depth holds a page, (o, lse) are left untouched, as in the reference.
 */
TVMB200_API int tvmb200_cache_cross_attention(tvmb200_cache_t c, int64_t layer_id, double sm_scale, const void* q, void* o,
                                              float* lse, int64_t rows, tvmb200_stream_t stream);
/*!
 * \brief attention_with_shared_kv(source_layer_id, sm_scale, q, current_k, current_v, o)
 *  (`vm.builtin.attention_kv_cache_attention_with_shared_kv`, kv_state.cc:97-104; AttentionWithSharedKV,
 *  paged_kv_cache.cc:1487-1529): a layer without its own KV queries the K / V of `source_layer_id` after that layer's
 *  attention_with_fused_qkv of the same step; current_k / current_v (already rotated) are only read when the batch
 *  appends after the attention.
 */
TVMB200_API int tvmb200_cache_attention_with_shared_kv(tvmb200_cache_t c, int64_t source_layer_id, double sm_scale,
                                                       const void* q, const void* current_k, const void* current_v,
                                                       void* o, int64_t rows, tvmb200_stream_t stream);
/*!
 * \brief merge_attn_output_inplace(o_self, lse_self, o_cross, lse_cross)
 *  (`vm.builtin.attention_kv_cache_merge_attn_output_inplace`, kv_state.cc:109-115; MergeAttnOutputInplace,
 *  paged_kv_cache.cc:1553-1559 = f_merge_inplace[1]): (o_self, lse_self) <- merge with (o_cross, lse_cross).
 *  o_*: [n, num_heads, head_dim], lse_*: [n, num_heads] f32.
 */
TVMB200_API int tvmb200_cache_merge_attn_output_inplace(tvmb200_cache_t c, void* o_self, float* lse_self,
                                                        const void* o_cross, const float* lse_cross, int64_t n,
                                                        int64_t num_heads, int64_t head_dim, tvmb200_stream_t stream);
/*! \brief k_out, v_out: [num_layers, end-start, Hkv, D] device tensors (DebugGetKV, paged_kv_cache.cc:1690-1725). */
TVMB200_API int tvmb200_cache_debug_get_kv(tvmb200_cache_t c, int64_t seq_id, int64_t start_pos, int64_t end_pos,
                                           void* k_out, void* v_out, tvmb200_stream_t stream);

/*!
 * \brief Register the cache under the Relax VM's global function names (`vm.builtin.paged_attention_kv_cache_create`,
 *  `vm.builtin.kv_state_*`, `vm.builtin.attention_kv_cache_*`: src/runtime/vm/kv_state.cc:33-116,
 *  paged_kv_cache.cc:2535-2639) in the tvm-ffi global table of this process, so that a compiled Relax / MLC model,
 *  which calls them by name, runs on tvm_b200's cache and kernels unmodified.  allow_override != 0 replaces the
 *  reference's own registrations when libtvm_runtime is loaded too.  Returns the number of names registered, < 0 on
 *  error (libtvm_ffi.so not loaded, or a name exists and allow_override == 0).
 */
TVMB200_API int tvmb200_register_vm_builtins(int allow_override);

/* ---- introspection (parity tests, integration glue) ---- */
/*! \brief out6 = {num_qo_heads, num_kv_heads, head_dim, dtype, num_layers, layer_id_begin_offset}. */
/*!
 * \brief Disaggregated prefill -> decode (enable_kv_transfer of the reference's create call, paged_kv_cache.cc:376-405;
 *  DisaggPrepareRecv / DisaggMarkSend :1220-1301; the transfers in AttentionWithFusedQKV :1374-1394).  The reference puts
 *  the page pools on the NVSHMEM symmetric heap and addresses a receiver by its PE number; here a receiver is a
 *  peer-mapped device pointer to its page pool of the layer (tvmb200_cache_pages on the receiving side, exported over
 *  CUDA IPC / symmetric memory / peer access), registered per PE.  `local_tp_rank` and `remote_num_kv_heads` drive the
 *  gather / scatter head mapping of kv_transfer.cu:54-66 when sender and receiver shard the KV heads differently.
 *  Transfers run on a private stream behind the step's rotary / append and are joined back into the caller's stream at
 *  the last layer of the step.
 */
TVMB200_API int tvmb200_cache_enable_kv_transfer(tvmb200_cache_t c, int32_t local_tp_rank, int32_t num_pe,
                                                 int32_t remote_num_kv_heads);
TVMB200_API int tvmb200_cache_set_remote_pages(tvmb200_cache_t c, int32_t pe, int64_t local_layer, void* peer_mapped_pages);
/*! \brief vm.builtin.kv_cache_disagg_prepare_recv: reserves `append_length` slots for the sequence (a BeginForward) and
 *  returns them run-length compressed, [n, begin_1, length_1, ..., begin_n, length_n]; *out_len = entries needed. */
TVMB200_API int tvmb200_cache_disagg_prepare_recv(tvmb200_cache_t c, int64_t seq_id, int64_t append_length, int64_t* out,
                                                  int64_t capacity, int64_t* out_len);
/*! \brief vm.builtin.kv_cache_disagg_mark_send: tokens from `begin` on go to the slots of the (compressed) map on the
 *  receiving TP group that starts at PE `recver_pe_offset`. */
TVMB200_API int tvmb200_cache_disagg_mark_send(tvmb200_cache_t c, int64_t seq_id, int64_t begin,
                                               const int64_t* compressed_remote_position_map, int64_t n,
                                               int32_t recver_pe_offset);

TVMB200_API int tvmb200_cache_shape(tvmb200_cache_t c, int64_t* out6);
/*! \brief The cache's own kernel-set context (borrowed; valid while the cache lives).  It starts as a copy of the
 *  settings current at tvmb200_cache_create; tvmb200_context_enter(ctx) + tvmb200_set_rope_scaling(...) changes
 *  them for this cache only. */
TVMB200_API int tvmb200_cache_context(tvmb200_cache_t c, tvmb200_context_t* out);
/*! \brief device pointer of pages_[local_layer]: [num_total_pages, 2, Hkv, page, D]. */
TVMB200_API int tvmb200_cache_pages(tvmb200_cache_t c, int64_t local_layer, void** dev_ptr, int64_t* num_total_pages);
/*! \brief Start (1) / stop (0) recording the callback sequence with every int32 array and scalar argument. */
TVMB200_API int tvmb200_cache_set_trace(tvmb200_cache_t c, int32_t on);
/*! \brief The recorded trace as a JSON array (valid until the next call on this cache); clears the recording. */
TVMB200_API int tvmb200_cache_take_trace(tvmb200_cache_t c, const char** json);

#ifdef __cplusplus
}
#endif
#endif /* TVM_B200_CACHE_H_ */
