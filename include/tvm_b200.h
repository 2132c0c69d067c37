/*
 * tvm_b200.h -- C ABI of the B200-native PagedKVCache attention kernel set.
 *
 * Every entry point replaces one callback that the reference's C++ cache object
 * (PagedAttentionKVCacheObj, /root/reference/src/runtime/vm/paged_kv_cache.cc) receives
 * through `vm.builtin.paged_attention_kv_cache_create` (paged_kv_cache.cc:2535-2639).
 * The reference implements the callbacks as TIR PrimFuncs
 * (python/tvm/relax/frontend/nn/llm/_{page,decode,prefill}_kernels.py, tree_attn.py,
 * position_embedding.py); here they are hand-written sm_100a CUDA.
 *
 * Conventions
 *  - plain pointers and sizes only; all tensor pointers are DEVICE pointers that already
 *    include the DLTensor byte_offset; tensors are compact row-major.
 *  - `dtype` is the element type of q/k/v/o/pages: TVMB200_F16 or TVMB200_BF16.
 *  - every call is asynchronous on `stream` (a cudaStream_t), never synchronises the device.
 *  - return value: 0 on success, non-zero on error; the message is in tvmb200_last_error()
 *    (thread local).  Nothing is launched when an error is returned.
 *  - there is NO CPU fallback: without a CUDA device every call fails.
 *
 * The same library also exports the tvm-ffi packed-function symbols (`__tvm_ffi_<name>`,
 * TVMFFISafeCallType) with the reference's positional signatures; see INTEGRATION.md.
 */
#ifndef TVM_B200_H_
#define TVM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVMB200_F16 0
#define TVMB200_BF16 1

/* mask kinds of tvmb200_attention_prefill* (host side picks them from the reference's arguments) */
#define TVMB200_MASK_NONE 0   /* col < kv_len                                             */
#define TVMB200_MASK_CAUSAL 1 /* col < kv_len - qo_len + row + 1  (_kernel_common.py:130) */

typedef void* tvmb200_stream_t;

#if defined(__GNUC__)
#define TVMB200_API __attribute__((visibility("default")))
#else
#define TVMB200_API
#endif

/*! \brief last error message of the calling thread ("" if none). */
TVMB200_API const char* tvmb200_last_error(void);
/*! \brief library version string. */
TVMB200_API const char* tvmb200_version(void);
/*! \brief number of kernels launched by this library in this process (all threads). */
TVMB200_API int64_t tvmb200_launch_count(void);

/*!
 * \brief Kernel-set context = what the reference compiles INTO one set of PrimFuncs (the `rope_scaling` dict, the
 *  rotary_dim / theta / scale of fused_rope, `layer_sliding_window_size`; kv_cache.py:690-736) plus the device scratch
 *  the launches need (split-KV partials, the work counter of the persistent prefill kernel, the block ticket of the
 *  peer gather), keyed by (device, stream).  The reference builds one kernel set per model; here a process holds
 *    - one DEFAULT context, used by every entry point below and by the `__tvm_ffi_*` module symbols,
 *    - one context per host cache (tvm_b200_cache.h), created from the settings current at its creation,
 *    - any number made with tvmb200_context_create (tvm-ffi closures bound to one: `bind_context`).
 *  tvmb200_context_enter(c) makes `c` the calling thread's current context -- the setters below and every launch on
 *  this thread then use it -- and returns the previous one (NULL = default); pass that back to leave the scope.
 *  Launches of different contexts, or of one context on different streams, never share scratch, so they may run
 *  concurrently on one device.  A new context copies the settings of the creator's current one.
 */
typedef struct tvmb200_context_s* tvmb200_context_t;
TVMB200_API int tvmb200_context_create(tvmb200_context_t* out);
TVMB200_API void tvmb200_context_retain(tvmb200_context_t c);
/*! \brief drops one reference; the last one frees the context's device scratch. */
TVMB200_API void tvmb200_context_release(tvmb200_context_t c);
TVMB200_API tvmb200_context_t tvmb200_context_enter(tvmb200_context_t c);

/*!
 * \brief Pre-size the split-KV workspace (partial O / LSE) of the current context for launches on (device, stream);
 *        tvmb200_reserve_workspace = the NULL stream.  Optional: the workspace starts at 32 MiB, which covers every
 *        decode shape up to batch 256 x 64 heads, and grows on demand, but growing calls cudaMalloc, which must not
 *        happen inside CUDA-graph capture.
 */
TVMB200_API int tvmb200_reserve_workspace(int device_id, int64_t bytes);
TVMB200_API int tvmb200_reserve_workspace_stream(int device_id, int64_t bytes, tvmb200_stream_t stream);

/*!
 * \brief Per-layer sliding window size compiled into the reference's `*_sliding_window` prefill
 *        flavour (kv_cache.py:494,703 `layer_sliding_window_size`, default 1024).  Only used by the
 *        tvm-ffi symbol `batch_prefill_paged_kv_sliding_window`; the C entry point takes it explicitly.
 */
TVMB200_API void tvmb200_set_layer_sliding_window_size(int32_t size);

/*!
 * \brief RoPE frequency scaling the reference compiles into every PrimFunc that rotates (`rope_scaling` dict ->
 *  switch_rope_freq_func, position_embedding.py:257-299: fused_rope and the inline-RoPE paths of the attention kernels,
 *  _kernel_common.py:115-127).  A loaded library needs it as state of the calling thread's current context (the
 *  default context unless tvmb200_context_enter was used), for every call made afterwards with that context:
 *    TVMB200_ROPE_SCALING_NONE   rope_freq_default
 *    TVMB200_ROPE_SCALING_LLAMA3 rope_freq_llama3 (position_embedding.py:130-160; Llama-3.1: factor 8, low 1, high 4,
 *                                original_max_position_embeddings 8192) -- tvmb200_split_rotary[_append], rotary_mode = 1
 *                                of the attention entries and the fused decode step
 *    TVMB200_ROPE_SCALING_GPTJ   rope_freq_gptj + interleaved pairs (position_embedding.py:70-76, :509-514)
 *    TVMB200_ROPE_SCALING_LLAMA4 rope_freq_llama4 (position_embedding.py:79-127; high == low selects its threshold branch)
 *    TVMB200_ROPE_SCALING_YARN   rope_freq_yarn (position_embedding.py:192-254), set by tvmb200_set_rope_scaling_yarn;
 *                                inv_theta_log_scale <= 0 means 1 / (2 ln rope_theta) (kv_cache.py:355-366)
 *  gptj / llama4 / yarn are implemented by tvmb200_split_rotary[_append] only (rope mode "normal", where the cache
 *  holds rotated K): while one of them is set, rotary_mode = 1 and the fused decode step with apply_rope > 0 are
 *  rejected, and the host cache runs decode steps as split_rotary + append + decode.  longrope (`rope_ext_factors`) is
 *  rejected: the reference's own longrope PrimFunc does not build at this commit, so it cannot be pinned.
 */
#define TVMB200_ROPE_SCALING_NONE 0
#define TVMB200_ROPE_SCALING_LLAMA3 1
#define TVMB200_ROPE_SCALING_GPTJ 2
#define TVMB200_ROPE_SCALING_LLAMA4 3
#define TVMB200_ROPE_SCALING_YARN 5
TVMB200_API int tvmb200_set_rope_scaling(int32_t kind, float factor, float low_freq_factor, float high_freq_factor,
                                         float original_max_position_embeddings);
TVMB200_API int tvmb200_set_rope_scaling_yarn(float factor, float original_max_position_embeddings, float beta_fast,
                                              float beta_slow, float inv_theta_log_scale);
/*! \brief The kind set by the two functions above. */
TVMB200_API int32_t tvmb200_get_rope_scaling_kind(void);

/*!
 * \brief Prefill implementation selector (test / profiling hook): 0 = auto (tcgen05 path for eligible shapes
 *        with >= 2048 folded rows, generic mma.sync path otherwise), 1 = force generic, 2 = force tcgen05
 *        wherever it is eligible (head_dim 128, GQA group 1 / 2 / 4 / 8 / 16; every mask; inline RoPE and sliding
 *        windows through the pre-pass below).
 */
TVMB200_API void tvmb200_set_prefill_impl(int impl);

/*!
 * \brief Prefill launches so far by path (test hook: "no silent fallback"): out[0] the mma.sync kernel, out[1] the tcgen05
 *        kernel, out[2] the tcgen05 kernel behind the gather / rotate pre-pass, out[3] the launches of out[1] that cut every
 *        item's KV range into parts merged by a second kernel (few long items, e.g. a token tree against a 32K context).
 *        The pre-pass serves rotary_mode = 1
 *        (_kernel_common.py:115-127) and the `_sliding_window` flavours (_kernel_common.py:147-170): it writes rotated
 *        q and position-ordered, rotated K / V into the context's scratch (tvmb200_reserve_workspace sizes it ahead of
 *        time: 2 * nnz_pages * 16 * Hkv * D * 2 + n * Hq * D * 2 bytes), then the attention runs as a ragged tcgen05 launch.
 */
TVMB200_API void tvmb200_debug_prefill_path_counts(int64_t out[4]);
/*! \brief Pre-pass scratch above this many bytes keeps such a call on the mma.sync kernel (default 2 GiB). */
TVMB200_API void tvmb200_set_prefill_prepass_cap(int64_t bytes);

/*!
 * \brief f_transpose_append  (ctor arg 13; _page_kernels.py:40-74; called paged_kv_cache.cc:1371,1399)
 *  pages[pos/page_size, 0|1, h, pos%page_size, :] = k|v[t, h, :]  for pos = position_map[t] != -1.
 *  pages: [num_pages, 2, num_kv_heads, page_size, head_dim]; k, v: [ntoken, num_kv_heads, head_dim].
 */
TVMB200_API int tvmb200_transpose_append(void* pages, const void* k, const void* v, const int32_t* position_map,
                             int64_t ntoken, int64_t num_pages, int32_t num_kv_heads,
                             int32_t page_size, int32_t head_dim, int dtype, tvmb200_stream_t stream);

/*!
 * \brief f_debug_get_kv  (ctor arg 26; _page_kernels.py:106-136; called paged_kv_cache.cc:1718)
 *  k_out|v_out[layer_id, p, h, :] = pages[pos/page_size, 0|1, h, pos%page_size, :], pos = position_map[p].
 *  k_out, v_out: [num_layers, seqlen, num_kv_heads, head_dim].
 */
TVMB200_API int tvmb200_debug_get_kv(const void* pages, const int32_t* position_map, void* k_out, void* v_out,
                         int64_t layer_id, int64_t num_layers, int64_t seqlen, int64_t num_pages,
                         int32_t num_kv_heads, int32_t page_size, int32_t head_dim, int dtype,
                         tvmb200_stream_t stream);

/*!
 * \brief f_copy_single_page  (ctor arg 25; _page_kernels.py:169-189; called paged_kv_cache.cc:728)
 *  pages[tgt, :, :, 0:copy_length, :] = pages[src, :, :, 0:copy_length, :].
 */
TVMB200_API int tvmb200_copy_single_page(void* pages, int64_t src_page_id, int64_t tgt_page_id,
                             int64_t copy_length, int64_t num_pages, int32_t num_kv_heads,
                             int32_t page_size, int32_t head_dim, int dtype, tvmb200_stream_t stream);

/*!
 * \brief f_compact_copy  (ctor arg 27; _page_kernels.py:235-263; called paged_kv_cache.cc:759)
 *  for each sequence b, for i in [indptr[b], indptr[b+1]) IN ORDER: slot dst[i] <- slot src[i]
 *  (slot = page_id*page_size + offset).  src_dst_pos: [2, total_copy_length].
 */
TVMB200_API int tvmb200_compact_kv_copy(void* pages, const int32_t* copy_length_indptr,
                            const int32_t* copy_src_dst_pos, int32_t batch_size,
                            int32_t total_copy_length, int64_t num_pages, int32_t num_kv_heads,
                            int32_t page_size, int32_t head_dim, int dtype, tvmb200_stream_t stream);

/*!
 * \brief f_split_rotary  (ctor arg 24; position_embedding.py:444-565; called paged_kv_cache.cc:1360)
 *  Split fused qkv [n, Hq+2Hkv, D] into q [n,Hq,D], k,v [n,Hkv,D]; when apply_rope > 0 rotate q and k
 *  (neox half-split) at position_map[t]: freq = pos*rope_scale / rope_theta^((2d mod rd)/rd).
 *  rotary_dim <= 0 means head_dim.
 */
TVMB200_API int tvmb200_split_rotary(const void* qkv, const int32_t* position_map, void* q, void* k, void* v,
                         int64_t ntoken, int32_t num_qo_heads, int32_t num_kv_heads,
                         int32_t head_dim, int32_t rotary_dim, int64_t apply_rope, float rope_scale,
                         float rope_theta, int dtype, tvmb200_stream_t stream);

/*!
 * \brief f_attention_decode for a KV-head shard of one multi-GPU box, fused with the re-assembly of the per-head
 *  outputs.  The reference shards this path with Disco tensor parallelism and gathers the heads with
 *  `runtime.disco.allgather` -> ncclAllGather behind the kernel (src/runtime/extra/disco/nccl/nccl.cc:136-144;
 *  sharded attention: tests/python/disco/test_ccl.py:557-700).  Here the kernel that produces O stores this rank's
 *  `num_qo_heads` heads into EVERY rank's gathered buffer `peer_outputs[i]` ([batch, world*num_qo_heads, D], heads
 *  [rank*num_qo_heads, (rank+1)*num_qo_heads)) through NVLink peer pointers, then writes `epoch` to
 *  `peer_flags[i][rank]` with release semantics at system scope.  `output` / `lse` still receive the local result.
 *  The same launch then waits until `peer_flags[rank][r]` has reached `epoch` for every r (16-byte peer stores, one
 *  fence + ticket per block, raise-then-wait in the last block: no cycle), so when the call's stream work completes the
 *  gathered buffer of THIS rank is complete and stream-ordered consumers may read it; tvmb200_wait_peer_flags remains for
 *  consumers on other streams.
 *  Epochs must increase by one per call (wrap-around safe); pointers are peer-mapped device pointers of one process
 *  per GPU (e.g. torch.distributed._symmetric_memory buffer_ptrs).  world == 1 degenerates to a local copy.
 */
TVMB200_API int tvmb200_attention_decode_gather(const void* q, const void* pages, const int32_t* page_indptr,
                                    const int32_t* page_values, const int32_t* length_info,
                                    const int32_t* k_rope_pos_offset, const int32_t* q_rope_position,
                                    void* output, float* lse, int32_t batch_size, int32_t nnz_pages,
                                    int64_t num_pages, int32_t num_qo_heads, int32_t num_kv_heads,
                                    int32_t page_size, int32_t head_dim, int sliding_window, int rotary_mode,
                                    float rope_scale, float rope_theta, float sm_scale, int dtype,
                                    void* const* peer_outputs, uint32_t* const* peer_flags, int32_t world,
                                    int32_t rank, uint32_t epoch, tvmb200_stream_t stream);

/*! \brief tvmb200_attention_decode_fused_qkv and tvmb200_attention_decode_gather in one: the whole decode step of a KV-head
 *  shard (rotary + append + decode) with the head re-assembly over peer memory, two launches in total. */
TVMB200_API int tvmb200_attention_decode_fused_qkv_gather(
    const void* qkv, const int32_t* q_rope_position, const int32_t* append_position_map, void* pages,
    const int32_t* page_indptr, const int32_t* page_values, const int32_t* length_info, const int32_t* k_rope_pos_offset,
    void* output, float* lse, int32_t batch_size, int32_t nnz_pages, int64_t num_pages, int32_t num_qo_heads,
    int32_t num_kv_heads, int32_t page_size, int32_t head_dim, int sliding_window, int64_t apply_rope, float rope_scale,
    float rope_theta, float sm_scale, int dtype, void* const* peer_outputs, uint32_t* const* peer_flags, int32_t world,
    int32_t rank, uint32_t epoch, tvmb200_stream_t stream);

/*!
 * \brief nvshmem.KVTransfer (src/runtime/extra/contrib/nvshmem/kv_transfer.cu:38-83, :139-257): push the k / v rows
 *  [ntokens, local_num_kv_heads, head_dim] of freshly computed tokens into the page pool of the receiving cache(s),
 *  slot remote_position_map[t] (-1 = skip), TP group starting at PE remote_tp_group_pe_offset[t].  The reference
 *  addresses receivers through the NVSHMEM symmetric heap; here `remote_pages[pe]` is the peer-mapped device pointer of
 *  PE pe's pool [*, 2, remote_num_kv_heads, page_size, head_dim] (NVLink stores, system-scope fence before the kernel
 *  ends = nvshmem_quiet).  Head mapping between differently sharded sender / receiver: kv_transfer.cu:54-66.
 */
TVMB200_API int tvmb200_kv_transfer(void* const* remote_pages, const void* k, const void* v,
                                    const int32_t* remote_position_map, const int32_t* remote_tp_group_pe_offset,
                                    int64_t ntokens, int32_t local_num_kv_heads, int32_t remote_num_kv_heads,
                                    int32_t page_size, int32_t head_dim, int32_t local_tp_rank, int32_t num_pe, int dtype,
                                    tvmb200_stream_t stream);
/*! \brief nvshmem.KVTransferPageToPage (kv_transfer.cu:84-130, :259-325): the same for rows already in the local pool
 *  (slot local_position_map[t] of `local_pages`). */
TVMB200_API int tvmb200_kv_transfer_page_to_page(void* const* remote_pages, const void* local_pages,
                                                 const int32_t* remote_position_map, const int32_t* local_position_map,
                                                 const int32_t* remote_tp_group_pe_offset, int64_t ntokens,
                                                 int32_t local_num_kv_heads, int32_t remote_num_kv_heads,
                                                 int32_t page_size, int32_t head_dim, int32_t local_tp_rank,
                                                 int32_t num_pe, int dtype, tvmb200_stream_t stream);
/*! \brief cudaDeviceEnablePeerAccess(device -> peer) for single-process multi-GPU hosts and tests. */
TVMB200_API int tvmb200_enable_peer_access(int32_t device, int32_t peer);

/*! \brief Block the stream until flags[r] has reached `epoch` for every r < world (see tvmb200_attention_decode_gather). */
TVMB200_API int tvmb200_wait_peer_flags(const uint32_t* flags, int32_t world, uint32_t epoch, tvmb200_stream_t stream);

/*!
 * \brief The whole decode step of AttentionWithFusedQKV in ONE launch (+ the split-KV merge): f_split_rotary,
 *  f_transpose_append and f_attention_decode (paged_kv_cache.cc:1360, :1371, decode at :2261-2290) for a batch in which
 *  every sequence appends exactly one token.  q and the new k are read from the fused
 *  qkv [batch, Hq+2Hkv, D] and rotated at q_rope_position[b] when apply_rope > 0 (RoPE mode "normal"; the cached K is
 *  stored rotated) exactly like f_split_rotary; the new k / v are written to slot append_position_map[b] of `pages` --
 *  which must be the last slot of sequence b, as BeginForward lays it out -- by the work item that owns that page, which
 *  also patches them into its staged copy, so the attention sees them.  The q / k / v temporaries are never
 *  materialised.  Results agree with the three-call sequence within the fp tolerance (the rotation is the same
 *  arithmetic; the pages end up bit-identical).  head_dim 128 only.
 */
TVMB200_API int tvmb200_attention_decode_fused_qkv(const void* qkv, const int32_t* q_rope_position,
                                       const int32_t* append_position_map, void* pages,
                                       const int32_t* page_indptr, const int32_t* page_values,
                                       const int32_t* length_info, const int32_t* k_rope_pos_offset,
                                       void* output, float* lse, int32_t batch_size, int32_t nnz_pages,
                                       int64_t num_pages, int32_t num_qo_heads, int32_t num_kv_heads,
                                       int32_t page_size, int32_t head_dim, int sliding_window,
                                       int64_t apply_rope, float rope_scale, float rope_theta, float sm_scale,
                                       int dtype, tvmb200_stream_t stream);

/*!
 * \brief f_split_rotary immediately followed by f_transpose_append, in ONE launch (SURVEY 8(f) "next": fused
 *  rotary + append).  Replaces the back-to-back callback pair of AttentionWithFusedQKV when the append precedes
 *  the attention (paged_kv_cache.cc:1360 then :1371): q, k, v are written exactly as tvmb200_split_rotary writes
 *  them and the same k / v registers are scattered to slot append_position_map[t] (page = slot / page_size) of
 *  `pages` when it is >= 0, so pages, q, k, v are bit-identical to the two-call sequence.
 */
TVMB200_API int tvmb200_split_rotary_append(const void* qkv, const int32_t* q_rope_position_map,
                                const int32_t* append_position_map, void* q, void* k, void* v, void* pages,
                                int64_t ntoken, int64_t num_pages, int32_t num_qo_heads, int32_t num_kv_heads,
                                int32_t page_size, int32_t head_dim, int32_t rotary_dim, int64_t apply_rope,
                                float rope_scale, float rope_theta, int dtype, tvmb200_stream_t stream);

/*!
 * \brief f_merge_inplace  (ctor arg 23; _decode_kernels.py:414-526; called paged_kv_cache.cc:2292,2329,1557)
 *  (V,S) <- merge((V,S),(V',S')) with base-2 LSE.  v, v_other: [N,H,D]; s, s_other: [N,H] float32.
 */
TVMB200_API int tvmb200_merge_state_inplace(void* v, float* s, const void* v_other, const float* s_other,
                                int64_t n, int32_t num_heads, int32_t head_dim, int dtype,
                                tvmb200_stream_t stream);

/*!
 * \brief f_attention_decode (+ sliding-window flavour)  (ctor args 17, 19; _decode_kernels.py:49-411;
 *        adapter attn_backend.h:507-515)
 *  One query token per sequence against that sequence's paged KV.  Split-KV over CTAs with an
 *  in-kernel LSE merge pass.  length_info is [B] (sliding_window == 0) or [3,B] (sliding_window != 0).
 *  lse is base-2; empty KV gives O = 0, lse = -5e4.
 */
TVMB200_API int tvmb200_attention_decode(const void* q, const void* pages, const int32_t* page_indptr,
                             const int32_t* page_values, const int32_t* length_info,
                             const int32_t* k_rope_pos_offset, const int32_t* q_rope_position,
                             void* output, float* lse, int32_t batch_size, int32_t nnz_pages,
                             int64_t num_pages, int32_t num_qo_heads, int32_t num_kv_heads,
                             int32_t page_size, int32_t head_dim, int sliding_window,
                             int rotary_mode, float rope_scale, float rope_theta, float sm_scale,
                             int dtype, tvmb200_stream_t stream);

/*!
 * \brief f_attention_prefill (+ sliding-window flavour)  (ctor args 16, 18; _prefill_kernels.py:54-391;
 *        adapter attn_backend.h:234-243).  Ragged Q vs paged KV.
 *  layer_sliding_window_size is the reference's compile-time constant; it only matters when
 *  sliding_window != 0 and causal > 0 (_kernel_common.py:138-144).
 */
TVMB200_API int tvmb200_attention_prefill_paged(const void* q, const int32_t* q_indptr, const void* pages,
                                    const int32_t* page_indptr, const int32_t* page_values,
                                    const int32_t* length_info, const int32_t* k_rope_pos_offset,
                                    const int32_t* q_rope_position, void* output, float* lse,
                                    int32_t batch_size, int32_t total_q_len, int32_t nnz_pages,
                                    int64_t num_pages, int32_t num_qo_heads, int32_t num_kv_heads,
                                    int32_t page_size, int32_t head_dim, int sliding_window,
                                    int32_t layer_sliding_window_size, int causal, int rotary_mode,
                                    float rope_scale, float rope_theta, float sm_scale, int dtype,
                                    tvmb200_stream_t stream);

/*!
 * \brief f_attention_prefill_ragged  (ctor arg 15; _prefill_kernels.py:677-923; adapter
 *        attn_backend.h:388-396; called paged_kv_cache.cc:2187).  Ragged Q vs ragged K,V.
 */
TVMB200_API int tvmb200_attention_prefill_ragged(const void* q, const int32_t* q_indptr, const void* k,
                                     const void* v, const int32_t* kv_indptr,
                                     const int32_t* q_rope_position,
                                     const int32_t* k_rope_pos_offset, void* output, float* lse,
                                     int32_t batch_size, int32_t total_q_len, int32_t total_kv_len,
                                     int32_t num_qo_heads, int32_t num_kv_heads, int32_t head_dim,
                                     int causal, int rotary_mode, float rope_scale,
                                     float rope_theta, float sm_scale, int dtype,
                                     tvmb200_stream_t stream);

/*!
 * \brief f_attention_prefill_with_tree_mask  (ctor arg 21; tree_attn.py:68-603; adapter
 *        attn_backend.h:665-673).  Ragged self-attention under a token-tree mask.
 *  mask: [tree_size, 2] int32 rows (dfs_order, subtree_end); mn_indptr: [B+1].
 */
TVMB200_API int tvmb200_attention_prefill_tree_ragged(const void* q, const int32_t* q_indptr, const void* k,
                                          const void* v, const int32_t* kv_indptr,
                                          const int32_t* q_rope_position, const int32_t* mn_indptr,
                                          const int32_t* mask, void* output, float* lse,
                                          int32_t batch_size, int32_t total_q_len,
                                          int32_t total_kv_len, int32_t num_qo_heads,
                                          int32_t num_kv_heads, int32_t head_dim, int rotary_mode,
                                          float rope_scale, float rope_theta, float sm_scale,
                                          int dtype, tvmb200_stream_t stream);

/*!
 * \brief f_attention_prefill_with_tree_mask_paged_kv  (ctor arg 20; tree_attn.py:606-1259; adapter
 *        attn_backend.h:618-627).  Ragged Q vs paged KV, tree mask on the trailing tree_len columns.
 *  rotary_mode must be 0 (the reference asserts it, tree_attn.py:699,930).
 */
TVMB200_API int tvmb200_attention_prefill_tree_paged(const void* q, const int32_t* q_indptr, const void* pages,
                                         const int32_t* page_indptr, const int32_t* page_values,
                                         const int32_t* length_info,
                                         const int32_t* k_rope_pos_offset,
                                         const int32_t* q_rope_position, void* output, float* lse,
                                         int32_t batch_size, int32_t total_q_len, int32_t nnz_pages,
                                         int64_t num_pages, int32_t num_qo_heads,
                                         int32_t num_kv_heads, int32_t page_size, int32_t head_dim,
                                         int rotary_mode, float rope_scale, float rope_theta,
                                         float sm_scale, const int32_t* tree_order_indptr,
                                         const int32_t* tree_order, int dtype,
                                         tvmb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TVM_B200_H_ */
