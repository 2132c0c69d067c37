// Host side of the paged KV cache: sequence / block / page bookkeeping, per-step auxiliary arrays, kernel
// orchestration.  C ABI in include/tvm_b200_cache.h.
//
// It restates, with its own structure, the algorithms of the reference's PagedAttentionKVCacheObj so that every
// int32 array a kernel receives is bit-identical to the reference's on the same call sequence:
//   block tree / fork / popn / remove      src/runtime/vm/paged_kv_cache.cc:574-862, attn_utils.h:101-230
//   BeginForward aux-array construction    paged_kv_cache.cc:884-1212, attn_utils.h:242-332
//   token-tree mask (DFS intervals)        paged_kv_cache.cc:1800-1928
//   sliding window / page reservation      paged_kv_cache.cc:1935-2044
//   tree commit (compaction lists)         paged_kv_cache.cc:1566-1659
//   merged aux copy + 16-byte aligned views paged_kv_cache.cc:2373-2525, attn_utils.h:817-1052
//   per-layer kernel sequence              paged_kv_cache.cc:1303-1402, 2161-2299
// B200 specifics: one pinned staging buffer and ONE cudaMemcpyAsync per step on a private copy stream, ordered
// against the caller's compute stream with events; kernels are the sm_100a C-ABI entry points of tvm_b200.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <functional>
#include <memory>
#include <numeric>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges cost nothing unless a profiler injects itself
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/tvm_b200_cache.h"

namespace tvmb200 {
int set_error(const char* fmt, ...);

namespace host {

constexpr int kMaxBlockDepth = 2;     // attn_utils.h:50
constexpr int kMaxTreeSize = 256;     // attn_utils.h:52
constexpr int32_t kTempPageId = -1;   // attn_utils.h:58

[[noreturn]] static void fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw std::runtime_error(buf);
}
#define HCHECK(cond, ...) \
  do {                    \
    if (!(cond)) fail(__VA_ARGS__); \
  } while (0)
#define HCUDA(expr)                                                                         \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) fail("CUDA error %s (%s)", cudaGetErrorString(e__), #expr);     \
  } while (0)

struct Block {
  std::vector<int32_t> page_ids;
  int32_t seq_length = 0;
  int32_t start_pos = 0;
  int32_t sink_length = 0;
  int32_t sliding_window_offset = 0;
  int32_t parent_idx = -1;
  int external_ref_cnt = 0;
  void Reset() { *this = Block(); }
};

struct Sequence {
  int32_t last_block_idx = -1;
  int32_t seq_length = 0;
  int sliding_window_size = -1;
  int last_block_attn_sink_size = 0;
  bool is_chain = true;
  std::vector<int32_t> tree_parent;
  std::vector<int32_t> tree_depth;
  bool committed = true;
  // disaggregation (attn_utils.h:154-159 KVTransferMetadata): tokens from `kv_send_start` on are pushed to the receiver
  int64_t kv_send_start = INT64_MAX;
  std::vector<int32_t> kv_remote_pos;   // receiver slot of token kv_send_start + i
  int32_t kv_recver_pe_offset = -1;
  std::vector<int32_t> kv_local_pos;    // slots of the tokens already cached here that still have to be sent
};

using IVec = std::vector<int32_t>;

// one int32 array inside the merged aux buffer
struct View {
  int64_t offset = 0;  // element offset in the merged buffer (multiple of 4 = 16 bytes)
  int64_t size = 0;
  int64_t rows = 0;    // 0 = 1-D; otherwise [rows, size / rows]
};

class Cache {
 public:
  explicit Cache(const tvmb200_cache_config& c);
  ~Cache();

  // ---- sequence management ----
  void Clear();
  void AddSequence(int64_t seq_id);
  void RemoveSequence(int64_t seq_id);
  void ForkSequence(int64_t parent, int64_t child, int64_t fork_pos);
  void PopN(int64_t seq_id, int32_t n);
  void EnableSlidingWindowForSeq(int64_t seq_id, int32_t window, int32_t sink);
  void BeginForward(const int64_t* seq_ids, const int64_t* lens, int n, const int64_t* tree, int tree_size);
  void BeginForwardImpl(const int64_t* seq_ids, const int64_t* lens, int n, const int64_t* tree, int tree_size);
  void EndForward() {}
  // disaggregation (paged_kv_cache.cc:1220-1301)
  void EnableKVTransfer(int32_t local_tp_rank, int32_t num_pe, int32_t remote_num_kv_heads);
  void SetRemotePages(int32_t pe, int64_t local_layer, void* peer_ptr);
  std::vector<int64_t> DisaggPrepareRecv(int64_t seq_id, int64_t append_length);
  void DisaggMarkSend(int64_t seq_id, int64_t begin, const int64_t* compressed_remote_position_map, int64_t n,
                      int32_t recver_pe_offset);
  void CommitAcceptedTokenTreeNodes(const int64_t* seq_ids, const int64_t* leaves, int n);
  bool Empty() const {
    return seq_map_.empty() && free_blocks_.size() == blocks_.size() &&
           free_pages_.size() == static_cast<size_t>(num_total_pages_);
  }
  int32_t NumAvailablePages() const { return static_cast<int32_t>(free_pages_.size()); }
  int32_t TotalSequenceLength() const {
    int32_t t = 0;
    for (const auto& kv : seq_map_) t += kv.second.seq_length;
    return t;
  }
  void AttentionWithFusedQKV(int64_t layer_id, double sm_scale, const void* qkv, void* o, int64_t rows,
                             cudaStream_t stream);
  void SelfAttention(int64_t layer_id, double sm_scale, const void* q, const void* k, const void* v, void* o, float* lse,
                     int64_t rows, cudaStream_t stream);
  void CrossAttention(int64_t layer_id, double sm_scale, const void* q, void* o, float* lse, int64_t rows,
                      cudaStream_t stream);
  void AttentionWithSharedKV(int64_t source_layer_id, double sm_scale, const void* q, const void* cur_k, const void* cur_v,
                             void* o, int64_t rows, cudaStream_t stream);
  void MergeAttnOutputInplace(void* o_self, float* lse_self, const void* o_cross, const float* lse_cross, int64_t n,
                              int64_t num_heads, int64_t head_dim, cudaStream_t stream);
  void DebugGetKV(int64_t seq_id, int64_t start, int64_t end, void* k_out, void* v_out, cudaStream_t stream);
  void QueryPositions(const int32_t** ptr, int64_t* n, cudaStream_t stream) {
    HCHECK(batch_valid_, "get_query_positions: call begin_forward first (the last one did not complete)");
    SyncAux(stream);
    *ptr = dev(v_q_rope_pos_);
    *n = v_q_rope_pos_.size;
  }
  void* Pages(int64_t layer, int64_t* np) const {
    HCHECK(layer >= 0 && layer < num_layers_, "layer %ld out of range", (long)layer);
    *np = num_total_pages_;
    return pages_.empty() ? nullptr : pages_[layer];
  }
  void Shape(int64_t* out) const {
    out[0] = num_qo_heads_;
    out[1] = num_kv_heads_;
    out[2] = head_dim_;
    out[3] = dtype_;
    out[4] = num_layers_;
    out[5] = layer_begin_;
  }
  tvmb200_context_t Context() const { return ctx_; }
  void SetTrace(bool on) {
    tracing_ = on;
    trace_.clear();
  }
  const char* TakeTrace() {
    trace_json_ = "[";
    for (size_t i = 0; i < trace_.size(); ++i) trace_json_ += (i ? "," : "") + trace_[i];
    trace_json_ += "]";
    trace_.clear();
    return trace_json_.c_str();
  }

 private:
  // ---- configuration ----
  int64_t page_size_, num_layers_, layer_begin_, num_qo_heads_, num_kv_heads_, head_dim_;
  int64_t num_total_pages_, prefill_chunk_, reserved_seqs_, layer_sws_;
  bool support_sw_, support_layer_sw_;
  std::vector<int32_t> attn_kinds_;
  int rope_mode_;
  double rotary_scale_, rotary_theta_;
  int dtype_, device_;
  size_t esize_ = 2;
  // this cache's kernel-set context (tvm_b200.h): the rope scaling / layer window current at creation, own scratch
  tvmb200_context_t ctx_ = nullptr;

  // ---- page / block / sequence state ----
  std::vector<int32_t> free_pages_;
  std::unordered_map<int64_t, Sequence> seq_map_;
  std::vector<Block> blocks_;
  std::vector<int32_t> free_blocks_;

  // ---- current batch ----
  bool dirty_ = false;
  bool batch_valid_ = false;  // the last BeginForward ran to completion
  int64_t cur_batch_ = 0;
  std::vector<int64_t> cur_seq_ids_, cur_lens_;
  std::vector<bool> is_chain_on_depths_ = std::vector<bool>(kMaxBlockDepth, true);
  int num_depths_ = 0;
  bool append_before_attn_ = false;
  std::vector<bool> use_decode_kernel_;
  bool is_decode_request_ = true;

  // host aux arrays (BeginForward)
  IVec qo_indptr_[kMaxBlockDepth], page_indptr_[kMaxBlockDepth], page_indices_[kMaxBlockDepth];
  IVec page_indptr_sw_[kMaxBlockDepth], page_indices_sw_[kMaxBlockDepth];
  IVec last_page_len_[kMaxBlockDepth], sw_offset_[kMaxBlockDepth], sink_size_[kMaxBlockDepth];
  IVec k_rope_off_[kMaxBlockDepth], k_rope_off_sw_[kMaxBlockDepth];
  IVec tree_mask_[kMaxBlockDepth], tree_mn_indptr_[kMaxBlockDepth];
  IVec k_ragged_rope_off_, q_rope_pos_, append_pos_, cur_len_indptr_;
  IVec commit_indptr_, commit_src_, commit_dst_;

  // merged aux buffer: host staging (pinned when on a device) + device copy, and the views into it
  IVec stage_;
  // two device copies + two pinned staging buffers, alternated per step, so that the H2D copy of step i+1 does not
  // have to wait for the kernels of step i (which still read the other buffer)
  int32_t* stage_pinned_[2] = {nullptr, nullptr};
  int32_t* aux_dev_[2] = {nullptr, nullptr};
  cudaEvent_t ev_aux_copied_[2] = {nullptr, nullptr};   // H2D from stage_pinned_[i] has completed
  cudaEvent_t ev_aux_readers_[2] = {nullptr, nullptr};  // last kernels that read aux_dev_[i] have completed
  bool aux_used_[2] = {false, false};
  int aux_cur_ = 0;
  int64_t aux_capacity_ = 0;
  int32_t* compact_pinned_ = nullptr;
  int32_t* compact_dev_ = nullptr;
  View v_q_rope_pos_, v_qo_indptr_[kMaxBlockDepth], v_page_indptr_[kMaxBlockDepth], v_page_indices_[kMaxBlockDepth];
  View v_page_indptr_sw_[kMaxBlockDepth], v_page_indices_sw_[kMaxBlockDepth], v_length_info_[kMaxBlockDepth];
  View v_length_info_sw_[kMaxBlockDepth], v_k_rope_off_[kMaxBlockDepth], v_k_rope_off_sw_[kMaxBlockDepth];
  View v_cur_len_indptr_, v_k_ragged_rope_off_, v_append_pos_, v_tree_mask_[kMaxBlockDepth], v_tree_mn_[kMaxBlockDepth];
  int64_t stage_off_ = 0;
  int64_t total_append_ = 0;

  // device memory
  std::vector<void*> pages_;
  void *tmp_q_ = nullptr, *tmp_k_ = nullptr, *tmp_v_ = nullptr, *tmp_o_ = nullptr;
  float *tmp_lse_ = nullptr, *merged_lse_ = nullptr;
  int32_t* dbg_pos_dev_ = nullptr;
  cudaStream_t copy_stream_ = nullptr;
  cudaEvent_t ev_copy_ = nullptr, ev_compute_ = nullptr, ev_attn_done_ = nullptr;

  // trace
  bool tracing_ = false;
  std::vector<std::string> trace_;
  std::string trace_json_;

  bool planning_only() const { return device_ < 0; }
  int32_t* dev(const View& v) const { return aux_dev_[aux_cur_] ? aux_dev_[aux_cur_] + v.offset : nullptr; }
  const int32_t* hostv(const View& v) const { return stage_.data() + v.offset; }

  int32_t GetFreePage() {
    HCHECK(!free_pages_.empty(), "The KV cache is full. No page can be allocated.");
    int32_t p = free_pages_.back();
    free_pages_.pop_back();
    return p;
  }
  int32_t GetFreeBlock() {
    if (!free_blocks_.empty()) {
      int32_t b = free_blocks_.back();
      free_blocks_.pop_back();
      blocks_[b].Reset();
      return b;
    }
    blocks_.emplace_back();
    return static_cast<int32_t>(blocks_.size()) - 1;
  }
  Sequence MakeSequence(int32_t last_block) {
    Sequence s;
    ++blocks_[last_block].external_ref_cnt;
    s.last_block_idx = last_block;
    for (int32_t b = last_block; b != -1; b = blocks_[b].parent_idx) s.seq_length += blocks_[b].seq_length;
    return s;
  }
  std::vector<int32_t> BlockTrace(const Sequence& s) const {
    std::vector<int32_t> t;
    for (int32_t b = s.last_block_idx; b != -1; b = blocks_[b].parent_idx) t.push_back(b);
    std::reverse(t.begin(), t.end());
    return t;
  }
  Sequence& Seq(int64_t id) {
    auto it = seq_map_.find(id);
    HCHECK(it != seq_map_.end(), "The sequence \"%ld\" cannot be found in KV cache.", (long)id);
    return it->second;
  }
  int32_t LayerSwOffset(int64_t len) const {
    return len <= layer_sws_ ? 0 : static_cast<int32_t>((len - layer_sws_) % page_size_);
  }
  int32_t LayerSwNumPages(int64_t len) const {
    if (len == 0) return 0;
    int64_t w = std::min(len, layer_sws_);
    return static_cast<int32_t>((LayerSwOffset(len) + w + page_size_ - 1) / page_size_);
  }
  void ReserveAppendLength(Sequence* seq, int64_t append_length);
  void SlideWindow(Sequence* seq);
  void ConstructTokenTreeMask(const std::vector<Sequence*>& seqs, const int64_t* tree, int tree_size,
                              const std::vector<std::vector<int32_t>>& ids_on_depths,
                              const std::vector<std::vector<int32_t>>& trailing);
  void CopySinglePage(int32_t src, int32_t tgt, int64_t len);
  void CompactKVCopy();
  // the staging buffer is sized for reserved_num_seqs sequences (like the reference's merged aux buffer,
  // attn_utils.h:817-1052); a batch beyond that is refused before anything is written past its end
  void StageRoom(int64_t n) const {
    HCHECK(stage_off_ + (n + 3) / 4 * 4 <= static_cast<int64_t>(stage_.size()),
           "auxiliary buffer overflow: the batch needs more than the %ld int32 reserved for reserved_num_seqs = %ld",
           (long)stage_.size(), (long)reserved_seqs_);
  }
  View Put(const IVec& v) {
    View r;
    StageRoom(static_cast<int64_t>(v.size()));
    r.offset = stage_off_;
    r.size = static_cast<int64_t>(v.size());
    if (!v.empty()) std::memcpy(stage_.data() + stage_off_, v.data(), v.size() * 4);
    stage_off_ += (r.size + 3) / 4 * 4;
    return r;
  }
  View Put3(const IVec& a, const IVec& b, const IVec& c) {
    View r;
    const int64_t n = static_cast<int64_t>(a.size());
    StageRoom(3 * n);
    r.offset = stage_off_;
    r.size = 3 * n;
    r.rows = 3;
    std::memcpy(stage_.data() + stage_off_, a.data(), n * 4);
    std::memcpy(stage_.data() + stage_off_ + n, b.data(), n * 4);
    std::memcpy(stage_.data() + stage_off_ + 2 * n, c.data(), n * 4);
    stage_off_ += (3 * n + 3) / 4 * 4;
    return r;
  }
  void BuildAuxViews();
  void SyncAux(cudaStream_t compute);
  void EnsureScratch(cudaStream_t compute);
  std::set<cudaStream_t> scratch_streams_;
  // ---- disaggregation ----
  bool kv_transfer_enabled_ = false, transfer_kv_ = false, page_to_page_transfer_kv_ = false;
  int32_t kv_local_tp_rank_ = 0, kv_num_pe_ = 0, kv_remote_num_kv_heads_ = 0;
  std::vector<std::vector<void*>> remote_pages_;  // [local layer][pe] peer-mapped page pools of the receivers
  IVec kv_tx_remote_pos_, kv_tx_recver_, kv_p2p_local_pos_, kv_p2p_remote_pos_, kv_p2p_recver_;
  View v_kv_tx_remote_pos_, v_kv_tx_recver_, v_kv_p2p_local_pos_, v_kv_p2p_remote_pos_, v_kv_p2p_recver_;
  cudaStream_t kv_transfer_stream_ = nullptr;
  cudaEvent_t ev_kv_ready_ = nullptr, ev_kv_sent_ = nullptr;
  bool kv_sent_pending_ = false, kv_aux_grown_ = false;
  int64_t CheckLayer(int64_t layer_id) const;
  void MarkAttentionDone(cudaStream_t st);
  void SelfAttnInternal(const void* q, const void* k, const void* v, void* o, float* lse, double sm_scale, cudaStream_t st);
  bool CrossAttnInternal(int64_t layer_id, const void* q, void* o, float* lse, double sm_scale, bool is_first, bool causal,
                         const void* fused_qkv, cudaStream_t st);
  void AttentionInternal(int64_t layer_id, const void* q, const void* k, const void* v, void* o, double sm_scale,
                         const void* fused_qkv, cudaStream_t st);

  // ---- trace helpers ----
  struct Arg {
    std::string s;
  };
  static Arg TI(const View& v, const int32_t* host) {  // int32 tensor with values
    std::ostringstream o;
    if (v.rows)
      o << "{\"t\":\"int32\",\"shape\":[" << v.rows << "," << v.size / v.rows << "],\"v\":[";
    else
      o << "{\"t\":\"int32\",\"shape\":[" << v.size << "],\"v\":[";
    for (int64_t i = 0; i < v.size; ++i) o << (i ? "," : "") << host[i];
    o << "]}";
    return {o.str()};
  }
  Arg TI(const View& v) const { return TI(v, hostv(v)); }
  Arg TF(std::initializer_list<int64_t> shape, const char* dt = nullptr) const {
    std::ostringstream o;
    o << "{\"t\":\"" << (dt ? dt : (dtype_ == TVMB200_F16 ? "float16" : "bfloat16")) << "\",\"shape\":[";
    int i = 0;
    for (int64_t s : shape) o << (i++ ? "," : "") << s;
    o << "],\"v\":null}";
    return {o.str()};
  }
  static Arg SI(int64_t v) { return {"{\"s\":" + std::to_string(v) + "}"}; }
  static Arg SF(double v) {
    char b[64];
    snprintf(b, sizeof(b), "{\"s\":%.17g}", v);
    return {b};
  }
  // the argument strings (every int32 array as text) are only built while a trace is being recorded
#define TRACE(fn, ...)                       \
  do {                                       \
    if (tracing_) Trace(fn, __VA_ARGS__);    \
  } while (0)
  void Trace(const char* fn, std::initializer_list<Arg> args) {
    if (!tracing_) return;
    std::string s = std::string("{\"fn\":\"") + fn + "\",\"args\":[";
    int i = 0;
    for (const Arg& a : args) s += (i++ ? "," : "") + a.s;
    s += "]}";
    trace_.push_back(std::move(s));
  }
  static void Rc(int rc) {
    if (rc != 0) throw std::runtime_error(tvmb200_last_error());
  }
};

// --------------------------------------------------------------------------------------------------------------------
Cache::Cache(const tvmb200_cache_config& c)
    : page_size_(c.page_size),
      num_layers_(c.num_layers),
      layer_begin_(c.layer_id_begin_offset),
      num_qo_heads_(c.num_qo_heads),
      num_kv_heads_(c.num_kv_heads),
      head_dim_(c.head_dim),
      prefill_chunk_(c.prefill_chunk_size),
      reserved_seqs_(c.reserved_num_seqs),
      layer_sws_(c.layer_sliding_window_size > 0 ? c.layer_sliding_window_size : 1024),
      rope_mode_(c.rope_mode),
      rotary_scale_(c.rotary_scale),
      rotary_theta_(c.rotary_theta),
      dtype_(c.dtype),
      device_(c.device_id) {
  HCHECK(page_size_ > 0 && num_layers_ > 0 && num_qo_heads_ > 0 && num_kv_heads_ > 0 && head_dim_ > 0, "bad cache shape");
  HCHECK(dtype_ == TVMB200_F16 || dtype_ == TVMB200_BF16, "unsupported KV dtype %d (float16 / bfloat16)", dtype_);
  const int64_t total_layers = layer_begin_ + num_layers_;
  attn_kinds_.assign(total_layers, TVMB200_ATTN_MHA);
  if (c.attn_kinds) attn_kinds_.assign(c.attn_kinds, c.attn_kinds + total_layers);
  bool any_layer_sw = false;
  for (int32_t k : attn_kinds_) {
    HCHECK(k == TVMB200_ATTN_MHA || k == TVMB200_ATTN_MHA_SLIDING,
           "attention kind %d is outside the MHA/GQA hot path (MLA / linear attention are not built)", k);
    any_layer_sw |= k == TVMB200_ATTN_MHA_SLIDING;
  }
  // paged_kv_cache.cc:335-344
  support_sw_ = any_layer_sw ? false : c.support_sliding_window != 0;
  support_layer_sw_ = any_layer_sw;
  if (c.support_sliding_window && rope_mode_ != TVMB200_ROPE_NONE) rope_mode_ = TVMB200_ROPE_INLINE;
  // paged_kv_cache.cc:2615-2619
  num_total_pages_ = (c.total_token_capacity + page_size_ - 1) / page_size_ + 1;
  if (c.support_sliding_window) num_total_pages_ += reserved_seqs_ * 2;
  Clear();
  HCHECK(tvmb200_context_create(&ctx_) == 0, "%s", tvmb200_last_error());
  {
    tvmb200_context_t prev = tvmb200_context_enter(ctx_);
    tvmb200_set_layer_sliding_window_size(static_cast<int32_t>(layer_sws_));
    tvmb200_context_enter(prev);
  }

  // worst-case size of the merged aux buffer (the reference allocates a flat 32 Mi-element buffer, attn_utils.h:825)
  auto al = [](int64_t n) { return (n + 3) / 4 * 4; };
  int64_t per_depth = 2 * al(reserved_seqs_ + 1) + 2 * al(reserved_seqs_ + 1) + 2 * al(num_total_pages_) +
                      2 * al(3 * reserved_seqs_) + 2 * al(reserved_seqs_) + al(kMaxTreeSize * 2 * reserved_seqs_) +
                      al(reserved_seqs_ + 1);
  // q_rope_position + append_position maps, plus the two prefill_chunk-sized regions BuildAuxViews skips to keep the
  // reference's byte offsets (its kv-transfer maps)
  aux_capacity_ = al(prefill_chunk_) * 4 + al(reserved_seqs_ + 1) + al(reserved_seqs_) + kMaxBlockDepth * per_depth + 64;
  stage_.assign(aux_capacity_, 0);

  if (!planning_only()) {
    HCHECK(page_size_ == 16, "the sm_100a kernels are built for 16-slot pages, got page_size %ld", (long)page_size_);
    HCUDA(cudaSetDevice(device_));
    const size_t page_bytes = static_cast<size_t>(num_total_pages_) * 2 * num_kv_heads_ * page_size_ * head_dim_ * esize_;
    for (int64_t l = 0; l < num_layers_; ++l) {
      void* p = nullptr;
      HCUDA(cudaMalloc(&p, page_bytes));
      pages_.push_back(p);
    }
    const size_t qb = static_cast<size_t>(prefill_chunk_) * num_qo_heads_ * head_dim_ * esize_;
    const size_t kb = static_cast<size_t>(prefill_chunk_) * num_kv_heads_ * head_dim_ * esize_;
    HCUDA(cudaMalloc(&tmp_q_, qb));
    HCUDA(cudaMalloc(&tmp_k_, kb));
    HCUDA(cudaMalloc(&tmp_v_, kb));
    HCUDA(cudaMalloc(&tmp_o_, qb));
    HCUDA(cudaMalloc(reinterpret_cast<void**>(&tmp_lse_), static_cast<size_t>(prefill_chunk_) * num_qo_heads_ * 4));
    HCUDA(cudaMalloc(reinterpret_cast<void**>(&merged_lse_), static_cast<size_t>(prefill_chunk_) * num_qo_heads_ * 4));
    for (int i = 0; i < 2; ++i) {
      HCUDA(cudaMalloc(reinterpret_cast<void**>(&aux_dev_[i]), aux_capacity_ * 4));
      HCUDA(cudaMallocHost(reinterpret_cast<void**>(&stage_pinned_[i]), aux_capacity_ * 4));
      HCUDA(cudaEventCreateWithFlags(&ev_aux_copied_[i], cudaEventDisableTiming));
      HCUDA(cudaEventCreateWithFlags(&ev_aux_readers_[i], cudaEventDisableTiming));
    }
    const int64_t ccap = al(reserved_seqs_ + 1) + al(2 * std::min<int64_t>(kMaxTreeSize * reserved_seqs_, prefill_chunk_)) + 16;
    HCUDA(cudaMalloc(reinterpret_cast<void**>(&compact_dev_), ccap * 4));
    HCUDA(cudaMallocHost(reinterpret_cast<void**>(&compact_pinned_), ccap * 4));
    HCUDA(cudaMalloc(reinterpret_cast<void**>(&dbg_pos_dev_), static_cast<size_t>(num_total_pages_) * page_size_ * 4));
    HCUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    HCUDA(cudaEventCreateWithFlags(&ev_copy_, cudaEventDisableTiming));
    HCUDA(cudaEventCreateWithFlags(&ev_compute_, cudaEventDisableTiming));
    HCUDA(cudaEventCreateWithFlags(&ev_attn_done_, cudaEventDisableTiming));
  }
}

Cache::~Cache() {
  if (!planning_only()) {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
  }
  tvmb200_context_release(ctx_);
  if (planning_only()) return;
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  for (void* p : pages_) cudaFree(p);
  if (kv_transfer_stream_) cudaStreamDestroy(kv_transfer_stream_);
  if (ev_kv_ready_) cudaEventDestroy(ev_kv_ready_);
  if (ev_kv_sent_) cudaEventDestroy(ev_kv_sent_);
  cudaFree(tmp_q_);
  cudaFree(tmp_k_);
  cudaFree(tmp_v_);
  cudaFree(tmp_o_);
  cudaFree(tmp_lse_);
  cudaFree(merged_lse_);
  for (int i = 0; i < 2; ++i) {
    cudaFree(aux_dev_[i]);
    cudaFreeHost(stage_pinned_[i]);
    if (ev_aux_copied_[i]) cudaEventDestroy(ev_aux_copied_[i]);
    if (ev_aux_readers_[i]) cudaEventDestroy(ev_aux_readers_[i]);
  }
  cudaFree(compact_dev_);
  cudaFree(dbg_pos_dev_);
  cudaFreeHost(compact_pinned_);
  if (copy_stream_) cudaStreamDestroy(copy_stream_);
  if (ev_copy_) cudaEventDestroy(ev_copy_);
  if (ev_compute_) cudaEventDestroy(ev_compute_);
  if (ev_attn_done_) cudaEventDestroy(ev_attn_done_);
}

// NVTX ranges named like the reference's (paged_kv_cache.cc:2374 "SyncAuxArrayToDevice"; the VM wraps every builtin in
// "RelaxVM: <name>", vm.cc:551): a timeline of a reference-driven run and of this cache line up by name.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

void Cache::Clear() {
  seq_map_.clear();
  free_pages_.clear();
  for (int64_t p = num_total_pages_ - 1; p >= 0; --p) free_pages_.push_back(static_cast<int32_t>(p));  // LIFO: page 0 first
  blocks_.clear();
  free_blocks_.clear();
  dirty_ = false;
  batch_valid_ = false;
}

void Cache::AddSequence(int64_t seq_id) {
  HCHECK(seq_map_.find(seq_id) == seq_map_.end(), "The sequence \"%ld\" is already in the KV cache.", (long)seq_id);
  int32_t b = GetFreeBlock();
  seq_map_.insert({seq_id, MakeSequence(b)});
  dirty_ = true;
}

void Cache::RemoveSequence(int64_t seq_id) {
  auto it = seq_map_.find(seq_id);
  HCHECK(it != seq_map_.end(), "The sequence \"%ld\" cannot be found in KV cache.", (long)seq_id);
  int32_t b = it->second.last_block_idx;
  HCHECK(blocks_[b].external_ref_cnt >= 1, "block reference count underflow");
  while (b != -1 && blocks_[b].external_ref_cnt == 1) {
    for (int32_t p : blocks_[b].page_ids) free_pages_.push_back(p);
    free_blocks_.push_back(b);
    b = blocks_[b].parent_idx;
  }
  if (b != -1) {
    HCHECK(blocks_[b].external_ref_cnt > 1, "block reference count underflow");
    --blocks_[b].external_ref_cnt;
  }
  seq_map_.erase(it);
  dirty_ = true;
}

void Cache::ForkSequence(int64_t parent_id, int64_t child_id, int64_t fork_pos) {
  auto pit = seq_map_.find(parent_id);
  HCHECK(pit != seq_map_.end(), "The parent sequence \"%ld\" cannot be found in KV cache.", (long)parent_id);
  HCHECK(seq_map_.find(child_id) == seq_map_.end(), "The child sequence \"%ld\" is already in the KV cache.", (long)child_id);
  HCHECK(fork_pos >= -1, "The forked position should be non-negative, or -1 for last position as default.");
  Sequence& parent = pit->second;
  HCHECK(fork_pos <= parent.seq_length, "The forked position should not exceed the total length of parent sequence.");
  HCHECK(parent.committed,
         "The parent sequence's token tree computed in the last round of forward has not been committed with accepted nodes.");
  if (fork_pos == -1) fork_pos = parent.seq_length;
  if (parent.sliding_window_size != -1) {
    const int32_t sink = parent.seq_length - blocks_[parent.last_block_idx].seq_length + parent.last_block_attn_sink_size;
    HCHECK(fork_pos <= sink,
           "The parent sequence \"%ld\" is enabled with sliding window and thus only can be forked within sink size = %d. "
           "But the forked position = %ld.", (long)parent_id, sink, (long)fork_pos);
  }
  if (fork_pos == parent.seq_length && fork_pos % page_size_ == 0 && blocks_[parent.last_block_idx].seq_length > 0) {
    // keep the parent decodable: give it a fresh empty tail block
    int32_t nb = GetFreeBlock();
    blocks_[nb].start_pos = parent.seq_length;
    blocks_[nb].parent_idx = parent.last_block_idx;
    blocks_[nb].external_ref_cnt = 1;
    parent.last_block_idx = nb;
  }
  int32_t child_block = GetFreeBlock();
  std::vector<int32_t> trace = BlockTrace(parent);
  int64_t in_block = fork_pos;
  for (int32_t fb : trace) {
    if (fb != trace.back()) {
      HCHECK(blocks_[fb].seq_length > 0, "empty interior block");
      HCHECK(blocks_[fb].seq_length % page_size_ == 0, "interior block is not page aligned");
      if (blocks_[fb].seq_length <= in_block) {
        in_block -= blocks_[fb].seq_length;
        continue;
      }
    }
    const int32_t in_page = static_cast<int32_t>(in_block % page_size_);
    const int32_t moved_offset = static_cast<int32_t>(in_block - in_page);
    const int32_t moved_pages = moved_offset / static_cast<int32_t>(page_size_);
    if (moved_pages == 0) {
      const int32_t pb = blocks_[fb].parent_idx;
      if (pb != -1) ++blocks_[pb].external_ref_cnt;
      blocks_[child_block].parent_idx = pb;
    } else {
      // split the forked block: its leading full pages become a shared parent block
      const int32_t pb = GetFreeBlock();
      blocks_[pb].parent_idx = blocks_[fb].parent_idx;
      blocks_[fb].parent_idx = pb;
      blocks_[child_block].parent_idx = pb;
      blocks_[pb].external_ref_cnt = 2;
      auto first = blocks_[fb].page_ids.begin();
      blocks_[pb].page_ids.assign(first, first + moved_pages);
      blocks_[fb].page_ids.erase(first, first + moved_pages);
      blocks_[pb].start_pos = blocks_[fb].start_pos;
      blocks_[fb].start_pos += moved_offset;
      blocks_[pb].seq_length = moved_offset;
      blocks_[fb].seq_length -= moved_offset;
      if (parent.sliding_window_size != -1 && fb == parent.last_block_idx) {
        HCHECK(moved_offset <= parent.last_block_attn_sink_size, "fork splits inside the sliding window");
        parent.last_block_attn_sink_size -= moved_offset;
      }
    }
    blocks_[child_block].start_pos = static_cast<int32_t>(fork_pos - in_page);
    blocks_[child_block].seq_length = in_page;
    if (in_page > 0) {
      const int32_t src = blocks_[fb].page_ids[0];
      const int32_t tgt = GetFreePage();
      blocks_[child_block].page_ids.push_back(tgt);
      CopySinglePage(src, tgt, in_page);
    }
    break;
  }
  seq_map_.insert({child_id, MakeSequence(child_block)});
  dirty_ = true;
}

void Cache::CopySinglePage(int32_t src, int32_t tgt, int64_t len) {
  for (int64_t l = 0; l < num_layers_; ++l)
    TRACE("copy_single_page", {TF({num_total_pages_, 2, num_kv_heads_, page_size_, head_dim_}), SI(src), SI(tgt), SI(len)});
  if (planning_only()) return;
  // runs on the copy stream, after the last attention's appends; the next SyncAux orders it before any later
  // attention (paged_kv_cache.cc:721-734)
  HCUDA(cudaStreamWaitEvent(copy_stream_, ev_attn_done_, 0));
  for (int64_t l = 0; l < num_layers_; ++l)
    Rc(tvmb200_copy_single_page(pages_[l], src, tgt, len, num_total_pages_, static_cast<int32_t>(num_kv_heads_),
                                static_cast<int32_t>(page_size_), static_cast<int32_t>(head_dim_), dtype_, copy_stream_));
}

void Cache::EnableSlidingWindowForSeq(int64_t seq_id, int32_t window, int32_t sink) {
  HCHECK(support_sw_ || support_layer_sw_, "The KV cache does not support sliding window.");
  Sequence& s = Seq(seq_id);
  HCHECK(sink >= 0, "The specified attention sink size is expected to be non negative");
  HCHECK(window > 0, "The specified sliding window size should be positive.");
  HCHECK(sink < window, "The attn sink size should be less than the sliding window size.");
  HCHECK(s.sliding_window_size == -1, "A sequence cannot be enabled twice for sliding window.");
  const int32_t prefix = s.seq_length - blocks_[s.last_block_idx].seq_length;
  HCHECK(prefix >= 0, "negative prefix length");
  s.last_block_attn_sink_size = std::max(sink - prefix, 0);
  s.sliding_window_size = window;
}

void Cache::PopN(int64_t seq_id, int32_t n) {
  auto it = seq_map_.find(seq_id);
  HCHECK(it != seq_map_.end(), "The sequence \"%ld\" cannot be found in KV cache.", (long)seq_id);
  HCHECK(n >= 0, "The length of popping %d cannot be negative.", n);
  HCHECK(n <= it->second.seq_length,
         "The sequence only has length %d, while the length of pop is %d which exceeds the whole sequence length.",
         it->second.seq_length, n);
  if (n == 0) return;
  int32_t b = it->second.last_block_idx;
  HCHECK(blocks_[b].external_ref_cnt >= 1, "block reference count underflow");
  while (b != -1 && blocks_[b].external_ref_cnt == 1) {
    if (n > blocks_[b].seq_length) {
      n -= blocks_[b].seq_length;
      it->second.seq_length -= blocks_[b].seq_length;
      for (int32_t p : blocks_[b].page_ids) free_pages_.push_back(p);
      free_blocks_.push_back(b);
      b = blocks_[b].parent_idx;
      it->second.last_block_idx = b;
      continue;
    }
    int64_t cur = static_cast<int64_t>(blocks_[b].page_ids.size());
    const int64_t tgt = (blocks_[b].seq_length - n + page_size_ - 1) / page_size_;
    while (cur > tgt) {
      free_pages_.push_back(blocks_[b].page_ids.back());
      blocks_[b].page_ids.pop_back();
      --cur;
    }
    it->second.seq_length -= n;
    blocks_[b].seq_length -= n;
    n = 0;
    break;
  }
  if (n) {
    // the rest lives in a shared block: re-create the sequence as a fork of itself at the shorter length
    const int64_t tmp_id = -1 - seq_id;
    HCHECK(seq_map_.find(tmp_id) == seq_map_.end(), "temporary sequence id collision");
    ForkSequence(seq_id, tmp_id, it->second.seq_length - n);
    RemoveSequence(seq_id);
    auto t = seq_map_.find(tmp_id);
    Sequence moved = t->second;
    seq_map_.erase(t);
    seq_map_.insert({seq_id, moved});
  }
  dirty_ = true;
}

void Cache::SlideWindow(Sequence* seq) {
  if (seq->sliding_window_size == -1 || !support_sw_) return;
  if (seq->seq_length <= seq->sliding_window_size) return;
  const int32_t slide = seq->seq_length - seq->sliding_window_size;
  Block& blk = blocks_[seq->last_block_idx];
  if (seq->last_block_attn_sink_size > 0 && blk.sink_length == 0) {
    HCHECK(blk.sliding_window_offset == 0, "unexpected sliding window offset");
    blk.sink_length = seq->last_block_attn_sink_size;
    blk.sliding_window_offset = seq->last_block_attn_sink_size;
  }
  const int32_t ps = static_cast<int32_t>(page_size_);
  const int32_t sink_pages = (blk.sink_length + ps - 1) / ps;
  int32_t page_idx = (blk.sliding_window_offset + slide) / ps;
  const int32_t page_off = (blk.sliding_window_offset + slide) % ps;
  while (page_idx > sink_pages) {
    if (blk.page_ids[sink_pages] != kTempPageId) free_pages_.push_back(blk.page_ids[sink_pages]);
    blk.page_ids.erase(blk.page_ids.begin() + sink_pages);
    --page_idx;
  }
  HCHECK(page_idx == sink_pages - 1 || page_idx == sink_pages, "sliding window bookkeeping is inconsistent");
  seq->seq_length = seq->sliding_window_size;
  blk.seq_length -= slide;
  blk.sliding_window_offset = page_idx * ps + page_off;
  HCHECK(blk.seq_length >= blk.sink_length && blk.sliding_window_offset >= blk.sink_length, "sliding window underflow");
  HCHECK((blk.sliding_window_offset + (blk.seq_length - blk.sink_length) + ps - 1) / ps ==
             static_cast<int32_t>(blk.page_ids.size()), "sliding window page count mismatch");
}

void Cache::ReserveAppendLength(Sequence* seq, int64_t append_length) {
  Block& blk = blocks_[seq->last_block_idx];
  HCHECK(append_length > 0, "Append with length 0 is not allowed.");
  HCHECK(blk.external_ref_cnt == 1, "The block is %d-time referenced by other blocks, thus cannot accept new KV values.",
         blk.external_ref_cnt - 1);
  const int64_t cur = static_cast<int64_t>(blk.page_ids.size());
  const int64_t tgt = (blk.seq_length - blk.sink_length + blk.sliding_window_offset + append_length + page_size_ - 1) / page_size_;
  for (int64_t i = cur; i < tgt; ++i) {
    if (free_pages_.empty() && seq->sliding_window_size != -1 && support_sw_)
      blk.page_ids.push_back(kTempPageId);  // borrowed until the window slides (paged_kv_cache.cc:2023-2027)
    else
      blk.page_ids.push_back(GetFreePage());
  }
  blk.seq_length += static_cast<int32_t>(append_length);
  SlideWindow(seq);
  if (support_sw_)
    for (int32_t& p : blk.page_ids)
      if (p == kTempPageId) p = GetFreePage();
  dirty_ = true;
}

void Cache::ConstructTokenTreeMask(const std::vector<Sequence*>& seqs, const int64_t* tree, int tree_size,
                                   const std::vector<std::vector<int32_t>>& ids_on_depths,
                                   const std::vector<std::vector<int32_t>>& trailing) {
  auto on_depth = [&](int i, int d) -> bool {
    if (!append_before_attn_) return true;
    return ids_on_depths[d][i] == seqs[i]->last_block_idx || (d + 1 == kMaxBlockDepth && !trailing[i].empty());
  };
  for (int d = 0; d < num_depths_; ++d) {
    IVec& mn = tree_mn_indptr_[d];
    IVec& mask = tree_mask_[d];
    std::vector<bool> here(cur_batch_, false);
    mn.clear();
    mask.clear();
    std::fill(is_chain_on_depths_.begin(), is_chain_on_depths_.end(), true);
    bool is_chain = true;
    mn.push_back(0);
    int64_t off = 0;
    for (int i = 0; i < cur_batch_; ++i) {
      const int64_t len = cur_lens_[i];
      here[i] = on_depth(i, d);
      if (!here[i]) {
        mn.push_back(mn.back());
        off += len;
        continue;
      }
      HCHECK(static_cast<int64_t>(seqs[i]->tree_parent.size()) <= blocks_[seqs[i]->last_block_idx].seq_length,
             "The token tree size is larger than the sequence length of the last block.");
      HCHECK(off + len <= tree_size, "Invalid token tree size.");
      for (int64_t k = 0; k < len; ++k) seqs[i]->tree_parent.push_back(static_cast<int32_t>(tree[off + k]));
      off += len;
      HCHECK(static_cast<int>(seqs[i]->tree_parent.size()) <= kMaxTreeSize,
             "The tree size is %ld which exceeds the maximum tree size limit %d", (long)len, kMaxTreeSize);
      mn.push_back(mn.back() + static_cast<int32_t>(seqs[i]->tree_parent.size()));
    }
    HCHECK(off == tree_size,
           "Invalid token tree size. The sum of \"append_lengths\" is %ld while there are %d elements in \"token_tree_parent_ptr\".",
           (long)off, tree_size);
    for (int i = 0; i < cur_batch_; ++i) {
      if (!here[i]) continue;
      Sequence* s = seqs[i];
      const int n = static_cast<int>(s->tree_parent.size());
      std::vector<int32_t> depth;
      depth.reserve(n);
      s->is_chain = true;
      s->committed = false;
      std::vector<std::vector<int>> children(n);
      std::vector<int> roots;
      for (int k = 0; k < n; ++k) {
        const int32_t par = s->tree_parent[k];
        HCHECK(par < k, "Invalid token tree. The parent of node %d in tree %d is %d, which is not smaller than %d", k, i, par, k);
        HCHECK(par >= -1, "Invalid token tree. The parent of node %d in tree %d is %d", k, i, par);
        if (par != k - 1) {
          s->is_chain = false;
          is_chain = false;
        }
        if (par != -1) {
          children[par].push_back(k);
          depth.push_back(depth[par] + 1);
        } else {
          depth.push_back(0);
          roots.push_back(k);
        }
      }
      // DFS pre-order numbering: node -> [order, end of subtree)
      std::vector<std::pair<int, int>> iv(n);
      int order = 0;
      std::function<int(int)> dfs = [&](int u) -> int {
        iv[u].first = order++;
        int ub = iv[u].first + 1;
        for (int ch : children[u]) ub = std::max(ub, dfs(ch));
        iv[u].second = ub;
        return ub;
      };
      for (int r : roots) dfs(r);
      for (int k = 0; k < n; ++k) {
        mask.push_back(iv[k].first);
        mask.push_back(iv[k].second);
      }
      s->tree_depth = std::move(depth);
    }
    is_chain_on_depths_[d] = is_chain;
    if (!append_before_attn_) break;
  }
}

// BeginForward is transactional: when it throws (unknown sequence, cache full, aux overflow, invalid tree) every
// sequence length, block, page list and the free-page stack are restored to what they were before the call, so the
// cache stays usable; there is just no current batch (the attention / commit entries refuse to run until a
// begin_forward completes).  The reference is not transactional (its errors are fatal ICHECKs).
void Cache::BeginForward(const int64_t* seq_ids, const int64_t* lens, int n, const int64_t* tree, int tree_size) {
  NvtxRange nvtx_range("vm.builtin.kv_state_begin_forward");
  batch_valid_ = false;
  HCHECK(n > 0, "begin_forward: the batch is empty");
  for (int i = 0; i < n; ++i) {
    Seq(seq_ids[i]);
    HCHECK(lens[i] > 0, "Append with length 0 is not allowed.");
    for (int j = 0; j < i; ++j) HCHECK(seq_ids[j] != seq_ids[i], "begin_forward: sequence %ld appears twice in the batch", (long)seq_ids[i]);
  }
  // snapshot of what the call may mutate: the batch's Sequence records and their last blocks, the free-page stack
  struct Saved {
    int64_t id;
    Sequence seq;
    int32_t block;
    Block blk;
  };
  std::vector<Saved> saved;
  saved.reserve(n);
  for (int i = 0; i < n; ++i) {
    const Sequence& s = seq_map_.at(seq_ids[i]);
    saved.push_back(Saved{seq_ids[i], s, s.last_block_idx, Block()});
    // a decode step of a plain sequence can only push pages onto its last block: remember the scalars and the page
    // count; a sliding-window sequence may also drop pages from the middle, so keep the list
    const Block& b = blocks_[s.last_block_idx];
    Block& kb = saved.back().blk;
    kb.seq_length = b.seq_length;
    kb.start_pos = b.start_pos;
    kb.sink_length = b.sink_length;
    kb.sliding_window_offset = b.sliding_window_offset;
    kb.parent_idx = b.parent_idx;
    kb.external_ref_cnt = static_cast<int>(b.page_ids.size());  // (re-used field: the page count before the call)
    if (s.sliding_window_size != -1) kb.page_ids = b.page_ids;
  }
  const std::vector<int32_t> free_before = support_sw_ ? free_pages_ : std::vector<int32_t>();
  const size_t free_size_before = free_pages_.size();
  try {
    BeginForwardImpl(seq_ids, lens, n, tree, tree_size);
  } catch (...) {
    for (auto it = saved.rbegin(); it != saved.rend(); ++it) {  // reverse: pages go back in the order they came
      const Saved& sv = *it;
      seq_map_.at(sv.id) = sv.seq;
      Block& b = blocks_[sv.block];
      const size_t pages_before = static_cast<size_t>(sv.blk.external_ref_cnt);
      if (sv.seq.sliding_window_size != -1) {
        b.page_ids = sv.blk.page_ids;
      } else {
        // pages were handed out from the back of the free stack in push order: give them back in reverse
        while (b.page_ids.size() > pages_before) {
          if (!support_sw_ && b.page_ids.back() != kTempPageId) free_pages_.push_back(b.page_ids.back());
          b.page_ids.pop_back();
        }
      }
      b.seq_length = sv.blk.seq_length;
      b.start_pos = sv.blk.start_pos;
      b.sink_length = sv.blk.sink_length;
      b.sliding_window_offset = sv.blk.sliding_window_offset;
    }
    if (support_sw_) free_pages_ = free_before;
    if (free_pages_.size() != free_size_before) fail("begin_forward rollback lost pages (%zu != %zu)", free_pages_.size(), free_size_before);
    throw;
  }
}

void Cache::BeginForwardImpl(const int64_t* seq_ids, const int64_t* lens, int n, const int64_t* tree, int tree_size) {
  cur_batch_ = n;
  cur_seq_ids_.assign(seq_ids, seq_ids + n);
  cur_lens_.assign(lens, lens + n);
  std::vector<Sequence*> seqs;
  std::vector<int32_t> last_len_before;
  is_decode_request_ = true;
  k_ragged_rope_off_.clear();
  for (int i = 0; i < n; ++i) {
    Sequence& s = Seq(seq_ids[i]);
    seqs.push_back(&s);
    last_len_before.push_back(blocks_[s.last_block_idx].seq_length);
    int32_t k_off = s.seq_length;
    if (!s.committed) k_off -= static_cast<int32_t>(s.tree_parent.size());
    k_ragged_rope_off_.push_back(k_off);
    s.seq_length += static_cast<int32_t>(lens[i]);
    if (lens[i] != 1) is_decode_request_ = false;
  }

  // ---- block ids per depth (attn_utils.h:242-279) ----
  std::vector<std::vector<int32_t>> traces, trailing;
  int depths = 0;
  for (int i = 0; i < n; ++i) {
    std::vector<int32_t> t = BlockTrace(*seqs[i]);
    if (static_cast<int>(t.size()) <= kMaxBlockDepth) {
      depths = std::max<int>(depths, static_cast<int>(t.size()));
      trailing.emplace_back();
      traces.push_back(std::move(t));
    } else {
      depths = std::max(depths, kMaxBlockDepth);
      trailing.emplace_back(t.begin() + kMaxBlockDepth, t.end());
      t.resize(kMaxBlockDepth);
      traces.push_back(std::move(t));
    }
  }
  std::vector<std::vector<int32_t>> ids_on_depths(depths, std::vector<int32_t>(n, -1));
  for (int d = 0; d < depths; ++d)
    for (int i = 0; i < n; ++i)
      if (d < static_cast<int>(traces[i].size())) ids_on_depths[d][i] = traces[i][d];
  num_depths_ = std::min(depths, kMaxBlockDepth);

  // ---- coalescing decision per depth (attn_utils.h:294-332) ----
  using Chunk = std::pair<int32_t, int32_t>;  // (block id, query rows that attend to it)
  std::vector<std::vector<Chunk>> chunks(num_depths_);
  use_decode_kernel_.clear();
  for (int d = 0; d < num_depths_; ++d) {
    const std::vector<int32_t>& ids = ids_on_depths[d];
    const bool enable_coalesce = d != kMaxBlockDepth - 1;
    std::vector<Chunk> plain, merged;
    int cur = ids[0];
    int run = static_cast<int>(lens[0]);
    int pages_merged = 0;
    int pages_plain = ids[0] != -1 ? static_cast<int>(blocks_[ids[0]].page_ids.size()) : 0;
    for (int i = 1; i < n; ++i) {
      if (ids[i] != -1) pages_plain += static_cast<int>(blocks_[ids[i]].page_ids.size());
      plain.emplace_back(ids[i - 1], static_cast<int32_t>(lens[i - 1]));
      if (ids[i] == cur) {
        run += static_cast<int>(lens[i]);
      } else {
        merged.emplace_back(cur, run);
        if (cur != -1) pages_merged += static_cast<int>(blocks_[cur].page_ids.size());
        cur = ids[i];
        run = static_cast<int>(lens[i]);
      }
    }
    plain.emplace_back(ids.back(), static_cast<int32_t>(lens[n - 1]));
    merged.emplace_back(cur, run);
    if (cur != -1) pages_merged += static_cast<int>(blocks_[cur].page_ids.size());
    const double ratio = pages_merged > 0 ? 1.0 * pages_plain / pages_merged : 0.0;
    const bool use_decode = is_decode_request_ && ratio < 32;
    chunks[d] = (use_decode || !enable_coalesce) ? plain : merged;
    use_decode_kernel_.push_back(use_decode);
  }
  if (num_depths_ == kMaxBlockDepth)
    HCHECK(static_cast<int64_t>(chunks[num_depths_ - 1].size()) == cur_batch_, "max-depth blocks must not coalesce");

  append_before_attn_ = !support_sw_ && use_decode_kernel_.back();
  bool has_previous_tree = false;
  for (Sequence* s : seqs) has_previous_tree |= !s->committed;
  if (has_previous_tree) append_before_attn_ = true;

  if (tree != nullptr) {
    HCHECK(!support_sw_, "Tree attention does not support sliding window.");
    HCHECK(rope_mode_ != TVMB200_ROPE_INLINE, "Tree attention does not support inline RoPE mode.");
    ConstructTokenTreeMask(seqs, tree, tree_size, ids_on_depths, trailing);
  } else {
    for (int i = 0; i < n; ++i) {
      HCHECK(seqs[i]->committed,
             "The input batch does not form a tree, in which case the sequences in the input batch are expected to have "
             "their accepted tokens token tree nodes committed. Please invoke CommitAcceptedTokenTreeNodes for sequence %ld",
             (long)seq_ids[i]);
      seqs[i]->is_chain = true;
      seqs[i]->tree_parent.clear();
      seqs[i]->tree_depth.clear();
    }
    std::fill(is_chain_on_depths_.begin(), is_chain_on_depths_.end(), true);
  }

  if (append_before_attn_)
    for (int i = 0; i < n; ++i) ReserveAppendLength(seqs[i], lens[i]);

  const int32_t ps = static_cast<int32_t>(page_size_);
  for (int d = 0; d < num_depths_; ++d) {
    IVec &qo = qo_indptr_[d], &pi = page_indptr_[d], &pv = page_indices_[d], &pis = page_indptr_sw_[d],
         &pvs = page_indices_sw_[d], &lpl = last_page_len_[d], &swo = sw_offset_[d], &snk = sink_size_[d],
         &kro = k_rope_off_[d], &kros = k_rope_off_sw_[d];
    for (IVec* v : {&qo, &pi, &pv, &pis, &pvs, &lpl, &swo, &snk, &kro, &kros}) v->clear();
    qo.push_back(0);
    pi.push_back(0);
    pis.push_back(0);
    for (int i = 0; i < static_cast<int>(chunks[d].size()); ++i) {
      const int32_t bid = chunks[d][i].first;
      qo.push_back(qo.back() + chunks[d][i].second);
      if (bid == -1) {
        pi.push_back(pi.back());
        pis.push_back(pis.back());
        lpl.push_back(0);
        swo.push_back(0);
        snk.push_back(0);
        kro.push_back(0);
        kros.push_back(0);
        continue;
      }
      const Block& blk = blocks_[bid];
      int32_t npages = static_cast<int32_t>(blk.page_ids.size());
      int32_t total_len = blk.seq_length;
      int32_t last_id = bid;
      for (int32_t p : blk.page_ids) pv.push_back(p);
      if (d == kMaxBlockDepth - 1) {
        // deepest kernel depth also swallows every deeper ("trailing") block of the sequence
        for (int32_t tid : trailing[i]) {
          const Block& tb = blocks_[tid];
          for (int32_t p : tb.page_ids) pv.push_back(p);
          npages += static_cast<int32_t>(tb.page_ids.size());
          total_len += tb.seq_length;
          last_id = tid;
        }
      }
      pi.push_back(pi.back() + npages);
      const int32_t n_sw = std::min(npages, LayerSwNumPages(total_len));
      pis.push_back(pis.back() + n_sw);
      for (int k = static_cast<int>(pv.size()) - n_sw; k < static_cast<int>(pv.size()); ++k) pvs.push_back(pv[k]);
      const Block& lb = blocks_[last_id];
      lpl.push_back(total_len == 0 ? 0 : (total_len - lb.sink_length + lb.sliding_window_offset - 1) % ps + 1);
      swo.push_back(support_layer_sw_ ? LayerSwOffset(total_len) : lb.sliding_window_offset);
      snk.push_back(lb.sink_length);
      kro.push_back(blk.start_pos);
      if (support_layer_sw_)
        kros.push_back(static_cast<int32_t>(std::max<int64_t>(0, blk.start_pos + total_len - layer_sws_)));
    }
  }

  if (!append_before_attn_)
    for (int i = 0; i < n; ++i) ReserveAppendLength(seqs[i], lens[i]);

  // ---- token -> rope position, token -> KV slot (paged_kv_cache.cc:1150-1183) ----
  q_rope_pos_.clear();
  append_pos_.clear();
  kv_tx_remote_pos_.clear();
  kv_tx_recver_.clear();
  kv_p2p_local_pos_.clear();
  kv_p2p_remote_pos_.clear();
  kv_p2p_recver_.clear();
  transfer_kv_ = page_to_page_transfer_kv_ = false;
  for (int i = 0; i < n; ++i) {
    const int64_t len = lens[i];
    const Block& blk = blocks_[seqs[i]->last_block_idx];
    for (int64_t pos = 0; pos < len; ++pos) {
      if (seqs[i]->tree_depth.empty()) {
        q_rope_pos_.push_back(static_cast<int32_t>(k_ragged_rope_off_[i] + pos));
      } else {
        const int64_t off_in_tree = static_cast<int64_t>(seqs[i]->tree_parent.size()) - len;
        HCHECK(off_in_tree >= 0, "token tree shorter than the append length");
        q_rope_pos_.push_back(k_ragged_rope_off_[i] + seqs[i]->tree_depth[off_in_tree + pos]);
      }
      const int32_t pos_in_block = static_cast<int32_t>(blk.seq_length - len + pos);
      if (last_len_before[i] + pos < blk.sink_length) {
        const int32_t o = static_cast<int32_t>(last_len_before[i] + pos);
        append_pos_.push_back(blk.page_ids[o / ps] * ps + o % ps);
      } else if (pos_in_block < blk.sink_length) {
        append_pos_.push_back(-1);  // the slot is pinned by the attention sink
      } else {
        const int32_t o = pos_in_block - blk.sink_length + blk.sliding_window_offset;
        append_pos_.push_back(blk.page_ids[o / ps] * ps + o % ps);
      }
      // paged_kv_cache.cc:1184-1196: where the token goes on the receiving side (if it is sent at all)
      const int64_t pos_in_seq = seqs[i]->seq_length - len + pos;
      if (pos_in_seq < seqs[i]->kv_send_start) {
        kv_tx_remote_pos_.push_back(-1);
        kv_tx_recver_.push_back(-1);
      } else {
        transfer_kv_ = true;
        const size_t at = static_cast<size_t>(pos_in_seq - seqs[i]->kv_send_start);
        HCHECK(at < seqs[i]->kv_remote_pos.size(), "sequence %ld sends more tokens than the receiver prepared", (long)seq_ids[i]);
        kv_tx_remote_pos_.push_back(seqs[i]->kv_remote_pos[at]);
        kv_tx_recver_.push_back(seqs[i]->kv_recver_pe_offset);
      }
    }
    if (!seqs[i]->kv_local_pos.empty()) {  // :1197-1210 rows cached before mark_send: page-to-page, once
      page_to_page_transfer_kv_ = true;
      for (size_t k = 0; k < seqs[i]->kv_local_pos.size(); ++k) {
        kv_p2p_local_pos_.push_back(seqs[i]->kv_local_pos[k]);
        kv_p2p_remote_pos_.push_back(seqs[i]->kv_remote_pos[k]);
        kv_p2p_recver_.push_back(seqs[i]->kv_recver_pe_offset);
      }
      seqs[i]->kv_local_pos.clear();
    }
  }
  BuildAuxViews();
  batch_valid_ = true;
}

// layout of the merged buffer = order of SyncAuxArrayToDevice (paged_kv_cache.cc:2392-2512)
void Cache::BuildAuxViews() {
  cur_len_indptr_.clear();
  cur_len_indptr_.push_back(0);
  for (int64_t l : cur_lens_) cur_len_indptr_.push_back(cur_len_indptr_.back() + static_cast<int32_t>(l));
  total_append_ = cur_len_indptr_.back();
  HCHECK(total_append_ == static_cast<int64_t>(append_pos_.size()), "append position map size mismatch");
  HCHECK(total_append_ <= prefill_chunk_, "the batch has %ld tokens, more than prefill_chunk_size %ld", (long)total_append_, (long)prefill_chunk_);
  stage_off_ = 0;
  v_q_rope_pos_ = Put(q_rope_pos_);
  for (int d = 0; d < num_depths_; ++d) v_qo_indptr_[d] = Put(qo_indptr_[d]);
  for (int d = 0; d < num_depths_; ++d) v_page_indptr_[d] = Put(page_indptr_[d]);
  for (int d = 0; d < num_depths_; ++d) v_page_indices_[d] = Put(page_indices_[d]);
  if (support_layer_sw_) {
    for (int d = 0; d < num_depths_; ++d) v_page_indptr_sw_[d] = Put(page_indptr_sw_[d]);
    for (int d = 0; d < num_depths_; ++d) v_page_indices_sw_[d] = Put(page_indices_sw_[d]);
  }
  for (int d = 0; d < num_depths_; ++d) {
    v_length_info_[d] = support_sw_ ? Put3(last_page_len_[d], sw_offset_[d], sink_size_[d]) : Put(last_page_len_[d]);
    if (support_layer_sw_) v_length_info_sw_[d] = Put3(last_page_len_[d], sw_offset_[d], sink_size_[d]);
  }
  for (int d = 0; d < num_depths_; ++d) {
    v_k_rope_off_[d] = Put(k_rope_off_[d]);
    if (support_layer_sw_) v_k_rope_off_sw_[d] = Put(k_rope_off_sw_[d]);
  }
  v_cur_len_indptr_ = Put(cur_len_indptr_);
  v_k_ragged_rope_off_ = Put(k_ragged_rope_off_);
  v_append_pos_ = Put(append_pos_);
  // the five kv-transfer maps (SyncAuxArrayToDevice steps 10-14, paged_kv_cache.cc:2484-2505): two of total_append_
  // elements (-1 = not sent), three page-to-page ones that are empty except right after a mark_send
  v_kv_tx_remote_pos_ = Put(kv_tx_remote_pos_);
  v_kv_tx_recver_ = Put(kv_tx_recver_);
  v_kv_p2p_local_pos_ = Put(kv_p2p_local_pos_);
  v_kv_p2p_remote_pos_ = Put(kv_p2p_remote_pos_);
  v_kv_p2p_recver_ = Put(kv_p2p_recver_);
  for (int d = 0; d < num_depths_; ++d) {
    if (!is_chain_on_depths_[d]) {
      v_tree_mask_[d] = Put(tree_mask_[d]);
      v_tree_mask_[d].rows = static_cast<int64_t>(tree_mask_[d].size()) / 2;
      v_tree_mn_[d] = Put(tree_mn_indptr_[d]);
    }
  }
  HCHECK(stage_off_ <= aux_capacity_, "auxiliary buffer overflow (%ld > %ld)", (long)stage_off_, (long)aux_capacity_);
  dirty_ = true;
}

// Device scratch of this cache's kernel set on (device, compute stream): the split-KV partials of decode and -- for
// caches that rotate inline or slide (rope mode "inline", sliding-window support, per-layer windows: the flavours the
// tcgen05 prefill reaches through its gather / rotate pre-pass) -- rotated q of a full prefill chunk plus position-
// ordered K / V of every page of one layer.  Reserved the FIRST time a stream is seen, so that no later callback
// allocates (the stream is a per-call argument: it cannot be done in the constructor).
void Cache::EnsureScratch(cudaStream_t compute) {
  if (scratch_streams_.count(compute)) return;
  int64_t bytes = int64_t(32) << 20;
  if (rope_mode_ == TVMB200_ROPE_INLINE || support_sw_ || support_layer_sw_) {
    const int64_t row = head_dim_ * 2;
    const int64_t want = prefill_chunk_ * num_qo_heads_ * row + 2 * num_total_pages_ * page_size_ * num_kv_heads_ * row +
                         (reserved_seqs_ + 64) * 4 + 4096;
    if (want > bytes && want <= (int64_t(2) << 30)) bytes = want;
  }
  HCHECK(tvmb200_reserve_workspace_stream(device_, bytes, compute) == 0, "%s", tvmb200_last_error());
  scratch_streams_.insert(compute);
}

void Cache::SyncAux(cudaStream_t compute) {
  NvtxRange nvtx_range("SyncAuxArrayToDevice");
  if (!planning_only()) EnsureScratch(compute);
  if (!dirty_ || planning_only()) {
    dirty_ = false;
    return;
  }
  const int i = aux_cur_ ^ 1;  // the buffer pair not used by the previous step
  if (aux_used_[i]) {
    HCUDA(cudaEventSynchronize(ev_aux_copied_[i]));                     // host: the old DMA out of stage_pinned_[i] is done
    HCUDA(cudaStreamWaitEvent(copy_stream_, ev_aux_readers_[i], 0));   // device: kernels that read aux_dev_[i] are done
  }
  std::memcpy(stage_pinned_[i], stage_.data(), static_cast<size_t>(stage_off_) * 4);
  HCUDA(cudaMemcpyAsync(aux_dev_[i], stage_pinned_[i], static_cast<size_t>(stage_off_) * 4, cudaMemcpyHostToDevice, copy_stream_));
  HCUDA(cudaEventRecord(ev_aux_copied_[i], copy_stream_));
  // the copy stream also carries page copies / compactions issued since the last step: one wait orders them all
  HCUDA(cudaStreamWaitEvent(compute, ev_aux_copied_[i], 0));
  aux_used_[i] = true;
  aux_cur_ = i;
  dirty_ = false;
}

// MHASelfAttnInternal (paged_kv_cache.cc:2182-2206): the new tokens against themselves, causal or tree-masked.
void Cache::SelfAttnInternal(const void* q, const void* k, const void* v, void* o, float* lse, double sm_scale,
                             cudaStream_t st) {
  const int64_t n = total_append_;
  const int32_t hq = static_cast<int32_t>(num_qo_heads_), hkv = static_cast<int32_t>(num_kv_heads_),
                d = static_cast<int32_t>(head_dim_);
  const bool plan = planning_only();
  const int rot = rope_mode_ == TVMB200_ROPE_INLINE;
  if (is_chain_on_depths_[0]) {
    TRACE("prefill_ragged", {TF({n, hq, d}), TI(v_cur_len_indptr_), TF({n, hkv, d}), TF({n, hkv, d}), TI(v_cur_len_indptr_),
                             TI(v_q_rope_pos_), TI(v_k_ragged_rope_off_), TF({n, hq, d}), TF({n, hq}, "float32"), SI(1),
                             SI(rot), SF(rotary_scale_), SF(rotary_theta_), SF(sm_scale)});
    if (!plan)
      Rc(tvmb200_attention_prefill_ragged(q, dev(v_cur_len_indptr_), k, v, dev(v_cur_len_indptr_), dev(v_q_rope_pos_),
                                          dev(v_k_ragged_rope_off_), o, lse, static_cast<int32_t>(cur_batch_),
                                          static_cast<int32_t>(n), static_cast<int32_t>(n), hq, hkv, d, 1, rot,
                                          static_cast<float>(rotary_scale_), static_cast<float>(rotary_theta_),
                                          static_cast<float>(sm_scale), dtype_, st));
  } else {
    TRACE("tree_ragged", {TF({n, hq, d}), TI(v_cur_len_indptr_), TF({n, hkv, d}), TF({n, hkv, d}), TI(v_cur_len_indptr_),
                          TI(v_q_rope_pos_), TI(v_tree_mn_[0]), TI(v_tree_mask_[0]), TF({n, hq, d}), TF({n, hq}, "float32"),
                          SI(rot), SF(rotary_scale_), SF(rotary_theta_), SF(sm_scale)});
    if (!plan)
      Rc(tvmb200_attention_prefill_tree_ragged(q, dev(v_cur_len_indptr_), k, v, dev(v_cur_len_indptr_), dev(v_q_rope_pos_),
                                               dev(v_tree_mn_[0]), dev(v_tree_mask_[0]), o, lse,
                                               static_cast<int32_t>(cur_batch_), static_cast<int32_t>(n),
                                               static_cast<int32_t>(n), hq, hkv, d, rot, static_cast<float>(rotary_scale_),
                                               static_cast<float>(rotary_theta_), static_cast<float>(sm_scale), dtype_, st));
  }
}

// MHACrossAttnInternal (paged_kv_cache.cc:2216-2299): q against the cached pages of every block depth; the first kernel
// writes (o, lse), every later one writes the temporaries and is merged in place.  fused_qkv != nullptr: the decode
// depth runs as the fused split_rotary + transpose_append + decode launch on the un-split qkv.
bool Cache::CrossAttnInternal(int64_t layer_id, const void* q, void* o, float* lse_out, double sm_scale, bool is_first,
                              bool causal, const void* fused_qkv, cudaStream_t st) {
  const int64_t local = layer_id - layer_begin_;
  const int64_t n = total_append_;
  const int32_t hq = static_cast<int32_t>(num_qo_heads_), hkv = static_cast<int32_t>(num_kv_heads_),
                d = static_cast<int32_t>(head_dim_), ps = static_cast<int32_t>(page_size_);
  const bool plan = planning_only();
  void* pages = plan ? nullptr : pages_[local];
  const int rot = rope_mode_ == TVMB200_ROPE_INLINE;
  const int64_t apply_rope = rope_mode_ == TVMB200_ROPE_NORMAL;
  const bool layer_sw = attn_kinds_[layer_id] == TVMB200_ATTN_MHA_SLIDING;
  const bool sw_flavour = support_sw_ || layer_sw;  // the `_sliding_window` kernels take [3,B] length_info
  bool cross_done = false;
  for (int dd = 0; dd < num_depths_; ++dd) {
    if (page_indices_[dd].empty()) continue;
    void* out = is_first ? o : tmp_o_;
    float* lse = is_first ? lse_out : tmp_lse_;
    const View &pip = layer_sw ? v_page_indptr_sw_[dd] : v_page_indptr_[dd];
    const View &piv = layer_sw ? v_page_indices_sw_[dd] : v_page_indices_[dd];
    const View &li = layer_sw ? v_length_info_sw_[dd] : v_length_info_[dd];
    const View &kro = layer_sw ? v_k_rope_off_sw_[dd] : v_k_rope_off_[dd];
    const double theta = layer_sw ? 10000.0 : rotary_theta_;
    const double scale = layer_sw ? 1.0 : rotary_scale_;
    const int32_t B = static_cast<int32_t>(v_qo_indptr_[dd].size - 1);
    const int32_t nnz = static_cast<int32_t>(piv.size);
    if (append_before_attn_ && !is_chain_on_depths_[dd]) {
      TRACE("tree_paged", {TF({n, hq, d}), TI(v_qo_indptr_[dd]), TF({num_total_pages_, 2, hkv, ps, d}), TI(pip), TI(piv), TI(li),
                           TI(kro), TI(v_q_rope_pos_), TF({n, hq, d}), TF({n, hq}, "float32"), SI(rot), SF(scale), SF(theta),
                           SF(sm_scale), TI(v_tree_mn_[dd]), TI(v_tree_mask_[dd])});
      if (!plan)
        Rc(tvmb200_attention_prefill_tree_paged(q, dev(v_qo_indptr_[dd]), pages, dev(pip), dev(piv), dev(li), dev(kro),
                                                dev(v_q_rope_pos_), out, lse, B, static_cast<int32_t>(n), nnz,
                                                num_total_pages_, hq, hkv, ps, d, rot, static_cast<float>(scale),
                                                static_cast<float>(theta), static_cast<float>(sm_scale), dev(v_tree_mn_[dd]),
                                                dev(v_tree_mask_[dd]), dtype_, st));
    } else if (use_decode_kernel_[dd]) {
      TRACE(sw_flavour ? "decode_sliding_window" : "decode",
            {TF({n, hq, d}), TF({num_total_pages_, 2, hkv, ps, d}), TI(pip), TI(piv), TI(li), TI(kro), TI(v_q_rope_pos_),
             TF({n, hq, d}), TF({n, hq}, "float32"), SI(rot), SF(scale), SF(theta), SF(sm_scale)});
      if (fused_qkv != nullptr)
        Rc(tvmb200_attention_decode_fused_qkv(fused_qkv, dev(v_q_rope_pos_), dev(v_append_pos_), pages, dev(pip), dev(piv),
                                              dev(li), dev(kro), out, lse, B, nnz, num_total_pages_, hq, hkv, ps, d, 0,
                                              apply_rope, static_cast<float>(scale), static_cast<float>(theta),
                                              static_cast<float>(sm_scale), dtype_, st));
      else if (!plan)
        Rc(tvmb200_attention_decode(q, pages, dev(pip), dev(piv), dev(li), dev(kro), dev(v_q_rope_pos_), out, lse, B, nnz,
                                    num_total_pages_, hq, hkv, ps, d, sw_flavour ? 1 : 0, rot, static_cast<float>(scale),
                                    static_cast<float>(theta), static_cast<float>(sm_scale), dtype_, st));
    } else {
      TRACE(sw_flavour ? "prefill_sliding_window" : "prefill",
            {TF({n, hq, d}), TI(v_qo_indptr_[dd]), TF({num_total_pages_, 2, hkv, ps, d}), TI(pip), TI(piv), TI(li), TI(kro),
             TI(v_q_rope_pos_), TF({n, hq, d}), TF({n, hq}, "float32"), SI(causal ? 1 : 0), SI(rot), SF(scale), SF(theta),
             SF(sm_scale)});
      if (!plan)
        Rc(tvmb200_attention_prefill_paged(q, dev(v_qo_indptr_[dd]), pages, dev(pip), dev(piv), dev(li), dev(kro),
                                           dev(v_q_rope_pos_), out, lse, B, static_cast<int32_t>(n), nnz, num_total_pages_,
                                           hq, hkv, ps, d, sw_flavour ? 1 : 0, sw_flavour ? static_cast<int32_t>(layer_sws_) : 0,
                                           causal ? 1 : 0, rot, static_cast<float>(scale), static_cast<float>(theta),
                                           static_cast<float>(sm_scale), dtype_, st));
    }
    if (!is_first) {
      TRACE("merge", {TF({n, hq, d}), TF({n, hq}, "float32"), TF({n, hq, d}), TF({n, hq}, "float32")});
      if (!plan) Rc(tvmb200_merge_state_inplace(o, lse_out, tmp_o_, tmp_lse_, n, hq, d, dtype_, st));
    } else {
      is_first = false;
    }
    cross_done = true;
  }
  return cross_done;
}

// AttentionInternal (paged_kv_cache.cc:2161-2180)
void Cache::AttentionInternal(int64_t layer_id, const void* q, const void* k, const void* v, void* o, double sm_scale,
                              const void* fused_qkv, cudaStream_t st) {
  bool is_first = true;
  if (!append_before_attn_) {
    is_first = false;
    SelfAttnInternal(q, k, v, o, merged_lse_, sm_scale, st);
  }
  const bool self_done = !is_first;
  const bool causal = !append_before_attn_ && attn_kinds_[layer_id] == TVMB200_ATTN_MHA_SLIDING;
  const bool cross_done = CrossAttnInternal(layer_id, q, o, merged_lse_, sm_scale, is_first, causal, fused_qkv, st);
  HCHECK(self_done || cross_done, "Both self-attention and cross-attention are not computed.");
}

int64_t Cache::CheckLayer(int64_t layer_id) const {
  HCHECK(batch_valid_, "no batch to attend over: call begin_forward first (the last one did not complete)");
  const int64_t local = layer_id - layer_begin_;
  HCHECK(local >= 0 && local < num_layers_, "layer_id %ld is outside this cache's layers [%ld, %ld)", (long)layer_id,
         (long)layer_begin_, (long)(layer_begin_ + num_layers_));
  return local;
}

void Cache::MarkAttentionDone(cudaStream_t st) {
  if (planning_only()) return;
  HCUDA(cudaEventRecord(ev_attn_done_, st));
  HCUDA(cudaEventRecord(ev_aux_readers_[aux_cur_], st));
}

void Cache::AttentionWithFusedQKV(int64_t layer_id, double sm_scale, const void* qkv, void* o, int64_t rows,
                                  cudaStream_t st) {
  NvtxRange nvtx_range("vm.builtin.attention_kv_cache_attention_with_fused_qkv");
  const int64_t local = CheckLayer(layer_id);
  const int64_t n = total_append_;
  HCHECK(n <= rows, "qkv has %ld rows but the batch appends %ld tokens", (long)rows, (long)n);
  SyncAux(st);
  const int32_t hq = static_cast<int32_t>(num_qo_heads_), hkv = static_cast<int32_t>(num_kv_heads_),
                d = static_cast<int32_t>(head_dim_), ps = static_cast<int32_t>(page_size_);
  const bool plan = planning_only();
  void* pages = plan ? nullptr : pages_[local];

  // Part 2: split fused qkv (+ RoPE when the mode is "normal")
  const int64_t apply_rope = rope_mode_ == TVMB200_ROPE_NORMAL;
  TRACE("split_rotary", {TF({n, hq + 2 * hkv, d}), TI(v_q_rope_pos_), TF({n, hq, d}), TF({n, hkv, d}), TF({n, hkv, d}), SI(apply_rope)});
  auto trace_append = [&]() {
    TRACE("transpose_append", {TF({num_total_pages_, 2, hkv, ps, d}), TF({n, hkv, d}), TF({n, hkv, d}), TI(v_append_pos_)});
  };
  auto append = [&]() {
    trace_append();
    if (!plan) Rc(tvmb200_transpose_append(pages, tmp_k_, tmp_v_, dev(v_append_pos_), n, num_total_pages_, hkv, ps, d, dtype_, st));
  };
  // A plain decode step (one new token per sequence, a single block depth, K cached rotated or un-rotated, D = 128) runs
  // f_split_rotary + f_transpose_append + f_attention_decode as ONE launch (tvmb200_attention_decode_fused_qkv); the call
  // trace still records the three callbacks in the reference's order.
  const bool fuse_step = !plan && append_before_attn_ && num_depths_ == 1 && use_decode_kernel_[0] &&
                         is_chain_on_depths_[0] && !page_indices_[0].empty() && rope_mode_ != TVMB200_ROPE_INLINE && d == 128 &&
                         attn_kinds_[layer_id] != TVMB200_ATTN_MHA_SLIDING && !support_sw_ &&
                         static_cast<int64_t>(v_qo_indptr_[0].size - 1) == n && !transfer_kv_ &&
                         (apply_rope == 0 || tvmb200_get_rope_scaling_kind() <= TVMB200_ROPE_SCALING_LLAMA3);
  // the previous layer's transfer still reads tmp_k_ / tmp_v_ (paged_kv_cache.cc:1356-1359)
  if (!plan && kv_sent_pending_) {
    HCUDA(cudaStreamWaitEvent(st, ev_kv_sent_, 0));
    kv_sent_pending_ = false;
  }
  if (append_before_attn_ && fuse_step) {
    trace_append();
  } else if (append_before_attn_) {
    // the reference's f_split_rotary -> f_transpose_append pair (paged_kv_cache.cc:1360, :1371) as one launch
    trace_append();
    if (!plan)
      Rc(tvmb200_split_rotary_append(qkv, dev(v_q_rope_pos_), dev(v_append_pos_), tmp_q_, tmp_k_, tmp_v_, pages, n,
                                     num_total_pages_, hq, hkv, ps, d, 0, apply_rope, static_cast<float>(rotary_scale_),
                                     static_cast<float>(rotary_theta_), dtype_, st));
  } else if (!plan) {
    Rc(tvmb200_split_rotary(qkv, dev(v_q_rope_pos_), tmp_q_, tmp_k_, tmp_v_, n, hq, hkv, d, 0, apply_rope,
                            static_cast<float>(rotary_scale_), static_cast<float>(rotary_theta_), dtype_, st));
  }

  // Part 4: KV transfer (paged_kv_cache.cc:1374-1394) on the transfer stream, overlapping the attention below: first
  // the rows this cache already held when the sequence was marked (page to page), then this step's fresh k / v rows
  if (page_to_page_transfer_kv_ || transfer_kv_) {
    HCHECK(kv_transfer_enabled_, "a sequence is marked for KV transfer but the cache was not set up for it (tvmb200_cache_enable_kv_transfer)");
    TRACE("kv_transfer", {SI(page_to_page_transfer_kv_ ? static_cast<int64_t>(kv_p2p_local_pos_.size()) : 0), SI(transfer_kv_ ? n : 0)});
    if (!plan) {
      void* const* table = remote_pages_[local].data();
      HCUDA(cudaEventRecord(ev_kv_ready_, st));
      HCUDA(cudaStreamWaitEvent(kv_transfer_stream_, ev_kv_ready_, 0));
      if (page_to_page_transfer_kv_)
        Rc(tvmb200_kv_transfer_page_to_page(table, pages, dev(v_kv_p2p_remote_pos_), dev(v_kv_p2p_local_pos_),
                                            dev(v_kv_p2p_recver_), static_cast<int64_t>(kv_p2p_local_pos_.size()), hkv,
                                            kv_remote_num_kv_heads_, ps, d, kv_local_tp_rank_, kv_num_pe_, dtype_,
                                            kv_transfer_stream_));
      if (transfer_kv_)
        Rc(tvmb200_kv_transfer(table, tmp_k_, tmp_v_, dev(v_kv_tx_remote_pos_), dev(v_kv_tx_recver_), n, hkv,
                               kv_remote_num_kv_heads_, ps, d, kv_local_tp_rank_, kv_num_pe_, dtype_, kv_transfer_stream_));
      HCUDA(cudaEventRecord(ev_kv_sent_, kv_transfer_stream_));
      kv_sent_pending_ = true;
    }
  }

  // Part 5: attention
  AttentionInternal(layer_id, tmp_q_, tmp_k_, tmp_v_, o, sm_scale, fuse_step ? qkv : nullptr, st);
  if (!append_before_attn_) append();
  if (!plan && kv_sent_pending_ && local == num_layers_ - 1) {
    // last layer of the step: the caller's stream owns the transfers from here on (EndForward has no stream argument)
    HCUDA(cudaStreamWaitEvent(st, ev_kv_sent_, 0));
    kv_sent_pending_ = false;
  }
  MarkAttentionDone(st);
}

void Cache::EnableKVTransfer(int32_t local_tp_rank, int32_t num_pe, int32_t remote_num_kv_heads) {
  HCHECK(num_pe >= 1 && num_pe <= 64 && local_tp_rank >= 0 && remote_num_kv_heads > 0, "bad KV-transfer geometry");
  kv_transfer_enabled_ = true;
  kv_local_tp_rank_ = local_tp_rank;
  kv_num_pe_ = num_pe;
  kv_remote_num_kv_heads_ = remote_num_kv_heads;
  remote_pages_.assign(static_cast<size_t>(num_layers_), std::vector<void*>(static_cast<size_t>(num_pe), nullptr));
  // room for the three page-to-page maps (up to every cached token of the batch's sequences): the merged aux buffers grow
  const int64_t extra = 3 * ((num_total_pages_ * page_size_ + 3) / 4 * 4);
  if (!kv_aux_grown_) {
    aux_capacity_ += extra;
    stage_.resize(static_cast<size_t>(aux_capacity_), 0);
    if (!planning_only()) {
      HCUDA(cudaSetDevice(device_));
      HCUDA(cudaDeviceSynchronize());
      for (int i = 0; i < 2; ++i) {
        HCUDA(cudaFree(aux_dev_[i]));
        HCUDA(cudaFreeHost(stage_pinned_[i]));
        HCUDA(cudaMalloc(reinterpret_cast<void**>(&aux_dev_[i]), aux_capacity_ * 4));
        HCUDA(cudaMallocHost(reinterpret_cast<void**>(&stage_pinned_[i]), aux_capacity_ * 4));
        aux_used_[i] = false;
      }
      dirty_ = true;
    }
    kv_aux_grown_ = true;
  }
  if (!planning_only() && kv_transfer_stream_ == nullptr) {
    HCUDA(cudaStreamCreateWithFlags(&kv_transfer_stream_, cudaStreamNonBlocking));
    HCUDA(cudaEventCreateWithFlags(&ev_kv_ready_, cudaEventDisableTiming));
    HCUDA(cudaEventCreateWithFlags(&ev_kv_sent_, cudaEventDisableTiming));
  }
}

void Cache::SetRemotePages(int32_t pe, int64_t local_layer, void* peer_ptr) {
  HCHECK(kv_transfer_enabled_, "tvmb200_cache_enable_kv_transfer first");
  HCHECK(pe >= 0 && pe < kv_num_pe_ && local_layer >= 0 && local_layer < num_layers_, "remote pages: pe %d / layer %ld out of range", pe, (long)local_layer);
  remote_pages_[static_cast<size_t>(local_layer)][static_cast<size_t>(pe)] = peer_ptr;
}

// DisaggPrepareRecv (paged_kv_cache.cc:1220-1248): reserve the slots through BeginForward and hand their ids back,
// run-length compressed as [n, begin_1, length_1, ..., begin_n, length_n]
std::vector<int64_t> Cache::DisaggPrepareRecv(int64_t seq_id, int64_t append_length) {
  HCHECK(append_length > 0, "disagg_prepare_recv: append length %ld", (long)append_length);
  const int64_t ids[1] = {seq_id}, lens[1] = {append_length};
  BeginForward(ids, lens, 1, nullptr, 0);
  HCHECK(static_cast<int64_t>(append_pos_.size()) == append_length, "append position map size mismatch");
  std::vector<int64_t> out{1, append_pos_[0]};
  for (int64_t i = 1; i < append_length; ++i) {
    if (append_pos_[i] != append_pos_[i - 1] + 1) {
      out.push_back(append_pos_[i - 1] - out.back() + 1);
      ++out[0];
      out.push_back(append_pos_[i]);
    }
  }
  out.push_back(append_pos_.back() - out.back() + 1);
  return out;
}

// DisaggMarkSend (paged_kv_cache.cc:1250-1301)
void Cache::DisaggMarkSend(int64_t seq_id, int64_t begin, const int64_t* comp, int64_t n, int32_t recver_pe_offset) {
  HCHECK(kv_transfer_enabled_, "disagg_mark_send: the cache was not set up for KV transfer (tvmb200_cache_enable_kv_transfer)");
  auto it = seq_map_.find(seq_id);
  HCHECK(it != seq_map_.end(), "The sequence \"%ld\" cannot be found in KV cache.", (long)seq_id);
  HCHECK(n >= 1 && comp[0] >= 0 && n == 2 * comp[0] + 1, "disagg_mark_send: malformed compressed position map");
  Sequence& seq = it->second;
  seq.kv_send_start = begin;
  seq.kv_remote_pos.clear();
  for (int64_t i = 0; i < comp[0]; ++i)
    for (int64_t j = 0; j < comp[2 * i + 2]; ++j) seq.kv_remote_pos.push_back(static_cast<int32_t>(comp[2 * i + 1] + j));
  seq.kv_recver_pe_offset = recver_pe_offset;
  seq.kv_local_pos.clear();
  if (begin >= seq.seq_length) return;
  // tokens [begin, seq_length) are cached already: their slots, oldest first
  HCHECK(static_cast<int64_t>(seq.kv_remote_pos.size()) > seq.seq_length - begin, "Need at least one token to prefill");
  const int64_t want = seq.seq_length - begin;
  std::vector<int32_t> rev;
  for (int32_t b = seq.last_block_idx; b != -1 && static_cast<int64_t>(rev.size()) < want; b = blocks_[b].parent_idx) {
    const Block& blk = blocks_[b];
    for (int32_t i = blk.seq_length - 1; i >= 0 && static_cast<int64_t>(rev.size()) < want; --i) {
      const int32_t off = i < blk.sink_length ? i : i - blk.sink_length + blk.sliding_window_offset;
      rev.push_back(blk.page_ids[off / page_size_] * static_cast<int32_t>(page_size_) + off % static_cast<int32_t>(page_size_));
    }
  }
  seq.kv_local_pos.assign(rev.rbegin(), rev.rend());
}

// SelfAttention (paged_kv_cache.cc:1404-1445): q [n, Hq, D] against the step's own k / v [n, Hkv, D] (not the cache).
void Cache::SelfAttention(int64_t layer_id, double sm_scale, const void* q, const void* k, const void* v, void* o,
                          float* lse, int64_t rows, cudaStream_t st) {
  NvtxRange nvtx_range("vm.builtin.attention_kv_cache_self_attention");
  CheckLayer(layer_id);
  HCHECK(attn_kinds_[layer_id] == TVMB200_ATTN_MHA, "self_attention: layer %ld is not an MHA layer (MLA is not built)",
         (long)layer_id);
  HCHECK(rows == total_append_, "self_attention: the tensors have %ld rows but the batch appends %ld tokens", (long)rows,
         (long)total_append_);
  SyncAux(st);
  SelfAttnInternal(q, k, v, o, lse, sm_scale, st);
  MarkAttentionDone(st);
}

// CrossAttention (paged_kv_cache.cc:1447-1485): q against the cached KV of the layer, no causal mask.
void Cache::CrossAttention(int64_t layer_id, double sm_scale, const void* q, void* o, float* lse, int64_t rows,
                           cudaStream_t st) {
  NvtxRange nvtx_range("vm.builtin.attention_kv_cache_cross_attention");
  CheckLayer(layer_id);
  HCHECK(attn_kinds_[layer_id] == TVMB200_ATTN_MHA, "cross_attention: layer %ld is not an MHA layer (MLA is not built)",
         (long)layer_id);
  HCHECK(rows == total_append_, "cross_attention: the tensors have %ld rows but the batch appends %ld tokens", (long)rows,
         (long)total_append_);
  SyncAux(st);
  CrossAttnInternal(layer_id, q, o, lse, sm_scale, /*is_first=*/true, /*causal=*/false, nullptr, st);
  MarkAttentionDone(st);
}

// AttentionWithSharedKV (paged_kv_cache.cc:1487-1529): another logical layer queries the K / V of `source_layer_id`
// (pages already hold the step's tokens when the append ran before the source layer's attention).
void Cache::AttentionWithSharedKV(int64_t source_layer_id, double sm_scale, const void* q, const void* cur_k,
                                  const void* cur_v, void* o, int64_t rows, cudaStream_t st) {
  NvtxRange nvtx_range("vm.builtin.attention_kv_cache_attention_with_shared_kv");
  CheckLayer(source_layer_id);
  HCHECK(rows == total_append_, "attention_with_shared_kv: the tensors have %ld rows but the batch appends %ld tokens",
         (long)rows, (long)total_append_);
  SyncAux(st);
  AttentionInternal(source_layer_id, q, cur_k, cur_v, o, sm_scale, nullptr, st);
  MarkAttentionDone(st);
}

// MergeAttnOutputInplace (paged_kv_cache.cc:1553-1559): f_merge_inplace_[1] on caller tensors.
void Cache::MergeAttnOutputInplace(void* o_self, float* lse_self, const void* o_cross, const float* lse_cross, int64_t n,
                                   int64_t num_heads, int64_t head_dim, cudaStream_t st) {
  NvtxRange nvtx_range("vm.builtin.attention_kv_cache_merge_attn_output_inplace");
  TRACE("merge", {TF({n, num_heads, head_dim}), TF({n, num_heads}, "float32"), TF({n, num_heads, head_dim}),
                  TF({n, num_heads}, "float32")});
  if (!planning_only())
    Rc(tvmb200_merge_state_inplace(o_self, lse_self, o_cross, lse_cross, n, static_cast<int32_t>(num_heads),
                                   static_cast<int32_t>(head_dim), dtype_, st));
}

void Cache::CommitAcceptedTokenTreeNodes(const int64_t* seq_ids, const int64_t* leaves, int n) {
  NvtxRange nvtx_range("vm.builtin.attention_kv_cache_commit_accepted_token_tree_nodes");
  // the reference indexes cur_append_lengths_ / cur_seq_ids_ / append_position_map_host_ of the last BeginForward with
  // the positions of seq_ids (paged_kv_cache.cc:1625-1650): the sequences must be that batch's, in its order
  HCHECK(batch_valid_, "commit_accepted_token_tree_nodes: the last begin_forward did not complete");
  HCHECK(n <= cur_batch_, "commit_accepted_token_tree_nodes: %d sequences, but the last begin_forward had %ld", n, (long)cur_batch_);
  std::vector<Sequence*> seqs;
  bool is_chain = true;
  for (int i = 0; i < n; ++i) {
    HCHECK(seq_ids[i] == cur_seq_ids_[i], "commit_accepted_token_tree_nodes: sequence %ld at position %d is not sequence %ld "
           "of the last begin_forward", (long)seq_ids[i], i, (long)cur_seq_ids_[i]);
    Sequence& s = Seq(seq_ids[i]);
    HCHECK(s.tree_depth.size() == s.tree_parent.size(), "token tree of sequence %ld is incomplete", (long)seq_ids[i]);
    seqs.push_back(&s);
    is_chain = s.is_chain;
    HCHECK(leaves[i] == -1 || !s.committed, "The accepted nodes of sequence %ld are already committed.", (long)seq_ids[i]);
    HCHECK(leaves[i] >= -1, "Invalid tree index %ld which is less than -1", (long)leaves[i]);
    HCHECK(leaves[i] < static_cast<int64_t>(s.tree_parent.size()),
           "Invalid tree index %ld which is larger than or equals to the append length %zu of the sequence", (long)leaves[i],
           s.tree_parent.size());
  }
  if (!is_chain) {
    commit_indptr_.assign(1, 0);
    commit_src_.clear();
    commit_dst_.clear();
    for (int i = 0; i < n; ++i) {
      if (leaves[i] == -1) {
        commit_indptr_.push_back(commit_indptr_.back());
        continue;
      }
      std::vector<int32_t> path;
      for (int node = static_cast<int>(leaves[i]); node != -1; node = seqs[i]->tree_parent[node]) path.push_back(node);
      HCHECK(static_cast<int>(path.size()) == seqs[i]->tree_depth[leaves[i]] + 1, "token tree path / depth mismatch");
      std::vector<int32_t> dst(path.size());
      std::iota(dst.rbegin(), dst.rend(), 0);
      while (!path.empty() && path.back() == dst.back()) {  // leading nodes already sit in their final slots
        path.pop_back();
        dst.pop_back();
      }
      std::reverse(path.begin(), path.end());
      std::reverse(dst.begin(), dst.end());
      for (size_t p = 0; p < path.size(); ++p) {
        commit_src_.push_back(append_pos_[cur_len_indptr_[i] + path[p]]);
        commit_dst_.push_back(append_pos_[cur_len_indptr_[i] + dst[p]]);
      }
      commit_indptr_.push_back(commit_indptr_.back() + static_cast<int32_t>(path.size()));
    }
    CompactKVCopy();
  }
  for (int i = 0; i < n; ++i) {
    const int64_t pop = cur_lens_[i] - (leaves[i] != -1 ? (seqs[i]->tree_depth[leaves[i]] + 1) : 0);
    PopN(cur_seq_ids_[i], static_cast<int32_t>(pop));
    Sequence& s = Seq(cur_seq_ids_[i]);  // PopN may have re-created the map entry
    s.committed = true;
    s.tree_parent.clear();
    s.tree_depth.clear();
  }
}

void Cache::CompactKVCopy() {
  NvtxRange nvtx_range("CompactKVCopy");
  const int total = commit_indptr_.back();
  if (total == 0) return;
  View vi, vsd;
  vi.size = static_cast<int64_t>(commit_indptr_.size());
  vsd.offset = (vi.size + 3) / 4 * 4;
  vsd.size = 2 * total;
  vsd.rows = 2;
  IVec tmp(vsd.offset + vsd.size);
  std::memcpy(tmp.data(), commit_indptr_.data(), vi.size * 4);
  std::memcpy(tmp.data() + vsd.offset, commit_src_.data(), total * 4);
  std::memcpy(tmp.data() + vsd.offset + total, commit_dst_.data(), total * 4);
  for (int64_t l = 0; l < num_layers_; ++l)
    TRACE("compact_copy", {TF({num_total_pages_, 2, num_kv_heads_, page_size_, head_dim_}), TI(vi, tmp.data()),
                           TI(vsd, tmp.data() + vsd.offset), SI(cur_batch_)});
  if (planning_only()) return;
  HCUDA(cudaStreamWaitEvent(copy_stream_, ev_attn_done_, 0));
  HCUDA(cudaStreamSynchronize(copy_stream_));  // the pinned staging buffer may still be in flight from the last commit
  std::memcpy(compact_pinned_, tmp.data(), tmp.size() * 4);
  HCUDA(cudaMemcpyAsync(compact_dev_, compact_pinned_, tmp.size() * 4, cudaMemcpyHostToDevice, copy_stream_));
  for (int64_t l = 0; l < num_layers_; ++l)
    Rc(tvmb200_compact_kv_copy(pages_[l], compact_dev_, compact_dev_ + vsd.offset, static_cast<int32_t>(cur_batch_), total,
                               num_total_pages_, static_cast<int32_t>(num_kv_heads_), static_cast<int32_t>(page_size_),
                               static_cast<int32_t>(head_dim_), dtype_, copy_stream_));
  dirty_ = true;  // the next attention waits for the copy stream (as the reference relies on, paged_kv_cache.cc:767-769)
}

void Cache::DebugGetKV(int64_t seq_id, int64_t start, int64_t end, void* k_out, void* v_out, cudaStream_t st) {
  NvtxRange nvtx_range("vm.builtin.attention_kv_cache_debug_get_kv");
  const Sequence& s = Seq(seq_id);
  HCHECK(start >= 0, "DebugGetKV does not accept negative start_pos %ld", (long)start);
  HCHECK(end <= s.seq_length, "DebugGetKV does not accept out-of-range end_pos");
  HCHECK(start < end, "DebugGetKV does not accept \"start_pos >= end_pos\"");
  IVec pos;
  const int32_t ps = static_cast<int32_t>(page_size_);
  for (int32_t bid : BlockTrace(s)) {
    const Block& b = blocks_[bid];
    for (int i = 0; i < b.seq_length; ++i) {
      const int32_t off = i < b.sink_length ? i : i - b.sink_length + b.sliding_window_offset;
      pos.push_back(b.page_ids[off / ps] * ps + off % ps);
    }
  }
  // the reference dumps layer by layer and stops at the first non-MHA layer (paged_kv_cache.cc:1715-1717, indexed by the
  // local layer id as there); here nothing is dumped in that case
  for (int64_t l = 0; l < num_layers_; ++l)
    HCHECK(attn_kinds_[l] == TVMB200_ATTN_MHA, "Only MHA is supported for DebugGetKV");
  View v;
  v.size = end - start;
  for (int64_t l = 0; l < num_layers_; ++l)
    TRACE("debug_get_kv", {TF({num_total_pages_, 2, num_kv_heads_, page_size_, head_dim_}), TI(v, pos.data() + start),
                           TF({num_layers_, end - start, num_kv_heads_, head_dim_}),
                           TF({num_layers_, end - start, num_kv_heads_, head_dim_}), SI(l)});
  if (planning_only()) return;
  // copies and compactions issued on the copy stream must land first
  HCUDA(cudaEventRecord(ev_copy_, copy_stream_));
  HCUDA(cudaStreamWaitEvent(st, ev_copy_, 0));
  HCUDA(cudaMemcpyAsync(dbg_pos_dev_, pos.data() + start, (end - start) * 4, cudaMemcpyHostToDevice, st));
  HCUDA(cudaStreamSynchronize(st));  // `pos` is pageable host memory
  for (int64_t l = 0; l < num_layers_; ++l)
    Rc(tvmb200_debug_get_kv(pages_[l], dbg_pos_dev_, k_out, v_out, l, num_layers_, end - start, num_total_pages_,
                            static_cast<int32_t>(num_kv_heads_), static_cast<int32_t>(page_size_),
                            static_cast<int32_t>(head_dim_), dtype_, st));
}

}  // namespace host
}  // namespace tvmb200

// --------------------------------------------------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------------------------------------------------
using tvmb200::host::Cache;
struct tvmb200_cache_s {
  std::unique_ptr<Cache> impl;
};

// every entry runs with the cache's own kernel-set context current on the calling thread
struct CacheScope {
  explicit CacheScope(tvmb200_context_t c) : prev(tvmb200_context_enter(c)) {}
  ~CacheScope() { tvmb200_context_enter(prev); }
  tvmb200_context_t prev;
};
#define CACHE_API_BEGIN() try {
#define CACHE_API_BEGIN_C(c)                                                     \
  try {                                                                          \
    if (!(c)) throw std::runtime_error("null cache handle");                     \
    CacheScope scope__((c)->impl->Context());
#define CACHE_API_END()                                   \
  return 0;                                               \
  }                                                       \
  catch (const std::exception& e) {                       \
    return tvmb200::set_error("%s", e.what());            \
  }

extern "C" {
int tvmb200_cache_create(const tvmb200_cache_config* cfg, tvmb200_cache_t* out) {
  CACHE_API_BEGIN();
  if (!cfg || !out) throw std::runtime_error("tvmb200_cache_create: null argument");
  auto* h = new tvmb200_cache_s();
  try {
    h->impl.reset(new Cache(*cfg));
  } catch (...) {
    delete h;
    throw;
  }
  *out = h;
  CACHE_API_END();
}
void tvmb200_cache_destroy(tvmb200_cache_t c) { delete c; }
int tvmb200_cache_clear(tvmb200_cache_t c) { CACHE_API_BEGIN_C(c); c->impl->Clear(); CACHE_API_END(); }
int tvmb200_cache_add_sequence(tvmb200_cache_t c, int64_t s) { CACHE_API_BEGIN_C(c); c->impl->AddSequence(s); CACHE_API_END(); }
int tvmb200_cache_remove_sequence(tvmb200_cache_t c, int64_t s) { CACHE_API_BEGIN_C(c); c->impl->RemoveSequence(s); CACHE_API_END(); }
int tvmb200_cache_fork_sequence(tvmb200_cache_t c, int64_t p, int64_t ch, int64_t pos) {
  CACHE_API_BEGIN_C(c); c->impl->ForkSequence(p, ch, pos); CACHE_API_END();
}
int tvmb200_cache_popn(tvmb200_cache_t c, int64_t s, int32_t n) { CACHE_API_BEGIN_C(c); c->impl->PopN(s, n); CACHE_API_END(); }
int tvmb200_cache_begin_forward(tvmb200_cache_t c, const int64_t* seq_ids, const int64_t* lens, int32_t n,
                                const int64_t* tree, int32_t tree_size) {
  CACHE_API_BEGIN_C(c); c->impl->BeginForward(seq_ids, lens, n, tree, tree_size); CACHE_API_END();
}
int tvmb200_cache_end_forward(tvmb200_cache_t c) { CACHE_API_BEGIN_C(c); c->impl->EndForward(); CACHE_API_END(); }
int tvmb200_cache_enable_kv_transfer(tvmb200_cache_t c, int32_t local_tp_rank, int32_t num_pe, int32_t remote_num_kv_heads) {
  CACHE_API_BEGIN_C(c); c->impl->EnableKVTransfer(local_tp_rank, num_pe, remote_num_kv_heads); CACHE_API_END();
}
int tvmb200_cache_set_remote_pages(tvmb200_cache_t c, int32_t pe, int64_t local_layer, void* peer_mapped_pages) {
  CACHE_API_BEGIN_C(c); c->impl->SetRemotePages(pe, local_layer, peer_mapped_pages); CACHE_API_END();
}
int tvmb200_cache_disagg_prepare_recv(tvmb200_cache_t c, int64_t seq_id, int64_t append_length, int64_t* out, int64_t capacity,
                                      int64_t* out_len) {
  CACHE_API_BEGIN_C(c);
  const std::vector<int64_t> v = c->impl->DisaggPrepareRecv(seq_id, append_length);
  *out_len = static_cast<int64_t>(v.size());
  if (static_cast<int64_t>(v.size()) > capacity) throw std::runtime_error("disagg_prepare_recv: the compressed position map needs " + std::to_string(v.size()) + " entries");
  std::copy(v.begin(), v.end(), out);
  CACHE_API_END();
}
int tvmb200_cache_disagg_mark_send(tvmb200_cache_t c, int64_t seq_id, int64_t begin, const int64_t* compressed_remote_position_map,
                                   int64_t n, int32_t recver_pe_offset) {
  CACHE_API_BEGIN_C(c); c->impl->DisaggMarkSend(seq_id, begin, compressed_remote_position_map, n, recver_pe_offset); CACHE_API_END();
}
int tvmb200_cache_enable_sliding_window_for_seq(tvmb200_cache_t c, int64_t s, int32_t w, int32_t sink) {
  CACHE_API_BEGIN_C(c); c->impl->EnableSlidingWindowForSeq(s, w, sink); CACHE_API_END();
}
int tvmb200_cache_commit_accepted_token_tree_nodes(tvmb200_cache_t c, const int64_t* s, const int64_t* l, int32_t n) {
  CACHE_API_BEGIN_C(c); c->impl->CommitAcceptedTokenTreeNodes(s, l, n); CACHE_API_END();
}
int tvmb200_cache_empty(tvmb200_cache_t c, int32_t* out) { CACHE_API_BEGIN_C(c); *out = c->impl->Empty(); CACHE_API_END(); }
int tvmb200_cache_get_num_available_pages(tvmb200_cache_t c, int32_t* out) {
  CACHE_API_BEGIN_C(c); *out = c->impl->NumAvailablePages(); CACHE_API_END();
}
int tvmb200_cache_get_total_sequence_length(tvmb200_cache_t c, int32_t* out) {
  CACHE_API_BEGIN_C(c); *out = c->impl->TotalSequenceLength(); CACHE_API_END();
}
int tvmb200_cache_get_query_positions(tvmb200_cache_t c, const int32_t** p, int64_t* n, tvmb200_stream_t st) {
  CACHE_API_BEGIN_C(c); c->impl->QueryPositions(p, n, static_cast<cudaStream_t>(st)); CACHE_API_END();
}
int tvmb200_cache_attention_with_fused_qkv(tvmb200_cache_t c, int64_t layer, double sm_scale, const void* qkv, void* o,
                                           int64_t rows, tvmb200_stream_t st) {
  CACHE_API_BEGIN_C(c); c->impl->AttentionWithFusedQKV(layer, sm_scale, qkv, o, rows, static_cast<cudaStream_t>(st)); CACHE_API_END();
}
int tvmb200_cache_self_attention(tvmb200_cache_t c, int64_t layer, double sm_scale, const void* q, const void* k,
                                 const void* v, void* o, float* lse, int64_t rows, tvmb200_stream_t st) {
  CACHE_API_BEGIN_C(c); c->impl->SelfAttention(layer, sm_scale, q, k, v, o, lse, rows, static_cast<cudaStream_t>(st)); CACHE_API_END();
}
int tvmb200_cache_cross_attention(tvmb200_cache_t c, int64_t layer, double sm_scale, const void* q, void* o, float* lse,
                                  int64_t rows, tvmb200_stream_t st) {
  CACHE_API_BEGIN_C(c); c->impl->CrossAttention(layer, sm_scale, q, o, lse, rows, static_cast<cudaStream_t>(st)); CACHE_API_END();
}
int tvmb200_cache_attention_with_shared_kv(tvmb200_cache_t c, int64_t source_layer, double sm_scale, const void* q,
                                           const void* cur_k, const void* cur_v, void* o, int64_t rows, tvmb200_stream_t st) {
  CACHE_API_BEGIN_C(c);
  c->impl->AttentionWithSharedKV(source_layer, sm_scale, q, cur_k, cur_v, o, rows, static_cast<cudaStream_t>(st));
  CACHE_API_END();
}
int tvmb200_cache_merge_attn_output_inplace(tvmb200_cache_t c, void* o_self, float* lse_self, const void* o_cross,
                                            const float* lse_cross, int64_t n, int64_t num_heads, int64_t head_dim,
                                            tvmb200_stream_t st) {
  CACHE_API_BEGIN_C(c);
  c->impl->MergeAttnOutputInplace(o_self, lse_self, o_cross, lse_cross, n, num_heads, head_dim, static_cast<cudaStream_t>(st));
  CACHE_API_END();
}
int tvmb200_cache_debug_get_kv(tvmb200_cache_t c, int64_t s, int64_t a, int64_t b, void* k, void* v, tvmb200_stream_t st) {
  CACHE_API_BEGIN_C(c); c->impl->DebugGetKV(s, a, b, k, v, static_cast<cudaStream_t>(st)); CACHE_API_END();
}
int tvmb200_cache_pages(tvmb200_cache_t c, int64_t layer, void** p, int64_t* np) {
  CACHE_API_BEGIN_C(c); *p = c->impl->Pages(layer, np); CACHE_API_END();
}
int tvmb200_cache_shape(tvmb200_cache_t c, int64_t* out6) { CACHE_API_BEGIN_C(c); c->impl->Shape(out6); CACHE_API_END(); }
int tvmb200_cache_context(tvmb200_cache_t c, tvmb200_context_t* out) { CACHE_API_BEGIN_C(c); *out = c->impl->Context(); CACHE_API_END(); }
int tvmb200_cache_set_trace(tvmb200_cache_t c, int32_t on) { CACHE_API_BEGIN_C(c); c->impl->SetTrace(on != 0); CACHE_API_END(); }
int tvmb200_cache_take_trace(tvmb200_cache_t c, const char** json) { CACHE_API_BEGIN_C(c); *json = c->impl->TakeTrace(); CACHE_API_END(); }
}
