// KV transfer between caches on different GPUs (disaggregated prefill -> decode), over NVLink peer memory.
//
// Reference: src/runtime/extra/contrib/nvshmem/kv_transfer.cu -- KVTransfer (:38-83, k / v rows of freshly computed tokens
// pushed into the remote page pool with nvshmemx_putmem_nbi_warp + nvshmem_quiet) and KVTransferPageToPage (:84-130, rows
// already cached locally), registered as nvshmem.KVTransfer / nvshmem.KVTransferPageToPage (:327) and called by the cache
// from AttentionWithFusedQKV (paged_kv_cache.cc:1374-1394).
//
// There is no NVSHMEM here and none is needed inside one NVLink domain: the "symmetric heap" is a table of peer-mapped
// device pointers, one per processing element (CUDA IPC handles, torch symmetric memory or plain peer access supply
// them), and a put is a 16-byte store through such a pointer.  One warp moves one (token, kv head) pair -- K row and
// V row, 2 x head_dim x 2 bytes -- as 16-byte vectors; a system-scope fence before the kernel ends plays nvshmem_quiet.
// The head mapping between a sender with `local_num_kv_heads` per rank and a receiver with `remote_num_kv_heads` per
// rank (gather when the receiver's shards are wider, scatter when they are narrower) is the reference's, line by line
// in meaning: kv_transfer.cu:54-66.
#include "common.cuh"

namespace tvmb200 {

namespace {

constexpr int kMaxPe = 64;
struct PeTable {
  void* pages[kMaxPe];  // page pool [remote_num_pages, 2, remote_num_kv_heads, page_size, head_dim] of every PE
};

struct TransferParams {
  const void* k;                          // [ntokens, local_num_kv_heads, head_dim]        (fresh rows)
  const void* v;
  const void* local_pages;                // [*, 2, local_num_kv_heads, page_size, head_dim] (page-to-page)
  const int32_t* remote_position_map;     // [ntokens] slot in the remote pool, or -1
  const int32_t* local_position_map;      // [ntokens] slot in the local pool, or -1         (page-to-page)
  const int32_t* remote_tp_group_pe_offset;  // [ntokens] first PE of the receiving TP group
  int64_t ntokens;
  int local_num_kv_heads, remote_num_kv_heads, page_size, row_vecs, local_tp_rank, num_pe;
};

template <bool PAGE_TO_PAGE>
__global__ void __launch_bounds__(256)
kv_transfer_kernel(const TransferParams a, const PeTable pe) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t items = a.ntokens * a.local_num_kv_heads;
  for (int64_t it = warp; it < items; it += nwarps) {
    const int64_t tok = it / a.local_num_kv_heads;
    const int h = static_cast<int>(it - tok * a.local_num_kv_heads);
    const int rpos = a.remote_position_map[tok];
    if (rpos < 0) continue;
    int lpos = 0;
    if (PAGE_TO_PAGE) {
      lpos = a.local_position_map[tok];
      if (lpos < 0) continue;
    }
    int remote_pe, remote_h;
    if (a.local_num_kv_heads <= a.remote_num_kv_heads) {  // gather: several sender ranks fill one receiver rank
      const int gather = a.remote_num_kv_heads / a.local_num_kv_heads;
      remote_pe = a.remote_tp_group_pe_offset[tok] + a.local_tp_rank / gather;
      remote_h = (a.local_tp_rank % gather) * a.local_num_kv_heads + h;
    } else {                                              // scatter: one sender rank feeds several receiver ranks
      const int scatter = a.local_num_kv_heads / a.remote_num_kv_heads;
      remote_pe = a.remote_tp_group_pe_offset[tok] + a.local_tp_rank * scatter + h / a.remote_num_kv_heads;
      remote_h = h % a.remote_num_kv_heads;
    }
    if (remote_pe < 0 || remote_pe >= a.num_pe || pe.pages[remote_pe] == nullptr) __trap();  // a table the caller never filled
    const int64_t rpage = rpos / a.page_size, roff = rpos - rpage * a.page_size;
    uint4* dst = static_cast<uint4*>(pe.pages[remote_pe]);
    const int64_t dst_k = (((rpage * 2 + 0) * a.remote_num_kv_heads + remote_h) * a.page_size + roff) * a.row_vecs;
    const int64_t dst_v = (((rpage * 2 + 1) * a.remote_num_kv_heads + remote_h) * a.page_size + roff) * a.row_vecs;
    const uint4 *src_k, *src_v;
    if (PAGE_TO_PAGE) {
      const int64_t lpage = lpos / a.page_size, loff = lpos - lpage * a.page_size;
      const uint4* lp = static_cast<const uint4*>(a.local_pages);
      src_k = lp + (((lpage * 2 + 0) * a.local_num_kv_heads + h) * a.page_size + loff) * a.row_vecs;
      src_v = lp + (((lpage * 2 + 1) * a.local_num_kv_heads + h) * a.page_size + loff) * a.row_vecs;
    } else {
      src_k = static_cast<const uint4*>(a.k) + (tok * a.local_num_kv_heads + h) * a.row_vecs;
      src_v = static_cast<const uint4*>(a.v) + (tok * a.local_num_kv_heads + h) * a.row_vecs;
    }
    // lanes [0, row_vecs) carry K, lanes [16, 16 + row_vecs) carry V (head_dim <= 128: row_vecs <= 16)
    const int j = lane & 15;
    if (j < a.row_vecs) {
      if (lane < 16) dst[dst_k + j] = ldg_nc_v4(src_k + j);
      else dst[dst_v + j] = ldg_nc_v4(src_v + j);
    }
  }
  __threadfence_system();  // every put of this thread is visible to the receiving GPU before the kernel completes
}

int launch(const TransferParams& a, void* const* remote_pages, bool p2p, cudaStream_t st) {
  PeTable pe = {};
  for (int i = 0; i < a.num_pe; ++i) pe.pages[i] = remote_pages[i];
  const int64_t items = a.ntokens * a.local_num_kv_heads;
  const int64_t want = (items + 7) / 8, cap = static_cast<int64_t>(num_sms()) * 4;
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  if (p2p) kv_transfer_kernel<true><<<grid, 256, 0, st>>>(a, pe);
  else kv_transfer_kernel<false><<<grid, 256, 0, st>>>(a, pe);
  TVMB200_LAUNCH_OK();
  return 0;
}

int check(const char* who, void* const* remote_pages, int64_t ntokens, int32_t local_h, int32_t remote_h, int32_t page_size,
          int32_t head_dim, int32_t local_tp_rank, int32_t num_pe, int dtype) {
  TVMB200_CHECK(dtype == TVMB200_F16 || dtype == TVMB200_BF16, "%s: unsupported dtype %d", who, dtype);
  TVMB200_CHECK(remote_pages != nullptr && num_pe >= 1 && num_pe <= kMaxPe, "%s: %d processing elements (1..%d) / null table", who, num_pe, kMaxPe);
  TVMB200_CHECK(ntokens >= 0 && page_size > 0 && local_tp_rank >= 0, "%s: bad sizes", who);
  TVMB200_CHECK(head_dim % 8 == 0 && head_dim >= 8 && head_dim <= 128, "%s: head_dim %d unsupported (multiple of 8, <= 128)", who, head_dim);
  TVMB200_CHECK(local_h > 0 && remote_h > 0 && (local_h % remote_h == 0 || remote_h % local_h == 0),
                "%s: %d local and %d remote kv heads per rank do not divide each other", who, local_h, remote_h);
  return 0;
}

}  // namespace
}  // namespace tvmb200

using namespace tvmb200;

extern "C" int tvmb200_kv_transfer(void* const* remote_pages, const void* k, const void* v,
                                   const int32_t* remote_position_map, const int32_t* remote_tp_group_pe_offset,
                                   int64_t ntokens, int32_t local_num_kv_heads, int32_t remote_num_kv_heads,
                                   int32_t page_size, int32_t head_dim, int32_t local_tp_rank, int32_t num_pe, int dtype,
                                   tvmb200_stream_t stream) {
  if (int rc = check("kv_transfer", remote_pages, ntokens, local_num_kv_heads, remote_num_kv_heads, page_size, head_dim,
                     local_tp_rank, num_pe, dtype)) return rc;
  if (ntokens == 0) return 0;
  TVMB200_CHECK(k != nullptr && v != nullptr && remote_position_map != nullptr && remote_tp_group_pe_offset != nullptr, "kv_transfer: null argument");
  TransferParams a = {};
  a.k = k;
  a.v = v;
  a.remote_position_map = remote_position_map;
  a.remote_tp_group_pe_offset = remote_tp_group_pe_offset;
  a.ntokens = ntokens;
  a.local_num_kv_heads = local_num_kv_heads;
  a.remote_num_kv_heads = remote_num_kv_heads;
  a.page_size = page_size;
  a.row_vecs = head_dim / 8;
  a.local_tp_rank = local_tp_rank;
  a.num_pe = num_pe;
  return launch(a, remote_pages, false, static_cast<cudaStream_t>(stream));
}

extern "C" int tvmb200_kv_transfer_page_to_page(void* const* remote_pages, const void* local_pages,
                                                const int32_t* remote_position_map, const int32_t* local_position_map,
                                                const int32_t* remote_tp_group_pe_offset, int64_t ntokens,
                                                int32_t local_num_kv_heads, int32_t remote_num_kv_heads, int32_t page_size,
                                                int32_t head_dim, int32_t local_tp_rank, int32_t num_pe, int dtype,
                                                tvmb200_stream_t stream) {
  if (int rc = check("kv_transfer_page_to_page", remote_pages, ntokens, local_num_kv_heads, remote_num_kv_heads, page_size,
                     head_dim, local_tp_rank, num_pe, dtype)) return rc;
  if (ntokens == 0) return 0;
  TVMB200_CHECK(local_pages != nullptr && remote_position_map != nullptr && local_position_map != nullptr &&
                remote_tp_group_pe_offset != nullptr, "kv_transfer_page_to_page: null argument");
  TransferParams a = {};
  a.local_pages = local_pages;
  a.remote_position_map = remote_position_map;
  a.local_position_map = local_position_map;
  a.remote_tp_group_pe_offset = remote_tp_group_pe_offset;
  a.ntokens = ntokens;
  a.local_num_kv_heads = local_num_kv_heads;
  a.remote_num_kv_heads = remote_num_kv_heads;
  a.page_size = page_size;
  a.row_vecs = head_dim / 8;
  a.local_tp_rank = local_tp_rank;
  a.num_pe = num_pe;
  return launch(a, remote_pages, true, static_cast<cudaStream_t>(stream));
}

// peer access between two devices of one process (tests and single-process multi-GPU hosts; separate processes map each
// other's pools through CUDA IPC / symmetric memory instead)
extern "C" int tvmb200_enable_peer_access(int32_t device, int32_t peer) {
  int can = 0;
  TVMB200_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
  TVMB200_CHECK(can, "device %d cannot access device %d as a peer", device, peer);
  int prev = 0;
  TVMB200_CUDA(cudaGetDevice(&prev));
  TVMB200_CUDA(cudaSetDevice(device));
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  cudaSetDevice(prev);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  TVMB200_CUDA(e);
  return 0;
}
