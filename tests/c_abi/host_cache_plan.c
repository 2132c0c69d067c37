/* C99 host of the C ABI (no Python, no C++, no torch): what a cgo / JNI / Rust binding sees.
 * Drives a planning-only cache (device_id = -1: bookkeeping + call trace, nothing is launched, so it runs without a GPU)
 * through add_sequence -> begin_forward -> attention_with_fused_qkv -> fork -> decode, and checks the error convention.
 * Build: gcc -std=c99 -Wall -Wextra -pedantic -Werror -Iinclude tests/c_abi/host_cache_plan.c -Ltvm_b200/lib -ltvm_b200 */
#include <stdio.h>
#include <string.h>

#include "tvm_b200.h"
#include "tvm_b200_cache.h"

#define CHECK(expr)                                                                  \
  do {                                                                               \
    if ((expr) != 0) {                                                               \
      fprintf(stderr, "FAILED %s: %s\n", #expr, tvmb200_last_error());               \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main(void) {
  tvmb200_cache_config cfg;
  tvmb200_cache_t cache = NULL;
  int64_t seqs[2] = {0, 1}, lens[2] = {37, 5}, one[2] = {1, 1};
  int32_t n = 0, empty = 0;
  const char* trace = NULL;

  memset(&cfg, 0, sizeof(cfg));
  cfg.reserved_num_seqs = 4;
  cfg.total_token_capacity = 256;
  cfg.prefill_chunk_size = 128;
  cfg.page_size = 16;
  cfg.num_layers = 1;
  cfg.num_qo_heads = 32;
  cfg.num_kv_heads = 8;
  cfg.head_dim = 128;
  cfg.rope_mode = TVMB200_ROPE_NORMAL;
  cfg.rotary_scale = 1.0;
  cfg.rotary_theta = 1e4;
  cfg.dtype = TVMB200_F16;
  cfg.device_id = -1;
  CHECK(tvmb200_cache_create(&cfg, &cache));
  CHECK(tvmb200_cache_empty(cache, &empty));
  if (!empty) return 2;
  CHECK(tvmb200_cache_add_sequence(cache, 0));
  CHECK(tvmb200_cache_add_sequence(cache, 1));
  CHECK(tvmb200_cache_set_trace(cache, 1));
  CHECK(tvmb200_cache_begin_forward(cache, seqs, lens, 2, NULL, 0));
  CHECK(tvmb200_cache_attention_with_fused_qkv(cache, 0, 0.088388, NULL, NULL, 42, NULL));
  CHECK(tvmb200_cache_end_forward(cache));
  CHECK(tvmb200_cache_take_trace(cache, &trace));
  if (strstr(trace, "\"fn\":\"split_rotary\"") == NULL || strstr(trace, "\"fn\":\"prefill_ragged\"") == NULL ||
      strstr(trace, "\"fn\":\"transpose_append\"") == NULL) {
    fprintf(stderr, "unexpected plan: %s\n", trace);
    return 3;
  }
  CHECK(tvmb200_cache_get_total_sequence_length(cache, &n));
  if (n != 42) return 4;
  CHECK(tvmb200_cache_get_num_available_pages(cache, &n));
  if (n != 17 - 3 - 1) return 5; /* 256 / 16 + 1 pages; 37 tokens take 3, 5 tokens take 1 */
  CHECK(tvmb200_cache_fork_sequence(cache, 0, 2, 20));
  seqs[1] = 2;
  CHECK(tvmb200_cache_begin_forward(cache, seqs, one, 2, NULL, 0));
  CHECK(tvmb200_cache_attention_with_fused_qkv(cache, 0, 0.088388, NULL, NULL, 2, NULL));
  CHECK(tvmb200_cache_end_forward(cache));
  CHECK(tvmb200_cache_take_trace(cache, &trace));
  if (strstr(trace, "\"fn\":\"decode\"") == NULL) {
    fprintf(stderr, "unexpected decode plan: %s\n", trace);
    return 6;
  }
  /* error convention: non-zero return, message in tvmb200_last_error(), nothing launched, the cache stays usable */
  if (tvmb200_cache_add_sequence(cache, 0) == 0) return 7;
  if (strstr(tvmb200_last_error(), "already in the KV cache") == NULL) return 8;
  if (tvmb200_cache_remove_sequence(cache, 9) == 0) return 9;
  if (tvmb200_set_rope_scaling(99, 1.f, 1.f, 4.f, 8192.f) == 0) return 10;
  CHECK(tvmb200_cache_remove_sequence(cache, 2));
  CHECK(tvmb200_cache_clear(cache));
  CHECK(tvmb200_cache_empty(cache, &empty));
  if (!empty) return 11;
  tvmb200_cache_destroy(cache);
  printf("c abi ok: %s\n", tvmb200_version());
  return 0;
}
