"""Host cache vs the REAL reference (no GPU needed): every scenario program captured from the reference's C++
PagedAttentionKVCacheObj is replayed on a planning-only cache, and the callback sequence with every int32 auxiliary
array (page_indptr, page_values, length_info, rope offsets, append / compaction position maps, tree masks) and every
scalar must be bit-identical."""
import pytest

from tests.golden_replay import replay, scenario_names


@pytest.mark.parametrize("name", scenario_names())
def test_aux_arrays_bit_exact(built_lib, name):
    replay(name, device=None)


def test_error_behaviour(built_lib):
    from tvm_b200 import capi
    from tvm_b200.kv_cache import PagedKVCache

    c = PagedKVCache(reserved_num_seqs=4, total_token_capacity=64, prefill_chunk_size=128, num_layers=1, num_qo_heads=4,
                     num_kv_heads=1, head_dim=128, device=None)
    c.add_sequence(0)
    with pytest.raises(capi.TvmB200Error, match="already in the KV cache"):
        c.add_sequence(0)
    with pytest.raises(capi.TvmB200Error, match="cannot be found"):
        c.remove_sequence(7)
    with pytest.raises(capi.TvmB200Error, match="does not support sliding window"):
        c.enable_sliding_window_for_seq(0, 8, 2)
    c.begin_forward([0], [80 - 16])  # capacity 64 tokens + 1 spare page = 5 pages
    with pytest.raises(capi.TvmB200Error, match="KV cache is full"):
        c.begin_forward([0], [32])
    with pytest.raises(capi.TvmB200Error, match="Invalid token tree"):
        c2 = PagedKVCache(reserved_num_seqs=4, total_token_capacity=64, prefill_chunk_size=128, num_layers=1,
                          num_qo_heads=4, num_kv_heads=1, head_dim=128, device=None)
        c2.add_sequence(0)
        c2.begin_forward([0], [3], [-1, 2, 0])
