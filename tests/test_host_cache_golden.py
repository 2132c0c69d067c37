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


def _plan_cache(**kw):
    from tvm_b200.kv_cache import PagedKVCache

    cfg = dict(reserved_num_seqs=4, total_token_capacity=1024, prefill_chunk_size=512, num_layers=2, num_qo_heads=4,
               num_kv_heads=1, head_dim=128, device=None)
    cfg.update(kw)
    return PagedKVCache(**cfg)


def test_limits_and_preconditions_like_the_reference(built_lib):
    """The reference's ICHECKs on the hot path (paged_kv_cache.cc:606-640, 772-830, 1566-1590, 1800-1870, 1404-1485),
    with its messages: maximum tree size 256, tree shape, fork / popn / sliding-window preconditions, the split entries."""
    from tvm_b200 import capi

    E = capi.TvmB200Error
    c = _plan_cache()
    c.add_sequence(0)
    c.begin_forward([0], [20])
    c.attention_with_fused_qkv(0, 1.0, None, None)
    c.end_forward()
    # kTreeAttnMaxTreeSize = 256 (attn_utils.h:52): exactly 256 nodes pass, 257 do not
    chain = list(range(-1, 255))
    c.begin_forward([0], [256], chain)
    c.commit_accepted_token_tree_nodes([0], [-1])
    assert c.get_total_sequence_length() == 20
    with pytest.raises(E, match="exceeds the maximum tree size limit 256"):
        c.begin_forward([0], [257], list(range(-1, 256)))
    c = _plan_cache()
    c.add_sequence(0)
    c.begin_forward([0], [20])
    c.end_forward()
    with pytest.raises(E, match="not smaller than"):
        c.begin_forward([0], [3], [-1, 1, 0])
    # a begin_forward that failed half-way leaves no batch: every consumer refuses it (this sequence used to crash)
    with pytest.raises(E, match="did not complete"):
        c.commit_accepted_token_tree_nodes([0], [0])
    with pytest.raises(E, match="did not complete"):
        c.attention_with_fused_qkv(0, 1.0, None, None)
    c = _plan_cache()
    c.add_sequence(0)
    c.add_sequence(1)
    c.begin_forward([0, 1], [20, 3])
    c.end_forward()
    with pytest.raises(E, match="already committed"):
        c.commit_accepted_token_tree_nodes([0], [0])
    with pytest.raises(E, match="is not sequence 0 of the last begin_forward"):
        c.commit_accepted_token_tree_nodes([1], [-1])
    with pytest.raises(E, match="3 sequences, but the last begin_forward had 2"):
        c.commit_accepted_token_tree_nodes([0, 1, 1], [-1, -1, -1])
    # a token tree that was not committed blocks the fork of its sequence
    c2 = _plan_cache()
    c2.add_sequence(0)
    c2.begin_forward([0], [4], [-1, 0, 0, 1])
    with pytest.raises(E, match="has not been committed|not been committed"):
        c2.fork_sequence(0, 1, 2)
    with pytest.raises(E, match="larger than or equals to the append length|Invalid tree index"):
        c2.commit_accepted_token_tree_nodes([0], [4])
    c2.commit_accepted_token_tree_nodes([0], [3])
    assert c2.get_total_sequence_length() == 3          # the path 0 -> 1 -> 3
    # fork / popn preconditions
    with pytest.raises(E, match="should not exceed the total length of parent"):
        c2.fork_sequence(0, 1, 9)
    with pytest.raises(E, match="non-negative, or -1"):
        c2.fork_sequence(0, 1, -2)
    with pytest.raises(E, match="already in the KV cache"):
        c2.fork_sequence(0, 0, 1)
    with pytest.raises(E, match="cannot be negative"):
        c2.popn(0, -1)
    with pytest.raises(E):
        c2.popn(0, 4)
    with pytest.raises(E, match="Append with length 0 is not allowed"):
        c2.begin_forward([0], [0])
    with pytest.raises(E, match="more than prefill_chunk_size"):
        c2.begin_forward([0], [513])
    # sliding window
    s = _plan_cache(support_sliding_window=True, rope_mode=2)
    s.add_sequence(0)
    with pytest.raises(E, match="should be less than the sliding window size"):
        s.enable_sliding_window_for_seq(0, 8, 8)
    with pytest.raises(E, match="should be positive"):
        s.enable_sliding_window_for_seq(0, 0, 0)
    with pytest.raises(E, match="non negative"):
        s.enable_sliding_window_for_seq(0, 8, -1)
    s.enable_sliding_window_for_seq(0, 24, 4)
    with pytest.raises(E, match="cannot be enabled twice"):
        s.enable_sliding_window_for_seq(0, 24, 4)
    s.begin_forward([0], [40])
    s.end_forward()
    assert s.get_total_sequence_length() == 24          # the window slid
    with pytest.raises(E, match="only can be forked within sink size|within sink size"):
        s.fork_sequence(0, 1, 10)
    with pytest.raises(E, match="Tree attention does not support sliding window"):
        s.begin_forward([0], [2], [-1, 0])
    # the split entries: rows must be the batch's append length, MHA layers only, layer ids inside the cache
    m = _plan_cache(attn_kinds=[3, 0], layer_sliding_window_size=16, rope_mode=0)
    m.add_sequence(0)
    m.begin_forward([0], [5])
    with pytest.raises(E, match="5 tokens|appends 5"):
        m.self_attention(1, 1.0, 4, None, None, None, None)
    with pytest.raises(E, match="not an MHA layer"):
        m.cross_attention(0, 1.0, 5, None, None)
    with pytest.raises(E, match="outside this cache's layers"):
        m.attention_with_shared_kv(2, 1.0, 5, None, None, None)
    m.self_attention(1, 1.0, 5, None, None, None, None)
    m.end_forward()


def test_batch_beyond_reserved_num_seqs_is_refused_not_overrun(built_lib):
    """The merged aux buffer is sized for reserved_num_seqs sequences (as the reference's is, attn_utils.h:817-1052): a batch
    that needs more is a loud error, checked before every write into the staging buffer, and -- begin_forward being
    transactional -- the refused batch leaves lengths and free pages untouched."""
    from tvm_b200 import capi

    N = 1200
    for sliding in (False, True):
        c = _plan_cache(reserved_num_seqs=1, total_token_capacity=16 * N, prefill_chunk_size=N, num_layers=1,
                        support_sliding_window=sliding, rope_mode=2)
        for i in range(N):
            c.add_sequence(i)
        free0 = c.get_num_available_pages()
        with pytest.raises(capi.TvmB200Error, match="auxiliary buffer overflow"):
            c.begin_forward(list(range(N)), [1] * N)
        assert c.get_total_sequence_length() == 0 and c.get_num_available_pages() == free0
        with pytest.raises(capi.TvmB200Error, match="did not complete"):
            c.attention_with_fused_qkv(0, 1.0, None, None)
        c.begin_forward([0], [1])           # the cache stays usable
        c.attention_with_fused_qkv(0, 1.0, None, None)
        c.end_forward()
    ok_ = _plan_cache(reserved_num_seqs=N, total_token_capacity=16 * N, prefill_chunk_size=N, num_layers=1)
    for i in range(N):
        ok_.add_sequence(i)
    ok_.begin_forward(list(range(N)), [1] * N)
    ok_.attention_with_fused_qkv(0, 1.0, None, None)
    ok_.end_forward()


def test_full_prefill_chunk_is_accepted(built_lib):
    """a batch with total_append == prefill_chunk_size must fit the merged aux buffer (the two
    kv-transfer regions BuildAuxViews skips are budgeted too)."""
    for chunk in (512, 8192):
        c = _plan_cache(reserved_num_seqs=4, total_token_capacity=32768, prefill_chunk_size=chunk, num_layers=1)
        c.add_sequence(0)
        c.begin_forward([0], [chunk])
        c.attention_with_fused_qkv(0, 1.0, None, None)
        c.end_forward()
        assert c.get_total_sequence_length() == chunk
        c.add_sequence(1)
        c.add_sequence(2)
        c.begin_forward([1, 2], [chunk // 2, chunk - chunk // 2])
        c.attention_with_fused_qkv(0, 1.0, None, None)
        c.end_forward()


def test_failed_begin_forward_rolls_back(built_lib):
    """a begin_forward that throws leaves sequence lengths, block lengths and the free-page stack
    exactly as they were; popn / remove / fork afterwards behave as if the call had never happened."""
    from tvm_b200 import capi

    E = capi.TvmB200Error
    for sliding in (False, True):
        c = _plan_cache(reserved_num_seqs=4, total_token_capacity=64, prefill_chunk_size=64, num_layers=1,
                        support_sliding_window=sliding, rope_mode=2 if sliding else 1)
        twin = _plan_cache(reserved_num_seqs=4, total_token_capacity=64, prefill_chunk_size=64, num_layers=1,
                           support_sliding_window=sliding, rope_mode=2 if sliding else 1)
        for m in (c, twin):
            m.add_sequence(0)
            m.add_sequence(1)
            m.begin_forward([0, 1], [20, 30])
            m.attention_with_fused_qkv(0, 1.0, None, None)
            m.end_forward()
        state0 = (c.get_total_sequence_length(), c.get_num_available_pages())
        with pytest.raises(E, match="cannot be found"):      # seq 0 would have been lengthened before seq 7 is looked up
            c.begin_forward([0, 7], [3, 3])
        assert (c.get_total_sequence_length(), c.get_num_available_pages()) == state0
        with pytest.raises(E, match="full|prefill_chunk_size|No page"):  # seq 0 gets pages, seq 1 cannot
            c.begin_forward([0, 1], [10, 60])
        assert (c.get_total_sequence_length(), c.get_num_available_pages()) == state0
        with pytest.raises(E, match="empty"):
            c.begin_forward([], [])
        with pytest.raises(E, match="Invalid token tree|does not support"):
            c.begin_forward([0, 1], [2, 2], [-1, 0, -1, 5])
        assert (c.get_total_sequence_length(), c.get_num_available_pages()) == state0
        # from here on both caches must plan identically
        for m in (c, twin):
            m.set_trace(True)
            m.fork_sequence(0, 2, 17)
            m.begin_forward([0, 2, 1], [1, 2, 1])
            m.attention_with_fused_qkv(0, 1.0, None, None)
            m.end_forward()
            m.popn(1, 5)
            m.remove_sequence(0)
            m.begin_forward([2, 1], [1, 1])
            m.attention_with_fused_qkv(0, 1.0, None, None)
            m.end_forward()
        assert c.take_trace() == twin.take_trace()
        assert c.get_num_available_pages() == twin.get_num_available_pages()


def test_disaggregation_bookkeeping_planning_only(built_lib):
    """DisaggPrepareRecv / DisaggMarkSend (paged_kv_cache.cc:1220-1301) on planning-only caches: the compressed position map
    is [n, begin_1, length_1, ...] over the receiver's append slots, the sender's next forward carries one transfer entry
    with the fresh rows (and, when the sequence was marked after a partial prefill, the cached rows page to page), the
    un-marked sequences send nothing, and the fused decode step is not taken while a transfer is due."""
    from tvm_b200 import capi

    recv = _plan_cache(reserved_num_seqs=4, total_token_capacity=256, prefill_chunk_size=128, num_layers=2)
    recv.add_sequence(5)
    recv.add_sequence(6)
    recv.begin_forward([6], [3])                 # seq 6 takes the first page, so seq 5's slots do not start at 0
    recv.attention_with_fused_qkv(0, 1.0, None, None)
    recv.end_forward()
    m = recv.disagg_prepare_recv(5, 40)
    assert m[0] * 2 + 1 == len(m) and sum(m[2::2]) == 40
    slots = [b + i for b, n in zip(m[1::2], m[2::2]) for i in range(n)]
    assert len(set(slots)) == 40 and min(slots) >= 16
    assert recv.get_total_sequence_length() == 43
    recv.end_forward()

    send = _plan_cache(reserved_num_seqs=4, total_token_capacity=256, prefill_chunk_size=128, num_layers=2)
    send.add_sequence(5)
    send.add_sequence(9)
    with pytest.raises(capi.TvmB200Error, match="enable_kv_transfer|set up"):
        send.disagg_mark_send(5, 0, m, 1)
    send.enable_kv_transfer(local_tp_rank=0, num_pe=2)
    with pytest.raises(capi.TvmB200Error, match="malformed"):
        send.disagg_mark_send(5, 0, m[:-1], 1)
    with pytest.raises(capi.TvmB200Error, match="cannot be found"):
        send.disagg_mark_send(77, 0, m, 1)
    send.disagg_mark_send(5, 0, m, 1)
    send.set_trace(True)
    send.begin_forward([5, 9], [25, 7])          # seq 9 is not marked: its rows stay here
    for layer in range(2):
        send.attention_with_fused_qkv(layer, 1.0, None, None)
    send.end_forward()
    tr = send.take_trace()
    names = [c["fn"] if isinstance(c, dict) else c[0] for c in tr]
    assert names.count("kv_transfer") == 2       # once per layer
    send.begin_forward([5], [15])                # the rest of the prefill
    send.attention_with_fused_qkv(0, 1.0, None, None)
    send.attention_with_fused_qkv(1, 1.0, None, None)
    send.end_forward()
    with pytest.raises(capi.TvmB200Error, match="more tokens than the receiver prepared"):
        send.begin_forward([5], [1])             # 41st token: the receiver reserved 40
    # marking a sequence that already holds tokens: they go page to page with the next forward
    send.disagg_mark_send(9, 2, [1, 100, 20], 1)
    send.take_trace()
    send.begin_forward([9], [4])
    send.attention_with_fused_qkv(0, 1.0, None, None)
    send.attention_with_fused_qkv(1, 1.0, None, None)
    send.end_forward()
    tr = send.take_trace()
    kv = [c for c in tr if (c["fn"] if isinstance(c, dict) else c[0]) == "kv_transfer"]
    assert len(kv) == 2
