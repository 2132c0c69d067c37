"""GPU parity of the tcgen05 prefill kernel on the flavours that used to run on the mma.sync kernel only: token-tree masks
(ragged a6 / paged a7, in-kernel ancestor test on the trailing tree columns), the per-layer sliding-window mask, and --
through the gather / rotate pre-pass (prefill_prepass.cu) -- inline RoPE (a9) and the `_sliding_window` flavours with
their [3, B] length_info (a11).  The kernel is forced on (auto dispatch only takes it from 2048 folded rows); every case
also checks the launch went down the tensor-core path it names: no silent fallback."""
import numpy as np
import pytest

from tests.test_kernels_gpu import _dfs_mask, _run_paged_prefill, _run_ragged

pytestmark = pytest.mark.gpu
DTYPES = ["float16", "bfloat16"]
GENERIC, TC05, PREPASS = 0, 1, 2


@pytest.fixture()
def tc05(built_lib):
    from tvm_b200 import capi

    capi.lib()
    capi.set_prefill_impl(2)
    yield capi
    capi.set_prefill_impl(0)


class took:
    """with took(capi, PREPASS): ...  -- exactly the launches inside went down that path, none down the mma.sync kernel"""

    def __init__(self, capi, path, n=1):
        self.capi, self.path, self.n = capi, path, n

    def __enter__(self):
        self.before = self.capi.prefill_path_counts()

    def __exit__(self, et, ev, tb):
        if et is None:
            after = self.capi.prefill_path_counts()
            d = [a - b for a, b in zip(after, self.before)]
            assert d[GENERIC] == 0 and d[self.path] == self.n, f"prefill paths taken (generic, tcgen05, pre-pass): {d}"


def _random_tree(rng, n):
    return [-1] + [int(rng.integers(-1 if k > 3 else 0, k)) for k in range(1, n)]


def _tree_arrays(trees):
    ind = np.zeros(len(trees) + 1, np.int32)
    ind[1:] = np.cumsum([len(t) for t in trees])
    return ind, np.concatenate([_dfs_mask(t) for t in trees])


@pytest.mark.parametrize("dtype", DTYPES)
def test_tc05_tree_ragged(tc05, dtype):
    rng = np.random.default_rng(60)
    trees = [[-1, 0, 0, 1], [(k - 1) // 2 if k else -1 for k in range(64)], [-1, 0, 1, 2, 3, 4, 5], _random_tree(rng, 200),
             [-1] + [0] * 129, _random_tree(rng, 257)]
    lens = [len(t) for t in trees]
    with took(tc05, TC05):
        _run_ragged(tc05, rng, lens, lens, 32, 8, 128, dtype, tree=_tree_arrays(trees))


@pytest.mark.parametrize("dtype", DTYPES)
def test_tc05_tree_ragged_inline_rope(tc05, dtype):
    """tree + rotary_mode 1: K rows are rotated at q_rope_position[kv row] (tree_attn.py:429) by the pre-pass"""
    rng = np.random.default_rng(61)
    trees = [_random_tree(rng, 40), [(k - 1) // 2 if k else -1 for k in range(31)], _random_tree(rng, 130)]
    lens = [len(t) for t in trees]
    with took(tc05, PREPASS):
        _run_ragged(tc05, rng, lens, lens, 32, 8, 128, dtype, rotary_mode=1, tree=_tree_arrays(trees))


@pytest.mark.parametrize("dtype", DTYPES)
def test_tc05_tree_paged(tc05, dtype):
    """the tree occupies the trailing columns of the cached KV (tree region crossing 64- and 128-column tile borders)"""
    rng = np.random.default_rng(62)
    trees = [[-1, 0, 0, 1], [(k - 1) // 2 if k else -1 for k in range(15)], _random_tree(rng, 64), _random_tree(rng, 150),
             _random_tree(rng, 9)]
    sizes = [len(t) for t in trees]
    with took(tc05, TC05):
        _run_paged_prefill(tc05, rng, sizes, [4 + 20, 15 + 200, 64 + 90, 150 + 1000, 9], 32, 8, 128, dtype,
                           tree=_tree_arrays(trees))


@pytest.mark.parametrize("dtype", DTYPES)
def test_tc05_tree_paged_second_round_rows(tc05, dtype):
    """fewer query rows than tree nodes (a later round of the same tree: rows are the LAST nodes, tree_attn.py:57)"""
    rng = np.random.default_rng(63)
    trees = [_random_tree(rng, 30), _random_tree(rng, 70)]
    with took(tc05, TC05):
        _run_paged_prefill(tc05, rng, [12, 33], [30 + 50, 70 + 300], 32, 8, 128, dtype, tree=_tree_arrays(trees))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("causal", [0, 1])
def test_tc05_paged_inline_rope(tc05, dtype, causal):
    rng = np.random.default_rng(64)
    with took(tc05, PREPASS):
        _run_paged_prefill(tc05, rng, [3, 64, 5, 200, 1], [16, 300, 77, 1000, 1], 32, 8, 128, dtype, causal=causal,
                           rotary_mode=1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rotary_mode", [1, 0])
def test_tc05_paged_sliding_window_slots(tc05, dtype, rotary_mode):
    """[3, B] length_info: attention sinks + a window that has slid (position -> slot remap), non-causal as the cache
    issues it (the new tokens are appended after the attention)"""
    rng = np.random.default_rng(65)
    with took(tc05, PREPASS):
        _run_paged_prefill(tc05, rng, [4, 9, 2, 70], [100, 64, 33, 400], 32, 8, 128, dtype, causal=0,
                           rotary_mode=rotary_mode, sliding=[(37, 4), (16, 16), (0, 0), (130, 7)])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sws", [3, 40, 1024])
def test_tc05_paged_layer_sliding_window_mask(tc05, dtype, sws):
    """causal > 0 on the `_sliding_window` flavour = the per-layer window mask (_kernel_common.py:138-144): a LOWER
    bound per row; rows farther than the window from every key are fully masked (O = 0, lse = -5e4)"""
    rng = np.random.default_rng(66)
    with took(tc05, PREPASS):
        _run_paged_prefill(tc05, rng, [2, 30, 7], [50, 200, 90], 32, 8, 128, dtype, causal=1, rotary_mode=1,
                           sliding=[(0, 0), (0, 0), (20, 3)], layer_sws=sws)


@pytest.mark.parametrize("dtype", DTYPES)
def test_tc05_ragged_inline_rope(tc05, dtype):
    rng = np.random.default_rng(67)
    with took(tc05, PREPASS):
        _run_ragged(tc05, rng, [10, 70, 300], [10, 70, 300], 32, 8, 128, dtype, rotary_mode=1)


def test_tc05_prepass_cap_falls_back_loudly_countable(tc05):
    """a scratch cap below the need keeps the call on the mma.sync kernel -- visible in the path counters"""
    rng = np.random.default_rng(68)
    tc05.set_prefill_prepass_cap(1024)
    try:
        before = tc05.prefill_path_counts()
        _run_ragged(tc05, rng, [10, 70], [10, 70], 32, 8, 128, "float16", rotary_mode=1)
        after = tc05.prefill_path_counts()
        assert after[GENERIC] - before[GENERIC] == 1 and after[PREPASS] == before[PREPASS]
    finally:
        tc05.set_prefill_prepass_cap(2 << 30)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("causal", [0, 1])
def test_tc05_kv_split_long_context_few_items(tc05, dtype, causal):
    """few 256-row items against long contexts: every item's KV range is cut into parts (fp32 partials + merge kernel);
    sequences shorter than the part count, a query longer than one tile pair, causal rows that see nothing in late parts"""
    rng = np.random.default_rng(69)
    before = tc05.prefill_path_counts()
    with took(tc05, TC05):
        _run_paged_prefill(tc05, rng, [64, 10, 70, 3], [4096 + 64, 3000, 5000, 40], 32, 8, 128, dtype, causal=causal)
    assert tc05.prefill_path_counts()[3] == before[3] + 1, "the launch did not split its KV ranges"


@pytest.mark.parametrize("dtype", DTYPES)
def test_tc05_kv_split_tree_paged(tc05, dtype):
    """C5's shape in small: token trees against a long committed context, split KV + tree mask in the last part"""
    rng = np.random.default_rng(70)
    trees = [_random_tree(rng, 64), [(k - 1) // 2 if k else -1 for k in range(64)]]
    before = tc05.prefill_path_counts()
    with took(tc05, TC05):
        _run_paged_prefill(tc05, rng, [64, 64], [64 + 4000, 64 + 2500], 32, 8, 128, dtype, tree=_tree_arrays(trees))
    assert tc05.prefill_path_counts()[3] == before[3] + 1
