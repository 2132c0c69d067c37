"""GPU parity tests: every C-ABI entry point (include/tvm_b200.h) against the CPU oracle on the same
seeded inputs.  Bar: bit-exact for copies / index work; max-abs 2e-3 / rtol 1e-2 for O and LSE
(north_star tolerance).  Scenario shapes follow the reference's own test file
(tests/python/relax/test_runtime_builtin_paged_attention_kv_cache_cpu.py)."""
import numpy as np
import pytest

from oracle import kernels as ok
from tests.util import assert_close, make_paged_cache, rand16, to_dev, to_np

pytestmark = pytest.mark.gpu

DTYPES = ["float16", "bfloat16"]


def _i32(x):
    return to_dev(np.asarray(x, np.int32))


@pytest.fixture(scope="module")
def capi(built_lib):
    from tvm_b200 import capi

    capi.lib()
    return capi


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hkv,d", [(8, 128), (4, 64), (1, 128)])
def test_transpose_append_bit_exact(capi, dtype, hkv, d):
    import torch

    rng = np.random.default_rng(0)
    P, page = 37, 16
    n = 100
    pages = rand16(rng, (P, 2, hkv, page, d), dtype)
    k = rand16(rng, (n, hkv, d), dtype)
    v = rand16(rng, (n, hkv, d), dtype)
    pm = rng.permutation(P * page)[:n].astype(np.int32)
    pm[::7] = -1  # "do not append"
    want = pages.copy()
    ok.transpose_append(want, k, v, pm)
    dp = to_dev(pages, dtype)
    capi.transpose_append(dp, to_dev(k, dtype), to_dev(v, dtype), _i32(pm))
    torch.cuda.synchronize()
    assert np.array_equal(ok.to_bits16(to_np(dp), dtype), ok.to_bits16(want, dtype))
    # debug_get_kv round trip (the reference's verify_cached_kv)
    live = pm[pm >= 0]
    ko = torch.zeros((2, len(live), hkv, d), dtype=dp.dtype, device="cuda")
    vo = torch.zeros_like(ko)
    capi.debug_get_kv(dp, _i32(live), ko, vo, 1)
    torch.cuda.synchronize()
    wk, wv = ok.debug_get_kv(want, live)
    assert np.array_equal(to_np(ko[1]), wk) and np.array_equal(to_np(vo[1]), wv)
    assert float(ko[0].abs().sum()) == 0.0


def test_transpose_append_empty(capi):
    import torch

    pages = torch.zeros((3, 2, 8, 16, 128), dtype=torch.float16, device="cuda")
    k = torch.zeros((0, 8, 128), dtype=torch.float16, device="cuda")
    capi.transpose_append(pages, k, k, torch.zeros((0,), dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()


@pytest.mark.parametrize("dtype", DTYPES)
def test_copy_single_page_and_compact(capi, dtype):
    import torch

    rng = np.random.default_rng(1)
    P, hkv, page, d = 9, 8, 16, 128
    pages = rand16(rng, (P, 2, hkv, page, d), dtype)
    want = pages.copy()
    dp = to_dev(pages, dtype)
    for src, tgt, ln in [(2, 3, 2), (0, 8, 16), (5, 1, 0), (4, 6, 15)]:
        ok.copy_single_page(want, src, tgt, ln)
        capi.copy_single_page(dp, src, tgt, ln)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(dp), want)
    # compaction incl. a chain where a destination is a later source (serial order matters)
    indptr = np.array([0, 3, 3, 5], np.int32)
    src_dst = np.array([[20, 21, 5, 100, 33], [5, 20, 6, 101, 34]], np.int32)
    ok.compact_kv_copy(want, indptr, src_dst, 3)
    capi.compact_kv_copy(dp, _i32(indptr), _i32(src_dst), 3)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(dp), want)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("apply_rope", [0, 1])
@pytest.mark.parametrize("theta", [1e4, 5e5])
def test_split_rotary(capi, dtype, apply_rope, theta):
    import torch

    rng = np.random.default_rng(2)
    n, hq, hkv, d = 53, 32, 8, 128
    qkv = rand16(rng, (n, hq + 2 * hkv, d), dtype)
    pos = rng.integers(0, 4096, n).astype(np.int32)
    wq, wk, wv = ok.split_rotary(qkv, pos, hq, hkv, apply_rope, theta, 1.0, dtype)
    tdt = to_dev(qkv, dtype).dtype
    q = torch.empty((n, hq, d), dtype=tdt, device="cuda")
    k = torch.empty((n, hkv, d), dtype=tdt, device="cuda")
    v = torch.empty((n, hkv, d), dtype=tdt, device="cuda")
    capi.split_rotary(to_dev(qkv, dtype), _i32(pos), q, k, v, apply_rope, 1.0, theta)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(v), wv)  # v is a pure copy
    if apply_rope == 0:
        assert np.array_equal(to_np(q), wq) and np.array_equal(to_np(k), wk)
    else:
        # fp32 sin/cos of arguments up to 4096 rad: allow 1 dtype ulp of slack on top of the tolerance
        assert_close("q", to_np(q), wq, atol=2e-2 if dtype == "bfloat16" else 4e-3)
        assert_close("k", to_np(k), wk, atol=2e-2 if dtype == "bfloat16" else 4e-3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("apply_rope", [0, 1])
def test_split_rotary_append_fused(capi, dtype, apply_rope):
    """The one-launch f_split_rotary + f_transpose_append must leave q, k, v AND the pages bit-identical to the
    two-call sequence (and to the oracle's append of the oracle-rotated... of the GPU-rotated k / v), skipped
    slots (-1) included."""
    import torch

    rng = np.random.default_rng(4)
    n, hq, hkv, d, npages = 37, 32, 8, 128, 9
    qkv = to_dev(rand16(rng, (n, hq + 2 * hkv, d), dtype), dtype)
    pos = _i32(rng.integers(0, 4096, n).astype(np.int32))
    slots = rng.permutation(npages * 16)[:n].astype(np.int32)
    slots[[3, 11]] = -1
    pages0 = rand16(rng, (npages, 2, hkv, 16, d), dtype)
    tdt = qkv.dtype
    outs = []
    for fused in (False, True):
        q = torch.empty((n, hq, d), dtype=tdt, device="cuda")
        k = torch.empty((n, hkv, d), dtype=tdt, device="cuda")
        v = torch.empty((n, hkv, d), dtype=tdt, device="cuda")
        pages = to_dev(pages0, dtype)
        if fused:
            capi.split_rotary_append(qkv, pos, _i32(slots), q, k, v, pages, apply_rope, 1.0, 5e5)
        else:
            capi.split_rotary(qkv, pos, q, k, v, apply_rope, 1.0, 5e5)
            capi.transpose_append(pages, k, v, _i32(slots))
        torch.cuda.synchronize()
        outs.append((q, k, v, pages))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    want = pages0.copy()
    ok.transpose_append(want, to_np(outs[1][1]), to_np(outs[1][2]), slots)
    assert np.array_equal(to_np(outs[1][3]), want)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("apply_rope,rotary_dim", [(1, 0), (1, 64), (0, 0)])
def test_split_rotary_large_batches_match_the_per_token_kernel(capi, dtype, apply_rope, rotary_dim):
    """>= 1024 tokens take the warp-per-token kernel (page_kernels.cu split_rotary_warp_kernel); its q / k / v and the
    appended pages must be bit-identical to the one-CTA-per-token kernel, reached by feeding the same tokens in chunks
    of 512, with and without the fused append, skipped slots (-1) and a partial rotary_dim included."""
    import torch

    rng = np.random.default_rng(41)
    n, hq, hkv, d = 2600, 32, 8, 128
    npages = n // 16 + 8
    qkv = to_dev(rand16(rng, (n, hq + 2 * hkv, d), dtype), dtype)
    pos = _i32(rng.integers(0, 8192, n).astype(np.int32))
    slots = rng.permutation(npages * 16)[:n].astype(np.int32)
    slots[::37] = -1
    dslots = _i32(slots)
    pages0 = to_dev(rand16(rng, (npages, 2, hkv, 16, d), dtype), dtype)
    tdt = qkv.dtype
    res = []
    for chunk in (n, 512):
        q = torch.empty((n, hq, d), dtype=tdt, device="cuda")
        k = torch.empty((n, hkv, d), dtype=tdt, device="cuda")
        v = torch.empty((n, hkv, d), dtype=tdt, device="cuda")
        q2, k2, v2 = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        pages = pages0.clone()
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            capi.split_rotary(qkv[a:b], pos[a:b], q[a:b], k[a:b], v[a:b], apply_rope, 1.0, 5e5, rotary_dim)
            capi.split_rotary_append(qkv[a:b], pos[a:b], dslots[a:b], q2[a:b], k2[a:b], v2[a:b], pages, apply_rope, 1.0, 5e5,
                                     rotary_dim)
        torch.cuda.synchronize()
        assert torch.equal(q, q2) and torch.equal(k, k2) and torch.equal(v, v2)
        res.append((q, k, v, pages))
    for a, b in zip(*res):
        assert torch.equal(a, b)
    want = to_np(pages0).copy()
    ok.transpose_append(want, to_np(res[0][1]), to_np(res[0][2]), slots)
    assert np.array_equal(to_np(res[0][3]), want)
    if rotary_dim == 0:
        wq, wk, wv = ok.split_rotary(to_np(qkv), to_np(pos), hq, hkv, apply_rope, 5e5, 1.0, dtype)
        assert np.array_equal(to_np(res[0][2]), wv)
        assert_close("q", to_np(res[0][0]), wq, atol=3.2e-2 if dtype == "bfloat16" else 4e-3)
        assert_close("k", to_np(res[0][1]), wk, atol=3.2e-2 if dtype == "bfloat16" else 4e-3)


@pytest.mark.parametrize("dtype", DTYPES)
def test_merge_state_inplace(capi, dtype):
    import torch

    rng = np.random.default_rng(3)
    n, h, d = 77, 32, 128
    v = rand16(rng, (n, h, d), dtype)
    vo = rand16(rng, (n, h, d), dtype)
    s = (rng.standard_normal((n, h)) * 4).astype(np.float32)
    so = (rng.standard_normal((n, h)) * 4).astype(np.float32)
    so[3] = ok.NEG_INIT  # empty other side: merge must be a no-op
    s[5] = ok.NEG_INIT
    wv, ws = ok.merge_state_inplace(v, s, vo, so, dtype)
    dv, ds = to_dev(v, dtype), to_dev(s)
    capi.merge_state_inplace(dv, ds, to_dev(vo, dtype), to_dev(so))
    torch.cuda.synchronize()
    assert_close("merged v", to_np(dv), wv)
    assert_close("merged s", to_np(ds), ws)
    assert np.array_equal(to_np(dv)[3], v[3])


# ---------------------------------------------------------------------------------------------------
def _run_decode(capi, rng, kv_lens, hq, hkv, d, dtype, rotary_mode=0, sliding=None, theta=1e4):
    import torch

    B = len(kv_lens)
    c = make_paged_cache(rng, kv_lens, hkv, d, dtype, sliding=sliding)
    q = rand16(rng, (B, hq, d), dtype)
    kpos = rng.integers(0, 50, B).astype(np.int32)
    qpos = (kpos + np.array(kv_lens) - 1).astype(np.int32)
    sm = d ** -0.5
    wo, wl = ok.attention_decode(q, c["pages"], c["page_indptr"], c["page_values"], c["length_info"], kpos, qpos,
                                 rotary_mode, 1.0, theta, sm, dtype)
    dq = to_dev(q, dtype)
    o = torch.full((B, hq, d), float("nan"), dtype=dq.dtype, device="cuda")
    lse = torch.full((B, hq), float("nan"), dtype=torch.float32, device="cuda")
    capi.attention_decode(dq, to_dev(c["pages"], dtype), _i32(c["page_indptr"]), _i32(c["page_values"]),
                          _i32(c["length_info"]), _i32(kpos), _i32(qpos), o, lse, rotary_mode, 1.0, theta, sm)
    torch.cuda.synchronize()
    assert_close("decode O", to_np(o), wo)
    assert_close("decode LSE", to_np(lse), wl)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hq,hkv,d", [(32, 8, 128), (32, 4, 128), (8, 8, 128), (32, 8, 64), (8, 1, 128)])
def test_decode_small(capi, dtype, hq, hkv, d):
    rng = np.random.default_rng(10)
    _run_decode(capi, rng, [11, 21, 31, 41], hq, hkv, d, dtype)  # C1 decode lengths


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hq,hkv", [(16, 1), (32, 2), (64, 2), (24, 1)])
def test_decode_gqa_groups_above_eight(capi, dtype, hq, hkv):
    """GQA groups of 16 / 32 / 24 query heads per kv head (Llama-3.1-405B: 128 / 8): cut into virtual heads of 8, each a
    pass over the same pages; split-KV lengths included"""
    rng = np.random.default_rng(16)
    _run_decode(capi, rng, [11, 300, 1, 2049, 64], hq, hkv, 128, dtype)
    _run_decode(capi, rng, [40, 7], hq, hkv, 128, dtype, rotary_mode=1)


def test_unsupported_shapes_fail_loudly(capi):
    """limits raise, nothing is computed by a fallback: a GQA group that is neither <= 8 nor a multiple of 8, head_dim 96,
    8-slot pages, a float32 cache"""
    import torch

    from tvm_b200.capi import TvmB200Error

    rng = np.random.default_rng(17)
    with pytest.raises(TvmB200Error, match="group size 12"):
        _run_decode(capi, rng, [11, 21], 12, 1, 128, "float16")
    with pytest.raises(TvmB200Error, match="head_dim 96"):
        _run_decode(capi, rng, [11, 21], 8, 2, 96, "float16")
    i32 = lambda *a: torch.zeros(a, dtype=torch.int32, device="cuda")  # noqa: E731
    q = torch.zeros((1, 8, 128), dtype=torch.float16, device="cuda")
    lse = torch.zeros((1, 8), dtype=torch.float32, device="cuda")
    pages8 = torch.zeros((4, 2, 2, 8, 128), dtype=torch.float16, device="cuda")
    with pytest.raises(TvmB200Error, match="page_size 8"):
        capi.attention_decode(q, pages8, i32(2), i32(1), i32(1), i32(1), i32(1), q.clone(), lse, 0, 1.0, 1e4, 1.0)
    with pytest.raises(TvmB200Error, match="page_size 8"):
        capi.attention_prefill_paged(q, i32(2), pages8, i32(2), i32(1), i32(1), i32(1), i32(1), q.clone(), lse, 0, 0, 1.0, 1e4, 1.0)
    pages32 = torch.zeros((4, 2, 2, 16, 128), dtype=torch.float32, device="cuda")
    with pytest.raises(TvmB200Error, match="dtype"):
        capi.attention_decode(q.float(), pages32, i32(2), i32(1), i32(1), i32(1), i32(1), q.float(), lse, 0, 1.0, 1e4, 1.0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_decode_ragged_lengths_and_empty(capi, dtype):
    rng = np.random.default_rng(11)
    # 0 = sequence with no pages at this depth (Appendix C.4): O = 0, lse = -5e4
    _run_decode(capi, rng, [1, 16, 17, 0, 255, 256, 257, 1000, 0, 5], 32, 8, 128, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_decode_split_kv_long(capi, dtype):
    rng = np.random.default_rng(12)
    _run_decode(capi, rng, [8192, 3, 4097], 32, 8, 128, dtype)  # few long sequences => split-KV + merge


@pytest.mark.parametrize("dtype", DTYPES)
def test_decode_batch64(capi, dtype):
    rng = np.random.default_rng(13)
    _run_decode(capi, rng, list(rng.integers(1, 700, 64)), 32, 8, 128, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("apply_rope", [1, 0])
@pytest.mark.parametrize("kv_lens", [[11, 16, 17, 300, 1], [4096, 2049, 33]])
def test_decode_fused_qkv_matches_the_three_call_step(capi, dtype, apply_rope, kv_lens):
    """tvmb200_attention_decode_fused_qkv = split_rotary + transpose_append + attention_decode: the pages must end up
    bit-identical to the three-call sequence, O / LSE within the tolerance of it AND of the oracle's restatement of
    the three callbacks (split-KV and single-chunk sequences, new token at the start / middle / end of a page)."""
    import torch

    rng = np.random.default_rng(19)
    hq, hkv, d = 32, 8, 128
    B = len(kv_lens)
    c = make_paged_cache(rng, kv_lens, hkv, d, dtype)
    qkv = rand16(rng, (B, hq + 2 * hkv, d), dtype)
    qpos = (np.array(kv_lens) - 1 + 7).astype(np.int32)      # position of the new token (an arbitrary rope offset)
    slots = np.array([int(c["page_values"][c["page_indptr"][b + 1] - 1]) * 16 + (kv_lens[b] - 1) % 16 for b in range(B)],
                     np.int32)
    kofs = np.full(B, 7, np.int32)
    sm, theta = d ** -0.5, 5e5
    # oracle: the three callbacks
    wq, wk, wv = ok.split_rotary(qkv, qpos, hq, hkv, apply_rope, theta, 1.0, dtype)
    wpages = c["pages"].copy()
    ok.transpose_append(wpages, wk, wv, slots)
    wo, wl = ok.attention_decode(wq, wpages, c["page_indptr"], c["page_values"], c["length_info"], kofs, qpos, 0, 1.0,
                                 theta, sm, dtype)
    dqkv = to_dev(qkv, dtype)
    ip, pv, li = _i32(c["page_indptr"]), _i32(c["page_values"]), _i32(c["length_info"])
    tdt = dqkv.dtype
    # three calls
    p3 = to_dev(c["pages"], dtype)
    q = torch.empty((B, hq, d), dtype=tdt, device="cuda")
    k = torch.empty((B, hkv, d), dtype=tdt, device="cuda")
    v = torch.empty((B, hkv, d), dtype=tdt, device="cuda")
    o3 = torch.empty((B, hq, d), dtype=tdt, device="cuda")
    l3 = torch.empty((B, hq), dtype=torch.float32, device="cuda")
    capi.split_rotary_append(dqkv, _i32(qpos), _i32(slots), q, k, v, p3, apply_rope, 1.0, theta)
    capi.attention_decode(q, p3, ip, pv, li, _i32(kofs), _i32(qpos), o3, l3, 0, 1.0, theta, sm)
    # one call
    p1 = to_dev(c["pages"], dtype)
    o1 = torch.full((B, hq, d), float("nan"), dtype=tdt, device="cuda")
    l1 = torch.full((B, hq), float("nan"), dtype=torch.float32, device="cuda")
    n0 = capi.launch_count()
    capi.attention_decode_fused_qkv(dqkv, _i32(qpos), _i32(slots), p1, ip, pv, li, _i32(kofs), o1, l1, apply_rope, 1.0,
                                    theta, sm)
    torch.cuda.synchronize()
    assert capi.launch_count() - n0 <= 2          # decode (+ merge when a sequence is split)
    # ... and the same step with the single-rank peer gather (this rank is its own peer)
    p2 = to_dev(c["pages"], dtype)
    o2, l2 = torch.empty_like(o1), torch.empty_like(l1)
    gathered = torch.full((B, hq, d), float("nan"), dtype=tdt, device="cuda")
    flags = torch.zeros(64, dtype=torch.int32, device="cuda")
    capi.attention_decode_fused_qkv_gather(dqkv, _i32(qpos), _i32(slots), p2, ip, pv, li, _i32(kofs), o2, l2, apply_rope,
                                           1.0, theta, sm, [gathered.data_ptr()], [flags.data_ptr()], 0, 1)
    capi.wait_peer_flags(flags, 1, 1)
    torch.cuda.synchronize()
    assert torch.equal(p2, p1) and torch.equal(o2, o1) and torch.equal(l2, l1) and torch.equal(gathered, o1)
    assert torch.equal(p1, p3)                     # the appended k / v are the very same values
    assert_close("fused O vs three calls", to_np(o1), to_np(o3))
    assert_close("fused LSE vs three calls", to_np(l1), to_np(l3))
    assert_close("fused O vs oracle", to_np(o1), wo)
    assert_close("fused LSE vs oracle", to_np(l1), wl)


def test_decode_step_is_cuda_graph_capturable(capi):
    """the per-layer sequence (fused rotary + append, split-KV decode, merge -- the last two launched as
    programmatic dependents) captured into one CUDA graph and replayed gives the eager result bit for bit; nothing in
    the path allocates or synchronises once the workspace is sized."""
    import torch

    rng = np.random.default_rng(17)
    dtype, hq, hkv, d = "bfloat16", 32, 8, 128
    kv_lens = [4096, 17, 2000, 900]
    B = len(kv_lens)
    c = make_paged_cache(rng, kv_lens, hkv, d, dtype)
    pages = to_dev(c["pages"], dtype)
    qkv = to_dev(rand16(rng, (B, hq + 2 * hkv, d), dtype), dtype)
    qpos = _i32(np.array(kv_lens, np.int32) - 1)
    slots = _i32(np.array([int(c["page_values"][c["page_indptr"][b + 1] - 1]) * 16 + (kv_lens[b] - 1) % 16
                           for b in range(B)], np.int32))
    kofs = _i32(np.zeros(B, np.int32))
    ip, pv, li = _i32(c["page_indptr"]), _i32(c["page_values"]), _i32(c["length_info"])
    tdt = qkv.dtype
    q = torch.empty((B, hq, d), dtype=tdt, device="cuda")
    k = torch.empty((B, hkv, d), dtype=tdt, device="cuda")
    v = torch.empty((B, hkv, d), dtype=tdt, device="cuda")
    o = torch.empty((B, hq, d), dtype=tdt, device="cuda")
    lse = torch.empty((B, hq), dtype=torch.float32, device="cuda")

    def step():
        capi.split_rotary_append(qkv, qpos, slots, q, k, v, pages, 1, 1.0, 5e5)
        capi.attention_decode(q, pages, ip, pv, li, kofs, qpos, o, lse, 0, 1.0, 5e5, d ** -0.5)

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step()  # sizes the workspace outside the capture
        torch.cuda.synchronize()
        want_o, want_lse = o.clone(), lse.clone()
        o.zero_()
        lse.zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            step()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(o, want_o) and torch.equal(lse, want_lse)


@pytest.mark.parametrize("kv_lens", [[11, 70, 300, 0, 16], [8192, 3, 4097]])
def test_decode_gather_single_rank(capi, kv_lens):
    """tvmb200_attention_decode_gather with world = 1 (this rank is its own peer): the gathered buffer, the local
    output and the LSE must be bit-identical to plain attention_decode, whether or not the sequences are split, and the
    rank's flag must carry the epoch afterwards (two consecutive epochs: the block counter resets itself)."""
    import torch

    rng = np.random.default_rng(16)
    dtype, hq, hkv, d = "bfloat16", 32, 8, 128
    B = len(kv_lens)
    c = make_paged_cache(rng, kv_lens, hkv, d, dtype)
    q = to_dev(rand16(rng, (B, hq, d), dtype), dtype)
    pages = to_dev(c["pages"], dtype)
    kpos = _i32(np.zeros(B, np.int32))
    qpos = _i32(np.maximum(np.array(kv_lens) - 1, 0).astype(np.int32))
    args = (q, pages, _i32(c["page_indptr"]), _i32(c["page_values"]), _i32(c["length_info"]), kpos, qpos)
    o = torch.empty((B, hq, d), dtype=q.dtype, device="cuda")
    lse = torch.empty((B, hq), dtype=torch.float32, device="cuda")
    capi.attention_decode(*args, o, lse, 0, 1.0, 1e4, d ** -0.5)
    flags = torch.zeros(64, dtype=torch.int32, device="cuda")
    for epoch in (1, 2):
        gathered = torch.full((B, hq, d), float("nan"), dtype=q.dtype, device="cuda")
        o2 = torch.full_like(o, float("nan"))
        lse2 = torch.full_like(lse, float("nan"))
        capi.attention_decode_gather(*args, o2, lse2, 0, 1.0, 1e4, d ** -0.5, [gathered.data_ptr()], [flags.data_ptr()],
                                     0, epoch)
        capi.wait_peer_flags(flags, 1, epoch)
        torch.cuda.synchronize()
        assert torch.equal(o2, o) and torch.equal(lse2, lse) and torch.equal(gathered, o)
        assert int(flags[0]) == epoch


@pytest.mark.parametrize("dtype", DTYPES)
def test_decode_sliding_window(capi, dtype):
    rng = np.random.default_rng(14)
    # (slots in pages, (sliding offset, sink)): visible = sink + [off, slots)
    _run_decode(capi, rng, [40, 100, 64, 33], 32, 8, 128, dtype, sliding=[(0, 0), (37, 4), (16, 16), (20, 0)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_decode_inline_rope(capi, dtype):
    rng = np.random.default_rng(15)
    _run_decode(capi, rng, [11, 70, 300], 32, 8, 128, dtype, rotary_mode=1)


LLAMA31 = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
           "original_max_position_embeddings": 8192}


@pytest.fixture()
def llama3_rope(capi):
    """llama3 frequency scaling (rope_freq_llama3, position_embedding.py:130-160) on both sides: kernels and oracle"""
    capi.set_rope_scaling(LLAMA31)
    ok.set_rope_scaling(LLAMA31)
    yield capi
    capi.set_rope_scaling(None)
    ok.set_rope_scaling(None)


def test_rope_llama3_against_the_reference_golden(llama3_rope):
    """tests/golden/rope_llama3.npz: outputs of the reference's own fused_rope and inline-RoPE decode built with the
    Llama-3.1 rope_scaling (oracle/ref_harness/gen_golden_rope.py)."""
    import torch
    from pathlib import Path

    capi = llama3_rope
    g = np.load(Path(__file__).parent / "golden" / "rope_llama3.npz")
    theta, scale = float(g["params"][0]), float(g["params"][1])
    n, hq, hkv, d = g["qkv"].shape[0], 8, 2, 128
    q = torch.empty((n, hq, d), dtype=torch.float16, device="cuda")
    k = torch.empty((n, hkv, d), dtype=torch.float16, device="cuda")
    v = torch.empty((n, hkv, d), dtype=torch.float16, device="cuda")
    capi.split_rotary(torch.from_numpy(g["qkv"]).cuda(), _i32(g["pos"]), q, k, v, 1, scale, theta)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(v), g["v"].astype(np.float32))
    for i, pos in enumerate(g["pos"]):  # the angle is a float32: its ulp grows with the position (see the oracle test)
        atol = 4e-3 + 3e-7 * float(pos)
        assert_close(f"q[{i}]", to_np(q)[i], g["q"][i].astype(np.float32), atol=atol)
        assert_close(f"k[{i}]", to_np(k)[i], g["k"][i].astype(np.float32), atol=atol)
    B = g["qd"].shape[0]
    o = torch.empty((B, hq, d), dtype=torch.float16, device="cuda")
    lse = torch.empty((B, hq), dtype=torch.float32, device="cuda")
    capi.attention_decode(torch.from_numpy(g["qd"]).cuda(), torch.from_numpy(g["pages"]).cuda(), _i32(g["page_indptr"]),
                          _i32(g["page_values"]), _i32(g["length_info"]), _i32(g["kofs"]), _i32(g["qpos"]), o, lse, 1,
                          scale, theta, d ** -0.5)
    torch.cuda.synchronize()
    assert_close("decode O", to_np(o), g["o"].astype(np.float32), atol=6e-3)
    assert_close("decode LSE", to_np(lse), g["lse"], atol=2e-2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_rope_llama3_inline_paths_vs_oracle(llama3_rope, dtype):
    """every kernel with an inline-RoPE path under llama3 scaling: decode, paged prefill, ragged prefill"""
    rng = np.random.default_rng(18)
    _run_decode(llama3_rope, rng, [11, 70, 300], 32, 8, 128, dtype, rotary_mode=1, theta=5e5)
    _run_paged_prefill(llama3_rope, rng, [3, 17, 40], [20, 100, 333], 32, 8, 128, dtype, causal=0, rotary_mode=1)
    _run_ragged(llama3_rope, rng, [10, 65, 130], [10, 65, 130], 32, 8, 128, dtype, causal=1, rotary_mode=1)


def test_rope_scaling_rejects_unknown_kinds(capi):
    with pytest.raises(Exception, match="not implemented"):
        capi.set_rope_scaling({"rope_type": "longrope"})
    L = capi.lib()
    assert L.tvmb200_set_rope_scaling(7, 1.0, 1.0, 4.0, 8192.0) != 0
    assert b"unsupported" in L.tvmb200_last_error()


# ---------------------------------------------------------------------------------------------------
def _run_ragged(capi, rng, q_lens, kv_lens, hq, hkv, d, dtype, causal=1, rotary_mode=0, tree=None, q_scale=1.0,
                v_scale=1.0):
    import torch

    B = len(q_lens)
    qi = np.zeros(B + 1, np.int32)
    qi[1:] = np.cumsum(q_lens)
    ki = np.zeros(B + 1, np.int32)
    ki[1:] = np.cumsum(kv_lens)
    n, m = int(qi[-1]), int(ki[-1])
    q = ok.round_dtype(rand16(rng, (n, hq, d), dtype) * np.float32(q_scale), dtype)
    k = rand16(rng, (m, hkv, d), dtype)
    v = ok.round_dtype(rand16(rng, (m, hkv, d), dtype) * np.float32(v_scale), dtype)
    kofs = rng.integers(0, 30, B).astype(np.int32)
    qpos = np.concatenate([kofs[b] + kv_lens[b] - q_lens[b] + np.arange(q_lens[b]) for b in range(B)]).astype(np.int32)
    sm = d ** -0.5
    dq = to_dev(q, dtype)
    o = torch.full((n, hq, d), float("nan"), dtype=dq.dtype, device="cuda")
    lse = torch.full((n, hq), float("nan"), dtype=torch.float32, device="cuda")
    if tree is None:
        wo, wl = ok.attention_prefill_ragged(q, qi, k, v, ki, qpos, kofs, causal, rotary_mode, 1.0, 1e4, sm, dtype)
        capi.attention_prefill_ragged(dq, _i32(qi), to_dev(k, dtype), to_dev(v, dtype), _i32(ki), _i32(qpos),
                                      _i32(kofs), o, lse, causal, rotary_mode, 1.0, 1e4, sm)
    else:
        mn, mask = tree
        wo, wl = ok.attention_prefill_ragged(q, qi, k, v, ki, qpos, kofs, 0, rotary_mode, 1.0, 1e4, sm, dtype,
                                             mn_indptr=mn, tree_mask=mask)
        capi.attention_prefill_tree_ragged(dq, _i32(qi), to_dev(k, dtype), to_dev(v, dtype), _i32(ki), _i32(qpos),
                                           _i32(mn), _i32(mask), o, lse, rotary_mode, 1.0, 1e4, sm)
    torch.cuda.synchronize()
    assert_close("ragged O", to_np(o), wo)
    assert_close("ragged LSE", to_np(lse), wl)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hq,hkv,d", [(32, 8, 128), (32, 4, 64), (8, 8, 128)])
def test_prefill_ragged_c1(capi, dtype, hq, hkv, d):
    rng = np.random.default_rng(20)
    _run_ragged(capi, rng, [10, 20, 30, 40], [10, 20, 30, 40], hq, hkv, d, dtype)  # C1 prefill


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("causal", [0, 1])
def test_prefill_ragged_uneven(capi, dtype, causal):
    rng = np.random.default_rng(21)
    _run_ragged(capi, rng, [1, 129, 64, 300, 7], [5, 129, 200, 300, 7], 32, 8, 128, dtype, causal=causal)


@pytest.mark.parametrize("dtype", DTYPES)
def test_prefill_ragged_inline_rope(capi, dtype):
    rng = np.random.default_rng(22)
    _run_ragged(capi, rng, [10, 70], [10, 70], 32, 8, 128, dtype, rotary_mode=1)


def _dfs_mask(parents):
    """(dfs_order, subtree_end) rows as ConstructTokenTreeMask emits them (paged_kv_cache.cc:1900-1918)."""
    n = len(parents)
    children = [[] for _ in range(n)]
    roots = []
    for i, p in enumerate(parents):
        (roots if p < 0 else children[p]).append(i)
    order = [0] * n
    end = [0] * n
    cnt = 0

    def visit(u):
        nonlocal cnt
        order[u] = cnt
        cnt += 1
        for c in children[u]:
            visit(c)
        end[u] = cnt

    for r in roots:
        visit(r)
    return np.array([[order[i], end[i]] for i in range(n)], np.int32)


@pytest.mark.parametrize("dtype", DTYPES)
def test_prefill_tree_ragged(capi, dtype):
    rng = np.random.default_rng(23)
    trees = [[-1, 0, 0, 1], [(k - 1) // 2 if k else -1 for k in range(64)], [-1, 0, 1, 2, 3, 4, 5]]
    masks = [_dfs_mask(t) for t in trees]
    mn = np.zeros(len(trees) + 1, np.int32)
    mn[1:] = np.cumsum([len(t) for t in trees])
    lens = [len(t) for t in trees]
    _run_ragged(capi, rng, lens, lens, 32, 8, 128, dtype, tree=(mn, np.concatenate(masks)))


def _run_paged_prefill(capi, rng, q_lens, kv_lens, hq, hkv, d, dtype, causal=0, rotary_mode=0, sliding=None,
                       layer_sws=0, tree=None, nan_tail=False):
    import torch

    B = len(q_lens)
    qi = np.zeros(B + 1, np.int32)
    qi[1:] = np.cumsum(q_lens)
    n = int(qi[-1])
    c = make_paged_cache(rng, kv_lens, hkv, d, dtype, sliding=sliding)
    q = rand16(rng, (n, hq, d), dtype)
    kofs = rng.integers(0, 30, B).astype(np.int32)
    qpos = np.concatenate([kofs[b] + kv_lens[b] + np.arange(q_lens[b]) for b in range(B)]).astype(np.int32)
    sm = d ** -0.5
    dq = to_dev(q, dtype)
    o = torch.full((n, hq, d), float("nan"), dtype=dq.dtype, device="cuda")
    lse = torch.full((n, hq), float("nan"), dtype=torch.float32, device="cuda")
    dpages = to_dev(c["pages"], dtype)
    if nan_tail:  # poison every slot past kv_len of each sequence's last page (never visible to the oracle)
        for b in range(B):
            if kv_lens[b] % 16:
                last = int(c["page_values"][c["page_indptr"][b + 1] - 1])
                dpages[last, :, :, kv_lens[b] % 16:, :] = float("nan")
    args = (dq, _i32(qi), dpages, _i32(c["page_indptr"]), _i32(c["page_values"]),
            _i32(c["length_info"]), _i32(kofs), _i32(qpos), o, lse)
    if tree is None:
        wo, wl = ok.attention_prefill_paged(q, qi, c["pages"], c["page_indptr"], c["page_values"], c["length_info"],
                                            kofs, qpos, causal, rotary_mode, 1.0, 1e4, sm, dtype,
                                            sliding_window_size=layer_sws)
        capi.attention_prefill_paged(*args, causal, rotary_mode, 1.0, 1e4, sm, layer_sliding_window_size=layer_sws)
    else:
        ti, to = tree
        wo, wl = ok.attention_prefill_paged(q, qi, c["pages"], c["page_indptr"], c["page_values"], c["length_info"],
                                            kofs, qpos, 0, rotary_mode, 1.0, 1e4, sm, dtype, tree_indptr=ti,
                                            tree_order=to)
        capi.attention_prefill_tree_paged(*args, rotary_mode, 1.0, 1e4, sm, _i32(ti), _i32(to))
    torch.cuda.synchronize()
    # rows whose mask hides every column are outside the comparison (see DESIGN.md "fully masked rows")
    assert_close("paged prefill O", to_np(o), wo)
    assert_close("paged prefill LSE", to_np(lse), wl)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("causal", [0, 1])
def test_prefill_paged(capi, dtype, causal):
    rng = np.random.default_rng(30)
    # causal=1 needs kv_len >= qo_len (cache already holds the new tokens)
    _run_paged_prefill(capi, rng, [3, 17, 1, 64, 5], [16, 18, 0 if not causal else 1, 300, 77], 32, 8, 128, dtype,
                       causal=causal)


@pytest.mark.parametrize("dtype", DTYPES)
def test_prefill_paged_hd64_gqa8(capi, dtype):
    rng = np.random.default_rng(31)
    _run_paged_prefill(capi, rng, [10, 33], [100, 47], 32, 4, 64, dtype, causal=0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_prefill_paged_sliding_and_inline_rope(capi, dtype):
    rng = np.random.default_rng(32)
    _run_paged_prefill(capi, rng, [4, 9, 2], [100, 64, 33], 32, 8, 128, dtype, causal=0, rotary_mode=1,
                       sliding=[(37, 4), (16, 16), (0, 0)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_prefill_paged_tree(capi, dtype):
    rng = np.random.default_rng(33)
    trees = [[-1, 0, 0, 1], [(k - 1) // 2 if k else -1 for k in range(15)]]
    masks = [_dfs_mask(t) for t in trees]
    ti = np.zeros(len(trees) + 1, np.int32)
    ti[1:] = np.cumsum([len(t) for t in trees])
    # the tree occupies the trailing columns of the cached KV; q rows = the tree nodes
    _run_paged_prefill(capi, rng, [4, 15], [4 + 20, 15 + 200], 32, 8, 128, dtype, tree=(ti, np.concatenate(masks)))


def test_kat_paged_prefill_layer_sliding_window(capi):
    """Known-answer test of the reference, literal copy of its inputs:
    tests/python/relax/test_runtime_builtin_paged_attention_kv_cache_tir.py:96-156
    (head_dim 64, 2 kv / 4 qo heads, V rows = 1,3,5, zero Q/K, window 3 => output [4, 5])."""
    import math

    import torch

    head_dim, hkv, hq, page = 64, 2, 4, 16
    pages = np.zeros((1, 2, hkv, page, head_dim), np.float32)
    for position, value in enumerate([1, 3, 5]):
        pages[0, 1, :, position, :] = value
    q = np.zeros((2, hq, head_dim), np.float32)
    dq = to_dev(q, "float16")
    o = torch.zeros((2, hq, head_dim), dtype=torch.float16, device="cuda")
    lse = torch.zeros((2, hq), dtype=torch.float32, device="cuda")
    capi.attention_prefill_paged(dq, _i32([0, 2]), to_dev(pages, "float16"), _i32([0, 1]), _i32([0]),
                                 _i32([[3], [0], [0]]), _i32([0]), _i32([3, 4]), o, lse, 1, 0, 1.0, 10000.0,
                                 1 / math.sqrt(head_dim), layer_sliding_window_size=3)
    torch.cuda.synchronize()
    np.testing.assert_allclose(to_np(o)[:, 0, 0], [4.0, 5.0], rtol=1e-3, atol=1e-3)
    # and the oracle gives the same known answer
    wo, _ = ok.attention_prefill_paged(q, [0, 2], pages, [0, 1], [0], np.array([[3], [0], [0]]), [0], [3, 4], 1, 0,
                                       1.0, 1e4, 1 / math.sqrt(head_dim), "float16", sliding_window_size=3)
    np.testing.assert_allclose(wo[:, 0, 0], [4.0, 5.0], rtol=1e-3, atol=1e-3)


def test_errors_are_loud(capi):
    import torch

    pages = torch.zeros((3, 2, 8, 16, 128), dtype=torch.float32, device="cuda")
    with pytest.raises(capi.TvmB200Error):
        capi.transpose_append(pages, pages, pages, pages)  # unsupported dtype
    cpu = torch.zeros((3, 2, 8, 16, 128), dtype=torch.float16)
    with pytest.raises(capi.TvmB200Error):
        capi.transpose_append(cpu, cpu, cpu, cpu)  # no CPU fallback
