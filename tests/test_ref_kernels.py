"""The oracle restatement vs the reference's OWN compiled CPU kernels (oracle/_ref, built from /root/reference by
oracle/ref_harness/emit_ref_kernels.py) on the C2 head shape.  Skipped where oracle/_ref was not built."""
import numpy as np
import pytest

from oracle import cpu_ref
from oracle import kernels as ok


def test_ref_decode_step_matches_oracle():
    mod = cpu_ref._ref_module("float16", 32, 8, 128)
    if mod is None:
        pytest.skip("oracle/_ref not built (needs the reference build, see oracle/ref_harness/)")
    import torch

    B, L, Hq, Hkv, D = 2, 100, 32, 8, 128
    inp = cpu_ref._decode_inputs(B, L, Hq, Hkv, D, "float16", seed=3)
    t = {k: torch.from_numpy(v.astype(np.float16) if v.dtype == np.float32 else v.copy()) for k, v in inp.items()}
    for nm, h in (("q", Hq), ("k", Hkv), ("v", Hkv), ("o", Hq)):
        t[nm] = torch.zeros((B, h, D), dtype=torch.float16)
    t["lse"] = torch.zeros((B, Hq), dtype=torch.float32)
    cpu_ref._one_step_ref(mod, t, Hq, Hkv, D)
    pages = inp["pages"].copy()
    inp2 = dict(inp, pages=pages)
    wo, wl = cpu_ref._one_step_port(inp2, Hq, Hkv, D, "float16")
    assert np.array_equal(t["pages"].float().numpy()[..., :, :][:, 1], pages[:, 1])  # V append bit-exact
    np.testing.assert_allclose(t["pages"].float().numpy(), pages, atol=2e-3, rtol=2e-3)  # K went through RoPE
    np.testing.assert_allclose(t["o"].float().numpy(), wo, atol=2e-3, rtol=2e-3)
    np.testing.assert_allclose(t["lse"].numpy(), wl, atol=2e-3, rtol=2e-3)


def _ragged_inputs(seed=5, lens=(10, 65, 130), Hq=32, Hkv=8, D=128):
    rng = np.random.default_rng(seed)
    n = sum(lens)
    ip = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    q = ok.round_dtype(rng.standard_normal((n, Hq, D)).astype(np.float32), "float16")
    k = ok.round_dtype(rng.standard_normal((n, Hkv, D)).astype(np.float32), "float16")
    v = ok.round_dtype(rng.standard_normal((n, Hkv, D)).astype(np.float32), "float16")
    qpos = np.concatenate([np.arange(x) for x in lens]).astype(np.int32)
    return dict(q=q, k=k, v=v, ip=ip, qpos=qpos, kofs=np.zeros(len(lens), np.int32))


def ref_ragged_prefill(mod, inp, rotary_mode=0, theta=5e5):
    """The reference's own _attention_prefill_ragged_cpu (compiled, oracle/_ref) on CPU tensors -> (O, LSE)."""
    import torch

    n, Hq, D = inp["q"].shape
    t = lambda a: torch.from_numpy(a.astype(np.float16) if a.dtype == np.float32 else a.copy())  # noqa: E731
    o, lse = torch.zeros((n, Hq, D), dtype=torch.float16), torch.zeros((n, Hq), dtype=torch.float32)
    mod["batch_prefill_ragged_kv_cpu"](t(inp["q"]), t(inp["ip"]), t(inp["k"]), t(inp["v"]), t(inp["ip"]), t(inp["qpos"]),
                                       t(inp["kofs"]), o, lse, 1, rotary_mode, 1.0, theta, D ** -0.5)
    return o.float().numpy(), lse.numpy()


def ref_merge(mod, v, s, v2, s2):
    import torch

    tv, ts = torch.from_numpy(v.astype(np.float16)), torch.from_numpy(s.copy())
    mod["merge_state_inplace_cpu"](tv, ts, torch.from_numpy(v2.astype(np.float16)), torch.from_numpy(s2.copy()))
    return tv.float().numpy(), ts.numpy()


@pytest.mark.parametrize("rotary_mode", [0, 1])
def test_ref_ragged_prefill_and_merge_match_oracle(rotary_mode):
    mod = cpu_ref._ref_module("float16", 32, 8, 128)
    if mod is None:
        pytest.skip("oracle/_ref not built (needs the reference build, see oracle/ref_harness/)")
    inp = _ragged_inputs()
    o, lse = ref_ragged_prefill(mod, inp, rotary_mode)
    wo, wl = ok.attention_prefill_ragged(inp["q"], inp["ip"], inp["k"], inp["v"], inp["ip"], inp["qpos"], inp["kofs"], 1,
                                         rotary_mode, 1.0, 5e5, 128 ** -0.5, "float16")
    np.testing.assert_allclose(o, wo, atol=2e-3, rtol=1e-2)
    np.testing.assert_allclose(lse, wl, atol=2e-3, rtol=1e-2)
    rng = np.random.default_rng(9)
    o2 = ok.round_dtype(rng.standard_normal(o.shape).astype(np.float32), "float16")
    lse2 = (lse + rng.standard_normal(lse.shape).astype(np.float32) * 3).astype(np.float32)
    mv, ms = ref_merge(mod, o, lse, o2, lse2)
    wv, ws = ok.merge_state_inplace(o.copy(), lse.copy(), o2, lse2, "float16")
    np.testing.assert_allclose(mv, wv, atol=2e-3, rtol=1e-2)
    np.testing.assert_allclose(ms, ws, atol=2e-3, rtol=1e-2)


# ---- the remaining attention kernels of the reference at the Llama-3-8B head shape ---------------------------------
HQ, HKV, D, THETA, SM = 32, 8, 128, 5e5, 128 ** -0.5


def _t(a):
    import torch

    a = np.asarray(a)
    return torch.from_numpy(np.ascontiguousarray(a.astype(np.float16) if a.dtype in (np.float32, np.float64) else a))


def _out(n):
    import torch

    return torch.zeros((n, HQ, D), dtype=torch.float16), torch.zeros((n, HQ), dtype=torch.float32)


def _need_ref():
    mod = cpu_ref._ref_module("float16", HQ, HKV, D)
    if mod is None:
        pytest.skip("oracle/_ref not built (needs the reference build, see oracle/ref_harness/)")
    try:
        mod["batch_prefill_paged_kv_cpu"]
    except Exception:
        pytest.skip("oracle/_ref holds the decode-step kernels only (re-run oracle/ref_harness/emit_ref_kernels.py)")
    return mod


def _dfs_mask(parents):
    """(dfs_order, subtree_end) rows as ConstructTokenTreeMask emits them (paged_kv_cache.cc:1900-1918)."""
    n = len(parents)
    children, roots = [[] for _ in range(n)], []
    for i, p in enumerate(parents):
        (roots if p < 0 else children[p]).append(i)
    order, end, cnt = [0] * n, [0] * n, [0]

    def visit(u):
        order[u] = cnt[0]
        cnt[0] += 1
        for c in children[u]:
            visit(c)
        end[u] = cnt[0]

    for r in roots:
        visit(r)
    return np.array([[order[i], end[i]] for i in range(n)], np.int32)


def _paged_case(rng, q_lens, kv_lens, sliding=None):
    from tests.util import make_paged_cache, rand16

    B = len(q_lens)
    qi = np.zeros(B + 1, np.int32)
    qi[1:] = np.cumsum(q_lens)
    c = make_paged_cache(rng, kv_lens, HKV, D, "float16", sliding=sliding)
    q = rand16(rng, (int(qi[-1]), HQ, D), "float16")
    kofs = rng.integers(0, 30, B).astype(np.int32)
    vis = list(kv_lens) if sliding is None else [L - s[0] + s[1] for L, s in zip(kv_lens, sliding)]
    qpos = np.concatenate([kofs[b] + vis[b] + np.arange(q_lens[b]) for b in range(B)]).astype(np.int32)
    return c, q, qi, kofs, qpos


# Each case_* runs one of the reference's compiled kernels on seeded inputs and returns (inputs, O, LSE); the CPU tests
# below compare the oracle with them, tests/test_zzz_ref_kernels_gpu.py compares the CUDA kernels with them.
def case_paged_prefill(mod, causal, rotary_mode):
    rng = np.random.default_rng(21)
    q_lens, kv_lens = [3, 17, 40], [20, 100, 333]
    if causal:  # the query rows are the last q_len cached tokens
        kv_lens = [a + b for a, b in zip(kv_lens, q_lens)]
    c, q, qi, kofs, qpos = _paged_case(rng, q_lens, kv_lens)
    if causal:
        qpos = np.concatenate([kofs[b] + kv_lens[b] - q_lens[b] + np.arange(q_lens[b]) for b in range(3)]).astype(np.int32)
    o, lse = _out(q.shape[0])
    mod["batch_prefill_paged_kv_cpu"](_t(q), _t(qi), _t(c["pages"]), _t(c["page_indptr"]), _t(c["page_values"]),
                                      _t(c["length_info"]), _t(kofs), _t(qpos), o, lse, causal, rotary_mode, 1.0, THETA, SM)
    return dict(c=c, q=q, qi=qi, kofs=kofs, qpos=qpos), o.float().numpy(), lse.numpy()


SLIDING_SLOTS, SLIDING = [70, 300, 40], [(37, 4), (16, 16), (0, 0)]   # (sliding_window_offset, sink_size) per sequence


def case_sliding_decode(mod, rotary_mode):
    rng = np.random.default_rng(22)
    c, q, qi, kofs, qpos = _paged_case(rng, [1, 1, 1], SLIDING_SLOTS, SLIDING)
    qpos = (qpos - 1).astype(np.int32)   # decode: the query is the last visible token
    o, lse = _out(3)
    mod["batch_decode_paged_kv_sliding_window_cpu"](_t(q), _t(c["pages"]), _t(c["page_indptr"]), _t(c["page_values"]),
                                                    _t(c["length_info"]), _t(kofs), _t(qpos), o, lse, rotary_mode, 1.0,
                                                    THETA, SM)
    return dict(c=c, q=q, qi=qi, kofs=kofs, qpos=qpos), o.float().numpy(), lse.numpy()


def case_sliding_prefill(mod, rotary_mode):
    rng = np.random.default_rng(25)
    c, q, qi, kofs, qpos = _paged_case(rng, [5, 9, 2], SLIDING_SLOTS, SLIDING)
    o, lse = _out(q.shape[0])
    # emit_ref_kernels.py builds this flavour with the reference's default layer window (1024): wider than these caches
    mod["batch_prefill_paged_kv_sliding_window_cpu"](_t(q), _t(qi), _t(c["pages"]), _t(c["page_indptr"]),
                                                     _t(c["page_values"]), _t(c["length_info"]), _t(kofs), _t(qpos), o, lse,
                                                     0, rotary_mode, 1.0, THETA, SM)
    return dict(c=c, q=q, qi=qi, kofs=kofs, qpos=qpos), o.float().numpy(), lse.numpy()


TREES = [[-1, 0, 0, 1], [(k - 1) // 2 if k else -1 for k in range(15)], [-1, 0, 1, -1, 3, 3, 0]]


def _tree_arrays():
    masks = np.concatenate([_dfs_mask(t) for t in TREES])
    lens = [len(t) for t in TREES]
    return masks, lens, np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)


def case_tree_ragged(mod):
    """tree_attn_cpu: the new tokens among themselves, token trees given as (dfs order, subtree end) rows."""
    from tests.util import rand16

    rng = np.random.default_rng(23)
    masks, lens, ip = _tree_arrays()
    n = int(ip[-1])
    q, k, v = (rand16(rng, (n, h, D), "float16") for h in (HQ, HKV, HKV))
    qpos = np.concatenate([30 + np.arange(x) for x in lens]).astype(np.int32)
    o, lse = _out(n)
    mod["batch_tree_attn_cpu"](_t(q), _t(ip), _t(k), _t(v), _t(ip), _t(qpos), _t(ip), _t(masks), o, lse, 0, 1.0, THETA, SM)
    return dict(q=q, k=k, v=v, ip=ip, qpos=qpos, masks=masks), o.float().numpy(), lse.numpy()


def case_tree_paged(mod):
    """tree_attn_with_paged_kv_cache_cpu: the tree occupies the trailing columns of each sequence's cached KV."""
    rng = np.random.default_rng(26)
    masks, lens, ip = _tree_arrays()
    kv_lens = [x + extra for x, extra in zip(lens, (20, 200, 0))]
    c, q, qi, kofs, qpos = _paged_case(rng, lens, kv_lens)
    o, lse = _out(q.shape[0])
    mod["tree_attn_paged_kv_cpu"](_t(q), _t(qi), _t(c["pages"]), _t(c["page_indptr"]), _t(c["page_values"]),
                                  _t(c["length_info"]), _t(kofs), _t(qpos), o, lse, 0, 1.0, THETA, SM, _t(ip), _t(masks))
    return dict(c=c, q=q, qi=qi, kofs=kofs, qpos=qpos, ip=ip, masks=masks), o.float().numpy(), lse.numpy()


def _same(o, lse, wo, wl):
    np.testing.assert_allclose(o, wo, atol=2e-3, rtol=1e-2)
    np.testing.assert_allclose(lse, wl, atol=2e-3, rtol=1e-2)


@pytest.mark.parametrize("causal,rotary_mode", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_ref_paged_prefill_matches_oracle(causal, rotary_mode):
    x, o, lse = case_paged_prefill(_need_ref(), causal, rotary_mode)
    c = x["c"]
    _same(o, lse, *ok.attention_prefill_paged(x["q"], x["qi"], c["pages"], c["page_indptr"], c["page_values"], c["length_info"],
                                             x["kofs"], x["qpos"], causal, rotary_mode, 1.0, THETA, SM, "float16"))


@pytest.mark.parametrize("rotary_mode", [0, 1])
def test_ref_sliding_window_decode_and_prefill_match_oracle(rotary_mode):
    """The `_sliding_window` flavours: length_info [3, B] = (last_page_len, sliding_window_offset, sink_size)."""
    mod = _need_ref()
    x, o, lse = case_sliding_decode(mod, rotary_mode)
    c = x["c"]
    _same(o, lse, *ok.attention_decode(x["q"], c["pages"], c["page_indptr"], c["page_values"], c["length_info"], x["kofs"],
                                      x["qpos"], rotary_mode, 1.0, THETA, SM, "float16"))
    x, o, lse = case_sliding_prefill(mod, rotary_mode)
    c = x["c"]
    _same(o, lse, *ok.attention_prefill_paged(x["q"], x["qi"], c["pages"], c["page_indptr"], c["page_values"], c["length_info"],
                                             x["kofs"], x["qpos"], 0, rotary_mode, 1.0, THETA, SM, "float16",
                                             sliding_window_size=1024))


def test_ref_tree_attention_matches_oracle():
    mod = _need_ref()
    x, o, lse = case_tree_ragged(mod)
    _same(o, lse, *ok.attention_prefill_ragged(x["q"], x["ip"], x["k"], x["v"], x["ip"], x["qpos"], None, 0, 0, 1.0, THETA, SM,
                                              "float16", mn_indptr=x["ip"], tree_mask=x["masks"]))
    x, o, lse = case_tree_paged(mod)
    c = x["c"]
    _same(o, lse, *ok.attention_prefill_paged(x["q"], x["qi"], c["pages"], c["page_indptr"], c["page_values"], c["length_info"],
                                             x["kofs"], x["qpos"], 0, 0, 1.0, THETA, SM, "float16", tree_indptr=x["ip"],
                                             tree_order=x["masks"]))


def test_ref_empty_and_boundary_lengths_match_oracle():
    """Edge cases against the reference's own kernels: sequences with no cached KV (O = 0, lse = -5e4), KV lengths around
    the page size, sequences with no query rows, a one-token ragged prefill."""
    mod = _need_ref()
    rng = np.random.default_rng(24)
    kv = [0, 1, 15, 16, 17, 32, 33, 0]
    c, q, qi, kofs, qpos = _paged_case(rng, [1] * len(kv), kv)
    qpos = np.maximum(qpos - 1, 0).astype(np.int32)
    o, lse = _out(len(kv))
    mod["batch_decode_paged_kv_cpu"](_t(q), _t(c["pages"]), _t(c["page_indptr"]), _t(c["page_values"]), _t(c["length_info"]),
                                     _t(kofs), _t(qpos), o, lse, 0, 1.0, THETA, SM)
    wo, wl = ok.attention_decode(q, c["pages"], c["page_indptr"], c["page_values"], c["length_info"], kofs, qpos, 0, 1.0,
                                 THETA, SM, "float16")
    np.testing.assert_allclose(o.float().numpy(), wo, atol=2e-3, rtol=1e-2)
    np.testing.assert_allclose(lse.numpy(), wl, atol=2e-3, rtol=1e-2)
    assert (lse.numpy()[[0, 7]] == -5e4).all() and (o.float().numpy()[[0, 7]] == 0).all()
    q_lens, kv_lens = [3, 0, 5, 2], [0, 20, 16, 1]
    c, q, qi, kofs, qpos = _paged_case(rng, q_lens, kv_lens)
    o, lse = _out(q.shape[0])
    mod["batch_prefill_paged_kv_cpu"](_t(q), _t(qi), _t(c["pages"]), _t(c["page_indptr"]), _t(c["page_values"]),
                                      _t(c["length_info"]), _t(kofs), _t(qpos), o, lse, 0, 0, 1.0, THETA, SM)
    wo, wl = ok.attention_prefill_paged(q, qi, c["pages"], c["page_indptr"], c["page_values"], c["length_info"], kofs, qpos,
                                        0, 0, 1.0, THETA, SM, "float16")
    np.testing.assert_allclose(o.float().numpy(), wo, atol=2e-3, rtol=1e-2)
    np.testing.assert_allclose(lse.numpy(), wl, atol=2e-3, rtol=1e-2)
    assert (lse.numpy()[:3] == -5e4).all()
    inp = _ragged_inputs(seed=6, lens=(1, 0, 16, 17))
    o, lse = ref_ragged_prefill(mod, inp, 1)
    wo, wl = ok.attention_prefill_ragged(inp["q"], inp["ip"], inp["k"], inp["v"], inp["ip"], inp["qpos"], inp["kofs"], 1, 1, 1.0,
                                         THETA, SM, "float16")
    np.testing.assert_allclose(o, wo, atol=2e-3, rtol=1e-2)
    np.testing.assert_allclose(lse, wl, atol=2e-3, rtol=1e-2)
