"""GPU parity at BASELINE.json's full configuration sizes (C2..C5), through the C ABI.

The oracle (NumPy, float64) cannot restate a whole full-size call in seconds, so each test runs the CUDA path at
the FULL size and checks it with size-independent properties of the domain:
  * a sample of sequences / rows is restated by the oracle on exactly the same page table and inputs (attention
    rows are independent of one another: any subset is a complete check of those rows);
  * causal prefix property: the first n tokens of a causal prefill do not depend on the tokens behind them, so the
    full-size output restricted to a prefix must equal the oracle's prefill of that prefix alone;
  * batch-composition invariance: a sequence's decode output must not depend on which other sequences share the
    launch (the split-KV plan changes with the batch), within the fp tolerance.
Tolerance: north_star's max-abs 2e-3 / rtol 1e-2 (tests/util.py::assert_close)."""
import numpy as np
import pytest

from oracle import kernels as ok
from tests.test_kernels_gpu import _dfs_mask, _i32
from tests.util import assert_close, make_paged_cache, rand16, to_dev, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture()
def capi(built_lib):
    from tvm_b200 import capi as c

    c.lib()
    return c


class _GpuCache:
    """Full-size paged cache generated ON the device (a C5 cache is 4 GiB; NumPy would need 17 GiB of float32 and a
    minute of RNG); `sub(b)` brings the pages of one sequence to the host as the oracle's compact sub-cache."""

    def __init__(self, rng, B, L, hkv, d, dtype, seed):
        import torch

        from tests.util import torch_dtype

        self.B, self.L = B, L
        ppseq = -(-L // 16)
        total = B * ppseq + 3
        g = torch.Generator(device="cuda")
        g.manual_seed(seed)
        self.pages = torch.randn((total, 2, hkv, 16, d), generator=g, device="cuda", dtype=torch_dtype(dtype))
        perm = rng.permutation(total).astype(np.int32)
        self.page_values = perm[: B * ppseq].copy()
        self.page_indptr = (np.arange(B + 1) * ppseq).astype(np.int32)
        self.length_info = np.full(B, ((L - 1) % 16) + 1, np.int32)

    def sub(self, b):
        import torch

        ids = self.page_values[self.page_indptr[b]:self.page_indptr[b + 1]]
        host = self.pages[torch.from_numpy(ids.astype(np.int64)).cuda()].float().cpu().numpy()
        return dict(pages=host, page_indptr=np.array([0, len(ids)], np.int32),
                    page_values=np.arange(len(ids), dtype=np.int32), length_info=self.length_info[b:b + 1])

    def as_dict(self):
        return dict(pages=self.pages, page_indptr=self.page_indptr, page_values=self.page_values,
                    length_info=self.length_info)


def _decode_full_vs_sample(capi, rng, B, L, hq, hkv, dtype, sample):
    import torch

    d = 128
    gc = _GpuCache(rng, B, L, hkv, d, dtype, seed=int(rng.integers(1 << 30)))
    c = gc.as_dict()
    q = rand16(rng, (B, hq, d), dtype)
    kpos = np.zeros(B, np.int32)
    qpos = np.full(B, L - 1, np.int32)
    sm = d ** -0.5
    dq, dpages = to_dev(q, dtype), c["pages"]
    o = torch.full((B, hq, d), float("nan"), dtype=dq.dtype, device="cuda")
    lse = torch.full((B, hq), float("nan"), dtype=torch.float32, device="cuda")
    capi.attention_decode(dq, dpages, _i32(c["page_indptr"]), _i32(c["page_values"]), _i32(c["length_info"]),
                          _i32(kpos), _i32(qpos), o, lse, 0, 1.0, 5e5, sm)
    torch.cuda.synchronize()
    go, gl = to_np(o), to_np(lse)
    assert np.isfinite(go).all() and np.isfinite(gl).all()
    # (1) sampled sequences restated by the oracle on the same page table
    for b in sample:
        sc = gc.sub(b)
        wo, wl = ok.attention_decode(q[b:b + 1], sc["pages"], sc["page_indptr"], sc["page_values"], sc["length_info"],
                                     kpos[b:b + 1], qpos[b:b + 1], 0, 1.0, 5e5, sm, dtype)
        assert_close(f"decode O seq {b}", go[b:b + 1], wo)
        assert_close(f"decode LSE seq {b}", gl[b:b + 1], wl)
    # (2) batch-composition invariance: the same sequences launched alone (different split-KV plan)
    sel = np.array(sample)
    ip = c["page_indptr"]
    vals = np.concatenate([c["page_values"][ip[b]:ip[b + 1]] for b in sel])
    sip = np.zeros(len(sel) + 1, np.int32)
    sip[1:] = np.cumsum([ip[b + 1] - ip[b] for b in sel])
    o2 = torch.empty((len(sel), hq, d), dtype=dq.dtype, device="cuda")
    lse2 = torch.empty((len(sel), hq), dtype=torch.float32, device="cuda")
    capi.attention_decode(to_dev(q[sel], dtype), dpages, _i32(sip), _i32(vals), _i32(c["length_info"][sel]),
                          _i32(kpos[sel]), _i32(qpos[sel]), o2, lse2, 0, 1.0, 5e5, sm)
    torch.cuda.synchronize()
    assert_close("decode O alone vs in batch", to_np(o2), go[sel])
    assert_close("decode LSE alone vs in batch", to_np(lse2), gl[sel])


def test_c2_decode_batch64_ctx4096_bf16(capi):
    """C2: Llama-3-8B attention shape, batch 64 decode at 4K context, bf16 paged KV."""
    _decode_full_vs_sample(capi, np.random.default_rng(100), 64, 4096, 32, 8, "bfloat16", [0, 17, 63])


@pytest.mark.parametrize("tp", [2, 8])
def test_c4_decode_70b_head_shard(capi, tp):
    """C4: Llama-3-70B GQA (64 q / 8 kv heads) decode at 8K context, one rank's KV-head group of a tp-way shard
    (batch 32 of the 256 here: the per-rank page table is the same for every batch slice)."""
    _decode_full_vs_sample(capi, np.random.default_rng(101 + tp), 32, 8192, 64 // tp, 8 // tp, "bfloat16", [3, 31])


def test_c5_decode_ctx32k_batch32(capi):
    """C5 decode half: split-KV decode at 32K context, batch 32."""
    _decode_full_vs_sample(capi, np.random.default_rng(105), 32, 32768, 32, 8, "bfloat16", [5, 30])


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_c3_ragged_prefill_16x2048_prefix_property(capi, dtype):
    """C3: ragged causal prefill 16 x 2048 at full size (tcgen05 path by auto dispatch).  Causal prefix property:
    rows [0, 192) of sequences 0, 7, 15 must equal the oracle's causal prefill of those 192 tokens alone; and the LAST
    64 rows of sequence 9 (which see the whole 2048-token context) are restated directly."""
    import torch

    rng = np.random.default_rng(106)
    nseq, L, hq, hkv, d = 16, 2048, 32, 8, 128
    n = nseq * L
    q, k, v = rand16(rng, (n, hq, d), dtype), rand16(rng, (n, hkv, d), dtype), rand16(rng, (n, hkv, d), dtype)
    ip = (np.arange(nseq + 1) * L).astype(np.int32)
    qpos = np.tile(np.arange(L, dtype=np.int32), nseq)
    kofs = np.zeros(nseq, np.int32)
    sm = d ** -0.5
    dq = to_dev(q, dtype)
    o = torch.full((n, hq, d), float("nan"), dtype=dq.dtype, device="cuda")
    lse = torch.full((n, hq), float("nan"), dtype=torch.float32, device="cuda")
    n0 = capi.launch_count()
    capi.attention_prefill_ragged(dq, _i32(ip), to_dev(k, dtype), to_dev(v, dtype), _i32(ip), _i32(qpos), _i32(kofs),
                                  o, lse, 1, 0, 1.0, 5e5, sm)
    torch.cuda.synchronize()
    assert capi.launch_count() - n0 == 1
    go, gl = to_np(o), to_np(lse)
    assert np.isfinite(go).all() and np.isfinite(gl).all()
    P = 192
    sip = np.array([0, P], np.int32)
    for b in (0, 7, 15):
        s0 = b * L
        wo, wl = ok.attention_prefill_ragged(q[s0:s0 + P], sip, k[s0:s0 + P], v[s0:s0 + P], sip, qpos[:P], kofs[:1], 1, 0,
                                             1.0, 5e5, sm, dtype)
        assert_close(f"prefix O seq {b}", go[s0:s0 + P], wo)
        assert_close(f"prefix LSE seq {b}", gl[s0:s0 + P], wl)
    # last 64 rows of sequence 9: q rows [L-64, L) against all L keys (causal offset = kv_len - q_len)
    b, T = 9, 64
    s0 = b * L
    wo, wl = ok.attention_prefill_ragged(q[s0 + L - T:s0 + L], np.array([0, T], np.int32), k[s0:s0 + L], v[s0:s0 + L],
                                         np.array([0, L], np.int32), qpos[L - T:L], kofs[:1], 1, 0, 1.0, 5e5, sm, dtype)
    assert_close("tail O seq 9", go[s0 + L - T:s0 + L], wo)
    assert_close("tail LSE seq 9", gl[s0 + L - T:s0 + L], wl)
    # ... and 96 rows picked at random over all sequences and positions (mid-sequence rows of interior work items, tile
    # borders included): each restated by the oracle as a one-row causal query over its own prefix
    picks = [(int(rng.integers(0, nseq)), int(t)) for t in
             list(rng.integers(1, L, 80)) + [127, 128, 129, 255, 256, 1023, 1024, 1025, 2047, 63, 64, 65, 511, 512, 1535, 1536]]
    one = np.array([0, 1], np.int32)
    for b, t in picks:
        s0 = b * L
        wo, wl = ok.attention_prefill_ragged(q[s0 + t:s0 + t + 1], one, k[s0:s0 + t + 1], v[s0:s0 + t + 1],
                                             np.array([0, t + 1], np.int32), qpos[t:t + 1], kofs[:1], 1, 0, 1.0, 5e5, sm, dtype)
        assert_close(f"row O seq {b} pos {t}", go[s0 + t:s0 + t + 1], wo)
        assert_close(f"row LSE seq {b} pos {t}", gl[s0 + t:s0 + t + 1], wl)


def _random_tree(rng, n):
    """parent array of a random n-node token tree in topological (parent before child) order"""
    return [-1] + [int(rng.integers(0, i)) for i in range(1, n)]


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_c5_tree_prefill_64_node_trees(capi, dtype):
    """C5 prefill half: speculative tree attention with 64-node trees, batch 32.  The cache applies the tree mask on
    the ragged self-attention of the new nodes (f_attention_prefill_ragged with tree mask) and a mask-free paged
    prefill over the committed context, then merges: both callbacks at C5's batch / tree size, restated by the oracle."""
    import torch

    rng = np.random.default_rng(107)
    B, nodes, hq, hkv, d = 32, 64, 32, 8, 128
    trees = [_random_tree(rng, nodes) for _ in range(B)]
    trees[0] = list(range(-1, nodes - 1))                                   # a chain (= causal)
    trees[1] = [-1] + [0] * (nodes - 1)                                     # a star
    trees[2] = [(i - 1) // 2 if i else -1 for i in range(nodes)]            # a complete binary tree
    masks = np.concatenate([_dfs_mask(t) for t in trees])
    mn = (np.arange(B + 1) * nodes).astype(np.int32)
    n = B * nodes
    q, k, v = rand16(rng, (n, hq, d), dtype), rand16(rng, (n, hkv, d), dtype), rand16(rng, (n, hkv, d), dtype)
    qpos = np.concatenate([1000 + np.array(_depths(t)) for t in trees]).astype(np.int32)
    sm = d ** -0.5
    dq = to_dev(q, dtype)
    o = torch.full((n, hq, d), float("nan"), dtype=dq.dtype, device="cuda")
    lse = torch.full((n, hq), float("nan"), dtype=torch.float32, device="cuda")
    capi.attention_prefill_tree_ragged(dq, _i32(mn), to_dev(k, dtype), to_dev(v, dtype), _i32(mn), _i32(qpos), _i32(mn),
                                       _i32(masks), o, lse, 0, 1.0, 5e5, sm)
    torch.cuda.synchronize()
    wo, wl = ok.attention_prefill_ragged(q, mn, k, v, mn, qpos, np.zeros(B, np.int32), 0, 0, 1.0, 5e5, sm, dtype,
                                         mn_indptr=mn, tree_mask=masks)
    assert_close("tree ragged O", to_np(o), wo)
    assert_close("tree ragged LSE", to_np(lse), wl)
    # mask-free paged prefill of the 64 nodes over a committed context (4 of the 32 sequences at 4K context here; the
    # 32K-context case is covered by test_c5_paged_prefill_ctx32k below)
    Bs = 4
    kv_lens = [4096, 1000, 17, 2048]
    c = make_paged_cache(rng, kv_lens, hkv, d, dtype)
    qi = (np.arange(Bs + 1) * nodes).astype(np.int32)
    q2 = q[: Bs * nodes]
    qpos2 = np.concatenate([kv_lens[b] + np.array(_depths(trees[b])) for b in range(Bs)]).astype(np.int32)
    kofs = np.zeros(Bs, np.int32)
    o2 = torch.full((Bs * nodes, hq, d), float("nan"), dtype=dq.dtype, device="cuda")
    lse2 = torch.full((Bs * nodes, hq), float("nan"), dtype=torch.float32, device="cuda")
    capi.attention_prefill_paged(to_dev(q2, dtype), _i32(qi), to_dev(c["pages"], dtype), _i32(c["page_indptr"]),
                                 _i32(c["page_values"]), _i32(c["length_info"]), _i32(kofs), _i32(qpos2), o2, lse2, 0, 0,
                                 1.0, 5e5, sm)
    torch.cuda.synchronize()
    wo2, wl2 = ok.attention_prefill_paged(q2, qi, c["pages"], c["page_indptr"], c["page_values"], c["length_info"], kofs,
                                          qpos2, 0, 0, 1.0, 5e5, sm, dtype)
    assert_close("tree context O", to_np(o2), wo2)
    assert_close("tree context LSE", to_np(lse2), wl2)
    # the merge of the two partial results is f_merge_inplace, covered bit-for-bit by the golden replay (tree_attn)


def _depths(parents):
    dep = []
    for i, p in enumerate(parents):
        dep.append(0 if p < 0 else dep[p] + 1)
    return dep


def test_c5_paged_prefill_ctx32k(capi):
    """C5: the 64 tree nodes of a sequence against a 32K-token committed context (mask-free paged prefill on the
    tcgen05 path), batch 32 at full size; two sampled sequences restated by the oracle."""
    import torch

    rng = np.random.default_rng(108)
    B, nodes, L, hq, hkv, d, dtype = 32, 64, 32768, 32, 8, 128, "bfloat16"
    gc = _GpuCache(rng, B, L, hkv, d, dtype, seed=7)
    c = gc.as_dict()
    n = B * nodes
    q = rand16(rng, (n, hq, d), dtype)
    qi = (np.arange(B + 1) * nodes).astype(np.int32)
    qpos = np.tile(L + np.arange(nodes, dtype=np.int32), B)
    kofs = np.zeros(B, np.int32)
    sm = d ** -0.5
    dq = to_dev(q, dtype)
    o = torch.full((n, hq, d), float("nan"), dtype=dq.dtype, device="cuda")
    lse = torch.full((n, hq), float("nan"), dtype=torch.float32, device="cuda")
    capi.attention_prefill_paged(dq, _i32(qi), c["pages"], _i32(c["page_indptr"]), _i32(c["page_values"]),
                                 _i32(c["length_info"]), _i32(kofs), _i32(qpos), o, lse, 0, 0, 1.0, 5e5, sm)
    torch.cuda.synchronize()
    go, gl = to_np(o), to_np(lse)
    assert np.isfinite(go).all() and np.isfinite(gl).all()
    for b in (2, 29):
        sc = gc.sub(b)
        rows = slice(b * nodes, (b + 1) * nodes)
        wo, wl = ok.attention_prefill_paged(q[rows], np.array([0, nodes], np.int32), sc["pages"], sc["page_indptr"],
                                            sc["page_values"], sc["length_info"], kofs[:1], qpos[rows], 0, 0, 1.0, 5e5,
                                            sm, dtype)
        assert_close(f"ctx32k O seq {b}", go[rows], wo)
        assert_close(f"ctx32k LSE seq {b}", gl[rows], wl)
