"""GPU parity THROUGH THE DROP-IN BOUNDARY: the 13 tvm-ffi packed callbacks (and their TIR-named twins) called the way
the reference's C++ cache calls them --

  * positional signatures of src/runtime/vm/attn_backend.h:234-243 (paged prefill), :388-396 (ragged prefill), :507-515
    (decode), :618-627 (tree paged), :665-673 (tree ragged) and paged_kv_cache.cc:1360-1373, 728, 759, 1718, 2292;
  * every int32 array is a `byte_offset` VIEW of one merged device buffer at a 16-byte-aligned offset
    (CachedPagedKVCacheAuxDataManager, attn_utils.h:1027-1052), q / k / v / o are views of larger temp buffers
    (paged_kv_cache.cc:1340-1345) -- a callback that ignored DLTensor.byte_offset would read the guard pattern;
  * kernels launch on the tvm-ffi ENVIRONMENT stream (TVMFFIEnvGetStream), here a non-default torch stream whose only
    ordering against the input upload is stream order -- a callback on any other stream reads stale inputs;
  * shapes the packed layer infers itself (batch from page_indptr, nnz_pages from page_values, num_pages from pages,
    the sliding flavour from length_info.ndim) are exercised with values that differ from each other.

Checked against the CPU oracle (oracle/kernels.py): bit-exact for copies / index work, max-abs 2e-3 / rtol 1e-2 for O and
LSE.  Also here: two kernel sets (contexts) with different rope scalings interleaved on two streams (re-entrancy)."""
import numpy as np
import pytest

from oracle import kernels as ok
from tests.util import MergedAux, assert_close, ffi_view, make_paged_cache, rand16, to_np

pytestmark = pytest.mark.gpu

DTYPES = ["float16", "bfloat16"]
HQ, HKV, D = 32, 8, 128
SM = D ** -0.5


@pytest.fixture(scope="module")
def mod(built_lib):
    from tvm_b200 import ffi

    return ffi.module()


class Arena:
    """One device temp buffer per dtype; tensors are carved out of it as byte_offset views (never at offset 0), the way
    the reference's cache hands out views of temp_attn_{q,k,v,output}_device_.  Uploads run on `stream`."""

    def __init__(self, stream, nbytes=96 << 20):
        import torch

        self.stream = stream
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        self.buf.fill_(0x7F)
        torch.cuda.synchronize()
        self.off = 256
        self.keep = []

    def _take(self, nbytes):
        off = self.off
        self.off += (nbytes + 255) // 256 * 256 + 256
        assert self.off <= self.buf.numel(), "arena too small"
        return off

    def put(self, x, dtype):
        """Upload numpy `x` (already rounded to `dtype`) behind whatever the stream is doing; returns (ffi view, torch view)."""
        import torch

        tdt = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32, "int32": torch.int32}[dtype]
        host = torch.from_numpy(np.ascontiguousarray(x)).to(tdt).pin_memory()
        off = self._take(host.numel() * host.element_size())
        tv = self.buf[off: off + host.numel() * host.element_size()].view(tdt).view(host.shape)
        with torch.cuda.stream(self.stream):
            tv.copy_(host, non_blocking=True)
        self.keep.append(host)
        return ffi_view(self.buf, off, host.shape, dtype), tv

    def empty(self, shape, dtype):
        import torch

        tdt = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}[dtype]
        n = int(np.prod(shape)) * tdt.itemsize
        off = self._take(n)
        tv = self.buf[off: off + n].view(tdt).view(tuple(shape))
        return ffi_view(self.buf, off, shape, dtype), tv


@pytest.fixture()
def env(mod):
    """(module, arena, stream context): a fresh non-default stream that is busy for a while before every upload."""
    import torch
    import tvm_ffi

    s = torch.cuda.Stream()
    arena = Arena(s)                   # (its constructor synchronises the device)
    with torch.cuda.stream(s):
        torch.cuda._sleep(20_000_000)  # ~10 ms: uploads and kernels queued behind it are ordered by the stream alone

    class Env:
        pass

    e = Env()
    e.mod, e.arena, e.stream = mod, arena, s
    e.scope = lambda: tvm_ffi.use_torch_stream(torch.cuda.stream(s))
    e.sync = lambda: s.synchronize()
    return e


def _aux(**arrays):
    return MergedAux(arrays)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["f_transpose_append", "tir_kv_cache_transpose_append"])
def test_append_and_debug_get_kv(env, dtype, name):
    rng = np.random.default_rng(0)
    P, n = 37, 100
    pages = rand16(rng, (P, 2, HKV, 16, D), dtype)
    k, v = rand16(rng, (n, HKV, D), dtype), rand16(rng, (n, HKV, D), dtype)
    pm = rng.permutation(P * 16)[:n].astype(np.int32)
    pm[::7] = -1
    live = pm[pm >= 0]
    want = pages.copy()
    ok.transpose_append(want, k, v, pm)
    a = env.arena
    fp, tp = a.put(pages, dtype)
    fk, _ = a.put(k, dtype)
    fv, _ = a.put(v, dtype)
    aux = _aux(pm=pm, live=live)
    fko, tko = a.empty((2, len(live), HKV, D), dtype)
    fvo, tvo = a.empty((2, len(live), HKV, D), dtype)
    with env.scope():
        env.mod[name](fp, fk, fv, aux["pm"])
        env.mod["f_debug_get_kv" if name.startswith("f_") else "tir_kv_cache_debug_get_kv"](fp, aux["live"], fko, fvo, 1)
    env.sync()
    assert np.array_equal(ok.to_bits16(to_np(tp), dtype), ok.to_bits16(want, dtype))
    wk, wv = ok.debug_get_kv(want, live)
    assert np.array_equal(to_np(tko[1]), wk) and np.array_equal(to_np(tvo[1]), wv)


@pytest.mark.parametrize("dtype", DTYPES)
def test_copy_single_page_and_compact_copy(env, dtype):
    rng = np.random.default_rng(1)
    pages = rand16(rng, (9, 2, HKV, 16, D), dtype)
    want = pages.copy()
    fp, tp = env.arena.put(pages, dtype)
    indptr = np.array([0, 3, 3, 5], np.int32)
    src_dst = np.array([[20, 21, 5, 100, 33], [5, 20, 6, 101, 34]], np.int32)
    aux = _aux(indptr=indptr, src_dst=src_dst)
    with env.scope():
        for src, tgt, ln in [(2, 3, 2), (0, 8, 16), (5, 1, 0), (4, 6, 15)]:
            ok.copy_single_page(want, src, tgt, ln)
            env.mod["f_copy_single_page"](fp, src, tgt, ln)
        ok.compact_kv_copy(want, indptr, src_dst, 3)
        env.mod["f_compact_copy"](fp, aux["indptr"], aux["src_dst"], 3)
    env.sync()
    assert np.array_equal(to_np(tp), want)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("theta", [1e4, 5e5])
def test_split_rotary_reference_signature(env, dtype, theta):
    """6 positional arguments like the reference's fused_rope; theta / scale come from the kernel set they are bound to."""
    from tvm_b200 import ffi

    rng = np.random.default_rng(2)
    n = 53
    qkv = rand16(rng, (n, HQ + 2 * HKV, D), dtype)
    pos = rng.integers(0, 4096, n).astype(np.int32)
    ks = ffi.KernelSet(rope_theta=theta)
    a = env.arena
    fqkv, _ = a.put(qkv, dtype)
    aux = _aux(pos=pos)
    outs = [a.empty((n, h, D), dtype) for h in (HQ, HKV, HKV)]
    for apply_rope in (1, 0):
        wq, wk, wv = ok.split_rotary(qkv, pos, HQ, HKV, apply_rope, theta, 1.0, dtype)
        with env.scope():
            ks["f_split_rotary"](fqkv, aux["pos"], outs[0][0], outs[1][0], outs[2][0], apply_rope)
        env.sync()
        assert np.array_equal(to_np(outs[2][1]), wv)
        if apply_rope == 0:
            assert np.array_equal(to_np(outs[0][1]), wq) and np.array_equal(to_np(outs[1][1]), wk)
        else:
            assert_close("q", to_np(outs[0][1]), wq, atol=4e-3 if dtype == "float16" else 3.2e-2)
            assert_close("k", to_np(outs[1][1]), wk, atol=4e-3 if dtype == "float16" else 3.2e-2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_merge_inplace(env, dtype):
    rng = np.random.default_rng(3)
    N, H = 77, HQ
    v, vo = rand16(rng, (N, H, D), dtype), rand16(rng, (N, H, D), dtype)
    s = rng.standard_normal((N, H)).astype(np.float32) * 3
    so = rng.standard_normal((N, H)).astype(np.float32) * 3
    so[::5] = -5e4  # an empty partner (the reference's "no cached KV" result) must be a no-op
    wv, ws = ok.merge_state_inplace(v, s, vo, so, dtype)
    a = env.arena
    fv, tv = a.put(v, dtype)
    fs, ts = a.put(s, "float32")
    fvo, _ = a.put(vo, dtype)
    fso, _ = a.put(so, "float32")
    with env.scope():
        env.mod["f_merge_inplace"](fv, fs, fvo, fso)
    env.sync()
    assert_close("v", to_np(tv), wv)
    assert_close("s", to_np(ts), ws)
    assert np.array_equal(to_np(tv)[::5], v[::5])


# ---------------------------------------------------------------------------------------------------------------------
def _decode_case(env, dtype, name, kv_lens, sliding=None, rotary_mode=0, extra_pages=5):
    rng = np.random.default_rng(4)
    B = len(kv_lens)
    c = make_paged_cache(rng, kv_lens, HKV, D, dtype, extra_pages=extra_pages, sliding=sliding)
    q = rand16(rng, (B, HQ, D), dtype)
    kro = rng.integers(0, 64, B).astype(np.int32)
    qpos = (kro + np.array(kv_lens)).astype(np.int32)
    wo, wl = ok.attention_decode(q, c["pages"], c["page_indptr"], c["page_values"], c["length_info"], kro, qpos,
                                 rotary_mode, 1.0, 1e4, SM, dtype)
    a = env.arena
    fq, _ = a.put(q, dtype)
    fp, _ = a.put(c["pages"], dtype)
    aux = _aux(indptr=c["page_indptr"], values=c["page_values"], li=c["length_info"], kro=kro, qpos=qpos)
    fo, to = a.empty((B, HQ, D), dtype)
    fl, tl = a.empty((B, HQ), "float32")
    with env.scope():
        env.mod[name](fq, fp, aux["indptr"], aux["values"], aux["li"], aux["kro"], aux["qpos"], fo, fl, rotary_mode,
                      1.0, 1e4, SM)
    env.sync()
    assert_close("o", to_np(to), wo)
    assert_close("lse", to_np(tl), wl)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["f_attention_decode", "batch_decode_paged_kv"])
def test_decode(env, dtype, name):
    # ragged lengths, an empty sequence, one long enough to be split over several CTAs (split-KV + merge)
    _decode_case(env, dtype, name, [1, 16, 17, 0, 700, 33, 2049])


@pytest.mark.parametrize("dtype", DTYPES)
def test_decode_inline_rope(env, dtype):
    _decode_case(env, dtype, "f_attention_decode", [40, 7, 300], rotary_mode=1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["f_attention_decode_sliding_window", "batch_decode_paged_kv_sliding_window"])
def test_decode_sliding_window_flavour(env, dtype, name):
    # length_info [3, B] = (last_page_len, sliding_offset, sink) -> the packed layer picks the flavour from ndim
    _decode_case(env, dtype, name, [64, 100, 37], sliding=[(0, 0), (19, 4), (5, 5)], rotary_mode=1)


def _prefill_paged_case(env, dtype, name, q_lens, kv_lens, causal, sliding=None, rotary_mode=0, tree=None):
    rng = np.random.default_rng(5)
    B = len(q_lens)
    c = make_paged_cache(rng, kv_lens, HKV, D, dtype, sliding=sliding)
    qi = np.concatenate([[0], np.cumsum(q_lens)]).astype(np.int32)
    n = int(qi[-1])
    q = rand16(rng, (n, HQ, D), dtype)
    kro = rng.integers(0, 32, B).astype(np.int32)
    qpos = np.concatenate([kro[b] + kv_lens[b] - q_lens[b] + np.arange(q_lens[b]) for b in range(B)]).astype(np.int32)
    kw = {}
    if tree is not None:
        kw = dict(tree_indptr=tree[0], tree_order=tree[1])
    wo, wl = ok.attention_prefill_paged(q, qi, c["pages"], c["page_indptr"], c["page_values"], c["length_info"], kro,
                                        qpos, causal, rotary_mode, 1.0, 1e4, SM, dtype, **kw)
    a = env.arena
    fq, _ = a.put(q, dtype)
    fp, _ = a.put(c["pages"], dtype)
    arrays = dict(qi=qi, indptr=c["page_indptr"], values=c["page_values"], li=c["length_info"], kro=kro, qpos=qpos)
    if tree is not None:
        arrays.update(ti=tree[0], to=tree[1])
    aux = _aux(**arrays)
    fo, to = a.empty((n, HQ, D), dtype)
    fl, tl = a.empty((n, HQ), "float32")
    with env.scope():
        if tree is None:
            env.mod[name](fq, aux["qi"], fp, aux["indptr"], aux["values"], aux["li"], aux["kro"], aux["qpos"], fo, fl,
                          causal, rotary_mode, 1.0, 1e4, SM)
        else:
            env.mod[name](fq, aux["qi"], fp, aux["indptr"], aux["values"], aux["li"], aux["kro"], aux["qpos"], fo, fl,
                          rotary_mode, 1.0, 1e4, SM, aux["ti"], aux["to"])
    env.sync()
    assert_close("o", to_np(to), wo)
    assert_close("lse", to_np(tl), wl)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("causal", [0, 1])
@pytest.mark.parametrize("name", ["f_attention_prefill", "batch_prefill_paged_kv"])
def test_prefill_paged(env, dtype, causal, name):
    _prefill_paged_case(env, dtype, name, [5, 64, 1, 130], [40, 64, 17, 450], causal)


@pytest.mark.parametrize("dtype", DTYPES)
def test_prefill_paged_tcgen05_size(env, dtype):
    """Enough folded rows (n * group >= 2048) for the packed call to take the tcgen05 path."""
    _prefill_paged_case(env, dtype, "f_attention_prefill", [300, 260], [300 + 128, 260 + 517], 0)
    _prefill_paged_case(env, dtype, "f_attention_prefill", [300, 260], [300 + 128, 260 + 517], 1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["f_attention_prefill_sliding_window", "batch_prefill_paged_kv_sliding_window"])
def test_prefill_paged_sliding_window_flavour(env, dtype, name):
    _prefill_paged_case(env, dtype, name, [3, 20], [70, 120], 0, sliding=[(9, 2), (0, 0)], rotary_mode=1)


def _random_tree(rng, n):
    """(order, subtree_end) per node of a random tree in the host's DFS numbering (paged_kv_cache.cc:1900-1918)."""
    parent = [-1] + [int(rng.integers(-1 if k > 3 else 0, k)) for k in range(1, n)]
    children = [[] for _ in range(n)]
    roots = []
    for k, p in enumerate(parent):
        (roots if p == -1 else children[p]).append(k)
    iv = np.zeros((n, 2), np.int32)
    order = [0]

    def dfs(u):
        iv[u, 0] = order[0]
        order[0] += 1
        ub = iv[u, 0] + 1
        for ch in children[u]:
            ub = max(ub, dfs(ch))
        iv[u, 1] = ub
        return ub

    for r in roots:
        dfs(r)
    return iv


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["f_attention_prefill_with_tree_mask_paged_kv", "tree_attn_paged_kv"])
def test_tree_paged(env, dtype, name):
    rng = np.random.default_rng(6)
    sizes = [7, 64, 20]
    ti = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    to = np.concatenate([_random_tree(rng, s) for s in sizes]).astype(np.int32)
    # the tree nodes are the LAST tree_size columns of each sequence's KV (the tree was appended before the attention)
    _prefill_paged_case(env, dtype, name, sizes, [7 + 33, 64 + 300, 20], 0, tree=(ti, to))


def _prefill_ragged_case(env, dtype, name, lens, causal=1, rotary_mode=0, tree=None):
    rng = np.random.default_rng(7)
    B = len(lens)
    qi = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n = int(qi[-1])
    q, k, v = (rand16(rng, (n, h, D), dtype) for h in (HQ, HKV, HKV))
    kro = rng.integers(0, 32, B).astype(np.int32)
    qpos = np.concatenate([kro[b] + np.arange(lens[b]) for b in range(B)]).astype(np.int32)
    kw = {}
    if tree is not None:
        kw = dict(mn_indptr=tree[0], tree_mask=tree[1])
    wo, wl = ok.attention_prefill_ragged(q, qi, k, v, qi, qpos, kro, causal, rotary_mode, 1.0, 1e4, SM, dtype, **kw)
    a = env.arena
    fq, _ = a.put(q, dtype)
    fk, _ = a.put(k, dtype)
    fv, _ = a.put(v, dtype)
    arrays = dict(qi=qi, ki=qi.copy(), qpos=qpos, kro=kro)
    if tree is not None:
        arrays.update(mn=tree[0], mask=tree[1])
    aux = _aux(**arrays)
    fo, to = a.empty((n, HQ, D), dtype)
    fl, tl = a.empty((n, HQ), "float32")
    with env.scope():
        if tree is None:
            env.mod[name](fq, aux["qi"], fk, fv, aux["ki"], aux["qpos"], aux["kro"], fo, fl, causal, rotary_mode, 1.0,
                          1e4, SM)
        else:
            env.mod[name](fq, aux["qi"], fk, fv, aux["ki"], aux["qpos"], aux["mn"], aux["mask"], fo, fl, rotary_mode,
                          1.0, 1e4, SM)
    env.sync()
    assert_close("o", to_np(to), wo)
    assert_close("lse", to_np(tl), wl)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["f_attention_prefill_ragged", "batch_prefill_ragged_kv"])
def test_prefill_ragged(env, dtype, name):
    _prefill_ragged_case(env, dtype, name, [10, 20, 30, 40])            # the reference's C1 scenario (generic path)
    _prefill_ragged_case(env, dtype, name, [257, 1, 300, 128])          # >= 2048 folded rows: tcgen05 path


@pytest.mark.parametrize("dtype", DTYPES)
def test_prefill_ragged_inline_rope(env, dtype):
    _prefill_ragged_case(env, dtype, "f_attention_prefill_ragged", [33, 5, 70], rotary_mode=1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["f_attention_prefill_with_tree_mask", "batch_tree_attn"])
def test_tree_ragged(env, dtype, name):
    rng = np.random.default_rng(8)
    sizes = [7, 64, 20, 1]
    mn = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    mask = np.concatenate([_random_tree(rng, s) for s in sizes]).astype(np.int32)
    _prefill_ragged_case(env, dtype, name, sizes, tree=(mn, mask))


# ---------------------------------------------------------------------------------------------------------------------
def test_stale_input_proves_the_stream_matters(mod):
    """Control experiment for the tests above: the same call made on the WRONG stream (the env stream left at the legacy
    default while the upload is queued behind a sleeping side stream) must NOT see the uploaded data -- i.e. the parity
    tests really depend on the packed functions launching on TVMFFIEnvGetStream."""
    import torch
    import tvm_ffi

    s = torch.cuda.Stream()
    arena = Arena(s, nbytes=8 << 20)
    with torch.cuda.stream(s):
        torch.cuda._sleep(400_000_000)  # ~0.2 s
    rng = np.random.default_rng(9)
    v = rand16(rng, (64, HQ, D), "float16")
    s_ = np.zeros((64, HQ), np.float32)
    fv, tv = arena.put(np.zeros_like(v), "float16")
    fs, _ = arena.put(s_, "float32")
    fvo, _ = arena.put(v, "float16")
    fso, _ = arena.put(s_ + 20.0, "float32")           # the partner dominates: v <- v_other
    other = torch.cuda.Stream()
    with tvm_ffi.use_torch_stream(torch.cuda.stream(other)):
        mod["f_merge_inplace"](fv, fs, fvo, fso)       # runs at once on `other`: inputs are still the 0x7F fill
    other.synchronize()
    early = to_np(tv).copy()
    s.synchronize()
    assert not np.allclose(early, v, atol=1e-2), "the merge saw data that had not been uploaded yet"


def test_two_kernel_sets_interleaved_on_two_streams(mod):
    """Re-entrancy (callbacks are re-entrant per device): two kernel sets with DIFFERENT rope scalings
    (default vs llama3) and split-KV decodes with different data, interleaved call by call on two streams.  Settings and
    scratch (split-KV partials, counters) are per kernel set and per stream, so nothing bleeds across."""
    import torch
    import tvm_ffi

    from tvm_b200 import ffi

    llama3 = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
              "original_max_position_embeddings": 8192}
    sets = [ffi.KernelSet(rope_theta=5e5), ffi.KernelSet(rope_theta=5e5, rope_scaling=llama3)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    arenas = [Arena(s, nbytes=64 << 20) for s in streams]
    dtype = "bfloat16"
    n = 64
    work = []
    for i in range(2):
        rng = np.random.default_rng(20 + i)
        qkv = rand16(rng, (n, HQ + 2 * HKV, D), dtype)
        pos = rng.integers(0, 100_000, n).astype(np.int32)
        kv_lens = [1500 + 300 * i, 900, 2100]
        c = make_paged_cache(rng, kv_lens, HKV, D, dtype)
        q = rand16(rng, (3, HQ, D), dtype)
        kro = np.zeros(3, np.int32)
        qpos = np.array(kv_lens, np.int32) - 1
        a = arenas[i]
        w = dict(qkv=qkv, pos=pos, c=c, q=q, kro=kro, qpos=qpos)
        w["fqkv"], _ = a.put(qkv, dtype)
        w["outs"] = [a.empty((n, h, D), dtype) for h in (HQ, HKV, HKV)]
        w["fq"], _ = a.put(q, dtype)
        w["fp"], _ = a.put(c["pages"], dtype)
        w["aux"] = MergedAux(dict(pos=pos, indptr=c["page_indptr"], values=c["page_values"], li=c["length_info"],
                                  kro=kro, qpos=qpos))
        w["fo"], w["to"] = a.empty((3, HQ, D), dtype)
        w["fl"], w["tl"] = a.empty((3, HQ), "float32")
        work.append(w)
    for _ in range(5):  # interleave: set 0 on stream 0, set 1 on stream 1, ...
        for i in range(2):
            w, ks = work[i], sets[i]
            with tvm_ffi.use_torch_stream(torch.cuda.stream(streams[i])):
                ks["f_split_rotary"](w["fqkv"], w["aux"]["pos"], w["outs"][0][0], w["outs"][1][0], w["outs"][2][0], 1)
                ks["f_attention_decode"](w["fq"], w["fp"], w["aux"]["indptr"], w["aux"]["values"], w["aux"]["li"],
                                         w["aux"]["kro"], w["aux"]["qpos"], w["fo"], w["fl"], 1, 1.0, 5e5, SM)
    torch.cuda.synchronize()
    try:
        for i, rs in enumerate([None, llama3]):
            w = work[i]
            ok.set_rope_scaling(rs)
            wq, wk, _ = ok.split_rotary(w["qkv"], w["pos"], HQ, HKV, 1, 5e5, 1.0, dtype)
            assert_close(f"q[{i}]", to_np(w["outs"][0][1]), wq, atol=3.2e-2)
            assert_close(f"k[{i}]", to_np(w["outs"][1][1]), wk, atol=3.2e-2)
            c = w["c"]
            wo, wl = ok.attention_decode(w["q"], c["pages"], c["page_indptr"], c["page_values"], c["length_info"],
                                         w["kro"], w["qpos"], 1, 1.0, 5e5, SM, dtype)
            assert_close(f"o[{i}]", to_np(w["to"]), wo)
            assert_close(f"lse[{i}]", to_np(w["tl"]), wl)
        # and the two scalings really differ on these positions (the test would be vacuous otherwise)
        ok.set_rope_scaling(None)
        q_default, _, _ = ok.split_rotary(work[1]["qkv"], work[1]["pos"], HQ, HKV, 1, 5e5, 1.0, dtype)
        assert np.abs(q_default - to_np(work[1]["outs"][0][1])).max() > 0.1
    finally:
        ok.set_rope_scaling(None)


def test_errors_surface_as_ffi_exceptions(env):
    """The reference's binders raise before anything is launched; so do the packed functions -- on the GPU box too."""
    import torch

    z = torch.zeros((2, HQ, D), dtype=torch.float16, device="cuda")
    lse = torch.zeros((2, HQ), dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError, match="shape mismatch"):
        env.mod["f_merge_inplace"](z, lse, z, torch.zeros((2, HQ + 1), dtype=torch.float32, device="cuda"))
    with pytest.raises(TypeError, match="expects 13 arguments"):
        env.mod["f_attention_decode"](z, z)
    pages = torch.zeros((4, 2, HKV, 16, D), dtype=torch.float16, device="cuda")
    i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device="cuda")  # noqa: E731
    with pytest.raises(Exception, match="Inline rotary mode is not supported in tree attention"):
        env.mod["f_attention_prefill_with_tree_mask_paged_kv"](z, i32(2), pages, i32(2), i32(1), i32(1), i32(1), i32(2), z,
                                                               lse, 1, 1.0, 1e4, SM, i32(2), i32(1, 2))
