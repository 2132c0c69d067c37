"""KV transfer between caches (disaggregated prefill -> decode) over peer-mapped device pointers: SURVEY section 8(f).3.

Mirrors the reference's own tests, which need NVSHMEM + MPI and are skipped upstream:
  * tests/python/relax/nvshmem/test_runtime_builtin_kv_cache_transfer_kernel.py -- nvshmem.KVTransfer and
    nvshmem.KVTransferPageToPage on literal position maps (the same maps are used here), plus the gather / scatter head
    mappings of kv_transfer.cu:54-66 against the oracle's restatement;
  * tests/python/relax/nvshmem/test_runtime_builtin_kv_cache_transfer.py -- a receiving cache reserves slots
    (disagg_prepare_recv), the sending cache is told where they are (disagg_mark_send) and prefills; afterwards the
    receiver decodes as if it had prefilled itself.  Checked against a third cache that did everything locally: KV dumps
    bit-identical, decode outputs bit-identical.
The single-GPU versions put sender and receiver on the same device (a PE is just a pointer); the two-GPU version (skipped
on a one-GPU box) moves the rows over NVLink peer stores."""
import numpy as np
import pytest

from oracle import kernels as ok
from tests.util import rand16, to_dev, to_np

pytestmark = pytest.mark.gpu
DTYPES = ["float16", "bfloat16"]
POSITIONS = [0, 1, 2, 3, 4, 5, 10, 11, 12, 15, 16, 17, 18, 19, 25, 27]  # the reference test's literal map


@pytest.fixture()
def capi(built_lib):
    from tvm_b200 import capi as c

    c.lib()
    return c


def _i32(a, device="cuda"):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a, np.int32)).to(device)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("page_size", [4, 16])
def test_kv_transfer_kernel(capi, dtype, page_size):
    import torch

    rng = np.random.default_rng(80)
    hkv, d, num_pages, n = 4, 128, 100, len(POSITIONS)
    pages0 = rand16(rng, (num_pages, 2, hkv, page_size, d), dtype)
    k, v = rand16(rng, (n, hkv, d), dtype), rand16(rng, (n, hkv, d), dtype)
    pos = np.array(POSITIONS, np.int32)
    pos[3] = -1                                                   # a token that is not sent
    pe = np.ones(n, np.int32)                                     # the receiver is PE 1; PE 0 (the sender) must stay untouched
    other = to_dev(pages0, dtype)
    remote = to_dev(pages0, dtype)
    capi.kv_transfer([other.data_ptr(), remote.data_ptr()], to_dev(k, dtype), to_dev(v, dtype), _i32(pos), _i32(pe), hkv, page_size)
    torch.cuda.synchronize()
    want = [pages0.copy(), pages0.copy()]
    ok.kv_transfer(want, k, v, pos, pe)
    assert np.array_equal(to_np(other), want[0]) and np.array_equal(to_np(remote), want[1])
    for i, p in enumerate(POSITIONS):                             # ... and literally what the reference test asserts
        if i != 3:
            assert np.array_equal(to_np(remote)[p // page_size, 0, :, p % page_size], k[i])
            assert np.array_equal(to_np(remote)[p // page_size, 1, :, p % page_size], v[i])


@pytest.mark.parametrize("dtype", DTYPES)
def test_kv_transfer_page_to_page_kernel(capi, dtype):
    import torch

    rng = np.random.default_rng(81)
    hkv, d, num_pages, page_size = 4, 128, 100, 4
    local = rand16(rng, (num_pages, 2, hkv, page_size, d), dtype)
    remote0 = rand16(rng, (num_pages, 2, hkv, page_size, d), dtype)
    rpos = np.array(POSITIONS, np.int32)
    lpos = np.array(list(reversed(POSITIONS)), np.int32)
    lpos[5] = -1
    pe = np.zeros(len(POSITIONS), np.int32)
    dremote, dlocal = to_dev(remote0, dtype), to_dev(local, dtype)
    capi.kv_transfer_page_to_page([dremote.data_ptr()], dlocal, _i32(rpos), _i32(lpos), _i32(pe), hkv)
    torch.cuda.synchronize()
    want = [remote0.copy()]
    ok.kv_transfer_page_to_page(want, local, rpos, lpos, pe)
    assert np.array_equal(to_np(dremote), want[0])
    assert np.array_equal(to_np(dlocal), local)


@pytest.mark.parametrize("local_h,remote_h,senders,receivers", [(2, 4, 2, 1), (4, 2, 1, 2), (1, 8, 8, 1), (8, 2, 2, 8)])
def test_kv_transfer_head_gather_and_scatter(capi, local_h, remote_h, senders, receivers):
    """sender TP ranks with local_h kv heads each -> receiver TP group (starting at PE 1) with remote_h heads per rank"""
    import torch

    rng = np.random.default_rng(82)
    dtype, d, page_size, num_pages, n = "float16", 64, 16, 12, 40
    npe = 1 + receivers
    pools0 = [rand16(rng, (num_pages, 2, remote_h, page_size, d), dtype) for _ in range(npe)]
    pools = [to_dev(p, dtype) for p in pools0]
    want = [p.copy() for p in pools0]
    pos = rng.permutation(num_pages * page_size)[:n].astype(np.int32)
    pos[::7] = -1
    off = np.ones(n, np.int32)
    for rank in range(senders):
        k, v = rand16(rng, (n, local_h, d), dtype), rand16(rng, (n, local_h, d), dtype)
        capi.kv_transfer([p.data_ptr() for p in pools], to_dev(k, dtype), to_dev(v, dtype), _i32(pos), _i32(off), remote_h,
                         page_size, local_tp_rank=rank)
        ok.kv_transfer(want, k, v, pos, off, local_tp_rank=rank)
    torch.cuda.synchronize()
    for got, w in zip(pools, want):
        assert np.array_equal(to_np(got), w)


def test_kv_transfer_rejects_bad_geometry(capi):
    import torch

    from tvm_b200.capi import TvmB200Error

    k = torch.zeros((4, 3, 128), dtype=torch.float16, device="cuda")
    pos = torch.zeros(4, dtype=torch.int32, device="cuda")
    with pytest.raises(TvmB200Error, match="do not divide"):
        capi.kv_transfer([k.data_ptr()], k, k, pos, pos, 2, 16)
    with pytest.raises(TvmB200Error, match="processing elements"):
        capi.kv_transfer([], k, k, pos, pos, 3, 16)


# ---------------------------------------------------------------------------------------------------------------------
PREFILL_OPS = [[(0, 6)], [(1, 8)], [(2, 11)], [(3, 16)], [(4, 19), (5, 20)], [(6, 21), (7, 24)], [(2, 5), (4, 7), (8, 24)],
               [(6, 13)], [(8, 19)], [(0, 1)], [(1, 3), (3, 8), (5, 12), (7, 11)]]      # the reference test's operation lists
DECODE_OPS = [[(s, 1) for s in range(9)], [(s, 1) for s in range(9)], [(s, 1) for s in (0, 2, 4, 6, 8)],
              [(s, 1) for s in (4, 5, 6, 7, 8)]]


def _make_cache(device, rope_mode, dtype="float16", layers=2):
    from tvm_b200.kv_cache import PagedKVCache

    return PagedKVCache(reserved_num_seqs=16, total_token_capacity=1024, prefill_chunk_size=128, num_layers=layers,
                        num_qo_heads=8, num_kv_heads=2, head_dim=128, rope_mode=rope_mode, rotary_theta=1e4, dtype=dtype,
                        device=device)


def _forward(cache, batch, seed, device, layers=2, dtype="float16"):
    """one begin_forward / attention over all layers / end_forward; returns the outputs [layers][n, Hq, D] (numpy)"""
    import torch

    n = sum(l for _, l in batch)
    rng = np.random.default_rng(seed)
    qkv = rand16(rng, (layers, n, 8 + 2 * 2, 128), dtype)
    cache.begin_forward([s for s, _ in batch], [l for _, l in batch])
    outs = []
    with torch.cuda.device(device):
        for layer in range(layers):
            o = torch.full((n, 8, 128), float("nan"), dtype=torch.float16, device=f"cuda:{device}")
            cache.attention_with_fused_qkv(layer, 128 ** -0.5, to_dev(qkv[layer], dtype, f"cuda:{device}"), o)
            outs.append(o)
        cache.end_forward()
        torch.cuda.synchronize(device)
    return [to_np(o) for o in outs]


def _dump(cache, seq, length, device, layers=2):
    import torch

    with torch.cuda.device(device):
        k = torch.zeros((layers, length, 2, 128), dtype=torch.float16, device=f"cuda:{device}")
        v = torch.zeros_like(k)
        cache.debug_get_kv(seq, 0, length, k, v)
        torch.cuda.synchronize(device)
    return to_np(k), to_np(v)


def _run_disaggregated(send_dev, recv_dev, rope_mode):
    import torch

    from tvm_b200 import capi

    prefill_len = {s: 0 for s in range(9)}
    for batch in PREFILL_OPS:
        for s, l in batch:
            prefill_len[s] += l
    with torch.cuda.device(recv_dev):
        recv = _make_cache(recv_dev, rope_mode)
    with torch.cuda.device(send_dev):
        send, local = _make_cache(send_dev, rope_mode), _make_cache(send_dev, rope_mode)
    if send_dev != recv_dev:
        capi.enable_peer_access(send_dev, recv_dev)
    # receiver (the "decode instance"): reserve the slots
    maps = {}
    with torch.cuda.device(recv_dev):
        for s, l in prefill_len.items():
            recv.add_sequence(s)
            maps[s] = recv.disagg_prepare_recv(s, l)
            assert maps[s][0] * 2 + 1 == len(maps[s]) and sum(maps[s][2::2]) == l
            recv.end_forward()
    # sender (the "prefill instance"): PE 0 is itself, PE 1 the receiver
    with torch.cuda.device(send_dev):
        send.enable_kv_transfer(local_tp_rank=0, num_pe=2)
        for layer in range(2):
            send.set_remote_pages(1, layer, recv.pages_ptr(layer)[0])
        for s in prefill_len:
            send.add_sequence(s)
            local.add_sequence(s)
            send.disagg_mark_send(s, 0, maps[s], 1)
    seed = 1000
    for batch in PREFILL_OPS:
        with torch.cuda.device(send_dev):
            got = _forward(send, batch, seed, send_dev)
            want = _forward(local, batch, seed, send_dev)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), "sending must not change the sender's own attention"
        seed += 1
    torch.cuda.synchronize(send_dev)
    # the receiver now holds what a local prefill would have produced ...
    for s, l in prefill_len.items():
        with torch.cuda.device(recv_dev):
            gk, gv = _dump(recv, s, l, recv_dev)
        with torch.cuda.device(send_dev):
            wk, wv = _dump(local, s, l, send_dev)
        assert np.array_equal(gv, wv), f"sequence {s}: transferred V differs"
        assert np.array_equal(gk, wk), f"sequence {s}: transferred K differs"
    # ... and decodes exactly like it
    for batch in DECODE_OPS:
        with torch.cuda.device(recv_dev):
            got = _forward(recv, batch, seed, recv_dev)
        with torch.cuda.device(send_dev):
            want = _forward(local, batch, seed, send_dev)
        for a, b in zip(got, want):
            assert np.isfinite(a).all() and np.array_equal(a, b)
        seed += 1


@pytest.mark.parametrize("rope_mode", [1, 0])
def test_disaggregated_prefill_then_decode_same_gpu(built_lib, rope_mode):
    _run_disaggregated(0, 0, rope_mode)


def test_disaggregated_prefill_then_decode_two_gpus(built_lib):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one NVLink domain (gpurun --gpus 2)")
    _run_disaggregated(0, 1, 1)


def test_mark_send_of_a_partly_prefilled_sequence_goes_page_to_page(built_lib):
    """mark_send with begin inside the cached part: the cached rows travel page to page with the next forward
    (paged_kv_cache.cc:1269-1300, :1197-1210), the new rows as usual"""
    import torch

    send, local, recv = _make_cache(0, 1), _make_cache(0, 1), _make_cache(0, 1)
    for c in (send, local, recv):
        c.add_sequence(7)
    first, second, begin = 37, 20, 5
    _forward(send, [(7, first)], 1, 0)
    _forward(local, [(7, first)], 1, 0)
    # the receiver already holds the first `begin` tokens (say, from a prefix cache) and reserves the rest
    _forward(recv, [(7, begin)], 1, 0)  # (different values than the sender's: only slots [begin, ...) are compared below)
    m = recv.disagg_prepare_recv(7, first + second - begin)
    recv.end_forward()
    send.enable_kv_transfer(0, 1)
    for layer in range(2):
        send.set_remote_pages(0, layer, recv.pages_ptr(layer)[0])
    send.disagg_mark_send(7, begin, m, 0)
    _forward(send, [(7, second)], 2, 0)
    _forward(local, [(7, second)], 2, 0)
    torch.cuda.synchronize()
    gk, gv = _dump(recv, 7, first + second, 0)
    wk, wv = _dump(local, 7, first + second, 0)
    assert np.array_equal(gk[:, begin:], wk[:, begin:]) and np.array_equal(gv[:, begin:], wv[:, begin:])


@pytest.mark.parametrize("page_to_page", [0, 1])
def test_bound_packed_functions_have_the_nvshmem_signatures(built_lib, page_to_page):
    """bind_kv_transfer(pool base pointers per PE, own PE, tp rank, page_to_page) -> a packed function with the reference's
    nvshmem.KVTransfer / KVTransferPageToPage signature (kv_transfer.cu:139-257, :259-325): `remote_pages` is the caller's
    own pool view of the LAYER (as in the reference test: a view at a byte offset into a [layers, ...] pool), whose offset
    inside the pool selects the layer on every PE."""
    import torch
    import tvm_ffi
    from tvm_ffi import Shape

    from tvm_b200.build import LIB

    rng = np.random.default_rng(83)
    mod = tvm_ffi.load_module(str(LIB))
    layers, layer_id, num_pages, hkv, page, d, n = 4, 1, 100, 4, 4, 128, len(POSITIONS)
    pool0 = rand16(rng, (layers, num_pages, 2, hkv, page, d), "float16")
    local = to_dev(pool0, "float16")          # PE 0: the sender's own pool
    remote = to_dev(pool0, "float16")         # PE 1: the receiver's
    f = mod["bind_kv_transfer"](Shape([local.data_ptr(), remote.data_ptr()]), 0, 0, page_to_page)
    pos = np.array(POSITIONS, np.int32)
    pe = np.ones(n, np.int32)
    want = pool0.copy()
    layer_view = tvm_ffi.from_dlpack(local[layer_id])
    if not page_to_page:
        k, v = rand16(rng, (n, hkv, d), "float16"), rand16(rng, (n, hkv, d), "float16")
        f(layer_view, tvm_ffi.from_dlpack(to_dev(k, "float16")), tvm_ffi.from_dlpack(to_dev(v, "float16")),
          tvm_ffi.from_dlpack(_i32(pos)), tvm_ffi.from_dlpack(_i32(pe)), None)
        w = [None, want[layer_id]]
        ok.kv_transfer(w, k, v, pos, pe)
    else:
        lpos = np.array(list(reversed(POSITIONS)), np.int32)
        f(layer_view, layer_view, tvm_ffi.from_dlpack(_i32(pos)), tvm_ffi.from_dlpack(_i32(lpos)), tvm_ffi.from_dlpack(_i32(pe)), None)
        w = [None, want[layer_id]]
        ok.kv_transfer_page_to_page(w, pool0[layer_id], pos, lpos, pe)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(remote), want), "the receiver's pool differs (wrong layer offset or rows)"
    assert np.array_equal(to_np(local), pool0), "the sender's own pool was touched"
