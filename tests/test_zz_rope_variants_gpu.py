"""f_split_rotary under the remaining rope_scaling types of the reference (gptj / llama4 / yarn; switch_rope_freq_func,
position_embedding.py:257-299): the sm_100a kernel against outputs of the reference's own fused_rope
(tests/golden/rope_variants.npz, oracle/ref_harness/gen_golden_rope_variants.py) and against the oracle on larger inputs;
in-kernel rotations are rejected while one of these types is set."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import kernels as ok
from tests.util import assert_close, rand16, to_dev, to_np

pytestmark = pytest.mark.gpu

CASES = ["gptj", "gptj_rd64", "llama4", "llama4_equal_factors", "yarn"]


@pytest.fixture()
def capi(built_lib):
    from tvm_b200 import capi as c

    c.lib()
    yield c
    c.set_rope_scaling(None)
    ok.set_rope_scaling(None)


def _golden():
    g = np.load(Path(__file__).parent / "golden" / "rope_variants.npz")
    return g, json.loads(bytes(g["meta"]).decode())


def _i32(x):
    return to_dev(np.asarray(x, np.int32))


@pytest.mark.parametrize("name", CASES)
def test_split_rotary_variant_against_the_reference_golden(capi, name):
    import torch

    g, meta = _golden()
    m = meta[name]
    capi.set_rope_scaling(m["rope_scaling"])
    assert capi.get_rope_scaling_kind() == capi.ROPE_SCALING_KINDS[m["rope_scaling"]["rope_type"]]
    n, hq, hkv, d = g[f"{name}_qkv"].shape[0], 8, 2, 128
    q = torch.empty((n, hq, d), dtype=torch.float16, device="cuda")
    k = torch.empty((n, hkv, d), dtype=torch.float16, device="cuda")
    v = torch.empty((n, hkv, d), dtype=torch.float16, device="cuda")
    capi.split_rotary(torch.from_numpy(g[f"{name}_qkv"]).cuda(), _i32(g[f"{name}_pos"]), q, k, v, 1, m["scale"], m["theta"],
                      m["rotary_dim"] or 0)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(v), g[f"{name}_v"].astype(np.float32))
    for i, pos in enumerate(g[f"{name}_pos"]):  # the angle is a float32: its ulp grows with the position
        atol = 4e-3 + 3e-7 * float(pos)
        assert_close(f"{name} q[{i}]", to_np(q)[i], g[f"{name}_q"][i].astype(np.float32), atol=atol)
        assert_close(f"{name} k[{i}]", to_np(k)[i], g[f"{name}_k"][i].astype(np.float32), atol=atol)


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("name", CASES)
def test_split_rotary_variant_vs_oracle_and_append(capi, name, dtype):
    """Llama-3-8B head shape, 37 tokens at positions < 4096; the `_append` entry equals split_rotary + transpose_append."""
    import torch

    _, meta = _golden()
    m = meta[name]
    capi.set_rope_scaling(m["rope_scaling"])
    ok.set_rope_scaling(m["rope_scaling"])
    rng = np.random.default_rng(5)
    n, hq, hkv, d, P = 37, 32, 8, 128, 6
    tdt = torch.float16 if dtype == "float16" else torch.bfloat16
    qkv = rand16(rng, (n, hq + 2 * hkv, d), dtype)
    pos = rng.integers(0, 4096, n).astype(np.int32)
    want_q, want_k, want_v = ok.split_rotary(qkv, pos, hq, hkv, 1, m["theta"], m["scale"], dtype, m["rotary_dim"])
    q = torch.empty((n, hq, d), dtype=tdt, device="cuda")
    k = torch.empty((n, hkv, d), dtype=tdt, device="cuda")
    v = torch.empty((n, hkv, d), dtype=tdt, device="cuda")
    dq = to_dev(qkv, dtype)
    capi.split_rotary(dq, _i32(pos), q, k, v, 1, m["scale"], m["theta"], m["rotary_dim"] or 0)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(v), want_v)
    assert_close(f"{name} q", to_np(q), want_q, atol=8e-3 if dtype == "float16" else 4e-2)
    assert_close(f"{name} k", to_np(k), want_k, atol=8e-3 if dtype == "float16" else 4e-2)
    # apply_rope = 0 is a plain split under every scaling
    q0 = torch.empty_like(q)
    capi.split_rotary(dq, _i32(pos), q0, torch.empty_like(k), torch.empty_like(v), 0, m["scale"], m["theta"], 0)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(q0), qkv[:, :hq])
    # fused with the append
    slots = rng.permutation(P * 16)[:n].astype(np.int32)
    slots[3] = -1
    pages_a = torch.zeros((P, 2, hkv, 16, d), dtype=tdt, device="cuda")
    pages_b = torch.zeros_like(pages_a)
    q2, k2, v2 = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    capi.split_rotary_append(dq, _i32(pos), _i32(slots), q2, k2, v2, pages_a, 1, m["scale"], m["theta"], m["rotary_dim"] or 0)
    capi.transpose_append(pages_b, k, v, _i32(slots))
    torch.cuda.synchronize()
    assert torch.equal(q2, q) and torch.equal(k2, k) and torch.equal(v2, v) and torch.equal(pages_a, pages_b)


def test_in_kernel_rotations_are_rejected_under_a_variant(capi):
    import torch

    capi.set_rope_scaling({"rope_type": "gptj"})
    hq, hkv, d = 32, 8, 128
    q = torch.zeros((1, hq, d), dtype=torch.float16, device="cuda")
    pages = torch.zeros((2, 2, hkv, 16, d), dtype=torch.float16, device="cuda")
    o, lse = torch.zeros_like(q), torch.zeros((1, hq), dtype=torch.float32, device="cuda")
    one, zero = _i32([0, 1]), _i32([0])
    with pytest.raises(capi.TvmB200Error, match="only implemented by split_rotary"):
        capi.attention_decode(q, pages, one, zero, _i32([5]), zero, _i32([4]), o, lse, 1, 1.0, 1e4, d ** -0.5)
    capi.attention_decode(q, pages, one, zero, _i32([5]), zero, _i32([4]), o, lse, 0, 1.0, 1e4, d ** -0.5)  # no rotation: fine
    with pytest.raises(capi.TvmB200Error, match="only implemented by split_rotary"):
        capi.attention_prefill_ragged(q, one, pages[0, 0, :, :1].reshape(1, hkv, d).contiguous(),
                                      pages[0, 1, :, :1].reshape(1, hkv, d).contiguous(), one, zero, zero, o, lse, 1, 1, 1.0,
                                      1e4, d ** -0.5)
    torch.cuda.synchronize()


@pytest.mark.parametrize("name", ["gptj", "yarn"])
def test_host_cache_decode_step_under_a_variant(capi, name):
    """rope mode "normal": the cache rotates in split_rotary; a decode step must take the split_rotary + append + decode
    route (the fused launch rotates in-kernel with the default / llama3 frequencies only) and give what the three C-ABI
    calls give on the same page table."""
    import torch

    from tvm_b200.kv_cache import PagedKVCache

    _, meta = _golden()
    m = meta[name]
    capi.set_rope_scaling(m["rope_scaling"])
    hq, hkv, d, dt = 32, 8, 128, torch.float16
    cache = PagedKVCache(reserved_num_seqs=4, total_token_capacity=1024, prefill_chunk_size=256, num_layers=1,
                         num_qo_heads=hq, num_kv_heads=hkv, head_dim=d, rope_mode=1, rotary_theta=m["theta"], dtype="float16")
    torch.manual_seed(3)
    lens = [40, 75]
    for sid in (0, 1):
        cache.add_sequence(sid)
    qkv = torch.randn((sum(lens), hq + 2 * hkv, d), device="cuda", dtype=dt)
    o = torch.empty((sum(lens), hq, d), device="cuda", dtype=dt)
    cache.begin_forward([0, 1], lens)
    cache.attention_with_fused_qkv(0, d ** -0.5, qkv, o)
    cache.end_forward()
    qkv1 = torch.randn((2, hq + 2 * hkv, d), device="cuda", dtype=dt)
    o1 = torch.full((2, hq, d), float("nan"), device="cuda", dtype=dt)
    cache.begin_forward([0, 1], [1, 1])
    cache.attention_with_fused_qkv(0, d ** -0.5, qkv1, o1)
    cache.end_forward()
    torch.cuda.synchronize()
    assert torch.isfinite(o1).all()
    # the k / v the step appended must be what split_rotary gives under this scaling (the fused launch would have
    # rotated k with the default frequencies)
    q = torch.empty((2, hq, d), device="cuda", dtype=dt)
    k = torch.empty((2, hkv, d), device="cuda", dtype=dt)
    v = torch.empty_like(k)
    capi.split_rotary(qkv1, _i32(lens), q, k, v, 1, 1.0, m["theta"], 0)
    torch.cuda.synchronize()
    dbg_k = torch.empty((1, 1, hkv, d), device="cuda", dtype=dt)
    dbg_v = torch.empty_like(dbg_k)
    for sid, ln in zip((0, 1), lens):
        cache.debug_get_kv(sid, ln, ln + 1, dbg_k, dbg_v)   # the token appended by the decode step
        torch.cuda.synchronize()
        assert torch.equal(dbg_k[0, 0], k[sid]) and torch.equal(dbg_v[0, 0], v[sid])
