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
