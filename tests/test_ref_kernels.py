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
