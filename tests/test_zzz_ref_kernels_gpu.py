"""The sm_100a kernels against the reference's OWN compiled CPU kernels on the GPU box: oracle/_ref holds
llama_rope_with_position_map, _kv_cache_transpose_append, _attention_decode_cpu, _attention_prefill_ragged_cpu and
_merge_state_inplace_cpu built from /root/reference at the Llama-3-8B head shape (32 q / 8 kv heads, D 128, fp16) by the
reference's `c` target (oracle/ref_harness/emit_ref_kernels.py); the .so travels with the repository snapshot, nothing of
/root/reference is read here.  Tolerance = north_star: max-abs 2e-3 / rtol 1e-2 on O and LSE, appended V bit-exact.
Skipped where oracle/_ref was not built."""
import numpy as np
import pytest

from oracle import cpu_ref
from tests.golden_replay import late_scenario_names
from tests.test_cache_gpu_golden import run_scenario
from tests.test_ref_kernels import _ragged_inputs, ref_merge, ref_ragged_prefill
from tests.util import assert_close, to_dev, to_np

pytestmark = pytest.mark.gpu
HQ, HKV, D, THETA = 32, 8, 128, 5e5


@pytest.fixture()
def ref_mod():
    mod = cpu_ref._ref_module("float16", HQ, HKV, D)
    if mod is None:
        pytest.skip("oracle/_ref not built (needs the reference build, see oracle/ref_harness/)")
    return mod


def _i32(x):
    return to_dev(np.asarray(x, np.int32))


def test_decode_step_vs_the_reference_kernels(built_lib, ref_mod):
    """split_rotary + transpose_append + decode of one layer: the reference's three PrimFuncs on the CPU, ours on the GPU."""
    import torch

    from tvm_b200 import capi

    B, L = 4, 300
    inp = cpu_ref._decode_inputs(B, L, HQ, HKV, D, "float16", seed=11)
    t = {k: torch.from_numpy(v.astype(np.float16) if v.dtype == np.float32 else v.copy()) for k, v in inp.items()}
    for nm, h in (("q", HQ), ("k", HKV), ("v", HKV), ("o", HQ)):
        t[nm] = torch.zeros((B, h, D), dtype=torch.float16)
    t["lse"] = torch.zeros((B, HQ), dtype=torch.float32)
    cpu_ref._one_step_ref(ref_mod, t, HQ, HKV, D, THETA)

    pages = to_dev(inp["pages"], "float16")
    qkv = to_dev(inp["qkv"], "float16")
    q = torch.empty((B, HQ, D), dtype=torch.float16, device="cuda")
    k = torch.empty((B, HKV, D), dtype=torch.float16, device="cuda")
    v = torch.empty_like(k)
    o = torch.empty_like(q)
    lse = torch.empty((B, HQ), dtype=torch.float32, device="cuda")
    capi.split_rotary_append(qkv, _i32(inp["qpos"]), _i32(inp["apos"]), q, k, v, pages, 1, 1.0, THETA)
    capi.attention_decode(q, pages, _i32(inp["page_indptr"]), _i32(inp["page_values"]), _i32(inp["length_info"]),
                          _i32(inp["kofs"]), _i32(inp["qpos"]), o, lse, 0, 1.0, THETA, D ** -0.5)
    torch.cuda.synchronize()
    ref_pages = t["pages"].float().numpy()
    assert np.array_equal(to_np(pages)[:, 1], ref_pages[:, 1]), "appended V differs from the reference"
    assert_close("pages K (rotated)", to_np(pages)[:, 0], ref_pages[:, 0])
    assert_close("q (rotated)", to_np(q), t["q"].float().numpy())
    assert_close("decode O", to_np(o), t["o"].float().numpy())
    assert_close("decode LSE", to_np(lse), t["lse"].numpy())


@pytest.mark.parametrize("rotary_mode", [0, 1])
def test_ragged_prefill_and_merge_vs_the_reference_kernels(built_lib, ref_mod, rotary_mode):
    import torch

    from tvm_b200 import capi

    inp = _ragged_inputs()
    want_o, want_lse = ref_ragged_prefill(ref_mod, inp, rotary_mode, THETA)
    n = inp["q"].shape[0]
    o = torch.empty((n, HQ, D), dtype=torch.float16, device="cuda")
    lse = torch.empty((n, HQ), dtype=torch.float32, device="cuda")
    capi.attention_prefill_ragged(to_dev(inp["q"], "float16"), _i32(inp["ip"]), to_dev(inp["k"], "float16"),
                                  to_dev(inp["v"], "float16"), _i32(inp["ip"]), _i32(inp["qpos"]), _i32(inp["kofs"]), o, lse, 1,
                                  rotary_mode, 1.0, THETA, D ** -0.5)
    torch.cuda.synchronize()
    assert_close("ragged prefill O", to_np(o), want_o)
    assert_close("ragged prefill LSE", to_np(lse), want_lse)
    # merge the reference's (O, LSE) with a second state: reference kernel on the CPU, ours on the GPU, same inputs
    rng = np.random.default_rng(9)
    o2 = rng.standard_normal(want_o.shape).astype(np.float16).astype(np.float32)
    lse2 = (want_lse + rng.standard_normal(want_lse.shape).astype(np.float32) * 3).astype(np.float32)
    mv, ms = ref_merge(ref_mod, want_o, want_lse, o2, lse2)
    dv, ds = to_dev(want_o, "float16"), to_dev(want_lse.copy())
    capi.merge_state_inplace(dv, ds, to_dev(o2, "float16"), to_dev(lse2))
    torch.cuda.synchronize()
    assert_close("merge V", to_np(dv), mv)
    assert_close("merge S", to_np(ds), ms)


@pytest.mark.parametrize("impl", [0, 2])
@pytest.mark.parametrize("name", late_scenario_names())
def test_late_scenario_matches_reference(built_lib, name, impl):
    """Fixtures captured from the reference late in the round (tests/golden_replay.py::LATE_SCENARIOS)."""
    run_scenario(name, impl)


# ---- paged / sliding-window / tree-mask kernels: same seeded cases as tests/test_ref_kernels.py, CUDA in the oracle's place
def _paged_args(x):
    import torch

    c, n = x["c"], x["q"].shape[0]
    o = torch.full((n, HQ, D), float("nan"), dtype=torch.float16, device="cuda")
    lse = torch.full((n, HQ), float("nan"), dtype=torch.float32, device="cuda")
    return (to_dev(x["q"], "float16"), _i32(x["qi"]), to_dev(c["pages"], "float16"), _i32(c["page_indptr"]),
            _i32(c["page_values"]), _i32(c["length_info"]), _i32(x["kofs"]), _i32(x["qpos"]), o, lse)


def _need_full_ref(ref_mod):
    try:
        ref_mod["batch_prefill_paged_kv_cpu"]
    except Exception:
        pytest.skip("oracle/_ref holds the decode-step kernels only (re-run oracle/ref_harness/emit_ref_kernels.py)")
    return ref_mod


@pytest.mark.parametrize("causal,rotary_mode", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_paged_prefill_vs_the_reference_kernels(built_lib, ref_mod, causal, rotary_mode):
    import torch

    from tests.test_ref_kernels import case_paged_prefill
    from tvm_b200 import capi

    x, want_o, want_lse = case_paged_prefill(_need_full_ref(ref_mod), causal, rotary_mode)
    args = _paged_args(x)
    capi.attention_prefill_paged(*args, causal, rotary_mode, 1.0, THETA, D ** -0.5)
    torch.cuda.synchronize()
    assert_close("paged prefill O", to_np(args[8]), want_o)
    assert_close("paged prefill LSE", to_np(args[9]), want_lse)


@pytest.mark.parametrize("rotary_mode", [0, 1])
def test_sliding_window_flavours_vs_the_reference_kernels(built_lib, ref_mod, rotary_mode):
    import torch

    from tests.test_ref_kernels import case_sliding_decode, case_sliding_prefill
    from tvm_b200 import capi

    mod = _need_full_ref(ref_mod)
    x, want_o, want_lse = case_sliding_decode(mod, rotary_mode)
    q, _qi, pages, pip, piv, li, kofs, qpos, o, lse = _paged_args(x)
    capi.attention_decode(q, pages, pip, piv, li, kofs, qpos, o, lse, rotary_mode, 1.0, THETA, D ** -0.5)
    torch.cuda.synchronize()
    assert_close("sliding decode O", to_np(o), want_o)
    assert_close("sliding decode LSE", to_np(lse), want_lse)
    x, want_o, want_lse = case_sliding_prefill(mod, rotary_mode)
    args = _paged_args(x)
    capi.attention_prefill_paged(*args, 0, rotary_mode, 1.0, THETA, D ** -0.5, layer_sliding_window_size=1024)
    torch.cuda.synchronize()
    assert_close("sliding prefill O", to_np(args[8]), want_o)
    assert_close("sliding prefill LSE", to_np(args[9]), want_lse)


def test_tree_attention_vs_the_reference_kernels(built_lib, ref_mod):
    import torch

    from tests.test_ref_kernels import case_tree_paged, case_tree_ragged
    from tvm_b200 import capi

    mod = _need_full_ref(ref_mod)
    x, want_o, want_lse = case_tree_ragged(mod)
    n = x["q"].shape[0]
    o = torch.full((n, HQ, D), float("nan"), dtype=torch.float16, device="cuda")
    lse = torch.full((n, HQ), float("nan"), dtype=torch.float32, device="cuda")
    capi.attention_prefill_tree_ragged(to_dev(x["q"], "float16"), _i32(x["ip"]), to_dev(x["k"], "float16"),
                                       to_dev(x["v"], "float16"), _i32(x["ip"]), _i32(x["qpos"]), _i32(x["ip"]), _i32(x["masks"]),
                                       o, lse, 0, 1.0, THETA, D ** -0.5)
    torch.cuda.synchronize()
    assert_close("tree ragged O", to_np(o), want_o)
    assert_close("tree ragged LSE", to_np(lse), want_lse)
    x, want_o, want_lse = case_tree_paged(mod)
    args = _paged_args(x)
    capi.attention_prefill_tree_paged(*args, 0, 1.0, THETA, D ** -0.5, _i32(x["ip"]), _i32(x["masks"]))
    torch.cuda.synchronize()
    assert_close("tree paged O", to_np(args[8]), want_o)
    assert_close("tree paged LSE", to_np(args[9]), want_lse)
