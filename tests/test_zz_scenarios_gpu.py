"""GPU replay of the fixtures beyond the reference's own scenario tests (tests/test_cache_gpu_golden.py has those): the
randomised prefill / decode / fork / popn / remove programs, the per-layer sliding-window caches (attn_kinds with
MHA_SLIDING, ..._cpu.py:610-650), attention_with_shared_kv (..._cpu.py:654-703) and the self_attention /
cross_attention / merge_attn_output_inplace entries (kv_state.cc:84-115) -- callback traces bit-exact, outputs within the
north_star tolerance of what the reference's own cache + CPU kernels produced on the same inputs."""
import pytest

from tests.golden_replay import extra_scenario_names
from tests.test_cache_gpu_golden import run_scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl", [0, 2])
@pytest.mark.parametrize("name", extra_scenario_names())
def test_extra_scenario_matches_reference(built_lib, name, impl):
    run_scenario(name, impl)


def test_split_attention_through_the_global_names(built_lib):
    """self_attention + cross_attention + merge_attn_output_inplace and attention_with_shared_kv called by their
    vm.builtin names (kv_state.cc:84-115) equal attention_with_fused_qkv on the same step (rope mode none)."""
    import torch
    from tvm_ffi import Shape

    from tests.test_vm_builtins import _create, _register
    from tvm_b200 import ffi

    f = _register()
    torch.manual_seed(1)
    hq, hkv, d, dt = 32, 8, 128, torch.float16
    cache = _create(f, torch.zeros((), dtype=dt, device="cuda"), rope_mode=0)
    for sid in (0, 1):
        f["vm.builtin.kv_state_add_sequence"](cache, sid)
    sm = d ** -0.5

    def fused(seq_ids, lens, qkv):
        o = torch.full((qkv.shape[0], hq, d), float("nan"), device="cuda", dtype=dt)
        with ffi.torch_stream():
            f["vm.builtin.kv_state_begin_forward"](cache, Shape(seq_ids), Shape(lens))
            f["vm.builtin.attention_kv_cache_attention_with_fused_qkv"](cache, 0, sm, qkv, o)
            f["vm.builtin.kv_state_end_forward"](cache)
        return o

    fused([0, 1], [40, 75], torch.randn((115, hq + 2 * hkv, d), device="cuda", dtype=dt))
    # a second chunk, first through the split entries (nothing is appended), then rolled back and run fused
    lens = [9, 33]
    n = sum(lens)
    qkv = torch.randn((n, hq + 2 * hkv, d), device="cuda", dtype=dt)
    q, k, v = (x.contiguous() for x in (qkv[:, :hq], qkv[:, hq:hq + hkv], qkv[:, hq + hkv:]))
    o_self, o_cross = torch.zeros((n, hq, d), device="cuda", dtype=dt), torch.zeros((n, hq, d), device="cuda", dtype=dt)
    lse_self = torch.full((n, hq), -5e4, device="cuda", dtype=torch.float32)
    lse_cross = torch.full_like(lse_self, -5e4)
    o_shared = torch.full((n, hq, d), float("nan"), device="cuda", dtype=dt)
    with ffi.torch_stream():
        f["vm.builtin.kv_state_begin_forward"](cache, Shape([0, 1]), Shape(lens))
        f["vm.builtin.attention_kv_cache_self_attention"](cache, 0, sm, q, k, v, o_self, lse_self)
        f["vm.builtin.attention_kv_cache_cross_attention"](cache, 0, sm, q, o_cross, lse_cross)
        ret = f["vm.builtin.attention_kv_cache_merge_attn_output_inplace"](cache, o_self, lse_self, o_cross, lse_cross)
        f["vm.builtin.attention_kv_cache_attention_with_shared_kv"](cache, 0, sm, q, k, v, o_shared)
        f["vm.builtin.kv_state_end_forward"](cache)
    torch.cuda.synchronize()
    assert len(ret) == 2
    assert torch.from_dlpack(ret[0]).data_ptr() == o_self.data_ptr()
    assert torch.from_dlpack(ret[1]).data_ptr() == lse_self.data_ptr()
    for sid, ln in zip((0, 1), lens):
        f["vm.builtin.kv_state_popn"](cache, sid, ln)
    want = fused([0, 1], lens, qkv)
    torch.cuda.synchronize()
    assert torch.isfinite(want).all()
    # shared-KV runs the very same kernel sequence as the fused call; self + cross + merge rounds the merged pair once more
    assert torch.equal(o_shared, want)
    assert (o_self.float() - want.float()).abs().max().item() <= 2e-3
