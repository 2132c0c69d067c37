"""tvm_b200's host cache registered under the Relax VM's global function names
(`vm.builtin.paged_attention_kv_cache_create`, `vm.builtin.kv_state_*`, `vm.builtin.attention_kv_cache_*`:
src/runtime/vm/kv_state.cc:33-116, paged_kv_cache.cc:2535-2639), so a compiled model that looks them up by name runs on the
sm_100a cache and kernels.  The CPU part checks the registration and the loud failure without a CUDA tensor; the GPU part
drives a prefill + decode + fork + popn scenario through the global names exactly as the reference's own test calls them
(tests/python/relax/test_runtime_builtin_paged_attention_kv_cache_cpu.py:211-300) and compares with the Python face."""
import numpy as np
import pytest

NAMES = [
    "vm.builtin.paged_attention_kv_cache_create", "vm.builtin.kv_state_clear", "vm.builtin.kv_state_add_sequence",
    "vm.builtin.kv_state_remove_sequence", "vm.builtin.kv_state_fork_sequence", "vm.builtin.kv_state_popn",
    "vm.builtin.kv_state_begin_forward", "vm.builtin.kv_state_end_forward",
    "vm.builtin.attention_kv_cache_enable_sliding_window_for_seq",
    "vm.builtin.attention_kv_cache_commit_accepted_token_tree_nodes", "vm.builtin.attention_kv_cache_empty",
    "vm.builtin.attention_kv_cache_get_num_available_pages", "vm.builtin.attention_kv_cache_get_total_sequence_length",
    "vm.builtin.attention_kv_cache_get_query_positions", "vm.builtin.attention_kv_cache_debug_get_kv",
    "vm.builtin.attention_kv_cache_attention_with_fused_qkv", "vm.builtin.attention_kv_cache_self_attention",
    "vm.builtin.attention_kv_cache_cross_attention", "vm.builtin.attention_kv_cache_attention_with_shared_kv",
    "vm.builtin.attention_kv_cache_merge_attn_output_inplace",
]


def _register():
    import tvm_ffi

    from tvm_b200 import ffi

    n = int(ffi.module()["register_vm_builtins"](1))
    assert n >= len(NAMES)
    return {name: tvm_ffi.get_global_func(name) for name in NAMES}


def _create(f, init, *, seqs=8, tokens=4096, chunk=512, layers=1, hq=32, hkv=8, d=128, rope_mode=1, theta=1e4):
    from tvm_ffi import Shape

    none_fn = None
    return f["vm.builtin.paged_attention_kv_cache_create"](
        Shape([seqs, tokens, chunk, 16, 0]), Shape([0, layers]), hq, hkv, d, d, Shape([0] * layers), False, rope_mode, 1.0,
        theta, None, init, *([none_fn] * 15))


def test_registers_every_vm_builtin_name(built_lib):
    f = _register()
    for name in NAMES:
        assert f[name] is not None, name


def test_create_without_cuda_tensor_fails_loudly(built_lib):
    import torch

    f = _register()
    with pytest.raises(Exception, match="CUDA device"):
        _create(f, torch.zeros((), dtype=torch.float16))
    with pytest.raises(Exception, match="cache returned by"):
        f["vm.builtin.kv_state_add_sequence"](3, 0)


def test_split_attention_entries_reject_cpu_tensors(built_lib):
    import torch

    f = _register()
    q = torch.zeros((2, 4, 128), dtype=torch.float16)
    lse = torch.zeros((2, 4), dtype=torch.float32)
    # no CPU fallback behind any of the entries
    with pytest.raises(Exception, match="must be a CUDA tensor"):
        f["vm.builtin.attention_kv_cache_self_attention"](0, 0, 1.0, q, q, q, q, lse)
    with pytest.raises(Exception, match="must be a CUDA tensor"):
        f["vm.builtin.attention_kv_cache_cross_attention"](0, 0, 1.0, q, q, lse)
    with pytest.raises(Exception, match="must be a CUDA tensor"):
        f["vm.builtin.attention_kv_cache_attention_with_shared_kv"](0, 0, 1.0, q, q, q, q)
    with pytest.raises(Exception, match="must be a CUDA tensor"):
        f["vm.builtin.attention_kv_cache_merge_attn_output_inplace"](0, q, lse, q, lse)


@pytest.mark.gpu
def test_scenario_through_the_global_names_matches_the_python_face(built_lib):
    import torch
    from tvm_ffi import Shape

    from tvm_b200 import ffi
    from tvm_b200.kv_cache import PagedKVCache

    f = _register()
    torch.manual_seed(0)
    hq, hkv, d, dt = 32, 8, 128, torch.float16
    init = torch.zeros((), dtype=dt, device="cuda")
    cache = _create(f, init)
    ref = PagedKVCache(reserved_num_seqs=8, total_token_capacity=4096, prefill_chunk_size=512, num_layers=1,
                       num_qo_heads=hq, num_kv_heads=hkv, head_dim=d, rope_mode=1, rotary_theta=1e4, dtype="float16")
    assert bool(f["vm.builtin.attention_kv_cache_empty"](cache))

    def step(seq_ids, lens):
        n = sum(lens)
        qkv = torch.randn((n, hq + 2 * hkv, d), device="cuda", dtype=dt)
        o1 = torch.full((n, hq, d), float("nan"), device="cuda", dtype=dt)
        o2 = torch.full_like(o1, float("nan"))
        with ffi.torch_stream():
            f["vm.builtin.kv_state_begin_forward"](cache, Shape(seq_ids), Shape(lens))
            f["vm.builtin.attention_kv_cache_attention_with_fused_qkv"](cache, 0, d ** -0.5, qkv, o1)
            pos = f["vm.builtin.attention_kv_cache_get_query_positions"](cache)
            f["vm.builtin.kv_state_end_forward"](cache)
        ref.begin_forward(seq_ids, lens)
        ref.attention_with_fused_qkv(0, d ** -0.5, qkv, o2)
        ref.end_forward()
        torch.cuda.synchronize()
        assert torch.equal(o1, o2) and torch.isfinite(o1).all()
        pos_t = torch.from_dlpack(pos)
        assert pos_t.shape == (n,) and pos_t.dtype == torch.int32
        return pos_t.cpu().numpy()

    for sid in (0, 1, 2):
        f["vm.builtin.kv_state_add_sequence"](cache, sid)
        ref.add_sequence(sid)
    pos = step([0, 1, 2], [37, 300, 5])                      # prefill
    assert np.array_equal(pos, np.concatenate([np.arange(37), np.arange(300), np.arange(5)]))
    pos = step([0, 1, 2], [1, 1, 1])                          # decode
    assert np.array_equal(pos, [37, 300, 5])
    f["vm.builtin.kv_state_fork_sequence"](cache, 1, 3, 100)  # fork at 100, then extend the child
    ref.fork_sequence(1, 3, 100)
    step([3], [20])
    f["vm.builtin.kv_state_popn"](cache, 0, 8)
    ref.popn(0, 8)
    pos = step([0, 3], [2, 1])
    assert np.array_equal(pos, [30, 31, 120])
    assert int(f["vm.builtin.attention_kv_cache_get_total_sequence_length"](cache)) == 32 + 301 + 6 + 121
    assert int(f["vm.builtin.attention_kv_cache_get_num_available_pages"](cache)) > 0
    f["vm.builtin.kv_state_remove_sequence"](cache, 2)
    f["vm.builtin.kv_state_clear"](cache)
    assert bool(f["vm.builtin.attention_kv_cache_empty"](cache))
