"""The C-ABI library loads without a GPU and exports every symbol declared in include/*.h, plus the tvm-ffi packed
functions under the reference's callback names.  No compute call is made here."""
import ctypes
import re

import pytest


def test_exports_every_declared_symbol(built_lib):
    from tvm_b200 import capi

    L = capi.lib()
    names = capi.declared_symbols()
    assert len(names) >= 35 and "tvmb200_attention_decode" in names and "tvmb200_cache_begin_forward" in names
    for n in names:
        assert getattr(L, n) is not None, n
    assert b"sm_100a" in L.tvmb200_version()


def test_tvm_ffi_module_exports_reference_callback_names(built_lib):
    from tvm_b200 import ffi

    m = ffi.module()
    for n in ffi.CALLBACKS + ffi.TIR_NAMES + ffi.STATE:
        assert m[n] is not None
    assert int(m["launch_count"]()) >= 0


def test_compute_fails_loudly_without_gpu(built_lib):
    import torch

    from tvm_b200 import capi, ffi

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cpu = torch.zeros((3, 2, 8, 16, 128), dtype=torch.float16)
    with pytest.raises(capi.TvmB200Error, match="no CPU fallback"):
        capi.transpose_append(cpu, cpu, cpu, cpu)
    pm = torch.zeros((3,), dtype=torch.int32)
    kv = torch.zeros((3, 8, 128), dtype=torch.float16)
    with pytest.raises(Exception, match="CUDA device"):
        ffi.module()["f_transpose_append"](cpu, kv, kv, pm)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from tvm_b200 import capi

    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(capi.TvmB200Error, match="no CPU fallback"):
        capi.lib()


def test_packed_functions_validate_like_the_tir_binders(built_lib):
    """The generated TIR binders of the reference check argument count, kind, dtype and shape before anything runs
    (src/tirx/transform/tvm_ffi_binder.cc:704-807); the packed functions do the same, ahead of the device check -- so this
    runs without a GPU."""
    import torch

    from tvm_b200 import ffi

    m = ffi.module()
    f16 = lambda *s: torch.zeros(s, dtype=torch.float16)  # noqa: E731
    i32 = lambda *s: torch.zeros(s, dtype=torch.int32)  # noqa: E731
    f32 = lambda *s: torch.zeros(s, dtype=torch.float32)  # noqa: E731
    pages = f16(2, 2, 1, 16, 128)
    q, o, lse = f16(2, 4, 128), f16(2, 4, 128), f32(2, 4)
    dec = (q, pages, i32(3), i32(2), i32(2), i32(2), i32(2), o, lse)
    cases = [
        (TypeError, "expects 4 arguments, got 3", "f_transpose_append", (pages, f16(3, 1, 128), f16(3, 1, 128))),
        (TypeError, "must be a Tensor", "f_transpose_append", (pages, f16(3, 1, 128), f16(3, 1, 128), 5)),
        (ValueError, "dtype mismatch: k_data vs pages", "f_transpose_append", (pages, f32(3, 1, 128), f16(3, 1, 128), i32(3))),
        (ValueError, "position_map must be int32", "f_transpose_append", (pages, f16(3, 1, 128), f16(3, 1, 128), f32(3))),
        (ValueError, "shape mismatch", "f_transpose_append", (pages, f16(3, 1, 64), f16(3, 1, 128), i32(3))),
        (TypeError, "expects 13 arguments, got 12", "f_attention_decode", dec + (0, 1.0, 1e4)),
        (ValueError, "lse must be float32", "f_attention_decode", dec[:8] + (f16(2, 4), 0, 1.0, 1e4, 0.08)),
        (TypeError, "rotary_mode) must be an int", "f_attention_decode", dec + ("x", 1.0, 1e4, 0.08)),
        (ValueError, "shape mismatch", "f_merge_inplace", (o, lse, o, f32(2, 5))),
        (ValueError, "rope_ext_factors (longrope", "f_split_rotary",
         (f16(2, 6, 128), i32(2), f16(2, 4, 128), f16(2, 1, 128), f16(2, 1, 128), f32(128))),
        # everything valid: the device check is what is left
        (ValueError, "no CPU fallback", "f_attention_decode", dec + (0, 1.0, 1e4, 0.08)),
        (ValueError, "no CPU fallback", "f_split_rotary", (f16(2, 6, 128), i32(2), f16(2, 4, 128), f16(2, 1, 128), f16(2, 1, 128), 1)),
    ]
    for exc, fragment, name, args in cases:
        with pytest.raises(exc, match=re.escape(fragment)):
            m[name](*args)


LLAMA31 = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
           "original_max_position_embeddings": 8192}


def test_contexts_keep_their_own_rope_scaling(built_lib):
    """A kernel-set context carries what the reference compiles into one set of PrimFuncs (include/tvm_b200.h): the
    setters act on the calling thread's current context; a host cache snapshots the settings current at its creation."""
    from tvm_b200 import capi
    from tvm_b200.kv_cache import PagedKVCache

    L = capi.lib()
    kind = lambda: int(L.tvmb200_get_rope_scaling_kind())  # noqa: E731
    capi.set_rope_scaling(None)
    ctx = capi.Context()
    with ctx:
        capi.set_rope_scaling(LLAMA31)
        assert kind() == 1
        inner = capi.Context()            # a new context copies its creator's current settings
    assert kind() == 0                    # the default context never saw it
    with inner:
        assert kind() == 1
    try:
        capi.set_rope_scaling({"rope_type": "gptj"})
        cache = PagedKVCache(reserved_num_seqs=2, total_token_capacity=64, prefill_chunk_size=16, num_layers=1,
                             num_qo_heads=4, num_kv_heads=1, head_dim=128, device=None)
    finally:
        capi.set_rope_scaling(None)
    with cache.context():
        assert kind() == 2                # the cache kept the scaling that was current when it was created
        capi.set_rope_scaling(LLAMA31)    # ... and can be given its own
    with cache.context():
        assert kind() == 1
    assert kind() == 0


def test_kernel_set_binds_packed_functions_to_a_context(built_lib):
    import torch

    from tvm_b200 import ffi

    a = ffi.KernelSet(rope_theta=5e5, rope_scaling=LLAMA31)
    b = ffi.KernelSet()
    assert set(a.callbacks()) == set(ffi.CALLBACKS)
    with pytest.raises(ValueError, match="exports no packed function"):
        a["f_no_such_thing"]
    with pytest.raises(Exception, match="not supported"):
        ffi.KernelSet(rope_scaling={"rope_type": "longrope"})
    # bound functions validate exactly like the module symbols
    with pytest.raises(TypeError, match="expects 4 arguments, got 1"):
        b["f_transpose_append"](torch.zeros(1))
    if not torch.cuda.is_available():
        f16 = lambda *s: torch.zeros(s, dtype=torch.float16)  # noqa: E731
        with pytest.raises(ValueError, match="no CPU fallback"):
            a["f_split_rotary"](f16(2, 6, 128), torch.zeros(2, dtype=torch.int32), f16(2, 4, 128), f16(2, 1, 128), f16(2, 1, 128), 1)
    del a, b  # releases the contexts (last reference frees their scratch)
