"""The C-ABI library loads without a GPU and exports every symbol declared in include/*.h, plus the tvm-ffi packed
functions under the reference's callback names.  No compute call is made here."""
import ctypes

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
