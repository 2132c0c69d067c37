"""TEST INFRASTRUCTURE ONLY -- runs INSIDE the reference process (source /tmp/tvm_ref/env.sh first): loads
tvm_b200/lib/libtvm_b200.so with the reference's own runtime (vendored tvm-ffi, libtvm_runtime loaded, the vm.builtin.* names
already registered by src/runtime/vm/kv_state.cc) and checks INTEGRATION.md routes A and A2 as far as a CPU-only build allows:
the packed functions resolve under the reference's callback names, argument errors surface as the reference-style
exceptions, register_vm_builtins(0) is refused while the reference's registrations exist, register_vm_builtins(1) replaces
them, and the replaced entries are tvm_b200's (no CPU fallback: a CPU `init` tensor is refused)."""
import os
import sys

import numpy as np
import tvm
import tvm_ffi

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
LIB = os.path.join(ROOT, "tvm_b200", "lib", "libtvm_b200.so")


def expect(exc_types, fragment, fn, *args):
    try:
        fn(*args)
    except exc_types as e:  # noqa: PERF203
        assert fragment in str(e), f"expected {fragment!r} in {e!r}"
        return
    raise AssertionError(f"no error raised (expected {fragment!r})")


def main():
    assert tvm.get_global_func("vm.builtin.paged_attention_kv_cache_create", allow_missing=True) is not None
    mod = tvm.runtime.load_module(LIB)
    for name in ["f_transpose_append", "f_attention_decode", "f_attention_prefill", "f_attention_prefill_ragged",
                 "f_merge_inplace", "f_split_rotary", "f_copy_single_page", "f_compact_copy", "f_debug_get_kv",
                 "f_attention_prefill_with_tree_mask", "f_attention_prefill_with_tree_mask_paged_kv",
                 "f_attention_decode_sliding_window", "f_attention_prefill_sliding_window", "batch_decode_paged_kv",
                 "fused_rope", "launch_count", "set_rope_params", "set_rope_scaling", "set_rope_scaling_yarn",
                 "register_vm_builtins"]:
        assert mod.get_function(name) is not None, name
    assert int(mod["launch_count"]()) == 0
    mod["set_rope_params"](1e4, 1.0)
    cpu = tvm.runtime.tensor(np.zeros((2, 1, 128), "float16"))
    pos = tvm.runtime.tensor(np.zeros((2,), "int32"))
    pages = tvm.runtime.tensor(np.zeros((1, 2, 1, 16, 128), "float16"))
    expect((ValueError, RuntimeError), "CUDA", mod["f_transpose_append"], pages, cpu, cpu, pos)   # no CPU fallback
    expect((TypeError, ValueError), "expects 4 arguments", mod["f_transpose_append"], pages, cpu, cpu)
    reg = mod["register_vm_builtins"]
    expect(Exception, "already registered", reg, 0)
    n = int(reg(1))
    assert n >= 24, n
    add = tvm.get_global_func("vm.builtin.kv_state_add_sequence")
    expect(TypeError, "cache returned by vm.builtin.paged_attention_kv_cache_create of tvm_b200", add, 3, 0)
    create = tvm.get_global_func("vm.builtin.paged_attention_kv_cache_create")
    S = tvm_ffi.Shape
    expect(Exception, "no CPU fallback", create, S([4, 256, 128, 16, 0]), S([0, 1]), 32, 8, 128, 128, S([0]), False, 1, 1.0, 1e4,
           None, tvm.runtime.tensor(np.zeros((), "float16")), *([None] * 15))
    print(f"reference runtime ok: tvm {tvm.__version__}, tvm_ffi {tvm_ffi.__version__}, {n} vm.builtin names replaced")


if __name__ == "__main__":
    sys.exit(main())
