"""TEST / BENCH INFRASTRUCTURE ONLY -- emits oracle/_ref/ref_kernels_<dtype>_hq<..>_hkv<..>_d<..>.so:
the reference's OWN CPU TIR PrimFuncs of the hot path (llama_rope_with_position_map, _kv_cache_transpose_append,
_attention_decode_cpu, _attention_prefill_ragged_cpu, _merge_state_inplace_cpu, and the paged / sliding-window / tree-mask
prefill and page-copy kernels), compiled
from /root/reference by the reference's own `c` target and gcc -O3 -Dhalf=_Float16 (no LLVM in the image).
Run inside the reference env:  source /tmp/tvm_ref/env.sh && python oracle/ref_harness/emit_ref_kernels.py
Only the resulting .so (git-ignored) is kept; it is loaded by oracle/cpu_ref.py through tvm-ffi as the CPU baseline."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tvm  # noqa: E402
from refenv import kernel_primfuncs  # noqa: E402  (also registers the exp2 lowering shim)

OUT = os.path.join(HERE, "..", "_ref")


def emit(dtype="float16", hq=32, hkv=8, d=128, theta=5e5, scale=1.0):
    pfs = kernel_primfuncs(1, hq, hkv, d, dtype, theta, scale)
    # indices = refenv.CALLBACK_ORDER; the first five are the decode-step / prefill hot path bench.py times, the rest pin
    # the oracle at this head shape (tests/test_ref_kernels.py)
    names = {0: "tir_kv_cache_transpose_append", 3: "batch_decode_paged_kv_cpu", 6: "batch_prefill_ragged_kv_cpu",
             9: "merge_state_inplace_cpu", 10: "fused_rope", 2: "batch_prefill_paged_kv_cpu",
             4: "batch_prefill_paged_kv_sliding_window_cpu", 5: "batch_decode_paged_kv_sliding_window_cpu",
             7: "batch_tree_attn_cpu", 8: "tree_attn_paged_kv_cpu", 1: "tir_kv_cache_debug_get_kv",
             11: "copy_single_page_cpu", 12: "compact_kv_copy_cpu"}
    funcs = {}
    for i, nm in names.items():
        funcs[nm] = pfs[i].with_attr("global_symbol", nm)
    mod = tvm.IRModule(funcs)
    lib = tvm.tirx.build(mod, target=tvm.target.Target("c"))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"ref_kernels_{dtype}_hq{hq}_hkv{hkv}_d{d}.so")
    lib.export_library(path, options=["-O3", "-Dhalf=_Float16", "-lm"])
    print("wrote", path)


if __name__ == "__main__":
    emit()
