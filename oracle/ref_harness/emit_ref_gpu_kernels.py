"""TEST / BENCH INFRASTRUCTURE ONLY -- emits oracle/_ref/ref_gpu_kernels_<dtype>_hq<..>_hkv<..>_d<..>.so:
the reference's OWN GPU TIR PrimFuncs of the hot path (python/tvm/relax/frontend/nn/llm/: _attention_decode
_decode_kernels.py:181-411, _attention_prefill / _attention_prefill_ragged _prefill_kernels.py:217-391, 795-923,
tree_attn / tree_attn_with_paged_kv_cache tree_attn.py:264-603, 798-1259, _merge_state_inplace, _kv_cache_transpose_append,
llama_rope_with_position_map, copy / compact kernels), scheduled and built exactly like the reference's own GPU test does
(tests/python/relax/test_runtime_builtin_paged_attention_kv_cache_tir.py:203-241: dl.gpu.Fallback for the unscheduled
ones, tvm.tirx.build) for sm_100a, in float16 AND bfloat16.  These are what the north_star tolerance (2e-3 / 1e-2) is
defined against and the "kernel to beat" of bench.py's ref_gpu sub-records.

Needs the USE_CUDA=ON build of the reference (oracle/ref_harness/build_tvm_cuda.sh); no GPU is needed to compile, only a
loadable libcuda (the toolkit's stub) so that the reference's CUDA module factory compiles the source (NVRTC, its default) instead
of storing it:  oracle/ref_harness/pack_ref_cuda.sh   (runs this script with the right environment)
Only the resulting .so files (git-ignored, they travel to the GPU box) are kept; oracle/ref_gpu_server.py loads them inside
the reference's own runtime (oracle/_ref/tvm_cuda, packed by pack_ref_cuda.sh)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "_ref")

import tvm  # noqa: E402
from tvm.s_tir import dlight as dl  # noqa: E402
from tvm.relax.frontend.nn.llm.kv_cache import (  # noqa: E402
    _attention_decode,
    _attention_prefill,
    _attention_prefill_ragged,
    _compact_kv_copy,
    _copy_single_page,
    _kv_cache_debug_get_kv,
    _kv_cache_transpose_append,
    _merge_state_inplace,
    llama_rope_with_position_map,
    tree_attn,
    tree_attn_with_paged_kv_cache,
)


def emit(dtype, hq=32, hkv=8, d=128, theta=5e5, scale=1.0, page_size=16, num_layers=1, layer_sws=1024):
    target = tvm.target.Target({"kind": "cuda", "arch": "sm_100a", "max_threads_per_block": 1024,
                                "max_shared_memory_per_block": 49152, "thread_warp_size": 32})
    rs = {}
    named = {
        "tir_kv_cache_transpose_append": _kv_cache_transpose_append(hkv, d, dtype),
        "tir_kv_cache_debug_get_kv": _kv_cache_debug_get_kv(num_layers, hkv, d, dtype),
        "batch_prefill_paged_kv": _attention_prefill(hkv, hq, d, dtype, False, rs, target),
        "batch_decode_paged_kv": _attention_decode(hkv, hq, d, dtype, False, rs, target),
        "batch_prefill_paged_kv_sliding_window": _attention_prefill(hkv, hq, d, dtype, True, rs, target,
                                                                    sliding_window_size=layer_sws),
        "batch_decode_paged_kv_sliding_window": _attention_decode(hkv, hq, d, dtype, True, rs, target),
        "batch_prefill_ragged_kv": _attention_prefill_ragged(hkv, hq, d, d, dtype, rs, target),
        "batch_tree_attn": tree_attn(hkv, hq, d, dtype, rs, target),
        "tree_attn_paged_kv": tree_attn_with_paged_kv_cache(hkv, hq, d, dtype, rs, target),
        "merge_state_inplace": _merge_state_inplace(hq, d, dtype, target),
        "fused_rope": llama_rope_with_position_map(theta, scale, d, hq, hkv, dtype, rs),
        "copy_single_page": _copy_single_page(hkv, page_size, d, dtype, target),
        "compact_kv_copy": _compact_kv_copy(hkv, d, dtype, target),
    }
    funcs = {}
    for name, pf in named.items():
        m = tvm.IRModule({"main": pf})
        with target:
            m = dl.ApplyDefaultSchedule(dl.gpu.Fallback())(m)
        funcs[name] = m["main"].with_attr("global_symbol", name)
    mod = tvm.IRModule(funcs)
    with target:  # the compile callback reads the arch from the current target scope (python/tvm/support/nvcc.py:893)
        lib = tvm.tirx.build(mod, target=target)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"ref_gpu_kernels_{dtype}_hq{hq}_hkv{hkv}_d{d}.so")
    # no LLVM in the image: the host side is C source (target `c`), which spells the 16-bit types `half` / `bfloat16`
    lib.export_library(path, options=["-O2", "-Dhalf=_Float16", "-Dbfloat16=__bf16"])
    print("wrote", path, os.path.getsize(path))
    return lib


if __name__ == "__main__":
    for dt in (sys.argv[1:] or ["float16", "bfloat16"]):
        emit(dt)
