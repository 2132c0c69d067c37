"""TEST INFRASTRUCTURE ONLY -- runs INSIDE the reference process (source /tmp/tvm_ref/env.sh first).

Builds the reference's own CPU TIR PrimFuncs (python/tvm/relax/frontend/nn/llm/*) with the reference's own
`c` target + gcc (the image has no LLVM, so BASELINE config 1's `llvm` target cannot exist),
and constructs the reference's own C++ PagedAttentionKVCacheObj around them, optionally wrapping every
callback in a logging closure so that the int32 aux arrays it receives can be captured as golden vectors.
No reference source is copied: everything is imported from /root/reference at run time.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np
import tvm
import tvm_ffi
from tvm.relax.frontend.nn.llm.kv_cache import (
    AttnKind,
    _attention_decode_cpu,
    _attention_prefill_cpu,
    _attention_prefill_ragged_cpu,
    _compact_kv_copy_cpu,
    _copy_single_page_cpu,
    _kv_cache_debug_get_kv,
    _kv_cache_transpose_append,
    _merge_state_inplace_cpu,
    llama_rope_with_position_map,
    tree_attn_cpu,
    tree_attn_with_paged_kv_cache_cpu,
)

# shim 2 of SURVEY 8(c): the `c` target has no lowering rule for tirx.exp2
tvm.ir.register_intrin_lowering(
    "tirx.exp2", target="c", f=lambda op: tvm.tirx.call_pure_extern("float32", "exp2f", op.args[0]), level=99
)

CALLBACK_ORDER = [
    "transpose_append", "debug_get_kv", "prefill", "decode", "prefill_sliding_window", "decode_sliding_window",
    "prefill_ragged", "tree_ragged", "tree_paged", "merge", "split_rotary", "copy_single_page", "compact_copy",
]


def kernel_primfuncs(num_layers, num_qo_heads, num_kv_heads, head_dim, dtype, rope_theta, rope_scale, page_size=16,
                     layer_sliding_window_size=1024):
    rs = {}
    return [
        _kv_cache_transpose_append(num_kv_heads, head_dim, dtype),
        _kv_cache_debug_get_kv(num_layers, num_kv_heads, head_dim, dtype),
        _attention_prefill_cpu(num_kv_heads, num_qo_heads, head_dim, dtype, False, rs),
        _attention_decode_cpu(num_kv_heads, num_qo_heads, head_dim, dtype, False, rs),
        _attention_prefill_cpu(num_kv_heads, num_qo_heads, head_dim, dtype, True, rs,
                               sliding_window_size=layer_sliding_window_size),
        _attention_decode_cpu(num_kv_heads, num_qo_heads, head_dim, dtype, True, rs),
        _attention_prefill_ragged_cpu(num_kv_heads, num_qo_heads, head_dim, head_dim, dtype, rs),
        tree_attn_cpu(num_kv_heads, num_qo_heads, head_dim, dtype, rs),
        tree_attn_with_paged_kv_cache_cpu(num_kv_heads, num_qo_heads, head_dim, dtype, rs),
        _merge_state_inplace_cpu(dtype),
        llama_rope_with_position_map(rope_theta, rope_scale, head_dim, num_qo_heads, num_kv_heads, dtype, rs),
        _copy_single_page_cpu(num_kv_heads, page_size, head_dim, dtype),
        _compact_kv_copy_cpu(num_kv_heads, head_dim, dtype),
    ]


def build_c(primfunc, out_path=None):
    """tvm.tirx.build(target='c') + export_library(gcc -O3 -Dhalf=_Float16) + load; returns (fn, so_path, symbol)."""
    target = tvm.target.Target("c")
    mod = tvm.IRModule({"main": primfunc})
    lib = tvm.tirx.build(mod["main"], target=target)
    if out_path is None:
        out_path = os.path.join(tempfile.mkdtemp(prefix="refk_"), "k.so")
    lib.export_library(out_path, options=["-O3", "-Dhalf=_Float16", "-lm"])
    loaded = tvm.runtime.load_module(out_path)
    sym = primfunc.attrs["global_symbol"] if primfunc.attrs and "global_symbol" in primfunc.attrs else "main"
    try:
        fn = loaded.get_function(sym)
    except Exception:
        fn = loaded.get_function("main")
    return fn, out_path, sym, loaded


class RefCache:
    """The reference's C++ cache over its own CPU kernels, every callback wrapped in a logging closure."""

    def __init__(self, num_layers=1, num_qo_heads=8, num_kv_heads=2, head_dim=128, dtype="float16", rope_mode=1,
                 support_sliding_window=False, reserved_nseq=32, max_total_seq=2048, prefill_chunk=512, page_size=16,
                 rope_scale=1.0, rope_theta=1e4, layer_sliding_window_size=None, attn_kinds=None, kernels=None,
                 layer_begin=0):
        self.cfg = dict(num_layers=num_layers, num_qo_heads=num_qo_heads, num_kv_heads=num_kv_heads,
                        head_dim=head_dim, dtype=dtype, rope_mode=int(rope_mode),
                        support_sliding_window=int(support_sliding_window), reserved_nseq=reserved_nseq,
                        max_total_seq=max_total_seq, prefill_chunk=prefill_chunk, page_size=page_size,
                        rope_scale=rope_scale, rope_theta=rope_theta)
        self.dev = tvm.cpu()
        self.trace = []
        self._keep = []
        pfs = None if kernels is not None else kernel_primfuncs(num_layers, num_qo_heads, num_kv_heads, head_dim, dtype, rope_theta, rope_scale,
                               page_size, layer_sliding_window_size or 1024)
        # kernels: the compiled callbacks (raw_fns) of an earlier RefCache with the same shape configuration
        if kernels is None:
            kernels = {}
            for name, pf in zip(CALLBACK_ORDER, pfs):
                fn, _, _, mod = build_c(pf)
                kernels[name] = (fn, mod)
        self.raw_fns = kernels
        fns = {name: self._wrap(name, fn) for name, (fn, _mod) in kernels.items()}
        self.fns = fns
        g = tvm.get_global_func
        self.f = {n: g("vm.builtin." + n) for n in [
            "kv_state_clear", "kv_state_add_sequence", "kv_state_remove_sequence", "kv_state_fork_sequence",
            "kv_state_popn", "kv_state_begin_forward", "kv_state_end_forward",
            "attention_kv_cache_enable_sliding_window_for_seq", "attention_kv_cache_commit_accepted_token_tree_nodes",
            "attention_kv_cache_attention_with_fused_qkv", "attention_kv_cache_empty",
            "attention_kv_cache_get_num_available_pages", "attention_kv_cache_get_total_sequence_length",
            "attention_kv_cache_debug_get_kv", "attention_kv_cache_get_query_positions",
            "attention_kv_cache_self_attention", "attention_kv_cache_cross_attention",
            "attention_kv_cache_attention_with_shared_kv", "attention_kv_cache_merge_attn_output_inplace"]}
        cache_config = [reserved_nseq, max_total_seq, prefill_chunk, page_size, int(support_sliding_window)]
        if layer_sliding_window_size is not None:
            cache_config.append(layer_sliding_window_size)
        if attn_kinds is None:
            attn_kinds = [int(AttnKind.MHA)] * (layer_begin + num_layers)
        self.cache = g("vm.builtin.paged_attention_kv_cache_create")(
            tvm_ffi.Shape(cache_config), tvm_ffi.Shape([layer_begin, layer_begin + num_layers]), num_qo_heads, num_kv_heads, head_dim, head_dim,
            tvm_ffi.Shape(attn_kinds), False, int(rope_mode), rope_scale, rope_theta, None,
            tvm.runtime.empty((), dtype, device=self.dev),
            fns["transpose_append"], None, ["tirx", fns["prefill_ragged"]], ["tirx", fns["prefill"]],
            ["tirx", fns["decode"]], ["tirx", fns["prefill_sliding_window"]], ["tirx", fns["decode_sliding_window"]],
            ["tirx", fns["tree_paged"]], ["tirx", fns["tree_ragged"]], [], [fns["merge"], fns["merge"]], fns["split_rotary"],
            fns["copy_single_page"], fns["debug_get_kv"], fns["compact_copy"])

    def _wrap(self, name, fn):
        trace = self.trace

        def logged(*args):
            rec = {"fn": name, "args": []}
            for a in args:
                if hasattr(a, "numpy") and hasattr(a, "shape"):
                    arr = a.numpy() if str(a.dtype) == "int32" else None
                    rec["args"].append({"t": str(a.dtype), "shape": [int(s) for s in a.shape],
                                        "v": arr.reshape(-1).tolist() if arr is not None else None})
                else:
                    rec["args"].append({"s": float(a) if isinstance(a, float) else int(a)})
            trace.append(rec)
            return fn(*args)

        return tvm_ffi.convert(logged)

    # -- thin API mirroring vm.builtin.* -------------------------------------------------------------------------
    def call(self, name, *args):
        return self.f[name](self.cache, *args)
