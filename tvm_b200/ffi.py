"""The sm_100a library as a tvm-ffi module: packed functions under the reference's callback names.

`module()["f_attention_decode"]` etc. are `tvm_ffi.Function`s with the exact positional signatures the
reference's PagedAttentionKVCacheObj calls (src/runtime/vm/attn_backend.h:234-243, 388-396, 507-515,
618-627, 665-673; paged_kv_cache.cc:1360-1373, 728, 759, 1718, 2292), so they can be handed to
`vm.builtin.paged_attention_kv_cache_create` as `["tirx", fn]` tuples.
Kernels launch on the tvm-ffi environment stream (TVMFFIEnvGetStream), like the reference's kernels.
"""
from __future__ import annotations

from . import capi

CALLBACKS = [
    "f_transpose_append", "f_attention_decode", "f_attention_decode_sliding_window", "f_attention_prefill",
    "f_attention_prefill_sliding_window", "f_attention_prefill_ragged", "f_attention_prefill_with_tree_mask",
    "f_attention_prefill_with_tree_mask_paged_kv", "f_merge_inplace", "f_split_rotary", "f_copy_single_page",
    "f_debug_get_kv", "f_compact_copy",
]
# global_symbols of the reference PrimFuncs the callbacks replace
TIR_NAMES = [
    "tir_kv_cache_transpose_append", "batch_decode_paged_kv", "batch_decode_paged_kv_sliding_window",
    "batch_prefill_paged_kv", "batch_prefill_paged_kv_sliding_window", "batch_prefill_ragged_kv", "batch_tree_attn",
    "tree_attn_paged_kv", "merge_state_inplace", "fused_rope", "copy_single_page", "tir_kv_cache_debug_get_kv",
    "compact_kv_copy",
]
STATE = ["set_rope_params", "set_rope_scaling", "set_rope_scaling_yarn", "set_layer_sliding_window_size", "launch_count", "register_vm_builtins"]
CONTEXT = ["context_create", "context_release", "bind_context"]

_mod = None


def module():
    """tvm_ffi.load_module(libtvm_b200.so); fails loudly when the library is not built."""
    global _mod
    if _mod is None:
        import tvm_ffi

        if not capi.LIB_PATH.exists():
            raise capi.TvmB200Error(f"{capi.LIB_PATH} is missing: run `python -m tvm_b200.build` (no CPU fallback)")
        _mod = tvm_ffi.load_module(str(capi.LIB_PATH))
    return _mod


def torch_stream(ctx=None):
    """Context manager that makes the tvm-ffi environment stream follow torch's current stream."""
    import tvm_ffi

    return tvm_ffi.use_torch_stream(ctx) if ctx is not None else tvm_ffi.use_torch_stream()


def register_globals(prefix: str = "tvm_b200.") -> None:
    """Register every callback as a tvm-ffi global function `<prefix><name>` (how a Relax-VM process
    would look them up: tvm_ffi.get_global_func)."""
    import tvm_ffi

    m = module()
    for name in CALLBACKS + STATE:
        tvm_ffi.register_global_func(prefix + name, m[name], override=True)


class KernelSet:
    """One kernel set = the 13 callbacks (plus their TIR-named twins) bound to their own context, the counterpart of the
    reference compiling its PrimFuncs with a `rope_scaling` dict, `rotary_dim`, `rope_theta` / `rope_scale` and
    `layer_sliding_window_size` baked in (kv_cache.py:690-736).  Two kernel sets never share settings or device
    scratch, so two models -- or two streams -- can use the library concurrently in one process.

        ks = KernelSet(rope_theta=5e5, rope_scaling={"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0,
                                                      "high_freq_factor": 4.0, "original_max_position_embeddings": 8192})
        ks["f_attention_decode"](q, pages, ...)
    """

    def __init__(self, rope_theta: float = 1e4, rope_scale: float = 1.0, rotary_dim: int = 0, rope_scaling: dict | None = None,
                 layer_sliding_window_size: int = 1024):
        m = module()
        self._release = m["context_release"]
        self._ctx = m["context_create"]()
        self._bind = m["bind_context"]
        self._fns = {}
        self["set_rope_params"](float(rope_theta), float(rope_scale), int(rotary_dim))
        self["set_layer_sliding_window_size"](int(layer_sliding_window_size))
        rs = dict(rope_scaling or {})
        kind = rs.get("rope_type", rs.get("type", "default"))
        if kind in ("default", None, "none"):
            self["set_rope_scaling"](0, 1.0, 0.0, 0.0, 0.0)
        elif kind in ("llama3", "llama4", "gptj"):
            code = {"llama3": 1, "gptj": 2, "llama4": 3}[kind]
            self["set_rope_scaling"](code, float(rs.get("factor", 1.0)), float(rs.get("low_freq_factor", 0.0)),
                                     float(rs.get("high_freq_factor", 0.0)),
                                     float(rs.get("original_max_position_embeddings", 0.0)))
        elif kind == "yarn":
            self["set_rope_scaling_yarn"](float(rs["factor"]), float(rs["original_max_position_embeddings"]),
                                          float(rs.get("beta_fast", 32.0)), float(rs.get("beta_slow", 1.0)),
                                          float(rs.get("inv_theta_log_scale", 0.0)))
        else:
            raise capi.TvmB200Error(f"rope_scaling type {kind!r} is not supported (default, llama3, gptj, llama4, yarn)")

    def __getitem__(self, name: str):
        f = self._fns.get(name)
        if f is None:
            f = self._fns[name] = self._bind(self._ctx, name)
        return f

    def callbacks(self) -> dict:
        return {n: self[n] for n in CALLBACKS}

    def __del__(self):
        try:
            self._release(self._ctx)
        except Exception:  # interpreter shutdown
            pass
