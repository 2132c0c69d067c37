"""tvm_b200 -- B200-native (sm_100a) PagedKVCache attention kernel set behind the reference's callbacks.

The package holds only what the hot path needs:
  csrc/        hand-written CUDA kernels + the C ABI (include/tvm_b200.h) + tvm-ffi packed functions
  capi.py      ctypes binding of the C ABI (plain pointers and sizes)
  ffi.py       the same library loaded as a tvm-ffi module (`f_attention_decode`, ... packed functions)
  kv_cache.py  host mirror of the reference's PagedAttentionKVCacheObj (src/runtime/vm/paged_kv_cache.cc)
  build.py     in-tree nvcc build (tvm_b200/lib/libtvm_b200.so)

There is no CPU fallback: every compute entry point fails loudly without the CUDA library / a GPU.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
__version__ = "0.1"
