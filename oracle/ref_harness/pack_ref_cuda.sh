#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- assembles what the GPU box needs to run the UNMODIFIED reference next to tvm_b200:
#   oracle/_ref/tvm_cuda/lib/   libtvm_ffi.so, libtvm_runtime.so, libtvm_runtime_cuda.so, libtvm_runtime_extra.so
#                               (built by build_tvm_cuda.sh from /root/reference; the KV cache object lives in _extra)
#   oracle/_ref/tvm_cuda/py/    the vendored tvm-ffi 0.1.14 Python package (build_tvm.sh), bound to that libtvm_ffi.so
#   oracle/_ref/ref_gpu_kernels_{float16,bfloat16}_hq32_hkv8_d128.so   the reference's GPU TIR kernels (sm_100a fatbin)
# oracle/_ref/ is git-ignored (built artefacts only) but travels to the GPU box with the snapshot.  Nothing under
# tvm_b200/ ever loads it; oracle/ref_gpu_server.py (a separate process: two libtvm_ffi versions must not share one) does.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
CUDA_BUILD=${1:-/tmp/tvm_ref_cuda}
CPU_BUILD=${2:-/tmp/tvm_ref}
DST="$HERE/../_ref/tvm_cuda"
rm -rf "$DST"; mkdir -p "$DST/lib" "$DST/py"
for l in libtvm_ffi.so libtvm_ffi_testing.so libtvm_runtime.so libtvm_runtime_cuda.so libtvm_runtime_extra.so; do
  cp "$CUDA_BUILD/build/lib/$l" "$DST/lib/"
  strip --strip-unneeded "$DST/lib/$l" || true
done
cp -r "$CPU_BUILD/ref_py/tvm_ffi" "$DST/py/tvm_ffi"
cp -r "$CPU_BUILD/ref_py/apache_tvm_ffi-0.1.14.dist-info" "$DST/py/"
rm -rf "$DST/py/tvm_ffi/include" "$DST/py/tvm_ffi/__pycache__"
cp "$DST/lib/libtvm_ffi.so" "$DST/py/tvm_ffi/lib/libtvm_ffi.so"
# the reference's GPU kernels: compiled here by the reference's DEFAULT backend (NVRTC -> cubin, no fast-math; its nvcc
# mode adds --use_fast_math, whose __sinf / __cosf lose the RoPE angle at large positions); a stub libcuda lets the
# reference's CUDA module factory load without a GPU
STUB=$(mktemp -d); ln -sf /usr/local/cuda/lib64/stubs/libcuda.so "$STUB/libcuda.so.1"
( cd /tmp && LD_LIBRARY_PATH="$STUB:${LD_LIBRARY_PATH:-}" TVM_CUDA_COMPILE_MODE=nvrtc TVM_LIBRARY_PATH="$CUDA_BUILD/build/lib" \
  PYTHONPATH="$CPU_BUILD/ref_py:/root/reference/python" python "$HERE/emit_ref_gpu_kernels.py" 2>&1 | grep -v "Cannot parse Arm" )
du -sh "$DST" "$HERE"/../_ref/*.so
