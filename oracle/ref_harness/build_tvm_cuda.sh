#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference (apache/tvm at /root/reference) WITH its CUDA runtime and
# codegen into a scratch directory (default /tmp/tvm_ref_cuda), ~27 min on 8 cores, no GPU needed.  Used for
#   (1) oracle/_ref/ref_gpu_kernels_*.so = the reference's own GPU TIR kernels for sm_100a (emit_ref_gpu_kernels.py), and
#   (2) oracle/_ref/tvm_cuda/ = the reference's runtime libraries (C++ PagedAttentionKVCacheObj + CUDA device API) that
#       travel to the GPU box, where oracle/ref_gpu_server.py drives them (pack_ref_cuda.sh).
# The CPU-only build of build_tvm.sh (/tmp/tvm_ref) supplies the tvm-ffi Python package; run that first.
set -euo pipefail
REF=${REF:-/root/reference}
OUT=${1:-/tmp/tvm_ref_cuda}
mkdir -p "$OUT"
if [ ! -f "$OUT/build/lib/libtvm_runtime_cuda.so" ]; then
  cmake -S "$REF" -B "$OUT/build" -G Ninja -DCMAKE_BUILD_TYPE=Release -DUSE_LLVM=OFF -DUSE_CUDA=ON -DUSE_NCCL=OFF \
        -DCMAKE_CUDA_ARCHITECTURES=100a -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc \
        -DUSE_GTEST=OFF -DUSE_Z3=OFF -DUSE_CCACHE=OFF -DUSE_RPC=OFF
  ninja -C "$OUT/build" -j"${JOBS:-$(nproc)}"
fi
echo "reference CUDA build ready: $OUT/build/lib"
