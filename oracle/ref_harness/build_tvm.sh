#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference (apache/tvm at
# /root/reference) into a scratch directory so that its own CPU TIR kernels and
# its own C++ PagedAttentionKVCacheObj can be run in this container to
#   (1) generate the golden fixtures under tests/golden/ (gen_golden.py), and
#   (2) emit oracle/_ref/*.so = the reference's CPU PrimFuncs compiled by the
#       reference's own `c` target (emit_ref_kernels.py).
# Nothing here is part of the product path; no reference source is copied into
# the repository.  Recipe verified in SURVEY.md section 8(c) addendum.
#
# usage: oracle/ref_harness/build_tvm.sh [scratch_dir]   (default /tmp/tvm_ref)
set -euo pipefail
REF=${REF:-/root/reference}
OUT=${1:-/tmp/tvm_ref}
PY=$(command -v python3)
mkdir -p "$OUT"

# 1. libtvm_{compiler,runtime,runtime_extra,ffi}.so, CPU only, no LLVM (~24 min on 8 cores)
if [ ! -f "$OUT/build/lib/libtvm_runtime_extra.so" ] && [ ! -f "$OUT/build/libtvm_runtime_extra.so" ]; then
  cmake -S "$REF" -B "$OUT/build" -G Ninja -DCMAKE_BUILD_TYPE=Release \
        -DUSE_LLVM=OFF -DUSE_CUDA=OFF -DUSE_GTEST=OFF -DUSE_Z3=OFF -DUSE_CCACHE=OFF -DUSE_RPC=OFF
  ninja -C "$OUT/build" -j"${JOBS:-$(nproc)}"
fi

# 2. the vendored tvm-ffi python extension (pip has 0.1.9; the reference needs >= 0.1.13)
if ! ls "$OUT/ffi_build"/core*.so >/dev/null 2>&1 && ! ls "$OUT/ffi_build"/*/core*.so >/dev/null 2>&1; then
  cmake -S "$REF/3rdparty/tvm-ffi" -B "$OUT/ffi_build" -G Ninja -DCMAKE_BUILD_TYPE=Release \
        -DTVM_FFI_BUILD_PYTHON_MODULE=ON -DPython_EXECUTABLE="$PY"
  ninja -C "$OUT/ffi_build" -j"${JOBS:-$(nproc)}"
fi

# 3. assemble a python package dir that shadows pip's apache-tvm-ffi 0.1.9
PKG="$OUT/ref_py"
rm -rf "$PKG"; mkdir -p "$PKG"
cp -r "$REF/3rdparty/tvm-ffi/python/tvm_ffi" "$PKG/tvm_ffi"
CORE=$(find "$OUT/ffi_build" -name 'core*.so' | head -1)
cp "$CORE" "$PKG/tvm_ffi/"
mkdir -p "$PKG/tvm_ffi/lib"
LIBDIR=$(dirname "$(find "$OUT/build" -name 'libtvm_runtime_extra.so' | head -1)")
cp "$(find "$OUT/build" -name 'libtvm_ffi.so' | head -1)" "$PKG/tvm_ffi/lib/"
mkdir -p "$PKG/tvm_ffi/include"
cp -r "$REF/3rdparty/tvm-ffi/include/." "$PKG/tvm_ffi/include/"
cp -r "$REF/3rdparty/tvm-ffi/3rdparty/dlpack/include/dlpack" "$PKG/tvm_ffi/include/" 2>/dev/null || \
  cp -r "$(find "$REF/3rdparty" -type d -path '*dlpack/include/dlpack' | head -1)" "$PKG/tvm_ffi/include/"
cat > "$PKG/tvm_ffi/_version.py" <<'PYEOF'
__version__ = version = "0.1.14"
__version_tuple__ = version_tuple = (0, 1, 14)
PYEOF
DI="$PKG/apache_tvm_ffi-0.1.14.dist-info"; mkdir -p "$DI"
printf 'Metadata-Version: 2.1\nName: apache-tvm-ffi\nVersion: 0.1.14\n' > "$DI/METADATA"
printf 'tvm_ffi/lib/libtvm_ffi.so,,\n' > "$DI/RECORD"
cat > "$OUT/env.sh" <<ENVEOF
export TVM_LIBRARY_PATH=$LIBDIR
export PYTHONPATH=$PKG:$REF/python
ENVEOF
echo "reference build ready: source $OUT/env.sh"
