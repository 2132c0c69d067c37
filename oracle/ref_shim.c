/* TEST / BENCH INFRASTRUCTURE ONLY.  The reference's `c`-target code calls its runtime's workspace allocator
 * (TVMBackendAllocWorkspace / TVMBackendFreeWorkspace, include/tvm/runtime/c_backend_api.h); outside a TVM process
 * (bench.py's cpu_baseline leg on the GPU box) this shim provides them with plain aligned malloc/free so that
 * oracle/_ref/*.so -- the reference's own CPU kernels -- load under pip tvm-ffi. */
#include <stdint.h>
#include <stdlib.h>

void* TVMBackendAllocWorkspace(int device_type, int device_id, uint64_t nbytes, int dtype_code_hint, int dtype_bits_hint) {
  (void)device_type; (void)device_id; (void)dtype_code_hint; (void)dtype_bits_hint;
  void* p = 0;
  if (posix_memalign(&p, 64, nbytes ? nbytes : 64) != 0) return 0;
  return p;
}
int TVMBackendFreeWorkspace(int device_type, int device_id, void* ptr) {
  (void)device_type; (void)device_id;
  free(ptr);
  return 0;
}
