// Micro-benchmark (tuning aid, not part of the product): cycles per tcgen05.mma for the shapes / operand modes
// the prefill kernel can choose from.  One CTA per SM, one issuing thread, REP back-to-back MMAs, then one
// tcgen05.commit; clock64 around issue+completion.  Operands are zero-filled shared memory / TMEM (timing is
// data independent; power is not, so read cycles, not seconds).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tvm_b200/csrc -I include scripts/umma_bench.cu -o gpurun_out/umma_bench
#include "tc05.cuh"

#include <cstdio>

using namespace tvmb200;

enum Mode { SS = 0, TS = 1, MIX = 2 };

struct Cfg {
  int mode, n_qk, n_pv, reps, chain, with_ld;
};

__global__ void __launch_bounds__(192, 1) umma_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5;
  // layout: [0,64K) Q-like A operand (2 x 32 KiB), [64K,192K) four 32 KiB K/V slots, then barriers
  for (int i = tid; i < (192 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(sgen)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t bar0 = sbase + 192 * 1024, tptr = bar0 + 64;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tc05::tmem_alloc(tptr, 512);
    tc05::tmem_relinquish();
  }
  fence_proxy_async();
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + 192 * 1024 + 64);
  // zero TMEM so TS-mode A operands are finite
  if (warp < 4) {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0;
    for (int col = 0; col < 512; col += 32) tc05::st32(tmem + ((warp * 32u) << 16) + col, z);
    tc05::wait_st();
  }
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();

  const uint64_t dkm = tc05::make_smem_desc(0, 16, 1024);
  const uint64_t dmn = tc05::make_smem_desc(0, 16384, 1024);
  const uint32_t kmaj_hi = uint32_t(dkm >> 32), kmaj_lo = uint32_t(dkm);
  const uint32_t mn_hi = uint32_t(dmn >> 32), mn_lo = uint32_t(dmn);
  const uint32_t q_lo = kmaj_lo + (sbase >> 4), k_lo = kmaj_lo + ((sbase + 65536) >> 4), v_lo = mn_lo + ((sbase + 65536) >> 4);
  const uint32_t idesc_qk = tc05::make_idesc(1, 1, 0, 0, 128, c.n_qk);
  const uint32_t idesc_pv = tc05::make_idesc(1, 1, 0, 1, 128, c.n_pv);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    t0 = clock64();
    if (tc05::elect_one()) {
      for (int r = 0; r < c.reps; ++r) {
        const int t = r & 1;
        if (c.mode == SS || c.mode == MIX) {
          // S_t[buf] = Q_t K^T : 8 K-slices of 16, accumulate chain of `chain` (8 = like the kernel)
          const uint32_t d = tmem + t * 128 + ((r >> 1) & 1) * 64 * (c.n_qk <= 64);
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            const uint32_t off = ((s >> 2) * 16384 + (s & 3) * 32) >> 4;
            tc05::mma_ss_w(d, q_lo + ((t * 32768) >> 4) + off, kmaj_hi, k_lo + off, kmaj_hi, idesc_qk,
                           (s % c.chain) != 0);
          }
        }
        if (c.mode == TS || c.mode == MIX) {
          const uint32_t d = tmem + 256 + (c.n_pv <= 128 ? t * 128 : 0);
          const uint32_t p0 = tmem + t * 128;
          const int ks = c.mode == MIX ? c.n_qk / 16 : 8;
          for (int s = 0; s < ks; ++s)
            tc05::mma_ts_w(d, p0 + (s & 3) * 8, v_lo + ((32768 + (s & 7) * 16 * 128) >> 4), mn_hi, idesc_pv, 1u);
        }
      }
      tc05::commit(bar0);
    }
    __syncwarp();
    mbar_wait(bar0, 0);
    t1 = clock64();
  } else if (c.with_ld && warp >= 2) {
    // background TMEM reads like the softmax warps do (lanes of warp%4)
    uint32_t v[32];
    uint32_t acc = 0;
    const uint32_t la = ((warp & 3) * 32u) << 16;
    for (int it = 0; it < c.with_ld; ++it) {
      tc05::ld32(tmem + la + (it & 7) * 32, v);
      tc05::wait_ld();
      acc += v[it & 31];
    }
    if (acc == 0x12345678u) out[1000] = acc;
  }
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc05::fence_after_sync();
    tc05::tmem_dealloc(tmem, 512);
  }
  if (tid == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 2048 * sizeof(long long));
  const size_t smem = 192 * 1024 + 1024 + 256;
  cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Case { const char* name; Cfg c; int mma_per_rep; double mac_per_rep; };
  const int R = 512;
  Case cases[] = {
      {"SS N=64  chain8", {SS, 64, 128, R, 8, 0}, 8, 8.0 * 128 * 64 * 16},
      {"SS N=64  chain1", {SS, 64, 128, R, 1, 0}, 8, 8.0 * 128 * 64 * 16},
      {"SS N=128 chain8", {SS, 128, 128, R, 8, 0}, 8, 8.0 * 128 * 128 * 16},
      {"SS N=256 chain8", {SS, 256, 128, R, 8, 0}, 8, 8.0 * 128 * 256 * 16},
      {"SS N=32  chain8", {SS, 32, 128, R, 8, 0}, 8, 8.0 * 128 * 32 * 16},
      {"TS N=128 (PV)  ", {TS, 64, 128, R, 8, 0}, 8, 8.0 * 128 * 128 * 16},
      {"TS N=64        ", {TS, 64, 64, R, 8, 0}, 8, 8.0 * 128 * 64 * 16},
      {"TS N=256       ", {TS, 64, 256, R, 8, 0}, 8, 8.0 * 128 * 256 * 16},
      {"MIX QK64+PV4   ", {MIX, 64, 128, R, 8, 0}, 12, 8.0 * 128 * 64 * 16 + 4.0 * 128 * 128 * 16},
      {"MIX QK128+PV8  ", {MIX, 128, 128, R, 8, 0}, 16, 8.0 * 128 * 128 * 16 + 8.0 * 128 * 128 * 16},
      {"MIX QK64+PV4+ld", {MIX, 64, 128, R, 8, 4000}, 12, 8.0 * 128 * 64 * 16 + 4.0 * 128 * 128 * 16},
  };
  for (auto& cs : cases) {
    for (int grid : {1, 148}) {
      cudaMemset(d_out, 0, 2048 * sizeof(long long));
      umma_kernel<<<grid, 192, smem>>>(cs.c, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s: %s\n", cs.name, cudaGetErrorString(e));
        return 1;
      }
      long long h[148];
      cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%s grid %3d: %8.1f clk/rep  %6.1f clk/mma  %7.1f MAC/clk/SM (nominal 4096)\n", cs.name, grid, mx / R,
             mx / R / cs.mma_per_rep, cs.mac_per_rep * R / mx);
    }
  }
  return 0;
}
