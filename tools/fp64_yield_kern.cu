// Kernels of the yield-bit experiment (tools/fp64_yield_run.cpp loads the cubin, patched or not, through the driver API).
// Per scheduler: 2 consumer warps running the un-fused leapfrog body + 1 or 2 filler warps running Philox rounds
// (IMAD.WIDE / LOP3 alternating: what ptxas emits for the producers of klb_hmc_ws_kernel).
#include <cuda_runtime.h>
extern "C" __global__ void __launch_bounds__(512, 1)
yk(double* out, int trans, int nfill, double c, double eps, long long* cyc, unsigned long long* nf) {
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
    double x[32], p[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { x[i] = 0.001 * (i + threadIdx.x); p[i] = 0.002 * (i + 1); }
    long long t0 = clock64();
    for (int t = 0; t < trans; ++t)
      for (int s = 0; s < 10; ++s) {
#pragma unroll
        for (int b = 0; b < 32; b += 4) {
          double tt[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) tt[q] = __dmul_rn(eps, p[b + q]);
#pragma unroll
          for (int q = 0; q < 4; ++q) x[b + q] = __dadd_rn(tt[q], x[b + q]);
#pragma unroll
          for (int q = 0; q < 4; ++q) tt[q] = __dmul_rn(c, x[b + q]);
#pragma unroll
          for (int q = 0; q < 4; ++q) p[b + q] = __dadd_rn(p[b + q], tt[q]);
#pragma unroll
          for (int q = 0; q < 4; ++q) p[b + q] = __dadd_rn(p[b + q], tt[q]);
        }
      }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += x[i] + p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    __syncwarp();
    if (lane == 0) atomicAdd((int*)&done, 1);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    unsigned a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 16 + i;
    unsigned long long n = 0;
    if (warp - 8 < nfill) {
      while (done < 8) {
#pragma unroll 1
        for (int r = 0; r < 16; ++r) {
          const unsigned b = 0x9E3779B9u + r, cc = 0xBB67AE85u;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            unsigned long long p0 = (unsigned long long)a[i] * 0xD2511F53ull, p1 = (unsigned long long)a[i + 1] * 0xCD9E8D57ull;
            a[i] = (unsigned)(p1 >> 32) ^ (unsigned)p0 ^ b;
            a[i + 1] = (unsigned)(p0 >> 32) ^ (unsigned)p1 ^ cc;
          }
        }
        n += 16;
      }
    }
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= a[i];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
    if (lane == 0 && blockIdx.x == 0 && warp == 8) *nf = n;
  }
}
