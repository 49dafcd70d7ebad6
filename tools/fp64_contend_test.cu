// How much do integer "producer" warps on the same scheduler slow a register-resident fp64 leapfrog body?
// 2 consumer warps per scheduler run the leapfrog body of tools/fp64_leapfrog_test.cu; P producer warps per scheduler
// run Philox4x32-10 rounds, either free-running (mode 0) or paced by named barriers exactly like the warp-specialised
// HMC kernel (mode 1: one producer per consumer, 16 Philox calls per lane per 10 leapfrog steps, FULL/EMPTY handshakes).
// Prints consumer cycles per fp64 warp-instruction per scheduler (ideal 2.0).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void philox(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
    c1 = (unsigned)p1; c3 = (unsigned)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int trans, int ncons_warps, int nprod_warps, double c, double eps, long long* cyc) {
  __shared__ volatile int done;
  extern __shared__ double stage_raw[];
  double (*stage)[32 * 32] = (double (*)[32 * 32])stage_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (warp < ncons_warps) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    double x[32], p[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { x[i] = 0.001 * (i + threadIdx.x); p[i] = 0.002 * (i + 1); }
    long long t0 = clock64();
    for (int t = 0; t < trans; ++t) {
      if (MODE == 1) {
        bar_sync(1 + warp, 64);  // FULL
#pragma unroll
        for (int i = 0; i < 32; ++i) p[i] = stage[warp][i * 32 + lane];
        bar_arrive(9 + warp, 64);  // EMPTY
      }
      for (int s = 0; s < 10; ++s) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          x[i] = __dadd_rn(__dmul_rn(eps, p[i]), x[i]);
          const double tt = __dmul_rn(c, x[i]);
          p[i] = __dadd_rn(p[i], tt);
          p[i] = __dadd_rn(p[i], tt);
        }
      }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += x[i] + p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (lane == 0) atomicAdd((int*)&done, 1);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    unsigned acc = 0;
    if (MODE == 0) {
      unsigned n = 0;
      while (warp - ncons_warps < nprod_warps && done < ncons_warps) {
        unsigned c0 = n++, c1 = threadIdx.x, c2 = blockIdx.x, c3 = 7;
        philox(c0, c1, c2, c3, 1234u, 5678u);
        acc ^= c0 ^ c1 ^ c2 ^ c3;
      }
    } else {
      const int cw = warp - ncons_warps;  // the consumer this producer feeds
      if (cw < ncons_warps && cw >= 0) {
        for (int t = 0; t < trans; ++t) {
          if (t > 0) bar_sync(9 + cw, 64);
          for (int u = 0; u < 16; ++u) {
            unsigned c0 = u, c1 = threadIdx.x, c2 = blockIdx.x, c3 = t;
            philox(c0, c1, c2, c3, 1234u, 5678u);
            stage[cw][(2 * u) * 32 + lane] = __longlong_as_double(((long long)c0 << 32 | c1) >> 12 | 0x3FF0000000000000ll) - 1.5;
            stage[cw][(2 * u + 1) * 32 + lane] = __longlong_as_double(((long long)c2 << 32 | c3) >> 12 | 0x3FF0000000000000ll) - 1.5;
          }
          bar_arrive(1 + cw, 64);
        }
      }
    }
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
  }
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int trans = 200;
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int P = 0; P <= 2; ++P) {
    k<0><<<148, 512>>>(out, trans, 8, 4 * P, -0.05, 0.05, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("free-running producers/sched %d : %.2f cycles per fp64 warp-instr per scheduler\n", P, (double)c / ((double)trans * 10 * 160 * 2));
  }
  {
    k<1><<<148, 512, 65536>>>(out, trans, 8, 8, -0.05, 0.05, cyc);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("paced 2C+2P per scheduler : %.2f cycles per fp64 warp-instr per scheduler (%.0f cycles per transition)\n",
           (double)c / ((double)trans * 10 * 160 * 2), (double)c / trans);
  }
  {
    k<1><<<148, 512, 65536>>>(out, trans, 4, 4, -0.05, 0.05, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("paced 1C+1P per scheduler : %.2f cycles per fp64 warp-instr per scheduler (%.0f cycles per transition)\n",
           (double)c / ((double)trans * 10 * 160 * 1), (double)c / trans);
  }
  return 0;
}
