// Does the register-resident leapfrog body itself run at the fp64 pipe's rate?  32 elements of x and p per lane
// (NV = 16 units), the exact un-fused operation sequence of the kernels (t = c*x; p += t; p += t; u = eps*p; x += u),
// 1 / 2 / 3 warps per scheduler, nothing else in the kernel.  Prints cycles per fp64 warp-instruction per
// scheduler (ideal 2.0).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/fp64_leapfrog_test tools/fp64_leapfrog_test.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NE>
__global__ void __launch_bounds__(384, 1) k(double* out, int steps, double c, double eps, long long* cyc) {
  double x[NE], p[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) { x[i] = 0.001 * (i + threadIdx.x); p[i] = 0.002 * (i + 1); }
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      x[i] = __dadd_rn(__dmul_rn(eps, p[i]), x[i]);
      const double t = __dmul_rn(c, x[i]);
      p[i] = __dadd_rn(p[i], t);
      p[i] = __dadd_rn(p[i], t);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NE; ++i) s += x[i] + p[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int steps = 2000;
  for (int w = 1; w <= 3; ++w) {
    k<32><<<148, 128 * w>>>(out, steps, -0.05, 0.05, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("NE 32  warps/sched %d : %.2f cycles per fp64 warp-instr per scheduler (%.0f cycles per leapfrog step)\n", w,
           (double)c / ((double)steps * 32 * 5 * w), (double)c / steps);
  }
  for (int w = 1; w <= 3; ++w) {
    k<16><<<148, 128 * w>>>(out, steps, -0.05, 0.05, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("NE 16  warps/sched %d : %.2f cycles per fp64 warp-instr per scheduler (%.0f cycles per leapfrog step)\n", w,
           (double)c / ((double)steps * 16 * 5 * w), (double)c / steps);
  }
  return 0;
}
