// Same-warp interleave: K integer instructions (IMAD.WIDE / LOP3, Philox-like) per 4 fp64 instructions of the
// leapfrog body, 1 or 2 warps per scheduler.  If an fp64 warp-instruction blocked the issue port for one cycle
// only, K <= 4 would be free (cycles per group of 4 fp64 stay 8); if it holds the port for both of its pipe
// cycles the cost is 8 + K.
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void __launch_bounds__(256, 1) k(double* out, int steps, double c, double eps, long long* cyc) {
  double x[16], p[16];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = threadIdx.x * 8 + i + blockIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = 0.001 * (i + threadIdx.x); p[i] = 0.002 * (i + 1); }
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      x[i] = __dadd_rn(__dmul_rn(eps, p[i]), x[i]);
      p[i] = __dadd_rn(p[i], __dmul_rn(c, x[i]));
#pragma unroll
      for (int q = 0; q < K; q += 2) {   // 2 instructions per pass (IMAD.WIDE + LOP3), 8 independent chains
        const int r = (i * (K / 2) + q / 2) & 7;
        const unsigned long long m = (unsigned long long)0xD2511F53u * u[r];
        u[r] = (unsigned)(m >> 32) ^ (unsigned)m ^ 0x9E3779B9u;
      }
    }
  }
  long long t1 = clock64();
  double s = u[0] ^ u[1] ^ u[2] ^ u[3] ^ u[4] ^ u[5] ^ u[6] ^ u[7];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + p[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int K>
void run(double* out, long long* cyc) {
  const int steps = 4000;
  for (int w = 1; w <= 2; ++w) {
    k<K><<<148, 128 * w>>>(out, steps, -0.05, 0.05, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("K=%d int per 4 fp64, warps/sched %d : %.2f cycles per group of 4 fp64 per scheduler\n", K, w,
           (double)c / ((double)steps * 16 * w));
  }
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  run<0>(out, cyc); run<2>(out, cyc); run<4>(out, cyc); run<8>(out, cyc);
  return 0;
}
