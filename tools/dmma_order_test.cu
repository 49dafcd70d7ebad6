// Which accumulation order does the fp64 tensor-core instruction use?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -fmad=false -o dmma_order_test tools/dmma_order_test.cu
// D = A(8x4) B(4x8) + C with mma.sync.aligned.m8n8k4.row.col.f64; compares every output element with
//   H0: fma chain in increasing k      fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c))))
//   H1: fma chain in decreasing k
//   H2: exact sum rounded once (all products exact, single rounding)
//   H3: pairwise ((a0b0+a1b1)+(a2b2+a3b3))+c with fma
// and prints the number of mismatching elements per hypothesis over many random trials.  Used to decide
// whether the dense-target matrix-vector products can run on DMMA and stay bit-identical to the oracle.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__global__ void dmma(const double* A, const double* B, const double* C, double* D, int trials) {
  const int T = threadIdx.x;
  for (int t = 0; t < trials; ++t) {
    const double a = A[t * 32 + (T / 4) * 4 + (T % 4)];          // A[row][k]
    const double b = B[t * 32 + (T % 4) * 8 + (T / 4)];          // B[k][n]
    double c0 = C[t * 64 + (T / 4) * 8 + 2 * (T % 4)], c1 = C[t * 64 + (T / 4) * 8 + 2 * (T % 4) + 1];
    double d0, d1;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
    D[t * 64 + (T / 4) * 8 + 2 * (T % 4)] = d0;
    D[t * 64 + (T / 4) * 8 + 2 * (T % 4) + 1] = d1;
  }
}

int main() {
  const int trials = 20000;
  double *hA = new double[trials * 32], *hB = new double[trials * 32], *hC = new double[trials * 64], *hD = new double[trials * 64];
  srand(7);
  auto rnd = [] { return (rand() / (double)RAND_MAX - 0.5) * exp((rand() % 40 - 20) * 0.5); };
  for (int i = 0; i < trials * 32; ++i) { hA[i] = rnd(); hB[i] = rnd(); }
  for (int i = 0; i < trials * 64; ++i) hC[i] = rnd();
  double *A, *B, *C, *D;
  cudaMalloc(&A, trials * 32 * 8); cudaMalloc(&B, trials * 32 * 8); cudaMalloc(&C, trials * 64 * 8); cudaMalloc(&D, trials * 64 * 8);
  cudaMemcpy(A, hA, trials * 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, trials * 32 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(C, hC, trials * 64 * 8, cudaMemcpyHostToDevice);
  dmma<<<1, 32>>>(A, B, C, D, trials);
  if (cudaMemcpy(hD, D, trials * 64 * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
  long bad[4] = {0, 0, 0, 0}, total = 0;
  for (int t = 0; t < trials; ++t)
    for (int m = 0; m < 8; ++m)
      for (int n = 0; n < 8; ++n) {
        const double* a = hA + t * 32 + m * 4;
        double b[4]; for (int k = 0; k < 4; ++k) b[k] = hB[t * 32 + k * 8 + n];
        const double c = hC[t * 64 + m * 8 + n], d = hD[t * 64 + m * 8 + n];
        const double h0 = fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], fma(a[0], b[0], c))));
        const double h1 = fma(a[0], b[0], fma(a[1], b[1], fma(a[2], b[2], fma(a[3], b[3], c))));
        const long double ex = (long double)a[0] * b[0] + (long double)a[1] * b[1] + (long double)a[2] * b[2] + (long double)a[3] * b[3] + c;
        const double h2 = (double)ex;
        const double h3 = fma(a[0], b[0], a[1] * b[1]) + fma(a[2], b[2], a[3] * b[3]) + c;
        bad[0] += d != h0; bad[1] += d != h1; bad[2] += d != h2; bad[3] += d != h3; ++total;
      }
  printf("elements %ld  mismatches: H0(fma chain k up) %ld  H1(k down) %ld  H2(~exact, long double) %ld  H3(pairwise) %ld\n",
         total, bad[0], bad[1], bad[2], bad[3]);
  return 0;
}
