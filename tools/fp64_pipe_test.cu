// Microbenchmark of the B200 fp64 vector pipe: dependent-issue latency and per-scheduler throughput of
// DADD / DMUL / DFMA as a function of the independent operations in flight per warp (ILP) and of the warps
// per scheduler.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/fp64_pipe_test tools/fp64_pipe_test.cu
// Output: cycles per warp-instruction per scheduler (the pipe's ideal is 2.0: 16 lanes per scheduler).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = a + i + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) v[i] = __dadd_rn(v[i], b);
      else if (OP == 1) v[i] = __dmul_rn(v[i], b);
      else v[i] = __fma_rn(v[i], b, a);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, int OP>
void run(int warps_per_smsp, double* out, long long* cyc) {
  const int iters = 4096;
  const int threads = 32 * 4 * warps_per_smsp;   // one CTA per SM, warps spread over the 4 schedulers
  k<ILP, OP><<<148, threads>>>(out, iters, 1.0, 1.0000001, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_inst = (double)c / ((double)iters * ILP * warps_per_smsp);
  printf("op %s  ILP %2d  warps/sched %d : %7.2f cycles per warp-instr per scheduler  (%.1f cycles per dependent step)\n",
         OP == 0 ? "DADD" : OP == 1 ? "DMUL" : "DFMA", ILP, warps_per_smsp, per_inst, (double)c / iters);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int w = 1; w <= 4; w *= 2) {
    run<1, 0>(w, out, cyc); run<2, 0>(w, out, cyc); run<4, 0>(w, out, cyc); run<6, 0>(w, out, cyc);
    run<8, 0>(w, out, cyc); run<12, 0>(w, out, cyc); run<16, 0>(w, out, cyc); run<24, 0>(w, out, cyc);
  }
  for (int w = 1; w <= 4; w *= 2) { run<1, 1>(w, out, cyc); run<8, 1>(w, out, cyc); run<16, 1>(w, out, cyc); run<24, 1>(w, out, cyc); }
  for (int w = 1; w <= 4; w *= 2) { run<1, 2>(w, out, cyc); run<8, 2>(w, out, cyc); run<16, 2>(w, out, cyc); run<24, 2>(w, out, cyc); }
  return 0;
}
