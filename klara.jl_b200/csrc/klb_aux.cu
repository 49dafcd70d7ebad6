// klb_aux.cu -- small support kernels: tuner-record reset and the device self-tests that the
// parity suite uses to compare the device math/RNG primitives with the oracle bit by bit.
#include "klb_kernels.cuh"

// tuner_state / reset!(tune, sampler, tuner):  (step, 0, 0, tuner.period, NaN)
//                                              src/samplers/samplers.jl:29-45, 79-90
__global__ void klb_fill_tune_kernel(double* step, long long* cnt, double* rate, long long n, double step0,
                                     long long period) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  step[c] = step0;
  cnt[3 * c] = 0; cnt[3 * c + 1] = 0; cnt[3 * c + 2] = period;
  rate[c] = klb_u2d(0x7FF8000000000000ULL);
}
void klb_launch_fill_tune(double* step, long long* cnt, double* rate, long long n, double step0, long long period,
                          cudaStream_t s) {
  klb_fill_tune_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(step, cnt, rate, n, step0, period);
}

__global__ void klb_debug_normals_kernel(const uint64_t* gtab, uint64_t seed, uint64_t chain, uint64_t t, long long n,
                                         double* out) {
  __shared__ uint64_t tab[KLB_TAB_LEN];
  for (int i = threadIdx.x; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  const klb_stream st = klb_stream_make(seed, chain, t);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = klb_normal(&st, (uint32_t)i, tab);
}
void klb_launch_debug_normals(const uint64_t* tab, uint64_t seed, uint64_t chain, uint64_t t, long long n, double* out,
                              cudaStream_t s) {
  klb_debug_normals_kernel<<<148, 256, 0, s>>>(tab, seed, chain, t, n, out);
}

__global__ void klb_debug_math_kernel(const uint64_t* gtab, int op, long long n, const double* in, double* out) {
  __shared__ uint64_t tab[KLB_TAB_LEN];
  for (int i = threadIdx.x; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = op == 0 ? klb_exp(in[i], tab) : klb_log(in[i], tab);
}
void klb_launch_debug_math(const uint64_t* tab, int op, long long n, const double* in, double* out, cudaStream_t s) {
  klb_debug_math_kernel<<<148, 256, 0, s>>>(tab, op, n, in, out);
}

__global__ void klb_debug_uniform_kernel(uint64_t seed, uint64_t chain, uint64_t t, double* out) {
  const klb_stream st = klb_stream_make(seed, chain, t);
  out[0] = klb_accept_uniform(&st);
}
void klb_launch_debug_uniform(uint64_t seed, uint64_t chain, uint64_t t, double* out, cudaStream_t s) {
  klb_debug_uniform_kernel<<<1, 1, 0, s>>>(seed, chain, t, out);
}
