// Cost of one extra non-fp64 instruction of a given type issued between the fp64 instructions of the leapfrog body
// (same warp, 8 independent chains of the extra type, 2 warps per scheduler).  Prints cycles per leapfrog step for
// 64 fp64 instructions + N extra, and the marginal cycles per extra instruction.
#include <cstdio>
#include <cuda_runtime.h>

enum { T_NONE, T_LOP3, T_IADD, T_IMAD, T_IMADWIDE, T_FFMA, T_SHF };

template <int TYPE>
__device__ __forceinline__ void extra(unsigned (&u)[8], unsigned long long (&v)[8], float (&f)[8], int r) {
  if (TYPE == T_LOP3) asm volatile("lop3.b32 %0, %0, %1, 0x9E3779B9, 0x96;" : "+r"(u[r]) : "r"(u[(r + 1) & 7]));
  if (TYPE == T_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[r]) : "r"(u[(r + 1) & 7]));
  if (TYPE == T_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(u[r]) : "r"(u[(r + 1) & 7]));
  if (TYPE == T_IMADWIDE) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(v[r]) : "r"((unsigned)v[r]));
  if (TYPE == T_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f[r]) : "f"(f[(r + 1) & 7]));
  if (TYPE == T_SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(u[r]) : "r"(u[(r + 1) & 7]));
}

template <int TYPE, int PER>   // PER extra instructions per element (4 fp64 instructions)
__global__ void __launch_bounds__(256, 1) k(double* out, int steps, double c, double eps, long long* cyc) {
  double x[16], p[16];
  unsigned u[8]; unsigned long long v[8]; float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { u[i] = threadIdx.x * 8 + i + blockIdx.x; v[i] = u[i]; f[i] = 1.0f + 1e-3f * u[i]; }
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = 0.001 * (i + threadIdx.x); p[i] = 0.002 * (i + 1); }
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      x[i] = __dadd_rn(__dmul_rn(eps, p[i]), x[i]);
      if (PER >= 1) extra<TYPE>(u, v, f, (i * PER) & 7);
      if (PER >= 2) extra<TYPE>(u, v, f, (i * PER + 1) & 7);
      p[i] = __dadd_rn(p[i], __dmul_rn(c, x[i]));
      if (PER >= 3) extra<TYPE>(u, v, f, (i * PER + 2) & 7);
      if (PER >= 4) extra<TYPE>(u, v, f, (i * PER + 3) & 7);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (double)u[i] + (double)v[i] + (double)f[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + p[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

static double base = 0;
template <int TYPE, int PER>
void run(const char* name, double* out, long long* cyc) {
  const int steps = 4000;
  k<TYPE, PER><<<148, 256>>>(out, steps, -0.05, 0.05, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_step = (double)c / steps / 2;   // two warps share the scheduler
  if (PER == 0) base = per_step;
  printf("%-10s %d extra per 4 fp64: %.1f cycles per step per warp", name, PER, per_step);
  if (PER) printf("  -> %.2f cycles per extra instruction", (per_step - base) / (16.0 * PER));
  printf("\n");
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  run<T_NONE, 0>("none", out, cyc);
  run<T_LOP3, 2>("LOP3", out, cyc); run<T_LOP3, 4>("LOP3", out, cyc);
  run<T_IADD, 2>("IADD", out, cyc); run<T_IADD, 4>("IADD", out, cyc);
  run<T_IMAD, 2>("IMAD", out, cyc); run<T_IMAD, 4>("IMAD", out, cyc);
  run<T_IMADWIDE, 2>("IMAD.WIDE", out, cyc); run<T_IMADWIDE, 4>("IMAD.WIDE", out, cyc);
  run<T_FFMA, 2>("FFMA", out, cyc); run<T_FFMA, 4>("FFMA", out, cyc);
  run<T_SHF, 2>("SHF", out, cyc); run<T_SHF, 4>("SHF", out, cyc);
  return 0;
}
