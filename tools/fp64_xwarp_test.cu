// Which kinds of instructions, issued by OTHER warps of the same scheduler, take fp64 issue time away from two
// register-resident leapfrog warps?  Per scheduler: 2 consumer warps (the un-fused leapfrog body, 160 fp64
// instructions per step) + P "filler" warps that run 8 independent dependency chains of one instruction kind until
// the consumers are done.  Prints, per kind, the consumers' cycles per fp64 warp-instruction (ideal 2.0) and the
// cost of one filler warp-instruction in cycles of consumer time:  (T - T0) / (filler instructions per scheduler).
#include <cstdio>
#include <cuda_runtime.h>

enum { K_LOP3, K_IADD, K_IMAD, K_IMADWIDE, K_IMADHI, K_SHF, K_FFMA, K_LDS, K_PHILOX, K_NKIND };
static const char* names[] = {"lop3", "iadd", "imad.lo", "imad.wide", "imad.hi", "shf", "ffma", "lds", "philox round (2 wide + 2 lop3)"};

template <int KIND>
__device__ __forceinline__ void filler(unsigned (&a)[8], unsigned b, unsigned c, const unsigned* sm) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
    if (KIND == K_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
    if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
    if (KIND == K_IMADWIDE) {
      unsigned long long w;
      asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w) : "r"(a[i]));
      asm volatile("mov.b64 {%0, _}, %1;" : "=r"(a[i]) : "l"(w));
    }
    if (KIND == K_IMADHI) asm volatile("mul.hi.u32 %0, %0, 0xD2511F53;" : "+r"(a[i]));
    if (KIND == K_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
    if (KIND == K_FFMA) { float f = __uint_as_float(a[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); a[i] = __float_as_uint(f); }
    if (KIND == K_LDS) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a[i]) : "r"((unsigned)__cvta_generic_to_shared(sm) + ((a[i] & 0xff) << 2)));
  }
  if (KIND == K_PHILOX) {   // 4 independent Philox rounds: 8 mul.wide + 8 lop3
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      unsigned long long p0, p1;
      asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p0) : "r"(a[i]));
      asm volatile("mul.wide.u32 %0, %1, 0xCD9E8D57;" : "=l"(p1) : "r"(a[i + 1]));
      a[i] = (unsigned)(p1 >> 32) ^ (unsigned)p0 ^ b;
      a[i + 1] = (unsigned)(p0 >> 32) ^ (unsigned)p1 ^ c;
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(512, 1) k(double* out, int trans, int nfill, double c, double eps, long long* cyc, unsigned long long* nf) {
  __shared__ volatile int done;
  __shared__ unsigned sm[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 256) sm[threadIdx.x] = threadIdx.x * 2654435761u;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    unsigned a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 8 + i;
    unsigned long long n = 0;
    if (warp - 8 < nfill) {
      while (done < 8) {
#pragma unroll 1
        for (int r = 0; r < 16; ++r) filler<KIND>(a, 0x9E3779B9u + r, 0xBB67AE85u, sm);
        n += 16;
      }
    }
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= a[i];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
    if (lane == 0 && blockIdx.x == 0 && warp == 8) *nf = n;
  }
}

template <int KIND>
static void run(double* out, long long* cyc, unsigned long long* nf, double t0pi) {
  const int trans = 100;
  for (int nfill = 4; nfill <= 8; nfill += 4) {
    k<KIND><<<148, 512>>>(out, trans, nfill, -0.05, 0.05, cyc, nf);
    cudaDeviceSynchronize();
    long long c; unsigned long long n;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&n, nf, 8, cudaMemcpyDeviceToHost);
    const double nfp = (double)trans * 10 * 160 * 2;                 // fp64 warp-instructions per scheduler
    const double per_iter = KIND == K_IMADWIDE ? 8 : (KIND == K_PHILOX ? 16 : 8);   // filler warp-instr per filler() call
    const double nint = (double)n * per_iter * (nfill / 4);
    printf("%-34s %d filler warp(s)/sched: %.2f cycles per fp64 instr; filler IPC %.2f/sched; cost %.2f cycles per filler instr\n",
           names[KIND], nfill / 4, c / nfp, nint / c, (c - t0pi * nfp) / nint);
  }
}

int main() {
  double* out; long long* cyc; unsigned long long* nf;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&nf, 8);
  k<K_LOP3><<<148, 512>>>(out, 100, 0, -0.05, 0.05, cyc, nf);
  cudaDeviceSynchronize();
  long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
  const double t0pi = c0 / (100.0 * 10 * 160 * 2);
  printf("no filler: %.3f cycles per fp64 warp-instr per scheduler\n", t0pi);
  run<K_LOP3>(out, cyc, nf, t0pi); run<K_IADD>(out, cyc, nf, t0pi); run<K_IMAD>(out, cyc, nf, t0pi);
  run<K_IMADWIDE>(out, cyc, nf, t0pi); run<K_IMADHI>(out, cyc, nf, t0pi); run<K_SHF>(out, cyc, nf, t0pi);
  run<K_FFMA>(out, cyc, nf, t0pi); run<K_LDS>(out, cyc, nf, t0pi); run<K_PHILOX>(out, cyc, nf, t0pi);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
