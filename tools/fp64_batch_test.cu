// Round 2 follow-up of fp64_xwarp_test.cu: does the ORDER of the producers' instructions matter?
// ptxas emits Philox4x32 as a strict alternation IMAD.WIDE / LOP3 (two different pipes: the stream can issue every
// cycle and each instruction costs ~1.5 cycles of fp64 time).  Here the same work is issued in same-pipe batches
// (B mul.wide back to back, then B lop3 back to back; volatile asm keeps the order): a homogeneous run is throttled
// by its own 16-lane pipe to one issue per two cycles and should fall into the shadow of the fp64 issue.
// Per scheduler: 2 consumer warps (un-fused leapfrog body) + 1 or 2 filler warps.  Prints the consumers' cycles per
// fp64 warp-instruction and the cost of one filler warp-instruction in cycles of consumer time.
#include <cstdio>
#include <cuda_runtime.h>

enum { K_ALT, K_B4, K_B8, K_B16, K_SPLIT, K_B8_NOVOL, K_NKIND };
static const char* names[] = {"philox alternating (wide, lop3, wide, lop3 ...)", "philox batches of 4", "philox batches of 8",
                              "philox batches of 16", "warp A only wide, warp B only lop3", "philox plain C (ptxas order)"};

#define WIDE(p, x, M) asm volatile("mul.wide.u32 %0, %1, " #M ";" : "=l"(p) : "r"(x))
#define MIX(o0, o1, p0, p1, k0, k1)                                                                      \
  asm volatile("{\n\t.reg .b32 l0, h0, l1, h1;\n\tmov.b64 {l0, h0}, %2;\n\tmov.b64 {l1, h1}, %3;\n\t"   \
               "lop3.b32 %0, h1, l0, %4, 0x96;\n\tlop3.b32 %1, h0, l1, %5, 0x96;\n\t}"                   \
               : "=r"(o0), "=r"(o1) : "l"(p0), "l"(p1), "r"(k0), "r"(k1))

template <int KIND>
__device__ __forceinline__ void filler(unsigned (&a)[16], unsigned b, unsigned c, int warp) {
  unsigned long long p[16];
  if (KIND == K_ALT) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      unsigned lo0, hi0, lo1, hi1;
      WIDE(p[i], a[i], 0xD2511F53);
      asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo0), "=r"(hi0) : "l"(p[i]));
      asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i + 1]) : "r"(hi0), "r"(a[i + 1]), "r"(c));
      WIDE(p[i + 1], a[i + 1], 0xCD9E8D57);
      asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo1), "=r"(hi1) : "l"(p[i + 1]));
      asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"(hi1), "r"(lo0), "r"(b));
    }
  }
  if (KIND == K_B4 || KIND == K_B8 || KIND == K_B16) {
    constexpr int B = KIND == K_B4 ? 4 : (KIND == K_B8 ? 8 : 16);
#pragma unroll
    for (int g = 0; g < 16; g += B) {
#pragma unroll
      for (int i = 0; i < B; i += 2) { WIDE(p[g + i], a[g + i], 0xD2511F53); WIDE(p[g + i + 1], a[g + i + 1], 0xCD9E8D57); }
#pragma unroll
      for (int i = 0; i < B; i += 2) MIX(a[g + i], a[g + i + 1], p[g + i], p[g + i + 1], b, c);
    }
  }
  if (KIND == K_SPLIT) {
    if (warp & 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) { WIDE(p[i], a[i], 0xD2511F53); asm volatile("mov.b64 {%0, _}, %1;" : "=r"(a[i]) : "l"(p[i])); }
    }
  }
  if (KIND == K_B8_NOVOL) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      unsigned long long p0 = (unsigned long long)a[i] * 0xD2511F53ull, p1 = (unsigned long long)a[i + 1] * 0xCD9E8D57ull;
      a[i] = (unsigned)(p1 >> 32) ^ (unsigned)p0 ^ b;
      a[i + 1] = (unsigned)(p0 >> 32) ^ (unsigned)p1 ^ c;
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(512, 1) k(double* out, int trans, int nfill, double c, double eps, long long* cyc, unsigned long long* nf) {
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
        for (int r = 0; r < 16; ++r) filler<KIND>(a, 0x9E3779B9u + r, 0xBB67AE85u, warp);
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

template <int KIND>
static void run(double* out, long long* cyc, unsigned long long* nf, double t0pi) {
  const int trans = 100;
  for (int nfill = 4; nfill <= 8; nfill += 4) {
    k<KIND><<<148, 512>>>(out, trans, nfill, -0.05, 0.05, cyc, nf);
    cudaDeviceSynchronize();
    long long c; unsigned long long n;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&n, nf, 8, cudaMemcpyDeviceToHost);
    const double nfp = (double)trans * 10 * 160 * 2;                 // fp64 warp-instructions per scheduler
    const double per_iter = KIND == K_SPLIT ? 16 : 32;               // filler warp-instr per filler() call
    const double nint = (double)n * per_iter * (nfill / 4);
    printf("%-50s %d filler warp(s)/sched: %.2f cycles per fp64 instr; filler IPC %.2f/sched; cost %.2f cycles per filler instr\n",
           names[KIND], nfill / 4, c / nfp, nint / c, (c - t0pi * nfp) / nint);
  }
}

int main() {
  double* out; long long* cyc; unsigned long long* nf;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&nf, 8);
  k<K_ALT><<<148, 512>>>(out, 100, 0, -0.05, 0.05, cyc, nf);
  cudaDeviceSynchronize();
  long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
  const double t0pi = c0 / (100.0 * 10 * 160 * 2);
  printf("no filler: %.3f cycles per fp64 warp-instr per scheduler\n", t0pi);
  run<K_ALT>(out, cyc, nf, t0pi); run<K_B4>(out, cyc, nf, t0pi); run<K_B8>(out, cyc, nf, t0pi);
  run<K_B16>(out, cyc, nf, t0pi); run<K_SPLIT>(out, cyc, nf, t0pi); run<K_B8_NOVOL>(out, cyc, nf, t0pi);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
