// fp64_yield_run <cubin>: runs kernel `yk` of the cubin (driver API) with 0, 1 and 2 filler warps per scheduler.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s; cuGetErrorString(r_, &s); printf("%s: %s\n", #x, s); exit(1); } } while (0)
int main(int argc, char** argv) {
  CK(cuInit(0));
  CUdevice dev; CK(cuDeviceGet(&dev, 0));
  CUcontext ctx; CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
  CUmodule mod; CK(cuModuleLoad(&mod, argv[1]));
  CUfunction f; CK(cuModuleGetFunction(&f, mod, "yk"));
  CUdeviceptr out, cyc, nf;
  CK(cuMemAlloc(&out, 148 * 1024 * 8)); CK(cuMemAlloc(&cyc, 8)); CK(cuMemAlloc(&nf, 8));
  int trans = 100; double c = -0.05, eps = 0.05;
  double t0pi = 0;
  for (int nfill = 0; nfill <= 8; nfill += 4) {
    void* args[] = {&out, &trans, &nfill, &c, &eps, &cyc, &nf};
    CK(cuLaunchKernel(f, 148, 1, 1, 512, 1, 1, 0, 0, args, 0));
    CK(cuCtxSynchronize());
    long long cy; unsigned long long n;
    CK(cuMemcpyDtoH(&cy, cyc, 8)); CK(cuMemcpyDtoH(&n, nf, 8));
    const double nfp = (double)trans * 10 * 160 * 2;
    if (nfill == 0) { t0pi = cy / nfp; printf("%s: no filler %.3f cycles per fp64 warp-instr\n", argv[1], t0pi); continue; }
    const double nint = (double)n * 32 * (nfill / 4);
    printf("%s: %d filler warp(s)/sched: %.2f cycles per fp64 instr; filler IPC %.2f/sched; cost %.2f cycles per filler instr\n",
           argv[1], nfill / 4, cy / nfp, nint / cy, (cy - t0pi * nfp) / nint);
  }
  return 0;
}
