// klb_hmc_ws_inst.cu -- instantiation + dispatch of the warp-specialised HMC kernels (NV = 8, 16; W = 1).
// Compiled once per arithmetic: -DKLB_INST_FMA={0,1}.
#include "klb_hmc_ws.cuh"

template <class T, int NV, bool FULL>
static int go(const KArgs* A, int* regs, int* bps, cudaStream_t s) {
  auto kern = klb_hmc_ws_kernel<T, NV, (KLB_INST_FMA != 0), FULL>;
  if (A) {
    const unsigned grid = (unsigned)((A->nchains + 3) / 4);
    kern<<<grid, 256, 0, s>>>(*A);
    return 0;
  }
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return -2;
  *regs = fa.numRegs;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, 256, 0) != cudaSuccess) return -2;
  return 0;
}
template <class T>
static int by_nv(const KArgs* A, int nv, int full, int* regs, int* bps, cudaStream_t s) {
  if (nv == 16) return full ? go<T, 16, true>(A, regs, bps, s) : go<T, 16, false>(A, regs, bps, s);
  if (nv == 8) return full ? go<T, 8, true>(A, regs, bps, s) : go<T, 8, false>(A, regs, bps, s);
  return -1;
}
#if KLB_INST_FMA
int klb_hmc_ws_1(const KArgs* A, int target, int nv, int full, int* regs, int* bps, cudaStream_t s) {
#else
int klb_hmc_ws_0(const KArgs* A, int target, int nv, int full, int* regs, int* bps, cudaStream_t s) {
#endif
  switch (target) {
    case 0: return by_nv<TgtIso>(A, nv, full, regs, bps, s);
    case 1: return by_nv<TgtShifted>(A, nv, full, regs, bps, s);
    case 3: return by_nv<TgtRosen>(A, nv, full, regs, bps, s);
  }
  return -1;
}
