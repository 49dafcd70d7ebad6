// klb_hmc_ws_inst.cu -- instantiation + dispatch of the warp-specialised HMC kernels (W = 1: NV = 8, 16; W = 4: NV = 16).
// Compiled once per arithmetic: -DKLB_INST_FMA={0,1}.
#include "klb_hmc_ws.cuh"

template <class T, int NV, int W, bool FULL>
static int go(const KArgs* A, int* regs, int* bps, cudaStream_t s) {
  auto kern = klb_hmc_ws_kernel<T, NV, W, (KLB_INST_FMA != 0), FULL>;
  // W = 4: the four momentum stages of the chain go to dynamic shared memory (with the static part the CTA is past 48 KB)
  const size_t dyn = (W == 1) ? 0 : (size_t)4 * NV * 32 * sizeof(double2);
  if (dyn && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return -2;
  if (A) {
    const unsigned grid = (W == 1) ? (unsigned)((A->nchains + 3) / 4) : (unsigned)A->nchains;
    kern<<<grid, 256, dyn, s>>>(*A);
    return 0;
  }
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return -2;
  *regs = fa.numRegs;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, 256, dyn) != cudaSuccess) return -2;
  return 0;
}
template <class T>
static int by_nv(const KArgs* A, int gw, int nv, int full, int* regs, int* bps, cudaStream_t s) {
  if (gw == 4 && nv == 16) return full ? go<T, 16, 4, true>(A, regs, bps, s) : go<T, 16, 4, false>(A, regs, bps, s);
  if (gw != 1) return -1;
  if (nv == 16) return full ? go<T, 16, 1, true>(A, regs, bps, s) : go<T, 16, 1, false>(A, regs, bps, s);
  if (nv == 8) return full ? go<T, 8, 1, true>(A, regs, bps, s) : go<T, 8, 1, false>(A, regs, bps, s);
  return -1;
}
#if KLB_INST_FMA
int klb_hmc_ws_1(const KArgs* A, int target, int gw, int nv, int full, int* regs, int* bps, cudaStream_t s) {
#else
int klb_hmc_ws_0(const KArgs* A, int target, int gw, int nv, int full, int* regs, int* bps, cudaStream_t s) {
#endif
  switch (target) {
    case 0: return by_nv<TgtIso>(A, gw, nv, full, regs, bps, s);
    case 1: return by_nv<TgtShifted>(A, gw, nv, full, regs, bps, s);
    case 3: return by_nv<TgtRosen>(A, gw, nv, full, regs, bps, s);
  }
  return -1;
}
