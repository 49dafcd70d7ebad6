// klb_nuts_inst.cu -- instantiation + dispatch of the NUTS kernels (klb_nuts.cuh), the default geometries of every dim.
// Compiled once per arithmetic: -DKLB_INST_FMA={0,1}.
#include "klb_nuts.cuh"

#define KLB_NUTS_GEOMS(X) X(1, 1) X(1, 2) X(1, 4) X(1, 8) X(1, 16) X(4, 16)

template <class T, int W, int NV>
static int go(const KArgs* A, int* regs, int* bps, cudaStream_t s) {
  auto kern = klb_nuts_kernel<T, NV, W, (KLB_INST_FMA != 0)>;
  if (A) {
    const unsigned grid = (unsigned)((A->nchains + (KLB_WPB / W) - 1) / (KLB_WPB / W));
    kern<<<grid, 32 * KLB_WPB, 0, s>>>(*A);
    return 0;
  }
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return -2;
  *regs = fa.numRegs;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, 32 * KLB_WPB, 0) != cudaSuccess) return -2;
  return 0;
}
template <class T>
static int by_geo(const KArgs* A, int W, int NV, int* regs, int* bps, cudaStream_t s) {
#define X(w_, nv_) \
  if (W == w_ && NV == nv_) return go<T, w_, nv_>(A, regs, bps, s);
  KLB_NUTS_GEOMS(X)
#undef X
  return -1;
}
#if KLB_INST_FMA
int klb_nuts_1(const KArgs* A, int target, int W, int NV, int* regs, int* bps, cudaStream_t s) {
#else
int klb_nuts_0(const KArgs* A, int target, int W, int NV, int* regs, int* bps, cudaStream_t s) {
#endif
  switch (target) {
    case 0: return by_geo<TgtIso>(A, W, NV, regs, bps, s);
    case 1: return by_geo<TgtShifted>(A, W, NV, regs, bps, s);
    case 3: return by_geo<TgtRosen>(A, W, NV, regs, bps, s);
  }
  return -1;
}
