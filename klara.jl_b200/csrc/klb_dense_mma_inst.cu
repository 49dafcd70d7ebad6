// klb_dense_mma_inst.cu -- instantiation + dispatch of the DMMA dense HMC kernels (dim = 64 NT, NT = 1,2,4,8)
#include "klb_dense_mma.cuh"

static size_t mma_smem(int d) {
  return ((KLB_TAB_LEN * 8 + 15) & ~15) + (size_t)KLB_MMA_MC * (d + 4) * 8 + (size_t)KLB_MMA_STAGES * KLB_MMA_KB * (d + 4) * 8 +
         sizeof(DenseShared<KLB_MMA_MC>);
}
template <int NT, bool F>
static int go(const DArgs* D, int* regs, int* bps, cudaStream_t st) {
  auto kern = klb_dense_mma_kernel<NT, F>;
  const size_t sm = mma_smem(64 * NT);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return -2;
  if (D) {
    const unsigned grid = (unsigned)((D->k.nchains + KLB_MMA_MC - 1) / KLB_MMA_MC);
    kern<<<grid, KLB_DENSE_THREADS, sm, st>>>(*D);
    return 0;
  }
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return -2;
  *regs = fa.numRegs;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, KLB_DENSE_THREADS, sm) != cudaSuccess) return -2;
  return 0;
}
template <bool F>
static int by_dim(const DArgs* D, int dim, int* regs, int* bps, cudaStream_t st) {
  switch (dim) {
    case 64: return go<1, F>(D, regs, bps, st);
    case 128: return go<2, F>(D, regs, bps, st);
    case 256: return go<4, F>(D, regs, bps, st);
    case 512: return go<8, F>(D, regs, bps, st);
  }
  return -1;
}
int klb_dense_mma_launch(const DArgs& D, int fma, cudaStream_t s) {
  return fma ? by_dim<true>(&D, (int)D.k.dim, nullptr, nullptr, s) : by_dim<false>(&D, (int)D.k.dim, nullptr, nullptr, s);
}
int klb_dense_mma_attrs(int fma, int dim, int* regs, int* bps) {
  return fma ? by_dim<true>(nullptr, dim, regs, bps, 0) : by_dim<false>(nullptr, dim, regs, bps, 0);
}
