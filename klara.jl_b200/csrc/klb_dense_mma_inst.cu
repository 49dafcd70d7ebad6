// klb_dense_mma_inst.cu -- instantiation + dispatch of the DMMA dense HMC kernels (dim = 64 NT, NT = 1,2,4,8)
#include "klb_dense_mma.cuh"

static size_t mma_smem(int d) {
  return ((KLB_TAB_LEN * 8 + 15) & ~15) + (size_t)KLB_MMA_MC * (d + 4) * 8 + (size_t)KLB_MMA_STAGES * KLB_MMA_KB * (d + 4) * 8 +
         sizeof(DenseShared<KLB_MMA_MC>) + 4 * sizeof(uint64_t);
}
template <int NT, bool F, int CL>
static int go(const DArgs* D, int* regs, int* bps, cudaStream_t st) {
  auto kern = klb_dense_mma_kernel<NT, F, CL>;
  const size_t sm = mma_smem(64 * NT);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return -2;
  if (D) {
    unsigned grid = (unsigned)((D->k.nchains + KLB_MMA_MC - 1) / KLB_MMA_MC);
    grid = (grid + CL - 1) / CL * CL;                       // whole clusters; surplus CTAs carry dead chains only
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(KLB_DENSE_THREADS); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, *D) == cudaSuccess ? 0 : -2;
  }
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return -2;
  *regs = fa.numRegs;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, KLB_DENSE_THREADS, sm) != cudaSuccess) return -2;
  return 0;
}
template <bool F, int CL>
static int by_dim(const DArgs* D, int dim, int* regs, int* bps, cudaStream_t st) {
  switch (dim) {
    case 64: return go<1, F, CL>(D, regs, bps, st);
    case 128: return go<2, F, CL>(D, regs, bps, st);
    case 256: return go<4, F, CL>(D, regs, bps, st);
    case 512: return go<8, F, CL>(D, regs, bps, st);
  }
  return -1;
}
int klb_dense_mma_launch(const DArgs& D, int fma, int cluster, cudaStream_t s) {
  const int dim = (int)D.k.dim;
  if (cluster == 4) return fma ? by_dim<true, 4>(&D, dim, nullptr, nullptr, s) : by_dim<false, 4>(&D, dim, nullptr, nullptr, s);
  if (cluster == 2) return fma ? by_dim<true, 2>(&D, dim, nullptr, nullptr, s) : by_dim<false, 2>(&D, dim, nullptr, nullptr, s);
  return fma ? by_dim<true, 1>(&D, dim, nullptr, nullptr, s) : by_dim<false, 1>(&D, dim, nullptr, nullptr, s);
}
int klb_dense_mma_attrs(int fma, int dim, int* regs, int* bps) {
  return fma ? by_dim<true, 1>(nullptr, dim, regs, bps, 0) : by_dim<false, 1>(nullptr, dim, regs, bps, 0);
}
