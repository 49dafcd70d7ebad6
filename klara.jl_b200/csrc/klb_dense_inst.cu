// klb_dense_inst.cu -- instantiation + host dispatch of the dense-precision chain kernels
#include "klb_dense.cuh"

static size_t chain_smem(int d) {
  return ((KLB_TAB_LEN * 8 + 15) & ~15) + 3 * (size_t)klb_dense_dl(d) * KLB_DENSE_MC * 8 + sizeof(DenseShared<KLB_DENSE_MC>);
}
static size_t init_smem(int d) { return 2 * (size_t)klb_dense_dl(d) * KLB_DENSE_MC * 8; }

template <int S, bool F, bool DA = false>
static int go(const DArgs* D, int dim, int* regs, int* bps, cudaStream_t st) {
  auto kern = klb_dense_kernel<S, KLB_DENSE_MC, F, DA>;
  const size_t sm = chain_smem(D ? (int)D->k.dim : dim);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return -2;
  if (D) {
    const unsigned grid = (unsigned)((D->k.nchains + KLB_DENSE_MC - 1) / KLB_DENSE_MC);
    kern<<<grid, KLB_DENSE_THREADS, sm, st>>>(*D);
    return 0;
  }
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return -2;
  *regs = fa.numRegs;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, KLB_DENSE_THREADS, sm) != cudaSuccess) return -2;
  return 0;
}
static int dispatch(const DArgs* D, int sampler, int fma, int da, int dim, int* regs, int* bps, cudaStream_t st) {
  if (da) {                                     // DualAveragingMCTuner: HMC only (checked by klb_job_create)
    if (sampler != 2) return -1;
    return fma ? go<2, true, true>(D, dim, regs, bps, st) : go<2, false, true>(D, dim, regs, bps, st);
  }
  switch (sampler * 2 + (fma ? 1 : 0)) {
    case 0: return go<0, false>(D, dim, regs, bps, st);
    case 1: return go<0, true>(D, dim, regs, bps, st);
    case 2: return go<1, false>(D, dim, regs, bps, st);
    case 3: return go<1, true>(D, dim, regs, bps, st);
    case 4: return go<2, false>(D, dim, regs, bps, st);
    case 5: return go<2, true>(D, dim, regs, bps, st);
  }
  return -1;
}
int klb_dense_launch(const DArgs& D, int sampler, int fma, cudaStream_t s) {
  return dispatch(&D, sampler, fma, D.k.tuner == 2, 0, nullptr, nullptr, s);
}
int klb_dense_attrs(int sampler, int fma, int da, int dim, int* regs, int* bps) {
  return dispatch(nullptr, sampler, fma, da, dim, regs, bps, 0);
}
int klb_dense_init(const DArgs& D, int fma, int check_grad, unsigned long long* flag, cudaStream_t s) {
  const size_t sm = init_smem((int)D.k.dim);
  const unsigned grid = (unsigned)((D.k.nchains + KLB_DENSE_MC - 1) / KLB_DENSE_MC);
  if (fma) {
    auto kern = klb_dense_init_kernel<KLB_DENSE_MC, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return -2;
    kern<<<grid, KLB_DENSE_THREADS, sm, s>>>(D, check_grad, flag);
  } else {
    auto kern = klb_dense_init_kernel<KLB_DENSE_MC, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return -2;
    kern<<<grid, KLB_DENSE_THREADS, sm, s>>>(D, check_grad, flag);
  }
  return 0;
}
