// klb_glm_inst.cu -- instantiation + host dispatch of the thread-per-chain kernels (klb_glm.cuh):
// padded dimension DP in {2, 4, 8, 16} x sampler x arithmetic.
#include "klb_glm.cuh"

template <int SAMPLER, int DP, bool FMA>
static int glm_go(const GArgs* G, size_t dyn, int* regs, int* bps, cudaStream_t s) {
  auto kern = klb_glm_kernel<SAMPLER, DP, FMA>;
  if (dyn > 0 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return -2;
  if (G) {
    const unsigned grid = (unsigned)((G->k.nchains + KLB_GLM_THREADS - 1) / KLB_GLM_THREADS);
    kern<<<grid, KLB_GLM_THREADS, dyn, s>>>(*G);
    return 0;
  }
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return -2;
  *regs = fa.numRegs;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, KLB_GLM_THREADS, dyn) != cudaSuccess) return -2;
  return 0;
}
template <int SAMPLER, bool FMA>
static int glm_dp(const GArgs* G, int dp, size_t dyn, int* regs, int* bps, cudaStream_t s) {
  switch (dp) {
    case 2: return glm_go<SAMPLER, 2, FMA>(G, dyn, regs, bps, s);
    case 4: return glm_go<SAMPLER, 4, FMA>(G, dyn, regs, bps, s);
    case 8: return glm_go<SAMPLER, 8, FMA>(G, dyn, regs, bps, s);
    case 16: return glm_go<SAMPLER, 16, FMA>(G, dyn, regs, bps, s);
  }
  return -1;
}
static int glm_any(const GArgs* G, int sampler, int fma, int dp, size_t dyn, int* regs, int* bps, cudaStream_t s) {
  switch (sampler * 2 + (fma ? 1 : 0)) {
    case 0: return glm_dp<0, false>(G, dp, dyn, regs, bps, s);
    case 1: return glm_dp<0, true>(G, dp, dyn, regs, bps, s);
    case 2: return glm_dp<1, false>(G, dp, dyn, regs, bps, s);
    case 3: return glm_dp<1, true>(G, dp, dyn, regs, bps, s);
    case 4: return glm_dp<2, false>(G, dp, dyn, regs, bps, s);
    case 5: return glm_dp<2, true>(G, dp, dyn, regs, bps, s);
    case 6: return glm_dp<3, false>(G, dp, dyn, regs, bps, s);   // NUTS
    case 7: return glm_dp<3, true>(G, dp, dyn, regs, bps, s);
  }
  return -1;
}
int klb_glm_launch(const GArgs& G, int sampler, int fma, int dp, size_t dyn, cudaStream_t s) {
  return glm_any(&G, sampler, fma, dp, dyn, nullptr, nullptr, s);
}
int klb_glm_attrs(int sampler, int fma, int dp, size_t dyn, int* regs, int* bps) {
  return glm_any(nullptr, sampler, fma, dp, dyn, regs, bps, 0);
}

template <int DP, bool FMA>
static int glm_init_go(const GArgs& G, size_t dyn, int cg, unsigned long long* flag, cudaStream_t s) {
  auto kern = klb_glm_init_kernel<DP, FMA>;
  if (dyn > 0 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return -2;
  const unsigned grid = (unsigned)((G.k.nchains + KLB_GLM_THREADS - 1) / KLB_GLM_THREADS);
  kern<<<grid, KLB_GLM_THREADS, dyn, s>>>(G, cg, flag);
  return 0;
}
template <bool FMA>
static int glm_init_dp(const GArgs& G, int dp, size_t dyn, int cg, unsigned long long* flag, cudaStream_t s) {
  switch (dp) {
    case 2: return glm_init_go<2, FMA>(G, dyn, cg, flag, s);
    case 4: return glm_init_go<4, FMA>(G, dyn, cg, flag, s);
    case 8: return glm_init_go<8, FMA>(G, dyn, cg, flag, s);
    case 16: return glm_init_go<16, FMA>(G, dyn, cg, flag, s);
  }
  return -1;
}
int klb_glm_init(const GArgs& G, int fma, int dp, size_t dyn, int check_grad, unsigned long long* flag, cudaStream_t s) {
  return fma ? glm_init_dp<true>(G, dp, dyn, check_grad, flag, s) : glm_init_dp<false>(G, dp, dyn, check_grad, flag, s);
}
