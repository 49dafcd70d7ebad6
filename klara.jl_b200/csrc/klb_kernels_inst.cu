// klb_kernels_inst.cu -- explicit instantiation + host dispatch of the chain kernels.
// Compiled once per (sampler, arithmetic) pair: -DKLB_INST_SAMPLER={0,1,2} -DKLB_INST_FMA={0,1},
// and once with -DKLB_INST_INIT for the initialize! kernels.
//
// Geometries (W warps per chain, NV double2 units per thread; capacity 64 W NV elements):
//   (1,1) 64   (1,2) 128   (1,4) 256   (1,8) 512   (1,16) 1024    <- one warp per chain, defaults by dim
//   (4,16) 4096                                                     <- four warps per chain for dim > 1024
//   (2,8) 1024  (4,4) 1024                                          <- team alternatives kept for experiments
// Chain kernels exist in a masked version (any dim <= capacity) and a FULL version (dim == capacity).
#include "klb_kernels.cuh"

#define KLB_CAT2(a, b, c, d) a##b##c##d
#define KLB_CAT(a, b, c, d) KLB_CAT2(a, b, c, d)

#define KLB_GEOMS(X) X(1, 1) X(1, 2) X(1, 4) X(1, 8) X(1, 16) X(4, 16) X(2, 8) X(4, 4)

#if defined(KLB_INST_INIT)

template <class T, bool FMA>
static int init_geo(const KArgs& A, int W, int NV, int cg, unsigned long long* flag, cudaStream_t s) {
#define X(w_, nv_)                                                                          \
  if (W == w_ && NV == nv_) {                                                               \
    const unsigned grid = (unsigned)((A.nchains + (KLB_WPB / w_) - 1) / (KLB_WPB / w_));    \
    klb_init_kernel<T, nv_, w_, FMA><<<grid, 32 * KLB_WPB, 0, s>>>(A, cg, flag);            \
    return 0;                                                                               \
  }
  KLB_GEOMS(X)
#undef X
  return -1;
}
template <bool FMA>
static int init_t(const KArgs& A, int target, int W, int NV, int cg, unsigned long long* flag, cudaStream_t s) {
  switch (target) {
    case 0: return init_geo<TgtIso, FMA>(A, W, NV, cg, flag, s);
    case 1: return init_geo<TgtShifted, FMA>(A, W, NV, cg, flag, s);
    case 3: return init_geo<TgtRosen, FMA>(A, W, NV, cg, flag, s);
  }
  return -1;
}
int klb_launch_init(const KArgs& A, int target, int W, int NV, int fma, int check_grad, unsigned long long* flag,
                    cudaStream_t s) {
  return fma ? init_t<true>(A, target, W, NV, check_grad, flag, s)
             : init_t<false>(A, target, W, NV, check_grad, flag, s);
}

#else

#define KLB_FN(name) KLB_CAT(name, KLB_INST_SAMPLER, _, KLB_INST_FMA)

template <class T, int W, int NV, bool FULL>
static int go2(const KArgs* A, int* regs, int* bps, cudaStream_t s) {
  auto kern = klb_chain_kernel<KLB_INST_SAMPLER, T, NV, W, (KLB_INST_FMA != 0), FULL>;
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
template <class T, int W, int NV>
static int go(const KArgs* A, int full, int* regs, int* bps, cudaStream_t s) {
  return full ? go2<T, W, NV, true>(A, regs, bps, s) : go2<T, W, NV, false>(A, regs, bps, s);
}
template <class T>
static int by_geo(const KArgs* A, int W, int NV, int full, int* regs, int* bps, cudaStream_t s) {
#define X(w_, nv_) \
  if (W == w_ && NV == nv_) return go<T, w_, nv_>(A, full, regs, bps, s);
  KLB_GEOMS(X)
#undef X
  return -1;
}
// A != null: launch.  A == null: query registers / occupancy.
int KLB_FN(klb_chain_)(const KArgs* A, int target, int W, int NV, int full, int* regs, int* bps, cudaStream_t s) {
  switch (target) {
    case 0: return by_geo<TgtIso>(A, W, NV, full, regs, bps, s);
    case 1: return by_geo<TgtShifted>(A, W, NV, full, regs, bps, s);
    case 3: return by_geo<TgtRosen>(A, W, NV, full, regs, bps, s);
  }
  return -1;
}
#endif
