// klb_kernels_inst.cu -- explicit instantiation + host dispatch of the chain kernels.
// Compiled once per (sampler, arithmetic) pair: -DKLB_INST_SAMPLER={0,1,2} -DKLB_INST_FMA={0,1},
// and once with -DKLB_INST_INIT for the initialize! kernels and the table upload.
#include "klb_kernels.cuh"

#define KLB_CAT2(a, b, c, d) a##b##c##d
#define KLB_CAT(a, b, c, d) KLB_CAT2(a, b, c, d)

#if defined(KLB_INST_INIT)

template <class T, int NV, bool FMA>
static void launch_init(const KArgs& A, int check_grad, unsigned long long* flag, cudaStream_t s) {
  const unsigned grid = (unsigned)((A.nchains + KLB_WPB - 1) / KLB_WPB);
  klb_init_kernel<T, NV, FMA><<<grid, 32 * KLB_WPB, 0, s>>>(A, check_grad, flag);
}
template <class T, bool FMA>
static int init_nv(const KArgs& A, int nv, int cg, unsigned long long* flag, cudaStream_t s) {
  switch (nv) {
    case 1: launch_init<T, 1, FMA>(A, cg, flag, s); return 0;
    case 2: launch_init<T, 2, FMA>(A, cg, flag, s); return 0;
    case 4: launch_init<T, 4, FMA>(A, cg, flag, s); return 0;
    case 8: launch_init<T, 8, FMA>(A, cg, flag, s); return 0;
    case 16: launch_init<T, 16, FMA>(A, cg, flag, s); return 0;
  }
  return -1;
}
template <bool FMA>
static int init_t(const KArgs& A, int target, int nv, int cg, unsigned long long* flag, cudaStream_t s) {
  switch (target) {
    case 0: return init_nv<TgtIso, FMA>(A, nv, cg, flag, s);
    case 1: return init_nv<TgtShifted, FMA>(A, nv, cg, flag, s);
    case 3: return init_nv<TgtRosen, FMA>(A, nv, cg, flag, s);
  }
  return -1;
}
int klb_launch_init(const KArgs& A, int target, int nv, int fma, int check_grad, unsigned long long* flag,
                    cudaStream_t s) {
  return fma ? init_t<true>(A, target, nv, check_grad, flag, s) : init_t<false>(A, target, nv, check_grad, flag, s);
}

#else

#define KLB_FN(name) KLB_CAT(name, KLB_INST_SAMPLER, _, KLB_INST_FMA)

template <class T, int NV>
static int go(const KArgs* A, int* regs, int* bps, cudaStream_t s) {
  auto kern = klb_chain_kernel<KLB_INST_SAMPLER, T, NV, (KLB_INST_FMA != 0)>;
  if (A) {
    const unsigned grid = (unsigned)((A->nchains + KLB_WPB - 1) / KLB_WPB);
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
static int by_nv(const KArgs* A, int nv, int* regs, int* bps, cudaStream_t s) {
  switch (nv) {
    case 1: return go<T, 1>(A, regs, bps, s);
    case 2: return go<T, 2>(A, regs, bps, s);
    case 4: return go<T, 4>(A, regs, bps, s);
    case 8: return go<T, 8>(A, regs, bps, s);
    case 16: return go<T, 16>(A, regs, bps, s);
  }
  return -1;
}
// A != null: launch.  A == null: query registers / occupancy.
int KLB_FN(klb_chain_)(const KArgs* A, int target, int nv, int* regs, int* bps, cudaStream_t s) {
  switch (target) {
    case 0: return by_nv<TgtIso>(A, nv, regs, bps, s);
    case 1: return by_nv<TgtShifted>(A, nv, regs, bps, s);
    case 3: return by_nv<TgtRosen>(A, nv, regs, bps, s);
  }
  return -1;
}
#endif
