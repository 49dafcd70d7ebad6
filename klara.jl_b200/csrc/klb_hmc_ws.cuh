// klb_hmc_ws.cuh -- warp-specialised HMC chain kernel (one chain per consumer warp, dim 257..1024).
//
// The fused kernel of klb_kernels.cuh leaves the fp64 pipe ~58 % busy: every warp alternates between an
// fp64-dense leapfrog phase and integer / latency-bound phases (Philox + ziggurat, the slow-path pass, the
// accept test), and with 244 registers only two warps per scheduler are resident to cover for each other
// (profiles/r1_summary.md).  Here the two kinds of work live in different warps of the same CTA:
//   * one warpgroup of consumers (setmaxnreg.inc): one chain each -- positions and momenta in registers,
//     leapfrog, reductions, Metropolis test, tuner, stores: almost pure fp64 issue;
//   * one warpgroup of producers (setmaxnreg.dec): producer s generates the next transition's momentum
//     (Philox4x32-10 + ziggurat, slow path included) and accept uniform of consumer s into shared memory.
// Producer s and consumer s meet at two named barriers (FULL / EMPTY, 64 threads each).  Because warp w and
// warp 4+w share a scheduler (the producers are warps 0-3, the consumers warps 4-7), every scheduler holds
// 2 consumers + 2 producers (2 CTAs per SM): one consumer's tail overlaps the other's leapfrog, and the producers'
// integer instructions run in between (measured cost and ceiling: profiles/r1_summary.md, fp64 pipe microbenchmarks).
// The counter-based RNG makes this legal: a draw depends on (seed, chain, transition) only.
// Results are bit-identical to klb_chain_kernel (same per-element operations, same reduction order).
#pragma once
#include "klb_kernels.cuh"

#ifndef KLB_WS_CONSUMER_REGS
#define KLB_WS_CONSUMER_REGS 192
#endif
#ifndef KLB_WS_PRODUCER_REGS
#define KLB_WS_PRODUCER_REGS 64
#endif

__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <class T, int NV, bool FMA, bool FULL>
__global__ void __launch_bounds__(256, 2)
klb_hmc_ws_kernel(const KArgs A) {
  constexpr int W = 1;
  __shared__ uint64_t tab[KLB_TAB_LEN];
  __shared__ double2 zstage[4][NV * 32];
  __shared__ unsigned short zqueue[4][KLB_QCAP];
  __shared__ double uacc[4];
  for (int i = threadIdx.x; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = A.tab[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = warp & 3;
#ifndef KLB_WS_PRODUCER_LOW
#define KLB_WS_PRODUCER_LOW 1
#endif
  // which warp-id range the producers take makes no measurable difference (profiles/r1_summary.md)
  const bool producer = KLB_WS_PRODUCER_LOW ? (warp < 4) : (warp >= 4);
  const long long c = (long long)blockIdx.x * 4 + slot;
  const bool live = c < A.nchains;                 // a dead slot retires its consumer AND its producer
  const int d = (int)A.dim;
  double2* const zbuf = zstage[slot];
  const int bar_full = 1 + slot, bar_empty = 5 + slot;

  if (producer) {
    setmaxnreg_dec<KLB_WS_PRODUCER_REGS>();
    if (!live) return;
    for (long long it = 0; it < A.nt; ++it) {
      const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)c,
                                            A.t0 + 1ull + (unsigned long long)it);
      if (it > 0) bar_sync(bar_empty, 64);         // the consumer has taken the previous transition's draws
      randn_stage<NV, W, FULL>(st, d, 0, lane, tab, zbuf, zqueue[slot]);
      if (lane == 0) uacc[slot] = klb_accept_uniform(&st);
      bar_arrive(bar_full, 64);
    }
    return;
  }

  setmaxnreg_inc<KLB_WS_CONSUMER_REGS>();
  if (!live) return;
  double* const xcol = A.state + c * A.ld;
  double x[2 * NV];
  load_chain<NV, W, FULL>(x, xcol, d, 0, lane);
  double lt_cur = A.lt[c];
  Tune tn;
  tn.step = A.tune_step[c];
  tn.accepted = A.tune_cnt[3 * c]; tn.proposed = A.tune_cnt[3 * c + 1]; tn.totproposed = A.tune_cnt[3 * c + 2];
  tn.rate = A.tune_rate[c];
  const bool saving = (A.out_value != nullptr) || (A.out_lt != nullptr) || (A.out_grad != nullptr) ||
                      (A.out_accept != nullptr);
  long long count = A.count0;
  long long thin = (A.i0 > A.burnin) ? klb_mod(A.i0 - A.burnin - 1, A.thinning) : 0;

  for (long long it = 0; it < A.nt; ++it) {
    const long long irun = A.i0 + it;
    double y[2 * NV];
    bar_sync(bar_full, 64);                                                  // momentum[:] = randn(d) is ready
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double2 v = zbuf[j * 32 + lane];
      y[2 * j] = v.x; y[2 * j + 1] = v.y;
    }
    const double u_acc = uacc[slot];
    if (it + 1 < A.nt) bar_arrive(bar_empty, 64);                            // buffer may be refilled

    const double step = tn.step;
    const double h = __dmul_rn(0.5, step);
    double acc[3][4] = {};
#pragma unroll
    for (int j = 0; j < NV; ++j) {                                           // old kinetic energy
      acc[0][j & 3] = dotacc(y[2 * j], y[2 * j], acc[0][j & 3]);
      acc[0][j & 3] = dotacc(y[2 * j + 1], y[2 * j + 1], acc[0][j & 3]);
    }
    // leapfrog! (src/samplers/samplers.jl:122-134): see klb_chain_kernel for the exact-rewrite notes
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = Geo<NV, W>::elem(j, 0, lane);
      T::template kick<FMA, false>(A, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], h, y[2 * j], y[2 * j + 1]);
    }
    for (int s = 1; s < A.nleaps; ++s) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int i = Geo<NV, W>::elem(j, 0, lane);
        x[2 * j] = Ar<FMA>::ma(step, y[2 * j], x[2 * j]);
        x[2 * j + 1] = Ar<FMA>::ma(step, y[2 * j + 1], x[2 * j + 1]);
        T::template kick<FMA, true>(A, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], h, y[2 * j], y[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = Geo<NV, W>::elem(j, 0, lane);
      x[2 * j] = Ar<FMA>::ma(step, y[2 * j], x[2 * j]);
      x[2 * j + 1] = Ar<FMA>::ma(step, y[2 * j + 1], x[2 * j + 1]);
      T::template kick<FMA, false>(A, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], h, y[2 * j], y[2 * j + 1]);
      acc[1][j & 3] = T::template lt_acc<FMA>(A, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], acc[1][j & 3]);
      acc[2][j & 3] = dotacc(y[2 * j], y[2 * j], acc[2][j & 3]);
      acc[2][j & 3] = dotacc(y[2 * j + 1], y[2 * j + 1], acc[2][j & 3]);
    }
    double sums[3];
    team_allsum<3, 1>(acc, sums, nullptr, 0, lane, 0);
    const double lt_new = T::lt_fin(A, sums[1]);
    const double oldh = __dsub_rn(lt_cur, __dmul_rn(0.5, sums[0]));          // hamiltonian()
    const double newh = __dsub_rn(lt_new, __dmul_rn(0.5, sums[2]));
    const double ratio = __dsub_rn(newh, oldh);
    bool accept;
    if (ratio >= 0.0) accept = true;                                         // min(1., exp(ratio)) = 1 > rand()
    else {
      const double ex = klb_exp(ratio, tab);
      const double a = (ex != ex) ? ex : (ex < 1.0 ? ex : 1.0);
      accept = u_acc < a;
    }
    if (A.counters_on) { tn.proposed += 1; if (accept) tn.accepted += 1; }
    tuner_block<2>(A, tn, tab);
    if (accept) {
      store_chain<NV, W, FULL>(x, xcol, d, 0, lane);
      lt_cur = lt_new;
    } else {
      load_chain<NV, W, FULL>(x, xcol, d, 0, lane);
    }
    if (irun > A.burnin) {                                                   // in(i, postrange) -> save
      if (thin == 0) {
        if (saving) {
          const long long col = c * A.npost + count;
          if (A.out_value) store_chain<NV, W, FULL>(x, A.out_value + col * A.ld, d, 0, lane);
          if (A.out_grad) {
            double gbuf[2 * NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
              const int i = Geo<NV, W>::elem(j, 0, lane);
              T::template grad<FMA>(A, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], gbuf[2 * j], gbuf[2 * j + 1]);
              if (!valid<FULL>(i + 1, d)) gbuf[2 * j + 1] = 0.0;
              if (!valid<FULL>(i, d)) gbuf[2 * j] = 0.0;
            }
            store_chain<NV, W, FULL>(gbuf, A.out_grad + col * A.ld, d, 0, lane);
          }
          if (lane == 0) {
            if (A.out_lt) A.out_lt[col] = lt_cur;
            if (A.out_accept) A.out_accept[col] = accept ? 1 : 0;
          }
        }
        count += 1;
      }
      thin = (thin + 1 == A.thinning) ? 0 : thin + 1;
    }
  }
  if (lane == 0) {
    A.lt[c] = lt_cur;
    A.tune_step[c] = tn.step;
    A.tune_cnt[3 * c] = tn.accepted; A.tune_cnt[3 * c + 1] = tn.proposed; A.tune_cnt[3 * c + 2] = tn.totproposed;
    A.tune_rate[c] = tn.rate;
  }
}
