// klb_nuts.cuh -- the reference's multivariate NUTS transition (src/samplers/iterate/NUTS.jl:230-457 with build_tree!,
// src/samplers/NUTS.jl:514-628 Vanilla, :781-927 DualAveraging) for the elementwise targets.
//
// What is restated is what the reference COMPUTES, not the textbook algorithm.  The reference builds its sampler state as
// MuvNUTSState(pstate, pstate, pstate, pstate, ...) (NUTS.jl:198-225): the plus end, the minus end, the proposal and the
// second-subtree proposal are one mutable object, every leaf returns sstate.pstateprime / sstate.momentumprime for all of
// them and every level of the recursion keeps n', s' in the one shared sstate.  The test suite holds a model of that code,
// statement by statement with reference semantics, and shows that it equals the state machine below (DESIGN.md section
// 6b; tests/test_oracle_nuts.py), which the CPU checker restates and this kernel follows bit for bit:
//   * one moving point (registers x) that every leaf advances in place by one leapfrog step of size v*step;
//   * one running momentum (registers y); a direction that has not been used yet starts its doubling from the pristine
//     initial momentum (kept in the warp's staging buffer for the whole transition);
//   * uturn(x - x, ...) never fires; an inner node draws its rand(), returns n' = 2 n'(second half), s' = s'(second half),
//     and adds up a', na' (first half + second half: pairwise, in tree order);
//   * job.pstate takes the moving point whenever s' && rand() < n'/n: the state column is written then, and re-read at the
//     end of the transition (the next transition starts from job.pstate, and so does the stored sample).
// Uniforms: slot q = 0, 1, 2, ... of KLB_TAG_ACCEPT of the transition's stream, in the order the reference calls rand().
// One warp per chain (or four, dim 1025..4096): every warp of a team evaluates the control flow for itself from the
// team's reduced sums, so the only synchronisation is the exchange of the lane accumulators (double-buffered).
#pragma once
#include "klb_kernels.cuh"

#define KLB_NUTS_MAXLEVELS 10   /* maxndoublings <= 10: at most 1023 leapfrog steps per transition */

__device__ __forceinline__ double klb_seq_uniform(const klb_stream& st, unsigned& q) {
  uint64_t w0, w1;
  klb_stream_draw(&st, q++, KLB_TAG_ACCEPT, 0u, &w0, &w1);
  return klb_u01(w0);
}

template <class T, int NV, int W, bool FMA>
__global__ void __launch_bounds__(32 * KLB_WPB)
klb_nuts_kernel(const KArgs A) {
  constexpr bool FULL = false;
  constexpr int CPB = KLB_WPB / W;
  constexpr int NLOC = 4 / W;
  constexpr bool S9 = !(NV == 16 && W == 4);
  __shared__ __align__(16) uint64_t tab[KLB_TAB_LEN + (S9 ? 512 : 0)];
  __shared__ double red[W == 1 ? 1 : CPB][2][W == 1 ? 1 : 2 * 4 * 32];   // exchange of the lane accumulators (teams only), alternating halves
  __shared__ double2 zstage[KLB_WPB][NV * 32];     // the transition's initial momentum (pristine)
  __shared__ unsigned short zqueue[KLB_WPB][KLB_QCAP];
  for (int i = threadIdx.x; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = A.tab[i];
  __syncthreads();
  if (S9) {
    zig_build9(tab);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, slot = warp / W, w = warp % W;
  const long long c = (long long)blockIdx.x * CPB + slot;
  if (c >= A.nchains) return;                      // whole teams leave together
  double2* const zbuf = zstage[warp];
  unsigned short* const queue = zqueue[warp];
  const int bar_id = 1 + slot;
  const bool writer = (w == 0) && (lane == 0);
  const int d = (int)A.dim;
  double* const xcol = A.state + c * A.ld;
  int rb = 0;                                       // which half of `red` the next exchange uses

  double x[2 * NV];
  load_chain<NV, W, FULL>(x, xcol, d, w, lane);
  double lt_cur = A.lt[c];
  Tune tn;
  tn.step = A.tune_step[c];
  tn.accepted = A.tune_cnt[3 * c]; tn.proposed = A.tune_cnt[3 * c + 1]; tn.totproposed = A.tune_cnt[3 * c + 2];
  tn.rate = A.tune_rate[c];
  const bool saving = (A.out_value != nullptr) || (A.out_lt != nullptr) || (A.out_grad != nullptr) ||
                      (A.out_accept != nullptr) || (A.out_ndoublings != nullptr) || (A.out_nuts_a != nullptr) ||
                      (A.out_nuts_na != nullptr);
  long long count = A.count0;
  long long thin = (A.i0 > A.burnin) ? klb_mod(A.i0 - A.burnin - 1, A.thinning) : 0;

  for (long long it = 0; it < A.nt; ++it) {
    const long long irun = A.i0 + it;
    const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)c,
                                          A.t0 + 1ull + (unsigned long long)it);
    double y[2 * NV];
    randn_stage<NV, W, FULL, S9>(st, d, w, lane, tab, zbuf, queue);            // momentum[:] = randn(size)        :242
    stage_load<NV>(y, zbuf, lane);
    double oldh;
    {
      double a1[1][NLOC] = {};
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int q = AccIdx<W>::local(j);
        a1[0][q] = dotacc(y[2 * j], y[2 * j], a1[0][q]);
        a1[0][q] = dotacc(y[2 * j + 1], y[2 * j + 1], a1[0][q]);
      }
      double k0[1];
      team_allsum<1, W>(a1, k0, red[W == 1 ? 0 : slot][rb], w, lane, bar_id);
      rb ^= 1;
      oldh = __dsub_rn(lt_cur, __dmul_rn(0.5, k0[0]));                         // hamiltonian(job.pstate.logtarget, momentum)   :244
    }
    unsigned q = 0u;
    const double u = __dadd_rn(klb_log(klb_seq_uniform(st, q), tab), oldh);     // log(rand()) + oldhamiltonian     :261
    const double step = tn.step;
    bool used_plus = false, used_minus = false, y_is_initial = true;
    bool s = true, update = false;
    int j = 0;
    long long n = 1, na = 1;
    double a_sum = klb_u2d(0x7FF8000000000000ULL);
    double lt_e = lt_cur;

    while (s && j < A.nuts_maxndoublings) {
      const bool fwd = klb_seq_uniform(st, q) < 0.5;                           // v = rand(Bool) ? 1 : -1            :264
      if (!(fwd ? used_plus : used_minus) && !y_is_initial) stage_load<NV>(y, zbuf, lane);
      y_is_initial = false;
      const double step_v = fwd ? step : -step;                                // v*sstate.tune.step
      const double h = __dmul_rn(0.5, step_v);
      // ---- build_tree!(..., j): up to 2^j leaves; the recursion is unwound after every leaf
      double saved_a[KLB_NUTS_MAXLEVELS + 1];
      long long saved_na[KLB_NUTS_MAXLEVELS + 1];
      long long nprime = 0;
      bool sprime = false;
      for (unsigned leaf = 0u;; ++leaf) {
        double acc[2][NLOC] = {};
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) {                                      // leapfrog!  samplers.jl:122-134    NUTS.jl:527
          const int i = Geo<NV, W>::elem(jj, w, lane);
          const int ql = AccIdx<W>::local(jj);
          const bool va = valid<FULL>(i, d), vb = valid<FULL>(i + 1, d);
          T::template kick<FMA, false>(A, i, va, vb, x[2 * jj], x[2 * jj + 1], h, y[2 * jj], y[2 * jj + 1]);
          x[2 * jj] = Ar<FMA>::ma(step_v, y[2 * jj], x[2 * jj]);
          x[2 * jj + 1] = Ar<FMA>::ma(step_v, y[2 * jj + 1], x[2 * jj + 1]);
          T::template kick<FMA, false>(A, i, va, vb, x[2 * jj], x[2 * jj + 1], h, y[2 * jj], y[2 * jj + 1]);
          acc[0][ql] = T::template lt_acc<FMA>(A, i, va, vb, x[2 * jj], x[2 * jj + 1], acc[0][ql]);   // logtarget!(pstateprime)
          acc[1][ql] = dotacc(y[2 * jj], y[2 * jj], acc[1][ql]);
          acc[1][ql] = dotacc(y[2 * jj + 1], y[2 * jj + 1], acc[1][ql]);
        }
        double sums[2];
        team_allsum<2, W>(acc, sums, red[W == 1 ? 0 : slot][rb], w, lane, bar_id);
        rb ^= 1;
        lt_e = T::lt_fin(A, sums[0]);
        const double hprime = __dsub_rn(lt_e, __dmul_rn(0.5, sums[1]));
        long long nn = (u <= hprime) ? 1 : 0;                                  // :532
        const bool ss = u < __dadd_rn((double)A.nuts_maxdelta, hprime);        // :533
        double a = 0.0;
        if (A.tuner == 2) {                                                    // min(1, exp(H' - H0))               :818
          const double ex = klb_exp(__dsub_rn(hprime, oldh), tab);
          a = (ex != ex) ? ex : (ex < 1.0 ? ex : 1.0);
        }
        long long nna = 1;
        bool descend = false;
        for (int k = 1; k <= j; ++k) {
          if ((leaf >> (k - 1)) & 1u) {             // a second half is complete: its rand(), n' doubles, sums add up
            (void)klb_seq_uniform(st, q);
            nn = 2 * nn;
            a = __dadd_rn(saved_a[k], a);
            nna = saved_na[k] + nna;
          } else if (ss) {                          // a first half that did not stop: on to its second half
            saved_a[k] = a; saved_na[k] = nna;
            descend = true;
            break;
          }                                         // a first half that stopped is returned as it is
        }
        if (!descend) { nprime = nn; sprime = ss; a_sum = a; na = nna; break; }
      }
      if (fwd) used_plus = true; else used_minus = true;
      if (j >= 1) { used_plus = true; used_minus = true; }                     // NUTS.jl:541-549 rebinds both ends
      if (sprime) {
        const double r = klb_seq_uniform(st, q);
        if (r < __ddiv_rn((double)nprime, (double)n)) {                        // job.pstate <- pstateprime          :355-375
          store_chain<NV, W, FULL>(x, xcol, d, w, lane);
          lt_cur = lt_e;
          update = true;
        }
      }
      j += 1;
      n += nprime;
      s = sprime;                                                              // && !uturn(E - E, ...) = true       :377-381
    }

    // counters and tuner                                                        iterate/NUTS.jl:238-240, 402-447
    if (A.counters_on) { tn.proposed += 1; if (update) tn.accepted += 1; }
    if (A.tuner == 2) da_block<true, (W == 1 ? 0 : 1), true>(A, c, tn, (int)na, __ddiv_rn(a_sum, (double)na), tab, writer);
    else tuner_block<2>(A, tn, tab, c);

    // the next transition starts from job.pstate, and so does the stored sample
    load_chain<NV, W, FULL>(x, xcol, d, w, lane);
    if (irun > A.burnin) {
      if (thin == 0) {
        if (saving) {
          const long long col = c * A.npost + count;
          if (A.out_value) store_chain<NV, W, FULL>(x, A.out_value + col * A.ld, d, w, lane);
          if (A.out_grad) {
            double g[2 * NV];
#pragma unroll
            for (int jj = 0; jj < NV; ++jj) {
              const int i = Geo<NV, W>::elem(jj, w, lane);
              T::template grad<FMA>(A, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * jj], x[2 * jj + 1], g[2 * jj], g[2 * jj + 1]);
              if (!valid<FULL>(i + 1, d)) g[2 * jj + 1] = 0.0;
              if (!valid<FULL>(i, d)) g[2 * jj] = 0.0;
            }
            store_chain<NV, W, FULL>(g, A.out_grad + col * A.ld, d, w, lane);
          }
          if (writer) {
            if (A.out_lt) A.out_lt[col] = lt_cur;
            if (A.out_accept) A.out_accept[col] = update ? 1 : 0;
            if (A.out_ndoublings) A.out_ndoublings[col] = (unsigned char)j;
            if (A.out_nuts_a) A.out_nuts_a[col] = a_sum;                        // :a, :na        iterate/NUTS.jl:393-399
            if (A.out_nuts_na) A.out_nuts_na[col] = (int)na;
          }
        }
        count += 1;
      }
      thin = (thin + 1 == A.thinning) ? 0 : thin + 1;
    }
  }

  if (writer) {
    A.lt[c] = lt_cur;
    A.tune_step[c] = tn.step;
    A.tune_cnt[3 * c] = tn.accepted; A.tune_cnt[3 * c + 1] = tn.proposed; A.tune_cnt[3 * c + 2] = tn.totproposed;
    A.tune_rate[c] = tn.rate;
  }
}
