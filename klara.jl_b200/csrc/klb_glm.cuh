// klb_glm.cuh -- data-dependent, low-dimensional targets: Bayesian logistic regression.
//
// Reference closures (doc/examples/swiss/HMC/noadaptation/analytical.jl:11-20, the same three functions
// in doc/examples/swiss/MALA/analytical.jl), hyper-parameters v = [lambda, X, y, p] in model-vertex order:
//   ploglikelihood(p, v) = dot(Xp, y) - sum(log.(1+exp.(Xp))),   Xp = X*p
//   plogprior(p, v)      = -0.5*(dot(p, p)/lambda + length(p)*log(2*pi*lambda))
//   pgradlogtarget(p, v) = X'*(y - 1./(1+exp.(-X*p))) - p/lambda
//   logtarget            = loglikelihood + logprior            BasicContMuvParameter.jl:185-190
//
// Geometry.  dim is tiny (4 for the swiss bank-note data) while one target evaluation walks the whole
// data set (ndata x dim), so the warp-per-chain layout of klb_kernels.cuh would leave 28 of 32 lanes idle.
// Here ONE THREAD owns one chain: position, cached gradient, momentum / proposal live in registers
// (5 x DP doubles), the design matrix X (row-major, row pitch DP, zero padded) and y sit in shared memory
// and every thread of the CTA reads the same row at the same time (one broadcast LDS.128 per two
// coefficients), so the kernel is bound by the fp64 pipe: per data row 2 DP DFMA + one exp (+ one log for
// the log-likelihood) + one division.
//
// Order of the floating-point operations (the contract shared with the CPU oracle, `nv = 0`):
// X*p, X'*r, dot(.,.) are fma chains in increasing index (they are BLAS calls in the reference: fma kernels
// of unspecified order), sum() is a chain of additions in increasing index; everything elementwise is
// un-fused in arith = reference.  The cached gradient (pstate.gradlogtarget, iterate/HMC.jl:139) is not
// stored in HBM: it is a pure function of the position and is re-evaluated when a launch starts.
//
// Transitions: HMC iterate/HMC.jl:124-224 + samplers.jl:101-134, MALA iterate/MALA.jl:78-152,
// MH iterate/MH.jl:72-141 (symmetric branch); tuner block and save as in klb_kernels.cuh.
// NUTS (SAMPLER = 3; doc/examples/swiss/NUTS/{noadaptation,dualaveraging}/analytical.jl): the reference's multivariate
// transition as klb_nuts.cuh resolves it (one moving point, one running momentum, the pristine initial momentum for a
// direction that has not been used yet), evaluated by the chain's one thread: the moving point
// carries its gradient (the cached pstate.gradlogtarget of the reference's one shared state object), so a leaf costs ONE
// pass over the data (gradient and log-target of the new point share X*p).
#pragma once
#include "klb_kernels.cuh"

#define KLB_GLM_MAXD 16
#define KLB_GLM_THREADS 64
#ifndef KLB_NUTS_MAXLEVELS
#define KLB_NUTS_MAXLEVELS 10
#endif

// uniform number q of the transition: slot q of KLB_TAG_ACCEPT, consumed in the order the reference calls rand()
__device__ __forceinline__ double glm_seq_uniform(const klb_stream& st, unsigned& q) {
  uint64_t w0, w1;
  klb_stream_draw(&st, q++, KLB_TAG_ACCEPT, 0u, &w0, &w1);
  return klb_u01(w0);
}

struct GArgs {
  KArgs k;
  const double* X;     // ndata x DP row-major, zero padded (device)
  const double* y;     // ndata
  long long ndata;
  double lambda;       // prior variance
  double logc;         // log(2*pi*lambda), evaluated once on the host with the same klb_log
  int data_in_smem;    // X and y are staged in shared memory (they fit), else read through L1/L2
};

template <int DP, bool FMA>
struct Glm {
  // log-target and / or gradient at x; d = dim <= DP
  template <bool WANT_LT, bool WANT_G>
  static __device__ __forceinline__ void eval(const GArgs& G, const double* __restrict__ Xs,
                                              const double* __restrict__ ys, const uint64_t* tab, int d,
                                              const double (&x)[DP], double& lt, double (&g)[DP]) {
    double a = 0.0, sl = 0.0;
#pragma unroll
    for (int j = 0; j < DP; ++j) g[j] = 0.0;
    const int n = (int)G.ndata;
    // Two data rows per trip.  Everything a row needs before it touches the accumulators -- Xp_i, exp, log, the
    // division -- is independent of the other row, and with exp / log inlined ptxas interleaves the two dependent
    // chains (one row at a time the fp64 pipe sat at 29 %: stall_wait 2.5 per issue, profiles/r1_summary.md, column L).
    // The accumulators (a, sl, g[j]) still take row i before row i + 1: the oracle's order, bit for bit.
    int i = 0;
    for (; i + 1 < n; i += 2) {
      const double* row = Xs + (size_t)i * DP;
      double xr[2][DP];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < DP; j += 2) {
          const double2 v = *reinterpret_cast<const double2*>(row + r * DP + j);
          xr[r][j] = v.x; xr[r][j + 1] = v.y;
        }
      double xp[2] = {0.0, 0.0};
#pragma unroll
      for (int j = 0; j < DP; ++j)
        if (j < d) { xp[0] = __fma_rn(xr[0][j], x[j], xp[0]); xp[1] = __fma_rn(xr[1][j], x[j], xp[1]); }   // Xp = v[2]*p
      const double y0 = ys[i], y1 = ys[i + 1];
      if (WANT_LT) {
        const double l0 = klb_log_inline(__dadd_rn(1.0, klb_exp_inline(xp[0], tab)), tab);
        const double l1 = klb_log_inline(__dadd_rn(1.0, klb_exp_inline(xp[1], tab)), tab);
        a = __fma_rn(xp[0], y0, a); a = __fma_rn(xp[1], y1, a);             // dot(Xp, v[3])
        sl = __dadd_rn(sl, l0); sl = __dadd_rn(sl, l1);                     // sum(log.(1+exp.(Xp)))
      }
      if (WANT_G) {
        const double r0 = __dsub_rn(y0, __ddiv_rn(1.0, __dadd_rn(1.0, klb_exp_inline(-xp[0], tab))));
        const double r1 = __dsub_rn(y1, __ddiv_rn(1.0, __dadd_rn(1.0, klb_exp_inline(-xp[1], tab))));
#pragma unroll
        for (int j = 0; j < DP; ++j)
          if (j < d) { g[j] = __fma_rn(xr[0][j], r0, g[j]); g[j] = __fma_rn(xr[1][j], r1, g[j]); }   // v[2]'*r
      }
    }
    for (; i < n; ++i) {                                                    // odd row count: the last row alone
      const double* row = Xs + (size_t)i * DP;
      double xr[DP];
#pragma unroll
      for (int j = 0; j < DP; j += 2) {
        const double2 v = *reinterpret_cast<const double2*>(row + j);
        xr[j] = v.x; xr[j + 1] = v.y;
      }
      double xp = 0.0;
#pragma unroll
      for (int j = 0; j < DP; ++j)
        if (j < d) xp = __fma_rn(xr[j], x[j], xp);
      const double yi = ys[i];
      if (WANT_LT) {
        a = __fma_rn(xp, yi, a);
        sl = __dadd_rn(sl, klb_log(__dadd_rn(1.0, klb_exp(xp, tab)), tab));
      }
      if (WANT_G) {
        const double r = __dsub_rn(yi, __ddiv_rn(1.0, __dadd_rn(1.0, klb_exp(-xp, tab))));
#pragma unroll
        for (int j = 0; j < DP; ++j)
          if (j < d) g[j] = __fma_rn(xr[j], r, g[j]);
      }
    }
    if (WANT_LT) {
      double pp = 0.0;
#pragma unroll
      for (int j = 0; j < DP; ++j)
        if (j < d) pp = __fma_rn(x[j], x[j], pp);
      pp = __ddiv_rn(pp, G.lambda);
      const double dd = (double)d;
      const double inner = FMA ? __fma_rn(dd, G.logc, pp) : __dadd_rn(pp, __dmul_rn(dd, G.logc));
      lt = __dadd_rn(__dsub_rn(a, sl), __dmul_rn(-0.5, inner));
    }
    if (WANT_G) {
#pragma unroll
      for (int j = 0; j < DP; ++j) g[j] = (j < d) ? __dsub_rn(g[j], __ddiv_rn(x[j], G.lambda)) : 0.0;
    }
  }
};

// z <- randn(d): the same streams as every other kernel (pair k -> elements 2k, 2k+1)
template <int DP>
__device__ __forceinline__ void glm_randn(const klb_stream& st, int d, const uint64_t* tab, double (&z)[DP]) {
#pragma unroll
  for (int k = 0; k < DP / 2; ++k) {
    z[2 * k] = 0.0; z[2 * k + 1] = 0.0;
    if (2 * k < d) {
      uint64_t w0, w1;
      klb_stream_draw(&st, (uint32_t)k, KLB_TAG_NORMAL, 0u, &w0, &w1);
      double v;
      z[2 * k] = klb_zig_fast(w0, tab, &v) ? v : klb_normal_from_word(w0, 2u * k, &st, tab);
      if (2 * k + 1 < d) z[2 * k + 1] = klb_zig_fast(w1, tab, &v) ? v : klb_normal_from_word(w1, 2u * k + 1u, &st, tab);
    }
  }
}

template <int DP>
__device__ __forceinline__ void glm_load(double (&q)[DP], const double* __restrict__ col, int ld) {
#pragma unroll
  for (int j = 0; j < DP; j += 2) {
    double2 v = make_double2(0.0, 0.0);
    if (j < ld) v = *reinterpret_cast<const double2*>(col + j);
    q[j] = v.x; q[j + 1] = v.y;
  }
}
// the pad row (odd dim) is written as exactly 0
template <int DP>
__device__ __forceinline__ void glm_store(const double (&q)[DP], double* __restrict__ col, int d) {
#pragma unroll
  for (int j = 0; j < DP; j += 2)
    if (j < d) *reinterpret_cast<double2*>(col + j) = make_double2(q[j], (j + 1 < d) ? q[j + 1] : 0.0);
}

// stage X, y and the math tables; returns the pointers the evaluation reads
__device__ __forceinline__ void glm_stage(const GArgs& G, int DP, uint64_t* tab, double* dyn, const double*& Xs,
                                          const double*& ys) {
  for (int i = threadIdx.x; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = G.k.tab[i];
  if (G.data_in_smem) {
    const long long nx = G.ndata * DP;
    for (long long i = threadIdx.x; i < nx; i += blockDim.x) dyn[i] = G.X[i];
    for (long long i = threadIdx.x; i < G.ndata; i += blockDim.x) dyn[nx + i] = G.y[i];
    Xs = dyn; ys = dyn + nx;
  } else {
    Xs = G.X; ys = G.y;
  }
  __syncthreads();
}

// (Capping the registers at 128 for 8 CTAs = 16 warps per SM instead of 6 was measured: 339 ms against 328 ms on the
// HMC case of tools/glm_perf.py -- more warps do not help, the per-row exp / division chains are issue-bound.)
template <int SAMPLER, int DP, bool FMA>
__global__ void __launch_bounds__(KLB_GLM_THREADS) klb_glm_kernel(const GArgs G) {
  const KArgs& A = G.k;
  __shared__ uint64_t tab[KLB_TAB_LEN];
  extern __shared__ double2 glm_dyn2[];
  const double *Xs, *ys;
  glm_stage(G, DP, tab, reinterpret_cast<double*>(glm_dyn2), Xs, ys);

  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= A.nchains) return;
  const int d = (int)A.dim, ld = (int)A.ld;
  double* const xcol = A.state + c * A.ld;

  double x[DP], g[DP];
  glm_load<DP>(x, xcol, ld);
  double lt_cur = A.lt[c];
  if (SAMPLER != 0) {                        // the cached gradient of the current state
    double dummy;
    Glm<DP, FMA>::template eval<false, true>(G, Xs, ys, tab, d, x, dummy, g);
  }
  Tune tn;
  tn.step = A.tune_step[c];
  tn.accepted = A.tune_cnt[3 * c]; tn.proposed = A.tune_cnt[3 * c + 1]; tn.totproposed = A.tune_cnt[3 * c + 2];
  tn.rate = A.tune_rate[c];
  long long count = A.count0;
  long long thin = (A.i0 > A.burnin) ? klb_mod(A.i0 - A.burnin - 1, A.thinning) : 0;

  for (long long it = 0; it < A.nt; ++it) {
    const long long irun = A.i0 + it;
    const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)c,
                                          A.t0 + 1ull + (unsigned long long)it);
    bool accept = false;
    double lt_new = 0.0, ratio, a_prob = 1.0;
    int nl = A.nleaps;
    double xs[DP], gs[DP];
    double z[DP];
    glm_randn<DP>(st, d, tab, z);
    const double step = tn.step;
    const double h = __dmul_rn(0.5, step);
    int ndoub = 0;
    long long nuts_na = 1;
    double nuts_a = klb_u2d(0x7FF8000000000000ULL);
    if (SAMPLER == 3) {
      // ---------------------------------------------------------------- NUTS   iterate/NUTS.jl:230-457, NUTS.jl:514-628, :781-927
      double y[DP];                                                            // the running momentum (momentumprime)
      double k0 = 0.0;
#pragma unroll
      for (int j = 0; j < DP; ++j) { k0 = dotacc(z[j], z[j], k0); xs[j] = x[j]; gs[j] = g[j]; y[j] = z[j]; }
      const double oldh = __dsub_rn(lt_cur, __dmul_rn(0.5, k0));               // hamiltonian(job.pstate.logtarget, momentum)   :244
      unsigned q = 0u;
      const double u = __dadd_rn(klb_log(glm_seq_uniform(st, q), tab), oldh);  // log(rand()) + oldhamiltonian     :261
      bool used_plus = false, used_minus = false, y_is_initial = true, s = true;
      long long n = 1;
      double lt_e = lt_cur;
      while (s && ndoub < A.nuts_maxndoublings) {
        const bool fwd = glm_seq_uniform(st, q) < 0.5;                         // v = rand(Bool) ? 1 : -1            :264
        if (!(fwd ? used_plus : used_minus) && !y_is_initial) {
#pragma unroll
          for (int j = 0; j < DP; ++j) y[j] = z[j];
        }
        y_is_initial = false;
        const double step_v = fwd ? step : -step;                              // v*sstate.tune.step
        const double hv = __dmul_rn(0.5, step_v);
        // ---- build_tree!(..., ndoub): up to 2^ndoub leaves; the recursion is unwound after every leaf
        double saved_a[KLB_NUTS_MAXLEVELS + 1];
        long long saved_na[KLB_NUTS_MAXLEVELS + 1];
        long long nprime = 0;
        bool sprime = false;
        for (unsigned leaf = 0u;; ++leaf) {
#pragma unroll
          for (int j = 0; j < DP; ++j) {                                       // leapfrog!  samplers.jl:122-134    NUTS.jl:527
            y[j] = Ar<FMA>::ma(hv, gs[j], y[j]);
            xs[j] = Ar<FMA>::ma(step_v, y[j], xs[j]);
          }
          Glm<DP, FMA>::template eval<true, true>(G, Xs, ys, tab, d, xs, lt_e, gs);   // gradlogtarget!, logtarget!(pstateprime)
          double k1 = 0.0;
#pragma unroll
          for (int j = 0; j < DP; ++j) { y[j] = Ar<FMA>::ma(hv, gs[j], y[j]); k1 = dotacc(y[j], y[j], k1); }
          const double hprime = __dsub_rn(lt_e, __dmul_rn(0.5, k1));
          long long nn = (u <= hprime) ? 1 : 0;                                // :532
          const bool ss = u < __dadd_rn((double)A.nuts_maxdelta, hprime);      // :533
          double a = 0.0;
          if (A.tuner == 2) {                                                  // min(1, exp(H' - H0))               :818
            const double ex = klb_exp(__dsub_rn(hprime, oldh), tab);
            a = (ex != ex) ? ex : (ex < 1.0 ? ex : 1.0);
          }
          long long nna = 1;
          bool descend = false;
          for (int k = 1; k <= ndoub; ++k) {
            if ((leaf >> (k - 1)) & 1u) {             // a second half is complete: its rand(), n' doubles, sums add up
              (void)glm_seq_uniform(st, q);
              nn = 2 * nn;
              a = __dadd_rn(saved_a[k], a);
              nna = saved_na[k] + nna;
            } else if (ss) {                          // a first half that did not stop: on to its second half
              saved_a[k] = a; saved_na[k] = nna;
              descend = true;
              break;
            }                                         // a first half that stopped is returned as it is
          }
          if (!descend) { nprime = nn; sprime = ss; nuts_a = a; nuts_na = nna; break; }
        }
        if (fwd) used_plus = true; else used_minus = true;
        if (ndoub >= 1) { used_plus = true; used_minus = true; }               // NUTS.jl:541-549 rebinds both ends
        if (sprime) {
          const double r = glm_seq_uniform(st, q);
          if (r < __ddiv_rn((double)nprime, (double)n)) {                      // job.pstate <- pstateprime          :355-375
#pragma unroll
            for (int j = 0; j < DP; ++j) { x[j] = xs[j]; g[j] = gs[j]; }
            lt_cur = lt_e;
            accept = true;
          }
        }
        ndoub += 1;
        n += nprime;
        s = sprime;                                                            // && !uturn(E - E, ...) = true       :377-381
      }
    } else if (SAMPLER == 2) {
      // ---------------------------------------------------------------- HMC
      double k0 = 0.0;
#pragma unroll
      for (int j = 0; j < DP; ++j) { k0 = dotacc(z[j], z[j], k0); xs[j] = x[j]; gs[j] = g[j]; }
      const double oldh = __dsub_rn(lt_cur, __dmul_rn(0.5, k0));             // hamiltonian()
      if (A.tuner == 2) nl = da_nleaps(A, c, step);                          // DualAveragingMCTuner: per chain
      for (int s = 1; s <= nl; ++s) {                                        // leapfrog!  samplers.jl:122-134
#pragma unroll
        for (int j = 0; j < DP; ++j) {
          z[j] = Ar<FMA>::ma(h, gs[j], z[j]);
          xs[j] = Ar<FMA>::ma(step, z[j], xs[j]);
        }
        // the last gradient evaluation also yields logtarget!(proposal): same Xp, same bits
        if (s == nl) Glm<DP, FMA>::template eval<true, true>(G, Xs, ys, tab, d, xs, lt_new, gs);
        else { double dummy; Glm<DP, FMA>::template eval<false, true>(G, Xs, ys, tab, d, xs, dummy, gs); }
#pragma unroll
        for (int j = 0; j < DP; ++j) z[j] = Ar<FMA>::ma(h, gs[j], z[j]);
      }
      double k1 = 0.0;
#pragma unroll
      for (int j = 0; j < DP; ++j) k1 = dotacc(z[j], z[j], k1);
      const double newh = __dsub_rn(lt_new, __dmul_rn(0.5, k1));
      ratio = __dsub_rn(newh, oldh);
      if (ratio >= 0.0) accept = true;
      else {
        const double ex = klb_exp(ratio, tab);
        const double a = (ex != ex) ? ex : (ex < 1.0 ? ex : 1.0);
        a_prob = a;
        accept = klb_accept_uniform(&st) < a;
      }
    } else if (SAMPLER == 1) {
      // ---------------------------------------------------------------- MALA
      const double sq = __dsqrt_rn(step);
      const double hinv = __ddiv_rn(0.5, step);
      const DivBy by_step(step);
      double mu[DP];
      double e1 = 0.0, e2 = 0.0;
#pragma unroll
      for (int j = 0; j < DP; ++j) {
        mu[j] = Ar<FMA>::ma(h, g[j], x[j]);
        xs[j] = Ar<FMA>::ma(sq, z[j], mu[j]);
      }
      Glm<DP, FMA>::template eval<true, true>(G, Xs, ys, tab, d, xs, lt_new, gs);
#pragma unroll
      for (int j = 0; j < DP; ++j) {
        if (j < d) {
          const double df = __dsub_rn(mu[j], xs[j]);
          e1 = __dadd_rn(e1, FMA ? __dmul_rn(__dmul_rn(df, hinv), df) : __dmul_rn(0.5, by_step(__dmul_rn(df, df))));
          const double m2 = Ar<FMA>::ma(h, gs[j], xs[j]);
          const double db = __dsub_rn(m2, x[j]);
          e2 = __dadd_rn(e2, FMA ? __dmul_rn(__dmul_rn(db, hinv), db) : __dmul_rn(0.5, by_step(__dmul_rn(db, db))));
        }
      }
      ratio = __dsub_rn(lt_new, lt_cur);
      ratio = __dadd_rn(ratio, e1);
      ratio = __dsub_rn(ratio, e2);
      accept = (ratio > 0.0) || (ratio > klb_log(klb_accept_uniform(&st), tab));
    } else {
      // ---------------------------------------------------------------- MH
#pragma unroll
      for (int j = 0; j < DP; ++j) xs[j] = Ar<FMA>::ma(__ldg(A.sigma + j), z[j], x[j]);
      Glm<DP, FMA>::template eval<true, false>(G, Xs, ys, tab, d, xs, lt_new, gs);
      ratio = __dsub_rn(lt_new, lt_cur);
      accept = (ratio > 0.0) || (ratio > klb_log(klb_accept_uniform(&st), tab));
    }

    if (A.counters_on) { tn.proposed += 1; if (accept) tn.accepted += 1; }
    if (SAMPLER == 3) {                                        // iterate/NUTS.jl:402-447: tune! is fed a/na of the last doubling
      if (A.tuner == 2) da_block<false, 0, true>(A, c, tn, (int)nuts_na, __ddiv_rn(nuts_a, (double)nuts_na), tab, true);
      else tuner_block<2>(A, tn, tab, c);
    } else if (SAMPLER == 2 && A.tuner == 2) da_block<false>(A, c, tn, nl, a_prob, tab, true);
    else tuner_block<SAMPLER>(A, tn, tab, c);
    if (SAMPLER != 3 && accept) {                              // (NUTS moved job.pstate inside its loop)
#pragma unroll
      for (int j = 0; j < DP; ++j) { x[j] = xs[j]; if (SAMPLER != 0) g[j] = gs[j]; }
      lt_cur = lt_new;
    }

    if (irun > A.burnin) {                                     // in(i, postrange) -> save(job, count)
      if (thin == 0) {
        const long long col = c * A.npost + count;
        if (A.out_value) glm_store<DP>(x, A.out_value + col * A.ld, d);
        if (A.out_grad) glm_store<DP>(g, A.out_grad + col * A.ld, d);
        if (A.out_lt) A.out_lt[col] = lt_cur;
        if (A.out_accept) A.out_accept[col] = accept ? 1 : 0;
        if (SAMPLER == 3 && A.out_ndoublings) A.out_ndoublings[col] = (unsigned char)ndoub;
        if (SAMPLER == 3 && A.out_nuts_a) A.out_nuts_a[col] = nuts_a;          // :a, :na        iterate/NUTS.jl:393-399
        if (SAMPLER == 3 && A.out_nuts_na) A.out_nuts_na[col] = (int)nuts_na;
        count += 1;
      }
      thin = (thin + 1 == A.thinning) ? 0 : thin + 1;
    }
  }

  glm_store<DP>(x, xcol, d);
  A.lt[c] = lt_cur;
  A.tune_step[c] = tn.step;
  A.tune_cnt[3 * c] = tn.accepted; A.tune_cnt[3 * c + 1] = tn.proposed; A.tune_cnt[3 * c + 2] = tn.totproposed;
  A.tune_rate[c] = tn.rate;
}

// initialize!: lt[c] = logtarget(x_c), finiteness of the log-target (and gradient)   HMC.jl:106-120
template <int DP, bool FMA>
__global__ void __launch_bounds__(KLB_GLM_THREADS) klb_glm_init_kernel(const GArgs G, int check_grad,
                                                                       unsigned long long* flag) {
  const KArgs& A = G.k;
  __shared__ uint64_t tab[KLB_TAB_LEN];
  extern __shared__ double2 glm_dyn2[];
  const double *Xs, *ys;
  glm_stage(G, DP, tab, reinterpret_cast<double*>(glm_dyn2), Xs, ys);
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= A.nchains) return;
  const int d = (int)A.dim;
  double x[DP], g[DP], lt = 0.0;
  glm_load<DP>(x, A.state + c * A.ld, (int)A.ld);
  Glm<DP, FMA>::template eval<true, true>(G, Xs, ys, tab, d, x, lt, g);
  bool ok = isfinite(lt);
  if (check_grad) {
#pragma unroll
    for (int j = 0; j < DP; ++j)
      if (j < d) ok = ok && isfinite(g[j]);
  }
  A.lt[c] = lt;
  if (!ok) atomicMin(flag, (unsigned long long)(A.chain_offset + c + 1));
}

// host side (klb_glm_inst.cu)
int klb_glm_launch(const GArgs& G, int sampler, int fma, int dp, size_t dyn_smem, cudaStream_t s);
int klb_glm_init(const GArgs& G, int fma, int dp, size_t dyn_smem, int check_grad, unsigned long long* flag,
                 cudaStream_t s);
int klb_glm_attrs(int sampler, int fma, int dp, size_t dyn_smem, int* regs, int* bps);
