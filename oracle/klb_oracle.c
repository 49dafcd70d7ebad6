/* klb_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of Klara.jl's serial MCMC hot path, one C function per
 * reference function, written to be read side by side with the Julia source
 * (paths relative to the reference checkout, commit ffa4f6d0):
 *
 *   orc_run_chain      <- run(job::BasicMCJob)                 src/jobs/BasicMCJob.jl:212-244
 *   orc_save           <- save / copy!(nstate, state, i)       src/jobs/BasicMCJob.jl:210,
 *                         src/nstates/ParameterNStates/BasicContMuvParameterNState.jl:89-119
 *   orc_hamiltonian    <- hamiltonian(logtarget, momentum)     src/samplers/samplers.jl:103
 *   orc_leapfrog       <- leapfrog!(... Multivariate ...)      src/samplers/samplers.jl:122-134
 *   orc_iterate_hmc    <- iterate!(job, HMC, Multivariate)     src/samplers/iterate/HMC.jl:124-224
 *   orc_iterate_mala   <- iterate!(job, MALA, Multivariate)    src/samplers/iterate/MALA.jl:78-152
 *   orc_iterate_mh     <- iterate!(job, MH, Multivariate)      src/samplers/iterate/MH.jl:72-141 (symmetric)
 *   orc_rate / orc_reset_burnin                                src/tuners/tuners.jl:27-32
 *   orc_tune           <- tune!(tune, ::AcceptanceRateMCTuner) src/tuners/AcceptanceRateMCTuner.jl:46
 *   orc_logistic       <- logistic(x,l,k,x0,y0)                src/stats/logistic.jl:11
 *   orc_logistic_rate_score / orc_erf_rate_score               src/tuners/AcceptanceRateMCTuner.jl:9,17
 *   orc_tuner_state    <- tuner_state(...)                     src/samplers/samplers.jl:29-45
 *   orc_da_state / orc_da_tune / orc_da_nleaps / orc_da_block  <- DualAveragingMCTuner for HMC:
 *                         src/tuners/DualAveragingMCTuner.jl:95-101, src/samplers/HMC.jl:124-133,192-223,
 *                         src/samplers/iterate/HMC.jl:125-127,142-144,225-248, src/samplers/samplers.jl:170-202
 *   orc_upto / orc_logtarget / orc_gradlogtarget               src/variables/parameters/BasicContMuvParameter.jl:174-201,264-279
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (klara.jl_b200/) never does.
 *
 * PARITY STATUS.  The reference cannot be executed (Julia 0.6, no julia binary) and it
 * never seeds its RNG, so there is no golden chain to compare with: sampler-result
 * parity is "parity unpinned" by the reference itself.  What IS pinned, bit-exactly,
 * against the reference's own known-answer tests (tests/test_oracle_kat.py):
 *   logistic(0.7,3,4,2.1,1.4)             test/common.jl:6
 *   logistic_rate_score(0.25), (0.5,11)   test/AcceptanceRateMCTuner.jl:8-9
 *   erf_rate_score(-0.1), (0.93,2)        test/AcceptanceRateMCTuner.jl:13-14
 *   function-defined normal target values test/BasicContMuvParameter.jl:539-563
 *   NState column layout                  test/ParameterNStates.jl:137-146
 * and the random-number primitives against their published vectors (the Philox4x32 round
 * function and key schedule through the 10-round Random123 KAT; the contract runs 7 rounds) and against libm / theory.
 *
 * Deliberate replacement (documented in DESIGN.md): Julia's global MersenneTwister
 * randn()/rand() are replaced by counter-based Philox streams keyed by
 * (seed, chain, transition, element); the arithmetic of every other line follows the
 * reference's un-fused evaluation order when cfg->arith == 0.
 *
 * Reductions (dot, sum) use the canonical order documented in DESIGN.md ("reduction
 * order"), which is also what the device kernels execute; Julia's own dot/sum order is
 * BLAS/SIMD dependent and not specified.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../klara.jl_b200/csrc/klb_math.h"

enum { ORC_MH = 0, ORC_MALA = 1, ORC_HMC = 2, ORC_NUTS = 3 };
enum { ORC_ISO = 0, ORC_SHIFTED = 1, ORC_DENSE = 2, ORC_ROSEN = 3, ORC_LOGIT = 4 };
enum { ORC_VANILLA = 0, ORC_ACCRATE = 1, ORC_DUALAVG = 2 };

typedef struct {
  int32_t sampler, target, tuner, arith;      /* arith 0 = reference order (un-fused), 1 = fma-contracted */
  int64_t nchains, dim, nsteps, burnin, thinning;
  double step;                                /* leapstep / driftstep */
  int32_t nleaps;
  double target_rate, score_k;
  int64_t period;
  int32_t verbose;
  uint32_t monitor;                           /* bit0 value, bit1 logtarget, bit2 gradlogtarget */
  uint32_t diagnostics;                       /* bit0 accept */
  uint64_t seed, chain_offset, t0;
  int32_t score;                              /* AcceptanceRateMCTuner score: 0 logistic_rate_score, 1 erf_rate_score */
  int32_t nv;                                 /* reduction geometry: double2 units per lane (1,2,4,8,16,64);
                                                 0 = sequential order (one thread per chain: the data-dependent
                                                 targets, klb_glm.cuh) */
  int32_t nthreads;                           /* OpenMP threads over chains; 1 = map(run, jobs) semantics */
  /* DualAveragingMCTuner(targetrate, nadapt; e0bar, h0bar, gamma, t0, kappa, period, verbose)
   * src/tuners/DualAveragingMCTuner.jl:53-95 (targetrate = target_rate above) */
  int64_t da_nadapt, da_t0;
  double da_eps0bar, da_h0bar, da_gamma, da_kappa;
  double* da;                                 /* per chain, in/out: the DualAveragingMCTune fields below (8 doubles) */
  /* NUTS(leapstep; maxδ, maxndoublings)   src/samplers/NUTS.jl:228-241; step above = leapstep */
  int32_t nuts_maxdelta, nuts_maxndoublings;
  uint8_t* nuts_ndoublings;                   /* out, npost x nchains: the :ndoublings diagnostic (diagnostics bit1); may be NULL */
  double* nuts_a;                             /* out, npost x nchains: :a  (diagnostics bit2; NUTS + DualAveragingMCTuner, NUTS.jl:317) */
  int32_t* nuts_na;                           /* out, npost x nchains: :na (diagnostics bit3) */
} orc_config;

/* DualAveragingMCTune minus the BasicMCTune part (src/tuners/DualAveragingMCTuner.jl:1-13); sstate.count rides along */
typedef struct { double lambda, mu, epsbar, hbar, hweight, epsweight, nleaps, count; } orc_da;

/* per-chain BasicMCTune (src/tuners/tuners.jl:5-25) */
typedef struct { double step; int64_t accepted, proposed, totproposed; double rate; } orc_tune;

/* one chain's BasicContMuvParameterState (src/states/ParameterStates/BasicContMuvParameterState.jl:62-97) */
typedef struct { double* value; double logtarget; double* gradlogtarget; int accept; int ndoublings; double nuts_a; int64_t nuts_na; } orc_pstate;

typedef struct {
  const orc_config* cfg;
  int64_t d, dp;              /* dim, padded dim = 64*nv */
  const double* mu;           /* shifted-iso mean (padded with 0) */
  const double* sigma;        /* MH proposal std-devs (padded with 0) */
  const double* C;            /* dense precision, d x d row-major (symmetric) */
  double ra, rb, rscale;      /* rosenbrock */
  /* Bayesian logistic regression (doc/examples/swiss): v = [lambda, X, y, p] */
  int64_t ndata;
  const double* X;            /* ndata x d, row-major */
  const double* y;            /* ndata */
  double lambda;              /* prior variance v[1] */
} orc_model;

/* ------------------------------------------------------------------ scalars */
double orc_logistic(double x, double l, double k, double x0, double y0) {
  /* l/(1+exp(-k*(x-x0)))+y0                                  src/stats/logistic.jl:11 */
  return l / (1 + klb_exp(-k * (x - x0), KLB_TAB)) + y0;
}
double orc_logistic_rate_score(double x, double k) { return orc_logistic(x, 2., k, 0., 0.); }
/* erf(k*x)+1                                               src/tuners/AcceptanceRateMCTuner.jl:17
 * klb_erf: double-double series shared with the device (identical bits); pinned to the reference's known-answer
 * values and to mpmath in tests/test_oracle_kat.py */
double orc_erf_rate_score(double x, double k) { return klb_erf(k * x, KLB_TAB) + 1; }
double orc_erf(double x) { return klb_erf(x, KLB_TAB); }

double orc_exp(double x) { return klb_exp(x, KLB_TAB); }
double orc_log(double x) { return klb_log(x, KLB_TAB); }
/* rounds = 10: the published generator (Random123 known-answer vectors); rounds = 7: the contract's (klb_math.h) */
void orc_philox(const uint32_t c[4], const uint32_t k[2], int rounds, uint32_t out[4]) {
  klb_philox4x32_r(rounds, c[0], c[1], c[2], c[3], k[0], k[1], out);
}
int orc_philox_rounds(void) { return KLB_PHILOX_ROUNDS; }
void orc_normals(uint64_t seed, uint64_t chain, uint64_t t, int64_t n, double* out) {
  klb_stream s = klb_stream_make(seed, chain, t);
  for (int64_t i = 0; i < n; ++i) out[i] = klb_normal(&s, (uint32_t)i, KLB_TAB);
}
double orc_uniform(uint64_t seed, uint64_t chain, uint64_t t) {
  klb_stream s = klb_stream_make(seed, chain, t);
  return klb_accept_uniform(&s);
}
/* units per lane the library uses for `dim` (klb_job_plan().nv): 1,2,4,8,16 for dim <= 1024, 64 up to 4096 */
/* uniform number q of transition t: slot q of KLB_TAG_ACCEPT (q = 0 is the accept uniform of HMC / MALA / MH; NUTS draws
 * q = 0, 1, 2, ... in the order the reference calls rand()) */
double orc_uniform_seq(uint64_t seed, uint64_t chain, uint64_t t, uint32_t q) {
  klb_stream st = klb_stream_make(seed, chain, t);
  uint64_t w0, w1;
  klb_stream_draw(&st, q, KLB_TAG_ACCEPT, 0u, &w0, &w1);
  return klb_u01(w0);
}
int orc_plan_nv(int64_t dim) {
  int nv = 1;
  while (64 * (int64_t)nv < dim && nv < 16) nv *= 2;
  if (64 * (int64_t)nv >= dim) return nv;
  return dim <= 4096 ? 64 : -1;
}

/* ------------------------------------------------- canonical reduction order
 * Lane l (0..31) owns double2 units k = l + 32 m, m = 0..nv-1, i.e. elements 2k, 2k+1.
 * Each lane keeps four accumulators; unit m adds its two addends, even element first,
 * into accumulator m & 3; lane value = (acc0+acc1)+(acc2+acc3); lanes are combined by
 * the xor butterfly 16, 8, 4, 2, 1.  `e` holds the addends (padded length 64*nv). */
static double orc_reduce_n(const double* e, int nv, int64_t n) {
  double lane[32], nxt[32];
  if (nv == 0) {                     /* sequential order (thread-per-chain kernels) */
    double acc = 0.;
    for (int64_t i = 0; i < n; ++i) acc = acc + e[i];
    return acc;
  }
  for (int l = 0; l < 32; ++l) {
    double acc[4] = {0., 0., 0., 0.};
    for (int m = 0; m < nv; ++m) {
      int k = l + 32 * m;
      acc[m & 3] = acc[m & 3] + e[2 * k];
      acc[m & 3] = acc[m & 3] + e[2 * k + 1];
    }
    lane[l] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
  for (int s = 16; s >= 1; s >>= 1) {
    for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ s];
    memcpy(lane, nxt, sizeof lane);
  }
  return lane[0];
}
static double orc_reduce(const double* e, int nv) { return orc_reduce_n(e, nv, 0); }
/* same order with the products folded in by fma */
static double orc_reduce_fma_n(const double* a, const double* b, int nv, int64_t n) {
  double lane[32], nxt[32];
  if (nv == 0) {
    double acc = 0.;
    for (int64_t i = 0; i < n; ++i) acc = fma(a[i], b[i], acc);
    return acc;
  }
  for (int l = 0; l < 32; ++l) {
    double acc[4] = {0., 0., 0., 0.};
    for (int m = 0; m < nv; ++m) {
      int k = l + 32 * m;
      acc[m & 3] = fma(a[2 * k], b[2 * k], acc[m & 3]);
      acc[m & 3] = fma(a[2 * k + 1], b[2 * k + 1], acc[m & 3]);
    }
    lane[l] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
  for (int s = 16; s >= 1; s >>= 1) {
    for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ s];
    memcpy(lane, nxt, sizeof lane);
  }
  return lane[0];
}

/* dot(a, b) over padded vectors.  Julia's dot on Vector{Float64} is BLAS ddot: an fma kernel with an
 * unspecified order on every FMA-capable CPU.  Here: fma-accumulated in the canonical order, in both
 * arithmetic modes (only the elementwise broadcast expressions of the reference are un-fused). */
static double orc_reduce_fma(const double* a, const double* b, int nv) { return orc_reduce_fma_n(a, b, nv, 0); }
static double orc_dot_m(const orc_model* M, const double* a, const double* b, double* scratch) {
  (void)scratch;
  return orc_reduce_fma_n(a, b, M->cfg->nv, M->d);
}
/* exported for tests */
double orc_dot(const double* a, const double* b, int64_t d, int nv, int arith) {
  int64_t dp = 64 * (int64_t)nv;
  double* pa = calloc(3 * dp, sizeof(double)); double* pb = pa + dp; double* sc = pb + dp;
  memcpy(pa, a, d * sizeof(double)); memcpy(pb, b, d * sizeof(double));
  (void)arith; (void)sc;
  double r = orc_reduce_fma(pa, pb, nv);
  free(pa);
  return r;
}

/* a*b + c as the elementwise expressions evaluate it: two roundings in reference arithmetic, one in fma
 * arithmetic (exported for the test that the build really keeps them apart) */
double orc_ma(double a, double b, double c, int arith) { return arith == 1 ? fma(a, b, c) : a * b + c; }

/* ------------------------------------------------------------ target library
 * Device targets are descriptors, not closures (a Julia closure cannot run on the GPU);
 * the definitions follow the reference's examples:
 *   iso      plogtarget(z) = -dot(z, z), pgradlogtarget(z) = -2*z            README.md:153-155
 *   shifted  -(x-mu).(x-mu), -2(x-mu)                    test/BasicContMuvParameter.jl:539-563
 *   dense    -p.(C p), -2 C p        doc/examples/BivariateNormal/MALA/function/analytical.jl:8-9
 *   rosen    this repo's paired Rosenbrock; not in the reference
 *   logit    Bayesian logistic regression with a N(0, lambda I) prior, v = [lambda, X, y, p]:
 *              ploglikelihood(p, v) = dot(Xp, y) - sum(log.(1+exp.(Xp))),  Xp = X*p
 *              plogprior(p, v)      = -0.5*(dot(p, p)/lambda + length(p)*log(2*pi*lambda))
 *              pgradlogtarget(p, v) = X'*(y - 1./(1+exp.(-X*p))) - p/lambda
 *            doc/examples/swiss/HMC/noadaptation/analytical.jl:11-20 (same closures in swiss/MALA/analytical.jl);
 *            logtarget = loglikelihood + logprior      BasicContMuvParameter.jl:185-190
 *            X*p, X'*r and dot are BLAS calls in the reference (fma kernels, unspecified order): here fma
 *            chains in increasing index; sum() is a chain of additions in increasing index.
 */
static void orc_gradlogtarget(const orc_model* M, orc_pstate* s, double* scratch);

static void orc_logtarget(const orc_model* M, orc_pstate* s, double* scratch) {
  const double* x = s->value;
  const int fm = M->cfg->arith == 1;
  switch (M->cfg->target) {
    case ORC_ISO:
      s->logtarget = -orc_dot_m(M, x, x, scratch);
      break;
    case ORC_SHIFTED: {
      double* df = scratch + M->dp;
      for (int64_t i = 0; i < M->dp; ++i) df[i] = (i < M->d) ? x[i] - M->mu[i] : 0.;
      s->logtarget = -orc_dot_m(M, df, df, scratch);
      break;
    }
    case ORC_DENSE: {
      /* (C x)_i accumulated by fma in increasing j (both arith modes: the reference's gemv order
       * is BLAS-defined, i.e. unspecified), then -x.(Cx) in the canonical order */
      double* cx = scratch + M->dp;
      for (int64_t i = 0; i < M->dp; ++i) {
        double acc = 0.;
        if (i < M->d) for (int64_t j = 0; j < M->d; ++j) acc = fma(M->C[i * M->d + j], x[j], acc);
        cx[i] = acc;
      }
      s->logtarget = -orc_dot_m(M, x, cx, scratch);
      break;
    }
    case ORC_ROSEN: {
      /* -(scale) * sum_k [ rb*(b - a^2)^2 + (ra - a)^2 ],  (a,b) = (x[2k], x[2k+1]) */
      for (int64_t k = 0; 2 * k < M->dp; ++k) {
        double a = x[2 * k], b = x[2 * k + 1], term = 0.;
        if (2 * k + 1 < M->d) {
          double u = fm ? fma(-a, a, b) : b - a * a;
          double v = M->ra - a;
          term = fm ? fma(M->rb * u, u, v * v) : M->rb * (u * u) + v * v;
        }
        scratch[2 * k] = term; scratch[2 * k + 1] = 0.;
      }
      s->logtarget = -(M->rscale * orc_reduce(scratch, M->cfg->nv));
      break;
    }
    case ORC_LOGIT: {
      const int64_t d = M->d;
      double a = 0., sl = 0.;
      for (int64_t i = 0; i < M->ndata; ++i) {
        double xp = 0.;
        for (int64_t j = 0; j < d; ++j) xp = fma(M->X[i * d + j], x[j], xp);      /* Xp = v[2]*p */
        a = fma(xp, M->y[i], a);                                                  /* dot(Xp, v[3]) */
        sl = sl + klb_log(1 + klb_exp(xp, KLB_TAB), KLB_TAB);                     /* sum(log.(1+exp.(Xp))) */
      }
      const double loglik = a - sl;
      const double pp = orc_dot_m(M, x, x, scratch) / M->lambda;
      const double lc = klb_log((2 * 3.141592653589793) * M->lambda, KLB_TAB);    /* log(2*pi*v[1]) */
      const double logprior = -0.5 * (fm ? fma((double)d, lc, pp) : pp + (double)d * lc);
      s->logtarget = loglik + logprior;
      break;
    }
  }
}

static void orc_gradlogtarget(const orc_model* M, orc_pstate* s, double* scratch) {
  const double* x = s->value; double* g = s->gradlogtarget;
  const int fm = M->cfg->arith == 1;
  switch (M->cfg->target) {
    case ORC_ISO:
      for (int64_t i = 0; i < M->dp; ++i) g[i] = -2 * x[i];
      break;
    case ORC_SHIFTED:
      for (int64_t i = 0; i < M->dp; ++i) g[i] = (i < M->d) ? -2 * (x[i] - M->mu[i]) : 0.;
      break;
    case ORC_DENSE:
      for (int64_t i = 0; i < M->dp; ++i) {
        double acc = 0.;
        if (i < M->d) for (int64_t j = 0; j < M->d; ++j) acc = fma(M->C[i * M->d + j], x[j], acc);
        g[i] = -2 * acc;
      }
      break;
    case ORC_ROSEN:
      for (int64_t k = 0; 2 * k < M->dp; ++k) {
        double a = x[2 * k], b = x[2 * k + 1], ga = 0., gb = 0.;
        if (2 * k + 1 < M->d) {
          double u = fm ? fma(-a, a, b) : b - a * a;
          double v = M->ra - a;
          double t = ((4 * M->rb) * a) * u;
          ga = M->rscale * (fm ? fma(2., v, t) : t + 2 * v);
          gb = -(M->rscale * ((2 * M->rb) * u));
        }
        g[2 * k] = ga; g[2 * k + 1] = gb;
      }
      break;
    case ORC_LOGIT: {
      const int64_t d = M->d;
      for (int64_t j = 0; j < M->dp; ++j) g[j] = 0.;
      for (int64_t i = 0; i < M->ndata; ++i) {
        double xp = 0.;
        for (int64_t j = 0; j < d; ++j) xp = fma(M->X[i * d + j], x[j], xp);
        /* -v[2]*p = (-X)*p = -(X*p) exactly (negation commutes with every rounding of the fma chain) */
        const double r = M->y[i] - 1. / (1 + klb_exp(-xp, KLB_TAB));
        for (int64_t j = 0; j < d; ++j) g[j] = fma(M->X[i * d + j], r, g[j]);     /* v[2]'*r */
      }
      for (int64_t j = 0; j < d; ++j) g[j] = g[j] - x[j] / M->lambda;
      break;
    }
  }
  (void)scratch; (void)fm;
}

/* uptogradlogtarget! = logtarget! then gradlogtarget!   BasicContMuvParameter.jl:264-279 */
static void orc_upto(const orc_model* M, orc_pstate* s, double* scratch) {
  orc_logtarget(M, s, scratch);
  orc_gradlogtarget(M, s, scratch);
}

/* ------------------------------------------------------------------- tuners */
static void orc_rate(orc_tune* t) { t->rate = (double)t->accepted / (double)t->proposed; }
static void orc_reset_burnin(orc_tune* t) {
  t->totproposed += t->proposed;
  t->accepted = 0; t->proposed = 0; t->rate = NAN;
}
static void orc_tune_step(orc_tune* t, const orc_config* c) {
  /* tune!: tune.step *= tuner.score(tune.rate - tuner.targetrate)   AcceptanceRateMCTuner.jl:46 */
  t->step *= c->score == 1 ? orc_erf_rate_score(t->rate - c->target_rate, c->score_k)
                           : orc_logistic_rate_score(t->rate - c->target_rate, c->score_k);
}
/* tuner_state: BasicMCTune(step, 0, 0, tuner.period); MH gets step 1.   samplers.jl:29-45 */
void orc_tuner_state(const orc_config* c, orc_tune* t) {
  t->step = (c->sampler == ORC_MH) ? 1. : c->step;
  t->accepted = 0; t->proposed = 0; t->totproposed = c->period; t->rate = NAN;
}
static int orc_counters_on(const orc_config* c) {
  if (c->sampler == ORC_MH) return c->verbose != 0;                    /* iterate/MH.jl:73-75 */
  if (c->sampler == ORC_NUTS) return c->verbose != 0;                  /* iterate/NUTS.jl:238-240 */
  return (c->tuner == ORC_VANILLA && c->verbose) || c->tuner == ORC_ACCRATE ||
         (c->tuner == ORC_DUALAVG && c->verbose);                      /* iterate/HMC.jl:129-133 */
}

/* ---- DualAveragingMCTuner for HMC
 * tuner_state (src/samplers/HMC.jl:124-133) + sampler_state (HMC.jl:192-213):
 *   step = leapstep, lambda = nleaps*leapstep, ebar = e0bar, hbar = h0bar, counters (0, 0, period), then
 *   step = initialize_step!(...), mu = log(10*step), count = 0.
 * initialize_step! (src/samplers/samplers.jl:170-202) returns step0 unchanged for every function-defined target:
 * it compares hamiltonian(pstate.logtarget, momentum) of the FRESH proposal state with the old one, but its
 * leapfrog! only calls gradlogtarget! (samplers.jl:131), so pstate.logtarget still holds the NaN the state was
 * constructed with (BasicContMuvParameterState.jl:107) -> ratio = NaN -> a = 2*(exp(NaN) > 0.5)-1 = -1 ->
 * `while exp(NaN)^-1 > 2` is false -> the doubling loop (and its reference to the undefined `moment`, :195) is
 * never entered.  Its randn(d) draw and scratch leapfrog leave no trace in a counter-based RNG.
 * reset!(tune, ::HMC, ::DualAveragingMCTuner) (HMC.jl:217-223) sets step = 1 (sic), not leapstep; `first` selects. */
void orc_da_state(const orc_config* c, orc_tune* t, orc_da* d, int first) {
  t->step = first ? c->step : 1.;
  t->accepted = 0; t->proposed = 0; t->totproposed = c->period; t->rate = NAN;
  d->lambda = (c->sampler == ORC_NUTS) ? NAN : (double)c->nleaps * c->step;      /* NUTS: λ=NaN   NUTS.jl:260-269 */
  d->epsbar = c->da_eps0bar; d->hbar = c->da_h0bar;
  d->hweight = NAN; d->epsweight = NAN; d->nleaps = 0.;
  d->mu = klb_log(10 * t->step, KLB_TAB);
  d->count = 0.;
}
/* tune!(tune, tuner, count, a)                              src/tuners/DualAveragingMCTuner.jl:95-101 */
static void orc_da_tune(orc_tune* t, orc_da* d, const orc_config* c, double a) {
  const double count = d->count;
  d->hweight = 1 / (count + (double)c->da_t0);
  d->hbar = (1 - d->hweight) * d->hbar + d->hweight * (c->target_rate - a);
  t->step = klb_exp(d->mu - sqrt(count) * d->hbar / c->da_gamma, KLB_TAB);
  d->epsweight = klb_pow_pos(count, -c->da_kappa, KLB_TAB);
  d->epsbar = klb_exp((1 - d->epsweight) * klb_log(d->epsbar, KLB_TAB) + d->epsweight * klb_log(t->step, KLB_TAB), KLB_TAB);
}
/* nleaps = max(1, Int(round(lambda/step)))                  src/samplers/iterate/HMC.jl:142-144
 * round = ties to even.  Int() of a NaN / infinite / huge quotient throws InexactError in the reference (a chain
 * whose adaptation has diverged kills the job); here such a chain takes one leapfrog step per transition. */
static int64_t orc_da_nleaps(double lambda, double step) {
  const double q = rint(lambda / step);
  if (!(q >= 1.)) return 1;
  if (q > 2147483647.) return 1;
  return (int64_t)q;
}
/* the DualAveragingMCTuner branch of the burn-in block         src/samplers/iterate/HMC.jl:225-248 */
static void orc_da_block(const orc_config* c, orc_tune* t, orc_da* d, double a) {
  if (d->count <= (double)c->da_nadapt) {
    orc_da_tune(t, d, c, a);
    if (c->verbose && t->proposed % c->period == 0) { orc_rate(t); orc_reset_burnin(t); }
  } else {
    t->step = d->epsbar;
  }
}
/* the burn-in tuner block shared by HMC (:203-224) and MALA (:130-152); MH (:126-140) never tunes */
static void orc_tuner_block(const orc_config* c, orc_tune* t) {
  if (!orc_counters_on(c)) return;
  if (t->totproposed <= c->burnin && t->proposed % c->period == 0) {
    orc_rate(t);
    if (c->tuner == ORC_ACCRATE && c->sampler != ORC_MH) orc_tune_step(t, c);
    orc_reset_burnin(t);
  }
}

/* ------------------------------------------------------------------ samplers */
typedef struct {
  orc_pstate sp;        /* sstate.pstate: the proposal */
  double* momentum;     /* HMC momentum / MALA mu / MH scratch normals */
  double* z;            /* normals */
  double* scratch;      /* 3*dp */
} orc_sstate;

static double orc_hamiltonian(const orc_model* M, double logtarget, const double* p, double* scratch) {
  return logtarget - 0.5 * orc_dot_m(M, p, p, scratch);
}

static void orc_leapfrog(const orc_model* M, orc_pstate* sp, double* mom, double step, double* scratch) {
  const double h = 0.5 * step;                    /* Julia folds 0.5*step*g left to right */
  if (M->cfg->arith == 1) {
    for (int64_t i = 0; i < M->dp; ++i) mom[i] = fma(h, sp->gradlogtarget[i], mom[i]);
    for (int64_t i = 0; i < M->dp; ++i) sp->value[i] = fma(step, mom[i], sp->value[i]);
    orc_gradlogtarget(M, sp, scratch);
    for (int64_t i = 0; i < M->dp; ++i) mom[i] = fma(h, sp->gradlogtarget[i], mom[i]);
  } else {
    for (int64_t i = 0; i < M->dp; ++i) mom[i] = mom[i] + h * sp->gradlogtarget[i];
    for (int64_t i = 0; i < M->dp; ++i) sp->value[i] = sp->value[i] + step * mom[i];
    orc_gradlogtarget(M, sp, scratch);
    for (int64_t i = 0; i < M->dp; ++i) mom[i] = mom[i] + h * sp->gradlogtarget[i];
  }
}

static void orc_randn(const orc_model* M, const klb_stream* st, double* z) {
  for (int64_t i = 0; i < M->dp; ++i) z[i] = (i < M->d) ? klb_normal(st, (uint32_t)i, KLB_TAB) : 0.;
}

static void orc_iterate_hmc(const orc_model* M, orc_pstate* ps, orc_sstate* ss, orc_tune* tune, orc_da* da,
                            const klb_stream* st) {
  const orc_config* c = M->cfg;
  if (c->tuner == ORC_DUALAVG) da->count += 1;                         /* job.sstate.count += 1   iterate/HMC.jl:125-127 */
  if (orc_counters_on(c)) tune->proposed += 1;
  orc_randn(M, st, ss->momentum);
  double oldh = orc_hamiltonian(M, ps->logtarget, ss->momentum, ss->scratch);
  memcpy(ss->sp.value, ps->value, M->dp * sizeof(double));
  memcpy(ss->sp.gradlogtarget, ps->gradlogtarget, M->dp * sizeof(double));
  int64_t nleaps = c->nleaps;
  if (c->tuner == ORC_DUALAVG) { nleaps = orc_da_nleaps(da->lambda, tune->step); da->nleaps = (double)nleaps; }
  for (int64_t i = 0; i < nleaps; ++i) orc_leapfrog(M, &ss->sp, ss->momentum, tune->step, ss->scratch);
  orc_logtarget(M, &ss->sp, ss->scratch);
  double newh = orc_hamiltonian(M, ss->sp.logtarget, ss->momentum, ss->scratch);
  double ratio = newh - oldh;
  double e = klb_exp(ratio, KLB_TAB);
  double a = (e != e) ? e : (e < 1. ? e : 1.);       /* min(1., exp(ratio)) keeps NaN */
  if (klb_accept_uniform(st) < a) {
    memcpy(ps->value, ss->sp.value, M->dp * sizeof(double));
    memcpy(ps->gradlogtarget, ss->sp.gradlogtarget, M->dp * sizeof(double));
    ps->logtarget = ss->sp.logtarget;
    ps->accept = 1;
    if (orc_counters_on(c)) tune->accepted += 1;
  } else {
    ps->accept = 0;
  }
  if (c->tuner == ORC_DUALAVG) orc_da_block(c, tune, da, a); else orc_tuner_block(c, tune);
}

/* ---- NUTS, multivariate                src/samplers/iterate/NUTS.jl:230-457, src/samplers/NUTS.jl:514-628, :781-927
 * The reference constructs its sampler state as MuvNUTSState(pstate, pstate, pstate, pstate, ...) (NUTS.jl:198-225): the
 * plus end, the minus end, the proposal and the second-subtree proposal are ONE mutable object, every leaf of build_tree!
 * returns sstate.pstateprime / sstate.momentumprime for all of them (:533-539) and every level keeps n', s' in the one
 * shared sstate (:541-549).  Resolving that aliasing by Julia's reference semantics (oracle/nuts_alias.py holds the
 * statement-by-statement model and tests/test_oracle_nuts.py checks the two against each other) leaves:
 *   - one moving point E (sp below) that every leaf advances in place by one leapfrog step of size v*step;
 *   - one running momentum (momentumprime); each direction owns a pristine copy of the initial momentum until it is first
 *     used: the first leaf of a doubling in a direction that has not been used reads that copy, afterwards that end (at
 *     j >= 1 both ends, :541-549) IS momentumprime;
 *   - uturn(E.value - E.value, ...) = (0 < 0) never fires; the rand() of an inner node only decides a self-copy, but is
 *     consumed; an inner node returns n' = 2 n'(second half), s' = s'(second half) (the first half's n' lived in the
 *     same field), a' and na' do add up (they are locals, :879-880);
 *   - job.pstate takes E's value, gradient and log-target whenever `s' && rand() < n'/n` (iterate/NUTS.jl:355-375).
 * Uniforms are consumed in order from the transition's stream: slot q of KLB_TAG_ACCEPT, q = 0 the slice variable, then
 * per doubling the direction (rand(Bool) := uniform < 0.5 -> +1), the inner nodes in post-order, the acceptance test. */
static double orc_sequ(const klb_stream* st, uint32_t* q) {
  uint64_t w0, w1;
  klb_stream_draw(st, (*q)++, KLB_TAG_ACCEPT, 0u, &w0, &w1);
  return klb_u01(w0);
}
#define ORC_NUTS_MAXLEVELS 16
static void orc_nuts_tree(const orc_model* M, orc_sstate* ss, double step_v, int j, double u, double oldh,
                          const klb_stream* st, uint32_t* q, int64_t* n_out, int* s_out, double* a_out, int64_t* na_out) {
  const orc_config* c = M->cfg;
  double saved_a[ORC_NUTS_MAXLEVELS + 1]; int64_t saved_na[ORC_NUTS_MAXLEVELS + 1];
  for (uint64_t leaf = 0;; ++leaf) {
    orc_leapfrog(M, &ss->sp, ss->momentum, step_v, ss->scratch);                       /* NUTS.jl:527 */
    orc_logtarget(M, &ss->sp, ss->scratch);
    const double hprime = orc_hamiltonian(M, ss->sp.logtarget, ss->momentum, ss->scratch);
    int64_t n = (u <= hprime) ? 1 : 0;                                                 /* :532 */
    const int s = u < (double)c->nuts_maxdelta + hprime;                               /* :533 */
    const double e = klb_exp(hprime - oldh, KLB_TAB);
    double a = (e != e) ? e : (e < 1. ? e : 1.);                                       /* min(1, exp(H' - H0))   :818 */
    int64_t na = 1;
    int k = 1, descend = 0;
    while (k <= j) {
      if ((leaf >> (k - 1)) & 1u) {                 /* the block just finished was a second half   :580-603 */
        (void)orc_sequ(st, q);
        n = 2 * n;
        a = saved_a[k] + a; na = saved_na[k] + na;
        ++k;
      } else if (s) {                               /* a first half that did not stop: on to the second */
        saved_a[k] = a; saved_na[k] = na;
        descend = 1;
        break;
      } else ++k;                                   /* a first half that stopped is returned as it is */
    }
    if (!descend) { *n_out = n; *s_out = s; *a_out = a; *na_out = na; return; }
  }
}
static void orc_iterate_nuts(const orc_model* M, orc_pstate* ps, orc_sstate* ss, orc_tune* tune, orc_da* da,
                             const klb_stream* st) {
  const orc_config* c = M->cfg;
  if (c->tuner == ORC_DUALAVG) da->count += 1;                                         /* iterate/NUTS.jl:234-236 */
  if (c->verbose) tune->proposed += 1;                                                 /* :238-240 */
  orc_randn(M, st, ss->z);                                                             /* momentum[:] = randn(size) */
  const double oldh = orc_hamiltonian(M, ps->logtarget, ss->z, ss->scratch);
  memcpy(ss->sp.value, ps->value, M->dp * sizeof(double));                             /* E = copies of job.pstate   :246-251 */
  memcpy(ss->sp.gradlogtarget, ps->gradlogtarget, M->dp * sizeof(double));
  int aliased_plus = 0, aliased_minus = 0, j = 0, s = 1, update = 0;
  int64_t n = 1, na = 1;
  double a = NAN;
  uint32_t q = 0;
  const double u = klb_log(orc_sequ(st, &q), KLB_TAB) + oldh;                          /* :261 */
  while (s && j < c->nuts_maxndoublings) {
    const int v = (orc_sequ(st, &q) < 0.5) ? 1 : -1;                                   /* rand(Bool) ? 1 : -1   :264 */
    if (!(v == 1 ? aliased_plus : aliased_minus)) memcpy(ss->momentum, ss->z, M->dp * sizeof(double));
    int64_t nprime; int sprime;
    orc_nuts_tree(M, ss, (double)v * tune->step, j, u, oldh, st, &q, &nprime, &sprime, &a, &na);
    if (v == 1) aliased_plus = 1; else aliased_minus = 1;
    if (j >= 1) aliased_plus = aliased_minus = 1;
    if (sprime && orc_sequ(st, &q) < (double)nprime / (double)n) {                     /* :355-375 */
      memcpy(ps->value, ss->sp.value, M->dp * sizeof(double));
      memcpy(ps->gradlogtarget, ss->sp.gradlogtarget, M->dp * sizeof(double));
      ps->logtarget = ss->sp.logtarget;
      update = 1;
    }
    j += 1; n += nprime; s = sprime;                                                   /* :377-381 */
  }
  ps->accept = update; ps->ndoublings = j;                                             /* :384-399 */
  ps->nuts_a = a; ps->nuts_na = na;                                                    /* :a, :na of the LAST doubling   :393-399 */
  if (c->verbose && update) tune->accepted += 1;                                       /* :402-404 */
  if (c->tuner == ORC_DUALAVG) {                                                       /* :424-447 */
    da->nleaps = (double)na;
    if (da->count <= (double)c->da_nadapt) orc_da_tune(tune, da, c, a / (double)na); else tune->step = da->epsbar;
    if (c->verbose && tune->totproposed <= c->burnin && tune->proposed % c->period == 0) { orc_rate(tune); orc_reset_burnin(tune); }
  } else if (c->verbose) {                                                             /* :406-423 */
    if (tune->totproposed <= c->burnin && tune->proposed % c->period == 0) { orc_rate(tune); orc_reset_burnin(tune); }
  }
}

static void orc_iterate_mala(const orc_model* M, orc_pstate* ps, orc_sstate* ss, orc_tune* tune,
                             const klb_stream* st) {
  const orc_config* c = M->cfg;
  const int fm = c->arith == 1;
  double* mu = ss->momentum; double* e = ss->scratch + 2 * M->dp;
  if (orc_counters_on(c)) tune->proposed += 1;
  const double step = tune->step, h = 0.5 * step, sq = sqrt(step), hinv = 0.5 / step;
  orc_randn(M, st, ss->z);
  for (int64_t i = 0; i < M->dp; ++i)
    mu[i] = fm ? fma(h, ps->gradlogtarget[i], ps->value[i]) : ps->value[i] + h * ps->gradlogtarget[i];
  for (int64_t i = 0; i < M->dp; ++i)
    ss->sp.value[i] = fm ? fma(sq, ss->z[i], mu[i]) : mu[i] + sq * ss->z[i];
  orc_upto(M, &ss->sp, ss->scratch);
  double ratio = ss->sp.logtarget - ps->logtarget;
  for (int64_t i = 0; i < M->dp; ++i) {
    double df = mu[i] - ss->sp.value[i];
    e[i] = fm ? (df * hinv) * df : 0.5 * ((df * df) / step);
  }
  ratio += orc_reduce_n(e, c->nv, M->d);
  for (int64_t i = 0; i < M->dp; ++i)
    mu[i] = fm ? fma(h, ss->sp.gradlogtarget[i], ss->sp.value[i]) : ss->sp.value[i] + h * ss->sp.gradlogtarget[i];
  for (int64_t i = 0; i < M->dp; ++i) {
    double df = mu[i] - ps->value[i];
    e[i] = fm ? (df * hinv) * df : 0.5 * ((df * df) / step);
  }
  ratio -= orc_reduce_n(e, c->nv, M->d);
  /* the counter-based uniform makes Julia's short-circuit (no draw when ratio > 0) unobservable */
  if (ratio > 0 || ratio > klb_log(klb_accept_uniform(st), KLB_TAB)) {
    memcpy(ps->value, ss->sp.value, M->dp * sizeof(double));
    memcpy(ps->gradlogtarget, ss->sp.gradlogtarget, M->dp * sizeof(double));
    ps->logtarget = ss->sp.logtarget;
    ps->accept = 1;
    if (orc_counters_on(c)) tune->accepted += 1;
  } else {
    ps->accept = 0;
  }
  orc_tuner_block(c, tune);
}

static void orc_iterate_mh(const orc_model* M, orc_pstate* ps, orc_sstate* ss, orc_tune* tune,
                           const klb_stream* st) {
  const orc_config* c = M->cfg;
  if (orc_counters_on(c)) tune->proposed += 1;
  orc_randn(M, st, ss->z);
  /* rand(MvNormal(x, sigma)) = unwhiten (sigma .* z) then add the mean        iterate/MH.jl:77-79 */
  for (int64_t i = 0; i < M->dp; ++i)
    ss->sp.value[i] = (c->arith == 1) ? fma(M->sigma[i], ss->z[i], ps->value[i])
                                      : M->sigma[i] * ss->z[i] + ps->value[i];
  orc_logtarget(M, &ss->sp, ss->scratch);
  double ratio = ss->sp.logtarget - ps->logtarget;
  if (ratio > 0 || ratio > klb_log(klb_accept_uniform(st), KLB_TAB)) {
    memcpy(ps->value, ss->sp.value, M->dp * sizeof(double));
    ps->logtarget = ss->sp.logtarget;
    ps->accept = 1;
    if (orc_counters_on(c)) tune->accepted += 1;
  } else {
    ps->accept = 0;
  }
  orc_tuner_block(c, tune);
}

/* tparams of the logit target: {lambda, ndata, X (ndata x d, row-major), y (ndata)} */
static void orc_model_logit(orc_model* M, const double* tparams) {
  M->lambda = tparams[0];
  M->ndata = (int64_t)tparams[1];
  M->X = tparams + 2;
  M->y = tparams + 2 + M->ndata * M->d;
}

/* ----------------------------------------------------------------- the job */
typedef struct {
  double* value;       /* d x npost x nchains */
  double* logtarget;   /* npost x nchains */
  double* grad;        /* d x npost x nchains */
  uint8_t* accept;     /* npost x nchains */
} orc_output;

static void orc_save(const orc_model* M, const orc_pstate* ps, const orc_output* o,
                     int64_t chain, int64_t npost, int64_t count /* 1-based column */) {
  const orc_config* c = M->cfg;
  int64_t col = chain * npost + (count - 1);
  if ((c->monitor & 1u) && o->value) memcpy(o->value + col * M->d, ps->value, M->d * sizeof(double));
  if ((c->monitor & 2u) && o->logtarget) o->logtarget[col] = ps->logtarget;
  if ((c->monitor & 4u) && o->grad) memcpy(o->grad + col * M->d, ps->gradlogtarget, M->d * sizeof(double));
  if ((c->diagnostics & 1u) && o->accept) o->accept[col] = (uint8_t)ps->accept;
  if ((c->diagnostics & 2u) && c->nuts_ndoublings) c->nuts_ndoublings[col] = (uint8_t)ps->ndoublings;
  if ((c->diagnostics & 4u) && c->nuts_a) c->nuts_a[col] = ps->nuts_a;
  if ((c->diagnostics & 8u) && c->nuts_na) c->nuts_na[col] = (int32_t)ps->nuts_na;
}

int64_t orc_npoststeps(int64_t burnin, int64_t thinning, int64_t nsteps) {
  /* length((burnin+1):thinning:nsteps)                       src/ranges/BasicMCRange.jl:14-25 */
  if (nsteps <= burnin) return 0;
  return (nsteps - burnin - 1) / thinning + 1;
}

/* Run all chains: the semantics of run(::Vector{MCJob}) = map(run, job) (src/jobs/jobs.jl:212).
 * x (d x nchains, in/out), logtarget (nchains, in/out; must already hold logtarget(x) when
 * `initialized`, else it is computed = initialize!), tune (nchains, in/out).
 * tparams: mu (d) for SHIFTED, C (d*d) for DENSE, {a, b, scale} for ROSEN; sigma (d) for MH.
 * Returns 0, or -(1+chain) if that chain's initial log-target/gradient is not finite. */
int orc_run(const orc_config* cfg, const double* tparams, const double* sigma,
            double* x, double* logtarget, orc_tune* tune, int initialized,
            double* out_value, double* out_logtarget, double* out_grad, uint8_t* out_accept) {
  const int64_t d = cfg->dim, dp = cfg->nv ? 64 * (int64_t)cfg->nv : cfg->dim, N = cfg->nchains;
  if (dp < d) return -1000000;
  const int64_t npost = orc_npoststeps(cfg->burnin, cfg->thinning, cfg->nsteps);
  orc_output out = {out_value, out_logtarget, out_grad, out_accept};
  double* mu_p = calloc(2 * dp, sizeof(double)); double* sg_p = mu_p + dp;
  orc_model M; memset(&M, 0, sizeof M);
  M.cfg = cfg; M.d = d; M.dp = dp; M.mu = mu_p; M.sigma = sg_p; M.C = NULL;
  if (cfg->target == ORC_SHIFTED && tparams) memcpy(mu_p, tparams, d * sizeof(double));
  if (cfg->target == ORC_DENSE) M.C = tparams;
  if (cfg->target == ORC_ROSEN) { M.ra = tparams[0]; M.rb = tparams[1]; M.rscale = tparams[2]; }
  if (cfg->target == ORC_LOGIT) orc_model_logit(&M, tparams);
  if (sigma) memcpy(sg_p, sigma, d * sizeof(double));
  int bad = 0;
  int nth = cfg->nthreads > 0 ? cfg->nthreads : 1;
  (void)nth;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nth)
  for (int64_t c = 0; c < N; ++c) {
    double* buf = calloc(9 * dp, sizeof(double));
    orc_pstate ps = {buf, 0., buf + dp, 0, 0};
    orc_sstate ss; ss.sp.value = buf + 2 * dp; ss.sp.gradlogtarget = buf + 3 * dp; ss.sp.logtarget = NAN; ss.sp.accept = 0;
    ss.momentum = buf + 4 * dp; ss.z = buf + 5 * dp; ss.scratch = buf + 6 * dp;
    memcpy(ps.value, x + c * d, d * sizeof(double));
    /* initialize!: first target (+ gradient) evaluation, finiteness asserts   HMC.jl:106-120 */
    if (cfg->sampler == ORC_MH) orc_logtarget(&M, &ps, ss.scratch); else orc_upto(&M, &ps, ss.scratch);
    if (initialized) ps.logtarget = logtarget[c];
    int ok = isfinite(ps.logtarget);
    if (cfg->sampler != ORC_MH) for (int64_t i = 0; i < d; ++i) ok = ok && isfinite(ps.gradlogtarget[i]);
    if (!ok) {
#pragma omp critical
      { if (bad == 0 || -(int)(1 + c) > bad) bad = -(int)(1 + c); }
      free(buf);
      continue;
    }
    orc_tune tn = tune[c];
    orc_da da; memset(&da, 0, sizeof da);
    if (cfg->tuner == ORC_DUALAVG && cfg->da) da = ((const orc_da*)cfg->da)[c];
    int64_t count = 0;
    for (int64_t i = 1; i <= cfg->nsteps; ++i) {                       /* BasicMCJob.jl:219 */
      klb_stream st = klb_stream_make(cfg->seed, cfg->chain_offset + (uint64_t)c, cfg->t0 + (uint64_t)i);
      switch (cfg->sampler) {                                          /* BasicMCJob.jl:224 */
        case ORC_HMC: orc_iterate_hmc(&M, &ps, &ss, &tn, &da, &st); break;
        case ORC_NUTS: orc_iterate_nuts(&M, &ps, &ss, &tn, &da, &st); break;
        case ORC_MALA: orc_iterate_mala(&M, &ps, &ss, &tn, &st); break;
        default: orc_iterate_mh(&M, &ps, &ss, &tn, &st); break;
      }
      if (i > cfg->burnin && (i - cfg->burnin - 1) % cfg->thinning == 0) {   /* in(i, postrange) :226 */
        count += 1;
        orc_save(&M, &ps, &out, c, npost, count);
      }
    }
    memcpy(x + c * d, ps.value, d * sizeof(double));
    logtarget[c] = ps.logtarget;
    tune[c] = tn;
    if (cfg->tuner == ORC_DUALAVG && cfg->da) ((orc_da*)cfg->da)[c] = da;
    free(buf);
  }
  free(mu_p);
  return bad;
}

/* single evaluations for the closure-wiring KATs (test/BasicContMuvParameter.jl:539-563) */
int orc_eval_target(const orc_config* cfg, const double* tparams, const double* x,
                    double* logtarget, double* grad) {
  const int64_t d = cfg->dim, dp = cfg->nv ? 64 * (int64_t)cfg->nv : cfg->dim;
  if (dp < d) return -1;
  double* buf = calloc(6 * dp, sizeof(double));
  orc_model M; memset(&M, 0, sizeof M);
  M.cfg = cfg; M.d = d; M.dp = dp; M.mu = buf + 4 * dp; M.sigma = buf + 5 * dp;
  if (cfg->target == ORC_SHIFTED) memcpy(buf + 4 * dp, tparams, d * sizeof(double));
  if (cfg->target == ORC_DENSE) M.C = tparams;
  if (cfg->target == ORC_ROSEN) { M.ra = tparams[0]; M.rb = tparams[1]; M.rscale = tparams[2]; }
  if (cfg->target == ORC_LOGIT) orc_model_logit(&M, tparams);
  orc_pstate ps = {buf, 0., buf + dp, 0};
  memcpy(ps.value, x, d * sizeof(double));
  orc_upto(&M, &ps, buf + 2 * dp);
  *logtarget = ps.logtarget;
  memcpy(grad, ps.gradlogtarget, d * sizeof(double));
  free(buf);
  return 0;
}

/* ------------------------------------------------------------ post-hoc statistics
 * ess(v) = len * mcvar(v, :iid) / mcvar(v, :imse)                src/stats/convergence/ess.jl:3-14
 * mcvar(v, :iid)  = var(v)/length(v)                             src/stats/variance/mcvar.jl:5
 * mcvar(v, :imse, maxlag = length(v)-1): Geyer's initial monotone sequence estimator over
 *   acv = StatsBase.autocov(v, 0:maxlag) (demeaned, each lag divided by length(v))   mcvar.jl:75-105
 * Summation order (unspecified in Julia: pairwise sum / BLAS dot): sequential in t, fma-accumulated,
 * the same as the device kernel.  Lags are evaluated pairwise until the first non-positive G_j, which is
 * all the estimator ever reads. */
/* out[5] = {mean (mean.jl:9), mcvar iid (mcvar.jl:5), mcvar imse (mcvar.jl:75-105), ess (ess.jl:3), iact (iact.jl:3)} */
void orc_stats_series(const double* v, int64_t n, int64_t stride, double out[5]) {
  for (int q = 0; q < 5; ++q) out[q] = NAN;
  if (n < 1) return;
  double s = 0.;
  for (int64_t t = 0; t < n; ++t) s = s + v[t * stride];
  const double mu = s / (double)n;
  out[0] = mu;
  if (n < 4) return;
  double s0 = 0.;
  for (int64_t t = 0; t < n; ++t) { double z = v[t * stride] - mu; s0 = fma(z, z, s0); }
  const double iidvar = (s0 / (double)(n - 1)) / (double)n;       /* var(v)/length(v) */
  const double acv0 = s0 / (double)n;
  const int64_t maxlag = n - 1;
  const int64_t k = (int64_t)floor((double)(maxlag - 1) / 2.);
  double sumg = 0., gprev = 0.;
  for (int64_t j = 0; j <= k; ++j) {
    double a = 0., b = 0.;
    const int64_t l0 = 2 * j, l1 = 2 * j + 1;
    for (int64_t t = 0; t + l0 < n; ++t) a = fma(v[t * stride] - mu, v[(t + l0) * stride] - mu, a);
    for (int64_t t = 0; t + l1 < n; ++t) b = fma(v[t * stride] - mu, v[(t + l1) * stride] - mu, b);
    double g = a / (double)n + b / (double)n;                     /* acv[2j+1]+acv[2j+2] (1-based) */
    if (g <= 0) break;                                            /* m = j */
    if (j > 0 && g > gprev) g = gprev;                            /* monotone sequence */
    sumg = sumg + g;
    gprev = g;
  }
  const double mcvar = (-acv0 + 2 * sumg) / (double)n;
  out[1] = iidvar;
  out[2] = mcvar;
  out[3] = (double)n * iidvar / mcvar;                            /* src/stats/convergence/ess.jl:3 */
  out[4] = mcvar / iidvar;                                        /* src/stats/convergence/iact.jl:3 */
}

double orc_ess_series(const double* v, int64_t n, int64_t stride, double* iact_out) {
  double o[5];
  orc_stats_series(v, n, stride, o);
  if (iact_out) *iact_out = o[4];
  return o[3];
}

/* value: (nchains, npost, d); out: (5, nchains, d) in the order of orc_stats_series */
void orc_stats(const double* value, int64_t nchains, int64_t npost, int64_t d, double* out, int nthreads) {
  (void)nthreads;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t c = 0; c < nchains; ++c)
    for (int64_t i = 0; i < d; ++i) {
      double o[5];
      orc_stats_series(value + c * npost * d + i, npost, d, o);
      for (int q = 0; q < 5; ++q) out[((int64_t)q * nchains + c) * d + i] = o[q];
    }
}

/* acceptance(v::AbstractArray{Bool}) = mean(v) (diag != 0), or the change-count form of acceptance.jl:3-14 over the
 * saved states of each chain (value (nchains, npost, d)); out: nchains */
void orc_acceptance(const unsigned char* accept, const double* value, int64_t nchains, int64_t npost, int64_t d,
                    double* out) {
  for (int64_t c = 0; c < nchains; ++c) {
    int64_t cnt = 0;
    if (accept) {
      for (int64_t t = 0; t < npost; ++t) cnt += accept[c * npost + t] ? 1 : 0;
    } else {
      cnt = npost > 0 ? 1 : 0;
      for (int64_t t = 1; t < npost; ++t) {
        int diff = 0;
        for (int64_t i = 0; i < d && !diff; ++i)
          diff = value[(c * npost + t) * d + i] != value[(c * npost + t - 1) * d + i];
        cnt += diff;
      }
    }
    out[c] = (double)cnt / (double)npost;
  }
}

/* value: (nchains, npost, d) as stored by orc_run; ess: (nchains, d) */
void orc_ess(const double* value, int64_t nchains, int64_t npost, int64_t d, double* ess, int nthreads) {
  (void)nthreads;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t c = 0; c < nchains; ++c)
    for (int64_t i = 0; i < d; ++i) ess[c * d + i] = orc_ess_series(value + c * npost * d + i, npost, d, NULL);
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
