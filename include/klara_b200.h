/* klara_b200.h -- C ABI of libklara_b200.so: the B200-native replacement for the MCMC
 * inner loop of JuliaStats/Klara.jl (one BasicContMuvParameter, MH / MALA / HMC, batched
 * over independent chains).
 *
 * The reference has no FFI on this path; its extension points are Julia closures and
 * multiple dispatch.  Each entry point below therefore names the reference routine(s)
 * whose work it takes over (paths relative to the Klara.jl checkout, commit ffa4f6d0);
 * INTEGRATION.md shows the `ccall` stubs a Klara maintainer would add, and
 * klara.jl_b200/ the Python mirror of Klara's user surface used by the tests.
 *
 * Conventions
 *   - every call returns KLB_OK (0) or a negative KLB_E* code and never throws; the
 *     message of the last failure on the calling thread is klb_last_error()
 *   - host buffers are caller-owned and only touched during the call
 *   - device memory is library-owned, freed by klb_job_destroy
 *   - a job handle is not thread-safe (like the reference's mutable BasicMCJob);
 *     distinct handles are independent
 *   - matrices are Julia-style column-major: the state is `dim x nchains` (one chain =
 *     one contiguous column), monitored values are `dim x npoststeps x nchains`
 *     (BasicContMuvParameterNState.value is `size x n`,
 *      src/nstates/ParameterNStates/BasicContMuvParameterNState.jl:1-21, one per chain)
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     KLB_ECUDA
 */
#ifndef KLARA_B200_H
#define KLARA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KLB_VERSION 100 /* 0.1.0 */

/* error codes */
#define KLB_OK 0
#define KLB_EINVAL (-1)     /* invalid configuration / argument (the reference's @assert failures) */
#define KLB_ECUDA (-2)      /* CUDA runtime error, or no device */
#define KLB_ENOTFINITE (-3) /* initial log-target / gradient not finite (src/samplers/HMC.jl:113-114) */
#define KLB_ESTATE (-4)     /* call out of order (e.g. run before set_state, output overrun) */
#define KLB_EUNSUPPORTED (-5)
#define KLB_ENOMEM (-6)

/* sampler: src/samplers/MH.jl:47-66 (symmetric normal random walk), MALA.jl:61-70, HMC.jl:89-100 */
#define KLB_SAMPLER_MH 0
#define KLB_SAMPLER_MALA 1
#define KLB_SAMPLER_HMC 2
#define KLB_SAMPLER_NUTS 3   /* NUTS(leapstep; maxδ, maxndoublings), src/samplers/NUTS.jl:228-241, iterate/NUTS.jl:230-457: the
                                 multivariate transition as the reference computes it; elementwise
                                 targets; VanillaMCTuner or DualAveragingMCTuner */

/* target descriptors (device-resident replacements of the logtarget / gradlogtarget closures of
 * BasicContMuvParameter, src/variables/parameters/BasicContMuvParameter.jl:174-201) */
#define KLB_TARGET_ISO 0        /* -z.z, -2z                        README.md:153-155 */
#define KLB_TARGET_SHIFTED_ISO 1 /* -(z-mu).(z-mu), -2(z-mu)        test/BasicContMuvParameter.jl:539-563 */
#define KLB_TARGET_DENSE 2      /* -z'Cz, -2Cz   doc/examples/BivariateNormal/MALA/function/analytical.jl:8-9 (dim <= 512) */
#define KLB_TARGET_ROSENBROCK 3 /* -scale*sum_k [b(x_2k+1 - x_2k^2)^2 + (a - x_2k)^2]  (this repo) */
#define KLB_TARGET_LOGIT 4      /* Bayesian logistic regression, N(0, lambda I) prior, hyper-parameters [lambda, X, y]:
                                   dot(Xp,y) - sum(log.(1+exp.(Xp))) - 0.5*(dot(p,p)/lambda + d*log(2*pi*lambda)),
                                   X'*(y - 1./(1+exp.(-X*p))) - p/lambda
                                   doc/examples/swiss/HMC/noadaptation/analytical.jl:11-20 (dim <= 16) */

/* tuner: src/tuners/VanillaMCTuner.jl:6-16, src/tuners/AcceptanceRateMCTuner.jl:25-46 */
#define KLB_TUNER_VANILLA 0
#define KLB_TUNER_ACCEPTANCE_RATE 1
/* DualAveragingMCTuner(targetrate, nadapt; ε0bar, h0bar, γ, t0, κ, period, verbose) for HMC:
 * src/tuners/DualAveragingMCTuner.jl:53-101, src/samplers/HMC.jl:124-133,192-223, iterate/HMC.jl:125-127,142-144,225-248.
 * Per-chain step AND per-chain number of leapfrog steps, nleaps = max(1, round(λ/step)), with every target (the
 * dense-precision tile kernel runs the longest trajectory of a tile and masks the chains that have finished theirs). */
#define KLB_TUNER_DUAL_AVERAGING 2
#define KLB_SCORE_LOGISTIC 0
#define KLB_SCORE_ERF 1

/* arithmetic: 0 = every product and sum rounded separately, in the reference's evaluation order
 * ((0.5*step)*g first, then the addition: src/samplers/samplers.jl:130-133); 1 = a*b+c contracted
 * to fma (faster, differs from the reference by <= 1 ulp per operation) */
#define KLB_ARITH_REFERENCE 0
#define KLB_ARITH_FMA 1

/* monitor / diagnostics bit masks (outopts[:monitor], outopts[:diagnostics], src/jobs/jobs.jl:9-43) */
#define KLB_MONITOR_VALUE 1u
#define KLB_MONITOR_LOGTARGET 2u
#define KLB_MONITOR_GRADLOGTARGET 4u
#define KLB_DIAG_ACCEPT 1u
#define KLB_DIAG_NDOUBLINGS 2u /* NUTS: the :ndoublings diagnostic (src/samplers/NUTS.jl:285, iterate/NUTS.jl:389-391) */
#define KLB_DIAG_NUTS_A 4u     /* NUTS + DualAveragingMCTuner: :a  (src/samplers/NUTS.jl:317,344, iterate/NUTS.jl:393-395) */
#define KLB_DIAG_NUTS_NA 8u    /* NUTS + DualAveragingMCTuner: :na (iterate/NUTS.jl:397-399) */

/* outopts[:destination]: :nstate (device-resident buffer) or :none */
#define KLB_DEST_NSTATE 0
#define KLB_DEST_NONE 1

/* which = argument of klb_job_set_target_f64 */
#define KLB_PARAM_MU 0     /* dim doubles */
#define KLB_PARAM_C 1      /* dim*dim doubles, symmetric precision matrix */
#define KLB_PARAM_SIGMA 2  /* dim doubles: MH(sigma::Vector) proposal standard deviations (src/samplers/MH.jl:64) */
#define KLB_PARAM_ROSEN 3  /* 3 doubles: a, b, scale */
#define KLB_PARAM_LOGIT_X 4      /* ndata*dim doubles: design matrix Data(:X), row-major (row i = observation i) */
#define KLB_PARAM_LOGIT_Y 5      /* ndata doubles: outcomes Data(:y) */
#define KLB_PARAM_LOGIT_LAMBDA 6 /* 1 double: prior variance Hyperparameter(:lambda) (> 0) */

/* field = argument of klb_job_output / klb_job_device_ptr */
#define KLB_OUT_VALUE 0          /* double  dim x npost x nchains */
#define KLB_OUT_LOGTARGET 1      /* double  npost x nchains */
#define KLB_OUT_GRADLOGTARGET 2  /* double  dim x npost x nchains */
#define KLB_OUT_ACCEPT 3         /* uint8   npost x nchains */
#define KLB_OUT_STATE 4          /* double  dim x nchains      current pstate.value */
#define KLB_OUT_STATE_LOGTARGET 5 /* double nchains            current pstate.logtarget */
#define KLB_OUT_TUNE_STEP 6      /* double  nchains            sstate.tune.step */
#define KLB_OUT_TUNE_COUNTERS 7  /* int64   3 x nchains        accepted, proposed, totproposed */
#define KLB_OUT_TUNE_RATE 8      /* double  nchains            sstate.tune.rate (NaN after reset_burnin!) */
#define KLB_OUT_ESS 9            /* double  dim x nchains      filled by klb_job_ess (device_ptr only after that call) */
#define KLB_OUT_TUNE_RATES 11    /* double  nperiods x nchains  verbose tuners only: the acceptance rate of every burn-in period of
                                                               every chain -- what the reference prints per period (iterate/HMC.jl:
                                                               211-221, MALA.jl:138-148, MH.jl:126-139); nperiods = burnin / period
                                                               (nadapt / period for DualAveragingMCTuner); NaN = period not closed */
#define KLB_OUT_NDOUBLINGS 12    /* uint8   npost x nchains    NUTS: doublings of the stored transitions (KLB_DIAG_NDOUBLINGS) */
#define KLB_OUT_NUTS_A 13        /* double  npost x nchains    NUTS + dual averaging, :a: sum of min(1, exp(H' - H0)) over the leaves of
                                                               the last doubling of the stored transitions (KLB_DIAG_NUTS_A) */
#define KLB_OUT_NUTS_NA 14       /* int32   npost x nchains    ... and :na, the number of those leaves (KLB_DIAG_NUTS_NA); the example's
                                                               mean(diags[:a]./diags[:na]), doc/examples/swiss/NUTS/dualaveraging/analytical.jl:50 */
#define KLB_OUT_TUNE_DA 10       /* double  8 x nchains        DualAveragingMCTune: λ, μ, εbar, hbar, hweight, εweight,
                                                               nleaps of the last transition, sstate.count */

typedef struct klb_job klb_job; /* opaque, library-owned: one batched BasicMCJob */

/* Everything BasicMCJob(model, sampler, mcrange, v0; tuner, outopts) is built from
 * (src/jobs/BasicMCJob.jl:24-104), flattened. */
typedef struct {
  uint32_t struct_size;   /* = sizeof(klb_config), for ABI evolution */
  int32_t sampler;        /* KLB_SAMPLER_* */
  int32_t target;         /* KLB_TARGET_* */
  int32_t tuner;          /* KLB_TUNER_* */
  int32_t arith;          /* KLB_ARITH_* */
  int64_t nchains;        /* chains in THIS job (= this rank's shard) */
  int64_t dim;
  int64_t nsteps, burnin, thinning; /* BasicMCRange, src/ranges/BasicMCRange.jl:7-33 */
  double step;            /* HMC leapstep / MALA driftstep (> 0); unused by MH */
  int32_t nleaps;         /* HMC (> 0) */
  double target_rate;     /* AcceptanceRateMCTuner.targetrate in (0,1) */
  double score_k;         /* steepness k of the score function (defaults: 7 logistic, 3 erf) */
  int64_t period;         /* tuner period (> 0, default 100) */
  int32_t verbose;        /* tuner.verbose: switches the acceptance counters on (iterate/HMC.jl:129-133) */
  uint32_t monitor;       /* KLB_MONITOR_* */
  uint32_t diagnostics;   /* KLB_DIAG_* */
  int32_t destination;    /* KLB_DEST_* */
  uint64_t seed;          /* Philox key */
  int64_t chain_offset;   /* global index of this job's first chain (RNG streams use global indices,
                             so results do not depend on how chains are sharded over GPUs) */
  int32_t device;         /* CUDA device ordinal */
  int32_t score;          /* AcceptanceRateMCTuner.score: KLB_SCORE_LOGISTIC (logistic_rate_score, 2/(1+exp(-k x))) or
                             KLB_SCORE_ERF (erf_rate_score, erf(k x)+1); k = score_k   AcceptanceRateMCTuner.jl:9,17 */
  /* DualAveragingMCTuner only (target_rate, period, verbose above are shared with the other tuners) */
  int64_t da_nadapt;      /* nadapt > 0: transitions during which the step adapts */
  int64_t da_t0;          /* t0 > 0 (default 10) */
  double da_eps0bar;      /* ε0bar > 0 (default 1) */
  double da_h0bar;        /* h0bar (default 0) */
  double da_gamma;        /* γ (default 0.05) */
  double da_kappa;        /* κ (default 0.75) */
  /* NUTS only (step above = leapstep) */
  int32_t nuts_maxdelta;       /* maxδ > 0 (default 1000) */
  int32_t nuts_maxndoublings;  /* maxndoublings in 1..10 (default 5) */
} klb_config;

/* geometry the library chose for a job (needed by the oracle to reproduce the reduction order) */
typedef struct {
  int32_t nv;               /* double2 units per lane of the canonical reduction order: lane l owns
                               elements 2(l+32m), 2(l+32m)+1, m < nv (DESIGN.md "reduction order");
                               0 = sequential order (thread-per-chain kernels of KLB_TARGET_LOGIT) */
  int32_t warps_per_block;
  int32_t warps_per_chain;  /* W: a chain is owned by W warps; warp w holds the units m = w (mod W) */
  int32_t regs_per_thread;
  int32_t blocks_per_sm;
  int64_t ld;               /* leading dimension of the device-resident state / value / gradient columns
                               (dim rounded up to even, so every column is 16-byte aligned); host copies
                               made by klb_job_output are dense (leading dimension = dim) */
  int64_t npoststeps;       /* length((burnin+1):thinning:nsteps) */
  int64_t transitions_done; /* global transition counter t (RNG counter word) */
  int64_t saved;            /* job.count: samples stored since the last reset */
} klb_plan;

int klb_version(void);
const char* klb_last_error(void);
int klb_device_count(void);

/* BasicMCJob constructor minus initialize! (src/jobs/BasicMCJob.jl:24-104): validates the
 * configuration exactly like the reference's @asserts (HMC.jl:93-96, MALA.jl:64-67,
 * tuners.jl:12-20, VanillaMCTuner.jl:10-13, AcceptanceRateMCTuner.jl:31-35, BasicMCRange.jl:19-21),
 * allocates device state / tuner records / output (initialize_output, src/jobs/jobs.jl:188-210). */
int klb_job_create(const klb_config* cfg, klb_job** out);

/* Target / proposal parameters: what reaches the closures through parameter.states
 * (hyper-parameters, BasicContMuvParameter.jl:497-501) or through MH(sigma) (MH.jl:64). */
int klb_job_set_target_f64(klb_job* job, int which, const double* host, int64_t n);

/* initialize! (HMC.jl:106-120, MALA.jl:76-90, MH.jl:72-85) and reset(job, x) (BasicMCJob.jl:198-201):
 * upload x0 (dim x nchains), evaluate log-target (+ gradient) of every chain, fail with
 * KLB_ENOTFINITE naming the first offending chain; also resets the tuner records and the
 * output cursor like reset(job). */
int klb_job_set_state(klb_job* job, const double* x0);
/* same, x0 already in device memory of the job's device (no host copy) */
int klb_job_set_state_device(klb_job* job, const double* x0_dev);

/* The synthetic initial value of the benchmark configurations, generated on the device and then initialised like
 * klb_job_set_state:  x0[i, c] = N(0,1) of the Philox stream (cfg.seed, global chain chain_offset + c, transition 0,
 * element i)  (SURVEY.md section 8d).  A function of the global chain index only, so a job's input does not depend
 * on how its chains are sharded over GPUs.  (The reference's examples start from user vectors, README.md:47.) */
int klb_job_set_state_synthetic(klb_job* job);

/* Position the job's RNG streams: the next transition will be number t + 1.  The reference's unseeded global RNG
 * (iterate/HMC.jl:135,165) cannot seek; counter-based streams can, which is what makes a run reproducible from any
 * point and lets shards of one logical job be recomputed independently. */
int klb_job_seek(klb_job* job, uint64_t t);

/* run(job) (BasicMCJob.jl:212-244): all nsteps transitions of all chains, burn-in tuning,
 * thinning, saving.  Blocking. */
int klb_job_run(klb_job* job);
/* launch only (asynchronous on the job's stream); klb_job_sync waits */
int klb_job_run_async(klb_job* job);
int klb_job_sync(klb_job* job);

/* Transitions per kernel launch (0 = the whole run in one launch, the default).  nt = 1 is the
 * one-launch-per-iterate! lockstep schedule; results do not depend on the chunking. */
int klb_job_set_chunk(klb_job* job, int64_t nt);

/* reset(job) (BasicMCJob.jl:187-196): tuner records <- (sampler step, 0, 0, period, NaN), count <- 0.
 * The chain state is kept (pstate persists), the RNG counter keeps advancing.
 * DualAveragingMCTuner: reset!(tune, ::HMC, tuner) sets step = 1 (HMC.jl:218) and re-runs initialize_step!, which
 * in the reference throws (undefined `moment`, samplers.jl:195) as soon as the job has made one transition (the
 * proposal state's log-target is then no longer NaN); so klb_job_reset / klb_job_set_state / klb_job_run_host(x0)
 * of a dual-averaging job that has run return KLB_EUNSUPPORTED -- create a new job.  Before the first transition
 * they do what the reference does: step = 1, μ = log(10). */
int klb_job_reset(klb_job* job);

/* output(job) (BasicMCJob.jl:279) and job.pstate / job.sstate.tune: copy one field to host. */
int klb_job_output(klb_job* job, int field, void* host_dst, int64_t nbytes);
/* device address and byte size of a field (for zero-copy consumers: NCCL all-gather, torch views) */
int klb_job_device_ptr(klb_job* job, int field, void** dev_ptr, int64_t* nbytes);

/* run(job) from a host initial value to host results in ONE call:
 *     klb_job_set_state(job, x0); klb_job_run(job); klb_job_output(job, field_i, dst_i, nbytes_i) for every i
 * with identical results, but the chains are processed in `nslices` contiguous slices, each on its own CUDA
 * stream, so that the host->device copy of slice s+1, the kernels of slice s and the device->host copies of
 * slice s-1 overlap (the chains are independent: src/jobs/jobs.jl:212).  x0 may be NULL: reset(job) + run(job)
 * from the current state.  nslices <= 0 lets the library choose (about 32 MiB of state per slice, at most 16).  Host buffers
 * should be pinned (klb_host_alloc) for the copies to overlap; pageable memory works but serialises.
 * Fails with KLB_ENOTFINITE like klb_job_set_state (the run of the offending job is then discarded). */
typedef struct {
  int32_t field;     /* KLB_OUT_* (not KLB_OUT_ESS) */
  int32_t reserved;
  void* host_dst;
  int64_t nbytes;    /* full size of the field, as for klb_job_output */
} klb_host_field;
int klb_job_run_host(klb_job* job, const double* x0, const klb_host_field* fields, int32_t nfields, int32_t nslices);
/* The same call in two halves, so that several jobs (devices) can be in flight at once: _async enqueues the whole
 * pipeline and returns; the host buffers must stay valid until _finish, which waits, reports KLB_ENOTFINITE and commits
 * (or rolls back) the job's bookkeeping.  _wait waits and reports the same status but leaves the run pending, _abort
 * discards it: together they let a caller commit several jobs all-or-nothing (klb_multi_run_host).  Other calls on the
 * job are refused while a run is pending. */
int klb_job_run_host_async(klb_job* job, const double* x0, const klb_host_field* fields, int32_t nfields, int32_t nslices);
int klb_job_run_host_finish(klb_job* job);
int klb_job_run_host_wait(klb_job* job);
int klb_job_run_host_abort(klb_job* job);

/* ess(output(job)) = ess(chain, :imse) for every coordinate of every chain (src/stats/convergence/ess.jl:3-14,
 * src/stats/variance/mcvar.jl:5,75-105), computed on the device over the monitored values; host_ess
 * (dim x nchains doubles) may be NULL to leave the result on the device (KLB_OUT_ESS). */
int klb_job_ess(klb_job* job, double* host_ess);

/* The other post-hoc estimators of the monitored output, computed on the device by the same pass:
 *   KLB_STAT_MEAN        mean(s)                    src/stats/mean.jl:7-11
 *   KLB_STAT_MCVAR_IID   mcvar(s, Val{:iid})        src/stats/variance/mcvar.jl:5,15-16
 *   KLB_STAT_MCVAR_IMSE  mcvar(s, Val{:imse})       src/stats/variance/mcvar.jl:75-105
 *   KLB_STAT_ESS         ess(s)                     src/stats/convergence/ess.jl:3-14
 *   KLB_STAT_IACT        iact(s)                    src/stats/convergence/iact.jl:3-5
 * each dim x nchains doubles (coordinate fastest), and per chain (nchains doubles)
 *   KLB_STAT_ACCEPTANCE        acceptance(s)                     mean of the :accept diagnostic  src/stats/acceptance.jl:28-34
 *   KLB_STAT_ACCEPTANCE_VALUE  acceptance(s, diagnostics=false)  fraction of saved states that differ from the
 *                              previous saved state (first one counted)                          src/stats/acceptance.jl:3-14,33
 * host_dst may be NULL (result stays on the device).  Series shorter than 4 samples give NaN for all but the mean. */
#define KLB_STAT_MEAN 0
#define KLB_STAT_MCVAR_IID 1
#define KLB_STAT_MCVAR_IMSE 2
#define KLB_STAT_ESS 3
#define KLB_STAT_IACT 4
#define KLB_STAT_ACCEPTANCE 5
#define KLB_STAT_ACCEPTANCE_VALUE 6
int klb_job_stat(klb_job* job, int stat, double* host_dst);

int klb_job_plan(klb_job* job, klb_plan* out);
/* the configuration the job was created with */
int klb_job_config(klb_job* job, klb_config* out);
/* kernels launched by this job so far */
int64_t klb_job_launches(klb_job* job);
/* CUDA-event time (ms) of the last klb_job_run / run_async+sync kernel sequence */
double klb_job_last_run_ms(klb_job* job);
void* klb_job_stream(klb_job* job);

void klb_job_destroy(klb_job* job);

/* ---- multi-GPU ------------------------------------------------------------------------------------------------
 * run(job::Vector{MCJob}) = map(run, job) (src/jobs/jobs.jl:212) is the reference's whole multi-job facility: the
 * chains of one logical job are independent, so they shard over GPUs -- device g of G owns the contiguous chain block
 * [g N/G, (g+1) N/G), RNG streams use global chain indices (klb_config.chain_offset), results do not depend on G, and
 * nothing is exchanged while sampling.  One closing all-gather leaves, on EVERY device, the final state
 * (KLB_OUT_STATE, ld x N), log-target (KLB_OUT_STATE_LOGTARGET), tuner step (KLB_OUT_TUNE_STEP) and
 * {accepted, proposed, totproposed} (KLB_OUT_TUNE_COUNTERS) of ALL chains.  It is done by the copy engines over
 * NVLink / NVSwitch peer-to-peer (no kernel, no SM time), asynchronously to the job's stream.
 *
 * klb_multi_*: one process drives `ngpus` devices (the `ngpus` field of SURVEY.md section 8b).  cfg->nchains is the
 * TOTAL chain count; cfg->device is ignored; devices = NULL uses devices 0 .. ngpus-1 (ngpus <= 0: all of them); the
 * same ordinal may be listed twice.  Host arrays are those of the logical job (klb_job_* layouts with nchains = N). */
typedef struct klb_multi klb_multi;
int klb_multi_create(const klb_config* cfg, int32_t ngpus, const int32_t* devices, klb_multi** out);
int klb_multi_ngpus(klb_multi* m);
int klb_multi_job(klb_multi* m, int32_t g, klb_job** job);              /* the shard's job (klb_job_ess, klb_job_stat, ...) */
int klb_multi_set_target_f64(klb_multi* m, int which, const double* host, int64_t n);
int klb_multi_set_state(klb_multi* m, const double* x0);               /* dim x N */
int klb_multi_set_state_synthetic(klb_multi* m);
int klb_multi_reset(klb_multi* m);
int klb_multi_seek(klb_multi* m, uint64_t t);
int klb_multi_run(klb_multi* m);                                        /* run on every device + closing all-gather; blocking */
int klb_multi_run_async(klb_multi* m);
int klb_multi_sync(klb_multi* m);
/* klb_job_run_host for the logical job: x0 (dim x N) and the fields' host arrays are those of all N chains; every device
 * runs its shard's pipeline concurrently (klb_job_run_host_async on each, then _finish), then the closing all-gather */
int klb_multi_run_host(klb_multi* m, const double* x0, const klb_host_field* fields, int32_t nfields, int32_t nslices);
int klb_multi_output(klb_multi* m, int field, void* host_dst, int64_t nbytes);   /* output(job): shards concatenated */
int klb_multi_gathered(klb_multi* m, int32_t g, int field, void** dev_ptr, int64_t* nbytes);  /* device g's copy of the all-gather */
int klb_multi_gathered_output(klb_multi* m, int32_t g, int field, void* host_dst, int64_t nbytes);
double klb_multi_last_run_ms(klb_multi* m);                             /* max over devices */
void klb_multi_destroy(klb_multi* m);

/* klb_gather_*: one end of the same all-gather per PROCESS (one process per GPU, e.g. torchrun).  Every rank creates
 * its end on its shard's job, exports a KLB_GATHER_HANDLE_BYTES handle (a CUDA IPC memory handle plus the shard's
 * position), the caller all-gathers the handles with whatever it has (torch.distributed, MPI, a file) and connects.
 * klb_gather_push_async copies this rank's shard into every rank's buffers after the work queued on the job's stream;
 * klb_gather_sync waits for this rank's copies.  Once every rank's klb_gather_sync has returned (the inter-process
 * barrier is the caller's), every rank holds the complete all-gather.  first_chain = global index of chain 0 of the
 * logical job. */
#define KLB_GATHER_HANDLE_BYTES 128
typedef struct klb_gather klb_gather;
int klb_gather_create(klb_job* job, int32_t world, int32_t rank, int64_t nchains_total, int64_t first_chain, klb_gather** out);
int klb_gather_handle(klb_gather* g, void* handle_out);
int klb_gather_connect(klb_gather* g, const void* handles /* world x KLB_GATHER_HANDLE_BYTES, rank order */);
int klb_gather_push_async(klb_gather* g);
int klb_gather_sync(klb_gather* g);
int klb_gather_join(klb_gather* g);   /* the job's stream waits for the last push (for device-side timing) */
int klb_gather_device_ptr(klb_gather* g, int field, void** dev_ptr, int64_t* nbytes);
int klb_gather_output(klb_gather* g, int field, void* host_dst, int64_t nbytes);
int klb_gather_disconnect(klb_gather* g);   /* unmap the peers' buffers; then (after the caller's barrier) klb_gather_destroy */
void klb_gather_destroy(klb_gather* g);

/* Measured throughput of one device for the roofline denominators that MEASURED_PEAKS.json does not hold:
 * KLB_PEAK_FP64 = fp64 results per second of a stream of independent DFMA (DADD / DMUL issue at the same rate),
 * KLB_PEAK_DMMA = flop per second of a stream of independent mma.sync.m8n8k4.f64 (the fp64 tensor pipe). */
#define KLB_PEAK_FP64 0
#define KLB_PEAK_DMMA 1
int klb_device_peak(int device, int kind, double* per_second);

/* pinned host memory for end-to-end pipelines */
int klb_host_alloc(void** p, int64_t nbytes);
int klb_host_free(void* p);

/* Device self-tests used by the parity suite (run on the GPU, results copied to host):
 * n standard normals of stream (seed, chain, t) / op(x[i]) with op 0 = exp, 1 = log, 2 = erf, 3 = the shared-divisor
 * division of the MALA kernels on pairs (x[2k] / x[2k+1], written to both slots) / the accept uniform. */
int klb_debug_normals(int device, uint64_t seed, uint64_t chain, uint64_t t, int64_t n, double* host_out);
int klb_debug_math(int device, int op, int64_t n, const double* host_in, double* host_out);
int klb_debug_uniform(int device, uint64_t seed, uint64_t chain, uint64_t t, double* host_out);

#ifdef __cplusplus
}
#endif
#endif /* KLARA_B200_H */
