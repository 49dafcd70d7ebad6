// klb_api.cu -- the C ABI of libklara_b200.so (include/klara_b200.h): job objects, validation,
// device memory, launches.  No CPU compute path exists here: without a CUDA device every
// compute entry point fails with KLB_ECUDA.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>

#include "../../include/klara_b200.h"
#include "klb_kernels.cuh"
#include "klb_dense.cuh"
#include "klb_dense_mma.cuh"
#include "klb_nuts.cuh"
#include "klb_glm.cuh"

// per-(sampler, arithmetic) dispatchers, klb_kernels_inst.cu
int klb_chain_0_0(const KArgs*, int, int, int, int, int*, int*, cudaStream_t);
int klb_chain_0_1(const KArgs*, int, int, int, int, int*, int*, cudaStream_t);
int klb_chain_1_0(const KArgs*, int, int, int, int, int*, int*, cudaStream_t);
int klb_chain_1_1(const KArgs*, int, int, int, int, int*, int*, cudaStream_t);
int klb_chain_2_0(const KArgs*, int, int, int, int, int*, int*, cudaStream_t);
int klb_chain_2_1(const KArgs*, int, int, int, int, int*, int*, cudaStream_t);
// warp-specialised HMC kernels, klb_hmc_ws_inst.cu
int klb_hmc_ws_0(const KArgs*, int target, int gw, int nv, int full, int* regs, int* bps, cudaStream_t);
int klb_hmc_ws_1(const KArgs*, int target, int gw, int nv, int full, int* regs, int* bps, cudaStream_t);
// klb_kernels_inst.cu (-DKLB_INST_INIT) / klb_aux.cu
int klb_launch_init(const KArgs& A, int target, int W, int NV, int fma, int check_grad, unsigned long long* flag,
                    cudaStream_t s);
void klb_launch_fill_tune(double* step, long long* cnt, double* rate, long long n, double step0, long long period,
                          cudaStream_t s);
void klb_launch_fill_da(double* da, long long n, double lambda, double mu, double epsbar, double hbar, cudaStream_t s);
void klb_launch_debug_normals(const uint64_t* tab, uint64_t seed, uint64_t chain, uint64_t t, long long n, double* out,
                              cudaStream_t s);
void klb_launch_debug_math(const uint64_t* tab, int op, long long n, const double* in, double* out, cudaStream_t s);
void klb_launch_debug_uniform(uint64_t seed, uint64_t chain, uint64_t t, double* out, cudaStream_t s);
void klb_launch_ess(const double* value, long long ld, long long npost, long long nchains, int dim, double* ess,
                    cudaStream_t s);
void klb_launch_stats(const double* value, long long ld, long long npost, long long nchains, int dim,
                      double* const stats[5], cudaStream_t s);
void klb_launch_acceptance(const unsigned char* accept, const double* value, long long ld, long long npost,
                           long long nchains, int dim, double* out, cudaStream_t s);

void klb_launch_synth_state(const uint64_t* tab, uint64_t seed, uint64_t chain_offset, long long nchains, int dim,
                            long long ld, double* state, cudaStream_t s);
int klb_measure_peak(int kind, double* per_second);

#define KLB_MAX_SLICES 16
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(e_ == cudaErrorMemoryAllocation ? KLB_ENOMEM : KLB_ECUDA, "%s: %s", #call, \
                  cudaGetErrorString(e_));                                                     \
  } while (0)

struct klb_job {
  klb_config cfg;
  int gw, gnv;      // team geometry: warps per chain, double2 units per thread
  int nv;           // = gw * gnv: units per lane of the canonical reduction order
  long long ld;     // even leading dimension of the state / value / grad columns
  long long npost;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  // device buffers
  double* state;
  double* lt;
  double* tune_step;
  long long* tune_cnt;
  double* tune_rate;
  double* tune_da;  // DualAveragingMCTuner: 8 doubles per chain
  double* tune_rates;   // verbose tuners: nperiods burn-in acceptance rates per chain (chain-major), else null
  long long nperiods;
  bool da, constructed;   // constructed: the first klb_job_set_state (= BasicMCJob constructor) has happened
  double* out_value;
  double* out_lt;
  double* out_grad;
  unsigned char* out_accept;
  unsigned char* out_ndoublings;   // NUTS, KLB_DIAG_NDOUBLINGS
  double* out_nuts_a;              // NUTS + DualAveragingMCTuner, KLB_DIAG_NUTS_A
  int* out_nuts_na;                // NUTS + DualAveragingMCTuner, KLB_DIAG_NUTS_NA
  double* mu;
  double* sigma;
  double* Cm;       // dense precision matrix (d x d), KLB_TARGET_DENSE only
  bool dense, have_C;
  bool dense_mma;   // HMC + dense + dim in {64,128,256,512}: matrix-vector products on the fp64 tensor pipe
  uint64_t* tab;
  unsigned long long* flag;
  double* ess;      // dim x nchains, allocated on first klb_job_ess
  double* stat[5];  // mean, mcvar(:iid), mcvar(:imse), ess, iact: dim x nchains each, allocated on first klb_job_stat
  double* accrate;  // nchains
  unsigned long long stat_epoch;  // t_global + 1 when stat[] was filled (0 = never)
  double rosen[3];
  // KLB_TARGET_LOGIT (klb_glm.cuh): one thread per chain, design matrix padded to row pitch gdp
  bool glm, have_X, have_y;
  int gdp;            // padded dimension of the kernel instance: 2, 4, 8 or 16
  double* gX;         // ndata x gdp
  double* gy;         // ndata
  long long ndata, ny;
  double lambda;
  bool have_mu, have_sigma, have_rosen, have_state;
  unsigned long long t_global;  // transitions done since creation (RNG counter)
  long long count;              // job.count
  long long launches;
  long long chunk;              // transitions per launch (0 = whole run)
  int regs, bps;
  bool timed;
  // klb_job_run_host: chain slices pipelined over their own streams (H2D | kernels | D2H overlap)
  // a pipelined run enqueued by klb_job_run_host_async and not yet finished
  bool host_pending, host_had_x0, host_was_constructed;
  unsigned long long* flag_host;   // pinned: the finiteness flag comes back asynchronously
  cudaStream_t sl_stream[KLB_MAX_SLICES];
  cudaEvent_t sl_done[KLB_MAX_SLICES];
  int nsl_streams;
};

// HMC with one warp per chain and 8 or 16 units per lane (dim 257..1024), or four warps per chain (dim 1025..4096), runs
// the warp-specialised kernel (klb_hmc_ws.cuh) unless KLB_HMC_WS=0 (experiments: the fused single-role kernel of
// klb_kernels.cuh; KLB_HMC_WS=1: only the one-warp geometries).
static bool use_hmc_ws(int sampler, int gw, int gnv) {
  const char* env = getenv("KLB_HMC_WS");
  if (sampler != KLB_SAMPLER_HMC || (env && env[0] == '0')) return false;
  if (gw == 4) return gnv == 16 && !(env && env[0] == '1');
  return gw == 1 && (gnv == 8 || gnv == 16);
}

// NUTS kernels, klb_nuts_inst.cu
int klb_nuts_0(const KArgs*, int target, int W, int NV, int* regs, int* bps, cudaStream_t);
int klb_nuts_1(const KArgs*, int target, int W, int NV, int* regs, int* bps, cudaStream_t);

static int chain_dispatch(int sampler, int fma, const KArgs* A, int target, int gw, int gnv, int full, int* regs, int* bps,
                          cudaStream_t s) {
  if (sampler == KLB_SAMPLER_NUTS)
    return fma ? klb_nuts_1(A, target, gw, gnv, regs, bps, s) : klb_nuts_0(A, target, gw, gnv, regs, bps, s);
  if (use_hmc_ws(sampler, gw, gnv))
    return fma ? klb_hmc_ws_1(A, target, gw, gnv, full, regs, bps, s) : klb_hmc_ws_0(A, target, gw, gnv, full, regs, bps, s);
  switch (sampler * 2 + (fma ? 1 : 0)) {
    case 0: return klb_chain_0_0(A, target, gw, gnv, full, regs, bps, s);
    case 1: return klb_chain_0_1(A, target, gw, gnv, full, regs, bps, s);
    case 2: return klb_chain_1_0(A, target, gw, gnv, full, regs, bps, s);
    case 3: return klb_chain_1_1(A, target, gw, gnv, full, regs, bps, s);
    case 4: return klb_chain_2_0(A, target, gw, gnv, full, regs, bps, s);
    case 5: return klb_chain_2_1(A, target, gw, gnv, full, regs, bps, s);
  }
  return -1;
}

// Team geometry by dim: (warps per chain, double2 units per thread), capacity 64*W*NV elements.
// KLB_GEOM="W,NV" overrides the default (experiments; must be an instantiated pair with enough capacity).
static int plan_geom(long long dim, int* gw, int* gnv) {
  static const int geoms[][2] = {{1, 1}, {1, 2}, {1, 4}, {1, 8}, {1, 16}, {4, 16}};
  const char* env = getenv("KLB_GEOM");
  if (env) {
    int w = 0, nv = 0;
    if (sscanf(env, "%d,%d", &w, &nv) == 2 && w > 0 && nv > 0 && 64ll * w * nv >= dim) {
      *gw = w; *gnv = nv;
      return 0;
    }
  }
  for (auto& g : geoms)
    if (64ll * g[0] * g[1] >= dim) { *gw = g[0]; *gnv = g[1]; return 0; }
  return -1;
}

static long long npoststeps(long long burnin, long long thinning, long long nsteps) {
  return nsteps <= burnin ? 0 : (nsteps - burnin - 1) / thinning + 1;
}

// for the other translation units of the library (klb_multi.cu)
int klb_set_error(int code, const char* msg) { return fail(code, "%s", msg); }

extern "C" {

int klb_version(void) { return KLB_VERSION; }
const char* klb_last_error(void) { return g_err; }

int klb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static void free_job(klb_job* j) {
  if (!j) return;
  cudaSetDevice(j->cfg.device);
  cudaFree(j->state); cudaFree(j->lt); cudaFree(j->tune_step); cudaFree(j->tune_cnt); cudaFree(j->tune_rate); cudaFree(j->tune_da);
  cudaFree(j->tune_rates);
  if (j->flag_host) cudaFreeHost(j->flag_host);
  cudaFree(j->out_value); cudaFree(j->out_lt); cudaFree(j->out_grad); cudaFree(j->out_accept); cudaFree(j->out_ndoublings);
  cudaFree(j->out_nuts_a); cudaFree(j->out_nuts_na);
  cudaFree(j->mu); cudaFree(j->sigma); cudaFree(j->Cm); cudaFree(j->tab); cudaFree(j->flag); cudaFree(j->ess); cudaFree(j->accrate);
  cudaFree(j->gX); cudaFree(j->gy);
  for (int q = 0; q < 5; ++q) if (q != KLB_STAT_ESS) cudaFree(j->stat[q]);
  for (int q = 0; q < j->nsl_streams; ++q) { cudaStreamDestroy(j->sl_stream[q]); cudaEventDestroy(j->sl_done[q]); }
  if (j->ev0) cudaEventDestroy(j->ev0);
  if (j->ev1) cudaEventDestroy(j->ev1);
  if (j->stream) cudaStreamDestroy(j->stream);
  delete j;
}

// wait for everything the job has in flight: its own stream and the slice streams of klb_job_run_host
static int sync_all(klb_job* j) {
  CK(cudaSetDevice(j->cfg.device));
  cudaError_t first = cudaStreamSynchronize(j->stream);
  for (int q = 0; q < j->nsl_streams; ++q) {
    const cudaError_t e = cudaStreamSynchronize(j->sl_stream[q]);
    if (first == cudaSuccess) first = e;
  }
  CK(first);
  return KLB_OK;
}

static void fill_args(const klb_job* j, KArgs& A) {
  const klb_config& c = j->cfg;
  memset(&A, 0, sizeof A);
  A.state = j->state; A.lt = j->lt;
  A.tune_step = j->tune_step; A.tune_cnt = j->tune_cnt; A.tune_rate = j->tune_rate;
  A.out_value = j->out_value; A.out_lt = j->out_lt; A.out_grad = j->out_grad; A.out_accept = j->out_accept;
  A.out_nuts_a = j->out_nuts_a; A.out_nuts_na = j->out_nuts_na;
  A.out_ndoublings = j->out_ndoublings; A.nuts_maxdelta = c.nuts_maxdelta; A.nuts_maxndoublings = c.nuts_maxndoublings;
  A.mu = j->mu; A.sigma = j->sigma; A.tab = j->tab;
  A.ra = j->rosen[0]; A.rb = j->rosen[1]; A.rscale = j->rosen[2];
  A.nchains = c.nchains; A.dim = c.dim; A.ld = j->ld;
  A.burnin = c.burnin; A.thinning = c.thinning; A.npost = j->npost; A.period = c.period;
  A.nleaps = c.nleaps; A.tuner = c.tuner;
  // counters advance for AcceptanceRateMCTuner or a verbose tuner (iterate/HMC.jl:129-133);
  // for MH only when verbose (iterate/MH.jl:73-75)
  // and for NUTS (iterate/NUTS.jl:238-240)
  A.counters_on = (c.sampler == KLB_SAMPLER_MH || c.sampler == KLB_SAMPLER_NUTS)
                      ? (c.verbose != 0) : ((c.tuner == KLB_TUNER_ACCEPTANCE_RATE) || c.verbose != 0);
  A.target_rate = c.target_rate; A.score_k = c.score_k; A.score = c.score;
  A.seed = c.seed; A.chain_offset = (unsigned long long)c.chain_offset;
  A.out_rate = j->tune_rates; A.nperiods = j->nperiods;
  A.tune_da = j->tune_da; A.da_nadapt = c.da_nadapt; A.da_t0 = c.da_t0; A.da_gamma = c.da_gamma; A.da_kappa = c.da_kappa;
}

// dynamic shared memory of the thread-per-chain kernels: X and y are staged when they fit
static size_t glm_smem(const klb_job* j) {
  const size_t need = (size_t)j->ndata * (size_t)(j->gdp + 1) * sizeof(double);
  return need <= 160 * 1024 ? need : 0;
}
static void fill_glm_args(const klb_job* j, const KArgs& A, GArgs& G) {
  G.k = A;
  G.X = j->gX; G.y = j->gy; G.ndata = j->ndata; G.lambda = j->lambda;
  G.logc = klb_log((2 * 3.141592653589793) * j->lambda, KLB_TAB);     // log(2*pi*v[1]), same klb_log as the device
  G.data_in_smem = glm_smem(j) > 0;
}

int klb_job_create(const klb_config* cfg, klb_job** out) {
  if (!cfg || !out) return fail(KLB_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(klb_config))
    return fail(KLB_EINVAL, "klb_config.struct_size = %u, library expects %zu", cfg->struct_size, sizeof(klb_config));
  const klb_config& c = *cfg;
  if (c.sampler < 0 || c.sampler > 3) return fail(KLB_EINVAL, "unknown sampler %d", c.sampler);
  if (c.target != KLB_TARGET_ISO && c.target != KLB_TARGET_SHIFTED_ISO && c.target != KLB_TARGET_ROSENBROCK &&
      c.target != KLB_TARGET_DENSE && c.target != KLB_TARGET_LOGIT)
    return fail(KLB_EINVAL, "unknown target %d", c.target);
  if (c.target == KLB_TARGET_LOGIT && c.dim > KLB_GLM_MAXD)
    return fail(KLB_EUNSUPPORTED, "the logistic-regression kernels hold one chain per thread: dim <= %d", KLB_GLM_MAXD);
  if (c.target == KLB_TARGET_DENSE && c.dim > KLB_DENSE_MAXD)
    return fail(KLB_EUNSUPPORTED, "the dense-precision kernels need dim <= %d", KLB_DENSE_MAXD);
  if (c.tuner != KLB_TUNER_VANILLA && c.tuner != KLB_TUNER_ACCEPTANCE_RATE && c.tuner != KLB_TUNER_DUAL_AVERAGING)
    return fail(KLB_EINVAL, "unknown tuner %d", c.tuner);
  if (c.tuner == KLB_TUNER_DUAL_AVERAGING) {
    if (c.sampler != KLB_SAMPLER_HMC && c.sampler != KLB_SAMPLER_NUTS)
      return fail(KLB_EINVAL, "DualAveragingMCTuner tunes HMC (and NUTS) only; MALA / MH have no tuner_state method for it");
    // DualAveragingMCTuner asserts (src/tuners/DualAveragingMCTuner.jl:76-80)
    if (!(c.target_rate > 0 && c.target_rate < 1)) return fail(KLB_EINVAL, "Target acceptance rate should be between 0 and 1");
    if (c.da_nadapt <= 0) return fail(KLB_EINVAL, "Number of adaptation steps should be positive");
    if (!(c.da_eps0bar > 0)) return fail(KLB_EINVAL, "ε0bar should be positive");
    if (c.da_t0 <= 0) return fail(KLB_EINVAL, "t0 should be positive");
  }
  if (c.nsteps >= 0x7fffffffll) return fail(KLB_EINVAL, "nsteps must be below 2^31");
  if (c.arith != KLB_ARITH_REFERENCE && c.arith != KLB_ARITH_FMA) return fail(KLB_EINVAL, "unknown arith %d", c.arith);
  if (c.nchains <= 0) return fail(KLB_EINVAL, "nchains must be positive");
  if (c.dim <= 0) return fail(KLB_EINVAL, "dim must be positive");
  if (c.target == KLB_TARGET_ROSENBROCK && (c.dim & 1)) return fail(KLB_EINVAL, "paired Rosenbrock needs an even dim");
  int gw = 0, gnv = 0;
  if (plan_geom(c.dim, &gw, &gnv) != 0)
    return fail(KLB_EUNSUPPORTED, "dim %lld > 4096 is not supported by the register-resident chain kernels", (long long)c.dim);
  const int nv = c.target == KLB_TARGET_LOGIT ? 0 : gw * gnv;   // 0: sequential reduction order (klb_glm.cuh)
  // BasicMCRange asserts (src/ranges/BasicMCRange.jl:19-21)
  if (c.burnin < 0) return fail(KLB_EINVAL, "Number of burn-in iterations should be non-negative");
  if (c.thinning < 1) return fail(KLB_EINVAL, "Thinning should be >= 1");
  if (c.nsteps <= c.burnin)
    return fail(KLB_EINVAL, "Total number of MCMC iterations should be greater than number of burn-in iterations");
  // sampler asserts (HMC.jl:93-96, MALA.jl:64-67)
  if (c.sampler == KLB_SAMPLER_HMC) {
    if (!(c.step > 0)) return fail(KLB_EINVAL, "Leapfrog step is not positive");
    if (c.nleaps <= 0) return fail(KLB_EINVAL, "Number of leapfrog steps is not positive");
  }
  if (c.sampler == KLB_SAMPLER_MALA && !(c.step > 0)) return fail(KLB_EINVAL, "Drift step is not positive");
  if (c.sampler == KLB_SAMPLER_NUTS) {                                 // NUTS.jl:233-237
    if (!(c.step > 0)) return fail(KLB_EINVAL, "Leapfrog step is not positive");
    if (c.nuts_maxdelta <= 0) return fail(KLB_EINVAL, "maxδ is not positive");
    if (c.nuts_maxndoublings <= 0) return fail(KLB_EINVAL, "Maximum number of doublings is not positive");
    if (c.nuts_maxndoublings > KLB_NUTS_MAXLEVELS)
      return fail(KLB_EUNSUPPORTED, "maxndoublings <= %d (2^maxndoublings - 1 leapfrog steps per transition)", KLB_NUTS_MAXLEVELS);
    if (c.tuner == KLB_TUNER_ACCEPTANCE_RATE)
      return fail(KLB_EINVAL, "NUTS has sampler states for VanillaMCTuner and DualAveragingMCTuner only (src/samplers/NUTS.jl:271-345)");
    if (c.target == KLB_TARGET_DENSE)
      return fail(KLB_EUNSUPPORTED, "the NUTS kernels cover the elementwise targets (iso, shifted iso, Rosenbrock) and the "
                                    "logistic-regression target, not the dense-precision one");
  }
  // tuner asserts (VanillaMCTuner.jl:10-13, AcceptanceRateMCTuner.jl:31-35)
  if (c.period <= 0) return fail(KLB_EINVAL, "Adaptation period should be positive");
  if (c.tuner == KLB_TUNER_ACCEPTANCE_RATE && !(c.target_rate > 0 && c.target_rate < 1))
    return fail(KLB_EINVAL, "Target acceptance rate should be between 0 and 1");
  if (c.tuner == KLB_TUNER_ACCEPTANCE_RATE && c.score != KLB_SCORE_LOGISTIC && c.score != KLB_SCORE_ERF)
    return fail(KLB_EINVAL, "unknown score function %d", c.score);
  if ((c.monitor & KLB_MONITOR_GRADLOGTARGET) && c.sampler == KLB_SAMPLER_MH)
    return fail(KLB_EINVAL, "MH does not evaluate gradlogtarget; it cannot be monitored");
  if (c.monitor & ~7u) return fail(KLB_EINVAL, "unknown monitor bits");
  if (c.diagnostics & ~15u) return fail(KLB_EINVAL, "unknown diagnostics bits");
  if ((c.diagnostics & (KLB_DIAG_NUTS_A | KLB_DIAG_NUTS_NA)) &&
      !(c.sampler == KLB_SAMPLER_NUTS && c.tuner == KLB_TUNER_DUAL_AVERAGING))
    return fail(KLB_EINVAL, ":a and :na are diagnostics of NUTS with DualAveragingMCTuner (src/samplers/NUTS.jl:317,344)");
  if ((c.diagnostics & KLB_DIAG_NDOUBLINGS) && c.sampler != KLB_SAMPLER_NUTS)
    return fail(KLB_EINVAL, ":ndoublings is a diagnostic of NUTS");
  if (c.destination != KLB_DEST_NSTATE && c.destination != KLB_DEST_NONE) return fail(KLB_EINVAL, "unknown destination");
  if (c.chain_offset < 0 || (unsigned long long)c.chain_offset + (unsigned long long)c.nchains > 0xFFFFFFFFull)
    return fail(KLB_EINVAL, "global chain indices must fit 32 bits");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(KLB_ECUDA, "no CUDA device available (this library has no CPU path)");
  }
  if (c.device < 0 || c.device >= ndev) return fail(KLB_EINVAL, "device %d out of range (%d devices)", c.device, ndev);
  CK(cudaSetDevice(c.device));

  klb_job* j = new (std::nothrow) klb_job();
  if (!j) return fail(KLB_ENOMEM, "host allocation failed");
  memset(j, 0, sizeof *j);
  j->cfg = c;
  j->nv = nv; j->gw = gw; j->gnv = gnv;
  j->ld = (c.dim + 1) & ~1ll;
  j->npost = npoststeps(c.burnin, c.thinning, c.nsteps);
  j->rosen[0] = 1.0; j->rosen[1] = 100.0; j->rosen[2] = 0.05;
  j->have_rosen = true;
  const size_t N = (size_t)c.nchains, d = (size_t)j->ld, P = (size_t)j->npost, pad = nv ? 64 * (size_t)nv : KLB_GLM_MAXD;
#define CKJ(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      free_job(j);                                                                             \
      return fail(e_ == cudaErrorMemoryAllocation ? KLB_ENOMEM : KLB_ECUDA, "%s: %s", #call, \
                  cudaGetErrorString(e_));                                                     \
    }                                                                                          \
  } while (0)
  CKJ(cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking));
  CKJ(cudaEventCreate(&j->ev0));
  CKJ(cudaEventCreate(&j->ev1));
  CKJ(cudaMalloc(&j->state, N * d * sizeof(double)));
  CKJ(cudaMemset(j->state, 0, N * d * sizeof(double)));
  CKJ(cudaMalloc(&j->lt, N * sizeof(double)));
  CKJ(cudaMalloc(&j->tune_step, N * sizeof(double)));
  CKJ(cudaMalloc(&j->tune_cnt, 3 * N * sizeof(long long)));
  CKJ(cudaMalloc(&j->tune_rate, N * sizeof(double)));
  j->da = c.tuner == KLB_TUNER_DUAL_AVERAGING;
  if (j->da) CKJ(cudaMalloc(&j->tune_da, 8 * N * sizeof(double)));
  // verbose: one rate per chain and burn-in period (the reference prints them: iterate/HMC.jl:211-221, iterate/MH.jl:126-139);
  // dual averaging reports while count <= nadapt, every `period` proposals
  // (NUTS: the burn-in condition holds for both tuners, iterate/NUTS.jl:406-447)
  j->nperiods = c.verbose ? ((j->da && c.sampler != KLB_SAMPLER_NUTS) ? c.da_nadapt : c.burnin) / c.period : 0;
  if (j->nperiods > 0 && (size_t)j->nperiods * N * sizeof(double) <= ((size_t)1 << 32)) {
    CKJ(cudaMalloc(&j->tune_rates, (size_t)j->nperiods * N * sizeof(double)));
    CKJ(cudaMemset(j->tune_rates, 0xff, (size_t)j->nperiods * N * sizeof(double)));   // NaN until a period closes
  } else j->nperiods = 0;
  CKJ(cudaMalloc(&j->mu, pad * sizeof(double)));
  CKJ(cudaMalloc(&j->sigma, pad * sizeof(double)));
  CKJ(cudaMemset(j->mu, 0, pad * sizeof(double)));
  CKJ(cudaMemset(j->sigma, 0, pad * sizeof(double)));
  CKJ(cudaMalloc(&j->tab, sizeof(KLB_TAB)));
  CKJ(cudaMemcpy(j->tab, KLB_TAB, sizeof(KLB_TAB), cudaMemcpyHostToDevice));
  CKJ(cudaMalloc(&j->flag, sizeof(unsigned long long)));
  CKJ(cudaHostAlloc((void**)&j->flag_host, sizeof(unsigned long long), cudaHostAllocDefault));
  if (c.destination == KLB_DEST_NSTATE) {      // initialize_output, src/jobs/jobs.jl:188-210
    if (c.monitor & KLB_MONITOR_VALUE) CKJ(cudaMalloc(&j->out_value, N * P * d * sizeof(double)));
    if (j->out_value && (c.dim & 1)) CKJ(cudaMemset(j->out_value, 0, N * P * d * sizeof(double)));
    if (c.monitor & KLB_MONITOR_LOGTARGET) CKJ(cudaMalloc(&j->out_lt, N * P * sizeof(double)));
    if (c.monitor & KLB_MONITOR_GRADLOGTARGET) CKJ(cudaMalloc(&j->out_grad, N * P * d * sizeof(double)));
    if (c.diagnostics & KLB_DIAG_ACCEPT) CKJ(cudaMalloc(&j->out_accept, N * P));
    if (c.diagnostics & KLB_DIAG_NDOUBLINGS) CKJ(cudaMalloc(&j->out_ndoublings, N * P));
    if (c.diagnostics & KLB_DIAG_NUTS_A) CKJ(cudaMalloc(&j->out_nuts_a, N * P * sizeof(double)));
    if (c.diagnostics & KLB_DIAG_NUTS_NA) CKJ(cudaMalloc(&j->out_nuts_na, N * P * sizeof(int)));
  }
  j->dense = c.target == KLB_TARGET_DENSE;
  if (j->dense) {                        // rows padded to the even length of the state columns (zero pad column)
    CKJ(cudaMalloc(&j->Cm, (size_t)c.dim * d * sizeof(double)));
    CKJ(cudaMemset(j->Cm, 0, (size_t)c.dim * d * sizeof(double)));
  }
  j->glm = c.target == KLB_TARGET_LOGIT;
  j->lambda = 100.0;                               // v0 of the swiss examples; KLB_PARAM_LOGIT_LAMBDA overrides
  for (j->gdp = 2; j->gdp < c.dim; j->gdp *= 2) {}
  {
    const char* env = getenv("KLB_DENSE_MMA");     // KLB_DENSE_MMA=0 forces the DFMA register-tile kernel (experiments)
    // (DualAveragingMCTuner: chains of a tile run different numbers of leapfrog steps; the DFMA tile kernel masks them,
    // the cluster pipeline of the tensor-pipe kernel wants every CTA to run the same number of matrix products)
    j->dense_mma = j->dense && c.sampler == KLB_SAMPLER_HMC && !(env && env[0] == '0') &&
                   c.tuner != KLB_TUNER_DUAL_AVERAGING && (c.dim == 64 || c.dim == 128 || c.dim == 256 || c.dim == 512);
  }
  if (j->glm ? klb_glm_attrs(c.sampler, c.arith, j->gdp, 0, &j->regs, &j->bps) != 0
      : j->dense_mma ? klb_dense_mma_attrs(c.arith, (int)c.dim, &j->regs, &j->bps) != 0
      : j->dense ? klb_dense_attrs(c.sampler, c.arith, c.tuner == KLB_TUNER_DUAL_AVERAGING, (int)c.dim, &j->regs, &j->bps) != 0
               : chain_dispatch(c.sampler, c.arith, nullptr, c.target, gw, gnv, c.dim == 64ll * gw * gnv, &j->regs, &j->bps, j->stream) != 0) {
    cudaGetLastError();
    free_job(j);
    return fail(KLB_ECUDA, "no sm_100a kernel image for geometry (%d,%d) loadable on this device", gw, gnv);
  }
#undef CKJ
  *out = j;
  return KLB_OK;
}

void klb_job_destroy(klb_job* job) { free_job(job); }

int klb_job_set_target_f64(klb_job* j, int which, const double* host, int64_t n) {
  if (!j || !host) return fail(KLB_EINVAL, "null argument");
  CK(cudaSetDevice(j->cfg.device));
  // A kernel of klb_job_run_async may still be reading the old parameters; and the cached pstate.logtarget (for HMC
  // the accepted state's energy) belongs to the old target: the state has to be initialised again afterwards
  // (the reference rebuilds the job when a hyper-parameter changes: parameter.states is fixed at construction,
  // src/jobs/BasicMCJob.jl:51).
  { int rc = sync_all(j); if (rc) return rc; }
  j->have_state = false;
  switch (which) {
    case KLB_PARAM_MU:
      if (n != j->cfg.dim) return fail(KLB_EINVAL, "mu needs dim = %lld values", (long long)j->cfg.dim);
      CK(cudaMemcpy(j->mu, host, n * sizeof(double), cudaMemcpyHostToDevice));
      j->have_mu = true;
      return KLB_OK;
    case KLB_PARAM_SIGMA:
      if (n != j->cfg.dim) return fail(KLB_EINVAL, "sigma needs dim = %lld values", (long long)j->cfg.dim);
      CK(cudaMemcpy(j->sigma, host, n * sizeof(double), cudaMemcpyHostToDevice));
      j->have_sigma = true;
      return KLB_OK;
    case KLB_PARAM_ROSEN:
      if (n != 3) return fail(KLB_EINVAL, "rosenbrock needs 3 values (a, b, scale)");
      memcpy(j->rosen, host, 3 * sizeof(double));
      return KLB_OK;
    case KLB_PARAM_C: {
      const int64_t d = j->cfg.dim;
      if (!j->dense) return fail(KLB_EINVAL, "this job's target takes no precision matrix");
      if (n != d * d) return fail(KLB_EINVAL, "C needs dim*dim = %lld values", (long long)(d * d));
      for (int64_t a = 0; a < d; ++a)          // the kernels read C by rows where the math says columns
        for (int64_t b = a + 1; b < d; ++b)
          if (memcmp(&host[a * d + b], &host[b * d + a], sizeof(double)) != 0)
            return fail(KLB_EINVAL, "C must be exactly symmetric (C[%lld][%lld] != C[%lld][%lld])", (long long)a,
                        (long long)b, (long long)b, (long long)a);
      CK(cudaMemcpy2D(j->Cm, (size_t)j->ld * sizeof(double), host, (size_t)d * sizeof(double), (size_t)d * sizeof(double), (size_t)d,
                      cudaMemcpyHostToDevice));
      j->have_C = true;
      return KLB_OK;
    }
  }
  if (which == KLB_PARAM_LOGIT_X || which == KLB_PARAM_LOGIT_Y || which == KLB_PARAM_LOGIT_LAMBDA) {
    if (!j->glm) return fail(KLB_EINVAL, "this job's target takes no regression data");
    const int64_t d = j->cfg.dim;
    if (which == KLB_PARAM_LOGIT_LAMBDA) {
      if (n != 1 || !(host[0] > 0) || !std::isfinite(host[0])) return fail(KLB_EINVAL, "lambda needs 1 positive value");
      j->lambda = host[0];
      return KLB_OK;
    }
    if (which == KLB_PARAM_LOGIT_X) {
      if (n <= 0 || n % d != 0) return fail(KLB_EINVAL, "X needs ndata*dim values (dim = %lld)", (long long)d);
      const int64_t nd = n / d;
      if (nd > 0x7fffffff) return fail(KLB_EINVAL, "too many observations");
      double* packed = (double*)calloc((size_t)nd * j->gdp, sizeof(double));   // row pitch gdp, zero padded
      if (!packed) return fail(KLB_ENOMEM, "host allocation failed");
      for (int64_t i = 0; i < nd; ++i) memcpy(packed + i * j->gdp, host + i * d, d * sizeof(double));
      cudaFree(j->gX); j->gX = nullptr; j->have_X = false;
      cudaError_t e = cudaMalloc(&j->gX, (size_t)nd * j->gdp * sizeof(double));
      if (e == cudaSuccess) e = cudaMemcpy(j->gX, packed, (size_t)nd * j->gdp * sizeof(double), cudaMemcpyHostToDevice);
      free(packed);
      CK(e);
      j->ndata = nd; j->have_X = true;
      return KLB_OK;
    }
    if (n <= 0) return fail(KLB_EINVAL, "y needs ndata values");
    cudaFree(j->gy); j->gy = nullptr; j->have_y = false;
    CK(cudaMalloc(&j->gy, (size_t)n * sizeof(double)));
    CK(cudaMemcpy(j->gy, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    j->ny = n; j->have_y = true;
    return KLB_OK;
  }
  return fail(KLB_EINVAL, "unknown parameter id %d", which);
}

// KArgs of the chain slice [c0, c0 + nc): every per-chain array advanced to the slice, RNG streams keep their
// global chain indices
static void slice_args(const klb_job* j, KArgs& A, long long c0, long long nc) {
  const long long ld = j->ld, P = j->npost;
  A.state += c0 * ld; A.lt += c0;
  A.tune_step += c0; A.tune_cnt += 3 * c0; A.tune_rate += c0;
  if (A.tune_da) A.tune_da += 8 * c0;
  if (A.out_rate) A.out_rate += c0 * j->nperiods;              // DualAveragingMCTune record: indexed by the slice-local chain id
  if (A.out_value) A.out_value += c0 * P * ld;
  if (A.out_lt) A.out_lt += c0 * P;
  if (A.out_grad) A.out_grad += c0 * P * ld;
  if (A.out_accept) A.out_accept += c0 * P;
  if (A.out_ndoublings) A.out_ndoublings += c0 * P;
  if (A.out_nuts_a) A.out_nuts_a += c0 * P;
  if (A.out_nuts_na) A.out_nuts_na += c0 * P;
  A.nchains = nc;
  A.chain_offset += (unsigned long long)c0;
}

// initialize!: log-target (+ gradient) of every chain of A, non-finite chains flagged     HMC.jl:106-120
static int launch_init(klb_job* j, const KArgs& A, cudaStream_t s) {
  const klb_config& c = j->cfg;
  if (j->glm) {
    GArgs G; fill_glm_args(j, A, G);
    if (klb_glm_init(G, c.arith, j->gdp, glm_smem(j), c.sampler != KLB_SAMPLER_MH, j->flag, s) != 0)
      return fail(KLB_ECUDA, "logistic-regression init kernel launch failed");
  } else if (j->dense) {
    DArgs D; D.k = A; D.Cm = j->Cm; D.nv = j->nv;
    if (klb_dense_init(D, c.arith, c.sampler != KLB_SAMPLER_MH, j->flag, s) != 0)
      return fail(KLB_ECUDA, "dense init kernel launch failed");
  } else if (klb_launch_init(A, c.target, j->gw, j->gnv, c.arith, c.sampler != KLB_SAMPLER_MH, j->flag, s) != 0)
    return fail(KLB_EINVAL, "no init kernel for this configuration");
  j->launches += 1;
  CK(cudaGetLastError());
  return KLB_OK;
}

// all nsteps transitions of the chains of A (A.nt / i0 / count0 / t0 filled here), in launches of `chunk`
static int launch_run(klb_job* j, KArgs A, cudaStream_t s) {
  const klb_config& c = j->cfg;
  long long done = 0, saved = 0;
  const long long chunk = j->chunk > 0 ? j->chunk : c.nsteps;
  while (done < c.nsteps) {
    const long long nt = (c.nsteps - done) < chunk ? (c.nsteps - done) : chunk;
    A.nt = nt; A.i0 = done + 1; A.count0 = saved; A.t0 = j->t_global + (unsigned long long)done;
    if (j->glm) {
      GArgs G; fill_glm_args(j, A, G);
      if (klb_glm_launch(G, c.sampler, c.arith, j->gdp, glm_smem(j), s) != 0)
        return fail(KLB_ECUDA, "logistic-regression kernel launch failed");
    } else if (j->dense) {
      DArgs D; D.k = A; D.Cm = j->Cm; D.nv = j->nv;
      const char* cl = getenv("KLB_DENSE_CLUSTER");    // thread-block clusters of 2 (default) or 4 CTAs share every slab of C; 1 = no clusters
      const int cluster = (cl && cl[0] == '4') ? 4 : (cl && cl[0] == '1') ? 1 : 2;   // default: pairs
      if ((j->dense_mma ? klb_dense_mma_launch(D, c.arith, cluster, s) : klb_dense_launch(D, c.sampler, c.arith, s)) != 0)
        return fail(KLB_ECUDA, "dense kernel launch failed");
    } else if (chain_dispatch(c.sampler, c.arith, &A, c.target, j->gw, j->gnv, c.dim == 64ll * j->gw * j->gnv, nullptr, nullptr, s) != 0)
      return fail(KLB_EINVAL, "no kernel for this configuration");
    CK(cudaGetLastError());
    j->launches += 1;
    done += nt;
    saved = done <= c.burnin ? 0 : (done - c.burnin - 1) / c.thinning + 1;
  }
  return KLB_OK;
}

// tuner_state: BasicMCTune(step, 0, 0, tuner.period); MH gets step 1 (src/samplers/samplers.jl:29-45).
// DualAveragingMCTuner: the constructor keeps leapstep (initialize_step! returns it unchanged, see the oracle's note),
// reset! sets step = 1 (src/samplers/HMC.jl:217-223); mu = log(10*step) either way.
static double tune_step0(const klb_job* j) {
  if (j->cfg.sampler == KLB_SAMPLER_MH) return 1.0;
  if (j->da && j->constructed) return 1.0;
  return j->cfg.step;
}
static int da_refuses_reset(const klb_job* j) {
  if (j->da && j->t_global > 0)
    return fail(KLB_EUNSUPPORTED, "reset of a DualAveragingMCTuner job that has run: the reference's reset! throws "
                                  "(undefined `moment`, src/samplers/samplers.jl:195); create a new job");
  return KLB_OK;
}
static void fill_tune_slice(klb_job* j, size_t c0, size_t nc, cudaStream_t s) {
  const double step0 = tune_step0(j);
  klb_launch_fill_tune(j->tune_step + c0, j->tune_cnt + 3 * c0, j->tune_rate + c0, (long long)nc, step0, j->cfg.period, s);
  j->launches += 1;
  if (j->tune_rates) cudaMemsetAsync(j->tune_rates + c0 * (size_t)j->nperiods, 0xff, nc * (size_t)j->nperiods * sizeof(double), s);
  if (j->da) {
    klb_launch_fill_da(j->tune_da + 8 * c0, (long long)nc,
                       j->cfg.sampler == KLB_SAMPLER_NUTS ? klb_u2d(0x7FF8000000000000ULL)       // λ = NaN   NUTS.jl:260-269
                                                          : (double)j->cfg.nleaps * j->cfg.step,
                       klb_log(10 * step0, KLB_TAB), j->cfg.da_eps0bar, j->cfg.da_h0bar, s);
    j->launches += 1;
  }
}

static int reset_tune(klb_job* j) {
  fill_tune_slice(j, 0, (size_t)j->cfg.nchains, j->stream);
  j->constructed = true;
  CK(cudaGetLastError());
  j->count = 0;
  return KLB_OK;
}

static int check_params(klb_job* j) {
  const klb_config& c = j->cfg;
  if (c.target == KLB_TARGET_SHIFTED_ISO && !j->have_mu) return fail(KLB_ESTATE, "set KLB_PARAM_MU before the state");
  if (c.sampler == KLB_SAMPLER_MH && !j->have_sigma) return fail(KLB_ESTATE, "set KLB_PARAM_SIGMA before the state");
  if (j->dense && !j->have_C) return fail(KLB_ESTATE, "set KLB_PARAM_C before the state");
  if (j->glm) {
    if (!j->have_X || !j->have_y) return fail(KLB_ESTATE, "set KLB_PARAM_LOGIT_X and KLB_PARAM_LOGIT_Y before the state");
    if (j->ny != j->ndata) return fail(KLB_EINVAL, "X has %lld rows, y has %lld entries", j->ndata, j->ny);
    if (klb_glm_attrs(c.sampler, c.arith, j->gdp, glm_smem(j), &j->regs, &j->bps) != 0) {
      cudaGetLastError();
      return fail(KLB_ECUDA, "no sm_100a kernel image for the logistic-regression kernel (dp %d)", j->gdp);
    }
  }
  return KLB_OK;
}

static int init_state(klb_job* j) {
  const klb_config& c = j->cfg;
  { int rc = check_params(j); if (rc) return rc; }
  KArgs A;
  fill_args(j, A);
  const unsigned long long none = std::numeric_limits<unsigned long long>::max();
  CK(cudaMemcpyAsync(j->flag, &none, sizeof none, cudaMemcpyHostToDevice, j->stream));
  { int rc = launch_init(j, A, j->stream); if (rc) return rc; }
  unsigned long long f = 0;
  CK(cudaMemcpyAsync(&f, j->flag, sizeof f, cudaMemcpyDeviceToHost, j->stream));
  CK(cudaStreamSynchronize(j->stream));
  if (f != none) {
    j->have_state = false;
    return fail(KLB_ENOTFINITE, "Log-target%s not finite: initial value out of support (chain %llu)",
                c.sampler != KLB_SAMPLER_MH ? " or its gradient" : "", f - 1);
  }
  j->have_state = true;
  return reset_tune(j);
}


int klb_job_set_state(klb_job* j, const double* x0) {
  if (!j || !x0) return fail(KLB_EINVAL, "null argument");
  if (j->host_pending) return fail(KLB_ESTATE, "a pipelined run is pending: call klb_job_run_host_finish first");
  { int rc = da_refuses_reset(j); if (rc) return rc; }
  CK(cudaSetDevice(j->cfg.device));
  CK(cudaMemcpy2DAsync(j->state, (size_t)j->ld * 8, x0, (size_t)j->cfg.dim * 8, (size_t)j->cfg.dim * 8,
                       (size_t)j->cfg.nchains, cudaMemcpyHostToDevice, j->stream));
  return init_state(j);
}

int klb_job_set_state_device(klb_job* j, const double* x0_dev) {
  if (!j || !x0_dev) return fail(KLB_EINVAL, "null argument");
  if (j->host_pending) return fail(KLB_ESTATE, "a pipelined run is pending: call klb_job_run_host_finish first");
  { int rc = da_refuses_reset(j); if (rc) return rc; }
  CK(cudaSetDevice(j->cfg.device));
  if (x0_dev != j->state)   // x0_dev is a dense dim x nchains matrix
    CK(cudaMemcpy2DAsync(j->state, (size_t)j->ld * 8, x0_dev, (size_t)j->cfg.dim * 8, (size_t)j->cfg.dim * 8,
                         (size_t)j->cfg.nchains, cudaMemcpyDeviceToDevice, j->stream));
  return init_state(j);
}

// x0 = the synthetic initial value of the benchmark configurations, generated on the device (header)
int klb_job_set_state_synthetic(klb_job* j) {
  if (!j) return fail(KLB_EINVAL, "null argument");
  if (j->host_pending) return fail(KLB_ESTATE, "a pipelined run is pending: call klb_job_run_host_finish first");
  { int rc = da_refuses_reset(j); if (rc) return rc; }
  CK(cudaSetDevice(j->cfg.device));
  klb_launch_synth_state(j->tab, j->cfg.seed, (uint64_t)j->cfg.chain_offset, j->cfg.nchains, (int)j->cfg.dim, j->ld,
                         j->state, j->stream);
  j->launches += 1;
  CK(cudaGetLastError());
  return init_state(j);
}

// Position the RNG: the next transition will be number t + 1 of the job's streams.  Counter-based generators
// can seek; the reference's global MersenneTwister cannot.
int klb_job_seek(klb_job* j, uint64_t t) {
  if (!j) return fail(KLB_EINVAL, "null argument");
  if (t >> 48) return fail(KLB_EINVAL, "the transition counter has 48 bits");
  if (j->da && j->t_global > 0) return fail(KLB_EUNSUPPORTED, "a DualAveragingMCTuner job that has run cannot be repositioned");
  if (j->da && t > 0) return fail(KLB_EUNSUPPORTED, "a DualAveragingMCTuner job starts at transition 0 (reset! rule)");
  { int rc = sync_all(j); if (rc) return rc; }
  j->t_global = t;
  return KLB_OK;
}

int klb_job_set_chunk(klb_job* j, int64_t nt) {
  if (!j || nt < 0) return fail(KLB_EINVAL, "bad chunk");
  j->chunk = nt;
  return KLB_OK;
}

int klb_job_run_async(klb_job* j) {
  if (!j) return fail(KLB_EINVAL, "null argument");
  if (j->host_pending) return fail(KLB_ESTATE, "a pipelined run is pending: call klb_job_run_host_finish first");
  if (!j->have_state) return fail(KLB_ESTATE, "klb_job_set_state must succeed before klb_job_run");
  const klb_config& c = j->cfg;
  // a second run without reset would write past column npoststeps of the NState (a BoundsError in the reference)
  if (j->count != 0) return fail(KLB_ESTATE, "output already holds %lld samples: call klb_job_reset first", j->count);
  CK(cudaSetDevice(c.device));
  KArgs A;
  fill_args(j, A);
  CK(cudaEventRecord(j->ev0, j->stream));
  { int rc = launch_run(j, A, j->stream); if (rc) return rc; }
  j->t_global += (unsigned long long)c.nsteps;
  CK(cudaEventRecord(j->ev1, j->stream));
  j->count = j->npost;
  j->timed = true;
  return KLB_OK;
}

int klb_job_sync(klb_job* j) {
  if (!j) return fail(KLB_EINVAL, "null argument");
  CK(cudaSetDevice(j->cfg.device));
  CK(cudaStreamSynchronize(j->stream));
  return KLB_OK;
}

int klb_job_run(klb_job* j) {
  int rc = klb_job_run_async(j);
  if (rc) return rc;
  return klb_job_sync(j);
}

static int field_ptr(klb_job* j, int field, void** p, size_t* nb, size_t* cols);

// set_state + run + output in one pipelined call (header: klb_job_run_host)
int klb_job_run_host_finish(klb_job* j);
// enqueue half of klb_job_run_host (header: klb_job_run_host_async)
int klb_job_run_host_async(klb_job* j, const double* x0, const klb_host_field* fields, int32_t nfields, int32_t nslices) {
  if (!j || (nfields > 0 && !fields) || nfields < 0) return fail(KLB_EINVAL, "null argument");
  if (j->host_pending) return fail(KLB_ESTATE, "a pipelined run is pending: call klb_job_run_host_finish first");
  const klb_config& c = j->cfg;
  CK(cudaSetDevice(c.device));
  if (!x0 && !j->have_state) return fail(KLB_ESTATE, "no initial value: pass x0 or call klb_job_set_state first");
  { int rc = da_refuses_reset(j); if (rc) return rc; }
  { int rc = check_params(j); if (rc) return rc; }
  const size_t N = (size_t)c.nchains, d = (size_t)c.dim, ld = (size_t)j->ld;
  // validate the requested fields before anything is enqueued
  struct Fld { char* p; size_t per_chain, cols_per_chain; };
  Fld fl[32];
  if (nfields > 32) return fail(KLB_EINVAL, "at most 32 output fields");
  for (int q = 0; q < nfields; ++q) {
    void* p; size_t nb, cols;
    int rc = field_ptr(j, fields[q].field, &p, &nb, &cols);
    if (rc) return rc;
    if (fields[q].field == KLB_OUT_ESS) return fail(KLB_EINVAL, "KLB_OUT_ESS is produced by klb_job_ess, not by a run");
    if (!fields[q].host_dst) return fail(KLB_EINVAL, "null host buffer for field %d", fields[q].field);
    if ((size_t)fields[q].nbytes != nb)
      return fail(KLB_EINVAL, "field %d holds %zu bytes, caller passed %lld", fields[q].field, nb, (long long)fields[q].nbytes);
    fl[q].p = (char*)p; fl[q].per_chain = nb / N; fl[q].cols_per_chain = cols / N;
  }
  int S = nslices;
  if (S <= 0) {                                    // auto: ~4 MiB of state per slice, at least 512 chains, at most 16 slices
    const size_t bytes = N * d * 8;                // (C3 on one B200: 16 slices 74.1 ms, 8: 75.3, 4: 78.7, serial: 94.3;
    size_t mib = 4;                                //  small shards -- 8192 chains per GPU at N = 8 -- need as many slices
    if (const char* env = getenv("KLB_SLICE_MIB")) //  for their copies to hide behind the kernels of the other slices)
      if (atoi(env) > 0) mib = (size_t)atoi(env);
    S = (int)std::min<size_t>(KLB_MAX_SLICES, bytes / (mib << 20));
    while (S > 1 && N / (size_t)S < 512) --S;
  }
  if (S < 1) S = 1;
  if (S > KLB_MAX_SLICES) S = KLB_MAX_SLICES;
  if ((size_t)S > N) S = (int)N;
  while (j->nsl_streams < S) {
    CK(cudaStreamCreateWithFlags(&j->sl_stream[j->nsl_streams], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&j->sl_done[j->nsl_streams], cudaEventDisableTiming));
    j->nsl_streams += 1;
  }
  const unsigned long long none = std::numeric_limits<unsigned long long>::max();
  // Everything below is asynchronous on the slice streams and touches caller-owned host buffers, so no error path may
  // return before those streams are idle; and nothing of the job's bookkeeping (RNG counter, output cursor,
  // `constructed`) moves until the whole pipeline has succeeded -- a bad x0 or a failed launch leaves a job that
  // klb_job_set_state / klb_job_run_host can restart (a DualAveragingMCTuner job included: t_global stays 0).
  int rc = KLB_OK;
#define CKR(call)                                                                                              \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess && rc == KLB_OK)                                                                     \
      rc = fail(e_ == cudaErrorMemoryAllocation ? KLB_ENOMEM : KLB_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)
  const bool was_constructed = j->constructed;
  if (x0) CKR(cudaMemcpyAsync(j->flag, &none, sizeof none, cudaMemcpyHostToDevice, j->stream));
  CKR(cudaEventRecord(j->ev0, j->stream));         // slices start after everything already queued on the job stream
  for (int q = 0; q < S && rc == KLB_OK; ++q) {
    const size_t c0 = N * (size_t)q / (size_t)S, c1 = N * (size_t)(q + 1) / (size_t)S, nc = c1 - c0;
    cudaStream_t st = j->sl_stream[q];
    CKR(cudaStreamWaitEvent(st, j->ev0, 0));
    KArgs A;
    fill_args(j, A);
    slice_args(j, A, (long long)c0, (long long)nc);
    if (x0 && rc == KLB_OK) {                      // initialize! / reset(job, x) of the slice
      CKR(cudaMemcpy2DAsync(j->state + c0 * ld, ld * 8, x0 + c0 * d, d * 8, d * 8, nc, cudaMemcpyHostToDevice, st));
      if (rc == KLB_OK) rc = launch_init(j, A, st);
    }
    if (rc != KLB_OK) break;
    // reset(job): tuner records of the slice                                  samplers.jl:29-45
    fill_tune_slice(j, c0, nc, st);
    CKR(cudaGetLastError());
    if (rc == KLB_OK) rc = launch_run(j, A, st);
    for (int f = 0; f < nfields && rc == KLB_OK; ++f) {   // output(job), slice by slice
      const size_t per = fl[f].per_chain, cpc = fl[f].cols_per_chain;
      char* dst = (char*)fields[f].host_dst + c0 * per;
      if (cpc && ld != d)                          // odd dim: device columns are padded
        CKR(cudaMemcpy2DAsync(dst, d * 8, fl[f].p + c0 * cpc * ld * 8, ld * 8, d * 8, nc * cpc, cudaMemcpyDeviceToHost, st));
      else
        CKR(cudaMemcpyAsync(dst, fl[f].p + c0 * per, nc * per, cudaMemcpyDeviceToHost, st));
    }
    CKR(cudaEventRecord(j->sl_done[q], st));
    if (rc == KLB_OK) CKR(cudaStreamWaitEvent(j->stream, j->sl_done[q], 0));
  }
  *j->flag_host = none;
  if (rc == KLB_OK) {
    CKR(cudaEventRecord(j->ev1, j->stream));
    if (x0) CKR(cudaMemcpyAsync(j->flag_host, j->flag, sizeof(unsigned long long), cudaMemcpyDeviceToHost, j->stream));
  }
#undef CKR
  if (rc != KLB_OK) {                              // a failed enqueue: idle the streams (host buffers are the caller's), roll back
    char keep[sizeof g_err];
    memcpy(keep, g_err, sizeof keep);
    sync_all(j);
    memcpy(g_err, keep, sizeof keep);
    j->have_state = false; j->count = 0; j->constructed = was_constructed; j->timed = false;
    return rc;
  }
  j->host_pending = true; j->host_had_x0 = x0 != nullptr; j->host_was_constructed = was_constructed;
  return KLB_OK;
}

// wait for the slices and look at the finiteness flag; the run stays pending (several jobs can be inspected before any
// of them is committed: klb_multi_run_host is all-or-nothing)
int klb_job_run_host_wait(klb_job* j) {
  if (!j) return fail(KLB_EINVAL, "null argument");
  if (!j->host_pending) return fail(KLB_ESTATE, "no pipelined run is pending");
  const klb_config& c = j->cfg;
  const unsigned long long none = std::numeric_limits<unsigned long long>::max();
  int rc = sync_all(j);                            // idle streams: host buffers are the caller's
  if (rc == KLB_OK && j->host_had_x0 && *j->flag_host != none)
    rc = fail(KLB_ENOTFINITE, "Log-target%s not finite: initial value out of support (chain %llu)",
              c.sampler != KLB_SAMPLER_MH ? " or its gradient" : "", *j->flag_host - 1);
  return rc;
}

static void host_rollback(klb_job* j) {            // the state buffers hold a discarded run
  j->host_pending = false;
  j->have_state = false;
  j->count = 0;
  j->constructed = j->host_was_constructed;
  j->timed = false;
}

// discard a pending run (as if its start had been rejected)
int klb_job_run_host_abort(klb_job* j) {
  if (!j) return fail(KLB_EINVAL, "null argument");
  if (!j->host_pending) return fail(KLB_ESTATE, "no pipelined run is pending");
  char keep[sizeof g_err];
  memcpy(keep, g_err, sizeof keep);
  sync_all(j);
  memcpy(g_err, keep, sizeof keep);
  host_rollback(j);
  return KLB_OK;
}

// second half: wait, then commit the job's bookkeeping -- or roll it back when the run failed
int klb_job_run_host_finish(klb_job* j) {
  const int rc = klb_job_run_host_wait(j);
  if (rc == KLB_EINVAL || (rc == KLB_ESTATE && !(j && j->host_pending))) return rc;
  if (rc != KLB_OK) { host_rollback(j); return rc; }
  j->host_pending = false;
  j->t_global += (unsigned long long)j->cfg.nsteps;
  j->count = j->npost;
  j->timed = true;
  j->constructed = true;
  j->have_state = true;
  return KLB_OK;
}

int klb_job_run_host(klb_job* j, const double* x0, const klb_host_field* fields, int32_t nfields, int32_t nslices) {
  const int rc = klb_job_run_host_async(j, x0, fields, nfields, nslices);
  return rc ? rc : klb_job_run_host_finish(j);
}

int klb_job_reset(klb_job* j) {
  if (!j) return fail(KLB_EINVAL, "null argument");
  if (j->host_pending) return fail(KLB_ESTATE, "a pipelined run is pending: call klb_job_run_host_finish first");
  { int rc = da_refuses_reset(j); if (rc) return rc; }
  CK(cudaSetDevice(j->cfg.device));
  return reset_tune(j);
}

// *cols > 0: the field is a matrix of `cols` columns of cfg.dim doubles stored with leading dimension ld
static int field_ptr(klb_job* j, int field, void** p, size_t* nb, size_t* cols) {
  const size_t N = (size_t)j->cfg.nchains, d = (size_t)j->cfg.dim, P = (size_t)j->npost;
  *cols = 0;
  switch (field) {
    case KLB_OUT_VALUE: *p = j->out_value; *nb = N * P * d * 8; *cols = N * P; break;
    case KLB_OUT_LOGTARGET: *p = j->out_lt; *nb = N * P * 8; break;
    case KLB_OUT_GRADLOGTARGET: *p = j->out_grad; *nb = N * P * d * 8; *cols = N * P; break;
    case KLB_OUT_ACCEPT: *p = j->out_accept; *nb = N * P; break;
    case KLB_OUT_NDOUBLINGS: *p = j->out_ndoublings; *nb = N * P; break;
    case KLB_OUT_NUTS_A: *p = j->out_nuts_a; *nb = N * P * 8; break;
    case KLB_OUT_NUTS_NA: *p = j->out_nuts_na; *nb = N * P * 4; break;
    case KLB_OUT_STATE: *p = j->state; *nb = N * d * 8; *cols = N; break;
    case KLB_OUT_STATE_LOGTARGET: *p = j->lt; *nb = N * 8; break;
    case KLB_OUT_TUNE_STEP: *p = j->tune_step; *nb = N * 8; break;
    case KLB_OUT_TUNE_COUNTERS: *p = j->tune_cnt; *nb = 3 * N * 8; break;
    case KLB_OUT_TUNE_RATE: *p = j->tune_rate; *nb = N * 8; break;
    case KLB_OUT_ESS: *p = j->ess; *nb = N * d * 8; break;
    case KLB_OUT_TUNE_DA: *p = j->tune_da; *nb = 8 * N * 8; break;
    case KLB_OUT_TUNE_RATES: *p = j->tune_rates; *nb = (size_t)j->nperiods * N * 8; break;
    default: return fail(KLB_EINVAL, "unknown field %d", field);
  }
  if (!*p) return fail(KLB_ESTATE, "field %d is not monitored by this job", field);
  return KLB_OK;
}

int klb_job_output(klb_job* j, int field, void* host_dst, int64_t nbytes) {
  if (!j || !host_dst) return fail(KLB_EINVAL, "null argument");
  if (j->host_pending) return fail(KLB_ESTATE, "a pipelined run is pending: call klb_job_run_host_finish first");
  void* p; size_t nb, cols;
  int rc = field_ptr(j, field, &p, &nb, &cols);
  if (rc) return rc;
  if ((size_t)nbytes != nb) return fail(KLB_EINVAL, "field %d holds %zu bytes, caller passed %lld", field, nb, (long long)nbytes);
  CK(cudaSetDevice(j->cfg.device));
  if (cols && j->ld != j->cfg.dim)   // odd dim: device columns are padded to an even leading dimension
    CK(cudaMemcpy2DAsync(host_dst, (size_t)j->cfg.dim * 8, p, (size_t)j->ld * 8, (size_t)j->cfg.dim * 8, cols,
                         cudaMemcpyDeviceToHost, j->stream));
  else
    CK(cudaMemcpyAsync(host_dst, p, nb, cudaMemcpyDeviceToHost, j->stream));
  CK(cudaStreamSynchronize(j->stream));
  return KLB_OK;
}

int klb_job_device_ptr(klb_job* j, int field, void** dev_ptr, int64_t* nbytes) {
  if (!j || !dev_ptr || !nbytes) return fail(KLB_EINVAL, "null argument");
  size_t nb, cols;
  int rc = field_ptr(j, field, dev_ptr, &nb, &cols);
  if (rc) return rc;
  *nbytes = (int64_t)(cols ? cols * (size_t)j->ld * 8 : nb);   // matrices: `cols` columns with leading dimension plan.ld
  return KLB_OK;
}

int klb_job_ess(klb_job* j, double* host_ess) {
  if (!j) return fail(KLB_EINVAL, "null argument");
  if (!j->out_value) return fail(KLB_ESTATE, "ess needs the monitored values (outopts[:monitor] must include :value)");
  if (j->count != j->npost) return fail(KLB_ESTATE, "run the job before asking for its effective sample size");
  CK(cudaSetDevice(j->cfg.device));
  const size_t N = (size_t)j->cfg.nchains, d = (size_t)j->cfg.dim;
  if (!j->ess) CK(cudaMalloc(&j->ess, N * d * sizeof(double)));
  for (size_t c0 = 0; c0 < N; c0 += 32768) {   // gridDim.y <= 65535: chains go in blocks
    const size_t nc = (N - c0) < 32768 ? (N - c0) : 32768;
    klb_launch_ess(j->out_value + c0 * (size_t)j->npost * (size_t)j->ld, j->ld, j->npost, (long long)nc, (int)d,
                   j->ess + c0 * d, j->stream);
    j->launches += 1;
  }
  CK(cudaGetLastError());
  if (host_ess) CK(cudaMemcpyAsync(host_ess, j->ess, N * d * sizeof(double), cudaMemcpyDeviceToHost, j->stream));
  CK(cudaStreamSynchronize(j->stream));
  return KLB_OK;
}

// mean / mcvar / ess / iact / acceptance of the monitored output, on the device (src/stats/*.jl)
int klb_job_stat(klb_job* j, int stat, double* host_dst) {
  if (!j) return fail(KLB_EINVAL, "null argument");
  if (stat < 0 || stat > KLB_STAT_ACCEPTANCE_VALUE) return fail(KLB_EINVAL, "unknown statistic %d", stat);
  if (j->count != j->npost) return fail(KLB_ESTATE, "run the job before asking for statistics of its output");
  CK(cudaSetDevice(j->cfg.device));
  const size_t N = (size_t)j->cfg.nchains, d = (size_t)j->cfg.dim;
  if (stat == KLB_STAT_ACCEPTANCE || stat == KLB_STAT_ACCEPTANCE_VALUE) {
    const bool diag = stat == KLB_STAT_ACCEPTANCE;
    if (diag && !j->out_accept)
      return fail(KLB_ESTATE, "acceptance needs the :accept diagnostic (outopts[:diagnostics] must include :accept)");
    if (!diag && !j->out_value)
      return fail(KLB_ESTATE, "acceptance(diagnostics=false) needs the monitored values (outopts[:monitor] must include :value)");
    if (!j->accrate) CK(cudaMalloc(&j->accrate, N * sizeof(double)));
    klb_launch_acceptance(diag ? j->out_accept : nullptr, j->out_value, j->ld, j->npost, (long long)N, (int)d, j->accrate,
                          j->stream);
    j->launches += 1;
    CK(cudaGetLastError());
    if (host_dst) CK(cudaMemcpyAsync(host_dst, j->accrate, N * sizeof(double), cudaMemcpyDeviceToHost, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    return KLB_OK;
  }
  if (!j->out_value) return fail(KLB_ESTATE, "statistics need the monitored values (outopts[:monitor] must include :value)");
  if (j->stat_epoch != j->t_global + 1) {
    if (!j->ess) CK(cudaMalloc(&j->ess, N * d * sizeof(double)));
    j->stat[KLB_STAT_ESS] = j->ess;
    for (int q = 0; q < 5; ++q)
      if (!j->stat[q]) CK(cudaMalloc(&j->stat[q], N * d * sizeof(double)));
    for (size_t c0 = 0; c0 < N; c0 += 32768) {   // gridDim.y <= 65535: chains go in blocks
      const size_t nc = (N - c0) < 32768 ? (N - c0) : 32768;
      double* st[5];
      for (int q = 0; q < 5; ++q) st[q] = j->stat[q] + c0 * d;
      klb_launch_stats(j->out_value + c0 * (size_t)j->npost * (size_t)j->ld, j->ld, j->npost, (long long)nc, (int)d, st,
                       j->stream);
      j->launches += 1;
    }
    CK(cudaGetLastError());
    j->stat_epoch = j->t_global + 1;
  }
  if (host_dst) CK(cudaMemcpyAsync(host_dst, j->stat[stat], N * d * sizeof(double), cudaMemcpyDeviceToHost, j->stream));
  CK(cudaStreamSynchronize(j->stream));
  return KLB_OK;
}

int klb_job_plan(klb_job* j, klb_plan* out) {
  if (!j || !out) return fail(KLB_EINVAL, "null argument");
  out->nv = j->nv;
  out->ld = j->ld;
  out->warps_per_block = j->glm ? KLB_GLM_THREADS / 32 : use_hmc_ws(j->cfg.sampler, j->gw, j->gnv) && !j->dense ? 8 : KLB_WPB;
  out->warps_per_chain = j->glm ? 0 : j->gw;       // 0: one THREAD per chain
  out->regs_per_thread = j->regs;
  out->blocks_per_sm = j->bps;
  out->npoststeps = j->npost;
  out->transitions_done = (int64_t)j->t_global;
  out->saved = j->count;
  return KLB_OK;
}

int klb_job_config(klb_job* j, klb_config* out) {
  if (!j || !out) return fail(KLB_EINVAL, "null argument");
  *out = j->cfg;
  return KLB_OK;
}

int64_t klb_job_launches(klb_job* j) { return j ? j->launches : 0; }

double klb_job_last_run_ms(klb_job* j) {
  if (!j || !j->timed) return -1.0;
  cudaSetDevice(j->cfg.device);
  float ms = 0.f;
  if (cudaEventSynchronize(j->ev1) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, j->ev0, j->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}

void* klb_job_stream(klb_job* j) { return j ? (void*)j->stream : nullptr; }

int klb_host_alloc(void** p, int64_t nbytes) {
  if (!p || nbytes <= 0) return fail(KLB_EINVAL, "bad argument");
  CK(cudaHostAlloc(p, (size_t)nbytes, cudaHostAllocDefault));
  return KLB_OK;
}
int klb_host_free(void* p) {
  CK(cudaFreeHost(p));
  return KLB_OK;
}

// fp64 / DMMA throughput of one device, measured with a stream of independent instructions (roofline denominators)
int klb_device_peak(int device, int kind, double* per_second) {
  if (!per_second || (kind != KLB_PEAK_FP64 && kind != KLB_PEAK_DMMA)) return fail(KLB_EINVAL, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(KLB_ECUDA, "no CUDA device available (this library has no CPU path)");
  }
  if (device < 0 || device >= ndev) return fail(KLB_EINVAL, "device out of range");
  CK(cudaSetDevice(device));
  if (klb_measure_peak(kind, per_second) != 0) { cudaGetLastError(); return fail(KLB_ECUDA, "peak measurement failed"); }
  return KLB_OK;
}

// ---------------------------------------------------------------- device self-tests
// device scratch of one self-test call, released on every return path
struct DbgBufs {
  void* p[3] = {nullptr, nullptr, nullptr};
  ~DbgBufs() { for (void* q : p) if (q) cudaFree(q); }
};
static int dbg_setup(int device, DbgBufs& b) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(KLB_ECUDA, "no CUDA device available (this library has no CPU path)");
  }
  if (device < 0 || device >= ndev) return fail(KLB_EINVAL, "device out of range");
  CK(cudaSetDevice(device));
  CK(cudaMalloc(&b.p[0], sizeof(KLB_TAB)));
  CK(cudaMemcpy(b.p[0], KLB_TAB, sizeof(KLB_TAB), cudaMemcpyHostToDevice));
  return KLB_OK;
}

int klb_debug_normals(int device, uint64_t seed, uint64_t chain, uint64_t t, int64_t n, double* host_out) {
  if (!host_out || n <= 0) return fail(KLB_EINVAL, "bad argument");
  DbgBufs b;
  int rc = dbg_setup(device, b);
  if (rc) return rc;
  CK(cudaMalloc(&b.p[1], (size_t)n * 8));
  klb_launch_debug_normals((const uint64_t*)b.p[0], seed, chain, t, n, (double*)b.p[1], 0);
  CK(cudaGetLastError());
  CK(cudaMemcpy(host_out, b.p[1], (size_t)n * 8, cudaMemcpyDeviceToHost));
  return KLB_OK;
}

int klb_debug_math(int device, int op, int64_t n, const double* host_in, double* host_out) {
  if (!host_in || !host_out || n <= 0) return fail(KLB_EINVAL, "bad argument");
  DbgBufs b;
  int rc = dbg_setup(device, b);
  if (rc) return rc;
  CK(cudaMalloc(&b.p[1], (size_t)n * 8));
  CK(cudaMalloc(&b.p[2], (size_t)n * 8));
  CK(cudaMemcpy(b.p[1], host_in, (size_t)n * 8, cudaMemcpyHostToDevice));
  klb_launch_debug_math((const uint64_t*)b.p[0], op, n, (const double*)b.p[1], (double*)b.p[2], 0);
  CK(cudaGetLastError());
  CK(cudaMemcpy(host_out, b.p[2], (size_t)n * 8, cudaMemcpyDeviceToHost));
  return KLB_OK;
}

int klb_debug_uniform(int device, uint64_t seed, uint64_t chain, uint64_t t, double* host_out) {
  if (!host_out) return fail(KLB_EINVAL, "bad argument");
  DbgBufs b;
  int rc = dbg_setup(device, b);
  if (rc) return rc;
  CK(cudaMalloc(&b.p[1], 8));
  klb_launch_debug_uniform(seed, chain, t, (double*)b.p[1], 0);
  CK(cudaGetLastError());
  CK(cudaMemcpy(host_out, b.p[1], 8, cudaMemcpyDeviceToHost));
  return KLB_OK;
}

}  // extern "C"
