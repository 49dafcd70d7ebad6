// klb_aux.cu -- small support kernels: tuner-record reset and the device self-tests that the
// parity suite uses to compare the device math/RNG primitives with the oracle bit by bit.
#include <cstdlib>
#include "klb_kernels.cuh"

#ifndef KLB_ESS_DEFAULT_VARIANT
#define KLB_ESS_DEFAULT_VARIANT 7   /* klb_ess_win_kernel<32, 16, true>; 0: klb_ess_tile_kernel (round 1: eight lags per pass, centred tile) */
#endif

// tuner_state / reset!(tune, sampler, tuner):  (step, 0, 0, tuner.period, NaN)
//                                              src/samplers/samplers.jl:29-45, 79-90
__global__ void klb_fill_tune_kernel(double* step, long long* cnt, double* rate, long long n, double step0,
                                     long long period) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  step[c] = step0;
  cnt[3 * c] = 0; cnt[3 * c + 1] = 0; cnt[3 * c + 2] = period;
  rate[c] = klb_u2d(0x7FF8000000000000ULL);
}
void klb_launch_fill_tune(double* step, long long* cnt, double* rate, long long n, double step0, long long period,
                          cudaStream_t s) {
  klb_fill_tune_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(step, cnt, rate, n, step0, period);
}

// DualAveragingMCTune record of tuner_state / sampler_state / reset! for HMC (src/samplers/HMC.jl:124-133, 192-213,
// 217-223): [lambda, mu, epsbar, hbar, hweight = NaN, epsweight = NaN, nleaps = 0, count = 0]
__global__ void klb_fill_da_kernel(double* da, long long n, double lambda, double mu, double epsbar, double hbar) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double* r = da + 8 * c;
  const double nan = klb_u2d(0x7FF8000000000000ULL);
  r[0] = lambda; r[1] = mu; r[2] = epsbar; r[3] = hbar; r[4] = nan; r[5] = nan; r[6] = 0.0; r[7] = 0.0;
}
void klb_launch_fill_da(double* da, long long n, double lambda, double mu, double epsbar, double hbar, cudaStream_t s) {
  klb_fill_da_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(da, n, lambda, mu, epsbar, hbar);
}

__global__ void klb_debug_normals_kernel(const uint64_t* gtab, uint64_t seed, uint64_t chain, uint64_t t, long long n,
                                         double* out) {
  __shared__ uint64_t tab[KLB_TAB_LEN];
  for (int i = threadIdx.x; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  const klb_stream st = klb_stream_make(seed, chain, t);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = klb_normal(&st, (uint32_t)i, tab);
}
void klb_launch_debug_normals(const uint64_t* tab, uint64_t seed, uint64_t chain, uint64_t t, long long n, double* out,
                              cudaStream_t s) {
  klb_debug_normals_kernel<<<148, 256, 0, s>>>(tab, seed, chain, t, n, out);
}

__global__ void klb_debug_math_kernel(const uint64_t* gtab, int op, long long n, const double* in, double* out) {
  __shared__ uint64_t tab[KLB_TAB_LEN];
  for (int i = threadIdx.x; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (op == 0) out[i] = klb_exp(in[i], tab);
    else if (op == 1) out[i] = klb_log(in[i], tab);
    else if (op == 2) out[i] = klb_erf(in[i], tab);
    else { const DivBy by(in[i | 1]); out[i] = by(in[i & ~1ll]); }            // op 3: pairs (a, b) -> a / b through DivBy, twice
  }
}
void klb_launch_debug_math(const uint64_t* tab, int op, long long n, const double* in, double* out, cudaStream_t s) {
  klb_debug_math_kernel<<<148, 256, 0, s>>>(tab, op, n, in, out);
}

__global__ void klb_debug_uniform_kernel(uint64_t seed, uint64_t chain, uint64_t t, double* out) {
  const klb_stream st = klb_stream_make(seed, chain, t);
  out[0] = klb_accept_uniform(&st);
}
void klb_launch_debug_uniform(uint64_t seed, uint64_t chain, uint64_t t, double* out, cudaStream_t s) {
  klb_debug_uniform_kernel<<<1, 1, 0, s>>>(seed, chain, t, out);
}

// ------------------------------------------------------------------ post-hoc statistics on device
// ess(chain, :imse) for every (coordinate, chain) series of the monitored values:
//   ess = len * mcvar(:iid) / mcvar(:imse)           src/stats/convergence/ess.jl:3-14
//   mcvar(:iid) = var(v)/len                          src/stats/variance/mcvar.jl:5
//   mcvar(:imse): Geyer's initial monotone sequence estimator over autocov(v, 0:len-1)   mcvar.jl:75-105
// One thread per series; consecutive threads take consecutive coordinates, so every load of sample s is a
// coalesced row segment of the `ld x npost x nchains` value array and the CTA's working set (128 coordinates
// x npost samples) stays in L1 across the lag passes.  Lags are evaluated pairwise until the first
// non-positive G_j -- all the estimator reads.  Sequential, fma-accumulated sums: the oracle's order.
// Every estimator the reference derives from one series shares the same sums, so one pass fills them all
// (any pointer may be null):
//   mean   = mean(s, i)                                src/stats/mean.jl:9
//   iid    = mcvar(s, Val{:iid}, i) = var(v)/len       src/stats/variance/mcvar.jl:5
//   imse   = mcvar(s, Val{:imse}, i)                   src/stats/variance/mcvar.jl:75-105
//   ess    = len*iid/imse                              src/stats/convergence/ess.jl:3
//   iact   = imse/iid                                  src/stats/convergence/iact.jl:3
struct KlbStatPtrs { double *mean, *iid, *imse, *ess, *iact; };
__device__ __forceinline__ void stat_store(const KlbStatPtrs& S, long long idx, double mean, double iid, double imse,
                                           double ess, double iact) {
  if (S.mean) S.mean[idx] = mean;
  if (S.iid) S.iid[idx] = iid;
  if (S.imse) S.imse[idx] = imse;
  if (S.ess) S.ess[idx] = ess;
  if (S.iact) S.iact[idx] = iact;
}

__global__ void __launch_bounds__(128)
klb_ess_kernel(const double* __restrict__ value, long long ld, long long npost, int dim, const KlbStatPtrs S) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const long long c = blockIdx.y;
  if (i >= dim) return;
  const double* v = value + c * npost * ld + i;
  const long long n = npost;
  const double qnan = klb_u2d(0x7FF8000000000000ULL);
  double out = qnan, o_mean = qnan, o_iid = qnan, o_imse = qnan, o_iact = qnan;
  if (n >= 1) {
    double s = 0.0;
    for (long long t = 0; t < n; ++t) s = __dadd_rn(s, v[t * ld]);
    o_mean = __ddiv_rn(s, (double)n);
  }
  if (n >= 4) {
    const double mu = o_mean;
    const double dn = (double)n;
    const long long k = (n - 2) / 2;                                // floor((maxlag-1)/2), maxlag = n-1
    double sumg = 0.0, gprev = 0.0, s0 = 0.0;
    bool done = false;
    // eight lags per pass over the series: a sliding register window holds z[t+L .. t+L+7], so a pass costs two
    // loads per eight fma; every lag's sum still runs sequentially in t (the oracle's order)
    for (long long L = 0; !done && L <= 2 * k + 1; L += 8) {
      double acc[8], w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { acc[q] = 0.0; w[q] = (L + q < n) ? __dsub_rn(v[(L + q) * ld], mu) : 0.0; }
      for (long long t = 0; t + L < n; ++t) {
        const double zt = __dsub_rn(v[t * ld], mu);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = __fma_rn(zt, w[q], acc[q]);
#pragma unroll
        for (int q = 0; q < 7; ++q) w[q] = w[q + 1];
        w[7] = (t + L + 8 < n) ? __dsub_rn(v[(t + L + 8) * ld], mu) : 0.0;
      }
      if (L == 0) s0 = acc[0];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long j = L / 2 + q;
        if (!done && j <= k) {
          double g = __dadd_rn(__ddiv_rn(acc[2 * q], dn), __ddiv_rn(acc[2 * q + 1], dn));
          if (g <= 0.0) done = true;
          else {
            if (j > 0 && g > gprev) g = gprev;
            sumg = __dadd_rn(sumg, g);
            gprev = g;
          }
        }
      }
    }
    const double iidvar = __ddiv_rn(__ddiv_rn(s0, (double)(n - 1)), dn);
    const double acv0 = __ddiv_rn(s0, dn);
    const double mcvar = __ddiv_rn(__dadd_rn(-acv0, __dmul_rn(2.0, sumg)), dn);
    out = __ddiv_rn(__dmul_rn(dn, iidvar), mcvar);
    o_iid = iidvar; o_imse = mcvar; o_iact = __ddiv_rn(mcvar, iidvar);
  }
  stat_store(S, c * dim + i, o_mean, o_iid, o_imse, out, o_iact);
}
// Same estimator with the CTA's series staged once in shared memory: tile[t][TC] holds the centred samples of
// TC coordinates of one chain (npost x TC x 8 bytes).  The global-memory version above re-reads the values on
// every lag pass; with 16 CTAs per SM its working set does not fit L2 and the kernel is DRAM bound on ~5 x
// the stored bytes (profiles/r1_summary.md).  Here the values cross DRAM once.  Same sums in the same order.
template <int TC>
__global__ void __launch_bounds__(TC)
klb_ess_tile_kernel(const double* __restrict__ value, long long ld, long long npost, int dim, const KlbStatPtrs S) {
  extern __shared__ double tile[];                              // [npost][TC]
  const int i = blockIdx.x * TC + threadIdx.x;
  const long long c = blockIdx.y;
  const bool act = i < dim;
  const double* v = value + c * npost * ld + (act ? i : 0);
  const int n = (int)npost;
  // stage the column with 8-byte cp.async copies: all npost row segments are in flight at once (a plain
  // load / store loop keeps one or two loads in flight per warp and is latency bound on DRAM)
  if (act) {
    for (int t = 0; t < n; ++t) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(tile + t * TC + threadIdx.x);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(v + (long long)t * ld) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  double s = 0.0;
  if (act)
    for (int t = 0; t < n; ++t) s = __dadd_rn(s, tile[t * TC + threadIdx.x]);
  if (!act) return;                                              // every thread only ever reads its own column
  const double qnan = klb_u2d(0x7FF8000000000000ULL);
  double out = qnan, o_iid = qnan, o_imse = qnan, o_iact = qnan;
  const double o_mean = n >= 1 ? __ddiv_rn(s, (double)n) : qnan;
  if (n >= 4) {
    const double dn = (double)n;
    const double mu = o_mean;
    double* z = tile + threadIdx.x;
    for (int t = 0; t < n; ++t) z[t * TC] = __dsub_rn(z[t * TC], mu);
    const int k = (n - 2) / 2;
    double sumg = 0.0, gprev = 0.0, s0 = 0.0;
    bool done = false;
    for (int L = 0; !done && L <= 2 * k + 1; L += 8) {
      double acc[8], w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { acc[q] = 0.0; w[q] = (L + q < n) ? z[(L + q) * TC] : 0.0; }
      int t = 0;
      // main part, eight steps of t per trip while every index stays inside the series: the sixteen loads of a trip
      // are issued ahead of its 64 DFMA (the rolled loop waited ~30 cycles of shared-memory latency for z[t] in every
      // step, with two warps per scheduler to hide it), and the window z[t+L .. t+L+7] lives in w[(t + q) & 7] with
      // compile-time slot indices instead of being shifted (14 MOVs per step).  Same sums in the same order.
      for (; t + L + 15 < n; t += 8) {
        double zt[8], wn[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { zt[u] = z[(t + u) * TC]; wn[u] = z[(t + u + L + 8) * TC]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = __fma_rn(zt[u], w[(u + q) & 7], acc[q]);
          w[u] = wn[u];
        }
      }
      for (; t + L < n; ++t) {                 // tail (the window is back in slot order after whole trips)
        const double zt = z[t * TC];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = __fma_rn(zt, w[q], acc[q]);
#pragma unroll
        for (int q = 0; q < 7; ++q) w[q] = w[q + 1];
        w[7] = (t + L + 8 < n) ? z[(t + L + 8) * TC] : 0.0;
      }
      if (L == 0) s0 = acc[0];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = L / 2 + q;
        if (!done && j <= k) {
          double g = __dadd_rn(__ddiv_rn(acc[2 * q], dn), __ddiv_rn(acc[2 * q + 1], dn));
          if (g <= 0.0) done = true;
          else {
            if (j > 0 && g > gprev) g = gprev;
            sumg = __dadd_rn(sumg, g);
            gprev = g;
          }
        }
      }
    }
    const double iidvar = __ddiv_rn(__ddiv_rn(s0, (double)(n - 1)), dn);
    const double acv0 = __ddiv_rn(s0, dn);
    const double mcvar = __ddiv_rn(__dadd_rn(-acv0, __dmul_rn(2.0, sumg)), dn);
    out = __ddiv_rn(__dmul_rn(dn, iidvar), mcvar);
    o_iid = iidvar; o_imse = mcvar; o_iact = __ddiv_rn(mcvar, iidvar);
  }
  stat_store(S, c * dim + i, o_mean, o_iid, o_imse, out, o_iact);
}

// The same estimator with a window of WIN lags per pass and, with RAW, the samples left as they were staged (the
// centring pass over the tile -- a load, a subtraction and a store per sample -- is replaced by a subtraction where a
// sample is read: z_t = v_t - mean is the same double either way).  Why: with eight lags per pass the warp's slowest
// lane decides how many passes run (Geyer's rule stops at the first non-positive pair, and with 100 samples the
// estimated autocovariances are noisy: three or four passes per warp are common), and every pass reads the series
// twice from shared memory -- the kernel was bound by the shared-memory pipe, not by HBM or the fp64 pipe.  A pass over
// WIN lags costs the same 2 loads per step as a pass over 8.  Trips of WIN steps keep the window slots compile-time
// constants; the trips that reach the end of the series take predicated loads (an out-of-range sample counts as 0:
// its products are +-0 and leave the sums as they are).  Same sums in the same order as klb_ess_tile_kernel.
template <int TC, int WIN, bool RAW>
__global__ void __launch_bounds__(TC)
klb_ess_win_kernel(const double* __restrict__ value, long long ld, long long npost, int dim, const KlbStatPtrs S) {
  extern __shared__ double tile[];                              // [npost][TC]
  const int i = blockIdx.x * TC + threadIdx.x;
  const long long c = blockIdx.y;
  const bool act = i < dim;
  const double* v = value + c * npost * ld + (act ? i : 0);
  const int n = (int)npost;
  if (act) {
    for (int t = 0; t < n; ++t) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(tile + t * TC + threadIdx.x);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(v + (long long)t * ld) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  if (!act) return;                                              // every thread only ever reads its own column
  double* const z = tile + threadIdx.x;
  double s = 0.0;
  for (int t = 0; t < n; ++t) s = __dadd_rn(s, z[t * TC]);
  const double qnan = klb_u2d(0x7FF8000000000000ULL);
  double out = qnan, o_iid = qnan, o_imse = qnan, o_iact = qnan;
  const double o_mean = n >= 1 ? __ddiv_rn(s, (double)n) : qnan;
  if (n >= 4) {
    const double dn = (double)n;
    const double mu = o_mean;
    if (!RAW)
      for (int t = 0; t < n; ++t) z[t * TC] = __dsub_rn(z[t * TC], mu);
#define KLB_ESS_Z(idx) (RAW ? __dsub_rn(z[(idx) * TC], mu) : z[(idx) * TC])
    const int k = (n - 2) / 2;
    double sumg = 0.0, gprev = 0.0, s0 = 0.0;
    bool done = false;
    for (int L = 0; !done && L <= 2 * k + 1; L += WIN) {
      double acc[WIN], w[WIN];
#pragma unroll
      for (int q = 0; q < WIN; ++q) { acc[q] = 0.0; w[q] = (L + q < n) ? KLB_ESS_Z(L + q) : 0.0; }
      int t = 0;
      for (; t + L + 2 * WIN - 1 < n; t += WIN) {               // every index of the trip lies inside the series
#pragma unroll
        for (int h = 0; h < WIN; h += 8) {
          double zt[8], wn[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { zt[u] = KLB_ESS_Z(t + h + u); wn[u] = KLB_ESS_Z(t + h + u + L + WIN); }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int q = 0; q < WIN; ++q) acc[q] = __fma_rn(zt[u], w[(h + u + q) & (WIN - 1)], acc[q]);
            w[(h + u) & (WIN - 1)] = wn[u];
          }
        }
      }
      for (; t + L < n; t += WIN) {                             // the trips that reach the end of the series
#pragma unroll
        for (int h = 0; h < WIN; h += 8) {
          double zt[8], wn[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            zt[u] = (t + h + u < n) ? KLB_ESS_Z(t + h + u) : 0.0;
            wn[u] = (t + h + u + L + WIN < n) ? KLB_ESS_Z(t + h + u + L + WIN) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int q = 0; q < WIN; ++q) acc[q] = __fma_rn(zt[u], w[(h + u + q) & (WIN - 1)], acc[q]);
            w[(h + u) & (WIN - 1)] = wn[u];
          }
        }
      }
#undef KLB_ESS_Z
      if (L == 0) s0 = acc[0];
#pragma unroll
      for (int q = 0; q < WIN / 2; ++q) {
        const int j = L / 2 + q;
        if (!done && j <= k) {
          double g = __dadd_rn(__ddiv_rn(acc[2 * q], dn), __ddiv_rn(acc[2 * q + 1], dn));
          if (g <= 0.0) done = true;
          else {
            if (j > 0 && g > gprev) g = gprev;
            sumg = __dadd_rn(sumg, g);
            gprev = g;
          }
        }
      }
    }
    const double iidvar = __ddiv_rn(__ddiv_rn(s0, (double)(n - 1)), dn);
    const double acv0 = __ddiv_rn(s0, dn);
    const double mcvar = __ddiv_rn(__dadd_rn(-acv0, __dmul_rn(2.0, sumg)), dn);
    out = __ddiv_rn(__dmul_rn(dn, iidvar), mcvar);
    o_iid = iidvar; o_imse = mcvar; o_iact = __ddiv_rn(mcvar, iidvar);
  }
  stat_store(S, c * dim + i, o_mean, o_iid, o_imse, out, o_iact);
}

template <int TC, int WIN, bool RAW>
static bool launch_ess_win(const double* value, long long ld, long long npost, long long nchains, int dim,
                           const KlbStatPtrs& S, cudaStream_t s) {
  const size_t sm = (size_t)npost * TC * sizeof(double);
  if (sm > (size_t)(100 * 1024)) return false;                    // npost = 100: 2 CTAs of 128 series per SM, or 4 of 64, or 8 of 32
  auto kern = klb_ess_win_kernel<TC, WIN, RAW>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  dim3 grid((unsigned)((dim + TC - 1) / TC), (unsigned)nchains);
  kern<<<grid, TC, sm, s>>>(value, ld, npost, dim, S);
  return true;
}

template <int TC>
static bool launch_ess_tile(const double* value, long long ld, long long npost, long long nchains, int dim,
                            const KlbStatPtrs& S, cudaStream_t s) {
  const size_t sm = (size_t)npost * TC * sizeof(double);
  if (sm > 100 * 1024) return false;                              // two CTAs per SM
  auto kern = klb_ess_tile_kernel<TC>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  dim3 grid((unsigned)((dim + TC - 1) / TC), (unsigned)nchains);
  kern<<<grid, TC, sm, s>>>(value, ld, npost, dim, S);
  return true;
}

// stats[5] = {mean, mcvar_iid, mcvar_imse, ess, iact}: device arrays of nchains*dim doubles, null to skip
void klb_launch_stats(const double* value, long long ld, long long npost, long long nchains, int dim,
                      double* const stats[5], cudaStream_t s) {
  const KlbStatPtrs S = {stats[0], stats[1], stats[2], stats[3], stats[4]};
  // Default: one warp per CTA (32 series), sixteen lags per pass, raw tile.  Measured on the C3 output (65 536 x 1024 series of
  // 100 samples, one B200, tools/ess_variants.py; every variant bit-identical to the oracle): 128 series per CTA with 8 / 16 /
  // 32 lags per pass 34.2 / 33.0 / 54.6 ms (the time follows the number of lags evaluated, wasted ones included: 32 per pass
  // is too many), 64 series per CTA 30.5 ms, 32 series per CTA 29.1 ms: eight independent one-warp CTAs per SM overlap their
  // load, mean and lag phases better than two CTAs of four warps that move in step.
  // KLB_ESS_VARIANT selects the others (experiments): window of lags per pass / raw or centred tile / series per CTA.
  const char* ev = getenv("KLB_ESS_VARIANT");
  const int variant = ev ? atoi(ev) : KLB_ESS_DEFAULT_VARIANT;
  switch (variant) {
    case 1: if (launch_ess_win<128, 8, true>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 2: if (launch_ess_win<128, 16, false>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 3: if (launch_ess_win<128, 16, true>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 4: if (launch_ess_win<128, 32, false>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 5: if (launch_ess_win<128, 32, true>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 6: if (launch_ess_win<64, 16, true>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 7: if (launch_ess_win<32, 16, true>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 8: if (launch_ess_win<64, 32, true>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 9: if (launch_ess_win<32, 8, true>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 10: if (launch_ess_win<32, 16, false>(value, ld, npost, nchains, dim, S, s)) return; break;
    case 11: if (launch_ess_win<32, 8, false>(value, ld, npost, nchains, dim, S, s)) return; break;
    default: break;
  }
  if (launch_ess_tile<128>(value, ld, npost, nchains, dim, S, s)) return;
  if (launch_ess_tile<64>(value, ld, npost, nchains, dim, S, s)) return;
  if (launch_ess_tile<32>(value, ld, npost, nchains, dim, S, s)) return;
  dim3 grid((unsigned)((dim + 127) / 128), (unsigned)nchains);    // very long chains: stream from global memory
  klb_ess_kernel<<<grid, 128, 0, s>>>(value, ld, npost, dim, S);
}
void klb_launch_ess(const double* value, long long ld, long long npost, long long nchains, int dim, double* ess,
                    cudaStream_t s) {
  double* const stats[5] = {nullptr, nullptr, nullptr, ess, nullptr};
  klb_launch_stats(value, ld, npost, nchains, dim, stats, s);
}

// acceptance(s; diagnostics=true): mean of the :accept diagnostic of one chain          src/stats/acceptance.jl:28-34
// acceptance(s; diagnostics=false): 1 + #{t >= 2 : value[:, t] != value[:, t-1]}, over n  src/stats/acceptance.jl:3-14,33
// One warp per chain; the value variant compares consecutive saved states coordinate by coordinate.
__global__ void __launch_bounds__(256)
klb_acceptance_kernel(const unsigned char* __restrict__ accept, const double* __restrict__ value, long long ld,
                      long long npost, long long nchains, int dim, double* __restrict__ out) {
  const long long c = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= nchains) return;
  long long cnt = 0;
  if (accept) {
    const unsigned char* a = accept + c * npost;
    for (long long t = lane; t < npost; t += 32) cnt += a[t] ? 1 : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  } else {
    const double* v = value + c * npost * ld;
    cnt = npost > 0 ? 1 : 0;
    for (long long t = 1; t < npost; ++t) {
      bool diff = false;
      for (int i = lane; i < dim; i += 32) diff |= (v[t * ld + i] != v[(t - 1) * ld + i]);
      if (__any_sync(0xffffffffu, diff)) cnt += 1;
    }
  }
  if (lane == 0) out[c] = __ddiv_rn((double)cnt, (double)npost);
}
void klb_launch_acceptance(const unsigned char* accept, const double* value, long long ld, long long npost,
                           long long nchains, int dim, double* out, cudaStream_t s) {
  klb_acceptance_kernel<<<(unsigned)((nchains + 7) / 8), 256, 0, s>>>(accept, value, ld, npost, nchains, dim, out);
}

// ------------------------------------------------------------------ synthetic initial state
// x0[i, c] = N(0,1) of stream (seed, global chain c, transition 0, element i): the initial value of every benchmark
// configuration, a function of the GLOBAL chain index only, so that a job's input does not depend on how its chains
// are sharded over GPUs.  One thread per double2 unit (klb_normal: the scalar procedure the oracle calls).
__global__ void klb_synth_state_kernel(const uint64_t* gtab, uint64_t seed, uint64_t chain_offset, long long nchains,
                                       int dim, long long ld, double* state) {
  __shared__ uint64_t tab[KLB_TAB_LEN];
  for (int i = threadIdx.x; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  const long long upc = ld / 2;                                // units per chain (ld is even)
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < nchains * upc; u += (long long)gridDim.x * blockDim.x) {
    const long long c = u / upc;
    const int i = 2 * (int)(u - c * upc);
    const klb_stream st = klb_stream_make(seed, chain_offset + (uint64_t)c, 0ull);
    uint64_t w0, w1;
    klb_stream_draw(&st, (uint32_t)(i >> 1), KLB_TAG_NORMAL, 0u, &w0, &w1);
    double2 v;
    v.x = klb_normal_from_word(w0, (uint32_t)i, &st, tab);
    v.y = (i + 1 < dim) ? klb_normal_from_word(w1, (uint32_t)i + 1u, &st, tab) : 0.0;   // pad row of an odd dim stays 0
    *reinterpret_cast<double2*>(state + c * ld + i) = v;
  }
}
void klb_launch_synth_state(const uint64_t* tab, uint64_t seed, uint64_t chain_offset, long long nchains, int dim,
                            long long ld, double* state, cudaStream_t s) {
  klb_synth_state_kernel<<<148 * 8, 256, 0, s>>>(tab, seed, chain_offset, nchains, dim, ld, state);
}

// ------------------------------------------------------------------ measured peaks (roofline denominators)
// MEASURED_PEAKS.json holds HBM and bf16 figures only; the kernels of this library are bound by the fp64 pipe
// (DADD / DMUL / DFMA: one warp instruction per two cycles per scheduler) or by the fp64 tensor pipe (DMMA), so
// bench.py measures both denominators live with these two streams of independent instructions.
__global__ void __launch_bounds__(256) klb_peak_fp64_kernel(double* out, int iters, double a, double b) {
  double v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = 1e-3 * (threadIdx.x + q);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __fma_rn(v[q], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += v[q];
  if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) klb_peak_dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int q = 0; q < 8; ++q) { c[q][0] = 1e-3 * (threadIdx.x + q); c[q][1] = 2e-3 * q; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
  if (s == 12345.678) out[0] = s;
}
// kind 0: fp64 results per second (one per lane and DFMA / DADD / DMUL instruction); kind 1: DMMA flop per second
int klb_measure_peak(int kind, double* per_second) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return -1;
  double* out = nullptr;
  if (cudaMalloc(&out, 8) != cudaSuccess) return -1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = sms * 8, iters = kind == 0 ? 4000 : 2000;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, 0);
    if (kind == 0) klb_peak_fp64_kernel<<<grid, 256>>>(out, iters, 0.999999, 1e-9);
    else klb_peak_dmma_kernel<<<grid, 256>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, 0);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = 0.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double work = kind == 0 ? (double)grid * 256 * iters * 64.0                      // 64 DFMA per thread and iteration
                                  : (double)grid * 8 * iters * 32.0 * 512.0;               // 32 DMMA (8x8x4: 512 flop) per warp
    if (rep > 0 && ms > 0.f && work / (ms * 1e-3) > best) best = work / (ms * 1e-3);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  if (cudaGetLastError() != cudaSuccess || best == 0.0) return -1;
  *per_second = best;
  return 0;
}
