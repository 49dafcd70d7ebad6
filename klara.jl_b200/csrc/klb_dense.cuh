// klb_dense.cuh -- chain kernels for the dense-precision Gaussian target
//     logtarget(z) = -z'Cz,  gradlogtarget(z) = -2 C z      (C symmetric, d x d, shared by all chains)
// reference: doc/examples/BivariateNormal/MALA/function/analytical.jl:8-9 (closures fed by the
// Hyperparameter C through parameter.states, BasicContMuvParameter.jl:497-501); BASELINE config C4.
//
// The elementwise kernels of klb_kernels.cuh do not apply: every gradient evaluation is a d x d matrix-vector
// product per chain, i.e. a GEMM over the chains of a CTA.  Layout of one CTA (256 threads, MC chains):
//   * thread t owns elements 2t, 2t+1 of EVERY chain of the CTA (p, C z and the cached gradient of MC chains
//     in registers: a 2 x MC register tile);
//   * the positions of the MC chains live in shared memory, xs[j*MC + r] (r fastest), so that one
//     broadcast LDS.128 delivers x_j of two chains to the whole CTA;
//   * (C z)_i = sum_j C[j][i] z_j is accumulated by fma in increasing j -- each thread streams the two
//     columns it owns (C symmetric: column i = row i; row j is read coalesced by the CTA, 4 KB per row at
//     d = 512) and applies every loaded C value to all MC chains: MC x 2 DFMA per 16-byte load;
//   * reductions use the canonical order of DESIGN.md: addends go to a shared-memory scratch array and one
//     warp per chain reduces them exactly like the one-warp-per-chain kernels (and the oracle) do.
// The accumulation order (increasing j, one fma per term) is the oracle's, so parity stays bit-exact; on
// B200 the fp64 tensor pipe (DMMA) has the same peak as the DFMA pipe, so nothing is lost by not using it.
#pragma once
#include "klb_kernels.cuh"

#define KLB_DENSE_THREADS 256
#define KLB_DENSE_MAXD 512  /* 2 elements per thread */
// Odd dim: rows of C on the device and the per-element arrays in shared memory are padded to the even length
// dl = dim + 1.  The pad column of C is zero, the pad element of every chain is +0.0 and never moves (its momentum /
// proposal noise is 0, not a draw; (C z)_pad = 0), and the canonical reductions skip it (i + 1 < dim), so every sum
// has exactly the oracle's addends.
__host__ __device__ __forceinline__ int klb_dense_dl(int d) { return (d + 1) & ~1; }

template <int MC>
struct DenseShared {
  double lt_cur[MC], lt_new[MC], k0[MC], k1[MC], s1[MC], s2[MC], step[MC];
  long long accepted[MC], proposed[MC], totproposed[MC];
  double rate[MC];
  int accept[MC];
  // DualAveragingMCTuner: the chain's leapfrog count of this transition and a = min(1, exp(ratio))
  int nl[MC];
  double aprob[MC];
};

// canonical reduction of one chain's addends by one warp.  MODE 0: dot product sum of a[i]*b[i]
// (fma-accumulated, see dotacc); MODE 1: plain sum of a[i].
template <int MC, bool FMA, int MODE>
__device__ __forceinline__ double dense_reduce(const double* a, const double* b, int r, int d, int nv, int lane) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int m = 0; m < nv; ++m) {
    const int i = 2 * (lane + 32 * m);
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    if (i < d) { a0 = a[(size_t)i * MC + r]; if (MODE == 0) b0 = b[(size_t)i * MC + r]; }
    if (i + 1 < d) { a1 = a[(size_t)(i + 1) * MC + r]; if (MODE == 0) b1 = b[(size_t)(i + 1) * MC + r]; }
    if (MODE == 0) {
      acc[m & 3] = dotacc(a0, b0, acc[m & 3]);
      acc[m & 3] = dotacc(a1, b1, acc[m & 3]);
    } else {
      acc[m & 3] = __dadd_rn(acc[m & 3], a0);
      acc[m & 3] = __dadd_rn(acc[m & 3], a1);
    }
  }
  double v = __dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3]));
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

// acc[r][0..1] = (C x_r)[2t, 2t+1] for the MC chains whose positions are in xs.
// C is streamed from L2 with plain read-only loads (4 rows in flight per thread).  A private cp.async ring in
// shared memory (16 rows in flight) was measured too: it removes the long-scoreboard stalls but adds an
// LDGSTS + LDS pair per row to an L1/shared pipe that is already the co-bottleneck (8 wavefronts per
// 16 DFMA), and was 10 % slower (profiles/r1_summary.md, dense section).
template <int MC>
__device__ __forceinline__ void dense_matvec(double (&acc)[MC][2], const double* __restrict__ Cm, const double* xs,
                                             int d, int i0, bool active) {
  const int dl = klb_dense_dl(d);
#pragma unroll
  for (int r = 0; r < MC; ++r) { acc[r][0] = 0.0; acc[r][1] = 0.0; }
  if (!active) return;
  const double* col = Cm + i0;
#pragma unroll 4
  for (int j = 0; j < d; ++j) {
    const double2 c = __ldg(reinterpret_cast<const double2*>(col + (size_t)j * dl));
    const double2* xj = reinterpret_cast<const double2*>(xs + (size_t)j * MC);
#pragma unroll
    for (int r2 = 0; r2 < MC / 2; ++r2) {
      const double2 xv = xj[r2];
      acc[2 * r2][0] = __fma_rn(c.x, xv.x, acc[2 * r2][0]);
      acc[2 * r2][1] = __fma_rn(c.y, xv.x, acc[2 * r2][1]);
      acc[2 * r2 + 1][0] = __fma_rn(c.x, xv.y, acc[2 * r2 + 1][0]);
      acc[2 * r2 + 1][1] = __fma_rn(c.y, xv.y, acc[2 * r2 + 1][1]);
    }
  }
}

struct DArgs {
  KArgs k;
  const double* Cm;  // d rows of dl = klb_dense_dl(d) values, row-major, symmetric
  int nv;            // canonical reduction units per lane for this dim
};

// DA (HMC with DualAveragingMCTuner): every chain has its own step AND its own number of leapfrog steps
// (iterate/HMC.jl:142-144).  The CTA runs max(nl) steps; a chain that has done its nl steps stops moving (its updates
// are predicated off; the matrix-vector product of the tile recomputes the same C x for it), so each chain sees exactly
// its own trajectory.
template <int SAMPLER, int MC, bool FMA, bool DA = false>
__global__ void __launch_bounds__(KLB_DENSE_THREADS)
klb_dense_kernel(const DArgs D) {
  static_assert(!DA || SAMPLER == 2, "DualAveragingMCTuner tunes HMC");
  const KArgs& A = D.k;
  extern __shared__ __align__(16) unsigned char dsm[];
  const int d = (int)A.dim, dl = klb_dense_dl(d);
  // dynamic shared memory: tab | xs[dl*MC] | sc[dl*MC] | sc2[dl*MC] | DenseShared
  uint64_t* tab = reinterpret_cast<uint64_t*>(dsm);
  double* xs = reinterpret_cast<double*>(dsm + ((KLB_TAB_LEN * 8 + 15) & ~15));
  double* sc = xs + (size_t)dl * MC;
  double* sc2 = sc + (size_t)dl * MC;
  DenseShared<MC>& S = *reinterpret_cast<DenseShared<MC>*>(sc2 + (size_t)dl * MC);

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = A.tab[i];
  const long long c0 = (long long)blockIdx.x * MC;
  const int i0 = 2 * t;
  const bool active = i0 < d;                     // this thread owns elements i0 and (unless it is the pad) i0 + 1
  const bool pad_b = i0 + 1 >= d;
  const double* Cm = D.Cm;

  // per-chain scalars
  if (t < MC) {
    const long long c = c0 + t;
    const bool live = c < A.nchains;
    S.lt_cur[t] = live ? A.lt[c] : 0.0;
    S.step[t] = live ? A.tune_step[c] : 1.0;
    S.accepted[t] = live ? A.tune_cnt[3 * c] : 0;
    S.proposed[t] = live ? A.tune_cnt[3 * c + 1] : 0;
    S.totproposed[t] = live ? A.tune_cnt[3 * c + 2] : 0;
    S.rate[t] = live ? A.tune_rate[c] : 0.0;
    if (DA) S.nl[t] = live ? da_nleaps(A, c, S.step[t]) : 0;
  }
  // positions -> registers and shared memory; chains beyond nchains are padded with zeros
  double x[MC][2], gc[MC][2];   // current position, cached raw C x of the current position
#pragma unroll
  for (int r = 0; r < MC; ++r) {
    double2 v = make_double2(0.0, 0.0);
    if (active && c0 + r < A.nchains) v = *reinterpret_cast<const double2*>(A.state + (c0 + r) * A.ld + i0);
    x[r][0] = v.x; x[r][1] = v.y;
    if (active) { xs[(size_t)i0 * MC + r] = v.x; xs[(size_t)(i0 + 1) * MC + r] = v.y; }
  }
  __syncthreads();
  dense_matvec<MC>(gc, Cm, xs, d, i0, active);   // gradient cache of the starting point (= uptogradlogtarget!)
  __syncthreads();   // every thread has read xs before the first MALA / MH proposal overwrites it (racecheck, round 2:
                     // the later transitions are separated by the barrier that closes a transition, the first was not)

  const bool saving = (A.out_value != nullptr) || (A.out_lt != nullptr) || (A.out_grad != nullptr) ||
                      (A.out_accept != nullptr);
  long long count = A.count0;
  long long thin = (A.i0 > A.burnin) ? klb_mod(A.i0 - A.burnin - 1, A.thinning) : 0;

  for (long long it = 0; it < A.nt; ++it) {
    const long long irun = A.i0 + it;
    const unsigned long long tglob = A.t0 + 1ull + (unsigned long long)it;
    double y[MC][2], acc[MC][2];    // HMC: momentum, MALA/MH: proposal ; acc = raw C * (proposal position)
    double xr[MC][2];               // HMC integrates x in place: pre-transition copy for the reject branch
#pragma unroll
    for (int r = 0; r < MC; ++r) { xr[r][0] = x[r][0]; xr[r][1] = x[r][1]; }

    // ---------------- normals: unit k = t of every chain
#pragma unroll
    for (int r = 0; r < MC; ++r) {
      y[r][0] = 0.0; y[r][1] = 0.0;
      if (active) {
        const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)(c0 + r), tglob);
        uint64_t w0, w1;
        klb_stream_draw(&st, (unsigned)t, KLB_TAG_NORMAL, 0u, &w0, &w1);
        double a, b;
        if (!klb_zig_fast(w0, tab, &a)) a = klb_normal_from_word(w0, (unsigned)i0, &st, tab);
        if (pad_b) b = 0.0;
        else if (!klb_zig_fast(w1, tab, &b)) b = klb_normal_from_word(w1, (unsigned)i0 + 1u, &st, tab);
        y[r][0] = a; y[r][1] = b;
      }
    }

    if (SAMPLER == 2) {
      // ------------------------------------------------------------ HMC
      // old kinetic energy
      if (active) {
#pragma unroll
        for (int r = 0; r < MC; ++r) { sc[(size_t)i0 * MC + r] = y[r][0]; sc[(size_t)(i0 + 1) * MC + r] = y[r][1]; }
      }
      __syncthreads();
      for (int r = warp; r < MC; r += KLB_DENSE_THREADS / 32) {
        const double v = dense_reduce<MC, FMA, 0>(sc, sc, r, d, D.nv, lane);
        if (lane == 0) S.k0[r] = v;
      }
      // leapfrog: the cached gradient of the current point opens the first step
#pragma unroll
      for (int r = 0; r < MC; ++r) { acc[r][0] = gc[r][0]; acc[r][1] = gc[r][1]; }
      int nlmax = A.nleaps;
      if (DA) {
        nlmax = 0;
#pragma unroll
        for (int r = 0; r < MC; ++r) nlmax = max(nlmax, S.nl[r]);
      }
      for (int s = 1; s <= nlmax; ++s) {
        __syncthreads();                       // xs readers of the previous matvec / reduction are done
#pragma unroll
        for (int r = 0; r < MC; ++r) {
          if (DA && s > S.nl[r]) continue;     // this chain's trajectory is complete
          const double step = S.step[r];
          const double h = __dmul_rn(0.5, step);
          const double ga = __dmul_rn(-2.0, acc[r][0]), gb = __dmul_rn(-2.0, acc[r][1]);
          y[r][0] = Ar<FMA>::ma(h, ga, y[r][0]);                 // p += (h g)
          y[r][1] = Ar<FMA>::ma(h, gb, y[r][1]);
          x[r][0] = Ar<FMA>::ma(step, y[r][0], x[r][0]);         // x += step p
          x[r][1] = Ar<FMA>::ma(step, y[r][1], x[r][1]);
          if (active) { xs[(size_t)i0 * MC + r] = x[r][0]; xs[(size_t)(i0 + 1) * MC + r] = x[r][1]; }
        }
        __syncthreads();
        dense_matvec<MC>(acc, Cm, xs, d, i0, active);            // g = -2 C x
#pragma unroll
        for (int r = 0; r < MC; ++r) {
          if (DA && s > S.nl[r]) continue;
          const double h = __dmul_rn(0.5, S.step[r]);
          y[r][0] = Ar<FMA>::ma(h, __dmul_rn(-2.0, acc[r][0]), y[r][0]);
          y[r][1] = Ar<FMA>::ma(h, __dmul_rn(-2.0, acc[r][1]), y[r][1]);
        }
      }
      // log-target of the proposal: -x.(C x); new kinetic energy
      if (active) {
#pragma unroll
        for (int r = 0; r < MC; ++r) {
          sc[(size_t)i0 * MC + r] = y[r][0]; sc[(size_t)(i0 + 1) * MC + r] = y[r][1];
          sc2[(size_t)i0 * MC + r] = acc[r][0]; sc2[(size_t)(i0 + 1) * MC + r] = acc[r][1];
        }
      }
      __syncthreads();
      for (int r = warp; r < MC; r += KLB_DENSE_THREADS / 32) {
        const double k1 = dense_reduce<MC, FMA, 0>(sc, sc, r, d, D.nv, lane);
        const double xcx = dense_reduce<MC, FMA, 0>(xs, sc2, r, d, D.nv, lane);
        if (lane == 0) {
          const long long c = c0 + r;
          const double lt_new = -xcx;
          const double oldh = __dsub_rn(S.lt_cur[r], __dmul_rn(0.5, S.k0[r]));
          const double newh = __dsub_rn(lt_new, __dmul_rn(0.5, k1));
          const double ratio = __dsub_rn(newh, oldh);
          const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)c, tglob);
          bool acc_ = false;
          double a_prob = 1.0;
          if (ratio >= 0.0) acc_ = true;
          else {
            const double ex = klb_exp(ratio, tab);
            const double a = (ex != ex) ? ex : (ex < 1.0 ? ex : 1.0);
            a_prob = a;
            acc_ = klb_accept_uniform(&st) < a;
          }
          S.lt_new[r] = lt_new;
          S.accept[r] = (acc_ && c < A.nchains) ? 1 : 0;
          if (DA) S.aprob[r] = a_prob;
        }
      }
    } else if (SAMPLER == 1) {
      // ------------------------------------------------------------ MALA
      double mu[MC][2];
#pragma unroll
      for (int r = 0; r < MC; ++r) {
        const double step = S.step[r];
        const double h = __dmul_rn(0.5, step), sq = __dsqrt_rn(step), hinv = __ddiv_rn(0.5, step);
        const DivBy by_step(step);
        const double ga = __dmul_rn(-2.0, gc[r][0]), gb = __dmul_rn(-2.0, gc[r][1]);
        mu[r][0] = Ar<FMA>::ma(h, ga, x[r][0]); mu[r][1] = Ar<FMA>::ma(h, gb, x[r][1]);
        const double ya = Ar<FMA>::ma(sq, y[r][0], mu[r][0]), yb = Ar<FMA>::ma(sq, y[r][1], mu[r][1]);
        y[r][0] = ya; y[r][1] = yb;
        const double da = __dsub_rn(mu[r][0], ya), db = __dsub_rn(mu[r][1], yb);
        const double ea = FMA ? __dmul_rn(__dmul_rn(da, hinv), da) : __dmul_rn(0.5, by_step(__dmul_rn(da, da)));
        const double eb = FMA ? __dmul_rn(__dmul_rn(db, hinv), db) : __dmul_rn(0.5, by_step(__dmul_rn(db, db)));
        if (active) {
          xs[(size_t)i0 * MC + r] = ya; xs[(size_t)(i0 + 1) * MC + r] = yb;       // proposal positions
          sc[(size_t)i0 * MC + r] = ea; sc[(size_t)(i0 + 1) * MC + r] = eb;
        }
      }
      __syncthreads();
      dense_matvec<MC>(acc, Cm, xs, d, i0, active);               // C y
      if (active) {
#pragma unroll
        for (int r = 0; r < MC; ++r) { sc2[(size_t)i0 * MC + r] = acc[r][0]; sc2[(size_t)(i0 + 1) * MC + r] = acc[r][1]; }
      }
      __syncthreads();
      for (int r = warp; r < MC; r += KLB_DENSE_THREADS / 32) {
        const double s1 = dense_reduce<MC, FMA, 1>(sc, sc, r, d, D.nv, lane);
        const double ycy = dense_reduce<MC, FMA, 0>(xs, sc2, r, d, D.nv, lane);
        if (lane == 0) { S.s1[r] = s1; S.lt_new[r] = -ycy; }
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < MC; ++r) {
        const double step = S.step[r];
        const double h = __dmul_rn(0.5, step), hinv = __ddiv_rn(0.5, step);
        const DivBy by_step(step);
        const double ga = __dmul_rn(-2.0, acc[r][0]), gb = __dmul_rn(-2.0, acc[r][1]);
        const double ma = Ar<FMA>::ma(h, ga, y[r][0]), mb = Ar<FMA>::ma(h, gb, y[r][1]);        // mu' = y + (h g(y))
        const double da = __dsub_rn(ma, x[r][0]), db = __dsub_rn(mb, x[r][1]);
        const double ea = FMA ? __dmul_rn(__dmul_rn(da, hinv), da) : __dmul_rn(0.5, by_step(__dmul_rn(da, da)));
        const double eb = FMA ? __dmul_rn(__dmul_rn(db, hinv), db) : __dmul_rn(0.5, by_step(__dmul_rn(db, db)));
        if (active) { sc[(size_t)i0 * MC + r] = ea; sc[(size_t)(i0 + 1) * MC + r] = eb; }
      }
      __syncthreads();
      for (int r = warp; r < MC; r += KLB_DENSE_THREADS / 32) {
        const double s2 = dense_reduce<MC, FMA, 1>(sc, sc, r, d, D.nv, lane);
        if (lane == 0) {
          const long long c = c0 + r;
          double ratio = __dsub_rn(S.lt_new[r], S.lt_cur[r]);
          ratio = __dadd_rn(ratio, S.s1[r]);
          ratio = __dsub_rn(ratio, s2);
          const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)c, tglob);
          const bool acc_ = (ratio > 0.0) || (ratio > klb_log(klb_accept_uniform(&st), tab));
          S.accept[r] = (acc_ && c < A.nchains) ? 1 : 0;
        }
      }
    } else {
      // ------------------------------------------------------------ MH (normal random walk)
      const double2 sg = active ? __ldg(reinterpret_cast<const double2*>(A.sigma + i0)) : make_double2(0.0, 0.0);
#pragma unroll
      for (int r = 0; r < MC; ++r) {
        y[r][0] = Ar<FMA>::ma(sg.x, y[r][0], x[r][0]);
        y[r][1] = Ar<FMA>::ma(sg.y, y[r][1], x[r][1]);
        if (active) { xs[(size_t)i0 * MC + r] = y[r][0]; xs[(size_t)(i0 + 1) * MC + r] = y[r][1]; }
      }
      __syncthreads();
      dense_matvec<MC>(acc, Cm, xs, d, i0, active);
      if (active) {
#pragma unroll
        for (int r = 0; r < MC; ++r) { sc2[(size_t)i0 * MC + r] = acc[r][0]; sc2[(size_t)(i0 + 1) * MC + r] = acc[r][1]; }
      }
      __syncthreads();
      for (int r = warp; r < MC; r += KLB_DENSE_THREADS / 32) {
        const double ycy = dense_reduce<MC, FMA, 0>(xs, sc2, r, d, D.nv, lane);
        if (lane == 0) {
          const long long c = c0 + r;
          S.lt_new[r] = -ycy;
          const double ratio = __dsub_rn(-ycy, S.lt_cur[r]);
          const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)c, tglob);
          const bool acc_ = (ratio > 0.0) || (ratio > klb_log(klb_accept_uniform(&st), tab));
          S.accept[r] = (acc_ && c < A.nchains) ? 1 : 0;
        }
      }
    }

    // ---------------- per-chain epilogue: counters, tuner (one thread per chain)
    __syncthreads();
    if (t < MC) {
      const int r = t;
      if (A.counters_on) { S.proposed[r] += 1; if (S.accept[r]) S.accepted[r] += 1; }
      if (S.accept[r]) S.lt_cur[r] = S.lt_new[r];
      Tune tn;
      tn.step = S.step[r]; tn.accepted = S.accepted[r]; tn.proposed = S.proposed[r]; tn.totproposed = S.totproposed[r];
      tn.rate = S.rate[r];
      if (DA) {
        if (c0 + r < A.nchains) {
          da_block<false>(A, c0 + r, tn, S.nl[r], S.aprob[r], tab, true);
          S.nl[r] = da_nleaps(A, c0 + r, tn.step);                 // of the next transition
        }
      } else tuner_block<SAMPLER>(A, tn, tab, c0 + r);
      S.step[r] = tn.step; S.accepted[r] = tn.accepted; S.proposed[r] = tn.proposed; S.totproposed[r] = tn.totproposed;
      S.rate[r] = tn.rate;
    }
    // ---------------- accept / reject, state and output
    const bool do_save = (irun > A.burnin) && (thin == 0);
#pragma unroll
    for (int r = 0; r < MC; ++r) {
      if (S.accept[r] != 0) {
        if (SAMPLER != 2) { x[r][0] = y[r][0]; x[r][1] = y[r][1]; }   // HMC: x already is the end of the trajectory
        gc[r][0] = acc[r][0]; gc[r][1] = acc[r][1];
      } else if (SAMPLER == 2) {
        x[r][0] = xr[r][0]; x[r][1] = xr[r][1];
      }
    }
    __syncthreads();   // S.lt_cur / S.step visible to everyone before the next transition
    if (do_save && saving && active) {
#pragma unroll
      for (int r = 0; r < MC; ++r) {
        const long long c = c0 + r;
        if (c < A.nchains) {
          const long long col = c * A.npost + count;
          if (A.out_value) *reinterpret_cast<double2*>(A.out_value + col * A.ld + i0) = make_double2(x[r][0], x[r][1]);
          if (A.out_grad)
            *reinterpret_cast<double2*>(A.out_grad + col * A.ld + i0) =
                make_double2(__dmul_rn(-2.0, gc[r][0]), __dmul_rn(-2.0, gc[r][1]));
          if (t == 0) {
            if (A.out_lt) A.out_lt[col] = S.lt_cur[r];
            if (A.out_accept) A.out_accept[col] = (unsigned char)S.accept[r];
          }
        }
      }
    }
    if (irun > A.burnin) {
      if (thin == 0) count += 1;
      thin = (thin + 1 == A.thinning) ? 0 : thin + 1;
    }
  }

  // final state
#pragma unroll
  for (int r = 0; r < MC; ++r) {
    const long long c = c0 + r;
    if (active && c < A.nchains) *reinterpret_cast<double2*>(A.state + c * A.ld + i0) = make_double2(x[r][0], x[r][1]);
  }
  if (t < MC && c0 + t < A.nchains) {
    const long long c = c0 + t;
    A.lt[c] = S.lt_cur[t];
    A.tune_step[c] = S.step[t];
    A.tune_cnt[3 * c] = S.accepted[t]; A.tune_cnt[3 * c + 1] = S.proposed[t]; A.tune_cnt[3 * c + 2] = S.totproposed[t];
    A.tune_rate[c] = S.rate[t];
  }
}

// initialize!: lt[c] = -x.(C x), finiteness of the log-target and of the gradient -2 C x
template <int MC, bool FMA>
__global__ void __launch_bounds__(KLB_DENSE_THREADS)
klb_dense_init_kernel(const DArgs D, int check_grad, unsigned long long* flag) {
  const KArgs& A = D.k;
  extern __shared__ __align__(16) unsigned char dsm[];
  const int d = (int)A.dim, dl = klb_dense_dl(d);
  double* xs = reinterpret_cast<double*>(dsm);
  double* sc2 = xs + (size_t)dl * MC;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long c0 = (long long)blockIdx.x * MC;
  const int i0 = 2 * t;
  const bool active = i0 < d;
#pragma unroll
  for (int r = 0; r < MC; ++r) {
    double2 v = make_double2(0.0, 0.0);
    if (active && c0 + r < A.nchains) v = *reinterpret_cast<const double2*>(A.state + (c0 + r) * A.ld + i0);
    if (active) { xs[(size_t)i0 * MC + r] = v.x; xs[(size_t)(i0 + 1) * MC + r] = v.y; }
  }
  __syncthreads();
  double acc[MC][2];
  dense_matvec<MC>(acc, D.Cm, xs, d, i0, active);
  unsigned bad = 0u;
#pragma unroll
  for (int r = 0; r < MC; ++r) {
    if (active) { sc2[(size_t)i0 * MC + r] = acc[r][0]; sc2[(size_t)(i0 + 1) * MC + r] = acc[r][1]; }
    if (check_grad && active && !(isfinite(acc[r][0]) && isfinite(acc[r][1]))) bad |= 1u << r;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < MC; ++r)
    if (((bad >> r) & 1u) && c0 + r < A.nchains) atomicMin(flag, (unsigned long long)(A.chain_offset + c0 + r + 1));
  for (int r = warp; r < MC; r += KLB_DENSE_THREADS / 32) {
    const double xcx = dense_reduce<MC, FMA, 0>(xs, sc2, r, d, D.nv, lane);
    if (lane == 0 && c0 + r < A.nchains) {
      A.lt[c0 + r] = -xcx;
      if (!isfinite(xcx)) atomicMin(flag, (unsigned long long)(A.chain_offset + c0 + r + 1));
    }
  }
}

#define KLB_DENSE_MC 8
int klb_dense_launch(const DArgs& D, int sampler, int fma, cudaStream_t s);
int klb_dense_init(const DArgs& D, int fma, int check_grad, unsigned long long* flag, cudaStream_t s);
int klb_dense_attrs(int sampler, int fma, int da, int dim, int* regs, int* bps);
