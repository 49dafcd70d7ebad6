// klb_hmc_ws.cuh -- warp-specialised HMC chain kernel (W = 1: one chain per consumer warp, dim 257..1024;
// W = 4: one chain per consumer warpgroup, dim 1025..4096).
//
// The fused kernel of klb_kernels.cuh leaves the fp64 pipe ~58 % busy: every warp alternates between an
// fp64-dense leapfrog phase and integer / latency-bound phases (Philox + ziggurat, the slow-path pass, the
// accept test), and with 244 registers only two warps per scheduler are resident to cover for each other
// (profiles/r1_summary.md).  Here the two kinds of work live in different warps of the same CTA:
//   * one warpgroup of consumers (setmaxnreg.inc): one chain each -- positions and momenta in registers,
//     leapfrog, reductions, Metropolis test, tuner, stores: almost pure fp64 issue;
//   * one warpgroup of producers (setmaxnreg.dec): producer s generates the next transition's momentum
//     (Philox4x32-7 + ziggurat, slow path included) and accept uniform of consumer s into shared memory.
// Producer s and consumer s meet at two named barriers (FULL / EMPTY, 64 threads each).  Because warp w and
// warp 4+w share a scheduler (the producers are warps 0-3, the consumers warps 4-7), every scheduler holds
// 2 consumers + 2 producers (2 CTAs per SM): one consumer's tail overlaps the other's leapfrog, and the producers'
// integer instructions run in between (measured cost and ceiling: profiles/r1_summary.md, fp64 pipe microbenchmarks).
// The counter-based RNG makes this legal: a draw depends on (seed, chain, transition) only.
// Results are bit-identical to klb_chain_kernel (same per-element operations, same reduction order).
//
// W = 4 (one chain per CTA): consumer w owns the units j*4 + w of the chain and is fed by producer w, exactly like a W = 1
// pair; the four consumers meet once per transition at the exchange of the lane accumulators (team_allsum, named barrier
// over the 128 consumer threads, double-buffered).  After it every consumer holds the same three sums, so each evaluates
// the Metropolis test and the tuner for itself on its own copy of the chain's scalars -- no broadcast, no second barrier
// (the dual-averaging record lives in global memory: there the four read, meet, and consumer 0 writes).
#pragma once
#include <type_traits>
#include "klb_kernels.cuh"

#ifndef KLB_WS_EXP
#define KLB_WS_EXP 0
#endif
// Register split of the two warpgroups (4 x consumer + 4 x producer = 1024 per lane slot, 2 CTAs per SM).  With the
// per-chain scalars parked in shared memory the consumers need < 184; the producers spill at 64 (a local-memory reload
// at the head of every Philox iteration) and do not at 72.  Measured on C3 (reference / fma arithmetic):
// 192/64 13.89 / 12.27 ms, 184/72 13.61 / 11.42 ms, 176/80 13.85 / 11.31 ms per 40 transitions.
#ifndef KLB_WS_CONSUMER_REGS
#define KLB_WS_CONSUMER_REGS 184
#endif
#ifndef KLB_WS_PRODUCER_REGS
#define KLB_WS_PRODUCER_REGS 72
#endif

__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// One inner leapfrog step of the isotropic target, x += step p; p += 2 (c x) with c = -2h (see TgtIso::kick),
// software-pipelined by hand: the five operations of an element form a dependent chain (fp64 latency ~8 cycles,
// issue every 2), and with only a handful of free registers ptxas used to emit the chain of one element back to
// back (DMUL t; DADD p,t; DADD p,t), leaving the warp stalled ~6 cycles per link whenever the partner consumer
// could not fill in.  Here each phase runs over a block of KLB_WS_BLK elements before the next phase starts, so
// dependent instructions are KLB_WS_BLK issue slots apart.  Same operations per element, same bits.
#ifndef KLB_WS_BLK
#define KLB_WS_BLK 4
#endif
template <bool FMA, int NE>
__device__ __forceinline__ void iso_leap_step(double (&x)[NE], double (&y)[NE], double step, double c) {
  constexpr int B = KLB_WS_BLK > 0 ? KLB_WS_BLK : 1;   // (KLB_WS_BLK = 0 never gets here)
#pragma unroll
  for (int b = 0; b < NE; b += B) {
    if (FMA) {
#pragma unroll
      for (int k = 0; k < B; ++k) x[b + k] = __fma_rn(step, y[b + k], x[b + k]);
#pragma unroll
      for (int k = 0; k < B; ++k) y[b + k] = __fma_rn(c, x[b + k], y[b + k]);
#pragma unroll
      for (int k = 0; k < B; ++k) y[b + k] = __fma_rn(c, x[b + k], y[b + k]);
    } else {
      double t[B];
#pragma unroll
      for (int k = 0; k < B; ++k) t[k] = __dmul_rn(step, y[b + k]);
#pragma unroll
      for (int k = 0; k < B; ++k) x[b + k] = __dadd_rn(t[k], x[b + k]);
#pragma unroll
      for (int k = 0; k < B; ++k) t[k] = __dmul_rn(c, x[b + k]);
#pragma unroll
      for (int k = 0; k < B; ++k) y[b + k] = __dadd_rn(y[b + k], t[k]);
#pragma unroll
      for (int k = 0; k < B; ++k) y[b + k] = __dadd_rn(y[b + k], t[k]);
    }
  }
}

// The same for the shifted target, gradient -2 (x - mu): x += step p; v = x - mu; p += 2 (c v), c = -2h.  (h (-2 v) and
// (-2h) v are the same real product, rounded once: scaling by 2 is exact.)  Without the blocking ptxas serialises the
// NV = 16 loop onto one temporary (tools/sass_depdist.py: 89 % of its fp64 instructions at distance 1).  mu comes from
// L1 (read-only path), one 16-byte load per unit and leapfrog step.  The caller passes `mu` plus a zero it has just read
// from shared memory through a volatile pointer: with a loop-invariant address ptxas hoists all 2 NV values of mu out of
// the leapfrog loop, spills them (LDL in the loop) and the step collapses onto one temporary all the same.
template <bool FMA, int NV, int W>
__device__ __forceinline__ void shifted_leap_step(const double* __restrict__ mu, double (&x)[2 * NV], double (&y)[2 * NV],
                                                  double step, double c, int w, int lane) {
  constexpr int B = (KLB_WS_BLK >= 2 && KLB_WS_BLK % 2 == 0) ? KLB_WS_BLK : 4;
#pragma unroll
  for (int b = 0; b < 2 * NV; b += B) {
    double m[B];
#pragma unroll
    for (int k = 0; k < B; k += 2) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(mu + Geo<NV, W>::elem((b + k) / 2, w, lane)));   // padded: always in range
      m[k] = v.x; m[k + 1] = v.y;
    }
    if (FMA) {
#pragma unroll
      for (int k = 0; k < B; ++k) x[b + k] = __fma_rn(step, y[b + k], x[b + k]);
#pragma unroll
      for (int k = 0; k < B; ++k) m[k] = __dsub_rn(x[b + k], m[k]);
#pragma unroll
      for (int k = 0; k < B; ++k) y[b + k] = __fma_rn(c, m[k], y[b + k]);
#pragma unroll
      for (int k = 0; k < B; ++k) y[b + k] = __fma_rn(c, m[k], y[b + k]);
    } else {
      double t[B];
#pragma unroll
      for (int k = 0; k < B; ++k) t[k] = __dmul_rn(step, y[b + k]);
#pragma unroll
      for (int k = 0; k < B; ++k) x[b + k] = __dadd_rn(t[k], x[b + k]);
#pragma unroll
      for (int k = 0; k < B; ++k) m[k] = __dsub_rn(x[b + k], m[k]);
#pragma unroll
      for (int k = 0; k < B; ++k) t[k] = __dmul_rn(c, m[k]);
#pragma unroll
      for (int k = 0; k < B; ++k) y[b + k] = __dadd_rn(y[b + k], t[k]);
#pragma unroll
      for (int k = 0; k < B; ++k) y[b + k] = __dadd_rn(y[b + k], t[k]);
    }
  }
}

// What the targets read of the argument block.  For the shifted target `mu` carries a zero offset read through a volatile
// pointer, a different read for each phase of a transition (opening half-kick | inner steps | closing step and log-target):
// otherwise the loads of mu -- invariant over the transitions, the same addresses in every phase -- are hoisted / merged
// into 2 NV values that stay live (or get spilled) through the leapfrog loop.
struct WsArgs {
  const double* mu;
  double ra, rb, rscale;
};

struct WsChain {          // per consumer: the chain's scalars between trajectories
  double lt_cur, step, rate, u_acc;
  long long accepted, proposed, totproposed, count, thin;
};

#define KLB_WS_BAR_TEAM 9   /* W = 4: the four consumers of a chain */
extern __shared__ __align__(16) unsigned char klb_ws_dyn[];   // W = 4: the momentum stages (4 x NV x 32 double2 = 32 KB at NV = 16)

template <class T, int NV, int W, bool FMA, bool FULL>
__global__ void __launch_bounds__(256, 2)
klb_hmc_ws_kernel(const KArgs A) {
  static_assert(W == 1 || W == 4, "one chain per consumer warp or per consumer warpgroup");
  __shared__ WsChain wchain[4];
  __shared__ __align__(16) uint64_t tab[KLB_TAB_LEN + 512];   // + the sign-flipped ziggurat pairs (zig_build9)
  __shared__ double2 zstage[W == 1 ? 4 : 1][W == 1 ? NV * 32 : 1];
  __shared__ unsigned short zqueue[4][KLB_QCAP];
  __shared__ double uacc[4];
  __shared__ int ws_zero;                                      // see shifted_leap_step
  if (threadIdx.x == 0) ws_zero = 0;
  __shared__ double red[W == 1 ? 1 : 2][W == 1 ? 1 : 3 * 4 * 32];   // W = 4: exchange of the lane accumulators, by transition parity
  for (int i = threadIdx.x; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = A.tab[i];
  __syncthreads();
  zig_build9(tab);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = warp & 3;
#ifndef KLB_WS_PRODUCER_LOW
#define KLB_WS_PRODUCER_LOW 1
#endif
  // which warp-id range the producers take makes no measurable difference (profiles/r1_summary.md)
  const bool producer = KLB_WS_PRODUCER_LOW ? (warp < 4) : (warp >= 4);
  const int w = (W == 1) ? 0 : slot;               // warp within the chain's team
  const long long c = (W == 1) ? (long long)blockIdx.x * 4 + slot : (long long)blockIdx.x;
  const bool live = c < A.nchains;                 // a dead slot retires its consumer AND its producer
  const int d = (int)A.dim;
  double2* const zbuf = (W == 1) ? zstage[W == 1 ? slot : 0] : reinterpret_cast<double2*>(klb_ws_dyn) + slot * (NV * 32);
  const int bar_full = 1 + slot, bar_empty = 5 + slot;

  if (producer) {
    setmaxnreg_dec<KLB_WS_PRODUCER_REGS>();
    if (!live) return;
    for (long long it = 0; it < A.nt; ++it) {
      const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)c,
                                            A.t0 + 1ull + (unsigned long long)it);
      if (it > 0) bar_sync(bar_empty, 64);         // the consumer has taken the previous transition's draws
#if KLB_WS_EXP != 2      // timing experiment 2: producers idle (consumer-bound time; results meaningless)
      randn_stage<NV, W, FULL, true>(st, d, w, lane, tab, zbuf, zqueue[slot]);
#endif
      if (lane == 0) uacc[slot] = klb_accept_uniform(&st);
      bar_arrive(bar_full, 64);
    }
    return;
  }

  setmaxnreg_inc<KLB_WS_CONSUMER_REGS>();
  if (!live) return;
  double* const xcol = A.state + c * A.ld;
  double x[2 * NV];
  load_chain<NV, W, FULL>(x, xcol, d, w, lane);
  // Per-chain scalars that are only touched between trajectories (log-target, tuner record, output cursor, the
  // accept uniform) are parked in shared memory, so that during the leapfrog loop the registers hold x, p and
  // enough temporaries for the software-pipelined step (iso_leap_step).  Every lane of the warp computes the same
  // values, so lane 0 alone writes them back (W = 4: every consumer keeps its own, identical, copy).
  const bool writer = lane == 0 && w == 0;                                     // of the chain's records in global memory
  volatile WsChain* const cs = &wchain[slot];
  if (lane == 0) {
    cs->lt_cur = A.lt[c];
    cs->step = A.tune_step[c];
    cs->accepted = A.tune_cnt[3 * c]; cs->proposed = A.tune_cnt[3 * c + 1]; cs->totproposed = A.tune_cnt[3 * c + 2];
    cs->rate = A.tune_rate[c];
    cs->count = A.count0;
    cs->thin = (A.i0 > A.burnin) ? klb_mod(A.i0 - A.burnin - 1, A.thinning) : 0;
  }
  __syncwarp();
  const bool saving = (A.out_value != nullptr) || (A.out_lt != nullptr) || (A.out_grad != nullptr) ||
                      (A.out_accept != nullptr);
  const int nt = (int)A.nt;                                                  // klb_job_create: nsteps < 2^31

  for (int it = 0; it < nt; ++it) {
    double y[2 * NV];
    bar_sync(bar_full, 64);                                                  // momentum[:] = randn(d) is ready
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double2 v = zbuf[j * 32 + lane];
      y[2 * j] = v.x; y[2 * j + 1] = v.y;
    }
    if (lane == 0) cs->u_acc = uacc[slot];
    __syncwarp();                                                            // lane 0's store is ordered before every lane's read below
    if (it + 1 < nt) bar_arrive(bar_empty, 64);                              // buffer may be refilled

    const double step = cs->step;
    const double h = __dmul_rn(0.5, step);
    const int nl = (A.tuner == 2) ? da_nleaps(A, c, step) : A.nleaps;         // DualAveragingMCTuner: per chain
    double k0lane;                                                           // W = 4: the warp's one accumulator
    {
      double a0[4 / W] = {};
#pragma unroll
      for (int j = 0; j < NV; ++j) {                                         // old kinetic energy
        const int q = AccIdx<W>::local(j);
        a0[q] = dotacc(y[2 * j], y[2 * j], a0[q]);
        a0[q] = dotacc(y[2 * j + 1], y[2 * j + 1], a0[q]);
      }
      if (W == 1) k0lane = __dadd_rn(__dadd_rn(a0[0], a0[1 % (4 / W)]), __dadd_rn(a0[2 % (4 / W)], a0[3 % (4 / W)]));  // the lane value of team_allsum
      else {
        // One accumulator = one chain of 2 NV dependent fma.  Left in a register, the compiler sinks the whole chain below
        // the leapfrog loop and keeps the untouched momenta alive for it (+34 registers; the hand-pipelined step then
        // collapses to one temporary).  The volatile store of the team exchange pins it here.  Nobody reads this half of
        // `red` any more: see the note at the exchange.
        k0lane = a0[0];
        *static_cast<volatile double*>(&red[W == 1 ? 0 : (it & 1)][(W == 1 ? 0 : w) * 32 + lane]) = k0lane;
      }
    }
    constexpr bool kOpaqueMu = std::is_same<T, TgtShifted>::value && NV == 16;
    const WsArgs B0 = {A.mu + (kOpaqueMu ? *static_cast<volatile int*>(&ws_zero) : 0), A.ra, A.rb, A.rscale};
    // leapfrog! (src/samplers/samplers.jl:122-134): see klb_chain_kernel for the exact-rewrite notes
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = Geo<NV, W>::elem(j, w, lane);
      T::template kick<FMA, false>(B0, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], h, y[2 * j], y[2 * j + 1]);
    }
#if KLB_WS_EXP == 1      // timing experiment 1: consumers skip the inner leapfrog steps (producer-bound time)
    for (int s = nl; s < nl; ++s) {
#else
    for (int s = 1; s < nl; ++s) {
#endif
      if (KLB_WS_BLK > 0 && std::is_same<T, TgtIso>::value) {
        iso_leap_step<FMA, 2 * NV>(x, y, step, __dmul_rn(-2.0, h));
        continue;
      }
      if (KLB_WS_BLK > 0 && NV == 16 && std::is_same<T, TgtShifted>::value) {
        shifted_leap_step<FMA, NV, W>(A.mu + *static_cast<volatile int*>(&ws_zero), x, y, step, __dmul_rn(-2.0, h), w, lane);
        continue;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int i = Geo<NV, W>::elem(j, w, lane);
        x[2 * j] = Ar<FMA>::ma(step, y[2 * j], x[2 * j]);
        x[2 * j + 1] = Ar<FMA>::ma(step, y[2 * j + 1], x[2 * j + 1]);
        T::template kick<FMA, true>(B0, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], h, y[2 * j], y[2 * j + 1]);
      }
    }
    const WsArgs B1 = {A.mu + (kOpaqueMu ? *static_cast<volatile int*>(&ws_zero) : 0), A.ra, A.rb, A.rscale};
    double acc[2][4 / W] = {};
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = Geo<NV, W>::elem(j, w, lane);
      const int q = AccIdx<W>::local(j);
      x[2 * j] = Ar<FMA>::ma(step, y[2 * j], x[2 * j]);
      x[2 * j + 1] = Ar<FMA>::ma(step, y[2 * j + 1], x[2 * j + 1]);
      T::template kick<FMA, false>(B1, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], h, y[2 * j], y[2 * j + 1]);
      acc[0][q] = T::template lt_acc<FMA>(B1, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], acc[0][q]);
      acc[1][q] = dotacc(y[2 * j], y[2 * j], acc[1][q]);
      acc[1][q] = dotacc(y[2 * j + 1], y[2 * j + 1], acc[1][q]);
    }
    // canonical reduction (team_allsum) of (old kinetic energy, log-target sum, new kinetic energy)
    double sums[3];
    if (W == 1) {
      constexpr int M = 4 / W;
      sums[0] = k0lane;
      sums[1] = __dadd_rn(__dadd_rn(acc[0][0], acc[0][1 % M]), __dadd_rn(acc[0][2 % M], acc[0][3 % M]));
      sums[2] = __dadd_rn(__dadd_rn(acc[1][0], acc[1][1 % M]), __dadd_rn(acc[1][2 % M], acc[1][3 % M]));
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        double t[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) t[v] = __shfl_xor_sync(0xffffffffu, sums[v], o);
#pragma unroll
        for (int v = 0; v < 3; ++v) sums[v] = __dadd_rn(sums[v], t[v]);
      }
    } else {
      // a consumer that is a transition ahead writes the other half of `red`; it cannot get two ahead, because the
      // barrier of the transition in between needs the slowest consumer, which by then has read this half
      // team_allsum<3, 4> with the first value already in place: accumulator q of value v sits at red[(4 v + q) * 32 + lane]
      double* const rb = red[W == 1 ? 0 : (it & 1)];
      rb[(4 + w) * 32 + lane] = acc[0][0];
      rb[(8 + w) * 32 + lane] = acc[1][0];
      bar_sync(KLB_WS_BAR_TEAM, 128);
#pragma unroll
      for (int v = 0; v < 3; ++v)
        sums[v] = __dadd_rn(__dadd_rn(rb[(4 * v) * 32 + lane], rb[(4 * v + 1) * 32 + lane]),
                            __dadd_rn(rb[(4 * v + 2) * 32 + lane], rb[(4 * v + 3) * 32 + lane]));
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        double t[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) t[v] = __shfl_xor_sync(0xffffffffu, sums[v], o);
#pragma unroll
        for (int v = 0; v < 3; ++v) sums[v] = __dadd_rn(sums[v], t[v]);
      }
    }
    const double lt_new = T::lt_fin(B1, sums[1]);
    double lt_cur = cs->lt_cur;
    const double oldh = __dsub_rn(lt_cur, __dmul_rn(0.5, sums[0]));          // hamiltonian()
    const double newh = __dsub_rn(lt_new, __dmul_rn(0.5, sums[2]));
    const double ratio = __dsub_rn(newh, oldh);
    bool accept;
    double a_prob = 1.0;
    if (ratio >= 0.0) accept = true;                                         // min(1., exp(ratio)) = 1 > rand()
    else {
      const double ex = klb_exp(ratio, tab);
      const double a = (ex != ex) ? ex : (ex < 1.0 ? ex : 1.0);
      a_prob = a;
      accept = cs->u_acc < a;
    }
    __syncwarp();                  // every lane has read this transition's scalars (lt_cur, step, u_acc) before lane 0 rewrites them
    if (A.counters_on || A.tuner == 2) {                                     // tuner record: lives in shared memory
      Tune tn;
      tn.step = step; tn.accepted = cs->accepted; tn.proposed = cs->proposed; tn.totproposed = cs->totproposed;
      tn.rate = cs->rate;
      if (A.counters_on) { tn.proposed += 1; if (accept) tn.accepted += 1; }
      if (A.tuner == 2) da_block<true, (W == 1 ? 0 : KLB_WS_BAR_TEAM)>(A, c, tn, nl, a_prob, tab, writer);
      else tuner_block<2>(A, tn, tab, c);
      __syncwarp();
      if (lane == 0) {
        cs->step = tn.step; cs->accepted = tn.accepted; cs->proposed = tn.proposed; cs->totproposed = tn.totproposed;
        cs->rate = tn.rate;
      }
    }
    if (accept) {
      store_chain<NV, W, FULL>(x, xcol, d, w, lane);
      lt_cur = lt_new;
      if (lane == 0) cs->lt_cur = lt_new;
    } else {
      load_chain<NV, W, FULL>(x, xcol, d, w, lane);
    }
    if (A.i0 + it > A.burnin) {                                              // in(i, postrange) -> save
      const long long thin = cs->thin, count = cs->count;
      __syncwarp();
      if (thin == 0) {
        if (saving) {
          const long long col = c * A.npost + count;
          if (A.out_value) store_chain<NV, W, FULL>(x, A.out_value + col * A.ld, d, w, lane);
          if (A.out_grad) {
            double gbuf[2 * NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
              const int i = Geo<NV, W>::elem(j, w, lane);
              T::template grad<FMA>(B1, i, valid<FULL>(i, d), valid<FULL>(i + 1, d), x[2 * j], x[2 * j + 1], gbuf[2 * j], gbuf[2 * j + 1]);
              if (!valid<FULL>(i + 1, d)) gbuf[2 * j + 1] = 0.0;
              if (!valid<FULL>(i, d)) gbuf[2 * j] = 0.0;
            }
            store_chain<NV, W, FULL>(gbuf, A.out_grad + col * A.ld, d, w, lane);
          }
          if (writer) {
            if (A.out_lt) A.out_lt[col] = lt_cur;
            if (A.out_accept) A.out_accept[col] = accept ? 1 : 0;
          }
        }
        if (lane == 0) cs->count = count + 1;
      }
      if (lane == 0) cs->thin = (thin + 1 == A.thinning) ? 0 : thin + 1;
    }
    __syncwarp();
  }
  if (writer) {
    A.lt[c] = cs->lt_cur;
    A.tune_step[c] = cs->step;
    A.tune_cnt[3 * c] = cs->accepted; A.tune_cnt[3 * c + 1] = cs->proposed; A.tune_cnt[3 * c + 2] = cs->totproposed;
    A.tune_rate[c] = cs->rate;
  }
}
