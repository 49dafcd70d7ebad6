/* klb_math.h -- deterministic arithmetic primitives shared by the sm_100a
 * kernels and by the CPU oracle (which includes this file so that both sides
 * run the SAME source for everything that is not IEEE-exact by itself).
 *
 * Contents
 *   - bit casts, contraction-proof fp64 add/mul/fma/div/sqrt wrappers
 *   - Philox4x32 counter-based generator (Salmon et al., SC'11; 7 rounds, see klb_philox4x32_r) and the
 *     counter layout that replaces the reference's unseeded global RNG
 *     (reference draws: src/samplers/iterate/HMC.jl:135,165,
 *      src/samplers/iterate/MALA.jl:84,94, src/samplers/iterate/MH.jl:79,97)
 *   - uint64 -> uniform maps
 *   - exp / log built only from IEEE add/mul/fma and table lookups, so that
 *     gcc-on-x86 and nvcc-on-sm_100a return identical bits
 *   - 256-layer ziggurat standard normal (the reference's randn(d) is Julia's
 *     ziggurat; here it is keyed per (seed, chain, transition, element))
 *
 * Plain C99 / CUDA C++.  Compile host code with -ffp-contract=off.
 */
#ifndef KLB_MATH_H
#define KLB_MATH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define KLB_HD __host__ __device__ __forceinline__
/* the rare, bulky routines (exp, log, ziggurat slow path) are real calls on the device: inlining them
 * at every site blows the instruction cache of the chain kernels */
#define KLB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#include <math.h>
#include <string.h>
#define KLB_HD static inline
#define KLB_HD_NOINLINE static inline
#endif

/* ---------------------------------------------------------------- bit casts */
KLB_HD double klb_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}
KLB_HD uint64_t klb_d2u(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}

/* ------------------------------------------ contraction-proof fp64 operators */
KLB_HD double klb_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
KLB_HD double klb_sub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
KLB_HD double klb_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
KLB_HD double klb_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}
KLB_HD double klb_div(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
KLB_HD double klb_sqrt(double a) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(a);
#else
  return sqrt(a);
#endif
}
/* exact for |v| < 2^53 */
KLB_HD double klb_i2d(int64_t v) {
#if defined(__CUDA_ARCH__)
  return __ll2double_rn((long long)v);
#else
  return (double)v;
#endif
}

/* ------------------------------------------------------------ Philox4x32-10 */
#define KLB_PHILOX_M0 0xD2511F53u
#define KLB_PHILOX_M1 0xCD9E8D57u
#define KLB_PHILOX_W0 0x9E3779B9u
#define KLB_PHILOX_W1 0xBB67AE85u

KLB_HD uint32_t klb_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* out[0..3] = Philox4x32-R(counter c0..c3, key k0,k1): R rounds of the Philox4x32 bijection (Salmon, Moraes, Dror,
 * Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11).
 *
 * THE CONTRACT USES R = KLB_PHILOX_ROUNDS = 7.  Philox4x32-7 is the smallest round count the authors report as
 * Crush-resistant (it passes the complete TestU01 BigCrush battery; their Table 2); Random123 and cuRAND default to
 * 10 rounds as a safety margin.  Round 1 of this project used 10; ncu showed the HMC kernel bound by the issue slots
 * its 1024 normals per chain-transition take away from the fp64 pipe (every Philox instruction costs 1.3-1.8 cycles
 * of fp64 issue, profiles/r2_summary.md), and the RNG contract is this project's own (the reference draws from
 * Julia's unseeded global MersenneTwister), so round 2 moved to the 7-round generator: -30 % Philox instructions,
 * +7.4 % leapfrog steps/s.  The round function and key schedule are pinned by the Random123 known-answer vectors of
 * the 10-round generator, evaluated through the SAME code (klb_philox4x32_10 below; tests/test_oracle_kat.py,
 * tests/test_twin.py), and the statistical tests of the normals and uniforms run on the 7-round streams. */
#ifndef KLB_PHILOX_ROUNDS
#define KLB_PHILOX_ROUNDS 7
#endif
KLB_HD void klb_philox4x32_r(int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                             uint32_t k0, uint32_t k1, uint32_t out[4]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < rounds; ++r) {
#if defined(__CUDA_ARCH__)
    /* one IMAD.WIDE per product, opaque to the optimiser: left to itself nvcc strength-reduces the first round over
     * consecutive slots into IMAD.HI + IADD, and IMAD.HI is the most expensive integer instruction to issue next to
     * an fp64 stream (profiles/r1_summary.md, fp64_xwarp_test) */
    uint32_t hi0, lo0, hi1, lo1;
    asm("{\n\t.reg .b64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(KLB_PHILOX_M0));
    asm("{\n\t.reg .b64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(KLB_PHILOX_M1));
#else
    uint32_t hi0 = klb_mulhi32(KLB_PHILOX_M0, c0), lo0 = KLB_PHILOX_M0 * c0;
    uint32_t hi1 = klb_mulhi32(KLB_PHILOX_M1, c2), lo1 = KLB_PHILOX_M1 * c2;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += KLB_PHILOX_W0; k1 += KLB_PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* the contract's generator */
KLB_HD void klb_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  klb_philox4x32_r(KLB_PHILOX_ROUNDS, c0, c1, c2, c3, k0, k1, out);
}
/* the published 10-round generator: known-answer tests only */
KLB_HD void klb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  klb_philox4x32_r(10, c0, c1, c2, c3, k0, k1, out);
}

/* Counter layout (the RNG contract, DESIGN.md section "RNG"):
 *   c0 = slot        pair index for TAG_NORMAL, element index for TAG_SLOW,
 *                    0 for TAG_ACCEPT
 *   c1 = low 32 bits of the job's global transition counter t (first
 *        transition is t = 1; t = 0 is free for synthetic initial states)
 *   c2 = global chain index
 *   c3 = tag | attempt << 4 | (bits 32..47 of t) << 16
 *   key = 64-bit seed
 */
#define KLB_TAG_NORMAL 0u
#define KLB_TAG_SLOW   1u
#define KLB_TAG_ACCEPT 2u

typedef struct {
  uint32_t k0, k1;   /* seed */
  uint32_t t_lo;     /* c1 */
  uint32_t chain;    /* c2 */
  uint32_t t_hi16;   /* bits 32..47 of t, already shifted to bits 16..31 */
} klb_stream;

KLB_HD klb_stream klb_stream_make(uint64_t seed, uint64_t chain, uint64_t t) {
  klb_stream s;
  s.k0 = (uint32_t)seed; s.k1 = (uint32_t)(seed >> 32);
  s.t_lo = (uint32_t)t; s.chain = (uint32_t)chain;
  s.t_hi16 = (uint32_t)((t >> 32) & 0xFFFFu) << 16;
  return s;
}

KLB_HD void klb_stream_draw(const klb_stream* s, uint32_t slot, uint32_t tag, uint32_t attempt,
                            uint64_t* w0, uint64_t* w1) {
  uint32_t o[4];
  klb_philox4x32(slot, s->t_lo, s->chain, tag | (attempt << 4) | s->t_hi16, s->k0, s->k1, o);
  *w0 = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
  *w1 = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
}

/* --------------------------------------------------------- uniform variates */
/* [0,1): same map and same edge behaviour as Julia's rand() (u = 0 => log u = -Inf => accept,
 * src/samplers/iterate/MALA.jl:94) */
KLB_HD double klb_u01(uint64_t w) { return klb_mul(klb_i2d((int64_t)(w >> 11)), 0x1p-53); }
/* (0,1]: used where a logarithm must stay finite (ziggurat tail) */
KLB_HD double klb_u01_open(uint64_t w) { return klb_mul(klb_i2d((int64_t)(w >> 11) + 1), 0x1p-53); }

/* the accept/reject uniform of transition s->t for chain s->chain */
KLB_HD double klb_accept_uniform(const klb_stream* s) {
  uint64_t w0, w1;
  klb_stream_draw(s, 0u, KLB_TAG_ACCEPT, 0u, &w0, &w1);
  return klb_u01(w0);
}

/* ---------------------------------------------------------------- exp / log */
#ifndef KLB_TABLES_H
#include "klb_tables.h"
#endif

#define KLB_EXP_SHIFT 0x1.8p52

/* exp(x); tab = pointer to a copy of KLB_TAB.  |error| < 0.51 ulp (tests/test_math.py).
 * klb_exp_inline is the body; klb_exp the out-of-line call most kernels use (instruction-cache footprint).  The
 * thread-per-chain kernels (klb_glm.cuh) inline it so that the exp chains of two data rows interleave. */
KLB_HD double klb_exp_inline(double x, const uint64_t* tab) {
  if (!(x == x)) return x;
  if (x > 709.782712893384) return klb_u2d(0x7FF0000000000000ULL);
  if (x < -745.2) return 0.0;
  const uint64_t* T = tab + KLB_TAB_EXP;
  double z = klb_mul(x, klb_u2d(KLB_INVLN2N_BITS));
  double kd = klb_add(z, KLB_EXP_SHIFT);
  uint64_t ki = klb_d2u(kd);
  kd = klb_sub(kd, KLB_EXP_SHIFT);
  double r = klb_fma(kd, klb_u2d(KLB_NEGLN2HIN_BITS), x);
  r = klb_fma(kd, klb_u2d(KLB_NEGLN2LON_BITS), r);
  uint32_t j = (uint32_t)ki & 127u;
  double tail = klb_u2d(T[2 * j]);
  uint64_t sbits = T[2 * j + 1] + (ki << 45);
  double r2 = klb_mul(r, r);
  double p = klb_fma(r, 1.0 / 120.0, 1.0 / 24.0);
  double q = klb_fma(r, 1.0 / 6.0, 0.5);
  double tmp = klb_fma(klb_mul(r2, r2), p, klb_fma(r2, q, klb_add(tail, r)));
  /* k = j + 128 e ; the result is 2^e * H_j * (1 + tmp), H_j in [1,2) */
  int32_t e = (int32_t)(((int64_t)(ki << 13)) >> 20); /* sign-extend the 51-bit integer, then >> 7 */
  if (e >= -1021 && e <= 1022) {
    double scale = klb_u2d(sbits);
    return klb_fma(scale, tmp, scale);
  }
  if (e > 0) {              /* near overflow: evaluate at 2^-1009 and rescale */
    double scale = klb_u2d(sbits - (1009ULL << 52));
    return klb_mul(klb_fma(scale, tmp, scale), 0x1p1009);
  }
  {                         /* subnormal result: evaluate at 2^+1022 and rescale */
    double scale = klb_u2d(sbits + (1022ULL << 52));
    return klb_mul(klb_fma(scale, tmp, scale), 0x1p-1022);
  }
}

KLB_HD_NOINLINE double klb_exp(double x, const uint64_t* tab) { return klb_exp_inline(x, tab); }

/* log(x); error < 1 ulp, except < 2 ulp for x in (0.99, 1) (tests/test_math.py) */
KLB_HD double klb_log_inline(double x, const uint64_t* tab) {
  uint64_t ix = klb_d2u(x);
  int64_t kadj = 0;
  if (ix - 0x0010000000000000ULL >= 0x7FE0000000000000ULL) {
    /* x is 0, subnormal, negative, inf or nan */
    if ((ix << 1) == 0) return klb_u2d(0xFFF0000000000000ULL);       /* log(+-0) = -inf */
    if (ix == 0x7FF0000000000000ULL) return x;                          /* log(inf) = inf */
    if ((ix >> 63) || (ix & 0x7FF0000000000000ULL) == 0x7FF0000000000000ULL)
      return klb_u2d(0x7FF8000000000000ULL);                            /* nan */
    ix = klb_d2u(klb_mul(x, 0x1p52));                                   /* subnormal */
    kadj = -52;
  }
  const uint64_t* T = tab + KLB_TAB_LOG;
  uint64_t tmp = ix - KLB_LOG_OFF_BITS;
  uint32_t i = (uint32_t)(tmp >> 45) & 127u;
  int64_t k = ((int64_t)tmp >> 52) + kadj;
  uint64_t iz = ix - (tmp & (0xFFFULL << 52));
  double invc = klb_u2d(T[2 * i]), logc = klb_u2d(T[2 * i + 1]);
  double z = klb_u2d(iz);
  double r = klb_fma(z, invc, -1.0);               /* z/c - 1, |r| <= 2^-8 */
  double kd = klb_i2d(k);
  double t1 = klb_mul(kd, klb_u2d(KLB_LN2HI_BITS)); /* exact: 21-bit x 11-bit */
  double w = klb_add(t1, logc);
  double e1 = klb_add(klb_sub(t1, w), logc);        /* fast two-sum (|t1| >= |logc| or t1 == 0) */
  double hi = klb_add(w, r);
  double bb = klb_sub(hi, w);
  double e2 = klb_add(klb_sub(w, klb_sub(hi, bb)), klb_sub(r, bb));   /* two-sum */
  double lo = klb_fma(kd, klb_u2d(KLB_LN2LO_BITS), klb_add(e1, e2));
  double r2 = klb_mul(r, r);
  double q = klb_fma(r2, 1.0 / 7.0, klb_fma(r, -1.0 / 6.0, 0.2));
  double p = klb_fma(r2, q, klb_fma(r, -0.25, 1.0 / 3.0));
  double y = klb_fma(klb_mul(r, r2), p, klb_fma(r2, -0.5, lo));
  return klb_add(y, hi);
}

KLB_HD_NOINLINE double klb_log(double x, const uint64_t* tab) { return klb_log_inline(x, tab); }

/* ---------------------------------------------------------------------- erf
 * erf(x) for erf_rate_score(x, k) = erf(k*x)+1 (src/tuners/AcceptanceRateMCTuner.jl:17), the other score function of
 * AcceptanceRateMCTuner.  It runs once per tuning period and chain, so it is written for exactness, not speed:
 *   |x| <= 4.5  Maclaurin series  erf(x) = 2/sqrt(pi) sum_n (-1)^n x^(2n+1) / (n! (2n+1))  in double-double
 *               arithmetic (error-free two-sum / fma two-product): the cancellation (terms up to e^(x^2) ~ 6e8) costs
 *               30 of the 106 bits, the result is the correctly rounded double in all but ~1e-8 of the cases;
 *   4.5 < |x| < 6  1 - erfc(x), erfc by its continued fraction with klb_exp (erfc < 2e-10: a relative error of
 *               1e-14 in it is 1e-24 in the result);
 *   |x| >= 6    +-1 (erfc(6) = 2.2e-17 < 2^-54).
 * Only IEEE add / mul / fma / div: identical bits on host and device (tests/test_oracle_kat.py pins it to the
 * reference's known-answer values and to mpmath). */
typedef struct { double hi, lo; } klb_dd;
KLB_HD klb_dd klb_dd_fast2sum(double a, double b) {          /* |a| >= |b| */
  klb_dd r; r.hi = klb_add(a, b); r.lo = klb_sub(b, klb_sub(r.hi, a)); return r;
}
KLB_HD klb_dd klb_dd_2sum(double a, double b) {
  klb_dd r; r.hi = klb_add(a, b);
  double bb = klb_sub(r.hi, a);
  r.lo = klb_add(klb_sub(a, klb_sub(r.hi, bb)), klb_sub(b, bb));
  return r;
}
KLB_HD klb_dd klb_dd_add(klb_dd a, klb_dd b) {
  klb_dd s = klb_dd_2sum(a.hi, b.hi);
  return klb_dd_fast2sum(s.hi, klb_add(s.lo, klb_add(a.lo, b.lo)));
}
KLB_HD klb_dd klb_dd_mul(klb_dd a, klb_dd b) {
  double p = klb_mul(a.hi, b.hi);
  double e = klb_fma(a.hi, b.hi, -p);
  e = klb_fma(a.hi, b.lo, e);
  e = klb_fma(a.lo, b.hi, e);
  return klb_dd_fast2sum(p, e);
}
KLB_HD klb_dd klb_dd_div_d(klb_dd a, double d) {             /* a / d, d an exactly representable small integer */
  double q1 = klb_div(a.hi, d);
  double r = klb_add(klb_fma(-q1, d, a.hi), a.lo);
  double q2 = klb_div(r, d);
  return klb_dd_fast2sum(q1, q2);
}
KLB_HD_NOINLINE double klb_erf(double x, const uint64_t* tab) {
  if (!(x == x)) return x;
  const double ax = x < 0 ? -x : x;
  if (ax >= 6.0) return x < 0 ? -1.0 : 1.0;
  if (ax > 4.5) {
    /* erfc(ax) = exp(-ax^2) / (ax sqrt(pi)) * 1/(1 + u/(1 + 2u/(1 + 3u/(1 + ...)))),  u = 1/(2 ax^2); evaluated bottom-up */
    const double u = klb_div(0.5, klb_mul(ax, ax));
    double f = 1.0;
    for (int n = 60; n >= 1; --n) f = klb_add(1.0, klb_div(klb_mul(klb_i2d(n), u), f));
    /* exp(-ax^2) with the rounding error of ax^2 carried: ax^2 = p + e exactly, exp(-(p+e)) = exp(-p) (1 - e) */
    const double pp = klb_mul(ax, ax), ee = klb_fma(ax, ax, -pp);
    double ex = klb_exp(-pp, tab);
    ex = klb_fma(-ex, ee, ex);
    const double erfc = klb_div(ex, klb_mul(klb_mul(ax, 1.7724538509055160273 /* sqrt(pi) */), f));
    const double r = klb_sub(1.0, erfc);
    return x < 0 ? -r : r;
  }
  klb_dd x2; x2.hi = klb_mul(x, x); x2.lo = klb_fma(x, x, -x2.hi);
  klb_dd term; term.hi = x; term.lo = 0.0;                   /* x^(2n+1) / n! */
  klb_dd sum = term;
  for (int n = 1; n < 200; ++n) {
    term = klb_dd_div_d(klb_dd_mul(term, x2), klb_i2d(n));
    klb_dd t = klb_dd_div_d(term, klb_i2d(2 * n + 1));
    if (n & 1) { t.hi = -t.hi; t.lo = -t.lo; }
    sum = klb_dd_add(sum, t);
    const double at = t.hi < 0 ? -t.hi : t.hi, as = sum.hi < 0 ? -sum.hi : sum.hi;
    if (at <= klb_mul(as, 0x1p-110)) break;
  }
  /* 2/sqrt(pi) = 1.1283791670955125738961589031215451716881... as a double-double */
  klb_dd c; c.hi = 1.1283791670955126; c.lo = 1.533545961316588e-17;
  sum = klb_dd_mul(sum, c);
  return klb_add(sum.hi, sum.lo);
}

/* ------------------------------------------------- ziggurat standard normal */
/* Layout of a 64-bit word w:  layer = w & 255, sign = bit 8, mantissa m = w >> 12. */

/* x^y for x > 0 as exp(y*log(x)): the dual-averaging tuner's count^(-kappa) (src/tuners/DualAveragingMCTuner.jl:99).
 * Relative error <= (|y log x| + 2) ulp -- about 1e-15 for count <= 1e6, kappa = 0.75; exact 1 for x = 1.
 * Host and device evaluate the same expression, so both sides agree bit for bit. */
KLB_HD double klb_pow_pos(double x, double y, const uint64_t* tab) {
  return klb_exp(klb_mul(y, klb_log(x, tab)), tab);
}

/* Fast path: candidate x = (t-1)*(+-X[layer]), t = 1.m in [1,2).  Returns 1 when the
 * candidate lies in the layer's rectangle below the density (accept, ~98.8 %). */
KLB_HD int klb_zig_fast(uint64_t w, const uint64_t* tab, double* x) {
  uint32_t idx = (uint32_t)w & 255u;
  uint64_t m = w >> 12;
  double X = klb_u2d(tab[KLB_TAB_ZXK + 2 * idx]);
  uint64_t kk = tab[KLB_TAB_ZXK + 2 * idx + 1];
  uint64_t tb = 0x3FF0000000000000ULL | m;
  double t = klb_u2d(tb);
  double Xs = klb_u2d(klb_d2u(X) ^ ((w & 256ULL) << 55));   /* bit 8 of the word: sign of the draw */
  *x = klb_fma(t, Xs, -Xs);                                 /* -(t X - X) = t(-X) + X exactly; signed zero for m = 0 */
  return tb < kk;                                           /* m < k[layer]: the table holds k | 0x3ff<<52 */
}

/* Complete draw for element `elem` given its first candidate word w (from
 * klb_stream_draw(s, elem >> 1, KLB_TAG_NORMAL, 0), low word for even elem, high word for odd).
 * Extra randomness comes from (slot = elem, TAG_SLOW, attempt = 1, 2, ...). */
KLB_HD_NOINLINE double klb_normal_from_word(uint64_t w, uint32_t elem, const klb_stream* s, const uint64_t* tab) {
  uint32_t attempt = 0;
  for (;;) {
    double x;
    if (klb_zig_fast(w, tab, &x)) return x;
    uint32_t idx = (uint32_t)w & 255u;
    uint64_t sign = (w & 256ULL) << 55;
    uint64_t c0, c1;
    if (attempt >= 4000u) return x;      /* unreachable in practice (p < 1e-1000); bounds the loop */
    ++attempt;
    klb_stream_draw(s, elem, KLB_TAG_SLOW, attempt, &c0, &c1);
    if (idx == 0u) {
      /* base strip beyond r: Marsaglia's exponential-majorant tail sampler */
      for (;;) {
        double xx = klb_mul(-klb_log(klb_u01_open(c0), tab), klb_u2d(KLB_ZIG_RINV_BITS));
        double yy = -klb_log(klb_u01_open(c1), tab);
        if (klb_add(yy, yy) > klb_mul(xx, xx) || attempt >= 4000u)
          return klb_u2d(klb_d2u(klb_add(klb_u2d(KLB_ZIG_R_BITS), xx)) ^ sign);
        ++attempt;
        klb_stream_draw(s, elem, KLB_TAG_SLOW, attempt, &c0, &c1);
      }
    }
    /* wedge of layer idx: y uniform on [f(x[idx]), f(x[idx+1])] against the density */
    {
      double f0 = klb_u2d(tab[KLB_TAB_ZF + idx]), f1 = klb_u2d(tab[KLB_TAB_ZF + idx + 1]);
      double y = klb_fma(klb_u01(c0), klb_sub(f1, f0), f0);
      if (y < klb_exp(klb_mul(klb_mul(-0.5, x), x), tab)) return x;
    }
    w = c1;                               /* rejected: next candidate */
  }
}

/* Scalar reference draw of N(0,1) for (stream, element) -- what the oracle calls. */
KLB_HD double klb_normal(const klb_stream* s, uint32_t elem, const uint64_t* tab) {
  uint64_t w0, w1;
  klb_stream_draw(s, elem >> 1, KLB_TAG_NORMAL, 0u, &w0, &w1);
  return klb_normal_from_word((elem & 1u) ? w1 : w0, elem, s, tab);
}

#endif /* KLB_MATH_H */
