// klb_dense_mma.cuh -- HMC on the dense-precision Gaussian target with the matrix-vector products on the
// fp64 tensor pipe (mma.sync.aligned.m8n8k4.f64 = DMMA.8x8x4 on sm_100a).
//
// Why it is legal: tools/dmma_order_test.cu (run on the B200) shows that the instruction accumulates exactly
// like the fma chain in increasing k, so chaining the k-steps through the C operand reproduces the oracle's
// `acc = fma(C[i][j], x[j], acc)`, j = 0..d-1, bit for bit.
// Why it pays: the DFMA register tile of klb_dense.cuh needs 8 L1/shared wavefronts per 16 DFMA (512 FMA); a warp
// tile of 16 chains x (8 NT) columns reuses every A fragment NT times and every B fragment twice:
// (2 + NT) fragment loads (2 wavefronts each) per 2 NT DMMA (256 FMA each), i.e. 10x fewer wavefronts per FMA
// at NT = 8.
//
// CTA = 256 threads = 8 warps, 16 chains.  Warp w owns columns [w*8NT, (w+1)*8NT) (d = 64 NT) of all 16 chains:
// thread (g = lane/4, q = lane%4) holds, for tile (mt, nt), elements (chain mt*8+g, column w*8NT + nt*8 + 2q + {0,1})
// of p, of C x and of the cached gradient -- the D-fragment layout of the instruction, and a double2 unit of
// the RNG / reduction contracts.  Positions live in shared memory (xs[chain][col], the A operand), C is
// streamed from L2 in 16-row slabs through a two-stage cp.async buffer shared by the CTA (the B operand).
#pragma once
#include "klb_dense.cuh"

#define KLB_MMA_MC 16
#define KLB_MMA_KB 16     /* rows of C per slab */
#define KLB_MMA_STAGES 2  /* ring depth.  4 stages x 8 rows (same bytes in flight) measured the same: the ring (132 KB)
                             is what shared memory leaves after the positions; see profiles/r1_summary.md */

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- thread-block-cluster path: one L2 read of a slab feeds every CTA of the cluster -----------------------
// cp.async.bulk ... .multicast::cluster (the 1-D form of the TMA copy) writes the same bytes at the same
// shared-memory offset of every CTA in the mask and signals each CTA's own mbarrier.  Each CTA issues 1/CL of
// the rows of a slab; FULL barriers count bytes (expect_tx), EMPTY barriers count one release per warp of every CTA.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
// Release of a stage by a consumer warp: an arrive on the EMPTY barrier of CTA `cta` of the cluster.  Default semantics
// (.release at CTA scope), the form CUTLASS's ClusterBarrier::arrive uses for consumer_release: the warp's fragment loads
// have returned before it gets here (their values fed the DMMAs above), so nothing needs flushing.  Round 1 wrote
// `.release.cluster`, which ptxas expands to MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of every
// arrive: 20 % of the kernel's stall samples (`stall_membar` 2.6 per issue, profiles/r2_kernel_metrics.csv column D).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* b, unsigned cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(b)), "r"(cta) : "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast(void* dst, const void* src, unsigned bytes, uint64_t* bar,
                                                   unsigned short mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int CL>
struct McPipe {
  uint64_t* full;        // [2]  bytes of a whole slab have landed in this CTA
  uint64_t* empty;       // [2]  every warp of all CL CTAs has finished reading the stage
  unsigned full_par;     // bit s: parity to wait for on full[s]  (every thread)
  unsigned empty_par;    // bit s: parity to wait for on empty[s] (thread 0)
  unsigned rank;
};

// thread 0 of every CTA: rows [rank*KB/CL, (rank+1)*KB/CL) of slab s -> stage s&1 of every CTA
template <int CL>
__device__ __forceinline__ void mc_issue_slab(McPipe<CL>& P, const double* __restrict__ Cm, double* slab, int d, int ldc,
                                              int s) {
  if (threadIdx.x != 0) return;
  const int st = s & 1;
  mbar_wait(P.empty + st, (P.empty_par >> st) & 1u);
  P.empty_par ^= 1u << st;
  mbar_expect_tx(P.full + st, (unsigned)(KLB_MMA_KB * d * 8));
  double* dst = slab + (size_t)st * KLB_MMA_KB * ldc;
  const double* src = Cm + (size_t)s * KLB_MMA_KB * d;
  constexpr int RPC = KLB_MMA_KB / CL;
#pragma unroll
  for (int r = 0; r < RPC; ++r) {
    const int row = (int)P.rank * RPC + r;
    bulk_g2s_multicast(dst + (size_t)row * ldc, src + (size_t)row * d, (unsigned)(d * 8), P.full + st,
                       (unsigned short)((1u << CL) - 1u));
  }
}

// slab s of C (KB rows) -> ring buffer s % STAGES, one commit group
__device__ __forceinline__ void mma_issue_slab(const double* __restrict__ Cm, double* slab, int d, int ldc, int s) {
  double* dst = slab + (size_t)(s % KLB_MMA_STAGES) * KLB_MMA_KB * ldc;
  const double* src = Cm + (size_t)s * KLB_MMA_KB * d;
  const int chunks = KLB_MMA_KB * (d / 2);                  // 16-byte chunks per slab
  for (int ch = threadIdx.x; ch < chunks; ch += KLB_DENSE_THREADS) {
    const int row = ch / (d / 2), cp = ch - row * (d / 2);
    cp_async16(dst + (size_t)row * ldc + 2 * cp, src + (size_t)row * d + 2 * cp);
  }
  cp_async_commit();
}
__device__ __forceinline__ void mma_prefetch(const double* __restrict__ Cm, double* slab, int d, int ldc) {
#pragma unroll
  for (int s = 0; s < KLB_MMA_STAGES - 1; ++s) mma_issue_slab(Cm, slab, d, ldc, s);
}

// acc[mt][nt][0..1] = (C x)[chain mt*8+g][col ...] for the 16 chains in xs.  `prefetched`: the first
// STAGES-1 slabs are already in flight (issued at the end of the previous product).  `prefetch_next`: issue them
// again on the way out, so the next product starts without an exposed L2 round trip (only legal when the caller
// does not use the slab region as scratch before that product).
template <int NT, int CL>
__device__ __forceinline__ void mma_matvec(double (&acc)[2][NT][2], const double* __restrict__ Cm, const double* xs,
                                           double* slab, int d, int ldx, int ldc, int w, int g, int q,
                                           bool prefetched, bool prefetch_next, McPipe<CL>& P) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
  const int nslab = d / KLB_MMA_KB;
  if (CL == 1) {
    if (!prefetched) mma_prefetch(Cm, slab, d, ldc);
  } else if (!prefetched) {
    // the slab region was used as generic-proxy scratch: order it before the async-proxy fills, cluster-wide
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cluster_sync_all();
    mc_issue_slab<CL>(P, Cm, slab, d, ldc, 0);
  }
  for (int s = 0; s < nslab; ++s) {
    if (CL == 1) {
      cp_async_wait<KLB_MMA_STAGES - 2>();                     // this thread's chunks of slab s have landed
      __syncthreads();                                         // ... everybody's; and slab s-1 is consumed by all
      if (s + KLB_MMA_STAGES - 1 < nslab) mma_issue_slab(Cm, slab, d, ldc, s + KLB_MMA_STAGES - 1);
      else cp_async_commit();                                  // empty group keeps the wait arithmetic uniform
    } else {
      if (s + 1 < nslab) mc_issue_slab<CL>(P, Cm, slab, d, ldc, s + 1);
      mbar_wait(P.full + (s & 1), (P.full_par >> (s & 1)) & 1u);
      P.full_par ^= 1u << (s & 1);
    }
    const double* cs = slab + (size_t)(s % KLB_MMA_STAGES) * KLB_MMA_KB * ldc + (size_t)w * 8 * NT + g;
    const double* xa = xs + (size_t)g * ldx + (size_t)s * KLB_MMA_KB + q;
#pragma unroll
    for (int kk = 0; kk < KLB_MMA_KB / 4; ++kk) {
      const double a0 = xa[kk * 4], a1 = xa[(size_t)8 * ldx + kk * 4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double b = cs[(size_t)(kk * 4 + q) * ldc + nt * 8];
        dmma884(acc[0][nt][0], acc[0][nt][1], a0, b);
        dmma884(acc[1][nt][0], acc[1][nt][1], a1, b);
      }
    }
    if (CL > 1) {
      // this warp is done with stage s&1: tell every issuer of the cluster (no CTA-wide barrier in the loop:
      // the warps of a CTA may run up to one slab apart)
      __syncwarp();
      if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int cta = 0; cta < CL; ++cta) mbar_arrive_cluster(P.empty + (s & 1), (unsigned)cta);
      }
    }
  }
  if (CL == 1) {
    cp_async_wait<0>();
    __syncthreads();                                           // the ring and xs are free
    if (prefetch_next) mma_prefetch(Cm, slab, d, ldc);
  } else if (prefetch_next) {
    mc_issue_slab<CL>(P, Cm, slab, d, ldc, 0);
  }
}

template <int NT, bool FMA, int CL>
__global__ void __launch_bounds__(KLB_DENSE_THREADS)
klb_dense_mma_kernel(const DArgs D) {
  static_assert(CL == 1 || KLB_MMA_STAGES == 2, "the multicast pipeline is two-stage");
  constexpr int MC = KLB_MMA_MC;
  const KArgs& A = D.k;
  extern __shared__ __align__(16) unsigned char dsm[];
  const int d = 64 * NT;
  const int ldx = d + 4, ldc = d + 4;                        // = 4 (mod 16): conflict-free fragment loads
  // dynamic shared memory: tab | xs[16*ldx] | slab ring[STAGES*KB*ldc] (also the reduction scratch) | DenseShared
  uint64_t* tab = reinterpret_cast<uint64_t*>(dsm);
  double* xs = reinterpret_cast<double*>(dsm + ((KLB_TAB_LEN * 8 + 15) & ~15));
  double* slab = xs + (size_t)MC * ldx;
  DenseShared<MC>& S = *reinterpret_cast<DenseShared<MC>*>(slab + (size_t)KLB_MMA_STAGES * KLB_MMA_KB * ldc);
  uint64_t* mbars = reinterpret_cast<uint64_t*>(&S + 1);     // full[2], empty[2] (cluster path)
  double* scA = slab;                                        // [16][d] scratch, valid between matvecs
  double* scB = slab + (size_t)MC * d;

  const int t = threadIdx.x, lane = t & 31, w = t >> 5, g = lane >> 2, q = lane & 3;
  for (int i = t; i < KLB_TAB_LEN; i += blockDim.x) tab[i] = A.tab[i];
  const long long c0 = (long long)blockIdx.x * MC;
  const double* Cm = D.Cm;

  if (t < MC) {
    const long long c = c0 + t;
    const bool live = c < A.nchains;
    S.lt_cur[t] = live ? A.lt[c] : 0.0;
    S.step[t] = live ? A.tune_step[c] : 1.0;
    S.accepted[t] = live ? A.tune_cnt[3 * c] : 0;
    S.proposed[t] = live ? A.tune_cnt[3 * c + 1] : 0;
    S.totproposed[t] = live ? A.tune_cnt[3 * c + 2] : 0;
    S.rate[t] = live ? A.tune_rate[c] : 0.0;
  }
  McPipe<CL> P;
  P.full = mbars; P.empty = mbars + 2; P.full_par = 0u; P.empty_par = 3u; P.rank = 0u;
  if (CL > 1) {
    P.rank = cluster_rank();
    if (t == 0) {
      mbar_init(P.full, 1u); mbar_init(P.full + 1, 1u);
      mbar_init(P.empty, (unsigned)(CL * (KLB_DENSE_THREADS / 32))); mbar_init(P.empty + 1, (unsigned)(CL * (KLB_DENSE_THREADS / 32)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();                                      // every CTA's barriers exist before anyone signals them
  }
  // element ownership: chain r(mt) = mt*8+g, column col(nt) = w*8NT + nt*8 + 2q (+1)
  auto colof = [&](int nt) { return w * 8 * NT + nt * 8 + 2 * q; };
  // positions -> shared memory
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int r = mt * 8 + g;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      double2 v = make_double2(0.0, 0.0);
      if (c0 + r < A.nchains) v = *reinterpret_cast<const double2*>(A.state + (c0 + r) * A.ld + colof(nt));
      *reinterpret_cast<double2*>(xs + (size_t)r * ldx + colof(nt)) = v;
    }
  }
  __syncthreads();
  double gc[2][NT][2];
  mma_matvec<NT, CL>(gc, Cm, xs, slab, d, ldx, ldc, w, g, q, false, false, P);   // cached C x of the starting point
  if (CL > 1) __syncthreads();

  const bool saving = (A.out_value != nullptr) || (A.out_lt != nullptr) || (A.out_grad != nullptr) ||
                      (A.out_accept != nullptr);
  long long count = A.count0;
  long long thin = (A.i0 > A.burnin) ? klb_mod(A.i0 - A.burnin - 1, A.thinning) : 0;

  for (long long it = 0; it < A.nt; ++it) {
    const long long irun = A.i0 + it;
    const unsigned long long tglob = A.t0 + 1ull + (unsigned long long)it;
    double p[2][NT][2], acc[2][NT][2];

    // ---------------- momentum: unit k = col/2 of chain r
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r = mt * 8 + g;
      const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)(c0 + r), tglob);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = colof(nt);
        uint64_t w0, w1;
        klb_stream_draw(&st, (unsigned)(col >> 1), KLB_TAG_NORMAL, 0u, &w0, &w1);
        double a, b;
        if (!klb_zig_fast(w0, tab, &a)) a = klb_normal_from_word(w0, (unsigned)col, &st, tab);
        if (!klb_zig_fast(w1, tab, &b)) b = klb_normal_from_word(w1, (unsigned)col + 1u, &st, tab);
        p[mt][nt][0] = a; p[mt][nt][1] = b;
        *reinterpret_cast<double2*>(scA + (size_t)r * d + col) = make_double2(a, b);
      }
    }
    __syncthreads();
    // old kinetic energy (canonical order: one warp per chain over the scratch copy of p)
    for (int r = w; r < MC; r += KLB_DENSE_THREADS / 32) {
      double a4[4] = {0.0, 0.0, 0.0, 0.0};
      for (int m = 0; m < D.nv; ++m) {
        const int i = 2 * (lane + 32 * m);
        double2 v = make_double2(0.0, 0.0);
        if (i < d) v = *reinterpret_cast<const double2*>(scA + (size_t)r * d + i);
        a4[m & 3] = dotacc(v.x, v.x, a4[m & 3]);
        a4[m & 3] = dotacc(v.y, v.y, a4[m & 3]);
      }
      double v = __dadd_rn(__dadd_rn(a4[0], a4[1]), __dadd_rn(a4[2], a4[3]));
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, s));
      if (lane == 0) S.k0[r] = v;
    }
    // ---------------- leapfrog
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = gc[mt][nt][0]; acc[mt][nt][1] = gc[mt][nt][1]; }
    for (int s = 1; s <= A.nleaps; ++s) {
      __syncthreads();                       // scratch / xs readers are done
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = mt * 8 + g;
        const double step = S.step[r];
        const double h = __dmul_rn(0.5, step);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          double2* xp = reinterpret_cast<double2*>(xs + (size_t)r * ldx + colof(nt));
          double2 xv = *xp;
          p[mt][nt][0] = Ar<FMA>::ma(h, __dmul_rn(-2.0, acc[mt][nt][0]), p[mt][nt][0]);   // p += (h g)
          p[mt][nt][1] = Ar<FMA>::ma(h, __dmul_rn(-2.0, acc[mt][nt][1]), p[mt][nt][1]);
          xv.x = Ar<FMA>::ma(step, p[mt][nt][0], xv.x);                                  // x += step p
          xv.y = Ar<FMA>::ma(step, p[mt][nt][1], xv.y);
          *xp = xv;
        }
      }
      __syncthreads();
      mma_matvec<NT, CL>(acc, Cm, xs, slab, d, ldx, ldc, w, g, q, s > 1, s < A.nleaps, P);   // g = -2 C x
      if (CL > 1) __syncthreads();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const double h = __dmul_rn(0.5, S.step[mt * 8 + g]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          p[mt][nt][0] = Ar<FMA>::ma(h, __dmul_rn(-2.0, acc[mt][nt][0]), p[mt][nt][0]);
          p[mt][nt][1] = Ar<FMA>::ma(h, __dmul_rn(-2.0, acc[mt][nt][1]), p[mt][nt][1]);
        }
      }
    }
    // ---------------- log-target of the proposal and new kinetic energy (the slab region is free again)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r = mt * 8 + g;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = colof(nt);
        *reinterpret_cast<double2*>(scA + (size_t)r * d + col) = make_double2(p[mt][nt][0], p[mt][nt][1]);
        *reinterpret_cast<double2*>(scB + (size_t)r * d + col) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
      }
    }
    __syncthreads();
    for (int r = w; r < MC; r += KLB_DENSE_THREADS / 32) {
      double k4[4] = {0.0, 0.0, 0.0, 0.0}, l4[4] = {0.0, 0.0, 0.0, 0.0};
      for (int m = 0; m < D.nv; ++m) {
        const int i = 2 * (lane + 32 * m);
        double2 pv = make_double2(0.0, 0.0), cv = make_double2(0.0, 0.0), xv = make_double2(0.0, 0.0);
        if (i < d) {
          pv = *reinterpret_cast<const double2*>(scA + (size_t)r * d + i);
          cv = *reinterpret_cast<const double2*>(scB + (size_t)r * d + i);
          xv = *reinterpret_cast<const double2*>(xs + (size_t)r * ldx + i);
        }
        k4[m & 3] = dotacc(pv.x, pv.x, k4[m & 3]);
        k4[m & 3] = dotacc(pv.y, pv.y, k4[m & 3]);
        l4[m & 3] = dotacc(xv.x, cv.x, l4[m & 3]);
        l4[m & 3] = dotacc(xv.y, cv.y, l4[m & 3]);
      }
      double k1 = __dadd_rn(__dadd_rn(k4[0], k4[1]), __dadd_rn(k4[2], k4[3]));
      double xcx = __dadd_rn(__dadd_rn(l4[0], l4[1]), __dadd_rn(l4[2], l4[3]));
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const double tk = __shfl_xor_sync(0xffffffffu, k1, s), tx = __shfl_xor_sync(0xffffffffu, xcx, s);
        k1 = __dadd_rn(k1, tk); xcx = __dadd_rn(xcx, tx);
      }
      if (lane == 0) {
        const long long c = c0 + r;
        const double lt_new = -xcx;
        const double oldh = __dsub_rn(S.lt_cur[r], __dmul_rn(0.5, S.k0[r]));
        const double newh = __dsub_rn(lt_new, __dmul_rn(0.5, k1));
        const double ratio = __dsub_rn(newh, oldh);
        bool acc_ = false;
        if (ratio >= 0.0) acc_ = true;
        else {
          const klb_stream st = klb_stream_make(A.seed, A.chain_offset + (unsigned long long)c, tglob);
          const double ex = klb_exp(ratio, tab);
          const double a = (ex != ex) ? ex : (ex < 1.0 ? ex : 1.0);
          acc_ = klb_accept_uniform(&st) < a;
        }
        S.lt_new[r] = lt_new;
        S.accept[r] = (acc_ && c < A.nchains) ? 1 : 0;
      }
    }
    __syncthreads();
    // ---------------- per-chain epilogue: counters, tuner
    if (t < MC) {
      const int r = t;
      if (A.counters_on) { S.proposed[r] += 1; if (S.accept[r]) S.accepted[r] += 1; }
      if (S.accept[r]) S.lt_cur[r] = S.lt_new[r];
      Tune tn;
      tn.step = S.step[r]; tn.accepted = S.accepted[r]; tn.proposed = S.proposed[r]; tn.totproposed = S.totproposed[r];
      tn.rate = S.rate[r];
      tuner_block<2>(A, tn, tab, c0 + r);
      S.step[r] = tn.step; S.accepted[r] = tn.accepted; S.proposed[r] = tn.proposed; S.totproposed[r] = tn.totproposed;
      S.rate[r] = tn.rate;
    }
    // ---------------- accept: publish x to the state column, keep C x; reject: restore x from the column
    const bool do_save = (irun > A.burnin) && (thin == 0);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r = mt * 8 + g;
      const long long c = c0 + r;
      const bool live = c < A.nchains;
      const bool acc_ = S.accept[r] != 0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = colof(nt);
        double2* xp = reinterpret_cast<double2*>(xs + (size_t)r * ldx + col);
        if (acc_) {
          gc[mt][nt][0] = acc[mt][nt][0]; gc[mt][nt][1] = acc[mt][nt][1];
          if (live) *reinterpret_cast<double2*>(A.state + c * A.ld + col) = *xp;
        } else if (live) {
          *xp = *reinterpret_cast<const double2*>(A.state + c * A.ld + col);
        } else {
          *xp = make_double2(0.0, 0.0);
        }
      }
    }
    __syncthreads();
    if (do_save && saving) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = mt * 8 + g;
        const long long c = c0 + r;
        if (c < A.nchains) {
          const long long colidx = c * A.npost + count;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int col = colof(nt);
            if (A.out_value)
              *reinterpret_cast<double2*>(A.out_value + colidx * A.ld + col) =
                  *reinterpret_cast<const double2*>(xs + (size_t)r * ldx + col);
            if (A.out_grad)
              *reinterpret_cast<double2*>(A.out_grad + colidx * A.ld + col) =
                  make_double2(__dmul_rn(-2.0, gc[mt][nt][0]), __dmul_rn(-2.0, gc[mt][nt][1]));
          }
          if (w == 0 && q == 0) {
            if (A.out_lt) A.out_lt[colidx] = S.lt_cur[r];
            if (A.out_accept) A.out_accept[colidx] = (unsigned char)S.accept[r];
          }
        }
      }
    }
    if (irun > A.burnin) {
      if (thin == 0) count += 1;
      thin = (thin + 1 == A.thinning) ? 0 : thin + 1;
    }
  }

  if (CL > 1) cluster_sync_all();     // nobody leaves while a peer may still signal its barriers
  // the state columns are current (written on every accept); per-chain scalars
  if (t < MC && c0 + t < A.nchains) {
    const long long c = c0 + t;
    A.lt[c] = S.lt_cur[t];
    A.tune_step[c] = S.step[t];
    A.tune_cnt[3 * c] = S.accepted[t]; A.tune_cnt[3 * c + 1] = S.proposed[t]; A.tune_cnt[3 * c + 2] = S.totproposed[t];
    A.tune_rate[c] = S.rate[t];
  }
}

int klb_dense_mma_launch(const DArgs& D, int fma, int cluster, cudaStream_t s);   // -1 when dim is not 64, 128, 256 or 512
int klb_dense_mma_attrs(int fma, int dim, int* regs, int* bps);
