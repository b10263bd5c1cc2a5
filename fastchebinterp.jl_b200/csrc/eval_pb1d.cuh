// 1-D evaluation in a PRODUCT BASIS: one FP64 multiply-add per coefficient instead of the Clenshaw recurrence's
// multiply-add + subtract, i.e. past the 50 %-of-peak ceiling of the reference's algorithm.
//
// Replaces, for batches of points of a 1-D polynomial (BASELINE.json configs[0]: order 200),
//   (c::ChebPoly{1})(x)            /root/reference/src/eval.jl:63-70 -> evaluate, eval.jl:19-31
//   chebgradient(c, x)             eval.jl:168-177 -> Jevaluate, eval.jl:82-101 (scale :151)
//
// Mathematics.  With k = L hi + lo (0 <= lo < L, 0 <= hi < H, L H >= n) and w = T_L(z):
//     T_{L hi}(z) = T_hi(w),      T_{L hi}(z) T_lo(z) = (T_{L hi + lo}(z) + T_{|L hi - lo|}(z)) / 2,
// so the functions T_hi(w) T_lo(z) have the distinct degrees L hi + lo: a basis of the polynomials of degree < L H.
// The plan (eval_pb1d.cu, long double on the host) rewrites the Chebyshev coefficients once,
//     sum_k c_k T_k(z) = sum_hi T_hi(w) * ( sum_lo D[hi][lo] T_lo(z) ),
// by the triangular recursion  D[hi][0] = c_{L hi},  D[hi][lo] = 2 c_{L hi + lo} - D[hi+1][L-lo]  (hi >= 1),
// D[0][lo] = c_lo - D[1][L-lo] / 2.  L = AMAX is a compile-time constant (8 or 16): a thread keeps T_0..T_{L-1}(z) of
// its points in registers, and the next step of the same recurrence IS w; the inner sums are one DFMA per
// coefficient with the coefficient a constant-bank operand, the outer factor T_hi(w) advances by its own recurrence
// one row ahead, the outer sum is one DFMA per row:
//     n + L + 2 H  FP64 operations per point instead of 2 n   (order 200, L = 16, H = 13: 236 instead of 402).
// (The first version kept the HIGH factor T_a(w) in registers and ran the low one on the fly; it needed a separate
// recurrence for w = T_M(z) ahead of everything else: n + A + 3 M operations and a dependent prologue twice as long.)
// The gradient needs T_lo'(z), T_hi'(w) and w' = T_L'(z):  dv/dz = w' sum_hi T_hi'(w) g_hi + sum_hi T_hi(w) g'_hi.
//
// Numerics.  z is the oracle's z (same IEEE sequence, eval.jl:65); the summation order differs from Clenshaw's, so
// results agree with the oracle within rounding, not bit for bit.  Both algorithms lose accuracy like k^2 eps next
// to the end points; for coefficient vectors that do not decay (dense random) the two can differ there by more than
// 16 eps sum|c| (so does the oracle from the exact value).  AUTO therefore takes this path only after a plan-time
// self-check against the bit-exact Clenshaw kernel on probe points clustered at the end points (eval_pb1d.cu).
#pragma once
#include <cstdint>
#include "ptx_util.cuh"

namespace fastcheb {

// rows hi of D processed together (their coefficients are interleaved, see the kernel): 8 independent accumulation
// chains per thread whatever the number of components (the gradient kernels carry two inner sums per component: the
// scalar one takes 2 rows x 2 points x 2 sums, so its tables are laid out in groups of 2, not 4)
__host__ __device__ constexpr int pb_group(int ec, bool jac = false) { return ec == 1 ? (jac ? 2 : 4) : (ec == 2 ? 2 : 1); }

struct Pb1dParams {
  const void* pts;  // npts reals
  void* out_v;      // values:      out_v[p * v_stride + e0 + e]
  void* out_J;      // derivatives: out_J[p * J_stride + e0 + e]   (mode 0)
  long long* bad;   // first out-of-domain point index (atomicMin), may be null
  long long npts, idx_base, v_stride, J_stride;
  int e0, Etot, M;  // M = H, the number of rows hi
  double lb, ub;
  // Jacobian contraction (AD rules, chainrules.jl:4-39), as in the other evaluation kernels: 0 = store J;
  // 1 = pullback out_J[p] = sum_e J[e] tan[p*Etot + e]; 2 = pushforward out_J[p*Etot + e] = J[e] tan[p]
  const void* tan;
  int mode;
  int spin;      // always 0 (see the kernel)
  int last_len;  // number of non-zero entries of the last row of D: n - L (H - 1)
  int vec;       // bit 0: pts is 16-byte aligned (vector loads); bit 1: mode 0, contiguous 16-byte-aligned results (vector stores)
};

// The coefficients travel IN THE KERNEL PARAMETERS (constant bank 0; up to 18 KB of the 32 KB parameter space): every
// multiply-add of a warp uses the same coefficient, and a constant-bank / uniform-register operand costs no
// register-file write, whereas a broadcast LDS writes the value into all 32 lanes' registers — measured
// (profiles/r01_micro_fp64.jsonl, lds128_dfma_p2) that return path caps DFMAs fed two per loaded double at 69 % of the
// FP64 peak, which is where the first version of this kernel sat (63 % pipe utilisation, profiles/r02_pb1d_*).
// Layout of D: floor(H / pb_group(EC)) groups [group][lo][j][e] (hi = pb_group(EC) * group + j) from offset 0, the
// H % pb_group(EC) remaining rows [t][lo][e] from pb_tail_base (a fixed offset: all offsets are compile-time).
// H <= pb_mmax(AMAX) rows; the groups sit at compile-time offsets, the H % BU single rows after the LAST POSSIBLE group
__host__ __device__ constexpr int pb_mmax(int amax) { return amax == 8 ? 16 : 32; }
__host__ __device__ constexpr int pb_tail_base(int amax, int ec) { return pb_mmax(amax) * amax * ec; }
__host__ __device__ constexpr int pb_cap(int amax, int ec, bool jac = false) { return (pb_mmax(amax) + pb_group(ec, jac) - 1) * amax * ec; }

// Gradient kernels carry a second table, the product-basis coefficients of the DERIVATIVE series (see the kernel), in
// the same layout.  Both must fit the kernel parameter space (32 764 bytes): Float64, 16-entry tables and 4 components
// leave room for 31 of the 32 rows of the second table (single rows, no tail section: the table is simply cut short —
// gradients of such polynomials with more than 496 coefficients stay on the Clenshaw kernel).
template <typename R>
__host__ __device__ constexpr int pb_cap2(int amax, int ec) {
  const int cap = pb_cap(amax, ec, true);
  const int room = (32764 - 256 - cap * (int)sizeof(R)) / (int)sizeof(R);
  return cap <= room ? cap : room / (amax * ec) * (amax * ec);
}
template <typename R, int CAP, int CAP2 = 1>
struct Pb1dArgs {
  Pb1dParams p;
  R D[CAP];
  R D2[CAP2];
};

// K contiguous reals as 16-byte (or, when K reals are only 8 bytes more, 8-byte) vectors; `ptr` aligned accordingly
template <typename R, int K>
__device__ __forceinline__ void pb_vec_load(const R* __restrict__ ptr, R (&x)[K]) {
  constexpr int B = K * (int)sizeof(R);
  if constexpr (B % 16 == 0) {
    union { uint4 v[B / 16]; R r[K]; } u;
#pragma unroll
    for (int i = 0; i < B / 16; ++i) u.v[i] = reinterpret_cast<const uint4*>(ptr)[i];
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = u.r[k];
  } else if constexpr (B % 8 == 0) {
    union { uint2 v[B / 8]; R r[K]; } u;
#pragma unroll
    for (int i = 0; i < B / 8; ++i) u.v[i] = reinterpret_cast<const uint2*>(ptr)[i];
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = u.r[k];
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = ptr[k];
  }
}
template <typename R, int K>
__device__ __forceinline__ void pb_vec_store(R* __restrict__ ptr, const R (&x)[K]) {
  constexpr int B = K * (int)sizeof(R);
  if constexpr (B % 16 == 0) {
    union { uint4 v[B / 16]; R r[K]; } u;
#pragma unroll
    for (int k = 0; k < K; ++k) u.r[k] = x[k];
#pragma unroll
    for (int i = 0; i < B / 16; ++i) reinterpret_cast<uint4*>(ptr)[i] = u.v[i];
  } else if constexpr (B % 8 == 0) {
    union { uint2 v[B / 8]; R r[K]; } u;
#pragma unroll
    for (int k = 0; k < K; ++k) u.r[k] = x[k];
#pragma unroll
    for (int i = 0; i < B / 8; ++i) reinterpret_cast<uint2*>(ptr)[i] = u.v[i];
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) ptr[k] = x[k];
  }
}

// resident CTAs of 256 threads the register allocation aims at: the short-table (AMAX = 8) scalar Float64 value kernel
// needs 82 registers, which the allocation granularity turns into 2 CTAs per SM — capped at 80 it runs 3 (24 warps:
// the order-64 shapes are latency-bound, profiles/r02_pb1d_order64_ncu_summary.txt)
template <typename R, int AMAX, int EC, bool JAC>
__host__ __device__ constexpr int pb_min_ctas() {
  return (sizeof(R) == 8 && AMAX == 8 && EC == 1 && !JAC) ? 3 : 2;
}

template <typename R, int AMAX, int EC, int P, bool JAC>
__global__ void __launch_bounds__(256, pb_min_ctas<R, AMAX, EC, JAC>()) eval_pb1d_kernel(const __grid_constant__ Pb1dArgs<R, pb_cap(AMAX, EC, JAC), JAC ? pb_cap2<R>(AMAX, EC) : 1> args) {
  const Pb1dParams& prm = args.p;
  const R* Ds = args.D;
  [[maybe_unused]] const R* Ds2 = args.D2;  // the derivative series (JAC)
  const int M = prm.M;
  const R* __restrict__ pts = reinterpret_cast<const R*>(prm.pts);
  R* __restrict__ out_v = reinterpret_cast<R*>(prm.out_v);
  R* __restrict__ out_J = reinterpret_cast<R*>(prm.out_J);
  const R lb = (R)prm.lb, ub = (R)prm.ub;
  // The host takes this path only when the span passes SpanDiv's test (span_fast): the quotient by the span is the
  // reciprocal + exact-remainder correction (still the IEEE quotient) with no per-point test and no division subroutine
  const R span = rsub(ub, lb), rcp = rdiv(R(1), span);
  auto div_span = [&](R a) {
    const R q = rmul(a, rcp);
    return rfma(rfma(-span, q, a), rcp, q);
  };
  // derivative scale 2/(ub-lb) (eval.jl:151), any numerator: the same correction while nothing can underflow or overflow
  // on the way (numerator exponent within +-400, Float32 +-40), the true division otherwise
  [[maybe_unused]] auto div_any = [&](R a) {
    bool safe;
    if constexpr (sizeof(R) == 8) {
      const unsigned e = ((unsigned)(__double_as_longlong((double)a) >> 52)) & 0x7ffu;
      safe = e > 1023u - 400u && e < 1023u + 400u;
    } else {
      const unsigned e = (__float_as_uint((float)a) >> 23) & 0xffu;
      safe = e > 127u - 40u && e < 127u + 40u;
    }
    return safe ? div_span(a) : rdiv(a, span);
  };
  // A thread owns P CONSECUTIVE points (as in eval_lo1d.cuh): with 16-byte-aligned contiguous buffers (prm.vec) the
  // coordinates arrive and the results leave as 8- or 16-byte vectors, with one index computation per thread — at order 64
  // the per-point work around the series (index arithmetic, domain test, division test, scalar loads and stores) was as
  // large as the series itself.
  const long long tile_pts = (long long)blockDim.x * P;
  const long long ntiles = (prm.npts + tile_pts - 1) / tile_pts;
  const long long toff = (long long)threadIdx.x * P;
  const bool vec_in = (prm.vec & 1) != 0, vec_out = (prm.vec & 2) != 0;
  auto count = [&](long long i0) {
    const long long left = prm.npts - i0;
    return left >= P ? P : (left > 0 ? (int)left : 0);
  };
  auto fetch = [&](long long tile, R (&x)[P]) {
    const long long i0 = tile * tile_pts + toff;
    const int c = count(i0);
    if (vec_in && c == P) {
      pb_vec_load<R, P>(pts + i0, x);
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) x[p] = p < c ? pts[i0 + p] : lb;
    }
  };
  // the coordinates of the next tile are loaded while the current one is computed (a tile is ~1000 pipe cycles per
  // warp; an unprefetched global load at its head leaves the FP64 pipe idle for a quarter of the time)
  R xn[P];
  fetch(blockIdx.x, xn);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i0 = tile * tile_pts + toff;
    const int cnt = count(i0);
    bool live[P];
    R z[P], w[P];
    unsigned badmask = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const R x = xn[p];
      const bool ok = (lb <= x) && (x <= ub);                            // eval.jl:64 (closed box, NaN fails)
      z[p] = ok ? rsub(div_span(rmul(rsub(x, lb), R(2))), R(1)) : R(0);  // eval.jl:65, the oracle's operation order
      live[p] = ok && p < cnt;
      if (!ok && p < cnt) badmask |= 1u << p;
    }
    const bool all = cnt == P && badmask == 0;
    if (badmask && prm.bad) atomicMin(prm.bad, i0 + (__ffs(badmask) - 1) + prm.idx_base);
    fetch(tile + gridDim.x, xn);

    // T_lo(z) for lo < L = AMAX of every point, in registers; one more step of the same recurrence is w = T_L(z), the
    // argument of the outer factor: T_{L hi}(z) = T_hi(w).  (Doubling formulas T_{2m} = 2 T_m^2 - 1 would cut the
    // dependency depth to log2 L but measured less accurate — they failed the 16 eps sum|c| check on the order-200 README
    // polynomial — and no faster.)
    // The gradient is NOT formed from derivative tables (T_lo', T_hi', w': three operations per recurrence step and three
    // folds per row): the Chebyshev coefficients of p' are computed at plan time, rewritten in the same product basis
    // (second table, Ds2), and p'(z) is a second set of inner sums over the SAME T_lo(z), T_hi(w) — n more multiply-adds
    // and one more fold per row, nothing else.
    R Ta[P][AMAX];
    {
      R z2[P];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        z2[p] = rmul(R(2), z[p]);
        Ta[p][0] = R(1);
        Ta[p][1] = z[p];
      }
#pragma unroll
      for (int a = 2; a < AMAX; ++a)
#pragma unroll
        for (int p = 0; p < P; ++p) Ta[p][a] = rfma(z2[p], Ta[p][a - 1], -Ta[p][a - 2]);
#pragma unroll
      for (int p = 0; p < P; ++p) w[p] = rfma(z2[p], Ta[p][AMAX - 1], -Ta[p][AMAX - 2]);
    }

    // A loop that never runs (spin = 0).  Without it the tile loop is an INNERMOST loop, and ptxas hoists the ~250
    // loop-invariant constant-bank loads of the coefficients out of it: the uniform registers overflow into vector
    // registers and from there into local memory (128 registers + 184 bytes of spills instead of 114 registers).
#pragma unroll 1
    for (int r = 0; r < prm.spin; ++r) __nanosleep(1);
    // sums over the rows hi, two interleaved accumulators each (even / odd rows) so that consecutive rows do not chain
    R v[2][P][EC];
    [[maybe_unused]] R dv[2][JAC ? P : 1][EC];  // sum_hi T_hi(w) g'_hi, g' from the derivative series
    R tbm[P], tb[P], z2[P];                      // T_{hi-1}(w), T_hi(w); z2 = 2 w
#pragma unroll
    for (int p = 0; p < P; ++p) {
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        v[0][p][e] = v[1][p][e] = R(0);
        if constexpr (JAC) dv[0][p][e] = dv[1][p][e] = R(0);
      }
      z2[p] = rmul(R(2), w[p]);
      tbm[p] = w[p];  // T_{-1} = T_1: the recurrence T_{b+1} = 2 w T_b - T_{b-1} then also yields T_1 from T_0
      tb[p] = R(1);
    }
    // T_hi(w) of the next row: taken BEFORE the row's contraction, so the recurrence step overlaps it
    auto next_tb = [&](R (&t)[P]) {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        t[p] = tb[p];
        const R tn = rfma(z2[p], tb[p], -tbm[p]);
        tbm[p] = tb[p];
        tb[p] = tn;
      }
    };
    // fold g (and g') of one row into the sums
    auto fold = [&](int par, const R (&t)[P], const R (&g)[P][EC], [[maybe_unused]] const R (&gd)[JAC ? P : 1][EC]) {
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          v[par][p][e] = rfma(t[p], g[p][e], v[par][p][e]);
          if constexpr (JAC) dv[par][p][e] = rfma(t[p], gd[p][e], dv[par][p][e]);
        }
    };
    constexpr int BU = pb_group(EC, JAC), MAXG = pb_mmax(AMAX) / BU;
    const int ngroups = M / BU;
    // the group loop is unrolled completely (with a uniform exit), so that every coefficient is a constant-bank
    // operand at a compile-time offset
#pragma unroll
    for (int gi = 0; gi < MAXG; ++gi) {
      if (gi >= ngroups) break;
      const R* Dg = Ds + gi * BU * AMAX * EC;
      [[maybe_unused]] const R* Dg2 = Ds2 + (JAC ? gi * BU * AMAX * EC : 0);
      R g[BU][P][EC];
      [[maybe_unused]] R gd[BU][JAC ? P : 1][EC];
      R tbv[BU][P];
#pragma unroll
      for (int j = 0; j < BU; ++j) next_tb(tbv[j]);
#pragma unroll
      for (int j = 0; j < BU; ++j)
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            g[j][p][e] = Dg[j * EC + e];  // lo = 0: T_0(z) = 1
            if constexpr (JAC) gd[j][p][e] = Dg2[j * EC + e];
          }
#pragma unroll
      for (int a = 1; a < AMAX; ++a) {
#pragma unroll
        for (int j = 0; j < BU; ++j)
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            const R d = Dg[(a * BU + j) * EC + e];
#pragma unroll
            for (int p = 0; p < P; ++p) g[j][p][e] = rfma(d, Ta[p][a], g[j][p][e]);
            if constexpr (JAC) {
              const R d2 = Dg2[(a * BU + j) * EC + e];
#pragma unroll
              for (int p = 0; p < P; ++p) gd[j][p][e] = rfma(d2, Ta[p][a], gd[j][p][e]);
            }
          }
      }
#pragma unroll
      for (int j = 0; j < BU; ++j) fold((gi * BU + j) & 1, tbv[j], g[j], gd[j]);
    }
    // the M % BU remaining rows, one at a time ([t][a][e] at pb_tail_base)
    if constexpr (BU > 1) {
      const int ntail = M - ngroups * BU;
#pragma unroll
      for (int t = 0; t < BU - 1; ++t) {
        if (t >= ntail) break;
        const R* Dt = Ds + pb_tail_base(AMAX, EC) + t * AMAX * EC;
        [[maybe_unused]] const R* Dt2 = Ds2 + (JAC ? pb_tail_base(AMAX, EC) + t * AMAX * EC : 0);
        R g[P][EC];
        [[maybe_unused]] R gd[JAC ? P : 1][EC];
        R tbt[P];
        next_tb(tbt);
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            g[p][e] = Dt[e];
            if constexpr (JAC) gd[p][e] = Dt2[e];
          }
        // the LAST row of D is short (its entries lo >= n - L (H - 1) are zero): stop at the next multiple of 4
        const int alim = t == ntail - 1 ? prm.last_len : AMAX;
#pragma unroll
        for (int a = 1; a < AMAX; ++a) {
          if ((a & 3) == 1 && a >= alim) break;  // uniform
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            const R d = Dt[a * EC + e];
#pragma unroll
            for (int p = 0; p < P; ++p) g[p][e] = rfma(d, Ta[p][a], g[p][e]);
            if constexpr (JAC) {
              const R d2 = Dt2[a * EC + e];
#pragma unroll
              for (int p = 0; p < P; ++p) gd[p][e] = rfma(d2, Ta[p][a], gd[p][e]);
            }
          }
        }
        fold(t & 1, tbt, g, gd);
      }
    }

    if (vec_out && all) {
      // contiguous results of P consecutive points (mode 0, v_stride = J_stride = EC): vector stores
      R ov[P * EC];
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int e = 0; e < EC; ++e) ov[p * EC + e] = v[0][p][e] + v[1][p][e];
      pb_vec_store<R, P * EC>(out_v + i0 * EC, ov);
      if constexpr (JAC) {
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int e = 0; e < EC; ++e) ov[p * EC + e] = div_any(rmul(dv[0][p][e] + dv[1][p][e], R(2)));  // eval.jl:151
        pb_vec_store<R, P * EC>(out_J + i0 * EC, ov);
      }
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        if (!live[p]) continue;
        const long long ip = i0 + p;
#pragma unroll
        for (int e = 0; e < EC; ++e) out_v[ip * prm.v_stride + prm.e0 + e] = v[0][p][e] + v[1][p][e];
        if constexpr (JAC) {
          R Js[EC];
#pragma unroll
          for (int e = 0; e < EC; ++e) Js[e] = div_any(rmul(dv[0][p][e] + dv[1][p][e], R(2)));  // eval.jl:151
          if (prm.mode == 0) {
#pragma unroll
            for (int e = 0; e < EC; ++e) out_J[ip * prm.J_stride + prm.e0 + e] = Js[e];
          } else if (prm.mode == 1) {  // pullback: real(J' dy), chainrules.jl:7,14
            const R* dy = reinterpret_cast<const R*>(prm.tan) + ip * prm.Etot + prm.e0;
            R acc = R(0);
#pragma unroll
            for (int e = 0; e < EC; ++e) acc = rfma(Js[e], dy[e], acc);
            out_J[ip] = acc;
          } else {  // pushforward: J dx, chainrules.jl:24
            const R dx = reinterpret_cast<const R*>(prm.tan)[ip];
#pragma unroll
            for (int e = 0; e < EC; ++e) out_J[ip * prm.Etot + prm.e0 + e] = rmul(Js[e], dx);
          }
        }
      }
    }
  }
}

}  // namespace fastcheb
