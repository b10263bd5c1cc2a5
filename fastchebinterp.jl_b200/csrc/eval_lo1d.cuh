// Low-order 1-D evaluation (n <= 17 coefficients): the memory-bound end of (c::ChebPoly{1})(x) eval.jl:63-70, the 1-D
// Clenshaw recurrence eval.jl:19-31 and its derivative eval.jl:88-101, scale :151.
//
// Bound: HBM.  A point costs 8-16 bytes of traffic per real (coordinate in, value / derivative out) against ~20-45 FP64
// operations, so what matters is bytes in flight and instructions per point, not the arithmetic: the generic Clenshaw
// kernel (eval_clenshaw.cuh: one 8-byte load and store per point and thread, coefficient line in shared memory, run-time
// dimension nest) spends ~78 instructions per point on an order-4 polynomial and reached 48 % of the HBM bandwidth
// (71 % once the next tile was prefetched, profiles/r02_loworder.jsonl).  Here
//  * a thread owns P CONSECUTIVE points: coordinates arrive as 16-byte vector loads, results leave as 16-byte vector
//    stores (a warp instruction moves 512 contiguous bytes), one index computation per thread instead of one per point;
//  * the coordinates of the next tile are loaded before the current one is evaluated;
//  * the coefficients travel in the kernel parameters (constant bank, uniform operands — no shared memory, no barrier),
//    zero-padded to a multiple of 4 above the leading coefficient: a zero-start recurrence step on a zero coefficient
//    leaves the state exactly zero, so padded steps change nothing and the loop is unrolled by 4 without a remainder.
// The operation sequence per point is the oracle's (explicit IEEE fma / sub / mul / div): results are bit-identical to
// the generic kernel and to the CPU oracle.
#pragma once
#include <cstdint>
#include "ptx_util.cuh"

#ifndef FC_LO_P8
#define FC_LO_P8 1
#endif

namespace fastcheb {

constexpr int kLoMaxN = 17;      // coefficients c_0 .. c_16: the memory-bound range (table capacity of the first instantiation)
// Longer series are FP64-bound, but the same kernel is also the faster Clenshaw kernel for them (4 independent
// recurrences per thread fed from uniform operands: FP64 pipe ~90 % busy = 45 % of peak against the generic kernel's ~30 %):
// AUTO uses it for every 1-D polynomial up to 513 coefficients that the product-basis path does not take (n < 40, or
// rejected by its self-check, e.g. dense random coefficients).  Table capacities 17 / 129 / 513 keep the parameter copy small.
constexpr int kLoMaxNAny = 513;
__host__ __device__ constexpr int lo_cap(int n) { return n <= 17 ? 17 : (n <= 129 ? 129 : 513); }
constexpr int kLoMinBatch = 4096;  // smaller batches are launch-latency-bound: the generic kernel serves them

struct Lo1dParams {
  const void* pts;  // npts reals, 16-byte aligned
  void* out_v;      // out_v[p * E + e], 16-byte aligned
  void* out_J;      // out_J[p * E + e], 16-byte aligned (JAC)
  long long* bad;   // first out-of-domain point index (atomicMin), may be null
  long long npts, idx_base;
  int ngroups;      // ceil((n - 1) / 4): recurrence steps j = 4 ngroups .. 1
  double lb, ub;
};

template <typename R, int EC, int NCAP>
struct Lo1dArgs {
  Lo1dParams p;
  R c[NCAP * EC];  // [j][e], zero above j = n - 1
};

// points per thread: 32 bytes of coordinates for scalar polynomials, 16 for vector-valued ones; the FP64-bound scalar
// Float64 value kernels of the longer series (table capacity > 17) take 64 bytes = 8 independent recurrences
template <typename R>
__host__ __device__ constexpr int lo_pts(int ec, int ncap = 17, bool jac = false) {
  return (ec == 1 ? ((FC_LO_P8 && ncap > 17 && !jac && sizeof(R) == 8) ? 64 : 32) : 16) / (int)sizeof(R);
}

template <typename R> struct LoVec;
template <> struct LoVec<double> {
  using type = double2;
  static constexpr int V = 2;
};
template <> struct LoVec<float> {
  using type = float4;
  static constexpr int V = 4;
};

// K contiguous reals at `ptr` (16-byte aligned when `vec`), of which the first `cnt` exist
template <typename R, int K>
__device__ __forceinline__ void lo_load(const R* __restrict__ ptr, int cnt, R fill, R (&x)[K]) {
  using VT = typename LoVec<R>::type;
  constexpr int V = LoVec<R>::V;
  static_assert(K % V == 0, "whole vectors");
  if (cnt >= K) {
#pragma unroll
    for (int v = 0; v < K / V; ++v) {
      const VT t = reinterpret_cast<const VT*>(ptr)[v];
      if constexpr (V == 2) {
        x[2 * v] = t.x;
        x[2 * v + 1] = t.y;
      } else {
        x[4 * v] = t.x;
        x[4 * v + 1] = t.y;
        x[4 * v + 2] = t.z;
        x[4 * v + 3] = t.w;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = k < cnt ? ptr[k] : fill;
  }
}

// K contiguous reals to `ptr`; `keep` = every one of them is to be written (full group, no dead point)
template <typename R, int K, int EC>
__device__ __forceinline__ void lo_store(R* __restrict__ ptr, bool keep, const bool (&live)[K / EC], const R (&o)[K]) {
  using VT = typename LoVec<R>::type;
  constexpr int V = LoVec<R>::V;
  static_assert(K % V == 0, "whole vectors");
  if (keep) {
#pragma unroll
    for (int v = 0; v < K / V; ++v) {
      VT t;
      if constexpr (V == 2) {
        t.x = o[2 * v];
        t.y = o[2 * v + 1];
      } else {
        t.x = o[4 * v];
        t.y = o[4 * v + 1];
        t.z = o[4 * v + 2];
        t.w = o[4 * v + 3];
      }
      reinterpret_cast<VT*>(ptr)[v] = t;
    }
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (live[k / EC]) ptr[k] = o[k];
  }
}

template <typename R, int EC, bool JAC, int NCAP>
__global__ void __launch_bounds__(256) eval_lo1d_kernel(const __grid_constant__ Lo1dArgs<R, EC, NCAP> args) {
  constexpr int P = lo_pts<R>(EC, NCAP, JAC);
  const Lo1dParams& prm = args.p;
  const R* __restrict__ pts = reinterpret_cast<const R*>(prm.pts);
  R* __restrict__ out_v = reinterpret_cast<R*>(prm.out_v);
  R* __restrict__ out_J = reinterpret_cast<R*>(prm.out_J);
  const R lb = (R)prm.lb, ub = (R)prm.ub;
  // The host takes this path only when the span passes SpanDiv's test (lo1d_ok): the quotient by the span is then the
  // reciprocal + exact-remainder correction (still the IEEE quotient), with no per-point test and no division subroutine.
  const R span = rsub(ub, lb), rcp = rdiv(R(1), span);
  auto div_span = [&](R a) {
    const R q = rmul(a, rcp);
    return rfma(rfma(-span, q, a), rcp, q);
  };
  // derivative scale 2/(ub-lb) (eval.jl:151): any numerator — the same correction is the IEEE quotient while nothing can
  // underflow or overflow on the way (numerator exponent within +-400, Float32 +-40; the span's is within +-500 / +-60)
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
  const long long tile_pts = (long long)blockDim.x * P;
  const long long ntiles = (prm.npts + tile_pts - 1) / tile_pts;
  const long long toff = (long long)threadIdx.x * P;
  auto count = [&](long long i0) {
    const long long left = prm.npts - i0;
    return left >= P ? P : (left > 0 ? (int)left : 0);
  };

  R xn[P];
  {
    const long long i0 = (long long)blockIdx.x * tile_pts + toff;
    const int c0 = count(i0);
    lo_load<R, P>(pts + (c0 ? i0 : 0), c0, lb, xn);
  }
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i0 = tile * tile_pts + toff;
    const int cnt = count(i0);
    R z[P], z2[P];
    bool live[P];
    unsigned badmask = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const R x = xn[p];
      const bool ok = (lb <= x) && (x <= ub);                            // eval.jl:64 (closed box, NaN fails)
      z[p] = ok ? rsub(div_span(rmul(rsub(x, lb), R(2))), R(1)) : R(0);  // eval.jl:65, the oracle's operation order
      z2[p] = rmul(R(2), z[p]);
      live[p] = ok && p < cnt;
      if (!ok && p < cnt) badmask |= 1u << p;
    }
    const bool all = cnt == P && badmask == 0;
    if (badmask && prm.bad) atomicMin(prm.bad, i0 + (__ffs(badmask) - 1) + prm.idx_base);
    {
      const long long n0 = (tile + gridDim.x) * tile_pts + toff;
      const int cn = count(n0);
      lo_load<R, P>(pts + (cn ? n0 : 0), cn, lb, xn);
    }

    R b1[P][EC], b2[P][EC];
    [[maybe_unused]] R d1[JAC ? P : 1][EC], d2[JAC ? P : 1][EC];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        b1[p][e] = b2[p][e] = R(0);
        if constexpr (JAC) d1[p][e] = d2[p][e] = R(0);
      }
    for (int g = prm.ngroups - 1; g >= 0; --g) {
#pragma unroll
      for (int u = 4; u >= 1; --u) {
        const int j = 4 * g + u;
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          const R cj = args.c[j * EC + e];
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const R nb = rsub(rfma(z2[p], b1[p][e], cj), b2[p][e]);  // eval.jl:28 / :96
            if constexpr (JAC) {
              const R nd = rsub(rfma(z2[p], d1[p][e], rmul(R(2), b1[p][e])), d2[p][e]);  // eval.jl:97
              d2[p][e] = d1[p][e];
              d1[p][e] = nd;
            }
            b2[p][e] = b1[p][e];
            b1[p][e] = nb;
          }
        }
      }
    }
    R ov[P * EC];
    [[maybe_unused]] R oj[JAC ? P * EC : 1];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        ov[p * EC + e] = rsub(rfma(z[p], b1[p][e], args.c[e]), b2[p][e]);  // eval.jl:31 / :101
        if constexpr (JAC)
          oj[p * EC + e] = div_any(rmul(rsub(rfma(z[p], d1[p][e], b1[p][e]), d2[p][e]), R(2)));  // eval.jl:101, :151
      }
    if (cnt > 0) {
      lo_store<R, P * EC, EC>(out_v + i0 * EC, all, live, ov);
      if constexpr (JAC) lo_store<R, P * EC, EC>(out_J + i0 * EC, all, live, oj);
    }
  }
}

}  // namespace fastcheb
