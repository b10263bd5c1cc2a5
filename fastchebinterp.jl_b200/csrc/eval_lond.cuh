// Low-order N-d evaluation (2 <= N <= 4, a few hundred coefficients): the same idea as eval_lo1d.cuh for the tensor-product
// nest eval.jl:33-54 (value) and eval.jl:102-133 (value + Jacobian), scale eval.jl:151.
//
// The generic Clenshaw kernel walks a run-time dimension nest over a shared-memory slab with one point coordinate per load;
// on a 5 x 5 Float64 polynomial it issues ~70 % of all cycles with the FP64 pipe 47 % busy and moves 34 % of the HBM
// bandwidth — neither bound.  Here a thread owns P consecutive points (P N coordinates in, P E values and P N E Jacobian
// entries out, all as 16-byte vectors), the next tile's coordinates are prefetched, and the coefficients sit in the kernel
// parameters (constant bank), dimension 1 zero-padded to 1 + 4 g entries so that its recurrence is unrolled by 4 with no
// remainder (a zero-start step on a zero coefficient is an exact no-op).  Operation order = the oracle's: bit-identical.
#pragma once
#include "eval_lo1d.cuh"

namespace fastcheb {

constexpr int kLoNdMaxN = 4;
constexpr int kLoNdCap = 2048;  // reals in the kernel parameters: prod(n_d) * E with n_1 padded (two table capacities: 512, 2048)
__host__ __device__ constexpr int lond_cap(long long tot) { return tot <= 512 ? 512 : 2048; }

struct LoNdParams {
  const void* pts;  // npts x N reals (AoS), 16-byte aligned
  void* out_v;      // out_v[p * E + e]
  void* out_J;      // out_J[(p * N + d) * E + e]
  long long* bad;
  long long npts, idx_base;
  int n[kLoNdMaxN];
  int stride[kLoNdMaxN];  // element strides of the padded tensor: 1, n1p, n1p n2, ...
  int ngroups1;           // (n1p - 1) / 4
  double lb[kLoNdMaxN], ub[kLoNdMaxN];
};

template <typename R, int CAP>
struct LoNdArgs {
  LoNdParams p;
  R c[CAP];  // [i_N]...[i_2][i_1 < n1p][e]
};

template <typename R, int N, int EC, int P, bool JAC, int CAP>
struct LoNest {
  const LoNdArgs<R, CAP>& a;
  R z[N][P], z2[N][P];
  __device__ __forceinline__ LoNest(const LoNdArgs<R, CAP>& args) : a(args) {}

  template <int D>
  __device__ __forceinline__ void level(int off, R (&out)[P][EC]) {
    R b1[P][EC], b2[P][EC];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int e = 0; e < EC; ++e) b1[p][e] = b2[p][e] = R(0);
    if constexpr (D == 1) {
      for (int g = a.p.ngroups1 - 1; g >= 0; --g) {
#pragma unroll
        for (int u = 4; u >= 1; --u) {
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            const R cj = a.c[(off + 4 * g + u) * EC + e];
#pragma unroll
            for (int p = 0; p < P; ++p) {
              const R nb = rsub(rfma(z2[0][p], b1[p][e], cj), b2[p][e]);  // eval.jl:28
              b2[p][e] = b1[p][e];
              b1[p][e] = nb;
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        const R c0 = a.c[off * EC + e];
#pragma unroll
        for (int p = 0; p < P; ++p) out[p][e] = rsub(rfma(z[0][p], b1[p][e], c0), b2[p][e]);  // eval.jl:31
      }
    } else {
      for (int j = a.p.n[D - 1] - 1; j >= 1; --j) {
        R cj[P][EC];
        level<D - 1>(off + j * a.p.stride[D - 1], cj);  // eval.jl:39,45-46,50
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            const R nb = rsub(rfma(z2[D - 1][p], b1[p][e], cj[p][e]), b2[p][e]);  // eval.jl:51
            b2[p][e] = b1[p][e];
            b1[p][e] = nb;
          }
      }
      R c0[P][EC];
      level<D - 1>(off, c0);
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int e = 0; e < EC; ++e) out[p][e] = rsub(rfma(z[D - 1][p], b1[p][e], c0[p][e]), b2[p][e]);  // eval.jl:54
    }
  }

  // J[..][q], q < D: partials wrt z_1..z_D (eval.jl:82-133)
  template <int D>
  __device__ __forceinline__ void level_jac(int off, R (&out)[P][EC], R (&J)[P][EC][N]) {
    R b1[P][EC], b2[P][EC], d1[P][EC], d2[P][EC];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int e = 0; e < EC; ++e) b1[p][e] = b2[p][e] = d1[p][e] = d2[p][e] = R(0);
    if constexpr (D == 1) {
      for (int g = a.p.ngroups1 - 1; g >= 0; --g) {
#pragma unroll
        for (int u = 4; u >= 1; --u) {
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            const R cj = a.c[(off + 4 * g + u) * EC + e];
#pragma unroll
            for (int p = 0; p < P; ++p) {
              const R nb = rsub(rfma(z2[0][p], b1[p][e], cj), b2[p][e]);                    // eval.jl:96
              const R nd = rsub(rfma(z2[0][p], d1[p][e], rmul(R(2), b1[p][e])), d2[p][e]);  // eval.jl:97
              b2[p][e] = b1[p][e];
              b1[p][e] = nb;
              d2[p][e] = d1[p][e];
              d1[p][e] = nd;
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        const R c0 = a.c[off * EC + e];
#pragma unroll
        for (int p = 0; p < P; ++p) {
          out[p][e] = rsub(rfma(z[0][p], b1[p][e], c0), b2[p][e]);  // eval.jl:101
          J[p][e][0] = rsub(rfma(z[0][p], d1[p][e], b1[p][e]), d2[p][e]);
        }
      }
    } else {
      R Jb1[P][EC][D - 1], Jb2[P][EC][D - 1];
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int e = 0; e < EC; ++e)
#pragma unroll
          for (int q = 0; q < D - 1; ++q) Jb1[p][e][q] = Jb2[p][e][q] = R(0);
      for (int j = a.p.n[D - 1] - 1; j >= 1; --j) {
        R cj[P][EC];
        R Jc[P][EC][N];
        level_jac<D - 1>(off + j * a.p.stride[D - 1], cj, Jc);  // eval.jl:109,115-116,124
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            const R nb = rsub(rfma(z2[D - 1][p], b1[p][e], cj[p][e]), b2[p][e]);              // eval.jl:125
            const R nd = rsub(rfma(z2[D - 1][p], d1[p][e], rmul(R(2), b1[p][e])), d2[p][e]);  // eval.jl:126
#pragma unroll
            for (int q = 0; q < D - 1; ++q) {
              const R nj = rsub(rfma(z2[D - 1][p], Jb1[p][e][q], Jc[p][e][q]), Jb2[p][e][q]);  // eval.jl:127
              Jb2[p][e][q] = Jb1[p][e][q];
              Jb1[p][e][q] = nj;
            }
            b2[p][e] = b1[p][e];
            b1[p][e] = nb;
            d2[p][e] = d1[p][e];
            d1[p][e] = nd;
          }
      }
      R c0[P][EC];
      R Jc[P][EC][N];
      level_jac<D - 1>(off, c0, Jc);
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          out[p][e] = rsub(rfma(z[D - 1][p], b1[p][e], c0[p][e]), b2[p][e]);  // eval.jl:132
#pragma unroll
          for (int q = 0; q < D - 1; ++q) J[p][e][q] = rsub(rfma(z[D - 1][p], Jb1[p][e][q], Jc[p][e][q]), Jb2[p][e][q]);
          J[p][e][D - 1] = rsub(rfma(z[D - 1][p], d1[p][e], b1[p][e]), d2[p][e]);
        }
    }
  }
};

// points per thread: 16 bytes per coordinate column (2 Float64 / 4 Float32 points), twice that for the light nests — every
// coefficient operand (index arithmetic + constant-bank load) then feeds twice the recurrences: the 5 x 5 Float64 value
// kernel issued 235 instructions per point of which 80 FP64 with 2 points per thread (ncu: issue slots 79 % busy)
template <typename R>
__host__ __device__ constexpr int lond_pts(int n, int ec, bool jac) {
  return (16 / (int)sizeof(R)) * ((jac ? n * ec <= 2 : n * ec <= 3) ? 2 : 1);
}

template <typename R, int N, int EC, bool JAC, int CAP>
__global__ void __launch_bounds__(256) eval_lond_kernel(const __grid_constant__ LoNdArgs<R, CAP> args) {
  constexpr int P = lond_pts<R>(N, EC, JAC);
  const LoNdParams& prm = args.p;
  const R* __restrict__ pts = reinterpret_cast<const R*>(prm.pts);
  R* __restrict__ out_v = reinterpret_cast<R*>(prm.out_v);
  R* __restrict__ out_J = reinterpret_cast<R*>(prm.out_J);
  // spans pass SpanDiv's test on the host (lond_ok): reciprocal + exact-remainder correction = the IEEE quotient
  R lb[N], ub[N], span[N], rcp[N];
#pragma unroll
  for (int d = 0; d < N; ++d) {
    lb[d] = (R)prm.lb[d];
    ub[d] = (R)prm.ub[d];
    span[d] = rsub(ub[d], lb[d]);
    rcp[d] = rdiv(R(1), span[d]);
  }
  auto div_span = [&](R x, int d) {
    const R q = rmul(x, rcp[d]);
    return rfma(rfma(-span[d], q, x), rcp[d], q);
  };
  [[maybe_unused]] auto div_any = [&](R x, int d) {
    bool safe;
    if constexpr (sizeof(R) == 8) {
      const unsigned e = ((unsigned)(__double_as_longlong((double)x) >> 52)) & 0x7ffu;
      safe = e > 1023u - 400u && e < 1023u + 400u;
    } else {
      const unsigned e = (__float_as_uint((float)x) >> 23) & 0xffu;
      safe = e > 127u - 40u && e < 127u + 40u;
    }
    return safe ? div_span(x, d) : rdiv(x, span[d]);
  };
  const long long tile_pts = (long long)blockDim.x * P;
  const long long ntiles = (prm.npts + tile_pts - 1) / tile_pts;
  const long long toff = (long long)threadIdx.x * P;
  auto count = [&](long long i0) {
    const long long left = prm.npts - i0;
    return left >= P ? P : (left > 0 ? (int)left : 0);
  };

  R xn[P * N];
  {
    const long long i0 = (long long)blockIdx.x * tile_pts + toff;
    const int c0 = count(i0);
    lo_load<R, P * N>(pts + (c0 ? i0 * N : 0), c0 * N, R(0), xn);
  }
  LoNest<R, N, EC, P, JAC, CAP> nest(args);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i0 = tile * tile_pts + toff;
    const int cnt = count(i0);
    bool live[P];
    unsigned badmask = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      bool ok = true;
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const R x = p < cnt ? xn[p * N + d] : lb[d];
        const bool in = (lb[d] <= x) && (x <= ub[d]);  // eval.jl:64 (closed box, NaN fails)
        ok = ok && in;
        const R zz = in ? rsub(div_span(rmul(rsub(x, lb[d]), R(2)), d), R(1)) : R(0);  // eval.jl:65
        nest.z[d][p] = zz;
        nest.z2[d][p] = rmul(R(2), zz);
      }
      live[p] = ok && p < cnt;
      if (!ok && p < cnt) badmask |= 1u << p;
    }
    const bool all = cnt == P && badmask == 0;
    if (badmask && prm.bad) atomicMin(prm.bad, i0 + (__ffs(badmask) - 1) + prm.idx_base);
    {
      const long long n0 = (tile + gridDim.x) * tile_pts + toff;
      const int cn = count(n0);
      lo_load<R, P * N>(pts + (cn ? n0 * N : 0), cn * N, R(0), xn);
    }
    R val[P][EC];
    R ov[P * EC];
    if constexpr (JAC) {
      R J[P][EC][N];
      nest.template level_jac<N>(0, val, J);
      R oj[P * N * EC];
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          ov[p * EC + e] = val[p][e];
#pragma unroll
          for (int d = 0; d < N; ++d) oj[(p * N + d) * EC + e] = div_any(rmul(J[p][e][d], R(2)), d);  // eval.jl:151
        }
      if (cnt > 0) {
        lo_store<R, P * EC, EC>(out_v + i0 * EC, all, live, ov);
        lo_store<R, P * N * EC, N * EC>(out_J + i0 * N * EC, all, live, oj);
      }
    } else {
      nest.template level<N>(0, val);
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int e = 0; e < EC; ++e) ov[p * EC + e] = val[p][e];
      if (cnt > 0) lo_store<R, P * EC, EC>(out_v + i0 * EC, all, live, ov);
    }
  }
}

}  // namespace fastcheb
