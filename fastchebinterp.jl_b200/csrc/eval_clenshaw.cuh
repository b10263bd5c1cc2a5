// CUDA-core tensor-product Clenshaw evaluation (value, and value + Jacobian) — the generic path.
//
// Replaces, for a whole batch of points, the reference's per-point
//   (c::ChebPoly{N})(x)            /root/reference/src/eval.jl:63-67   (domain check, affine map)
//   evaluate(x, c, Val{dim}, ...)  eval.jl:16-56                        (nested Clenshaw)
//   chebjacobian / Jevaluate       eval.jl:147-152 / :79-134            (value + K x N Jacobian)
//
// Design (not a translation of the reference's recursion over a Julia array):
//  * one thread owns P points and EC real components and keeps every Clenshaw state in registers;
//    the nest over dimensions is a compile-time recursion, so all loops have CTA-uniform trip counts;
//  * the device handle owns the coefficient layout: components are split into planes, dimension 1 is
//    padded to a 16-byte multiple with zeros, so the hot dimension-1 loop reads 16-byte vectors at a
//    warp-uniform (broadcast) shared-memory address;
//  * coefficients reach shared memory as contiguous "slabs" (all of dims 1..s for one index of the
//    outer dims) moved by the TMA engine (cp.async.bulk + mbarrier), multi-buffered when the tensor
//    is larger than shared memory; because every thread of the CTA walks the coefficient tensor in
//    the same order, the slab hand-over is a plain CTA barrier inside the dimension nest;
//  * the recurrences start from zero state (b_{n+1} = b_{n+2} = 0), which reproduces the reference's
//    special-cased n == 1, n == 2 and n >= 3 branches bit for bit, and every
//    operation is an explicit IEEE op (fma / sub / mul / div, no contraction), so the kernel is
//    bit-identical to the CPU oracle.
#pragma once
#include <cstdint>
#include <type_traits>
#include "ptx_util.cuh"

namespace fastcheb {

constexpr int kMaxKernelNdim = 6;  // dimension nests instantiated (eval_clenshaw_inst.cuh)

struct EvalParams {
  const void* coefs;  // planar-padded slabs of this component group (device)
  const void* pts;    // npts x N reals, AoS
  void* out_v;        // values:   out_v[p * v_stride + e]
  void* out_J;        // Jacobian: out_J[p * J_stride + e + Etot * d]
  long long* bad;     // first out-of-domain point index (atomicMin), may be null
  long long npts;
  long long idx_base;  // added to the reported out-of-domain index (chunked host calls)
  long long v_stride, J_stride;
  int e0;    // first real component handled by this launch
  int Etot;  // real components per element
  int n[kMaxNdim];
  int nvec1;             // padded n1 / V
  int stride[kMaxNdim];  // element strides inside one component plane of a slab
  int s;                 // dims 1..s live inside a slab
  int T;                 // slabs per polynomial
  int compStride;        // elements per component plane of a slab
  int slabElems;         // EC * compStride
  int nstage;            // shared-memory slab buffers (0: 1-D line read from global memory, GLOBAL instantiation)
  double lb[kMaxNdim], ub[kMaxNdim];
  // Jacobian contraction (AD rules, chainrules.jl:4-39): 0 = store J; 1 = pullback, out_J[p*N + d] (+)= sum_e J[e][d] *
  // tan[p*Etot + e] (accumulated over component groups, e0 == 0 writes); 2 = pushforward, out_J[p*Etot + e] =
  // sum_d J[e][d] * tan[p*N + d]
  const void* tan;
  int mode;
};

template <typename R> struct VecOf;
template <> struct VecOf<double> {
  using type = double2;
  static constexpr int V = 2;
  static __device__ __forceinline__ double get(const double2& v, int u) { return u == 0 ? v.x : v.y; }
};
template <> struct VecOf<float> {
  using type = float4;
  static constexpr int V = 4;
  static __device__ __forceinline__ float get(const float4& v, int u) {
    return u == 0 ? v.x : (u == 1 ? v.y : (u == 2 ? v.z : v.w));
  }
};

// GLOBAL (1-D only): the coefficient line does not fit shared memory (n1 > ~29 000 doubles); it is read straight from
// global memory / L2 with the same warp-uniform vector loads.
template <typename R, int N, int EC, int P, bool JAC, bool GLOBAL = false>
struct ClenshawCta {
  using VT = VecOf<R>;
  using Vec = typename VT::type;
  static constexpr int V = VT::V;

  const EvalParams& prm;
  R z[N][P], z2[N][P];
  const R* slab;  // current slab in shared memory
  R* bufs;
  uint64_t* full;
  long long visit, total_visits;

  __device__ __forceinline__ ClenshawCta(const EvalParams& p) : prm(p) {}

  // ---- slab pipeline -------------------------------------------------------------------------
  __device__ __forceinline__ void issue(long long v) {
    const int b = (int)(v % prm.nstage);
    const long long t = prm.T - 1 - (v % prm.T);  // slabs are visited in descending memory order
    const uint32_t bytes = (uint32_t)prm.slabElems * (uint32_t)sizeof(R);
    const char* src = reinterpret_cast<const char*>(prm.coefs) + t * (long long)bytes;
    char* dst = reinterpret_cast<char*>(bufs) + (size_t)b * bytes;
    mbar_arrive_expect_tx(&full[b], bytes);
    for (uint32_t o = 0; o < bytes; o += 65536u) {
      const uint32_t len = bytes - o < 65536u ? bytes - o : 65536u;
      bulk_g2s(dst + o, src + o, len, &full[b]);
    }
  }
  // CTA-uniform: hand over to the next slab in visiting order
  __device__ __forceinline__ void acquire() {
    if (prm.T == 1) return;  // resident tensor, loaded once in the prologue
    __syncthreads();         // every thread is done with the previous slab's buffer
    if (threadIdx.x == 0 && visit >= 1) {
      const long long nxt = visit - 1 + prm.nstage;
      if (nxt < total_visits) issue(nxt);
    }
    const int b = (int)(visit % prm.nstage);
    mbar_wait(&full[b], (uint32_t)((visit / prm.nstage) & 1));
    slab = bufs + (size_t)b * prm.slabElems;
    ++visit;
  }

  // ---- dimension 1: the hot loop ----------------------------------------------------------------
  // value only.  eval.jl:19-31 with zero-start recurrences.
  __device__ __forceinline__ void line(int off, R (&out)[P][EC]) {
    R b1[P][EC], b2[P][EC];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int c = 0; c < EC; ++c) b1[p][c] = b2[p][c] = R(0);
    const R* base = slab + off;
#pragma unroll 2
    for (int jv = prm.nvec1 - 1; jv >= 1; --jv) {
#pragma unroll
      for (int c = 0; c < EC; ++c) {
        const Vec v = *reinterpret_cast<const Vec*>(base + c * prm.compStride + jv * V);
#pragma unroll
        for (int u = V - 1; u >= 0; --u) {
          const R cj = VT::get(v, u);
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const R nb = rsub(rfma(z2[0][p], b1[p][c], cj), b2[p][c]);  // eval.jl:28
            b2[p][c] = b1[p][c];
            b1[p][c] = nb;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < EC; ++c) {
      const Vec v = *reinterpret_cast<const Vec*>(base + c * prm.compStride);
#pragma unroll
      for (int u = V - 1; u >= 1; --u) {
        const R cj = VT::get(v, u);
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const R nb = rsub(rfma(z2[0][p], b1[p][c], cj), b2[p][c]);
          b2[p][c] = b1[p][c];
          b1[p][c] = nb;
        }
      }
#pragma unroll
      for (int p = 0; p < P; ++p) out[p][c] = rsub(rfma(z[0][p], b1[p][c], VT::get(v, 0)), b2[p][c]);  // eval.jl:31
    }
  }

  // value + d/dz1.  eval.jl:82-101.
  __device__ __forceinline__ void line_jac(int off, R (&out)[P][EC], R (&J)[P][EC][N]) {
    R b1[P][EC], b2[P][EC], d1[P][EC], d2[P][EC];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int c = 0; c < EC; ++c) b1[p][c] = b2[p][c] = d1[p][c] = d2[p][c] = R(0);
    const R* base = slab + off;
#pragma unroll 2
    for (int jv = prm.nvec1 - 1; jv >= 1; --jv) {
#pragma unroll
      for (int c = 0; c < EC; ++c) {
        const Vec v = *reinterpret_cast<const Vec*>(base + c * prm.compStride + jv * V);
#pragma unroll
        for (int u = V - 1; u >= 0; --u) {
          const R cj = VT::get(v, u);
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const R nb = rsub(rfma(z2[0][p], b1[p][c], cj), b2[p][c]);                           // eval.jl:96
            const R nd = rsub(rfma(z2[0][p], d1[p][c], rmul(R(2), b1[p][c])), d2[p][c]);         // eval.jl:97
            b2[p][c] = b1[p][c];
            b1[p][c] = nb;
            d2[p][c] = d1[p][c];
            d1[p][c] = nd;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < EC; ++c) {
      const Vec v = *reinterpret_cast<const Vec*>(base + c * prm.compStride);
#pragma unroll
      for (int u = V - 1; u >= 1; --u) {
        const R cj = VT::get(v, u);
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const R nb = rsub(rfma(z2[0][p], b1[p][c], cj), b2[p][c]);
          const R nd = rsub(rfma(z2[0][p], d1[p][c], rmul(R(2), b1[p][c])), d2[p][c]);
          b2[p][c] = b1[p][c];
          b1[p][c] = nb;
          d2[p][c] = d1[p][c];
          d1[p][c] = nd;
        }
      }
#pragma unroll
      for (int p = 0; p < P; ++p) {
        out[p][c] = rsub(rfma(z[0][p], b1[p][c], VT::get(v, 0)), b2[p][c]);  // eval.jl:101
        J[p][c][0] = rsub(rfma(z[0][p], d1[p][c], b1[p][c]), d2[p][c]);
      }
    }
  }

  // ---- dimensions 2..N --------------------------------------------------------------------------
  // eval.jl:33-54.  D is the (1-based) dimension; `off` the element offset inside the current slab.
  template <int D>
  __device__ __forceinline__ void level(int off, R (&out)[P][EC]) {
    if constexpr (D == 1) {
      line(off, out);
    } else {
      R b1[P][EC], b2[P][EC];
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int c = 0; c < EC; ++c) b1[p][c] = b2[p][c] = R(0);
      const bool inner = D <= prm.s;
      for (int j = prm.n[D - 1] - 1; j >= 0; --j) {
        if (D == prm.s + 1) acquire();
        R cj[P][EC];
        level<D - 1>(inner ? off + j * prm.stride[D - 1] : 0, cj);  // eval.jl:39,45-46,50
        if (j > 0) {
#pragma unroll
          for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < EC; ++c) {
              const R nb = rsub(rfma(z2[D - 1][p], b1[p][c], cj[p][c]), b2[p][c]);  // eval.jl:51
              b2[p][c] = b1[p][c];
              b1[p][c] = nb;
            }
        } else {
#pragma unroll
          for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < EC; ++c) out[p][c] = rsub(rfma(z[D - 1][p], b1[p][c], cj[p][c]), b2[p][c]);  // eval.jl:54
        }
      }
    }
  }

  // eval.jl:102-133.  J[..][q] for q < D are the partials wrt z_1..z_D.
  template <int D>
  __device__ __forceinline__ void level_jac(int off, R (&out)[P][EC], R (&J)[P][EC][N]) {
    if constexpr (D == 1) {
      line_jac(off, out, J);
    } else {
      R b1[P][EC], b2[P][EC], d1[P][EC], d2[P][EC];
      R Jb1[P][EC][D - 1], Jb2[P][EC][D - 1];
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int c = 0; c < EC; ++c) {
          b1[p][c] = b2[p][c] = d1[p][c] = d2[p][c] = R(0);
#pragma unroll
          for (int q = 0; q < D - 1; ++q) Jb1[p][c][q] = Jb2[p][c][q] = R(0);
        }
      const bool inner = D <= prm.s;
      for (int j = prm.n[D - 1] - 1; j >= 0; --j) {
        if (D == prm.s + 1) acquire();
        R cj[P][EC];
        R Jc[P][EC][N];
        level_jac<D - 1>(inner ? off + j * prm.stride[D - 1] : 0, cj, Jc);  // eval.jl:109,115-116,124
        if (j > 0) {
#pragma unroll
          for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < EC; ++c) {
              const R nb = rsub(rfma(z2[D - 1][p], b1[p][c], cj[p][c]), b2[p][c]);                    // eval.jl:125
              const R nd = rsub(rfma(z2[D - 1][p], d1[p][c], rmul(R(2), b1[p][c])), d2[p][c]);        // eval.jl:126
#pragma unroll
              for (int q = 0; q < D - 1; ++q) {
                const R nj = rsub(rfma(z2[D - 1][p], Jb1[p][c][q], Jc[p][c][q]), Jb2[p][c][q]);       // eval.jl:127
                Jb2[p][c][q] = Jb1[p][c][q];
                Jb1[p][c][q] = nj;
              }
              b2[p][c] = b1[p][c];
              b1[p][c] = nb;
              d2[p][c] = d1[p][c];
              d1[p][c] = nd;
            }
        } else {
#pragma unroll
          for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < EC; ++c) {
              out[p][c] = rsub(rfma(z[D - 1][p], b1[p][c], cj[p][c]), b2[p][c]);  // eval.jl:132
#pragma unroll
              for (int q = 0; q < D - 1; ++q)
                J[p][c][q] = rsub(rfma(z[D - 1][p], Jb1[p][c][q], Jc[p][c][q]), Jb2[p][c][q]);
              J[p][c][D - 1] = rsub(rfma(z[D - 1][p], d1[p][c], b1[p][c]), d2[p][c]);
            }
        }
      }
    }
  }

  // ---- the CTA's life ---------------------------------------------------------------------------
  __device__ __forceinline__ void run(unsigned char* smem_raw) {
    full = reinterpret_cast<uint64_t*>(smem_raw);
    bufs = reinterpret_cast<R*>(smem_raw + 128);
    const long long tile_pts = (long long)blockDim.x * P;
    const long long ntiles = (prm.npts + tile_pts - 1) / tile_pts;
    if ((long long)blockIdx.x >= ntiles) return;
    const long long my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    total_visits = my_tiles * prm.T;
    visit = 0;
    if constexpr (GLOBAL) {
      slab = reinterpret_cast<const R*>(prm.coefs);
    } else {
      if (threadIdx.x == 0) {
        for (int b = 0; b < prm.nstage; ++b) mbar_init(&full[b], 1);
        fence_mbar_init();
        for (int b = 0; b < prm.nstage && b < total_visits; ++b) issue(b);
      }
      __syncthreads();
      if (prm.T == 1) {
        mbar_wait(&full[0], 0);
        slab = bufs;
      }
    }
    const R* __restrict__ pts = reinterpret_cast<const R*>(prm.pts);
    R* __restrict__ out_v = reinterpret_cast<R*>(prm.out_v);
    R* __restrict__ out_J = reinterpret_cast<R*>(prm.out_J);

    SpanDiv<R> sdiv[N];
#pragma unroll
    for (int d = 0; d < N; ++d) sdiv[d].init((R)prm.lb[d], (R)prm.ub[d]);

    // The coordinates of the NEXT tile are loaded before the current one is evaluated: low-order polynomials are
    // memory-bound (8-100 bytes per point against a few dozen operations), and an unprefetched load at the head of
    // every tile left a third of the resident warps waiting on it (ncu: long_scoreboard, 48 % of the HBM bandwidth on
    // a 1-D order-4 polynomial).
    // (Not for N >= 4: the extra P * N registers spill there, and those nests are never memory-bound.)
    constexpr bool PF = N <= 3;
    [[maybe_unused]] R xn[P][N];
    auto fetch = [&](long long tile) {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const long long i = tile * tile_pts + (long long)p * blockDim.x + threadIdx.x;
#pragma unroll
        for (int d = 0; d < N; ++d) xn[p][d] = (tile < ntiles && i < prm.npts) ? pts[i * N + d] : (R)prm.lb[d];
      }
    };
    if constexpr (PF) fetch(blockIdx.x);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      long long idx[P];
      bool live[P];
      long long first_bad = 0x7fffffffffffffffLL;
      [[maybe_unused]] R xc[P][N];
      if constexpr (PF) {
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int d = 0; d < N; ++d) xc[p][d] = xn[p][d];
        fetch(tile + gridDim.x);
      }
#pragma unroll
      for (int p = 0; p < P; ++p) {
        idx[p] = tile * tile_pts + (long long)p * blockDim.x + threadIdx.x;
        live[p] = idx[p] < prm.npts;
        bool ok = true;
#pragma unroll
        for (int d = 0; d < N; ++d) {
          const R lb = (R)prm.lb[d], ub = (R)prm.ub[d];
          R x;
          if constexpr (PF)
            x = xc[p][d];
          else
            x = live[p] ? pts[idx[p] * N + d] : lb;
          ok = ok && (lb <= x) && (x <= ub);                           // eval.jl:64 (closed box, NaN fails)
          R zz = rsub(sdiv[d].div(rmul(rsub(x, lb), R(2))), R(1));  // eval.jl:65, this operation order (SpanDiv: the IEEE quotient)
          if (!(lb <= x && x <= ub)) zz = R(0);
          z[d][p] = zz;
          z2[d][p] = rmul(R(2), zz);
        }
        if (live[p] && !ok) {
          live[p] = false;
          if (idx[p] < first_bad) first_bad = idx[p];
        }
      }
      if (first_bad != 0x7fffffffffffffffLL && prm.bad) atomicMin(prm.bad, first_bad + prm.idx_base);

      R val[P][EC];
      if constexpr (JAC) {
        R J[P][EC][N];
        level_jac<N>(0, val, J);
#pragma unroll
        for (int p = 0; p < P; ++p)
          if (live[p]) {
            R Js[EC][N];
#pragma unroll
            for (int c = 0; c < EC; ++c) {
              out_v[idx[p] * prm.v_stride + prm.e0 + c] = val[p][c];
#pragma unroll
              for (int d = 0; d < N; ++d) {
                const R span = rsub((R)prm.ub[d], (R)prm.lb[d]);
                Js[c][d] = rdiv(rmul(J[p][c][d], R(2)), span);  // eval.jl:151
              }
            }
            if (prm.mode == 0) {
#pragma unroll
              for (int c = 0; c < EC; ++c)
#pragma unroll
                for (int d = 0; d < N; ++d) out_J[idx[p] * prm.J_stride + prm.e0 + c + (long long)prm.Etot * d] = Js[c][d];
            } else if (prm.mode == 1) {  // pullback: real(J' * dy), chainrules.jl:7,14
              const R* dy = reinterpret_cast<const R*>(prm.tan) + idx[p] * prm.Etot + prm.e0;
#pragma unroll
              for (int d = 0; d < N; ++d) {
                R acc = prm.e0 == 0 ? R(0) : out_J[idx[p] * N + d];
#pragma unroll
                for (int c = 0; c < EC; ++c) acc = rfma(Js[c][d], dy[c], acc);
                out_J[idx[p] * N + d] = acc;
              }
            } else {  // pushforward: J * dx, chainrules.jl:24
              const R* dx = reinterpret_cast<const R*>(prm.tan) + idx[p] * N;
#pragma unroll
              for (int c = 0; c < EC; ++c) {
                R acc = R(0);
#pragma unroll
                for (int d = 0; d < N; ++d) acc = rfma(Js[c][d], dx[d], acc);
                out_J[idx[p] * prm.Etot + prm.e0 + c] = acc;
              }
            }
          }
      } else {
        level<N>(0, val);
#pragma unroll
        for (int p = 0; p < P; ++p)
          if (live[p]) {
#pragma unroll
            for (int c = 0; c < EC; ++c) out_v[idx[p] * prm.v_stride + prm.e0 + c] = val[p][c];
          }
      }
    }
  }
};

template <typename R, int N, int EC, int P, bool JAC, bool GLOBAL = false>
__global__ void __launch_bounds__((JAC || N * EC >= 9) ? 256 : 512) clenshaw_kernel(const __grid_constant__ EvalParams prm) {
  extern __shared__ __align__(128) unsigned char fastcheb_smem[];
  ClenshawCta<R, N, EC, P, JAC, GLOBAL> cta(prm);
  cta.run(fastcheb_smem);
}

}  // namespace fastcheb
