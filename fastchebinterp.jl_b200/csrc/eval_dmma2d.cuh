// FP64 tensor-core (DMMA) evaluation of a 2-D ChebPoly whose tile-record stream fits in shared memory
// (BASELINE configs[1]: order (64,64) is 38.6 KB): value, and value + Jacobian / gradient.
//
// Same mathematics and the same tile-record stream as dmma_kernel (eval_dmma.cuh; reference
// /root/reference/src/eval.jl:16-56, :79-134, :147-152):
//
//     v[p,e] = sum_{k2} T_{k2}(z2_p) * G[p,k2,e],     G[p,k2,e] = sum_{k1} T_{k1}(z1_p) * C[k1,k2,e]
//
// but a 2-D polynomial has only n2/8 column tiles per point, so everything the generic kernel does once
// per pass (per-point Chebyshev tables, ring hand-offs, the global loads of the points) is a large share
// of a pass there (59 % of FP64 peak at order 64).  This kernel removes those costs:
//
//   * the whole tile-record stream is loaded into shared memory ONCE per CTA (cp.async.bulk -> UBLKCP);
//     after that warps never synchronise with each other, so the latency-bound prologue of one warp
//     hides under the DMMAs of the others;
//   * the Chebyshev recurrences of a pass are run once per (point, dimension) — one lane each — into
//     shared memory instead of once per lane of a quad: T_k(z1) goes, 16 values at a time, through a
//     2 KB scratch block of the warp into the A fragments (registers), T_k(z2) stays in the warp's weight
//     table.  Both are rotated by a row-dependent amount inside 16-column blocks so that the column-wise
//     stores of the prologue AND the fragment / weight loads of the tile loop are bank-conflict free;
//   * when n2 = 8j+1 or 8j+2 (every power-of-two order), the 1-2 trailing columns are contracted with
//     DFMAs from the A fragments already in registers instead of a whole, mostly padded DMMA tile
//     (order 64: 8 tiles instead of 9);
//   * the coordinates of a warp's next unit are prefetched while it computes the current one.
//
// Jacobian: A'-fragments T'_k(z1) = k U_{k-1}(z1) (second-kind recurrence, run in its own lanes) give
// dv/dz1 through a second accumulator chain on the same B fragments; dv/dz2 folds G with T'_{k2}(z2).
#pragma once
#include "eval_dmma.cuh"

namespace fastcheb {

struct Dmma2dParams {
  const double* frag;  // tile-record stream of this component group
  const void* pts;
  void* out_v;
  void* out_J;
  long long* bad;
  long long npts, idx_base;
  long long v_stride, J_stride;
  int e0, Etot;
  int ntiles;    // tile records (all resident in shared memory)
  int nmma;      // leading tiles contracted on the tensor cores
  int ntail;     // 0..2 trailing columns k2 = 8*nmma + j, contracted with DFMAs from tile record `nmma`
  double lb[2], ub[2];
  const void* tan;
  int mode;
};

constexpr int kDmma2dThreads = 512;

// shared-memory bytes of a launch with `warps` warps: tile records + per warp [weight table | tail slots | scratch block]
__host__ __device__ constexpr size_t dmma2d_smem(int EC, int KS, bool JAC, int ntiles, int warps) {
  const size_t rows = JAC ? 8 : 16, nt = JAC ? 2 : 1, KA = 4 * (size_t)KS;
  return 128 + 8 * ((size_t)ntiles * (kDmmaMeta + EC * (KS * 32 + 8)) + (size_t)warps * nt * rows * (KA + 2 + 16));
}

template <typename R, int EC, int KS, bool JAC>
__global__ void __launch_bounds__(kDmma2dThreads) dmma2d_kernel(const __grid_constant__ Dmma2dParams prm) {
  constexpr int N = 2;
  constexpr int MT = JAC ? 1 : 2;
  constexpr int ROWS = MT * 8;
  constexpr int U = KS * 32 + 8;
  constexpr int TR = kDmmaMeta + EC * U;
  constexpr int KA = 4 * KS;  // A entries per point (k1 = 1..KA) and main weight-table columns (k2 = 0..KA-1)
  constexpr int NT = JAC ? 2 : 1;
  const R* const pts_in = reinterpret_cast<const R*>(prm.pts);
  R* const out_v = reinterpret_cast<R*>(prm.out_v);
  [[maybe_unused]] R* const out_J = reinterpret_cast<R*>(prm.out_J);

  extern __shared__ __align__(128) unsigned char fastcheb_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(fastcheb_smem);
  unsigned int* counter = reinterpret_cast<unsigned int*>(fastcheb_smem + 8);  // units drawn by this CTA's warps
  double* recs = reinterpret_cast<double*>(fastcheb_smem + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncw = blockDim.x >> 5;
  double* t2 = recs + (size_t)prm.ntiles * TR + (size_t)warp * NT * ROWS * KA;  // [NT][ROWS][KA], rotated
  double* tails = recs + (size_t)prm.ntiles * TR + (size_t)ncw * NT * ROWS * KA + (size_t)warp * NT * ROWS * 2;
  double* sA = recs + (size_t)prm.ntiles * TR + (size_t)ncw * NT * ROWS * (KA + 2) + (size_t)warp * NT * ROWS * 16;

  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    *counter = 0;
    fence_mbar_init();
    const uint32_t bytes = (uint32_t)prm.ntiles * TR * 8u;
    mbar_arrive_expect_tx(full, bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t len = bytes - o < 32768u ? bytes - o : 32768u;
      bulk_g2s(reinterpret_cast<char*>(recs) + o, reinterpret_cast<const char*>(prm.frag) + o, len, full);
    }
  }
  __syncthreads();

  const int c = lane & 3;   // column pair inside a tile / k index inside a k-step
  const int r = lane >> 2;  // point row inside an M tile
  // ---- the lane's prologue task: one (point row, dimension[, kind]) recurrence ----
  const int trow = lane & (ROWS - 1);
  const int tdim = JAC ? (lane >> 3) & 1 : lane >> 4;
  const int tkind = JAC ? lane >> 4 : 0;  // 1: derivative table T'_k = k U_{k-1}
  // rotation of a table row inside its 16-column blocks (see header): A scratch / weight table
  auto rotA = [](int row) { return 4 * (row & 3) + (row >> 2); };
  auto rotT = [](int row) { return (row & 1) + 8 * ((row >> 1) & 1) + 2 * (row >> 2); };
  // the task stores value k (1..KA) at position k-1 (A scratch) or k (weight table): position j + tdim for j = k-1
  const int trot = (tdim ? rotT(trow) : rotA(trow)) + tdim;
  double* const tbase = tdim ? t2 + (size_t)(tkind * ROWS + trow) * KA : sA + (tkind * ROWS + trow) * 16;
  double* const ttail = tails + (size_t)(tkind * ROWS + trow) * 2;
  const R tlb = (R)prm.lb[tdim], tub = (R)prm.ub[tdim];
  SpanDiv<R> tdiv;
  tdiv.init(tlb, tub);

  // Units of ROWS consecutive points; a CTA owns a contiguous range and its warps draw units from a shared-memory
  // counter (warps differ in speed — their share of the FP64 pipe — and with a static split the slow ones finished
  // last with the pipe mostly idle: 14.5 of 16 warps active on average).
  const long long nunits = (prm.npts + ROWS - 1) / ROWS;
  const long long ubeg = nunits * blockIdx.x / gridDim.x, uend = nunits * (blockIdx.x + 1) / gridDim.x;
  auto draw = [&]() -> long long {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(counter, 1u);
    return ubeg + __shfl_sync(0xffffffffu, t, 0);
  };
  auto load_x = [&](long long u) -> R {
    const long long i = u * ROWS + trow;
    return (u < uend && i < prm.npts) ? pts_in[i * N + tdim] : tlb;
  };
  long long unit = draw(), unit_next = draw();
  R xn = load_x(unit);
  bool first = true;

  for (; unit < uend; unit = unit_next, unit_next = draw()) {
    const long long base = unit * ROWS;
    const R x = xn;
    xn = load_x(unit_next);  // prefetch: consumed a whole unit later
    const bool in = (tlb <= x) && (x <= tub);  // eval.jl:64
    const double z = in ? (double)rsub(tdiv.div(rmul(rsub(x, tlb), R(2))), R(1)) : 0.0;  // eval.jl:65 (SpanDiv: the IEEE quotient)
    const unsigned inmask = __ballot_sync(0xffffffffu, in);
    const unsigned okrows = (inmask & (inmask >> ROWS)) & ((1u << ROWS) - 1u);  // both coordinates inside the box
    const long long left = prm.npts - base;
    const unsigned liverows = left >= ROWS ? ((1u << ROWS) - 1u) : ((1u << (int)left) - 1u);
    const unsigned badrows = liverows & ~okrows;
    if (badrows && lane == 0 && prm.bad) atomicMin(prm.bad, base + (__ffs(badrows) - 1) + prm.idx_base);

    // ---- tables: T_k (or k U_{k-1}) for k = 1..KA+1, one recurrence per lane; the A fragments are picked up from
    //      the warp's 16-column scratch block after every 16 steps ----
    double A[MT][KS];
    double Ad[JAC ? MT : 1][JAC ? KS : 1];
    __syncwarp();  // the previous unit has finished reading this warp's tables
    {
      const double z2 = 2.0 * z;
      double prev = tkind ? 0.0 : 1.0, cur = tkind ? 1.0 : z;
      if (tdim) tbase[(trot - 1) & 15] = tkind ? 0.0 : 1.0;  // k2 = 0: T_0 = 1, T'_0 = 0
#pragma unroll
      for (int q = 0; q < KS / 4; ++q) {
        if (q > 0) __syncwarp();  // the A loads of the previous round are done: the scratch block may be overwritten
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int j = 16 * q + jj;  // value index k = j + 1
          double val = cur;
          if constexpr (JAC) val = cur * (tkind ? (double)(j + 1) : 1.0);
          if (j == KA - 1 && tdim)
            ttail[0] = val;  // k2 = KA
          else
            tbase[(tdim ? (j & ~15) + (jj == 15 ? 16 : 0) : 0) + ((j + trot) & 15)] = val;
          const double nx = fma(z2, cur, -prev);
          prev = cur;
          cur = nx;
        }
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int row = mt * 8 + r;
          const int b0 = c + rotA(row);
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // T_k(z1), k = 1 + 4*ks + c, ks = 4q + i, of rows r (+8)
            A[mt][4 * q + i] = sA[row * 16 + ((b0 + 4 * i) & 15)];
            if constexpr (JAC) Ad[mt][4 * q + i] = sA[(ROWS + row) * 16 + ((b0 + 4 * i) & 15)];
          }
        }
      }
      double val = cur;  // k = KA + 1 (second tail column)
      if constexpr (JAC) val = cur * (tkind ? (double)(KA + 1) : 1.0);
      if (tdim) ttail[1] = val;
    }
    __syncwarp();
    if (first) {
      mbar_wait(full, 0);  // the tile records have landed (overlapped with the first prologue)
      first = false;
    }

    // weight-table rows of this lane
    const double* wrow[MT];
    int wrot[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      wrow[mt] = t2 + (size_t)(mt * 8 + r) * KA;
      wrot[mt] = rotT(mt * 8 + r);
    }
    auto wload = [&](int mt, int k, bool deriv) -> double {  // T_k(z2) (or T'_k) of row mt*8 + r, k = 8*tile + 2c + j
      const double* row = wrow[mt] + (deriv ? ROWS * KA : 0);
      return row[(k & ~15) + ((k + wrot[mt]) & 15)];
    };

    double acc[MT][EC];
    double accJ[JAC ? MT : 1][JAC ? EC : 1][JAC ? N : 1];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        acc[mt][e] = 0.0;
        if constexpr (JAC) accJ[mt][e][0] = accJ[mt][e][1] = 0.0;
      }
    // NTL column tiles at a time: NTL * MT (* 2 with the derivative) independent accumulator chains per warp, so that
    // one or two warps in this loop keep the DMMA pipe of their SM sub-partition busy while the others are in the
    // latency-bound prologue.  The chains are folded with the weights T_{k2}(z2) right after their last k-step, the
    // oldest chain first.
    auto tiles = [&](auto ntl, const double* rec, int tile) {
      constexpr int NTL = decltype(ntl)::value;
      double w[NTL][MT][2];
      [[maybe_unused]] double wd[NTL][JAC ? MT : 1][2];
#pragma unroll
      for (int t = 0; t < NTL; ++t) {
        const int k = 8 * (tile + t) + 2 * c;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          w[t][mt][0] = wload(mt, k, false);
          w[t][mt][1] = wload(mt, k + 1, false);
          if constexpr (JAC) {
            wd[t][mt][0] = wload(mt, k, true);
            wd[t][mt][1] = wload(mt, k + 1, true);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        const double* u = rec + kDmmaMeta + e * U;
        double g0[NTL][MT], g1[NTL][MT];
        [[maybe_unused]] double h0[NTL][JAC ? MT : 1], h1[NTL][JAC ? MT : 1];
#pragma unroll
        for (int t = 0; t < NTL; ++t) {
          const double2 c0 = *reinterpret_cast<const double2*>(u + t * TR + 2 * c);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            g0[t][mt] = c0.x;  // k1 = 0 term: T_0 = 1
            g1[t][mt] = c0.y;
            if constexpr (JAC) h0[t][mt] = h1[t][mt] = 0.0;
          }
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
          for (int t = 0; t < NTL; ++t) {
            const double bf = u[t * TR + 8 + ks * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              dmma884(g0[t][mt], g1[t][mt], A[mt][ks], bf);
              if constexpr (JAC) dmma884(h0[t][mt], h1[t][mt], Ad[mt][ks], bf);
            }
          }
        }
#pragma unroll
        for (int t = 0; t < NTL; ++t)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            acc[mt][e] = fma(g0[t][mt], w[t][mt][0], acc[mt][e]);
            acc[mt][e] = fma(g1[t][mt], w[t][mt][1], acc[mt][e]);
            if constexpr (JAC) {
              accJ[mt][e][0] = fma(h0[t][mt], w[t][mt][0], accJ[mt][e][0]);
              accJ[mt][e][0] = fma(h1[t][mt], w[t][mt][1], accJ[mt][e][0]);
              accJ[mt][e][1] = fma(g0[t][mt], wd[t][mt][0], accJ[mt][e][1]);
              accJ[mt][e][1] = fma(g1[t][mt], wd[t][mt][1], accJ[mt][e][1]);
            }
          }
      }
    };
    const double* rec = recs;
    int tile = 0;
    for (; tile + 2 <= prm.nmma; tile += 2, rec += 2 * TR) tiles(std::integral_constant<int, 2>{}, rec, tile);
    if (tile < prm.nmma) {
      tiles(std::integral_constant<int, 1>{}, rec, tile);
      rec += TR;
    }

    // ---- trailing columns on the CUDA cores: each lane contracts the k1 it holds, the quad sum finishes it ----
    for (int j = 0; j < prm.ntail; ++j) {
      const int k2 = 8 * prm.nmma + j;
      double wt[MT];
      [[maybe_unused]] double wtd[JAC ? MT : 1];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const double* tl = tails + (size_t)(mt * 8 + r) * 2;
        wt[mt] = k2 < KA ? wload(mt, k2, false) : tl[k2 - KA];
        if constexpr (JAC) wtd[mt] = k2 < KA ? wload(mt, k2, true) : tl[ROWS * 2 + k2 - KA];
      }
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        const double* u = rec + kDmmaMeta + e * U;  // rec = tile record `nmma`
        double s[MT];
        [[maybe_unused]] double sd[JAC ? MT : 1];
        const double c0 = c == 0 ? u[j] : 0.0;  // k1 = 0 coefficient, counted once per quad
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          s[mt] = c0;
          if constexpr (JAC) sd[mt] = 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const double bf = u[8 + ks * 32 + 4 * j + c];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            s[mt] = fma(A[mt][ks], bf, s[mt]);
            if constexpr (JAC) sd[mt] = fma(Ad[mt][ks], bf, sd[mt]);
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          acc[mt][e] = fma(s[mt], wt[mt], acc[mt][e]);
          if constexpr (JAC) {
            accJ[mt][e][0] = fma(sd[mt], wt[mt], accJ[mt][e][0]);
            accJ[mt][e][1] = fma(s[mt], wtd[mt], accJ[mt][e][1]);
          }
        }
      }
    }

    // ---- quad reduction and store ----
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        double x0 = acc[mt][e];
        x0 += __shfl_xor_sync(0xffffffffu, x0, 1);
        x0 += __shfl_xor_sync(0xffffffffu, x0, 2);
        acc[mt][e] = x0;
        if constexpr (JAC) {
#pragma unroll
          for (int d = 0; d < N; ++d) {
            double y = accJ[mt][e][d];
            y += __shfl_xor_sync(0xffffffffu, y, 1);
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            accJ[mt][e][d] = y;
          }
        }
      }
    if (c == 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int row = mt * 8 + r;
        const long long idx = base + row;
        if ((liverows & okrows) >> row & 1u) {
#pragma unroll
          for (int e = 0; e < EC; ++e) out_v[idx * prm.v_stride + prm.e0 + e] = (R)acc[mt][e];
          if constexpr (JAC) {
            double Js[EC][N];
#pragma unroll
            for (int e = 0; e < EC; ++e)
#pragma unroll
              for (int d = 0; d < N; ++d)
                Js[e][d] = __ddiv_rn(__dmul_rn(accJ[mt][e][d], 2.0), __dsub_rn(prm.ub[d], prm.lb[d]));  // eval.jl:151
            if (prm.mode == 0) {
#pragma unroll
              for (int e = 0; e < EC; ++e)
#pragma unroll
                for (int d = 0; d < N; ++d)
                  out_J[idx * prm.J_stride + prm.e0 + e + (long long)prm.Etot * d] = (R)Js[e][d];
            } else if (prm.mode == 1) {  // pullback real(J' * dy), chainrules.jl:7,14 (accumulated over component groups)
              const R* dy = reinterpret_cast<const R*>(prm.tan) + idx * prm.Etot + prm.e0;
#pragma unroll
              for (int d = 0; d < N; ++d) {
                double a = prm.e0 == 0 ? 0.0 : (double)out_J[idx * N + d];
#pragma unroll
                for (int e = 0; e < EC; ++e) a = fma(Js[e][d], (double)dy[e], a);
                out_J[idx * N + d] = (R)a;
              }
            } else {  // pushforward J * dx, chainrules.jl:24
              const R* dx = reinterpret_cast<const R*>(prm.tan) + idx * N;
#pragma unroll
              for (int e = 0; e < EC; ++e) {
                double a = 0.0;
#pragma unroll
                for (int d = 0; d < N; ++d) a = fma(Js[e][d], (double)dx[d], a);
                out_J[idx * prm.Etot + prm.e0 + e] = (R)a;
              }
            }
          }
        }
      }
    }
  }
  if (first) mbar_wait(full, 0);  // never leave with the CTA's bulk copy still in flight
}

}  // namespace fastcheb
