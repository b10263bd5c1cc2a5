// FP64 tensor-core (DMMA) evaluation of a >= 2-D ChebPoly: value, and value + Jacobian.
//
// Computes the same quantity as the reference's evaluate / Jevaluate (/root/reference/src/eval.jl:16-56,
// :79-134) and chebjacobian's column scaling (eval.jl:147-152), but as a contraction instead of a
// recursion:
//
//     v[p,e]      = sum_m  w[p,m] * G[p,m,e],        G[p,m,e] = sum_{k1} T_{k1}(z1_p) * C[k1,m,e]
//     dv/dz1      = sum_m  w[p,m] * G'[p,m,e],       G'       = sum_{k1} T'_{k1}(z1_p) * C[k1,m,e]
//     dv/dz_d     = sum_m  w_d[p,m] * G[p,m,e]       (d >= 2: T'_{k_d} replaces T_{k_d} in the weight)
//
// with m = (k2..kN) flattened and w[p,m] = prod_{d>=2} T_{k_d}(z_d,p).  G is a GEMM (points x k1) x
// (k1 x columns) and runs on the FP64 tensor cores (mma.sync m8n8k4 -> SASS DMMA); the k1 = 0 term
// (T_0 = 1) is folded into the accumulator initialisation, so the MMA K extent is n1-1 (32 for an
// order-32 dimension: exactly 8 k-steps, no padding).  The weighted reduction over m costs 2 DFMA
// per 8 DMMA.  Everything is linear, so each lane keeps partial sums over the columns it owns and a
// quad shuffle-reduction finishes a point.
//
// Layout: the handle stores the coefficients as a stream of 8-column tile records
//     [ 16 doubles of metadata | EC units of (8 + 32*KS) doubles ]
// metadata: for each weighted dimension d and tile column j the int32 offset (into a point's weight-table
// row) of T_{k_d}(z_d) for that column — so a lane finds the weights of its two columns with ND 8-byte
// loads and no index arithmetic; unit: the 8 k1=0 coefficients of the tile, then per k-step the 32 lane
// values B[k = lane%4][col = lane/4] = C[1 + 4*ks + lane%4, 8*tile + lane/4, e], so a warp's B load is one
// conflict-free LDS.64 and a stage of tiles is one contiguous bulk copy (cp.async.bulk -> UBLKCP) from L2
// into a shared-memory ring, signalled by mbarriers.  Every warp of the CTA consumes the same stage
// (16 points per warp); the last warp to release a stage buffer refills it, so no warp ever waits for
// another warp's progress, only for data.  The weights of tile t+1 are computed in the same basic block
// as the DMMAs of tile t, which lets ptxas hide them in the DMMA issue gaps.
#pragma once
#include <cstdint>
#include <type_traits>
#include "ptx_util.cuh"

namespace fastcheb {

struct DmmaParams {
  const double* frag;  // tile-record stream of this component group (always Float64)
  const void* pts;     // points, values and Jacobians are in the polynomial's real type R (Float64 or Float32)
  void* out_v;
  void* out_J;
  long long* bad;
  long long npts, idx_base;
  long long v_stride, J_stride;
  int e0, Etot;
  int n[kMaxNdim];
  int ntiles;         // 8-column tiles per pass
  int M;              // prod n2..nN (valid columns)
  int tilesPerStage;  // tile records per shared-memory stage
  int nstage;
  int tabStride;  // doubles per point row of the weight tables
  double lb[kMaxNdim], ub[kMaxNdim];
  const void* tan;  // Jacobian contraction operand (see EvalParams::mode in eval_clenshaw.cuh)
  int mode;
#ifdef FC_PROFILE
  long long* prof;  // experiment: per-warp cycle counters
#endif
};

#ifndef FC_DMMA_MT
#define FC_DMMA_MT 2
#endif
constexpr int kDmmaMT = FC_DMMA_MT;  // 8-point M tiles per warp (value kernels)
#ifndef FC_DMMA_MT_JAC
#define FC_DMMA_MT_JAC 1
#endif
constexpr int kDmmaMTJac = FC_DMMA_MT_JAC;  // Jacobian kernels: the G and G' chains already give two MMAs per B fragment
__host__ __device__ constexpr int dmma_mt(bool jac) { return jac ? kDmmaMTJac : kDmmaMT; }
constexpr int kDmmaMeta = 16;  // doubles of metadata per tile record (8 columns x up to 3 dims x int32, padded)

// Threads per CTA the instantiation is compiled for: the value kernels with few accumulators leave registers for more
// warps (the tile loop is latency-bound per warp when a tile has only 16 DMMAs, so more warps per SM pay directly).
__host__ __device__ constexpr int dmma_max_threads(int EC, int KS, bool JAC) {
  return JAC ? (kDmmaMTJac == 1 ? 512 : 256) : (KS <= 4 ? 768 : (KS <= 8 && EC < 3 ? 640 : 512));
}
// R: the real type of points and results.  Float32 polynomials run the same FP64 tensor-core arithmetic (their
// coefficients are widened once in the tile-record stream); the normalised coordinate is formed in R with the oracle's
// operation sequence and only then widened, results are rounded to R at the very end.
template <typename R, int N, int EC, int KS, bool JAC>
__global__ void __launch_bounds__(dmma_max_threads(EC, KS, JAC)) dmma_kernel(const __grid_constant__ DmmaParams prm) {
  constexpr int MT = dmma_mt(JAC);
  const R* const pts_in = reinterpret_cast<const R*>(prm.pts);
  R* const out_v = reinterpret_cast<R*>(prm.out_v);
  [[maybe_unused]] R* const out_J = reinterpret_cast<R*>(prm.out_J);
  constexpr int U = KS * 32 + 8;
  constexpr int TR = kDmmaMeta + EC * U;  // doubles per tile record
  constexpr int ROWS = MT * 8;
  constexpr int ND = N - 1;  // weighted dimensions 2..N
  constexpr int NTAB = JAC ? 2 : 1;
  extern __shared__ __align__(128) unsigned char fastcheb_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(fastcheb_smem);
  unsigned int* released = reinterpret_cast<unsigned int*>(full + 8);  // warps done with each stage buffer
  unsigned int* ready = released + 8;  // low 32 bits of (visit + 1) whose data is known to be in each buffer
  double* stages = reinterpret_cast<double*>(fastcheb_smem + 128);
  const int stageDoubles = prm.tilesPerStage * TR;
  double* tables = stages + (size_t)prm.nstage * stageDoubles;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncw = blockDim.x >> 5;  // every warp consumes
  const long long pass_pts = (long long)ncw * ROWS;
  const long long npass = (prm.npts + pass_pts - 1) / pass_pts;
  if ((long long)blockIdx.x >= npass) return;
  const long long my_pass = (npass - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int S = (prm.ntiles + prm.tilesPerStage - 1) / prm.tilesPerStage;  // stages per pass
  const int ns = prm.nstage;

  if (threadIdx.x == 0) {
    for (int b = 0; b < ns; ++b) {
      mbar_init(&full[b], 1);
      released[b] = 0;
      ready[b] = 0;
    }
    fence_mbar_init();
  }
  __syncthreads();

  // No producer warp and no warp that others wait on: visits 0..nstage-1 are issued in the prologue, and
  // the LAST warp to release a stage buffer (shared-memory counter) refills it with visit v + nstage.
  const long long total_visits = my_pass * S;
  auto issue = [&](long long u) {
    const int b = (int)(u % ns);
    const int s = (int)(u % S);
    const int t0 = s * prm.tilesPerStage;
    const int nt = prm.ntiles - t0 < prm.tilesPerStage ? prm.ntiles - t0 : prm.tilesPerStage;
    const uint32_t bytes = (uint32_t)nt * TR * 8u;
    const char* src = reinterpret_cast<const char*>(prm.frag) + (size_t)t0 * TR * 8u;
    char* dst = reinterpret_cast<char*>(stages + (size_t)b * stageDoubles);
    mbar_arrive_expect_tx(&full[b], bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t len = bytes - o < 32768u ? bytes - o : 32768u;
      bulk_g2s(dst + o, src + o, len, &full[b]);
    }
  };
  if (threadIdx.x == 0)
    for (long long u = 0; u < ns && u < total_visits; ++u) issue(u);
  // Wait until visit `u` has landed in buffer `bb`.  All warps ask at about the same time and 16 mbarrier
  // try_waits serialise in the sync unit, so the first warp through publishes a plain shared-memory flag
  // and the others take a one-LDS fast path.
#ifdef FC_PROFILE
  long long prof_wait = 0, prof_slow = 0, prof_issue = 0, prof_t0 = clock64();
#endif
  auto wait_stage = [&](long long u, int bb, uint32_t parity) {
    const unsigned tag = (unsigned)(u + 1);
    unsigned seen;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&ready[bb])) : "memory");
    if (seen != tag) {
#ifdef FC_PROFILE
      const long long t0 = clock64();
      mbar_wait(&full[bb], parity);
      prof_wait += clock64() - t0;
      prof_slow += 1;
#else
      mbar_wait(&full[bb], parity);
#endif
      asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&ready[bb])), "r"(tag) : "memory");
    }
  };

  const int c = lane & 3;   // column pair inside a tile / k index inside a k-step
  const int r = lane >> 2;  // point row inside an M tile
  // Weight tables of this warp (+ a derivative copy for the Jacobian), two layouts chosen at compile time from
  // measurements (profiles/r01_configs.jsonl):
  //   plain  [point row][k]      — one 8-byte load per M tile; fastest for the 3-D K=3 value kernel (91.9 % vs 86.9 %)
  //   paired [row r][k][mt]      — the MT points a lane owns (rows r, r+8, ..) side by side, one 16-byte load serves
  //                                both M tiles; wins wherever the weights are a larger share of the tile: Jacobians,
  //                                >= 3 weighted dimensions (4-D: 72 % vs 66 %), scalar-valued 3-D
  constexpr bool PAIRED = JAC || ND >= 3 || (ND == 2 && EC < 3);
  // (code generation of the tile loop is sensitive to how these pointers are formed: the two layouts are kept as two
  // literal code paths, and the plain one is the form that was measured at 91.9 %)
  double* tab = tables + (size_t)warp * ROWS * prm.tabStride * NTAB;
  double* dtab = tab + (size_t)ROWS * prm.tabStride;
  const double* trow[MT];
  const double* drow[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    if constexpr (PAIRED) {
      trow[mt] = tab + (size_t)r * prm.tabStride * MT + mt;
      drow[mt] = dtab + (size_t)r * prm.tabStride * MT + mt;
    } else {
      trow[mt] = tab + (size_t)(mt * 8 + r) * prm.tabStride;
      drow[mt] = dtab + (size_t)(mt * 8 + r) * prm.tabStride;
    }
  }
  struct TV {
    double v[MT];
  };
  auto tload = [&](const double* const (&rows)[MT], int k) {
    TV t;
    if constexpr (PAIRED && MT == 2) {
      const double2 q = *reinterpret_cast<const double2*>(rows[0] + 2 * k);
      t.v[0] = q.x;
      t.v[1] = q.y;
    } else if constexpr (PAIRED) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) t.v[mt] = rows[mt][k * MT];
    } else {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) t.v[mt] = rows[mt][k];
    }
    return t;
  };

  // weights of the two columns this lane owns in the tile whose record starts at `rec`
  struct Weights {
    double w[MT][2];
    double wd[JAC ? MT : 1][2][JAC ? N : 1];  // [..][d]: T' in dimension d+1 (d >= 1)
  };
  auto weights = [&](const double* rec, Weights& W) {
    const int2* meta = reinterpret_cast<const int2*>(rec);
    TV t0[ND], t1[ND];
    int2 ix[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      ix[d] = meta[d * 4 + c];
      t0[d] = tload(trow, ix[d].x);
      t1[d] = tload(trow, ix[d].y);
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      double w0 = t0[0].v[mt], w1 = t1[0].v[mt];
#pragma unroll
      for (int d = 1; d < ND; ++d) {
        w0 *= t0[d].v[mt];
        w1 *= t1[d].v[mt];
      }
      W.w[mt][0] = w0;
      W.w[mt][1] = w1;
    }
    if constexpr (JAC) {
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const TV x0 = tload(drow, ix[dd].x), x1 = tload(drow, ix[dd].y);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          double y0 = x0.v[mt], y1 = x1.v[mt];
#pragma unroll
          for (int d = 0; d < ND; ++d)
            if (d != dd) {
              y0 *= t0[d].v[mt];
              y1 *= t1[d].v[mt];
            }
          W.wd[mt][0][dd + 1] = y0;
          W.wd[mt][1][dd + 1] = y1;
        }
      }
    }
  };

  // ring position of the stage being consumed (same sequence in every warp)
  long long v = 0;
  int b = 0;
  uint32_t ph = 0;
  for (long long pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    // ---- points of this warp: row r of M tile mt is point base + mt*8 + r ----
    const long long base = pass * pass_pts + (long long)warp * ROWS;
    double A[MT][KS];
    double Ad[JAC ? MT : 1][JAC ? KS : 1];
    bool live[MT];
    long long idx[MT];
    long long first_bad = 0x7fffffffffffffffLL;
    __syncwarp();  // the previous pass has finished reading the tables
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      idx[mt] = base + mt * 8 + r;
      live[mt] = idx[mt] < prm.npts;
      bool ok = true;
      double z[N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const R lb = (R)prm.lb[d], ub = (R)prm.ub[d];
        const R x = live[mt] ? pts_in[idx[mt] * N + d] : lb;
        const bool in = (lb <= x) && (x <= ub);  // eval.jl:64
        ok = ok && in;
        z[d] = in ? (double)rsub(rdiv(rmul(rsub(x, lb), R(2)), rsub(ub, lb)), R(1)) : 0.0;  // eval.jl:65
      }
      if (live[mt] && !ok) {
        live[mt] = false;
        if (idx[mt] < first_bad) first_bad = idx[mt];
      }
      // A fragments: T_k(z1), k = 1 + 4*ks + c  (and T'_k)
      {
        const double z2 = 2.0 * z[0];
        double tkm = 1.0, tk = z[0];  // T_0, T_1
        double dkm = 0.0, dk = 1.0;   // T'_0, T'_1
#pragma unroll
        for (int k = 1; k <= 4 * KS; ++k) {
          if (((k - 1) & 3) == c) {
            A[mt][(k - 1) >> 2] = tk;
            if constexpr (JAC) Ad[mt][(k - 1) >> 2] = dk;
          }
          const double tn = fma(z2, tk, -tkm);
          if constexpr (JAC) {
            const double dn = fma(z2, dk, 2.0 * tk) - dkm;
            dkm = dk;
            dk = dn;
          }
          tkm = tk;
          tk = tn;
        }
      }
      // weight tables of dims 2..N: every lane of the quad runs the recurrence, lane c stores k % 4 == c
      int o = 0;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const double zz = z[d + 1], z2 = 2.0 * zz;
        double tkm = 1.0, tk = zz, dkm = 0.0, dk = 1.0;
        double* tr;
        double* dr;
        if constexpr (PAIRED) {
          tr = tab + ((size_t)r * prm.tabStride + o) * MT + mt;
          dr = dtab + ((size_t)r * prm.tabStride + o) * MT + mt;
        } else {
          tr = tab + (size_t)(mt * 8 + r) * prm.tabStride + o;
          dr = dtab + (size_t)(mt * 8 + r) * prm.tabStride + o;
        }
        const int nd = prm.n[d + 1];
        o += nd;
        if (c == 0) {
          tr[0] = 1.0;
          if constexpr (JAC) dr[0] = 0.0;
        }
        for (int k = 1; k < nd; ++k) {
          if ((k & 3) == c) {
            if constexpr (PAIRED) {
              tr[k * MT] = tk;
              if constexpr (JAC) dr[k * MT] = dk;
            } else {
              tr[k] = tk;
              if constexpr (JAC) dr[k] = dk;
            }
          }
          const double tn = fma(z2, tk, -tkm);
          if constexpr (JAC) {
            const double dn = fma(z2, dk, 2.0 * tk) - dkm;
            dkm = dk;
            dk = dn;
          }
          tkm = tk;
          tk = tn;
        }
      }
    }
    if (first_bad != 0x7fffffffffffffffLL && prm.bad) atomicMin(prm.bad, first_bad + prm.idx_base);
    __syncwarp();

    double acc[MT][EC];
    double accJ[JAC ? MT : 1][JAC ? EC : 1][JAC ? N : 1];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        acc[mt][e] = 0.0;
        if constexpr (JAC) {
#pragma unroll
          for (int d = 0; d < N; ++d) accJ[mt][e][d] = 0.0;
        }
      }

    // Software pipelining of the weighted reduction: the DFMAs that fold a finished G chain into the per-lane sums
    // depend on the chain's last DMMA, and an in-order warp that waits for that latency cannot issue the next
    // (independent) chain.  So chain e is folded only after chain e+1 has been issued; the last chain of a tile is
    // folded after the first chain of the next tile — W still holds that tile's weights then, and the weights of the
    // new tile are loaded right after (their first use is a whole chain away).
    double pg0[MT], pg1[MT];
    [[maybe_unused]] double ph0[JAC ? MT : 1], ph1[JAC ? MT : 1];
    Weights W;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      pg0[mt] = pg1[mt] = 0.0;
      W.w[mt][0] = W.w[mt][1] = 0.0;
      if constexpr (JAC) {
        ph0[mt] = ph1[mt] = 0.0;
#pragma unroll
        for (int d = 0; d < N; ++d) W.wd[mt][0][d] = W.wd[mt][1][d] = 0.0;
      }
    }
    auto fold = [&](auto ec) {  // fold the pending chain (component `ec`) with the weights in W
      constexpr int e = decltype(ec)::value;
#if defined(FC_EXP) && (FC_EXP & 2)  // experiment (bench_micro/dmma_eval_variants.cu): no FP64 fold, results kept live
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        acc[mt][e] = __hiloint2double(__double2hiint(acc[mt][e]) ^ __double2hiint(pg0[mt]),
                                      __double2loint(acc[mt][e]) ^ __double2loint(pg1[mt]));
      return;
#endif
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        acc[mt][e] = fma(pg0[mt], W.w[mt][0], acc[mt][e]);
        acc[mt][e] = fma(pg1[mt], W.w[mt][1], acc[mt][e]);
        if constexpr (JAC) {
          accJ[mt][e][0] = fma(ph0[mt], W.w[mt][0], accJ[mt][e][0]);
          accJ[mt][e][0] = fma(ph1[mt], W.w[mt][1], accJ[mt][e][0]);
#pragma unroll
          for (int d = 1; d < N; ++d) {
            accJ[mt][e][d] = fma(pg0[mt], W.wd[mt][0][d], accJ[mt][e][d]);
            accJ[mt][e][d] = fma(pg1[mt], W.wd[mt][1][d], accJ[mt][e][d]);
          }
        }
      }
    };
    const double* rec = nullptr;
    int tin = 0, nt = 0;  // tile index inside the current stage, tiles in the current stage
    for (int tile = 0; tile < prm.ntiles; ++tile) {
      if (tin == 0) {  // entering a stage
        wait_stage(v, b, ph);
        rec = stages + (size_t)b * stageDoubles;
        const int left = prm.ntiles - tile;
        nt = left < prm.tilesPerStage ? left : prm.tilesPerStage;
      }
      auto chain = [&](auto ec) {
        constexpr int e = decltype(ec)::value;
        const double* u = rec + kDmmaMeta + e * U;
        const double2 c0 = *reinterpret_cast<const double2*>(u + 2 * c);
        double g0[MT], g1[MT];
        [[maybe_unused]] double h0[JAC ? MT : 1], h1[JAC ? MT : 1];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          g0[mt] = c0.x;  // k1 = 0 term: T_0 = 1
          g1[mt] = c0.y;
          if constexpr (JAC) h0[mt] = h1[mt] = 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const double bf = u[8 + ks * 32 + lane];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            dmma884(g0[mt], g1[mt], A[mt][ks], bf);
            if constexpr (JAC) dmma884(h0[mt], h1[mt], Ad[mt][ks], bf);
          }
        }
        // the previous chain has long finished: fold it now, then this chain becomes the pending one
        fold(std::integral_constant<int, (e == 0 ? EC - 1 : e - 1)>{});
#if !(defined(FC_EXP) && (FC_EXP & 1))  // experiment: weights never loaded / multiplied
        if constexpr (e == 0) weights(rec, W);  // the pending chain of the previous tile is folded: W moves on to this tile
#endif
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          pg0[mt] = g0[mt];
          pg1[mt] = g1[mt];
          if constexpr (JAC) {
            ph0[mt] = h0[mt];
            ph1[mt] = h1[mt];
          }
        }
      };
      chain(std::integral_constant<int, 0>{});
      if constexpr (EC > 1) chain(std::integral_constant<int, 1>{});
      if constexpr (EC > 2) chain(std::integral_constant<int, 2>{});
      rec += TR;
      // ---- release the stage after its last tile; the last warp out refills the buffer
      if (++tin == nt) {
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          if (atomicAdd(&released[b], 1u) == (unsigned)ncw - 1u) {
            released[b] = 0;
            if (v + ns < total_visits) {
#ifdef FC_PROFILE
              const long long t0 = clock64();
#endif
              fence_proxy_async();
              issue(v + ns);
#ifdef FC_PROFILE
              prof_issue += clock64() - t0;
#endif
            }
          }
        }
        ++v;
        if (++b == ns) {
          b = 0;
          ph ^= 1u;
        }
        tin = 0;
      }
    }
    fold(std::integral_constant<int, EC - 1>{});  // the last chain of the pass

    // ---- quad reduction and store ----
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < EC; ++e) {
        double x = acc[mt][e];
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        acc[mt][e] = x;
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
#ifdef FC_PROFILE
    if (lane == 0 && blockIdx.x == 0 && pass + gridDim.x >= npass) {
      long long* q = prm.prof + warp * 4;
      q[0] = clock64() - prof_t0;
      q[1] = prof_wait;
      q[2] = prof_slow;
      q[3] = prof_issue;
    }
#endif
    if (c == 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        if (live[mt]) {
#pragma unroll
          for (int e = 0; e < EC; ++e) out_v[idx[mt] * prm.v_stride + prm.e0 + e] = (R)acc[mt][e];
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
                  out_J[idx[mt] * prm.J_stride + prm.e0 + e + (long long)prm.Etot * d] = (R)Js[e][d];
            } else if (prm.mode == 1) {  // pullback real(J' * dy), chainrules.jl:7,14 (accumulated over component groups)
              const R* dy = reinterpret_cast<const R*>(prm.tan) + idx[mt] * prm.Etot + prm.e0;
#pragma unroll
              for (int d = 0; d < N; ++d) {
                double a = prm.e0 == 0 ? 0.0 : (double)out_J[idx[mt] * N + d];
#pragma unroll
                for (int e = 0; e < EC; ++e) a = fma(Js[e][d], (double)dy[e], a);
                out_J[idx[mt] * N + d] = (R)a;
              }
            } else {  // pushforward J * dx, chainrules.jl:24
              const R* dx = reinterpret_cast<const R*>(prm.tan) + idx[mt] * N;
#pragma unroll
              for (int e = 0; e < EC; ++e) {
                double a = 0.0;
#pragma unroll
                for (int d = 0; d < N; ++d) a = fma(Js[e][d], (double)dx[d], a);
                out_J[idx[mt] * prm.Etot + prm.e0 + e] = (R)a;
              }
            }
          }
        }
    }
  }
}

// Julia layout -> tile-record stream of one component group (see eval_dmma.cuh).
struct FragDims {
  int nd;                 // weighted dims (ndim - 1)
  int n[kMaxNdim];        // n2..nN
  int toff[kMaxNdim];     // offset of each dim's table inside a weight-table row
};
template <typename S>
__global__ void __launch_bounds__(256) relayout_frag_kernel(const S* __restrict__ src, double* __restrict__ dst,
                                                            long long total, int E, int e0, int ec, int n1, int M,
                                                            int KS, const FragDims fd) {
  const int U = KS * 32 + 8;
  const int TR = kDmmaMeta + ec * U;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long tile = g / TR;
    const int w = (int)(g - tile * TR);
    if (w < kDmmaMeta) {
      // two int32 per double slot: meta[d*8 + col]
      int v2[2];
      for (int h = 0; h < 2; ++h) {
        const int i = 2 * w + h, d = i >> 3, col = i & 7;
        int val = 0;
        if (d < fd.nd) {
          long long m = tile * 8 + col;
          if (m >= M) m = M - 1;  // padding column: its coefficients are zero, any valid entry will do
          for (int q = 0; q < d; ++q) m /= fd.n[q];
          val = fd.toff[d] + (int)(d == fd.nd - 1 ? m : m % fd.n[d]);
        }
        v2[h] = val;
      }
      reinterpret_cast<int2*>(dst)[g] = make_int2(v2[0], v2[1]);
      continue;
    }
    const int e = (w - kDmmaMeta) / U, u = (w - kDmmaMeta) % U;
    int k1;
    long long m;
    if (u < 8) {
      k1 = 0;
      m = tile * 8 + u;
    } else {
      const int l = (u - 8) & 31, ks = (u - 8) >> 5;
      k1 = 1 + 4 * ks + (l & 3);
      m = tile * 8 + (l >> 2);
    }
    dst[g] = (k1 < n1 && m < M) ? (double)src[(e0 + e) + (long long)E * (k1 + (long long)n1 * m)] : 0.0;
  }
}

}  // namespace fastcheb
