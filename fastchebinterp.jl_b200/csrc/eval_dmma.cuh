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
// Layout: the handle stores the coefficients in B-fragment order — per 8-column tile and component
// a "unit" of 8 + 32*KS doubles: the 8 k1=0 coefficients, then for each k-step the 32 lane values
// B[k = lane%4][col = lane/4] = C[1 + 4*ks + lane%4, 8*tile + lane/4, e] — so a warp's B load is one
// conflict-free LDS.64 and a stage of tiles is one contiguous bulk copy.  One elected thread streams
// stages from L2 into a shared-memory ring with cp.async.bulk + full/empty mbarriers; all warps
// (16 points each) consume the same stage.
#pragma once
#include <cstdint>
#include "ptx_util.cuh"

namespace fastcheb {

struct DmmaParams {
  const double* frag;  // [tile][ec][U] fragment stream of this component group
  const double* pts;
  double* out_v;
  double* out_J;
  long long* bad;
  long long npts, idx_base;
  long long v_stride, J_stride;
  int e0, Etot;
  int n[kMaxNdim];
  int ntiles;         // 8-column tiles per pass
  int M;              // prod n2..nN (valid columns)
  int tilesPerStage;  // tiles per shared-memory stage
  int nstage;
  int tabStride;  // doubles per point row of the weight tables
  double lb[kMaxNdim], ub[kMaxNdim];
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kDmmaMT = 2;  // 8-point M tiles per consumer warp

template <int N, int EC, int KS, bool JAC>
__global__ void __launch_bounds__(JAC ? 256 : 512) dmma_kernel(const __grid_constant__ DmmaParams prm) {
  constexpr int MT = kDmmaMT;
  constexpr int U = KS * 32 + 8;
  constexpr int ROWS = MT * 8;
  extern __shared__ __align__(128) unsigned char fastcheb_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(fastcheb_smem);
  uint64_t* empty = full + 8;
  double* stages = reinterpret_cast<double*>(fastcheb_smem + 128);
  const int stageDoubles = prm.tilesPerStage * EC * U;
  double* tables = stages + (size_t)prm.nstage * stageDoubles;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncw = blockDim.x >> 5;  // every warp consumes
  const long long pass_pts = (long long)ncw * ROWS;
  const long long npass = (prm.npts + pass_pts - 1) / pass_pts;
  if ((long long)blockIdx.x >= npass) return;
  const long long my_pass = (npass - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int S = (prm.ntiles + prm.tilesPerStage - 1) / prm.tilesPerStage;  // stages per pass

  if (threadIdx.x == 0) {
    for (int b = 0; b < prm.nstage; ++b) {
      mbar_init(&full[b], 1);
      mbar_init(&empty[b], ncw);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // The producer role is folded into lane 0 of warp 0 (all warps compute): visits 0..nstage-1 are
  // issued here; the refill for visit u (>= nstage) is issued at the start of iteration u-nstage+2,
  // after every warp has released visit u-nstage (two iterations of slack, so the wait rarely spins).
  const long long total_visits = my_pass * S;
  auto issue = [&](long long u) {
    const int b = (int)(u % prm.nstage);
    const int s = (int)(u % S);
    const int t0 = s * prm.tilesPerStage;
    const int nt = prm.ntiles - t0 < prm.tilesPerStage ? prm.ntiles - t0 : prm.tilesPerStage;
    const uint32_t bytes = (uint32_t)nt * EC * U * 8u;
    const char* src = reinterpret_cast<const char*>(prm.frag) + (size_t)t0 * EC * U * 8u;
    char* dst = reinterpret_cast<char*>(stages + (size_t)b * stageDoubles);
    mbar_arrive_expect_tx(&full[b], bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t len = bytes - o < 32768u ? bytes - o : 32768u;
      bulk_g2s(dst + o, src + o, len, &full[b]);
    }
  };
  if (threadIdx.x == 0)
    for (long long u = 0; u < prm.nstage && u < total_visits; ++u) issue(u);

  // ---------------- consumers ----------------
  const int c = lane & 3;     // column pair inside a tile / k index inside a k-step
  const int r = lane >> 2;    // point row inside an M tile
  constexpr int ND = N - 1;   // weighted dimensions 2..N
  constexpr int NTAB = JAC ? 2 : 1;
  double* tab = tables + (size_t)warp * ROWS * prm.tabStride * NTAB;  // [row][tabStride] (+ derivative copy)
  double* dtab = tab + (size_t)ROWS * prm.tabStride;
  int toff[ND];
  {
    int o = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      toff[d] = o;
      o += prm.n[d + 1];
    }
  }

  long long v = 0;  // stage visit counter (same sequence as the producer's)
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
        const double lb = prm.lb[d], ub = prm.ub[d];
        const double x = live[mt] ? prm.pts[idx[mt] * N + d] : lb;
        const bool in = (lb <= x) && (x <= ub);  // eval.jl:64
        ok = ok && in;
        z[d] = in ? __dsub_rn(__ddiv_rn(__dmul_rn(__dsub_rn(x, lb), 2.0), __dsub_rn(ub, lb)), 1.0) : 0.0;  // eval.jl:65
      }
      if (live[mt] && !ok) {
        live[mt] = false;
        if (idx[mt] < first_bad) first_bad = idx[mt];
      }
      // A fragments: T_k(z1), k = 1 + 4*ks + c  (and T'_k)
      {
        const double z2 = 2.0 * z[0];
        double tkm = 1.0, tk = z[0];     // T_0, T_1
        double dkm = 0.0, dk = 1.0;      // T'_0, T'_1
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
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const double zz = z[d + 1], z2 = 2.0 * zz;
        double tkm = 1.0, tk = zz, dkm = 0.0, dk = 1.0;
        double* trow = tab + (size_t)(mt * 8 + r) * prm.tabStride + toff[d];
        double* drow = dtab + (size_t)(mt * 8 + r) * prm.tabStride + toff[d];
        const int nd = prm.n[d + 1];
        if (c == 0) {
          trow[0] = 1.0;
          if constexpr (JAC) drow[0] = 0.0;
        }
        for (int k = 1; k < nd; ++k) {
          if ((k & 3) == c) {
            trow[k] = tk;
            if constexpr (JAC) drow[k] = dk;
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

    // column counters of the two slots this lane owns: m = 8*tile + 2*c + slot
    int kk[2][ND];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      int m = 2 * c + sl;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const int nd = prm.n[d + 1];
        if (d < ND - 1) {
          kk[sl][d] = m % nd;
          m /= nd;
        } else {
          kk[sl][d] = m;  // last dimension keeps the overflow (padding columns)
        }
      }
    }

    int tile = 0;
    for (int s = 0; s < S; ++s, ++v) {
      const int b = (int)(v % prm.nstage);
      if (threadIdx.x == 0 && v >= 2) {
        const long long u = v - 2 + prm.nstage;
        if (u < total_visits) {
          const int bb = (int)((v - 2) % prm.nstage);
          mbar_wait(&empty[bb], (uint32_t)(((v - 2) / prm.nstage) & 1));
          issue(u);
        }
      }
      __syncwarp();
      mbar_wait(&full[b], (uint32_t)((v / prm.nstage) & 1));
      const double* sb = stages + (size_t)b * stageDoubles;
      const int nt = prm.ntiles - tile < prm.tilesPerStage ? prm.ntiles - tile : prm.tilesPerStage;
      for (int t = 0; t < nt; ++t, ++tile) {
        // weights of this tile's two columns for each M tile
        double w[MT][2];
        double wd[JAC ? MT : 1][2][JAC ? N : 1];  // [..][d]: weight with T' in dimension d+1 (d >= 1)
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          int ki[ND];
#pragma unroll
          for (int d = 0; d < ND; ++d) {
            const int nd = prm.n[d + 1];
            ki[d] = kk[sl][d] < nd ? kk[sl][d] : nd - 1;  // clamp padding columns (their coefficients are 0)
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const double* trow = tab + (size_t)(mt * 8 + r) * prm.tabStride;
            double tv[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) tv[d] = trow[toff[d] + ki[d]];
            double ww = tv[0];
#pragma unroll
            for (int d = 1; d < ND; ++d) ww *= tv[d];
            w[mt][sl] = ww;
            if constexpr (JAC) {
              const double* drow = dtab + (size_t)(mt * 8 + r) * prm.tabStride;
#pragma unroll
              for (int dd = 0; dd < ND; ++dd) {
                double x = drow[toff[dd] + ki[dd]];
#pragma unroll
                for (int d = 0; d < ND; ++d)
                  if (d != dd) x *= tv[d];
                wd[mt][sl][dd + 1] = x;
              }
            }
          }
          // advance this slot's column counters by 8
          kk[sl][0] += 8;
#pragma unroll
          for (int d = 0; d < ND - 1; ++d) {
            const int nd = prm.n[d + 1];
            while (kk[sl][d] >= nd) {
              kk[sl][d] -= nd;
              kk[sl][d + 1] += 1;
            }
          }
        }
        const double* tb = sb + (size_t)t * EC * U;
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          const double* u = tb + e * U;
          const double2 c0 = *reinterpret_cast<const double2*>(u + 2 * c);
          double g0[MT], g1[MT];
          double h0[JAC ? MT : 1], h1[JAC ? MT : 1];
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
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            acc[mt][e] = fma(g0[mt], w[mt][0], acc[mt][e]);
            acc[mt][e] = fma(g1[mt], w[mt][1], acc[mt][e]);
            if constexpr (JAC) {
              accJ[mt][e][0] = fma(h0[mt], w[mt][0], accJ[mt][e][0]);
              accJ[mt][e][0] = fma(h1[mt], w[mt][1], accJ[mt][e][0]);
#pragma unroll
              for (int d = 1; d < N; ++d) {
                accJ[mt][e][d] = fma(g0[mt], wd[mt][0][d], accJ[mt][e][d]);
                accJ[mt][e][d] = fma(g1[mt], wd[mt][1][d], accJ[mt][e][d]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
    }

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
    if (c == 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        if (live[mt]) {
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            prm.out_v[idx[mt] * prm.v_stride + prm.e0 + e] = acc[mt][e];
            if constexpr (JAC) {
#pragma unroll
              for (int d = 0; d < N; ++d)
                prm.out_J[idx[mt] * prm.J_stride + prm.e0 + e + (long long)prm.Etot * d] =
                    __ddiv_rn(__dmul_rn(accJ[mt][e][d], 2.0), __dsub_rn(prm.ub[d], prm.lb[d]));  // eval.jl:151
            }
          }
        }
    }
  }
}

}  // namespace fastcheb
