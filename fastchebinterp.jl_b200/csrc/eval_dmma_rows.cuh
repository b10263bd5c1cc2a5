// FP64 tensor-core (DMMA) evaluation of a 3-D / 4-D ChebPoly, values only, with the column tiles aligned to the rows
// of dimension 2 ("row" = all k2 for one (k3[,k4])).
//
// Same mathematics as dmma_kernel (eval_dmma.cuh; reference /root/reference/src/eval.jl:16-56), regrouped:
//
//     v[p,e] = sum_{rho=(k3,k4)} w_o[p,rho] * ( sum_{k2} T_{k2}(z2_p) * G[p,k2,rho,e] ),   w_o = T_{k3}(z3_p) T_{k4}(z4_p)
//     G[p,k2,rho,e] = sum_{k1} T_{k1}(z1_p) C[k1,k2,rho,e]                                  (the GEMM, on the tensor cores)
//
// In dmma_kernel a tile is 8 consecutive columns of the flattened index m = (k2,k3,..), so every tile needs its own
// weights: per tile a metadata load, 4-8 table loads and 2-4 multiplies per lane — a third of the issue slots when a
// tile is only 16 DMMAs (scalar-valued 3-D: 77 % of FP64 peak; 4-D complex: 75 %).  Here a row is cut into TPR full
// tiles plus 0-2 trailing columns (n2 = 8 TPR + {0,1,2}: every power-of-two order), so the inner weights T_{k2}(z2) of
// a lane's two columns are the same for every row and live in REGISTERS for the whole pass (2 TPR doubles per M tile),
// and the outer weight is one table load (and one multiply in 4-D) per ROW.  The trailing columns are contracted with
// DFMAs from the A fragments already in registers (as in eval_dmma2d.cuh), which also fills the latency gap before the
// row's last accumulator chain can be folded.  Per tile that leaves the B-fragment loads, the DMMAs and 2 DFMAs per
// chain and M tile.
//
// Stream (built once per handle by relayout_rows_kernel): for every row, TPR tile units per component
//     [ 8 coefficients k1 = 0 | KS x 32 lane values of the mma.sync B fragment ]          (as in eval_dmma.cuh, no metadata)
// laid out [tile][component], followed by the trailing columns [component][column][ c0 | 4 KS coefficients k1 = 1.. ]
// (padded to 16 bytes).  Whole rows stream through the same shared-memory ring as in dmma_kernel (cp.async.bulk ->
// UBLKCP, mbarriers, last warp out refills).
#pragma once
#include "eval_dmma.cuh"

namespace fastcheb {

struct DmmaRowsParams {
  const double* stream;
  const void* pts;
  void* out_v;
  long long* bad;
  long long npts, idx_base;
  long long v_stride;
  int e0;
  int n[kMaxNdim];
  int nrows;         // prod n3..nN
  int ntail;         // trailing columns of a row (0..2)
  int rowDoubles;    // doubles per row record
  int rowsPerStage;  // rows per shared-memory stage
  int nstage;
  int tabStride;  // doubles per point row of the outer weight tables (dims 3..N)
  double lb[kMaxNdim], ub[kMaxNdim];
};

__host__ __device__ constexpr int dmma_rows_tail_doubles(int EC, int KS, int ntail) {
  return (EC * ntail * (4 * KS + 1) + 1) / 2 * 2;
}
__host__ __device__ constexpr int dmma_rows_row_doubles(int EC, int KS, int TPR, int ntail) {
  return TPR * EC * (KS * 32 + 8) + dmma_rows_tail_doubles(EC, KS, ntail);
}
constexpr int kDmmaRowsThreads = 512;

template <typename R, int N, int EC, int KS, int TPR>
__global__ void __launch_bounds__(kDmmaRowsThreads) dmma_rows_kernel(const __grid_constant__ DmmaRowsParams prm) {
  static_assert(N == 3 || N == 4, "rows kernel: 3-D and 4-D");
  constexpr int MT = 2;
  constexpr int ROWS = MT * 8;
  constexpr int U = KS * 32 + 8;
  constexpr int ND = N - 2;  // outer dimensions 3..N
  const R* const pts_in = reinterpret_cast<const R*>(prm.pts);
  R* const out_v = reinterpret_cast<R*>(prm.out_v);
  extern __shared__ __align__(128) unsigned char fastcheb_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(fastcheb_smem);
  unsigned int* released = reinterpret_cast<unsigned int*>(full + 8);
  unsigned int* ready = released + 8;
  double* stages = reinterpret_cast<double*>(fastcheb_smem + 128);
  const int stageDoubles = prm.rowsPerStage * prm.rowDoubles;
  double* tables = stages + (size_t)prm.nstage * stageDoubles;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncw = blockDim.x >> 5;
  const long long pass_pts = (long long)ncw * ROWS;
  const long long npass = (prm.npts + pass_pts - 1) / pass_pts;
  if ((long long)blockIdx.x >= npass) return;
  const long long my_pass = (npass - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int S = (prm.nrows + prm.rowsPerStage - 1) / prm.rowsPerStage;  // stages per pass
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

  // ring protocol of dmma_kernel: visits 0..nstage-1 are issued in the prologue, the last warp to release a stage
  // buffer refills it with visit v + nstage
  const long long total_visits = my_pass * S;
  auto issue = [&](long long u) {
    const int b = (int)(u % ns);
    const int s = (int)(u % S);
    const int r0 = s * prm.rowsPerStage;
    const int nr = prm.nrows - r0 < prm.rowsPerStage ? prm.nrows - r0 : prm.rowsPerStage;
    const uint32_t bytes = (uint32_t)nr * prm.rowDoubles * 8u;
    const char* src = reinterpret_cast<const char*>(prm.stream) + (size_t)r0 * prm.rowDoubles * 8u;
    char* dst = reinterpret_cast<char*>(stages + (size_t)b * stageDoubles);
    mbar_arrive_expect_tx(&full[b], bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t len = bytes - o < 32768u ? bytes - o : 32768u;
      bulk_g2s(dst + o, src + o, len, &full[b]);
    }
  };
  if (threadIdx.x == 0)
    for (long long u = 0; u < ns && u < total_visits; ++u) issue(u);
  auto wait_stage = [&](long long u, int bb, uint32_t parity) {
    const unsigned tag = (unsigned)(u + 1);
    unsigned seen;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&ready[bb])) : "memory");
    if (seen != tag) {
      mbar_wait(&full[bb], parity);
      asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&ready[bb])), "r"(tag) : "memory");
    }
  };

  const int c = lane & 3;   // column pair inside a tile / k index inside a k-step
  const int r = lane >> 2;  // point row inside an M tile
  double* tab = tables + (size_t)warp * ROWS * prm.tabStride;  // [point row][k3.. | k4..]
  const int n3 = prm.n[2];

  long long v = 0;
  int b = 0;
  uint32_t ph = 0;
  for (long long pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const long long base = pass * pass_pts + (long long)warp * ROWS;
    double A[MT][KS];      // T_k(z1), k = 1 + 4 ks + c
    double W2[MT][TPR][2];  // T_k(z2), k = 8 t + 2c + {0,1}
    double Wt[MT][2];       // T_k(z2) of the trailing columns
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
      {  // A fragments
        const double z2 = 2.0 * z[0];
        double tkm = 1.0, tk = z[0];
#pragma unroll
        for (int k = 1; k <= 4 * KS; ++k) {
          if (((k - 1) & 3) == c) A[mt][(k - 1) >> 2] = tk;
          const double tn = fma(z2, tk, -tkm);
          tkm = tk;
          tk = tn;
        }
      }
      {  // inner weights: the lane keeps T_k(z2) of its own columns, k = 8t + 2c + h (and of the trailing columns)
        const double z2 = 2.0 * z[1];
        double tkm = 1.0, tk = z[1];  // T_0, T_1
        Wt[mt][0] = Wt[mt][1] = 0.0;
#pragma unroll
        for (int k = 0; k < 8 * TPR + 2; ++k) {
          const double cur = k == 0 ? 1.0 : tk;  // T_k
          if (k < 8 * TPR) {
            if (((k & 7) >> 1) == c) W2[mt][k >> 3][k & 1] = cur;
          } else {
            Wt[mt][k - 8 * TPR] = cur;
          }
          if (k >= 1) {
            const double tn = fma(z2, tk, -tkm);
            tkm = tk;
            tk = tn;
          }
        }
      }
      // outer weight tables (dims 3..N): every lane of the quad runs the recurrence, lane c stores k % 4 == c
      int o = 0;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const double zz = z[d + 2], z2 = 2.0 * zz;
        double tkm = 1.0, tk = zz;
        double* tr = tab + (size_t)(mt * 8 + r) * prm.tabStride + o;
        const int nd = prm.n[d + 2];
        o += nd;
        if (c == 0) tr[0] = 1.0;
        for (int k = 1; k < nd; ++k) {
          if ((k & 3) == c) tr[k] = tk;
          const double tn = fma(z2, tk, -tkm);
          tkm = tk;
          tk = tn;
        }
      }
    }
    if (first_bad != 0x7fffffffffffffffLL && prm.bad) atomicMin(prm.bad, first_bad + prm.idx_base);
    __syncwarp();

    double acc[MT][EC];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < EC; ++e) acc[mt][e] = 0.0;
    const double* trow[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) trow[mt] = tab + (size_t)(mt * 8 + r) * prm.tabStride;

    const double* rec = nullptr;
    int rin = 0, nr = 0;  // row index inside the current stage, rows in the current stage
    int k3 = 0, k4 = 0;
    for (int row = 0; row < prm.nrows; ++row) {
      if (rin == 0) {  // entering a stage
        wait_stage(v, b, ph);
        rec = stages + (size_t)b * stageDoubles;
        const int left = prm.nrows - row;
        nr = left < prm.rowsPerStage ? left : prm.rowsPerStage;
      }
      // outer weight of this row
      double wo[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        wo[mt] = trow[mt][k3];
        if constexpr (N == 4) wo[mt] *= trow[mt][n3 + k4];
      }
      double rs[MT][EC];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int e = 0; e < EC; ++e) rs[mt][e] = 0.0;
      // chains of the row: chain (t, e) is folded after the next one has been issued (see dmma_kernel)
      double pg0[MT], pg1[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) pg0[mt] = pg1[mt] = 0.0;
#pragma unroll
      for (int t = 0; t < TPR; ++t) {
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          const double* u = rec + (size_t)(t * EC + e) * U;
          const double2 c0 = *reinterpret_cast<const double2*>(u + 2 * c);
          double g0[MT], g1[MT];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            g0[mt] = c0.x;  // k1 = 0 term: T_0 = 1
            g1[mt] = c0.y;
          }
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const double bf = u[8 + ks * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) dmma884(g0[mt], g1[mt], A[mt][ks], bf);
          }
          if (t > 0 || e > 0) {  // fold the previous chain (tp, ep) of this row
            const int tp = e == 0 ? t - 1 : t, ep = e == 0 ? EC - 1 : e - 1;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              rs[mt][ep] = fma(pg0[mt], W2[mt][tp][0], rs[mt][ep]);
              rs[mt][ep] = fma(pg1[mt], W2[mt][tp][1], rs[mt][ep]);
            }
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            pg0[mt] = g0[mt];
            pg1[mt] = g1[mt];
          }
        }
      }
      // trailing columns on the CUDA cores (independent of the pending chain: covers its latency)
      {
        const double* tl = rec + (size_t)TPR * EC * U;
        for (int j = 0; j < prm.ntail; ++j) {
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            const double* q = tl + (size_t)(e * prm.ntail + j) * (4 * KS + 1);
            double s[MT];
            const double c0 = c == 0 ? q[0] : 0.0;  // counted once per quad
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) s[mt] = c0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              const double bf = q[1 + 4 * ks + c];
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) s[mt] = fma(A[mt][ks], bf, s[mt]);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) rs[mt][e] = fma(s[mt], Wt[mt][j], rs[mt][e]);
          }
        }
      }
      // the last chain of the row, then the row joins the point sums with its outer weight
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        rs[mt][EC - 1] = fma(pg0[mt], W2[mt][TPR - 1][0], rs[mt][EC - 1]);
        rs[mt][EC - 1] = fma(pg1[mt], W2[mt][TPR - 1][1], rs[mt][EC - 1]);
#pragma unroll
        for (int e = 0; e < EC; ++e) acc[mt][e] = fma(rs[mt][e], wo[mt], acc[mt][e]);
      }
      rec += prm.rowDoubles;
      if (++k3 == n3) {
        k3 = 0;
        ++k4;
      }
      // ---- release the stage after its last row; the last warp out refills the buffer
      if (++rin == nr) {
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          if (atomicAdd(&released[b], 1u) == (unsigned)ncw - 1u) {
            released[b] = 0;
            if (v + ns < total_visits) {
              fence_proxy_async();
              issue(v + ns);
            }
          }
        }
        ++v;
        if (++b == ns) {
          b = 0;
          ph ^= 1u;
        }
        rin = 0;
      }
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
      }
    if (c == 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        if (live[mt]) {
#pragma unroll
          for (int e = 0; e < EC; ++e) out_v[idx[mt] * prm.v_stride + prm.e0 + e] = (R)acc[mt][e];
        }
    }
  }
}

// Julia layout -> row-aligned stream of one component group (layout: see the head of this file).
template <typename S>
__global__ void __launch_bounds__(256) relayout_rows_kernel(const S* __restrict__ src, double* __restrict__ dst,
                                                            long long total, int E, int e0, int ec, int n1, int n2,
                                                            int KS, int TPR, int ntail, int rowDoubles) {
  const int U = KS * 32 + 8;
  const int tileDoubles = TPR * ec * U;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long row = g / rowDoubles;
    const int w = (int)(g - row * rowDoubles);
    int k1 = -1, k2 = 0, e = 0;
    if (w < tileDoubles) {
      const int t = w / (ec * U), rem = w % (ec * U);
      e = rem / U;
      const int u = rem % U;
      if (u < 8) {
        k1 = 0;
        k2 = 8 * t + u;
      } else {
        const int l = (u - 8) & 31, ks = (u - 8) >> 5;
        k1 = 1 + 4 * ks + (l & 3);
        k2 = 8 * t + (l >> 2);
      }
    } else {
      const int x = w - tileDoubles, per = 4 * KS + 1;
      if (x < ec * ntail * per) {
        e = x / (ntail * per);
        const int j = (x / per) % ntail;
        k1 = x % per;  // 0: the k1 = 0 coefficient, then k1 = 1..4 KS
        k2 = 8 * TPR + j;
      }
    }
    dst[g] = (k1 >= 0 && k1 < n1 && k2 < n2) ? (double)src[(e0 + e) + (long long)E * (k1 + (long long)n1 * (k2 + (long long)n2 * row))] : 0.0;
  }
}

}  // namespace fastcheb
