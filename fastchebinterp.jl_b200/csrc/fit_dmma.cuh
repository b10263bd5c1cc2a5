// FP64 tensor-core DCT-I for batches of short 1-D fits — the bandwidth-bound fit path.
//
// Replaces the arithmetic of chebcoefs for many independent lines (/root/reference/src/interp.jl:39-57:
// FFTW.r2r(vals, REDFT00) :43-44, the 1/(2(n-1)) scale :49 and the interior doubling :50-54), i.e.
// BASELINE.json configs[4] (10^5 independent order-64 fits).  Algorithmic traffic is one read of `vals`
// and one write of `coefs`; everything else (fold, transform, normalisation, finite check) happens on
// the way through shared memory.
//
// Transform: with N = n-1 and REDFT00  X_k = x_0 + (-1)^k x_N + 2 sum_{0<j<N} x_j cos(pi j k/N),
// cos(pi k (N-j)/N) = (-1)^k cos(pi k j/N), so the even-k outputs depend only on a_j = x_j + x_{N-j}
// and the odd-k outputs only on b_j = x_j - x_{N-j} (j <= N/2): two half-size dense transforms
//     C[2l]   = sum_j Me[l][j] a_j ,   C[2l+1] = sum_j Mo[l][j] b_j .
// The even half is itself a REDFT00 of the sequence a_0..a_{N/2}, so when 4 | N it is folded once more
// (k = 0 mod 4 from a_j + a_{N/2-j}, k = 2 mod 4 from a_j - a_{N/2-j}): 55 instead of 77 DMMAs per 8
// lines at n = 65.  The matrices (built on the host from exactly reduced long-double cosines, normalisation folded in)
// are the A operands of mma.sync m8n8k4 (SASS DMMA); 8 lines form the N extent of one MMA.  This works
// for every n (no power-of-two requirement) and is a direct sum, so a coefficient carries ~sqrt(n) eps
// of rounding, not an FFT's log n stages.
//
// Data movement: a warp owns a ring of shared-memory slots; one slot holds a unit of 8*NT lines that is
// contiguous in global memory, fetched with one cp.async.bulk (TMA engine) signalled on the warp's own
// mbarrier, transformed in place, and written back with one cp.async.bulk store.  Warps never
// synchronise with each other: no CTA barrier after the prologue.
//
// Shared-memory banks: lines are n reals apart; for n = 2^m + 1 (17, 33, 65, 129) that is 1 (mod 16), so
// MMA column q of line group t is mapped to line 2*NT*(q%4) + q/4 + 2t — the 16 lanes of a half warp
// then read 16 distinct 8-byte banks in the fold, and write 16 distinct banks in the epilogue.
#pragma once
#include <cstdint>
#include "ptx_util.cuh"

namespace fastcheb {

struct FitDmmaParams {
  const void* src;
  void* dst;
  const double* afrag;  // A fragments of parts O, B, C back to back, each [m-tile][k-step][32 lanes]
  long long nunits;     // full units (the host sends the remainder through the generic kernel)
  long long* bad;       // first non-finite flat real index (atomicMin), may be null
  long long idx_base;   // flat index of src[0] in the caller's array (chunked host calls)
  double* linemax;      // optional: max |c| of every line (nunits * 8*NT doubles, dead lines 0)
  int n, E;
  int U;   // fits per unit
  int LU;  // live lines per unit = U*E <= 8*NT
  int nslots;
  int slotBytes;  // unit bytes rounded up to 128
  int unitBytes;
};

__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Parts of the folded transform.  Every part is a small dense product on the tensor cores:
//   part O  (KO k-steps): k = 2l+1          from b_j  = x_j - x_{N-j}
//   part B  (KB k-steps): k = 2l            from a_j  = x_j + x_{N-j}                (one fold,  KC == 0)
//                         k = 4l            from aa_j = a_j + a_{N/2-j}              (two folds, KC  > 0)
//   part C  (KC k-steps): k = 4l+2          from ab_j = a_j - a_{N/2-j}              (two folds only)
// K* are compile-time, so a unit is one straight-line block with immediate shared-memory offsets.
// E1: scalar elements (E == 1).
template <typename R, int KO, int KB, int KC, int NT, bool E1>
__global__ void __launch_bounds__(NT == 1 ? 768 : 256) fit_dmma_kernel(const __grid_constant__ FitDmmaParams prm) {
  constexpr bool L2 = KC > 0;
  constexpr int MO = (KO + 1) / 2, MB = (KB + 1) / 2, MC = (KC + 1) / 2;  // 8-row m-tiles: ceil(4 K / 8)
  constexpr int KMAX = KO > KB ? KO : KB;
  constexpr int SB = L2 ? 4 : 2, OB = 0;  // output stride / offset of part B; part C: stride 4, offset 2
  extern __shared__ __align__(128) unsigned char fastcheb_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int S = prm.nslots;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fastcheb_smem);  // [nwarps][S], <= 128 entries
  constexpr int afragDoubles = (MO * KO + MB * KB + MC * KC) * 32;
  double* afrag = reinterpret_cast<double*>(fastcheb_smem + 1024);
  unsigned char* ring = fastcheb_smem + 1024 + ((afragDoubles * 8 + 127) / 128) * 128 + (size_t)warp * S * prm.slotBytes;

  for (int i = threadIdx.x; i < afragDoubles; i += blockDim.x) afrag[i] = prm.afrag[i];
  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[warp * S + s], 1);
    fence_mbar_init();
  }
  __syncthreads();  // the only CTA-wide barrier

  const long long gw = (long long)blockIdx.x * nwarps + warp;
  const long long tw = (long long)gridDim.x * nwarps;
  if (gw >= prm.nunits) return;
  const long long my_units = (prm.nunits - gw + tw - 1) / tw;
  const char* src = reinterpret_cast<const char*>(prm.src);
  char* dst = reinterpret_cast<char*>(prm.dst);
  auto issue = [&](long long i) {  // lane 0 only
    const int s = (int)(i % S);
    const long long u = gw + i * tw;
    mbar_arrive_expect_tx(&bars[warp * S + s], (uint32_t)prm.unitBytes);
    bulk_g2s(ring + (size_t)s * prm.slotBytes, src + u * (long long)prm.unitBytes, (uint32_t)prm.unitBytes,
             &bars[warp * S + s]);
  };
  if (lane == 0)
    for (long long i = 0; i < S - 1 && i < my_units; ++i) issue(i);

  const int n = prm.n, N = n - 1, H = N / 2, E = E1 ? 1 : prm.E;
  const int c = lane & 3, q = lane >> 2;
  // B-operand line of this lane in every line group, D-operand lines (columns 2c, 2c+1).  Dead lines (a unit
  // may hold fewer than 8*NT lines when E does not divide it) alias line 0: computed, never stored.
  auto line_of = [&](int t, int col) { return 2 * NT * (col & 3) + (col >> 2) + 2 * t; };
  auto line_base = [&](int L) {
    const int f = L / E;
    return (f * n) * E + (L - f * E);
  };
  int bbase[NT], dbase[NT][2];
  bool dlive[NT][2];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int L = line_of(t, q);
    bbase[t] = line_base(L < prm.LU ? L : 0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int Ld = line_of(t, 2 * c + h);
      dlive[t][h] = Ld < prm.LU;
      dbase[t][h] = line_base(dlive[t][h] ? Ld : 0);
    }
  }
  const double* aO = afrag + lane;
  const double* aB = aO + MO * KO * 32;
  const double* aC = aB + MB * KB * 32;
  const int E4 = 4 * E;
  // output element of row q in m-tile 0 of each part, and validity of the last m-tile's rows (earlier ones are full)
  const int kO0 = 2 * q + 1, kB0 = SB * q + OB, kC0 = 4 * q + 2;
  const bool lastO = 2 * (8 * (MO - 1) + q) + 1 <= N, lastB = SB * (8 * (MB - 1) + q) + OB <= N;
  const bool lastC = L2 && 4 * (8 * (MC - 1) + q) + 2 <= N;

  for (long long i = 0; i < my_units; ++i) {
    const int s = (int)(i % S);
    const uint32_t parity = (uint32_t)((i / S) & 1);
    const long long u = gw + i * tw;
    R* x = reinterpret_cast<R*>(ring + (size_t)s * prm.slotBytes);
    mbar_wait(&bars[warp * S + s], parity);

    double accO[NT][MO][2], accB[NT][MB][2], accC[NT][L2 ? MC : 1][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
      for (int mt = 0; mt < MO; ++mt) accO[t][mt][0] = accO[t][mt][1] = 0.0;
#pragma unroll
      for (int mt = 0; mt < MB; ++mt) accB[t][mt][0] = accB[t][mt][1] = 0.0;
#pragma unroll
      for (int mt = 0; mt < MC; ++mt) accC[t][mt][0] = accC[t][mt][1] = 0.0;
    }
    // fold pointers for j = 4 ks + c: x_j walks up and x_{N-j} down (and x_{H-j}, x_{H+j} for the second fold).
    // Beyond j = N/2 (possible only in the last k-step) the matrix columns are zero; the operand only has to be
    // finite, so those lanes read element 0.
    const R* p1[NT];
    const R* p2[NT];
    const R* p3[NT];
    const R* p4[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      p1[t] = x + bbase[t] + c * E;
      p2[t] = x + bbase[t] + (N - c) * E;
      p3[t] = x + bbase[t] + (H - c) * E;
      p4[t] = x + bbase[t] + (H + c) * E;
    }
#pragma unroll
    for (int ks = 0; ks < KMAX; ++ks) {
      const bool inr = 4 * ks + c <= H;
      double fo[NT], fb[NT], fc[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const R* z = x + bbase[t];
        const double x1 = (double)*(inr ? p1[t] : z), x2 = (double)*(inr ? p2[t] : z);
        fo[t] = __dsub_rn(x1, x2);
        fb[t] = __dadd_rn(x1, x2);  // a self-paired element (2j == N) counts twice: its matrix entry carries weight 1
        fc[t] = 0.0;
        if (L2 && ks < KB) {
          const double x3 = (double)*(inr ? p3[t] : z), x4 = (double)*(inr ? p4[t] : z);
          const double a2 = __dadd_rn(x3, x4);
          fc[t] = __dsub_rn(fb[t], a2);
          fb[t] = __dadd_rn(fb[t], a2);
        }
        p1[t] += E4;
        p2[t] -= E4;
        p3[t] -= E4;
        p4[t] += E4;
      }
      if (ks < KB) {
#pragma unroll
        for (int mt = 0; mt < MB; ++mt) {
          const double a = aB[(mt * KB + ks) * 32];
#pragma unroll
          for (int t = 0; t < NT; ++t) dmma884(accB[t][mt][0], accB[t][mt][1], a, fb[t]);
        }
      }
      if (ks < KC) {
#pragma unroll
        for (int mt = 0; mt < MC; ++mt) {
          const double a = aC[(mt * KC + ks) * 32];
#pragma unroll
          for (int t = 0; t < NT; ++t) dmma884(accC[t][mt][0], accC[t][mt][1], a, fc[t]);
        }
      }
      if (ks < KO) {
#pragma unroll
        for (int mt = 0; mt < MO; ++mt) {
          const double a = aO[(mt * KO + ks) * 32];
#pragma unroll
          for (int t = 0; t < NT; ++t) dmma884(accO[t][mt][0], accO[t][mt][1], a, fo[t]);
        }
      }
    }
    // Finite check (interp.jl:40) on the way out: c_0 = sum_j w_j x_j / 2N has only non-zero weights, so it is
    // non-finite whenever an input of the line is.  Rare path: find the first offending element of the line
    // (the slot still holds the inputs).
    if (prm.bad && q == 0) {
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (dlive[t][h] && !(fabs(accB[t][0][h]) <= 1.79769313486231570815e308)) {
            const int lb0 = dbase[t][h];
            for (int j = 0; j <= N; ++j)
              if (!(fabs((double)x[lb0 + j * E]) <= 1.79769313486231570815e308)) {
                atomicMin(prm.bad, prm.idx_base + u * (long long)(prm.unitBytes / (int)sizeof(R)) + lb0 + j * E);
                break;
              }
          }
    }
    __syncwarp();  // every lane has finished reading the slot: transform in place
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!dlive[t][h]) continue;
        R* y = x + dbase[t][h];
#pragma unroll
        for (int mt = 0; mt < MB; ++mt)
          if (mt < MB - 1 || lastB) y[(kB0 + 8 * SB * mt) * E] = (R)accB[t][mt][h];
#pragma unroll
        for (int mt = 0; mt < MC; ++mt)
          if (mt < MC - 1 || lastC) y[(kC0 + 32 * mt) * E] = (R)accC[t][mt][h];
#pragma unroll
        for (int mt = 0; mt < MO; ++mt)
          if (mt < MO - 1 || lastO) y[(kO0 + 16 * mt) * E] = (R)accO[t][mt][h];
      }
    if (prm.linemax) {
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double m = 0.0;
          if (dlive[t][h]) {
#pragma unroll
            for (int mt = 0; mt < MB; ++mt)
              if (mt < MB - 1 || lastB) m = fmax(m, fabs(accB[t][mt][h]));
#pragma unroll
            for (int mt = 0; mt < MC; ++mt)
              if (mt < MC - 1 || lastC) m = fmax(m, fabs(accC[t][mt][h]));
#pragma unroll
            for (int mt = 0; mt < MO; ++mt)
              if (mt < MO - 1 || lastO) m = fmax(m, fabs(accO[t][mt][h]));
          }
          m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 4));
          m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 8));
          m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 16));
          if (q == 0) prm.linemax[u * (8 * NT) + line_of(t, 2 * c + h)] = m;
        }
    }
    fence_proxy_async();  // generic-proxy writes of the slot -> visible to the bulk store
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(dst + u * (long long)prm.unitBytes, x, (uint32_t)prm.unitBytes);
      bulk_commit();
      // refill the slot whose store was issued one iteration ago with the unit S-1 ahead
      const long long nxt = i + S - 1;
      if (nxt < my_units) {
        bulk_wait_read<1>();
        issue(nxt);
      }
    }
  }
  if (lane == 0) bulk_wait_all();
}

}  // namespace fastcheb
