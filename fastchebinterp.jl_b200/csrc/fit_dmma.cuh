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
// (k = 0 mod 4 from a_j + a_{N/2-j}, k = 2 mod 4 from a_j - a_{N/2-j}), and the odd half splits by the parity
// of j: cos(pi (N-k) j/N) = (-1)^j cos(pi k j/N), so with P_l = sum_i Mp[l][i] b_{2i}, Q_l = sum_i Mq[l][i] b_{2i+1}
// (l < N/4)  C[2l+1] = P_l + Q_l and C[N-2l-1] = P_l - Q_l — both in the same lane of the MMA D fragment, so the
// recombination costs two DADDs and no shuffle: 39 DMMAs per 8 lines at n = 65 (one fold: 77, two: 55).  The matrices (built on the host from exactly reduced long-double cosines, normalisation folded in)
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

constexpr int kFitHeaderBytes = 2048;  // mbarriers [<= 128] | unit counter | per-slot unit numbers [<= 128 x 4 B]

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
  int lag;  // a slot is refilled `lag` (1 or 2) iterations after its store was issued: loads run nslots - lag units ahead
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
//   parts P, Q (KC k-steps each, two folds and SPLIT only, instead of part O): k = 2l+1 and N-2l-1 from b_{2i}, b_{2i+1}
// SPLIT is used where the kernel is bound by the tensor pipe and shared-memory wavefronts (Float32: 57 -> 68 % of HBM
// bandwidth at n = 65); the Float64 kernel is HBM-bound without it (96 %) and measured slower with it (82-91 %
// depending on the launch shape, although it then issues 29 % fewer DMMAs), so it keeps part O.
// K* are compile-time, so a unit is one straight-line block with immediate shared-memory offsets.
// E1: scalar elements (E == 1).
//
// MAP: which folded input sits in contraction slot (k-step ks, lane%4 = c).  Any bijection works as long as the
// matrices are laid out with the same one (fit_map_* below, shared with the host); what changes is the shared-memory
// bank pattern of the operand loads.  Lines are n reals apart, n = 65 = 1 (mod 16 and mod 32), so with the plain map
// (slot 4 ks + c) the 8 lines x 4 slots of a warp-wide load overlap 2- to 4-fold.  MAP 1 / 2 are for n = 65
// (KO, KB, KC = 8, 5, 4) with line q in MMA column q:
//   parts B, C:  j = 4c + ks%4 + 16 (ks/4)            Float64: 16 lanes -> 16 distinct 8-byte banks; Float32: 2-way
//   parts P, Q:  MAP 1 (Float32)  i = ks + 4c          2i = 2ks + 8c: 32 lanes -> 32 distinct 4-byte banks
//                MAP 2 (Float64)  i = 2c + ks%2 + 8 (ks/2)   2i = 4c + ...: 16 lanes -> 16 distinct 8-byte banks
__host__ __device__ constexpr int fit_map_j(int map, int ks, int c) {
  return map == 0 ? 4 * ks + c : 4 * c + (ks & 3) + 16 * (ks >> 2);
}
__host__ __device__ constexpr int fit_map_i(int map, int ks, int c) {
  return map == 0 ? 4 * ks + c : (map == 1 ? ks + 4 * c : 2 * c + (ks & 1) + 8 * (ks >> 1));
}
template <typename R, int KO, int KB, int KC, int NT, bool E1, int MAP = 0, bool SPLIT = false, bool DYN = false>
__global__ void __launch_bounds__(NT == 1 ? 768 : 256) fit_dmma_kernel(const __grid_constant__ FitDmmaParams prm) {
  static_assert(MAP == 0 || (KO == 8 && KB == 5 && KC == 4 && NT == 1 && SPLIT), "tuned slot maps are built for n = 65");
  static_assert(!DYN || (KO == 8 && KB == 5 && KC == 4 && NT == 1), "dynamic unit assignment is instantiated for n = 65");
  constexpr bool L2 = KC > 0;
  constexpr bool PQ = L2 && SPLIT;  // the odd half as parts P, Q instead of part O
  constexpr int MO = (KO + 1) / 2, MB = (KB + 1) / 2, MC = (KC + 1) / 2;  // 8-row m-tiles: ceil(4 K / 8)
  constexpr int KMAX = KO > KB ? KO : KB;
  constexpr int SB = L2 ? 4 : 2, OB = 0;  // output stride / offset of part B; part C: stride 4, offset 2
  extern __shared__ __align__(128) unsigned char fastcheb_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int S = prm.nslots;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fastcheb_smem);  // [nwarps][S], <= 128 entries
  // DYN: the CTA owns a contiguous range of units and its warps draw them from a shared-memory counter (small batches:
  // with 5-6 units per warp the static round-robin split leaves most warps idle during the last round)
  unsigned int* counter = reinterpret_cast<unsigned int*>(fastcheb_smem + 1024);
  unsigned int* uq = reinterpret_cast<unsigned int*>(fastcheb_smem + 1088) + warp * S;  // units in this warp's slots
  constexpr int afragDoubles = ((PQ ? 2 * MC * KC : MO * KO) + MB * KB + MC * KC) * 32;
  double* afrag = reinterpret_cast<double*>(fastcheb_smem + kFitHeaderBytes);
  unsigned char* ring =
      fastcheb_smem + kFitHeaderBytes + ((afragDoubles * 8 + 127) / 128) * 128 + (size_t)warp * S * prm.slotBytes;

  // Back-to-back fits (a 10^5-fit launch is ~25 us, ~9 of them fixed cost): let the next launch in the stream start its
  // CTAs and this set-up on SMs that are already done with the current one.  The matrices are constants written at
  // plan time; everything a predecessor may touch (src, dst, linemax, bad) comes after grid_dep_wait().
  grid_dep_launch_dependents();
  for (int i = threadIdx.x; i < afragDoubles; i += blockDim.x) afrag[i] = prm.afrag[i];
  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[warp * S + s], 1);
    fence_mbar_init();
  }
  if (DYN && threadIdx.x == 0) *counter = 0;
  __syncthreads();  // the only CTA-wide barrier
  grid_dep_wait();

  const long long gw = (long long)blockIdx.x * nwarps + warp;
  const long long tw = (long long)gridDim.x * nwarps;
  if (!DYN && gw >= prm.nunits) return;
  const long long my_units = DYN ? 0 : (prm.nunits - gw + tw - 1) / tw;
  const long long ubeg = prm.nunits * blockIdx.x / gridDim.x, uend = prm.nunits * (blockIdx.x + 1) / gridDim.x;
  const char* src = reinterpret_cast<const char*>(prm.src);
  char* dst = reinterpret_cast<char*>(prm.dst);
  // lane 0 only: put the load of this warp's i-th unit in flight; false when there is no i-th unit
  auto issue = [&](long long i) -> bool {
    const int s = (int)(i % S);
    long long u;
    if constexpr (DYN) {
      const unsigned d = atomicAdd(counter, 1u);
      u = ubeg + d;
      if (u >= uend) return false;
      uq[s] = d;
    } else {
      if (i >= my_units) return false;
      u = gw + i * tw;
    }
    mbar_arrive_expect_tx(&bars[warp * S + s], (uint32_t)prm.unitBytes);
    bulk_g2s(ring + (size_t)s * prm.slotBytes, src + u * (long long)prm.unitBytes, (uint32_t)prm.unitBytes,
             &bars[warp * S + s]);
    return true;
  };
  long long issued = 0;  // units of this warp whose load has been issued (warp-uniform)
  if (lane == 0)
    while (issued < S - prm.lag && issue(issued)) ++issued;
  if constexpr (DYN) issued = __shfl_sync(0xffffffffu, issued, 0);

  const int n = prm.n, N = n - 1, H = N / 2, E = E1 ? 1 : prm.E;
  const int c = lane & 3, q = lane >> 2;
  // B-operand line of this lane in every line group, D-operand lines (columns 2c, 2c+1).  Dead lines (a unit
  // may hold fewer than 8*NT lines when E does not divide it) alias line 0: computed, never stored.
  // tuned maps: columns 0..3 (and 4..7) on lines distinct mod 4 for the operand loads, and columns {0,2,4,6} /
  // {1,3,5,7} on lines {0,1,4,5} / {2,3,6,7} so that the strided result stores are at most 2-way conflicted
  auto line_of = [&](int t, int col) {
    return MAP != 0 ? ((col & 4) | ((col & 1) << 1) | ((col >> 1) & 1)) : 2 * NT * (col & 3) + (col >> 2) + 2 * t;
  };
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
  const double* aO = afrag + lane;  // one fold: O | B;  two folds: P | Q | B | C
  const double* aQ = aO + MC * KC * 32;
  const double* aB = PQ ? aQ + MC * KC * 32 : aO + MO * KO * 32;
  const double* aC = aB + MB * KB * 32;
  // output element of row q in m-tile 0 of each part, and validity of the last m-tile's rows (earlier ones are full)
  const int kO0 = 2 * q + 1, kB0 = SB * q + OB, kC0 = 4 * q + 2;
  const bool lastO = 2 * (8 * (MO - 1) + q) + 1 <= N, lastB = SB * (8 * (MB - 1) + q) + OB <= N;
  const bool lastC = L2 && 4 * (8 * (MC - 1) + q) + 2 <= N;
  const bool lastP = PQ && 4 * (8 * (MC - 1) + q) < N;  // row l = 8 mt + q of P, Q exists: l < N/4

  for (long long i = 0; DYN ? i < issued : i < my_units; ++i) {
    const int s = (int)(i % S);
    const uint32_t parity = (uint32_t)((i / S) & 1);
    if constexpr (DYN) __syncwarp();  // lane 0's uq[s] is visible
    const long long u = DYN ? ubeg + uq[s] : gw + i * tw;
    R* x = reinterpret_cast<R*>(ring + (size_t)s * prm.slotBytes);
    mbar_wait(&bars[warp * S + s], parity);

    // one fold: accO = part O;  two folds: accO[..][0..MC) = part P, accQ = part Q
    double accO[NT][PQ ? MC : MO][2], accQ[NT][PQ ? MC : 1][2], accB[NT][MB][2], accC[NT][L2 ? MC : 1][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
      for (int mt = 0; mt < (PQ ? MC : MO); ++mt) accO[t][mt][0] = accO[t][mt][1] = 0.0;
#pragma unroll
      for (int mt = 0; mt < (PQ ? MC : 1); ++mt) accQ[t][mt][0] = accQ[t][mt][1] = 0.0;
#pragma unroll
      for (int mt = 0; mt < MB; ++mt) accB[t][mt][0] = accB[t][mt][1] = 0.0;
#pragma unroll
      for (int mt = 0; mt < MC; ++mt) accC[t][mt][0] = accC[t][mt][1] = 0.0;
    }
    // fold operands of slot (ks, c): j = fit_map_j = jc + (compile-time part), x_j and x_{N-j} (and x_{H-j}, x_{H+j}
    // for the second fold).  Beyond j = N/2 (possible only in the last k-step of the plain map) the matrix columns are
    // zero; the operand only has to be finite, so those lanes read element 0.
    const int jc = fit_map_j(MAP, 0, c);
#pragma unroll
    for (int ks = 0; ks < (PQ ? KB : KMAX); ++ks) {
      const int j = jc + (fit_map_j(MAP, ks, 0) - fit_map_j(MAP, 0, 0));  // the maps are separable in (ks, c)
      const bool inr = MAP != 0 || j <= H;
      double fo[NT], fb[NT], fc[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const R* z = x + bbase[t];
        const double x1 = (double)*(inr ? z + j * E : z), x2 = (double)*(inr ? z + (N - j) * E : z);
        fo[t] = __dsub_rn(x1, x2);
        fb[t] = __dadd_rn(x1, x2);  // a self-paired element (2j == N) counts twice: its matrix entry carries weight 1
        fc[t] = 0.0;
        if (L2 && ks < KB) {
          const double x3 = (double)*(inr ? z + (H - j) * E : z), x4 = (double)*(inr ? z + (H + j) * E : z);
          const double a2 = __dadd_rn(x3, x4);
          fc[t] = __dsub_rn(fb[t], a2);
          fb[t] = __dadd_rn(fb[t], a2);
        }
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
      if (!PQ && ks < KO) {
#pragma unroll
        for (int mt = 0; mt < MO; ++mt) {
          const double a = aO[(mt * KO + ks) * 32];
#pragma unroll
          for (int t = 0; t < NT; ++t) dmma884(accO[t][mt][0], accO[t][mt][1], a, fo[t]);
        }
      }
    }
    if constexpr (PQ) {  // parts P, Q: b_{2i} and b_{2i+1}, i = fit_map_i < N/4 (beyond: zero matrix columns, read element 0)
      const int ic = fit_map_i(MAP, 0, c);
#pragma unroll
      for (int ks = 0; ks < KC; ++ks) {
        const int i = ic + (fit_map_i(MAP, ks, 0) - fit_map_i(MAP, 0, 0));
        const bool inr = MAP != 0 || 4 * i < N;
        double fp[NT], fq[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const R* z = x + bbase[t];
          const R* pe1 = z + 2 * i * E;
          const R* pe2 = z + (N - 2 * i) * E;
          const double e1 = (double)*(inr ? pe1 : z), e2 = (double)*(inr ? pe2 : z);
          const double o1 = (double)*(inr ? pe1 + E : z), o2 = (double)*(inr ? pe2 - E : z);
          fp[t] = __dsub_rn(e1, e2);
          fq[t] = __dsub_rn(o1, o2);
        }
#pragma unroll
        for (int mt = 0; mt < MC; ++mt) {
          const double ap = aO[(mt * KC + ks) * 32], aq = aQ[(mt * KC + ks) * 32];
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            dmma884(accO[t][mt][0], accO[t][mt][1], ap, fp[t]);
            dmma884(accQ[t][mt][0], accQ[t][mt][1], aq, fq[t]);
          }
        }
      }
      // C[2l+1] = P_l + Q_l, C[N-2l-1] = P_l - Q_l: recombine in place (accO <- k = 2l+1, accQ <- k = N-2l-1)
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int mt = 0; mt < MC; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double P = accO[t][mt][h], Q = accQ[t][mt][h];
            accO[t][mt][h] = __dadd_rn(P, Q);
            accQ[t][mt][h] = __dsub_rn(P, Q);
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
        if constexpr (PQ) {
#pragma unroll
          for (int mt = 0; mt < MC; ++mt)
            if (mt < MC - 1 || lastP) {
              y[(kO0 + 16 * mt) * E] = (R)accO[t][mt][h];
              y[(N - kO0 - 16 * mt) * E] = (R)accQ[t][mt][h];
            }
        } else {
#pragma unroll
          for (int mt = 0; mt < MO; ++mt)
            if (mt < MO - 1 || lastO) y[(kO0 + 16 * mt) * E] = (R)accO[t][mt][h];
        }
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
            if constexpr (PQ) {
#pragma unroll
              for (int mt = 0; mt < MC; ++mt)
                if (mt < MC - 1 || lastP) m = fmax(m, fmax(fabs(accO[t][mt][h]), fabs(accQ[t][mt][h])));
            } else {
#pragma unroll
              for (int mt = 0; mt < MO; ++mt)
                if (mt < MO - 1 || lastO) m = fmax(m, fabs(accO[t][mt][h]));
            }
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
      // refill the slot whose store was issued `lag` iterations ago (its source must have been read) with the unit
      // S - lag ahead.  lag = 2 costs a slot of look-ahead but the wait then almost never blocks: with lag = 1 a warp
      // that computes faster than its previous store drains stalls here instead of keeping loads in flight.
      const long long nxt = i + S - prm.lag;
      if (DYN ? nxt == issued : nxt < my_units) {  // DYN: the queue has never run dry
        if (prm.lag == 1)
          bulk_wait_read<1>();
        else
          bulk_wait_read<2>();
        if (issue(nxt)) ++issued;
      }
    }
    if constexpr (DYN) issued = __shfl_sync(0xffffffffu, issued, 0);
  }
  if (lane == 0) bulk_wait_all();
}

}  // namespace fastcheb
