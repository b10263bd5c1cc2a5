// N-d DCT-I fit with the WHOLE ARRAY RESIDENT IN SHARED MEMORY: one HBM round trip for all dimensions.
//
// Replaces chebcoefs for multidimensional arrays and for longer 1-D lines
// (/root/reference/src/interp.jl:39-57: FFTW.r2r(vals, REDFT00) over every dimension :43-44, the 1/prod 2(n-1)
// scale :49, the interior doubling :50-54) — BASELINE.json configs[1] (65 x 65) and the 33^3 / 17^4 arrays of
// configs[2..3], single fits (`chebinterp(vals, lb, ub)`) and batches of them alike.
//
// A CTA (several per SM for small arrays) owns one array at a time: it loads the array into shared memory, transforms
// it in place along dimension 1, 2, ... (a CTA barrier between dimensions), and writes the coefficients back.  A pass
// along dimension d treats the array as lines of n_d elements with element stride s_d = E n_1 .. n_{d-1}; 8 lines
// form the N extent of an FP64 tensor-core MMA (mma.sync m8n8k4, SASS DMMA), exactly as in the batched 1-D kernel
// (fit_dmma.cuh): REDFT00's symmetry splits the transform into odd outputs from b_j = x_j - x_{N-j} (part O) and
// even outputs from a_j = x_j + x_{N-j}, folded once more when 4 | N (parts B: k = 4l, C: k = 4l+2); the folded,
// normalised cosine matrices of every distinct n are A fragments in shared memory.  A warp takes 8 lines, forms the
// folds, runs the three small dense products and stores the outputs over the inputs (all reads of a line group
// happen before its writes; different warps own different lines).
//
// The k-step and m-tile counts are RUNTIME values here (one instantiation serves every n up to 8 MT_O + 1), so a
// 65 x 33 x 17 array needs no kernel of its own; the accumulators are sized for the largest case.
//
// Shared-memory banks: for n = 2^m + 1 every stride s_d (and the distance between neighbouring lines) is 1 modulo the
// number of banks, so element j of line L sits in bank (L + j).  The 8 lines of an MMA are therefore taken 4 apart
// from a block of 16 (Float64: column q <-> line 4 (q%4) + q/4 + 2t, t = 0, 1) or 32 lines (Float32: column q <-> line
// 4 q + t, t = 0..3): with contraction slot (k-step ks, lane%4 = c) <-> folded input j = 4 ks + c, the 16 (32) lanes of
// a shared-memory wavefront read banks 4 q' + c — all distinct — in every pass.
#pragma once
#include <cstdint>
#include "ptx_util.cuh"

namespace fastcheb {

constexpr int kFitNdMaxDims = 6;

struct FitNdDim {
  int n;        // size of this dimension (1: skipped)
  int stride;   // element stride of a line, in reals
  int KO, KB, KC, two;  // k-steps of parts O, B, C; two = even half folded twice
  int afrag;    // offset (doubles) of this dimension's A fragments in the shared-memory matrix area
};

struct FitNdParams {
  const void* src;
  void* dst;
  const double* afrag;  // all distinct matrices back to back
  int afragDoubles;
  long long howmany;
  long long* bad;       // first non-finite flat real index (atomicMin), may be null
  long long idx_base;
  double* maxabs;       // optional: max_k infnorm(C_k) per array
  int ndim, E, K, cplx;
  int reals;            // reals per array = E prod n (x lines per block for batches of 1-D lines)
  int fits, fit_reals;  // independent fits in one array (1-D line blocks: > 1) and reals per fit
  int nbuf;             // array buffers in shared memory (2: the next array loads while this one is transformed)
  int teams, team_warps;  // a CTA = teams x team_warps warps; a team owns one array at a time
  double* stats;          // optional, ONE array only: droptol's statistics (interp.jl:77-91) — [0] max, [1] min of the
  int hoff[kFitNdMaxDims];  // element norms, [2 + hoff[d] + k] != 0 iff the hyperplane k_d = k has an element >= tol * max
  int hsum;               // sum of the dimensions
  double stats_tol;
  FitNdDim dim[kFitNdMaxDims];
};

// One pass over 8 lines.  x: the array in shared memory; lane (c = lane%4, q = lane/4) feeds line `bl` (B operand,
// column q) and receives lines dl[0], dl[1] (D operand, columns 2c, 2c+1; dlive: the line exists).
// CKO / CKB / CKC != 0: the k-step counts are compile-time (the common sizes n = 17, 33, 65, 129): the k loop is then one
// straight-line block, every operand load of a k-step is issued before its DMMAs and the loads of the next k-step
// overlap them.  0: runtime counts from `d` (any n), with the loads of a k-step hoisted by hand.
// AREG: the A fragments come from REGISTERS (`areg`, loaded once per kernel: all dimensions of the array have this size),
// not from shared memory — the matrix loads were more than half of the shared-memory wavefronts of a line group, and
// the shared-memory pipe, not the tensor pipe, bounded the kernel (ncu: 5100 wavefronts per 65 x 65 array against 3100
// tensor-pipe cycles).
template <typename R, int MTO, int MTB, int MTC, int CKO, int CKB, int CKC, bool AREG = false>
__device__ __forceinline__ void fit_nd_lines(R* x, const FitNdDim& d, const double* __restrict__ af, int bl, const int (&dl)[2],
                                             const bool (&dlive)[2], int lane, const double* areg = nullptr) {
  constexpr bool CT = CKO != 0;
  static_assert(!AREG || (CT && CKC > 0), "register-resident matrices: compile-time, twice-folded sizes only");
  const int c = lane & 3, q = lane >> 2;
  const int N = d.n - 1, H = N / 2, s = d.stride;
  const int KO = CT ? CKO : d.KO, KB = CT ? CKB : d.KB, KC = CT ? CKC : d.KC;
  const int MO = (KO + 1) / 2, MB = (KB + 1) / 2, MC = (KC + 1) / 2;
  const bool two = CT ? (CKC > 0) : (d.two != 0);
  const int Jo = (N + 1) / 2;                  // inputs of part O (once-folded): j < Jo
  const int Jb = two ? H / 2 + 1 : N / 2 + 1;  // inputs of part B
  // once-folded: accO = part O (MO tiles).  twice-folded: the odd half is split by the parity of the input index,
  //   P_l = sum_i Mp[l][i] b_{2i},  Q_l = sum_i Mq[l][i] b_{2i+1}  (l, i < N/4),   C[2l+1] = P_l + Q_l,  C[N-2l-1] = P_l - Q_l
  // (cos(pi (N-k) j / N) = (-1)^j cos(pi k j / N)): accO[0..MC) = P, accQ = Q — 39 instead of 55 DMMAs per 8 lines at n = 65.
  constexpr int MTP = MTO > MTC ? MTO : MTC;  // accO holds part O (MTO tiles) or part P (MTC tiles)
  double accO[MTP][2], accQ[MTC][2], accB[MTB][2], accC[MTC][2];
#pragma unroll
  for (int m = 0; m < MTP; ++m) accO[m][0] = accO[m][1] = 0.0;
#pragma unroll
  for (int m = 0; m < MTB; ++m) accB[m][0] = accB[m][1] = 0.0;
#pragma unroll
  for (int m = 0; m < MTC; ++m) accC[m][0] = accC[m][1] = accQ[m][0] = accQ[m][1] = 0.0;
  const double* aO = af + lane;  // once-folded: O | B;  twice-folded: P | Q | B | C
  const double* aQ = aO + MC * KC * 32;
  const double* aB = two ? aQ + MC * KC * 32 : aO + MO * KO * 32;
  const double* aC = aB + MB * KB * 32;
  const R* z = x + bl;
  const int KMAX = two ? KB : (KO > KB ? KO : KB);  // twice-folded: KC <= KB
  auto step = [&](int ks) {
    const int j = 4 * ks + c;
    // beyond the folded range the matrix columns are zero: the operand only has to be finite, read element 0
    const int jj = ((!two && j < Jo) || j < Jb) ? j : 0;
    const R r1 = z[jj * s], r2 = z[(N - jj) * s];
    R r3 = R(0), r4 = R(0), p1 = R(0), p2 = R(0), q1 = R(0), q2 = R(0);
    if (two) {
      r3 = z[(H - jj) * s];
      r4 = z[(H + jj) * s];
      const int i2 = (ks < KC && 4 * j < N) ? 2 * j : 0;  // parts P, Q: slot i = j carries b_{2i}, b_{2i+1}
      p1 = z[i2 * s];
      p2 = z[(N - i2) * s];
      q1 = z[(i2 + 1) * s];
      q2 = z[(N - i2 - 1) * s];
    }
    // the A fragments of this k-step, all loaded before the first DMMA (the first entry of the matrices where the part
    // has no such tile; those values are never used)
    double vO[MTP], vQ[MTC], vB[MTB], vC[MTC];
    if constexpr (AREG) {
      // register layout = the matrices' order: P | Q | B | C, each [m][ks]
      constexpr int oQ = ((CKC + 1) / 2) * CKC, oB = 2 * oQ, oC = oB + ((CKB + 1) / 2) * CKB;
#pragma unroll
      for (int m = 0; m < MTC; ++m) {
        vO[m] = (m < MC && ks < KC) ? areg[m * CKC + ks] : 0.0;
        vQ[m] = (m < MC && ks < KC) ? areg[oQ + m * CKC + ks] : 0.0;
        vC[m] = (m < MC && ks < KC) ? areg[oC + m * CKC + ks] : 0.0;
      }
#pragma unroll
      for (int m = 0; m < MTB; ++m) vB[m] = (m < MB && ks < KB) ? areg[oB + m * CKB + ks] : 0.0;
    } else if (two) {
#pragma unroll
      for (int m = 0; m < MTC; ++m) {
        vO[m] = *((m < MC && ks < KC) ? aO + (m * KC + ks) * 32 : aO);
        vQ[m] = *((m < MC && ks < KC) ? aQ + (m * KC + ks) * 32 : aO);
        vC[m] = *((m < MC && ks < KC) ? aC + (m * KC + ks) * 32 : aO);
      }
    } else {
#pragma unroll
      for (int m = 0; m < MTO; ++m) vO[m] = *((m < MO && ks < KO) ? aO + (m * KO + ks) * 32 : aO);
    }
    if constexpr (!AREG) {
#pragma unroll
      for (int m = 0; m < MTB; ++m) vB[m] = *((m < MB && ks < KB) ? aB + (m * KB + ks) * 32 : aO);
    }
    const double x1 = (double)r1, x2 = (double)r2;
    double fb = __dadd_rn(x1, x2);
    if (two) {
      const double a2 = __dadd_rn((double)r3, (double)r4);
      const double fc = __dsub_rn(fb, a2);
      fb = __dadd_rn(fb, a2);
      if (ks < KC) {
        const double fp = __dsub_rn((double)p1, (double)p2), fq = __dsub_rn((double)q1, (double)q2);
#pragma unroll
        for (int m = 0; m < MTC; ++m)
          if (m < MC) {
            dmma884(accO[m][0], accO[m][1], vO[m], fp);
            dmma884(accQ[m][0], accQ[m][1], vQ[m], fq);
            dmma884(accC[m][0], accC[m][1], vC[m], fc);
          }
      }
    } else if (ks < KO) {
      const double fo = __dsub_rn(x1, x2);
#pragma unroll
      for (int m = 0; m < MTO; ++m)
        if (m < MO) dmma884(accO[m][0], accO[m][1], vO[m], fo);
    }
    if (ks < KB) {
#pragma unroll
      for (int m = 0; m < MTB; ++m)
        if (m < MB) dmma884(accB[m][0], accB[m][1], vB[m], fb);
    }
  };
  if constexpr (CT) {
#pragma unroll
    for (int ks = 0; ks < (CKC > 0 ? CKB : (CKO > CKB ? CKO : CKB)); ++ks) step(ks);
  } else {
#pragma unroll 2
    for (int ks = 0; ks < KMAX; ++ks) step(ks);
  }
  __syncwarp();  // every lane has read its inputs: transform in place
  const int SB = two ? 4 : 2;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (!dlive[h]) continue;
    R* y = x + dl[h];
    if (two) {
#pragma unroll
      for (int m = 0; m < MTC; ++m) {
        const int l = 8 * m + q;
        if (m < MC && 4 * l < N) {
          const double P = accO[m][h], Q = accQ[m][h];
          y[(2 * l + 1) * s] = (R)__dadd_rn(P, Q);
          y[(N - 2 * l - 1) * s] = (R)__dsub_rn(P, Q);
        }
        const int k = 4 * l + 2;
        if (m < MC && k <= N) y[k * s] = (R)accC[m][h];
      }
    } else {
#pragma unroll
      for (int m = 0; m < MTO; ++m) {
        const int k = 2 * (8 * m + q) + 1;
        if (m < MO && k <= N) y[k * s] = (R)accO[m][h];
      }
    }
#pragma unroll
    for (int m = 0; m < MTB; ++m) {
      const int k = SB * (8 * m + q);
      if (m < MB && k <= N) y[k * s] = (R)accB[m][h];
    }
  }
}

// One pass along a dimension: the team's warps take line groups round-robin.
// number of A-fragment values per lane of the register-resident form (parts P | Q | B | C)
__host__ __device__ constexpr int fit_nd_areg_count(int spec) { return spec == 65 ? 39 : (spec == 33 ? 12 : (spec == 17 ? 5 : 1)); }

// SPEC != 0: every transformed dimension has size SPEC (17, 33 or 65) and its matrices are in `areg`
template <typename R, int MTO, int MTB, int MTC, int SPEC = 0>
__device__ __forceinline__ void fit_nd_pass(R* x, const FitNdDim& d, const double* afd, int reals, int warp, int nwarps, int lane,
                                            const double* areg = nullptr, int nlines_live = -1) {
  const int s = d.stride;
  const int nlines = nlines_live >= 0 ? nlines_live : reals / d.n;
  constexpr int GB = sizeof(R) == 8 ? 2 : 4;  // groups per block of 8 GB lines
  // groups of the last, partial block that contain no line at all are skipped (the groups of a block start at lines
  // blk + 2 t (Float64) / blk + t (Float32), so the live ones are its first): 65 lines are 9 groups, not 10 (Float64) or 12
  // (Float32) — the pass phase is bound by shared-memory wavefronts, and a dead group costs as many as a live one
  const int rem = nlines % (8 * GB);
  const int ngroups = nlines / (8 * GB) * GB + (GB == 2 ? min(GB, (rem + 1) / 2) : min(GB, rem));
  for (int gi = warp; gi < ngroups; gi += nwarps) {
    // line L -> (outer o, inner i): base = o * n * s + i
    auto base_of = [&](int L) {
      const int o = L / s, i = L - o * s;
      return o * d.n * s + i;
    };
    const int c = lane & 3, q = lane >> 2;
    const int blk = (gi / GB) * 8 * GB, t = gi % GB;
    auto line_of = [&](int col) { return blk + (GB == 2 ? 4 * (col & 3) + (col >> 2) + 2 * t : 4 * col + t); };
    const int Lb = line_of(q);
    const int bl = base_of(Lb < nlines ? Lb : 0);
    int dl[2];
    bool dlive[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int Ld = line_of(2 * c + h);
      dlive[h] = Ld < nlines;
      dl[h] = base_of(dlive[h] ? Ld : 0);
    }
    if constexpr (SPEC == 65) {
      fit_nd_lines<R, 4, 3, 2, 8, 5, 4, true>(x, d, afd, bl, dl, dlive, lane, areg);
      continue;
    } else if constexpr (SPEC == 33) {
      fit_nd_lines<R, 2, 2, 1, 4, 3, 2, true>(x, d, afd, bl, dl, dlive, lane, areg);
      continue;
    } else if constexpr (SPEC == 17) {
      fit_nd_lines<R, 1, 1, 1, 2, 2, 1, true>(x, d, afd, bl, dl, dlive, lane, areg);
      continue;
    }
    // the common sizes run a fully unrolled copy of the pass
    if (d.KO == 8 && d.KB == 5 && d.KC == 4)                    // n = 65
      fit_nd_lines<R, 4, 3, 2, 8, 5, 4>(x, d, afd, bl, dl, dlive, lane);
    else if (d.KO == 4 && d.KB == 3 && d.KC == 2)               // n = 33
      fit_nd_lines<R, 2, 2, 1, 4, 3, 2>(x, d, afd, bl, dl, dlive, lane);
    else if (d.KO == 2 && d.KB == 2 && d.KC == 1)               // n = 17
      fit_nd_lines<R, 1, 1, 1, 2, 2, 1>(x, d, afd, bl, dl, dlive, lane);
    else if (MTB >= 5 && d.KO == 16 && d.KB == 9 && d.KC == 8)  // n = 129
      fit_nd_lines<R, 4, 5, 4, 16, 9, 8>(x, d, afd, bl, dl, dlive, lane);
    else
      fit_nd_lines<R, MTO, MTB, MTC, 0, 0, 0>(x, d, afd, bl, dl, dlive, lane);
  }
}

// Asynchronous copy of one array into shared memory (cp.async, SASS LDGSTS): the destination is shifted so that it has
// the source's alignment modulo 16 bytes, the body then moves in 16-byte pieces and only the ragged head / tail element
// by element — arrays of an odd number of reals (65 x 65 doubles = 33 800 B) start at every other 8-byte offset.
// Returns the shifted base.  Every thread commits one group.
template <typename R>
__device__ __forceinline__ R* fit_nd_load_async(unsigned char* buf16, const R* g, int reals, int tid, int nthr) {
  const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(g) & 15u);
  R* x = reinterpret_cast<R*>(buf16 + mis);
  int head = (int)(((16u - mis) & 15u) / sizeof(R));
  if (head > reals) head = reals;
  const int nvec = (int)(((size_t)(reals - head) * sizeof(R)) / 16);
  const int per = 16 / (int)sizeof(R);
  const int tail0 = head + nvec * per;
  for (int i = tid; i < nvec; i += nthr)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(x + head + i * per)), "l"(g + head + i * per) : "memory");
  for (int i = tid; i < head + (reals - tail0); i += nthr) {
    const int e = i < head ? i : tail0 + (i - head);
    if constexpr (sizeof(R) == 8)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(x + e)), "l"(g + e) : "memory");
    else
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(x + e)), "l"(g + e) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  return x;
}

// MTO/MTB/MTC: accumulator m-tiles the instantiation provides for part O (once-folded sizes, n <= 16 MTO + 1), part B and
// parts P / Q / C (twice-folded sizes).
// AFG: the matrices are too large for shared memory next to the array (n > 129: 130-200 KB) and are read from global
// memory instead — they are shared by every CTA and stay in L2.
// Two array buffers when they fit (prm.nbuf == 2): the next array's copy is in flight while the current one is
// transformed.
// TEAMS: a CTA is cut into prm.teams teams of prm.team_warps warps; a team owns one array at a time (its own buffers, its
// own named barrier), the matrices are shared by the whole CTA.  Small arrays (65 x 65: 10 line groups per pass) keep
// only 2 warps busy without idle barrier rounds, so an SM runs 6 such teams side by side in ONE CTA with ONE copy of the
// matrices, instead of 4 CTAs of 4 half-idle warps with a copy each.
// SPEC (17 / 33 / 65): all transformed dimensions have this size: the matrices live in registers (no shared-memory copy).
template <typename R, int MTO, int MTB, int MTC, bool AFG, int SPEC = 0>
__global__ void __launch_bounds__((MTB <= 4 && SPEC != 65) ? 512 : 256) fit_nd_kernel(const __grid_constant__ FitNdParams prm) {
  extern __shared__ __align__(16) unsigned char fastcheb_smem[];
  // [1 KB header: per team 8 partial maxima + flag] [matrices unless AFG] [array buffers, 16-byte aligned, 16 bytes of
  // slack]  (no static shared memory: the kernel is launched with the opt-in maximum of dynamic shared memory)
  const int team_thr = prm.team_warps * 32;
  const int team = threadIdx.x / team_thr;
  const int tid = threadIdx.x - team * team_thr, lane = tid & 31, warp = tid >> 5, nwarps = prm.team_warps;
  double* s_red = reinterpret_cast<double*>(fastcheb_smem) + team * 8;
  int& s_flag = *(reinterpret_cast<int*>(fastcheb_smem + 768) + team);
  double* af_s = reinterpret_cast<double*>(fastcheb_smem + 1024);
  const int reals = prm.reals;
  const size_t bufBytes = (((size_t)reals * sizeof(R) + 15) / 16) * 16 + 16;
  unsigned char* bufs = fastcheb_smem + 1024 + ((AFG || SPEC != 0) ? 0 : (((size_t)prm.afragDoubles * 8 + 15) / 16) * 16) +
                        (size_t)team * prm.nbuf * bufBytes;
  if constexpr (!AFG && SPEC == 0) {
    for (int i = threadIdx.x; i < prm.afragDoubles; i += blockDim.x) af_s[i] = prm.afrag[i];
    __syncthreads();  // the only CTA-wide barrier
  }
  // SPEC: this lane's A-fragment values of the one matrix set, in registers for the whole kernel
  [[maybe_unused]] double areg[fit_nd_areg_count(SPEC)];
  if constexpr (SPEC != 0) {
#pragma unroll
    for (int i = 0; i < fit_nd_areg_count(SPEC); ++i) areg[i] = prm.afrag[i * 32 + (threadIdx.x & 31)];
  }
  // barrier of this team (named barriers 1..15; the team is the whole CTA when prm.teams == 1)
  auto team_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(team_thr) : "memory"); };
  const R* src = reinterpret_cast<const R*>(prm.src);
  R* dst = reinterpret_cast<R*>(prm.dst);
  const bool dbl = prm.nbuf == 2;
  const long long first = (long long)blockIdx.x * prm.teams + team, stride = (long long)gridDim.x * prm.teams;
  R* xnext = nullptr;
  if (first < prm.howmany) xnext = fit_nd_load_async<R>(bufs, src + first * reals, reals, tid, team_thr);
  int it = 0;
  for (long long f = first; f < prm.howmany; f += stride, ++it) {
    const R* g = src + f * (long long)reals;
    R* x = xnext;
    const long long fn = f + stride;
    if (dbl && fn < prm.howmany) {
      // the other buffer was stored in the previous iteration (all threads passed that iteration's last barrier)
      xnext = fit_nd_load_async<R>(bufs + (size_t)((it + 1) & 1) * bufBytes, src + fn * (long long)reals, reals, tid, team_thr);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    team_sync();  // the array is in shared memory
    for (int dd = 0; dd < prm.ndim; ++dd) {
      const FitNdDim& d = prm.dim[dd];
      if (d.n <= 1) continue;  // DHT of length 1 = identity (interp.jl:43)
      const double* afd;  // matrices: shared memory (address space visible to the compiler) or, AFG, global memory / L2
      if constexpr (AFG)
        afd = prm.afrag + d.afrag;
      else
        afd = af_s + d.afrag;
      if constexpr (SPEC != 0)
        fit_nd_pass<R, MTO, MTB, MTC, SPEC>(x, d, nullptr, reals, warp, nwarps, lane, areg);
      else
        fit_nd_pass<R, MTO, MTB, MTC>(x, d, afd, reals, warp, nwarps, lane);
      team_sync();
    }
    // Finite check (interp.jl:40): c_{0..0} carries a non-zero weight of every input, so it is non-finite iff some
    // input is.  Rare path: the CTA scans the (not yet overwritten) input array for the first offending real.
    if (prm.bad) {
      if (tid == 0) s_flag = 0;
      team_sync();
      for (int i = tid; i < prm.fits * prm.E; i += team_thr) {
        const int fi = i / prm.E, e = i - fi * prm.E;
        if (!(fabs((double)x[(size_t)fi * prm.fit_reals + e]) <= 1.79769313486231570815e308)) s_flag = 1;
      }
      team_sync();
      if (s_flag) {
        long long first = 0x7fffffffffffffffLL;
        for (int i = tid; i < reals; i += team_thr)
          if (!(fabs((double)g[i]) <= 1.79769313486231570815e308)) {
            first = i;
            break;
          }
        if (first != 0x7fffffffffffffffLL) atomicMin(prm.bad, prm.idx_base + f * (long long)reals + first);
      }
    }
    if (prm.stats) {
      // droptol's statistics of a single array, while it is still in shared memory (saves the two reduction launches of
      // a chebinterp call, whose three global atomics per element cost ~30 us on a 65 x 65 array):
      //   phase 1: max / min of the element norms (interp.jl:78-79);
      //   phase 2, only when something can be dropped (min < tol * max): a hyperplane k_d = k survives iff one of its
      //   elements is >= tol * max (interp.jl:83-91) — every such element marks its N hyperplanes with a plain store.
      // stats[2 + hoff[d] + k] = max (marked) or 0 (not marked): the host compares them with tol * max as before.
      const int E = prm.E, nel = reals / E;
      double* s_w = reinterpret_cast<double*>(fastcheb_smem);  // header scratch (teams == 1): [0,16) maxima, [16,32) minima
      auto infnorm = [&](const R* ce) {
        double m = 0.0;
        for (int k = 0; k < prm.K; ++k) {
          const double a = prm.cplx ? hypot((double)ce[2 * k], (double)ce[2 * k + 1]) : fabs((double)ce[k]);
          m = a > m ? a : m;
        }
        return m;
      };
      double tm = 0.0, tn = 1.79769313486231570815e308;
      for (int el = tid; el < nel; el += team_thr) {
        const double a = infnorm(x + (size_t)el * E);
        tm = fmax(tm, a);
        tn = fmin(tn, a);
      }
      for (int off = 16; off; off >>= 1) {
        tm = fmax(tm, __shfl_xor_sync(0xffffffffu, tm, off));
        tn = fmin(tn, __shfl_xor_sync(0xffffffffu, tn, off));
      }
      if (lane == 0) {
        s_w[warp] = tm;
        s_w[16 + warp] = tn;
      }
      for (int i = tid; i < prm.hsum; i += team_thr) prm.stats[2 + i] = 0.0;
      team_sync();
      double gm = 0.0, gn = 1.79769313486231570815e308;
      for (int w = 0; w < nwarps; ++w) {
        gm = fmax(gm, s_w[w]);
        gn = fmin(gn, s_w[16 + w]);
      }
      if (tid == 0) {
        prm.stats[0] = gm;
        prm.stats[1] = gn;
      }
      const double abstol = gm * prm.stats_tol;  // interp.jl:78 (the host forms the same product)
      if (!(gn >= abstol)) {                     // interp.jl:79
        for (int el = tid; el < nel; el += team_thr) {
          if (!(infnorm(x + (size_t)el * E) >= abstol)) continue;
          int rem = el;
          for (int dd = 0; dd < prm.ndim; ++dd) {
            const int nd = prm.dim[dd].n, k = rem % nd;
            rem /= nd;
            prm.stats[2 + prm.hoff[dd] + k] = gm;
          }
        }
      }
      team_sync();  // the header scratch is free again (the max|c| reduction below uses it)
    }
    R* o = dst + f * (long long)reals;
    double mx = 0.0;
    if (prm.maxabs && prm.fits == 1) {
      // infnorm of every element (interp.jl:74-75): complex modulus, max over the K components
      const int E = prm.E;
      for (int el = tid; el < reals / E; el += team_thr) {
        const R* ce = x + (size_t)el * E;
        for (int k = 0; k < prm.K; ++k) {
          const double a = prm.cplx ? hypot((double)ce[2 * k], (double)ce[2 * k + 1]) : fabs((double)ce[k]);
          mx = a > mx ? a : mx;
        }
      }
    } else if (prm.maxabs) {
      // a block of 1-D lines: one thread scans one fit
      const int E = prm.E;
      for (int fi = tid; fi < prm.fits; fi += team_thr) {
        const R* cf = x + (size_t)fi * prm.fit_reals;
        double m = 0.0;
        for (int el = 0; el < prm.fit_reals / E; ++el)
          for (int k = 0; k < prm.K; ++k) {
            const R* ce = cf + (size_t)el * E;
            const double a = prm.cplx ? hypot((double)ce[2 * k], (double)ce[2 * k + 1]) : fabs((double)ce[k]);
            m = a > m ? a : m;
          }
        prm.maxabs[f * prm.fits + fi] = m;
      }
    }
    for (int i = tid; i < reals; i += team_thr) o[i] = x[i];
    if (prm.maxabs && prm.fits == 1) {
      for (int off = 16; off; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      if (lane == 0) s_red[warp] = mx;
      team_sync();
      if (tid == 0) {
        double m = 0.0;
        for (int w = 0; w < nwarps; ++w) m = fmax(m, s_red[w]);
        prm.maxabs[f] = m;
      }
    }
    team_sync();  // every thread has read this buffer: it may be refilled
    if (!dbl && fn < prm.howmany) xnext = fit_nd_load_async<R>(bufs, src + fn * (long long)reals, reals, tid, team_thr);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// The same transform as a PIPELINE for small arrays (three buffers + the matrices fit shared memory, <= 14 line groups
// per pass: 65 x 65, 33 x 33, 17^3 ...).  In fit_nd_kernel a CTA's warps load, transform and store in lock-step, and
// its few line groups leave warps idle at every barrier (65 lines = 10 groups: 4 warps run 3 rounds, the last
// half-empty): the tensor pipe was 45 % busy.  Here the warps are specialised:
//   * one LOADER warp keeps cp.async copies of the next arrays in flight (3 buffers; completion on an mbarrier per
//     buffer via cp.async.mbarrier.arrive),
//   * NC CONSUMER warps — as many as a pass has line groups, so a pass is ONE round — transform the array in place; the
//     barrier between dimensions is a named barrier of the consumers only,
//   * one STORER warp writes the coefficients back (fusing the finite check and max|c|) while the consumers are already
//     on the next array.
// The consumers never wait for memory unless HBM itself is the limit.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kFitPipeBufs = 3;

template <typename R, int MTO, int MTB, int MTC>
__global__ void __launch_bounds__(512) fit_nd_pipe_kernel(const __grid_constant__ FitNdParams prm) {
  extern __shared__ __align__(16) unsigned char fastcheb_smem[];
  // [1 KB header: mbarriers full[3] | done[3] | empty[3]] [matrices] [3 array buffers]
  uint64_t* full = reinterpret_cast<uint64_t*>(fastcheb_smem);
  uint64_t* done = full + kFitPipeBufs;
  uint64_t* empty = done + kFitPipeBufs;
  double* af_s = reinterpret_cast<double*>(fastcheb_smem + 1024);
  const int reals = prm.reals;
  const size_t bufBytes = (((size_t)reals * sizeof(R) + 15) / 16) * 16 + 16;
  unsigned char* bufs = fastcheb_smem + 1024 + (((size_t)prm.afragDoubles * 8 + 15) / 16) * 16;
  const int NC = prm.team_warps;  // consumer warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int b = 0; b < kFitPipeBufs; ++b) {
      mbar_init(&full[b], 32);   // the loader's 32 lanes, each after its own copies
      mbar_init(&done[b], NC);   // one arrival per consumer warp
      mbar_init(&empty[b], prm.teams);   // the storer warps (prm.teams of them)
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < prm.afragDoubles; i += blockDim.x) af_s[i] = prm.afrag[i];
  __syncthreads();  // the only CTA-wide barrier
  const R* src = reinterpret_cast<const R*>(prm.src);
  R* dst = reinterpret_cast<R*>(prm.dst);
  const long long first = blockIdx.x, stride = gridDim.x;
  auto xbase = [&](int b, const R* g) {
    return reinterpret_cast<R*>(bufs + (size_t)b * bufBytes + (reinterpret_cast<uintptr_t>(g) & 15u));
  };

  if (warp == NC) {
    // ---- loader
    long long i = 0;
    for (long long f = first; f < prm.howmany; f += stride, ++i) {
      const int b = (int)(i % kFitPipeBufs);
      const long long use = i / kFitPipeBufs;
      if (use > 0) mbar_wait(&empty[b], (uint32_t)((use - 1) & 1));  // the storer has drained the previous array
      const R* g = src + f * (long long)reals;
      R* x = xbase(b, g);
      const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(g) & 15u);
      int head = (int)(((16u - mis) & 15u) / sizeof(R));
      if (head > reals) head = reals;
      const int nvec = (int)(((size_t)(reals - head) * sizeof(R)) / 16);
      constexpr int per = 16 / (int)sizeof(R);
      const int tail0 = head + nvec * per;
      for (int k = lane; k < nvec; k += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(x + head + k * per)), "l"(g + head + k * per) : "memory");
      for (int k = lane; k < head + (reals - tail0); k += 32) {
        const int e = k < head ? k : tail0 + (k - head);
        if constexpr (sizeof(R) == 8)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(x + e)), "l"(g + e) : "memory");
        else
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(x + e)), "l"(g + e) : "memory");
      }
      cp_async_mbar_arrive(&full[b]);  // arrives when this lane's copies have landed
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp > NC && warp <= NC + prm.teams) {
    // ---- storers (NS warps share an array): coefficients out, finite check (interp.jl:40) and max|c| (interp.jl:74-78) on
    // the way.  One warp alone needs ~5000 cycles per 65 x 65 array (132 dependent LDS -> STG round trips per lane) and
    // was the bottleneck of the first version; the loop is batched 4 deep and split over the storer warps.
    const int NS = prm.teams, sw = warp - NC - 1;
    const int slane = sw * 32 + lane, sthr = NS * 32;
    long long i = 0;
    for (long long f = first; f < prm.howmany; f += stride, ++i) {
      const int b = (int)(i % kFitPipeBufs);
      const long long use = i / kFitPipeBufs;
      mbar_wait(&done[b], (uint32_t)(use & 1));
      const R* g = src + f * (long long)reals;
      const R* x = xbase(b, g);
      R* o = dst + f * (long long)reals;
      const int E = prm.E;
      // finite check first: the rare path scans the INPUT array, which (vals == coefs allowed) the copy below overwrites
      if (prm.bad) {
        bool bad = false;
        for (int e = 0; e < E; ++e) bad = bad || !(fabs((double)x[e]) <= 1.79769313486231570815e308);
        if (bad) {
          long long firstbad = 0x7fffffffffffffffLL;
          for (int k = slane; k < reals; k += sthr)
            if (!(fabs((double)g[k]) <= 1.79769313486231570815e308)) {
              firstbad = k;
              break;
            }
          if (firstbad != 0x7fffffffffffffffLL) atomicMin(prm.bad, prm.idx_base + f * (long long)reals + firstbad);
        }
      }
      double mx = 0.0;
      if (E == 1) {
        int k = slane;
        for (; k + 3 * sthr < reals; k += 4 * sthr) {
          const R v0 = x[k], v1 = x[k + sthr], v2 = x[k + 2 * sthr], v3 = x[k + 3 * sthr];
          mx = fmax(fmax(mx, fabs((double)v0)), fmax(fabs((double)v1), fmax(fabs((double)v2), fabs((double)v3))));
          o[k] = v0;
          o[k + sthr] = v1;
          o[k + 2 * sthr] = v2;
          o[k + 3 * sthr] = v3;
        }
        for (; k < reals; k += sthr) {
          const R v = x[k];
          mx = fmax(mx, fabs((double)v));
          o[k] = v;
        }
      } else {
        for (int el = slane; el < reals / E; el += sthr) {
          const R* ce = x + (size_t)el * E;
          for (int k = 0; k < prm.K; ++k) {
            const double a = prm.cplx ? hypot((double)ce[2 * k], (double)ce[2 * k + 1]) : fabs((double)ce[k]);
            mx = a > mx ? a : mx;
          }
          for (int e = 0; e < E; ++e) o[(size_t)el * E + e] = ce[e];
        }
      }
      if (prm.maxabs) {  // maxabs is zero-initialised by the host (fit_device_t): the storer warps combine with an atomic max
        for (int off = 16; off; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(prm.maxabs + f), (unsigned long long)__double_as_longlong(mx));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
    }
  } else if (warp < NC) {
    // ---- consumers
    const int cthr = NC * 32;
    long long i = 0;
    for (long long f = first; f < prm.howmany; f += stride, ++i) {
      const int b = (int)(i % kFitPipeBufs);
      const long long use = i / kFitPipeBufs;
      mbar_wait(&full[b], (uint32_t)(use & 1));
      R* x = xbase(b, src + f * (long long)reals);
      for (int dd = 0; dd < prm.ndim; ++dd) {
        const FitNdDim& d = prm.dim[dd];
        if (d.n <= 1) continue;
        fit_nd_pass<R, MTO, MTB, MTC>(x, d, af_s + d.afrag, reals, warp, NC, lane);
        asm volatile("bar.sync 1, %0;" ::"r"(cthr) : "memory");  // consumers only
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[b]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Arrays that do NOT fit shared memory (33^3 Float64 = 287 KB, 17^4 = 668 KB, 65^3 ...): the leading dimensions are
// transformed slab by slab with the resident kernel (a slab of dims 1..N-1 is contiguous), and the LAST dimension by
// this kernel on strided tiles: `rows` = n_N rows of `ncols` consecutive reals, `row_stride` apart in global memory,
// dense in shared memory with a row stride of tile_cols + 1 (= 1 mod 16: the bank model of fit_nd_lines holds with the
// columns as lines).  Two HBM round trips in total instead of one per dimension with O(n^2) scalar sums.
struct FitTileParams {
  void* data;           // in place
  const double* afrag;
  int afragDoubles;
  long long narrays, array_stride;  // reals between consecutive arrays
  long long inner;      // columns per array = reals per row of the last dimension
  int row_stride;       // = inner (reals)
  int tile_cols, tiles_per_array;
  double* maxabs;       // optional (real scalar elements only): atomic max per array, zero-initialised by the caller
  FitNdDim dim;         // the last dimension: n = rows, stride = tile_cols + 1
};

template <typename R, int MTO, int MTB, int MTC, bool AFG>
__global__ void __launch_bounds__(256) fit_nd_tile_kernel(const __grid_constant__ FitTileParams prm) {
  extern __shared__ __align__(16) unsigned char fastcheb_smem[];
  double* af_s = reinterpret_cast<double*>(fastcheb_smem);
  R* x = reinterpret_cast<R*>(fastcheb_smem + (AFG ? 0 : (((size_t)prm.afragDoubles * 8 + 15) / 16) * 16));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  if constexpr (!AFG)
    for (int i = tid; i < prm.afragDoubles; i += blockDim.x) af_s[i] = prm.afrag[i];
  const int rows = prm.dim.n, TS = prm.dim.stride;
  R* data = reinterpret_cast<R*>(prm.data);
  const long long ntiles = prm.narrays * prm.tiles_per_array;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long a = t / prm.tiles_per_array;
    const int cb = (int)(t - a * prm.tiles_per_array);
    const long long c0 = (long long)cb * prm.tile_cols;
    const int ncols = (int)((prm.inner - c0) < prm.tile_cols ? (prm.inner - c0) : prm.tile_cols);
    R* g = data + a * prm.array_stride + c0;
    __syncthreads();  // the previous tile has been stored (and, first time, the matrices are in place)
    for (int i = tid; i < rows * ncols; i += blockDim.x) {
      const int r = i / ncols, c = i - r * ncols;
      if constexpr (sizeof(R) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(x + r * TS + c)), "l"(g + (long long)r * prm.row_stride + c) : "memory");
      else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(x + r * TS + c)), "l"(g + (long long)r * prm.row_stride + c) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const double* afd;
    if constexpr (AFG)
      afd = prm.afrag;
    else
      afd = af_s;
    fit_nd_pass<R, MTO, MTB, MTC>(x, prm.dim, afd, rows * TS, warp, nwarps, lane, nullptr, ncols);
    __syncthreads();
    double mx = 0.0;
    for (int i = tid; i < rows * ncols; i += blockDim.x) {
      const int r = i / ncols, c = i - r * ncols;
      const R v = x[r * TS + c];
      mx = fmax(mx, fabs((double)v));
      g[(long long)r * prm.row_stride + c] = v;
    }
    if (prm.maxabs) {
      for (int off = 16; off; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(prm.maxabs + a), (unsigned long long)__double_as_longlong(mx));
    }
  }
}

}  // namespace fastcheb
