// C ABI: many independent 1-D ChebPolys evaluated at ONE SHARED set of points (include/fastcheb.h,
// fastcheb_eval_many_shared) — `[c.(xs) for c in cs]` with the same xs for every polynomial, e.g. all 10^5 fitted
// polynomials of BASELINE.json configs[4] sampled on a common grid.
//
// Reference per value: (::ChebPoly{1})(x) /root/reference/src/eval.jl:63-70 (domain check :64, affine map :65, Clenshaw
// :19-31); derivative: chebgradient eval.jl:174-177 (scale :151).
//
// With shared points the batch IS a matrix product,  V[f, m] = sum_k C[f, k] T_k(z_m)  (howmany x n times n x M), so it
// runs on the FP64 tensor cores (mma.sync m8n8k4, SASS DMMA) at the pipe's rate instead of the Clenshaw recurrence's
// 50 % ceiling:
//   1. cheb_table_kernel: T_k(z_m) (or T_k'(z_m)) for the M points once, into an L2-resident table; z is formed with
//      the oracle's operation sequence; the domain check happens here;
//   2. eval_shared_kernel: a CTA owns 128 rows (polynomial x real component); every warp keeps the A fragments of its
//      16 rows (all n coefficients) in REGISTERS for the whole launch and sweeps the points in tiles of 64 whose table
//      rows stream through shared memory (cp.async, double-buffered): per k-step 8 B fragments feed 16 DMMAs.  The
//      k = 0 term (T_0 = 1) initialises the accumulators.  Results leave as 16-byte stores, 64 contiguous bytes per row
//      and tile.
// Float32 data runs the same FP64 kernel (coefficients widened once in the A fragments, results rounded at the end).
// The summation order differs from Clenshaw's: results agree with the oracle within 16 eps sum|c|, not bit for bit.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "host_util.h"
#include "ptx_util.cuh"

namespace fastcheb {
namespace {

constexpr int kSharedKS = 16;   // k-steps: n - 1 <= 64
constexpr int kSharedPT = 64;   // points per tile
constexpr int kSharedHeader = 256;  // mbarriers + release counters

struct TableParams {
  const void* pts;
  double* table;   // [M][RS]: T_1..T_{n-1} (zero padded to 4 KS, row stride RS)
  long long* bad;
  long long M;
  int n, RS, der;
  double lb, ub;              // the bounds by value (host mode) ...
  const double *lbp, *ubp;    // ... or in device memory (device mode: no read-back, the call stays asynchronous)
};

template <typename R>
__global__ void __launch_bounds__(256) cheb_table_kernel(const __grid_constant__ TableParams prm) {
  const R lb = (R)(prm.lbp ? *prm.lbp : prm.lb), ub = (R)(prm.ubp ? *prm.ubp : prm.ub);
  SpanDiv<R> sdiv;
  sdiv.init(lb, ub);
  const R* pts = reinterpret_cast<const R*>(prm.pts);
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < prm.M; m += (long long)gridDim.x * blockDim.x) {
    const R x = pts[m];
    const bool ok = (lb <= x) && (x <= ub);                                            // eval.jl:64
    const double z = ok ? (double)rsub(sdiv.div(rmul(rsub(x, lb), R(2))), R(1)) : 0.0;  // eval.jl:65, oracle's sequence
    if (!ok && prm.bad) atomicMin(prm.bad, m);
    double* row = prm.table + m * prm.RS;
    const double z2 = 2.0 * z;
    double tkm = 1.0, tk = z, dkm = 0.0, dk = 1.0;  // T_0, T_1, T_0', T_1'
    for (int k = 1; k < prm.n; ++k) {
      row[k - 1] = prm.der ? dk : tk;
      const double tn = fma(z2, tk, -tkm);
      const double dn = fma(z2, dk, 2.0 * tk) - dkm;
      tkm = tk;
      tk = tn;
      dkm = dk;
      dk = dn;
    }
    for (int k = prm.n; k <= prm.RS; ++k) row[k - 1] = 0.0;
  }
}

struct SharedParams {
  const void* coefs;    // [howmany][n][E]
  const double* table;  // [M][RS]
  void* out;            // [howmany][M][E]
  long long rows;       // howmany * E
  long long M;
  int n, E, RS, der, nstage;
  double lb, ub;
  const double *lbp, *ubp;  // as in TableParams
};

// the general (ragged / vector-element / unaligned) form of the epilogue store, kept out of line
template <typename R>
__device__ __noinline__ void store_slow(R* orow, long long m, long long M, int E, double v0, double v1) {
  if (m < M) orow[m * E] = (R)v0;
  if (m + 1 < M) orow[(m + 1) * E] = (R)v1;
}

// RT: row tiles (of 8 rows) per warp.  With RT = 2 every B fragment feeds two DMMAs (two accumulator chains interleave by
// construction); the column tiles are then taken two at a time with k innermost so that only four accumulators are live
// next to the 2 x 16 A fragments (128 registers, two CTAs per SM — all eight column tiles at once would need ~250).
// KSU: k-steps executed, a compile-time constant (4, 8, 12 or 16 >= ceil((n - 1) / 4); coefficients and table rows are
// zero-padded): the tile loop is then one straight-line block whose B-fragment loads ptxas schedules across k-steps.
// (With a run-time count and a uniform exit per k-step every k-step was a basic block of its own: the loads' latency was
// exposed 16 times per tile, and branch / NOP slots were a quarter of the stall samples.)
template <typename R, int RT, int KSU, bool DER>
__global__ void __launch_bounds__(256, 2) eval_shared_kernel(const __grid_constant__ SharedParams prm) {
  constexpr int kSharedRows = 8 * 8 * RT;  // rows per CTA
  extern __shared__ __align__(128) unsigned char fastcheb_smem[];
  // [256 B header: full[NST] mbarriers | released[NST] counters] [NST tile buffers]
  uint64_t* full = reinterpret_cast<uint64_t*>(fastcheb_smem);
  unsigned int* released = reinterpret_cast<unsigned int*>(fastcheb_smem + 64);
  double* tiles = reinterpret_cast<double*>(fastcheb_smem + kSharedHeader);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = lane & 3, q = lane >> 2;
  const int n = prm.n, E = prm.E, RS = prm.RS, NST = prm.nstage;
  const R* coefs = reinterpret_cast<const R*>(prm.coefs);
  R* out = reinterpret_cast<R*>(prm.out);
  const long long ntiles = (prm.M + kSharedPT - 1) / kSharedPT;
  const long long nblocks = (prm.rows + kSharedRows - 1) / kSharedRows;
  const int tileDoubles = kSharedPT * RS;

  // Work = the (row block, point tile) pairs in row-block-major order; a CTA owns one CONTIGUOUS range of them (balanced
  // to within one tile), so the table tiles stream through the ring without a restart at a row-block boundary and the A
  // fragments are reloaded only when the row block changes.
  const long long total = nblocks * ntiles;
  const long long g0 = total * blockIdx.x / gridDim.x, g1 = total * (blockIdx.x + 1) / gridDim.x;
  if (g0 >= g1) return;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      released[s] = 0;
    }
    fence_mbar_init();
  }
  __syncthreads();  // the only CTA-wide barrier
  // A tile is one contiguous, 16-byte aligned piece of the table (the last tile is zero-padded by the host): ONE bulk
  // copy (TMA engine) signalled on the stage's mbarrier.  No producer warp and no barrier in the loop: the LAST warp to
  // release a stage (shared-memory counter) refills it with the tile NST visits ahead, so the warps of a CTA drift up to
  // NST - 1 tiles apart instead of meeting twice per tile.
  auto issue = [&](long long g) {  // one thread
    const int s = (int)((g - g0) % NST);
    const uint32_t bytes = (uint32_t)tileDoubles * 8u;
    const char* src = reinterpret_cast<const char*>(prm.table + (g % ntiles) * (long long)tileDoubles);
    char* dst = reinterpret_cast<char*>(tiles + (size_t)s * tileDoubles);
    mbar_arrive_expect_tx(&full[s], bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t len = bytes - o < 32768u ? bytes - o : 32768u;
      bulk_g2s(dst + o, src + o, len, &full[s]);
    }
  };
  if (tid == 0)
    for (long long g = g0; g < g0 + NST && g < g1; ++g) issue(g);
  // A fragments of this warp's RT x 8 rows: lane (c, q) holds C[row q][k = 1 + 4 ks + c]; the k = 0 column separately
  double A[RT][KSU], c0[RT][2];
  long long rowq[RT];
  bool lived[RT][2];
  R* orow[RT];  // this lane's output row of each row tile: point m at orow[m * E]
  long long cur = -1;
  // derivative pass: eval.jl:151's factor 2 / (ub - lb) is folded into the coefficients as they are loaded (one rounding
  // more than the reference's (d * 2) / span — this path is held to the tolerance, not to bit-identity)
  const double dscale = DER ? 2.0 / (double)rsub((R)(prm.ubp ? *prm.ubp : prm.ub), (R)(prm.lbp ? *prm.lbp : prm.lb)) : 1.0;
  {
    for (long long g = g0; g < g1; ++g) {
      const int b = (int)((g - g0) % NST);
      const uint32_t parity = (uint32_t)(((g - g0) / NST) & 1);
      const long long blk = g / ntiles, pt = g - blk * ntiles;
      if (blk != cur) {
        cur = blk;
#pragma unroll
        for (int t = 0; t < RT; ++t) {
          const long long r0 = blk * kSharedRows + warp * (8 * RT) + t * 8;
          rowq[t] = r0 + q;
          const bool liveq = rowq[t] < prm.rows;
          const long long f = liveq ? rowq[t] / E : 0;
          const int e = liveq ? (int)(rowq[t] - f * E) : 0;
          const R* cr = coefs + (f * n) * E + e;  // element k at cr[k * E]
          orow[t] = out + (f * prm.M) * E + e;
#pragma unroll
          for (int ks = 0; ks < KSU; ++ks) {
            const int k = 1 + 4 * ks + c;
            A[t][ks] = (liveq && k < n) ? (double)cr[(long long)k * E] * dscale : 0.0;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // D fragment: lane (c, q) holds out[row q][point 2c + h]: the k = 0 coefficient of row q initialises it
            lived[t][h] = liveq;
            c0[t][h] = (liveq && !DER) ? (double)cr[0] : 0.0;
          }
        }
      }
      mbar_wait(&full[b], parity);
      const double* T = tiles + (size_t)b * tileDoubles;
      const long long m0 = pt * kSharedPT;
      // epilogue of one column tile: lane (c, q) holds points 8 j + 2c, 8 j + 2c + 1 of row q
      // Fast path, decided once per tile and row: scalar elements, the whole tile inside M and the lane's pairs of points
      // 2-element aligned — one vector store and nothing else in the unrolled body (the body's size matters: the
      // instruction cache, not the tensor pipe, limited the first versions).  Everything else goes through store_slow.
      bool fastst[RT];
#pragma unroll
      for (int t = 0; t < RT; ++t)
        fastst[t] = lived[t][0] && E == 1 && m0 + kSharedPT <= prm.M &&
                    ((reinterpret_cast<uintptr_t>(orow[t] + m0 + 2 * c) & (2 * sizeof(R) - 1)) == 0);
      auto store = [&](int t, int j, double v0, double v1) {
        if (fastst[t]) {
          R* o = orow[t] + m0 + 8 * j + 2 * c;
          if constexpr (sizeof(R) == 8)
            *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
          else
            *reinterpret_cast<float2*>(o) = make_float2((float)v0, (float)v1);
        } else if (lived[t][0]) {
          store_slow<R>(orow[t], m0 + 8 * j + 2 * c, prm.M, E, v0, v1);
        }
      };
      if constexpr (RT == 1) {
        // one row tile per warp: all 8 column tiles of the point tile accumulate together (k outermost)
        double acc[kSharedPT / 8][2];
#pragma unroll
        for (int j = 0; j < kSharedPT / 8; ++j) {
          acc[j][0] = c0[0][0];
          acc[j][1] = c0[0][1];
        }
#pragma unroll
        for (int ks = 0; ks < KSU; ++ks) {
#pragma unroll
          for (int j = 0; j < kSharedPT / 8; ++j) {
            // B fragment: lane (c, q) holds T[point 8 j + q][k = 1 + 4 ks + c]
            const double bv = T[(8 * j + q) * RS + 4 * ks + c];
            dmma884(acc[j][0], acc[j][1], A[0][ks], bv);
          }
        }
#pragma unroll
        for (int j = 0; j < kSharedPT / 8; ++j) store(0, j, acc[j][0], acc[j][1]);
      } else {
        // RT row tiles per warp, column tiles two at a time (k innermost): every B fragment feeds RT DMMAs, 2 RT
        // independent accumulator chains per warp, and only those 2 RT accumulators are live next to the A fragments —
        // 128 registers, two CTAs per SM (all 8 column tiles at once need ~250)
        // (unrolled although `no_instruction` is 15 % of the stall samples: as a run-time loop of four iterations the kernel
        // drops from 68 % to 62 %)
#pragma unroll 2
        for (int j = 0; j < kSharedPT / 8; j += 2) {
          double acc[RT][2][2];
#pragma unroll
          for (int t = 0; t < RT; ++t)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              acc[t][jj][0] = c0[t][0];
              acc[t][jj][1] = c0[t][1];
            }
          const double* T0 = T + (8 * j + q) * RS + c;
          const double* T1 = T0 + 8 * RS;
#pragma unroll
          for (int ks = 0; ks < KSU; ++ks) {
            const double b0 = T0[4 * ks], b1 = T1[4 * ks];
#pragma unroll
            for (int t = 0; t < RT; ++t) {
              dmma884(acc[t][0][0], acc[t][0][1], A[t][ks], b0);
              dmma884(acc[t][1][0], acc[t][1][1], A[t][ks], b1);
            }
          }
#pragma unroll
          for (int t = 0; t < RT; ++t)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) store(t, j + jj, acc[t][jj][0], acc[t][jj][1]);
        }
      }
      // release the stage; the last warp out refills it
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&released[b], 1u) == (unsigned)(blockDim.x >> 5) - 1u) {
          released[b] = 0;
          if (g + NST < g1) {
            fence_proxy_async();
            issue(g + NST);
          }
        }
      }
    }
  }
}

}  // namespace
}  // namespace fastcheb

using namespace fastcheb;

extern "C" int fastcheb_eval_many_shared(int64_t howmany, int64_t n, int dtype, int is_complex, int ncomp, const void* coefs,
                                         const double* lb, const double* ub, const void* pts, int64_t M, void* out_v, void* out_d,
                                         int64_t* status_dev, int mem, int device, void* stream) {
  if (howmany < 0 || n < 1 || M < 0 || ncomp < 1) return fail(FASTCHEB_EINVAL, "bad size");
  if (dtype != FASTCHEB_F32 && dtype != FASTCHEB_F64) return fail(FASTCHEB_EINVAL, "bad dtype %d", dtype);
  if (mem != FASTCHEB_MEM_HOST && mem != FASTCHEB_MEM_DEVICE) return fail(FASTCHEB_EINVAL, "bad mem %d", mem);
  if (howmany == 0 || M == 0) return FASTCHEB_OK;
  if (!coefs || !lb || !ub || !pts || !out_v) return fail(FASTCHEB_EINVAL, "null buffer");
  if (n - 1 > 4 * kSharedKS)
    return fail(FASTCHEB_EUNSUPPORTED, "fastcheb_eval_many_shared covers n <= %d, got %lld", 4 * kSharedKS + 1, (long long)n);
  const int E = ncomp * (is_complex ? 2 : 1);
  const size_t rs = real_size(dtype);
  const DeviceInfo* di = device_info(device);
  if (!di) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  if (!dg.ok()) return fail(FASTCHEB_ECUDA, "cudaSetDevice(%d) failed", device);
  cudaStream_t st = (cudaStream_t)stream;
  // lb / ub: host doubles in host mode (passed to the kernels by value), device doubles in device mode (as in
  // fastcheb_eval_many; the kernels read them through the pointers: no read-back, the device-mode call never synchronises)
  double hlb = 0.0, hub = 0.0;
  const double *lbp = nullptr, *ubp = nullptr;
  if (mem == FASTCHEB_MEM_DEVICE) {
    lbp = lb;
    ubp = ub;
  } else {
    hlb = lb[0];
    hub = ub[0];
  }
  const int KS = (int)((n - 1 + 3) / 4);
  const int ksu = KS <= 4 ? 4 : (KS <= 8 ? 8 : (KS <= 12 ? 12 : 16));  // k-steps the instantiation executes
  // a table row holds the 4 ksu columns the kernel reads (zero beyond n - 1), padded to a stride = 4 (mod 16) doubles: the
  // B-fragment loads of a half warp hit 16 distinct banks
  int RS = 4 * ksu;
  while (RS % 16 != 4) RS += 2;
  const long long ntiles = (M + kSharedPT - 1) / kSharedPT;
  const size_t tableBytes = (size_t)ntiles * kSharedPT * RS * sizeof(double);
  const size_t cbytes = (size_t)howmany * n * E * rs, pbytes = (size_t)M * rs, obytes = (size_t)howmany * M * E * rs;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const bool host = mem == FASTCHEB_MEM_HOST;
  const size_t total = al(tableBytes) + (host ? al(cbytes) + al(pbytes) + al(obytes) : 0);
  char* scratch = nullptr;
  FC_CUDA(pool_alloc((void**)&scratch, total, st));
  struct Free {
    char* p;
    cudaStream_t st;
    ~Free() { cudaFreeAsync(p, st); }
  } fr{scratch, st};
  double* d_table = (double*)scratch;
  const void* d_c = coefs;
  const void* d_p = pts;
  char* d_o = nullptr;
  if (host) {
    char* q = scratch + al(tableBytes);
    FC_CUDA(cudaMemcpyAsync(q, coefs, cbytes, cudaMemcpyHostToDevice, st));
    d_c = q;
    q += al(cbytes);
    FC_CUDA(cudaMemcpyAsync(q, pts, pbytes, cudaMemcpyHostToDevice, st));
    d_p = q;
    q += al(pbytes);
    d_o = q;
  }
  FlagSlot fs;
  struct Rel {
    FlagSlot& f;
    ~Rel() { flag_release(f); }
  } rel{fs};
  long long* flag = (long long*)status_dev;
  if (host) {
    FC_TRY(flag_acquire(device, &fs));
    flag = fs.dev;
  }
  if (flag) FC_CUDA(set_flag_async(flag, kNoBad, st));
  FC_CUDA(cudaMemsetAsync(d_table + (size_t)M * RS, 0, tableBytes - (size_t)M * RS * sizeof(double), st));  // padding rows of the last tile
  const long long rows = (long long)howmany * E;
  // row tiles per warp: 2 (every B fragment feeds two DMMAs: Float64 73 % vs 66 % of peak at 10^3 points, Float32 data
  // 70 % vs 66 %); FASTCHEB_SHARED_RT=1 keeps one (experiments).  Orders below 49 use the one-tile instantiations.
  static const int rt_env = [] {
    const char* e = getenv("FASTCHEB_SHARED_RT");
    return e ? atoi(e) : 0;
  }();
  const int rt = rt_env == 1 ? 1 : 2;
  const int rte = (rt == 2 && ksu == 16) ? 2 : 1;
  const long long nblocks = (rows + 64 * rte - 1) / (64 * rte);
  // as many tile stages as two resident CTAs per SM leave room for (3 at n = 65), at least 2
  const size_t tileBytes = (size_t)kSharedPT * RS * sizeof(double);
  int nstage = 4;
  const size_t ctas = 2;  // resident CTAs per SM the kernel is built for (128 registers; three CTAs of 80 registers: 62 % vs 65 %)
  while (nstage > 2 && ctas * (kSharedHeader + nstage * tileBytes + 1024) > di->smem_optin) --nstage;
  const size_t smem = kSharedHeader + (size_t)nstage * tileBytes;
  for (int der = 0; der < (out_d ? 2 : 1); ++der) {
    TableParams tp;
    memset(&tp, 0, sizeof tp);
    tp.pts = d_p;
    tp.table = d_table;
    tp.bad = der == 0 ? flag : nullptr;
    tp.M = M;
    tp.n = (int)n;
    tp.RS = RS;
    tp.der = der;
    tp.lb = hlb;
    tp.ub = hub;
    tp.lbp = lbp;
    tp.ubp = ubp;
    // one warp per CTA: every lane writes its own table row (32 sectors per store instruction), so the kernel is bound by
    // the store path of the SMs it runs on — 1000 points on 4 SMs took 13.8 us, a warp per SM takes ~2
    const int tthreads = 32;
    const int tgrid = (int)std::max<long long>(1, std::min<long long>((M + tthreads - 1) / tthreads, (long long)di->sms * 32));
    if (dtype == FASTCHEB_F64)
      cheb_table_kernel<double><<<tgrid, tthreads, 0, st>>>(tp);
    else
      cheb_table_kernel<float><<<tgrid, tthreads, 0, st>>>(tp);
    FC_CUDA(cudaGetLastError());
    SharedParams sp;
    memset(&sp, 0, sizeof sp);
    sp.coefs = d_c;
    sp.table = d_table;
    sp.out = host ? (void*)(d_o) : (der ? out_d : out_v);
    sp.rows = rows;
    sp.M = M;
    sp.n = (int)n;
    sp.E = E;
    sp.RS = RS;
    sp.der = der;
    sp.nstage = nstage;
    sp.lb = hlb;
    sp.ub = hub;
    sp.lbp = lbp;
    sp.ubp = ubp;
    const void* kern;
#define FC_SH2(RR, DD)                                                                                                 \
  (rte == 2    ? (const void*)eval_shared_kernel<RR, 2, 16, DD>                                                        \
   : ksu == 4  ? (const void*)eval_shared_kernel<RR, 1, 4, DD>                                                         \
   : ksu == 8  ? (const void*)eval_shared_kernel<RR, 1, 8, DD>                                                         \
   : ksu == 12 ? (const void*)eval_shared_kernel<RR, 1, 12, DD>                                                        \
               : (const void*)eval_shared_kernel<RR, 1, 16, DD>)
#define FC_SH(RR) (der ? FC_SH2(RR, true) : FC_SH2(RR, false))
    if (dtype == FASTCHEB_F64)
      kern = FC_SH(double);
    else
      kern = FC_SH(float);
#undef FC_SH2
#undef FC_SH
    FC_CUDA(ensure_max_smem(kern));
    int bps = 1;
    FC_CUDA(cached_occupancy(&bps, kern, 256, smem));
    const long long slots = (long long)di->sms * std::max(1, bps);
    // one contiguous range of (row block, point tile) pairs per resident CTA
    const int grid = (int)std::max<long long>(1, std::min<long long>(nblocks * ntiles, slots));
    void* args[] = {&sp};
    FC_CUDA(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(256), args, smem, st));
    FC_CUDA(cudaGetLastError());
    note_launch(2);
    if (host) FC_CUDA(cudaMemcpyAsync(der ? out_d : out_v, d_o, obytes, cudaMemcpyDeviceToHost, st));
  }
  if (!host) return FASTCHEB_OK;  // asynchronous: the caller reads status_dev after synchronising
  FC_CUDA(cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaStreamSynchronize(st));
  if (*fs.host != kNoBad) {
    err_state().index = *fs.host;
    return fail(FASTCHEB_EDOMAIN, "point %lld not in domain (eval.jl:64)", *fs.host);
  }
  return FASTCHEB_OK;
}
