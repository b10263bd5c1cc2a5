// C ABI: many independent 1-D ChebPolys evaluated at their own points in one launch (include/fastcheb.h,
// fastcheb_eval_many) — the evaluation leg of BASELINE.json configs[4] (10^5 order-64 polynomials x 10^3 points).
//
// Reference: `[c.(xs) for (c, xs) in zip(cs, xss)]`, i.e. per polynomial (::ChebPoly{1})(x) eval.jl:63-70 and
// chebgradient eval.jl:174-177 (1-D Clenshaw eval.jl:19-31, derivative recurrence :88-101, scale :151).
//
// One warp owns one polynomial at a time: its n*E coefficients go to the warp's shared-memory region (they are the
// output of fastcheb_fit: back-to-back lines, element stride E), every lane runs the Clenshaw recurrence for P
// points with the coefficient read as a warp-wide broadcast.  The operation sequence is the oracle's (explicit IEEE
// fma / sub / mul / div), so results are bit-identical to it.
#include <algorithm>
#include <cstring>
#include "host_util.h"
#include "ptx_util.cuh"

namespace fastcheb {
namespace {

struct ManyParams {
  const void* coefs;   // [howmany][n][E]
  const void* pts;     // [howmany][M]
  void* out_v;         // [howmany][M][E]
  void* out_d;         // [howmany][M][E] or null
  const double* lb;    // per polynomial (bounds_per_poly) or one value
  const double* ub;
  long long* bad;      // first out-of-domain flat point index (atomicMin)
  long long howmany, M;
  int n, bounds_per_poly;
};

template <typename R> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename R, int E, int P, bool DER>
__global__ void __launch_bounds__(256) eval_many_kernel(const __grid_constant__ ManyParams prm) {
  extern __shared__ __align__(16) unsigned char fastcheb_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int n = prm.n, nE = n * E;
  R* cs = reinterpret_cast<R*>(fastcheb_smem) + (size_t)warp * ((nE + 1) & ~1);  // even stride: 16-byte aligned pairs
  const R* coefs = reinterpret_cast<const R*>(prm.coefs);
  const R* pts = reinterpret_cast<const R*>(prm.pts);
  R* out_v = reinterpret_cast<R*>(prm.out_v);
  R* out_d = reinterpret_cast<R*>(prm.out_d);
  const long long gw = (long long)blockIdx.x * nwarps + warp, tw = (long long)gridDim.x * nwarps;
  for (long long f = gw; f < prm.howmany; f += tw) {
    __syncwarp();
    for (int i = lane; i < nE; i += 32) cs[i] = coefs[f * nE + i];
    __syncwarp();
    const R lb = (R)prm.lb[prm.bounds_per_poly ? f : 0], ub = (R)prm.ub[prm.bounds_per_poly ? f : 0];
    SpanDiv<R> sdiv;
    sdiv.init(lb, ub);
    const R span = sdiv.span;
    // the points of the next chunk are loaded while the current one is evaluated (the recurrence is ~130 dependent FP64
    // operations per chunk; an unprefetched load at its head stalls the warp for a global-memory latency every time)
    R xn[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const long long ip = lane + 32 * p;
      xn[p] = ip < prm.M ? pts[f * prm.M + ip] : lb;
    }
    for (long long p0 = lane; p0 < prm.M; p0 += 32 * P) {
      R z[P], z2[P];
      bool live[P];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const long long ip = p0 + 32 * p;
        live[p] = ip < prm.M;
        const R x = xn[p];
        const long long in = ip + 32 * P;
        xn[p] = in < prm.M ? pts[f * prm.M + in] : lb;
        const bool ok = (lb <= x) && (x <= ub);                       // eval.jl:64 (closed box, NaN fails)
        z[p] = ok ? rsub(sdiv.div(rmul(rsub(x, lb), R(2))), R(1)) : R(0);  // eval.jl:65 (SpanDiv: the IEEE quotient)
        z2[p] = rmul(R(2), z[p]);
        if (live[p] && !ok) {
          live[p] = false;
          if (prm.bad) atomicMin(prm.bad, f * prm.M + ip);
        }
      }
      R b1[P][E], b2[P][E], d1[DER ? P : 1][E], d2[DER ? P : 1][E];
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int e = 0; e < E; ++e) {
          b1[p][e] = b2[p][e] = R(0);
          if constexpr (DER) d1[p][e] = d2[p][e] = R(0);
        }
      auto step = [&](int e, R cj) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const R nb = rsub(rfma(z2[p], b1[p][e], cj), b2[p][e]);                    // eval.jl:28 / :96
          if constexpr (DER) {
            const R nd = rsub(rfma(z2[p], d1[p][e], rmul(R(2), b1[p][e])), d2[p][e]);  // eval.jl:97
            d2[p][e] = d1[p][e];
            d1[p][e] = nd;
          }
          b2[p][e] = b1[p][e];
          b1[p][e] = nb;
        }
      };
      if constexpr (E == 1) {
        // scalar coefficients: two per shared-memory load (the warp's region is 16-byte aligned, pairs (c_{j-1}, c_j), j odd)
        int j = n - 1;
        if (j >= 1 && (j & 1) == 0) {
          step(0, cs[j]);
          --j;
        }
        for (; j >= 3; j -= 2) {
          const auto c2 = *reinterpret_cast<const typename Vec2<R>::type*>(cs + j - 1);
          step(0, c2.y);
          step(0, c2.x);
        }
        if (j == 1) step(0, cs[1]);
      } else {
        for (int j = n - 1; j >= 1; --j) {
#pragma unroll
          for (int e = 0; e < E; ++e) step(e, cs[j * E + e]);
        }
      }
#pragma unroll
      for (int p = 0; p < P; ++p) {
        if (!live[p]) continue;
        const long long o = (f * prm.M + p0 + 32 * p) * E;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          out_v[o + e] = rsub(rfma(z[p], b1[p][e], cs[e]), b2[p][e]);  // eval.jl:31 / :101
          if constexpr (DER)
            out_d[o + e] = rdiv(rmul(rsub(rfma(z[p], d1[p][e], b1[p][e]), d2[p][e]), R(2)), span);  // eval.jl:101, :151
        }
      }
    }
  }
}

template <typename R, int E, bool DER>
cudaError_t launch(const ManyParams& prm, int sms, cudaStream_t st) {
  constexpr int P = sizeof(R) == 8 ? (E >= 3 ? 2 : 4) : (E >= 3 ? 2 : 4);
  const int warps = 8;
  const size_t smem = (size_t)warps * (((size_t)prm.n * E + 1) & ~(size_t)1) * sizeof(R);
  auto kern = eval_many_kernel<R, E, P, DER>;
  cudaError_t e = ensure_max_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  int bps = 1;
  e = cached_occupancy(&bps, reinterpret_cast<const void*>(kern), warps * 32, smem);
  if (e != cudaSuccess) return e;
  const long long want = (prm.howmany + warps - 1) / warps;
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sms * std::max(1, bps)));
  kern<<<grid, warps * 32, smem, st>>>(prm);
  note_launch();
  return cudaGetLastError();
}

template <typename R, bool DER>
cudaError_t dispatch(int E, const ManyParams& prm, int sms, cudaStream_t st) {
  switch (E) {
    case 1: return launch<R, 1, DER>(prm, sms, st);
    case 2: return launch<R, 2, DER>(prm, sms, st);
    case 3: return launch<R, 3, DER>(prm, sms, st);
    case 4: return launch<R, 4, DER>(prm, sms, st);
    case 6: return launch<R, 6, DER>(prm, sms, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace
}  // namespace fastcheb

using namespace fastcheb;

extern "C" int fastcheb_eval_many(int64_t howmany, int64_t n, int dtype, int is_complex, int ncomp, const void* coefs,
                                  const double* lb, const double* ub, int bounds_per_poly, const void* pts, int64_t M,
                                  void* out_v, void* out_d, int64_t* status_dev, int mem, int device, void* stream) {
  if (howmany < 0 || n < 1 || M < 0 || ncomp < 1) return fail(FASTCHEB_EINVAL, "bad size");
  if (dtype != FASTCHEB_F32 && dtype != FASTCHEB_F64) return fail(FASTCHEB_EINVAL, "bad dtype %d", dtype);
  if (mem != FASTCHEB_MEM_HOST && mem != FASTCHEB_MEM_DEVICE) return fail(FASTCHEB_EINVAL, "bad mem %d", mem);
  if (howmany == 0 || M == 0) return FASTCHEB_OK;
  if (!coefs || !lb || !ub || !pts || !out_v) return fail(FASTCHEB_EINVAL, "null buffer");
  const int E = ncomp * (is_complex ? 2 : 1);
  if (!(E == 1 || E == 2 || E == 3 || E == 4 || E == 6))
    return fail(FASTCHEB_EUNSUPPORTED, "fastcheb_eval_many covers 1, 2, 3, 4 or 6 real components per element, got %d", E);
  const size_t rs = real_size(dtype);
  if ((size_t)n * E * rs * 8 > (size_t)96 * 1024)
    return fail(FASTCHEB_EUNSUPPORTED, "n = %lld too large for the batched 1-D evaluator (use one handle per polynomial)", (long long)n);
  const DeviceInfo* di = device_info(device);
  if (!di) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  if (!dg.ok()) return fail(FASTCHEB_ECUDA, "cudaSetDevice(%d) failed", device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cbytes = (size_t)howmany * n * E * rs, pbytes = (size_t)howmany * M * rs, obytes = pbytes * E;
  const size_t nb = bounds_per_poly ? (size_t)howmany : 1;
  ManyParams prm;
  memset(&prm, 0, sizeof prm);
  prm.howmany = howmany;
  prm.M = M;
  prm.n = (int)n;
  prm.bounds_per_poly = bounds_per_poly ? 1 : 0;
  // The status flag: the synchronous host path borrows a slot of the per-device pool (released when the call returns,
  // after the stream has been synchronised); the asynchronous device path writes only to the caller's status_dev
  // (may be NULL: no status) — a pool slot must not outlive the call while the kernel is still running.
  FlagSlot fs;
  struct Rel {
    FlagSlot& f;
    ~Rel() { flag_release(f); }
  } rel{fs};
  long long* flag = (long long*)status_dev;
  if (mem == FASTCHEB_MEM_HOST) {
    FC_TRY(flag_acquire(device, &fs));
    flag = fs.dev;
  }
  if (flag) FC_CUDA(set_flag_async(flag, kNoBad, st));
  prm.bad = flag;
  void* scratch = nullptr;
  struct Free {
    int dev;
    void*& p;
    cudaStream_t st;
    ~Free() {
      if (p) {
        cudaStreamSynchronize(st);  // copies / kernels enqueued on the buffer (error paths return early) are done
        ws_release(dev, p);
      }
    }
  } fr{device, scratch, st};
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  if (mem == FASTCHEB_MEM_HOST) {
    const size_t total = al(cbytes) + al(pbytes) + al(obytes) * (out_d ? 2 : 1) + 2 * al(nb * 8);
    scratch = ws_acquire(device, total);
    if (!scratch) return fail(FASTCHEB_ENOMEM, "device staging of %zu bytes", total);
    char* q = (char*)scratch;
    char* d_c = q;
    q += al(cbytes);
    char* d_p = q;
    q += al(pbytes);
    char* d_v = q;
    q += al(obytes);
    char* d_d = out_d ? q : nullptr;
    if (out_d) q += al(obytes);
    char* d_lb = q;
    q += al(nb * 8);
    char* d_ub = q;
    FC_CUDA(cudaMemcpyAsync(d_c, coefs, cbytes, cudaMemcpyHostToDevice, st));
    FC_CUDA(cudaMemcpyAsync(d_p, pts, pbytes, cudaMemcpyHostToDevice, st));
    FC_CUDA(cudaMemcpyAsync(d_lb, lb, nb * 8, cudaMemcpyHostToDevice, st));
    FC_CUDA(cudaMemcpyAsync(d_ub, ub, nb * 8, cudaMemcpyHostToDevice, st));
    prm.coefs = d_c;
    prm.pts = d_p;
    prm.out_v = d_v;
    prm.out_d = d_d;
    prm.lb = (const double*)d_lb;
    prm.ub = (const double*)d_ub;
  } else {
    prm.coefs = coefs;
    prm.pts = pts;
    prm.out_v = out_v;
    prm.out_d = out_d;
    prm.lb = lb;  // device pointers in device mode
    prm.ub = ub;
  }
  cudaError_t e;
  if (dtype == FASTCHEB_F64)
    e = out_d ? dispatch<double, true>(E, prm, di->sms, st) : dispatch<double, false>(E, prm, di->sms, st);
  else
    e = out_d ? dispatch<float, true>(E, prm, di->sms, st) : dispatch<float, false>(E, prm, di->sms, st);
  if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "batched 1-D evaluation launch failed: %s", cudaGetErrorString(e));
  if (mem == FASTCHEB_MEM_DEVICE) return FASTCHEB_OK;  // asynchronous: the caller reads status_dev after syncing
  FC_CUDA(cudaMemcpyAsync(out_v, prm.out_v, obytes, cudaMemcpyDeviceToHost, st));
  if (out_d) FC_CUDA(cudaMemcpyAsync(out_d, prm.out_d, obytes, cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaStreamSynchronize(st));
  if (*fs.host != kNoBad) {
    err_state().index = *fs.host;
    return fail(FASTCHEB_EDOMAIN, "point %lld not in domain (eval.jl:64)", *fs.host);
  }
  return FASTCHEB_OK;
}
