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

#ifndef FC_MANY_P1
#define FC_MANY_P1 8  // points per lane of the scalar Float64 value kernel
#endif

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
  int vec;  // rows of pts / out_v / out_d start 16-byte aligned: vector loads and stores
};

template <typename R> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// K contiguous reals as 16- or 8-byte vectors (`ptr` aligned accordingly), else one by one
template <typename R, int K>
__device__ __forceinline__ void many_vec_load(const R* __restrict__ ptr, R (&x)[K]) {
  constexpr int B = K * (int)sizeof(R);
  if constexpr (B % 16 == 0) {
    union { uint4 v[B / 16]; R r[K]; } u;
#pragma unroll
    for (int i = 0; i < B / 16; ++i) u.v[i] = reinterpret_cast<const uint4*>(ptr)[i];
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = u.r[k];
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = ptr[k];
  }
}
template <typename R, int K>
__device__ __forceinline__ void many_vec_store(R* __restrict__ ptr, const R (&x)[K]) {
  constexpr int B = K * (int)sizeof(R);
  if constexpr (B % 16 == 0) {
    union { uint4 v[B / 16]; R r[K]; } u;
#pragma unroll
    for (int k = 0; k < K; ++k) u.r[k] = x[k];
#pragma unroll
    for (int i = 0; i < B / 16; ++i) reinterpret_cast<uint4*>(ptr)[i] = u.v[i];
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) ptr[k] = x[k];
  }
}

// The M points of polynomial f, by one warp.  A lane owns P CONSECUTIVE points of every 32 P-point chunk: with 16-byte
// aligned rows (prm.vec) coordinates arrive and results leave as 16-byte vectors, one index computation per lane and chunk.
template <typename R, int E, int P, bool DER, bool FAST>
__device__ __forceinline__ void poly_points(const ManyParams& prm, long long f, int lane, const R* cs, R lb, R ub,
                                            const SpanDiv<R>& sdiv) {
  const R* __restrict__ pts = reinterpret_cast<const R*>(prm.pts) + f * prm.M;
  R* __restrict__ out_v = reinterpret_cast<R*>(prm.out_v) + f * prm.M * E;
  R* __restrict__ out_d = DER ? reinterpret_cast<R*>(prm.out_d) + f * prm.M * E : nullptr;
  const int n = prm.n;
  const R span = sdiv.span, rcp = sdiv.y;
  auto div_span = [&](R a) {
    if constexpr (FAST) {
      const R q = rmul(a, rcp);
      return rfma(rfma(-span, q, a), rcp, q);
    } else {
      return rdiv(a, span);
    }
  };
  // derivative scale (eval.jl:151), any numerator: the correction is the IEEE quotient while nothing can underflow or
  // overflow on the way (numerator exponent within +-400, Float32 +-40)
  [[maybe_unused]] auto div_any = [&](R a) {
    if constexpr (FAST) {
      bool safe;
      if constexpr (sizeof(R) == 8) {
        const unsigned e = ((unsigned)(__double_as_longlong((double)a) >> 52)) & 0x7ffu;
        safe = e > 1023u - 400u && e < 1023u + 400u;
      } else {
        const unsigned e = (__float_as_uint((float)a) >> 23) & 0xffu;
        safe = e > 127u - 40u && e < 127u + 40u;
      }
      if (safe) {
        const R q = rmul(a, rcp);
        return rfma(rfma(-span, q, a), rcp, q);
      }
    }
    return rdiv(a, span);
  };
  const bool vec = prm.vec != 0;
  auto count = [&](long long i0) {
    const long long left = prm.M - i0;
    return left >= P ? P : (left > 0 ? (int)left : 0);
  };
  auto fetch = [&](long long i0, R (&x)[P]) {
    const int c = count(i0);
    if (vec && c == P) {
      many_vec_load<R, P>(pts + i0, x);
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) x[p] = p < c ? pts[i0 + p] : lb;
    }
  };
  // the points of the next chunk are loaded while the current one is evaluated (the recurrence is ~130 dependent FP64
  // operations per chunk; an unprefetched load at its head stalls the warp for a global-memory latency every time)
  R xn[P];
  fetch((long long)lane * P, xn);
  for (long long i0 = (long long)lane * P; i0 < prm.M; i0 += 32 * P) {
    const int cnt = count(i0);
    R z[P], z2[P];
    bool live[P];
    unsigned badmask = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const R x = xn[p];
      const bool ok = (lb <= x) && (x <= ub);                               // eval.jl:64 (closed box, NaN fails)
      z[p] = ok ? rsub(div_span(rmul(rsub(x, lb), R(2))), R(1)) : R(0);     // eval.jl:65 (the IEEE quotient)
      z2[p] = rmul(R(2), z[p]);
      live[p] = ok && p < cnt;
      if (!ok && p < cnt) badmask |= 1u << p;
    }
    const bool all = cnt == P && badmask == 0;
    if (badmask && prm.bad) atomicMin(prm.bad, f * prm.M + i0 + (__ffs(badmask) - 1));
    fetch(i0 + 32 * P, xn);

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
    R ov[P * E];
    [[maybe_unused]] R od[DER ? P * E : 1];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        ov[p * E + e] = rsub(rfma(z[p], b1[p][e], cs[e]), b2[p][e]);  // eval.jl:31 / :101
        if constexpr (DER) od[p * E + e] = div_any(rmul(rsub(rfma(z[p], d1[p][e], b1[p][e]), d2[p][e]), R(2)));  // eval.jl:101, :151
      }
    if (vec && all) {
      many_vec_store<R, P * E>(out_v + i0 * E, ov);
      if constexpr (DER) many_vec_store<R, P * E>(out_d + i0 * E, od);
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        if (!live[p]) continue;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          out_v[(i0 + p) * E + e] = ov[p * E + e];
          if constexpr (DER) out_d[(i0 + p) * E + e] = od[p * E + e];
        }
      }
    }
  }
}

template <typename R, int E, int P, bool DER>
__global__ void __launch_bounds__(256) eval_many_kernel(const __grid_constant__ ManyParams prm) {
  extern __shared__ __align__(16) unsigned char fastcheb_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int n = prm.n, nE = n * E;
  R* cs = reinterpret_cast<R*>(fastcheb_smem) + (size_t)warp * ((nE + 1) & ~1);  // even stride: 16-byte aligned pairs
  const R* coefs = reinterpret_cast<const R*>(prm.coefs);
  const long long gw = (long long)blockIdx.x * nwarps + warp, tw = (long long)gridDim.x * nwarps;
  for (long long f = gw; f < prm.howmany; f += tw) {
    __syncwarp();
    for (int i = lane; i < nE; i += 32) cs[i] = coefs[f * nE + i];
    __syncwarp();
    const R lb = (R)prm.lb[prm.bounds_per_poly ? f : 0], ub = (R)prm.ub[prm.bounds_per_poly ? f : 0];
    SpanDiv<R> sdiv;
    sdiv.init(lb, ub);
    // The applicability test of the reciprocal division (SpanDiv) is made once per polynomial, not once per point: the
    // point loop exists in two copies.
    if (sdiv.fast)
      poly_points<R, E, P, DER, true>(prm, f, lane, cs, lb, ub, sdiv);
    else
      poly_points<R, E, P, DER, false>(prm, f, lane, cs, lb, ub, sdiv);
  }
}

template <typename R, int E, bool DER>
cudaError_t launch(const ManyParams& prm, int sms, cudaStream_t st) {
  // (8 points per lane measured faster only for the scalar Float64 value kernel: 8.68e10 -> 9.31e10 points/s; the derivative
  // kernel then needs 172 registers and slows down, Float32 does not move)
  constexpr int P = (E == 1 && !DER && sizeof(R) == 8) ? FC_MANY_P1 : (E >= 3 ? 2 : 4);
  const int warps = 8;
  const size_t smem = (size_t)warps * (((size_t)prm.n * E + 1) & ~(size_t)1) * sizeof(R);
  auto kern = eval_many_kernel<R, E, P, DER>;
  cudaError_t e = ensure_max_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  int bps = 1;
  e = cached_occupancy(&bps, reinterpret_cast<const void*>(kern), warps * 32, smem);
  if (e != cudaSuccess) return e;
  const long long want = (prm.howmany + warps - 1) / warps;
  // 4 CTAs per resident slot (the hardware scheduler balances the tail: +3 to +5 %; FASTCHEB_MANY_OVERSUB: experiments)
  static const int over = [] {
    const char* t = getenv("FASTCHEB_MANY_OVERSUB");
    return t ? std::max(1, std::min(16, atoi(t))) : 4;
  }();
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sms * std::max(1, bps) * over));
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
  {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    prm.vec = ((size_t)M * rs) % 16 == 0 && al16(prm.pts) && al16(prm.out_v) && (!prm.out_d || al16(prm.out_d));
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
