// Host side of the low-order N-d evaluation path (eval_lond.cuh): when it applies, padded coefficient copy, launch.
#include <algorithm>
#include <cstring>
#include "eval_lond.cuh"
#include "poly.h"

namespace fastcheb {
namespace {

template <typename R, int N, int EC, bool JAC, int CAP>
cudaError_t launch_cap(const LoNdParams& prm, const void* h_c, size_t c_bytes, int sms, cudaStream_t st) {
  auto kern = eval_lond_kernel<R, N, EC, JAC, CAP>;
  LoNdArgs<R, CAP> args;
  static_assert(sizeof(args) <= 32764, "kernel parameter space");
  args.p = prm;
  memcpy(args.c, h_c, std::min(c_bytes, sizeof(args.c)));
  const int threads = 256;
  int bps = 1;
  cudaError_t e = cached_occupancy(&bps, reinterpret_cast<const void*>(kern), threads, 0);
  if (e != cudaSuccess) return e;
  const long long tile = (long long)threads * lond_pts<R>(N, EC, JAC);
  const long long ntiles = (prm.npts + tile - 1) / tile;
  // 4 CTAs per resident slot: with exactly one CTA per slot the warps that finish first leave their slots empty (ncu: 18 of 24
  // warps active on average); the hardware scheduler balances the tail — measured +1 to +10 % (FASTCHEB_LO_OVERSUB: experiments)
  static const int over = [] {
    const char* t = getenv("FASTCHEB_LO_OVERSUB");
    return t ? std::max(1, std::min(16, atoi(t))) : 4;
  }();
  const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)sms * std::max(1, bps) * over));
  kern<<<grid, threads, 0, st>>>(args);
  note_launch();
  return cudaGetLastError();
}

template <typename R, int N, int EC, bool JAC>
cudaError_t launch(const LoNdParams& prm, const void* h_c, size_t c_bytes, int sms, cudaStream_t st) {
  if constexpr (JAC && N * EC > 6) {
    return cudaErrorInvalidValue;  // not instantiated: the Jacobian nest would spill (lond_ok keeps these on the generic kernel)
  } else {
    if (lond_cap((long long)(c_bytes / sizeof(R))) == 512) return launch_cap<R, N, EC, JAC, 512>(prm, h_c, c_bytes, sms, st);
    return launch_cap<R, N, EC, JAC, 2048>(prm, h_c, c_bytes, sms, st);
  }
}

template <typename R, int N, bool JAC>
cudaError_t launch_ec(int ec, const LoNdParams& prm, const void* h_c, size_t c_bytes, int sms, cudaStream_t st) {
  switch (ec) {
    case 1: return launch<R, N, 1, JAC>(prm, h_c, c_bytes, sms, st);
    case 2: return launch<R, N, 2, JAC>(prm, h_c, c_bytes, sms, st);
    case 3: return launch<R, N, 3, JAC>(prm, h_c, c_bytes, sms, st);
    case 4: return launch<R, N, 4, JAC>(prm, h_c, c_bytes, sms, st);
  }
  return cudaErrorInvalidValue;
}

template <typename R, bool JAC>
cudaError_t launch_n(int n, int ec, const LoNdParams& prm, const void* h_c, size_t c_bytes, int sms, cudaStream_t st) {
  switch (n) {
    case 2: return launch_ec<R, 2, JAC>(ec, prm, h_c, c_bytes, sms, st);
    case 3: return launch_ec<R, 3, JAC>(ec, prm, h_c, c_bytes, sms, st);
    case 4: return launch_ec<R, 4, JAC>(ec, prm, h_c, c_bytes, sms, st);
  }
  return cudaErrorInvalidValue;
}

bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

int n1_padded(const fastcheb_poly* p) { return 1 + ((int)p->dims[0] - 1 + 3) / 4 * 4; }

}  // namespace

// SpanDiv's test (ptx_util.cuh) on the host, for dimension d: the span in the polynomial's real type has a mantissa that
// is not all ones and an exponent well inside the range — the low-order kernels then divide by it without a per-point test
bool span_fast(const fastcheb_poly* p, int d) {
  if (p->dtype == FASTCHEB_F64) {
    const volatile double span = p->ub[d] - p->lb[d];
    const double sp = span;
    unsigned long long u;
    memcpy(&u, &sp, sizeof u);
    const unsigned e = (unsigned)(u >> 52) & 0x7ffu;
    return (u & 0xfffffffffffffull) != 0xfffffffffffffull && e > 1023u - 500u && e < 1023u + 500u;
  }
  const volatile float lbf = (float)p->lb[d], ubf = (float)p->ub[d];
  const volatile float span = ubf - lbf;
  const float sp = span;
  unsigned u;
  memcpy(&u, &sp, sizeof u);
  const unsigned e = (u >> 23) & 0xffu;
  return (u & 0x7fffffu) != 0x7fffffu && e > 127u - 60u && e < 127u + 60u;
}

bool lond_ok(const fastcheb_poly* p, const void* pts, int64_t npts, const void* out_v, int64_t v_stride, const void* out_J,
             int64_t J_stride, bool jac, int mode) {
  static const bool off = getenv("FASTCHEB_LOND") != nullptr && atoi(getenv("FASTCHEB_LOND")) == 0;  // experiments
  if (off || p->algo.load() != FASTCHEB_ALGO_AUTO) return false;  // FASTCHEB_ALGO_CLENSHAW keeps the generic kernel
  if (p->ndim < 2 || p->ndim > kLoNdMaxN || p->E > 4 || mode != 0 || npts < kLoMinBatch) return false;
  long long tot = (long long)n1_padded(p) * p->E;
  for (int d = 1; d < p->ndim; ++d) tot *= p->dims[d];
  if (tot > kLoNdCap) return false;
  // register budget of the Jacobian nest (2 or 4 points per thread): N (N + 1) E states per point
  if (jac && p->ndim * p->E > 6) return false;
  if (!aligned16(pts) || !aligned16(out_v) || v_stride != p->E) return false;
  if (jac && (!aligned16(out_J) || J_stride != (int64_t)p->ndim * p->E)) return false;
  for (int d = 0; d < p->ndim; ++d)
    if (!span_fast(p, d)) return false;
  return true;
}

int lond_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, void* out_J, bool jac, long long* bad_dev,
              long long idx_base, cudaStream_t st) {
  LoPlan& lo = p->lo;
  const size_t rs = real_size(p->dtype);
  const int N = p->ndim, E = p->E, n1 = (int)p->dims[0], n1p = n1_padded(p);
  long long lines = 1;
  for (int d = 1; d < N; ++d) lines *= p->dims[d];
  if (!lo.ready.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(lo.mu);
    if (!lo.ready.load(std::memory_order_relaxed)) {
      DeviceGuard dg(p->device);
      std::vector<unsigned char> raw(p->coef_bytes), h((size_t)lines * n1p * E * rs, 0);
      cudaStream_t s2 = cudaStreamPerThread;
      FC_CUDA(cudaMemcpyAsync(raw.data(), p->d_coefs, p->coef_bytes, cudaMemcpyDeviceToHost, s2));
      FC_CUDA(cudaStreamSynchronize(s2));
      for (long long l = 0; l < lines; ++l)  // dimension 1 padded with zeros: exact no-op recurrence steps
        memcpy(h.data() + (size_t)l * n1p * E * rs, raw.data() + (size_t)l * n1 * E * rs, (size_t)n1 * E * rs);
      lo.h_c.swap(h);
      lo.ngroups = (n1p - 1) / 4;
      lo.ready.store(1, std::memory_order_release);
    }
  }
  LoNdParams prm;
  memset(&prm, 0, sizeof prm);
  prm.pts = pts;
  prm.out_v = out_v;
  prm.out_J = out_J;
  prm.bad = bad_dev;
  prm.npts = npts;
  prm.idx_base = idx_base;
  prm.ngroups1 = lo.ngroups;
  int str = 1;
  for (int d = 0; d < N; ++d) {
    prm.n[d] = (int)p->dims[d];
    prm.stride[d] = str;
    str *= d == 0 ? n1p : (int)p->dims[d];
    prm.lb[d] = p->lb[d];
    prm.ub[d] = p->ub[d];
  }
  const int sms = p->di->sms;
  const void* hc = lo.h_c.data();
  const size_t cb = lo.h_c.size();
  const cudaError_t e = p->dtype == FASTCHEB_F64
                            ? (jac ? launch_n<double, true>(N, E, prm, hc, cb, sms, st) : launch_n<double, false>(N, E, prm, hc, cb, sms, st))
                            : (jac ? launch_n<float, true>(N, E, prm, hc, cb, sms, st) : launch_n<float, false>(N, E, prm, hc, cb, sms, st));
  if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "low-order N-d kernel launch failed: %s", cudaGetErrorString(e));
  return FASTCHEB_OK;
}

}  // namespace fastcheb
