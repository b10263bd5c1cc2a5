// Host side of the low-order 1-D evaluation path (eval_lo1d.cuh): when it applies, coefficient copy, launch.
#include <algorithm>
#include <cstring>
#include "eval_lo1d.cuh"
#include "poly.h"

namespace fastcheb {
namespace {

template <typename R, int EC, bool JAC, int NCAP>
cudaError_t launch_cap(const Lo1dParams& prm, const void* h_c, int sms, cudaStream_t st) {
  auto kern = eval_lo1d_kernel<R, EC, JAC, NCAP>;
  Lo1dArgs<R, EC, NCAP> args;
  static_assert(sizeof(args) <= 32764, "kernel parameter space");
  args.p = prm;
  memcpy(args.c, h_c, sizeof(args.c));
  const int threads = 256;
  int bps = 1;
  cudaError_t e = cached_occupancy(&bps, reinterpret_cast<const void*>(kern), threads, 0);
  if (e != cudaSuccess) return e;
  const long long tile = (long long)threads * lo_pts<R>(EC, NCAP, JAC);
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

template <typename R, int EC, bool JAC>
cudaError_t launch(int cap, const Lo1dParams& prm, const void* h_c, int sms, cudaStream_t st) {
  switch (cap) {
    case 17: return launch_cap<R, EC, JAC, 17>(prm, h_c, sms, st);
    case 129: return launch_cap<R, EC, JAC, 129>(prm, h_c, sms, st);
    case 513: return launch_cap<R, EC, JAC, 513>(prm, h_c, sms, st);
  }
  return cudaErrorInvalidValue;
}

template <typename R, bool JAC>
cudaError_t launch_ec(int ec, int cap, const Lo1dParams& prm, const void* h_c, int sms, cudaStream_t st) {
  switch (ec) {
    case 1: return launch<R, 1, JAC>(cap, prm, h_c, sms, st);
    case 2: return launch<R, 2, JAC>(cap, prm, h_c, sms, st);
    case 3: return launch<R, 3, JAC>(cap, prm, h_c, sms, st);
    case 4: return launch<R, 4, JAC>(cap, prm, h_c, sms, st);
  }
  return cudaErrorInvalidValue;
}

bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

}  // namespace

bool lo1d_ok(const fastcheb_poly* p, const void* pts, int64_t npts, const void* out_v, int64_t v_stride, const void* out_J,
             int64_t J_stride, bool jac, int mode) {
  static const bool off = getenv("FASTCHEB_LO1D") != nullptr && atoi(getenv("FASTCHEB_LO1D")) == 0;  // experiments
  if (off || p->algo.load() != FASTCHEB_ALGO_AUTO) return false;  // FASTCHEB_ALGO_CLENSHAW keeps the generic kernel
  if (p->ndim != 1 || p->dims[0] > kLoMaxNAny || p->E > 4 || mode != 0 || npts < kLoMinBatch) return false;
  if (!aligned16(pts) || !aligned16(out_v) || v_stride != p->E) return false;
  if (jac && (!aligned16(out_J) || J_stride != p->E)) return false;
  return span_fast(p, 0);
}

int lo1d_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, void* out_J, bool jac, long long* bad_dev,
              long long idx_base, cudaStream_t st) {
  LoPlan& lo = p->lo;
  const size_t rs = real_size(p->dtype);
  const int n = (int)p->dims[0], E = p->E;
  if (!lo.ready.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(lo.mu);
    if (!lo.ready.load(std::memory_order_relaxed)) {
      DeviceGuard dg(p->device);
      std::vector<unsigned char> h((size_t)lo_cap(n) * E * rs, 0);  // zero above c_{n-1}: exact no-op recurrence steps
      cudaStream_t s2 = cudaStreamPerThread;
      FC_CUDA(cudaMemcpyAsync(h.data(), p->d_coefs, (size_t)n * E * rs, cudaMemcpyDeviceToHost, s2));
      FC_CUDA(cudaStreamSynchronize(s2));
      lo.h_c.swap(h);
      lo.ngroups = (n - 1 + 3) / 4;
      lo.ready.store(1, std::memory_order_release);
    }
  }
  Lo1dParams prm;
  memset(&prm, 0, sizeof prm);
  prm.pts = pts;
  prm.out_v = out_v;
  prm.out_J = out_J;
  prm.bad = bad_dev;
  prm.npts = npts;
  prm.idx_base = idx_base;
  prm.ngroups = lo.ngroups;
  prm.lb = p->lb[0];
  prm.ub = p->ub[0];
  const int sms = p->di->sms;
  const cudaError_t e =
      p->dtype == FASTCHEB_F64
          ? (jac ? launch_ec<double, true>(E, lo_cap(n), prm, lo.h_c.data(), sms, st) : launch_ec<double, false>(E, lo_cap(n), prm, lo.h_c.data(), sms, st))
          : (jac ? launch_ec<float, true>(E, lo_cap(n), prm, lo.h_c.data(), sms, st) : launch_ec<float, false>(E, lo_cap(n), prm, lo.h_c.data(), sms, st));
  if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "low-order 1-D kernel launch failed: %s", cudaGetErrorString(e));
  return FASTCHEB_OK;
}

}  // namespace fastcheb
