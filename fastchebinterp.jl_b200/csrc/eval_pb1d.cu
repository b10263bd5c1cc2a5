// Host side of the 1-D product-basis evaluation path (eval_pb1d.cuh): coefficient rewrite, plan-time self-check,
// launch.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "eval_pb1d.cuh"
#include "poly.h"

namespace fastcheb {
namespace {

template <typename R, int AMAX, int EC, int P, bool JAC>
cudaError_t launch_p(const Pb1dParams& prm, const void* h_D, const void* h_D2, size_t d_count, int sms, cudaStream_t st) {
  auto kern = eval_pb1d_kernel<R, AMAX, EC, P, JAC>;
  constexpr int CAP = pb_cap(AMAX, EC, JAC);
  constexpr int CAP2 = JAC ? pb_cap2<R>(AMAX, EC) : 1;
  static_assert(!JAC || CAP2 == CAP || pb_group(EC, true) == 1, "a shortened second table needs the single-row layout");
  if (d_count > (size_t)CAP || (JAC && (!h_D2 || d_count > (size_t)CAP2))) return cudaErrorInvalidValue;
  Pb1dArgs<R, CAP, CAP2> args;
  static_assert(sizeof(args) <= 32764, "kernel parameter space");
  args.p = prm;
  memcpy(args.D, h_D, d_count * sizeof(R));
  if constexpr (JAC) memcpy(args.D2, h_D2, d_count * sizeof(R));
  int threads = 256;
  if (const char* t = getenv("FASTCHEB_PB_THREADS")) {
    const int v = atoi(t);
    if (v >= 32 && v <= 256 && v % 32 == 0) threads = v;
  }
  const long long per_sm = (prm.npts + (long long)sms * P - 1) / ((long long)sms * P);
  if (per_sm < threads) threads = (int)std::max<long long>(32, (per_sm + 31) / 32 * 32);
  int bps = 1;
  cudaError_t e = cached_occupancy(&bps, reinterpret_cast<const void*>(kern), threads, 0);
  if (e != cudaSuccess) return e;
  const long long tile = (long long)threads * P;
  const long long ntiles = (prm.npts + tile - 1) / tile;
  // 8 CTAs per resident slot: with exactly one per slot the warps that finish first leave their slots empty (ncu: 18 of 24
  // warps active on average) — the hardware scheduler balances the tail: order 64 46.3 -> 49.0 % of FP64 peak, order 200
  // 63.3 -> 65.6 %, gradients 64.0 -> 68.4 % / 56.2 -> 59.2 % (FASTCHEB_PB_OVERSUB: experiments)
  int over = 8;
  if (const char* t = getenv("FASTCHEB_PB_OVERSUB")) over = std::max(1, std::min(16, atoi(t)));
  const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)sms * std::max(1, bps) * over));
  kern<<<grid, threads, 0, st>>>(args);
  note_launch();
  return cudaGetLastError();
}

template <typename R, int AMAX, int EC, bool JAC>
cudaError_t launch(const Pb1dParams& prm, const void* h_D, const void* h_D2, size_t d_count, int sms, cudaStream_t st) {
  // points per thread: the T_lo(z) table is AMAX registers per point, the inner sums P * EC (twice that with the
  // derivative series of the gradient kernels) — 2 points unless the sums of a gradient kernel would not fit
  // (Float32: half the registers per point, and the FP32 pipe is fast enough for the kernel to be issue-bound — twice
  // the points per thread halve the non-FMA instructions per point)
  constexpr int EF = EC * (JAC ? 2 : 1);  // inner sums per point and row
  // (Float64 value kernel, 16-entry tables, 4 components: 2 points spill ~1 KB per thread)
  constexpr int P0 = ((JAC && EC > 1) || (sizeof(R) == 8 && AMAX == 16 && EC == 4)) ? 1 : 2;
  constexpr int P = P0 * ((sizeof(R) == 4 && EF <= 2) ? 2 : 1);
  if constexpr (P != P0) {
    static const bool small = getenv("FASTCHEB_PB_SMALLP") != nullptr;  // experiments
    if (small) return launch_p<R, AMAX, EC, P0, JAC>(prm, h_D, h_D2, d_count, sms, st);
  }
  return launch_p<R, AMAX, EC, P, JAC>(prm, h_D, h_D2, d_count, sms, st);
}

template <typename R, int AMAX, bool JAC>
cudaError_t launch_ec(int ec, const Pb1dParams& prm, const void* h_D, const void* h_D2, size_t d_count, int sms, cudaStream_t st) {
  switch (ec) {
    case 1: return launch<R, AMAX, 1, JAC>(prm, h_D, h_D2, d_count, sms, st);
    case 2: return launch<R, AMAX, 2, JAC>(prm, h_D, h_D2, d_count, sms, st);
    case 3: return launch<R, AMAX, 3, JAC>(prm, h_D, h_D2, d_count, sms, st);
    case 4: return launch<R, AMAX, 4, JAC>(prm, h_D, h_D2, d_count, sms, st);
  }
  return cudaErrorInvalidValue;
}

template <typename R>
cudaError_t launch_any(int amax, int ec, bool jac, const Pb1dParams& prm, const void* h_D, const void* h_D2, size_t d_count, int sms,
                       cudaStream_t st) {
  if (amax == 8)
    return jac ? launch_ec<R, 8, true>(ec, prm, h_D, h_D2, d_count, sms, st) : launch_ec<R, 8, false>(ec, prm, h_D, h_D2, d_count, sms, st);
  if (amax == 16)
    return jac ? launch_ec<R, 16, true>(ec, prm, h_D, h_D2, d_count, sms, st) : launch_ec<R, 16, false>(ec, prm, h_D, h_D2, d_count, sms, st);
  return cudaErrorInvalidValue;
}

}  // namespace

// (the kernel divides by the span with the reciprocal + exact-remainder correction unconditionally: spans that fail
// SpanDiv's applicability test — mantissa all ones, extreme exponents — stay on the Clenshaw kernels)
bool pb1d_shape_ok(const fastcheb_poly* p) {
  return p->ndim == 1 && p->dims[0] >= kPbMinN && p->dims[0] <= kPbMaxN && p->E <= 4 && span_fast(p, 0);
}

int pb1d_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, int64_t v_stride, void* out_J,
              int64_t J_stride, bool jac, long long* bad_dev, long long idx_base, cudaStream_t st, int mode,
              const void* tan) {
  const PbPlan& pb = p->pb;
  if (jac && pb.h_D2.empty())
    return fail(FASTCHEB_EUNSUPPORTED, "product-basis gradients: two coefficient tables of this shape exceed the kernel parameter space");
  Pb1dParams prm;
  memset(&prm, 0, sizeof prm);
  prm.pts = pts;
  prm.out_v = out_v;
  prm.out_J = out_J;
  prm.bad = bad_dev;
  prm.npts = npts;
  prm.idx_base = idx_base;
  prm.v_stride = v_stride;
  prm.J_stride = J_stride;
  prm.e0 = 0;
  prm.Etot = p->E;
  prm.M = pb.M;
  prm.last_len = (int)p->dims[0] - pb.AMAX * (pb.M - 1);
  prm.lb = p->lb[0];
  prm.ub = p->ub[0];
  prm.tan = tan;
  prm.mode = mode;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  prm.vec = (al16(pts) ? 1 : 0) |
            ((mode == 0 && v_stride == p->E && al16(out_v) && (!jac || (J_stride == p->E && al16(out_J)))) ? 2 : 0);
  const std::vector<unsigned char>& hD = (jac && !pb.h_DJ.empty()) ? pb.h_DJ : pb.h_D;
  const size_t d_count = hD.size() / real_size(p->dtype);
  const cudaError_t e = p->dtype == FASTCHEB_F64
                            ? launch_any<double>(pb.AMAX, p->E, jac, prm, hD.data(), pb.h_D2.empty() ? nullptr : pb.h_D2.data(), d_count, p->di->sms, st)
                            : launch_any<float>(pb.AMAX, p->E, jac, prm, hD.data(), pb.h_D2.empty() ? nullptr : pb.h_D2.data(), d_count, p->di->sms, st);
  if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "product-basis kernel launch failed: %s", cudaGetErrorString(e));
  return FASTCHEB_OK;
}

// Build the product-basis coefficients (and, unless `forced`, run the self-check that lets AUTO use them).
// Thread-safe and idempotent; the handle's observable state (coefficients, box) never changes.
int pb1d_plan(const fastcheb_poly* cp, bool forced) {
  fastcheb_poly* p = const_cast<fastcheb_poly*>(cp);
  PbPlan& pb = p->pb;
  std::lock_guard<std::mutex> lk(pb.mu);
  if (pb.state.load() != kPbNone && (pb.checked || forced)) return FASTCHEB_OK;
  if (!pb1d_shape_ok(p)) {
    pb.state.store(kPbRejected);
    pb.checked = true;
    return FASTCHEB_OK;
  }
  DeviceGuard dg(p->device);
  const int n = (int)p->dims[0], E = p->E;
  const size_t rs = real_size(p->dtype);
  cudaStream_t st = cudaStreamPerThread;
  if (pb.h_D.empty()) {
    int A = n <= 128 ? 8 : 16;
    if (const char* t = getenv("FASTCHEB_PB_A")) {  // experiments: force the table size
      const int v = atoi(t);
      if (v == 8 || v == 16) A = v;
    }
    const int M = (n + A - 1) / A;
    std::vector<unsigned char> raw((size_t)n * E * rs);
    FC_CUDA(cudaMemcpyAsync(raw.data(), p->d_coefs, raw.size(), cudaMemcpyDeviceToHost, st));
    FC_CUDA(cudaStreamSynchronize(st));
    auto coef = [&](int k, int e) -> long double {
      if (k >= n) return 0.0L;
      return p->dtype == FASTCHEB_F64 ? (long double)((const double*)raw.data())[(size_t)k * E + e]
                                      : (long double)((const float*)raw.data())[(size_t)k * E + e];
    };
    // D[hi][lo], hi < M rows (the runtime extent), lo < A (the compile-time modulus L of the kernel)
    std::vector<long double> D((size_t)(M + 1) * A);
    // kernel layout (eval_pb1d.cuh): full groups of G rows interleaved [group][lo][j][e] from offset 0, the M % G
    // remaining rows [t][lo][e] from the fixed offset pb_tail_base
    // (the scalar gradient kernel groups 2 rows, the value kernel 4: pb_group — it gets its own copy of D, h_DJ)
    auto count_of = [&](int G) {
      const int ngroups = M / G;
      return ngroups * G == M ? (size_t)M * A * E : (size_t)pb_tail_base(A, E) + (size_t)(M - ngroups * G) * A * E;
    };
    const int Gv = pb_group(E, false), Gj = pb_group(E, true);
    std::vector<unsigned char> out(count_of(Gv) * rs, 0), outj, out2;
    // gradient kernels: the same table for the derivative series p' = sum_k c'_k T_k, c'_{k-1} = c'_{k+1} + 2 k c_k
    // (c'_0 halved) — skipped where two tables do not fit the kernel parameter space (pb_cap2)
    const bool grad = count_of(Gj) <= (size_t)(rs == 8 ? pb_cap2<double>(A, E) : pb_cap2<float>(A, E));
    if (grad) out2.assign(count_of(Gj) * rs, 0);
    if (grad && Gj != Gv) outj.assign(count_of(Gj) * rs, 0);
    std::vector<long double> cd((size_t)n + 2);
    auto table = [&](auto&& cf, int e, int G, std::vector<unsigned char>& dst) {
      const int ngroups = M / G;
      std::fill(D.begin(), D.end(), 0.0L);
      for (int hi = M - 1; hi >= 0; --hi)
        for (int lo = 0; lo < A; ++lo) {
          const long double c = cf(A * hi + lo);
          long double v;
          if (lo == 0)
            v = c;
          else if (hi == 0)
            v = c - 0.5L * D[(size_t)1 * A + (A - lo)];
          else
            v = 2.0L * c - D[(size_t)(hi + 1) * A + (A - lo)];
          D[(size_t)hi * A + lo] = v;
        }
      for (int hi = 0; hi < M; ++hi)
        for (int lo = 0; lo < A; ++lo) {
          const size_t o = hi < ngroups * G
                               ? (((size_t)(hi / G) * A + lo) * G + (size_t)(hi % G)) * E + e
                               : (size_t)pb_tail_base(A, E) + ((size_t)(hi - ngroups * G) * A + lo) * E + e;
          if (p->dtype == FASTCHEB_F64)
            ((double*)dst.data())[o] = (double)D[(size_t)hi * A + lo];
          else
            ((float*)dst.data())[o] = (float)D[(size_t)hi * A + lo];
        }
    };
    for (int e = 0; e < E; ++e) {
      long double sum_abs = 0, sum_k2 = 0;
      table([&](int k) { return coef(k, e); }, e, Gv, out);
      if (!outj.empty()) table([&](int k) { return coef(k, e); }, e, Gj, outj);
      if (grad) {
        std::fill(cd.begin(), cd.end(), 0.0L);
        for (int k = n - 1; k >= 1; --k) cd[k - 1] = cd[k + 1] + 2.0L * (long double)k * coef(k, e);
        cd[0] *= 0.5L;
        table([&](int k) { return k < n ? cd[k] : 0.0L; }, e, Gj, out2);
      }
      for (int k = 0; k < n; ++k) {
        sum_abs += fabsl(coef(k, e));
        sum_k2 += fabsl(coef(k, e)) * (long double)k * (long double)k;
      }
      pb.sum_abs[e] = (double)sum_abs;
      pb.sum_k2[e] = (double)sum_k2;
    }
    pb.h_D2 = out2;
    pb.h_DJ = outj;
    pb.h_D = out;  // the launches pass the coefficients in the kernel parameters
    pb.AMAX = A;
    pb.M = M;
    pb.state.store(kPbBuilt);
  }
  if (forced || pb.checked) return FASTCHEB_OK;

  // Self-check against the bit-exact Clenshaw kernel on probe points: a uniform grid, Chebyshev-clustered points
  // and points at distances 2^-1 .. 2^-53 of the span from both end points (where both algorithms are least accurate).
  // AUTO uses this path only if values stay within a quarter of the north_star tolerance 16 eps sum|c| of the
  // Clenshaw kernel (derivatives: of the differentiated-series bound 16 eps sum|c_k| k^2 * 2/(ub-lb)).
  const int NP = 3 * 2048;
  std::vector<double> x((size_t)NP);
  const double lb = p->lb[0], ub = p->ub[0], span = ub - lb;
  for (int i = 0; i < 2048; ++i) {
    x[i] = lb + span * (double)i / 2047.0;
    x[2048 + i] = lb + span * 0.5 * (1.0 + cos(M_PI * (i + 0.5) / 2048.0));
    const double d = ldexp(1.0, -(1 + (52 * (i / 2)) / 1024)) * (1.0 + (i % 7) / 7.0) * 0.5;
    x[4096 + i] = (i & 1) ? ub - span * d : lb + span * d;
  }
  std::vector<unsigned char> hx((size_t)NP * rs);
  for (int i = 0; i < NP; ++i) {
    const double xi = std::min(std::max(x[i], lb), ub);
    if (p->dtype == FASTCHEB_F64) {
      ((double*)hx.data())[i] = xi;
    } else {
      float f = (float)xi;
      if ((double)f < lb) f = nextafterf(f, INFINITY);
      if ((double)f > ub) f = nextafterf(f, -INFINITY);
      ((float*)hx.data())[i] = f;
    }
  }
  const size_t vb = (size_t)NP * E * rs;
  unsigned char* dbuf = nullptr;
  FC_CUDA(pool_alloc((void**)&dbuf, hx.size() + 4 * vb, st));
  unsigned char *d_x = dbuf + 4 * vb, *d_v0 = dbuf, *d_J0 = dbuf + vb, *d_v1 = dbuf + 2 * vb, *d_J1 = dbuf + 3 * vb;
  std::vector<unsigned char> h((size_t)4 * vb);
  int rc = FASTCHEB_OK;
  const bool has_grad = !pb.h_D2.empty();
  cudaError_t ce = cudaMemcpyAsync(d_x, hx.data(), hx.size(), cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) rc = eval_clenshaw_device(p, d_x, NP, d_v0, E, d_J0, E, true, nullptr, 0, st, 0, nullptr);
  if (ce == cudaSuccess && rc == FASTCHEB_OK) rc = pb1d_eval(p, d_x, NP, d_v1, E, d_J1, E, has_grad, nullptr, 0, st, 0, nullptr);
  if (ce == cudaSuccess && rc == FASTCHEB_OK) ce = cudaMemcpyAsync(h.data(), dbuf, 4 * vb, cudaMemcpyDeviceToHost, st);
  const cudaError_t e2 = cudaStreamSynchronize(st);
  cudaFreeAsync(dbuf, st);
  if (rc != FASTCHEB_OK) return rc;
  if (ce != cudaSuccess || e2 != cudaSuccess)
    return fail(FASTCHEB_ECUDA, "product-basis self-check failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : e2));
  auto get = [&](size_t off, size_t i) -> double {
    return p->dtype == FASTCHEB_F64 ? ((const double*)(h.data() + off))[i] : (double)((const float*)(h.data() + off))[i];
  };
  // the tolerances are PER COMPONENT (16 eps sum_k |c_{k,e}| for the values of component e): the worst ratio over the
  // components decides (one sum over all components would be E times too lenient for a vector-valued polynomial)
  const double eps = p->dtype == FASTCHEB_F64 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
  bool finite = true;
  double rv = 0, rJ = 0;
  for (int e = 0; e < E; ++e) {
    double dv = 0, dJ = 0;
    for (size_t i = e; i < (size_t)NP * E; i += E) {
      const double a = fabs(get(2 * vb, i) - get(0, i)), b = has_grad ? fabs(get(3 * vb, i) - get(vb, i)) : 0.0;
      finite = finite && std::isfinite(a) && std::isfinite(b);
      dv = std::max(dv, a);
      dJ = std::max(dJ, b);
    }
    const double tv = 16 * eps * pb.sum_abs[e], tJ = 16 * eps * pb.sum_k2[e] * 2.0 / span;
    rv = std::max(rv, tv > 0 ? dv / tv : (dv > 0 ? INFINITY : 0.0));
    rJ = std::max(rJ, tJ > 0 ? dJ / tJ : (dJ > 0 ? INFINITY : 0.0));
  }
  pb.ok_val = finite && rv <= 0.25;
  pb.ok_jac = has_grad && finite && pb.ok_val && rJ <= 0.25;
  pb.check_dv = rv;
  pb.check_dJ = rJ;
  pb.checked = true;
  pb.state.store((pb.ok_val || pb.ok_jac) ? kPbAccepted : kPbRejected);
  return FASTCHEB_OK;
}

}  // namespace fastcheb
