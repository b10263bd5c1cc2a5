// C ABI: the coefficient fit — chebcoefs / droptol / chebinterp (include/fastcheb.h).
//   chebcoefs   /root/reference/src/interp.jl:39-71
//   droptol     interp.jl:77-93
//   chebinterp  interp.jl:95-99, 140-141 (default tol: :102)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>
#include "fit_direct.cuh"
#include "fit_dmma_launch.h"
#include "poly.h"

using namespace fastcheb;

namespace {

constexpr long long kFitDmmaMinLines = 256;  // below this a batch is latency-bound: generic kernel

std::mutex g_tab_mu;
std::map<std::pair<int, int>, double2*> g_ctab;  // (device, n) -> cos(pi r/(n-1)) as (hi, lo), r in [0, 2(n-1))

// exactly reduced cosine table: cos(pi/2) = 0, cos(pi) = -1 and the symmetries hold bit for bit
int cos_table(int dev, int n, const double2** out) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto key = std::make_pair(dev, n);
  auto it = g_ctab.find(key);
  if (it != g_ctab.end()) {
    *out = it->second;
    return FASTCHEB_OK;
  }
  const int m = n - 1;
  std::vector<double2> h((size_t)2 * m);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int r = 0; r < 2 * m; ++r) {
    const int rr = r <= m ? r : 2 * m - r;
    long double v;
    if (2 * rr == m)
      v = 0;
    else if (2 * rr < m)
      v = cosl(pi * (long double)rr / (long double)m);
    else
      v = -cosl(pi * (long double)(m - rr) / (long double)m);
    h[r].x = (double)v;
    h[r].y = (double)(v - (long double)h[r].x);
  }
  // adaptive order doubling and roots() visit a handful of sizes; a caller that sweeps many distinct n must not
  // accumulate tables for ever: beyond 64 cached tables the cache is dropped (after the device has drained)
  if (g_ctab.size() >= 64) {
    cudaDeviceSynchronize();
    for (auto& kv : g_ctab) cudaFree(kv.second);
    g_ctab.clear();
  }
  double2* d = nullptr;
  FC_CUDA(cudaMalloc(&d, h.size() * sizeof(double2)));
  FC_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(double2), cudaMemcpyHostToDevice));
  g_ctab[key] = d;
  *out = d;
  return FASTCHEB_OK;
}

int grid_for(long long total, int threads, int sms) {
  long long b = (total + threads - 1) / threads;
  return (int)std::max<long long>(1, std::min<long long>(b, (long long)sms * 32));
}

template <typename R>
int fit_passes(const DeviceInfo* di, int dev, int ndim, const int64_t* dims, int E, int64_t howmany, const R* vals,
               R* coefs, cudaStream_t st) {
  long long per_fit = E;
  for (int d = 0; d < ndim; ++d) per_fit *= dims[d];
  const long long total = per_fit * howmany;
  int npass = 0;
  for (int d = 0; d < ndim; ++d) npass += dims[d] > 1;
  if (npass == 0) {
    if ((const void*)vals != (void*)coefs)
      FC_CUDA(cudaMemcpyAsync(coefs, vals, (size_t)total * sizeof(R), cudaMemcpyDeviceToDevice, st));
    return FASTCHEB_OK;
  }
  // ping-pong so that the last pass lands in `coefs`; `vals` is never written
  // scratch comes from the device's stream-ordered pool and is freed in stream order: the call never blocks the host
  R* tmp = nullptr;
  R* tmp2 = nullptr;
  if (npass > 1 || (const void*)vals == (void*)coefs) {
    if (pool_alloc((void**)&tmp, (size_t)total * sizeof(R), st) != cudaSuccess) {
      cudaGetLastError();
      return fail(FASTCHEB_ENOMEM, "fit scratch of %lld bytes", total * (long long)sizeof(R));
    }
  }
  const R* src = vals;
  int rc = FASTCHEB_OK;
  int pass = 0;
  long long inner = E;
  for (int d = 0; d < ndim && rc == FASTCHEB_OK; ++d) {
    const int n = (int)dims[d];
    if (n > 1) {
      // destination: coefs if the remaining pass count (including this one) is odd, else tmp
      const int remaining = npass - pass;
      R* dst = (remaining % 2 == 1) ? coefs : tmp;
      if ((const void*)dst == (const void*)src) {
        // only possible on pass 0 with vals == coefs and odd npass: move vals out of the way first
        if (pool_alloc((void**)&tmp2, (size_t)total * sizeof(R), st) != cudaSuccess) {
          cudaGetLastError();
          tmp2 = nullptr;
          rc = fail(FASTCHEB_ENOMEM, "fit scratch of %lld bytes", total * (long long)sizeof(R));
          break;
        }
        cudaMemcpyAsync(tmp2, src, (size_t)total * sizeof(R), cudaMemcpyDeviceToDevice, st);
        src = tmp2;
      }
      const double2* ctab = nullptr;
      if ((rc = cos_table(dev, n, &ctab)) != FASTCHEB_OK) break;
      // long lines in a small array: one warp per output element (see fit_direct.cuh)
      if (n >= 256 && total < (long long)di->sms * 2048)
        dct1_direct_warp_kernel<R><<<grid_for(total * 32, 256, di->sms), 256, 0, st>>>(src, dst, ctab, total, n, inner);
      else
        dct1_direct_kernel<R><<<grid_for(total, 256, di->sms), 256, 0, st>>>(src, dst, ctab, total, n, inner);
      if (cudaGetLastError() != cudaSuccess) {
        rc = fail(FASTCHEB_ECUDA, "DCT-I kernel launch failed");
        break;
      }
      note_launch();
      src = dst;
      ++pass;
    }
    inner *= n;
  }
  if (tmp) cudaFreeAsync(tmp, st);
  if (tmp2) cudaFreeAsync(tmp2, st);
  return rc;
}

template <typename R>
int fit_device_t(const DeviceInfo* di, int dev, int ndim, const int64_t* dims, int cplx, int K, int64_t howmany,
                 const R* vals, R* coefs, double* maxabs_dev, long long* bad_dev, long long idx_base, cudaStream_t st,
                 double* stats_dev = nullptr, const int* stats_hoff = nullptr, double stats_tol = 0.0, bool* stats_done = nullptr) {
  const int E = K * (cplx ? 2 : 1);
  long long elems = 1;
  for (int d = 0; d < ndim; ++d) elems *= dims[d];
  const int dt = sizeof(R) == 8 ? FASTCHEB_F64 : FASTCHEB_F32;
  if (maxabs_dev) FC_CUDA(cudaMemsetAsync(maxabs_dev, 0, (size_t)howmany * sizeof(double), st));
  // The bandwidth-bound path: many short 1-D lines go through the tensor-core transform, which also
  // does the finite check and (for real coefficients) the per-fit maximum on the way.
  long long done = 0;
  if (ndim == 1 && (long long)E * howmany >= kFitDmmaMinLines)
    FC_TRY(fit_dmma_1d(di, dev, (int)dims[0], E, howmany, dt, vals, coefs, bad_dev, idx_base, (maxabs_dev && !cplx) ? maxabs_dev : nullptr,
                       st, &done));
  if (done > 0 && maxabs_dev && cplx) {
    const long long nchunks = ((elems + 1023) / 1024) * done;
    coef_max_kernel<R><<<grid_for(nchunks * 32, 256, di->sms), 256, 0, st>>>(coefs, elems, done, K, cplx, maxabs_dev);
    FC_CUDA(cudaGetLastError());
    note_launch();
  }
  const long long rest = howmany - done;
  if (rest == 0) return FASTCHEB_OK;
  const long long off = done * elems * E;  // reals handled by the tensor-core path
  const R* rvals = vals + off;
  R* rcoefs = coefs + off;
  // N-d arrays and longer lines: the resident tensor-core kernel (whole array in shared memory, one HBM round trip;
  // finite check and per-array maximum fused)
  bool nd_done = false, nd_max = false;
  FC_TRY(fit_nd(di, dev, ndim, dims, K, cplx, rest, dt, rvals, rcoefs, bad_dev, idx_base + off, maxabs_dev ? maxabs_dev + done : nullptr,
                st, &nd_done, &nd_max, stats_dev, stats_hoff, stats_tol, stats_done));
  if (nd_done) {
    if (nd_max && maxabs_dev) {
      const long long nchunks = ((elems + 1023) / 1024) * rest;
      coef_max_kernel<R><<<grid_for(nchunks * 32, 256, di->sms), 256, 0, st>>>(rcoefs, elems, rest, K, cplx, maxabs_dev + done);
      FC_CUDA(cudaGetLastError());
      note_launch();
    }
    return FASTCHEB_OK;
  }
  if (bad_dev) {
    nonfinite_kernel<R><<<grid_for(rest * elems * E, 256, di->sms), 256, 0, st>>>(rvals, rest * elems * E, idx_base + off, bad_dev);  // interp.jl:40
    FC_CUDA(cudaGetLastError());
    note_launch();
  }
  FC_TRY(fit_passes<R>(di, dev, ndim, dims, E, rest, rvals, rcoefs, st));
  if (maxabs_dev) {
    const long long nchunks = ((elems + 1023) / 1024) * rest;
    coef_max_kernel<R><<<grid_for(nchunks * 32, 256, di->sms), 256, 0, st>>>(rcoefs, elems, rest, K, cplx, maxabs_dev + done);
    FC_CUDA(cudaGetLastError());
    note_launch();
  }
  return FASTCHEB_OK;
}

int fit_device(const DeviceInfo* di, int dev, int ndim, const int64_t* dims, int dtype, int cplx, int K, int64_t howmany,
               const void* vals, void* coefs, double* maxabs_dev, long long* bad_dev, cudaStream_t st,
               long long idx_base = 0, double* stats_dev = nullptr, const int* stats_hoff = nullptr, double stats_tol = 0.0,
               bool* stats_done = nullptr) {
  if (stats_done) *stats_done = false;
  if (dtype == FASTCHEB_F64)
    return fit_device_t<double>(di, dev, ndim, dims, cplx, K, howmany, (const double*)vals, (double*)coefs, maxabs_dev,
                                bad_dev, idx_base, st, stats_dev, stats_hoff, stats_tol, stats_done);
  return fit_device_t<float>(di, dev, ndim, dims, cplx, K, howmany, (const float*)vals, (float*)coefs, maxabs_dev,
                             bad_dev, idx_base, st, stats_dev, stats_hoff, stats_tol, stats_done);
}

int check_args(int ndim, const int64_t* dims, int dtype, int ncomp, int mem) {
  if (!dims) return fail(FASTCHEB_EINVAL, "null dims");
  if (ndim < 1 || ndim > FASTCHEB_MAX_NDIM) return fail(FASTCHEB_EINVAL, "ndim = %d outside 1..%d", ndim, FASTCHEB_MAX_NDIM);
  if (dtype != FASTCHEB_F32 && dtype != FASTCHEB_F64) return fail(FASTCHEB_EINVAL, "bad dtype %d", dtype);
  if (ncomp < 1) return fail(FASTCHEB_EINVAL, "ncomp = %d < 1", ncomp);
  if (mem != FASTCHEB_MEM_HOST && mem != FASTCHEB_MEM_DEVICE) return fail(FASTCHEB_EINVAL, "bad mem %d", mem);
  for (int d = 0; d < ndim; ++d)
    if (dims[d] < 1 || dims[d] > (1 << 30)) return fail(FASTCHEB_EINVAL, "dims[%d] = %lld", d, (long long)dims[d]);
  return FASTCHEB_OK;
}

// stats[0] = max (init 0), stats[1] = min (init +inf), hyperplane maxima (init 0)
__global__ void __launch_bounds__(256) norm_init_kernel(double* __restrict__ stats, int nd) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += gridDim.x * blockDim.x) stats[i] = i == 1 ? INFINITY : 0.0;
}

// droptol's shape decision from device coefficients (interp.jl:77-93), in two halves so that a caller can put other
// work (and its own status read-back) between the enqueue and the single host synchronisation.
struct DroptolJob {
  NormParams np;
  size_t nd = 0;
  double* d_stats = nullptr;  // stream-ordered pool allocation
  double* h = nullptr;        // PINNED host mirror (cached): a pageable target would make the read-back a blocking staged copy
  std::vector<double> h_fallback;
  ~DroptolJob() { pin_release(h); }
};

int droptol_enqueue(DroptolJob& job, const DeviceInfo* di, int ndim, const int64_t* dims, int dtype, int cplx, int K,
                    const void* d_coefs, cudaStream_t st) {
  long long elems = 1;
  int hsum = 0;
  memset(&job.np, 0, sizeof job.np);
  job.np.ndim = ndim;
  for (int d = 0; d < ndim; ++d) {
    elems *= dims[d];
    job.np.n[d] = (int)dims[d];
    job.np.hoff[d] = hsum;
    hsum += (int)dims[d];
  }
  job.nd = 2 + (size_t)hsum;
  if (!job.h && !(job.h = (double*)pin_acquire(job.nd * sizeof(double)))) {
    job.h_fallback.resize(job.nd);
  }
  if (!job.d_stats && pool_alloc((void**)&job.d_stats, job.nd * sizeof(double), st) != cudaSuccess) {
    cudaGetLastError();
    job.d_stats = nullptr;
    return fail(FASTCHEB_ENOMEM, "droptol scratch");
  }
  if (!d_coefs) return FASTCHEB_OK;  // layout + scratch only: the fit kernel itself produces the statistics
  norm_init_kernel<<<(int)std::min<size_t>((job.nd + 255) / 256, 64), 256, 0, st>>>(job.d_stats, (int)job.nd);
  const int grid = grid_for(elems, 256, di->sms);
  if (dtype == FASTCHEB_F64)
    coef_norm_kernel<double><<<grid, 256, 0, st>>>((const double*)d_coefs, elems, 1, K, cplx, job.d_stats, job.d_stats + 2, job.np);
  else
    coef_norm_kernel<float><<<grid, 256, 0, st>>>((const float*)d_coefs, elems, 1, K, cplx, job.d_stats, job.d_stats + 2, job.np);
  note_launch(2);
  if (cudaGetLastError() != cudaSuccess) {
    cudaFreeAsync(job.d_stats, st);
    job.d_stats = nullptr;
    return fail(FASTCHEB_ECUDA, "droptol reduction launch failed");
  }
  return FASTCHEB_OK;
}

// blocks until the statistics are on the host (the one synchronisation of droptol), then decides the shape
int droptol_finish(DroptolJob& job, int ndim, const int64_t* dims, double tol, int64_t* newdims, double* maxabs_out,
                   cudaStream_t st) {
  double* h = job.h ? job.h : job.h_fallback.data();
  cudaError_t e = cudaMemcpyAsync(h, job.d_stats, job.nd * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFreeAsync(job.d_stats, st);
  job.d_stats = nullptr;
  if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "droptol reduction failed: %s", cudaGetErrorString(e));
  for (int d = 0; d < ndim; ++d) newdims[d] = dims[d];
  const double mx = h[0], mn = h[1];
  if (maxabs_out) *maxabs_out = mx;
  const double abstol = mx * tol;        // interp.jl:78
  if (mn >= abstol) return FASTCHEB_OK;  // interp.jl:79: nothing to drop
  for (int d = 0; d < ndim; ++d) {       // interp.jl:83-91
    int64_t n = dims[d];
    while (n > 1 && !(h[2 + job.np.hoff[d] + (n - 1)] >= abstol)) --n;
    newdims[d] = n;
  }
  return FASTCHEB_OK;
}

int droptol_device(const DeviceInfo* di, int dev, int ndim, const int64_t* dims, int dtype, int cplx, int K,
                   const void* d_coefs, double tol, int64_t* newdims, double* maxabs_out, cudaStream_t st) {
  (void)dev;
  DroptolJob job;
  FC_TRY(droptol_enqueue(job, di, ndim, dims, dtype, cplx, K, d_coefs, st));
  return droptol_finish(job, ndim, dims, tol, newdims, maxabs_out, st);
}

// points of a tensor-product grid, AoS: pts[g*N + d] = axis_d[index of g along d] (g column-major over the grid)
struct GridParams {
  int ndim;
  int n[kMaxNdim];
  int off[kMaxNdim];
};
template <typename R>
__global__ void __launch_bounds__(256) grid_points_kernel(const double* __restrict__ axes, R* __restrict__ pts,
                                                          long long total, const GridParams gp) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    long long rem = g;
    for (int d = 0; d < gp.ndim; ++d) {
      const int k = (int)(rem % gp.n[d]);
      rem /= gp.n[d];
      pts[g * gp.ndim + d] = (R)axes[gp.off[d] + k];
    }
  }
}

// A cached device buffer that goes back to the cache when the call returns.  The stream is synchronised first: on
// the success paths it already is (cheap), on early error returns copies or kernels may still be using the buffer.
struct HostStage {
  int dev;
  cudaStream_t st;
  void* d = nullptr;
  HostStage(int dev_, cudaStream_t st_) : dev(dev_), st(st_) {}
  ~HostStage() {
    if (d) {
      cudaStreamSynchronize(st);
      ws_release(dev, d);
    }
  }
};

}  // namespace

extern "C" {

int fastcheb_fit(int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp, int64_t howmany, const void* vals,
                 void* coefs, double* maxabs, int mem, int device, void* stream) {
  FC_TRY(check_args(ndim, dims, dtype, ncomp, mem));
  if (howmany < 0) return fail(FASTCHEB_EINVAL, "howmany < 0");
  if (howmany == 0) return FASTCHEB_OK;
  if (!vals || !coefs) return fail(FASTCHEB_EINVAL, "null buffer");
  const DeviceInfo* di = device_info(device);
  if (!di) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  if (!dg.ok()) return fail(FASTCHEB_ECUDA, "cudaSetDevice(%d) failed", device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rs = real_size(dtype);
  const int E = ncomp * (is_complex ? 2 : 1);
  long long elems = 1;
  for (int d = 0; d < ndim; ++d) elems *= dims[d];
  const size_t bytes = (size_t)elems * E * howmany * rs;
  FlagSlot fs;
  FC_TRY(flag_acquire(device, &fs));
  struct Rel {
    FlagSlot& f;
    ~Rel() { flag_release(f); }
  } rel{fs};
  FC_CUDA(flag_prepare(fs, st));

  if (mem == FASTCHEB_MEM_DEVICE) {
    FC_TRY(fit_device(di, device, ndim, dims, dtype, is_complex, ncomp, howmany, vals, coefs, maxabs, fs.dev, st));
    FC_CUDA(cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
    FC_CUDA(cudaStreamSynchronize(st));
  } else {
    // Host buffers: the fits are independent, so a large batch is cut into chunks that flow through two
    // internal streams — the D2H copy of chunk i overlaps the H2D copy and the transform of chunk i+1 (PCIe is
    // full duplex).  A single fit is never split.
    const size_t fit_bytes = (size_t)elems * E * rs;
    int64_t chunk = howmany;
    if (bytes > (size_t(64) << 20) && howmany >= 4)
      chunk = std::max<int64_t>(1, std::min<int64_t>((howmany + 3) / 4, (int64_t)((size_t(128) << 20) / fit_bytes)));
    const int nbuf = howmany > chunk ? 2 : 1;
    // two internal streams only for multi-chunk batches: the flag initialisation above is then ordered before them by a
    // synchronisation; a single chunk (single fits, small batches) runs on the caller's stream with ONE synchronisation
    // at the end of the call
    if (nbuf > 1 && fs.dirty) FC_CUDA(cudaStreamSynchronize(st));
    cudaStream_t ss[2] = {nullptr, nullptr};
    void* dbuf[2] = {nullptr, nullptr};
    const size_t cb = (size_t)chunk * fit_bytes, cb_al = (cb + 255) / 256 * 256;
    const size_t mxb = maxabs ? (size_t)chunk * sizeof(double) : 0;
    StreamPair sp;  // the two chunk streams come from the per-device pool (no stream creation per call)
    sp.dev = -1;
    auto cleanup = [&]() {
      streams_release(sp);
      sp.dev = -1;
      for (int i = 0; i < 2; ++i)
        if (dbuf[i]) ws_release(device, dbuf[i]);
    };
    if (nbuf > 1) {
      FC_TRY(streams_acquire(device, &sp));
      ss[0] = sp.s[0];
      ss[1] = sp.s[1];
    }
    for (int i = 0; i < nbuf; ++i) {
      dbuf[i] = ws_acquire(device, 2 * cb_al + mxb);
      if (!dbuf[i]) {
        cleanup();
        return fail(FASTCHEB_ENOMEM, "device staging of %zu bytes", 2 * cb_al + mxb);
      }
    }
    int rc = FASTCHEB_OK;
    cudaError_t ce = cudaSuccess;
    int64_t done = 0;
    for (int it = 0; rc == FASTCHEB_OK && ce == cudaSuccess && done < howmany; ++it) {
      const int b = it % nbuf;
      cudaStream_t s2 = nbuf == 1 ? st : ss[b];
      const int64_t m = std::min<int64_t>(chunk, howmany - done);
      char* d_in = (char*)dbuf[b];
      char* d_out = d_in + cb_al;
      double* d_mx = maxabs ? (double*)(d_out + cb_al) : nullptr;
      ce = cudaMemcpyAsync(d_in, (const char*)vals + (size_t)done * fit_bytes, (size_t)m * fit_bytes, cudaMemcpyHostToDevice, s2);
      if (ce != cudaSuccess) break;
      rc = fit_device(di, device, ndim, dims, dtype, is_complex, ncomp, m, d_in, d_out, d_mx, fs.dev, s2,
                      (long long)done * elems * E);
      if (rc != FASTCHEB_OK) break;
      ce = cudaMemcpyAsync((char*)coefs + (size_t)done * fit_bytes, d_out, (size_t)m * fit_bytes, cudaMemcpyDeviceToHost, s2);
      if (ce == cudaSuccess && maxabs)
        ce = cudaMemcpyAsync(maxabs + done, d_mx, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, s2);
      done += m;
    }
    if (nbuf == 1 && ce == cudaSuccess && rc == FASTCHEB_OK)
      ce = cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st);  // one synchronisation for all
    for (int i = 0; i < nbuf; ++i) {
      cudaError_t e2 = cudaStreamSynchronize(nbuf == 1 ? st : ss[i]);
      if (ce == cudaSuccess) ce = e2;
    }
    cleanup();
    if (rc != FASTCHEB_OK) return rc;
    if (ce != cudaSuccess) return fail(FASTCHEB_ECUDA, "host-staged fit failed: %s", cudaGetErrorString(ce));
    if (nbuf > 1) {
      FC_CUDA(cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
      FC_CUDA(cudaStreamSynchronize(st));
    }
  }
  flag_observe(fs);
  if (*fs.host != kNoBad) {
    err_state().index = *fs.host;
    return fail(FASTCHEB_ENONFINITE, "non-finite interpolant value at flat index %lld (interp.jl:40)", *fs.host);
  }
  return FASTCHEB_OK;
}

int fastcheb_fit_async(int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp, int64_t howmany,
                       const void* vals_dev, void* coefs_dev, double* maxabs_dev, int64_t* status_dev, int device,
                       void* stream) {
  FC_TRY(check_args(ndim, dims, dtype, ncomp, FASTCHEB_MEM_DEVICE));
  if (howmany < 0) return fail(FASTCHEB_EINVAL, "howmany < 0");
  if (howmany == 0) return FASTCHEB_OK;
  if (!vals_dev || !coefs_dev) return fail(FASTCHEB_EINVAL, "null buffer");
  const DeviceInfo* di = device_info(device);
  if (!di) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  if (!dg.ok()) return fail(FASTCHEB_ECUDA, "cudaSetDevice(%d) failed", device);
  cudaStream_t st = (cudaStream_t)stream;
  if (status_dev) FC_CUDA(set_flag_async((long long*)status_dev, kNoBad, st));
  return fit_device(di, device, ndim, dims, dtype, is_complex, ncomp, howmany, vals_dev, coefs_dev, maxabs_dev,
                    (long long*)status_dev, st);
}

int fastcheb_droptol_shape(int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp, const void* coefs,
                           double tol, int64_t* newdims, int mem, int device, void* stream) {
  FC_TRY(check_args(ndim, dims, dtype, ncomp, mem));
  if (!coefs || !newdims) return fail(FASTCHEB_EINVAL, "null buffer");
  const DeviceInfo* di = device_info(device);
  if (!di) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  cudaStream_t st = (cudaStream_t)stream;
  const int E = ncomp * (is_complex ? 2 : 1);
  long long elems = 1;
  for (int d = 0; d < ndim; ++d) elems *= dims[d];
  const size_t bytes = (size_t)elems * E * real_size(dtype);
  HostStage in(device, st);
  const void* d_c = coefs;
  if (mem == FASTCHEB_MEM_HOST) {
    in.d = ws_acquire(device, bytes);
    if (!in.d) return fail(FASTCHEB_ENOMEM, "device staging of %zu bytes", bytes);
    FC_CUDA(cudaMemcpyAsync(in.d, coefs, bytes, cudaMemcpyHostToDevice, st));
    d_c = in.d;
  }
  return droptol_device(di, device, ndim, dims, dtype, is_complex, ncomp, d_c, tol, newdims, nullptr, st);
}

int fastcheb_interp(fastcheb_poly** out, int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp,
                    const void* vals, int mem, const double* lb, const double* ub, double tol, int device,
                    void* stream) {
  if (!out || !lb || !ub || !vals) return fail(FASTCHEB_EINVAL, "null argument");
  *out = nullptr;
  FC_TRY(check_args(ndim, dims, dtype, ncomp, mem));
  const DeviceInfo* di = device_info(device);
  if (!di) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rs = real_size(dtype);
  const int E = ncomp * (is_complex ? 2 : 1);
  long long elems = 1;
  for (int d = 0; d < ndim; ++d) elems *= dims[d];
  const size_t bytes = (size_t)elems * E * rs;
  if (tol < 0) tol = dtype == FASTCHEB_F64 ? DBL_EPSILON : (double)FLT_EPSILON;  // interp.jl:102

  FlagSlot fs;
  FC_TRY(flag_acquire(device, &fs));
  struct Rel {
    FlagSlot& f;
    ~Rel() { flag_release(f); }
  } rel{fs};
  FC_CUDA(flag_prepare(fs, st));
  HostStage in(device, st), co(device, st), blk(device, st);
  const void* d_vals = vals;
  if (mem == FASTCHEB_MEM_HOST) {
    in.d = ws_acquire(device, bytes);
    if (!in.d) return fail(FASTCHEB_ENOMEM, "device staging of %zu bytes", bytes);
    FC_CUDA(cudaMemcpyAsync(in.d, vals, bytes, cudaMemcpyHostToDevice, st));
    d_vals = in.d;
  }
  co.d = ws_acquire(device, bytes);
  if (!co.d) return fail(FASTCHEB_ENOMEM, "device scratch of %zu bytes", bytes);
  // one host synchronisation for both the finite flag (interp.jl:40) and droptol's statistics (interp.jl:77-91); arrays
  // that the resident fit kernel handles get the statistics from that kernel (no reduction launches)
  DroptolJob job;
  const bool trim = tol > 0;  // tol == 0: abstol = 0 and nothing is ever dropped (interp.jl:78-79)
  bool stats_done = false;
  if (trim) FC_TRY(droptol_enqueue(job, di, ndim, dims, dtype, is_complex, ncomp, nullptr, st));  // scratch + layout
  {
    const int rc = fit_device(di, device, ndim, dims, dtype, is_complex, ncomp, 1, d_vals, co.d, nullptr, fs.dev, st, 0,
                              trim ? job.d_stats : nullptr, trim ? job.np.hoff : nullptr, tol, &stats_done);
    if (rc != FASTCHEB_OK) {
      if (job.d_stats) cudaFreeAsync(job.d_stats, st);
      return rc;
    }
  }
  if (trim && !stats_done) FC_TRY(droptol_enqueue(job, di, ndim, dims, dtype, is_complex, ncomp, co.d, st));
  if (cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st) != cudaSuccess) {
    if (job.d_stats) cudaFreeAsync(job.d_stats, st);
    return fail(FASTCHEB_ECUDA, "status read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  int64_t newdims[FASTCHEB_MAX_NDIM];
  if (trim) {
    FC_TRY(droptol_finish(job, ndim, dims, tol, newdims, nullptr, st));
  } else {
    FC_CUDA(cudaStreamSynchronize(st));
    for (int d = 0; d < ndim; ++d) newdims[d] = dims[d];
  }
  flag_observe(fs);
  if (*fs.host != kNoBad) {
    err_state().index = *fs.host;
    return fail(FASTCHEB_ENONFINITE, "non-finite interpolant value at flat index %lld (interp.jl:40)", *fs.host);
  }
  bool same = true;
  long long new_elems = 1;
  for (int d = 0; d < ndim; ++d) {
    same = same && newdims[d] == dims[d];
    new_elems *= newdims[d];
  }
  const void* d_final = co.d;
  if (!same) {
    const size_t nb = (size_t)new_elems * E * rs;
    blk.d = ws_acquire(device, nb);
    if (!blk.d) return fail(FASTCHEB_ENOMEM, "device scratch of %zu bytes", nb);
    BlockParams bp;
    memset(&bp, 0, sizeof bp);
    bp.ndim = ndim;
    for (int d = 0; d < ndim; ++d) {
      bp.n_old[d] = (int)dims[d];
      bp.n_new[d] = (int)newdims[d];
    }
    const long long tot = new_elems * E;
    const int grid = grid_for(tot, 256, di->sms);
    if (dtype == FASTCHEB_F64)
      leading_block_kernel<double><<<grid, 256, 0, st>>>((const double*)co.d, (double*)blk.d, tot, E, bp);  // interp.jl:92
    else
      leading_block_kernel<float><<<grid, 256, 0, st>>>((const float*)co.d, (float*)blk.d, tot, E, bp);
    FC_CUDA(cudaGetLastError());
    d_final = blk.d;
  }
  return create_from_device(out, ndim, newdims, dtype, is_complex, ncomp, d_final, lb, ub, device, st);
}

// chebinterp(c, order, lb, ub; tol) for an interpolant that is itself a device-resident ChebPoly (the loop body of
// roots, roots.jl:105-107, and of any re-expansion on a sub-box): Chebyshev points of the new box (interp.jl:4-36)
// -> batched evaluation of `p` -> N-d DCT-I fit -> droptol -> new handle, without leaving the device.
int fastcheb_reinterp(fastcheb_poly** out, const fastcheb_poly* p, const int64_t* order, const double* lb,
                      const double* ub, double tol, void* stream) {
  if (!out || !p || !order || !lb || !ub) return fail(FASTCHEB_EINVAL, "null argument");
  *out = nullptr;
  const int N = p->ndim, dev = p->device;
  const DeviceInfo* di = device_info(dev);
  if (!di) return FASTCHEB_ECUDA;
  DeviceGuard dg(dev);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rs = real_size(p->dtype);
  int64_t dims[FASTCHEB_MAX_NDIM];
  GridParams gp;
  memset(&gp, 0, sizeof gp);
  gp.ndim = N;
  long long total = 1;
  int asum = 0;
  for (int d = 0; d < N; ++d) {
    if (order[d] < 0) return fail(FASTCHEB_EINVAL, "invalid negative order %lld (interp.jl:10)", (long long)order[d]);
    if (!(lb[d] >= p->lb[d] && ub[d] <= p->ub[d]))
      return fail(FASTCHEB_EDOMAIN, "new box [%g, %g] leaves the domain of the polynomial in dimension %d", lb[d], ub[d], d + 1);
    dims[d] = order[d] + 1;
    gp.n[d] = (int)dims[d];
    gp.off[d] = asum;
    asum += (int)dims[d];
    total *= dims[d];
    if (total > (1LL << 31)) return fail(FASTCHEB_EUNSUPPORTED, "grid too large");
  }
  // the grid coordinates, computed like interp.jl:4-7 in the working precision (index 0 is the upper bound; order 0
  // gives the midpoint), clamped into the source box so that rounding at the faces never leaves the domain
  std::vector<double> axes((size_t)asum);
  for (int d = 0; d < N; ++d)
    for (int i = 0; i < gp.n[d]; ++i) {
      double x;
      if (p->dtype == FASTCHEB_F64) {
        const double cth = order[d] == 0 ? cos(1.0 * M_PI / 2.0) : cos((double)i * M_PI / (double)order[d]);
        x = lb[d] + (1.0 + cth) * (ub[d] - lb[d]) * 0.5;
      } else {
        const float cth = order[d] == 0 ? cosf((float)(1.0 * M_PI / 2.0)) : cosf((float)((double)i * M_PI / (double)order[d]));
        x = (double)((float)lb[d] + (1.0f + cth) * ((float)ub[d] - (float)lb[d]) * 0.5f);
      }
      axes[gp.off[d] + i] = std::min(std::max(x, p->lb[d]), p->ub[d]);
    }
  HostStage ax(dev, st), pts(dev, st), vals(dev, st);
  ax.d = ws_acquire(dev, axes.size() * sizeof(double));
  pts.d = ws_acquire(dev, (size_t)total * N * rs);
  vals.d = ws_acquire(dev, (size_t)total * p->E * rs);
  if (!ax.d || !pts.d || !vals.d) return fail(FASTCHEB_ENOMEM, "re-interpolation scratch");
  // `axes` is pageable host memory: cudaMemcpyAsync stages it before returning, so the vector may go out of scope
  FC_CUDA(cudaMemcpyAsync(ax.d, axes.data(), axes.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  const int grid = grid_for(total, 256, di->sms);
  if (p->dtype == FASTCHEB_F64)
    grid_points_kernel<double><<<grid, 256, 0, st>>>((const double*)ax.d, (double*)pts.d, total, gp);
  else
    grid_points_kernel<float><<<grid, 256, 0, st>>>((const double*)ax.d, (float*)pts.d, total, gp);
  FC_CUDA(cudaGetLastError());
  note_launch();
  FC_TRY(eval_device(p, pts.d, total, vals.d, p->E, nullptr, 0, false, nullptr, 0, st));
  return fastcheb_interp(out, N, dims, p->dtype, p->cplx, p->K, vals.d, FASTCHEB_MEM_DEVICE, lb, ub, tol, dev, stream);
}

}  // extern "C"
