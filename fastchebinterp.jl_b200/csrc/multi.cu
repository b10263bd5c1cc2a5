// C ABI: multi-GPU evaluation and fits driven from ONE host process (include/fastcheb.h, "multi-GPU, one process").
//
// The reference's user writes `c.(ys)` in one process (/root/reference/src/eval.jl:63-70, FastChebInterp.jl:37).  A
// fastcheb_multi handle replicates the polynomial on several GPUs of the node and block-partitions every batch of
// points over them (points are independent, there is no exchange step):
//   * one persistent worker thread + one non-blocking stream per device;
//   * MEM_HOST: each worker runs the chunked H2D -> kernel -> D2H pipeline of fastcheb_eval on its block of the
//     caller's single point array, results land directly in the caller's single result array;
//   * MEM_DEVICE: points and results live on devices[0]; with peer access the kernels of the other devices load
//     their block of points from, and store their results straight into, devices[0]'s HBM over NVLink/NVSwitch
//     (the gather is fused into the evaluation kernel's stores, as in the multi-process path of peer.cu); without
//     peer access the blocks are staged with cudaMemcpyPeerAsync;
//   * the out-of-domain status is the minimum over devices of the first offending GLOBAL point index, so the error
//     is the one the serial broadcast would raise (eval.jl:64).
// Coefficients reach the other devices from the host pointer (one upload each) or, for device-resident
// coefficients, by peer copies over NVLink.  Batches of independent fits are split per device the same way
// (a single fit is never split).
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "poly.h"

using namespace fastcheb;

namespace {

// One worker thread per device: runs jobs submitted by the calling thread, keeps its device current.
class Worker {
 public:
  explicit Worker(int dev) : dev_(dev), th_([this] { loop(); }) {}
  ~Worker() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    cv_.notify_all();
    th_.join();
  }
  void submit(std::function<int()> job) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = std::move(job);
      has_job_ = true;
      done_ = false;
    }
    cv_.notify_all();
  }
  // status of the job, plus the worker thread's error message / index (the C ABI's error state is thread-local)
  int wait(std::string* msg, int64_t* index) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return done_; });
    if (msg) *msg = msg_;
    if (index) *index = index_;
    return rc_;
  }

 private:
  void loop() {
    cudaSetDevice(dev_);
    for (;;) {
      std::function<int()> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return has_job_ || quit_; });
        if (quit_ && !has_job_) return;
        job = std::move(job_);
        has_job_ = false;
      }
      err_state().msg.clear();
      err_state().index = -1;
      const int rc = job();
      {
        std::lock_guard<std::mutex> lk(mu_);
        rc_ = rc;
        msg_ = err_state().msg;
        index_ = err_state().index;
        done_ = true;
      }
      cv_.notify_all();
    }
  }
  int dev_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::function<int()> job_;
  bool has_job_ = false, done_ = true, quit_ = false;
  int rc_ = 0;
  std::string msg_;
  int64_t index_ = -1;
  std::thread th_;  // last: starts after the other members exist
};

struct Dev {
  int id = 0;
  fastcheb_poly* poly = nullptr;
  cudaStream_t st = nullptr;
  long long* status = nullptr;  // device flag of asynchronous launches
  bool peer_to_root = false;    // kernels on this device may dereference devices[0]'s memory
  Worker* worker = nullptr;
};

void block_range(int64_t n, int parts, int i, int64_t* lo, int64_t* hi) {
  const int64_t base = n / parts, rem = n % parts;
  *lo = i * base + std::min<int64_t>(i, rem);
  *hi = *lo + base + (i < rem ? 1 : 0);
}

bool enable_peer(int from, int to) {
  if (from == to) return true;
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, from, to) != cudaSuccess || !can) {
    cudaGetLastError();
    return false;
  }
  DeviceGuard dg(from);
  const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
  cudaGetLastError();
  return e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
}

// run job(i) on every device's worker; returns the first non-OK status in device order and copies that worker's error
// state to the calling thread.  EDOMAIN / ENONFINITE indices are made global by the jobs themselves.
template <typename F>
int run_all(const std::vector<Dev>& devs, F job) {
  for (size_t i = 0; i < devs.size(); ++i) devs[i].worker->submit([i, &job] { return job((int)i); });
  int rc = FASTCHEB_OK;
  int64_t best = -1;
  std::string best_msg;
  for (size_t i = 0; i < devs.size(); ++i) {
    std::string m;
    int64_t idx = -1;
    const int r = devs[i].worker->wait(&m, &idx);
    if (r == FASTCHEB_OK) continue;
    const bool indexed = r == FASTCHEB_EDOMAIN || r == FASTCHEB_ENONFINITE;
    // an indexed error (first offending element) wins with the smallest global index; any other failure wins outright
    if (rc == FASTCHEB_OK || (indexed && (rc == FASTCHEB_EDOMAIN || rc == FASTCHEB_ENONFINITE) && idx < best) ||
        (!indexed && (rc == FASTCHEB_EDOMAIN || rc == FASTCHEB_ENONFINITE))) {
      rc = r;
      best = idx;
      best_msg = m;
    }
  }
  if (rc != FASTCHEB_OK) {
    err_state().msg = best_msg;
    err_state().index = best;
  }
  return rc;
}

}  // namespace

struct fastcheb_multi {
  std::vector<Dev> devs;
  std::mutex mu;  // one batch at a time per handle (the workers hold one job each)
  int E = 1, ndim = 1, dtype = FASTCHEB_F64;
};

namespace {

int make_workers(std::vector<Dev>& devs, int ngpus, const int* devices) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return fail(FASTCHEB_ECUDA, "no usable CUDA device; libfastcheb_cuda has no CPU fallback");
  }
  if (ngpus < 1 || ngpus > count) return fail(FASTCHEB_EINVAL, "ngpus = %d, the node has %d device(s)", ngpus, count);
  devs.resize(ngpus);
  for (int i = 0; i < ngpus; ++i) {
    devs[i].id = devices ? devices[i] : i;
    if (!device_info(devs[i].id)) return FASTCHEB_ECUDA;
    for (int j = 0; j < i; ++j)
      if (devs[j].id == devs[i].id) return fail(FASTCHEB_EINVAL, "device %d listed twice", devs[i].id);
  }
  for (int i = 0; i < ngpus; ++i) {
    DeviceGuard dg(devs[i].id);
    FC_CUDA(cudaStreamCreateWithFlags(&devs[i].st, cudaStreamNonBlocking));
    FC_CUDA(cudaMalloc(&devs[i].status, sizeof(long long)));
    devs[i].peer_to_root = enable_peer(devs[i].id, devs[0].id);
    if (i > 0) enable_peer(devs[0].id, devs[i].id);
    devs[i].worker = new Worker(devs[i].id);
  }
  return FASTCHEB_OK;
}

void free_workers(std::vector<Dev>& devs) {
  for (Dev& d : devs) {
    delete d.worker;
    d.worker = nullptr;
    DeviceGuard dg(d.id);
    if (d.poly) fastcheb_destroy(d.poly);
    if (d.st) cudaStreamDestroy(d.st);
    if (d.status) cudaFree(d.status);
    d.poly = nullptr;
    d.st = nullptr;
    d.status = nullptr;
  }
  cudaGetLastError();
}

// which = 0: values, 1: values + Jacobian (separate arrays), 2: interleaved (v, J) tuple records
int multi_eval(fastcheb_multi* m, const void* pts, int64_t npts, void* out_v, void* out_J, int which, int mem) {
  if (!m) return fail(FASTCHEB_EINVAL, "null handle");
  if (npts < 0) return fail(FASTCHEB_EINVAL, "npts < 0");
  if (npts == 0) return FASTCHEB_OK;
  if (!pts || !out_v || (which == 1 && !out_J)) return fail(FASTCHEB_EINVAL, "null buffer");
  if (mem != FASTCHEB_MEM_HOST && mem != FASTCHEB_MEM_DEVICE) return fail(FASTCHEB_EINVAL, "bad mem %d", mem);
  std::lock_guard<std::mutex> lk(m->mu);
  const size_t rs = real_size(m->dtype);
  const int N = m->ndim, E = m->E, G = (int)m->devs.size();
  const size_t in_pp = (size_t)N * rs, v_pp = (size_t)E * rs * (which == 2 ? 1 + N : 1), J_pp = (size_t)E * N * rs;
  auto job = [&](int i) -> int {
    const Dev& d = m->devs[i];
    int64_t lo, hi;
    block_range(npts, G, i, &lo, &hi);
    const int64_t n = hi - lo;
    if (n == 0) return FASTCHEB_OK;
    const char* p_in = (const char*)pts + (size_t)lo * in_pp;
    char* p_v = (char*)out_v + (size_t)lo * v_pp;
    char* p_J = out_J ? (char*)out_J + (size_t)lo * J_pp : nullptr;
    int rc;
    if (mem == FASTCHEB_MEM_HOST) {
      // the chunked H2D -> kernel -> D2H pipeline of the single-GPU entry, on this device's block
      if (which == 0)
        rc = fastcheb_eval(d.poly, p_in, n, p_v, FASTCHEB_MEM_HOST, d.st);
      else if (which == 1)
        rc = fastcheb_eval_jac(d.poly, p_in, n, p_v, p_J, FASTCHEB_MEM_HOST, d.st);
      else
        rc = fastcheb_eval_jac_tuple(d.poly, p_in, n, p_v, FASTCHEB_MEM_HOST, d.st);
      if (rc == FASTCHEB_EDOMAIN) err_state().index += lo;
      return rc;
    }
    // device-resident batch on devices[0]
    void *l_in = nullptr, *l_v = nullptr, *l_J = nullptr;
    const bool direct = i == 0 || d.peer_to_root;
    const void* k_in = p_in;
    void* k_v = p_v;
    void* k_J = p_J;
    if (!direct) {  // no peer access: stage the block through this device's memory
      l_in = ws_acquire(d.id, (size_t)n * in_pp);
      l_v = ws_acquire(d.id, (size_t)n * v_pp);
      l_J = which == 1 ? ws_acquire(d.id, (size_t)n * J_pp) : nullptr;
      if (!l_in || !l_v || (which == 1 && !l_J)) {
        ws_release(d.id, l_in);
        ws_release(d.id, l_v);
        ws_release(d.id, l_J);
        return fail(FASTCHEB_ENOMEM, "device staging on device %d", d.id);
      }
      cudaMemcpyPeerAsync(l_in, d.id, p_in, m->devs[0].id, (size_t)n * in_pp, d.st);
      k_in = l_in;
      k_v = l_v;
      k_J = l_J;
    }
    if (which == 0) {
      rc = fastcheb_eval_async(d.poly, k_in, n, k_v, (int64_t*)d.status, d.st);
    } else if (which == 1) {
      rc = fastcheb_eval_jac_async(d.poly, k_in, n, k_v, k_J, (int64_t*)d.status, d.st);
    } else {
      if (d.status) set_flag_async(d.status, kNoBad, d.st);
      rc = eval_device(d.poly, k_in, n, k_v, (int64_t)E * (1 + N), (char*)k_v + (size_t)E * rs, (int64_t)E * (1 + N), true,
                       d.status, 0, d.st);
    }
    long long flag = kNoBad;
    cudaError_t ce = cudaSuccess;
    if (rc == FASTCHEB_OK) {
      if (!direct) {
        cudaMemcpyPeerAsync(p_v, m->devs[0].id, l_v, d.id, (size_t)n * v_pp, d.st);
        if (which == 1) cudaMemcpyPeerAsync(p_J, m->devs[0].id, l_J, d.id, (size_t)n * J_pp, d.st);
      }
      ce = cudaMemcpyAsync(&flag, d.status, sizeof flag, cudaMemcpyDeviceToHost, d.st);
    }
    const cudaError_t e2 = cudaStreamSynchronize(d.st);
    if (ce == cudaSuccess) ce = e2;
    if (!direct) {
      ws_release(d.id, l_in);
      ws_release(d.id, l_v);
      ws_release(d.id, l_J);
    }
    if (rc != FASTCHEB_OK) return rc;
    if (ce != cudaSuccess) return fail(FASTCHEB_ECUDA, "multi-GPU evaluation failed on device %d: %s", d.id, cudaGetErrorString(ce));
    if (flag != kNoBad) {
      err_state().index = flag + lo;
      return fail(FASTCHEB_EDOMAIN, "point %lld not in domain (eval.jl:64)", flag + lo);
    }
    return FASTCHEB_OK;
  };
  const int rc = run_all(m->devs, job);
  if (rc == FASTCHEB_EDOMAIN) {
    const int64_t idx = err_state().index;
    fail(FASTCHEB_EDOMAIN, "point %lld not in domain (eval.jl:64)", (long long)idx);
    err_state().index = idx;
  }
  return rc;
}

}  // namespace

extern "C" {

int fastcheb_multi_create(fastcheb_multi** out, int ngpus, const int* devices, int ndim, const int64_t* dims, int dtype,
                          int is_complex, int ncomp, const void* coefs, int coefs_mem, const double* lb, const double* ub) {
  if (!out || !dims || !coefs || !lb || !ub) return fail(FASTCHEB_EINVAL, "null argument");
  *out = nullptr;
  fastcheb_multi* m = new (std::nothrow) fastcheb_multi();
  if (!m) return fail(FASTCHEB_ENOMEM, "out of host memory");
  int rc = make_workers(m->devs, ngpus, devices);
  if (rc == FASTCHEB_OK) {
    // replicate the polynomial: host coefficients are uploaded once per device; device-resident coefficients (on
    // devices[0]) travel by peer copy over NVLink into a scratch buffer the new handle re-lays out from
    size_t bytes = (size_t)ncomp * (is_complex ? 2 : 1) * real_size(dtype);
    for (int d = 0; d < ndim && d < FASTCHEB_MAX_NDIM; ++d) bytes *= (size_t)std::max<int64_t>(dims[d], 1);
    const int root = m->devs[0].id;
    auto job = [&](int i) -> int {
      Dev& d = m->devs[i];
      if (coefs_mem == FASTCHEB_MEM_HOST || i == 0)
        return fastcheb_create(&d.poly, ndim, dims, dtype, is_complex, ncomp, coefs, coefs_mem, lb, ub, d.id);
      void* tmp = ws_acquire(d.id, bytes);
      if (!tmp) return fail(FASTCHEB_ENOMEM, "coefficient staging on device %d", d.id);
      int r = FASTCHEB_OK;
      if (cudaMemcpyPeer(tmp, d.id, coefs, root, bytes) != cudaSuccess)
        r = fail(FASTCHEB_ECUDA, "peer copy of the coefficients to device %d failed: %s", d.id,
                 cudaGetErrorString(cudaGetLastError()));
      if (r == FASTCHEB_OK)
        r = fastcheb_create(&d.poly, ndim, dims, dtype, is_complex, ncomp, tmp, FASTCHEB_MEM_DEVICE, lb, ub, d.id);
      ws_release(d.id, tmp);
      return r;
    };
    rc = run_all(m->devs, job);
  }
  if (rc != FASTCHEB_OK) {
    const ErrState keep = err_state();
    free_workers(m->devs);
    delete m;
    err_state() = keep;
    return rc;
  }
  m->E = m->devs[0].poly->E;
  m->ndim = ndim;
  m->dtype = dtype;
  *out = m;
  return FASTCHEB_OK;
}

int fastcheb_multi_destroy(fastcheb_multi* m) {
  if (!m) return FASTCHEB_OK;
  free_workers(m->devs);
  delete m;
  return FASTCHEB_OK;
}

int fastcheb_multi_ngpus(const fastcheb_multi* m) { return m ? (int)m->devs.size() : -1; }

int fastcheb_multi_device(const fastcheb_multi* m, int i) {
  return (m && i >= 0 && i < (int)m->devs.size()) ? m->devs[i].id : -1;
}

const fastcheb_poly* fastcheb_multi_poly(const fastcheb_multi* m, int i) {
  return (m && i >= 0 && i < (int)m->devs.size()) ? m->devs[i].poly : nullptr;
}

int fastcheb_multi_eval(fastcheb_multi* m, const void* pts, int64_t npts, void* out, int mem) {
  return multi_eval(m, pts, npts, out, nullptr, 0, mem);
}

int fastcheb_multi_eval_jac(fastcheb_multi* m, const void* pts, int64_t npts, void* out_v, void* out_J, int mem) {
  return multi_eval(m, pts, npts, out_v, out_J, 1, mem);
}

int fastcheb_multi_eval_grad(fastcheb_multi* m, const void* pts, int64_t npts, void* out_v, void* out_g, int mem) {
  if (!m) return fail(FASTCHEB_EINVAL, "null handle");
  if (m->devs[0].poly->K != 1)
    return fail(FASTCHEB_EDIMMISMATCH, "chebgradient needs a scalar-valued polynomial, got %d components (eval.jl:170)",
                m->devs[0].poly->K);
  return multi_eval(m, pts, npts, out_v, out_g, 1, mem);
}

int fastcheb_multi_eval_jac_tuple(fastcheb_multi* m, const void* pts, int64_t npts, void* out_vJ, int mem) {
  return multi_eval(m, pts, npts, out_vJ, nullptr, 2, mem);
}

// Batches of independent fits, block-partitioned by fit over the devices (no collective; a single fit
// stays on one GPU).  MEM_HOST: every device stages its block of the caller's arrays; MEM_DEVICE: vals / coefs /
// maxabs live on devices[0] and the blocks travel by peer copies over NVLink.
int fastcheb_multi_fit(int ngpus, const int* devices, int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp,
                       int64_t howmany, const void* vals, void* coefs, double* maxabs, int mem) {
  if (!dims || ndim < 1 || ndim > FASTCHEB_MAX_NDIM) return fail(FASTCHEB_EINVAL, "bad dims");
  if (howmany < 0) return fail(FASTCHEB_EINVAL, "howmany < 0");
  if (howmany == 0) return FASTCHEB_OK;
  if (!vals || !coefs) return fail(FASTCHEB_EINVAL, "null buffer");
  if (mem != FASTCHEB_MEM_HOST && mem != FASTCHEB_MEM_DEVICE) return fail(FASTCHEB_EINVAL, "bad mem %d", mem);
  std::vector<Dev> devs;
  int rc = make_workers(devs, ngpus, devices);
  if (rc == FASTCHEB_OK) {
    size_t fit_bytes = (size_t)ncomp * (is_complex ? 2 : 1) * real_size(dtype);
    size_t fit_reals = (size_t)ncomp * (is_complex ? 2 : 1);
    for (int d = 0; d < ndim; ++d) {
      fit_bytes *= (size_t)std::max<int64_t>(dims[d], 1);
      fit_reals *= (size_t)std::max<int64_t>(dims[d], 1);
    }
    const int G = ngpus, root = devs[0].id;
    auto job = [&](int i) -> int {
      const Dev& d = devs[i];
      int64_t lo, hi;
      block_range(howmany, G, i, &lo, &hi);
      const int64_t n = hi - lo;
      if (n == 0) return FASTCHEB_OK;
      const char* p_in = (const char*)vals + (size_t)lo * fit_bytes;
      char* p_out = (char*)coefs + (size_t)lo * fit_bytes;
      double* p_mx = maxabs ? maxabs + lo : nullptr;
      int r;
      if (mem == FASTCHEB_MEM_HOST || i == 0) {
        r = fastcheb_fit(ndim, dims, dtype, is_complex, ncomp, n, p_in, p_out, p_mx, mem, d.id, d.st);
      } else {
        char* buf = (char*)ws_acquire(d.id, 2 * (size_t)n * fit_bytes + (size_t)n * sizeof(double));
        if (!buf) return fail(FASTCHEB_ENOMEM, "fit staging on device %d", d.id);
        char* l_out = buf + (size_t)n * fit_bytes;
        double* l_mx = (double*)(l_out + (size_t)n * fit_bytes);
        cudaMemcpyPeerAsync(buf, d.id, p_in, root, (size_t)n * fit_bytes, d.st);
        r = fastcheb_fit(ndim, dims, dtype, is_complex, ncomp, n, buf, l_out, maxabs ? l_mx : nullptr, FASTCHEB_MEM_DEVICE,
                         d.id, d.st);
        if (r == FASTCHEB_OK) {
          cudaMemcpyPeerAsync(p_out, root, l_out, d.id, (size_t)n * fit_bytes, d.st);
          if (maxabs) cudaMemcpyPeerAsync(p_mx, root, l_mx, d.id, (size_t)n * sizeof(double), d.st);
        }
        if (cudaStreamSynchronize(d.st) != cudaSuccess && r == FASTCHEB_OK)
          r = fail(FASTCHEB_ECUDA, "multi-GPU fit failed on device %d: %s", d.id, cudaGetErrorString(cudaGetLastError()));
        ws_release(d.id, buf);
      }
      if (r == FASTCHEB_ENONFINITE) err_state().index += lo * (int64_t)fit_reals;
      return r;
    };
    rc = run_all(devs, job);
  }
  const ErrState keep = err_state();
  free_workers(devs);
  err_state() = keep;
  return rc;
}

}  // extern "C"
