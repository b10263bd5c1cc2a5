// C ABI: the ChebPoly handle and batched evaluation (include/fastcheb.h).
#include "poly.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include "eval_clenshaw_launch.h"
#include "eval_dmma_launch.h"

using namespace fastcheb;

namespace {

// Julia layout -> planar-padded slabs of one component group.
//   dst[((t*ec + c)*Ms + inner)*n1p + k1] = src[(e0+c) + E*(k1 + n1*(inner + Ms*t))], zero for k1 >= n1
template <typename R>
__global__ void __launch_bounds__(256) relayout_pp_kernel(const R* __restrict__ src, R* __restrict__ dst,
                                                          long long total, int E, int e0, int ec, int n1, int n1p,
                                                          long long Ms) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int k1 = (int)(g % n1p);
    long long rem = g / n1p;
    const long long inner = rem % Ms;
    rem /= Ms;
    const int c = (int)(rem % ec);
    const long long t = rem / ec;
    dst[g] = k1 < n1 ? src[(e0 + c) + (long long)E * (k1 + (long long)n1 * (inner + Ms * t))] : R(0);
  }
}

int grid_for(long long total, int threads, int sms) {
  long long b = (total + threads - 1) / threads;
  return (int)std::max<long long>(1, std::min<long long>(b, (long long)sms * 16));
}

int plan_clenshaw(fastcheb_poly* p, cudaStream_t st) {
  const size_t rs = real_size(p->dtype);
  const int V = 16 / (int)rs;
  const int n1 = (int)p->dims[0];
  const int n1p = (n1 + V - 1) / V * V;
  const size_t budget = p->di->smem_optin - 1024 - 128;
  for (int e0 = 0; e0 < p->E;) {
    const int left = p->E - e0;
    int ec = left <= kClenshawMaxEC ? left : (left == 5 ? 3 : (left % 4 == 0 ? 4 : (left % 3 == 0 ? 3 : 4)));
    ClenshawGroup g;
    // the largest slab level that fits; when not even one dimension-1 line of `ec` components fits shared memory,
    // retry with fewer components per launch
    int s = 0;
    for (;; ec = ec > 2 ? 2 : 1) {
      for (s = p->ndim; s >= 1; --s) {
        double elems = (double)ec * n1p;
        for (int d = 1; d < s; ++d) elems *= (double)p->dims[d];
        long long T = 1;
        for (int d = s; d < p->ndim; ++d) T *= p->dims[d];
        const double bytes = elems * rs;
        if (T == 1 ? bytes <= (double)budget : 2 * bytes <= (double)budget) break;
      }
      if (s >= 1 || ec == 1) break;
    }
    g.e0 = e0;
    g.ec = ec;
    g.n1p = n1p;
    g.nvec1 = n1p / V;
    const bool global_line = s < 1;  // one line of ONE component exceeds shared memory (n1 > ~29 000 doubles)
    if (global_line) {
      if (p->ndim != 1) {
        // The handle is still built (coefficients, droptol, fits and roots work); only evaluation is refused.
        for (auto& q : p->groups) pool_free_all(&q.d_pp, 1);
        p->groups.clear();
        p->eval_ok = false;
        return FASTCHEB_OK;
      }
      s = 1;  // 1-D: the GLOBAL instantiation reads the line from global memory / L2
    }
    long long Ms = 1, T = 1;
    for (int d = 1; d < s; ++d) Ms *= p->dims[d];
    for (int d = s; d < p->ndim; ++d) T *= p->dims[d];
    if (T == 1) s = p->ndim;
    g.s = s;
    g.T = (int)T;
    g.compStride = (int)(Ms * n1p);
    g.slabElems = ec * g.compStride;
    const size_t slabBytes = (size_t)g.slabElems * rs;
    g.nstage = global_line ? 0 : (T == 1 ? 1 : (int)std::min<long long>(std::min<long long>(4, T), (long long)(budget / slabBytes)));
    g.smem = 128 + (size_t)g.nstage * slabBytes;
    long long str = 1;
    for (int d = 0; d < p->ndim; ++d) {
      g.stride[d] = d < s ? (int)str : 0;
      str *= (d == 0 ? n1p : p->dims[d]);
    }
    g.pp_bytes = slabBytes * (size_t)T;
    FC_CUDA(pool_alloc(&g.d_pp, g.pp_bytes, st));
    const long long total = (long long)(g.pp_bytes / rs);
    const int grid = grid_for(total, 256, p->di->sms);
    if (p->dtype == FASTCHEB_F64)
      relayout_pp_kernel<double><<<grid, 256, 0, st>>>((const double*)p->d_coefs, (double*)g.d_pp, total, p->E, e0, ec,
                                                       n1, n1p, Ms);
    else
      relayout_pp_kernel<float><<<grid, 256, 0, st>>>((const float*)p->d_coefs, (float*)g.d_pp, total, p->E, e0, ec, n1,
                                                      n1p, Ms);
    FC_CUDA(cudaGetLastError());
    note_launch();
    p->groups.push_back(g);
    e0 += ec;
  }
  return FASTCHEB_OK;
}

typedef cudaError_t (*ClenshawFn)(int, int, int, const EvalParams&, int, int, size_t, cudaStream_t, int*);
ClenshawFn clenshaw_fn(int dtype, bool jac) {
  if (dtype == FASTCHEB_F64) return jac ? clenshaw_f64_jac : clenshaw_f64_val;
  return jac ? clenshaw_f32_jac : clenshaw_f32_val;
}

int eval_clenshaw(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, int64_t v_stride, void* out_J,
                  int64_t J_stride, bool jac, long long* bad_dev, long long idx_base, cudaStream_t st, int mode,
                const void* tan) {
  if (p->ndim > kMaxKernelNdim)
    return fail(FASTCHEB_EUNSUPPORTED, "ndim = %d: evaluation kernels cover ndim <= %d", p->ndim, kMaxKernelNdim);
  if (!p->eval_ok)
    return fail(FASTCHEB_EUNSUPPORTED,
                "evaluation of an N-d polynomial whose dimension-1 line (%lld coefficients) exceeds shared memory is not "
                "covered; coefficients, droptol and re-interpolation of this handle work",
                (long long)p->dims[0]);
  ClenshawFn fn = clenshaw_fn(p->dtype, jac);
  for (const ClenshawGroup& g : p->groups) {
    const int P = clenshaw_pts(p->ndim, g.ec, jac);
    EvalParams prm;
    memset(&prm, 0, sizeof prm);
    prm.coefs = g.d_pp;
    prm.pts = pts;
    prm.out_v = out_v;
    prm.out_J = out_J;
    prm.bad = (&g == &p->groups[0]) ? bad_dev : nullptr;  // the domain check is per point, not per group
    prm.npts = npts;
    prm.idx_base = idx_base;
    prm.v_stride = v_stride;
    prm.J_stride = J_stride;
    prm.mode = mode;
    prm.tan = tan;
    prm.e0 = g.e0;
    prm.Etot = p->E;
    for (int d = 0; d < p->ndim; ++d) {
      prm.n[d] = (int)p->dims[d];
      prm.stride[d] = g.stride[d];
      prm.lb[d] = p->lb[d];
      prm.ub[d] = p->ub[d];
    }
    prm.nvec1 = g.nvec1;
    prm.s = g.s;
    prm.T = g.T;
    prm.compStride = g.compStride;
    prm.slabElems = g.slabElems;
    prm.nstage = g.nstage;
    // block size: the widest block the instantiation allows, narrowed for small batches so that the
    // points spread over all SMs
    const bool heavy = !jac && p->ndim * g.ec >= kClenshawHeavy;
    int threads = jac ? kClenshawMaxThreadsJac : (heavy ? kClenshawMaxThreadsHeavy : kClenshawMaxThreadsVal);
    const long long per_sm = (npts + (long long)p->di->sms * P - 1) / ((long long)p->di->sms * P);
    if (per_sm < threads) threads = (int)std::max<long long>(32, (per_sm + 31) / 32 * 32);
    int bps = 1;
    cudaError_t e = fn(kOpOccupancy, p->ndim, g.ec, prm, threads, 0, g.smem, st, &bps);
    if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "occupancy query failed: %s", cudaGetErrorString(e));
    if (bps < 1) return fail(FASTCHEB_ECUDA, "Clenshaw kernel does not fit an SM (smem %zu B)", g.smem);
    const long long tile = (long long)threads * P;
    const long long ntiles = (npts + tile - 1) / tile;
    static const int over = [] {  // CTAs per resident slot (FASTCHEB_CL_OVERSUB: experiments)
      const char* t = getenv("FASTCHEB_CL_OVERSUB");
      return t ? std::max(1, std::min(16, atoi(t))) : 1;
    }();
    const int grid = (int)std::min<long long>(ntiles, (long long)p->di->sms * bps * over);
    e = fn(kOpLaunch, p->ndim, g.ec, prm, threads, grid, g.smem, st, nullptr);
    if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "Clenshaw kernel launch failed: %s", cudaGetErrorString(e));
  }
  return FASTCHEB_OK;
}

// 1-D product basis: forced by fastcheb_set_algo, or (AUTO) accepted by the plan-time self-check; the plan is built on
// the first batch of at least kPbMinBatch points
bool use_pb(const fastcheb_poly* p, bool jac, int64_t npts) {
  const int algo = p->algo.load();
  if (algo == FASTCHEB_ALGO_PRODUCT) return p->pb.state.load() != kPbNone && p->pb.state.load() != kPbRejected;
  if (algo != FASTCHEB_ALGO_AUTO || !pb1d_shape_ok(p)) return false;
  int st = p->pb.state.load();
  if ((st == kPbNone || st == kPbBuilt) && npts >= kPbMinBatch) {
    if (pb1d_plan(p, false) != FASTCHEB_OK) return false;
    st = p->pb.state.load();
  }
  return st == kPbAccepted && (jac ? p->pb.ok_jac : p->pb.ok_val);
}

bool use_dmma(const fastcheb_poly* p, bool jac) {
  if (!p->dmma.ok) return false;
  if (p->algo == FASTCHEB_ALGO_CLENSHAW || p->algo == FASTCHEB_ALGO_PRODUCT) return false;
  (void)jac;
  if (p->algo == FASTCHEB_ALGO_AUTO) {
    // Measured on B200 (tools/bench_configs.py with BENCH_SHAPES): the MMA K extent is padded to 4 KS = 16, 32 or 64
    // slots, and the Clenshaw kernel runs at ~30 % of FP64 peak whatever the shape, so the tensor-core path only wins
    // when dimension 1 fills most of its slots and there are a few column tiles to amortise a pass over
    // (5x8, 9x9, 5x40, 9x9x9: Clenshaw 1.1-1.7x faster; 33x9: equal; 17x17, 17^3, 6^4, 9^4 and up: DMMA 1.1-2.9x).
    const int n1 = (int)p->dims[0];
    if (5 * (n1 - 1) < 3 * 4 * p->dmma.KS || p->dmma.M < 16) return p->ndim >= 4 && n1 >= 6;
  }
  return true;
}

}  // namespace

namespace fastcheb {

int eval_clenshaw_device(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, int64_t v_stride, void* out_J,
                         int64_t J_stride, bool jac, long long* bad_dev, long long idx_base, cudaStream_t st, int mode,
                         const void* tan) {
  return eval_clenshaw(p, pts, npts, out_v, v_stride, out_J, J_stride, jac, bad_dev, idx_base, st, mode, tan);
}

int eval_device(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, int64_t v_stride, void* out_J,
                int64_t J_stride, bool jac, long long* bad_dev, long long idx_base, cudaStream_t st, int mode,
                const void* tan) {
  if (npts == 0) return FASTCHEB_OK;
  if (use_pb(p, jac, npts))
    return pb1d_eval(p, pts, npts, out_v, v_stride, out_J, J_stride, jac, bad_dev, idx_base, st, mode, tan);
  if (lo1d_ok(p, pts, npts, out_v, v_stride, out_J, J_stride, jac, mode))  // n < 40, or not taken by the product basis
    return lo1d_eval(p, pts, npts, out_v, out_J, jac, bad_dev, idx_base, st);
  // Scalar 2-D polynomials with n1 <= 17: the low-order nest measured equal to the tensor-core path on 17 x 17 and 1.2-1.26x
  // faster on 17 x 33 and 13 x 60 (tools/probe_crossover.py; 33 x 33, 12^3, 6^4 and vector-valued shapes stay on the tensor
  // cores).  FASTCHEB_LOND_FIRST: experiments, the nest wherever it applies.
  static const bool lond_first = getenv("FASTCHEB_LOND_FIRST") != nullptr;
  const bool small2d = p->ndim == 2 && p->E == 1 && p->dims[0] <= 17;
  if ((lond_first || small2d) && lond_ok(p, pts, npts, out_v, v_stride, out_J, J_stride, jac, mode))
    return lond_eval(p, pts, npts, out_v, out_J, jac, bad_dev, idx_base, st);
  if (use_dmma(p, jac))
    return dmma_eval(p, pts, npts, out_v, v_stride, out_J, J_stride, jac, bad_dev, idx_base, st, mode, tan);
  if (lond_ok(p, pts, npts, out_v, v_stride, out_J, J_stride, jac, mode))
    return lond_eval(p, pts, npts, out_v, out_J, jac, bad_dev, idx_base, st);
  return eval_clenshaw(p, pts, npts, out_v, v_stride, out_J, J_stride, jac, bad_dev, idx_base, st, mode, tan);
}

int create_from_device(fastcheb_poly** out, int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp,
                       const void* d_src, const double* lb, const double* ub, int device, cudaStream_t st) {
  fastcheb_poly* p = new (std::nothrow) fastcheb_poly();
  if (!p) return fail(FASTCHEB_ENOMEM, "out of host memory");
  p->ndim = ndim;
  p->dtype = dtype;
  p->cplx = is_complex ? 1 : 0;
  p->K = ncomp;
  p->E = ncomp * (is_complex ? 2 : 1);
  p->device = device;
  p->di = device_info(device);
  p->nelems = 1;
  for (int d = 0; d < ndim; ++d) {
    p->dims[d] = dims[d];
    p->lb[d] = lb[d];
    p->ub[d] = ub[d];
    p->nelems *= dims[d];
  }
  p->coef_bytes = (size_t)p->nelems * p->E * real_size(dtype);
  int rc = FASTCHEB_OK;
  do {
    if (pool_alloc(&p->d_coefs, p->coef_bytes, st) != cudaSuccess) {
      cudaGetLastError();
      rc = fail(FASTCHEB_ENOMEM, "cudaMalloc of %zu coefficient bytes failed", p->coef_bytes);
      break;
    }
    if (cudaMemcpyAsync(p->d_coefs, d_src, p->coef_bytes, cudaMemcpyDefault, st) != cudaSuccess) {
      rc = fail(FASTCHEB_ECUDA, "coefficient upload failed: %s", cudaGetErrorString(cudaGetLastError()));
      break;
    }
    if (ndim <= kMaxKernelNdim) {
      if ((rc = plan_clenshaw(p, st)) != FASTCHEB_OK) break;
      if ((rc = dmma_plan(p, st)) != FASTCHEB_OK) break;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) {
      rc = fail(FASTCHEB_ECUDA, "handle creation failed: %s", cudaGetErrorString(cudaGetLastError()));
      break;
    }
  } while (0);
  if (rc != FASTCHEB_OK) {
    fastcheb_destroy(p);
    return rc;
  }
  *out = p;
  return FASTCHEB_OK;
}

}  // namespace fastcheb

// ------------------------------------------------------------------------------------------------
extern "C" {

int fastcheb_version(void) { return FASTCHEB_VERSION; }

int64_t fastcheb_launch_count(void) { return (int64_t)launch_count(); }

int fastcheb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

size_t fastcheb_last_error(char* buf, size_t n) {
  const std::string& m = err_state().msg;
  if (buf && n) {
    const size_t k = std::min(n - 1, m.size());
    memcpy(buf, m.data(), k);
    buf[k] = 0;
  }
  return m.size();
}

int fastcheb_error_index(int64_t* idx) {
  if (!idx) return fail(FASTCHEB_EINVAL, "null idx");
  *idx = err_state().index;
  return FASTCHEB_OK;
}

int fastcheb_create(fastcheb_poly** out, int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp,
                    const void* coefs, int coefs_mem, const double* lb, const double* ub, int device) {
  if (!out || !dims || !coefs || !lb || !ub) return fail(FASTCHEB_EINVAL, "null argument");
  *out = nullptr;
  if (ndim < 1 || ndim > FASTCHEB_MAX_NDIM)
    return fail(FASTCHEB_EINVAL, "ndim = %d outside 1..%d", ndim, FASTCHEB_MAX_NDIM);
  if (dtype != FASTCHEB_F32 && dtype != FASTCHEB_F64) return fail(FASTCHEB_EINVAL, "bad dtype %d", dtype);
  if (ncomp < 1) return fail(FASTCHEB_EINVAL, "ncomp = %d < 1", ncomp);
  if (coefs_mem != FASTCHEB_MEM_HOST && coefs_mem != FASTCHEB_MEM_DEVICE)
    return fail(FASTCHEB_EINVAL, "bad coefs_mem %d", coefs_mem);
  double total = 1;
  for (int d = 0; d < ndim; ++d) {
    if (dims[d] < 1) return fail(FASTCHEB_EINVAL, "dims[%d] = %lld < 1", d, (long long)dims[d]);
    total *= (double)dims[d];
  }
  if (total * ncomp * 2 > 2.0e9) return fail(FASTCHEB_EUNSUPPORTED, "coefficient tensor too large");
  if (!device_info(device)) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  if (!dg.ok()) return fail(FASTCHEB_ECUDA, "cudaSetDevice(%d) failed", device);
  return create_from_device(out, ndim, dims, dtype, is_complex, ncomp, coefs, lb, ub, device, 0);
}

int fastcheb_destroy(fastcheb_poly* p) {
  if (!p) return FASTCHEB_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != p->device) cudaSetDevice(p->device);
  std::vector<void*> bufs{p->d_coefs, p->dmma.d_frag};
  for (auto& g : p->groups) bufs.push_back(g.d_pp);
  pool_free_all(bufs.data(), (int)bufs.size());
  if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
  cudaGetLastError();
  delete p;
  return FASTCHEB_OK;
}

int fastcheb_ndim(const fastcheb_poly* p) { return p ? p->ndim : -1; }
int fastcheb_dims(const fastcheb_poly* p, int64_t* dims_out) {
  if (!p || !dims_out) return fail(FASTCHEB_EINVAL, "null argument");
  for (int d = 0; d < p->ndim; ++d) dims_out[d] = p->dims[d];
  return FASTCHEB_OK;
}
int fastcheb_dtype(const fastcheb_poly* p) { return p ? p->dtype : -1; }
int fastcheb_is_complex(const fastcheb_poly* p) { return p ? p->cplx : -1; }
int fastcheb_ncomp(const fastcheb_poly* p) { return p ? p->K : -1; }
int fastcheb_device(const fastcheb_poly* p) { return p ? p->device : -1; }

int fastcheb_get_coefs(const fastcheb_poly* p, void* dst, int dst_mem) {
  if (!p || !dst) return fail(FASTCHEB_EINVAL, "null argument");
  (void)dst_mem;
  DeviceGuard dg(p->device);
  FC_CUDA(cudaMemcpy(dst, p->d_coefs, p->coef_bytes, cudaMemcpyDefault));
  return FASTCHEB_OK;
}

int fastcheb_set_algo(fastcheb_poly* p, int algo) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  if ((algo == FASTCHEB_ALGO_DMMA || algo == FASTCHEB_ALGO_DMMA_RING) && !p->dmma.ok)
    return fail(FASTCHEB_EUNSUPPORTED, "the DMMA path needs ndim >= 2, n1 >= 5 and at least 8 coefficient columns");
  if (algo == FASTCHEB_ALGO_PRODUCT) {
    if (!pb1d_shape_ok(p))
      return fail(FASTCHEB_EUNSUPPORTED, "the product-basis path covers 1-D polynomials with %d <= n <= %d, at most 4 real components and a span ub - lb that "
                  "passes the reciprocal-division test (mantissa not all ones, exponent within +-500)",
                  kPbMinN, kPbMaxN);
    FC_TRY(pb1d_plan(p, true));
  } else if (algo != FASTCHEB_ALGO_AUTO && algo != FASTCHEB_ALGO_CLENSHAW && algo != FASTCHEB_ALGO_DMMA &&
             algo != FASTCHEB_ALGO_DMMA_RING) {
    return fail(FASTCHEB_EINVAL, "bad algo %d", algo);
  }
  p->algo.store(algo);
  return FASTCHEB_OK;
}

int fastcheb_get_algo(const fastcheb_poly* p, int jac) {
  if (!p) return -1;
  if (use_pb(p, jac != 0, kPbMinBatch)) return FASTCHEB_ALGO_PRODUCT;  // AUTO: builds and self-checks the plan if needed
  return use_dmma(p, jac != 0) ? FASTCHEB_ALGO_DMMA : FASTCHEB_ALGO_CLENSHAW;
}

int fastcheb_product_check(const fastcheb_poly* p, double* dv_over_tol, double* dJ_over_tol) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  if (!pb1d_shape_ok(p)) return fail(FASTCHEB_EUNSUPPORTED, "not a shape of the product-basis path");
  FC_TRY(pb1d_plan(p, false));
  if (dv_over_tol) *dv_over_tol = p->pb.check_dv;
  if (dJ_over_tol) *dJ_over_tol = p->pb.check_dJ;
  return FASTCHEB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// synchronous evaluation entry points
namespace {

struct OutSpec {
  // per point: `v_reals` reals at v, `J_reals` reals at J; strides in reals
  void* v;
  void* J;
  int64_t v_stride, J_stride;
  bool interleaved;  // v and J share one buffer (tuple layout)
};

int check_flag(long long flag, const fastcheb_poly* p) {
  if (flag == kNoBad) return FASTCHEB_OK;
  err_state().index = flag;
  return fail(FASTCHEB_EDOMAIN, "point %lld not in domain (eval.jl:64)", flag);
}

int eval_sync(const fastcheb_poly* p, const void* pts, int64_t npts, const OutSpec& o, bool jac, int mem,
              void* stream) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  if (npts < 0) return fail(FASTCHEB_EINVAL, "npts < 0");
  if (npts == 0) return FASTCHEB_OK;
  if (!pts || !o.v || (jac && !o.J)) return fail(FASTCHEB_EINVAL, "null buffer");
  if (mem != FASTCHEB_MEM_HOST && mem != FASTCHEB_MEM_DEVICE) return fail(FASTCHEB_EINVAL, "bad mem %d", mem);
  DeviceGuard dg(p->device);
  if (!dg.ok()) return fail(FASTCHEB_ECUDA, "cudaSetDevice(%d) failed", p->device);
  cudaStream_t st = (cudaStream_t)stream;
  FlagSlot fs;
  FC_TRY(flag_acquire(p->device, &fs));
  struct Rel {
    FlagSlot& f;
    ~Rel() { flag_release(f); }
  } rel{fs};
  const size_t rs = real_size(p->dtype);

  if (mem == FASTCHEB_MEM_DEVICE) {
    FC_CUDA(flag_prepare(fs, st));
    FC_TRY(eval_device(p, pts, npts, o.v, o.v_stride, o.J, o.J_stride, jac, fs.dev, 0, st));
    FC_CUDA(cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
    FC_CUDA(cudaStreamSynchronize(st));
    return (flag_observe(fs), check_flag(*fs.host, p));
  }

  // host buffers: chunked H2D -> kernel -> D2H pipeline on two cached internal streams
  const int N = p->ndim, E = p->E;
  const size_t in_pp = (size_t)N * rs;
  size_t out_pp;  // bytes per point of output, all buffers
  if (o.interleaved)
    out_pp = (size_t)o.v_stride * rs;
  else
    out_pp = (size_t)E * rs + (jac ? (size_t)E * N * rs : 0);
  // chunks of ~64 MB (points + results): small enough that the first upload and the last download are a few percent of
  // the call, large enough (>= 1 M points) that the persistent kernel stays efficient
  int64_t chunk = std::min<int64_t>(npts, std::max<int64_t>(1 << 16, (int64_t)((size_t(64) << 20) / (in_pp + out_pp))));
  const int nbuf = npts > chunk ? 2 : 1;
  const size_t in_bytes = (size_t)chunk * in_pp;
  const size_t v_pp = o.interleaved ? out_pp : (size_t)E * rs, J_pp = (!o.interleaved && jac) ? (size_t)E * N * rs : 0;
  const size_t v_bytes = (size_t)chunk * v_pp, J_bytes = (size_t)chunk * J_pp;
  const size_t in_al = (in_bytes + 255) / 256 * 256, v_al = (v_bytes + 255) / 256 * 256;

  // a call that fits one chunk (single points, small batches) runs on the caller's stream with plain copies: setting up
  // the pipeline costs more than such a call itself
  if (nbuf == 1) {
    void* dbuf = ws_acquire(p->device, in_al + v_al + J_bytes);
    if (!dbuf) return fail(FASTCHEB_ENOMEM, "device staging buffer of %zu bytes", in_al + v_al + J_bytes);
    char* d_in = (char*)dbuf;
    char* d_v = d_in + in_al;
    char* d_J = d_v + v_al;
    int rc = FASTCHEB_OK;
    cudaError_t ce = flag_prepare(fs, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_in, pts, (size_t)npts * in_pp, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) {
      if (o.interleaved)
        rc = eval_device(p, d_in, npts, d_v, o.v_stride, (char*)d_v + (size_t)E * rs, o.J_stride, jac, fs.dev, 0, st);
      else
        rc = eval_device(p, d_in, npts, d_v, E, d_J, (int64_t)E * N, jac, fs.dev, 0, st);
    }
    if (rc == FASTCHEB_OK && ce == cudaSuccess) ce = cudaMemcpyAsync(o.v, d_v, (size_t)npts * v_pp, cudaMemcpyDeviceToHost, st);
    if (rc == FASTCHEB_OK && ce == cudaSuccess && J_pp)
      ce = cudaMemcpyAsync(o.J, d_J, (size_t)npts * J_pp, cudaMemcpyDeviceToHost, st);
    if (rc == FASTCHEB_OK && ce == cudaSuccess) ce = cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st);
    const cudaError_t e2 = cudaStreamSynchronize(st);  // one synchronisation for the whole call
    if (ce == cudaSuccess) ce = e2;
    ws_release(p->device, dbuf);
    if (rc != FASTCHEB_OK) return rc;
    if (ce != cudaSuccess) return fail(FASTCHEB_ECUDA, "host-staged evaluation failed: %s", cudaGetErrorString(ce));
    return (flag_observe(fs), check_flag(*fs.host, p));
  }

  // Large batches.  PAGEABLE caller memory (a Julia Vector, a numpy array) would make every cudaMemcpyAsync block the
  // host (H2D until staged, D2H until complete), which serialises upload, kernel and download; it is therefore copied
  // by this thread (par_memcpy) through cached PINNED staging buffers, so that the DMA transfers of chunk i overlap the
  // kernel of chunk i-1 exactly as they do for pinned caller memory.
  const bool stage_in = !is_pinned_host(pts);
  const bool stage_out = !is_pinned_host(o.v) || (J_pp && !is_pinned_host(o.J));
  StreamPair sp;
  FC_TRY(streams_acquire(p->device, &sp));
  void* dbuf[2] = {nullptr, nullptr};
  char* hin[2] = {nullptr, nullptr};
  char* hout[2] = {nullptr, nullptr};
  auto cleanup = [&]() {
    for (int i = 0; i < 2; ++i) {
      cudaStreamSynchronize(sp.s[i]);
      if (dbuf[i]) ws_release(p->device, dbuf[i]);
      pin_release(hin[i]);
      pin_release(hout[i]);
    }
    streams_release(sp);
  };
  for (int i = 0; i < 2; ++i) {
    dbuf[i] = ws_acquire(p->device, in_al + v_al + J_bytes);
    if (stage_in) hin[i] = (char*)pin_acquire(in_bytes);
    if (stage_out) hout[i] = (char*)pin_acquire(v_al + J_bytes);
    if (!dbuf[i] || (stage_in && !hin[i]) || (stage_out && !hout[i])) {
      cleanup();
      return fail(FASTCHEB_ENOMEM, "staging buffers of %zu bytes", in_al + v_al + J_bytes);
    }
  }
  int rc = FASTCHEB_OK;
  cudaError_t ce = cudaSuccess;
  if (fs.dirty) {
    ce = set_flag_async(fs.dev, kNoBad, sp.s[0]);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(sp.s[0]);  // ordered before the second stream's work
  }
  struct Pending {
    int64_t first = 0, m = 0;
    bool live = false;
  } pend[2];
  // results of the chunk that last used buffer b: wait for its download, then (pageable output) copy them out
  auto retire = [&](int b) {
    if (!pend[b].live) return;
    pend[b].live = false;
    const cudaError_t e = cudaEventSynchronize(sp.ev[b]);
    if (ce == cudaSuccess) ce = e;
    if (stage_out && e == cudaSuccess) {
      par_memcpy((char*)o.v + (size_t)pend[b].first * v_pp, hout[b], (size_t)pend[b].m * v_pp);
      if (J_pp) par_memcpy((char*)o.J + (size_t)pend[b].first * J_pp, hout[b] + v_al, (size_t)pend[b].m * J_pp);
    }
  };
  int64_t done = 0;
  for (int it = 0; ce == cudaSuccess && rc == FASTCHEB_OK && done < npts; ++it) {
    const int b = it & 1;
    const int64_t m = std::min<int64_t>(chunk, npts - done);
    retire(b);
    if (ce != cudaSuccess) break;
    char* d_in = (char*)dbuf[b];
    char* d_v = d_in + in_al;
    char* d_J = d_v + v_al;
    const char* src = (const char*)pts + (size_t)done * in_pp;
    if (stage_in) {
      par_memcpy(hin[b], src, (size_t)m * in_pp);
      src = hin[b];
    }
    ce = cudaMemcpyAsync(d_in, src, (size_t)m * in_pp, cudaMemcpyHostToDevice, sp.s[b]);
    if (ce != cudaSuccess) break;
    if (o.interleaved)
      rc = eval_device(p, d_in, m, d_v, o.v_stride, (char*)d_v + (size_t)E * rs, o.J_stride, jac, fs.dev, done, sp.s[b]);
    else
      rc = eval_device(p, d_in, m, d_v, E, d_J, (int64_t)E * N, jac, fs.dev, done, sp.s[b]);
    if (rc != FASTCHEB_OK) break;
    ce = cudaMemcpyAsync(stage_out ? hout[b] : (char*)o.v + (size_t)done * v_pp, d_v, (size_t)m * v_pp, cudaMemcpyDeviceToHost, sp.s[b]);
    if (ce == cudaSuccess && J_pp)
      ce = cudaMemcpyAsync(stage_out ? hout[b] + v_al : (char*)o.J + (size_t)done * J_pp, d_J, (size_t)m * J_pp,
                           cudaMemcpyDeviceToHost, sp.s[b]);
    if (ce == cudaSuccess) ce = cudaEventRecord(sp.ev[b], sp.s[b]);
    pend[b].first = done;
    pend[b].m = m;
    pend[b].live = ce == cudaSuccess;
    done += m;
  }
  // drain in issue order
  const int last = (int)(((npts + chunk - 1) / chunk - 1) & 1);
  retire(last ^ 1);
  retire(last);
  if (ce == cudaSuccess && rc == FASTCHEB_OK) ce = cudaMemcpy(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost);
  cleanup();
  if (rc != FASTCHEB_OK) return rc;
  if (ce != cudaSuccess) return fail(FASTCHEB_ECUDA, "host-staged evaluation failed: %s", cudaGetErrorString(ce));
  return (flag_observe(fs), check_flag(*fs.host, p));
}

}  // namespace

extern "C" {

int fastcheb_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out, int mem, void* stream) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  OutSpec o{out, nullptr, p->E, 0, false};
  return eval_sync(p, pts, npts, o, false, mem, stream);
}

int fastcheb_eval_jac(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, void* out_J, int mem,
                      void* stream) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  OutSpec o{out_v, out_J, p->E, (int64_t)p->E * p->ndim, false};
  return eval_sync(p, pts, npts, o, true, mem, stream);
}

int fastcheb_eval_grad(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, void* out_g, int mem,
                       void* stream) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  if (p->K != 1)
    return fail(FASTCHEB_EDIMMISMATCH, "chebgradient needs a scalar-valued polynomial, got %d components (eval.jl:170)",
                p->K);
  return fastcheb_eval_jac(p, pts, npts, out_v, out_g, mem, stream);
}

int fastcheb_eval_jac_tuple(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_vJ, int mem,
                            void* stream) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  const int64_t rec = (int64_t)p->E * (1 + p->ndim);
  OutSpec o{out_vJ, out_vJ ? (char*)out_vJ + (size_t)p->E * real_size(p->dtype) : nullptr, rec, rec, true};
  return eval_sync(p, pts, npts, o, true, mem, stream);
}

int fastcheb_eval_async(const fastcheb_poly* p, const void* pts_dev, int64_t npts, void* out_dev, int64_t* status_dev,
                        void* stream) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  if (npts < 0 || (npts > 0 && (!pts_dev || !out_dev))) return fail(FASTCHEB_EINVAL, "bad buffer");
  DeviceGuard dg(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (status_dev) FC_CUDA(set_flag_async((long long*)status_dev, kNoBad, st));
  return eval_device(p, pts_dev, npts, out_dev, p->E, nullptr, 0, false, (long long*)status_dev, 0, st);
}

int fastcheb_eval_jac_async(const fastcheb_poly* p, const void* pts_dev, int64_t npts, void* out_v_dev, void* out_J_dev,
                            int64_t* status_dev, void* stream) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  if (npts < 0 || (npts > 0 && (!pts_dev || !out_v_dev || !out_J_dev))) return fail(FASTCHEB_EINVAL, "bad buffer");
  DeviceGuard dg(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (status_dev) FC_CUDA(set_flag_async((long long*)status_dev, kNoBad, st));
  return eval_device(p, pts_dev, npts, out_v_dev, p->E, out_J_dev, (int64_t)p->E * p->ndim, true,
                     (long long*)status_dev, 0, st);
}

}  // extern "C"

// ---- batched AD rules (chainrules.jl:4-39): the Jacobian is contracted inside the kernel, never written to HBM ----
namespace {
// mode 1: pullback  xbar[p, :] = real(J_p' * dy[p, :])   (tan: npts x E reals, out_t: npts x ndim)
// mode 2: pushforward ydot[p, :] = J_p * dx[p, :]         (tan: npts x ndim,    out_t: npts x E reals)
int eval_tangent(const fastcheb_poly* p, const void* pts, int64_t npts, const void* tan, void* out_v, void* out_t,
                 int mode, int mem, void* stream) {
  if (!p) return fail(FASTCHEB_EINVAL, "null handle");
  if (npts < 0) return fail(FASTCHEB_EINVAL, "npts < 0");
  if (npts == 0) return FASTCHEB_OK;
  if (!pts || !tan || !out_v || !out_t) return fail(FASTCHEB_EINVAL, "null buffer");
  if (mem != FASTCHEB_MEM_HOST && mem != FASTCHEB_MEM_DEVICE) return fail(FASTCHEB_EINVAL, "bad mem %d", mem);
  DeviceGuard dg(p->device);
  if (!dg.ok()) return fail(FASTCHEB_ECUDA, "cudaSetDevice(%d) failed", p->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rs = real_size(p->dtype);
  const int N = p->ndim, E = p->E;
  const size_t tan_pp = (size_t)(mode == 1 ? E : N) * rs, out_pp = (size_t)(mode == 1 ? N : E) * rs;
  FlagSlot fs;
  FC_TRY(flag_acquire(p->device, &fs));
  struct Rel {
    FlagSlot& f;
    ~Rel() { flag_release(f); }
  } rel{fs};
  FC_CUDA(flag_prepare(fs, st));
  if (mem == FASTCHEB_MEM_DEVICE) {
    FC_TRY(eval_device(p, pts, npts, out_v, E, out_t, 0, true, fs.dev, 0, st, mode, tan));
  } else {
    const size_t per = (size_t)N * rs + tan_pp + (size_t)E * rs + out_pp;
    const int64_t chunk = std::min<int64_t>(npts, std::max<int64_t>(1 << 16, (int64_t)((size_t(256) << 20) / per)));
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t b_in = al((size_t)chunk * N * rs), b_tan = al((size_t)chunk * tan_pp), b_v = al((size_t)chunk * E * rs);
    char* buf = (char*)ws_acquire(p->device, b_in + b_tan + b_v + al((size_t)chunk * out_pp));
    if (!buf) return fail(FASTCHEB_ENOMEM, "device staging buffer");
    int rc = FASTCHEB_OK;
    cudaError_t ce = cudaSuccess;
    for (int64_t done = 0; done < npts && rc == FASTCHEB_OK && ce == cudaSuccess; done += chunk) {
      const int64_t m = std::min<int64_t>(chunk, npts - done);
      char *d_in = buf, *d_tan = buf + b_in, *d_v = d_tan + b_tan, *d_t = d_v + b_v;
      ce = cudaMemcpyAsync(d_in, (const char*)pts + (size_t)done * N * rs, (size_t)m * N * rs, cudaMemcpyHostToDevice, st);
      if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(d_tan, (const char*)tan + (size_t)done * tan_pp, (size_t)m * tan_pp, cudaMemcpyHostToDevice, st);
      if (ce != cudaSuccess) break;
      rc = eval_device(p, d_in, m, d_v, E, d_t, 0, true, fs.dev, done, st, mode, d_tan);
      if (rc != FASTCHEB_OK) break;
      ce = cudaMemcpyAsync((char*)out_v + (size_t)done * E * rs, d_v, (size_t)m * E * rs, cudaMemcpyDeviceToHost, st);
      if (ce == cudaSuccess)
        ce = cudaMemcpyAsync((char*)out_t + (size_t)done * out_pp, d_t, (size_t)m * out_pp, cudaMemcpyDeviceToHost, st);
    }
    cudaStreamSynchronize(st);
    ws_release(p->device, buf);
    if (rc != FASTCHEB_OK) return rc;
    if (ce != cudaSuccess) return fail(FASTCHEB_ECUDA, "host-staged tangent evaluation failed: %s", cudaGetErrorString(ce));
  }
  FC_CUDA(cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaStreamSynchronize(st));
  return (flag_observe(fs), check_flag(*fs.host, p));
}
}  // namespace

extern "C" {

int fastcheb_eval_vjp(const fastcheb_poly* p, const void* pts, int64_t npts, const void* dy, void* out_v, void* out_xbar,
                      int mem, void* stream) {
  return eval_tangent(p, pts, npts, dy, out_v, out_xbar, 1, mem, stream);
}

int fastcheb_eval_jvp(const fastcheb_poly* p, const void* pts, int64_t npts, const void* dx, void* out_v, void* out_ydot,
                      int mem, void* stream) {
  return eval_tangent(p, pts, npts, dx, out_v, out_ydot, 2, mem, stream);
}

}  // extern "C"
