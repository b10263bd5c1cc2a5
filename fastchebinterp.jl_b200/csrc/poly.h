// The device-resident ChebPoly handle (mirrors struct ChebPoly{N,T,Td}, FastChebInterp.jl:37-41).
#pragma once
#include <atomic>
#include <mutex>
#include <vector>
#include "host_util.h"

namespace fastcheb {

// Clenshaw-path plan for one group of <= 4 real components
struct ClenshawGroup {
  int e0 = 0, ec = 0;
  void* d_pp = nullptr;  // planar-padded slabs
  size_t pp_bytes = 0;
  int n1p = 0, nvec1 = 0;
  int s = 0, T = 1;
  int compStride = 0, slabElems = 0, nstage = 1;
  int stride[FASTCHEB_MAX_NDIM] = {0};
  size_t smem = 0;
};

// DMMA-path plan (FP64, ndim >= 2); see eval_dmma.cuh
struct DmmaGroup {
  int e0 = 0, ec = 0;
  size_t offset = 0;  // doubles into d_frag
};
struct DmmaPlan {
  bool ok = false;
  void* d_frag = nullptr;  // coefficients in B-fragment order, one stream per component group
  size_t frag_bytes = 0;
  int KS = 0;         // 4-wide k steps covering k1 = 1..n1-1
  int M = 0;          // columns per component = prod(n2..nN)
  int ntiles = 0;     // ceil(M / 8)
  int tabStride = 0;  // doubles per point row of the weight tables
  std::vector<DmmaGroup> groups;
};

// 1-D product-basis plan (eval_pb1d.cuh), built lazily on the first large batch and guarded by a self-check against
// the Clenshaw kernel; internally synchronised, never changes what the handle evaluates to beyond rounding
constexpr int kPbMinN = 40, kPbMaxN = 512;   // dimension-1 sizes the path covers
constexpr int64_t kPbMinBatch = 8192;        // AUTO plans it on the first batch of at least this many points
enum PbState { kPbNone = 0, kPbBuilt = 1, kPbAccepted = 2, kPbRejected = 3 };
struct PbPlan {
  std::mutex mu;
  std::atomic<int> state{kPbNone};
  bool checked = false, ok_val = false, ok_jac = false;
  std::vector<unsigned char> h_D;  // product-basis coefficients (kernel layout, eval_pb1d.cuh), passed as kernel parameters
  std::vector<unsigned char> h_D2;  // the same for the derivative series p' (gradient kernels); empty: gradients not covered
  std::vector<unsigned char> h_DJ;  // h_D in the gradient kernel's row grouping where that differs (scalar polynomials), else empty
  int AMAX = 0, M = 0;
  double sum_abs[8] = {0}, sum_k2[8] = {0};  // per real component e: sum_k |c_{k,e}|, sum_k |c_{k,e}| k^2 (the tolerances are per component)
  double check_dv = 0, check_dJ = 0;   // self-check result: worst difference over the tolerance
};

// low-order 1-D path (eval_lo1d.cuh): host copy of the zero-padded coefficients, made on the first large batch
struct LoPlan {
  std::mutex mu;
  std::atomic<int> ready{0};
  std::vector<unsigned char> h_c;  // 1-D: [17][E] reals; N-d: the tensor with dimension 1 zero-padded to 1 + 4 ngroups
  int ngroups = 0;
};

}  // namespace fastcheb

struct fastcheb_poly {
  int ndim = 0;
  int64_t dims[FASTCHEB_MAX_NDIM] = {0};
  int dtype = FASTCHEB_F64, cplx = 0, K = 1, E = 1;
  double lb[FASTCHEB_MAX_NDIM] = {0}, ub[FASTCHEB_MAX_NDIM] = {0};
  int device = 0;
  const fastcheb::DeviceInfo* di = nullptr;
  void* d_coefs = nullptr;  // Julia layout copy (fastcheb_get_coefs, re-layouts)
  size_t coef_bytes = 0;
  int64_t nelems = 0;  // prod dims
  std::vector<fastcheb::ClenshawGroup> groups;
  fastcheb::DmmaPlan dmma;
  std::atomic<int> algo{FASTCHEB_ALGO_AUTO};
  mutable fastcheb::PbPlan pb;
  mutable fastcheb::LoPlan lo;
  bool eval_ok = true;  // false: N-d polynomial whose dimension-1 line does not fit shared memory (plan_clenshaw)
};

namespace fastcheb {
// enqueue a batched evaluation on `st`; all pointers are device pointers on p->device.
int eval_device(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, int64_t v_stride, void* out_J,
                int64_t J_stride, bool jac, long long* bad_dev, long long idx_base, cudaStream_t st, int mode = 0,
                const void* tan = nullptr);
// the CUDA-core Clenshaw path regardless of the handle's algorithm selection (self-checks of the other paths)
int eval_clenshaw_device(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, int64_t v_stride, void* out_J,
                         int64_t J_stride, bool jac, long long* bad_dev, long long idx_base, cudaStream_t st, int mode,
                         const void* tan);
// 1-D product-basis path (eval_pb1d.cu)
bool pb1d_shape_ok(const fastcheb_poly* p);
int pb1d_plan(const fastcheb_poly* p, bool forced);
int pb1d_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, int64_t v_stride, void* out_J,
              int64_t J_stride, bool jac, long long* bad_dev, long long idx_base, cudaStream_t st, int mode, const void* tan);
// low-order 1-D path (eval_lo1d.cu): bit-identical to the Clenshaw kernel, vector loads / stores; lo1d_ok: shape, batch
// size, alignment and strides allow it
bool lo1d_ok(const fastcheb_poly* p, const void* pts, int64_t npts, const void* out_v, int64_t v_stride, const void* out_J,
             int64_t J_stride, bool jac, int mode);
int lo1d_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, void* out_J, bool jac, long long* bad_dev,
              long long idx_base, cudaStream_t st);
// low-order N-d path (eval_lond.cu), 2 <= N <= 4: the same for small tensor-product nests (taken where AUTO would use the
// generic Clenshaw kernel)
bool span_fast(const fastcheb_poly* p, int d);
bool lond_ok(const fastcheb_poly* p, const void* pts, int64_t npts, const void* out_v, int64_t v_stride, const void* out_J,
             int64_t J_stride, bool jac, int mode);
int lond_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, void* out_J, bool jac, long long* bad_dev,
              long long idx_base, cudaStream_t st);
// build the handle from coefficients already on the device (Julia layout); takes no ownership of d_src
int create_from_device(fastcheb_poly** out, int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp,
                       const void* d_src, const double* lb, const double* ub, int device, cudaStream_t st);
}  // namespace fastcheb
