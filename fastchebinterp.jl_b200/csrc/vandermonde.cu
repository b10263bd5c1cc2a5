// C ABI: Chebyshev-Vandermonde assembly for least-squares regression (include/fastcheb.h, fastcheb_vandermonde) —
// SURVEY.md 8f N2.
//
// Reference: _chebvandermonde / chebvandermonde, /root/reference/src/regression.jl:5-46.  The generic reference
// routine evaluates the whole interpolant once per basis function and point (O(m P^2), its own TODO at :6-9); this
// kernel writes the m x P matrix in one pass, O(m P): one thread per point walks the column-major multi-index with
// the three-term recurrence of every dimension carried in registers, so consecutive threads write consecutive
// addresses of every column (the matrix is column-major, Julia's Array layout).  Bound: HBM write bandwidth
// (sizeof(T) * m * P bytes written, N * sizeof(T) * m read).
//
// 1-D follows regression.jl:27-46 literally (T_i = 2x * T_{i-1} - T_{i-2}, a rounded product then a rounded
// difference; the domain test is on the normalised coordinate, :33-34), so its matrix is bit-identical to the
// reference's specialised method.  N-d forms the same products with one rounding per factor.
#include <algorithm>
#include <cstring>
#include "host_util.h"
#include "ptx_util.cuh"

namespace fastcheb {
namespace {

struct VdmParams {
  const void* pts;  // m x N, AoS
  void* A;          // m x P column-major, leading dimension lda
  long long m, lda;
  long long* bad;
  int ndim;
  int n[kMaxNdim];
  double lb[kMaxNdim], ub[kMaxNdim];
};

template <typename R, int N>
__global__ void __launch_bounds__(256) vandermonde_kernel(const __grid_constant__ VdmParams prm) {
  const R* pts = reinterpret_cast<const R*>(prm.pts);
  R* A = reinterpret_cast<R*>(prm.A);
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < prm.m; j += (long long)gridDim.x * blockDim.x) {
    R z[N], z2[N];
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      const R lb = (R)prm.lb[d], ub = (R)prm.ub[d];
      const R x = pts[j * N + d];
      z[d] = rsub(rdiv(rmul(rsub(x, lb), R(2)), rsub(ub, lb)), R(1));  // regression.jl:32 / eval.jl:65
      ok = ok && (R(-1) <= z[d]) && (z[d] <= R(1));                     // regression.jl:33 (NaN fails)
      z2[d] = rmul(R(2), z[d]);
    }
    if (!ok) {
      if (prm.bad) atomicMin(prm.bad, j);
      continue;
    }
    // odometer over dims 2..N: (k_d, T_{k_d}, T_{k_d - 1}); `outer` = prod_{d >= 2} T_{k_d}
    int k[N];
    R tc[N], tp[N];
#pragma unroll
    for (int d = 0; d < N; ++d) {
      k[d] = 0;
      tc[d] = R(1);
      tp[d] = R(0);
    }
    R* col = A + j;
    const int n1 = prm.n[0];
    bool done = false;
    while (!done) {
      R outer = R(1);
#pragma unroll
      for (int d = 1; d < N; ++d) outer = rmul(outer, tc[d]);
      // dimension 1: T_0 = 1, T_1 = z, T_i = 2z T_{i-1} - T_{i-2}   (regression.jl:35-43)
      R t0 = R(1), t1 = z[0];
      *col = N == 1 ? t0 : rmul(outer, t0);
      col += prm.lda;
      if (n1 > 1) {
        *col = N == 1 ? t1 : rmul(outer, t1);
        col += prm.lda;
        for (int i = 2; i < n1; ++i) {
          const R t = rsub(rmul(z2[0], t1), t0);
          t0 = t1;
          t1 = t;
          *col = N == 1 ? t : rmul(outer, t);
          col += prm.lda;
        }
      }
      // advance the odometer
      done = true;
#pragma unroll
      for (int d = 1; d < N; ++d) {
        if (!done) break;
        if (k[d] + 1 < prm.n[d]) {
          const R t = k[d] == 0 ? z[d] : rsub(rmul(z2[d], tc[d]), tp[d]);
          tp[d] = tc[d];
          tc[d] = t;
          ++k[d];
          done = false;
        } else {
          k[d] = 0;
          tc[d] = R(1);
          tp[d] = R(0);
        }
      }
    }
  }
}

template <typename R>
cudaError_t launch(int N, const VdmParams& prm, int grid, cudaStream_t st) {
  switch (N) {
    case 1: vandermonde_kernel<R, 1><<<grid, 256, 0, st>>>(prm); break;
    case 2: vandermonde_kernel<R, 2><<<grid, 256, 0, st>>>(prm); break;
    case 3: vandermonde_kernel<R, 3><<<grid, 256, 0, st>>>(prm); break;
    case 4: vandermonde_kernel<R, 4><<<grid, 256, 0, st>>>(prm); break;
    case 5: vandermonde_kernel<R, 5><<<grid, 256, 0, st>>>(prm); break;
    case 6: vandermonde_kernel<R, 6><<<grid, 256, 0, st>>>(prm); break;
    default: return cudaErrorInvalidValue;
  }
  note_launch();
  return cudaGetLastError();
}

}  // namespace
}  // namespace fastcheb

using namespace fastcheb;

extern "C" int fastcheb_vandermonde(int ndim, const int64_t* dims, int dtype, const void* pts, int64_t m,
                                    const double* lb, const double* ub, void* A, int64_t lda, int mem, int device,
                                    void* stream) {
  if (!dims || !pts || !lb || !ub || !A) return fail(FASTCHEB_EINVAL, "null argument");
  if (ndim < 1 || ndim > FASTCHEB_MAX_NDIM) return fail(FASTCHEB_EINVAL, "ndim = %d outside 1..%d", ndim, FASTCHEB_MAX_NDIM);
  if (dtype != FASTCHEB_F32 && dtype != FASTCHEB_F64) return fail(FASTCHEB_EINVAL, "bad dtype %d", dtype);
  if (mem != FASTCHEB_MEM_HOST && mem != FASTCHEB_MEM_DEVICE) return fail(FASTCHEB_EINVAL, "bad mem %d", mem);
  if (m < 0 || lda < m) return fail(FASTCHEB_EINVAL, "bad m / lda");
  if (m == 0) return FASTCHEB_OK;
  long long P = 1;
  for (int d = 0; d < ndim; ++d) {
    if (dims[d] < 1) return fail(FASTCHEB_EINVAL, "order %lld must be nonnegative (regression.jl:30)", (long long)dims[d] - 1);
    P *= dims[d];
  }
  const DeviceInfo* di = device_info(device);
  if (!di) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  if (!dg.ok()) return fail(FASTCHEB_ECUDA, "cudaSetDevice(%d) failed", device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rs = real_size(dtype);
  FlagSlot fs;
  FC_TRY(flag_acquire(device, &fs));
  struct Rel {
    FlagSlot& f;
    ~Rel() { flag_release(f); }
  } rel{fs};
  FC_CUDA(set_flag_async(fs.dev, kNoBad, st));
  VdmParams prm;
  memset(&prm, 0, sizeof prm);
  prm.m = m;
  prm.lda = lda;
  prm.bad = fs.dev;
  prm.ndim = ndim;
  for (int d = 0; d < ndim; ++d) {
    prm.n[d] = (int)dims[d];
    prm.lb[d] = lb[d];
    prm.ub[d] = ub[d];
  }
  void* scratch = nullptr;
  const size_t pbytes = (size_t)m * ndim * rs, abytes = (size_t)lda * P * rs, pal = (pbytes + 255) / 256 * 256;
  if (mem == FASTCHEB_MEM_HOST) {
    scratch = ws_acquire(device, pal + abytes);
    if (!scratch) return fail(FASTCHEB_ENOMEM, "device staging of %zu bytes", pal + abytes);
    if (cudaMemcpyAsync(scratch, pts, pbytes, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      ws_release(device, scratch);
      return fail(FASTCHEB_ECUDA, "point upload failed");
    }
    prm.pts = scratch;
    prm.A = (char*)scratch + pal;
  } else {
    prm.pts = pts;
    prm.A = A;
  }
  const int grid = (int)std::max<long long>(1, std::min<long long>((m + 255) / 256, (long long)di->sms * 8));
  cudaError_t e = dtype == FASTCHEB_F64 ? launch<double>(ndim, prm, grid, st) : launch<float>(ndim, prm, grid, st);
  int rc = FASTCHEB_OK;
  if (e != cudaSuccess) rc = fail(FASTCHEB_ECUDA, "Vandermonde kernel launch failed: %s", cudaGetErrorString(e));
  if (rc == FASTCHEB_OK && mem == FASTCHEB_MEM_HOST &&
      cudaMemcpyAsync(A, prm.A, abytes, cudaMemcpyDeviceToHost, st) != cudaSuccess)
    rc = fail(FASTCHEB_ECUDA, "matrix download failed");
  if (rc == FASTCHEB_OK && (cudaMemcpyAsync(fs.host, fs.dev, sizeof(long long), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                            cudaStreamSynchronize(st) != cudaSuccess))
    rc = fail(FASTCHEB_ECUDA, "Vandermonde assembly failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (scratch) {
    cudaStreamSynchronize(st);
    ws_release(device, scratch);
  }
  if (rc != FASTCHEB_OK) return rc;
  if (*fs.host != kNoBad) {
    err_state().index = *fs.host;
    return fail(FASTCHEB_EDOMAIN, "point %lld not in domain (regression.jl:33)", *fs.host);
  }
  return FASTCHEB_OK;
}
