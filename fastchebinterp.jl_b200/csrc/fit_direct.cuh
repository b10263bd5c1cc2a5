// Generic DCT-I fit pass (any n, any rank, any stride) + coefficient-norm reductions for droptol.
//
// Replaces the arithmetic of chebcoefs (/root/reference/src/interp.jl:39-57): FFTW.r2r(vals, REDFT00)
// along one dimension (interp.jl:43-44), fused with that dimension's share of the normalisation
// (division by 2(n-1), interp.jl:49, and the doubling of interior indices, interp.jl:50-54).
//
// This is the correctness path for arbitrary sizes: a direct cosine sum with an exactly reduced
// cosine table, accumulated in FP64 in blocks (pairwise-style) so the error stays ~1e-15 * max|x|.
// The bandwidth-bound path for batched short transforms is fit_dmma.cuh.
#pragma once
#include <cstdint>
#include "ptx_util.cuh"

namespace fastcheb {

// One pass along a dimension of length n (m = n-1 >= 1).  The array is viewed as [outer][n][inner]
// reals (inner = E * prod of lower dims, outer = prod of higher dims * howmany).
// ctab[r] = cos(pi r / m) as an unevaluated double-double (hi, lo), r in [0, 2m).
// The sum is accumulated in double-double arithmetic (error-free two-product / two-sum), so a
// coefficient is the correctly rounded transform of the (rounded) inputs to ~1e-30 relative — the
// near-cancelling high-order coefficients that droptol (interp.jl:77-93) compares with eps*max|c|
// carry no O(eps*|x|) summation noise.
template <typename R>
__global__ void __launch_bounds__(256) dct1_direct_kernel(const R* __restrict__ src, R* __restrict__ dst,
                                                          const double2* __restrict__ ctab, long long total,
                                                          int n, long long inner) {
  const int m = n - 1;
  const long long ni = (long long)n * inner;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long o = g / ni;
    const long long rem = g - o * ni;
    const int k = (int)(rem / inner);
    const long long i = rem - (long long)k * inner;
    const R* x = src + o * ni + i;
    double sh = 0.0, sl = 0.0;  // running sum of x_j cos(pi j k/m), j = 1..m-1
    int r = 0;                  // (j*k) mod 2m
    for (int j = 1; j < m; ++j) {
      r += k;
      if (r >= 2 * m) r -= 2 * m;
      const double xv = (double)x[(long long)j * inner];
      const double2 c = ctab[r];
      const double p = __dmul_rn(xv, c.x);
      double pe = __fma_rn(xv, c.x, -p);  // exact product error
      pe = __fma_rn(xv, c.y, pe);
      const double t = __dadd_rn(sh, p);  // two-sum
      const double bb = __dsub_rn(t, sh);
      const double err = __dadd_rn(__dsub_rn(sh, __dsub_rn(t, bb)), __dsub_rn(p, bb));
      sh = t;
      sl = __dadd_rn(sl, __dadd_rn(err, pe));
    }
    // y = x0 + (-1)^k xm + 2 (sh + sl)        REDFT00, interp.jl:44
    const double x0 = (double)x[0];
    const double xm = (k & 1) ? -(double)x[(long long)m * inner] : (double)x[(long long)m * inner];
    double a = __dadd_rn(x0, xm);
    double ae = __dadd_rn(__dsub_rn(x0, __dsub_rn(a, __dsub_rn(a, x0))), __dsub_rn(xm, __dsub_rn(a, x0)));
    const double s2 = __dmul_rn(2.0, sh);
    const double y = __dadd_rn(a, s2);
    const double bb = __dsub_rn(y, a);
    const double ye = __dadd_rn(__dsub_rn(a, __dsub_rn(y, bb)), __dsub_rn(s2, bb));
    const double tail = __dadd_rn(__dadd_rn(ye, ae), __dmul_rn(2.0, sl));
    const double yy = __dadd_rn(y, tail);
    const double sc = (k == 0 || k == m) ? 0.5 / (double)m : 1.0 / (double)m;  // interp.jl:49-54
    dst[g] = (R)__dmul_rn(yy, sc);
  }
}

// The same pass with one WARP per output element, for long lines in small arrays (a single 1-D fit of n = 10001 has
// only 10001 outputs: one thread each leaves three quarters of the SMs idle for 10^4 dependent steps, 1.9 ms; this
// form takes ~0.1 ms).  Lane l sums j = 1+l, 33+l, ..; the 32 double-double partial sums are combined with two-sums.
template <typename R>
__global__ void __launch_bounds__(256) dct1_direct_warp_kernel(const R* __restrict__ src, R* __restrict__ dst,
                                                               const double2* __restrict__ ctab, long long total,
                                                               int n, long long inner) {
  const int m = n - 1;
  const long long ni = (long long)n * inner;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long g = warp0; g < total; g += nwarps) {
    const long long o = g / ni;
    const long long rem = g - o * ni;
    const int k = (int)(rem / inner);
    const long long i = rem - (long long)k * inner;
    const R* x = src + o * ni + i;
    double sh = 0.0, sl = 0.0;
    int r = (int)(((long long)(1 + lane) * k) % (2LL * m));  // (j*k) mod 2m at j = 1 + lane
    const int step = (int)((32LL * k) % (2LL * m));
    for (int j = 1 + lane; j < m; j += 32) {
      const double xv = (double)x[(long long)j * inner];
      const double2 c = ctab[r];
      const double p = __dmul_rn(xv, c.x);
      double pe = __fma_rn(xv, c.x, -p);
      pe = __fma_rn(xv, c.y, pe);
      const double t = __dadd_rn(sh, p);
      const double bb = __dsub_rn(t, sh);
      const double err = __dadd_rn(__dsub_rn(sh, __dsub_rn(t, bb)), __dsub_rn(p, bb));
      sh = t;
      sl = __dadd_rn(sl, __dadd_rn(err, pe));
      r += step;
      if (r >= 2 * m) r -= 2 * m;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const double oh = __shfl_down_sync(0xffffffffu, sh, off), ol = __shfl_down_sync(0xffffffffu, sl, off);
      const double t = __dadd_rn(sh, oh);
      const double bb = __dsub_rn(t, sh);
      const double err = __dadd_rn(__dsub_rn(sh, __dsub_rn(t, bb)), __dsub_rn(oh, bb));
      sh = t;
      sl = __dadd_rn(sl, __dadd_rn(err, ol));
    }
    if (lane == 0) {
      const double x0 = (double)x[0];
      const double xm = (k & 1) ? -(double)x[(long long)m * inner] : (double)x[(long long)m * inner];
      double a = __dadd_rn(x0, xm);
      double ae = __dadd_rn(__dsub_rn(x0, __dsub_rn(a, __dsub_rn(a, x0))), __dsub_rn(xm, __dsub_rn(a, x0)));
      const double s2 = __dmul_rn(2.0, sh);
      const double y = __dadd_rn(a, s2);
      const double bb = __dsub_rn(y, a);
      const double ye = __dadd_rn(__dsub_rn(a, __dsub_rn(y, bb)), __dsub_rn(s2, bb));
      const double tail = __dadd_rn(__dadd_rn(ye, ae), __dmul_rn(2.0, sl));
      const double yy = __dadd_rn(y, tail);
      const double sc = (k == 0 || k == m) ? 0.5 / (double)m : 1.0 / (double)m;  // interp.jl:49-54
      dst[g] = (R)__dmul_rn(yy, sc);
    }
  }
}

// first non-finite real of `vals` (interp.jl:40): atomicMin of the flat index
template <typename R>
__global__ void __launch_bounds__(256) nonfinite_kernel(const R* __restrict__ v, long long total, long long base,
                                                        long long* __restrict__ bad) {
  long long first = 0x7fffffffffffffffLL;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const double a = (double)v[g];
    if (!(fabs(a) <= 1.79769313486231570815e308) && g < first) first = g;
  }
  if (first != 0x7fffffffffffffffLL) atomicMin(bad, first + base);
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void atomic_min_nonneg(double* addr, double v) {
  atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// infnorm of one coefficient element (interp.jl:74-75): abs for numbers (complex modulus), max over
// the K entries of an SVector.
template <typename R>
__device__ __forceinline__ double elt_infnorm(const R* c, int K, int cplx) {
  double mx = 0.0;
  for (int k = 0; k < K; ++k) {
    const double a = cplx ? hypot((double)c[2 * k], (double)c[2 * k + 1]) : fabs((double)c[k]);
    mx = a > mx ? a : mx;
  }
  return mx;
}

// stats[f*2+0] = max infnorm, stats[f*2+1] = min infnorm of fit f; hmax (optional, single fit):
// per-dimension hyperplane maxima, hmax[off_d + k_d] (interp.jl:77-91).
struct NormParams {
  int ndim;
  int n[kMaxNdim];
  int hoff[kMaxNdim];
};
template <typename R>
__global__ void __launch_bounds__(256) coef_norm_kernel(const R* __restrict__ c, long long elems_per_fit,
                                                        long long howmany, int K, int cplx,
                                                        double* __restrict__ stats, double* __restrict__ hmax,
                                                        const NormParams np) {
  const int E = K * (cplx ? 2 : 1);
  const long long total = elems_per_fit * howmany;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long f = g / elems_per_fit;
    const double a = elt_infnorm(c + g * E, K, cplx);
    // warp-aggregate when the whole warp is inside one fit
    const long long f0 = __shfl_sync(__activemask(), f, 0);
    const unsigned act = __activemask();
    if (__all_sync(act, f == f0) && act == 0xffffffffu) {
      double mx = a, mn = a;
      for (int o = 16; o; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      if ((threadIdx.x & 31) == 0) {
        atomic_max_nonneg(&stats[2 * f], mx);
        atomic_min_nonneg(&stats[2 * f + 1], mn);
      }
    } else {
      atomic_max_nonneg(&stats[2 * f], a);
      atomic_min_nonneg(&stats[2 * f + 1], a);
    }
    if (hmax) {
      long long rem = g - f * elems_per_fit;
      for (int d = 0; d < np.ndim; ++d) {
        const int kd = (int)(rem % np.n[d]);
        rem /= np.n[d];
        atomic_max_nonneg(&hmax[np.hoff[d] + kd], a);
      }
    }
  }
}

// maxabs[f] = max_k infnorm(C_k) of fit f (maxabs zero-initialised).  Each warp reduces chunks of up to
// 1024 elements that lie inside one fit, then issues one atomic.
template <typename R>
__global__ void __launch_bounds__(256) coef_max_kernel(const R* __restrict__ c, long long elems_per_fit,
                                                       long long howmany, int K, int cplx,
                                                       double* __restrict__ maxabs) {
  const int E = K * (cplx ? 2 : 1);
  const long long chunks_per_fit = (elems_per_fit + 1023) / 1024;
  const long long nchunks = chunks_per_fit * howmany;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long ch = warp0; ch < nchunks; ch += nwarps) {
    const long long f = ch / chunks_per_fit;
    const long long k0 = (ch - f * chunks_per_fit) * 1024;
    const long long k1 = k0 + 1024 < elems_per_fit ? k0 + 1024 : elems_per_fit;
    double mx = 0.0;
    for (long long k = k0 + lane; k < k1; k += 32) mx = fmax(mx, elt_infnorm(c + (f * elems_per_fit + k) * E, K, cplx));
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) atomic_max_nonneg(&maxabs[f], mx);
  }
}

// copy the leading block newdims of a column-major array of E-real elements (interp.jl:92)
struct BlockParams {
  int ndim;
  int n_old[kMaxNdim];
  int n_new[kMaxNdim];
};
template <typename R>
__global__ void __launch_bounds__(256) leading_block_kernel(const R* __restrict__ src, R* __restrict__ dst,
                                                            long long total_new, int E, const BlockParams bp) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total_new;
       g += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(g % E);
    long long rem = g / E;
    long long so = 0, mul = 1;
    for (int d = 0; d < bp.ndim; ++d) {
      const int kd = (int)(rem % bp.n_new[d]);
      rem /= bp.n_new[d];
      so += kd * mul;
      mul *= bp.n_old[d];
    }
    dst[g] = src[so * E + e];
  }
}

}  // namespace fastcheb
