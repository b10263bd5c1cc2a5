/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU oracle for the FastChebInterp.jl hot path: a restatement of the reference's algorithm
 * (/root/reference/src/eval.jl, /root/reference/src/interp.jl) in plain C.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product library (libfastcheb_cuda) never links, loads or calls it.
 *
 * Parity pinning: see oracle/README.md — the oracle is checked (tests/test_oracle_golden.py) against
 * every known answer the reference publishes for this path (README.md:76-77, 83-84, 116-117,
 * 131-140, 145-146; test/runtests.jl:23-24, 56-59, 70-74, 134-144, 151-165, 176-178).
 * The reference itself (Julia) cannot run in this image, and libfftw3 is absent, so the
 * fit is restated from FFTW's published REDFT00 definition
 *   Y_k = X_0 + (-1)^k X_{n-1} + 2 sum_{j=1}^{n-2} X_j cos(pi j k/(n-1))
 * (FFTW manual "1d Real-even DFTs"; call site interp.jl:43-44) as an exactly-reduced long-double
 * cosine sum — a *truth-grade* transform, more accurate than any FFT.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC)
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ evaluation, both precisions */
#define REAL double
#define FMA fma
#define NAME(x) x##_f64
#include "cheb_oracle_impl.inc"
#undef REAL
#undef FMA
#undef NAME

#define REAL float
#define FMA fmaf
#define NAME(x) x##_f32
#include "cheb_oracle_impl.inc"
#undef REAL
#undef FMA
#undef NAME

/* all components of an element together (the SIMD the reference gets from SVector / Complex elements): Float64, the
 * component counts of the benchmark configs */
#define REAL double
#define FMA fma
#define NAME(x) x##_f64
#define EV 1
#include "cheb_oracle_vec.inc"
#undef EV
#define EV 2
#include "cheb_oracle_vec.inc"
#undef EV
#define EV 3
#include "cheb_oracle_vec.inc"
#undef EV
#define EV 4
#include "cheb_oracle_vec.inc"
#undef EV
#define EV 6
#include "cheb_oracle_vec.inc"
#undef EV
#undef REAL
#undef FMA
#undef NAME

/* values of all E components per pass; falls back to the component-by-component routine for other E */
int oracle_eval_vec_f64(int ndim, const int64_t* dims, int E, const double* coefs, const double* lb, const double* ub,
                        const double* pts, int64_t npts, double* out, int64_t* bad) {
  switch (E) {
    case 1: return oracle_eval_vec_v1_f64(ndim, dims, coefs, lb, ub, pts, npts, out, bad);
    case 2: return oracle_eval_vec_v2_f64(ndim, dims, coefs, lb, ub, pts, npts, out, bad);
    case 3: return oracle_eval_vec_v3_f64(ndim, dims, coefs, lb, ub, pts, npts, out, bad);
    case 4: return oracle_eval_vec_v4_f64(ndim, dims, coefs, lb, ub, pts, npts, out, bad);
    case 6: return oracle_eval_vec_v6_f64(ndim, dims, coefs, lb, ub, pts, npts, out, bad);
  }
  return oracle_eval_f64(ndim, dims, E, coefs, lb, ub, pts, npts, out, bad);
}

/* ------------------------------------------------------------------ fit: interp.jl:39-57 chebcoefs
 * vals/coefs: Julia column-major array of elements of E inline reals (E = K*(1|2)); every real
 * component is transformed independently (interp.jl:59-65 for SVector, FFTW r2r on complex arrays
 * transforms re/im separately).  Work is done in long double; the caller rounds to its own type.
 * Returns 0, or 2 if a non-finite value is found (DomainError, interp.jl:40; *bad = flat index). */
static void dct1_line_ld(const long double* x, long double* y, int64_t n, int64_t stride,
                         const long double* costab /* cos(pi r/(n-1)), r=0..2(n-1)-1 */) {
  const int64_t m = n - 1;
  for (int64_t k = 0; k < n; ++k) {
    long double s = x[0] + ((k & 1) ? -x[m * stride] : x[m * stride]);
    long double acc = 0;
    for (int64_t j = 1; j < m; ++j) acc += x[j * stride] * costab[(j * k) % (2 * m)];
    y[k] = s + 2 * acc;
  }
}

int oracle_fit_ld(int ndim, const int64_t* dims, int E, const long double* vals, long double* coefs,
                  int64_t* bad) {
  int64_t total = E;
  for (int d = 0; d < ndim; ++d) total *= dims[d];
  for (int64_t i = 0; i < total; ++i)
    if (!isfinite((double)vals[i])) { if (bad) *bad = i; return 2; }     /* interp.jl:40 */
  memcpy(coefs, vals, (size_t)total * sizeof(long double));
  int64_t stride = E; /* stride (in reals) of the current dimension */
  for (int d = 0; d < ndim; ++d) {
    const int64_t n = dims[d];
    if (n > 1) {                                  /* interp.jl:43: REDFT00 if n>1 else DHT(1)=identity */
      const int64_t m = n - 1;
      long double* costab = (long double*)malloc((size_t)(2 * m) * sizeof(long double));
      for (int64_t r = 0; r < 2 * m; ++r) {
        /* exact symmetry reduction so that cos(pi/2)=0, cos(pi)=-1 exactly */
        int64_t rr = r <= m ? r : 2 * m - r;      /* cos(pi r/m) = cos(pi (2m-r)/m) */
        long double v;
        if (2 * rr == m) v = 0;
        else if (2 * rr < m) v = cosl(M_PIl * (long double)rr / (long double)m);
        else v = -cosl(M_PIl * (long double)(m - rr) / (long double)m);
        costab[r] = v;
      }
      const int64_t outer = total / (stride * n);
#pragma omp parallel
      {
        long double* y = (long double*)malloc((size_t)n * sizeof(long double));
#pragma omp for collapse(2) schedule(static)
        for (int64_t o = 0; o < outer; ++o)
          for (int64_t i = 0; i < stride; ++i) {
            long double* line = coefs + o * stride * n + i;
            dct1_line_ld(line, y, n, stride, costab);                 /* interp.jl:44 */
            for (int64_t k = 0; k < n; ++k) {
              /* interp.jl:49 (divide by 2(n-1)) and :50-54 (double interior 2:n-1) */
              long double sc = (k == 0 || k == m) ? (long double)1 / (2 * m) : (long double)1 / m;
              line[k * stride] = y[k] * sc;
            }
          }
        free(y);
      }
      free(costab);
    }
    stride *= n;
  }
  return 0;
}

/* interp.jl:77-93 droptol.  coefs as above (E reals per element, infnorm over the E reals; for
 * complex elements use cplx=1: infnorm(x::Number)=abs(x) is the complex modulus, infnorm of an
 * SVector is the max modulus of its entries, interp.jl:74-75).
 * Writes the new sizes to newdims; returns 1 if anything is dropped, 0 if unchanged (early return :79). */
static double elt_infnorm(const double* c, int K, int cplx) {
  double m = 0;
  for (int k = 0; k < K; ++k) {
    double a = cplx ? hypot(c[2 * k], c[2 * k + 1]) : fabs(c[k]);
    if (a > m || a != a) m = a;
  }
  return m;
}

int oracle_droptol_shape(int ndim, const int64_t* dims, int K, int cplx, const double* coefs, double tol,
                         int64_t* newdims) {
  const int E = K * (cplx ? 2 : 1);
  int64_t total = 1;
  for (int d = 0; d < ndim; ++d) { total *= dims[d]; newdims[d] = dims[d]; }
  double mx = 0;
  for (int64_t i = 0; i < total; ++i) { double a = elt_infnorm(coefs + i * E, K, cplx); if (a > mx) mx = a; }
  const double abstol = mx * tol;                                     /* interp.jl:78 */
  int all_ge = 1;
  for (int64_t i = 0; i < total && all_ge; ++i) if (!(elt_infnorm(coefs + i * E, K, cplx) >= abstol)) all_ge = 0;
  if (all_ge) return 0;                                               /* interp.jl:79 */
  int64_t stride = 1;
  for (int d = 0; d < ndim; ++d) {                                    /* interp.jl:83-91 */
    int64_t n = dims[d];
    const int64_t outer = total / (stride * dims[d]);
    while (n > 1) {
      int any_ge = 0;
      for (int64_t o = 0; o < outer && !any_ge; ++o)
        for (int64_t i = 0; i < stride && !any_ge; ++i)
          if (elt_infnorm(coefs + (o * stride * dims[d] + (n - 1) * stride + i) * E, K, cplx) >= abstol) any_ge = 1;
      if (any_ge) break;
      --n;
    }
    newdims[d] = n;
    stride *= dims[d];
  }
  return 1;
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the timed CPU baseline wants all host cores */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  extern void omp_set_num_threads(int);
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
  {
#pragma omp single
    n = __builtin_omp_get_num_threads();
  }
#endif
  return n;
}
