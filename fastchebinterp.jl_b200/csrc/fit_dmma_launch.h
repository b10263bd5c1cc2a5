// Host-side entry points of the tensor-core DCT-I path for batches of short 1-D fits (fit_dmma.cuh).
#pragma once
#include <vector>
#include "host_util.h"

namespace fastcheb {
// Transform as many whole units of fits as the tensor-core path can take: `howmany` fits of n elements
// of E inline reals each, stored back to back (so line (f, e) has element stride E).  *done receives the
// number of leading fits handled (0 when the shape is not covered or the buffers are not 16-byte
// aligned); the caller sends the remaining fits through the generic kernel.  Non-finite inputs are
// reported through bad_dev (idx_base + flat real index, atomicMin).  fit_max_dev (optional, real coefficients
// only): max |c| of every handled fit.
int fit_dmma_1d(const DeviceInfo* di, int dev, int n, int E, long long howmany, int dtype, const void* vals,
                void* coefs, long long* bad_dev, long long idx_base, double* fit_max_dev, cudaStream_t st, long long* done);
// Folded, normalised REDFT00 matrices of size n in mma.sync A-fragment order (parts O | B | C, or — split, 4 | n-1 —
// P | Q | B | C; slot map `map`, see fit_dmma.cuh) and the k-step counts of the parts (kc == 0: the even half is folded
// once only, and there is no split).
void fit_afrag_host(int n, int map, bool split, std::vector<double>& h, int* ko, int* kb, int* kc);

// Resident N-d / long-line fit (fit_nd.cuh): `howmany` arrays of dims[0] x .. x dims[ndim-1] elements of E inline reals.
// *handled = false when the shape is not covered (an array does not fit shared memory, a dimension is too long); the
// caller then uses the generic per-dimension kernels.
int fit_nd(const DeviceInfo* di, int dev, int ndim, const int64_t* dims, int K, int cplx, long long howmany, int dtype,
           const void* vals, void* coefs, long long* bad_dev, long long idx_base, double* maxabs_dev, cudaStream_t st,
           bool* handled, bool* need_max = nullptr,  // *need_max: the caller still has to reduce max|c| per array
           double* stats_dev = nullptr, const int* stats_hoff = nullptr, double stats_tol = 0.0, bool* stats_done = nullptr);
// stats_dev (howmany == 1 only): droptol's statistics computed by the same kernel, layout as coef_norm_kernel
// (fit_direct.cuh): [0] max, [1] min, [2 + hoff[d] + k] >= stats_tol * max iff the hyperplane k_d = k keeps an element that
// large (what droptol asks of the hyperplane maxima); *stats_done = true when they were produced.
}  // namespace fastcheb
