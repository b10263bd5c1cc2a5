// Host-side entry points of the Clenshaw kernel family; one translation unit per (real type, value|jacobian)
// so the 64 instantiations build in parallel.
#pragma once
#include <cuda_runtime.h>
#include "eval_clenshaw.cuh"

namespace fastcheb {

// points per thread used by the instantiations (value kernels interleave 2 points for ILP)
// 1-D polynomials have tiny per-point state, so a thread carries more points: every broadcast coefficient load then
// feeds more independent recurrences
constexpr int clenshaw_pts(int ndim, int ec, bool jac) { return jac ? (ndim == 1 && ec <= 2 ? 2 : 1) : (ndim == 1 && ec <= 2 ? 4 : 2); }
constexpr int kClenshawMaxThreadsVal = 512;
constexpr int kClenshawMaxThreadsJac = 256;
constexpr int kClenshawMaxEC = 4;
constexpr int kClenshawHeavy = 9;  // ndim * ec at which value kernels drop to 256-thread blocks
constexpr int kClenshawMaxThreadsHeavy = 256;

enum ClenshawOp { kOpLaunch = 0, kOpOccupancy = 1 };

// op == kOpLaunch: launch.  op == kOpOccupancy: *blocks_per_sm = resident CTAs for (threads, smem).
cudaError_t clenshaw_f64_val(int op, int ndim, int ec, const EvalParams& prm, int threads, int grid, size_t smem,
                             cudaStream_t st, int* blocks_per_sm);
cudaError_t clenshaw_f64_jac(int op, int ndim, int ec, const EvalParams& prm, int threads, int grid, size_t smem,
                             cudaStream_t st, int* blocks_per_sm);
cudaError_t clenshaw_f32_val(int op, int ndim, int ec, const EvalParams& prm, int threads, int grid, size_t smem,
                             cudaStream_t st, int* blocks_per_sm);
cudaError_t clenshaw_f32_jac(int op, int ndim, int ec, const EvalParams& prm, int threads, int grid, size_t smem,
                             cudaStream_t st, int* blocks_per_sm);

}  // namespace fastcheb
