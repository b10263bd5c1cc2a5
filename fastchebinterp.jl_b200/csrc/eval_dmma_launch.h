// Host-side entry points of the DMMA evaluation path (eval_dmma.cuh).
#pragma once
#include "eval_dmma.cuh"
#include "eval_dmma2d.cuh"
#include "poly.h"

namespace fastcheb {

constexpr int kDmmaMaxEC = 3;

inline int dmma_max_warps(int ec, int ks, bool jac) { return dmma_max_threads(ec, ks, jac) / 32; }

// per-(KS, JAC) translation units: op 0 = launch, op 1 = occupancy query
cudaError_t dmma_dispatch_ks4_val(int op, int dtype, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps);
cudaError_t dmma_dispatch_ks8_val(int op, int dtype, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps);
cudaError_t dmma_dispatch_ks12_val(int op, int dtype, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps);
cudaError_t dmma_dispatch_ks16_val(int op, int dtype, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps);
cudaError_t dmma_dispatch_ks4_jac(int op, int dtype, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps);
cudaError_t dmma_dispatch_ks8_jac(int op, int dtype, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps);
cudaError_t dmma_dispatch_ks12_jac(int op, int dtype, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps);
cudaError_t dmma_dispatch_ks16_jac(int op, int dtype, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps);

// resident 2-D kernels (eval_dmma2d.cuh), one translation unit per (KS, JAC)
cudaError_t dmma2d_dispatch_ks4_val(int dtype, int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st);
cudaError_t dmma2d_dispatch_ks8_val(int dtype, int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st);
cudaError_t dmma2d_dispatch_ks12_val(int dtype, int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st);
cudaError_t dmma2d_dispatch_ks16_val(int dtype, int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st);
cudaError_t dmma2d_dispatch_ks4_jac(int dtype, int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st);
cudaError_t dmma2d_dispatch_ks8_jac(int dtype, int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st);
cudaError_t dmma2d_dispatch_ks12_jac(int dtype, int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st);
cudaError_t dmma2d_dispatch_ks16_jac(int dtype, int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st);

// decide whether the handle gets a DMMA plan and build the fragment stream (no-op when not applicable)
int dmma_plan(fastcheb_poly* p, cudaStream_t st);
int dmma_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, int64_t v_stride, void* out_J,
              int64_t J_stride, bool jac, long long* bad_dev, long long idx_base, cudaStream_t st, int mode = 0,
                const void* tan = nullptr);

}  // namespace fastcheb
