// Included by eval_dmma_ks{4,8,16}_{val,jac}.cu with FC_KS, FC_JAC, FC_FUNC defined.
#include "eval_dmma_launch.h"

namespace fastcheb {
namespace {
template <typename R, int N, int EC>
cudaError_t go(int op, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps) {
  auto kern = dmma_kernel<R, N, EC, FC_KS, FC_JAC>;
  cudaError_t e = ensure_max_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  if (op == 1) return cached_occupancy(bps, reinterpret_cast<const void*>(kern), threads, smem);
  kern<<<grid, threads, smem, st>>>(prm);
  note_launch();
  return cudaGetLastError();
}
template <typename R, int N>
cudaError_t go_ec(int op, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps) {
  switch (ec) {
    case 1: return go<R, N, 1>(op, prm, threads, grid, smem, st, bps);
    case 2: return go<R, N, 2>(op, prm, threads, grid, smem, st, bps);
    case 3: return go<R, N, 3>(op, prm, threads, grid, smem, st, bps);
  }
  return cudaErrorInvalidValue;
}
template <typename R>
cudaError_t go_n(int op, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st,
                 int* bps) {
  switch (ndim) {
    case 2: return go_ec<R, 2>(op, ec, prm, threads, grid, smem, st, bps);
    case 3: return go_ec<R, 3>(op, ec, prm, threads, grid, smem, st, bps);
    case 4: return go_ec<R, 4>(op, ec, prm, threads, grid, smem, st, bps);
  }
  return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t FC_FUNC(int op, int dtype, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem,
                    cudaStream_t st, int* bps) {
  return dtype == FASTCHEB_F64 ? go_n<double>(op, ndim, ec, prm, threads, grid, smem, st, bps)
                               : go_n<float>(op, ndim, ec, prm, threads, grid, smem, st, bps);
}
}  // namespace fastcheb
