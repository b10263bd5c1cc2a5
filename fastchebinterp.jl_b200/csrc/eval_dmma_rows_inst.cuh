// Included by eval_dmma_rows_ks{4,8,16}.cu with FC_KS and FC_FUNC defined.
#include "eval_dmma_launch.h"

namespace fastcheb {
namespace {
template <typename R, int N, int EC, int TPR>
cudaError_t go_rows(const DmmaRowsParams& prm, int threads, int grid, size_t smem, cudaStream_t st) {
  auto kern = dmma_rows_kernel<R, N, EC, FC_KS, TPR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptinMax);
  if (e != cudaSuccess) return e;
  kern<<<grid, threads, smem, st>>>(prm);
  note_launch();
  return cudaGetLastError();
}
template <typename R, int N, int EC>
cudaError_t go_rows_tpr(int tpr, const DmmaRowsParams& prm, int threads, int grid, size_t smem, cudaStream_t st) {
  switch (tpr) {
    case 1: return go_rows<R, N, EC, 1>(prm, threads, grid, smem, st);
    case 2: return go_rows<R, N, EC, 2>(prm, threads, grid, smem, st);
    case 4: return go_rows<R, N, EC, 4>(prm, threads, grid, smem, st);
  }
  return cudaErrorInvalidValue;
}
template <typename R, int N>
cudaError_t go_rows_ec(int ec, int tpr, const DmmaRowsParams& prm, int threads, int grid, size_t smem, cudaStream_t st) {
  switch (ec) {
    case 1: return go_rows_tpr<R, N, 1>(tpr, prm, threads, grid, smem, st);
    case 2: return go_rows_tpr<R, N, 2>(tpr, prm, threads, grid, smem, st);
    case 3: return go_rows_tpr<R, N, 3>(tpr, prm, threads, grid, smem, st);
  }
  return cudaErrorInvalidValue;
}
template <typename R>
cudaError_t go_rows_n(int ndim, int ec, int tpr, const DmmaRowsParams& prm, int threads, int grid, size_t smem,
                      cudaStream_t st) {
  switch (ndim) {
    case 3: return go_rows_ec<R, 3>(ec, tpr, prm, threads, grid, smem, st);
    case 4: return go_rows_ec<R, 4>(ec, tpr, prm, threads, grid, smem, st);
  }
  return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t FC_FUNC(int dtype, int ndim, int ec, int tpr, const DmmaRowsParams& prm, int threads, int grid, size_t smem,
                    cudaStream_t st) {
  return dtype == FASTCHEB_F64 ? go_rows_n<double>(ndim, ec, tpr, prm, threads, grid, smem, st)
                               : go_rows_n<float>(ndim, ec, tpr, prm, threads, grid, smem, st);
}
}  // namespace fastcheb
