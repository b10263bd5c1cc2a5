// Included by eval_dmma_ks{4,8,16}_{val,jac}.cu with FC_KS, FC_JAC, FC_FUNC defined.
#include "eval_dmma_launch.h"

namespace fastcheb {
namespace {
template <int N, int EC>
cudaError_t go(int op, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps) {
  auto kern = dmma_kernel<N, EC, FC_KS, FC_JAC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptinMax);
  if (e != cudaSuccess) return e;
  if (op == 1) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, kern, threads, smem);
  kern<<<grid, threads, smem, st>>>(prm);
  note_launch();
  return cudaGetLastError();
}
template <int N>
cudaError_t go_ec(int op, int ec, const DmmaParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps) {
  switch (ec) {
    case 1: return go<N, 1>(op, prm, threads, grid, smem, st, bps);
    case 2: return go<N, 2>(op, prm, threads, grid, smem, st, bps);
    case 3: return go<N, 3>(op, prm, threads, grid, smem, st, bps);
  }
  return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t FC_FUNC(int op, int ndim, int ec, const DmmaParams& prm, int threads, int grid, size_t smem,
                    cudaStream_t st, int* bps) {
  switch (ndim) {
    case 2: return go_ec<2>(op, ec, prm, threads, grid, smem, st, bps);
    case 3: return go_ec<3>(op, ec, prm, threads, grid, smem, st, bps);
    case 4: return go_ec<4>(op, ec, prm, threads, grid, smem, st, bps);
  }
  return cudaErrorInvalidValue;
}
}  // namespace fastcheb
