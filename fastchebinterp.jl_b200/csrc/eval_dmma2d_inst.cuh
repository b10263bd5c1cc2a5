// Included by eval_dmma2d_ks{4,8,16}_{val,jac}.cu with FC_KS, FC_JAC, FC_FUNC defined.
#include "eval_dmma_launch.h"

namespace fastcheb {
namespace {
template <typename R, int EC>
cudaError_t go2d(const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st) {
  auto kern = dmma2d_kernel<R, EC, FC_KS, FC_JAC>;
  cudaError_t e = ensure_max_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  kern<<<grid, threads, smem, st>>>(prm);
  note_launch();
  return cudaGetLastError();
}
template <typename R>
cudaError_t go2d_ec(int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st) {
  switch (ec) {
    case 1: return go2d<R, 1>(prm, threads, grid, smem, st);
    case 2: return go2d<R, 2>(prm, threads, grid, smem, st);
    case 3: return go2d<R, 3>(prm, threads, grid, smem, st);
  }
  return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t FC_FUNC(int dtype, int ec, const Dmma2dParams& prm, int threads, int grid, size_t smem, cudaStream_t st) {
  return dtype == FASTCHEB_F64 ? go2d_ec<double>(ec, prm, threads, grid, smem, st)
                               : go2d_ec<float>(ec, prm, threads, grid, smem, st);
}
}  // namespace fastcheb
