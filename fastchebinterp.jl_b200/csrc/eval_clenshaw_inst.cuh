// Included by eval_clenshaw_{f64,f32}_{val,jac}.cu with FC_REAL, FC_JAC, FC_FUNC defined.
#include "eval_clenshaw_launch.h"
#include "host_util.h"

namespace fastcheb {
namespace {
template <int N, int EC>
cudaError_t go(int op, const EvalParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps) {
  constexpr int P = clenshaw_pts(N, EC, FC_JAC);
  if constexpr (N == 1) {
    if (prm.nstage == 0) {  // the coefficient line stays in global memory (plan_clenshaw, poly.cu)
      auto kg = clenshaw_kernel<FC_REAL, N, EC, P, FC_JAC, true>;
      if (op == kOpOccupancy) return cached_occupancy(bps, reinterpret_cast<const void*>(kg), threads, smem);
      kg<<<grid, threads, smem, st>>>(prm);
      note_launch();
      return cudaGetLastError();
    }
  }
  auto kern = clenshaw_kernel<FC_REAL, N, EC, P, FC_JAC>;
  cudaError_t e = ensure_max_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  if (op == kOpOccupancy) return cached_occupancy(bps, reinterpret_cast<const void*>(kern), threads, smem);
  kern<<<grid, threads, smem, st>>>(prm);
  note_launch();
  return cudaGetLastError();
}
template <int N>
cudaError_t go_ec(int op, int ec, const EvalParams& prm, int threads, int grid, size_t smem, cudaStream_t st, int* bps) {
  switch (ec) {
    case 1: return go<N, 1>(op, prm, threads, grid, smem, st, bps);
    case 2: return go<N, 2>(op, prm, threads, grid, smem, st, bps);
    case 3: return go<N, 3>(op, prm, threads, grid, smem, st, bps);
    case 4: return go<N, 4>(op, prm, threads, grid, smem, st, bps);
  }
  return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t FC_FUNC(int op, int ndim, int ec, const EvalParams& prm, int threads, int grid, size_t smem,
                    cudaStream_t st, int* bps) {
  switch (ndim) {
    case 1: return go_ec<1>(op, ec, prm, threads, grid, smem, st, bps);
    case 2: return go_ec<2>(op, ec, prm, threads, grid, smem, st, bps);
    case 3: return go_ec<3>(op, ec, prm, threads, grid, smem, st, bps);
    case 4: return go_ec<4>(op, ec, prm, threads, grid, smem, st, bps);
    case 5: return go_ec<5>(op, ec, prm, threads, grid, smem, st, bps);
    case 6: return go_ec<6>(op, ec, prm, threads, grid, smem, st, bps);
  }
  return cudaErrorInvalidValue;
}
}  // namespace fastcheb
