// Experiment harness: runs the production DMMA evaluation kernel (eval_dmma.cuh) for the 3-D order-32 K=3
// workload with compile-time variants (-DFC_DMMA_MT=.., -DFC_VARIANT=..) and run-time launch shapes.
// usage: dmma_eval_variants [npts] ; prints one JSON line per (warps, tilesPerStage, nstage) combination.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../fastchebinterp.jl_b200/csrc/eval_dmma.cuh"
using namespace fastcheb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

int main(int argc, char** argv) {
  const long long npts = argc > 1 ? atoll(argv[1]) : (1ll << 22);
#ifndef FC_EC
#define FC_EC 3
#endif
  constexpr int N = 3, EC = FC_EC, KS = 8, n = 33;  // -DFC_EC=1: scalar-valued 3-D
  const int M = n * n, ntiles = (M + 7) / 8, U = KS * 32 + 8, TR = kDmmaMeta + EC * U;
  std::vector<double> coefs((size_t)EC * n * n * n), pts((size_t)npts * 3);
  srand(1);
  for (auto& c : coefs) c = rand() / (double)RAND_MAX * 2 - 1;
  for (auto& p : pts) p = rand() / (double)RAND_MAX * 2 - 1;
  double *d_c, *d_f, *d_p, *d_o;
  CK(cudaMalloc(&d_c, coefs.size() * 8)); CK(cudaMalloc(&d_f, (size_t)ntiles * TR * 8));
  CK(cudaMalloc(&d_p, pts.size() * 8)); CK(cudaMalloc(&d_o, (size_t)npts * EC * 8));
  CK(cudaMemcpy(d_c, coefs.data(), coefs.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_p, pts.data(), pts.size() * 8, cudaMemcpyHostToDevice));
  FragDims fd; memset(&fd, 0, sizeof fd); fd.nd = 2; fd.n[0] = n; fd.n[1] = n; fd.toff[0] = 0; fd.toff[1] = n;
  relayout_frag_kernel<double><<<512, 256>>>(d_c, d_f, (long long)ntiles * TR, EC, 0, EC, n, M, KS, fd);
  CK(cudaDeviceSynchronize());
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  auto kern = dmma_kernel<double, N, EC, KS, false>;
  const int rows = kDmmaMT * 8, tabStride = (2 * n) | 1;
  #ifdef FC_PROFILE
  for (int warps : {16}) for (int tps : {1, 4}) for (int ns : {3}) {
#else
  for (int warps : {16, 20}) for (int tps : {4}) for (int ns : {3}) {
#endif
    DmmaParams prm; memset(&prm, 0, sizeof prm);
    prm.frag = d_f; prm.pts = d_p; prm.out_v = d_o; prm.npts = npts; prm.v_stride = EC; prm.Etot = EC;
    prm.n[0] = prm.n[1] = prm.n[2] = n; prm.ntiles = ntiles; prm.M = M; prm.tilesPerStage = tps; prm.nstage = ns;
    prm.tabStride = tabStride;
    for (int d = 0; d < 3; ++d) { prm.lb[d] = -1; prm.ub[d] = 1; }
    const size_t smem = 128 + (size_t)ns * tps * TR * 8 + (size_t)warps * rows * tabStride * 8;
    if (smem > prop.sharedMemPerBlockOptin || warps * 32 > dmma_max_threads(EC, KS, false)) continue;
#ifdef FC_PROFILE
    long long* d_prof; CK(cudaMalloc(&d_prof, 16 * 4 * 8)); CK(cudaMemset(d_prof, 0, 16 * 4 * 8)); prm.prof = d_prof;
#endif
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, warps * 32, smem));
    const int grid = prop.multiProcessorCount * bps;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<grid, warps * 32, smem>>>(prm);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; ++i) kern<<<grid, warps * 32, smem>>>(prm);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
    const double rate = npts / (ms * 1e-3);
    printf("{\"mt\":%d,\"variant\":%d,\"warps\":%d,\"bps\":%d,\"tps\":%d,\"nstage\":%d,\"smem\":%zu,\"ms\":%.3f,\"pts_per_s\":%.4g,\"frac_of_37TF\":%.3f}\n",
           kDmmaMT,
#ifdef FC_VARIANT
           FC_VARIANT,
#else
           0,
#endif
           warps, bps, tps, ns, smem, ms, rate, rate * (74118.0 * EC) / 37.0e12);
    fflush(stdout);
#ifdef FC_PROFILE
    { long long hp[64]; CK(cudaMemcpy(hp, d_prof, sizeof hp, cudaMemcpyDeviceToHost));
      for (int w = 0; w < warps; ++w) printf("   warp %2d total %lld wait %lld (%.1f%%) slow %lld issue %lld\n", w, hp[4*w], hp[4*w+1], 100.0*hp[4*w+1]/hp[4*w], hp[4*w+2], hp[4*w+3]); }
#endif
  }
  return 0;
}
