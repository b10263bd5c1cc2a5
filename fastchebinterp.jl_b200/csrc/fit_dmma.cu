// Host side of the tensor-core DCT-I path: folded transform matrices in A-fragment order, launch.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>
#include "fit_dmma.cuh"
#include "fit_dmma_launch.h"

namespace fastcheb {
namespace {

struct AFrag {
  double* d = nullptr;
  int ko = 0, kb = 0, kc = 0;  // k-steps of parts O, B, C (fit_dmma.cuh); kc == 0: one fold only
  size_t doubles = 0;
};
std::mutex g_mu;
std::map<std::pair<int, int>, AFrag> g_afrag;  // (device, 4 n + slot map)

// cos(pi r / N) with r reduced exactly, so that the symmetries of the transform hold bit for bit
long double cospi_ratio(long long r, int N) {
  const long double pi = 3.14159265358979323846264338327950288L;
  r %= 2LL * N;
  if (r > N) r = 2LL * N - r;
  if (2 * r == N) return 0.0L;
  if (2 * r < N) return cosl(pi * (long double)r / (long double)N);
  return -cosl(pi * (long double)(N - r) / (long double)N);
}

// slot map of the launch (fit_dmma.cuh): the tuned, bank-conflict-free ones exist for n = 65
int fit_slot_map(int n, int dtype) {
  static const bool plain = [] {  // experiments: FASTCHEB_FIT_MAP=0 forces the plain map
    const char* e = getenv("FASTCHEB_FIT_MAP");
    return e && atoi(e) == 0;
  }();
  return (n == 65 && !plain && dtype == FASTCHEB_F32) ? 1 : 0;
}

}  // namespace

// Folded, normalised REDFT00 matrices of size n in A-fragment order (parts O | B | C, or P | Q | B | C when map != 0):
// the host-side builder shared by the batched 1-D kernel (fit_dmma.cuh) and the resident N-d kernel (fit_nd.cuh).
void fit_afrag_host(int n, int map, bool split, std::vector<double>& h, int* ko_out, int* kb_out, int* kc_out) {
  const int N = n - 1, H = N / 2;
  const bool two = N >= 4 && N % 4 == 0;  // the even half folds once more
  split = split && two;                   // parts P, Q exist for the twice-folded form only
  // folded inputs (= outputs) of each part
  const int Jo = (N + 1) / 2;
  const int Jb = two ? H / 2 + 1 : N / 2 + 1;
  const int Jc = two ? H / 2 : 0;
  const int ko = (Jo + 3) / 4, kb = (Jb + 3) / 4, kc = (Jc + 3) / 4;
  const int mo = (ko + 1) / 2, mb = (kb + 1) / 2, mc = (kc + 1) / 2;
  // two folds: the odd half is split once more by the parity of j (parts P and Q, fit_dmma.cuh) and replaces part O
  const size_t doubles = (size_t)(split ? 2 * mc * kc : mo * ko) * 32 + (size_t)(mb * kb + mc * kc) * 32;
  h.assign(doubles, 0.0);
  // REDFT00 (interp.jl:44) with the scaling of interp.jl:49-54 folded in: output k, folded input j carrying REDFT00
  // weight w (end points and self-paired middle elements count once, everything else twice)
  auto entry = [&](int k, int j, bool once) -> double {
    const long double s = (k == 0 || k == N) ? 0.5L / (long double)N : 1.0L / (long double)N;
    return (double)(s * (once ? 1.0L : 2.0L) * cospi_ratio((long long)j * k, N));
  };
  auto fill = [&](size_t off, int K, int M, int J, int kstride, int koff, int selfpair) {
    for (int mt = 0; mt < M; ++mt)
      for (int ks = 0; ks < K; ++ks)
        for (int l = 0; l < 32; ++l) {
          const int row = 8 * mt + l / 4, j = fit_map_j(map, ks, l % 4), k = kstride * row + koff;
          if (row < J && j < J && k <= N) h[off + (size_t)(mt * K + ks) * 32 + l] = entry(k, j, j == 0 || j == selfpair);
        }
  };
  size_t off = 0;
  if (split) {
    // parts P, Q: row l (k = 2l+1, l < N/4), column i: b_{2i} resp. b_{2i+1}; b_0 = x_0 - x_N carries weight 1
    for (int odd = 0; odd < 2; ++odd) {
      for (int mt = 0; mt < mc; ++mt)
        for (int ks = 0; ks < kc; ++ks)
          for (int l = 0; l < 32; ++l) {
            const int row = 8 * mt + l / 4, i = fit_map_i(map, ks, l % 4), j = 2 * i + odd;
            if (row < Jc && i < Jc) h[off + (size_t)(mt * kc + ks) * 32 + l] = entry(2 * row + 1, j, j == 0);
          }
      off += (size_t)mc * kc * 32;
    }
  } else {
    fill(off, ko, mo, Jo, 2, 1, -1);                                  // part O: k = 2l+1, b_j
    off += (size_t)mo * ko * 32;
  }
  if (two) {
    fill(off, kb, mb, Jb, 4, 0, H / 2);                               // part B: k = 4l, aa_j (self-paired at j = N/4)
    off += (size_t)mb * kb * 32;
    fill(off, kc, mc, Jc, 4, 2, -1);                                  // part C: k = 4l+2, ab_j
  } else {
    fill(off, kb, mb, Jb, 2, 0, N % 2 == 0 ? H : -1);                 // part B: k = 2l, a_j (self-paired at j = N/2)
  }
  *ko_out = ko;
  *kb_out = kb;
  *kc_out = kc;
}

namespace {

// map != 0 implies the split odd half (parts P, Q), map == 0 the unsplit one (part O)
int get_afrag(int dev, int n, int map, AFrag* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_pair(dev, 4 * n + map);
  auto it = g_afrag.find(key);
  if (it != g_afrag.end()) {
    *out = it->second;
    return FASTCHEB_OK;
  }
  AFrag a;
  std::vector<double> h;
  fit_afrag_host(n, map, map != 0, h, &a.ko, &a.kb, &a.kc);
  a.doubles = h.size();
  FC_CUDA(cudaMalloc(&a.d, h.size() * sizeof(double)));
  FC_CUDA(cudaMemcpy(a.d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  g_afrag[key] = a;
  *out = a;
  return FASTCHEB_OK;
}

template <typename R, int KO, int KB, int KC, int NT, bool E1, int MAP = 0, bool SPLIT = false, bool DYN = false>
cudaError_t launch(const FitDmmaParams& prm, int grid, int threads, size_t smem, cudaStream_t st, int* bps) {
  auto kern = fit_dmma_kernel<R, KO, KB, KC, NT, E1, MAP, SPLIT, DYN>;
  cudaError_t e = ensure_max_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  if (bps) return cached_occupancy(bps, reinterpret_cast<const void*>(kern), threads, smem);
  // programmatic dependent launch: see the head of fit_dmma_kernel
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, prm);
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

template <typename R>
cudaError_t dispatch(int ko, int kb, int kc, bool e1, int map, bool dyn, const FitDmmaParams& prm, int grid, int threads,
                     size_t smem, cudaStream_t st, int* bps) {
  if (ko == 8 && kb == 5 && kc == 4) {  // n = 65: tuned slot map of this precision, dynamic unit assignment
    constexpr int M = sizeof(R) == 4 ? 1 : 2;
    if (map != 0 && map != M) return cudaErrorInvalidValue;
#define FC_N65(MAPV, SPLITV)                                                                                        \
  return dyn ? (e1 ? launch<R, 8, 5, 4, 1, true, MAPV, SPLITV, true>(prm, grid, threads, smem, st, bps)            \
                   : launch<R, 8, 5, 4, 1, false, MAPV, SPLITV, true>(prm, grid, threads, smem, st, bps))          \
             : (e1 ? launch<R, 8, 5, 4, 1, true, MAPV, SPLITV, false>(prm, grid, threads, smem, st, bps)           \
                   : launch<R, 8, 5, 4, 1, false, MAPV, SPLITV, false>(prm, grid, threads, smem, st, bps))
    if (map != 0) {
      FC_N65(M, true);
    } else {
      FC_N65(0, false);
    }
#undef FC_N65
  }
#define FC_CASE(A, B, C)                                                                   \
  if (ko == A && kb == B && kc == C)                                                       \
    return e1 ? launch<R, A, B, C, 1, true>(prm, grid, threads, smem, st, bps)             \
              : launch<R, A, B, C, 1, false>(prm, grid, threads, smem, st, bps)
  // one fold (4 does not divide n-1): kb is ko or ko + 1; n <= 65 gives kb <= 9
  FC_CASE(1, 1, 0);
  FC_CASE(1, 2, 0);
  FC_CASE(2, 2, 0);
  FC_CASE(2, 3, 0);
  FC_CASE(3, 3, 0);
  FC_CASE(3, 4, 0);
  FC_CASE(4, 4, 0);
  FC_CASE(4, 5, 0);
  FC_CASE(5, 5, 0);
  FC_CASE(5, 6, 0);
  FC_CASE(6, 6, 0);
  FC_CASE(6, 7, 0);
  FC_CASE(7, 7, 0);
  FC_CASE(7, 8, 0);
  FC_CASE(8, 8, 0);
  FC_CASE(8, 9, 0);
  // two folds (n - 1 = 4t, t = 1..16): (ceil(t/2), ceil((t+1)/4), ceil(t/4))
  FC_CASE(1, 1, 1);
  FC_CASE(2, 1, 1);
  FC_CASE(2, 2, 1);
  FC_CASE(3, 2, 2);
  FC_CASE(4, 2, 2);
  FC_CASE(4, 3, 2);
  FC_CASE(5, 3, 3);
  FC_CASE(6, 3, 3);
  FC_CASE(6, 4, 3);
  FC_CASE(7, 4, 4);
  FC_CASE(8, 4, 4);
#undef FC_CASE
  return cudaErrorInvalidValue;
}

// maxabs[f] = max_e linemax[(f / U) * lines + (f % U) * E + e]
__global__ void __launch_bounds__(256) fit_max_kernel(const double* __restrict__ linemax, double* __restrict__ fitmax,
                                                      long long nfits, int U, int E, int lines) {
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < nfits; f += (long long)gridDim.x * blockDim.x) {
    const double* p = linemax + (f / U) * lines + (f % U) * E;
    double m = 0.0;
    for (int e = 0; e < E; ++e) m = fmax(m, p[e]);
    fitmax[f] = m;
  }
}

}  // namespace

int fit_dmma_1d(const DeviceInfo* di, int dev, int n, int E, long long howmany, int dtype, const void* vals,
                void* coefs, long long* bad_dev, long long idx_base, double* fit_max_dev, cudaStream_t st, long long* done) {
  *done = 0;
  if (n < 2 || n > 65 || E < 1) return FASTCHEB_OK;
  // launch shape: FASTCHEB_FIT_TUNE="warps,slots,nt" overrides the defaults (benchmark sweeps)
  // measured best on B200 (profiles/r01_fit_sweep.txt): Float64 is HBM-bound with 16 warps; Float32 units are half
  // the bytes, so 24 warps fit and the (then DMMA-bound) kernel wants them.  NT = 2 is not instantiated.
  int warps = dtype == FASTCHEB_F64 ? 16 : 24, S = 3, NT = 1, lag = 1;
  const long long dyn_threshold = 32;  // units per warp below which the static round-robin split loses to imbalance
  if (const char* t = getenv("FASTCHEB_FIT_TUNE")) {
    int a = 0, b = 0, c = 0, d = 1;
    const int got = sscanf(t, "%d,%d,%d,%d", &a, &b, &c, &d);
    if (got >= 3 && a >= 1 && a <= 24 && b >= 2 && a * b <= 128 && c == 1 && (d == 1 || d == 2) && b > d) {
      warps = a;
      S = b;
      NT = c;
      lag = d;
    }
  }
  const int rs = (int)real_size(dtype);
  const int lines = 8 * NT;
  if (E > lines) return FASTCHEB_OK;
  int U = lines / E;
  while (U >= 1 && ((long long)U * n * E * rs) % 16 != 0) --U;
  if (U < 1) return FASTCHEB_OK;
  if (((uintptr_t)vals | (uintptr_t)coefs) & 15) return FASTCHEB_OK;
  const long long nunits = howmany / U;
  if (nunits < 1) return FASTCHEB_OK;
  AFrag af;
  const int map = fit_slot_map(n, dtype);
  FC_TRY(get_afrag(dev, n, map, &af));
  FitDmmaParams prm;
  memset(&prm, 0, sizeof prm);
  prm.src = vals;
  prm.dst = coefs;
  prm.afrag = af.d;
  prm.nunits = nunits;
  prm.bad = bad_dev;
  prm.idx_base = idx_base;
  prm.n = n;
  prm.E = E;
  prm.U = U;
  prm.LU = U * E;
  prm.unitBytes = U * n * E * rs;
  prm.slotBytes = (prm.unitBytes + 127) / 128 * 128;
  const size_t afragBytes = (af.doubles * 8 + 127) / 128 * 128;
  const size_t budget = di->smem_optin - 1024;
  while (kFitHeaderBytes + afragBytes + (size_t)warps * S * prm.slotBytes > budget && S > lag + 1) --S;
  while (kFitHeaderBytes + afragBytes + (size_t)warps * S * prm.slotBytes > budget && warps > 1) --warps;
  prm.nslots = S;
  prm.lag = lag;
  const size_t smem = kFitHeaderBytes + afragBytes + (size_t)warps * S * prm.slotBytes;
  // few units per warp: draw them dynamically inside each CTA (n = 65 instantiations; FASTCHEB_FIT_DYN=0/1 overrides).
  // Measured: Float32 +7 % at 10^5 fits, +18 % at 3*10^4; Float64 unchanged at 10^5 and 10 % slower at 2^20, so it
  // keeps the static round-robin split.
  bool dyn = n == 65 && dtype == FASTCHEB_F32 && nunits < (long long)di->sms * warps * dyn_threshold;
  if (const char* t = getenv("FASTCHEB_FIT_DYN")) dyn = n == 65 && atoi(t) != 0;
  if (smem > budget) return FASTCHEB_OK;
  double* linemax = nullptr;
  if (fit_max_dev) {
    if (pool_alloc((void**)&linemax, (size_t)nunits * lines * sizeof(double), st) != cudaSuccess) {
      cudaGetLastError();
      return fail(FASTCHEB_ENOMEM, "fit line-maximum scratch");
    }
  }
  prm.linemax = linemax;
  const int threads = warps * 32;
  int bps = 0;
  cudaError_t e;
  auto go = [&](int* bp, int grid) {
    if (dtype == FASTCHEB_F64) return dispatch<double>(af.ko, af.kb, af.kc, E == 1, map, dyn, prm, grid, threads, smem, st, bp);
    return dispatch<float>(af.ko, af.kb, af.kc, E == 1, map, dyn, prm, grid, threads, smem, st, bp);
  };
  e = go(&bps, 0);
  if (e != cudaSuccess || bps < 1) {
    cudaGetLastError();
    if (linemax) cudaFreeAsync(linemax, st);
    return FASTCHEB_OK;  // not covered: the caller falls back to the generic kernel
  }
  const long long want = (nunits + warps - 1) / warps;
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)di->sms * bps));
  e = go(nullptr, grid);
  int rc = FASTCHEB_OK;
  if (e != cudaSuccess) rc = fail(FASTCHEB_ECUDA, "DCT-I tensor-core kernel launch failed: %s", cudaGetErrorString(e));
  if (rc == FASTCHEB_OK && fit_max_dev) {
    const long long nf = nunits * U;
    const int g2 = (int)std::max<long long>(1, std::min<long long>((nf + 255) / 256, (long long)di->sms * 8));
    fit_max_kernel<<<g2, 256, 0, st>>>(linemax, fit_max_dev, nf, U, E, lines);
    note_launch();
    if (cudaGetLastError() != cudaSuccess) rc = fail(FASTCHEB_ECUDA, "fit maximum kernel launch failed");
  }
  if (linemax) cudaFreeAsync(linemax, st);  // stream-ordered: the call stays asynchronous
  if (rc == FASTCHEB_OK) *done = nunits * U;
  return rc;
}

}  // namespace fastcheb
