// Host side of the resident N-d / long-line DCT-I fit (fit_nd.cuh): matrices, planning, launch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>
#include "fit_dmma_launch.h"
#include "fit_nd.cuh"

namespace fastcheb {
namespace {

struct NdMatrices {
  double* d = nullptr;  // the matrices of every distinct n of a shape, back to back
  int doubles = 0;
  std::vector<int> ns, off, ko, kb, kc;
};
std::mutex g_mu;
std::map<std::pair<int, std::vector<int>>, NdMatrices> g_mats;  // (device, sorted distinct n > 1)

int get_matrices(int dev, std::vector<int> ns, const NdMatrices** out) {
  std::sort(ns.begin(), ns.end());
  ns.erase(std::unique(ns.begin(), ns.end()), ns.end());
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_pair(dev, ns);
  auto it = g_mats.find(key);
  if (it != g_mats.end()) {
    *out = &it->second;
    return FASTCHEB_OK;
  }
  if (g_mats.size() >= 128) {  // a caller sweeping many shapes must not accumulate matrices for ever
    cudaDeviceSynchronize();
    for (auto& kv : g_mats) cudaFree(kv.second.d);
    g_mats.clear();
  }
  NdMatrices m;
  m.ns = ns;
  std::vector<double> all;
  for (int n : ns) {
    std::vector<double> h;
    int ko, kb, kc;
    fit_afrag_host(n, 0, true, h, &ko, &kb, &kc);  // twice-folded sizes: the odd half as parts P, Q (fit_nd.cuh)
    m.off.push_back((int)all.size());
    m.ko.push_back(ko);
    m.kb.push_back(kb);
    m.kc.push_back(kc);
    all.insert(all.end(), h.begin(), h.end());
  }
  m.doubles = (int)all.size();
  FC_CUDA(cudaMalloc(&m.d, std::max<size_t>(all.size(), 1) * sizeof(double)));
  FC_CUDA(cudaMemcpy(m.d, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice));
  auto ins = g_mats.emplace(key, std::move(m));
  *out = &ins.first->second;
  return FASTCHEB_OK;
}

template <typename R, int MTO, int MTB, int MTC, bool AFG = false, int SPEC = 0>
cudaError_t launch(const FitNdParams& prm, int grid, int threads, size_t smem, cudaStream_t st, int* bps) {
  auto kern = fit_nd_kernel<R, MTO, MTB, MTC, AFG, SPEC>;
  cudaError_t e = ensure_max_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  if (bps) return cached_occupancy(bps, reinterpret_cast<const void*>(kern), threads, smem);
  kern<<<grid, threads, smem, st>>>(prm);
  note_launch();
  return cudaGetLastError();
}

template <typename R>
cudaError_t launch_pipe(const FitNdParams& prm, int grid, int threads, size_t smem, cudaStream_t st, int* bps) {
  auto kern = fit_nd_pipe_kernel<R, 4, 4, 2>;
  cudaError_t e = ensure_max_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  if (bps) return cached_occupancy(bps, reinterpret_cast<const void*>(kern), threads, smem);
  kern<<<grid, threads, smem, st>>>(prm);
  note_launch();
  return cudaGetLastError();
}

// size class of the accumulators: 0: n <= 65 (any n), 1: n <= 129, 2: n <= 257 (sizes above 65: 4 | n-1 only, so that
// part O — MTO tiles — is needed for once-folded sizes n <= 65 only; twice-folded sizes use MTC tiles for P, Q, C)
template <typename R>
cudaError_t dispatch(int cls, const FitNdParams& prm, int grid, int threads, size_t smem, cudaStream_t st, int* bps) {
  switch (cls) {
    case 0: return launch<R, 4, 4, 2>(prm, grid, threads, smem, st, bps);
    case 17: return launch<R, 4, 4, 2, false, 17>(prm, grid, threads, smem, st, bps);  // matrices in registers
    case 33: return launch<R, 4, 4, 2, false, 33>(prm, grid, threads, smem, st, bps);
    case 65: return launch<R, 4, 4, 2, false, 65>(prm, grid, threads, smem, st, bps);
    case 1: return launch<R, 4, 5, 4>(prm, grid, threads, smem, st, bps);
    case 2: return launch<R, 4, 9, 8, true>(prm, grid, threads, smem, st, bps);  // matrices from global memory / L2
  }
  return cudaErrorInvalidValue;
}

}  // namespace

namespace {
template <typename R, int MTO, int MTB, int MTC, bool AFG>
cudaError_t launch_tile(const FitTileParams& prm, int grid, int threads, size_t smem, cudaStream_t st, int* bps) {
  auto kern = fit_nd_tile_kernel<R, MTO, MTB, MTC, AFG>;
  cudaError_t e = ensure_max_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  if (bps) return cached_occupancy(bps, reinterpret_cast<const void*>(kern), threads, smem);
  kern<<<grid, threads, smem, st>>>(prm);
  note_launch();
  return cudaGetLastError();
}
template <typename R>
cudaError_t dispatch_tile(int cls, const FitTileParams& prm, int grid, int threads, size_t smem, cudaStream_t st, int* bps) {
  switch (cls) {
    case 0: return launch_tile<R, 4, 4, 2, false>(prm, grid, threads, smem, st, bps);
    case 1: return launch_tile<R, 4, 5, 4, false>(prm, grid, threads, smem, st, bps);
    case 2: return launch_tile<R, 4, 9, 8, true>(prm, grid, threads, smem, st, bps);
  }
  return cudaErrorInvalidValue;
}

int size_class(int n) {
  if (n <= 65) return 0;
  if ((n - 1) % 4 != 0) return -1;  // the larger classes are sized for the twice-folded form
  return n <= 129 ? 1 : (n <= 257 ? 2 : -1);
}
}  // namespace

// An array too large for shared memory: dims 1..N-1 slab by slab (recursively), then the last dimension on strided
// tiles.  *handled = false (nothing launched) when some dimension is not covered.
static int fit_nd_split(const DeviceInfo* di, int dev, int ndim, const int64_t* dims, int K, int cplx, long long howmany, int dtype,
                        const void* vals, void* coefs, long long* bad_dev, long long idx_base, double* maxabs_dev, cudaStream_t st,
                        bool* handled, bool* need_max) {
  *handled = false;
  if (ndim < 2) return FASTCHEB_OK;
  const int E = K * (cplx ? 2 : 1);
  const size_t rs = real_size(dtype);
  for (int d = 0; d < ndim; ++d)
    if (dims[d] > 1 && size_class((int)dims[d]) < 0) return FASTCHEB_OK;
  const int nl = (int)dims[ndim - 1];
  long long inner = E;
  for (int d = 0; d + 1 < ndim; ++d) inner *= dims[d];
  if (inner * nl > (1LL << 31) || inner > (1LL << 30)) return FASTCHEB_OK;
  const size_t budget = std::min<size_t>(di->smem_optin, kSmemOptinMax) - 3072;
  const int cls = nl > 1 ? size_class(nl) : 0;
  const NdMatrices* mats = nullptr;
  size_t af_bytes = 0;
  int tile_cols = 0;
  if (nl > 1) {
    FC_TRY(get_matrices(dev, std::vector<int>{nl}, &mats));
    af_bytes = cls == 2 ? 0 : (((size_t)mats->doubles * 8 + 15) / 16) * 16;
    // a tile of nl rows x tile_cols columns (+1 pad column), two CTAs per SM when possible
    const long long room = ((long long)(budget / 2) - (long long)af_bytes) / (long long)(nl * rs) - 1;
    tile_cols = (int)std::min<long long>((inner + 15) / 16 * 16, room / 16 * 16);
    if (tile_cols < 16) return FASTCHEB_OK;
  }
  // leading dimensions: howmany * nl contiguous slabs (recursion: a slab that is itself too large splits again)
  bool lead = false, lead_max = false;
  FC_TRY(fit_nd(di, dev, ndim - 1, dims, K, cplx, howmany * nl, dtype, vals, coefs, bad_dev, idx_base, nullptr, st, &lead));
  (void)lead_max;
  if (!lead) return FASTCHEB_OK;  // nothing was launched
  *handled = true;
  if (nl > 1) {
    FitTileParams prm;
    memset(&prm, 0, sizeof prm);
    prm.data = coefs;
    prm.afrag = mats->d;
    prm.afragDoubles = mats->doubles;
    prm.narrays = howmany;
    prm.array_stride = inner * nl;
    prm.inner = inner;
    prm.row_stride = (int)inner;
    prm.tile_cols = tile_cols;
    prm.tiles_per_array = (int)((inner + tile_cols - 1) / tile_cols);
    prm.maxabs = (maxabs_dev && E == 1) ? maxabs_dev : nullptr;
    prm.dim.n = nl;
    prm.dim.stride = tile_cols + 1;
    prm.dim.KO = mats->ko[0];
    prm.dim.KB = mats->kb[0];
    prm.dim.KC = mats->kc[0];
    prm.dim.two = prm.dim.KC > 0;
    prm.dim.afrag = 0;
    const size_t smem = af_bytes + (size_t)nl * (tile_cols + 1) * rs + 16;
    int bps = 0;
    cudaError_t e = dtype == FASTCHEB_F64 ? dispatch_tile<double>(cls, prm, 0, 256, smem, st, &bps)
                                          : dispatch_tile<float>(cls, prm, 0, 256, smem, st, &bps);
    if (e != cudaSuccess || bps < 1)
      return fail(FASTCHEB_ECUDA, "tile DCT-I kernel not launchable: %s", cudaGetErrorString(e != cudaSuccess ? e : cudaErrorInvalidConfiguration));
    const long long ntiles = howmany * prm.tiles_per_array;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)di->sms * bps));
    e = dtype == FASTCHEB_F64 ? dispatch_tile<double>(cls, prm, grid, 256, smem, st, nullptr)
                              : dispatch_tile<float>(cls, prm, grid, 256, smem, st, nullptr);
    if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "tile DCT-I kernel launch failed: %s", cudaGetErrorString(e));
  }
  // vector / complex elements: the per-array maximum needs the element-wise infnorm: the caller's reduction kernel
  if (maxabs_dev && (E != 1 || nl == 1)) *need_max = true;
  return FASTCHEB_OK;
}

int fit_nd(const DeviceInfo* di, int dev, int ndim, const int64_t* dims, int K, int cplx, long long howmany, int dtype,
           const void* vals, void* coefs, long long* bad_dev, long long idx_base, double* maxabs_dev, cudaStream_t st,
           bool* handled, bool* need_max, double* stats_dev, const int* stats_hoff, double stats_tol, bool* stats_done) {
  *handled = false;
  if (stats_done) *stats_done = false;
  bool nm_local = false;
  if (!need_max) need_max = &nm_local;
  static const bool off = [] {  // experiments: FASTCHEB_FIT_ND=0 keeps the generic per-dimension kernels
    const char* e = getenv("FASTCHEB_FIT_ND");
    return e && atoi(e) == 0;
  }();
  if (off || howmany < 1 || ndim > kFitNdMaxDims) return FASTCHEB_OK;
  const int E = K * (cplx ? 2 : 1);
  const size_t rs = real_size(dtype);
  long long reals = E;
  int nmax = 1;
  std::vector<int> ns;
  for (int d = 0; d < ndim; ++d) {
    reals *= dims[d];
    if (reals > (1LL << 26)) return FASTCHEB_OK;
    nmax = std::max<int>(nmax, (int)dims[d]);
    if (dims[d] > 1) ns.push_back((int)dims[d]);
  }
  if (ns.empty()) return FASTCHEB_OK;  // nothing to transform: the caller copies
  int cls = nmax <= 65 ? 0 : (nmax <= 129 ? 1 : (nmax <= 257 ? 2 : -1));
  if (cls < 0) return FASTCHEB_OK;
  if (cls > 0)
    for (int n : ns)
      if (n > 65 && (n - 1) % 4 != 0) return FASTCHEB_OK;  // the larger classes are sized for the twice-folded form
  // batches of 1-D lines: a CTA takes LB lines at a time as one (n x LB) array whose second dimension is not
  // transformed, so that all of its warps have line groups to work on
  long long arrays = howmany, per_array = reals;
  const bool lines = ndim == 1;
  long long LB = 1;
  if (lines) {
    LB = std::min<long long>(howmany, cls == 0 ? 128 : 64);
    arrays = howmany / LB;
    per_array = reals * LB;
  }
  const NdMatrices* mats = nullptr;
  FC_TRY(get_matrices(dev, ns, &mats));
  // all transformed dimensions of size 17 or 33: the matrices live in registers (SPEC instantiations, fit_nd.cuh; 5 / 12
  // values per lane).  n = 65 (39 values: 166 registers, three CTAs per SM instead of four) measured no faster — 34.9 %
  // against 36-38 % of HBM on 65 x 65 — and is used only with FASTCHEB_FITND_SPEC=65 (experiments); 33^3 Float32: 14.8 %
  // against 13.3 %.
  static const int spec_env = [] {
    const char* e = getenv("FASTCHEB_FITND_SPEC");
    return e ? atoi(e) : -1;
  }();
  int spec = 0;
  if (cls == 0 && mats->ns.size() == 1 && spec_env != 0) {
    const int n0 = mats->ns[0];
    if (n0 == 17 || n0 == 33 || (n0 == 65 && spec_env == 65)) spec = n0;
  }
  const size_t af_bytes = (cls == 2 || spec) ? 0 : (((size_t)mats->doubles * 8 + 15) / 16) * 16;  // class 2 reads them from L2
  const size_t budget = std::min<size_t>(di->smem_optin, kSmemOptinMax) - 3072;
  if (lines) {  // as many lines per block as fit next to the matrices (a multiple of the 16- / 32-line MMA blocks)
    const long long room = ((long long)budget - (long long)af_bytes - 256) / (long long)(reals * rs);
    if (room < LB) LB = room / 32 * 32;
    if (LB < 32 && LB < howmany) return FASTCHEB_OK;
    arrays = howmany / LB;
    per_array = reals * LB;
  }
  // arrays too large for shared memory go slab by slab + strided tiles (33^3 Float64: 19 % of HBM).  Arrays that fit stay
  // resident even when only one CTA per SM holds them (33^3 Float32, 144 KB: 14.8 % resident, 7.6 % split);
  // FASTCHEB_FITND_SPLIT=bytes lowers the threshold (experiments)
  static const long long split_above = [] {
    const char* e = getenv("FASTCHEB_FITND_SPLIT");
    return (e && *e) ? atoll(e) : -1LL;
  }();
  const size_t resident_max = split_above >= 0 ? (size_t)split_above : budget;
  if (af_bytes + (size_t)per_array * rs + 1024 + 256 > std::min(budget, resident_max) && !(lines && af_bytes + (size_t)per_array * rs + 1280 <= budget)) {
    if (lines) return FASTCHEB_OK;
    return fit_nd_split(di, dev, ndim, dims, K, cplx, howmany, dtype, vals, coefs, bad_dev, idx_base, maxabs_dev, st, handled, need_max);
  }

  auto run = [&](const void* src, void* dst, long long narr, long long arr_reals, long long first_fit, long long fits_per_arr) -> int {
    FitNdParams prm;
    memset(&prm, 0, sizeof prm);
    prm.src = src;
    prm.dst = dst;
    prm.afrag = mats->d;
    prm.afragDoubles = mats->doubles;
    prm.howmany = narr;
    prm.bad = bad_dev;
    prm.idx_base = idx_base + first_fit * reals;
    prm.maxabs = maxabs_dev ? maxabs_dev + first_fit : nullptr;  // indexed by array (N-d) or by array * fits + fit (line blocks)
    prm.ndim = ndim;
    prm.E = E;
    prm.K = K;
    prm.cplx = cplx;
    prm.reals = (int)arr_reals;
    prm.fits = (int)fits_per_arr;
    prm.fit_reals = (int)reals;
    const bool want_stats = stats_dev && stats_hoff && howmany == 1 && narr == 1 && fits_per_arr == 1;
    if (want_stats) {
      prm.stats = stats_dev;
      prm.stats_tol = stats_tol;
      for (int d = 0; d < ndim; ++d) {
        prm.hoff[d] = stats_hoff[d];
        prm.hsum += (int)dims[d];
      }
    }
    int stride = E;
    for (int d = 0; d < ndim; ++d) {
      FitNdDim& fd = prm.dim[d];
      fd.n = (int)dims[d];
      fd.stride = stride;
      stride *= (int)dims[d];
      if (fd.n > 1) {
        const int i = (int)(std::lower_bound(mats->ns.begin(), mats->ns.end(), fd.n) - mats->ns.begin());
        fd.KO = mats->ko[i];
        fd.KB = mats->kb[i];
        fd.KC = mats->kc[i];
        fd.two = fd.KC > 0;
        fd.afrag = mats->off[i];
      }
    }
    // Small arrays in batches: the warp-specialised pipeline (loader warp, one consumer warp per line group, storer warp;
    // three buffers) — see fit_nd_pipe_kernel.  FASTCHEB_FITND_PIPE=0 keeps the lock-step kernel (experiments).
    {
      static const bool pipe_off = [] {  // measured slower than the lock-step kernel (30 % vs 38 % on 65 x 65): opt-in
        const char* e = getenv("FASTCHEB_FITND_PIPE");
        return !(e && atoi(e) == 1);
      }();
      const int gb0 = dtype == FASTCHEB_F64 ? 2 : 4;
      long long gmax = 0;
      for (int d = 0; d < ndim; ++d)
        if (dims[d] > 1) {
          const long long nlines = arr_reals / dims[d];
          gmax = std::max(gmax, (nlines + 8 * gb0 - 1) / (8 * gb0) * gb0);
        }
      const size_t bufB = (((size_t)arr_reals * rs + 15) / 16) * 16 + 16;
      const size_t af_pipe = (((size_t)mats->doubles * 8 + 15) / 16) * 16;
      const size_t smem_pipe = 1024 + af_pipe + (size_t)kFitPipeBufs * bufB;
      const long long rounds = (gmax + 13) / 14;
      if (!pipe_off && cls == 0 && fits_per_arr == 1 && rounds <= 2 && smem_pipe <= budget && narr >= 2LL * di->sms) {
        const int NC = (int)((gmax + rounds - 1) / rounds);
        int NS = 2;  // storer warps
        if (const char* t = getenv("FASTCHEB_FITND_STORERS")) NS = std::max(1, std::min(4, atoi(t)));
        if (NC + 1 + NS > 16) NS = 16 - NC - 1;
        prm.team_warps = NC;
        prm.teams = NS;
        prm.nbuf = kFitPipeBufs;
        const int threads = (NC + 1 + NS) * 32;
        int bps = 0;
        cudaError_t e = dtype == FASTCHEB_F64 ? launch_pipe<double>(prm, 0, threads, smem_pipe, st, &bps)
                                              : launch_pipe<float>(prm, 0, threads, smem_pipe, st, &bps);
        if (e == cudaSuccess && bps >= 1) {
          const int grid = (int)std::max<long long>(1, std::min<long long>(narr, (long long)di->sms * bps));
          e = dtype == FASTCHEB_F64 ? launch_pipe<double>(prm, grid, threads, smem_pipe, st, nullptr)
                                    : launch_pipe<float>(prm, grid, threads, smem_pipe, st, nullptr);
          if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "pipelined DCT-I kernel launch failed: %s", cudaGetErrorString(e));
          return FASTCHEB_OK;
        }
        cudaGetLastError();
      }
    }
    // Team geometry.  The passes of an array are separated by barriers of the team that owns it, so the line groups
    // of a pass should divide evenly among the team's warps (65 lines = 10 groups: 2 or 5 warps, not 4 or 8), and an
    // SM wants ~12-16 warps in flight: small arrays run several teams per CTA (one copy of the matrices), large ones
    // one team of up to 8 warps.  FASTCHEB_FITND_WARPS / _TEAMS / _NBUF override (experiments).
    const int gb = dtype == FASTCHEB_F64 ? 2 : 4;
    const size_t bufBytes = (((size_t)arr_reals * rs + 15) / 16) * 16 + 16;
    const size_t fixed = 1024 + af_bytes;
    const int max_warps = (cls == 0 && spec != 65) ? 16 : 8;  // launch bounds of the instantiation: 512 / 256 threads
    if (fixed + bufBytes > budget) return -1;
    auto waste = [&](int w) {
      long long slots = 0, work = 0;
      for (int d = 0; d < ndim; ++d)
        if (dims[d] > 1) {
          const long long nlines = arr_reals / dims[d];
          const long long groups = (nlines + 8 * gb - 1) / (8 * gb) * gb;
          slots += (groups + w - 1) / w * w * (long long)dims[d];  // rounds x warps, weighted by the work of a line
          work += groups * (long long)dims[d];
        }
      return (double)slots / (double)std::max<long long>(work, 1);
    };
    // Measured on B200 (65 x 65 Float64 batches, tools/probe_fit_nd.py): one team per CTA with 4-5 warps and as many
    // CTAs per SM as shared memory allows (4) gives 37-38 % of HBM bandwidth; 6 teams of 2 warps in one CTA 32 %, 8 warps
    // with two buffers 30-32 %.  So: one team per CTA, the warp count in 4..8 with the fewest idle barrier rounds
    // (small arrays: 4 or 5, so that several CTAs share an SM).
    int warps = 4, teams = 1, nbuf = 1;
    {
      // small arrays: 4 warps (128 threads x 128 registers: four CTAs share an SM; 5 warps would balance 10 line groups
      // exactly but leave room for three CTAs only — measured 34 % against 38 %)
      const bool small = fixed + bufBytes <= budget / 4 && narr >= di->sms;  // a single array: latency, not occupancy
      double best = 1e30;
      // large arrays (one CTA per SM): up to 16 warps — the n = 33 passes are short dependent chains of 12 DMMAs per line
      // group and want many groups in flight (33^3 Float32: 10.9 % of HBM with 8 warps, 13.3 % with 16)
      for (int w = small ? 4 : (fixed + bufBytes > budget / 2 ? max_warps : 8); w >= 4; --w) {
        const double cost = waste(w);
        if (cost < best - 1e-9) {
          best = cost;
          warps = w;
        }
      }
    }
    if (const char* t = getenv("FASTCHEB_FITND_WARPS")) {
      const int v = atoi(t);
      if (v >= 1 && v <= max_warps) {
        warps = v;
        teams = std::max<int>(1, std::min<int>({(int)((budget - fixed) / bufBytes), max_warps / v, 15}));
      }
    }
    if (const char* t = getenv("FASTCHEB_FITND_TEAMS")) {
      const int v = atoi(t);
      if (v >= 1 && v <= 15 && v * warps <= max_warps && fixed + (size_t)v * bufBytes <= budget) teams = v;
    }
    // a second buffer (the next array loads during the transform) only when the array is too large for a second CTA
    // on the SM anyway
    if (fixed + bufBytes > budget / 2 && fixed + (size_t)teams * 2 * bufBytes <= budget && narr > (long long)di->sms * teams) nbuf = 2;
    if (const char* t = getenv("FASTCHEB_FITND_NBUF")) {
      const int v = atoi(t);
      if (v == 1 || (v == 2 && fixed + (size_t)teams * 2 * bufBytes <= budget)) nbuf = v;
    }
    prm.nbuf = nbuf;
    prm.teams = teams;
    prm.team_warps = warps;
    if (teams != 1 || warps > 16) prm.stats = nullptr;  // the statistics use the CTA's header scratch
    const int threads = warps * teams * 32;
    const size_t smem = fixed + (size_t)teams * nbuf * bufBytes;
    int bps = 0;
    const int sel = spec ? spec : cls;
    cudaError_t e = dtype == FASTCHEB_F64 ? dispatch<double>(sel, prm, 0, threads, smem, st, &bps)
                                          : dispatch<float>(sel, prm, 0, threads, smem, st, &bps);
    if (e != cudaSuccess || bps < 1) {
      cudaGetLastError();
      if (getenv("FASTCHEB_DEBUG")) fprintf(stderr, "fit_nd: not launchable (%s, %d CTAs/SM, %zu B)\n", cudaGetErrorString(e), bps, smem);
      return -1;  // not launchable: the caller falls back
    }
    const int grid = (int)std::max<long long>(1, std::min<long long>((narr + teams - 1) / teams, (long long)di->sms * bps));
    e = dtype == FASTCHEB_F64 ? dispatch<double>(sel, prm, grid, threads, smem, st, nullptr)
                              : dispatch<float>(sel, prm, grid, threads, smem, st, nullptr);
    if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "resident DCT-I kernel launch failed: %s", cudaGetErrorString(e));
    if (prm.stats && stats_done) *stats_done = true;
    return FASTCHEB_OK;
  };

  if (!lines) {
    // Batches of SMALL arrays (33 x 33 Float64 = 8.7 KB: 5 line groups per pass on 4 warps, a barrier per pass and array):
    // a CTA takes B arrays at a time as one block — the passes enumerate the lines of the whole block (the arrays are one
    // more, untransformed, outermost dimension), so the warps see B times as many line groups per barrier.  B: blocks of
    // <= 36 KB (four CTAs per SM still fit), at most 16 arrays, and enough blocks left to fill the GPU.
    long long B = 1;
    static const int block_env = [] {  // experiments: FASTCHEB_FITND_BLOCK=1 disables, =n forces n arrays per block
      const char* e = getenv("FASTCHEB_FITND_BLOCK");
      return e ? atoi(e) : 0;
    }();
    const size_t abytes = (size_t)reals * rs;
    if (block_env > 0)
      B = std::min<long long>(block_env, howmany);
    else if (abytes <= 18 * 1024 && howmany >= 16LL * di->sms)
      B = std::max<long long>(1, std::min<long long>({(long long)(36 * 1024 / abytes), 16LL, howmany / (8LL * di->sms)}));
    if (af_bytes + (size_t)B * abytes + 1024 + 256 > budget) B = 1;
    if (B > 1) {
      const long long nb = howmany / B, rest = howmany - nb * B;
      int rc = run(vals, coefs, nb, B * reals, 0, B);
      if (rc == -1) {
        B = 1;  // not launchable as blocks: one array per CTA below
      } else {
        FC_TRY(rc);
        if (rest > 0) {
          rc = run((const char*)vals + (size_t)nb * B * reals * rs, (char*)coefs + (size_t)nb * B * reals * rs, 1, rest * reals,
                   nb * B, rest);
          if (rc == -1) return fail(FASTCHEB_ECUDA, "resident DCT-I kernel: tail block not launchable");
          FC_TRY(rc);
        }
        *handled = true;
        return FASTCHEB_OK;
      }
    }
    const int rc = run(vals, coefs, howmany, reals, 0, 1);
    if (rc == -1) return FASTCHEB_OK;
    FC_TRY(rc);
    *handled = true;
    return FASTCHEB_OK;
  }
  // 1-D: blocks of LB lines, then the remaining lines as one smaller block
  long long done = 0;
  if (arrays > 0) {
    const int rc = run(vals, coefs, arrays, per_array, 0, LB);
    if (rc == -1) return FASTCHEB_OK;
    FC_TRY(rc);
    done = arrays * LB;
  }
  if (done < howmany) {
    const long long rest = howmany - done;
    const int rc = run((const char*)vals + (size_t)done * reals * rs, (char*)coefs + (size_t)done * reals * rs, 1, reals * rest, done, rest);
    if (rc == -1) return fail(FASTCHEB_ECUDA, "resident DCT-I kernel: tail block not launchable");
    FC_TRY(rc);
  }
  *handled = true;
  return FASTCHEB_OK;
}

}  // namespace fastcheb
