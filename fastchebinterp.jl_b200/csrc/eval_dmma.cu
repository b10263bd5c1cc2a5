// Host side of the DMMA evaluation path: plan (fragment re-layout) and launch.
#include <algorithm>
#include <cstring>
#include "eval_dmma_launch.h"

namespace fastcheb {

namespace {
typedef cudaError_t (*DmmaFn)(int, int, int, int, const DmmaParams&, int, int, size_t, cudaStream_t, int*);
DmmaFn dmma_fn(int ks, bool jac) {
  switch (ks) {
    case 4: return jac ? dmma_dispatch_ks4_jac : dmma_dispatch_ks4_val;
    case 8: return jac ? dmma_dispatch_ks8_jac : dmma_dispatch_ks8_val;
    case 16: return jac ? dmma_dispatch_ks16_jac : dmma_dispatch_ks16_val;
  }
  return nullptr;
}
}  // namespace

int dmma_plan(fastcheb_poly* p, cudaStream_t st) {
  DmmaPlan& pl = p->dmma;
  pl.ok = false;
  if (p->ndim < 2 || p->ndim > 4) return FASTCHEB_OK;  // Float32 polynomials use the same FP64 tensor-core kernels
  const int n1 = (int)p->dims[0];
  if (n1 < 5 || n1 - 1 > 64) return FASTCHEB_OK;
  long long M = 1;
  int tab = 0;
  for (int d = 1; d < p->ndim; ++d) {
    M *= p->dims[d];
    tab += (int)p->dims[d];
  }
  if (M < 8 || M > (1 << 28)) return FASTCHEB_OK;  // fewer than one column tile: the Clenshaw path is the right tool
  if (tab > 1024) return FASTCHEB_OK;              // weight tables would not fit shared memory
  pl.KS = n1 - 1 <= 16 ? 4 : (n1 - 1 <= 32 ? 8 : 16);
  pl.M = (int)M;
  pl.ntiles = (int)((M + 7) / 8);
  pl.tabStride = tab | 1;  // odd row stride: spreads the rows over the banks
  const int U = pl.KS * 32 + 8;
  size_t off = 0;
  for (int e0 = 0; e0 < p->E;) {
    const int left = p->E - e0;
    const int ec = left <= kDmmaMaxEC ? left : (left == 4 ? 2 : 3);
    DmmaGroup g;
    g.e0 = e0;
    g.ec = ec;
    g.offset = off;
    off += (size_t)pl.ntiles * (kDmmaMeta + ec * U);
    pl.groups.push_back(g);
    e0 += ec;
  }
  pl.frag_bytes = off * sizeof(double);
  FC_CUDA(cudaMalloc(&pl.d_frag, pl.frag_bytes));
  FragDims fd;
  memset(&fd, 0, sizeof fd);
  fd.nd = p->ndim - 1;
  for (int d = 0, o = 0; d < fd.nd; ++d) {
    fd.n[d] = (int)p->dims[d + 1];
    fd.toff[d] = o;
    o += fd.n[d];
  }
  for (const DmmaGroup& g : pl.groups) {
    const long long total = (long long)pl.ntiles * (kDmmaMeta + g.ec * U);
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)p->di->sms * 16));
    if (p->dtype == FASTCHEB_F64)
      relayout_frag_kernel<double><<<grid, 256, 0, st>>>((const double*)p->d_coefs, (double*)pl.d_frag + g.offset, total,
                                                        p->E, g.e0, g.ec, n1, pl.M, pl.KS, fd);
    else
      relayout_frag_kernel<float><<<grid, 256, 0, st>>>((const float*)p->d_coefs, (double*)pl.d_frag + g.offset, total,
                                                       p->E, g.e0, g.ec, n1, pl.M, pl.KS, fd);
    FC_CUDA(cudaGetLastError());
    note_launch();
  }
  pl.ok = true;
  return FASTCHEB_OK;
}

int dmma_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, int64_t v_stride, void* out_J,
              int64_t J_stride, bool jac, long long* bad_dev, long long idx_base, cudaStream_t st, int mode,
                const void* tan) {
  const DmmaPlan& pl = p->dmma;
  DmmaFn fn = dmma_fn(pl.KS, jac);
  if (!fn) return fail(FASTCHEB_EUNSUPPORTED, "no DMMA instantiation for KS = %d", pl.KS);
  const int U = pl.KS * 32 + 8;
  const int rows = dmma_mt(jac) * 8;
  for (const DmmaGroup& g : pl.groups) {
    DmmaParams prm;
    memset(&prm, 0, sizeof prm);
    prm.frag = (const double*)pl.d_frag + g.offset;
    prm.pts = pts;
    prm.out_v = out_v;
    prm.out_J = out_J;
    prm.bad = (&g == &pl.groups[0]) ? bad_dev : nullptr;
    prm.npts = npts;
    prm.idx_base = idx_base;
    prm.v_stride = v_stride;
    prm.J_stride = J_stride;
    prm.mode = mode;
    prm.tan = tan;
    prm.e0 = g.e0;
    prm.Etot = p->E;
    for (int d = 0; d < p->ndim; ++d) {
      prm.n[d] = (int)p->dims[d];
      prm.lb[d] = p->lb[d];
      prm.ub[d] = p->ub[d];
    }
    prm.ntiles = pl.ntiles;
    prm.M = pl.M;
    prm.tabStride = pl.tabStride;
    // warps per CTA: as many as the instantiation allows, fewer for small batches (spread over the SMs)
    int warps = dmma_max_warps(g.ec, pl.KS, jac);
    const long long per_sm = (npts + (long long)p->di->sms * rows - 1) / ((long long)p->di->sms * rows);
    if (per_sm < warps) warps = (int)std::max<long long>(1, per_sm);
    const size_t unitBytes = (size_t)(kDmmaMeta + g.ec * U) * 8;  // one tile record
    const size_t budget = p->di->smem_optin - 1024 - 128;
    // Stage geometry (measured on B200, profiles/r01_dmma_variants.jsonl): crossing a stage costs a ring
    // hand-off, so stages of 4 tile records beat 1-2 by 10+ points of FP64 peak; 3 stages of look-ahead
    // are enough once a stage holds 4 tiles.
    int tps = 0, nstage = 0;
    for (;; --warps) {
      const size_t tabBytes = (size_t)warps * rows * pl.tabStride * 8 * (jac ? 2 : 1);
      if (tabBytes + 3 * unitBytes <= budget) {
        const size_t room = budget - tabBytes;
        tps = (int)std::max<size_t>(1, std::min<size_t>(4, room / (3 * unitBytes)));
        tps = std::min(tps, pl.ntiles);
        nstage = (int)std::min<size_t>(4, room / (tps * unitBytes));
        if (nstage >= 3) break;
      }
      if (warps == 1) return fail(FASTCHEB_EUNSUPPORTED, "DMMA path: shared memory too small for this shape");
    }
    prm.tilesPerStage = tps;
    prm.nstage = nstage;
    const size_t smem = 128 + (size_t)nstage * tps * unitBytes + (size_t)warps * rows * pl.tabStride * 8 * (jac ? 2 : 1);
    const int threads = warps * 32;
    int bps = 1;
    cudaError_t e = fn(1, p->dtype, p->ndim, g.ec, prm, threads, 0, smem, st, &bps);
    if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "DMMA occupancy query failed: %s", cudaGetErrorString(e));
    if (bps < 1) return fail(FASTCHEB_ECUDA, "DMMA kernel does not fit an SM (smem %zu B)", smem);
    const long long pass_pts = (long long)warps * rows;
    const long long npass = (npts + pass_pts - 1) / pass_pts;
    const int grid = (int)std::min<long long>(npass, (long long)p->di->sms * bps);
    e = fn(0, p->dtype, p->ndim, g.ec, prm, threads, grid, smem, st, nullptr);
    if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "DMMA kernel launch failed: %s", cudaGetErrorString(e));
  }
  return FASTCHEB_OK;
}

}  // namespace fastcheb
