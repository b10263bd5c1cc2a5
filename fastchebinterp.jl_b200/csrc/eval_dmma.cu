// Host side of the DMMA evaluation path: plan (fragment re-layout) and launch.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "eval_dmma_launch.h"

namespace fastcheb {

namespace {
typedef cudaError_t (*DmmaFn)(int, int, int, int, const DmmaParams&, int, int, size_t, cudaStream_t, int*);
DmmaFn dmma_fn(int ks, bool jac) {
  switch (ks) {
    case 4: return jac ? dmma_dispatch_ks4_jac : dmma_dispatch_ks4_val;
    case 8: return jac ? dmma_dispatch_ks8_jac : dmma_dispatch_ks8_val;
    case 12: return jac ? dmma_dispatch_ks12_jac : dmma_dispatch_ks12_val;
    case 16: return jac ? dmma_dispatch_ks16_jac : dmma_dispatch_ks16_val;
  }
  return nullptr;
}
typedef cudaError_t (*Dmma2dFn)(int, int, const Dmma2dParams&, int, int, size_t, cudaStream_t);
Dmma2dFn dmma2d_fn(int ks, bool jac) {
  switch (ks) {
    case 4: return jac ? dmma2d_dispatch_ks4_jac : dmma2d_dispatch_ks4_val;
    case 8: return jac ? dmma2d_dispatch_ks8_jac : dmma2d_dispatch_ks8_val;
    case 12: return jac ? dmma2d_dispatch_ks12_jac : dmma2d_dispatch_ks12_val;
    case 16: return jac ? dmma2d_dispatch_ks16_jac : dmma2d_dispatch_ks16_val;
  }
  return nullptr;
}

// Launch geometry of the resident 2-D kernel for one component group; warps == 0: does not apply (the weight
// tables hold k2 <= 4 KS + 1, and the tile records plus the per-warp tables must fit shared memory).
struct Plan2d {
  int warps = 0, nmma = 0, ntail = 0;
  size_t smem = 0;
};
Plan2d plan2d(const fastcheb_poly* p, int ec, bool jac, int64_t npts) {
  Plan2d q;
  const DmmaPlan& pl = p->dmma;
  const int n2 = (int)p->dims[1];
  if (p->ndim != 2 || n2 > 4 * pl.KS + 2 || p->algo == FASTCHEB_ALGO_DMMA_RING) return q;
  const int rem = n2 % 8;
  q.nmma = (rem == 1 || rem == 2) ? n2 / 8 : pl.ntiles;
  q.ntail = n2 - 8 * q.nmma > 0 ? n2 - 8 * q.nmma : 0;
  const int rows = jac ? 8 : 16;
  // fewer warps for small batches (spread over the SMs), as in the ring kernel
  static const int tuned = [] {  // experiments only: FASTCHEB_2D_WARPS=n caps the warps per CTA
    const char* e = getenv("FASTCHEB_2D_WARPS");
    const int v = e ? atoi(e) : 0;
    return v >= 1 && v <= kDmma2dThreads / 32 ? v : kDmma2dThreads / 32;
  }();
  int warps = tuned;
  const long long per_sm = (npts + (long long)p->di->sms * rows - 1) / ((long long)p->di->sms * rows);
  if (per_sm < warps) warps = (int)std::max<long long>(1, per_sm);
  const size_t budget = std::min<size_t>(p->di->smem_optin, kSmemOptinMax);
  for (; warps >= 1; --warps)
    if (dmma2d_smem(ec, pl.KS, jac, pl.ntiles, warps) <= budget) break;
  if (warps < 1 || (warps < 4 && per_sm >= 4)) return q;  // too few warps to hide the prologue: use the ring kernel
  q.warps = warps;
  q.smem = dmma2d_smem(ec, pl.KS, jac, pl.ntiles, warps);
  return q;
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
  pl.KS = n1 - 1 <= 16 ? 4 : (n1 - 1 <= 32 ? 8 : (n1 - 1 <= 48 ? 12 : 16));
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
  FC_CUDA(pool_alloc(&pl.d_frag, pl.frag_bytes, st));
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
    if (const Plan2d q = plan2d(p, g.ec, jac, npts); q.warps > 0) {  // 2-D, tile records resident in shared memory
      Dmma2dParams prm;
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
      prm.ntiles = pl.ntiles;
      prm.nmma = q.nmma;
      prm.ntail = q.ntail;
      for (int d = 0; d < 2; ++d) {
        prm.lb[d] = p->lb[d];
        prm.ub[d] = p->ub[d];
      }
      const long long nunits = (npts + rows - 1) / rows;
      // every CTA's warp 0 owns at least one unit (it is the warp that waits for the CTA's bulk copy)
      const int grid = (int)std::min<long long>((nunits + q.warps - 1) / q.warps, (long long)p->di->sms);
      cudaError_t e = dmma2d_fn(pl.KS, jac)(p->dtype, g.ec, prm, q.warps * 32, grid, q.smem, st);
      if (e != cudaSuccess) return fail(FASTCHEB_ECUDA, "2-D DMMA kernel launch failed: %s", cudaGetErrorString(e));
      continue;
    }
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
