// Thin inline-PTX wrappers used by the sm_100a kernels: mbarrier, 1-D bulk async copy (TMA engine,
// SASS UBLKCP), FP64 tensor-core MMA (SASS DMMA).  No library dependencies.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fastcheb {

constexpr int kMaxNdim = 6;  // == FASTCHEB_MAX_NDIM

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make mbarrier initialisation visible to the async proxy
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-B aligned), completion on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start
// (CTA launch, constant set-up) while its predecessor in the stream drains; it must not touch memory the predecessor
// may access before grid_dep_wait() returns (all prerequisite grids complete, their writes visible).
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// exact IEEE helpers (no contraction regardless of compiler flags), one spelling for both precisions
__device__ __forceinline__ double rsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float rsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double rmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float rmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double rdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float rdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double rfma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float rfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// Division by the loop-invariant span ub - lb of the affine map z = ((x - lb) * 2) / (ub - lb) - 1 (eval.jl:65).
// With y = RN(1 / span) — one IEEE division per thread — Markstein's sequence
//     q = RN(a y),   r = a - span q  (exact in an FMA),   q' = RN(q + r y)
// returns RN(a / span), the IEEE quotient, for EVERY a, provided y is the correctly rounded reciprocal (it is), the
// significand of span is not all ones, and nothing leaves the normal range (span's exponent is checked; a <= 2 span
// here, and for a < 2^-60 span both quotients round z to -1).  3 operations instead of the ~12 + MUFU of a division,
// and still bit-identical to the oracle; bench_micro/micro_divconst.cu checks 1.6e11 (a, span) pairs against the
// IEEE division (none differ).  Otherwise: the plain division.
template <typename R>
struct SpanDiv {
  R span, y;
  bool fast;
  __device__ __forceinline__ void init(R lb, R ub) {
    span = rsub(ub, lb);
    y = rdiv(R(1), span);
    if constexpr (sizeof(R) == 8) {
      const unsigned long long u = (unsigned long long)__double_as_longlong((double)span);
      const unsigned e = (unsigned)(u >> 52) & 0x7ffu;
      fast = (u & 0xfffffffffffffull) != 0xfffffffffffffull && e > 1023u - 500u && e < 1023u + 500u;
    } else {
      const unsigned u = __float_as_uint((float)span);
      const unsigned e = (u >> 23) & 0xffu;
      fast = (u & 0x7fffffu) != 0x7fffffu && e > 127u - 60u && e < 127u + 60u;
    }
  }
  __device__ __forceinline__ R div(R a) const {
    if (fast) {
      const R q = rmul(a, y);
      const R r = rfma(-span, q, a);
      return rfma(r, y, q);
    }
    return rdiv(a, span);
  }
};

// D(8x8) += A(8x4, row) * B(4x8, col), FP64.  Lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4)+{0,1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

}  // namespace fastcheb
