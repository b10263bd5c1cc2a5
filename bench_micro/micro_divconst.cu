// Brute-force check of the divisor-invariant division used by the evaluation kernels' affine map:
//   y = RN(1/b) (one IEEE division), then for every a:  q = RN(a*y), r = fma(-b, q, a) (exact), q' = fma(r, y, q)
// Markstein's theorem: q' = RN(a/b) whenever y is the correctly rounded reciprocal and b's significand is not all ones.
// Counts mismatches against the IEEE division over random and adversarial (a, b), Float64 and Float32.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
template <typename R> struct Ops;
template <> struct Ops<double> {
  using U = uint64_t;
  static __device__ double from_bits(U u) { return __longlong_as_double((long long)u); }
  static __device__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static constexpr int MB = 52;
  static constexpr U ONE = 0x3ff0000000000000ULL;
};
template <> struct Ops<float> {
  using U = uint32_t;
  static __device__ float from_bits(U u) { return __uint_as_float(u); }
  static __device__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static constexpr int MB = 23;
  static constexpr U ONE = 0x3f800000u;
};

// mode 0: random b significand; 1: b significand within 64 ulps of all ones (excluding all ones); 2: within 64 ulps above 1.0;
// 3: b significand all ones (expected to fail sometimes: reported separately)
template <typename R>
__global__ void check(unsigned long long* bad, unsigned long long* tested, int mode, uint64_t seed, int per_thread) {
  using O = Ops<R>;
  using U = typename O::U;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const U mant_mask = ((U)1 << O::MB) - 1;
  // one divisor per warp-sized group of 64 threads, exponent in [-40, 40]
  const uint64_t hb = mix(seed + (tid >> 6) * 0x9e3779b97f4a7c15ULL);
  U mant = (U)hb & mant_mask;
  if (mode == 1) mant = mant_mask - 1 - (U)(hb % 64);
  if (mode == 2) mant = (U)(hb % 64);
  if (mode == 3) mant = mant_mask;
  const int eb = (int)((hb >> 56) % 81) - 40;
  const U bbits = (U)(O::ONE + ((U)(eb + 0) << O::MB)) + 0;  // exponent field
  const R b = O::from_bits((U)((O::ONE & ~mant_mask) + ((U)((long long)eb) << O::MB)) | mant);
  (void)bbits;
  const R y = O::div((R)1, b);
  unsigned long long nbad = 0;
  uint64_t h = mix(seed ^ (tid * 0xd1342543de82ef95ULL));
  for (int i = 0; i < per_thread; ++i) {
    h = mix(h + i);
    // a = b * t with t in [0, 2]: random significand, exponent eb-30 .. eb+1; every 8th sample near a multiple of b
    U am = (U)h & mant_mask;
    int ea = eb + 1 - (int)((h >> 58) % 32);
    R a = O::from_bits((U)((O::ONE & ~mant_mask) + ((U)((long long)ea) << O::MB)) | am);
    if ((i & 7) == 0) a = O::mul(b, (R)(1 + (int)((h >> 40) % 7)) * (R)0.25);  // products of b: exact or nearly exact quotients
    const R q = O::mul(a, y);
    const R r = O::fma(-b, q, a);
    const R q1 = O::fma(r, y, q);
    const R ref = O::div(a, b);
    if (!(q1 == ref)) ++nbad;
  }
  atomicAdd(bad, nbad);
  atomicAdd(tested, (unsigned long long)per_thread);
}

int main() {
  unsigned long long *bad, *tested;
  cudaMalloc(&bad, 8); cudaMalloc(&tested, 8);
  const char* modes[] = {"random_b", "b_near_all_ones", "b_near_one", "b_all_ones(expected exceptions)"};
  for (int f32 = 0; f32 < 2; ++f32)
    for (int mode = 0; mode < 4; ++mode) {
      cudaMemset(bad, 0, 8); cudaMemset(tested, 0, 8);
      for (int rep = 0; rep < 8; ++rep) {
        if (f32) check<float><<<148 * 16, 256>>>(bad, tested, mode, 1234567 + 977 * rep, 4096);
        else check<double><<<148 * 16, 256>>>(bad, tested, mode, 7654321 + 977 * rep, 4096);
      }
      unsigned long long hb = 0, ht = 0;
      cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&ht, tested, 8, cudaMemcpyDeviceToHost);
      printf("{\"test\":\"divconst_%s_%s\",\"pairs\":%llu,\"mismatches\":%llu}\n", f32 ? "f32" : "f64", modes[mode], ht, hb);
    }
  return 0;
}
