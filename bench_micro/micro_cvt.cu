// Throughput of Float32 <-> Float64 conversions (SASS F2F.F64.F32 / F2F.F32.F64, XU pipe) against an integer
// widening (ALU pipe): the question behind the Float32 path of fit_dmma_kernel.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned long long* out, unsigned seed) {
  unsigned x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 8 + i;
  unsigned long long acc = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i] = x[i] * 1664525u + 1013904223u;
      const unsigned b = (x[i] & 0x3fffffffu) | 0x20000000u;  // a normal finite float
      if (MODE == 0) {
        acc ^= (unsigned long long)__double_as_longlong((double)__uint_as_float(b));
      } else if (MODE == 1) {
        const unsigned hi = ((b & 0x7fffffffu) >> 3) + 0x38000000u + (b & 0x80000000u);
        acc ^= ((unsigned long long)hi << 32) | (unsigned long long)(b << 29);
      } else if (MODE == 2) {  // baseline: the LCG and xor only
        acc ^= b;
      } else {  // F64 -> F32
        const double d = __longlong_as_double(((unsigned long long)(0x3ff00000u | (b >> 12)) << 32) | b);
        acc ^= __float_as_uint((float)d);
      }
    }
  }
  if (acc == 0x1234567ull) out[0] = acc;
}
int main() {
  unsigned long long* out;
  cudaMalloc(&out, 8);
  int sms = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[] = {"f2f_f64_f32", "int_widen", "baseline_lcg", "f2f_f32_f64"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<sms * 8, 256>>>(out, 1);
      if (mode == 1) k<1><<<sms * 8, 256>>>(out, 1);
      if (mode == 2) k<2><<<sms * 8, 256>>>(out, 1);
      if (mode == 3) k<3><<<sms * 8, 256>>>(out, 1);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr_per_smsp = (double)8 * 8 * ITERS * 8 / 4;  // 8 CTAs x 8 warps per SM, 8 conversions per iteration
    printf("{\"test\":\"%s\",\"ms\":%.4f,\"cycles_per_warp_conversion_per_smsp\":%.2f}\n", names[mode], ms,
           ms * 1e-3 * 1.965e9 / warp_instr_per_smsp);
  }
  return 0;
}
