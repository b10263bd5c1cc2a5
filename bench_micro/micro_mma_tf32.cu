// Issue rate of the warp-level mma.sync shapes on sm_100a: TF32 m16n8k8 and BF16 m16n8k16 (FP32 accumulate) against the
// FP64 m8n8k4 the fit kernels use.  Register operands only (no memory), ILP independent accumulator chains per warp.
// Prints one JSON line per (shape, warps per SM, ILP): instructions per SM per cycle and TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_tf32(float* out, int iters) {
  float c[ILP][4];
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_bf16(float* out, int iters) {
  float c[ILP][4];
  unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f003f00u, 0x3e803e80u, 0x3f403f40u}, b[2] = {0x3f803f80u, 0x3f003f00u + threadIdx.x};
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_f64(float* out, int iters) {
  double c[ILP][2];
  double a = 1.0 + threadIdx.x, b = 0.5;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

template <typename K>
void run(const char* name, K kern, int warps, int ilp, double flops_per_mma, int sms, double mhz, float* out) {
  const int iters = 4096;
  kern<<<sms, warps * 32>>>(out, 64);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<sms, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)sms * warps * ilp * iters;
  printf("{\"shape\": \"%s\", \"warps_per_sm\": %d, \"ilp\": %d, \"ms\": %.4f, \"mma_per_sm_per_cycle\": %.4f, \"tflops\": %.1f}\n", name, warps,
         ilp, ms, mmas / sms / (ms * 1e-3 * mhz * 1e6), mmas * flops_per_mma / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1e3;
  const int sms = pr.multiProcessorCount;
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  for (int warps : {4, 8, 16}) {
    run("tf32_m16n8k8", k_tf32<4>, warps, 4, 2.0 * 16 * 8 * 8, sms, mhz, out);
    run("tf32_m16n8k8", k_tf32<8>, warps, 8, 2.0 * 16 * 8 * 8, sms, mhz, out);
    run("bf16_m16n8k16", k_bf16<4>, warps, 4, 2.0 * 16 * 8 * 16, sms, mhz, out);
    run("bf16_m16n8k16", k_bf16<8>, warps, 8, 2.0 * 16 * 8 * 16, sms, mhz, out);
    run("f64_m8n8k4", k_f64<4>, warps, 4, 2.0 * 8 * 8 * 4, sms, mhz, out);
    run("f64_m8n8k4", k_f64<8>, warps, 8, 2.0 * 8 * 8 * 4, sms, mhz, out);
  }
  printf("{\"sms\": %d, \"clock_mhz_nominal\": %.0f}\n", sms, mhz);
  return 0;
}
