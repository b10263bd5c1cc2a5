// FP64 / FP32 pipe microbenchmarks for B200 (sm_100a).
//
// Purpose: measure the roofline denominators that MEASURED_PEAKS.json does not carry
// (FP64 DFMA peak, FP64 DMMA peak, FP32 FFMA peak) and the design inputs for the
// evaluation kernel (does DMMA overlap with DFMA?  what does a broadcast LDS cost
// next to a DFMA?).  Prints one JSON object per test on stdout.
//
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o micro_fp64 micro_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

// ---------------------------------------------------------------- DFMA
template <int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
  double x[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------- FFMA
template <int CH>
__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
  float x[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-6f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i];
  if (s == 12345.678f) out[0] = s;
}

// ---------------------------------------------------------------- DMMA m8n8k4
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// m16n8k8: A 4 regs, B 2 regs, C 4 regs
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
// m16n8k4: A 2 regs, B 1 reg, C 4 regs
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(b));
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
  double d0[CH], d1[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { d0[i] = threadIdx.x * 1e-9; d1[i] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma884(d0[i], d1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += d0[i] + d1[i];
  if (s == 12345.678) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, double a, double b) {
  double d[CH][4];
  double af[4] = {a, a * 0.5, a * 0.25, a * 0.125};
  double bf[2] = {b, b * 0.5};
#pragma unroll
  for (int i = 0; i < CH; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma1688(d[i], af, bf);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 12345.678) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma1684(double* out, double a, double b) {
  double d[CH][4];
  double af[2] = {a, a * 0.5};
#pragma unroll
  for (int i = 0; i < CH; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma1684(d[i], af, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------- DMMA + DFMA in the same warp
// NF DFMAs per thread per DMMA.  If the pipes are shared the time adds, else it overlaps.
template <int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, double a, double b) {
  double d0[4], d1[4], x[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { d0[i] = threadIdx.x * 1e-9; d1[i] = i; }
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dmma884(d0[i], d1[i], a, b);
#pragma unroll
      for (int j = 0; j < NF; ++j) x[(i * NF + j) & 7] = fma(x[(i * NF + j) & 7], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += d0[i] + d1[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

// As k_mix, but every DFMA consumes the accumulator of a DMMA chain (like the weighted fold of the evaluation kernels):
// GROUP DMMAs of one chain, then NF DFMAs that read that chain's result.
template <int GROUP, int NF>
__global__ void __launch_bounds__(256) k_mix_dep(double* out, double a, double b) {
  double d0[4], d1[4], x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { d0[i] = threadIdx.x * 1e-9; d1[i] = i; x[i] = i; }
  for (int it = 0; it < ITERS / GROUP; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int g = 0; g < GROUP; ++g) dmma884(d0[i], d1[i], a, b);
#pragma unroll
      for (int j = 0; j < NF; ++j) x[i] = fma(j & 1 ? d1[i] : d0[i], a, x[i]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += d0[i] + d1[i] + x[i];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------- DFMA fed by broadcast LDS
// Each DFMA group consumes coefficients loaded from shared memory at a warp-uniform address.
// W = doubles per LDS (1 -> LDS.64, 2 -> LDS.128), P = points (independent chains) per thread
// that reuse each loaded coefficient.
template <int W, int P>
__global__ void __launch_bounds__(256) k_lds_dfma(double* out, double a, int n) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = 1.0 / (1 + i);
  __syncthreads();
  double acc[P][4];
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = threadIdx.x * 1e-9 + p;
  double t[P];
#pragma unroll
  for (int p = 0; p < P; ++p) t[p] = a + p;
  for (int rep = 0; rep < ITERS / 64; ++rep) {
#pragma unroll 4
    for (int i = 0; i < n; i += 4 * W) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (W == 2) {
          double2 c = *reinterpret_cast<const double2*>(&sm[i + 2 * q]);
#pragma unroll
          for (int p = 0; p < P; ++p) { acc[p][q] = fma(c.x, t[p], acc[p][q]); acc[p][q] = fma(c.y, t[p], acc[p][q]); }
        } else {
          double c = sm[i + q];
#pragma unroll
          for (int p = 0; p < P; ++p) acc[p][q] = fma(c, t[p], acc[p][q]);
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) s += acc[p][q];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------- Clenshaw-style DFMA+DADD chain
template <int CH>
__global__ void __launch_bounds__(256) k_clenshaw(double* out, double a, double b) {
  double b1[CH], b2[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { b1[i] = threadIdx.x * 1e-9 + i; b2[i] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) { double nb = fma(a, b1[i], b) - b2[i]; b2[i] = b1[i]; b1[i] = nb; }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += b1[i] + b2[i];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
static float time_kernel(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("{\"test\":\"device\",\"name\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", prop.name, sms, clk_khz);
  double* out; CK(cudaMalloc(&out, 1024));
  const int threads = 256;
  const int reps = 20;

  auto report = [&](const char* name, double flops_per_launch, float ms, const char* extra) {
    printf("{\"test\":\"%s\",\"ms\":%.4f,\"tflops\":%.3f%s}\n", name, ms, flops_per_launch / (ms * 1e-3) / 1e12, extra);
    fflush(stdout);
  };

  for (int bps : {1, 2, 4, 8}) {
    int grid = sms * bps;
    double thr = (double)grid * threads;
    char ex[64]; snprintf(ex, sizeof ex, ",\"blocks_per_sm\":%d", bps);
    report("dfma_ch8", thr * 8 * ITERS * 2.0, time_kernel([&] { k_dfma<8><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), ex);
  }
  {
    int grid = sms * 4; double thr = (double)grid * threads;
    report("dfma_ch2", thr * 2 * ITERS * 2.0, time_kernel([&] { k_dfma<2><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("dfma_ch4", thr * 4 * ITERS * 2.0, time_kernel([&] { k_dfma<4><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("dfma_ch16", thr * 16 * ITERS * 2.0, time_kernel([&] { k_dfma<16><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), "");
    // single-warp-per-SMSP latency probe: 1 block of 128 threads per SM, 1 chain -> latency-bound
    report("dfma_latency_1chain_4warps", (double)sms * 128 * 1 * ITERS * 2.0,
           time_kernel([&] { k_dfma<1><<<sms, 128>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("ffma_ch8", thr * 8 * ITERS * 2.0, time_kernel([&] { k_ffma<8><<<grid, threads>>>((float*)out, 1.0000001f, 1e-9f); }, reps), "");
    report("ffma_ch16", thr * 16 * ITERS * 2.0, time_kernel([&] { k_ffma<16><<<grid, threads>>>((float*)out, 1.0000001f, 1e-9f); }, reps), "");
    // Clenshaw: count algorithmic flops = 2 per step (the FMA), executed = 3
    report("clenshaw_ch4_alg2flop", thr * 4 * ITERS * 2.0, time_kernel([&] { k_clenshaw<4><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("clenshaw_ch8_alg2flop", thr * 8 * ITERS * 2.0, time_kernel([&] { k_clenshaw<8><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), "");
  }
  for (int bps : {1, 2, 4}) {
    int grid = sms * bps; double warps = (double)grid * threads / 32;
    char ex[64]; snprintf(ex, sizeof ex, ",\"blocks_per_sm\":%d", bps);
    report("dmma884_ch4", warps * 4 * ITERS * 2.0 * 8 * 8 * 4, time_kernel([&] { k_dmma884<4><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), ex);
    report("dmma884_ch8", warps * 8 * ITERS * 2.0 * 8 * 8 * 4, time_kernel([&] { k_dmma884<8><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), ex);
    report("dmma1688_ch4", warps * 4 * ITERS * 2.0 * 16 * 8 * 8, time_kernel([&] { k_dmma1688<4><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), ex);
    report("dmma1684_ch4", warps * 4 * ITERS * 2.0 * 16 * 8 * 4, time_kernel([&] { k_dmma1684<4><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), ex);
  }
  {  // GROUP-DMMA chains each followed by NF dependent DFMAs; pipe_cycles = 16 per DMMA + 2 per DFMA if nothing is lost
    int grid = sms * 2; double thr = (double)grid * threads; double warps = thr / 32;
    auto rep2 = [&](const char* name, int group, int nf, float ms) {
      const double dm = warps * 4 * (ITERS / group) * group, df = warps * 4 * (ITERS / group) * nf;
      const double ideal_cycles = (dm * 16 + df * 2) / (sms * 4.0);
      printf("{\"test\":\"%s\",\"ms\":%.4f,\"pipe_busy_if_no_loss\":%.3f,\"extra_cycles_per_dfma\":%.2f}\n", name, ms,
             ideal_cycles / (ms * 1e-3 * 1.965e9), (ms * 1e-3 * 1.965e9 - ideal_cycles) * (sms * 4.0) / df);
      fflush(stdout);
    };
    rep2("mixdep_g8_f2", 8, 2, time_kernel([&] { k_mix_dep<8, 2><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps));
    rep2("mixdep_g8_f4", 8, 4, time_kernel([&] { k_mix_dep<8, 4><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps));
    rep2("mixdep_g16_f4", 16, 4, time_kernel([&] { k_mix_dep<16, 4><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps));
    rep2("mixdep_g16_f8", 16, 8, time_kernel([&] { k_mix_dep<16, 8><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps));
    rep2("mixdep_g4_f2", 4, 2, time_kernel([&] { k_mix_dep<4, 2><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps));
  }
  // DMMA dependent-issue latency: ONE warp per SM sub-partition (128 threads per SM), CH independent accumulator chains.
  // Throughput saturates once CH * 16 cycles (the issue interval of an m8n8k4 DMMA per sub-partition) covers the latency.
  {
    const double warps = (double)sms * 4;
    report("dmma884_1warp_per_smsp_ch1", warps * 1 * ITERS * 512.0, time_kernel([&] { k_dmma884<1><<<sms, 128>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("dmma884_1warp_per_smsp_ch2", warps * 2 * ITERS * 512.0, time_kernel([&] { k_dmma884<2><<<sms, 128>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("dmma884_1warp_per_smsp_ch3", warps * 3 * ITERS * 512.0, time_kernel([&] { k_dmma884<3><<<sms, 128>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("dmma884_1warp_per_smsp_ch4", warps * 4 * ITERS * 512.0, time_kernel([&] { k_dmma884<4><<<sms, 128>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("dmma884_1warp_per_smsp_ch6", warps * 6 * ITERS * 512.0, time_kernel([&] { k_dmma884<6><<<sms, 128>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("dmma884_1warp_per_smsp_ch8", warps * 8 * ITERS * 512.0, time_kernel([&] { k_dmma884<8><<<sms, 128>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("dmma884_2warp_per_smsp_ch2", 2 * warps * 2 * ITERS * 512.0, time_kernel([&] { k_dmma884<2><<<sms, 256>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("dmma884_2warp_per_smsp_ch4", 2 * warps * 4 * ITERS * 512.0, time_kernel([&] { k_dmma884<4><<<sms, 256>>>(out, 1.0000001, 1e-9); }, reps), "");
  }
  {
    int grid = sms * 4; double thr = (double)grid * threads; double warps = thr / 32;
    // report total flops (dmma + dfma); compare with dmma-only and dfma-only to see overlap
    report("mix_dmma_1dfma", warps * 4 * ITERS * 512.0 + thr * 4 * 1 * ITERS * 2.0, time_kernel([&] { k_mix<1><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("mix_dmma_2dfma", warps * 4 * ITERS * 512.0 + thr * 4 * 2 * ITERS * 2.0, time_kernel([&] { k_mix<2><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("mix_dmma_4dfma", warps * 4 * ITERS * 512.0 + thr * 4 * 4 * ITERS * 2.0, time_kernel([&] { k_mix<4><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), "");
    report("mix_dmma_8dfma", warps * 4 * ITERS * 512.0 + thr * 4 * 8 * ITERS * 2.0, time_kernel([&] { k_mix<8><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps), "");
  }
  {
    int n = 4096;  // doubles in smem (32 KB)
    int grid = sms * 4; double thr = (double)grid * threads;
    double steps = (double)(ITERS / 64) * n;  // coefficient uses per thread per point
    size_t smem = n * sizeof(double);
    report("lds64_dfma_p1", thr * steps * 1 * 2.0, time_kernel([&] { k_lds_dfma<1, 1><<<grid, threads, smem>>>(out, 1.0000001, n); }, reps), "");
    report("lds64_dfma_p2", thr * steps * 2 * 2.0, time_kernel([&] { k_lds_dfma<1, 2><<<grid, threads, smem>>>(out, 1.0000001, n); }, reps), "");
    report("lds64_dfma_p4", thr * steps * 4 * 2.0, time_kernel([&] { k_lds_dfma<1, 4><<<grid, threads, smem>>>(out, 1.0000001, n); }, reps), "");
    report("lds128_dfma_p1", thr * steps * 1 * 2.0, time_kernel([&] { k_lds_dfma<2, 1><<<grid, threads, smem>>>(out, 1.0000001, n); }, reps), "");
    report("lds128_dfma_p2", thr * steps * 2 * 2.0, time_kernel([&] { k_lds_dfma<2, 2><<<grid, threads, smem>>>(out, 1.0000001, n); }, reps), "");
    report("lds128_dfma_p4", thr * steps * 4 * 2.0, time_kernel([&] { k_lds_dfma<2, 4><<<grid, threads, smem>>>(out, 1.0000001, n); }, reps), "");
  }
  // sustained: run DFMA back-to-back for ~3 s to see the clock under FP64 load
  {
    int grid = sms * 4; double thr = (double)grid * threads;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int n = 0; float ms = 0;
    CK(cudaEventRecord(e0));
    do {
      for (int i = 0; i < 50; ++i) k_dfma<8><<<grid, threads>>>(out, 1.0000001, 1e-9);
      n += 50;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
    } while (ms < 3000);
    report("dfma_ch8_sustained3s", thr * 8 * ITERS * 2.0, ms / n, "");
    n = 0;
    CK(cudaEventRecord(e0));
    do {
      for (int i = 0; i < 50; ++i) k_dmma884<8><<<grid, threads>>>(out, 1.0000001, 1e-9);
      n += 50;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
    } while (ms < 3000);
    report("dmma884_ch8_sustained3s", thr / 32 * 8 * ITERS * 512.0, ms / n, "");
  }
  return 0;
}
