// FP64 peak microbenchmark for B200 (sm_100a): register-resident DMMA (mma.sync m8n8k4 f64)
// and DFMA loops, timed with CUDA events.  Output: one JSON line.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double seed) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = seed + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double seed) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = seed + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  const int iters = 20000;
  // DMMA: warps/SM sweep through blocks/SM (256 threads = 8 warps per block)
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = sms * bps;
    double ms8 = time_ms([&] { dmma_loop<8><<<grid, 256>>>(out, iters, 0.5); }, 5);
    double fl = 2.0 * 256.0 * 8 * (double)iters * 8 /*warps*/ * grid;
    printf(", \"dmma_tflops_acc8_bps%d\": %.3f", bps, fl / ms8 * 1e-9);
    double ms16 = time_ms([&] { dmma_loop<16><<<grid, 256>>>(out, iters, 0.5); }, 5);
    fl = 2.0 * 256.0 * 16 * (double)iters * 8 * grid;
    printf(", \"dmma_tflops_acc16_bps%d\": %.3f", bps, fl / ms16 * 1e-9);
  }
  {
    int grid = sms * 4;
    double ms = time_ms([&] { dfma_loop<16><<<grid, 256>>>(out, iters, 0.5); }, 5);
    double fl = 2.0 * 16 * (double)iters * 256.0 * grid;
    printf(", \"dfma_tflops\": %.3f", fl / ms * 1e-9);
  }
  // sustained DMMA (about 2 s back-to-back)
  {
    int grid = sms * 2;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 60; ++n) dmma_loop<16><<<grid, 256>>>(out, iters * 4, 0.5);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 256.0 * 16 * (double)iters * 4 * 8 * grid * n;
    printf(", \"dmma_tflops_sustained\": %.3f, \"sustained_ms\": %.1f", fl / ms * 1e-9, ms);
  }
  printf("}\n");
  return 0;
}
