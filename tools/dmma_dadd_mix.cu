// How much DMMA throughput does a plain FP64 instruction cost when it is issued by ANOTHER warp of the same SM
// sub-partition?  Warps 0-7 of a 512-thread CTA (two per sub-partition) stream DMMAs, warps 8-15 issue `burst`
// independent FP64 (or, as a control, integer) instructions every `period` cycles.  One CTA per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mix tools/dmma_dadd_mix.cu && /tmp/mix
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) mix_kernel(double* out, long long* cyc, int iters, int burst, int period, int kind) {
  const int warp = threadIdx.x >> 5;
  if (warp < 8) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const double a = 0.5 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * 512 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  } else {
    if (kind == 0) return;                       // no second stream at all
    double x[8];
    long long y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = threadIdx.x + i; }
    const double inc = 1.0 + threadIdx.x * 1e-12;
    // runs for about as long as the DMMA warps: iters * 16 DMMAs * 16 cycles * 2 warps per sub-partition
    const long long total = (long long)iters * 16 * 16 * 2;
    const long long start = clock64();
    long long next = start;
    while (next - start < total) {
      while (clock64() < next) { }
      if (kind == 1) {
        for (int r = 0; r < burst; r += 8)
#pragma unroll
          for (int i = 0; i < 8; ++i) if (r + i < burst) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(inc));
      } else if (kind == 2) {
        for (int r = 0; r < burst; r += 8)
#pragma unroll
          for (int i = 0; i < 8; ++i) if (r + i < burst) asm volatile("add.s64 %0, %0, 3;" : "+l"(y[i]));
      }
      next += period;
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + double(y[i]);
    out[blockIdx.x * 512 + threadIdx.x] = s;
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, iters = 20000;
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * 512); cudaMalloc(&cyc, sizeof(long long) * sms);
  struct Case { const char* name; int burst, period, kind; };
  const Case cases[] = {
      {"DMMA only", 0, 0, 0},
      {"+ polling warps, no instructions", 0, 512, 3},
      {"+ 1 integer add / 512 cycles per warp (control)", 1, 512, 2},
      {"+ 1 DADD / 512 cycles per warp (= 1 per 256 per sub-partition: the E/O rate)", 1, 512, 1},
      {"+ 2 DADD / 1024 cycles", 2, 1024, 1},
      {"+ 8 DADD / 4096 cycles (one burst per chunk)", 8, 4096, 1},
      {"+ 32 DADD / 16384 cycles", 32, 16384, 1},
      {"+ 1 DADD / 256 cycles per warp (2x the E/O rate)", 1, 256, 1},
      {"+ 1 DADD / 128 cycles per warp (4x)", 1, 128, 1},
  };
  for (const Case& c : cases) {
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      mix_kernel<<<sms, 512>>>(out, cyc, iters, c.burst, c.period, c.kind);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0 && ms < best) best = ms;
    }
    long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double fl = 2.0 * 256.0 * 16.0 * iters * 8.0 * sms;
    // DMMA rate from the DMMA warps' own clock (the polling warps may outlive them)
    printf("%-82s %7.2f TFLOP/s by events, %6.2f cycles per DMMA per sub-partition\n", c.name, fl / best * 1e-9,
           double(h) / (iters * 16.0 * 2.0));
  }
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
