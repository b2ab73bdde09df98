// DMMA throughput vs warps per SM sub-partition (register-resident loop).
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void dmma_loop(double* out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = 0.5 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(double* out, int threads, int sms) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dmma_loop<NACC><<<sms, threads>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); dmma_loop<NACC><<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 2.0 * 256.0 * NACC * (double)iters * (threads / 32) * sms;
  printf("warps/SM=%2d nacc=%2d  %.2f TFLOP/s\n", threads / 32, NACC, fl / ms * 1e-9);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
  for (int t : {32, 64, 128, 256, 384, 512}) { run<4>(out, t, p.multiProcessorCount); run<16>(out, t, p.multiProcessorCount); run<32>(out, t, p.multiProcessorCount); }
  return 0;
}
