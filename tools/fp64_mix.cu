// Plain FP64 pipe rates on B200 by opcode and warps per scheduler: DFMA, DADD, DMUL and the FFT's mix
// (2 DADD : 1 DMUL : 1 DFMA), register resident, 16 independent chains per thread.  One JSON line;
// rates in warp instructions per cycle per SM sub-partition at the measured duration and 1.965 GHz.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_mix tools/fp64_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int KIND>
__global__ void __launch_bounds__(128) loop_kernel(double* out, int iters, double seed) {
  constexpr int NACC = 16;
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i + threadIdx.x * 1e-9;
  const double a = seed + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x + 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (KIND == 0) c[i] = fma(c[i], a, b);
      else if (KIND == 1) c[i] = c[i] + b;
      else if (KIND == 2) c[i] = c[i] * a;
      else {   // 4 instructions: add, add, mul, fma
        c[i] = c[i] + b;
        c[i] = c[i] - a;
        c[i] = c[i] * a;
        c[i] = fma(c[i], a, b);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
double rate(double* out, int sms, int warps_per_smsp) {
  const int iters = 20000;
  const int grid = sms * warps_per_smsp;           // 128-thread CTAs: one warp per sub-partition each
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  loop_kernel<KIND><<<grid, 128>>>(out, iters, 0.5); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0)); loop_kernel<KIND><<<grid, 128>>>(out, iters, 0.5); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double instr_per_warp = 16.0 * iters * (KIND == 3 ? 4 : 1);
  const double cycles = best * 1e-3 * 1.965e9;
  return instr_per_warp * warps_per_smsp / cycles;   // warp instructions per cycle per sub-partition
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 128));
  printf("{\"unit\": \"warp instructions / cycle / SM sub-partition (0.5 = 16 lanes)\"");
  const char* names[4] = {"dfma", "dadd", "dmul", "mix_2add_1mul_1fma"};
  for (int w = 1; w <= 4; ++w) {
    double r[4] = {rate<0>(out, sms, w), rate<1>(out, sms, w), rate<2>(out, sms, w), rate<3>(out, sms, w)};
    for (int k = 0; k < 4; ++k) printf(", \"%s_w%d\": %.3f", names[k], w, r[k]);
  }
  printf("}\n");
  return 0;
}
