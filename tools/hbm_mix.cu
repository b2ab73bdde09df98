// HBM bandwidth for read:write mixes on B200: the fused gradient reads 8 B and writes 24 B per point,
// the copy that MEASURED_PEAKS / the profiling recipe quote is 1:1.  Output: one JSON line.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/hbm_mix tools/hbm_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int NW, bool STREAM>
__global__ void __launch_bounds__(512) mix_kernel(const double2* __restrict__ x, double2* __restrict__ y0,
                                                  double2* __restrict__ y1, double2* __restrict__ y2, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const double2 v = x[i];
    double2* ys[3] = {y0, y1, y2};
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const double2 w = make_double2(v.x + k, v.y - k);
      if (STREAM) __stcs(ys[k] + i, w); else ys[k][i] = w;
    }
  }
}

__global__ void __launch_bounds__(512) read_kernel(const double2* __restrict__ x, double* out, size_t n) {
  double s = 0.0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const double2 v = x[i];
    s += v.x + v.y;
  }
  if (s == 1.2345) out[0] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(a));
    for (int k = 0; k < 4; ++k) f();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms / 4 < best) best = ms / 4;
  }
  return best;
}

int main() {
  const size_t n = size_t(1) << 26;          // double2 elements: 1 GiB per array
  double2 *x, *y0, *y1, *y2; double* out;
  CK(cudaMalloc(&x, n * 16)); CK(cudaMalloc(&y0, n * 16)); CK(cudaMalloc(&y1, n * 16)); CK(cudaMalloc(&y2, n * 16));
  CK(cudaMalloc(&out, 8));
  CK(cudaMemset(x, 0, n * 16));
  const int grid = 148 * 8;
  const double gb = n * 16 * 1e-9;
  float t_r = time_ms([&] { read_kernel<<<grid, 512>>>(x, out, n); });
  float t_11 = time_ms([&] { mix_kernel<1, false><<<grid, 512>>>(x, y0, y1, y2, n); });
  float t_11s = time_ms([&] { mix_kernel<1, true><<<grid, 512>>>(x, y0, y1, y2, n); });
  float t_12 = time_ms([&] { mix_kernel<2, true><<<grid, 512>>>(x, y0, y1, y2, n); });
  float t_13 = time_ms([&] { mix_kernel<3, false><<<grid, 512>>>(x, y0, y1, y2, n); });
  float t_13s = time_ms([&] { mix_kernel<3, true><<<grid, 512>>>(x, y0, y1, y2, n); });
  float t_w = time_ms([&] { CK(cudaMemsetAsync(y0, 1, n * 16)); });
  printf("{\"read_only_gbs\": %.1f, \"copy_1r1w_gbs\": %.1f, \"copy_1r1w_stcs_gbs\": %.1f, \"mix_1r2w_stcs_gbs\": %.1f, "
         "\"mix_1r3w_gbs\": %.1f, \"mix_1r3w_stcs_gbs\": %.1f, \"memset_write_only_gbs\": %.1f}\n",
         gb / t_r * 1e3, 2 * gb / t_11 * 1e3, 2 * gb / t_11s * 1e3, 3 * gb / t_12 * 1e3, 4 * gb / t_13 * 1e3,
         4 * gb / t_13s * 1e3, gb / t_w * 1e3);
  return 0;
}
