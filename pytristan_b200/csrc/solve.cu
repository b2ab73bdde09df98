// SURVEY.md 8f rank 4 — the step after the path in a collocation solver: replace rows of an
// operator by boundary-condition rows, and solve A y = f along a grid axis for every grid line.
//
// The solve is expressed through the path itself: the n x n system matrix (n <= 4096, e.g.
// D2 - k^2 I with boundary rows) is inverted ONCE on the device (Gauss-Jordan with partial
// pivoting, FP64) and every later solve is pt_dense_apply with the inverse — the batched
// tensor-core contraction along the axis, no transposes, all grid lines at once.
#include "common.cuh"

namespace {

// rows[r] of D (n x n row-major) <- R[r]
__global__ void replace_rows_kernel(double* __restrict__ D, int n, const int32_t* __restrict__ rows,
                                    int nrows, const double* __restrict__ R) {
  const int64_t total = int64_t(nrows) * n;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int r = int(idx / n), c = int(idx % n);
    D[int64_t(rows[r]) * n + c] = R[idx];
  }
}

// M = [A | I], n x 2n row-major
__global__ void augment_kernel(const double* __restrict__ A, int n, double* __restrict__ M) {
  const int64_t total = int64_t(n) * 2 * n;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int r = int(idx / (2 * n)), c = int(idx % (2 * n));
    M[idx] = c < n ? A[int64_t(r) * n + c] : (c - n == r ? 1.0 : 0.0);
  }
}

// pivot of column c among rows >= c: first row of maximal |M[r][c]| (deterministic)
__global__ void __launch_bounds__(1024) pivot_kernel(const double* __restrict__ M, int n, int c, int* __restrict__ piv,
                                                      int* __restrict__ info) {
  __shared__ double sv[1024];
  __shared__ int si[1024];
  double best = -1.0;
  int bi = c;
  for (int r = c + threadIdx.x; r < n; r += blockDim.x) {
    const double v = fabs(M[int64_t(r) * 2 * n + c]);
    if (v > best) { best = v; bi = r; }
  }
  sv[threadIdx.x] = best; si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double o = sv[threadIdx.x + s];
      const int oi = si[threadIdx.x + s];
      if (o > sv[threadIdx.x] || (o == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = o; si[threadIdx.x] = oi; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *piv = si[0];
    if (!(sv[0] > 0.0) && *info == 0) *info = c + 1;   // exactly singular at step c
  }
}

// swap rows c and *piv
__global__ void swap_rows_kernel(double* __restrict__ M, int n, int c, const int* __restrict__ piv) {
  const int p = *piv;
  double* rc = M + int64_t(c) * 2 * n;
  double* rp = M + int64_t(p) * 2 * n;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < 2 * n; j += gridDim.x * blockDim.x) {
    const double a = rc[j], b = rp[j];
    if (p != c) rp[j] = a;
    rc[j] = b;
  }
}

// pivot row /= pivot (the pivot is read from the saved multiplier column: no race with the update)
__global__ void scale_row_kernel(double* __restrict__ M, int n, int c, const double* __restrict__ col) {
  const double pivot = col[c];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < 2 * n; j += gridDim.x * blockDim.x)
    M[int64_t(c) * 2 * n + j] /= pivot;
}

// every other row r: M[r][:] -= col[r] * M[c][:]   (row c is read-only here)
__global__ void __launch_bounds__(256) eliminate_kernel(double* __restrict__ M, int n, int c,
                                                        const double* __restrict__ col) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 2 * n) return;
  const double pr = M[int64_t(c) * 2 * n + j];
  for (int r = blockIdx.y; r < n; r += gridDim.y) {
    if (r == c) continue;
    double* x = M + int64_t(r) * 2 * n + j;
    *x = fma(-col[r], pr, *x);
  }
}

__global__ void extract_kernel(const double* __restrict__ M, int n, double* __restrict__ Ainv) {
  const int64_t total = int64_t(n) * n;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int r = int(idx / n), c = int(idx % n);
    Ainv[idx] = M[int64_t(r) * 2 * n + n + c];
  }
}

inline unsigned blocks_for(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = int64_t(pt::sm_count()) * 16;
  return unsigned(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int pt_replace_rows(double* D, int n, const int32_t* rows, int nrows, const double* R,
                               pt_stream_t stream) {
  PT_REQUIRE(D && rows && R, "pt_replace_rows: null pointer");
  PT_REQUIRE(n >= 1 && nrows >= 0, "pt_replace_rows: bad shape");
  if (nrows == 0) return PT_OK;
  replace_rows_kernel<<<blocks_for(int64_t(nrows) * n, 256), 256, 0, pt::as_stream(stream)>>>(D, n, rows, nrows, R);
  return pt::check_launch("pt_replace_rows");
}

extern "C" int64_t pt_invert_work_size(int n) { return n <= 0 ? 0 : int64_t(n) * 2 * n + n + 2; }

extern "C" int pt_invert(const double* A, int n, double* Ainv, double* work, int32_t* info, pt_stream_t stream) {
  PT_REQUIRE(A && Ainv && work && info, "pt_invert: null pointer");
  PT_REQUIRE(n >= 1 && n <= 4096, "pt_invert: n=%d outside [1, 4096]", n);
  cudaStream_t st = pt::as_stream(stream);
  double* M = work;
  double* col = work + int64_t(n) * 2 * n;
  int* piv = reinterpret_cast<int*>(col + n);
  PT_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
  augment_kernel<<<blocks_for(int64_t(n) * 2 * n, 256), 256, 0, st>>>(A, n, M);
  const unsigned row_blocks = blocks_for(2 * n, 256);
  const dim3 egrid((2 * n + 255) / 256, unsigned(n < 64 ? n : 64));
  for (int c = 0; c < n; ++c) {
    pivot_kernel<<<1, 1024, 0, st>>>(M, n, c, piv, info);
    swap_rows_kernel<<<row_blocks, 256, 0, st>>>(M, n, c, piv);
    // multipliers = column c after the swap (strided gather of n values)
    PT_CUDA(cudaMemcpy2DAsync(col, sizeof(double), M + c, size_t(2) * n * sizeof(double), sizeof(double), n,
                              cudaMemcpyDeviceToDevice, st));
    scale_row_kernel<<<row_blocks, 256, 0, st>>>(M, n, c, col);
    eliminate_kernel<<<egrid, 256, 0, st>>>(M, n, c, col);
  }
  extract_kernel<<<blocks_for(int64_t(n) * n, 256), 256, 0, st>>>(M, n, Ainv);
  return pt::check_launch("pt_invert");
}
