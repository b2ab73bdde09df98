// SURVEY.md 8f rank 4 — the step after the path in a collocation solver: replace rows of an
// operator by boundary-condition rows, and solve A y = f along a grid axis for every grid line.
//
// The solve is expressed through the path itself: the n x n system matrix (n <= 4096, e.g.
// D2 - k^2 I with boundary rows) is inverted ONCE on the device (Gauss-Jordan with partial
// pivoting, FP64) and every later solve is pt_dense_apply with the inverse — the batched
// tensor-core contraction along the axis, no transposes, all grid lines at once.
//
// The factorisation-based form (pt_lu_factor / pt_lu_plan / pt_lu_solve): P A = L U with partial
// pivoting, then blocked forward and backward substitution in which every step is again the hot path —
// the off-diagonal block rows of L and U and the inverted nb x nb diagonal blocks are packed operators
// applied by pt_dense_apply_ld to row ranges of the fields (the scheme of a GPU TRSM), so the solve
// never forms A^-1 and never transposes or repacks a field.
#include "dense_common.cuh"

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


// ---- LU with partial pivoting (in place, n x n row-major, leading dimension n) ---------------------
__global__ void iota_kernel(int32_t* __restrict__ perm, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) perm[i] = i;
}

// pivot of column c among rows >= c: first row of maximal |A[r][c]| (the rule of LAPACK's idamax)
__global__ void __launch_bounds__(1024) lu_pivot_kernel(const double* __restrict__ A, int n, int c, int* __restrict__ piv,
                                                         int32_t* __restrict__ perm, int* __restrict__ info) {
  __shared__ double sv[1024];
  __shared__ int si[1024];
  double best = -1.0;
  int bi = c;
  for (int r = c + threadIdx.x; r < n; r += blockDim.x) {
    const double v = fabs(A[int64_t(r) * n + c]);
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
    const int p = si[0];
    *piv = p;
    const int32_t t = perm[c]; perm[c] = perm[p]; perm[p] = t;
    if (!(sv[0] > 0.0) && *info == 0) *info = c + 1;   // exactly singular at step c
  }
}

// swap rows c and *piv over all n columns (multipliers included, as LAPACK's laswp does)
__global__ void lu_swap_kernel(double* __restrict__ A, int n, int c, const int* __restrict__ piv) {
  const int p = *piv;
  if (p == c) return;
  double* rc = A + int64_t(c) * n;
  double* rp = A + int64_t(p) * n;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double a = rc[j];
    rc[j] = rp[j];
    rp[j] = a;
  }
}

// multipliers l[r] = A[r][c] / A[c][c], r > c: stored in place and in col[] for the update
__global__ void lu_scale_kernel(double* __restrict__ A, int n, int c, double* __restrict__ col) {
  const double pivot = A[int64_t(c) * n + c];
  for (int r = c + 1 + blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const double l = pivot != 0.0 ? A[int64_t(r) * n + c] / pivot : 0.0;
    A[int64_t(r) * n + c] = l;
    col[r] = l;
  }
}

// trailing update A[r][j] -= l[r] * A[c][j], r > c, j > c   (row c and col[] are read-only here)
__global__ void __launch_bounds__(256) lu_update_kernel(double* __restrict__ A, int n, int c,
                                                        const double* __restrict__ col) {
  const int j = c + 1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double pr = A[int64_t(c) * n + j];
  for (int r = c + 1 + blockIdx.y; r < n; r += gridDim.y) {
    double* x = A + int64_t(r) * n + j;
    *x = fma(-col[r], pr, *x);
  }
}

// Diagonal block b of the factors in the layout the substitution kernel reads: TT[k][i] (k-major, m16 = m rounded
// up to 16, identity in the padding) = -L_bb[i][k] below the (unit) diagonal for the lower factor; for U the block
// with both index orders REVERSED (i -> m-1-i), which turns the upper triangle into a lower one: -U below the
// diagonal and 1/U_ii ON it (the kernel multiplies; one extra rounding, still componentwise backward stable).
__global__ void tri_pack_kernel(const double* __restrict__ LU, int n, int r0, int m, int m16, int lower,
                                double* __restrict__ TT) {
  const int total = m16 * m16;
  const double* T = LU + int64_t(r0) * n + r0;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx / m16, i = idx - k * m16;
    double v = (i == k) ? 1.0 : 0.0;
    if (i < m && k < m) {
      if (lower) { if (i > k) v = -T[int64_t(i) * n + k]; }
      else if (i > k) v = -T[int64_t(m - 1 - i) * n + (m - 1 - k)];
      else if (i == k) v = 1.0 / T[int64_t(m - 1 - i) * n + (m - 1 - i)];
    }
    TT[idx] = v;
  }
}

// Triangular solves with the m x m diagonal block for TRI_TL grid lines per CTA by substitution (componentwise
// backward stable, unlike a product with an inverted block).  The rows of the block's row range are staged in shared
// memory [row][line]; every warp owns 32 lines and takes them through sub-blocks of 16 rows:
//   (1) rows s0..s0+15 -= T[s0.., 0:s0] x[0:s0] for its 32 lines as a 16 x s0 x 32 product on the FP64 tensor cores
//       (DMMA m8n8k4 accumulating onto the rows themselves: A fragments of the negated triangle straight from L1,
//       B fragments from the tile, pitch = 4 mod 16);
//   (2) the 16 x 16 triangle by one thread per line in registers (120 FMAs);
//   the rows of the triangle a sub-block needs arrive by cp.async in a double-buffered panel one sub-block ahead.
// Only warp barriers in the loop.  128 lines = 4 warps per CTA and ONE CTA per SM (the tile of a 128-row block
// fills it): a warp issuing plain FP64 instructions beside a warp that streams DMMAs on the same sub-partition
// waits ~150 cycles per instruction for the pipe, so every sub-partition hosts exactly one warp.  In/Out address element (line nn, row r) at
// (nn / before) * ld + r * before + nn % before; `reversed` mirrors the row order (upper factor).
constexpr int TRI_PP = 20;                    // pitch of a staged triangle panel row (16 entries), = 4 mod 16
constexpr int TRI_MAXM = 128;

__device__ __forceinline__ void tri_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <bool UNIT, int TL>
__global__ void __launch_bounds__(TL) tri_solve_kernel(const double* __restrict__ TT, int m, int m16, int reversed,
                                                           const double* __restrict__ In, int64_t ld_in,
                                                           double* __restrict__ Out, int64_t ld_out, int64_t NN,
                                                           int64_t before) {
  constexpr int TRI_TL = TL, TRI_TLP = TL + 4;
  extern __shared__ __align__(16) double tile[];   // [m16][TRI_TLP], then two panel buffers
  double* panels = tile + m16 * TRI_TLP;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int L0 = (tid >> 5) * 32;
  const int64_t n0 = int64_t(blockIdx.x) * TRI_TL;

  // rows of the triangle that sub-block s0 needs: TT[k][s0 .. s0+15] for k < s0 + 16 (the last 16 k are the
  // 16 x 16 triangle itself) — 16-byte cp.async into panel buffer (s0 / 16) & 1
  auto issue_panel = [&](int s0) {
    double* pan = panels + ((s0 >> 4) & 1) * (m16 * TRI_PP);
    const int pieces = (s0 + 16) * 8;
    for (int pc = tid; pc < pieces; pc += TRI_TL) {
      const int k = pc >> 3, c = (pc & 7) * 2;
      ptdense::cp_async16(pan + k * TRI_PP + c, TT + int64_t(k) * m16 + s0 + c, 16);
    }
  };

  // the whole tile asynchronously (8-byte copies: the tile transposes / strides), zero padding by plain stores
  if (before == 1) {
    // lines are rows of memory: consecutive threads take consecutive rows of one line
    for (int row = tid; row < m16; row += TRI_TL) {
      const int ar = reversed ? m - 1 - row : row;
      for (int line = 0; line < TRI_TL; ++line) {
        const int64_t nn = n0 + line;
        double* dst = tile + row * TRI_TLP + line;
        if (row < m && nn < NN) ptdense::cp_async8(dst, In + nn * ld_in + ar, 8);
        else *dst = 0.0;
      }
    }
  } else {
    const int64_t nn = n0 + tid;
    const int64_t a = nn / before;
    const double* src = In + a * ld_in + (nn - a * before);
    const bool ok = nn < NN;
    for (int row = 0; row < m16; ++row) {
      double* dst = tile + row * TRI_TLP + tid;
      if (ok && row < m) ptdense::cp_async8(dst, src + int64_t(reversed ? m - 1 - row : row) * before, 8);
      else *dst = 0.0;
    }
  }
  issue_panel(0);
  ptdense::cp_async_commit();

  double* col = tile + tid;
  for (int s0 = 0; s0 < m16; s0 += 16) {
    ptdense::cp_async_wait<0>();
    __syncthreads();                       // panel s0 (and the tile) landed; everyone is done with the other buffer
    if (s0 + 16 < m16) issue_panel(s0 + 16);
    ptdense::cp_async_commit();
    const double* pan = panels + ((s0 >> 4) & 1) * (m16 * TRI_PP);
    if (s0 > 0) {
      double acc[2][4][2];
      // C fragment: row = s0 + 8 i + g, lines L0 + 8 j + 2 t4, +1 — the accumulators start as the rows themselves
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 v = *reinterpret_cast<const double2*>(tile + (s0 + 8 * i + g) * TRI_TLP + L0 + 8 * j + 2 * t4);
          acc[i][j][0] = v.x; acc[i][j][1] = v.y;
        }
      const double* ap = pan + t4 * TRI_PP + g;                     // A[row = s0 + 8 i + g][k = k0 + t4] = -T
      const double* bp = tile + t4 * TRI_TLP + L0 + g;              // B[k = k0 + t4][line = L0 + 8 j + g]
#pragma unroll 2
      for (int k0 = 0; k0 < s0; k0 += 4) {
        const double a0 = ap[0], a1 = ap[8];
        double bf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = bp[8 * j];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tri_dmma(acc[0][j][0], acc[0][j][1], a0, bf[j]);
          tri_dmma(acc[1][j][0], acc[1][j][1], a1, bf[j]);
        }
        ap += 4 * TRI_PP;
        bp += 4 * TRI_TLP;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<double2*>(tile + (s0 + 8 * i + g) * TRI_TLP + L0 + 8 * j + 2 * t4) =
              make_double2(acc[i][j][0], acc[i][j][1]);
      __syncwarp();
    }
    double t[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) t[r] = col[(s0 + r) * TRI_TLP];
    const double* tri = pan + s0 * TRI_PP;                           // the 16 x 16 triangle: tri[r * PP + r2] = -T[s0+r2][s0+r]
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const double x = UNIT ? t[r] : t[r] * tri[r * TRI_PP + r];
      t[r] = x;
#pragma unroll
      for (int r2 = r + 1; r2 < 16; ++r2) t[r2] = fma(tri[r * TRI_PP + r2], x, t[r2]);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) col[(s0 + r) * TRI_TLP] = t[r];
    __syncwarp();
  }
  __syncthreads();
  if (before == 1) {
    for (int row = tid; row < m; row += TRI_TL) {
      const int ar = reversed ? m - 1 - row : row;
      for (int line = 0; line < TRI_TL; ++line) {
        const int64_t nn = n0 + line;
        if (nn < NN) Out[nn * ld_out + ar] = tile[row * TRI_TLP + line];
      }
    }
  } else {
    const int64_t nn = n0 + tid;
    if (nn < NN) {
      const int64_t a = nn / before;
      double* dst = Out + a * ld_out + (nn - a * before);
      for (int row = 0; row < m; ++row) dst[int64_t(reversed ? m - 1 - row : row) * before] = tile[row * TRI_TLP + tid];
    }
  }
}

// W[a, i, b] = F[a, perm[i], b]: the row permutation of P A = L U applied to every grid line
__global__ void __launch_bounds__(256) permute_rows_kernel(const double* __restrict__ F, double* __restrict__ W,
                                                           const int32_t* __restrict__ perm, int n, int64_t after,
                                                           int64_t before) {
  const int64_t total = after * n * before;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = idx % before;
    const int64_t r = idx / before;
    const int i = int(r % n);
    const int64_t a = r / n;
    W[idx] = F[(a * n + perm[i]) * before + b];
  }
}

// Layout of an LU plan (doubles): for every block row b (m_b = min(nb, n - b*nb) rows, r0 = b*nb) the off-diagonal
// block rows L[b, 0:r0] (m_b x r0) and U[b, r0+m_b:n] in pt_operator_pack's layout and the two diagonal triangles in
// tri_pack_kernel's layout (m16 x m16 each).
struct LuBlock { int64_t l_off, lt_off, u_off, ut_off; int m, m16, r0, rest; };

inline int64_t packed_doubles(int n_out, int n_in) {
  if (n_out <= 0 || n_in <= 0) return 0;
  return int64_t((n_in + 3) / 4) * (8 * ((int64_t(n_out) + 7) / 8)) * 4;
}

// Block rows actually used for a requested maximum of nb rows: n split into ceil(n / nb) nearly equal parts of a
// multiple of 8 rows (n = 257, nb = 128: 88 + 88 + 81 instead of 128 + 128 + 1 — a one-row block costs two passes
// over the fields for nothing).
inline int lu_rows(int n, int nb) {
  const int parts = (n + nb - 1) / nb;
  const int bs = 8 * ((n + 8 * parts - 1) / (8 * parts));
  return bs < nb ? bs : nb;
}

inline int64_t lu_block(int n, int nb_max, int b, LuBlock* out) {   // returns the offset past block b
  const int nb = lu_rows(n, nb_max);
  int64_t off = 0;
  for (int i = 0; i <= b; ++i) {
    LuBlock k;
    k.r0 = i * nb;
    k.m = n - k.r0 < nb ? n - k.r0 : nb;
    k.rest = n - k.r0 - k.m;
    k.m16 = 16 * ((k.m + 15) / 16);
    k.l_off = off;    off += packed_doubles(k.m, k.r0);
    k.lt_off = off;   off += int64_t(k.m16) * k.m16;
    k.u_off = off;    off += packed_doubles(k.m, k.rest);
    k.ut_off = off;   off += int64_t(k.m16) * k.m16;
    if (i == b && out) *out = k;
  }
  return off;
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

// ---- factorisation-based solve -------------------------------------------------------------------
extern "C" int64_t pt_lu_work_size(int n) { return n <= 0 ? 0 : int64_t(n) + 2; }

extern "C" int pt_lu_factor(const double* A, int n, double* LU, int32_t* perm, double* work, int32_t* info,
                            pt_stream_t stream) {
  PT_REQUIRE(A && LU && perm && work && info, "pt_lu_factor: null pointer");
  PT_REQUIRE(n >= 1 && n <= 8192, "pt_lu_factor: n=%d outside [1, 8192]", n);
  cudaStream_t st = pt::as_stream(stream);
  double* col = work;
  int* piv = reinterpret_cast<int*>(work + n);
  PT_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
  if (A != LU) PT_CUDA(cudaMemcpyAsync(LU, A, sizeof(double) * size_t(n) * n, cudaMemcpyDeviceToDevice, st));
  iota_kernel<<<blocks_for(n, 256), 256, 0, st>>>(perm, n);
  for (int c = 0; c < n; ++c) {
    lu_pivot_kernel<<<1, 1024, 0, st>>>(LU, n, c, piv, perm, info);
    lu_swap_kernel<<<blocks_for(n, 256), 256, 0, st>>>(LU, n, c, piv);
    const int rem = n - c - 1;
    if (rem == 0) break;
    lu_scale_kernel<<<blocks_for(rem, 256), 256, 0, st>>>(LU, n, c, col);
    const dim3 ugrid((rem + 255) / 256, unsigned(rem < 128 ? rem : 128));
    lu_update_kernel<<<ugrid, 256, 0, st>>>(LU, n, c, col);
  }
  return pt::check_launch("pt_lu_factor");
}

extern "C" int64_t pt_lu_plan_size(int n, int nb) {
  if (n <= 0 || nb <= 0) return 0;
  const int bs = lu_rows(n, nb);
  return lu_block(n, nb, (n + bs - 1) / bs - 1, nullptr);
}

extern "C" int pt_lu_plan(const double* LU, int n, int nb, double* plan, pt_stream_t stream) {
  PT_REQUIRE(LU && plan, "pt_lu_plan: null pointer");
  PT_REQUIRE(n >= 1 && nb >= 16 && nb <= 128 && nb % 16 == 0, "pt_lu_plan: bad n=%d / block size %d (16..128, multiple of 16)", n, nb);
  cudaStream_t st = pt::as_stream(stream);
  const int bs = lu_rows(n, nb);
  const int nblk = (n + bs - 1) / bs;
  for (int b = 0; b < nblk; ++b) {
    LuBlock k;
    lu_block(n, nb, b, &k);
    const unsigned tb = blocks_for(int64_t(k.m16) * k.m16, 256);
    tri_pack_kernel<<<tb, 256, 0, st>>>(LU, n, k.r0, k.m, k.m16, 1, plan + k.lt_off);
    tri_pack_kernel<<<tb, 256, 0, st>>>(LU, n, k.r0, k.m, k.m16, 0, plan + k.ut_off);
    int rc = pt::check_launch("pt_lu_plan");
    if (rc != PT_OK) return rc;
    if (k.r0 > 0) {
      rc = pt_operator_pack(LU + int64_t(k.r0) * n, k.m, k.r0, n, plan + k.l_off, stream);
      if (rc != PT_OK) return rc;
    }
    if (k.rest > 0) {
      rc = pt_operator_pack(LU + int64_t(k.r0) * n + k.r0 + k.m, k.m, k.rest, n, plan + k.u_off, stream);
      if (rc != PT_OK) return rc;
    }
  }
  return PT_OK;
}

namespace {
int g_tri_tl = 0;      // tuning hook (pt_debug_set_tri_lines): 0 = by shape, 64 or 128 = lines per CTA

template <bool UNIT, int TL>
int launch_tri_solve_tl(const double* TT, const LuBlock& k, int reversed, const double* in, double* out, int64_t ld,
                        int64_t after, int64_t before, cudaStream_t st) {
  auto kern = tri_solve_kernel<UNIT, TL>;
  const size_t smem = size_t(k.m16) * (TL + 4 + 2 * TRI_PP) * sizeof(double);
  static bool attr_set[64] = {false};
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(size_t(TRI_MAXM) * (TL + 4 + 2 * TRI_PP) * sizeof(double))));
    attr_set[dev] = true;
  }
  const int64_t NN = after * before;
  const int64_t grid = (NN + TL - 1) / TL;
  if (grid > 2147483647LL) return pt::fail(PT_ERR_UNSUPPORTED, "pt_lu_solve: grid too large");
  kern<<<unsigned(grid), TL, smem, st>>>(TT, k.m, k.m16, reversed, in, ld, out, ld, NN, before);
  return pt::check_launch("pt_lu_solve");
}

// 64 lines per CTA: two or three CTAs share an SM, so that the loads and stores of one overlap the substitution of
// another (one warp per sub-partition and CTA); 128 lines when that would not fill the GPU anyway
template <bool UNIT>
int launch_tri_solve(const double* TT, const LuBlock& k, int reversed, const double* in, double* out, int64_t ld,
                     int64_t after, int64_t before, cudaStream_t st) {
  const int tl = g_tri_tl ? g_tri_tl : 64;
  if (tl == 64) return launch_tri_solve_tl<UNIT, 64>(TT, k, reversed, in, out, ld, after, before, st);
  return launch_tri_solve_tl<UNIT, 128>(TT, k, reversed, in, out, ld, after, before, st);
}
}  // namespace

extern "C" int pt_debug_set_tri_lines(int lines) {
  g_tri_tl = (lines == 64 || lines == 128) ? lines : 0;
  return PT_OK;
}

extern "C" int pt_lu_solve(const double* plan, const int32_t* perm, int n, int nb, const double* F, double* X,
                           double* work, int64_t after, int64_t before, pt_stream_t stream) {
  PT_REQUIRE(plan && perm && F && X && work, "pt_lu_solve: null pointer");
  PT_REQUIRE(n >= 1 && nb >= 16 && nb <= 128 && nb % 16 == 0, "pt_lu_solve: bad n=%d / block size %d", n, nb);
  PT_REQUIRE(after >= 0 && before >= 1, "pt_lu_solve: bad field shape");
  PT_REQUIRE(F != X && F != work && X != work, "pt_lu_solve: F, X and work must be distinct");
  if (after == 0) return PT_OK;
  cudaStream_t st = pt::as_stream(stream);
  const int bs = lu_rows(n, nb);
  const int nblk = (n + bs - 1) / bs;
  const int64_t ld = int64_t(n) * before;
  const int64_t total = after * ld;
  // X <- P F
  permute_rows_kernel<<<blocks_for(total, 256), 256, 0, st>>>(F, X, perm, n, after, before);
  int rc = pt::check_launch("pt_lu_solve");
  if (rc != PT_OK) return rc;
  // forward substitution L z = P F: z in `work`, block by block
  //   X_b -= L[b, 0:r0] z[0:r0]        (rows of X below the finished ones still hold P F; tensor cores)
  //   L_bb z_b = X_b                   (substitution)
  for (int b = 0; b < nblk; ++b) {
    LuBlock k;
    lu_block(n, nb, b, &k);
    double* xb = X + int64_t(k.r0) * before;
    double* zb = work + int64_t(k.r0) * before;
    if (k.r0 > 0) {
      rc = pt_dense_apply_ld(plan + k.l_off, k.m, k.r0, work, ld, xb, ld, after, before, -1.0, 1.0, stream);
      if (rc != PT_OK) return rc;
    }
    rc = launch_tri_solve<true>(plan + k.lt_off, k, 0, xb, zb, ld, after, before, st);
    if (rc != PT_OK) return rc;
  }
  // backward substitution U x = z: x in X, last block first
  //   z_b -= U[b, r0+m:n] x[r0+m:n]
  //   U_bb x_b = z_b
  for (int b = nblk - 1; b >= 0; --b) {
    LuBlock k;
    lu_block(n, nb, b, &k);
    double* xb = X + int64_t(k.r0) * before;
    double* zb = work + int64_t(k.r0) * before;
    if (k.rest > 0) {
      rc = pt_dense_apply_ld(plan + k.u_off, k.m, k.rest, X + int64_t(k.r0 + k.m) * before, ld, zb, ld, after, before,
                             -1.0, 1.0, stream);
      if (rc != PT_OK) return rc;
    }
    rc = launch_tri_solve<false>(plan + k.ut_off, k, 1, zb, xb, ld, after, before, st);
    if (rc != PT_OK) return rc;
  }
  return PT_OK;
}
