// K5/K6/K7 — on-device generation of nodes, grids and operators; permutation copy.
//
// Arithmetic that the reference performs with NumPy elementwise ops (grid.py:75-77, :205) is
// restated with the round-to-nearest intrinsics (__dmul_rn, __dadd_rn, ...) so that the compiler
// cannot contract a*b+c into an FMA: an FMA changes up to half of np.linspace's results
// (hard part 2).  The operator generators use the same operation order as
// oracle/operators.py for the same reason.
#include "common.cuh"
#include "glibc_cos.cuh"

namespace {

constexpr double kPi = 3.141592653589793;  // == numpy.pi == M_PI

// ---- K5 ------------------------------------------------------------------------------------------
__global__ void linspace_kernel(double* __restrict__ x, double start, double stop, int64_t n) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double delta = __dsub_rn(stop, start);
  const int64_t div = n - 1;
  double y;
  if (div > 0) {
    const double step = __ddiv_rn(delta, double(div));
    if (step == 0.0) y = __dmul_rn(__ddiv_rn(double(i), double(div)), delta);
    else y = __dmul_rn(double(i), step);
  } else {
    y = __dmul_rn(double(i), delta);
  }
  y = __dadd_rn(y, start);
  if (n > 1 && i == n - 1) y = stop;
  x[i] = y;
}

__device__ __forceinline__ double affine_map(double x, double xmin, double xmax) {
  // (1.0 - x) / 2.0 * (xmax - xmin) + xmin, one rounding per operation (grid.py:77)
  const double t = __ddiv_rn(__dsub_rn(1.0, x), 2.0);
  return __dadd_rn(__dmul_rn(t, __dsub_rn(xmax, xmin)), xmin);
}

__global__ void cheb_kernel(double* __restrict__ x, int64_t n, int mapped, double xmin, double xmax) {
  const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (j >= n) return;
  const double theta = __ddiv_rn(__dmul_rn(double(j), kPi), double(n - 1));
  double v = cos(theta);
  if (mapped) v = affine_map(v, xmin, xmax);
  x[j] = v;
}

// the same nodes with the host libm's cosine reproduced bit for bit (glibc_cos.cuh)
template <bool FMA>
__global__ void cheb_libm_kernel(double* __restrict__ x, int64_t n, int mapped, double xmin, double xmax) {
  const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (j >= n) return;
  const double theta = __ddiv_rn(__dmul_rn(double(j), kPi), double(n - 1));
  double v = ptglibc::glibc_cos<FMA>(theta);
  if (mapped) v = affine_map(v, xmin, xmax);
  x[j] = v;
}

template <bool FMA>
__global__ void cos_libm_kernel(const double* __restrict__ t, double* __restrict__ y, int64_t n) {
  const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (j < n) y[j] = ptglibc::glibc_cos<FMA>(t[j]);
}

__global__ void cheb_map_kernel(double* __restrict__ x, int64_t n, double xmin, double xmax) {
  const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (j < n) x[j] = affine_map(x[j], xmin, xmax);
}

constexpr int MAXDIM = 8;
struct ExpandParams {
  uint64_t stride[MAXDIM];
  uint64_t n[MAXDIM];
  uint64_t offset[MAXDIM];  // element offset of axis d inside `axes`
  uint64_t P;
  int ndim;
};

template <typename T, typename I>
__global__ void __launch_bounds__(256) expand_kernel(const T* __restrict__ axes, T* __restrict__ out,
                                                      const ExpandParams p) {
  const int d = blockIdx.y;
  const I P = I(p.P), stride = I(p.stride[d]), nd = I(p.n[d]);
  const T* __restrict__ ax = axes + p.offset[d];
  T* __restrict__ o = out + uint64_t(d) * p.P;
  for (I l = I(blockIdx.x) * blockDim.x + threadIdx.x; l < P; l += I(gridDim.x) * blockDim.x)
    o[l] = ax[(l / stride) % nd];
}

__global__ void axpoints_kernel(const double* __restrict__ row, int64_t stride, int64_t n,
                                double* __restrict__ x) {
  const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (j < n) x[j] = row[j * stride];
}

// ---- K6: Chebyshev collocation -------------------------------------------------------------------
__device__ __forceinline__ double cheb_c(int i, int n) {
  const double mag = (i == 0 || i == n - 1) ? 2.0 : 1.0;
  return (i & 1) ? -mag : mag;
}

// order-l off-diagonal entries from the order-(l-1) matrix held in D (in place; the diagonal of
// the previous order is only read here and rewritten by cheb_diag_kernel afterwards)
__global__ void cheb_offdiag_kernel(const double* __restrict__ x, int n, int l, double* D) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= int64_t(n) * n) return;
  const int i = int(idx / n), j = int(idx % n);
  if (i == j) return;
  const double cr = __ddiv_rn(cheb_c(i, n), cheb_c(j, n));
  const double dx = __dsub_rn(x[i], x[j]);
  if (l == 1) {
    D[idx] = __ddiv_rn(cr, dx);
  } else {
    const double t = __dsub_rn(__dmul_rn(cr, D[int64_t(i) * n + i]), D[idx]);
    D[idx] = __dmul_rn(__ddiv_rn(double(l), dx), t);
  }
}

// negative-sum trick, sequential left-to-right sum of the off-diagonal entries of each row
__global__ void cheb_diag_kernel(int n, double* D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* row = D + int64_t(i) * n;
  double s = 0.0;
  for (int j = 0; j < n; ++j)
    if (j != i) s = __dadd_rn(s, row[j]);
  D[int64_t(i) * n + i] = -s;
}

// flipping trick (Don & Solomonoff 1995): rows below the middle (and the right part of the middle
// row of an odd n) are the mirror image of the rows above it, D[i][j] = s * D[n-1-i][n-1-j]:
// the matrix is exactly centrosymmetric (s = +1) / centro-antisymmetric (s = -1)
__global__ void flip_kernel(double* D, int n, int sign) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= int64_t(n) * n) return;
  const int i = int(idx / n), j = int(idx % n);
  const int mi = n - 1 - i, mj = n - 1 - j;
  if (i > mi || (i == mi && j > mj)) D[idx] = __dmul_rn(double(sign), D[int64_t(mi) * n + mj]);
  else if (sign < 0 && i == mi && j == mj) D[idx] = 0.0;   // the centre is its own negative
}

// ---- K6: Fourier ---------------------------------------------------------------------------------
__device__ double four_col(int k, int n, double s, int order) {
  // entry for (i - j) mod n == k; k > n/2 folded by symmetry (order 1 odd, order 2 even)
  if (k == 0) {
    if (order == 1) return 0.0;
    const double nn = double(n) * double(n);
    const double s2 = __dmul_rn(s, s);
    if ((n & 1) == 0) return -__dmul_rn(s2, __dadd_rn(__ddiv_rn(nn, 12.0), 1.0 / 6.0));
    return -__dmul_rn(s2, __ddiv_rn(__dsub_rn(nn, 1.0), 12.0));
  }
  // antipodal node of an even n, odd order: cot(pi/2) is exactly 0 (the rounded pi/2 would leave
  // 6e-17); it also makes the matrix exactly centro-antisymmetric (dense_apply_eo.cu)
  if (order == 1 && 2 * k == n) return 0.0;
  const bool fold = 2 * k > n;
  const int kk = fold ? n - k : k;
  const double theta = __ddiv_rn(__dmul_rn(double(kk), kPi), double(n));
  double sn, cs;
  sincos(theta, &sn, &cs);
  const double sign = (kk & 1) ? -1.0 : 1.0;
  double v;
  if (order == 1) {
    if ((n & 1) == 0) v = __dmul_rn(__dmul_rn(__dmul_rn(s, 0.5), sign), __ddiv_rn(cs, sn));
    else v = __ddiv_rn(__dmul_rn(__dmul_rn(s, 0.5), sign), sn);
    return fold ? -v : v;
  }
  const double s2 = __dmul_rn(s, s);
  const double sn2 = __dmul_rn(sn, sn);
  if ((n & 1) == 0) v = -__ddiv_rn(__dmul_rn(s2, sign), __dmul_rn(2.0, sn2));
  else v = -__ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(s2, 0.5), sign), cs), sn2);
  return v;
}

__global__ void fourmat_kernel(int n, double s, int order, double* __restrict__ D) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= int64_t(n) * n) return;
  const int i = int(idx / n), j = int(idx % n);
  int k = i - j;
  if (k < 0) k += n;
  D[idx] = four_col(k, n, s, order);
}


// ---- K6: polynomial collocation on arbitrary nodes (barycentric weights) -------------------------
// w_j = 1 / prod_{k != j} (x_j - x_k) held as mantissa * 2^exponent (frexp after every multiply,
// exact scaling) so that n = 2048 nodes on [-1, 1] neither under- nor overflow.
__global__ void bary_weights_kernel(const double* __restrict__ x, int n, double* __restrict__ mant,
                                    int* __restrict__ expo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double xj = x[j];
  double m = 1.0;
  int e = 0;
  for (int k = 0; k < n; ++k) {
    if (k == j) continue;
    m = __dmul_rn(m, __dsub_rn(xj, x[k]));
    int ex;
    m = frexp(m, &ex);
    e += ex;
  }
  mant[j] = m;   // prod = m * 2^e = 1 / w_j
  expo[j] = e;
}

__global__ void bary_offdiag_kernel(const double* __restrict__ x, const double* __restrict__ mant,
                                    const int* __restrict__ expo, int n, int l, double* D) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= int64_t(n) * n) return;
  const int i = int(idx / n), j = int(idx % n);
  if (i == j) return;
  // w_j / w_i = prod_i / prod_j
  const double cr = ldexp(__ddiv_rn(mant[i], mant[j]), expo[i] - expo[j]);
  const double dx = __dsub_rn(x[i], x[j]);
  if (l == 1) {
    D[idx] = __ddiv_rn(cr, dx);
  } else {
    const double t = __dsub_rn(__dmul_rn(cr, D[int64_t(i) * n + i]), D[idx]);
    D[idx] = __dmul_rn(__ddiv_rn(double(l), dx), t);
  }
}

// ---- K6: Fourier, any order, by direct summation of the symbol (i k s)^p ----------------------------
__global__ void four_general_kernel(int n, double s, int order, double* __restrict__ col) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const int K = (n - 1) / 2;               // plain modes 1..K; n even adds the Nyquist mode n/2
  const bool even_p = (order % 2) == 0;
  const double sgn = even_p ? ((order / 2) % 2 ? -1.0 : 1.0) : (((order - 1) / 2) % 2 ? 1.0 : -1.0);
  double acc = 0.0;
  for (int k = 1; k <= K; ++k) {
    double ks = __dmul_rn(double(k), s), pw = 1.0;
    for (int q = 0; q < order; ++q) pw = __dmul_rn(pw, ks);
    const long long r = (2LL * k * m) % (2LL * n);        // theta = pi * r / n, exact reduction
    double sn, cs;
    sincospi(__ddiv_rn(double(r), double(n)), &sn, &cs);
    acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(2.0 * sgn, pw), even_p ? cs : sn));
  }
  if ((n % 2) == 0 && even_p) {
    double ks = __dmul_rn(double(n / 2), s), pw = 1.0;
    for (int q = 0; q < order; ++q) pw = __dmul_rn(pw, ks);
    acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(sgn, pw), (m % 2) ? -1.0 : 1.0));
  }
  col[m] = __ddiv_rn(acc, double(n));
}

__global__ void circulant_kernel(const double* __restrict__ col, int n, double* __restrict__ D) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= int64_t(n) * n) return;
  int k = int(idx / n) - int(idx % n);
  if (k < 0) k += n;
  D[idx] = col[k];
}

// ---- K6: Fornberg finite-difference weights --------------------------------------------------------
constexpr int FD_MAXW = 16;
constexpr int FD_MAXM = 6;

// weights for derivative `m` at z from the nodes xs[0..N) (Fornberg 1988), result in out[0..N)
__device__ void fornberg_row(const double* xs, int N, double z, int m, double* out) {
  double C[FD_MAXW][FD_MAXM + 1];
  for (int a = 0; a < N; ++a)
    for (int k = 0; k <= m; ++k) C[a][k] = 0.0;
  C[0][0] = 1.0;
  double c1 = 1.0;
  double c4 = __dsub_rn(xs[0], z);
  for (int i = 1; i < N; ++i) {
    const int mn = i < m ? i : m;
    double c2 = 1.0;
    const double c5 = c4;
    c4 = __dsub_rn(xs[i], z);
    for (int j = 0; j < i; ++j) {
      const double c3 = __dsub_rn(xs[i], xs[j]);
      c2 = __dmul_rn(c2, c3);
      if (j == i - 1) {
        for (int k = mn; k >= 1; --k) {
          const double num = __dsub_rn(__dmul_rn(double(k), C[i - 1][k - 1]), __dmul_rn(c5, C[i - 1][k]));
          C[i][k] = __ddiv_rn(__dmul_rn(c1, num), c2);
        }
        C[i][0] = __ddiv_rn(__dmul_rn(__dmul_rn(-c1, c5), C[i - 1][0]), c2);
      }
      for (int k = mn; k >= 1; --k) {
        const double num = __dsub_rn(__dmul_rn(c4, C[j][k]), __dmul_rn(double(k), C[j][k - 1]));
        C[j][k] = __ddiv_rn(num, c3);
      }
      C[j][0] = __ddiv_rn(__dmul_rn(c4, C[j][0]), c3);
    }
    c1 = c2;
  }
  for (int a = 0; a < N; ++a) out[a] = C[a][m];
}

__global__ void fornberg_kernel(const double* __restrict__ x, int n, int m, int w, int wb, int wmax,
                                int periodic, int uniform, double h, int32_t* __restrict__ start,
                                double* __restrict__ W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int hw = w / 2;
  double xs[FD_MAXW], wt[FD_MAXW];
  int s0, first, cnt, slot0;  // start, first node index, node count, first slot in the row
  if (periodic) {
    s0 = i - hw; first = i - hw; cnt = w; slot0 = 0;
  } else {
    s0 = i - hw;
    if (s0 < 0) s0 = 0;
    if (s0 > n - wmax) s0 = n - wmax;
    if (i - hw < 0) { first = 0; cnt = wb; }
    else if (i + hw > n - 1) { first = n - wb; cnt = wb; }
    else { first = i - hw; cnt = w; }
    slot0 = first - s0;
  }
  double z;
  if (uniform) {
    for (int a = 0; a < cnt; ++a) xs[a] = double(a);
    z = double(i - first);
  } else {
    const double period = __dmul_rn(double(n), h);  // periodic axes are equispaced: n*h
    for (int a = 0; a < cnt; ++a) {
      int j = first + a;
      double shift = 0.0;
      if (periodic) {
        if (j < 0) { j += n; shift = -period; }
        else if (j >= n) { j -= n; shift = period; }
      }
      xs[a] = __dadd_rn(x[j], shift);
    }
    z = x[i];
  }
  fornberg_row(xs, cnt, z, m, wt);
  if (uniform) {
    double hm = h;
    for (int k = 1; k < m; ++k) hm = __dmul_rn(hm, h);
    for (int a = 0; a < cnt; ++a) wt[a] = __ddiv_rn(wt[a], hm);
  }
  start[i] = s0;
  double* row = W + int64_t(i) * wmax;
  for (int s = 0; s < wmax; ++s) {
    const int a = s - slot0;
    row[s] = (a >= 0 && a < cnt) ? wt[a] : 0.0;
  }
}

__global__ void band_to_dense_kernel(const double* __restrict__ W, const int32_t* __restrict__ start,
                                     int wmax, int n, int periodic, double* __restrict__ D) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= int64_t(n) * n) return;
  const int i = int(idx / n), j = int(idx % n);
  double v = 0.0;
  const int s0 = start[i];
  for (int s = 0; s < wmax; ++s) {
    int c = s0 + s;
    if (periodic) { c %= n; if (c < 0) c += n; }
    if (c == j) v += W[int64_t(i) * wmax + s];
  }
  D[idx] = v;
}

__global__ void kron_lift_kernel(const double* __restrict__ D, int n, int64_t before, int64_t P,
                                 double* __restrict__ L) {
  const int64_t total = P * P;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / P, c = idx % P;
    const int64_t br = r % before, bc = c % before;
    const int64_t ir = (r / before) % n, ic = (c / before) % n;
    const int64_t ar = r / (before * n), ac = c / (before * n);
    L[idx] = (br == bc && ar == ac) ? D[ir * n + ic] : 0.0;
  }
}

// ---- K7: 3-D permutation copy ----------------------------------------------------------------------
struct PermParams {
  int64_t in_dim[3];      // input extents, slowest first
  int64_t in_stride[3];   // input strides (elements), slowest first
  int p[3];               // output axis j (slowest first) is input axis p[j]
};

// General case: out index (o2, o1, o0) -> in index; coalesced writes.
__global__ void __launch_bounds__(256) permute_naive_kernel(const double* __restrict__ in,
                                                             double* __restrict__ out, PermParams q) {
  const int64_t e0 = q.in_dim[q.p[2]], e1 = q.in_dim[q.p[1]], e2 = q.in_dim[q.p[0]];
  const int64_t total = e0 * e1 * e2;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t o0 = idx % e0, o1 = (idx / e0) % e1, o2 = idx / (e0 * e1);
    out[idx] = in[o2 * q.in_stride[q.p[0]] + o1 * q.in_stride[q.p[1]] + o0 * q.in_stride[q.p[2]]];
  }
}

// Tiled transpose of the two axes (x = input fastest axis, y = the input axis that becomes the
// output fastest axis) through a 32x33 shared tile; z enumerates the remaining axis.
__global__ void __launch_bounds__(256) permute_tiled_kernel(const double* __restrict__ in,
                                                             double* __restrict__ out, int64_t nx,
                                                             int64_t ny, int64_t nz, int64_t in_sy,
                                                             int64_t in_sz, int64_t out_sx,
                                                             int64_t out_sz) {
  __shared__ double tile[32][33];
  const int64_t tx = (nx + 31) / 32, ty = (ny + 31) / 32;
  for (int64_t blk = blockIdx.x; blk < tx * ty * nz; blk += gridDim.x) {
    const int64_t z = blk / (tx * ty);
    const int64_t r = blk % (tx * ty);
    const int64_t x0 = (r % tx) * 32, y0 = (r / tx) * 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t y = y0 + ly + 8 * k, x = x0 + lx;
      if (x < nx && y < ny) tile[ly + 8 * k][lx] = in[z * in_sz + y * in_sy + x];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t x = x0 + ly + 8 * k, y = y0 + lx;
      if (x < nx && y < ny) out[z * out_sz + x * out_sx + y] = tile[lx][ly + 8 * k];
    }
    __syncthreads();
  }
}

inline unsigned blocks_for(int64_t n, int threads) { return unsigned((n + threads - 1) / threads); }

template <typename T>
int launch_expand(const void* axes, void* out, const ExpandParams& p, cudaStream_t st) {
  if (p.P == 0) return PT_OK;
  int64_t bx = (int64_t(p.P) + 255) / 256;
  const int64_t cap = int64_t(pt::sm_count()) * 32;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)bx, (unsigned)p.ndim);
  if (p.P < (1ull << 31))
    expand_kernel<T, uint32_t><<<grid, 256, 0, st>>>(static_cast<const T*>(axes), static_cast<T*>(out), p);
  else
    expand_kernel<T, uint64_t><<<grid, 256, 0, st>>>(static_cast<const T*>(axes), static_cast<T*>(out), p);
  return pt::check_launch("pt_grid_expand");
}

}  // namespace

extern "C" int pt_gen_linspace(double* x, double start, double stop, int64_t n, pt_stream_t stream) {
  PT_REQUIRE(n >= 0, "pt_gen_linspace: n must be non-negative");
  if (n == 0) return PT_OK;
  PT_REQUIRE(x, "pt_gen_linspace: null pointer");
  linspace_kernel<<<blocks_for(n, 256), 256, 0, pt::as_stream(stream)>>>(x, start, stop, n);
  return pt::check_launch("pt_gen_linspace");
}

extern "C" int pt_gen_cheb(double* x, int64_t n, int mapped, double xmin, double xmax,
                           pt_stream_t stream) {
  PT_REQUIRE(n >= 0, "pt_gen_cheb: n must be non-negative");
  if (n == 0) return PT_OK;
  PT_REQUIRE(x, "pt_gen_cheb: null pointer");
  cheb_kernel<<<blocks_for(n, 256), 256, 0, pt::as_stream(stream)>>>(x, n, mapped, xmin, xmax);
  return pt::check_launch("pt_gen_cheb");
}

extern "C" int pt_gen_cheb_libm(double* x, int64_t n, int mapped, double xmin, double xmax, int libm,
                                pt_stream_t stream) {
  PT_REQUIRE(n >= 0, "pt_gen_cheb_libm: n must be non-negative");
  PT_REQUIRE(libm == PT_LIBM_GLIBC_FMA || libm == PT_LIBM_GLIBC_NOFMA, "pt_gen_cheb_libm: unknown libm model %d", libm);
  if (n == 0) return PT_OK;
  PT_REQUIRE(x, "pt_gen_cheb_libm: null pointer");
  if (libm == PT_LIBM_GLIBC_FMA)
    cheb_libm_kernel<true><<<blocks_for(n, 256), 256, 0, pt::as_stream(stream)>>>(x, n, mapped, xmin, xmax);
  else
    cheb_libm_kernel<false><<<blocks_for(n, 256), 256, 0, pt::as_stream(stream)>>>(x, n, mapped, xmin, xmax);
  return pt::check_launch("pt_gen_cheb_libm");
}

extern "C" int pt_cos_libm(const double* t, double* y, int64_t n, int libm, pt_stream_t stream) {
  PT_REQUIRE(n >= 0, "pt_cos_libm: n must be non-negative");
  PT_REQUIRE(libm == PT_LIBM_GLIBC_FMA || libm == PT_LIBM_GLIBC_NOFMA, "pt_cos_libm: unknown libm model %d", libm);
  if (n == 0) return PT_OK;
  PT_REQUIRE(t && y, "pt_cos_libm: null pointer");
  if (libm == PT_LIBM_GLIBC_FMA) cos_libm_kernel<true><<<blocks_for(n, 256), 256, 0, pt::as_stream(stream)>>>(t, y, n);
  else cos_libm_kernel<false><<<blocks_for(n, 256), 256, 0, pt::as_stream(stream)>>>(t, y, n);
  return pt::check_launch("pt_cos_libm");
}

extern "C" int pt_cheb_map(double* x, int64_t n, double xmin, double xmax, pt_stream_t stream) {
  PT_REQUIRE(n >= 0, "pt_cheb_map: n must be non-negative");
  if (n == 0) return PT_OK;
  PT_REQUIRE(x, "pt_cheb_map: null pointer");
  cheb_map_kernel<<<blocks_for(n, 256), 256, 0, pt::as_stream(stream)>>>(x, n, xmin, xmax);
  return pt::check_launch("pt_cheb_map");
}

extern "C" int pt_grid_expand(const void* axes, const int64_t* npts_host, int ndim, int elem_size,
                              void* out, pt_stream_t stream) {
  PT_REQUIRE(ndim >= 1 && ndim <= MAXDIM, "pt_grid_expand: ndim=%d outside [1,%d]", ndim, MAXDIM);
  PT_REQUIRE(npts_host, "pt_grid_expand: null npts");
  ExpandParams p;
  p.ndim = ndim;
  uint64_t stride = 1, off = 0;
  for (int d = 0; d < ndim; ++d) {
    PT_REQUIRE(npts_host[d] >= 0, "pt_grid_expand: negative npts[%d]", d);
    p.stride[d] = stride; p.n[d] = uint64_t(npts_host[d]); p.offset[d] = off;
    stride *= uint64_t(npts_host[d]);
    off += uint64_t(npts_host[d]);
  }
  p.P = stride;
  if (p.P == 0) return PT_OK;
  PT_REQUIRE(axes && out, "pt_grid_expand: null pointer");
  cudaStream_t st = pt::as_stream(stream);
  switch (elem_size) {
    case 1: return launch_expand<uint8_t>(axes, out, p, st);
    case 2: return launch_expand<uint16_t>(axes, out, p, st);
    case 4: return launch_expand<uint32_t>(axes, out, p, st);
    case 8: return launch_expand<uint64_t>(axes, out, p, st);
    case 16: return launch_expand<ulonglong2>(axes, out, p, st);
    default: return pt::fail(PT_ERR_UNSUPPORTED, "pt_grid_expand: elem_size=%d", elem_size);
  }
}

extern "C" int pt_grid_axpoints(const double* grid_row, int64_t stride, int64_t n, double* x,
                                pt_stream_t stream) {
  PT_REQUIRE(n >= 0 && stride >= 1, "pt_grid_axpoints: bad shape");
  if (n == 0) return PT_OK;
  PT_REQUIRE(grid_row && x, "pt_grid_axpoints: null pointer");
  axpoints_kernel<<<blocks_for(n, 256), 256, 0, pt::as_stream(stream)>>>(grid_row, stride, n, x);
  return pt::check_launch("pt_grid_axpoints");
}

extern "C" int pt_gen_chebmat(const double* x, int n, int order, double* D, pt_stream_t stream) {
  PT_REQUIRE(x && D, "pt_gen_chebmat: null pointer");
  PT_REQUIRE(n >= 2, "pt_gen_chebmat: need at least 2 nodes, got %d", n);
  PT_REQUIRE(order >= 1 && order <= 8, "pt_gen_chebmat: order=%d outside [1,8]", order);
  cudaStream_t st = pt::as_stream(stream);
  for (int l = 1; l <= order; ++l) {
    cheb_offdiag_kernel<<<blocks_for(int64_t(n) * n, 256), 256, 0, st>>>(x, n, l, D);
    cheb_diag_kernel<<<blocks_for(n, 64), 64, 0, st>>>(n, D);
  }
  return pt::check_launch("pt_gen_chebmat");
}


extern "C" int pt_flip_symmetrize(double* D, int n, int sign, pt_stream_t stream) {
  PT_REQUIRE(D, "pt_flip_symmetrize: null pointer");
  PT_REQUIRE(n >= 1 && (sign == 1 || sign == -1), "pt_flip_symmetrize: bad argument");
  flip_kernel<<<blocks_for(int64_t(n) * n, 256), 256, 0, pt::as_stream(stream)>>>(D, n, sign);
  return pt::check_launch("pt_flip_symmetrize");
}

extern "C" int pt_gen_barymat(const double* x, int n, int order, double* D, double* work,
                              pt_stream_t stream) {
  PT_REQUIRE(x && D && work, "pt_gen_barymat: null pointer");
  PT_REQUIRE(n >= 2, "pt_gen_barymat: need at least 2 nodes, got %d", n);
  PT_REQUIRE(order >= 1 && order <= 8, "pt_gen_barymat: order=%d outside [1,8]", order);
  cudaStream_t st = pt::as_stream(stream);
  double* mant = work;
  int* expo = reinterpret_cast<int*>(work + n);
  bary_weights_kernel<<<blocks_for(n, 64), 64, 0, st>>>(x, n, mant, expo);
  for (int l = 1; l <= order; ++l) {
    bary_offdiag_kernel<<<blocks_for(int64_t(n) * n, 256), 256, 0, st>>>(x, mant, expo, n, l, D);
    cheb_diag_kernel<<<blocks_for(n, 64), 64, 0, st>>>(n, D);
  }
  return pt::check_launch("pt_gen_barymat");
}

extern "C" int pt_gen_fourmat(int n, double period, int order, double* D, pt_stream_t stream) {
  PT_REQUIRE(D, "pt_gen_fourmat: null pointer");
  PT_REQUIRE(n >= 2, "pt_gen_fourmat: need at least 2 nodes, got %d", n);
  PT_REQUIRE(order == 1 || order == 2, "pt_gen_fourmat: order=%d (only 1 and 2 are generated)", order);
  PT_REQUIRE(period > 0.0, "pt_gen_fourmat: period must be positive");
  const double s = (2.0 * kPi) / period;
  fourmat_kernel<<<blocks_for(int64_t(n) * n, 256), 256, 0, pt::as_stream(stream)>>>(n, s, order, D);
  return pt::check_launch("pt_gen_fourmat");
}

extern "C" int pt_gen_fourmat_general(int n, double period, int order, double* D, double* work,
                                      pt_stream_t stream) {
  PT_REQUIRE(D && work, "pt_gen_fourmat_general: null pointer");
  PT_REQUIRE(n >= 2, "pt_gen_fourmat_general: need at least 2 nodes, got %d", n);
  PT_REQUIRE(order >= 1 && order <= 8, "pt_gen_fourmat_general: order=%d outside [1,8]", order);
  PT_REQUIRE(period > 0.0, "pt_gen_fourmat_general: period must be positive");
  const double s = (2.0 * kPi) / period;
  cudaStream_t st = pt::as_stream(stream);
  four_general_kernel<<<blocks_for(n, 64), 64, 0, st>>>(n, s, order, work);
  circulant_kernel<<<blocks_for(int64_t(n) * n, 256), 256, 0, st>>>(work, n, D);
  return pt::check_launch("pt_gen_fourmat_general");
}

extern "C" int pt_gen_fornberg(const double* x, int n, int order, int accuracy, int periodic,
                               int uniform, double h, int32_t* start, double* W, int wmax,
                               pt_stream_t stream) {
  PT_REQUIRE(start && W, "pt_gen_fornberg: null pointer");
  PT_REQUIRE(uniform || x, "pt_gen_fornberg: nodes required for a non-uniform axis");
  PT_REQUIRE(order >= 1 && order <= FD_MAXM, "pt_gen_fornberg: order=%d outside [1,%d]", order, FD_MAXM);
  PT_REQUIRE(accuracy >= 1, "pt_gen_fornberg: accuracy must be positive");
  const int w = 2 * ((order + 1) / 2) - 1 + accuracy;
  const int wb = order + accuracy;
  PT_REQUIRE(w % 2 == 1, "pt_gen_fornberg: accuracy=%d must be even (central stencils)", accuracy);
  const int need = periodic ? w : (w > wb ? w : wb);
  PT_REQUIRE(need <= FD_MAXW, "pt_gen_fornberg: stencil width %d exceeds %d", need, FD_MAXW);
  PT_REQUIRE(wmax == need, "pt_gen_fornberg: wmax=%d, expected %d", wmax, need);
  PT_REQUIRE(n >= need, "pt_gen_fornberg: n=%d smaller than the stencil width %d", n, need);
  fornberg_kernel<<<blocks_for(n, 64), 64, 0, pt::as_stream(stream)>>>(x, n, order, w, wb, wmax,
                                                                       periodic ? 1 : 0, uniform ? 1 : 0,
                                                                       h, start, W);
  return pt::check_launch("pt_gen_fornberg");
}

extern "C" int pt_band_to_dense(const double* W, const int32_t* start, int wmax, int n, int periodic,
                                double* D, pt_stream_t stream) {
  PT_REQUIRE(W && start && D, "pt_band_to_dense: null pointer");
  PT_REQUIRE(n >= 1 && wmax >= 1, "pt_band_to_dense: bad shape");
  band_to_dense_kernel<<<blocks_for(int64_t(n) * n, 256), 256, 0, pt::as_stream(stream)>>>(
      W, start, wmax, n, periodic ? 1 : 0, D);
  return pt::check_launch("pt_band_to_dense");
}

extern "C" int pt_kron_lift(const double* D, int n, int64_t before, int64_t after, double* L,
                            pt_stream_t stream) {
  PT_REQUIRE(D && L, "pt_kron_lift: null pointer");
  PT_REQUIRE(n >= 1 && before >= 1 && after >= 1, "pt_kron_lift: bad shape");
  const int64_t P = after * n * before;
  PT_REQUIRE(P <= 46340, "pt_kron_lift: lifted matrix %lld^2 is too large to materialise", (long long)P);
  int64_t blocks = (P * P + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  kron_lift_kernel<<<unsigned(blocks), 256, 0, pt::as_stream(stream)>>>(D, n, before, P, L);
  return pt::check_launch("pt_kron_lift");
}

extern "C" int pt_permute3d(const double* in, double* out, int64_t d2, int64_t d1, int64_t d0,
                            int perm, pt_stream_t stream) {
  PT_REQUIRE(in && out, "pt_permute3d: null pointer");
  PT_REQUIRE(d0 >= 0 && d1 >= 0 && d2 >= 0, "pt_permute3d: negative extent");
  PT_REQUIRE(perm >= 0 && perm < 27, "pt_permute3d: bad perm code %d", perm);
  PermParams q;
  q.in_dim[0] = d2; q.in_dim[1] = d1; q.in_dim[2] = d0;
  q.in_stride[0] = d1 * d0; q.in_stride[1] = d0; q.in_stride[2] = 1;
  q.p[0] = perm % 3; q.p[1] = (perm / 3) % 3; q.p[2] = perm / 9;
  PT_REQUIRE(q.p[0] != q.p[1] && q.p[0] != q.p[2] && q.p[1] != q.p[2], "pt_permute3d: perm %d is not a permutation", perm);
  const int64_t total = d0 * d1 * d2;
  if (total == 0) return PT_OK;
  cudaStream_t st = pt::as_stream(stream);
  const int64_t cap = int64_t(pt::sm_count()) * 16;
  if (q.p[2] == 2) {
    // fastest axis unchanged: rows are copied whole, the naive kernel is already coalesced
    int64_t blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    permute_naive_kernel<<<unsigned(blocks), 256, 0, st>>>(in, out, q);
    return pt::check_launch("pt_permute3d");
  }
  // fastest output axis is input axis y = p[2]; x = input axis 2 lands on output position jx
  const int y = q.p[2];
  const int z = 3 - 2 - y;  // remaining input axis (0 or 1)
  int64_t out_extent[3] = {q.in_dim[q.p[0]], q.in_dim[q.p[1]], q.in_dim[q.p[2]]};
  int64_t out_stride[3] = {out_extent[1] * out_extent[2], out_extent[2], 1};
  int jx = (q.p[0] == 2) ? 0 : 1, jz = (q.p[0] == z) ? 0 : 1;
  const int64_t nx = d0, ny = q.in_dim[y], nz = q.in_dim[z];
  const int64_t tiles = ((nx + 31) / 32) * ((ny + 31) / 32) * nz;
  int64_t blocks = tiles < cap ? tiles : cap;
  permute_tiled_kernel<<<unsigned(blocks), 256, 0, st>>>(in, out, nx, ny, nz, q.in_stride[y],
                                                        q.in_stride[z], out_stride[jx], out_stride[jz]);
  return pt::check_launch("pt_permute3d");
}
