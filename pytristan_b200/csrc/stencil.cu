// K3/K4 — banded (finite-difference) operator application, HBM-bound.
//
//   Y[a, i, b] = alpha * sum_s W[i][s] * U[a, start[i]+s, b] + beta * Y[a, i, b]
//
// Three kernels:
//   stencil_table      any band (per-row weights, one-sided closures, periodic wrap); one thread
//                      per output point.  Used for non-uniform (mapped) axes and as the
//                      boundary-row path of the two fast kernels.
//   stencil_strided    (K4) before > 1, canonical interior rows with uniform weights: threads
//                      run along the contiguous index b (16-byte vectors), march along the
//                      differentiated axis with a register sliding window: every U element is
//                      loaded once and reused w times, every Y element written once (16 B/pt).
//   stencil_contig     (K3) before == 1: a (rows x segment) tile with halo is staged in shared
//                      memory with coalesced 16-byte loads; each lane produces two adjacent
//                      points from five conflict-free LDS.128.
// Algorithmic traffic is 16 bytes per point per derivative (read U once, write Y once);
// DESIGN.md §4.2 states the roofline.
#include "common.cuh"

namespace {

constexpr int MAXW = 16;

constexpr int MAXBND = 8;    // boundary rows that can ride in the kernel parameters
constexpr int MAXBW = 10;    // widest boundary stencil kept there

struct UniWeights {
  double w[MAXW];
  double wb[MAXBND][MAXBW];  // boundary rows: top rows first, then bottom rows
  int sb[MAXBND];
  int n_bnd;                 // 0: boundary rows are not in the parameters
};

struct StencilParams {
  const double* __restrict__ W;
  const int32_t* __restrict__ start;
  const double* __restrict__ U;
  double* __restrict__ Y;
  int wmax, n;       // n = rows of U per slab (wrap/bounds)
  int n_out, shift;  // rows of Y per slab; output row i sits at input row i + shift
  int64_t after, before;
  int periodic;
  int lo, hi;  // canonical interior rows [lo, hi) (fast kernels); table rows elsewhere
  double alpha, beta;
  // peer-halo form (pt_stencil_apply_halo): the input slab is virtual — halo_lo rows read from Ulo
  // (the lower neighbour's memory), then the n - halo_lo - halo_hi own rows of U, then halo_hi rows
  // from Uhi.  lo_stride / hi_stride: distance between consecutive slabs `a` in those buffers.
  const double* __restrict__ Ulo;
  const double* __restrict__ Uhi;
  int halo_lo, halo_hi;
  int64_t lo_stride, hi_stride;
};

// address of (slab a, virtual input row j, column 0) in the peer-halo form
__device__ __forceinline__ const double* halo_row(const StencilParams& p, int64_t a, int j) {
  const int n_own = p.n - p.halo_lo - p.halo_hi;
  if (j < p.halo_lo) return p.Ulo + a * p.lo_stride + int64_t(j) * p.before;
  j -= p.halo_lo;
  if (j < n_own) return p.U + (a * n_own + j) * p.before;
  return p.Uhi + a * p.hi_stride + int64_t(j - n_own) * p.before;
}

__device__ __forceinline__ int wrap_index(int j, int n) {
  j %= n;
  return j < 0 ? j + n : j;
}

// one table row for one column: base points at U[a, 0, b]
__device__ __forceinline__ double table_point(const StencilParams& p, const double* base, int i) {
  const double* w = p.W + int64_t(i) * p.wmax;
  const int s0 = p.start[i];
  double acc = 0.0;
  for (int s = 0; s < p.wmax; ++s) {
    int j = s0 + s;
    if (p.periodic) j = wrap_index(j, p.n);
    acc = fma(w[s], base[int64_t(j) * p.before], acc);
  }
  return acc;
}
__device__ __forceinline__ double table_point_halo(const StencilParams& p, int64_t a, int64_t b, int i) {
  const double* w = p.W + int64_t(i) * p.wmax;
  const int s0 = p.start[i];
  double acc = 0.0;
  for (int s = 0; s < p.wmax; ++s) acc = fma(w[s], halo_row(p, a, s0 + s)[b], acc);
  return acc;
}

__device__ __forceinline__ void store_point(const StencilParams& p, double* y, double v) {
  v *= p.alpha;
  if (p.beta != 0.0) v = fma(p.beta, *y, v);
  *y = v;
}

// ---- generic table kernel ------------------------------------------------------------------------
// output rows i outside [skip_lo, skip_hi) of every slab; start[i] is relative to the input slab
template <bool HALO>
__global__ void __launch_bounds__(256) stencil_table_kernel(const StencilParams p, int skip_lo,
                                                             int skip_hi) {
  // rows that need the table per slab: i in [0, skip_lo) U [skip_hi, n)
  const int per_slab = skip_lo + (p.n_out - skip_hi);
  const int64_t total = p.after * int64_t(per_slab) * p.before;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = idx % p.before;
    const int64_t r = idx / p.before;
    const int ii = int(r % per_slab);
    const int64_t a = r / per_slab;
    const int i = ii < skip_lo ? ii : ii - skip_lo + skip_hi;
    const double v = HALO ? table_point_halo(p, a, b, i)
                          : table_point(p, p.U + a * int64_t(p.n) * p.before + b, i);
    store_point(p, p.Y + (a * int64_t(p.n_out) + i) * p.before + b, v);
  }
}

// ---- K4: strided axis, register sliding window -----------------------------------------------------
template <int V> struct Vec;
template <> struct Vec<1> { using T = double; };
template <> struct Vec<2> { using T = double2; };

__device__ __forceinline__ double vfma(double w, double x, double acc) { return fma(w, x, acc); }
__device__ __forceinline__ double2 vfma(double w, double2 x, double2 acc) {
  return make_double2(fma(w, x.x, acc.x), fma(w, x.y, acc.y));
}
__device__ __forceinline__ double vzero(double) { return 0.0; }
__device__ __forceinline__ double2 vzero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double vaxpby(double al, double v, double be, double y) {
  return fma(be, y, al * v);
}
__device__ __forceinline__ double2 vaxpby(double al, double2 v, double be, double2 y) {
  return make_double2(fma(be, y.x, al * v.x), fma(be, y.y, al * v.y));
}
__device__ __forceinline__ double vscale(double al, double v) { return al * v; }
__device__ __forceinline__ double2 vscale(double al, double2 v) {
  return make_double2(al * v.x, al * v.y);
}

// Thread = one vector column (V doubles along b) of one slab a, rows [c0, c1) of the canonical
// range.  grid.x = column chunks, grid.y = i-chunks, grid.z = a (folded).
template <int W, int V, bool HALO>
__global__ void __launch_bounds__(128) stencil_strided_kernel(const StencilParams p,
                                                               const UniWeights uw, int ichunk,
                                                               int64_t ncolv) {
  using T = typename Vec<V>::T;
  constexpr int HW = W / 2;
  const int64_t colv = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;  // vector column in slab
  if (colv >= ncolv) return;
  const int64_t a = blockIdx.z;
  const int64_t stride = p.before / V;  // row stride in vectors
  const T* __restrict__ u = reinterpret_cast<const T*>(p.U + a * int64_t(p.n) * p.before) + colv;
  T* __restrict__ y = reinterpret_cast<T*>(p.Y + a * int64_t(p.n_out) * p.before) + colv;
  const int c0 = p.lo + blockIdx.y * ichunk;
  const int c1 = min(c0 + ichunk, p.hi);
  if (c0 >= c1) return;
  const int n = p.n;
  const bool use_beta = p.beta != 0.0;

  double w[W];
#pragma unroll
  for (int s = 0; s < W; ++s) w[s] = uw.w[s];

  auto row = [&](int j) -> T {
    if (HALO) return reinterpret_cast<const T*>(halo_row(p, a, j))[colv];
    if (p.periodic) { if (j < 0) j += n; else if (j >= n) j -= n; }
    return u[int64_t(j) * stride];
  };

  T v[2 * W - 1];
#pragma unroll
  for (int s = 0; s < W - 1; ++s) v[s] = row(c0 + p.shift - HW + s);

  for (int i = c0; i < c1; i += W) {
    // W independent loads in flight per thread
#pragma unroll
    for (int q = 0; q < W; ++q) {
      const int j = i + q + p.shift + HW;
      if (i + q < c1) v[W - 1 + q] = row(j);
    }
#pragma unroll
    for (int q = 0; q < W; ++q) {
      if (i + q < c1) {
        T acc = vzero(T());
#pragma unroll
        for (int s = 0; s < W; ++s) acc = vfma(w[s], v[q + s], acc);
        T* dst = y + int64_t(i + q) * stride;
        if (use_beta) *dst = vaxpby(p.alpha, acc, p.beta, *dst);
        else *dst = vscale(p.alpha, acc);
      }
    }
#pragma unroll
    for (int s = 0; s < W - 1; ++s) v[s] = v[W + s];
  }
}

// ---- K3: contiguous axis, shared-memory tile ---------------------------------------------------------
// Block = 256 threads, tile = ROWS rows x TI points (TI even), halo HL = 4 doubles each side.
// V = 2: 16-byte cp.async staging (n_in, n_out, shift even, 16-byte aligned bases); V = 1: 8-byte.
// The tile is staged with cp.async (no register staging, all copies of a thread in flight at
// once); chunks that straddle the ends of the axis are filled by plain stores (zero or wrapped).
__device__ __forceinline__ void st_cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void st_cp_async8(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}

// one boundary row evaluated from the staged tile with the weights held in the kernel parameters
// (falls back to global memory for inputs that a tile boundary left outside the staged window)
__device__ __forceinline__ double boundary_from_tile(const StencilParams& p, const UniWeights& uw,
                                                     const double* srow, const double* rowp, int i,
                                                     int jbase, int pitch) {
  const int bi = i < p.lo ? i : p.lo + (i - p.hi);
  const int s0 = uw.sb[bi];
  double acc = 0.0;
  for (int s = 0; s < p.wmax; ++s) {
    const int j = s0 + s;
    const int l = j - jbase;
    const double u = (l >= 0 && l < pitch) ? srow[l] : rowp[j];
    acc = fma(uw.wb[bi][s], u, acc);
  }
  return acc;
}

template <int W, int V>
__global__ void __launch_bounds__(256) stencil_contig_kernel(const StencilParams p,
                                                              const UniWeights uw, int TI, int ROWS,
                                                              int tiles_per_row) {
  constexpr int HW = W / 2;
  constexpr int HL = 4;
  extern __shared__ __align__(16) double sm[];
  const int pitch = TI + 2 * HL;
  const int n = p.n;
  const int ti = int(blockIdx.x % tiles_per_row);
  const int64_t row0 = int64_t(blockIdx.x / tiles_per_row) * ROWS;
  const int i0 = ti * TI;
  const int tid = threadIdx.x;
  const int64_t rows_left = p.after - row0;
  const int nrows = rows_left < ROWS ? int(rows_left) : ROWS;
  const int jbase = i0 + p.shift - HL;   // input index of the first staged double of a row

  // ---- stage
  const int cpr = pitch / V;
  for (int r = 0; r < nrows; ++r) {
    const double* rowp = p.U + (row0 + r) * int64_t(n);
    double* srow = sm + r * pitch;
    for (int c = tid; c < cpr; c += 256) {
      const int j = jbase + c * V;
      double* dst = srow + c * V;
      if (V == 2) {
        if (j >= 0 && j + 1 < n) {
          st_cp_async16(dst, rowp + j);
        } else {
          double2 val = make_double2(0.0, 0.0);
          if (p.periodic) {
            val.x = rowp[wrap_index(j, n)];
            val.y = rowp[wrap_index(j + 1, n)];
          } else {
            if (j >= 0 && j < n) val.x = rowp[j];
            if (j + 1 >= 0 && j + 1 < n) val.y = rowp[j + 1];
          }
          *reinterpret_cast<double2*>(dst) = val;
        }
      } else {
        if (j >= 0 && j < n) st_cp_async8(dst, rowp + j);
        else *dst = p.periodic ? rowp[wrap_index(j, n)] : 0.0;
      }
    }
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();

  double w[W];
#pragma unroll
  for (int s = 0; s < W; ++s) w[s] = uw.w[s];
  const bool use_beta = p.beta != 0.0;

  // ---- compute: each item = 2 adjacent points
  auto rowp_of = [&](int r) { return p.U + (row0 + r) * int64_t(n); };
  const int ipr = TI / 2;  // items per row
  for (int r = 0; r < nrows; ++r) {
    const double* srow = sm + r * pitch;
    double* yrow = p.Y + (row0 + r) * int64_t(p.n_out);
    for (int it = tid; it < ipr; it += 256) {
      const int li = it * 2;
      const int i = i0 + li;
      if (i >= p.n_out) break;
      const double* s = srow + li;  // s[0] is input point i + shift - HL
      double v[10];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double2 t2 = *reinterpret_cast<const double2*>(s + 2 * k);
        v[2 * k] = t2.x; v[2 * k + 1] = t2.y;
      }
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int q = 0; q < W; ++q) {
        acc0 = fma(w[q], v[HL - HW + q], acc0);
        acc1 = fma(w[q], v[HL - HW + 1 + q], acc1);
      }
      double* yp = yrow + i;
      const bool ok0 = (i >= p.lo && i < p.hi);
      const bool ok1 = (i + 1 >= p.lo && i + 1 < p.hi);
      if (V == 2 && ok0 && ok1) {
        double2 o = make_double2(p.alpha * acc0, p.alpha * acc1);
        if (use_beta) {
          const double2 old = *reinterpret_cast<const double2*>(yp);
          o.x = fma(p.beta, old.x, o.x); o.y = fma(p.beta, old.y, o.y);
        }
        *reinterpret_cast<double2*>(yp) = o;
      } else {
        if (ok0) { double o = p.alpha * acc0; if (use_beta) o = fma(p.beta, yp[0], o); yp[0] = o; }
        if (ok1) { double o = p.alpha * acc1; if (use_beta) o = fma(p.beta, yp[1], o); yp[1] = o; }
      }
    }
  }

  // ---- boundary rows (one-sided closures): one thread per (row, boundary point), weights from
  // the kernel parameters, inputs from the same tile.  Kept out of the main loop so that no
  // warp of the streaming part diverges.
  if (uw.n_bnd > 0) {
    for (int t = tid; t < nrows * uw.n_bnd; t += 256) {
      const int r = t / uw.n_bnd;
      const int bi = t - r * uw.n_bnd;
      const int i = bi < p.lo ? bi : p.hi + (bi - p.lo);
      if (i < i0 || i >= i0 + TI) continue;   // another tile of the row owns this point
      const double v = boundary_from_tile(p, uw, sm + r * pitch, rowp_of(r), i, jbase, pitch);
      double* yp = p.Y + (row0 + r) * int64_t(p.n_out) + i;
      double o = p.alpha * v;
      if (use_beta) o = fma(p.beta, *yp, o);
      *yp = o;
    }
  }
}

inline bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

int launch_table(const StencilParams& p, int skip_lo, int skip_hi, cudaStream_t st) {
  const int per_slab = skip_lo + (p.n_out - skip_hi);
  const int64_t total = p.after * int64_t(per_slab) * p.before;
  if (total <= 0) return PT_OK;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = int64_t(pt::sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  if (p.Ulo || p.Uhi) stencil_table_kernel<true><<<unsigned(blocks), 256, 0, st>>>(p, skip_lo, skip_hi);
  else stencil_table_kernel<false><<<unsigned(blocks), 256, 0, st>>>(p, skip_lo, skip_hi);
  return pt::check_launch("pt_stencil_apply(table)");
}

template <int W>
int launch_strided(const StencilParams& p, const UniWeights& uw, cudaStream_t st) {
  const bool halo = p.Ulo || p.Uhi;
  const bool vec = (p.before % 2 == 0) && aligned16(p.U) && aligned16(p.Y) &&
                   (!halo || (aligned16(p.Ulo) && aligned16(p.Uhi) && !((p.lo_stride | p.hi_stride) & 1)));
  const int V = vec ? 2 : 1;
  const int64_t ncolv = p.before / V;
  const int64_t gx = (ncolv + 127) / 128;
  // split the marching axis only when there are too few columns to fill the GPU
  const int64_t want = int64_t(pt::sm_count()) * 16;
  const int rows = p.hi - p.lo;
  int64_t ysplit = 1;
  if (gx * p.after < want) ysplit = (want + gx * p.after - 1) / (gx * p.after);
  int ichunk = int((rows + ysplit - 1) / ysplit);
  if (ichunk < 8 * W) ichunk = 8 * W;         // keep halo re-reads below ~1/8
  if (ichunk > rows) ichunk = rows;
  const int gy = (rows + ichunk - 1) / ichunk;
  PT_REQUIRE(gx <= 2147483647LL && gy <= 65535, "pt_stencil_apply: grid too large");
  // grid.z is limited to 65535: fold larger `after` into several launches
  for (int64_t a0 = 0; a0 < p.after; a0 += 65535) {
    StencilParams q = p;
    const int64_t na = (p.after - a0 < 65535) ? p.after - a0 : 65535;
    q.U = p.U + a0 * int64_t(p.n - p.halo_lo - p.halo_hi) * p.before;
    q.Y = p.Y + a0 * int64_t(p.n_out) * p.before;
    if (p.Ulo) q.Ulo = p.Ulo + a0 * p.lo_stride;
    if (p.Uhi) q.Uhi = p.Uhi + a0 * p.hi_stride;
    q.after = na;
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)na);
    if (halo) {
      if (vec) stencil_strided_kernel<W, 2, true><<<grid, 128, 0, st>>>(q, uw, ichunk, ncolv);
      else stencil_strided_kernel<W, 1, true><<<grid, 128, 0, st>>>(q, uw, ichunk, ncolv);
    } else if (vec) stencil_strided_kernel<W, 2, false><<<grid, 128, 0, st>>>(q, uw, ichunk, ncolv);
    else stencil_strided_kernel<W, 1, false><<<grid, 128, 0, st>>>(q, uw, ichunk, ncolv);
    int rc = pt::check_launch("pt_stencil_apply(strided)");
    if (rc) return rc;
  }
  return PT_OK;
}

template <int W>
int launch_contig(const StencilParams& p, const UniWeights& uw, cudaStream_t st) {
  const bool vec = (p.n % 2 == 0) && (p.n_out % 2 == 0) && (p.shift % 2 == 0) && aligned16(p.U) && aligned16(p.Y);
  // tile: TI points per row segment (even), ROWS rows, about 4096 points per block
  int TI = p.n_out + (p.n_out & 1);
  if (TI > 2048) TI = 2048;
  int ROWS = 4096 / TI;
  if (ROWS > 5120 / (TI + 8)) ROWS = 5120 / (TI + 8);  // keep the tile under 40 KB of smem
  if (ROWS < 1) ROWS = 1;
  if (ROWS > p.after) ROWS = int(p.after);
  const int tiles_per_row = (p.n_out + TI - 1) / TI;
  const int64_t tiles = ((p.after + ROWS - 1) / ROWS) * tiles_per_row;
  PT_REQUIRE(tiles <= 2147483647LL, "pt_stencil_apply: grid too large");
  const size_t smem = size_t(ROWS) * (TI + 8) * sizeof(double);
  if (vec) stencil_contig_kernel<W, 2><<<unsigned(tiles), 256, smem, st>>>(p, uw, TI, ROWS, tiles_per_row);
  else stencil_contig_kernel<W, 1><<<unsigned(tiles), 256, smem, st>>>(p, uw, TI, ROWS, tiles_per_row);
  return pt::check_launch("pt_stencil_apply(contig)");
}

template <int W>
int launch_fast(const StencilParams& p, const UniWeights& uw, cudaStream_t st) {
  int rc = (p.before == 1) ? launch_contig<W>(p, uw, st) : launch_strided<W>(p, uw, st);
  if (rc || (p.before == 1 && uw.n_bnd > 0)) return rc;  // boundary rows were computed in the tile
  return launch_table(p, p.lo, p.hi, st);  // boundary rows
}

}  // namespace

namespace {
struct HaloArgs {
  const double* Ulo; const double* Uhi;
  int halo_lo, halo_hi;
  int64_t lo_stride, hi_stride;
};

int stencil_apply_impl(const double* W, const int32_t* start, int wmax, int n_in, int n_out, int shift,
                       const double* U, double* Y, int64_t after, int64_t before, int periodic,
                       const pt_uniform_rows* uni, double alpha, double beta, const HaloArgs& h,
                       pt_stream_t stream) {
  PT_REQUIRE(W && start && U && Y, "pt_stencil_apply: null pointer");
  PT_REQUIRE(wmax >= 1 && wmax <= MAXW, "pt_stencil_apply: wmax=%d outside [1,%d]", wmax, MAXW);
  PT_REQUIRE(n_in >= 1 && n_in >= wmax, "pt_stencil_apply: n=%d smaller than the stencil width %d", n_in, wmax);
  PT_REQUIRE(n_out >= 0 && shift >= 0 && shift + n_out <= n_in, "pt_stencil_apply: output rows [%d, %d) outside the input rows [0, %d)", shift, shift + n_out, n_in);
  PT_REQUIRE(!periodic || (n_out == n_in && shift == 0), "pt_stencil_apply: periodic wrap needs the whole axis");
  PT_REQUIRE(after >= 0 && before >= 1, "pt_stencil_apply: bad field shape");
  PT_REQUIRE(U != Y, "pt_stencil_apply: in-place application is not supported");
  if (after == 0 || n_out == 0) return PT_OK;
  StencilParams p;
  p.W = W; p.start = start; p.U = U; p.Y = Y;
  p.wmax = wmax; p.n = n_in; p.n_out = n_out; p.shift = shift; p.after = after; p.before = before;
  p.periodic = periodic ? 1 : 0;
  p.alpha = alpha; p.beta = beta;
  p.lo = 0; p.hi = 0;
  p.Ulo = h.Ulo; p.Uhi = h.Uhi; p.halo_lo = h.halo_lo; p.halo_hi = h.halo_hi;
  p.lo_stride = h.lo_stride; p.hi_stride = h.hi_stride;
  cudaStream_t st = pt::as_stream(stream);

  // canonical rows must read inside the input slab: i + shift - hw >= 0 and i + shift + hw < n_in
  const int w_uni = uni ? uni->w_uni : 0;
  const int lo = uni ? uni->lo : 0, hi = uni ? uni->hi : 0;
  const int hwu = w_uni / 2;
  const bool fast = uni != nullptr && uni->wuni != nullptr &&
                    (w_uni == 3 || w_uni == 5 || w_uni == 7 || w_uni == 9) &&
                    lo >= 0 && hi <= n_out && hi - lo >= 2 * w_uni &&
                    (periodic || (lo + shift >= hwu && hi + shift <= n_in - hwu));
  if (!fast) return launch_table(p, 0, 0, st);  // every row through the table

  UniWeights uw;
  for (int s = 0; s < MAXW; ++s) uw.w[s] = s < w_uni ? uni->wuni[s] : 0.0;
  uw.n_bnd = 0;
  const int n_bnd = lo + (n_out - hi);
  if (n_bnd > 0 && n_bnd <= MAXBND && wmax <= MAXBW && uni->n_bnd == n_bnd && uni->w_bnd && uni->start_bnd) {
    uw.n_bnd = n_bnd;
    for (int r = 0; r < n_bnd; ++r) {
      uw.sb[r] = uni->start_bnd[r];
      for (int s = 0; s < MAXBW; ++s) uw.wb[r][s] = s < wmax ? uni->w_bnd[r * wmax + s] : 0.0;
    }
  }
  p.lo = lo; p.hi = hi;
  switch (w_uni) {
    case 3: return launch_fast<3>(p, uw, st);
    case 5: return launch_fast<5>(p, uw, st);
    case 7: return launch_fast<7>(p, uw, st);
    default: return launch_fast<9>(p, uw, st);
  }
}
}  // namespace

extern "C" int pt_stencil_apply_ex(const double* W, const int32_t* start, int wmax, int n_in, int n_out,
                                   int shift, const double* U, double* Y, int64_t after, int64_t before,
                                   int periodic, const pt_uniform_rows* uni, double alpha, double beta,
                                   pt_stream_t stream) {
  const HaloArgs none{nullptr, nullptr, 0, 0, 0, 0};
  return stencil_apply_impl(W, start, wmax, n_in, n_out, shift, U, Y, after, before, periodic, uni,
                            alpha, beta, none, stream);
}

extern "C" int pt_stencil_apply_halo(const double* W, const int32_t* start, int wmax, int n_own,
                                     int halo_lo, int halo_hi, const double* U, const double* U_lo,
                                     int64_t lo_stride, const double* U_hi, int64_t hi_stride, double* Y,
                                     int64_t after, int64_t before, const pt_uniform_rows* uni,
                                     double alpha, double beta, pt_stream_t stream) {
  PT_REQUIRE(n_own >= 1 && halo_lo >= 0 && halo_hi >= 0, "pt_stencil_apply_halo: bad slab shape");
  PT_REQUIRE((halo_lo == 0 || U_lo) && (halo_hi == 0 || U_hi), "pt_stencil_apply_halo: halo buffer is null");
  PT_REQUIRE(before > 1, "pt_stencil_apply_halo: the sharded axis cannot be the contiguous one");
  if (halo_lo == 0 && halo_hi == 0)
    return pt_stencil_apply_ex(W, start, wmax, n_own, n_own, 0, U, Y, after, before, 0, uni, alpha, beta, stream);
  // an absent side still needs a non-null pointer to select the halo kernels; it is never read
  const HaloArgs h{halo_lo ? U_lo : U, halo_hi ? U_hi : U, halo_lo, halo_hi, lo_stride, hi_stride};
  return stencil_apply_impl(W, start, wmax, halo_lo + n_own + halo_hi, n_own, halo_lo, U, Y, after, before,
                            0, uni, alpha, beta, h, stream);
}

extern "C" int pt_stencil_apply(const double* W, const int32_t* start, int wmax, int n,
                                const double* U, double* Y, int64_t after, int64_t before,
                                int periodic, const pt_uniform_rows* uni, double alpha, double beta,
                                pt_stream_t stream) {
  return pt_stencil_apply_ex(W, start, wmax, n, n, 0, U, Y, after, before, periodic, uni, alpha, beta,
                             stream);
}

// ---- fused three-axis gradient: d/dx, d/dy, d/dz of a 3-D field in ONE read of U ------------------
// (8 B/pt read + 24 B/pt written = 32 B/pt instead of 3 x 16 B/pt.)
// 2.5-D blocking: a CTA owns a 64 (x) x 16 (y) tile and marches along z.  Every thread keeps the
// z-window of its two x-adjacent points in registers (dz), the current plane of the tile plus
// halos lives in a double-buffered shared tile (dx from the row, dy from the column).  Halo cells
// are prefetched one plane ahead.  Centred weights everywhere; the few one-sided boundary planes
// of each axis are rewritten afterwards by the table kernel.
namespace {

struct Grad3Params {
  const double* __restrict__ U;
  // peer-halo form: the HW planes below z = 0 / above z = nz - 1 live in the neighbours' memory
  // (Ulo points at plane -HW, Uhi at plane nz); nullptr = no neighbour (physical boundary)
  const double* __restrict__ Ulo;
  const double* __restrict__ Uhi;
  double* __restrict__ G[3];
  int nx, ny, nz;
  int zchunk;
  double w[3][MAXW];
};

template <int W, int TY, int TXT>
__global__ void __launch_bounds__(TXT * TY, 1024 / (TXT * TY)) stencil_grad3_kernel(const Grad3Params p) {
  constexpr int HW = W / 2;
  constexpr int HX = 4;              // x halo in doubles (even: keeps 16-byte alignment)
  constexpr int TXP = 2 * TXT;       // tile: TXT threads x double2 in x, TY rows in y
  constexpr int PITCH = TXP + 2 * HX;
  constexpr int ROWS = TY + 2 * HW;
  constexpr int PLANE = ROWS * PITCH;
  __shared__ __align__(16) double sm[2][PLANE];

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TXT + tx;
  const int xb = blockIdx.x * TXP, yb = blockIdx.y * TY;
  const int x0 = xb + 2 * tx, y = yb + ty;
  const int z0 = blockIdx.z * p.zchunk;
  const int z1 = min(z0 + p.zchunk, p.nz);
  const int64_t plane = int64_t(p.nx) * p.ny;
  const bool mine = (x0 < p.nx) && (y < p.ny);            // nx is even: x0 + 1 < nx as well
  const double2 zero2 = make_double2(0.0, 0.0);
  const int64_t col = int64_t(y) * p.nx + x0;
  // the two x-adjacent points of this thread in plane z; planes outside the slab come from the
  // neighbours (peer loads) or are zero
  auto ldz = [&](int z) -> double2 {
    const double* base;
    if (z < 0) base = p.Ulo ? p.Ulo + (z + HW) * plane : nullptr;
    else if (z >= p.nz) base = p.Uhi ? p.Uhi + (z - p.nz) * plane : nullptr;
    else base = p.U + z * plane;
    return (mine && base) ? *reinterpret_cast<const double2*>(base + col) : zero2;
  };

  // halo duty: 2*(TY) x-halo row pieces (2 double2 left, 2 right per row) + 2*HW full rows
  constexpr int XH = TY * 4;                 // double2 chunks of the x halos
  constexpr int YH = 2 * HW * (PITCH / 2);   // double2 chunks of the y halo rows
  static_assert(XH + YH <= TXT * TY, "one halo chunk per thread");
  int h_sm = -1;
  int64_t h_off = 0;
  bool h_ok = false;
  if (tid < XH + YH) {
    int row, col;  // position in the shared tile (row in [0,ROWS), col even in [0,PITCH))
    if (tid < XH) {
      row = HW + tid / 4;
      const int q = tid % 4;
      col = (q < 2) ? 2 * q : HX + TXP + 2 * (q - 2);
    } else {
      const int t = tid - XH;
      const int r = t / (PITCH / 2);
      row = r < HW ? r : TY + r;               // rows 0..HW-1 and TY+HW..TY+2HW-1
      col = 2 * (t % (PITCH / 2));
    }
    const int gx = xb - HX + col, gy = yb - HW + row;
    h_sm = row * PITCH + col;
    h_ok = (gx >= 0) && (gx + 1 < p.nx) && (gy >= 0) && (gy < p.ny);
    h_off = int64_t(gy) * p.nx + gx;
  }

  double wx[W], wy[W], wz[W];
#pragma unroll
  for (int s = 0; s < W; ++s) { wx[s] = p.w[0][s]; wy[s] = p.w[1][s]; wz[s] = p.w[2][s]; }

  // z window in registers; the plane that enters it is loaded one iteration before its first use (the
  // centred dz stencil needs plane z + HW in iteration z: loaded in that iteration, the first DFMA of the
  // dz sum waited on HBM — ncu: long_scoreboard on that DFMA was 10 % of all samples)
  double2 win[W];
#pragma unroll
  for (int s = 0; s < W - 1; ++s) win[s] = ldz(z0 - HW + s);
  double2 ahead = ldz(z0 + HW);
  double2 hreg = (h_ok) ? *reinterpret_cast<const double2*>(p.U + z0 * plane + h_off) : zero2;

  for (int z = z0; z < z1; ++z) {
    win[W - 1] = ahead;
    ahead = (z + 1 < z1) ? ldz(z + 1 + HW) : zero2;
    double* buf = sm[z & 1];
    *reinterpret_cast<double2*>(buf + (ty + HW) * PITCH + HX + 2 * tx) = win[HW];
    if (h_sm >= 0) {
      *reinterpret_cast<double2*>(buf + h_sm) = hreg;
      hreg = (h_ok && z + 1 < z1) ? *reinterpret_cast<const double2*>(p.U + (z + 1) * plane + h_off) : zero2;
    }
    __syncthreads();
    if (mine) {
      const double* rowp = buf + (ty + HW) * PITCH + 2 * tx;   // rowp[0] = point x0 - HX
      double v[10];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        // (the pair in the middle is this thread's own: it is in the z window already)
        const double2 t2 = (2 * k == HX) ? win[HW] : *reinterpret_cast<const double2*>(rowp + 2 * k);
        v[2 * k] = t2.x; v[2 * k + 1] = t2.y;
      }
      double2 gx2 = zero2, gy2 = zero2, gz2 = zero2;
#pragma unroll
      for (int s = 0; s < W; ++s) {
        gx2.x = fma(wx[s], v[HX - HW + s], gx2.x);
        gx2.y = fma(wx[s], v[HX - HW + 1 + s], gx2.y);
        const double2 c = (s == HW) ? win[HW]
                                    : *reinterpret_cast<const double2*>(buf + (ty + s) * PITCH + HX + 2 * tx);
        gy2.x = fma(wy[s], c.x, gy2.x);
        gy2.y = fma(wy[s], c.y, gy2.y);
        gz2.x = fma(wz[s], win[s].x, gz2.x);
        gz2.y = fma(wz[s], win[s].y, gz2.y);
      }
      const int64_t o = z * plane + int64_t(y) * p.nx + x0;
      // streaming stores: the outputs are never re-read, keep L2 for the halo re-reads of U
      __stcs(reinterpret_cast<double2*>(p.G[0] + o), gx2);
      __stcs(reinterpret_cast<double2*>(p.G[1] + o), gy2);
      __stcs(reinterpret_cast<double2*>(p.G[2] + o), gz2);
    }
#pragma unroll
    for (int s = 0; s < W - 1; ++s) win[s] = win[s + 1];
  }
}

// tile shape, tuning hook pt_debug_set_grad_ty(ty) / (ty + 100 for 64 threads along x).  Measured at
// 1024^3 on B200: 128 x 8 points (1 KB rows = one DRAM page per tile row) 6.55 ms, 128 x 16 6.83 ms,
// 64 x 16 7.20 ms, 64 x 32 7.49 ms
int g_grad_ty = 8;
int g_grad_tx = 64;  // threads along x (the tile is twice as wide)

inline int grad_tile_x() { return 2 * g_grad_tx; }   // (9-point stencils use 64; only the z-chunk heuristic sees the difference)

template <int W>
int launch_grad3(const Grad3Params& p, cudaStream_t st) {
  const int tx = (W > 7) ? 32 : g_grad_tx;        // the 128-wide tile of a 9-point stencil exceeds 48 KB
  const int ty = (W > 7 && g_grad_ty == 8) ? 16 : g_grad_ty;
  const int gx = (p.nx + 2 * tx - 1) / (2 * tx), gy = (p.ny + ty - 1) / ty;
  const int gz = (p.nz + p.zchunk - 1) / p.zchunk;
  if (gy > 65535 || gz > 65535) return pt::fail(PT_ERR_UNSUPPORTED, "pt_stencil_grad3: grid too large");
  dim3 grid(gx, gy, gz), block(tx, ty);
  if constexpr (W <= 7) {
    if (tx == 64) {
      if (ty == 8) stencil_grad3_kernel<W, 8, 64><<<grid, block, 0, st>>>(p);
      else stencil_grad3_kernel<W, 16, 64><<<grid, block, 0, st>>>(p);
      return pt::check_launch("pt_stencil_grad3");
    }
  }
  if (ty == 32) stencil_grad3_kernel<W, 32, 32><<<grid, block, 0, st>>>(p);
  else stencil_grad3_kernel<W, 16, 32><<<grid, block, 0, st>>>(p);
  return pt::check_launch("pt_stencil_grad3");
}

}  // namespace

extern "C" int pt_debug_set_grad_ty(int ty) {
  if (ty >= 100) { g_grad_tx = 64; g_grad_ty = (ty - 100 == 16) ? 16 : 8; }
  else { g_grad_tx = 32; g_grad_ty = (ty == 16) ? 16 : 32; }
  return PT_OK;
}

namespace {
int grad3_impl(const double* U, const double* U_lo, const double* U_hi, double* GX, double* GY, double* GZ,
               int nx, int ny, int nz, int64_t nfields, const double* const* W, const int32_t* const* start,
               const int* wmax, const pt_uniform_rows* const* uni, pt_stream_t stream) {
  PT_REQUIRE(U && GX && GY && GZ && W && start && wmax && uni, "pt_stencil_grad3: null pointer");
  PT_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1 && nfields >= 0, "pt_stencil_grad3: bad shape");
  PT_REQUIRE(U != GX && U != GY && U != GZ, "pt_stencil_grad3: in-place application is not supported");
  const int dims[3] = {nx, ny, nz};
  int w = 0;
  for (int k = 0; k < 3; ++k) {
    if (!uni[k] || !uni[k]->wuni) return pt::fail(PT_ERR_UNSUPPORTED, "pt_stencil_grad3: axis %d is not uniform", k);
    if (k == 0) w = uni[k]->w_uni;
    if (uni[k]->w_uni != w) return pt::fail(PT_ERR_UNSUPPORTED, "pt_stencil_grad3: stencil widths differ");
    const int hw = w / 2;
    // canonical rows: everything but the hw rows next to a physical boundary (a slab side with a
    // neighbour has no boundary rows)
    const int lo_k = (k == 2 && U_lo) ? 0 : hw, hi_k = (k == 2 && U_hi) ? dims[k] : dims[k] - hw;
    if (uni[k]->lo != lo_k || uni[k]->hi != hi_k || dims[k] < 2 * w)
      return pt::fail(PT_ERR_UNSUPPORTED, "pt_stencil_grad3: axis %d is not a plain non-periodic uniform axis", k);
    PT_REQUIRE(W[k] && start[k] && wmax[k] >= 1 && wmax[k] <= MAXW, "pt_stencil_grad3: bad band for axis %d", k);
  }
  if (!(w == 3 || w == 5 || w == 7 || w == 9)) return pt::fail(PT_ERR_UNSUPPORTED, "pt_stencil_grad3: width %d", w);
  if (nx % 2 || !aligned16(U) || !aligned16(GX) || !aligned16(GY) || !aligned16(GZ))
    return pt::fail(PT_ERR_UNSUPPORTED, "pt_stencil_grad3: needs even nx and 16-byte aligned fields");
  cudaStream_t st = pt::as_stream(stream);
  const int64_t P = int64_t(nx) * ny * nz;
  double* G[3] = {GX, GY, GZ};
  for (int64_t f = 0; f < nfields; ++f) {
    Grad3Params p;
    p.U = U + f * P;
    p.Ulo = U_lo; p.Uhi = U_hi;
    for (int k = 0; k < 3; ++k) {
      p.G[k] = G[k] + f * P;
      for (int s = 0; s < MAXW; ++s) p.w[k][s] = s < w ? uni[k]->wuni[s] : 0.0;
    }
    p.nx = nx; p.ny = ny; p.nz = nz;
    // z-chunks: enough CTAs for a few waves, halo re-read (w-1 planes per chunk) kept small
    const int64_t tiles = int64_t((nx + grad_tile_x() - 1) / grad_tile_x()) * ((ny + g_grad_ty - 1) / g_grad_ty);
    int chunks = int((int64_t(pt::sm_count()) * 8 + tiles - 1) / tiles);
    if (chunks < 1) chunks = 1;
    int zchunk = (nz + chunks - 1) / chunks;
    if (zchunk < 8 * w) zchunk = 8 * w;
    if (zchunk > nz) zchunk = nz;
    p.zchunk = zchunk;
    int rc;
    switch (w) {
      case 3: rc = launch_grad3<3>(p, st); break;
      case 5: rc = launch_grad3<5>(p, st); break;
      case 7: rc = launch_grad3<7>(p, st); break;
      default: rc = launch_grad3<9>(p, st); break;
    }
    if (rc) return rc;
    // one-sided boundary planes of each axis through the table
    for (int k = 0; k < 3; ++k) {
      StencilParams q;
      q.W = W[k]; q.start = start[k]; q.U = p.U; q.Y = p.G[k];
      q.wmax = wmax[k]; q.n = dims[k]; q.n_out = dims[k]; q.shift = 0;
      q.before = (k == 0) ? 1 : (k == 1 ? nx : int64_t(nx) * ny);
      q.after = P / (q.before * dims[k]);
      q.periodic = 0; q.alpha = 1.0; q.beta = 0.0;
      q.lo = uni[k]->lo; q.hi = uni[k]->hi;
      q.Ulo = nullptr; q.Uhi = nullptr; q.halo_lo = 0; q.halo_hi = 0; q.lo_stride = 0; q.hi_stride = 0;
      if (k == 2 && (U_lo || U_hi)) {
        // start[] of a slab is relative to the virtual slab (halo rows first): the table reads
        // them from wherever they live
        const int hw = w / 2;
        q.halo_lo = U_lo ? hw : 0; q.halo_hi = U_hi ? hw : 0;
        q.Ulo = U_lo ? U_lo : p.U; q.Uhi = U_hi ? U_hi : p.U;
        q.n = q.halo_lo + nz + q.halo_hi; q.shift = q.halo_lo;
      }
      rc = launch_table(q, q.lo, q.hi, st);
      if (rc) return rc;
    }
  }
  return PT_OK;
}
}  // namespace

extern "C" int pt_stencil_grad3(const double* U, double* GX, double* GY, double* GZ, int nx, int ny, int nz,
                                int64_t nfields, const double* const* W, const int32_t* const* start,
                                const int* wmax, const pt_uniform_rows* const* uni, pt_stream_t stream) {
  return grad3_impl(U, nullptr, nullptr, GX, GY, GZ, nx, ny, nz, nfields, W, start, wmax, uni, stream);
}

extern "C" int pt_stencil_grad3_halo(const double* U, const double* U_lo, const double* U_hi, double* GX,
                                     double* GY, double* GZ, int nx, int ny, int nz_own, const double* const* W,
                                     const int32_t* const* start, const int* wmax,
                                     const pt_uniform_rows* const* uni, pt_stream_t stream) {
  return grad3_impl(U, U_lo, U_hi, GX, GY, GZ, nx, ny, nz_own, 1, W, start, wmax, uni, stream);
}
