/*
 * pytristan_b200 — C-ABI of the B200 (sm_100a) differentiation hot path.
 *
 * Drop-in boundary.  The reference (pytristan 0.2a) is pure Python/NumPy and
 * has no FFI; every entry point below replaces a NumPy call site of the reference or the
 * operator layer its README announces (README.md:7-10,16,24-26).  The Python façade
 * (the modules under pytristan_b200/) binds these with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative pt_status on failure; the message of the
 *     last failure on the calling thread is returned by pt_last_error();
 *   - all data pointers are DEVICE pointers unless the name ends in _host; buffers are owned by
 *     the caller; nothing is allocated inside;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *     enqueued asynchronously on it;
 *   - a *field* on a grid with npts = (n0, .., n_{d-1}) is a flat FP64 vector in the reference's
 *     own linearisation, axis 0 fastest (reference: grid.py:85-88).  For the axis k being
 *     differentiated, before = prod(npts[:k]) (the stride of axis k), after = prod(npts[k+1:])
 *     times the batch.  Element (a, i, b) sits at (a*n + i)*before + b.
 */
#ifndef PYTRISTAN_B200_H
#define PYTRISTAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pt_stream_t;

enum pt_status {
  PT_OK = 0,
  PT_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, unsupported value) */
  PT_ERR_CUDA = -2,      /* a CUDA runtime call or launch failed */
  PT_ERR_UNSUPPORTED = -3 /* valid request the library has no kernel for */
};

/* scale vector addressing for pt_dense_apply's fused epilogue */
enum pt_scale_mode {
  PT_SCALE_NONE = 0,
  PT_SCALE_ROW = 1,    /* scale[i],            i = output index along the differentiated axis */
  PT_SCALE_AFTER = 2,  /* scale[a % scale_len], a = index over the slower axes (and batch)     */
  PT_SCALE_BEFORE = 3  /* scale[b],            b = index over the faster axes                  */
};

/* ---- library ---------------------------------------------------------------------------- */
const char* pt_last_error(void);
int pt_version(void);
/* SM count and compute capability of the current device. */
int pt_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- K5: node and grid generators (replace grid.py:75-77, :205, :83-88) ------------------- */
/* x[i] = fl(fl(i*step)+start), step = fl(fl(stop-start)/(n-1)), x[n-1] = stop: NumPy's
 * linspace arithmetic (numpy/_core/function_base.py), used by the reference at grid.py:205. */
int pt_gen_linspace(double* x, double start, double stop, int64_t n, pt_stream_t stream);
/* Chebyshev-Gauss-Lobatto nodes x_j = cos(fl(fl(j*pi)/(n-1))), descending (grid.py:75);
 * if mapped != 0 additionally fl(fl(fl(fl(1-x)/2)*fl(xmax-xmin))+xmin) (grid.py:77). */
int pt_gen_cheb(double* x, int64_t n, int mapped, double xmin, double xmax, pt_stream_t stream);
/* The same nodes with the HOST libm's cosine reproduced bit for bit on the device, so that device
 * generated nodes equal numpy.cos-generated ones (grid.py:75) at every node, not just within an
 * ulp: glibc's dbl-64 cos is not correctly rounded, and its x86-64 FMA and baseline builds differ
 * from each other.  libm selects the build the host runs (the Python side probes numpy.cos). */
enum pt_libm_model {
  PT_LIBM_GLIBC_FMA = 1,   /* glibc >= 2.28 dbl-64 s_sin.c, -mfma -mavx2 build (__cos_fma) */
  PT_LIBM_GLIBC_NOFMA = 2  /* same source, baseline build (__cos_sse2 / __cos_avx)          */
};
int pt_gen_cheb_libm(double* x, int64_t n, int mapped, double xmin, double xmax, int libm,
                     pt_stream_t stream);
/* y[i] = cos(t[i]) with that cosine (|t| < 105414350; beyond that the CUDA cosine). */
int pt_cos_libm(const double* t, double* y, int64_t n, int libm, pt_stream_t stream);
/* In-place affine map of reference nodes, the arithmetic of grid.py:77 bit for bit. */
int pt_cheb_map(double* x, int64_t n, double xmin, double xmax, pt_stream_t stream);
/* Expand ndim 1-D axis arrays into the reference Grid layout (grid.py:83-88): out is
 * (ndim, P) row-major, P = prod(npts), out[d][l] = axis_d[(l / prod(npts[:d])) % npts[d]].
 * `axes` holds the axis arrays back to back (sum(npts) elements of elem_size bytes);
 * npts_host is a HOST array of ndim entries.  elem_size in {1,2,4,8,16}: pure data movement. */
int pt_grid_expand(const void* axes, const int64_t* npts_host, int ndim, int elem_size,
                   void* out, pt_stream_t stream);
/* Strided gather of the 1-D nodes of one axis out of a Grid row (grid.py:277-280). */
int pt_grid_axpoints(const double* grid_row, int64_t stride, int64_t n, double* x,
                     pt_stream_t stream);

/* ---- K6: operator generators (operator layer announced at README.md:7-10,16) -------------- */
/* Chebyshev collocation differentiation matrix of the given order on the nodes x (any affine
 * image of the CGL set, either orientation).  D is n*n row-major.  Recursion of Welfert with
 * the negative-sum diagonal, sequential summation order (oracle/operators.py:cheb_mat). */
int pt_gen_chebmat(const double* x, int n, int order, double* D, pt_stream_t stream);
/* Flipping trick: overwrite the rows below the middle (and the right part of the middle row of an
 * odd n) by the mirror image of the rows above, D[i][j] = sign * D[n-1-i][n-1-j] (sign = +1 / -1
 * for even / odd derivative orders).  ChebMat applies it on CGL axes: the matrix is then exactly
 * centro(anti)symmetric, which the even-odd kernels below exploit (oracle/operators.py:cheb_mat). */
int pt_flip_symmetrize(double* D, int n, int sign, pt_stream_t stream);
/* Polynomial collocation differentiation matrix on ARBITRARY distinct nodes (a non-affine
 * mapper of the reference, grid.py:201-207, returns nodes only): barycentric weights
 * w_j = 1/prod(x_j - x_k) held as mantissa*2^exponent, D1_ij = (w_j/w_i)/(x_i-x_j), higher
 * orders by the same recursion as pt_gen_chebmat.  work: 2*n doubles.
 * (oracle/operators.py:bary_mat). */
int pt_gen_barymat(const double* x, int n, int order, double* D, double* work, pt_stream_t stream);
/* Fourier spectral differentiation matrix (order 1 or 2) for n equispaced periodic nodes on a
 * period `period`.  D is n*n row-major, circulant (oracle/operators.py:four_mat). */
int pt_gen_fourmat(int n, double period, int order, double* D, pt_stream_t stream);
/* Fourier differentiation matrix of any order 1..8 by direct summation of the symbol (i k s)^p
 * over the retained modes (Nyquist mode dropped for odd orders).  work: n doubles.
 * (oracle/operators.py:four_col_general). */
int pt_gen_fourmat_general(int n, double period, int order, double* D, double* work,
                           pt_stream_t stream);
/* Fornberg finite-difference weights, one row per node, in banded row-start layout:
 * start[i] = clip(i - floor(w/2), 0, n - wmax) (periodic: i - floor(w/2), wrapped by the
 * kernels), W[i*wmax + s] multiplies x[start[i] + s].  w = 2*floor((order+1)/2) - 1 + accuracy
 * nodes for interior rows, order+accuracy nodes for one-sided boundary rows,
 * wmax = max of the two (periodic: wmax = w).  If uniform != 0 the weights are generated on
 * integer offsets and scaled by h^-order (h given), so that all interior rows are identical.
 * (oracle/operators.py:findiff_band). */
int pt_gen_fornberg(const double* x, int n, int order, int accuracy, int periodic, int uniform,
                    double h, int32_t* start, double* W, int wmax, pt_stream_t stream);
/* Banded row-start form -> dense n*n row-major matrix (zeros off the band). */
int pt_band_to_dense(const double* W, const int32_t* start, int wmax, int n, int periodic,
                     double* D, pt_stream_t stream);
/* Explicit Kronecker lifting kron(I_after, kron(D, I_before)) for small grids; L is P*P
 * row-major with P = after*n*before. */
int pt_kron_lift(const double* D, int n, int64_t before, int64_t after, double* L,
                 pt_stream_t stream);

/* ---- K1/K2: dense operator application (FP64 tensor cores, DMMA m8n8k4) ------------------- */
/* Number of doubles of the packed (fragment-major) copy of an n_out*n_in operator. */
int64_t pt_packed_size(int n_out, int n_in);
/* Pack a row-major operator (leading dimension ldD) into the layout the apply kernels read:
 * [ceil(n_in/4)][8*ceil(n_out/8)][4], zero padded. */
int pt_operator_pack(const double* D, int n_out, int n_in, int64_t ldD, double* packed,
                     pt_stream_t stream);
/* Y[a,:,b] = alpha * scale * (D @ U[a,:,b]) + beta * Y[a,:,b] for all a < after, b < before.
 * U is (after, n_in, before), Y is (after, n_out, before), both C-contiguous.  before == 1
 * (axis 0 of a grid) selects the contiguous-axis kernel (K2), before > 1 the strided-panel
 * kernel (K1); neither transposes U or Y.  scale may be NULL (mode PT_SCALE_NONE). */
int pt_dense_apply(const double* packed, int n_out, int n_in, const double* U, double* Y,
                   int64_t after, int64_t before, double alpha, double beta,
                   const double* scale, int scale_mode, int64_t scale_len, pt_stream_t stream);

/* ---- K1/K2 for centrosymmetric operators: even-odd decomposition (half the tensor-core work) ---- */
/* Largest entrywise violation of J D J = sign * D (J = order reversal), relative to the entry:
 * 0 for an exactly centro(anti)symmetric matrix.  Written to the caller-owned DEVICE scalar
 * *worst_dev in stream order (nothing is allocated, nothing synchronises). */
int pt_operator_symmetry(const double* D, int n, int sign, double* worst_dev, pt_stream_t stream);
/* Doubles of each of the two packed half operators. */
int64_t pt_packed_eo_size(int n);
/* Ae[i][k] = (D[i][k] + D[i][n-1-k])/2 and Ao[i][k] = (D[i][k] - D[i][n-1-k])/2 for i, k < ceil(n/2)
 * in the fragment-major layout of the apply kernels (the middle column of an odd n holds D/2, 0). */
int pt_operator_pack_eo(const double* D, int n, double* packed_even, double* packed_odd, pt_stream_t stream);
/* Same contract as pt_dense_apply for an operator with J D J = sign * D: rows i < ceil(n/2) are
 * computed as Ae (u + Ju) + Ao (u - Ju) (exact algebra), rows n-1-i follow from the symmetry.
 * Two half-size products instead of one full-size product; E = u + Ju, O = u - Ju are formed in
 * the kernel from two staged field tiles, no extra pass over the field. */
int pt_dense_apply_eo(const double* packed_even, const double* packed_odd, int n, int sign,
                      const double* U, double* Y, int64_t after, int64_t before, double alpha,
                      double beta, const double* scale, int scale_mode, int64_t scale_len,
                      pt_stream_t stream);

/* The same operation by the warp-specialised kernel (csrc/dense_apply_eo_ws.cu): producer warps stage
 * the operator panels with the bulk-copy engine (cp.async.bulk, one 32 KB request per k chunk) and
 * form E / O once per 128 half-rows, consumer warps run LDS + DMMA only, mbarrier ring, no CTA
 * barrier.  Bit-identical to pt_dense_apply_eo for even n; for odd n the middle row / column are
 * rank-1 terms beside an exactly (n-1)/2-sized half problem.  The operator is packed once per
 * (128-row block, k chunk): pt_packed_eo_ws_size(n) doubles filled by pt_operator_pack_eo_ws.
 * Returns PT_ERR_UNSUPPORTED — nothing done, no message — for shapes it does not cover (n < 128,
 * fewer tiles than SMs, odd strides or odd n on the contiguous axis, unaligned fields): callers
 * then use pt_dense_apply_eo. */
int64_t pt_packed_eo_ws_size(int n);
int pt_operator_pack_eo_ws(const double* D, int n, double* packed_ws, pt_stream_t stream);
int pt_dense_apply_eo_ws(const double* packed_ws, int n, int sign, const double* U, double* Y,
                         int64_t after, int64_t before, double alpha, double beta, const double* scale,
                         int scale_mode, int64_t scale_len, pt_stream_t stream);

/* ---- FourMat through the FFT ----------------------------------------- */
/* Twiddle table of pt_fourier_apply: n complex doubles (2 n doubles), n a power of two.  Opaque to the
 * caller; the layout is one run of roots of unity per radix-4 level L = n, n/4, n/16, ...:
 * tw[n - L + i] = exp(-2 pi i * i / L) for i < 3L/4, so that neighbouring lanes read neighbouring entries
 * at every level.  Build it once per n with this call and keep it with the operator. */
int pt_fourier_twiddles(int n, double* tw, pt_stream_t stream);
/* Y[a,:,b] = alpha * scale[a % scale_len] * (D U[a,:,b]) + beta * Y[a,:,b] (scale may be NULL:
 * no scaling; it is the PT_SCALE_AFTER epilogue of pt_dense_apply) with D the order-`order` Fourier spectral
 * differentiation matrix on n equispaced periodic nodes of period `period` — the same linear map
 * as the dense matrix of pt_gen_fourmat / pt_gen_fourmat_general, evaluated by shared-memory FFTs
 * (O(n log n), HBM-bound: 16 B/point).  Returns PT_ERR_UNSUPPORTED (and does nothing) unless n is
 * a power of two in [4, 4096] and, on a strided axis (before > 1), before is even and the fields
 * are 16-byte aligned: callers then use pt_dense_apply. */
int pt_fourier_apply(const double* tw, int n, double period, int order, const double* U, double* Y,
                     int64_t after, int64_t before, double alpha, double beta, const double* scale,
                     int64_t scale_len, pt_stream_t stream);

/* ---- K3/K4: banded (finite-difference) application --------------------------------------- */
/* Optional description of the uniform part of a band (all pointers are HOST pointers).  Rows
 * lo <= i < hi are canonical interior rows: their weights are the w_uni (3, 5, 7 or 9) entries of
 * wuni applied at offset i - floor(w_uni/2); the kernels keep them in registers.  The remaining
 * n_bnd = lo + (n_out - hi) boundary rows (top rows first, then bottom rows) may be described by
 * w_bnd (n_bnd * wmax weights) and start_bnd (n_bnd first columns): with n_bnd <= 8 the
 * contiguous-axis kernel then computes them out of the same shared-memory tile (single launch). */
typedef struct pt_uniform_rows {
  int w_uni, lo, hi, n_bnd;
  const double* wuni;
  const double* w_bnd;
  const int32_t* start_bnd;
} pt_uniform_rows;

/* Y[a,i,b] = alpha * sum_s W[i*wmax+s] * U[a, start[i]+s, b] + beta * Y[a,i,b].
 * uni may be NULL (every row goes through the per-row table: mapped, non-uniform axes).
 * periodic != 0 wraps indices modulo n. */
int pt_stencil_apply(const double* W, const int32_t* start, int wmax, int n, const double* U,
                     double* Y, int64_t after, int64_t before, int periodic,
                     const pt_uniform_rows* uni, double alpha, double beta, pt_stream_t stream);

/* Rectangular variant for slab-sharded axes (multi-GPU): U holds n_in rows per slab (the local
 * rows plus halo planes), Y holds n_out rows; output row i corresponds to input row i + shift,
 * start[i] (one entry per OUTPUT row) is relative to the input slab, and canonical rows
 * lo <= i < hi read input rows i + shift - floor(w_uni/2) ... i + shift + floor(w_uni/2). */
int pt_stencil_apply_ex(const double* W, const int32_t* start, int wmax, int n_in, int n_out,
                        int shift, const double* U, double* Y, int64_t after, int64_t before,
                        int periodic, const pt_uniform_rows* uni, double alpha, double beta,
                        pt_stream_t stream);

/* Fused three-axis gradient of 3-D fields (nx fastest): GX, GY, GZ = d/dx, d/dy, d/dz of U in one
 * read of U (32 B/pt instead of 3 x 16 B/pt).  W/start/wmax/uni are arrays of 3 (axis 0, 1, 2)
 * with the same meaning as in pt_stencil_apply.  Returns PT_ERR_UNSUPPORTED (and does nothing)
 * unless every axis is a non-periodic uniform axis with the same stencil width, nx is even and
 * the fields are 16-byte aligned: callers then fall back to three pt_stencil_apply calls. */
int pt_stencil_grad3(const double* U, double* GX, double* GY, double* GZ, int nx, int ny, int nz,
                     int64_t nfields, const double* const* W, const int32_t* const* start,
                     const int* wmax, const pt_uniform_rows* const* uni, pt_stream_t stream);

/* The same fused gradient on a z-slab of a slab-sharded field (multi-GPU): the floor(w/2) planes
 * below / above the nz_own local planes are loaded from the neighbours' memory (U_lo points at
 * the first of them in the lower neighbour, U_hi at the upper neighbour's first plane; NULL = no
 * neighbour, i.e. a physical boundary with one-sided rows).  The z band (W[2], start[2], uni[2])
 * describes the LOCAL rows, start relative to the virtual slab as in pt_stencil_apply_halo. */
int pt_stencil_grad3_halo(const double* U, const double* U_lo, const double* U_hi, double* GX,
                          double* GY, double* GZ, int nx, int ny, int nz_own, const double* const* W,
                          const int32_t* const* start, const int* wmax,
                          const pt_uniform_rows* const* uni, pt_stream_t stream);

/* ---- K7: pencil-transpose pack/unpack (multi-GPU) ----------------------------------------- */
/* Generic 3-D permutation copy used on both sides of the all-to-all: in is (d2, d1, d0)
 * C-contiguous; out[perm] = in, perm one of the 6 axis orders encoded as p0 + 3*p1 + 9*p2
 * where output axis j (slowest first) is input axis p_j. */
int pt_permute3d(const double* in, double* out, int64_t d2, int64_t d1, int64_t d0, int perm,
                 pt_stream_t stream);

/* ---- multi-GPU over peer memory (one process per GPU, one NVSwitch box) --------- */
#define PT_IPC_HANDLE_BYTES 64
/* Exchange buffer: plain device allocation (IPC-exportable), zero filled. */
int pt_peer_alloc(int64_t bytes, void** ptr);
int pt_peer_free(void* ptr);
/* CUDA IPC handle of an exchange buffer (PT_IPC_HANDLE_BYTES bytes, moved between the ranks by
 * the host, e.g. torch.distributed.all_gather_object) and its mapping in another process. */
int pt_ipc_export(const void* ptr, void* handle64);
int pt_ipc_open(const void* handle64, void** ptr);
int pt_ipc_close(void* ptr);
/* Device-side barrier among nranks processes.  flags is a DEVICE array of nranks pointers,
 * flags[r] = rank r's flag array (nranks uint64 slots, zero initialised, peer mapped).  The kernel
 * stores `epoch` (strictly increasing, >= 1) into slot `rank` of every peer and waits until
 * every slot of its own array reached it: work enqueued before the barrier on any rank is
 * visible to work enqueued after it on every rank.  timeout_ns > 0 bounds the wait; on expiry
 * *status (device int, may be NULL) is set to 1 + the missing rank and the kernel returns. */
int pt_peer_barrier(uint64_t* const* flags, int rank, int nranks, uint64_t epoch,
                    int64_t timeout_ns, int* status, pt_stream_t stream);

/* Where the elements of a locally held (after, n, before) array go in a pencil redistribution:
 * element (a, i, b) -> base[i / chunk] + offset + (a / a_inner)*s_outer + (a % a_inner)*s_inner
 *                      + (i % chunk)*s_i + b      (all strides in doubles).
 * base is a DEVICE array of nranks peer-mapped buffers. */
typedef struct pt_scatter_map {
  int nranks;
  int chunk;
  int64_t a_inner, s_outer, s_inner, s_i, offset;
  double* const* base;
} pt_scatter_map;

/* pt_dense_apply whose epilogue is the redistribution: alpha * (D @ U[a,:,b]) + beta * C[a,:,b]
 * is stored through `map` straight into the buffers of the ranks that own it afterwards (K7 + C1
 * fused: the NVLink transfer proceeds tile by tile under the DMMA main loop).  C (local, laid
 * out like the un-scattered result (after, n_out, before)) may be NULL when beta == 0. */
int pt_dense_apply_scatter(const double* packed, int n_out, int n_in, const double* U,
                           const double* C, int64_t after, int64_t before, double alpha,
                           double beta, const pt_scatter_map* map, pt_stream_t stream);
/* pt_dense_apply_scatter for a centrosymmetric operator (pt_dense_apply_eo). */
int pt_dense_apply_eo_scatter(const double* packed_even, const double* packed_odd, int n, int sign,
                              const double* U, const double* C, int64_t after, int64_t before,
                              double alpha, double beta, const pt_scatter_map* map, pt_stream_t stream);
/* ... and by the warp-specialised kernel (pt_dense_apply_eo_ws; PT_ERR_UNSUPPORTED likewise). */
int pt_dense_apply_eo_ws_scatter(const double* packed_ws, int n, int sign, const double* U, const double* C,
                                 int64_t after, int64_t before, double alpha, double beta,
                                 const pt_scatter_map* map, pt_stream_t stream);
/* ... and the redistribution of the FIELD fused into the same kernel: besides alpha * (D U) + beta * C through
 * `map`, U itself is stored through `field_map` by the warps that load it (every element of U passes through
 * their registers exactly once), so that a pencil transposition of (field, running sum) is ONE kernel whose NVLink
 * traffic is spread under its main loop — no pt_peer_scatter pass.  PT_ERR_UNSUPPORTED as pt_dense_apply_eo_ws,
 * and for field maps with odd strides / chunk. */
int pt_dense_apply_eo_ws_scatter_copy(const double* packed_ws, int n, int sign, const double* U, const double* C,
                                      int64_t after, int64_t before, double alpha, double beta,
                                      const pt_scatter_map* map, const pt_scatter_map* field_map, pt_stream_t stream);
/* The same redistribution for an array that is only moved (the field itself): 16-byte peer
 * stores, no staging buffer, no pack/unpack pass.  before == 1 requires s_i == 1. */
int pt_peer_scatter(const double* src, int n, int64_t after, int64_t before,
                    const pt_scatter_map* map, pt_stream_t stream);

/* pt_stencil_apply along a slab-sharded axis WITHOUT a halo exchange: the halo_lo rows below and
 * the halo_hi rows above the n_own local rows are loaded from the neighbours' memory (U_lo points
 * at the first needed row of the lower neighbour, U_hi at the first row of the upper one;
 * lo_stride / hi_stride = doubles between consecutive slabs `a` there).  start[] (one entry per
 * output row) is relative to the virtual slab halo_lo + n_own + halo_hi, as in
 * pt_stencil_apply_ex with shift = halo_lo.  before must be > 1. */
int pt_stencil_apply_halo(const double* W, const int32_t* start, int wmax, int n_own, int halo_lo,
                          int halo_hi, const double* U, const double* U_lo, int64_t lo_stride,
                          const double* U_hi, int64_t hi_stride, double* Y, int64_t after,
                          int64_t before, const pt_uniform_rows* uni, double alpha, double beta,
                          pt_stream_t stream);

/* ---- boundary rows and solves along an axis ---------------------------- */
/* Rows rows[r] (r < nrows) of the n*n row-major device matrix D are overwritten by R[r] (nrows*n):
 * boundary-condition row replacement (Dirichlet: identity rows; Neumann/Robin: rows of D1, ...). */
int pt_replace_rows(double* D, int n, const int32_t* rows, int nrows, const double* R,
                    pt_stream_t stream);
/* Ainv = A^-1 (n*n row-major, n <= 4096) by Gauss-Jordan elimination with partial pivoting in FP64.
 * work: pt_invert_work_size(n) doubles.  *info (device int32) is 0 on success, c+1 if the matrix is
 * exactly singular at elimination step c.  A solve A y = f along a grid axis, for all grid lines
 * at once, is then pt_dense_apply with the packed inverse. */
int64_t pt_invert_work_size(int n);
int pt_invert(const double* A, int n, double* Ainv, double* work, int32_t* info, pt_stream_t stream);

/* pt_dense_apply (no scaling) on ROW RANGES along the axis of longer fields: ldu / ldy are the distances in
 * doubles between consecutive `a` slabs of U and Y (n_in*before / n_out*before for whole fields); U and Y point
 * at the first row of the range.  The building block of the block substitutions below. */
int pt_dense_apply_ld(const double* packed, int n_out, int n_in, const double* U, int64_t ldu, double* Y,
                      int64_t ldy, int64_t after, int64_t before, double alpha, double beta, pt_stream_t stream);

/* Factorisation-based solve along an axis: P A = L U, then blocked substitution.
 * pt_lu_factor: LU (n*n row-major, may alias A) receives the multipliers of the unit lower factor below the
 *   diagonal and U on and above it; perm[i] = row of A that is row i of P A; partial pivoting with the first
 *   row of maximal modulus (LAPACK's rule); work: pt_lu_work_size(n) doubles; *info (device int32) = c+1 if
 *   column c has no nonzero pivot, else 0.  n <= 8192.
 * pt_lu_plan: for block rows of nb rows (multiple of 16, <= 128) packs the off-diagonal block rows L[b, 0:r0] and
 *   U[b, r0+m:n] as pt_operator_pack operators and the diagonal triangles for the substitution kernel into plan
 *   (pt_lu_plan_size doubles).
 * pt_lu_solve: X[a, :, b] = A^-1 F[a, :, b] for every grid line: one row-permuting copy, then per block row one
 *   rectangular pt_dense_apply_ld (tensor cores: all of the O(n^2) work per line) and one triangular
 *   substitution with the nb*nb diagonal block (FP64 FMA, one thread per grid line — never a product with an
 *   inverted block: componentwise backward stable) in the forward and in the backward sweep.  work is a second
 *   field-sized buffer; F, X and work are distinct.  Nothing is allocated, nothing synchronises. */
int64_t pt_lu_work_size(int n);
int pt_lu_factor(const double* A, int n, double* LU, int32_t* perm, double* work, int32_t* info, pt_stream_t stream);
int64_t pt_lu_plan_size(int n, int nb);
int pt_lu_plan(const double* LU, int n, int nb, double* plan, pt_stream_t stream);
int pt_lu_solve(const double* plan, const int32_t* perm, int n, int nb, const double* F, double* X, double* work,
                int64_t after, int64_t before, pt_stream_t stream);
/* tuning hook: grid lines per CTA of the substitution kernel (64 or 128; 0 = default) */
int pt_debug_set_tri_lines(int lines);

/* ---- measurement helper ------------------------------------------------------------------ */
/* Register-resident DMMA loop: the FP64 tensor-core roofline denominator measured on the box.
 * Writes achieved TFLOP/s to *tflops_host. */
int pt_measure_dmma_peak(double* tflops_host, int iters);

#ifdef __cplusplus
}
#endif
#endif /* PYTRISTAN_B200_H */
