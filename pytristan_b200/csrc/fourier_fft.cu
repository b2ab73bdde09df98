// FourMat application through the FFT: Y[a,:,b] = alpha * D_p U[a,:,b] + beta*Y
// where D_p is the Fourier spectral differentiation matrix of order p on n equispaced periodic
// nodes — the same linear map as the dense FourMat (the symbol (i k s)^p on the modes |k| < n/2,
// the Nyquist mode kept for even p and dropped for odd p), evaluated in O(n log n) and bound by
// HBM (16 B/point) instead of by the FP64 tensor pipe.  n must be a power of two in [4, 4096].
//
// The operator is real, so two real sequences ride in one complex transform: z = u1 + i*u2,
// D z = D u1 + i*D u2.  On the contiguous axis (before == 1) the pair is two consecutive rows;
// on a strided axis (before even) it is two adjacent columns, which are one 16-byte load.
//
// Per complex sequence: forward decimation-in-frequency stages (natural order in, digit-reversed
// out), multiplication by the symbol at the digit-reversed position, inverse decimation-in-time
// stages (digit-reversed in, natural out) — no reordering pass.  Every stage is a radix-16
// butterfly held in registers (two radix-4 levels); the first one loads straight from global
// memory, the last one stores straight to it, and the last forward level, the symbol and the
// first inverse level are one register-resident "middle" stage, so n = 1024 makes only four
// round trips through shared memory (n <= 16: none).  The shared tile is indexed through
// i + i/8 + i/64, which makes all the butterfly access patterns bank-conflict free.
#include "common.cuh"

namespace {

// kernel shape: NT threads per CTA, TILE complex points staged per CTA iteration
// (S = max(1, TILE / n) sequences); chosen by pt_debug_set_fft_config, default below

struct FftParams {
  const double* __restrict__ U;
  double* __restrict__ Y;
  const double2* __restrict__ tw;   // tw[n - L + i] = exp(-2 pi i * i / L), i < 3L/4, for L = n, n/4, n/16, ...
  int64_t after, before;
  int64_t ngroups, groups_per_a;
  int order;
  double s;                         // 2 pi / period
  double alpha_over_n, beta;
  const double* __restrict__ scale; // optional: result of slab/row a is multiplied by scale[a % scale_len]
  int64_t scale_len;
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {   // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ double2 mul_pi(double2 a) { return make_double2(-a.y, a.x); }   // a * (+i)

// 4-point DFT, forward (w = -i) or inverse (w = +i), in place
template <bool INV>
__device__ __forceinline__ void dft4(double2& x0, double2& x1, double2& x2, double2& x3) {
  const double2 s02 = cadd(x0, x2), d02 = csub(x0, x2), s13 = cadd(x1, x3), d13 = csub(x1, x3);
  const double2 r = INV ? mul_pi(d13) : mul_mi(d13);
  x0 = cadd(s02, s13);
  x2 = csub(s02, s13);
  x1 = cadd(d02, r);
  x3 = csub(d02, r);
}

// Twiddle table: one run of 3L/4 roots of unity per radix-4 level L = N, N/4, N/16, ... (run L starts
// at N - L; N complex numbers in all), so that the lanes of a warp, which run along j, read
// consecutive entries at every level.  TWS: the table has been copied to shared memory.
// MODE bit 0: table in shared memory; bit 1: W^2j and W^3j derived from W^j (one load per level)
template <int TWS>
__device__ __forceinline__ double2 twiddle(const double2* __restrict__ tw, int idx) {
  return (TWS & 1) ? tw[idx] : __ldg(tw + idx);
}
template <int TWS>
__device__ __forceinline__ void twiddles3(const double2* __restrict__ t, int j, double2& w1, double2& w2, double2& w3) {
  w1 = twiddle<TWS>(t, j);
  if (TWS & 2) {
    w2 = make_double2(fma(w1.x, w1.x, -w1.y * w1.y), 2.0 * w1.x * w1.y);
    w3 = cmul(w1, w2);
  } else {
    w2 = twiddle<TWS>(t, 2 * j);
    w3 = twiddle<TWS>(t, 3 * j);
  }
}

// One radix-4 level on a block of length L at offset j inside its quarter.
// forward: DFT over the 4 elements, then twiddle w_L^(j p); inverse: conj twiddle, then inverse DFT.
template <int N, int L, bool INV, int TWS>
__device__ __forceinline__ void level4(double2& v0, double2& v1, double2& v2, double2& v3, int j,
                                       const double2* __restrict__ tw) {
  const double2* __restrict__ t = tw + (N - L);
  double2 w1, w2, w3;
  if (L > 4) twiddles3<TWS>(t, j, w1, w2, w3);
  if (INV) {
    if (L > 4) { v1 = cmulc(v1, w1); v2 = cmulc(v2, w2); v3 = cmulc(v3, w3); }
    dft4<true>(v0, v1, v2, v3);
  } else {
    dft4<false>(v0, v1, v2, v3);
    if (L > 4) { v1 = cmul(v1, w1); v2 = cmul(v2, w2); v3 = cmul(v3, w3); }
  }
}

// Radix-16 butterfly = the radix-4 levels of block lengths L and L/4 on the 16 elements
// v[a][b] = x[base + a*(L/4) + b*(L/16)], base = blk*L + j, j < L/16.
template <int N, int L, bool INV, int TWS>
__device__ __forceinline__ void core16(double2 (&v)[4][4], int j, const double2* __restrict__ tw) {
  constexpr int Q2 = L / 16;
  if (!INV) {
#pragma unroll
    for (int b = 0; b < 4; ++b) level4<N, L, false, TWS>(v[0][b], v[1][b], v[2][b], v[3][b], j + b * Q2, tw);
#pragma unroll
    for (int a = 0; a < 4; ++a) level4<N, L / 4, false, TWS>(v[a][0], v[a][1], v[a][2], v[a][3], j, tw);
  } else {
#pragma unroll
    for (int a = 0; a < 4; ++a) level4<N, L / 4, true, TWS>(v[a][0], v[a][1], v[a][2], v[a][3], j, tw);
#pragma unroll
    for (int b = 0; b < 4; ++b) level4<N, L, true, TWS>(v[0][b], v[1][b], v[2][b], v[3][b], j + b * Q2, tw);
  }
}

// radix-2 level on a block of length 2 (no twiddle)
__device__ __forceinline__ void level2(double2& v0, double2& v1) {
  const double2 s = cadd(v0, v1), d = csub(v0, v1);
  v0 = s; v1 = d;
}

__device__ __forceinline__ int padi(int i) { return i + (i >> 3) + (i >> 6); }

template <int LOG2N, int TILE>
struct FftPlan {
  static constexpr int N = 1 << LOG2N;
  static constexpr int LT_LOG = (LOG2N % 4 == 0) ? 4 : (LOG2N % 4);  // middle block length 16, 2, 4 or 8
  static constexpr int LT = 1 << LT_LOG;
  static constexpr int NR16 = (LOG2N - LT_LOG) / 4;                  // radix-16 stages before the middle (0, 1, 2)
  static constexpr int S = TILE / N > 0 ? TILE / N : 1;              // sequences per CTA iteration
  static constexpr int PITCH = N + N / 8 + N / 64 + 1;   // odd for n >= 64: staggers the sequences over the banks
  static constexpr size_t SMEM = NR16 == 0 ? 0 : size_t(S) * PITCH * sizeof(double2);
};

// base-4 digit reversal of x over `digits` base-4 digits
__device__ __forceinline__ int rev4(int x, int digits) {
  if (digits == 0) return 0;
  unsigned y = __brev(unsigned(x)) >> (32 - 2 * digits);
  return int(((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u));
}

// global access of element i of staged sequence sq: one complex number = two real fields
template <int LOG2N, bool STRIDED>
struct GlobalIO {
  static constexpr int N = 1 << LOG2N;
  const double* __restrict__ U;
  double* __restrict__ Y;
  int64_t base0;       // contiguous: first real row of the group; strided: double2 offset of (a, i=0, cb)
  int64_t hb;          // strided: complex columns per row
  int64_t left;        // contiguous: real rows left from base0; strided: complex columns left from cb
  double beta;
  const double* __restrict__ scale;   // PT_SCALE_AFTER vector or nullptr
  int64_t scale_len;
  double slab_scale;   // strided: scale of this group's slab a

  __device__ __forceinline__ double2 load(int sq, int i) const {
    if (!STRIDED) {
      const int64_t r = 2 * sq;
      double2 z = make_double2(0.0, 0.0);
      if (r < left) z.x = U[(base0 + r) * N + i];
      if (r + 1 < left) z.y = U[(base0 + r + 1) * N + i];
      return z;
    }
    if (sq >= left) return make_double2(0.0, 0.0);
    return reinterpret_cast<const double2*>(U)[base0 + int64_t(i) * hb + sq];
  }
  __device__ __forceinline__ void store(int sq, int i, double2 z) const {
    if (!STRIDED) {
      const int64_t r = 2 * sq;
      if (scale) {
        if (r < left) z.x *= scale[(base0 + r) % scale_len];
        if (r + 1 < left) z.y *= scale[(base0 + r + 1) % scale_len];
      }
      if (r < left) {
        double* y = Y + (base0 + r) * N + i;
        *y = beta != 0.0 ? fma(beta, *y, z.x) : z.x;
      }
      if (r + 1 < left) {
        double* y = Y + (base0 + r + 1) * N + i;
        *y = beta != 0.0 ? fma(beta, *y, z.y) : z.y;
      }
      return;
    }
    if (sq >= left) return;
    if (scale) { z.x *= slab_scale; z.y *= slab_scale; }
    double2* y = reinterpret_cast<double2*>(Y) + base0 + int64_t(i) * hb + sq;
    if (beta != 0.0) { const double2 o = *y; z.x = fma(beta, o.x, z.x); z.y = fma(beta, o.y, z.y); }
    *y = z;
  }
};

// item -> (sequence, index inside the sequence's item range): contiguous axes keep a sequence's
// items on adjacent lanes (coalesced along i), strided axes keep the sequences adjacent (coalesced
// along the columns).
template <int S, int PER_SEQ, bool STRIDED>
__device__ __forceinline__ void split_item(int item, int& sq, int& t) {
  if (STRIDED) { sq = item % S; t = item / S; }
  else { sq = item / PER_SEQ; t = item % PER_SEQ; }
}

// One radix-16 stage on blocks of length L.  GIN: elements come from global memory (first forward
// stage); GOUT: results go to global memory (last inverse stage); otherwise the shared tile.
template <int LOG2N, int L, bool INV, bool GIN, bool GOUT, bool STRIDED, int NT, int TILE, int TWS>
__device__ __forceinline__ void stage16(double2* data, const double2* __restrict__ tw,
                                        const GlobalIO<LOG2N, STRIDED>& io, int tid) {
  using P = FftPlan<LOG2N, TILE>;
  constexpr int N = P::N, Q1 = L / 4, Q2 = L / 16, PER_SEQ = N / 16;
  for (int item = tid; item < P::S * PER_SEQ; item += NT) {
    int sq, t;
    split_item<P::S, PER_SEQ, STRIDED>(item, sq, t);
    // lanes run along j when a sixteenth of the block spans a quarter-warp, otherwise along the
    // blocks: either way the 8 lanes of a 128-bit shared access fall into 8 different bank groups
    constexpr int NBLK = N / L;
    const int blk = Q2 >= 8 ? t / Q2 : t % NBLK, j = Q2 >= 8 ? t % Q2 : t / NBLK;
    const int base = blk * L + j;
    double2* x = data + sq * P::PITCH;
    double2 v[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = base + a * Q1 + b * Q2;
        v[a][b] = GIN ? io.load(sq, i) : x[padi(i)];
      }
    core16<N, L, INV, TWS>(v, j, tw);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = base + a * Q1 + b * Q2;
        if (GOUT) io.store(sq, i, v[a][b]);
        else x[padi(i)] = v[a][b];
      }
  }
}

// multiply the value of frequency k by alpha/n * (i k s)^order
__device__ __forceinline__ double2 apply_symbol(double2 z, int k, int n, const FftParams& p) {
  int kk = k < n / 2 ? k : k - n;
  if (k == n / 2) kk = (p.order & 1) ? 0 : n / 2;
  const double ks = double(kk) * p.s;
  double pw = p.alpha_over_n;
  for (int o = 0; o < p.order; ++o) pw *= ks;
  switch (p.order & 3) {
    case 0: return make_double2(z.x * pw, z.y * pw);
    case 1: return make_double2(-z.y * pw, z.x * pw);
    case 2: return make_double2(-z.x * pw, -z.y * pw);
    default: return make_double2(z.y * pw, -z.x * pw);
  }
}

// Middle stage on blocks of length LT: last forward levels, symbol, first inverse levels.
template <int LOG2N, bool GLOBAL, bool STRIDED, int NT, int TILE, int TWS>
__device__ __forceinline__ void stage_middle(double2* data, const double2* __restrict__ tw,
                                             const GlobalIO<LOG2N, STRIDED>& io, const FftParams& p, int tid) {
  using P = FftPlan<LOG2N, TILE>;
  constexpr int N = P::N, LT = P::LT, PER_SEQ = N / LT;
  constexpr int BLK_DIGITS = (LOG2N - P::LT_LOG) / 2;   // base-4 digits of the block index
  for (int item = tid; item < P::S * PER_SEQ; item += NT) {
    int sq, blk;
    split_item<P::S, PER_SEQ, STRIDED>(item, sq, blk);
    double2* x = data + sq * P::PITCH;
    const int base = blk * LT;
    const int k0 = rev4(blk, BLK_DIGITS);              // low frequency digits (from the earlier stages)
    if constexpr (LT == 16) {
      double2 w[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) w[a][b] = GLOBAL ? io.load(sq, base + 4 * a + b) : x[padi(base + 4 * a + b)];
      core16<N, 16, false, TWS>(w, 0, tw);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) w[a][b] = apply_symbol(w[a][b], k0 + (N / 16) * (a + 4 * b), N, p);
      core16<N, 16, true, TWS>(w, 0, tw);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (GLOBAL) io.store(sq, base + 4 * a + b, w[a][b]);
          else x[padi(base + 4 * a + b)] = w[a][b];
        }
    } else {
      double2 v[LT];
#pragma unroll
      for (int q = 0; q < LT; ++q) v[q] = GLOBAL ? io.load(sq, base + q) : x[padi(base + q)];
      if constexpr (LT == 8) {
        // radix-4 level on L = 8 (quarter 2: j = 0, 1), then radix-2 on the pairs
        level4<N, 8, false, TWS>(v[0], v[2], v[4], v[6], 0, tw);
        level4<N, 8, false, TWS>(v[1], v[3], v[5], v[7], 1, tw);
#pragma unroll
        for (int a = 0; a < 4; ++a) level2(v[2 * a], v[2 * a + 1]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int e = 0; e < 2; ++e) v[2 * a + e] = apply_symbol(v[2 * a + e], k0 + (N / 8) * (a + 4 * e), N, p);
#pragma unroll
        for (int a = 0; a < 4; ++a) level2(v[2 * a], v[2 * a + 1]);
        level4<N, 8, true, TWS>(v[0], v[2], v[4], v[6], 0, tw);
        level4<N, 8, true, TWS>(v[1], v[3], v[5], v[7], 1, tw);
      } else if constexpr (LT == 4) {
        level4<N, 4, false, TWS>(v[0], v[1], v[2], v[3], 0, tw);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = apply_symbol(v[q], k0 + (N / 4) * q, N, p);
        level4<N, 4, true, TWS>(v[0], v[1], v[2], v[3], 0, tw);
      } else {
        level2(v[0], v[1]);
#pragma unroll
        for (int q = 0; q < 2; ++q) v[q] = apply_symbol(v[q], k0 + (N / 2) * q, N, p);
        level2(v[0], v[1]);
      }
#pragma unroll
      for (int q = 0; q < LT; ++q) {
        if (GLOBAL) io.store(sq, base + q, v[q]);
        else x[padi(base + q)] = v[q];
      }
    }
  }
}

template <int NT, int TWS>
constexpr int fft_min_ctas() { return NT == 128 ? ((TWS & 1) ? 4 : 5) : 2; }

template <int LOG2N, bool STRIDED, int NT, int TILE, int TWS>
__global__ void __launch_bounds__(NT, fft_min_ctas<NT, TWS>()) fourier_fft_kernel(const FftParams p) {
  using P = FftPlan<LOG2N, TILE>;
  constexpr int N = P::N;
  extern __shared__ __align__(16) double2 fsm[];
  const int tid = threadIdx.x;
  double2* data = fsm;
  const double2* __restrict__ tw = p.tw;
  if (TWS & 1) {
    double2* tws = fsm + P::S * P::PITCH;
    for (int i = threadIdx.x; i < N; i += NT) tws[i] = __ldg(p.tw + i);
    __syncthreads();
    tw = tws;
  }
  const int64_t hb = p.before >> 1;

  for (int64_t g = blockIdx.x; g < p.ngroups; g += gridDim.x) {
    GlobalIO<LOG2N, STRIDED> io;
    io.U = p.U; io.Y = p.Y; io.beta = p.beta; io.hb = hb;
    io.scale = p.scale; io.scale_len = p.scale_len; io.slab_scale = 1.0;
    if (!STRIDED) {
      io.base0 = g * (2 * P::S);
      io.left = p.after - io.base0;
    } else {
      const int64_t a = g / p.groups_per_a, cb = (g - a * p.groups_per_a) * P::S;
      io.base0 = a * N * hb + cb;
      io.left = hb - cb;
      if (p.scale) io.slab_scale = p.scale[a % p.scale_len];
    }
    if constexpr (P::NR16 == 0) {
      stage_middle<LOG2N, true, STRIDED, NT, TILE, TWS>(data, tw, io, p, tid);
    } else {
      constexpr int L2 = N >= 256 ? N / 16 : 16;   // block length of the second radix-16 stage
      __syncthreads();   // the previous group's last stage has read the tile
      stage16<LOG2N, N, false, true, false, STRIDED, NT, TILE, TWS>(data, tw, io, tid);
      __syncthreads();
      if constexpr (P::NR16 == 2) {
        stage16<LOG2N, L2, false, false, false, STRIDED, NT, TILE, TWS>(data, tw, io, tid);
        __syncthreads();
      }
      stage_middle<LOG2N, false, STRIDED, NT, TILE, TWS>(data, tw, io, p, tid);
      __syncthreads();
      if constexpr (P::NR16 == 2) {
        stage16<LOG2N, L2, true, false, false, STRIDED, NT, TILE, TWS>(data, tw, io, tid);
        __syncthreads();
      }
      stage16<LOG2N, N, true, false, true, STRIDED, NT, TILE, TWS>(data, tw, io, tid);
    }
  }
}


// ---- n = 1024 on the contiguous axis: one warp per complex sequence, 1024 = 32 x 32 -------------
// Lane j holds x[j + 32 q], q < 32, in registers and takes it through the radix-4, radix-4, radix-2
// levels of block lengths 1024, 256, 64 (table twiddles); one trip through the warp's shared tile
// regroups the sequence into 32 blocks of 32 consecutive elements, one per lane, which go through the
// levels 32, 8, 2 (twiddles are compile-time constants), the symbol, and back; a second trip returns
// the stride-32 layout for the inverse levels 64, 256, 1024 and the stores.  Two shared round trips
// instead of four, no CTA barrier (warps are independent and drift apart in phase, which is what
// overlaps the loads of one with the butterflies of the others), a third of the twiddle loads.
// The price is 32 complex numbers per thread: 252 registers, 8 warps per SM (DESIGN.md 4.4 lists the
// variants that were measured and lost: 12 warps at 168 registers, twiddles in shared memory, the next
// row pair fetched into the tile by the bulk-copy engine or into L2 by prefetch instructions).

__device__ __forceinline__ double cos32(int m) {   // cos(2 pi m / 32); m is a compile-time constant after unrolling
  m &= 31;
  const int k = m <= 8 ? m : (m <= 16 ? 16 - m : (m <= 24 ? m - 16 : 32 - m));
  double q = 0.0;
  switch (k) {
    case 0: q = 1.0; break;
    case 1: q = 0.98078528040323043; break;
    case 2: q = 0.92387953251128674; break;
    case 3: q = 0.83146961230254524; break;
    case 4: q = 0.70710678118654757; break;
    case 5: q = 0.55557023301960218; break;
    case 6: q = 0.38268343236508978; break;
    case 7: q = 0.19509032201612828; break;
    default: q = 0.0;
  }
  return (m <= 8 || m >= 24) ? q : -q;
}

// a * W_32^m (forward) or a * conj(W_32^m) (inverse)
template <bool INV>
__device__ __forceinline__ double2 rot32(double2 a, int m) {
  m &= 31;
  if (m == 0) return a;
  if (m == 8) return INV ? mul_pi(a) : mul_mi(a);
  if (m == 16) return make_double2(-a.x, -a.y);
  if (m == 24) return INV ? mul_mi(a) : mul_pi(a);
  const double2 w = make_double2(cos32(m), -cos32(m - 8));
  return INV ? cmulc(a, w) : cmul(a, w);
}

// radix-4 level with constant twiddles W_32^(step * p)
template <bool INV>
__device__ __forceinline__ void level4c(double2& v0, double2& v1, double2& v2, double2& v3, int step) {
  if (INV) {
    v1 = rot32<true>(v1, step); v2 = rot32<true>(v2, 2 * step); v3 = rot32<true>(v3, 3 * step);
    dft4<true>(v0, v1, v2, v3);
  } else {
    dft4<false>(v0, v1, v2, v3);
    v1 = rot32<false>(v1, step); v2 = rot32<false>(v2, 2 * step); v3 = rot32<false>(v3, 3 * step);
  }
}

// z * alpha/n * (i ks)^order; ORDER = 0: any order (p.order), else the compile-time order
template <int ORDER>
__device__ __forceinline__ double2 symbol_ks(double2 z, double ks, const FftParams& p) {
  double pw = p.alpha_over_n;
  if (ORDER == 1) pw *= ks;
  else if (ORDER == 2) pw *= ks * ks;
  else {
#pragma unroll 1
    for (int o = 0; o < p.order; ++o) pw *= ks;
  }
  switch ((ORDER ? ORDER : p.order) & 3) {
    case 0: return make_double2(z.x * pw, z.y * pw);
    case 1: return make_double2(-z.y * pw, z.x * pw);
    case 2: return make_double2(-z.x * pw, -z.y * pw);
    default: return make_double2(z.y * pw, -z.x * pw);
  }
}

constexpr int W1K_WARPS = 4;                 // warps per CTA (independent of each other), 2 CTAs per SM
constexpr int W1K_PITCH = 33 * 32;           // complex slots per warp tile: i + i/32
constexpr size_t W1K_SMEM = size_t(W1K_WARPS) * W1K_PITCH * sizeof(double2);

template <int MODE, int ORDER>
__global__ void __launch_bounds__(W1K_WARPS * 32, 2) fourier_fft1024_warp_kernel(const FftParams p) {
  constexpr int N = 1024;
  extern __shared__ __align__(16) double2 fsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2* x = fsm + warp * W1K_PITCH;
  const double2* __restrict__ tw = p.tw;
  const int64_t nseq = (p.after + 1) >> 1;
  // low frequency digits of the block a lane owns in the middle stage: lane = 8 d1 + 2 d2 + d3
  const int klow = (lane >> 3) + 4 * ((lane >> 1) & 3) + 16 * (lane & 1);
  const double ks_low = double(klow) * p.s;
  const bool drop_nyquist = klow == 0 && (p.order & 1);      // odd orders: the mode n/2 is dropped

  for (int64_t seq = int64_t(blockIdx.x) * W1K_WARPS + warp; seq < nseq; seq += int64_t(gridDim.x) * W1K_WARPS) {
    const int64_t r0 = 2 * seq;
    const bool two = r0 + 1 < p.after;
    // Everything that depends on the lane only (twiddles, frequencies) is loop invariant, and hoisting it
    // out of the loop costs ~150 registers that are then spilled: make it look loop variant instead.
    int ln = lane;
    double ksl = ks_low;
    asm volatile("" : "+r"(ln), "+d"(ksl));
    const double* __restrict__ u0 = p.U + r0 * N + lane;
    const double* __restrict__ u1 = u0 + (two ? N : 0);
    double2 v[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) { v[q].x = u0[32 * q]; v[q].y = u1[32 * q]; }
    if (!two) {
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q].y = 0.0;
    }
    // forward levels 1024, 256 (radix 4) and 64 (radix 2) on the elements lane + 32 q
#pragma unroll
    for (int r = 0; r < 8; ++r) level4<N, 1024, false, MODE>(v[r], v[8 + r], v[16 + r], v[24 + r], ln + 32 * r, tw);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int r = 0; r < 2; ++r)
        level4<N, 256, false, MODE>(v[8 * a + r], v[8 * a + 2 + r], v[8 * a + 4 + r], v[8 * a + 6 + r], ln + 32 * r, tw);
    double2 w64 = twiddle<MODE>(tw + (N - 64), ln);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const double2 sm = cadd(v[2 * m], v[2 * m + 1]), df = csub(v[2 * m], v[2 * m + 1]);
      v[2 * m] = sm; v[2 * m + 1] = cmul(df, w64);
    }
    __syncwarp();                            // the previous sequence's last reads of the tile are done
#pragma unroll
    for (int q = 0; q < 32; ++q) x[lane + 33 * q] = v[q];
    __syncwarp();
    // middle: block `lane`, elements 32 lane + r
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = x[33 * lane + r];
#pragma unroll
    for (int j = 0; j < 8; ++j) level4c<false>(v[j], v[8 + j], v[16 + j], v[24 + j], j);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < 2; ++j) level4c<false>(v[8 * a + j], v[8 * a + 2 + j], v[8 * a + 4 + j], v[8 * a + 6 + j], 4 * j);
#pragma unroll
    for (int m = 0; m < 16; ++m) level2(v[2 * m], v[2 * m + 1]);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // frequency klow + 32 (a + 4 b + 16 c), taken in (-n/2, n/2]
          double ks = fma(double(32 * (a + 4 * b) - 512 * c), p.s, ksl);
          if (a == 0 && b == 0 && c == 1 && drop_nyquist) ks = 0.0;
          v[8 * a + 2 * b + c] = symbol_ks<ORDER>(v[8 * a + 2 * b + c], ks, p);
        }
#pragma unroll
    for (int m = 0; m < 16; ++m) level2(v[2 * m], v[2 * m + 1]);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < 2; ++j) level4c<true>(v[8 * a + j], v[8 * a + 2 + j], v[8 * a + 4 + j], v[8 * a + 6 + j], 4 * j);
#pragma unroll
    for (int j = 0; j < 8; ++j) level4c<true>(v[j], v[8 + j], v[16 + j], v[24 + j], j);
#pragma unroll
    for (int r = 0; r < 32; ++r) x[33 * lane + r] = v[r];
    __syncwarp();
    // inverse levels 64, 256, 1024 on the elements lane + 32 q
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = x[lane + 33 * q];
    w64 = twiddle<MODE>(tw + (N - 64), ln);   // reloaded: cheaper than 4 registers held across the middle stage
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const double2 t = cmulc(v[2 * m + 1], w64);
      v[2 * m + 1] = csub(v[2 * m], t); v[2 * m] = cadd(v[2 * m], t);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int r = 0; r < 2; ++r)
        level4<N, 256, true, MODE>(v[8 * a + r], v[8 * a + 2 + r], v[8 * a + 4 + r], v[8 * a + 6 + r], ln + 32 * r, tw);
#pragma unroll
    for (int r = 0; r < 8; ++r) level4<N, 1024, true, MODE>(v[r], v[8 + r], v[16 + r], v[24 + r], ln + 32 * r, tw);
    // stores: row r0 takes the real parts, row r0 + 1 the imaginary parts
    double s0 = 1.0, s1 = 1.0;
    if (p.scale) { s0 = p.scale[r0 % p.scale_len]; s1 = p.scale[(r0 + 1) % p.scale_len]; }
    double* __restrict__ y0 = p.Y + r0 * N + lane;
    double* __restrict__ y1 = y0 + N;
    const double beta = p.beta;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      double a0 = v[q].x, a1 = v[q].y;
      if (p.scale) { a0 *= s0; a1 *= s1; }
      if (beta != 0.0) { a0 = fma(beta, y0[32 * q], a0); if (two) a1 = fma(beta, y1[32 * q], a1); }
      y0[32 * q] = a0;
      if (two) y1[32 * q] = a1;
    }
  }
}

template <int ORDER>
int launch_fft1024_warp_o(FftParams& p, cudaStream_t st) {
  auto kern = fourier_fft1024_warp_kernel<2, ORDER>;
  static bool attr_set[64] = {false};
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(W1K_SMEM)));
    attr_set[dev] = true;
  }
  const int64_t nseq = (p.after + 1) / 2;
  int64_t grid = int64_t(pt::sm_count()) * 2;
  const int64_t need = (nseq + W1K_WARPS - 1) / W1K_WARPS;
  if (grid > need) grid = need;
  kern<<<unsigned(grid), W1K_WARPS * 32, W1K_SMEM, st>>>(p);
  return pt::check_launch("pt_fourier_apply");
}

int launch_fft1024_warp(FftParams& p, cudaStream_t st) {
  if (p.order == 1) return launch_fft1024_warp_o<1>(p, st);
  if (p.order == 2) return launch_fft1024_warp_o<2>(p, st);
  return launch_fft1024_warp_o<0>(p, st);
}

// entry idx of the table: run L covers [n - L, n - L/4)
__global__ void twiddle_kernel(double2* tw, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  int L = n;
  while (L >= 4 && idx >= n - L / 4) L >>= 2;
  double sn = 0.0, cs = 1.0;
  if (L >= 4) sincospi(2.0 * double(idx - (n - L)) / double(L), &sn, &cs);
  tw[idx] = make_double2(cs, -sn);
}

inline bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

int g_fft_config = -1;  // tuning hook (pt_debug_set_fft_config); -1 = choose by shape

template <int LOG2N, bool STRIDED, int NT, int TILE, int TWS>
int launch_fft_cfg(FftParams& p, cudaStream_t st) {
  using P = FftPlan<LOG2N, TILE>;
  constexpr size_t SMEM = P::SMEM + ((TWS & 1) ? size_t(P::N) * sizeof(double2) : 0);
  auto kern = fourier_fft_kernel<LOG2N, STRIDED, NT, TILE, TWS>;
  static bool attr_set[64] = {false};
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (SMEM > 48 * 1024 && dev >= 0 && dev < 64 && !attr_set[dev]) {
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
    attr_set[dev] = true;
  }
  if (STRIDED) {
    const int64_t hb = p.before / 2;
    p.groups_per_a = (hb + P::S - 1) / P::S;
    p.ngroups = p.after * p.groups_per_a;
  } else {
    p.groups_per_a = 1;
    p.ngroups = (p.after + 2 * P::S - 1) / (2 * P::S);
  }
  int64_t grid = int64_t(pt::sm_count()) * fft_min_ctas<NT, TWS>();
  if (grid > p.ngroups) grid = p.ngroups;
  kern<<<unsigned(grid), NT, SMEM, st>>>(p);
  return pt::check_launch("pt_fourier_apply");
}

template <int LOG2N, bool STRIDED>
int launch_fft(FftParams& p, cudaStream_t st) {
  // Tuning ids (pt_debug_set_fft_config): 0/1 = 256x4096 / 128x2048 threads x points with the twiddles
  // through L1; 3/4 = 128x2048 / 256x4096 with the table in shared memory; 5/11 = as 1/0 with derived
  // twiddles; 8 = the warp-per-sequence kernel.
  // Shared-memory twiddles need at least one shared round trip (n >= 32) and a table that leaves room
  // for the tiles (n <= 2048).
  constexpr bool CAN_TWS = LOG2N >= 5 && LOG2N <= 11;
  // Defaults measured on B200 (tools/perf_probe.py fftcfg): contiguous axis: shared-memory
  // twiddles up to n = 256, the warp kernel at n = 1024, derived twiddles through L1 elsewhere; strided
  // axis: 256 x 4096 with shared-memory twiddles up to n = 1024, derived twiddles above.
  constexpr int DEFAULT = LOG2N < 5 ? 0
                          : !STRIDED ? (LOG2N <= 8 ? 3 : (LOG2N == 10 ? 8 : 5))
                                     : (LOG2N <= 10 ? 4 : 11);
  int cfg = g_fft_config >= 0 ? g_fft_config : DEFAULT;
  if constexpr (LOG2N == 10 && !STRIDED) {
    if (cfg == 8) return launch_fft1024_warp(p, st);
  }
  if (cfg == 8) cfg = 1;   // (the warp kernel exists for n = 1024 on the contiguous axis only)
  if constexpr (CAN_TWS) {
    if (cfg == 3) return launch_fft_cfg<LOG2N, STRIDED, 128, 2048, 3>(p, st);
    if (cfg == 4) return launch_fft_cfg<LOG2N, STRIDED, 256, 4096, 3>(p, st);
  }
  if (cfg == 5) return launch_fft_cfg<LOG2N, STRIDED, 128, 2048, 2>(p, st);
  if (cfg == 11) return launch_fft_cfg<LOG2N, STRIDED, 256, 4096, 2>(p, st);
  if (cfg == 0 || cfg == 4) return launch_fft_cfg<LOG2N, STRIDED, 256, 4096, 0>(p, st);
  return launch_fft_cfg<LOG2N, STRIDED, 128, 2048, 0>(p, st);
}

template <bool STRIDED>
int dispatch_fft(int log2n, FftParams& p, cudaStream_t st) {
  switch (log2n) {
    case 2: return launch_fft<2, STRIDED>(p, st);
    case 3: return launch_fft<3, STRIDED>(p, st);
    case 4: return launch_fft<4, STRIDED>(p, st);
    case 5: return launch_fft<5, STRIDED>(p, st);
    case 6: return launch_fft<6, STRIDED>(p, st);
    case 7: return launch_fft<7, STRIDED>(p, st);
    case 8: return launch_fft<8, STRIDED>(p, st);
    case 9: return launch_fft<9, STRIDED>(p, st);
    case 10: return launch_fft<10, STRIDED>(p, st);
    case 11: return launch_fft<11, STRIDED>(p, st);
    default: return launch_fft<12, STRIDED>(p, st);
  }
}

}  // namespace

extern "C" int pt_debug_set_fft_config(int id) {
  g_fft_config = id;
  return PT_OK;
}

extern "C" int pt_fourier_twiddles(int n, double* tw, pt_stream_t stream) {
  PT_REQUIRE(tw && n >= 4 && (n & (n - 1)) == 0, "pt_fourier_twiddles: n=%d is not a power of two >= 4", n);
  twiddle_kernel<<<(n + 255) / 256, 256, 0, pt::as_stream(stream)>>>(reinterpret_cast<double2*>(tw), n);
  return pt::check_launch("pt_fourier_twiddles");
}

extern "C" int pt_fourier_apply(const double* tw, int n, double period, int order, const double* U, double* Y,
                                int64_t after, int64_t before, double alpha, double beta,
                                const double* scale, int64_t scale_len, pt_stream_t stream) {
  PT_REQUIRE(tw && U && Y, "pt_fourier_apply: null pointer");
  PT_REQUIRE(scale == nullptr || scale_len > 0, "pt_fourier_apply: scale_len must be > 0");
  PT_REQUIRE(order >= 1 && order <= 8, "pt_fourier_apply: order=%d outside [1,8]", order);
  PT_REQUIRE(period > 0.0, "pt_fourier_apply: period must be positive");
  PT_REQUIRE(after >= 0 && before >= 1, "pt_fourier_apply: bad field shape");
  PT_REQUIRE(U != Y, "pt_fourier_apply: in-place application is not supported");
  if (n < 4 || (n & (n - 1)) != 0 || n > 4096)
    return pt::fail(PT_ERR_UNSUPPORTED, "pt_fourier_apply: n=%d is not a power of two in [4, 4096]", n);
  const bool strided = before > 1;
  if (strided && ((before & 1) || !aligned16(U) || !aligned16(Y)))
    return pt::fail(PT_ERR_UNSUPPORTED, "pt_fourier_apply: a strided axis needs an even, 16-byte aligned stride");
  if (!aligned16(tw)) return pt::fail(PT_ERR_INVALID, "pt_fourier_apply: twiddle table must be 16-byte aligned");
  if (after == 0) return PT_OK;

  FftParams p;
  p.U = U; p.Y = Y; p.tw = reinterpret_cast<const double2*>(tw);
  int log2n = 0;
  while ((1 << log2n) < n) ++log2n;
  p.after = after; p.before = before;
  p.order = order; p.s = 6.283185307179586476925286766559 / period;
  p.alpha_over_n = alpha / double(n); p.beta = beta;
  p.scale = scale; p.scale_len = scale ? scale_len : 1;
  cudaStream_t st = pt::as_stream(stream);
  return strided ? dispatch_fft<true>(log2n, p, st) : dispatch_fft<false>(log2n, p, st);
}
