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
  const double2* __restrict__ tw;   // tw[m] = exp(-2 pi i m / n), m < n
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

// One radix-4 level on a block of length L at offset j inside its quarter.
// forward: DFT over the 4 elements, then twiddle w_L^(j p); inverse: conj twiddle, then inverse DFT.
template <int N, int L, bool INV>
__device__ __forceinline__ void level4(double2& v0, double2& v1, double2& v2, double2& v3, int j,
                                       const double2* __restrict__ tw) {
  const int m = j * (N / L);
  if (INV) {
    if (L > 4) { v1 = cmulc(v1, __ldg(tw + m)); v2 = cmulc(v2, __ldg(tw + 2 * m)); v3 = cmulc(v3, __ldg(tw + 3 * m)); }
    dft4<true>(v0, v1, v2, v3);
  } else {
    dft4<false>(v0, v1, v2, v3);
    if (L > 4) { v1 = cmul(v1, __ldg(tw + m)); v2 = cmul(v2, __ldg(tw + 2 * m)); v3 = cmul(v3, __ldg(tw + 3 * m)); }
  }
}

// Radix-16 butterfly = the radix-4 levels of block lengths L and L/4 on the 16 elements
// v[a][b] = x[base + a*(L/4) + b*(L/16)], base = blk*L + j, j < L/16.
template <int N, int L, bool INV>
__device__ __forceinline__ void core16(double2 (&v)[4][4], int j, const double2* __restrict__ tw) {
  constexpr int Q2 = L / 16;
  if (!INV) {
#pragma unroll
    for (int b = 0; b < 4; ++b) level4<N, L, false>(v[0][b], v[1][b], v[2][b], v[3][b], j + b * Q2, tw);
#pragma unroll
    for (int a = 0; a < 4; ++a) level4<N, L / 4, false>(v[a][0], v[a][1], v[a][2], v[a][3], j, tw);
  } else {
#pragma unroll
    for (int a = 0; a < 4; ++a) level4<N, L / 4, true>(v[a][0], v[a][1], v[a][2], v[a][3], j, tw);
#pragma unroll
    for (int b = 0; b < 4; ++b) level4<N, L, true>(v[0][b], v[1][b], v[2][b], v[3][b], j + b * Q2, tw);
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
template <int LOG2N, int L, bool INV, bool GIN, bool GOUT, bool STRIDED, int NT, int TILE>
__device__ __forceinline__ void stage16(double2* data, const double2* __restrict__ tw,
                                        const GlobalIO<LOG2N, STRIDED>& io) {
  using P = FftPlan<LOG2N, TILE>;
  constexpr int N = P::N, Q1 = L / 4, Q2 = L / 16, PER_SEQ = N / 16;
  for (int item = threadIdx.x; item < P::S * PER_SEQ; item += NT) {
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
    core16<N, L, INV>(v, j, tw);
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
template <int LOG2N, bool GLOBAL, bool STRIDED, int NT, int TILE>
__device__ __forceinline__ void stage_middle(double2* data, const double2* __restrict__ tw,
                                             const GlobalIO<LOG2N, STRIDED>& io, const FftParams& p) {
  using P = FftPlan<LOG2N, TILE>;
  constexpr int N = P::N, LT = P::LT, PER_SEQ = N / LT;
  constexpr int BLK_DIGITS = (LOG2N - P::LT_LOG) / 2;   // base-4 digits of the block index
  for (int item = threadIdx.x; item < P::S * PER_SEQ; item += NT) {
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
      core16<N, 16, false>(w, 0, tw);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) w[a][b] = apply_symbol(w[a][b], k0 + (N / 16) * (a + 4 * b), N, p);
      core16<N, 16, true>(w, 0, tw);
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
        level4<N, 8, false>(v[0], v[2], v[4], v[6], 0, tw);
        level4<N, 8, false>(v[1], v[3], v[5], v[7], 1, tw);
#pragma unroll
        for (int a = 0; a < 4; ++a) level2(v[2 * a], v[2 * a + 1]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int e = 0; e < 2; ++e) v[2 * a + e] = apply_symbol(v[2 * a + e], k0 + (N / 8) * (a + 4 * e), N, p);
#pragma unroll
        for (int a = 0; a < 4; ++a) level2(v[2 * a], v[2 * a + 1]);
        level4<N, 8, true>(v[0], v[2], v[4], v[6], 0, tw);
        level4<N, 8, true>(v[1], v[3], v[5], v[7], 1, tw);
      } else if constexpr (LT == 4) {
        level4<N, 4, false>(v[0], v[1], v[2], v[3], 0, tw);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = apply_symbol(v[q], k0 + (N / 4) * q, N, p);
        level4<N, 4, true>(v[0], v[1], v[2], v[3], 0, tw);
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

template <int LOG2N, bool STRIDED, int NT, int TILE>
__global__ void __launch_bounds__(NT, NT == 128 ? 5 : 2) fourier_fft_kernel(const FftParams p) {
  using P = FftPlan<LOG2N, TILE>;
  constexpr int N = P::N;
  extern __shared__ __align__(16) double2 fsm[];
  double2* data = fsm;
  const double2* __restrict__ tw = p.tw;
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
      stage_middle<LOG2N, true, STRIDED, NT, TILE>(data, tw, io, p);
    } else {
      constexpr int L2 = N >= 256 ? N / 16 : 16;   // block length of the second radix-16 stage
      __syncthreads();   // the previous group's last stage has read the tile
      stage16<LOG2N, N, false, true, false, STRIDED, NT, TILE>(data, tw, io);
      __syncthreads();
      if constexpr (P::NR16 == 2) {
        stage16<LOG2N, L2, false, false, false, STRIDED, NT, TILE>(data, tw, io);
        __syncthreads();
      }
      stage_middle<LOG2N, false, STRIDED, NT, TILE>(data, tw, io, p);
      __syncthreads();
      if constexpr (P::NR16 == 2) {
        stage16<LOG2N, L2, true, false, false, STRIDED, NT, TILE>(data, tw, io);
        __syncthreads();
      }
      stage16<LOG2N, N, true, false, true, STRIDED, NT, TILE>(data, tw, io);
    }
  }
}

__global__ void twiddle_kernel(double2* tw, int n) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  double sn, cs;
  sincospi(2.0 * double(m) / double(n), &sn, &cs);
  tw[m] = make_double2(cs, -sn);
}

inline bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

int g_fft_config = -1;  // tuning hook (pt_debug_set_fft_config); -1 = choose by shape

template <int LOG2N, bool STRIDED, int NT, int TILE>
int launch_fft_cfg(FftParams& p, cudaStream_t st) {
  using P = FftPlan<LOG2N, TILE>;
  auto kern = fourier_fft_kernel<LOG2N, STRIDED, NT, TILE>;
  static bool attr_set[64] = {false};
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (P::SMEM > 48 * 1024 && dev >= 0 && dev < 64 && !attr_set[dev]) {
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(P::SMEM)));
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
  int64_t grid = int64_t(pt::sm_count()) * (NT == 128 ? 5 : 2);
  if (grid > p.ngroups) grid = p.ngroups;
  kern<<<unsigned(grid), NT, P::SMEM, st>>>(p);
  return pt::check_launch("pt_fourier_apply");
}

template <int LOG2N, bool STRIDED>
int launch_fft(FftParams& p, cudaStream_t st) {
  // measured on B200 (tools/perf_probe.py fftcfg): 128 threads x 2048 points with 5 CTAs/SM wins for
  // the contiguous n = 1024 axis (both BASELINE Fourier axes), 256 x 4096 with 2 CTAs/SM elsewhere
  const int cfg = g_fft_config >= 0 ? g_fft_config : ((LOG2N == 10 && !STRIDED) ? 1 : 0);
  switch (cfg) {
    case 0: return launch_fft_cfg<LOG2N, STRIDED, 256, 4096>(p, st);
    default: return launch_fft_cfg<LOG2N, STRIDED, 128, 2048>(p, st);
  }
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
