// Shared device helpers of the dense-apply kernels (plain operators: dense_apply.cu; centrosymmetric
// operators: dense_apply_eo.cu and the warp-specialised dense_apply_eo_ws.cu).
#pragma once
#include "common.cuh"

namespace ptdense {

// Destination map of the scatter epilogue (pt_dense_apply_scatter): output element (a, i, b) is
// stored at base[i / chunk] + offset + (a / a_inner)*s_outer + (a % a_inner)*s_inner
// + (i % chunk)*s_i + b, where base[] are peer-mapped buffers of the ranks of a pencil row/column.
struct ScatterDev {
  double* const* base;
  int chunk;
  int64_t a_inner, s_outer, s_inner, s_i, offset;
};

struct DenseParams {
  const double* __restrict__ Dp;
  const double* __restrict__ U;
  double* __restrict__ Y;
  const double* __restrict__ C;   // beta input of the scatter epilogue (local, laid out like Y)
  const double* __restrict__ scale;
  ScatterDev sc;
  int n_out, n_in;
  int MP, KP;            // padded rows of the packed operator, number of k-panels
  int mt_per_block;      // m-tiles (8 rows) handled by one CTA
  int m_blocks;
  int64_t before;
  int64_t NN;            // number of field columns: after*before (K1) or after (K2)
  int64_t slab_in;       // n_in*before: distance between consecutive `a` slabs of U
  int64_t slab_out;      // n_out*before
  double alpha, beta;
  int scale_mode;
  int64_t scale_len;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// WARPS_M x WARPS_N warps, each owning up to WM m-tiles x WN n-tiles of 8x8.
template <int WARPS_M, int WARPS_N, int WM, int WN, int BKP, int STAGES, bool KCONTIG, bool VEC16>
struct Cfg {
  static constexpr int NT = WARPS_M * WARPS_N * 32;
  static constexpr int BM = WARPS_M * WM * 8;
  static constexpr int BN = WARPS_N * WN * 8;
  static constexpr int BK = BKP * 4;
  static constexpr int A_STAGE = BKP * BM * 4;                     // doubles
  static constexpr int PB = KCONTIG ? (BK + 4) : (BN + 4);          // pitch, = 4 mod 16
  static constexpr int B_STAGE = KCONTIG ? BN * PB : BK * PB;       // doubles
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr size_t SMEM = size_t(STAGES) * STAGE * sizeof(double);
};


// ---- per-warp compute over one staged k-chunk, specialised on the exact number of m-tiles (MT) the
// warp owns.  Predicated-off DMMAs still occupy the tensor pipe for their full 16 cycles (measured:
// ncu r01a), so a warp that owns fewer than WM tiles must branch to code that only issues its own.
template <class C, int MT, int WMAX, int WN, bool KCONTIG>
__device__ __forceinline__ void load_frags(const double* As, const double* Bs, int pp, double (&af)[WMAX],
                                           double (&bf)[WN]) {
#pragma unroll
  for (int i = 0; i < MT; ++i) af[i] = As[(pp * C::BM + i * 8) * 4];
#pragma unroll
  for (int j = 0; j < WN; ++j) bf[j] = KCONTIG ? Bs[j * 8 * C::PB + pp * 4] : Bs[pp * 4 * C::PB + j * 8];
}

template <class C, int MT, int WMAX, int WN, int BKP, bool KCONTIG>
__device__ __forceinline__ void compute_full(const double* As, const double* Bs, double (&acc)[WMAX][WN][2]) {
  // software pipelined: fragments of panel pp+1 are loaded before the DMMAs of panel pp issue
  double af[2][WMAX], bf[2][WN];
  load_frags<C, MT, WMAX, WN, KCONTIG>(As, Bs, 0, af[0], bf[0]);
#pragma unroll
  for (int pp = 0; pp < BKP; ++pp) {
    if (pp + 1 < BKP) load_frags<C, MT, WMAX, WN, KCONTIG>(As, Bs, pp + 1, af[(pp + 1) & 1], bf[(pp + 1) & 1]);
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[pp & 1][i], bf[pp & 1][j]);
  }
}

template <class C, int MT, int WMAX, int WN, bool KCONTIG>
__device__ __forceinline__ void compute_partial(const double* As, const double* Bs, double (&acc)[WMAX][WN][2],
                                                int npanels) {
#pragma unroll 1
  for (int pp = 0; pp < npanels; ++pp) {
    double af[WMAX], bf[WN];
    load_frags<C, MT, WMAX, WN, KCONTIG>(As, Bs, pp, af, bf);
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
}

template <class C, int MT, int WMAX, int WN, int BKP, bool KCONTIG>
struct MtDispatch {
  static __device__ __forceinline__ void full(int my_mt, const double* As, const double* Bs,
                                              double (&acc)[WMAX][WN][2]) {
    if (my_mt == MT) compute_full<C, MT, WMAX, WN, BKP, KCONTIG>(As, Bs, acc);
    else MtDispatch<C, MT - 1, WMAX, WN, BKP, KCONTIG>::full(my_mt, As, Bs, acc);
  }
  static __device__ __forceinline__ void partial(int my_mt, const double* As, const double* Bs,
                                                 double (&acc)[WMAX][WN][2], int npanels) {
    if (my_mt == MT) compute_partial<C, MT, WMAX, WN, KCONTIG>(As, Bs, acc, npanels);
    else MtDispatch<C, MT - 1, WMAX, WN, BKP, KCONTIG>::partial(my_mt, As, Bs, acc, npanels);
  }
};
template <class C, int WMAX, int WN, int BKP, bool KCONTIG>
struct MtDispatch<C, 0, WMAX, WN, BKP, KCONTIG> {
  static __device__ __forceinline__ void full(int, const double*, const double*, double (&)[WMAX][WN][2]) {}
  static __device__ __forceinline__ void partial(int, const double*, const double*, double (&)[WMAX][WN][2], int) {}
};


}  // namespace ptdense
