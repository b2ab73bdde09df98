// Shared pieces of the two even-odd kernels (dense_apply_eo.cu: every warp loads and computes,
// cp.async ring, 2-4 CTAs/SM; dense_apply_eo_ws.cu: warp-specialised, producers form E/O and the
// bulk-copy engine stages the operator panels, mbarrier ring, 1 CTA/SM): parameter block and the
// fused epilogue  Y = alpha*scale*(D U) + beta*C  with rows i and n-1-i written by the same thread.
#pragma once
#include "dense_common.cuh"

namespace ptdense {

struct EoParams {
  const double* __restrict__ Dpe;   // packed Ae (dense_apply_eo.cu) | packed panels (dense_apply_eo_ws.cu)
  const double* __restrict__ Dpo;   // packed Ao                     | unused
  const double* __restrict__ U;
  double* __restrict__ Y;
  const double* __restrict__ C;     // beta input (== Y unless scattering)
  const double* __restrict__ scale;
  ScatterDev sc;
  ScatterDev cp;                    // warp-specialised kernel, COPY: second map through which U itself is redistributed
  int n, H;                         // operator size; half size (rows and k of the half problem)
  int MPH, KPH;                     // padded half rows, k-panels of the half problem (KPH is even)
  int rt_per_block, m_blocks;       // row-tiles (8 half-rows) per tile, tiles along the half-rows
  int64_t total_tiles;              // m_blocks * column tiles
  int sign;                         // +1: J D J = D, -1: J D J = -D
  int64_t before, NN, slab;         // slab = n*before
  double alpha, beta;
  int scale_mode;
  int64_t scale_len;
  // odd n in the warp-specialised kernel: the middle row/column are rank-1 terms beside the
  // (n-1)/2 x (n-1)/2 half problem
  const double* __restrict__ cmid;  // cmid[i] = D[i][mid], i < H
  const double* __restrict__ rmid;  // rmid[k] = D[mid][k], k < H
  double dmm;                       // D[mid][mid]
  int mid;                          // (n-1)/2 when n is odd and the middle is split off, else -1
  int dbg;                          // timing experiments only (pt_debug_set_eo_ws): 3 = no epilogue
  long long* trace;                 // tuning hook (pt_debug_set_eo_ws_trace): per-CTA, per-tile clock64 stamps, or null
};

// column index nn -> (a, b) = (nn / before, nn % before): 32-bit division when it fits (a 64-bit
// division is ~10x the instructions and sits on the critical path of every tile's epilogue)
__device__ __forceinline__ void split_column(int64_t nn, int64_t before, int64_t& a, int64_t& b) {
  if ((uint64_t(nn) | uint64_t(before)) >> 32) {
    a = nn / before;
  } else {
    a = int64_t(uint32_t(nn) / uint32_t(before));
  }
  b = nn - a * before;
}

// Epilogue of one warp: acce/acco hold ze/zo for half-rows row0 + 8*i + g (i < my_rt) and columns
// ncol0 + 8*j + 2*t + {0, 1}.  Rows i get ze + zo, rows n-1-i get s*(ze - zo).  MIDFIX: n is odd and
// the middle column is not part of the half problem: ze += D[i][mid] * u[mid] first; cmid_s (the
// column D[:, mid]) and umid_s (u[mid] of the tile's 64 columns, indexed by column - ncol_tile) then
// come from shared memory (staged by the warp-specialised kernel), so no global load sits between
// the accumulators and the stores.
// MIDADD = false: the caller has already added the middle-column term to acce (the warp-specialised kernel does
// it with one DMMA per accumulator tile: plain FMAs issued here queue behind the other warps' DMMAs).
template <int WR, int WN, bool KCONTIG, bool VEC16, bool SCAT, bool MIDFIX, bool MIDADD = MIDFIX>
__device__ __forceinline__ void eo_epilogue(const EoParams& p, double (&acce)[WR][WN][2], double (&acco)[WR][WN][2],
                                            int my_rt, int row0, int64_t ncol0, int g, int t,
                                            const double* cmid_s = nullptr, const double* umid_s = nullptr) {
  const int n = p.n;
  const bool use_beta = (p.beta != 0.0);
  const double sgn = double(p.sign);
  // ---- interior tiles without scatter (the bulk of every large field): straight-line code, no
  // per-element bounds checks, running pointers.  Same operations in the same order as the general
  // path below, so the results are bit-identical.
  // (an odd n whose middle row is part of the half problem — cp.async kernel — keeps the general path:
  // there rows i and n-1-i coincide for i = mid)
  if (!SCAT && (KCONTIG || VEC16) && (MIDFIX || (n & 1) == 0) && my_rt == WR && row0 + WR * 8 <= p.H &&
      ncol0 + WN * 8 <= p.NN) {
    const int64_t step = 8 * p.before;
    const bool row_scale = p.scale_mode == PT_SCALE_ROW;
    const bool unit = p.scale_mode == PT_SCALE_NONE && p.alpha == 1.0;
    const bool pos = p.sign > 0;
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      const int64_t nn = ncol0 + j * 8 + 2 * t;
      int64_t l0, a0, b0;
      if (KCONTIG) {
        a0 = nn; b0 = 0;
        l0 = nn * p.slab;
      } else {
        split_column(nn, p.before, a0, b0);
        l0 = a0 * p.slab + b0;
      }
      double s0 = p.alpha, s1 = p.alpha;                  // column nn, column nn + 1
      if (p.scale_mode == PT_SCALE_AFTER) {
        s0 *= p.scale[a0 % p.scale_len];
        s1 = KCONTIG ? p.alpha * p.scale[(a0 + 1) % p.scale_len] : s0;     // K1: both columns in one slab
      } else if (p.scale_mode == PT_SCALE_BEFORE) {
        s0 *= p.scale[b0];
        s1 *= p.scale[KCONTIG ? 0 : b0 + 1];
      }
      double um0 = 0.0, um1 = 0.0;
      if (MIDADD) {
        const double2 um = *reinterpret_cast<const double2*>(umid_s + (j * 8 + 2 * t));
        um0 = um.x;
        um1 = um.y;
      }
      double* ylo = p.Y + l0 + int64_t(row0 + g) * p.before;
      double* yhi = p.Y + l0 + int64_t(n - 1 - row0 - g) * p.before;
#pragma unroll
      for (int i = 0; i < WR; ++i, ylo += step, yhi -= step) {
        double e0 = acce[i][j][0], e1 = acce[i][j][1];
        if (MIDADD) {
          const double cm = cmid_s[row0 + i * 8 + g];      // shared memory: no preloaded array (registers are tight)
          e0 = __fma_rn(cm, um0, e0);
          e1 = __fma_rn(cm, um1, e1);
        }
        double rl0 = s0, rl1 = s1, rh0 = s0, rh1 = s1;
        if (row_scale) {
          const double sl = p.scale[row0 + i * 8 + g], sh = p.scale[n - 1 - row0 - i * 8 - g];
          rl0 *= sl; rl1 *= sl; rh0 *= sh; rh1 *= sh;
        }
        // FP64 instructions are expensive here: whatever a warp issues on the FP64 pipe queues behind the
        // DMMAs of the warps that are still (or again) in their main loop.  So: the sign of the mirrored
        // rows is an operand swap (o - e = -(e - o) exactly), and alpha = 1 without scaling multiplies nothing.
        const double o0 = acco[i][j][0], o1 = acco[i][j][1];
        double v0 = __dadd_rn(e0, o0), v1 = __dadd_rn(e1, o1);
        double x0 = pos ? __dsub_rn(e0, o0) : __dsub_rn(o0, e0);
        double x1 = pos ? __dsub_rn(e1, o1) : __dsub_rn(o1, e1);
        if (!unit) {
          v0 = __dmul_rn(rl0, v0); v1 = __dmul_rn(rl1, v1);
          x0 = __dmul_rn(rh0, x0); x1 = __dmul_rn(rh1, x1);
        }
        if (KCONTIG) {          // columns nn and nn + 1 are n apart, rows contiguous
          if (use_beta) {
            v0 = __fma_rn(p.beta, ylo[0], v0); v1 = __fma_rn(p.beta, ylo[n], v1);
            x0 = __fma_rn(p.beta, yhi[0], x0); x1 = __fma_rn(p.beta, yhi[n], x1);
          }
          ylo[0] = v0; ylo[n] = v1;
          yhi[0] = x0; yhi[n] = x1;
        } else {
          if (use_beta) {
            const double2 o = *reinterpret_cast<const double2*>(ylo), q = *reinterpret_cast<const double2*>(yhi);
            v0 = __fma_rn(p.beta, o.x, v0); v1 = __fma_rn(p.beta, o.y, v1);
            x0 = __fma_rn(p.beta, q.x, x0); x1 = __fma_rn(p.beta, q.y, x1);
          }
          *reinterpret_cast<double2*>(ylo) = make_double2(v0, v1);
          *reinterpret_cast<double2*>(yhi) = make_double2(x0, x1);
        }
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < WN; ++j) {
    const int64_t nn = ncol0 + j * 8 + 2 * t;
    if (nn >= p.NN) continue;
    const bool two = (nn + 1 < p.NN);
    int64_t a0, b0, a1, b1;
    if (KCONTIG) {
      a0 = nn; b0 = 0; a1 = nn + 1; b1 = 0;
    } else {
      split_column(nn, p.before, a0, b0);
      if (b0 + 1 < p.before) { a1 = a0; b1 = b0 + 1; } else { a1 = a0 + 1; b1 = 0; }
    }
    double s0 = p.alpha, s1 = p.alpha;
    if (p.scale_mode == PT_SCALE_AFTER) {
      s0 *= p.scale[a0 % p.scale_len];
      if (two) s1 *= p.scale[a1 % p.scale_len];
    } else if (p.scale_mode == PT_SCALE_BEFORE) {
      s0 *= p.scale[b0];
      if (two) s1 *= p.scale[b1];
    }
    const int64_t l0 = a0 * p.slab + b0, l1 = a1 * p.slab + b1;
    double um0 = 0.0, um1 = 0.0;
    if (MIDADD) {
      const double2 um = *reinterpret_cast<const double2*>(umid_s + (j * 8 + 2 * t));   // umid_s already at the warp's columns
      um0 = um.x;
      um1 = um.y;
    }
    int64_t d0 = 0, d1 = 0;
    if (SCAT) {
      const int64_t q0 = a0 / p.sc.a_inner, q1 = a1 / p.sc.a_inner;
      d0 = p.sc.offset + q0 * p.sc.s_outer + (a0 - q0 * p.sc.a_inner) * p.sc.s_inner + b0;
      d1 = p.sc.offset + q1 * p.sc.s_outer + (a1 - q1 * p.sc.a_inner) * p.sc.s_inner + b1;
    }
    // running offsets of rows i (upwards) and n-1-i (downwards): 8 rows per accumulator tile
    int64_t off_lo = int64_t(row0 + g) * p.before;
    int64_t off_hi = int64_t(n - 1 - row0 - g) * p.before;
    const int64_t off_step = 8 * p.before;
#pragma unroll
    for (int i = 0; i < WR; ++i, off_lo += off_step, off_hi -= off_step) {
      if (i >= my_rt) continue;
      const int hrow = row0 + i * 8 + g;
      if (hrow >= p.H) continue;
      double e0 = acce[i][j][0], e1 = acce[i][j][1];
      if (MIDADD) {
        const double cm = cmid_s[hrow];
        e0 = __fma_rn(cm, um0, e0);
        e1 = __fma_rn(cm, um1, e1);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = half == 0 ? hrow : n - 1 - hrow;
        if (half == 1 && row == hrow) continue;          // middle row of an odd n: written once
        double w0, w1;
        if (half == 0) { w0 = __dadd_rn(e0, acco[i][j][0]); w1 = __dadd_rn(e1, acco[i][j][1]); }
        else { w0 = __dmul_rn(sgn, __dsub_rn(e0, acco[i][j][0])); w1 = __dmul_rn(sgn, __dsub_rn(e1, acco[i][j][1])); }
        double r0 = s0, r1 = s1;
        if (p.scale_mode == PT_SCALE_ROW) { const double sr = p.scale[row]; r0 *= sr; r1 *= sr; }
        double v0 = __dmul_rn(r0, w0), v1 = __dmul_rn(r1, w1);
        const int64_t off = half == 0 ? off_lo : off_hi;
        const double* c0 = (SCAT ? p.C : p.Y) + l0 + off;
        const double* c1 = (SCAT ? p.C : p.Y) + l1 + off;
        double *y0, *y1;
        if (SCAT) {
          const int dq = row / p.sc.chunk;
          double* db = p.sc.base[dq] + int64_t(row - dq * p.sc.chunk) * p.sc.s_i;
          y0 = db + d0; y1 = db + d1;
        } else {
          y0 = p.Y + l0 + off; y1 = p.Y + l1 + off;
        }
        if (!KCONTIG && VEC16) {
          if (use_beta) { const double2 o = *reinterpret_cast<const double2*>(c0); v0 = __fma_rn(p.beta, o.x, v0); v1 = __fma_rn(p.beta, o.y, v1); }
          *reinterpret_cast<double2*>(y0) = make_double2(v0, v1);
        } else {
          if (use_beta) v0 = __fma_rn(p.beta, *c0, v0);
          *y0 = v0;
          if (two) {
            if (use_beta) v1 = __fma_rn(p.beta, *c1, v1);
            *y1 = v1;
          }
        }
      }
    }
  }
}

// entry point of the warp-specialised kernel (dense_apply_eo_ws.cu); returns PT_ERR_UNSUPPORTED,
// without a message and without doing anything, for shapes it does not cover
int launch_eo_ws(EoParams& p, bool kcontig, bool scatter, bool copy, cudaStream_t stream);

}  // namespace ptdense
