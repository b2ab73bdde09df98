// K1/K2 for centrosymmetric operators — the even-odd decomposition on FP64 tensor cores.
//
// Fourier differentiation matrices, and Chebyshev collocation matrices on symmetric node sets,
// satisfy J D J = s D with J the order reversal and s = (-1)^order.  Writing, for k < H = ceil(n/2),
//     E[k] = u[k] + u[n-1-k],   O[k] = u[k] - u[n-1-k],
//     Ae[i][k] = (D[i][k] + D[i][n-1-k]) / 2,   Ao[i][k] = (D[i][k] - D[i][n-1-k]) / 2
// (column (n-1)/2 of Ae is D[i][k]/2 when n is odd: E doubles that element), the rows i < H are
//     y[i] = sum_k Ae[i][k] E[k] + sum_k Ao[i][k] O[k] = ze[i] + zo[i]            (exact algebra)
// and the symmetry gives the other half for free:   y[n-1-i] = s * (ze[i] - zo[i]).
// Two H x H products replace one n x n product: HALF the tensor-core work for the same result
// (Solomonoff 1992).  Everything is fused here: E and O are formed from the two staged field
// tiles when the B fragments are read (one DADD + one DSUB per fragment), ze and zo accumulate in
// the same thread, and the epilogue writes rows i and n-1-i — no extra pass over the field, no
// extra HBM traffic.  Layout conventions, staging (cp.async ring, fragment-major operator panels,
// pitch = 4 mod 16) and the fused alpha/beta/scale/scatter epilogue are those of dense_apply.cu.
#include "dense_eo_common.cuh"
#include <string.h>

using namespace ptdense;

namespace {

// WARPS_M x WARPS_N warps; a warp owns up to WR row-tiles of the half problem (each with an even and
// an odd accumulator set) x WN column tiles.
template <int WARPS_M, int WARPS_N, int WR, int WN, int BKP, int STAGES, bool KCONTIG>
struct EoCfg {
  static constexpr int NT = WARPS_M * WARPS_N * 32;
  static constexpr int BMH = WARPS_M * WR * 8;                      // half-rows per CTA
  static constexpr int BN = WARPS_N * WN * 8;
  static constexpr int BK = BKP * 4;
  static constexpr int A_STAGE = BKP * 2 * BMH * 4;                 // per panel PAIR: [even rows][odd rows] x 8
  // field tile pitch: no padding, XOR swizzles instead (a padded pitch that is not a multiple of
  // 128 bytes costs the 16-byte cp.async writes 2-3 shared-memory wavefronts each: ncu r02a).
  // K1 ([k][column], 8-byte fragment reads of rows 2t and 2t+1): column index XOR 4*((k >> 1) & 3)
  // puts the 16 lanes of a half-warp (4 k rows x 4 columns) in 16 different bank pairs.  K2
  // ([column][k], 16-byte fragment reads): for 16-wide k chunks the 16-byte pieces of odd columns
  // are stored XOR 4, so that the 8 lanes of a quarter-warp (2 columns x 4 pieces) cover all 8 bank
  // groups.
  static constexpr int PB = KCONTIG ? BK : BN;
  static constexpr int SWZ = (KCONTIG && BK == 16) ? 4 : 0;     // piece-index XOR of odd columns
  static constexpr int B_TILE = KCONTIG ? BN * PB : BK * PB;
  static constexpr int STAGE = A_STAGE + 2 * B_TILE;                // field tile + mirrored field tile
  static constexpr size_t SMEM = size_t(STAGES) * STAGE * sizeof(double);
};

// Fragments of one k-panel PAIR (8 consecutive k).  The operator is packed with the 8 k of a pair
// contiguous per row, so ONE 16-byte load gives a lane its A elements of both panels of the pair:
// the first panel takes k = 8P + 2t, the second k = 8P + 2t + 1 (any split of the 8 k into two
// groups of 4 is a valid pair of DMMA k-panels); the field fragments follow the same assignment.
template <class C, int RT, int WR, int WN, bool KCONTIG, bool VEC16>
__device__ __forceinline__ void eo_frags(const double* As, const double* B1, const double* B2, int pr, int swz,
                                         const int (&bcol)[WN], double2 (&ae)[WR], double2 (&ao)[WR],
                                         double2 (&be)[WN], double2 (&bo)[WN]) {
#pragma unroll
  for (int i = 0; i < RT; ++i) {
    ae[i] = *reinterpret_cast<const double2*>(As + (pr * 2 * C::BMH + i * 8) * 8);
    ao[i] = *reinterpret_cast<const double2*>(As + (pr * 2 * C::BMH + C::BMH + i * 8) * 8);
  }
#pragma unroll
  for (int j = 0; j < WN; ++j) {
    double f1a, f1b, f2a, f2b;     // a: k = 8P + 2t, b: k = 8P + 2t + 1 (B1/B2 already hold the lane's 2t offset)
    if (KCONTIG) {
      // swz = 1 for the odd columns of a swizzled tile: their pieces 0-3 and 4-7 are exchanged
      const double2 v1 = *reinterpret_cast<const double2*>(B1 + j * 8 * C::PB + (pr ^ swz) * 8);
      const double2 v2 = *reinterpret_cast<const double2*>(B2 + j * 8 * C::PB + (pr ^ swz) * 8);
      f1a = v1.x; f1b = v1.y;
      // mirrored tile: 16-byte copies keep the memory order inside a pair, i.e. the pair is swapped
      if (VEC16) { f2a = v2.y; f2b = v2.x; } else { f2a = v2.x; f2b = v2.y; }
    } else {
      // B1/B2 point at row 2t of the tile; bcol[j] is the swizzled column of this lane
      f1a = B1[pr * 8 * C::PB + bcol[j]]; f1b = B1[(pr * 8 + 1) * C::PB + bcol[j]];
      f2a = B2[pr * 8 * C::PB + bcol[j]]; f2b = B2[(pr * 8 + 1) * C::PB + bcol[j]];
    }
    be[j] = make_double2(__dadd_rn(f1a, f2a), __dadd_rn(f1b, f2b));
    bo[j] = make_double2(__dsub_rn(f1a, f2a), __dsub_rn(f1b, f2b));
  }
}

template <class C, int RT, int WR, int WN, int BKP, bool KCONTIG, bool VEC16>
__device__ __forceinline__ void eo_compute(const double* As, const double* B1, const double* B2, int npairs, int swz,
                                           const int (&bcol)[WN], double (&acce)[WR][WN][2], double (&acco)[WR][WN][2]) {
  static_assert(BKP % 2 == 0, "k-panels are processed in pairs");
  if (npairs == BKP / 2) {
#pragma unroll
    for (int pr = 0; pr < BKP / 2; ++pr) {
      double2 ae[WR], ao[WR], be[WN], bo[WN];
      eo_frags<C, RT, WR, WN, KCONTIG, VEC16>(As, B1, B2, pr, swz, bcol, ae, ao, be, bo);
      // consecutive DMMAs share their A operand (operand reuse), as in the plain kernel
#pragma unroll
      for (int i = 0; i < RT; ++i) {
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acce[i][j][0], acce[i][j][1], ae[i].x, be[j].x);
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acco[i][j][0], acco[i][j][1], ao[i].x, bo[j].x);
      }
#pragma unroll
      for (int i = 0; i < RT; ++i) {
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acce[i][j][0], acce[i][j][1], ae[i].y, be[j].y);
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acco[i][j][0], acco[i][j][1], ao[i].y, bo[j].y);
      }
    }
  } else {
#pragma unroll 1
    for (int pr = 0; pr < npairs; ++pr) {
      double2 ae[WR], ao[WR], be[WN], bo[WN];
      eo_frags<C, RT, WR, WN, KCONTIG, VEC16>(As, B1, B2, pr, swz, bcol, ae, ao, be, bo);
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) {
          dmma(acce[i][j][0], acce[i][j][1], ae[i].x, be[j].x);
          dmma(acco[i][j][0], acco[i][j][1], ao[i].x, bo[j].x);
          dmma(acce[i][j][0], acce[i][j][1], ae[i].y, be[j].y);
          dmma(acco[i][j][0], acco[i][j][1], ao[i].y, bo[j].y);
        }
    }
  }
}

// exact row-tile dispatch (predicated-off DMMAs still occupy the tensor pipe)
template <class C, int RT, int WR, int WN, int BKP, bool KCONTIG, bool VEC16>
struct EoDispatch {
  static __device__ __forceinline__ void run(int my_rt, const double* As, const double* B1, const double* B2,
                                             int npairs, int swz, const int (&bcol)[WN], double (&acce)[WR][WN][2],
                                             double (&acco)[WR][WN][2]) {
    if (my_rt == RT) eo_compute<C, RT, WR, WN, BKP, KCONTIG, VEC16>(As, B1, B2, npairs, swz, bcol, acce, acco);
    else EoDispatch<C, RT - 1, WR, WN, BKP, KCONTIG, VEC16>::run(my_rt, As, B1, B2, npairs, swz, bcol, acce, acco);
  }
};
template <class C, int WR, int WN, int BKP, bool KCONTIG, bool VEC16>
struct EoDispatch<C, 0, WR, WN, BKP, KCONTIG, VEC16> {
  static __device__ __forceinline__ void run(int, const double*, const double*, const double*, int, int,
                                             const int (&)[WN], double (&)[WR][WN][2], double (&)[WR][WN][2]) {}
};

template <int WARPS_M, int WARPS_N, int WR, int WN, int BKP, int STAGES, bool KCONTIG, bool VEC16, int MINB, bool SCAT>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, MINB) dense_apply_eo_kernel(const EoParams p) {
  using C = EoCfg<WARPS_M, WARPS_N, WR, WN, BKP, STAGES, KCONTIG>;
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wn = warp % WARPS_N, wm = warp / WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  const int rtiles_total = p.MPH >> 3;
  const int nchunks = (p.KPH + BKP - 1) / BKP;
  const int n = p.n;

  constexpr int EPC = VEC16 ? 2 : 1;
  constexpr int CPR1 = C::BN / EPC, RPP1 = C::NT / CPR1;            // K1: chunks per k-row, k-rows per pass
  constexpr int CPR2 = C::BK / EPC, RPP2 = C::NT / CPR2;            // K2: chunks per column, columns per pass
  static_assert(KCONTIG || (C::NT % CPR1 == 0 && C::BK % RPP1 == 0), "K1 loader shape");
  static_assert(!KCONTIG || (C::NT % CPR2 == 0 && C::BN % RPP2 == 0), "K2 loader shape");
  static_assert(!KCONTIG || RPP2 % 2 == 0, "the swizzle of a thread's columns must not depend on the pass");
  constexpr int APIECES = C::BMH * 4;                                // 16-byte pieces of one panel-PAIR half (8 doubles per row)
  constexpr int ASUB = (APIECES + C::NT - 1) / C::NT;
  const int b_cc = KCONTIG ? tid % CPR2 : tid % CPR1;
  const int b_r0 = KCONTIG ? tid / CPR2 : tid / CPR1;
  const int64_t a_panel = int64_t(p.MPH) * 8;             // one panel pair of the packed operator
  const int fast_chunks = min(p.KPH / BKP, n / C::BK);   // all panels valid and all k + EPC <= n

  // The CTA is persistent: it walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ... and its loader
  // runs STAGES-1 chunks ahead of the DMMAs ACROSS tile boundaries, so the pipeline is never refilled
  // from empty and the epilogue of a tile overlaps the loads of the next one.
  // ---- loader state (tile ld_tile, next chunk ld_chunk) ---------------------------------------------
  int64_t ld_tile = blockIdx.x;
  int ld_chunk = 0;
  int ld_m0h = 0, ld_rows = 0;
  int64_t ld_n0 = 0;
  const double* b_base = p.U;                   // K1: address of (a, k = 0, b) of this thread's column chunk
  bool b_col_ok = false;
  const double *ae_ptr = p.Dpe, *ao_ptr = p.Dpo, *b1_ptr = p.U, *b2_ptr = p.U;
  int64_t b_rstep = 0, b_cstep = 0;
  unsigned a_ok = 0, a_in = 0, b_ok = 0;

  auto setup_loader = [&](int64_t tile) {
    const int m_block = int(tile % p.m_blocks);
    const int64_t n_block = tile / p.m_blocks;
    const int rt0 = m_block * p.rt_per_block;
    ld_m0h = rt0 * 8;
    ld_rows = min(p.rt_per_block, rtiles_total - rt0) * 8;
    ld_n0 = n_block * C::BN;
    ae_ptr = p.Dpe + size_t(ld_m0h) * 8 + tid * 2;
    ao_ptr = p.Dpo + size_t(ld_m0h) * 8 + tid * 2;
    a_ok = 0; a_in = 0; b_ok = 0;
#pragma unroll
    for (int sub = 0; sub < ASUB; ++sub) {
      if (((tid + sub * C::NT) >> 2) < ld_rows) a_ok |= 1u << sub;
      if (tid + sub * C::NT < APIECES) a_in |= 1u << sub;
    }
    if (!KCONTIG) {
      const int64_t nn = ld_n0 + int64_t(b_cc) * EPC;
      b_col_ok = nn < p.NN;
      b_base = p.U;
      if (b_col_ok) {
        const int64_t a = nn / p.before;
        b_base = p.U + a * p.slab + (nn - a * p.before);
      }
      b1_ptr = b_base + int64_t(b_r0) * p.before;
      b2_ptr = b_base + int64_t(n - 1 - b_r0) * p.before;
      b_rstep = int64_t(RPP1) * p.before;            // tile 2 walks the rows downwards
      b_cstep = int64_t(C::BK) * p.before;
      b_ok = b_col_ok ? ~0u : 0u;
    } else {
      b1_ptr = p.U + (ld_n0 + b_r0) * n + b_cc * EPC;
      b2_ptr = p.U + (ld_n0 + b_r0) * n + (n - EPC - b_cc * EPC);
      b_rstep = int64_t(RPP2) * n;                   // next column: both tiles move the same way
      b_cstep = C::BK;
#pragma unroll
      for (int r = 0; r < C::BN / RPP2; ++r)
        if (ld_n0 + b_r0 + r * RPP2 < p.NN) b_ok |= 1u << r;
    }
  };

  // generic (bounds-checked) loader: the chunk that crosses the end of the k range, tiny operators
  auto load_stage_edge = [&](int stage, int chunk) {
    double* As = smem + size_t(stage) * C::STAGE;
    double* B1 = As + C::A_STAGE;
    double* B2 = B1 + C::B_TILE;
    const int kp0 = chunk * BKP;
    // operator panel pairs: [pr][even rows | odd rows] x 8
#pragma unroll
    for (int pr = 0; pr < BKP / 2; ++pr)
#pragma unroll
      for (int sub = 0; sub < ASUB; ++sub) {
        const int c = tid + sub * C::NT;
        if (APIECES % C::NT == 0 || c < APIECES) {
          const bool ok = (kp0 + 2 * pr < p.KPH) && ((c >> 2) < ld_rows);
          const size_t off = (size_t(kp0 / 2 + pr) * p.MPH + ld_m0h) * 8 + c * 2;
          cp_async16(As + (pr * 2 * C::BMH) * 8 + c * 2, ok ? p.Dpe + off : p.Dpe, ok ? 16 : 0);
          cp_async16(As + (pr * 2 * C::BMH + C::BMH) * 8 + c * 2, ok ? p.Dpo + off : p.Dpo, ok ? 16 : 0);
        }
      }
    // field tiles: rows k (tile 1) and n-1-k (tile 2) for the 16 values of k of this chunk
    const int k0 = kp0 * 4;
    if (!KCONTIG) {
#pragma unroll
      for (int r = 0; r < C::BK / RPP1; ++r) {
        const int kr = b_r0 + r * RPP1;
        const int k = k0 + kr;
        const bool ok = b_col_ok && (k < n);
        const double* s1 = ok ? b_base + int64_t(k) * p.before : p.U;
        const double* s2 = ok ? b_base + int64_t(n - 1 - k) * p.before : p.U;
        const int cs = (b_cc * EPC) ^ (4 * ((kr >> 1) & 3));      // K1 swizzle (EoCfg)
        double* d1 = B1 + kr * C::PB + cs;
        double* d2 = B2 + kr * C::PB + cs;
        if (VEC16) { cp_async16(d1, s1, ok ? 16 : 0); cp_async16(d2, s2, ok ? 16 : 0); }
        else { cp_async8(d1, s1, ok ? 8 : 0); cp_async8(d2, s2, ok ? 8 : 0); }
      }
    } else {
      const int k = k0 + b_cc * EPC;            // first k of this thread's piece
#pragma unroll
      for (int r = 0; r < C::BN / RPP2; ++r) {
        const int col = b_r0 + r * RPP2;
        const int64_t nn = ld_n0 + col;
        // the piece k .. k+EPC-1 and its mirror n-EPC-k .. n-1-k lie inside the row iff k+EPC <= n
        const bool ok = (nn < p.NN) && (k + EPC <= n);
        const double* row = p.U + nn * n;
        const int sw = (col & 1) ? C::SWZ : 0;
        const int slot = VEC16 ? ((b_cc ^ sw) * 2) : ((((b_cc >> 1) ^ sw) << 1) | (b_cc & 1));
        double* d1 = B1 + col * C::PB + slot;
        double* d2 = B2 + col * C::PB + slot;
        if (VEC16) {
          cp_async16(d1, ok ? row + k : p.U, ok ? 16 : 0);
          cp_async16(d2, ok ? row + (n - 2 - k) : p.U, ok ? 16 : 0);
        } else {
          cp_async8(d1, ok ? row + k : p.U, ok ? 8 : 0);
          cp_async8(d2, ok ? row + (n - 1 - k) : p.U, ok ? 8 : 0);
        }
      }
    }
  };

  // fast loader for chunks whose panels are all valid: running pointers, no index arithmetic in
  // the loop (the chunks of a tile are loaded in increasing order)
  auto load_stage = [&](int stage, int chunk) {
    if (chunk >= fast_chunks) { load_stage_edge(stage, chunk); return; }
    double* As = smem + size_t(stage) * C::STAGE;
    double* B1 = As + C::A_STAGE;
    double* B2 = B1 + C::B_TILE;
#pragma unroll
    for (int pr = 0; pr < BKP / 2; ++pr)
#pragma unroll
      for (int sub = 0; sub < ASUB; ++sub) {
        const bool ok = (a_ok >> sub) & 1u;
        if (APIECES % C::NT == 0 || ((a_in >> sub) & 1u)) {
          const int c2 = (tid + sub * C::NT) * 2;
          cp_async16(As + (pr * 2 * C::BMH) * 8 + c2, ok ? ae_ptr + pr * a_panel + sub * C::NT * 2 : p.Dpe, ok ? 16 : 0);
          cp_async16(As + (pr * 2 * C::BMH + C::BMH) * 8 + c2, ok ? ao_ptr + pr * a_panel + sub * C::NT * 2 : p.Dpo, ok ? 16 : 0);
        }
      }
    ae_ptr += (BKP / 2) * a_panel;
    ao_ptr += (BKP / 2) * a_panel;
    constexpr int BR = KCONTIG ? C::BN / RPP2 : C::BK / RPP1;
    constexpr int RPP = KCONTIG ? RPP2 : RPP1;
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      const bool ok = (b_ok >> (KCONTIG ? r : 0)) & 1u;
      int so = (b_r0 + r * RPP) * C::PB + b_cc * EPC;
      if (!KCONTIG) so = (b_r0 + r * RPP) * C::PB + ((b_cc * EPC) ^ (4 * (((b_r0 + r * RPP) >> 1) & 3)));
      if (KCONTIG && C::SWZ) {          // column parity is constant over r (RPP is even)
        const int sw = (b_r0 & 1) ? C::SWZ : 0;
        so = (b_r0 + r * RPP) * C::PB + (VEC16 ? ((b_cc ^ sw) * 2) : ((((b_cc >> 1) ^ sw) << 1) | (b_cc & 1)));
      }
      const double* s1 = ok ? b1_ptr + r * b_rstep : p.U;
      const double* s2 = ok ? (KCONTIG ? b2_ptr + r * b_rstep : b2_ptr - r * b_rstep) : p.U;
      if (VEC16) { cp_async16(B1 + so, s1, ok ? 16 : 0); cp_async16(B2 + so, s2, ok ? 16 : 0); }
      else { cp_async8(B1 + so, s1, ok ? 8 : 0); cp_async8(B2 + so, s2, ok ? 8 : 0); }
    }
    b1_ptr += b_cstep;
    b2_ptr -= b_cstep;
  };

  // issue the next chunk of the flattened (tile, chunk) sequence into ring slot `slot`
  auto issue_next = [&](int slot) {
    if (ld_tile < p.total_tiles) {
      load_stage(slot, ld_chunk);
      if (++ld_chunk == nchunks) {
        ld_chunk = 0;
        ld_tile += gridDim.x;
        if (ld_tile < p.total_tiles) setup_loader(ld_tile);
      }
    }
    cp_async_commit();
  };

  if (ld_tile < p.total_tiles) setup_loader(ld_tile);
  unsigned issued = 0, consumed = 0;            // ring positions
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) { issue_next(issued % STAGES); ++issued; }

  // B fragment: k = t, column = g.  In the mirrored tile of the contiguous-axis kernel a 16-byte
  // piece holds (u[n-2-k], u[n-1-k]): the slot of mirrored index k is k ^ 1.
  // (the pair swap of the mirrored tile is undone when the fragments are read, eo_frags)
  const int b1_off = KCONTIG ? ((wn * WN * 8 + g) * C::PB + 2 * t) : (2 * t * C::PB);
  const int b2_off = b1_off;
  int bcol[WN];                                       // K1: swizzled column of this lane in column tile j
#pragma unroll
  for (int j = 0; j < WN; ++j) bcol[j] = (wn * WN * 8 + j * 8 + g) ^ (4 * t);
  const int frag_swz = (C::SWZ && (g & 1)) ? 1 : 0;   // odd columns of a swizzled tile: pair index XOR 1
  const int nfull = p.KPH / BKP;
  const int tail = (p.KPH - nfull * BKP) / 2;        // panel pairs of the last, partial chunk

  for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int m_block = int(tile % p.m_blocks);
    const int64_t n_block = tile / p.m_blocks;
    const int rt0 = m_block * p.rt_per_block;
    const int rt_count = min(p.rt_per_block, rtiles_total - rt0);
    const int m0h = rt0 * 8;                      // first half-row of the tile
    const int64_t n0 = n_block * C::BN;
    const int q = (rt_count + WARPS_M - 1) / WARPS_M;
    const int my_rt0 = wm * q;
    const int my_rt = max(0, min(q, rt_count - my_rt0));
    const int a_off = my_rt0 * 64 + 2 * lane;        // 8 rows x 8 k per row-tile and panel pair; lane (g, t) -> row g, k 2t

    double acce[WR][WN][2], acco[WR][WN][2];
#pragma unroll
    for (int i = 0; i < WR; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) { acce[i][j][0] = acce[i][j][1] = 0.0; acco[i][j][0] = acco[i][j][1] = 0.0; }

    for (int kc = 0; kc < nchunks; ++kc) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      issue_next(issued % STAGES);
      ++issued;
      const double* As = smem + size_t(consumed % STAGES) * C::STAGE;
      const double* B1 = As + C::A_STAGE;
      const double* B2 = B1 + C::B_TILE;
      ++consumed;
      EoDispatch<C, WR, WR, WN, BKP, KCONTIG, VEC16>::run(my_rt, As + a_off, B1 + b1_off, B2 + b2_off,
                                                          kc < nfull ? BKP / 2 : tail, frag_swz, bcol, acce, acco);
    }

    // ---- epilogue: rows i (ze + zo) and n-1-i (s*(ze - zo)) --------------------------------------------
    eo_epilogue<WR, WN, KCONTIG, VEC16, SCAT, false>(p, acce, acco, my_rt, m0h + my_rt0 * 8, n0 + wn * WN * 8, g, t);
  }
  cp_async_wait<0>();
}

// Ae / Ao of a centrosymmetric operator in the fragment-major layout of the apply kernel
__global__ void pack_eo_kernel(const double* __restrict__ D, int n, int H, int MPH, int KPH,
                               double* __restrict__ pe, double* __restrict__ po) {
  // [KPH/2 panel pairs][MPH rows][8 consecutive k]
  const int64_t total = int64_t(KPH / 2) * MPH * 8;
  const int mid = (n & 1) ? (n - 1) / 2 : -1;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int kk = int(idx & 7);
    const int64_t r = idx >> 3;
    const int i = int(r % MPH), k = int(r / MPH) * 8 + kk;
    double ve = 0.0, vo = 0.0;
    if (i < H && k < H) {
      const double a = D[int64_t(i) * n + k], b = D[int64_t(i) * n + (n - 1 - k)];
      if (k == mid) { ve = __dmul_rn(0.5, a); vo = 0.0; }      // E[mid] = 2 u[mid], O[mid] = 0
      else { ve = __dmul_rn(0.5, __dadd_rn(a, b)); vo = __dmul_rn(0.5, __dsub_rn(a, b)); }
    }
    pe[idx] = ve;
    po[idx] = vo;
  }
}

// largest entrywise violation of J D J = s D, relative to the entry (atomicMax on the bit pattern of
// a non-negative double)
__global__ void symmetry_kernel(const double* __restrict__ D, int n, int sign, unsigned long long* worst) {
  const int64_t total = int64_t(n) * n;
  double w = 0.0;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int i = int(idx / n), j = int(idx % n);
    const double a = D[idx], b = double(sign) * D[int64_t(n - 1 - i) * n + (n - 1 - j)];
    const double d = fabs(a - b), m = fmax(fabs(a), fabs(b));
    if (d > 0.0) w = fmax(w, d / m);
  }
  atomicMax(worst, (unsigned long long)__double_as_longlong(w));
}

inline bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

int g_eo_persist = 1;   // tuning hook: pt_debug_set_eo_config(50 / 51) = one CTA per tile / persistent

template <int WARPS_M, int WARPS_N, int WR, int WN, int BKP, int STAGES, bool KCONTIG, bool VEC16, int MINB, bool SCAT>
int launch_eo(EoParams& p, cudaStream_t stream) {
  using C = EoCfg<WARPS_M, WARPS_N, WR, WN, BKP, STAGES, KCONTIG>;
  auto kern = dense_apply_eo_kernel<WARPS_M, WARPS_N, WR, WN, BKP, STAGES, KCONTIG, VEC16, MINB, SCAT>;
  static bool attr_set[64] = {false};
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(C::SMEM)));
    attr_set[dev] = true;
  }
  const int rtiles = p.MPH / 8;
  const int max_rt = WARPS_M * WR;
  p.m_blocks = (rtiles + max_rt - 1) / max_rt;
  p.rt_per_block = (rtiles + p.m_blocks - 1) / p.m_blocks;
  p.m_blocks = (rtiles + p.rt_per_block - 1) / p.rt_per_block;
  const int64_t n_blocks = (p.NN + C::BN - 1) / C::BN;
  p.total_tiles = n_blocks * p.m_blocks;
  if (p.total_tiles <= 0) return PT_OK;
  // persistent CTAs: one per resident slot (g_eo_persist = 0: one CTA per tile, for A/B timing)
  int64_t grid = int64_t(pt::sm_count()) * MINB;
  if (!g_eo_persist || grid > p.total_tiles) grid = p.total_tiles;
  if (grid > 2147483647LL) return pt::fail(PT_ERR_UNSUPPORTED, "pt_dense_apply_eo: grid too large");
  kern<<<unsigned(grid), C::NT, C::SMEM, stream>>>(p);
  return pt::check_launch("pt_dense_apply_eo");
}

int g_eo_config = -1;   // tuning hook (pt_debug_set_eo_config); -1 = choose by shape

template <bool KCONTIG, bool VEC16, bool SCAT>
int launch_eo_shape(EoParams& p, cudaStream_t stream) {
  // CTAs of the large tile; below one per SM the small tile fills the GPU better
  const int64_t ctas = ((p.H + 63) / 64) * ((p.NN + 63) / 64);
  // measured on B200 (tools/perf_probe.py eocfg): the 1 x 4 warp layout (half as many E/O adds per
  // DMMA) is 1-6 % faster than 2 x 2
  // half-rows that split into <= 6 row-tiles per tile (n = 257: 17 row-tiles = 6 + 6 + 5) fit the
  // 48-row tile with 3 CTAs/SM better than the 64-row one (measured 0.206 vs 0.222 ms)
  const int rtiles = (p.H + 7) / 8, mb = (rtiles + 7) / 8, rt_bal = (rtiles + mb - 1) / mb;
  const int cfg = g_eo_config >= 0 ? g_eo_config : (ctas < pt::sm_count() ? 1 : (rt_bal <= 6 ? 3 : 2));
  if (cfg == 1) return launch_eo<2, 2, 2, 2, 4, 3, KCONTIG, VEC16, 4, SCAT>(p, stream);   // 32 half-rows x 32 columns
  if (cfg == 2) return launch_eo<1, 4, 8, 2, 4, 3, KCONTIG, VEC16, 2, SCAT>(p, stream);   // 64 x 64, warps side by side
  if (cfg == 3) return launch_eo<1, 4, 6, 2, 2, 3, KCONTIG, VEC16, 3, SCAT>(p, stream);   // 48 x 64, 3 CTAs/SM, 8-wide k chunks
  return launch_eo<2, 2, 4, 4, 4, 3, KCONTIG, VEC16, 2, SCAT>(p, stream);                 // 64 half-rows x 64 columns
}

int apply_eo_impl(const double* pe, const double* po, int n, int sign, const double* U, double* Y, const double* C,
                  int64_t after, int64_t before, double alpha, double beta, const double* scale, int scale_mode,
                  int64_t scale_len, const pt_scatter_map* map, pt_stream_t stream) {
  PT_REQUIRE(pe && po && U, "pt_dense_apply_eo: null pointer");
  PT_REQUIRE(n >= 2, "pt_dense_apply_eo: n=%d", n);
  PT_REQUIRE(sign == 1 || sign == -1, "pt_dense_apply_eo: sign must be +1 or -1");
  PT_REQUIRE(after >= 0 && before >= 1, "pt_dense_apply_eo: bad field shape");
  PT_REQUIRE(scale_mode >= PT_SCALE_NONE && scale_mode <= PT_SCALE_BEFORE, "pt_dense_apply_eo: bad scale_mode %d", scale_mode);
  PT_REQUIRE(scale_mode == PT_SCALE_NONE || scale != nullptr, "pt_dense_apply_eo: scale is null");
  PT_REQUIRE(scale_mode != PT_SCALE_AFTER || scale_len > 0, "pt_dense_apply_eo: scale_len must be > 0");
  if (after == 0) return PT_OK;
  EoParams p;
  p.Dpe = pe; p.Dpo = po; p.U = U; p.Y = Y; p.C = C; p.scale = scale;
  p.sc = map ? ScatterDev{map->base, map->chunk, map->a_inner, map->s_outer, map->s_inner, map->s_i, map->offset}
             : ScatterDev{nullptr, 1, 1, 0, 0, 0, 0};
  p.n = n; p.H = (n + 1) / 2;
  p.MPH = 8 * ((p.H + 7) / 8);
  p.KPH = 2 * ((p.H + 7) / 8);           // k-panels, padded to whole pairs
  p.sign = sign;
  p.before = before; p.NN = after * before; p.slab = int64_t(n) * before;
  p.alpha = alpha; p.beta = beta; p.scale_mode = scale_mode; p.scale_len = scale_len > 0 ? scale_len : 1;
  p.m_blocks = 1; p.rt_per_block = 1; p.total_tiles = 0;
  p.cmid = nullptr; p.rmid = nullptr; p.dmm = 0.0; p.mid = -1; p.dbg = 0; p.trace = nullptr;
  p.cp = ScatterDev{nullptr, 1, 1, 0, 0, 0, 0};
  cudaStream_t st = pt::as_stream(stream);
  if (before == 1) {
    const bool vec = (n % 2 == 0) && aligned16(U);
    if (map) return vec ? launch_eo_shape<true, true, true>(p, st) : launch_eo_shape<true, false, true>(p, st);
    return vec ? launch_eo_shape<true, true, false>(p, st) : launch_eo_shape<true, false, false>(p, st);
  }
  if (map) {
    const bool even = !((map->s_outer | map->s_inner | map->s_i | map->offset) & 1);
    const bool vec = (before % 2 == 0) && aligned16(U) && (C == nullptr || aligned16(C)) && even;
    return vec ? launch_eo_shape<false, true, true>(p, st) : launch_eo_shape<false, false, true>(p, st);
  }
  const bool vec = (before % 2 == 0) && aligned16(U) && aligned16(Y);
  return vec ? launch_eo_shape<false, true, false>(p, st) : launch_eo_shape<false, false, false>(p, st);
}

}  // namespace

extern "C" int pt_debug_set_eo_config(int id) {
  if (id == 50 || id == 51) g_eo_persist = id - 50;
  else g_eo_config = id;
  return PT_OK;
}

extern "C" int64_t pt_packed_eo_size(int n) {
  if (n <= 0) return 0;
  const int64_t H = (int64_t(n) + 1) / 2;
  return ((H + 7) / 8) * (8 * ((H + 7) / 8)) * 8;
}

extern "C" int pt_operator_symmetry(const double* D, int n, int sign, double* worst_dev, pt_stream_t stream) {
  PT_REQUIRE(D && worst_dev, "pt_operator_symmetry: null pointer");
  PT_REQUIRE(n >= 1 && (sign == 1 || sign == -1), "pt_operator_symmetry: bad argument");
  // stream-ordered, nothing allocated, no synchronisation: the caller owns the device scalar (the
  // bit pattern of a non-negative double orders like the unsigned integer the kernel takes the
  // maximum of) and reads it back when it wants the verdict
  cudaStream_t st = pt::as_stream(stream);
  PT_CUDA(cudaMemsetAsync(worst_dev, 0, sizeof(double), st));
  int64_t blocks = (int64_t(n) * n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  symmetry_kernel<<<unsigned(blocks), 256, 0, st>>>(D, n, sign, reinterpret_cast<unsigned long long*>(worst_dev));
  return pt::check_launch("pt_operator_symmetry");
}

extern "C" int pt_operator_pack_eo(const double* D, int n, double* packed_even, double* packed_odd,
                                   pt_stream_t stream) {
  PT_REQUIRE(D && packed_even && packed_odd, "pt_operator_pack_eo: null pointer");
  PT_REQUIRE(n >= 2, "pt_operator_pack_eo: n=%d", n);
  const int H = (n + 1) / 2, MPH = 8 * ((H + 7) / 8), KPH = 2 * ((H + 7) / 8);
  const int64_t total = int64_t(KPH / 2) * MPH * 8;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_eo_kernel<<<unsigned(blocks), 256, 0, pt::as_stream(stream)>>>(D, n, H, MPH, KPH, packed_even, packed_odd);
  return pt::check_launch("pt_operator_pack_eo");
}

extern "C" int pt_dense_apply_eo(const double* packed_even, const double* packed_odd, int n, int sign,
                                 const double* U, double* Y, int64_t after, int64_t before, double alpha,
                                 double beta, const double* scale, int scale_mode, int64_t scale_len,
                                 pt_stream_t stream) {
  PT_REQUIRE(Y != nullptr, "pt_dense_apply_eo: null pointer");
  PT_REQUIRE(U != Y, "pt_dense_apply_eo: in-place application is not supported");
  return apply_eo_impl(packed_even, packed_odd, n, sign, U, Y, Y, after, before, alpha, beta, scale, scale_mode,
                       scale_len, nullptr, stream);
}

extern "C" int pt_dense_apply_eo_scatter(const double* packed_even, const double* packed_odd, int n, int sign,
                                         const double* U, const double* C, int64_t after, int64_t before,
                                         double alpha, double beta, const pt_scatter_map* map, pt_stream_t stream) {
  PT_REQUIRE(map && map->base, "pt_dense_apply_eo_scatter: null map");
  PT_REQUIRE(map->nranks >= 1 && map->chunk >= 1 && int64_t(map->chunk) * map->nranks >= n,
             "pt_dense_apply_eo_scatter: %d destinations x %d rows do not cover n=%d", map->nranks, map->chunk, n);
  PT_REQUIRE(map->a_inner >= 1, "pt_dense_apply_eo_scatter: a_inner must be >= 1");
  PT_REQUIRE(beta == 0.0 || C != nullptr, "pt_dense_apply_eo_scatter: beta != 0 needs C");
  return apply_eo_impl(packed_even, packed_odd, n, sign, U, nullptr, C ? C : U, after, before, alpha, beta, nullptr,
                       PT_SCALE_NONE, 0, map, stream);
}
