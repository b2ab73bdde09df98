// K1/K2, persistent variant — FP64 tensor-core kernel whose operator panels are staged by the
// bulk-copy engine (cp.async.bulk, SASS UBLKCP), whose field tiles are staged by cp.async, and
// whose pipeline is tracked entirely with mbarriers.
//
// Same math, tiles and fragment layouts as dense_apply.cu; what changes is the data movement and
// the synchronisation:
//   * operator panels (packed, contiguous, 3-4 KB each) are moved by 1-D bulk copies issued by one
//     thread; field tiles keep the coalesced 16-byte cp.async of every thread (many 128-512 byte
//     bulk copies measured much slower: the copy engine serves roughly one request per ~46 cycles);
//   * a ring of STAGES slots with one "full" mbarrier (transaction bytes of the bulk copies plus
//     one cp.async.mbarrier.arrive.noinc per thread) and one "empty" mbarrier (one arrival per
//     warp) each replaces cp.async.wait_group + __syncthreads: a warp only ever waits for data
//     or for the slot it is about to refill;
//   * the CTA is persistent: it walks over its tiles (m fastest, so CTAs that share field columns
//     run together) and the copy stream runs ahead across tile boundaries, so the epilogue of one
//     tile overlaps the loads of the next and there is no wave quantisation.
// Requirements (else pt_dense_apply falls back to the cp.async kernel): 16-byte aligned fields and
// an even contiguous extent (the 16-byte staging path).
#include "dense_common.cuh"

using namespace ptdense;

namespace {

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <int WARPS_M, int WARPS_N, int WM, int WN, int BKP, int STAGES, bool KCONTIG, int MINB>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, MINB)
dense_apply_tma_kernel(const DenseParams p, int64_t tiles_total) {
  using C = Cfg<WARPS_M, WARPS_N, WM, WN, BKP, STAGES, KCONTIG, true>;
  extern __shared__ __align__(128) double smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * C::STAGE);
  uint64_t* empty = full + STAGES;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wn = warp % WARPS_N;
  const int wm = warp / WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  const int mtiles_total = p.MP >> 3;
  const int nchunks = (p.KP + BKP - 1) / BKP;
  const int nfull = p.KP / BKP;
  const int tail = p.KP - nfull * BKP;

  // ---- one-time setup: barriers, zeroed field tiles (k/n edges read zeros, never garbage) -------
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, C::NT + 1); mbar_init(empty + s, WARPS_M * WARPS_N); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int s = 0; s < STAGES; ++s)
    for (int i = tid; i < C::B_STAGE; i += C::NT) smem[size_t(s) * C::STAGE + C::A_STAGE + i] = 0.0;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();

  // ---- producer side (every thread): stream of chunks over this CTA's tiles ----------------------
  constexpr int CPR1 = C::BN / 2, RPP1 = C::NT / CPR1;   // K1: 16-byte chunks per k-row, k-rows per pass
  constexpr int CPR2 = C::BK / 2, RPP2 = C::NT / CPR2;   // K2: chunks per column, columns per pass
  constexpr int BR = KCONTIG ? C::BN / RPP2 : C::BK / RPP1;
  constexpr int RPP = KCONTIG ? RPP2 : RPP1;
  const int b_cc = KCONTIG ? tid % CPR2 : tid % CPR1;
  const int b_r0 = KCONTIG ? tid / CPR2 : tid / CPR1;
  const int64_t b_rstep = KCONTIG ? int64_t(RPP2) * p.n_in : int64_t(RPP1) * p.before;
  const int64_t b_cstep = KCONTIG ? int64_t(C::BK) : int64_t(C::BK) * p.before;

  int64_t p_tile = blockIdx.x;   // tile whose chunk is issued next
  int p_kc = 0;
  unsigned p_g = 0;              // chunks issued so far
  const double* p_bptr = p.U;    // this thread's source of the first pass of the current chunk
  const double* p_bbase = p.U;   // ... of chunk 0 of the current tile (edge chunks recompute from it)
  unsigned p_ok = 0;             // per-pass validity (K2) or column validity (K1)
  int p_mt0 = 0, p_rows = 0;

  auto issue_next = [&]() {
    const int slot = p_g % STAGES;
    if (p_kc == 0) {             // new tile: per-thread bases
      const int m_block = int(p_tile % p.m_blocks);
      const int64_t n_block = p_tile / p.m_blocks;
      p_mt0 = m_block * p.mt_per_block;
      p_rows = min(p.mt_per_block, mtiles_total - p_mt0) * 8;
      const int64_t n0 = n_block * C::BN;
      p_ok = 0;
      if (!KCONTIG) {
        const int64_t nn = n0 + int64_t(b_cc) * 2;
        if (nn < p.NN) {
          const int64_t a = nn / p.before;
          p_ok = ~0u;
          p_bbase = p.U + a * p.slab_in + (nn - a * p.before) + int64_t(b_r0) * p.before;
        } else {
          p_bbase = p.U;
        }
      } else {
        p_bbase = p.U + (n0 + b_r0) * p.n_in + b_cc * 2;
#pragma unroll
        for (int r = 0; r < BR; ++r)
          if (n0 + b_r0 + r * RPP2 < p.NN) p_ok |= 1u << r;
      }
      p_bptr = p_bbase;
    }
    double* As = smem + size_t(slot) * C::STAGE;
    double* Bs = As + C::A_STAGE;
    const int kp0 = p_kc * BKP;
    const int k0 = kp0 * 4;
    const bool inside = (k0 + C::BK <= p.n_in);
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      bool ok = (p_ok >> (KCONTIG ? r : 0)) & 1u;
      if (!inside) ok = ok && (KCONTIG ? (k0 + b_cc * 2 < p.n_in) : (k0 + b_r0 + r * RPP < p.n_in));
      double* dst = Bs + (b_r0 + r * RPP) * C::PB + b_cc * 2;
      cp_async16(dst, ok ? p_bptr + r * b_rstep : p.U, ok ? 16 : 0);
    }
    p_bptr += b_cstep;
    cp_async_arrive_noinc(full + slot);
    if (tid == 0) {              // operator panels: one bulk copy each
      const int npan = min(BKP, p.KP - kp0);
      const unsigned a_bytes = unsigned(p_rows) * 32u;
      mbar_expect_tx(full + slot, unsigned(npan) * a_bytes);
      for (int pp = 0; pp < npan; ++pp)
        bulk_g2s(As + pp * C::BM * 4, p.Dp + (size_t(kp0 + pp) * p.MP + p_mt0 * 8) * 4, a_bytes, full + slot);
    }
    ++p_g;
    if (++p_kc == nchunks) { p_kc = 0; p_tile += gridDim.x; }
  };

  for (int s = 0; s < STAGES && p_tile < tiles_total; ++s) issue_next();

  // ---- consumers (all warps) -------------------------------------------------------------------------
  unsigned cg = 0;  // chunks consumed so far
  for (int64_t tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
    const int m_block = int(tile % p.m_blocks);
    const int64_t n_block = tile / p.m_blocks;
    const int mt0 = m_block * p.mt_per_block;
    const int mt_count = min(p.mt_per_block, mtiles_total - mt0);
    const int m0 = mt0 * 8;
    const int64_t n0 = n_block * C::BN;
    const int q = (mt_count + WARPS_M - 1) / WARPS_M;
    const int my_mt0 = wm * q;
    const int my_mt = max(0, min(q, mt_count - my_mt0));
    const int a_off = my_mt0 * 32 + lane;
    const int b_off = KCONTIG ? ((wn * WN * 8 + g) * C::PB + t) : (t * C::PB + wn * WN * 8 + g);

    double acc[WM][WN][2];
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    for (int kc = 0; kc < nchunks; ++kc, ++cg) {
      const int slot = cg % STAGES;
      mbar_wait(full + slot, (cg / STAGES) & 1u);
      const double* As = smem + size_t(slot) * C::STAGE + a_off;
      const double* Bs = smem + size_t(slot) * C::STAGE + C::A_STAGE + b_off;
      if (kc < nfull) MtDispatch<C, WM, WM, WN, BKP, KCONTIG>::full(my_mt, As, Bs, acc);
      else MtDispatch<C, WM, WM, WN, BKP, KCONTIG>::partial(my_mt, As, Bs, acc, tail);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
      // producer duty: refill the slot that every warp released one chunk ago (one chunk of slack
      // keeps a fast warp from waiting on the slowest warp of the current chunk)
      if (cg >= 1 && p_tile < tiles_total) {
        const unsigned prev = cg - 1;
        mbar_wait(empty + prev % STAGES, (prev / STAGES) & 1u);
        issue_next();
      }
    }

    // ---- epilogue (same as the cp.async kernel) -----------------------------------------------------
    const bool use_beta = (p.beta != 0.0);
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      const int64_t nn = n0 + (wn * WN + j) * 8 + 2 * t;
      if (nn >= p.NN) continue;
      const bool two = (nn + 1 < p.NN);
      int64_t a0, b0, a1, b1;
      if (KCONTIG) {
        a0 = nn; b0 = 0; a1 = nn + 1; b1 = 0;
      } else {
        a0 = nn / p.before; b0 = nn - a0 * p.before; a1 = a0; b1 = b0 + 1;   // before even, nn even
      }
      double s0 = p.alpha, s1 = p.alpha;
      if (p.scale_mode == PT_SCALE_AFTER) {
        s0 *= p.scale[a0 % p.scale_len];
        if (two) s1 *= p.scale[a1 % p.scale_len];
      } else if (p.scale_mode == PT_SCALE_BEFORE) {
        s0 *= p.scale[b0];
        if (two) s1 *= p.scale[b1];
      }
      double* y0 = p.Y + a0 * p.slab_out + b0;
      double* y1 = p.Y + a1 * p.slab_out + b1;
#pragma unroll
      for (int i = 0; i < WM; ++i) {
        if (i >= my_mt) continue;
        const int row = m0 + (my_mt0 + i) * 8 + g;
        if (row >= p.n_out) continue;
        double r0 = s0, r1 = s1;
        if (p.scale_mode == PT_SCALE_ROW) { const double sr = p.scale[row]; r0 *= sr; r1 *= sr; }
        double v0 = __dmul_rn(r0, acc[i][j][0]);   // explicit roundings: every kernel variant gives the same bits
        double v1 = __dmul_rn(r1, acc[i][j][1]);
        const int64_t off = int64_t(row) * p.before;
        if (!KCONTIG) {
          double2* dst = reinterpret_cast<double2*>(y0 + off);
          if (use_beta) { const double2 o = *dst; v0 = __fma_rn(p.beta, o.x, v0); v1 = __fma_rn(p.beta, o.y, v1); }
          *dst = make_double2(v0, v1);
        } else {
          if (use_beta) v0 = __fma_rn(p.beta, y0[off], v0);
          y0[off] = v0;
          if (two) {
            if (use_beta) v1 = __fma_rn(p.beta, y1[off], v1);
            y1[off] = v1;
          }
        }
      }
    }
  }
}

template <int WARPS_M, int WARPS_N, int WM, int WN, int BKP, int STAGES, bool KCONTIG, int MINB>
int launch_tma_cfg(DenseParams& p, cudaStream_t stream) {
  using C = Cfg<WARPS_M, WARPS_N, WM, WN, BKP, STAGES, KCONTIG, true>;
  auto kern = dense_apply_tma_kernel<WARPS_M, WARPS_N, WM, WN, BKP, STAGES, KCONTIG, MINB>;
  constexpr size_t SMEM = C::SMEM + 2 * STAGES * sizeof(uint64_t);
  static int blocks_per_sm[64] = {0};
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return pt::fail(PT_ERR_UNSUPPORTED, "device index");
  if (blocks_per_sm[dev] == 0) {
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
    int nb = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::NT, SMEM));
    blocks_per_sm[dev] = nb > 0 ? nb : 1;
  }
  const int mtiles = p.MP / 8;
  const int max_mt = WARPS_M * WM;
  p.m_blocks = (mtiles + max_mt - 1) / max_mt;
  p.mt_per_block = (mtiles + p.m_blocks - 1) / p.m_blocks;
  p.m_blocks = (mtiles + p.mt_per_block - 1) / p.mt_per_block;
  const int64_t n_blocks = (p.NN + C::BN - 1) / C::BN;
  const int64_t tiles = n_blocks * p.m_blocks;
  if (tiles <= 0) return PT_OK;
  int64_t grid = int64_t(pt::sm_count()) * blocks_per_sm[dev];
  if (grid > tiles) grid = tiles;
  kern<<<unsigned(grid), C::NT, SMEM, stream>>>(p, tiles);
  return pt::check_launch("pt_dense_apply(tma)");
}

}  // namespace

namespace ptdense {

// Returns PT_ERR_UNSUPPORTED (without setting an error message) when the shape needs the
// cp.async kernel.
int launch_tma(DenseParams& p, bool kcontig, int cfg, cudaStream_t stream) {
  if (cfg == 5) {
    return kcontig ? launch_tma_cfg<2, 2, 8, 4, 4, 4, true, 2>(p, stream)
                   : launch_tma_cfg<2, 2, 8, 4, 4, 4, false, 2>(p, stream);
  }
  return kcontig ? launch_tma_cfg<1, 4, 12, 2, 4, 3, true, 3>(p, stream)
                 : launch_tma_cfg<1, 4, 12, 2, 4, 3, false, 3>(p, stream);
}

}  // namespace ptdense
