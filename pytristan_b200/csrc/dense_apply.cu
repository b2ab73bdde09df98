// K1/K2 — dense differentiation-operator application on FP64 tensor cores (DMMA m8n8k4).
//
//   Y[a, :, b] = alpha * scale * (D @ U[a, :, b]) + beta * Y[a, :, b]
//
// U is (after, n_in, before) C-contiguous, i.e. a field in the reference Grid's linearisation
// (axis 0 fastest, reference grid.py:85-88) seen from the differentiated axis.  The operator D
// is always the A operand of the MMA (M = output index i along the axis); the field supplies
// the B operand with its columns being the flattened (a, b) index space:
//   K1 (before > 1): column nn = a*before + b, k-stride = before   -> smem tile [k][n]
//   K2 (before == 1): column nn = a,            k-stride = 1        -> smem tile [n][k]
// Neither case transposes U or Y in global memory.
//
// D is read from the packed fragment-major copy [k/4][8*ceil(n_out/8)][4] (pt_operator_pack):
// an (8-row, 4-k) block is 256 contiguous bytes, lane l of a warp reads double l, which is
// exactly the m8n8k4 A fragment (row = l/4, k = l%4) — conflict-free without padding.
// Field tiles are staged with cp.async (16-byte when alignment allows, else 8-byte) into a
// padded layout (pitch = 4 mod 16 doubles) that makes the B-fragment reads conflict-free.
//
// FP64 tensor cores on sm_100a are warp-level only (tcgen05 has no f64 kind), so accumulators
// live in registers; see DESIGN.md §4.1 for the roofline this kernel is judged against.
#include "dense_common.cuh"

using namespace ptdense;

namespace {

__device__ __forceinline__ unsigned smem_u32(const void* ptr) {
  return static_cast<unsigned>(__cvta_generic_to_shared(ptr));
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

// ATMA: the operator panels of a chunk (packed, contiguous: rows*32 bytes each) are staged by the
// bulk-copy engine (cp.async.bulk, SASS UBLKCP) issued by one thread and tracked by one mbarrier
// per stage; the field tile keeps the per-thread 16-byte cp.async.
// SCAT: the epilogue stores each output element straight into the buffer of the rank that owns it
// after the pencil redistribution (peer memory over NVLink) instead of into a local Y.
template <int WARPS_M, int WARPS_N, int WM, int WN, int BKP, int STAGES, bool KCONTIG, bool VEC16, int MINB, bool ATMA,
          bool SCAT>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, MINB)
dense_apply_kernel(const DenseParams p) {
  using C = Cfg<WARPS_M, WARPS_N, WM, WN, BKP, STAGES, KCONTIG, VEC16>;
  extern __shared__ __align__(128) double smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * C::STAGE);
  if (ATMA) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(full_bar + s)));
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
  }

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wn = warp % WARPS_N;
  const int wm = warp / WARPS_N;
  const int g = lane >> 2;  // fragment row (A) / column (B)
  const int t = lane & 3;   // fragment k index

  // block -> (m_block fastest, n_block): CTAs sharing a field tile run together (L2 reuse of U)
  const int m_block = blockIdx.x % p.m_blocks;
  const int64_t n_block = blockIdx.x / p.m_blocks;
  const int mtiles_total = p.MP >> 3;
  const int mt0 = m_block * p.mt_per_block;
  const int mt_count = min(p.mt_per_block, mtiles_total - mt0);
  const int m0 = mt0 * 8;
  const int rows_blk = mt_count * 8;
  const int64_t n0 = n_block * C::BN;

  // this warp's m-tiles: balanced split of the CTA's mt_count tiles over WARPS_M warps
  const int q = (mt_count + WARPS_M - 1) / WARPS_M;
  const int my_mt0 = wm * q;
  const int my_mt = max(0, min(q, mt_count - my_mt0));

  const int nchunks = (p.KP + BKP - 1) / BKP;

  // ---- per-thread constants of the field-tile loader --------------------------------------
  constexpr int EPC = VEC16 ? 2 : 1;  // doubles per cp.async
  // K1: columns are contiguous in memory.  One thread always serves the same column chunk.
  constexpr int CPR1 = C::BN / EPC;           // chunks per k-row
  constexpr int RPP1 = C::NT / CPR1;          // k-rows per pass
  static_assert(KCONTIG || (C::NT % CPR1 == 0 && C::BK % RPP1 == 0), "K1 loader shape");
  // K2: k is contiguous in memory.
  constexpr int CPR2 = C::BK / EPC;           // chunks per column (field row a)
  constexpr int RPP2 = C::NT / CPR2;          // columns per pass
  static_assert(!KCONTIG || (C::NT % CPR2 == 0 && C::BN % RPP2 == 0), "K2 loader shape");

  const double* b_src = p.U;
  bool b_col_ok = false;
  int b_cc = 0, b_r0 = 0;
  if (!KCONTIG) {
    b_cc = tid % CPR1;
    b_r0 = tid / CPR1;
    const int64_t nn = n0 + int64_t(b_cc) * EPC;
    b_col_ok = nn < p.NN;
    if (b_col_ok) {
      const int64_t a = nn / p.before;
      const int64_t b = nn - a * p.before;
      b_src = p.U + a * p.slab_in + b;
    }
  } else {
    b_cc = tid % CPR2;  // chunk along k
    b_r0 = tid / CPR2;  // first column served
  }

  // operator panels of one chunk through the bulk-copy engine (one thread, one mbarrier per stage)
  auto issue_panels = [&](int stage, int chunk) {
    if (tid != 0) return;
    const int kp0 = chunk * BKP;
    const int npan = min(BKP, p.KP - kp0);
    const unsigned bytes = unsigned(rows_blk) * 32u;
    uint64_t* bar = full_bar + stage;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "r"(unsigned(npan) * bytes)
                 : "memory");
    double* As = smem + size_t(stage) * C::STAGE;
    for (int pp = 0; pp < npan; ++pp)
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
              smem_u32(As + pp * C::BM * 4)),
          "l"(p.Dp + (size_t(kp0 + pp) * p.MP + m0) * 4), "r"(bytes), "r"(smem_u32(bar))
          : "memory");
  };

  // generic (bounds-checked) loader: used for the chunk that crosses the end of the k range
  auto load_stage_edge = [&](int stage, int chunk) {
    double* As = smem + size_t(stage) * C::STAGE;
    double* Bs = As + C::A_STAGE;
    const int kp0 = chunk * BKP;
    if (ATMA) {
      issue_panels(stage, chunk);
    } else {
      for (int idx = tid; idx < BKP * C::BM * 2; idx += C::NT) {
        const int pp = idx / (C::BM * 2);
        const int c = idx - pp * (C::BM * 2);
        const bool ok = (kp0 + pp < p.KP) && ((c >> 1) < rows_blk);
        const double* src = ok ? p.Dp + (size_t(kp0 + pp) * p.MP + m0) * 4 + c * 2 : p.Dp;
        cp_async16(As + (pp * C::BM * 4) + c * 2, src, ok ? 16 : 0);
      }
    }
    const int k0 = kp0 * 4;
    if (!KCONTIG) {
#pragma unroll
      for (int r = 0; r < C::BK / RPP1; ++r) {
        const int kr = b_r0 + r * RPP1;
        const int k = k0 + kr;
        const bool ok = b_col_ok && (k < p.n_in);
        const double* src = ok ? b_src + int64_t(k) * p.before : p.U;
        double* dst = Bs + kr * C::PB + b_cc * EPC;
        if (VEC16) cp_async16(dst, src, ok ? 16 : 0);
        else cp_async8(dst, src, ok ? 8 : 0);
      }
    } else {
      const int k = k0 + b_cc * EPC;
#pragma unroll
      for (int r = 0; r < C::BN / RPP2; ++r) {
        const int col = b_r0 + r * RPP2;
        const int64_t nn = n0 + col;
        const bool ok = (nn < p.NN) && (k < p.n_in);
        const double* src = ok ? p.U + nn * p.slab_in + k : p.U;
        double* dst = Bs + col * C::PB + b_cc * EPC;
        if (VEC16) {
          const int bytes = ok ? ((k + 1 < p.n_in) ? 16 : 8) : 0;
          cp_async16(dst, src, bytes);
        } else {
          cp_async8(dst, src, ok ? 8 : 0);
        }
      }
    }
  };

  // fast loader for chunks that lie entirely inside the k range: running pointers, no index
  // arithmetic in the loop (chunks are always loaded in increasing order).  The loader burst is
  // what keeps a warp away from the tensor pipe, so it is kept to a few instructions per copy.
  constexpr int ASUB = (C::BM * 2 + C::NT - 1) / C::NT;  // 16-byte pieces of one operator panel per thread
  const int64_t a_panel = int64_t(p.MP) * 4;
  const double* a_ptr = p.Dp + size_t(m0) * 4 + tid * 2;
  unsigned a_ok = 0, a_in = 0;  // row is valid / piece lies inside the panel
#pragma unroll
  for (int sub = 0; sub < ASUB; ++sub) {
    if (((tid + sub * C::NT) >> 1) < rows_blk) a_ok |= 1u << sub;
    if (tid + sub * C::NT < C::BM * 2) a_in |= 1u << sub;
  }
  const double* b_ptr;
  int64_t b_rstep;
  unsigned b_ok = 0;
  if (!KCONTIG) {
    b_ptr = b_src + int64_t(b_r0) * p.before;
    b_rstep = int64_t(RPP1) * p.before;
    b_ok = b_col_ok ? ~0u : 0u;
  } else {
    b_ptr = p.U + (n0 + b_r0) * p.slab_in + b_cc * EPC;
    b_rstep = int64_t(RPP2) * p.slab_in;
#pragma unroll
    for (int r = 0; r < C::BN / RPP2; ++r)
      if (n0 + b_r0 + r * RPP2 < p.NN) b_ok |= 1u << r;
  }
  const int64_t b_cstep = KCONTIG ? int64_t(C::BK) : int64_t(C::BK) * p.before;

  auto load_stage = [&](int stage, int chunk) {
    if ((chunk + 1) * C::BK > p.n_in) { load_stage_edge(stage, chunk); return; }
    double* As = smem + size_t(stage) * C::STAGE;
    double* Bs = As + C::A_STAGE;
    if (ATMA) {
      issue_panels(stage, chunk);
    } else {
#pragma unroll
      for (int pp = 0; pp < BKP; ++pp)
#pragma unroll
        for (int sub = 0; sub < ASUB; ++sub) {
          const bool ok = (a_ok >> sub) & 1u;
          if ((C::BM * 2) % C::NT == 0 || ((a_in >> sub) & 1u))
            cp_async16(As + pp * C::BM * 4 + (tid + sub * C::NT) * 2,
                       ok ? a_ptr + pp * a_panel + sub * C::NT * 2 : p.Dp, ok ? 16 : 0);
        }
      a_ptr += BKP * a_panel;
    }
    constexpr int BR = KCONTIG ? C::BN / RPP2 : C::BK / RPP1;
    constexpr int RPP = KCONTIG ? RPP2 : RPP1;
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      const bool ok = (b_ok >> (KCONTIG ? r : 0)) & 1u;
      double* dst = Bs + (b_r0 + r * RPP) * C::PB + b_cc * EPC;
      const double* src = ok ? b_ptr + r * b_rstep : p.U;
      if (VEC16) cp_async16(dst, src, ok ? 16 : 0);
      else cp_async8(dst, src, ok ? 8 : 0);
    }
    b_ptr += b_cstep;
  };

  double acc[WM][WN][2];
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  // ---- prologue: STAGES-1 chunks in flight ---------------------------------------------------
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }

  // fragment addressing: A fragment (8 rows x 4 k) = 32 consecutive doubles, lane l reads double l;
  // B fragment: k = t, column = g
  const int a_off = my_mt0 * 32 + lane;
  const int b_off = KCONTIG ? ((wn * WN * 8 + g) * C::PB + t) : (t * C::PB + wn * WN * 8 + g);
  const int nfull = p.KP / BKP;       // chunks whose BKP panels are all valid
  const int tail = p.KP - nfull * BKP;

  // ---- main loop ----------------------------------------------------------------------------
  for (int kc = 0; kc < nchunks; ++kc) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kc + STAGES - 1;
      if (nxt < nchunks) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    if (ATMA) mbar_wait(full_bar + kc % STAGES, unsigned(kc / STAGES) & 1u);
    const double* As = smem + size_t(kc % STAGES) * C::STAGE + a_off;
    const double* Bs = smem + size_t(kc % STAGES) * C::STAGE + C::A_STAGE + b_off;
    if (kc < nfull) MtDispatch<C, WM, WM, WN, BKP, KCONTIG>::full(my_mt, As, Bs, acc);
    else MtDispatch<C, WM, WM, WN, BKP, KCONTIG>::partial(my_mt, As, Bs, acc, tail);
  }
  cp_async_wait<0>();

  // ---- epilogue: Y = alpha*scale*acc + beta*Y ----------------------------------------------------
  const bool use_beta = (p.beta != 0.0);
  const bool unit = p.scale_mode == PT_SCALE_NONE && p.alpha == 1.0;
#pragma unroll
  for (int j = 0; j < WN; ++j) {
    const int64_t nn = n0 + (wn * WN + j) * 8 + 2 * t;  // this thread's first column
    if (nn >= p.NN) continue;
    const bool two = (nn + 1 < p.NN);
    int64_t a0, b0, a1, b1;  // (a, b) of the two columns
    if (KCONTIG) {
      a0 = nn; b0 = 0; a1 = nn + 1; b1 = 0;
    } else {
      a0 = nn / p.before; b0 = nn - a0 * p.before;
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
    // local addresses (Y, or the beta input C of the scatter form) and destination column offsets
    const int64_t l0 = a0 * p.slab_out + b0, l1 = a1 * p.slab_out + b1;
    int64_t d0 = 0, d1 = 0;
    if (SCAT) {
      const int64_t q0 = a0 / p.sc.a_inner, q1 = a1 / p.sc.a_inner;
      d0 = p.sc.offset + q0 * p.sc.s_outer + (a0 - q0 * p.sc.a_inner) * p.sc.s_inner + b0;
      d1 = p.sc.offset + q1 * p.sc.s_outer + (a1 - q1 * p.sc.a_inner) * p.sc.s_inner + b1;
    }
#pragma unroll
    for (int i = 0; i < WM; ++i) {
      if (i >= my_mt) continue;
      const int row = m0 + (my_mt0 + i) * 8 + g;
      if (row >= p.n_out) continue;
      double r0 = s0, r1 = s1;
      if (p.scale_mode == PT_SCALE_ROW) { const double sr = p.scale[row]; r0 *= sr; r1 *= sr; }
      // explicit roundings: every kernel variant gives the same bits.  alpha = 1 without scaling multiplies nothing
      // (1 * x = x exactly): a plain FP64 instruction issued here queues behind the DMMAs of the CTA that shares the SM
      double v0 = unit ? acc[i][j][0] : __dmul_rn(r0, acc[i][j][0]);
      double v1 = unit ? acc[i][j][1] : __dmul_rn(r1, acc[i][j][1]);
      const int64_t off = int64_t(row) * p.before;
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
        // before is even and nn is even: both columns are adjacent in the same slab
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

// fragment-major packing: packed[((j/4)*MP + i)*4 + j%4] = D[i*ldD + j], zero elsewhere
__global__ void pack_operator_kernel(const double* __restrict__ D, int n_out, int n_in,
                                     int64_t ldD, int MP, int KP, double* __restrict__ packed) {
  const int64_t total = int64_t(KP) * MP * 4;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int kk = int(idx & 3);
    const int64_t r = idx >> 2;
    const int i = int(r % MP);
    const int kp = int(r / MP);
    const int j = kp * 4 + kk;
    packed[idx] = (i < n_out && j < n_in) ? D[int64_t(i) * ldD + j] : 0.0;
  }
}

template <int WARPS_M, int WARPS_N, int WM, int WN, int BKP, int STAGES, bool KCONTIG, bool VEC16, int MINB = 1, bool ATMA = false,
          bool SCAT = false>
int launch_cfg(DenseParams& p, cudaStream_t stream) {
  using C = Cfg<WARPS_M, WARPS_N, WM, WN, BKP, STAGES, KCONTIG, VEC16>;
  auto kern = dense_apply_kernel<WARPS_M, WARPS_N, WM, WN, BKP, STAGES, KCONTIG, VEC16, MINB, ATMA, SCAT>;
  constexpr size_t SMEM_BYTES = C::SMEM + (ATMA ? STAGES * sizeof(uint64_t) : 0);
  static bool attr_set[64] = {false};
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
    attr_set[dev] = true;
  }
  const int mtiles = p.MP / 8;
  const int max_mt = WARPS_M * WM;
  p.m_blocks = (mtiles + max_mt - 1) / max_mt;
  p.mt_per_block = (mtiles + p.m_blocks - 1) / p.m_blocks;
  p.m_blocks = (mtiles + p.mt_per_block - 1) / p.mt_per_block;
  const int64_t n_blocks = (p.NN + C::BN - 1) / C::BN;
  const int64_t grid = n_blocks * p.m_blocks;
  if (grid <= 0) return PT_OK;
  if (grid > 2147483647LL) return pt::fail(PT_ERR_UNSUPPORTED, "pt_dense_apply: grid too large");
  kern<<<unsigned(grid), C::NT, SMEM_BYTES, stream>>>(p);
  return pt::check_launch("pt_dense_apply");
}

int g_dense_config = -1;  // tuning hook (pt_debug_set_dense_config); -1 = choose by shape
int g_dense_atma = 0;     // tuning hook (pt_debug_set_dense_tma(2)): operator panels via cp.async.bulk

// CTAs the large tiles would launch: below one per SM the GPU is mostly idle (single fields:
// BASELINE config 3 at B = 1), so the small 64x32 tile (config 2) is used to fill it
int64_t ctas_of(const DenseParams& p, int bm, int bn) {
  const int64_t mb = (p.n_out + bm - 1) / bm, nb = (p.NN + bn - 1) / bn;
  return mb * nb;
}

int choose_config(const DenseParams& p) {
  if (g_dense_config >= 0) return g_dense_config;
  if (ctas_of(p, 96, 64) < pt::sm_count()) return 2;
  // measured on B200 (tools/sweep_dense.py, profiles/r01_dense_config_sweep.md): the 96x64 tile
  // with 3 CTAs/SM wins everywhere except when n_out is a large multiple of 128, where the
  // 128x64 tile with 2 CTAs/SM has no ragged last block and slightly better operand reuse
  return (p.n_out >= 1024 && p.n_out % 128 == 0) ? 5 : 9;
}

template <bool KCONTIG, bool VEC16, bool SCAT = false>
int launch_shape(DenseParams& p, cudaStream_t stream) {
  const int cfg = choose_config(p);
  if (SCAT) {
    if (cfg == 5) return launch_cfg<2, 2, 8, 4, 4, 4, KCONTIG, VEC16, 2, false, SCAT>(p, stream);
    return launch_cfg<1, 4, 12, 2, 4, 3, KCONTIG, VEC16, 3, false, SCAT>(p, stream);
  }
  if (g_dense_atma) {
    if (cfg == 5) return launch_cfg<2, 2, 8, 4, 4, 4, KCONTIG, VEC16, 2, true>(p, stream);
    return launch_cfg<1, 4, 12, 2, 4, 3, KCONTIG, VEC16, 3, true>(p, stream);
  }
  switch (cfg) {
    case 2: return launch_cfg<2, 2, 4, 2, 4, 4, KCONTIG, VEC16, 4>(p, stream);   // 32x16 per warp,  64x32 CTA, 4 CTAs/SM
    case 3: return launch_cfg<2, 2, 4, 4, 4, 4, KCONTIG, VEC16, 4>(p, stream);   // 32x32 per warp,  64x64 CTA, 4 CTAs/SM
    case 5: return launch_cfg<2, 2, 8, 4, 4, 4, KCONTIG, VEC16, 2>(p, stream);   // 64x32 per warp, 128x64 CTA, 2 CTAs/SM
    default: return launch_cfg<1, 4, 12, 2, 4, 3, KCONTIG, VEC16, 3>(p, stream); // 96x16 per warp,  96x64 CTA, 3 CTAs/SM
  }
}

inline bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

}  // namespace

extern "C" int pt_debug_set_dense_tma(int mode) {
  g_dense_atma = (mode == 2);  // 2: cp.async kernel with bulk-copied operator panels (A/B timing; bit-identical)
  return PT_OK;
}

extern "C" int pt_debug_set_dense_config(int id) {
  g_dense_config = id;
  return PT_OK;
}

extern "C" int64_t pt_packed_size(int n_out, int n_in) {
  if (n_out <= 0 || n_in <= 0) return 0;
  const int64_t MP = 8 * ((int64_t(n_out) + 7) / 8);
  const int64_t KP = (int64_t(n_in) + 3) / 4;
  return KP * MP * 4;
}

extern "C" int pt_operator_pack(const double* D, int n_out, int n_in, int64_t ldD, double* packed,
                                pt_stream_t stream) {
  PT_REQUIRE(D && packed, "pt_operator_pack: null pointer");
  PT_REQUIRE(n_out > 0 && n_in > 0 && ldD >= n_in, "pt_operator_pack: bad shape %d x %d ld %lld",
             n_out, n_in, (long long)ldD);
  const int MP = 8 * ((n_out + 7) / 8);
  const int KP = (n_in + 3) / 4;
  const int64_t total = int64_t(KP) * MP * 4;
  const int threads = 256;
  int64_t nb = (total + threads - 1) / threads;
  const int blocks = int(nb < 148 * 16 ? nb : 148 * 16);
  pack_operator_kernel<<<blocks, threads, 0, pt::as_stream(stream)>>>(D, n_out, n_in, ldD, MP, KP,
                                                                     packed);
  return pt::check_launch("pt_operator_pack");
}

// Shared body of pt_dense_apply / pt_dense_apply_ld: ldu, ldy = distance in doubles between consecutive
// `a` slabs of U and Y (n_in*before and n_out*before for whole fields; larger when U / Y are row ranges
// along the axis of a longer field, as in the block substitutions of pt_lu_solve).
static int dense_apply_impl(const char* who, const double* packed, int n_out, int n_in, const double* U, int64_t ldu,
                            double* Y, int64_t ldy, int64_t after, int64_t before, double alpha, double beta,
                            const double* scale, int scale_mode, int64_t scale_len, pt_stream_t stream) {
  PT_REQUIRE(packed && U && Y, "%s: null pointer", who);
  PT_REQUIRE(n_out > 0 && n_in > 0, "%s: bad operator shape %d x %d", who, n_out, n_in);
  PT_REQUIRE(after >= 0 && before >= 1, "%s: bad field shape after=%lld before=%lld", who,
             (long long)after, (long long)before);
  PT_REQUIRE(ldu >= int64_t(n_in) * before && ldy >= int64_t(n_out) * before, "%s: slab strides %lld / %lld too small",
             who, (long long)ldu, (long long)ldy);
  PT_REQUIRE(scale_mode >= PT_SCALE_NONE && scale_mode <= PT_SCALE_BEFORE, "%s: bad scale_mode %d", who, scale_mode);
  PT_REQUIRE(scale_mode == PT_SCALE_NONE || scale != nullptr, "%s: scale is null", who);
  PT_REQUIRE(scale_mode != PT_SCALE_AFTER || scale_len > 0, "%s: scale_len must be > 0", who);
  PT_REQUIRE(U != Y, "%s: in-place application is not supported", who);
  if (after == 0) return PT_OK;

  DenseParams p;
  p.Dp = packed; p.U = U; p.Y = Y; p.C = Y; p.scale = scale;
  p.sc = ScatterDev{nullptr, 1, 1, 0, 0, 0, 0};
  p.n_out = n_out; p.n_in = n_in;
  p.MP = 8 * ((n_out + 7) / 8);
  p.KP = (n_in + 3) / 4;
  p.before = before;
  p.NN = after * before;
  p.slab_in = ldu;
  p.slab_out = ldy;
  p.alpha = alpha; p.beta = beta;
  p.scale_mode = scale_mode; p.scale_len = scale_len > 0 ? scale_len : 1;
  p.m_blocks = 1; p.mt_per_block = 1;
  cudaStream_t st = pt::as_stream(stream);
  if (before == 1) {
    const bool vec = (n_in % 2 == 0) && (ldu % 2 == 0) && aligned16(U);
    return vec ? launch_shape<true, true>(p, st) : launch_shape<true, false>(p, st);
  }
  const bool vec = (before % 2 == 0) && (ldu % 2 == 0) && (ldy % 2 == 0) && aligned16(U) && aligned16(Y);
  return vec ? launch_shape<false, true>(p, st) : launch_shape<false, false>(p, st);
}

extern "C" int pt_dense_apply(const double* packed, int n_out, int n_in, const double* U, double* Y,
                              int64_t after, int64_t before, double alpha, double beta,
                              const double* scale, int scale_mode, int64_t scale_len,
                              pt_stream_t stream) {
  return dense_apply_impl("pt_dense_apply", packed, n_out, n_in, U, int64_t(n_in) * before, Y, int64_t(n_out) * before,
                          after, before, alpha, beta, scale, scale_mode, scale_len, stream);
}

extern "C" int pt_dense_apply_ld(const double* packed, int n_out, int n_in, const double* U, int64_t ldu, double* Y,
                                 int64_t ldy, int64_t after, int64_t before, double alpha, double beta,
                                 pt_stream_t stream) {
  return dense_apply_impl("pt_dense_apply_ld", packed, n_out, n_in, U, ldu, Y, ldy, after, before, alpha, beta, nullptr,
                          PT_SCALE_NONE, 1, stream);
}

// ---- dense apply fused with the pencil redistribution (scatter epilogue over peer memory) -------
extern "C" int pt_dense_apply_scatter(const double* packed, int n_out, int n_in, const double* U,
                                      const double* C, int64_t after, int64_t before, double alpha,
                                      double beta, const pt_scatter_map* map, pt_stream_t stream) {
  PT_REQUIRE(packed && U && map && map->base, "pt_dense_apply_scatter: null pointer");
  PT_REQUIRE(n_out > 0 && n_in > 0, "pt_dense_apply_scatter: bad operator shape %d x %d", n_out, n_in);
  PT_REQUIRE(after >= 0 && before >= 1, "pt_dense_apply_scatter: bad field shape");
  PT_REQUIRE(map->nranks >= 1 && map->chunk >= 1 && int64_t(map->chunk) * map->nranks >= n_out,
             "pt_dense_apply_scatter: %d destinations x %d rows do not cover n_out=%d", map->nranks,
             map->chunk, n_out);
  PT_REQUIRE(map->a_inner >= 1, "pt_dense_apply_scatter: a_inner must be >= 1");
  PT_REQUIRE(beta == 0.0 || C != nullptr, "pt_dense_apply_scatter: beta != 0 needs C");
  if (after == 0) return PT_OK;
  DenseParams p;
  p.Dp = packed; p.U = U; p.Y = nullptr; p.C = C ? C : U; p.scale = nullptr;
  p.sc = ScatterDev{map->base, map->chunk, map->a_inner, map->s_outer, map->s_inner, map->s_i, map->offset};
  p.n_out = n_out; p.n_in = n_in;
  p.MP = 8 * ((n_out + 7) / 8);
  p.KP = (n_in + 3) / 4;
  p.before = before;
  p.NN = after * before;
  p.slab_in = int64_t(n_in) * before;
  p.slab_out = int64_t(n_out) * before;
  p.alpha = alpha; p.beta = beta;
  p.scale_mode = PT_SCALE_NONE; p.scale_len = 1;
  p.m_blocks = 1; p.mt_per_block = 1;
  cudaStream_t st = pt::as_stream(stream);
  if (before == 1) {
    const bool vec = (n_in % 2 == 0) && aligned16(U);
    return vec ? launch_shape<true, true, true>(p, st) : launch_shape<true, false, true>(p, st);
  }
  // 16-byte destination stores need every destination offset to be even (bases are 256-byte aligned)
  const bool even = !((map->s_outer | map->s_inner | map->s_i | map->offset) & 1);
  const bool vec = (before % 2 == 0) && aligned16(U) && (C == nullptr || aligned16(C)) && even;
  return vec ? launch_shape<false, true, true>(p, st) : launch_shape<false, false, true>(p, st);
}

// ---- FP64 tensor-core peak (roofline denominator) ----------------------------------------------
namespace {
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  const double a = 0.5 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int pt_measure_dmma_peak(double* tflops_host, int iters) {
  PT_REQUIRE(tflops_host && iters > 0, "pt_measure_dmma_peak: bad argument");
  const int grid = pt::sm_count() * 2;
  double* out = nullptr;
  PT_CUDA(cudaMalloc(&out, sizeof(double) * grid * 256));
  cudaEvent_t e0, e1;
  PT_CUDA(cudaEventCreate(&e0));
  PT_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    PT_CUDA(cudaEventRecord(e0));
    dmma_peak_kernel<<<grid, 256>>>(out, iters);
    PT_CUDA(cudaEventRecord(e1));
    PT_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    PT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  const double flops = 2.0 * 256.0 * 16.0 * double(iters) * 8.0 * grid;
  *tflops_host = flops / (best * 1e-3) * 1e-12;
  return PT_OK;
}
