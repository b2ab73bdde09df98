// K1/K2 for centrosymmetric operators, warp-specialised form (the default for large fields).
//
// Same algebra and the same per-accumulator summation order as dense_apply_eo.cu (results are
// bit-identical for even n); what changes is who does what inside the CTA:
//
//   * 8 CONSUMER warps (2 x 4, each 64 half-rows x 16 columns, an even and an odd accumulator set)
//     do nothing but LDS + DMMA: they wait on a "full" mbarrier per ring stage, read the operator
//     fragments and the E / O fragments, and release the stage through an "empty" mbarrier.  No
//     __syncthreads, no loads, no address arithmetic, no DADD in their instruction stream.
//   * 8 PRODUCER warps (two warpgroups, shrunk to 64 registers with setmaxnreg so that the consumers
//     can keep 192; two groups that alternate over the chunks, because a DADD issued behind two warps of
//     back-to-back DMMAs waits ~150 cycles for the FP64 pipe — ncu r02c — and the wait is per warp)
//     own the data movement: one elected lane asks the bulk-copy engine for the
//     operator panels of the chunk (ONE cp.async.bulk of 32 KB: the operator is packed per
//     (128-row block, k chunk) for exactly this), all lanes load the field tile and its mirror image
//     from global memory, form E = u + Ju and O = u - Ju once per 128 half-rows (the cp.async kernel
//     forms them per warp and per 64 half-rows, on the FP64 pipe the DMMAs need) and store them into
//     the stage (swizzled, conflict-free for both sides).
//   * odd n (config 2's n = 257): the half problem is (n-1)/2 x (n-1)/2 exactly (H = 128, no 17th
//     row tile, no 33rd k panel); the middle COLUMN is a rank-1 term added to the even accumulator in
//     the epilogue (one FMA per element), the middle ROW is a dot product the producers accumulate
//     from the E (or O) values they form anyway (two independent DFMA chains per thread: a dependent
//     chain waits for the FP64 pipe behind the DMMAs at every link); the eight partial sums of a column
//     travel through shared memory with the tile's last two chunks and the epilogue adds and stores them.
//     (Accumulating in the consumer warps was measured: the DFMAs stall the DMMA stream, 0.24 vs 0.19 ms.)
//   * COPY (pencil redistribution): the producers of the first row block also store the field tiles
//     they load — every element of U passes through their registers exactly once — through a second
//     scatter map into the peers' buffers.  The redistribution of the FIELD then needs no kernel of its
//     own (this CTA shape leaves no room on an SM for a concurrent copy kernel): its NVLink traffic is
//     spread under the whole main loop, like that of the result, which leaves through the epilogue.
//
// One CTA per SM (192 KB ring of 4 stages), persistent over tiles (m fastest).
#include "dense_eo_common.cuh"

using namespace ptdense;

namespace {

constexpr int CW_M = 2, CW_N = 4, WR = 8, WN = 2;       // consumer warps and their tiles
constexpr int NCONS = CW_M * CW_N * 32;                 // 256 consumer threads
constexpr int NPROD = 256;                              // 8 producer warps: two groups of four, alternating chunks
constexpr int PWARPS = NPROD / 32;
constexpr int NPASS = 4;                                // passes of a producer group over a chunk
constexpr int NT = NCONS + NPROD;
constexpr int BMH = CW_M * WR * 8;                      // 128 half-rows per tile
constexpr int BN = CW_N * WN * 8;                       // 64 columns per tile
constexpr int BK = 16;                                  // k per chunk = 2 panel pairs
constexpr int STAGES = 4;
constexpr int STAGGER = 2;                              // chunks by which the second consumer row half trails the first
constexpr int A_PAIR = 2 * BMH * 8;                     // doubles of one panel pair: [e|o][128][8]
constexpr int A_STAGE = 2 * A_PAIR;                     // 4096 doubles = 32 KB
constexpr int B_TILE = BN * BK;                         // 1024 doubles = 8 KB (E tile; O tile follows)
constexpr int STAGE = A_STAGE + 2 * B_TILE;             // 48 KB
constexpr size_t SMEM_RING = size_t(STAGES) * STAGE * sizeof(double);
constexpr size_t SMEM_FIXED = SMEM_RING + 2 * STAGES * sizeof(uint64_t) + (2 * BN + 2 * PWARPS * BN) * sizeof(double);
constexpr size_t SMEM_MAX = 227 * 1024;                 // + cmid (odd n): MPH doubles

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Wait for a phase.  A pipeline bug must not hang the GPU (a hung device costs the whole box): the spin
// is bounded — every failed try_wait has itself waited for the hardware's time limit, so 2^22 of them are
// far beyond any legitimate wait (tens of microseconds) — and the kernel traps instead, which surfaces as
// a CUDA error on the host.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .u32 spins;\n"
      "mov.u32 spins, 0;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.gt.u32 q, spins, 4194304;\n"
      "@q trap;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// non-blocking probe of a phase (the blocking form above may suspend the warp)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ double2 ldg2(const double* p) {
  return __ldg(reinterpret_cast<const double2*>(p));
}

// ---- consumer: DMMAs of one staged chunk, specialised on the exact number of row tiles ------------
// NPAIRS is a template parameter so that a full chunk is ONE straight-line block: the fragment loads of the second
// panel pair can be scheduled under the DMMAs of the first (behind a run-time `pr < npairs` they could not, and a
// consumer warp sat through the shared-memory latency once more per chunk).
template <int RT, bool KCONTIG, int NPAIRS>
__device__ __forceinline__ void ws_compute(const double* As, const double* Et, const double* Ot,
                                           int swz, const int (&bcol)[WN], double (&acce)[WR][WN][2],
                                           double (&acco)[WR][WN][2]) {
#pragma unroll
  for (int pr = 0; pr < NPAIRS; ++pr) {
    {
      double2 ae[WR], ao[WR], be[WN], bo[WN];
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        ae[i] = *reinterpret_cast<const double2*>(As + pr * A_PAIR + i * 64);
        ao[i] = *reinterpret_cast<const double2*>(As + pr * A_PAIR + BMH * 8 + i * 64);
      }
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        if (KCONTIG) {          // [column][16 k]: one 16-byte read gives k = 8P + 2t and 8P + 2t + 1
          be[j] = *reinterpret_cast<const double2*>(Et + j * 8 * BK + (pr ^ swz) * 8);
          bo[j] = *reinterpret_cast<const double2*>(Ot + j * 8 * BK + (pr ^ swz) * 8);
        } else {                // [k][64 columns], column XOR 4*((k >> 1) & 3)
          be[j] = make_double2(Et[pr * 8 * BN + bcol[j]], Et[(pr * 8 + 1) * BN + bcol[j]]);
          bo[j] = make_double2(Ot[pr * 8 * BN + bcol[j]], Ot[(pr * 8 + 1) * BN + bcol[j]]);
        }
      }
      // per accumulator: panel (k = 8P + 2t) before panel (k = 8P + 2t + 1), pairs in order — the
      // summation order of dense_apply_eo.cu
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
  }
}

template <int RT, bool KCONTIG>
struct WsDispatch {
  static __device__ __forceinline__ void run(int my_rt, const double* As, const double* Et, const double* Ot,
                                             int npairs, int swz, const int (&bcol)[WN], double (&acce)[WR][WN][2],
                                             double (&acco)[WR][WN][2]) {
    if (my_rt == RT) {
      if (npairs == 2) ws_compute<RT, KCONTIG, 2>(As, Et, Ot, swz, bcol, acce, acco);
      else ws_compute<RT, KCONTIG, 1>(As, Et, Ot, swz, bcol, acce, acco);
    } else WsDispatch<RT - 1, KCONTIG>::run(my_rt, As, Et, Ot, npairs, swz, bcol, acce, acco);
  }
};
template <bool KCONTIG>
struct WsDispatch<0, KCONTIG> {
  static __device__ __forceinline__ void run(int, const double*, const double*, const double*, int, int,
                                             const int (&)[WN], double (&)[WR][WN][2], double (&)[WR][WN][2]) {}
};

// destination of element (a, i, b) under a scatter map: the (a, b) part and the i part separately
__device__ __forceinline__ int64_t map_a_part(const ScatterDev& m, int64_t a, int64_t b) {
  const int64_t q = ((uint64_t(a) | uint64_t(m.a_inner)) >> 32) ? a / m.a_inner : int64_t(uint32_t(a) / uint32_t(m.a_inner));
  return m.offset + q * m.s_outer + (a - q * m.a_inner) * m.s_inner + b;
}
__device__ __forceinline__ double* map_i_part(const ScatterDev& m, int i) {
  const int q = i / m.chunk;
  return m.base[q] + int64_t(i - q * m.chunk) * m.s_i;
}

template <bool KCONTIG, bool SCAT, bool MIDFIX, bool COPY>
__global__ void __launch_bounds__(NT, 1) dense_apply_eo_ws_kernel(const EoParams p) {
  extern __shared__ __align__(128) double smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(smem) + SMEM_RING);
  uint64_t* empty = full + STAGES;
  double* umid_s = reinterpret_cast<double*>(empty + STAGES);       // [2 tile parities][BN]: u[mid] of the tile's columns
  double* midbuf = umid_s + 2 * BN;                                 // [2 tile parities][8 producer warps][BN]: middle-row partial sums
  double* cmid_s = midbuf + 2 * PWARPS * BN;                             // [MPH]: D[:, mid] (odd n only)
  double* rmid_s = cmid_s + p.MPH;                                  // [nchunks * BK + 1]: D[mid, :], then D[mid][mid]

  const int tid = threadIdx.x;
  const int n = p.n, H = p.H;
  const int npanel_pairs = (H + 7) >> 3;               // pairs of k-panels = groups of 8 k
  const int nchunks = (npanel_pairs + 1) >> 1;
  const int tail_pairs = npanel_pairs & 1;             // 1: the last chunk holds one pair only
  const int rtiles_total = (H + 7) >> 3;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, NPROD / 2 + 1); mbar_init(empty + s, CW_M * CW_N); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (MIDFIX) {
    for (int i = tid; i < p.MPH; i += NT) cmid_s[i] = p.cmid[i];
    for (int i = tid; i <= nchunks * BK; i += NT) rmid_s[i] = p.rmid[i];
  }
  __syncthreads();

  if (tid >= NCONS) {
    // =============================== producers ====================================================
    // Two groups of four warps; group g stages the chunks g, g + 2, g + 4, ... of this CTA's (tile, chunk)
    // sequence.  A thread keeps ONE register set: right after it has stored chunk i it issues the global
    // loads of chunk i + 2, which then have the other group's whole chunk to arrive.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");
    const int ptid = (tid - NCONS) & 127;
    const int group = (tid - NCONS) >> 7;
    // tile-independent part of this thread's addressing
    const int c2 = KCONTIG ? 2 * (ptid & 7) : 0;              // K2: first k of the 16-byte piece
    const int colp = ptid >> 3;                               // K2: column of pass 0
    const int kr0 = KCONTIG ? 0 : (ptid >> 5);                // K1: k row of pass 0
    const int sto = KCONTIG ? colp * BK + (((ptid & 7) ^ ((colp & 1) ? 4 : 0)) << 1)   // odd columns: piece XOR 4
                            : 2 * (ptid & 31);                // K1: column pair of the tile
    const int64_t pass_step = KCONTIG ? int64_t(16) * n : 4 * p.before;
    const int64_t chunk_step = KCONTIG ? int64_t(BK) : int64_t(BK) * p.before;
    // tile indices fit 32 bits (checked on the host): no 64-bit divisions in the loops
    const unsigned ntiles = unsigned(p.total_tiles), mblocks = unsigned(p.m_blocks);
    const unsigned my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t total_it = int64_t(my_tiles) * nchunks;

    // ---- position in the (tile, chunk) sequence + this thread's forward / mirrored source pointers ----
    unsigned tile = blockIdx.x, ti = 0;      // tile index; count of this CTA's tiles so far
    int mblock = 0;                          // row block of the tile
    int64_t n0 = 0;                          // first column of the tile
    int kc = 0;
    const double *fwd = p.U, *mir = p.U;
    unsigned ok = 0;          // K2: validity of the column of each pass; K1: all-or-nothing
    auto setup = [&]() {
      const unsigned nb = tile / mblocks;
      mblock = int(tile - nb * mblocks);
      n0 = int64_t(nb) * BN;
      ok = 0;
      kc = 0;
      if (KCONTIG) {
        fwd = p.U + (n0 + colp) * n + c2;
        mir = p.U + (n0 + colp) * n + (n - 2 - c2);
#pragma unroll
        for (int r = 0; r < NPASS; ++r)
          if (n0 + colp + 16 * r < p.NN) ok |= 1u << r;
      } else {
        const int64_t nn = n0 + sto;
        int64_t la = 0;
        if (nn < p.NN) {
          ok = 0xfu;
          int64_t a, b;
          split_column(nn, p.before, a, b);
          la = a * p.slab + b;
        }
        fwd = p.U + la + int64_t(kr0) * p.before;
        mir = p.U + la + int64_t(n - 1 - kr0) * p.before;
      }
    };
    auto advance = [&]() {
      fwd += chunk_step;
      mir -= chunk_step;
      if (++kc == nchunks) {
        tile += gridDim.x;
        ++ti;
        if (tile < ntiles) setup();
        else ok = 0;
      }
    };

    double2 f[NPASS], m[NPASS];
    double2 um = make_double2(0.0, 0.0);     // MIDFIX: u[mid] of this thread's column pair (k row 0 warps)
    auto load = [&]() {
      const int k0 = kc * BK;
      if (MIDFIX && kc == 0 && kr0 == 0 && ok) um = ldg2(fwd + int64_t(p.mid) * p.before);   // fwd is at row 0 here
#pragma unroll
      for (int r = 0; r < NPASS; ++r) {
        // K2: pieces k0 + c2, k0 + c2 + 1 of column colp + 16r; K1: row k0 + kr0 + 4r of the column pair
        const bool pok = KCONTIG ? (((ok >> r) & 1u) && (k0 + c2 < H)) : (ok && (k0 + kr0 + 4 * r < H));
        if (pok) {
          f[r] = ldg2(fwd + r * pass_step);
          m[r] = ldg2(KCONTIG ? mir + r * pass_step : mir - r * pass_step);
        } else {
          f[r] = make_double2(0.0, 0.0);
          m[r] = make_double2(0.0, 0.0);
        }
      }
    };

    double macc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};   // MIDFIX: middle-row partial sums, two independent chains
    setup();
    if (group == 1) advance();
    unsigned it = group;
    if (it < total_it) load();
    while (it < total_it) {
      // ---- chunk `it` = (tile, kc): registers -> E / O tiles of ring slot it % STAGES ----------------
      const int slot = it % STAGES;
      const unsigned round = it / STAGES;
      const int par = int(ti & 1);
      if (p.trace && ptid == 0) {
        const long long t0 = clock64();
        if (round > 0) mbar_wait(empty + slot, (round - 1) & 1u);
        if (ti < 16) {
          long long* tr = p.trace + (int64_t(blockIdx.x) * 16 + ti) * 8;
          tr[4] += clock64() - t0;                 // cycles the producers waited for a free slot
          if (kc < 2) tr[5 + group] = t0;          // this group reaches the tile
        }
      } else if (round > 0) mbar_wait(empty + slot, (round - 1) & 1u);
      double* As = smem + size_t(slot) * STAGE;
      double* Et = As + A_STAGE;
      double* Ot = Et + B_TILE;
      if (ptid == 0) {
        const int npairs = (kc == nchunks - 1 && tail_pairs) ? 1 : 2;
        const unsigned bytes = unsigned(npairs) * A_PAIR * sizeof(double);
        mbar_arrive_expect_tx(full + slot, bytes);
        bulk_g2s(As, p.Dpe + (size_t(mblock) * nchunks + kc) * A_STAGE, bytes, full + slot);
      }
      if (MIDFIX && kc == 0 && kr0 == 0)      // u[mid] of the tile's columns for the consumers' epilogue
        *reinterpret_cast<double2*>(umid_s + par * BN + sto) = um;
      if (COPY && mblock == 0) {
        // the field itself, redistributed: forward pieces cover rows i < H, mirrored pieces rows i >= n - H
        const int k0 = kc * BK;
        if (KCONTIG) {
          const int kf = k0 + c2, km = n - 2 - kf;           // s_i == 1 on the contiguous axis
          double* df = map_i_part(p.cp, kf);
          double* dm = map_i_part(p.cp, km);
#pragma unroll
          for (int r = 0; r < NPASS; ++r)
            if (((ok >> r) & 1u) && kf < H) {
              const int64_t ao = map_a_part(p.cp, n0 + colp + 16 * r, 0);
              *reinterpret_cast<double2*>(df + ao) = f[r];
              *reinterpret_cast<double2*>(dm + ao) = m[r];
            }
        } else if (ok) {
          int64_t a, b;
          split_column(n0 + sto, p.before, a, b);
          const int64_t ao = map_a_part(p.cp, a, b);
#pragma unroll
          for (int r = 0; r < NPASS; ++r) {
            const int kr = k0 + kr0 + 4 * r;
            if (kr < H) {
              *reinterpret_cast<double2*>(map_i_part(p.cp, kr) + ao) = f[r];
              *reinterpret_cast<double2*>(map_i_part(p.cp, n - 1 - kr) + ao) = m[r];
            }
          }
          if (MIDFIX && kc == 0 && kr0 == 0) *reinterpret_cast<double2*>(map_i_part(p.cp, p.mid) + ao) = um;
        }
      }
      // all FP64 work of the chunk in one burst (in place: E over f, O over m)
#pragma unroll
      for (int r = 0; r < NPASS; ++r) {
        const double2 a = f[r], b = m[r];
        if (KCONTIG) {      // the mirrored piece holds (u[n-2-k], u[n-1-k]): pair swapped
          f[r] = make_double2(__dadd_rn(a.x, b.y), __dadd_rn(a.y, b.x));
          m[r] = make_double2(__dsub_rn(a.x, b.y), __dsub_rn(a.y, b.x));
        } else {
          f[r] = make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y));
          m[r] = make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y));
        }
      }
      if (MIDFIX && mblock == 0) {            // rmid_s is zero beyond H, E / O are zero there too
#pragma unroll
        for (int r = 0; r < NPASS; ++r) {
          const double rk = rmid_s[kc * BK + kr0 + 4 * r];
          const double2 x = p.sign > 0 ? f[r] : m[r];
          macc[r & 1][0] = __fma_rn(rk, x.x, macc[r & 1][0]);
          macc[r & 1][1] = __fma_rn(rk, x.y, macc[r & 1][1]);
        }
      }
#pragma unroll
      for (int r = 0; r < NPASS; ++r) {
        if (KCONTIG) {
          *reinterpret_cast<double2*>(Et + sto + r * (16 * BK)) = f[r];       // 16 columns further: same parity
          *reinterpret_cast<double2*>(Ot + sto + r * (16 * BK)) = m[r];
        } else {
          const int kr = kr0 + 4 * r;
          const int so = kr * BN + (sto ^ (4 * ((kr >> 1) & 3)));
          *reinterpret_cast<double2*>(Et + so) = f[r];
          *reinterpret_cast<double2*>(Ot + so) = m[r];
        }
      }
      // the last chunk of a tile that THIS group stages: its partial sums of the middle row ride with it
      // (visible to the consumers once they have waited for the tile's last two stages)
      if (MIDFIX && mblock == 0 && kc + 2 >= nchunks) {
        double* buf = midbuf + ((par * 2 + group) * 4 + kr0) * BN + sto;
        *reinterpret_cast<double2*>(buf) = make_double2(__dadd_rn(macc[0][0], macc[1][0]), __dadd_rn(macc[0][1], macc[1][1]));
        macc[0][0] = macc[0][1] = macc[1][0] = macc[1][1] = 0.0;
      }
      mbar_arrive(full + slot);
      // ---- two chunks further ---------------------------------------------------------------------------
      it += 2;
      advance();
      advance();
      if (it < total_it) load();
    }
    return;
  }

  // ================================= consumers ======================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 192;\n");
  const int lane = tid & 31, warp = tid >> 5;
  const int wn = warp % CW_N, wm = warp / CW_N;
  const int g = lane >> 2, t = lane & 3;
  int bcol[WN];
#pragma unroll
  for (int j = 0; j < WN; ++j) bcol[j] = (wn * WN * 8 + j * 8 + g) ^ (4 * t);
  const int b_off = KCONTIG ? ((wn * WN * 8 + g) * BK + 2 * t) : (2 * t * BN);
  const int frag_swz = KCONTIG ? (g & 1) : 0;
  unsigned it = 0;
  bool ready = false;
  // The two consumer warps of a sub-partition (one per row half) would otherwise reach every tile's
  // epilogue together and leave the DMMA pipe idle for its whole duration.  The second row half starts two
  // chunks late (it waits for chunk 2 to land before it touches chunk 0); the offset then sustains itself:
  // while one warp stores, the other has the pipe to itself.  tuning hook: pt_debug_set_eo_ws(6) = no stagger
  const unsigned ntiles = unsigned(p.total_tiles), mblocks = unsigned(p.m_blocks);
  if (wm == 1 && p.dbg != 6 && int64_t(nchunks) * ((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) > STAGGER)
    mbar_wait(full + STAGGER, 0);
  unsigned ti = 0;                             // count of this CTA's tiles so far
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
    const unsigned nb = tile / mblocks;
    const int m_block = int(tile - nb * mblocks);
    const int64_t n0 = int64_t(nb) * BN;
    const int rt0 = m_block * (BMH / 8);
    const int rt_count = min(BMH / 8, rtiles_total - rt0);
    const int q = (rt_count + CW_M - 1) / CW_M;
    const int my_rt0 = wm * q;
    const int my_rt = max(0, min(q, rt_count - my_rt0));
    const int a_off = my_rt0 * 64 + 2 * lane;

    double acce[WR][WN][2], acco[WR][WN][2];
#pragma unroll
    for (int i = 0; i < WR; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) { acce[i][j][0] = acce[i][j][1] = 0.0; acco[i][j][0] = acco[i][j][1] = 0.0; }

    const bool tracing = p.trace && tid == 0 && ti < 16;
    long long* tr = tracing ? p.trace + (int64_t(blockIdx.x) * 16 + ti) * 8 : nullptr;
    long long t_wait = 0;
    if (tracing) tr[0] = clock64();                 // consumer starts the tile
    for (int kc = 0; kc < nchunks; ++kc, ++it) {
      const int slot = it % STAGES;
      // `ready` was probed while the previous chunk's DMMAs were issuing: the round trip to the barrier
      // unit (~100 cycles even when the phase has completed) is off the critical path
      if (!ready) {
        if (tracing) {
          const long long t0 = clock64();
          mbar_wait(full + slot, (it / STAGES) & 1u);
          t_wait += clock64() - t0;
        } else mbar_wait(full + slot, (it / STAGES) & 1u);
      }
      ready = mbar_test(full + (it + 1) % STAGES, ((it + 1) / STAGES) & 1u);   // next stage (a later phase can not complete early)
      const double* As = smem + size_t(slot) * STAGE;
      const double* Et = As + A_STAGE + b_off;
      const double* Ot = Et + B_TILE;
      const int npairs = (kc == nchunks - 1 && tail_pairs) ? 1 : 2;
      WsDispatch<WR, KCONTIG>::run(my_rt, As + a_off, Et, Ot, npairs, frag_swz, bcol, acce, acco);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
    }
    if (tracing) { tr[1] = clock64(); tr[3] = t_wait; }        // main loop done; cycles spent waiting for data
    if (p.dbg == 3 && acce[0][0][0] != 12345.678) continue;      // timing experiment: no epilogue at all
    if (MIDFIX && wm == 0 && m_block == 0 && lane < WN * 8) {
      // y[mid] = sum_k D[mid][k] X[k] + D[mid][mid] u[mid]: the producers left eight partial sums per column
      // (written before they arrived on the full barriers of the tile's last two chunks, which this warp has
      // waited for).  Finishing the row in the producer group of the last chunk instead was tried: no faster,
      // and it needs a wait on the OTHER group's full barrier, whose phase can be lapped — a deadlock hazard.
      const int par = int(ti & 1);
      const int col = wn * WN * 8 + lane;
      const int64_t nn = n0 + col;
      if (nn < p.NN) {
        const double* mb = midbuf + par * PWARPS * BN + col;
        // fixed pairwise order (a chain of 7 dependent DADDs is ~2x the latency of the 3-level tree)
        double v = __dadd_rn(__dadd_rn(__dadd_rn(mb[0], mb[BN]), __dadd_rn(mb[2 * BN], mb[3 * BN])),
                             __dadd_rn(__dadd_rn(mb[4 * BN], mb[5 * BN]), __dadd_rn(mb[6 * BN], mb[7 * BN])));
        v = __fma_rn(rmid_s[nchunks * BK], umid_s[par * BN + col], v);
        int64_t a, b;
        split_column(nn, p.before, a, b);
        double sc = p.alpha;
        if (p.scale_mode == PT_SCALE_AFTER) sc *= p.scale[a % p.scale_len];
        else if (p.scale_mode == PT_SCALE_BEFORE) sc *= p.scale[b];
        else if (p.scale_mode == PT_SCALE_ROW) sc *= p.scale[p.mid];
        double out = __dmul_rn(sc, v);
        const int64_t off = a * p.slab + int64_t(p.mid) * p.before + b;
        if (p.beta != 0.0) out = __fma_rn(p.beta, (SCAT ? p.C : p.Y)[off], out);
        if (SCAT) *(map_i_part(p.sc, p.mid) + map_a_part(p.sc, a, b)) = out;
        else p.Y[off] = out;
      }
    }
    if (tracing) tr[7] = clock64();                              // middle row stored
    if (MIDFIX) {
      // middle column: ze[i][col] += D[i][mid] * u[mid][col] as ONE DMMA per accumulator tile — an outer product
      // whose k = 1..3 terms are zero — instead of two FMAs per element in the epilogue: a plain FP64 instruction
      // issued here queues behind the DMMAs of the warp that shares the sub-partition, a DMMA takes its turn
      const double* um = umid_s + (ti & 1) * BN + wn * WN * 8;
      const double* cm = cmid_s + rt0 * 8 + my_rt0 * 8;
      double bm[WN];
#pragma unroll
      for (int j = 0; j < WN; ++j) bm[j] = t == 0 ? um[j * 8 + g] : 0.0;
#pragma unroll
      for (int i = 0; i < WR; ++i) {
        const double am = t == 0 ? cm[i * 8 + g] : 0.0;
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acce[i][j][0], acce[i][j][1], am, bm[j]);
      }
    }
    eo_epilogue<WR, WN, KCONTIG, true, SCAT, MIDFIX, false>(p, acce, acco, my_rt, rt0 * 8 + my_rt0 * 8, n0 + wn * WN * 8, g, t,
                                                            cmid_s, umid_s + (ti & 1) * BN + wn * WN * 8);
    if (tracing) tr[2] = clock64();                              // epilogue done
  }
}

// operator halves in the layout the bulk copy wants: [m_block][chunk][pair][e|o][128 rows][8 k]
__global__ void pack_eo_ws_kernel(const double* __restrict__ D, int n, int H, int m_blocks, int nchunks,
                                  double* __restrict__ out) {
  const int64_t total = int64_t(m_blocks) * nchunks * A_STAGE;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int kk = int(idx & 7);
    const int row = int((idx >> 3) & (BMH - 1));
    const int eo = int((idx >> 10) & 1);
    const int pair = int((idx >> 11) & 1);
    const int64_t blk = idx >> 12;
    const int kc = int(blk % nchunks), mb = int(blk / nchunks);
    const int i = mb * BMH + row, k = (kc * 2 + pair) * 8 + kk;
    double v = 0.0;
    if (i < H && k < H) {
      const double a = D[int64_t(i) * n + k], b = D[int64_t(i) * n + (n - 1 - k)];
      v = eo == 0 ? __dmul_rn(0.5, __dadd_rn(a, b)) : __dmul_rn(0.5, __dsub_rn(a, b));
    }
    out[idx] = v;
  }
  // odd n: middle column, middle row, middle element after the panels
  if (n & 1) {
    const int mid = (n - 1) / 2;
    double* cmid = out + total;
    double* rmid = cmid + int64_t(m_blocks) * BMH;
    for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < int64_t(m_blocks) * BMH;
         idx += int64_t(gridDim.x) * blockDim.x)
      cmid[idx] = idx < H ? D[idx * n + mid] : 0.0;
    for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx <= int64_t(nchunks) * BK;
         idx += int64_t(gridDim.x) * blockDim.x)
      rmid[idx] = idx < H ? D[int64_t(mid) * n + idx] : (idx == int64_t(nchunks) * BK ? D[int64_t(mid) * n + mid] : 0.0);
  }
}

inline bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

long long* g_ws_trace = nullptr;   // tuning hook: device buffer [grid][16 tiles][8] of clock64 stamps
int g_ws_enabled = 1;      // tuning hook: pt_debug_set_eo_config(60 / 61) = never / default

struct WsShape {
  int H, m_blocks, nchunks;
  int64_t panels;          // doubles of the packed panels
};
WsShape ws_shape(int n) {
  WsShape s;
  s.H = n / 2;
  s.m_blocks = (s.H + BMH - 1) / BMH;
  const int pairs = (s.H + 7) / 8;
  s.nchunks = (pairs + 1) / 2;
  s.panels = int64_t(s.m_blocks) * s.nchunks * A_STAGE;
  return s;
}

template <bool KCONTIG, bool SCAT, bool MIDFIX, bool COPY = false>
int launch_ws(EoParams& p, cudaStream_t stream) {
  auto kern = dense_apply_eo_ws_kernel<KCONTIG, SCAT, MIDFIX, COPY>;
  static bool attr_set[64] = {false};
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_MAX)));
    attr_set[dev] = true;
  }
  int64_t grid = pt::sm_count();
  if (grid > p.total_tiles) grid = p.total_tiles;
  const size_t smem_bytes = SMEM_FIXED + (MIDFIX ? size_t(p.MPH + p.KPH * 4 + 2) * sizeof(double) : 0);
  kern<<<unsigned(grid), NT, smem_bytes, stream>>>(p);
  return pt::check_launch("pt_dense_apply_eo_ws");
}

}  // namespace

namespace ptdense {

int launch_eo_ws(EoParams& p, bool kcontig, bool scatter, bool copy, cudaStream_t stream) {
  const bool mid = p.mid >= 0;
  if (copy) {          // result through the scatter epilogue AND the field itself through the producers
    if (kcontig) return mid ? PT_ERR_UNSUPPORTED : launch_ws<true, true, false, true>(p, stream);
    return mid ? launch_ws<false, true, true, true>(p, stream) : launch_ws<false, true, false, true>(p, stream);
  }
  if (kcontig) {
    if (mid) return PT_ERR_UNSUPPORTED;
    return scatter ? launch_ws<true, true, false>(p, stream) : launch_ws<true, false, false>(p, stream);
  }
  if (mid) return scatter ? launch_ws<false, true, true>(p, stream) : launch_ws<false, false, true>(p, stream);
  return scatter ? launch_ws<false, true, false>(p, stream) : launch_ws<false, false, false>(p, stream);
}

}  // namespace ptdense

namespace {

// shapes the warp-specialised kernel covers; everything else runs dense_apply_eo.cu
bool ws_eligible(int n, const double* U, const double* Y, const double* C, int64_t after, int64_t before,
                 const pt_scatter_map* map, const pt_scatter_map* copy_map) {
  if (!g_ws_enabled || n < 128) return false;
  const WsShape s = ws_shape(n);
  const int64_t NN = after * before;
  const int64_t tiles = int64_t(s.m_blocks) * ((NN + BN - 1) / BN);
  if (tiles < pt::sm_count()) return false;            // small fields: the small-tile cp.async kernel fills the GPU better
  if (tiles >= (int64_t(1) << 31)) return false;       // the kernel indexes tiles with 32 bits
  if (!aligned16(U)) return false;
  if (copy_map) {      // 16-byte stores of the field pieces: even strides, pieces never straddle two destinations
    if ((copy_map->s_outer | copy_map->s_inner | copy_map->offset | copy_map->chunk) & 1) return false;
    if (before > 1 && (copy_map->s_i & 1)) return false;
    if (before == 1 && copy_map->s_i != 1) return false;
  }
  if (before == 1) return (n % 2) == 0;
  if (before % 2) return false;
  if ((n & 1) && SMEM_FIXED + size_t(s.m_blocks * BMH + s.nchunks * BK + 2) * sizeof(double) > SMEM_MAX) return false;   // D[:, mid], D[mid, :] beside the ring
  if (map) {
    if ((map->s_outer | map->s_inner | map->s_i | map->offset) & 1) return false;
    return C == nullptr || aligned16(C);
  }
  return aligned16(Y);
}

int apply_ws_impl(const double* packed, int n, int sign, const double* U, double* Y, const double* C, int64_t after,
                  int64_t before, double alpha, double beta, const double* scale, int scale_mode, int64_t scale_len,
                  const pt_scatter_map* map, const pt_scatter_map* copy_map, pt_stream_t stream) {
  PT_REQUIRE(packed && U, "pt_dense_apply_eo_ws: null pointer");
  PT_REQUIRE(n >= 2, "pt_dense_apply_eo_ws: n=%d", n);
  PT_REQUIRE(sign == 1 || sign == -1, "pt_dense_apply_eo_ws: sign must be +1 or -1");
  PT_REQUIRE(after >= 0 && before >= 1, "pt_dense_apply_eo_ws: bad field shape");
  PT_REQUIRE(scale_mode >= PT_SCALE_NONE && scale_mode <= PT_SCALE_BEFORE, "pt_dense_apply_eo_ws: bad scale_mode %d", scale_mode);
  PT_REQUIRE(scale_mode == PT_SCALE_NONE || scale != nullptr, "pt_dense_apply_eo_ws: scale is null");
  PT_REQUIRE(scale_mode != PT_SCALE_AFTER || scale_len > 0, "pt_dense_apply_eo_ws: scale_len must be > 0");
  if (!ws_eligible(n, U, Y, C, after, before, map, copy_map)) return PT_ERR_UNSUPPORTED;
  if (after == 0) return PT_OK;
  const WsShape s = ws_shape(n);
  EoParams p;
  p.Dpe = packed; p.Dpo = nullptr; p.U = U; p.Y = Y; p.C = C; p.scale = scale;
  p.sc = map ? ScatterDev{map->base, map->chunk, map->a_inner, map->s_outer, map->s_inner, map->s_i, map->offset}
             : ScatterDev{nullptr, 1, 1, 0, 0, 0, 0};
  p.cp = copy_map ? ScatterDev{copy_map->base, copy_map->chunk, copy_map->a_inner, copy_map->s_outer, copy_map->s_inner,
                               copy_map->s_i, copy_map->offset}
                  : ScatterDev{nullptr, 1, 1, 0, 0, 0, 0};
  p.n = n; p.H = s.H; p.MPH = s.m_blocks * BMH; p.KPH = s.nchunks * 4;
  p.rt_per_block = BMH / 8; p.m_blocks = s.m_blocks;
  p.sign = sign;
  p.before = before; p.NN = after * before; p.slab = int64_t(n) * before;
  p.total_tiles = int64_t(s.m_blocks) * ((p.NN + BN - 1) / BN);
  p.alpha = alpha; p.beta = beta; p.scale_mode = scale_mode; p.scale_len = scale_len > 0 ? scale_len : 1;
  p.cmid = nullptr; p.rmid = nullptr; p.dmm = 0.0; p.mid = -1; p.dbg = g_ws_enabled; p.trace = g_ws_trace;
  if (n & 1) {
    p.mid = (n - 1) / 2;
    p.cmid = packed + s.panels;
    p.rmid = p.cmid + int64_t(s.m_blocks) * BMH;
  }
  return launch_eo_ws(p, before == 1, map != nullptr, copy_map != nullptr, pt::as_stream(stream));
}

}  // namespace

extern "C" int pt_debug_set_eo_ws_trace(long long* device_buffer) {
  g_ws_trace = device_buffer;
  return PT_OK;
}

extern "C" int pt_debug_set_eo_ws(int mode) {      // 0 off, 1 default; 2 / 3: timing experiments (wrong results)
  g_ws_enabled = mode;
  return PT_OK;
}

extern "C" int64_t pt_packed_eo_ws_size(int n) {
  if (n < 2) return 0;
  const WsShape s = ws_shape(n);
  return s.panels + int64_t(s.m_blocks) * BMH + int64_t(s.nchunks) * BK + 1;
}

extern "C" int pt_operator_pack_eo_ws(const double* D, int n, double* packed, pt_stream_t stream) {
  PT_REQUIRE(D && packed, "pt_operator_pack_eo_ws: null pointer");
  PT_REQUIRE(n >= 2, "pt_operator_pack_eo_ws: n=%d", n);
  const WsShape s = ws_shape(n);
  int64_t blocks = (s.panels + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  pack_eo_ws_kernel<<<unsigned(blocks), 256, 0, pt::as_stream(stream)>>>(D, n, s.H, s.m_blocks, s.nchunks, packed);
  return pt::check_launch("pt_operator_pack_eo_ws");
}

extern "C" int pt_dense_apply_eo_ws(const double* packed_ws, int n, int sign, const double* U, double* Y,
                                    int64_t after, int64_t before, double alpha, double beta, const double* scale,
                                    int scale_mode, int64_t scale_len, pt_stream_t stream) {
  PT_REQUIRE(Y != nullptr, "pt_dense_apply_eo_ws: null pointer");
  PT_REQUIRE(U != Y, "pt_dense_apply_eo_ws: in-place application is not supported");
  return apply_ws_impl(packed_ws, n, sign, U, Y, Y, after, before, alpha, beta, scale, scale_mode, scale_len, nullptr,
                       nullptr, stream);
}

extern "C" int pt_dense_apply_eo_ws_scatter(const double* packed_ws, int n, int sign, const double* U, const double* C,
                                            int64_t after, int64_t before, double alpha, double beta,
                                            const pt_scatter_map* map, pt_stream_t stream) {
  PT_REQUIRE(map && map->base, "pt_dense_apply_eo_ws_scatter: null map");
  PT_REQUIRE(map->nranks >= 1 && map->chunk >= 1 && int64_t(map->chunk) * map->nranks >= n,
             "pt_dense_apply_eo_ws_scatter: %d destinations x %d rows do not cover n=%d", map->nranks, map->chunk, n);
  PT_REQUIRE(map->a_inner >= 1, "pt_dense_apply_eo_ws_scatter: a_inner must be >= 1");
  PT_REQUIRE(beta == 0.0 || C != nullptr, "pt_dense_apply_eo_ws_scatter: beta != 0 needs C");
  return apply_ws_impl(packed_ws, n, sign, U, nullptr, C ? C : U, after, before, alpha, beta, nullptr, PT_SCALE_NONE, 0,
                       map, nullptr, stream);
}

extern "C" int pt_dense_apply_eo_ws_scatter_copy(const double* packed_ws, int n, int sign, const double* U,
                                                 const double* C, int64_t after, int64_t before, double alpha,
                                                 double beta, const pt_scatter_map* map, const pt_scatter_map* field_map,
                                                 pt_stream_t stream) {
  PT_REQUIRE(map && map->base && field_map && field_map->base, "pt_dense_apply_eo_ws_scatter_copy: null map");
  for (const pt_scatter_map* m : {map, field_map}) {
    PT_REQUIRE(m->nranks >= 1 && m->chunk >= 1 && int64_t(m->chunk) * m->nranks >= n,
               "pt_dense_apply_eo_ws_scatter_copy: %d destinations x %d rows do not cover n=%d", m->nranks, m->chunk, n);
    PT_REQUIRE(m->a_inner >= 1, "pt_dense_apply_eo_ws_scatter_copy: a_inner must be >= 1");
  }
  PT_REQUIRE(beta == 0.0 || C != nullptr, "pt_dense_apply_eo_ws_scatter_copy: beta != 0 needs C");
  return apply_ws_impl(packed_ws, n, sign, U, nullptr, C ? C : U, after, before, alpha, beta, nullptr, PT_SCALE_NONE, 0,
                       map, field_map, stream);
}
