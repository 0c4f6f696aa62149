// fp64 block-GEMM for sm_100a: C(i,j) = sum_k A(i,k) * B(k,j) over the present k-blocks.
//
// Replaces, in one kernel launch per matrixMultiply call:
//   - BLAS.gemmddd -> nativeBLAS.dgemm per block pair   (M/matrix/BLAS.scala:327-346)
//   - MLMatrix.multiply's fresh zeroed C per pair       (M/matrix/MLMatrix.scala:100-104)
//   - reduceByKey(LocalMatrix.add) over k               (M/execution/MatfastExecutionHelper.scala:255)
//
// Design (B200 has no fp64 kind in tcgen05; the fp64 tensor pipe is reached through
// mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4, measured 37.07 TFLOP/s, profiles/fp64_peaks_r01.jsonl):
//   - warp-specialised CTA: 1 producer warp + NCW consumer warps.
//   - producer feeds a STAGES-deep shared-memory ring with TMA bulk copies
//     (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, SASS UBLKCP); one copy
//     per contiguous tile line, landing in a padded layout so the DMMA fragment loads
//     (one 8-byte LDS per lane) are bank-conflict free.  full/empty mbarriers per stage.
//   - each operand tile is stored in one of two layouts, chosen per k-block from the block's
//     isTransposed flag (so "T"/"N" cost nothing and no block is ever re-materialised):
//        MODE_MN: s[k * (TILE+4) + mn]   (global lines contiguous along m / n)
//        MODE_K : s[mn * (BK+4)  + k ]   (global lines contiguous along k)
//   - ragged tiles (block edge, k tail, odd leading dimension) use the same ring: the producer
//     zero-fills what the bulk copies do not cover, or falls back to guarded element loads.
//   - consumers keep the whole K reduction (all k-blocks of the block row/column) in registers.
#include "kernels.h"

namespace matrel {
namespace {

constexpr int BK = 16;
constexpr int LDS_K = BK + 4;  // padded k-line: (LDS_K mod 16) == 4 -> conflict-free 8-byte fragment loads

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk (1-D) global -> shared copy, completion signalled on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int TILE>
struct OperandLayout {
  static constexpr int LDS_MN = TILE + 4;  // (TILE mod 16 == 0) -> (LDS_MN mod 16) == 4
  static constexpr int ELEMS = (TILE * LDS_K > BK * LDS_MN) ? TILE * LDS_K : BK * LDS_MN;
};

// Producer: fill one operand tile (TILE lines-or-columns x BK) of a stage.  Returns the number of
// bytes the bulk copies of THIS LANE will deliver (the caller warp-reduces it for expect_tx) and
// issues them after `issue` is true (two-pass: count first, arm the barrier, then copy).
//   g      : element (mn = 0, k = 0) of the tile inside the block's value array
//   ld     : leading dimension of the stored array
//   k_contig: true = MODE_K (global contiguous along k), false = MODE_MN
//   mv, kv : valid extent of the tile along mn / k   (1..TILE, 1..BK)
template <int TILE>
__device__ __forceinline__ uint32_t produce_operand(const double* __restrict__ g, int ld, bool k_contig, int mv, int kv,
                                                    double* s, uint32_t bar, int lane, bool issue) {
  using L = OperandLayout<TILE>;
  const int line_len = k_contig ? kv : mv;    // contiguous elements per global line
  const int nlines = k_contig ? mv : kv;
  const int sstride = k_contig ? LDS_K : L::LDS_MN;
  const bool fast = ((reinterpret_cast<uintptr_t>(g) & 15) == 0) && ((ld & 1) == 0) && ((line_len & 1) == 0);
  if (fast) {
    uint32_t bytes = 0;
    for (int l = lane; l < nlines; l += 32) {
      bytes += static_cast<uint32_t>(line_len) * 8u;
      if (issue) tma_bulk_g2s(smem_u32(s + l * sstride), g + static_cast<size_t>(l) * ld, line_len * 8u, bar);
    }
    if (!issue && (mv < TILE || kv < BK)) {
      // zero-fill the part of the tile the bulk copies do not touch
      if (k_contig) {
        for (int idx = lane; idx < TILE * BK; idx += 32) {
          const int mn = idx / BK, k = idx % BK;
          if (mn >= mv || k >= kv) s[mn * LDS_K + k] = 0.0;
        }
      } else {
        for (int idx = lane; idx < TILE * BK; idx += 32) {
          const int k = idx / TILE, mn = idx % TILE;
          if (mn >= mv || k >= kv) s[k * L::LDS_MN + mn] = 0.0;
        }
      }
    }
    return bytes;
  }
  if (!issue) {
    // generic path: guarded element loads (odd leading dimension / unaligned borrowed pointer)
    if (k_contig) {
#pragma unroll 4
      for (int idx = lane; idx < TILE * BK; idx += 32) {
        const int mn = idx / BK, k = idx % BK;
        double v = 0.0;
        if (mn < mv && k < kv) v = __ldg(g + static_cast<size_t>(mn) * ld + k);
        s[mn * LDS_K + k] = v;
      }
    } else {
#pragma unroll 4
      for (int idx = lane; idx < TILE * BK; idx += 32) {
        const int k = idx / TILE, mn = idx % TILE;
        double v = 0.0;
        if (mn < mv && k < kv) v = __ldg(g + static_cast<size_t>(k) * ld + mn);
        s[k * L::LDS_MN + mn] = v;
      }
    }
  }
  return 0;
}

// Consumer: one BK-deep stage for a warp tile of (MI*8) x (NJ*8).
template <int BM, int BN, int MI, int NJ, bool A_K, bool B_K>
__device__ __forceinline__ void consume_stage(const double* __restrict__ sA, const double* __restrict__ sB, int m_base,
                                              int n_base, int g, int t, double (&acc)[MI][NJ][2]) {
  constexpr int LDA = OperandLayout<BM>::LDS_MN;
  constexpr int LDB = OperandLayout<BN>::LDS_MN;
  // per-lane base pointers; every fragment is base + compile-time offset
  const double* pa = A_K ? (sA + (m_base + g) * LDS_K + t) : (sA + t * LDA + m_base + g);
  const double* pb = B_K ? (sB + (n_base + g) * LDS_K + t) : (sB + t * LDB + n_base + g);
#pragma unroll
  for (int ks = 0; ks < BK / 4; ++ks) {
    double a[MI], b[NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = A_K ? pa[i * 8 * LDS_K + ks * 4] : pa[ks * 4 * LDA + i * 8];
#pragma unroll
    for (int j = 0; j < NJ; ++j) b[j] = B_K ? pb[j * 8 * LDS_K + ks * 4] : pb[ks * 4 * LDB + j * 8];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

template <int BM, int BN, int WM, int WN, int STAGES>
struct GemmCfg {
  static constexpr int NCW = WM * WN;             // consumer warps
  static constexpr int THREADS = (NCW + 1) * 32;  // + 1 producer warp
  static constexpr int MI = BM / WM / 8;
  static constexpr int NJ = BN / WN / 8;
  static constexpr int A_ELEMS = OperandLayout<BM>::ELEMS;
  static constexpr int B_ELEMS = OperandLayout<BN>::ELEMS;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static constexpr size_t SMEM_BYTES = static_cast<size_t>(STAGES) * STAGE_ELEMS * 8 + 2 * STAGES * 8 + 16;
};

template <int BM, int BN, int WM, int WN, int STAGES>
__global__ void __launch_bounds__(GemmCfg<BM, BN, WM, WN, STAGES>::THREADS, 1)
    gemm_f64_dmma_kernel(const GemmOut* __restrict__ outs, const GemmPair* __restrict__ pairs,
                         const GemmTile* __restrict__ tiles) {
  using Cfg = GemmCfg<BM, BN, WM, WN, STAGES>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage_buf = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(STAGES) * Cfg::STAGE_ELEMS * 8);
  // bars[0..STAGES) = full, bars[STAGES..2*STAGES) = empty

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const GemmTile tile = tiles[blockIdx.x];
  const GemmOut out = outs[tile.out];
  const int m0 = tile.tm * BM, n0 = tile.tn * BN;
  const int mv = min(BM, out.m - m0), nv = min(BN, out.n - n0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[STAGES + s]), Cfg::NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == Cfg::NCW) {
    // ===================== producer warp =====================
    int it = 0;
    for (int p = 0; p < out.pair_count; ++p) {
      const GemmPair pr = pairs[out.pair_begin + p];
      const bool a_k = pr.aT != 0;   // row-major A block: lines contiguous along k
      const bool b_k = pr.bT == 0;   // column-major B block: lines contiguous along k
      for (int k0 = 0; k0 < pr.kdim; k0 += BK, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const int kv = min(BK, pr.kdim - k0);
        mbar_wait(smem_u32(&bars[STAGES + s]), ph ^ 1);
        double* sA = stage_buf + static_cast<size_t>(s) * Cfg::STAGE_ELEMS;
        double* sB = sA + Cfg::A_ELEMS;
        const double* gA = a_k ? pr.A + static_cast<size_t>(m0) * pr.lda + k0 : pr.A + static_cast<size_t>(k0) * pr.lda + m0;
        const double* gB = b_k ? pr.B + static_cast<size_t>(n0) * pr.ldb + k0 : pr.B + static_cast<size_t>(k0) * pr.ldb + n0;
        const uint32_t full = smem_u32(&bars[s]);
        // pass 1: manual stores / zero fill + byte count
        uint32_t bytes = produce_operand<BM>(gA, pr.lda, a_k, mv, kv, sA, full, lane, false) +
                         produce_operand<BN>(gB, pr.ldb, b_k, nv, kv, sB, full, lane, false);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(full, bytes);
        __syncwarp();
        // pass 2: issue the bulk copies
        produce_operand<BM>(gA, pr.lda, a_k, mv, kv, sA, full, lane, true);
        produce_operand<BN>(gB, pr.ldb, b_k, nv, kv, sB, full, lane, true);
      }
    }
  } else {
    // ===================== consumer warps =====================
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t = lane & 3;
    const int m_base = wm * (Cfg::MI * 8), n_base = wn * (Cfg::NJ * 8);
    double acc[Cfg::MI][Cfg::NJ][2];
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int it = 0;
    for (int p = 0; p < out.pair_count; ++p) {
      const GemmPair pr = pairs[out.pair_begin + p];
      const bool a_k = pr.aT != 0;
      const bool b_k = pr.bT == 0;
      const int nchunks = (pr.kdim + BK - 1) / BK;
      for (int c = 0; c < nchunks; ++c, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const double* sA = stage_buf + static_cast<size_t>(s) * Cfg::STAGE_ELEMS;
        const double* sB = sA + Cfg::A_ELEMS;
        mbar_wait(smem_u32(&bars[s]), ph);
        if (a_k) {
          if (b_k) consume_stage<BM, BN, Cfg::MI, Cfg::NJ, true, true>(sA, sB, m_base, n_base, g, t, acc);
          else     consume_stage<BM, BN, Cfg::MI, Cfg::NJ, true, false>(sA, sB, m_base, n_base, g, t, acc);
        } else {
          if (b_k) consume_stage<BM, BN, Cfg::MI, Cfg::NJ, false, true>(sA, sB, m_base, n_base, g, t, acc);
          else     consume_stage<BM, BN, Cfg::MI, Cfg::NJ, false, false>(sA, sB, m_base, n_base, g, t, acc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[STAGES + s]));
      }
    }

    // epilogue: C is column-major (ldc = out.m); lane holds rows g, columns 2t, 2t+1 of each 8x8
    double* C = out.C;
    const size_t ldc = static_cast<size_t>(out.m);
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i) {
      const int row = m0 + m_base + i * 8 + g;
      if (row < out.m) {
#pragma unroll
        for (int j = 0; j < Cfg::NJ; ++j) {
          const int col = n0 + n_base + j * 8 + 2 * t;
          if (col < out.n) C[row + ldc * col] = acc[i][j][0];
          if (col + 1 < out.n) C[row + ldc * (col + 1)] = acc[i][j][1];
        }
      }
    }
  }
}

template <int BM, int BN, int WM, int WN, int STAGES>
cudaError_t launch_variant(const GemmOut* d_outs, const GemmPair* d_pairs, const GemmTile* d_tiles, int ntiles,
                           cudaStream_t stream) {
  using Cfg = GemmCfg<BM, BN, WM, WN, STAGES>;
  auto kern = gemm_f64_dmma_kernel<BM, BN, WM, WN, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Cfg::SMEM_BYTES));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<ntiles, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(d_outs, d_pairs, d_tiles);
  return cudaGetLastError();
}

}  // namespace

int gemm_tile_m(int variant) { return variant == GEMM_64x64 ? 64 : 128; }
int gemm_tile_n(int variant) { return variant == GEMM_64x64 ? 64 : 128; }

cudaError_t launch_gemm_f64(const GemmOut* d_outs, const GemmPair* d_pairs, const GemmTile* d_tiles, int ntiles,
                            int variant, cudaStream_t stream) {
  if (ntiles <= 0) return cudaSuccess;
  if (variant == GEMM_64x64) return launch_variant<64, 64, 2, 2, 4>(d_outs, d_pairs, d_tiles, ntiles, stream);
  return launch_variant<128, 128, 2, 4, 4>(d_outs, d_pairs, d_tiles, ntiles, stream);
}

}  // namespace matrel
