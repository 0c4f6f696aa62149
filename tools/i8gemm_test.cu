// Standalone bring-up harness for the tcgen05 kind::i8 GEMM used by the Ozaki fp64 path:
//   C[M x N] (s32) = A[M x K] (s8, K contiguous) * B[N x K]^T (s8, K contiguous)
// TMA (2-D tensor maps, SWIZZLE_128B) -> 4-stage smem ring -> tcgen05.mma (one elected thread, accumulator in TMEM)
// -> tcgen05.ld epilogue.  Checks against a naive kernel and prints TOPS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/i8gemm_test tools/i8gemm_test.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int BM = 128, BN = 256, BKB = 128 /* bytes of K per stage */, UMMA_K = 32, STAGES = 4;
constexpr int A_BYTES = BM * BKB, B_BYTES = BN * BKB, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B operand tile: 8-row groups 1024 B apart, rows 128 B
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

#define TMEM_LD16(taddr, r)                                                                                      \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr))

template <int CS>
__global__ void __launch_bounds__(THREADS, 1) i8gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                            int32_t* __restrict__ C, int M, int N, int K) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);  // full[STAGES], empty[STAGES], acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;  // cluster along x = consecutive m tiles share one B tile
  const int nk = K / BKB;

  uint32_t cta_rank = 0;
  if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  if (threadIdx.x == 0) {
    // full: own producer's expect_tx arrival; empty: one tcgen05.commit arrival from every CTA of the cluster
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[STAGES + s]), CS); }
    mbar_init(smem_u32(&bars[2 * STAGES]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CS > 1) {  // peers' barriers must be initialised before anyone multicasts into them
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(smem_u32(&bars[STAGES + s]), ph ^ 1);
        const uint32_t full = smem_u32(&bars[s]);
        mbar_arrive_expect_tx(full, STAGE_BYTES);
        tma_2d(smem_u32(smem + s * STAGE_BYTES), &tmA, it * BKB, m0, full);
        if (CS == 1) {
          tma_2d(smem_u32(smem + s * STAGE_BYTES + A_BYTES), &tmB, it * BKB, n0, full);
        } else {  // this CTA fetches rows [rank * BN/CS, +BN/CS) of the B tile and multicasts them to the whole cluster
          const uint32_t dst = smem_u32(smem + s * STAGE_BYTES + A_BYTES + cta_rank * (B_BYTES / CS));
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
                       ::"r"(dst), "l"(&tmB), "r"(full), "h"(static_cast<uint16_t>((1u << CS) - 1)), "r"(it * BKB),
                       "r"(n0 + static_cast<int>(cta_rank) * (BN / CS)) : "memory");
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = S32 (2 @ bit 4), A,B = signed 8-bit (1 @ bits 7, 10), K-major both, N >> 3 @ 17, M >> 4 @ 24
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
      for (int it = 0; it < nk; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(smem_u32(&bars[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
        for (int k = 0; k < BKB / UMMA_K; ++k)
          umma_i8(tmem_base, umma_desc_k_sw128(a0 + k * UMMA_K), umma_desc_k_sw128(b0 + k * UMMA_K), idesc, (it | k) != 0);
        if (CS == 1) umma_commit(smem_u32(&bars[STAGES + s]));  // frees the stage when these MMAs retire
        else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                          ::"r"(smem_u32(&bars[STAGES + s])), "h"(static_cast<uint16_t>((1u << CS) - 1)) : "memory");
      }
      umma_commit(smem_u32(&bars[2 * STAGES]));    // accumulator complete
    }
  } else if (warp >= 4) {
    mbar_wait(smem_u32(&bars[2 * STAGES]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t r[16];
      TMEM_LD16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + c + j < N) C[static_cast<size_t>(row) * N + n0 + c + j] = static_cast<int32_t>(r[j]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CS > 1) {  // nobody leaves while a peer may still multicast into / arrive on this CTA's shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// ---- 2-SM variant: cta_group::2 MMA over a CTA pair (cluster 2x1): the pair computes a 256 x 256 tile, each CTA stages
// its own 128 rows of A and its own 128 columns of B (32 KB per stage instead of 48 KB) and owns 128 TMEM lanes. ----
constexpr int STAGES2 = 6;
constexpr int B2_BYTES = (BN / 2) * BKB, STAGE2_BYTES = A_BYTES + B2_BYTES;

__global__ void __launch_bounds__(THREADS, 1) i8gemm_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                int32_t* __restrict__ C, int M, int N, int K) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES);  // full[S], empty[S], acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int m0 = blockIdx.x * BM;              // blockIdx.x = 2 * pair + rank: this CTA's 128 rows
  const int n0 = blockIdx.y * BN;              // the pair's 256 columns
  const int nk = K / BKB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[STAGES2 + s]), 1); }
    mbar_init(smem_u32(&bars[2 * STAGES2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % STAGES2;
        const uint32_t ph = (it / STAGES2) & 1;
        mbar_wait(smem_u32(&bars[STAGES2 + s]), ph ^ 1);                      // own empty barrier
        const uint32_t full_leader = smem_u32(&bars[s]) & 0xFEFFFFFFu;          // the pair's full barrier lives in CTA 0
        if (leader) mbar_arrive_expect_tx(smem_u32(&bars[s]), 2 * STAGE2_BYTES);
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(smem + s * STAGE2_BYTES)), "l"(&tmA), "r"(full_leader), "r"(it * BKB), "r"(m0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(smem + s * STAGE2_BYTES + A_BYTES)), "l"(&tmB), "r"(full_leader), "r"(it * BKB),
                     "r"(n0 + static_cast<int>(cta_rank) * (BN / 2)) : "memory");
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // M = 256 across the pair (m_dim = 16), N = 256
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>((2 * BM) >> 4) << 24);
      for (int it = 0; it < nk; ++it) {
        const int s = it % STAGES2;
        const uint32_t ph = (it / STAGES2) & 1;
        mbar_wait(smem_u32(&bars[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(smem + s * STAGE2_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
        for (int k = 0; k < BKB / UMMA_K; ++k) {
          const uint64_t da = umma_desc_k_sw128(a0 + k * UMMA_K), db = umma_desc_k_sw128(b0 + k * UMMA_K);
          const uint32_t acc = (it | k) != 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bars[STAGES2 + s])), "h"(static_cast<uint16_t>(3)) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bars[2 * STAGES2])), "h"(static_cast<uint16_t>(3)) : "memory");
    }
  } else if (warp >= 4) {
    mbar_wait(smem_u32(&bars[2 * STAGES2]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t r[16];
      TMEM_LD16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + c + j < N) C[static_cast<size_t>(row) * N + n0 + c + j] = static_cast<int32_t>(r[j]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

__global__ void naive(const int8_t* A, const int8_t* B, int32_t* C, int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  int32_t s = 0;
  for (int k = 0; k < K; ++k) s += static_cast<int32_t>(A[static_cast<size_t>(m) * K + k]) * B[static_cast<size_t>(n) * K + k];
  C[static_cast<size_t>(m) * N + n] = s;
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn fn, void* base, uint64_t K, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {K, rows};
  cuuint64_t gstr[1] = {K};
  cuuint32_t box[2] = {BKB, box_rows};
  cuuint32_t est[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
  return m;
}

template <int CS>
static void launch(dim3 grid, size_t smem_bytes, const CUtensorMap& tA, const CUtensorMap& tB, int32_t* dC, int M, int N, int K) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, i8gemm_kernel<CS>, tA, tB, dC, M, N, K));
}

template <int CS>
static void run_cs(EncodeFn enc, size_t smem_bytes);
static void run_2sm(EncodeFn enc);

static int gM = 8192, gN = 8192, gK = 8192;
int main(int argc, char** argv) {
  if (argc > 3) { gM = atoi(argv[1]); gN = atoi(argv[2]); gK = atoi(argv[3]); }
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeFn enc = reinterpret_cast<EncodeFn>(p);
  const size_t smem_bytes = 1024 + STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16;
  run_cs<1>(enc, smem_bytes);
  run_cs<2>(enc, smem_bytes);
  run_2sm(enc);
  return 0;
}

template <int CS>
static void run_cs(EncodeFn enc, size_t smem_bytes) {
  CK(cudaFuncSetAttribute(i8gemm_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  for (int pass = 0; pass < 2; ++pass) {
    const int M = pass == 0 ? 512 : gM, N = pass == 0 ? 768 : gN, K = pass == 0 ? 640 : gK;
    std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
    uint32_t s = 12345;
    for (auto& v : hA) { s = s * 1664525u + 1013904223u; v = (int8_t)(s >> 24); }
    for (auto& v : hB) { s = s * 1664525u + 1013904223u; v = (int8_t)(s >> 24); }
    int8_t *dA, *dB; int32_t *dC, *dR;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&dR, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
    CUtensorMap tA = make_map(enc, dA, K, M, BM), tB = make_map(enc, dB, K, N, BN / CS);
    dim3 grid(M / BM, N / BN);
    launch<CS>(grid, smem_bytes, tA, tB, dC, M, N, K);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    if (pass == 0 || true) {
      naive<<<dim3((N + 127) / 128, M), 128>>>(dA, dB, dR, M, N, K);
      CK(cudaDeviceSynchronize());
      std::vector<int32_t> hC((size_t)M * N), hR((size_t)M * N);
      CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hR.data(), dR, hR.size() * 4, cudaMemcpyDeviceToHost));
      size_t bad = 0;
      for (size_t i = 0; i < hC.size(); ++i) if (hC[i] != hR[i]) { if (bad < 5) printf("mismatch at (%zu,%zu): got %d want %d\n", i / N, i % N, hC[i], hR[i]); ++bad; }
      printf("{\"test\": \"i8gemm\", \"cluster\": %d, \"M\": %d, \"N\": %d, \"K\": %d, \"mismatches\": %zu}\n", CS, M, N, K, bad);
    }
    if (pass == 1) {
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      for (int i = 0; i < 3; ++i) launch<CS>(grid, smem_bytes, tA, tB, dC, M, N, K);
      CK(cudaEventRecord(e0));
      const int reps = 20;
      for (int i = 0; i < reps; ++i) launch<CS>(grid, smem_bytes, tA, tB, dC, M, N, K);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
      printf("{\"bench\": \"i8gemm_tcgen05\", \"cluster\": %d, \"M\": %d, \"N\": %d, \"K\": %d, \"ms\": %.4f, \"tops\": %.1f}\n", CS, M, N, K, ms, 2.0 * M * N * K / ms * 1e-9);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR);
  }
}

static void run_2sm(EncodeFn enc) {
  const size_t smem_bytes = 1024 + STAGES2 * STAGE2_BYTES + (2 * STAGES2 + 1) * 8 + 16;
  CK(cudaFuncSetAttribute(i8gemm_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  for (int pass = 0; pass < 2; ++pass) {
    const int M = pass == 0 ? 512 : gM, N = pass == 0 ? 768 : gN, K = pass == 0 ? 640 : gK;
    std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
    uint32_t s = 777;
    for (auto& v : hA) { s = s * 1664525u + 1013904223u; v = (int8_t)(s >> 24); }
    for (auto& v : hB) { s = s * 1664525u + 1013904223u; v = (int8_t)(s >> 24); }
    int8_t *dA, *dB; int32_t *dC, *dR;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&dR, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
    CUtensorMap tA = make_map(enc, dA, K, M, BM), tB = make_map(enc, dB, K, N, BN / 2);
    dim3 grid(M / BM, N / BN);
    auto go = [&]() {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = 0;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, i8gemm_2sm_kernel, tA, tB, dC, M, N, K));
    };
    go();
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    naive<<<dim3((N + 127) / 128, M), 128>>>(dA, dB, dR, M, N, K);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> hC((size_t)M * N), hR((size_t)M * N);
    CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hR.data(), dR, hR.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < hC.size(); ++i) if (hC[i] != hR[i]) { if (bad < 5) printf("mismatch at (%zu,%zu): got %d want %d\n", i / N, i % N, hC[i], hR[i]); ++bad; }
    printf("{\"test\": \"i8gemm_2sm\", \"M\": %d, \"N\": %d, \"K\": %d, \"mismatches\": %zu}\n", M, N, K, bad);
    if (pass == 1) {
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      for (int i = 0; i < 3; ++i) go();
      CK(cudaEventRecord(e0));
      const int reps = 20;
      for (int i = 0; i < reps; ++i) go();
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
      printf("{\"bench\": \"i8gemm_tcgen05_2sm\", \"M\": %d, \"N\": %d, \"K\": %d, \"ms\": %.4f, \"tops\": %.1f}\n", M, N, K, ms, 2.0 * M * N * K / ms * 1e-9);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR);
  }
}
