// Measured tcgen05 tensor-pipe peaks on this B200 for the operand kinds the engine uses (BASELINE.md section 2 "builder measures"):
//   kind::i8   (Ozaki-II residue GEMMs)   -> int8 TOPS
//   kind::tf32 (3xTF32 fp32 path)          -> TF32 TFLOP/s
//   kind::f16  (cross-check against MEASURED_PEAKS.json's cuBLAS bf16 figure)
// One CTA per SM; one elected thread issues 128 x 256 x (32 bytes of K) MMAs back to back on operands that already sit in shared
// memory (random data, SWIZZLE_128B K-major tiles, no TMA traffic, two TMEM accumulators alternating), so the number is the
// ceiling of the MMA pipe itself with shared-memory operands under the board's power management.  "burst" = a ~30 ms launch on
// a cool chip, "sustained" = back-to-back launches for ~3 s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_peak tools/tc_peak.cu && ./tools/tc_peak > profiles/tc_peaks_r02.jsonl
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int BM = 128, BN = 256, KB = 128 /* bytes of K per operand tile row */, STAGES = 2;
constexpr int A_BYTES = BM * KB, B_BYTES = BN * KB;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else if (KIND == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

// KIND 0 = i8 (s32 acc), 1 = tf32 (f32 acc), 2 = bf16 (f32 acc)
template <int KIND>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, uint32_t seed) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  // random operand bytes (finite values for the float kinds: small-magnitude patterns)
  uint32_t x = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
  for (int i = threadIdx.x; i < STAGES * (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
    x = x * 1664525u + 1013904223u;
    uint32_t v = x;
    if (KIND == 1) v = (x & 0x007fe000u) | 0x3f000000u | (x & 0x80000000u);      // +-[0.5, 1) tf32
    if (KIND == 2) v = (x & 0x007f007fu) | 0x3f003f00u | (x & 0x80008000u);      // two bf16 in +-[0.5, 1)
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 32) {
    uint32_t idesc;
    if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10);        // D = s32, A/B = s8
    else if (KIND == 1) idesc = (1u << 4) | (2u << 7) | (2u << 10);   // D = f32, A/B = tf32
    else idesc = (1u << 4) | (1u << 7) | (1u << 10);                  // D = f32, A/B = bf16
    idesc |= (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
    // groups of 32 MMAs (8 "stages" of 4) per commit; two groups in flight (one per accumulator / barrier)
    uint32_t ph[2] = {0, 0};
    bool pending[2] = {false, false};
    for (int it = 0; it < iters; ++it) {
      const int b = it & 1;
      if (pending[b]) {
        mbar_wait(smem_u32(&bars[b]), ph[b]);
        ph[b] ^= 1;
        pending[b] = false;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t tacc = tmem_base + static_cast<uint32_t>(b * BN);
#pragma unroll 1
      for (int g = 0; g < 8; ++g) {
        const uint32_t a0 = smem_u32(smem + (g & 1) * (A_BYTES + B_BYTES)), b0 = a0 + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<KIND>(tacc, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc, (g | k) != 0);
      }
      umma_commit(smem_u32(&bars[b]));
      pending[b] = true;
    }
    for (int b = 0; b < 2; ++b)
      if (pending[b]) mbar_wait(smem_u32(&bars[b]), ph[b]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

template <int KIND>
void run(const char* name, const char* unit, int k_per_mma, int sms) {
  const size_t smem = 1024 + STAGES * (A_BYTES + B_BYTES) + 64;
  CK(cudaFuncSetAttribute(peak_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double ops_per_group = 32.0 * 2.0 * BM * BN * k_per_mma;  // per CTA per iteration
  // calibrate: ~30 ms launches
  int iters = 2000;
  peak_kernel<KIND><<<sms, 128, smem>>>(iters, 1u);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  peak_kernel<KIND><<<sms, 128, smem>>>(iters, 2u);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  iters = static_cast<int>(iters * 30.0 / ms);
  // burst: best of 5 after a 1 s pause
  double best = 0;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    peak_kernel<KIND><<<sms, 128, smem>>>(iters, 3u + r);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double t = ops_per_group * iters * sms / (ms * 1e-3) / 1e12;
    if (t > best) best = t;
  }
  printf("{\"bench\": \"%s_burst\", \"value\": %.1f, \"unit\": \"%s\", \"launch_ms\": %.2f, \"sms\": %d, \"mma\": \"128x256x%d cta_group::1, smem operands\"}\n",
         name, best, unit, ms, sms, k_per_mma);
  // sustained: back to back for ~3 s
  const auto t0 = std::chrono::steady_clock::now();
  int launches = 0;
  CK(cudaEventRecord(e0));
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 3.0) {
    for (int r = 0; r < 8; ++r) peak_kernel<KIND><<<sms, 128, smem>>>(iters, 11u + launches + r);
    launches += 8;
    CK(cudaStreamSynchronize(0));
  }
  // the last second only (clocks have settled)
  const int tail = 16;
  CK(cudaEventRecord(e0));
  for (int r = 0; r < tail; ++r) peak_kernel<KIND><<<sms, 128, smem>>>(iters, 99u + r);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("{\"bench\": \"%s_sustained\", \"value\": %.1f, \"unit\": \"%s\", \"window_ms\": %.1f, \"after_s\": 3.0, \"sms\": %d}\n", name,
         ops_per_group * iters * sms * tail / (ms * 1e-3) / 1e12, unit, ms, sms);
  fflush(stdout);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  run<0>("tcgen05_i8", "TOPS", 32, sms);
  run<1>("tcgen05_tf32", "TFLOP/s", 8, sms);
  run<2>("tcgen05_bf16", "TFLOP/s", 16, sms);
  return 0;
}
