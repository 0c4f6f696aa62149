// Compiled-host check of the multi-GPU boundary WITHOUT Python: one process drives the GPUs of the box through the C ABI
// (mr_init_grid -> mr_dmatrix_*): blocks are routed to their owners, the multiply pulls peers' blocks over NVLink and runs on
// every GPU, results are read back block by block and checked against a plain triple loop written here.  Also exercises the
// element-wise operator, the NCCL reductions and the re-partitioning all-to-all.  Built (g++ only) and run by
// tests/test_gpu_cpp_facade.py.  argv[1] = number of GPUs (default: min(2, visible)).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "matrel.h"

#define CK(expr)                                                                   \
  do {                                                                             \
    mr_status st_ = (expr);                                                        \
    if (st_ != MR_OK) {                                                            \
      std::printf("FAIL %s -> %d: %s\n", #expr, (int)st_, mr_last_error());        \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

static double urand(unsigned long long& s) {
  s = s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

// column-major blocks of an n x m matrix, blk-sized, put through the routed entry point
static int fill(mr_dmatrix* D, std::vector<double>& full, int n, int m, int blk, unsigned long long seed) {
  full.assign((size_t)n * m, 0.0);
  for (auto& v : full) v = urand(seed);
  for (int i = 0; i * blk < n; ++i)
    for (int j = 0; j * blk < m; ++j) {
      const int r = std::min(blk, n - i * blk), c = std::min(blk, m - j * blk);
      std::vector<double> vals((size_t)r * c);
      for (int cc = 0; cc < c; ++cc)
        for (int rr = 0; rr < r; ++rr) vals[rr + (size_t)r * cc] = full[(size_t)(i * blk + rr) * m + (j * blk + cc)];
      mr_block_desc d{};
      d.type = 1;
      d.numRows = r;
      d.numCols = c;
      d.values = vals.data();
      d.valuesLen = (int64_t)vals.size();
      CK(mr_dmatrix_put_block(D, i, j, &d));
      // pageable host memory: the copy is staged before the call returns, `vals` may go out of scope
    }
  return 0;
}

static int fetch(mr_dmatrix* D, std::vector<double>& full, int n, int m, int blk) {
  full.assign((size_t)n * m, 0.0);
  for (int i = 0; i * blk < n; ++i)
    for (int j = 0; j * blk < m; ++j) {
      int32_t has = 0;
      CK(mr_dmatrix_has_block(D, i, j, &has));
      if (!has) continue;
      mr_block_desc d{};
      CK(mr_dmatrix_get_block(D, i, j, &d));  // sizes
      std::vector<double> vals((size_t)d.valuesLen);
      d.values = vals.data();
      CK(mr_dmatrix_get_block(D, i, j, &d));
      if (d.type != 1 || d.isTransposed) {
        std::printf("FAIL block (%d,%d) type %d isT %d\n", i, j, (int)d.type, (int)d.isTransposed);
        return 1;
      }
      for (int cc = 0; cc < d.numCols; ++cc)
        for (int rr = 0; rr < d.numRows; ++rr) full[(size_t)(i * blk + rr) * m + (j * blk + cc)] = vals[rr + (size_t)d.numRows * cc];
    }
  return 0;
}

int main(int argc, char** argv) {
  int want_gpus = argc > 1 ? std::atoi(argv[1]) : 2;
  mr_grid* g = nullptr;
  mr_status st = mr_init_grid(nullptr, want_gpus, &g);
  if (st != MR_OK && want_gpus > 1) {  // fewer GPUs visible: the 1 x 1 grid runs the same code without pulls
    std::printf("note: %s; falling back to one GPU\n", mr_last_error());
    want_gpus = 1;
    CK(mr_init_grid(nullptr, 1, &g));
  } else if (st != MR_OK) {
    std::printf("FAIL mr_init_grid: %s\n", mr_last_error());
    return 1;
  }
  int32_t n_gpus = 0, pr = 0, pc = 0, has_nccl = 0;
  CK(mr_grid_info(g, &n_gpus, &pr, &pc, &has_nccl));
  std::printf("grid: %d GPU(s), %d x %d, nccl %d\n", n_gpus, pr, pc, has_nccl);

  const int n = 5 * 96 - 7, k = 4 * 96, m = 6 * 96 - 20, blk = 96;   // ragged edges, several blocks per rank
  mr_dmatrix *A = nullptr, *B = nullptr, *C = nullptr;
  CK(mr_dmatrix_create(g, n, k, blk, &A));
  CK(mr_dmatrix_create(g, k, m, blk, &B));
  std::vector<double> Af, Bf, Cf;
  if (fill(A, Af, n, k, blk, 1) || fill(B, Bf, k, m, blk, 2)) return 1;
  int32_t owner = -1;
  CK(mr_dmatrix_owner(A, 3, 2, &owner));
  if (owner != (3 % pr) * pc + (2 % pc)) return std::printf("FAIL owner\n"), 1;
  CK(mr_dmatrix_multiply(A, B, &C));
  if (fetch(C, Cf, n, m, blk)) return 1;
  double worst = 0.0, scale = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      double acc = 0.0;
      for (int t = 0; t < k; ++t) acc += Af[(size_t)i * k + t] * Bf[(size_t)t * m + j];
      worst = std::fmax(worst, std::fabs(acc - Cf[(size_t)i * m + j]));
      scale = std::fmax(scale, std::fabs(acc));
    }
  std::printf("multiply: max abs err %.3e (scale %.3e)\n", worst, scale);
  if (!(worst <= 1e-11 * scale)) return std::printf("FAIL multiply\n"), 1;

  // the reference's require message crosses the ABI
  mr_dmatrix* bad = nullptr;
  if (mr_dmatrix_multiply(B, B, &bad) != MR_EDIM) return std::printf("FAIL expected MR_EDIM, got: %s\n", mr_last_error()), 1;

  // element-wise on co-partitioned operands: C + C
  mr_dmatrix* S = nullptr;
  CK(mr_dmatrix_elementwise(0, C, C, &S));
  std::vector<double> Sf;
  if (fetch(S, Sf, n, m, blk)) return 1;
  for (size_t t = 0; t < Sf.size(); ++t)
    if (Sf[t] != 2.0 * Cf[t]) return std::printf("FAIL elementwise at %zu\n", t), 1;

  // sum through the local kernels + one ncclAllReduce
  double total = 0.0, want_total = 0.0;
  for (double v : Cf) want_total += v;
  st = mr_dmatrix_reduce_scalar(C, 0, &total);
  if (st == MR_ENCCL && n_gpus > 1 && !has_nccl) {
    std::printf("note: NCCL unavailable (%s); reductions skipped\n", mr_last_error());
  } else {
    CK(st);
    if (!(std::fabs(total - want_total) <= 1e-9 * std::fmax(1.0, std::fabs(want_total)) + 1e-9 * scale * 100)) {
      std::printf("FAIL sum %.17g vs %.17g\n", total, want_total);
      return 1;
    }
    // re-partition to the RowPartitioner layout (P x 1) and back: an all-to-all of whole blocks
    mr_dmatrix *R = nullptr, *Back = nullptr;
    CK(mr_dmatrix_repartition(C, n_gpus, 1, &R));
    CK(mr_dmatrix_repartition(R, pr, pc, &Back));
    std::vector<double> Rf, Bk;
    if (fetch(R, Rf, n, m, blk) || fetch(Back, Bk, n, m, blk)) return 1;
    for (size_t t = 0; t < Cf.size(); ++t)
      if (Rf[t] != Cf[t] || Bk[t] != Cf[t]) return std::printf("FAIL repartition at %zu\n", t), 1;
    // a product of re-partitioned operands: (A in row layout) is brought back to the grid by the multiply's caller
    CK(mr_dmatrix_free(R));
    CK(mr_dmatrix_free(Back));
  }
  CK(mr_grid_sync(g));
  CK(mr_dmatrix_free(S));
  CK(mr_dmatrix_free(C));
  CK(mr_dmatrix_free(B));
  CK(mr_dmatrix_free(A));
  CK(mr_grid_shutdown(g));
  std::printf("OK grid_smoke gpus=%d\n", n_gpus);
  return 0;
}
