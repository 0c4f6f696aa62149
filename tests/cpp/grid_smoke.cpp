// Compiled-host check of the multi-GPU boundary WITHOUT Python: one process drives the GPUs of the box through the C ABI
// (mr_init_grid -> mr_dmatrix_*): blocks are routed to their owners, the multiply pulls peers' blocks over NVLink and runs on
// every GPU, results are read back block by block and checked against a plain triple loop written here.  Also exercises the
// element-wise and scalar operators, the NCCL reductions, the re-partitioning all-to-all and the transpose (id swap + re-placement).  Built (g++ only) and run by
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

static int fetch(mr_dmatrix* D, std::vector<double>& full, int n, int m, int blk, int want_transposed = 0) {
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
      if (d.type != 1 || (d.isTransposed != 0) != (want_transposed != 0)) {
        std::printf("FAIL block (%d,%d) type %d isT %d\n", i, j, (int)d.type, (int)d.isTransposed);
        return 1;
      }
      for (int cc = 0; cc < d.numCols; ++cc)   // DenseMatrix.index (MLMatrix.scala:285-289)
        for (int rr = 0; rr < d.numRows; ++rr)
          full[(size_t)(i * blk + rr) * m + (j * blk + cc)] = d.isTransposed ? vals[cc + (size_t)d.numCols * rr] : vals[rr + (size_t)d.numRows * cc];
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

  // scalar maps on the blocks every GPU owns: (C + 1.5) * 2 is reproducible bit for bit on the host
  {
    mr_dmatrix *P1 = nullptr, *P2 = nullptr, *P3 = nullptr;
    CK(mr_dmatrix_scalar(0, C, 1.5, &P1));
    CK(mr_dmatrix_scalar(1, P1, 2.0, &P2));
    CK(mr_dmatrix_scalar(2, P2, 2.0, &P3));
    std::vector<double> Pf, Qf;
    if (fetch(P2, Pf, n, m, blk) || fetch(P3, Qf, n, m, blk)) return 1;
    for (size_t t = 0; t < Pf.size(); ++t) {
      const double w = (Cf[t] + 1.5) * 2.0;
      if (Pf[t] != w) return std::printf("FAIL scalar ops at %zu\n", t), 1;
      if (!(std::fabs(Qf[t] - w * w) <= 1e-14 * w * w)) return std::printf("FAIL power at %zu\n", t), 1;
    }
    if (mr_dmatrix_scalar(7, C, 1.0, &P3) != MR_EINVAL) return std::printf("FAIL unknown scalar op accepted\n"), 1;
    CK(mr_dmatrix_free(P1));
    CK(mr_dmatrix_free(P2));
    CK(mr_dmatrix_free(P3));
  }

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
    CK(mr_dmatrix_free(R));
    CK(mr_dmatrix_free(Back));
    // rowSum / colSum: local line sums + one ncclAllReduce of a vector; result blocks (i, 0) / (0, j) at their owners
    {
      mr_dmatrix *Rs = nullptr, *Cs = nullptr;
      CK(mr_dmatrix_axis_sum(C, 0, &Rs));
      CK(mr_dmatrix_axis_sum(C, 1, &Cs));
      std::vector<double> rs, cs;
      if (fetch(Rs, rs, n, 1, blk) || fetch(Cs, cs, 1, m, blk)) return 1;
      for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int j = 0; j < m; ++j) acc += Cf[(size_t)i * m + j];
        if (!(std::fabs(acc - rs[i]) <= 1e-12 * scale * m)) return std::printf("FAIL rowSum at %d: %.17g vs %.17g\n", i, rs[i], acc), 1;
      }
      for (int j = 0; j < m; ++j) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += Cf[(size_t)i * m + j];
        if (!(std::fabs(acc - cs[j]) <= 1e-12 * scale * n)) return std::printf("FAIL colSum at %d: %.17g vs %.17g\n", j, cs[j], acc), 1;
      }
      int64_t nb = 0;
      CK(mr_dmatrix_num_blocks(Rs, &nb));
      if (nb != (n + blk - 1) / blk) return std::printf("FAIL rowSum block count %lld\n", (long long)nb), 1;
      CK(mr_dmatrix_owner(Rs, 3, 0, &owner));
      if (owner != (3 % pr) * pc) return std::printf("FAIL rowSum owner\n"), 1;
      mr_dmatrix* none = nullptr;
      if (mr_dmatrix_axis_sum(C, 2, &none) != MR_EINVAL) return std::printf("FAIL axis 2 accepted\n"), 1;
      CK(mr_dmatrix_free(Rs));
      CK(mr_dmatrix_free(Cs));
      // project / selection: row 200 (block row 2, offset 8), column m - 1 (the ragged last block column), one entry
      mr_dmatrix *Pr = nullptr, *Pc = nullptr, *Se = nullptr;
      CK(mr_dmatrix_project(C, 1, 200, &Pr));
      CK(mr_dmatrix_project(C, 0, m - 1, &Pc));
      CK(mr_dmatrix_selection(C, 301, 97, &Se));
      std::vector<double> prow, pcol, sel;
      if (fetch(Pr, prow, 1, m, blk) || fetch(Pc, pcol, n, 1, blk) || fetch(Se, sel, 1, 1, blk)) return 1;
      for (int j = 0; j < m; ++j)
        if (prow[j] != Cf[(size_t)200 * m + j]) return std::printf("FAIL project row at %d\n", j), 1;
      for (int i = 0; i < n; ++i)
        if (pcol[i] != Cf[(size_t)i * m + (m - 1)]) return std::printf("FAIL project column at %d\n", i), 1;
      if (sel[0] != Cf[(size_t)301 * m + 97]) return std::printf("FAIL selection\n"), 1;
      if (mr_dmatrix_project(C, 1, n, &none) != MR_EINVAL) return std::printf("FAIL row index n accepted: %s\n", mr_last_error()), 1;
      for (mr_dmatrix* h : {Pr, Pc, Se}) CK(mr_dmatrix_free(h));
    }
    // transpose: block (i, j) becomes block (j, i), moves to its new owner, payload untouched, isTransposed set
    mr_dmatrix *Ct = nullptr, *Ctt = nullptr, *At = nullptr, *G = nullptr;
    CK(mr_dmatrix_transpose(C, &Ct));
    int64_t tr = 0, tc = 0;
    CK(mr_dmatrix_dims(Ct, &tr, &tc, nullptr));
    if (tr != m || tc != n) return std::printf("FAIL transpose dims\n"), 1;
    CK(mr_dmatrix_owner(Ct, 2, 3, &owner));
    if (owner != (2 % pr) * pc + (3 % pc)) return std::printf("FAIL transpose owner\n"), 1;
    std::vector<double> Tf, TTf;
    if (fetch(Ct, Tf, m, n, blk, 1)) return 1;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j)
        if (Tf[(size_t)j * n + i] != Cf[(size_t)i * m + j]) return std::printf("FAIL transpose at (%d, %d)\n", i, j), 1;
    CK(mr_dmatrix_transpose(Ct, &Ctt));          // the row-major source is normalised on the device first
    if (fetch(Ctt, TTf, n, m, blk, 1)) return 1;
    for (size_t t = 0; t < Cf.size(); ++t)
      if (TTf[t] != Cf[t]) return std::printf("FAIL double transpose at %zu\n", t), 1;
    // a transposed operand in a product: A^T (k x n) times C (n x m)
    CK(mr_dmatrix_transpose(A, &At));
    CK(mr_dmatrix_multiply(At, C, &G));
    std::vector<double> Gf;
    if (fetch(G, Gf, k, m, blk)) return 1;
    double gworst = 0.0, gscale = 0.0;
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < m; ++j) {
        double acc = 0.0;
        for (int t = 0; t < n; ++t) acc += Af[(size_t)t * k + i] * Cf[(size_t)t * m + j];
        gworst = std::fmax(gworst, std::fabs(acc - Gf[(size_t)i * m + j]));
        gscale = std::fmax(gscale, std::fabs(acc));
      }
    std::printf("A^T C: max abs err %.3e (scale %.3e)\n", gworst, gscale);
    if (!(gworst <= 1e-11 * gscale)) return std::printf("FAIL product with a transposed operand\n"), 1;
    for (mr_dmatrix* h : {Ct, Ctt, At, G}) CK(mr_dmatrix_free(h));
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
