// matrel.hpp -- header-only C++17 mirror of the reference's Dataset operator API over the C ABI
// (include/matrel.h).  Method names and argument order follow
// /root/reference/src/main/scala/org/apache/spark/sql/matfast/Dataset.scala:57-152; `require`
// failures surface as matfast::IllegalArgumentException carrying the reference's message text.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "matrel.h"

namespace matfast {

struct IllegalArgumentException : std::invalid_argument {
  using std::invalid_argument::invalid_argument;
};
struct MatrelError : std::runtime_error {
  int code;
  MatrelError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void check(mr_status st) {
  if (st == MR_OK) return;
  const std::string msg = mr_last_error();
  if (st == MR_EINVAL || st == MR_EDIM) throw IllegalArgumentException(msg);
  throw MatrelError(st, msg);
}

// MLMatrix.scala:234-241 / :525-543 -- host containers for block ingest/egress
struct DenseMatrix {
  int32_t numRows = 0, numCols = 0;
  std::vector<double> values;  // column-major; row-major when isTransposed
  bool isTransposed = false;
};
struct SparseMatrix {
  int32_t numRows = 0, numCols = 0;
  std::vector<int32_t> colPtrs, rowIndices;
  std::vector<double> values;
  bool isTransposed = false;
};

class MatfastSession {
 public:
  explicit MatfastSession(int device = -1, bool compat_bugs = true, void* stream = nullptr) {
    mr_options o{};
    o.device = device;
    o.compat_bugs = compat_bugs ? 1 : 0;
    o.stream = stream;
    check(mr_init(&o, &ctx_));
  }
  ~MatfastSession() {
    if (ctx_) mr_shutdown(ctx_);
  }
  MatfastSession(const MatfastSession&) = delete;
  MatfastSession& operator=(const MatfastSession&) = delete;
  void sync() { check(mr_sync(ctx_)); }
  mr_context* raw() const { return ctx_; }

 private:
  mr_context* ctx_ = nullptr;
};

class Dataset {
 public:
  explicit Dataset(MatfastSession& s) : session_(&s) {
    mr_matrix* h = nullptr;
    check(mr_matrix_create(s.raw(), &h));
    h_.reset(h, [](mr_matrix* m) { mr_matrix_free(m); });
  }
  // Seq(MatrixBlock(rid, cid, matrix)).toDS()
  void putBlock(int32_t rid, int32_t cid, const DenseMatrix& m) {
    mr_block_desc d{};
    d.type = 1;
    d.numRows = m.numRows;
    d.numCols = m.numCols;
    d.values = const_cast<double*>(m.values.data());
    d.valuesLen = static_cast<int64_t>(m.values.size());
    d.isTransposed = m.isTransposed;
    check(mr_matrix_put_block(h_.get(), rid, cid, &d));
  }
  void putBlock(int32_t rid, int32_t cid, const SparseMatrix& m) {
    mr_block_desc d{};
    d.type = 0;
    d.numRows = m.numRows;
    d.numCols = m.numCols;
    d.colPtrs = const_cast<int32_t*>(m.colPtrs.data());
    d.colPtrsLen = static_cast<int64_t>(m.colPtrs.size());
    d.rowIndices = const_cast<int32_t*>(m.rowIndices.data());
    d.rowIndicesLen = static_cast<int64_t>(m.rowIndices.size());
    d.values = const_cast<double*>(m.values.data());
    d.valuesLen = static_cast<int64_t>(m.values.size());
    d.isTransposed = m.isTransposed;
    check(mr_matrix_put_block(h_.get(), rid, cid, &d));
  }
  bool hasBlock(int32_t rid, int32_t cid) const {
    int32_t present = 0;
    check(mr_matrix_has_block(h_.get(), rid, cid, &present));
    return present != 0;
  }
  std::vector<std::pair<int32_t, int32_t>> blockIds() const {
    int64_t n = 0;
    check(mr_matrix_num_blocks(h_.get(), &n));
    std::vector<int32_t> r(n), c(n);
    if (n) check(mr_matrix_block_ids(h_.get(), r.data(), c.data(), n));
    std::vector<std::pair<int32_t, int32_t>> out(n);
    for (int64_t i = 0; i < n; ++i) out[i] = {r[i], c[i]};
    return out;
  }
  // dense blocks only; sparse results are fetched with the two-call mr_matrix_get_block protocol
  DenseMatrix getDenseBlock(int32_t rid, int32_t cid) const {
    mr_block_desc d{};
    check(mr_matrix_get_block(h_.get(), rid, cid, &d));
    if (d.type != 1) throw MatrelError(MR_ENOTSUP, "block is sparse");
    DenseMatrix m;
    m.numRows = d.numRows;
    m.numCols = d.numCols;
    m.isTransposed = d.isTransposed != 0;
    m.values.resize(static_cast<size_t>(d.valuesLen));
    d.values = m.values.data();
    check(mr_matrix_get_block(h_.get(), rid, cid, &d));
    return m;
  }

  // ---- operators (Dataset.scala:57-152) -----------------------------------------------------
  Dataset matrixMultiply(int64_t leftRowNum, int64_t leftColNum, const Dataset& right, int64_t rightRowNum,
                         int64_t rightColNum, int32_t blkSize) const {
    return binary(mr_matrix_multiply, leftRowNum, leftColNum, right, rightRowNum, rightColNum, blkSize);
  }
  Dataset addElement(int64_t lr, int64_t lc, const Dataset& right, int64_t rr, int64_t rc, int32_t blkSize) const {
    return binary(mr_add_element, lr, lc, right, rr, rc, blkSize);
  }
  Dataset multiplyElement(int64_t lr, int64_t lc, const Dataset& right, int64_t rr, int64_t rc, int32_t blkSize) const {
    return binary(mr_multiply_element, lr, lc, right, rr, rc, blkSize);
  }
  Dataset divideElement(int64_t lr, int64_t lc, const Dataset& right, int64_t rr, int64_t rc, int32_t blkSize) const {
    return binary(mr_divide_element, lr, lc, right, rr, rc, blkSize);
  }
  Dataset matrixRankOneUpdate(int64_t lr, int64_t lc, const Dataset& right, int64_t rr, int64_t rc, int32_t blkSize) const {
    return binary(mr_rank_one_update, lr, lc, right, rr, rc, blkSize);
  }
  Dataset transpose() const {
    mr_matrix* o = nullptr;
    check(mr_transpose(h_.get(), &o));
    return Dataset(*session_, o);
  }
  Dataset t() const { return transpose(); }
  Dataset addScalar(double alpha) const { return unary(mr_add_scalar, alpha); }
  Dataset multiplyScalar(double alpha) const { return unary(mr_multiply_scalar, alpha); }
  Dataset power(double alpha) const { return unary(mr_power, alpha); }
  mr_matrix* raw() const { return h_.get(); }

 private:
  Dataset(MatfastSession& s, mr_matrix* h) : session_(&s), h_(h, [](mr_matrix* m) { mr_matrix_free(m); }) {}
  using BinFn = mr_status (*)(mr_matrix*, int64_t, int64_t, mr_matrix*, int64_t, int64_t, int32_t, mr_matrix**);
  using UnFn = mr_status (*)(mr_matrix*, double, mr_matrix**);
  Dataset binary(BinFn f, int64_t lr, int64_t lc, const Dataset& right, int64_t rr, int64_t rc, int32_t blk) const {
    mr_matrix* o = nullptr;
    check(f(h_.get(), lr, lc, right.h_.get(), rr, rc, blk, &o));
    return Dataset(*session_, o);
  }
  Dataset unary(UnFn f, double alpha) const {
    mr_matrix* o = nullptr;
    check(f(h_.get(), alpha, &o));
    return Dataset(*session_, o);
  }
  MatfastSession* session_;
  std::shared_ptr<mr_matrix> h_;
};

}  // namespace matfast
