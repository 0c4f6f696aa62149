// C-ABI host layer of the B200 block-matrix engine (see include/matrel.h).
//
// This file is the replacement for the reference's L2 physical operators and their helpers:
//   MatfastExecution.scala   (MatrixMatrixMultiplicationExecution :688-726, MatrixTransposeExecution
//                             :215-236, MatrixElement*Execution :571-686, MatrixScalar*/Power :465-532,
//                             RankOneUpdateExecution :728-747)
//   MatfastExecutionHelper.scala (matrixMultiplyGeneral :235-263, multiplyOuterProductDuplicate* :175-221,
//                             add/multiply/divideWithPartitioner :64-173, matrixRankOneUpdate :265-285,
//                             genBlockCyclicPartitioner :46-62)
//   LocalMatrix.scala        (matrixMultiplication dispatch :889-914 and the per-block kernels)
//   MLMatrixSerializer.scala (block <-> 7-field struct, :26-69)
// The Spark shuffles (groupByKey / join / reduceByKey / zipPartitions) become index arithmetic over a
// device-resident block table; every arithmetic step is a CUDA kernel (gemm_f64.cu, ew.cu).
// There is no CPU compute path in this file: host code only validates, builds descriptor tables
// and launches.
#include "../../include/matrel.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "kernels.h"

using namespace matrel;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_last_error;

struct MrError {
  mr_status code;
  std::string msg;
};

[[noreturn]] void fail(mr_status code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw MrError{code, buf};
}

// Scala `require(cond, msg)` -> IllegalArgumentException("requirement failed: " + msg)
#define MR_REQUIRE(cond, code, ...)                                           \
  do {                                                                        \
    if (!(cond)) {                                                            \
      char _b[900];                                                           \
      snprintf(_b, sizeof(_b), __VA_ARGS__);                                  \
      fail((code), "requirement failed: %s", _b);                             \
    }                                                                         \
  } while (0)

#define CUDA_CHECK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) fail(MR_ECUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(_e), __FILE__, __LINE__, \
                                cudaGetErrorString(_e));                                         \
  } while (0)

template <typename F>
mr_status guarded(F&& f) {
  try {
    f();
    return MR_OK;
  } catch (const MrError& e) {
    g_last_error = e.msg;
    return e.code;
  } catch (const std::bad_alloc&) {
    g_last_error = "out of host memory";
    return MR_ENOMEM;
  } catch (const std::exception& e) {
    g_last_error = std::string("internal error: ") + e.what();
    return MR_EINVAL;
  } catch (...) {
    g_last_error = "unknown internal error";
    return MR_EINVAL;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// context, device buffers, blocks
// ------------------------------------------------------------------------------------------------
struct mr_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int compat_bugs = 1;
  int gemm_algo = 0;
  int ozaki_slices = 0;
  int crt_moduli = 0;
  int time_kernels = 0;
  int force_variant = -1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_alloc = nullptr, ev_order = nullptr;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;  // ingest / egress overlap with compute on `stream`
  // The chunks of a pipelined multiply are independent launches: issued round-robin on these side streams, the last
  // (partial) wave of one chunk overlaps the first waves of the next instead of leaving SMs idle between launches.
  static constexpr int kChunkStreams = 3;
  cudaStream_t chunk_stream[kChunkStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t chunk_join[kChunkStreams] = {nullptr, nullptr, nullptr};
  uint64_t ingest_seq = 0;
  // mapped pinned staging ring for descriptor tables (see upload())
  char* stage_host = nullptr;
  char* stage_dev = nullptr;
  size_t stage_cap = 0, stage_off = 0;
  int pipeline = 1;
  mr_stats stats{};
  std::mutex mu;
};

namespace {

// Completion event of an asynchronous producer of a block (an H2D copy on the ingest stream, or one chunk of a
// chunked multiply); consumers on other streams wait on it instead of on whole streams.
struct Ready {
  cudaEvent_t ev = nullptr;
  Ready() { cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); }
  ~Ready() {
    if (ev) cudaEventDestroy(ev);
  }
  Ready(const Ready&) = delete;
  Ready& operator=(const Ready&) = delete;
};
using ReadyPtr = std::shared_ptr<Ready>;

struct DevBuf {
  mr_context* ctx;
  ReadyPtr ready;  // set when another stream writes this buffer: the free must be ordered after it
  void* p = nullptr;
  size_t bytes = 0;
  bool owned = true;
  DevBuf(mr_context* c, size_t n) : ctx(c), bytes(n) {
    if (n == 0) return;
    cudaError_t e = cudaMallocAsync(&p, n, ctx->stream);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      fail(e == cudaErrorMemoryAllocation ? MR_ENOMEM : MR_ECUDA, "cudaMallocAsync(%zu bytes) failed: %s", n,
           cudaGetErrorString(e));
    }
  }
  DevBuf(mr_context* c, void* borrowed, size_t n) : ctx(c), p(borrowed), bytes(n), owned(false) {}
  ~DevBuf() {
    if (owned && p) {
      if (ready) cudaStreamWaitEvent(ctx->stream, ready->ev, 0);
      cudaFreeAsync(p, ctx->stream);
    }
  }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};
using Buf = std::shared_ptr<DevBuf>;

struct Span {  // a typed window into a (possibly shared) device buffer
  Buf buf;
  size_t off = 0;  // bytes
  template <typename T>
  T* ptr() const {
    return buf ? reinterpret_cast<T*>(static_cast<char*>(buf->p) + off) : nullptr;
  }
};

struct Block {
  uint8_t type = 1;  // 0 sparse, 1 dense (MLMatrixSerializer.scala:31,40)
  int32_t numRows = 0, numCols = 0;
  bool isT = false;
  Span values;
  int64_t valuesLen = 0;
  Span colPtrs, rowIndices;  // sparse only
  int64_t colPtrsLen = 0;
  mutable ReadyPtr ready;    // producer still in flight on another stream (nullptr = ordered on the context stream);
                             // dropped (under the context mutex) once the event is seen complete
  uint64_t seq = 0;              // ingest order of the block's H2D copy (0 = not produced by the ingest stream)
  mutable bool settled = false;  // the dropped event had completed: the block's data needs no further ordering
  bool dense() const { return type == 1; }
};

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

}  // namespace

struct mr_matrix {
  mr_context* ctx;
  std::map<std::pair<int32_t, int32_t>, Block> blocks;
};

namespace {
bool wait_ready(mr_context* ctx, const Block& b);
bool wait_ready_on(cudaStream_t stream, const Block& b);
}

namespace {

void note_launch(mr_context* ctx, int n = 1) { ctx->stats.kernel_launches += n; }

// Order the context stream after the producer of a block.  Returns true when the producer was still running.
bool wait_ready(mr_context* ctx, const Block& b) {
  if (!b.ready) return false;
  const cudaError_t q = cudaEventQuery(b.ready->ev);
  if (q == cudaSuccess) {
    b.ready.reset();  // never query this block again
    b.settled = true;
    return false;
  }
  (void)cudaGetLastError();
  cudaStreamWaitEvent(ctx->stream, b.ready->ev, 0);
  return true;
}
bool wait_ready_on(cudaStream_t stream, const Block& b) {  // same, ordering `stream` instead of the context stream
  if (!b.ready) return false;
  const cudaError_t q = cudaEventQuery(b.ready->ev);
  if (q == cudaSuccess) {
    b.ready.reset();
    b.settled = true;
    return false;
  }
  (void)cudaGetLastError();
  cudaStreamWaitEvent(stream, b.ready->ev, 0);
  return true;
}
// true when the block's producer has finished (and forgets the event so it is not queried again)
bool block_done(const Block& b) {
  if (!b.ready) return true;
  if (cudaEventQuery(b.ready->ev) == cudaSuccess) {
    b.ready.reset();
    b.settled = true;
    return true;
  }
  (void)cudaGetLastError();
  return false;
}
bool wait_ready_all(mr_context* ctx, const mr_matrix* m) {
  bool any = false;
  for (auto& kv : m->blocks) any = wait_ready(ctx, kv.second) || any;
  return any;
}

// Descriptor tables (a few hundred KB per operator) travel through a mapped pinned staging ring and an SM-driven copy
// kernel instead of cudaMemcpyAsync: on the H2D copy engine they would queue behind every block upload already
// submitted on the ingest stream, and the first chunk of a pipelined multiply could not start until ALL operands had
// landed.  Ring regions are only reused after a stream synchronisation (on wrap-around).
Buf upload_bytes(mr_context* ctx, const void* data, size_t bytes) {
  Buf b = std::make_shared<DevBuf>(ctx, std::max<size_t>((bytes + 3) / 4 * 4, 16));
  if (bytes == 0) return b;
  const size_t need = align_up(bytes);
  if (ctx->stage_host != nullptr && need <= ctx->stage_cap / 2) {
    if (ctx->stage_off + need > ctx->stage_cap) {
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      ctx->stage_off = 0;
    }
    std::memcpy(ctx->stage_host + ctx->stage_off, data, bytes);
    CUDA_CHECK(launch_copy_words(b->p, ctx->stage_dev + ctx->stage_off, bytes, ctx->stream));
    ctx->stage_off += need;
  } else {
    CUDA_CHECK(cudaMemcpyAsync(b->p, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->stats.h2d_bytes += static_cast<int64_t>(bytes);
  return b;
}

template <typename T>
Buf upload(mr_context* ctx, const std::vector<T>& v) {
  return upload_bytes(ctx, v.data(), v.size() * sizeof(T));
}

// A slab allocator for operator results: one device allocation, blocks are windows into it.
struct Slab {
  Buf buf;
  size_t used = 0;
  Slab(mr_context* ctx, size_t total) : buf(std::make_shared<DevBuf>(ctx, std::max<size_t>(total, kAlign))) {}
  Span take(size_t bytes) {
    Span s{buf, used};
    used += align_up(bytes);
    return s;
  }
};

Span upload_raw(mr_context* ctx, const void* host, size_t bytes);

Block dense_block(int32_t rows, int32_t cols, Span values, bool isT = false) {
  Block b;
  b.type = 1;
  b.numRows = rows;
  b.numCols = cols;
  b.isT = isT;
  b.values = std::move(values);
  b.valuesLen = static_cast<int64_t>(rows) * cols;
  return b;
}

// SparseMatrix.toDense (MLMatrix.scala:669-671) on the device: zero fill + scatter.
Block densify(mr_context* ctx, const Block& s) {
  const size_t bytes = static_cast<size_t>(s.numRows) * s.numCols * sizeof(double);
  Span v{std::make_shared<DevBuf>(ctx, std::max<size_t>(bytes, 16)), 0};
  if (bytes) CUDA_CHECK(cudaMemsetAsync(v.ptr<double>(), 0, bytes, ctx->stream));
  if (s.valuesLen > 0) {
    CUDA_CHECK(launch_sparse_to_dense(s.colPtrs.ptr<int32_t>(), s.rowIndices.ptr<int32_t>(), s.values.ptr<double>(),
                                      s.isT, v.ptr<double>(), s.numRows, s.numCols, ctx->stream));
    note_launch(ctx);
  }
  return dense_block(s.numRows, s.numCols, v, false);
}

const char* type_name(const Block& b) { return b.dense() ? "DenseMatrix" : "SparseMatrix"; }

int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// block product planning (LocalMatrix.matrixMultiplication, LocalMatrix.scala:889-914)
// ------------------------------------------------------------------------------------------------
struct GemmSrc {
  const Block* a;
  const Block* b;
  int32_t k;
};

struct OutPlan {
  int32_t rid, cid;
  int32_t m = -1, n = -1;
  std::vector<GemmPair> gemm;                            // dense x dense (possibly densified) pairs
  std::vector<GemmSrc> src;                              // the blocks behind each gemm pair (+ its k-block id)
  std::vector<std::pair<const Block*, const Block*>> spmm;  // sparse x dense pairs
  std::vector<std::pair<const Block*, const Block*>> spsp;  // sparse x sparse, both densities <= 0.1 (ascending k)
};

struct MultiplyPlanner {
  mr_context* ctx;
  std::deque<Block> temps;   // densified sparse operands kept alive until the launches are enqueued (deque: the
                             // plans hold pointers to these blocks, so growth must not move them)
  std::map<const Block*, size_t> densified;

  const Block& dense_of(const Block& b) {
    if (b.dense()) return b;
    auto it = densified.find(&b);
    if (it == densified.end()) {
      temps.push_back(densify(ctx, b));
      it = densified.emplace(&b, temps.size() - 1).first;
    }
    return temps[it->second];
  }

  void add_pair(OutPlan& o, const Block& a, const Block& b, int32_t k) {
    // shape checks: BLAS.gemmddd / gemmsdd `require`s (BLAS.scala:338-343, 363-366)
    MR_REQUIRE(a.numCols == b.numRows, MR_EDIM, "The columns of A don't match the rows of B. A: %d, B: %d", a.numCols,
               b.numRows);
    if (o.m < 0) {
      o.m = a.numRows;
      o.n = b.numCols;
    } else {
      // LocalMatrix.add `require`s on the partial products (LocalMatrix.scala:36-41)
      MR_REQUIRE(o.m == a.numRows, MR_EDIM,
                 "Matrix A and B must have the same number of rows. But found A.numRows = %d, B.numRows = %d", o.m,
                 a.numRows);
      MR_REQUIRE(o.n == b.numCols, MR_EDIM,
                 "Matrix A and B must have the same number of cols. But found A.numCols = %d, B.numCols = %d", o.n,
                 b.numCols);
    }
    if (a.dense()) {
      push_gemm(o, a, dense_of(b), k);  // dense x dense, dense x sparse.toDense (:891-892)
    } else if (b.dense()) {
      o.spmm.emplace_back(&a, &b);   // sparse x dense (:893-899; the n == 1 SpMV case is the same kernel)
    } else {
      const double s1 = a.valuesLen * 1.0 / (static_cast<double>(a.numRows) * a.numCols);
      const double s2 = b.valuesLen * 1.0 / (static_cast<double>(b.numRows) * b.numCols);
      if (s1 > 0.1) {
        push_gemm(o, dense_of(a), dense_of(b), k);  // :903-904
      } else if (s2 > 0.1) {
        o.spmm.emplace_back(&a, &dense_of(b));   // :906-907
      } else {
        o.spsp.emplace_back(&a, &b);  // LocalMatrix.multiplySparseSparse (:909-911, :143-323): see run_sparse_chains
      }
    }
  }

  void push_gemm(OutPlan& o, const Block& a, const Block& b, int32_t k) {
    GemmPair p{};
    p.A = a.values.ptr<double>();
    p.B = b.values.ptr<double>();
    p.aT = a.isT;
    p.bT = b.isT;
    p.lda = a.isT ? a.numCols : a.numRows;  // BLAS.scala:335
    p.ldb = b.isT ? b.numCols : b.numRows;  // BLAS.scala:336
    p.kdim = a.numCols;
    p.tmA = p.tmB = -1;
    o.gemm.push_back(p);
    o.src.push_back(GemmSrc{&a, &b, k});
  }
};

// gemm_algo 2: the dense pairs of every output block through the tcgen05 int8 Ozaki pipeline (gemm_ozaki.cu).
// Returns false when the problem does not fit its regular-grid assumptions or holds Inf/NaN (caller uses DMMA).
bool try_ozaki(mr_context* ctx, std::vector<OutPlan>& plans, const std::vector<double*>& cptr, int32_t blkSize, int64_t M,
               int64_t K, int64_t N, bool outer) {
  const bool tf32 = ctx->gemm_algo == 3;
  const bool crt = ctx->gemm_algo == 4;
  const int S_eff = std::min(7, std::max(2, ctx->ozaki_slices > 0 ? ctx->ozaki_slices : 7));
  // s32 accumulator bound: up to S pairs x K terms of |digit product| <= 2^14 land in one accumulator
  if (M <= 0 || K <= 0 || N <= 0 || M > INT32_MAX || N > INT32_MAX || K > INT32_MAX / 2) return false;
  if (crt ? K >= (1 << 17) : (!tf32 && K * S_eff >= (1 << 17))) return false;
  // Compact the block rows / columns that actually have output blocks (a rank of the process grid owns every pr-th
  // block row and pc-th block column: slicing and multiplying the absent ones would only produce zeros).
  std::map<int32_t, int32_t> crow, ccol;
  for (const OutPlan& o : plans)
    if (!o.gemm.empty()) {
      crow.emplace(o.rid, 0);
      ccol.emplace(o.cid, 0);
    }
  if (crow.empty()) return false;
  {
    int32_t i = 0;
    for (auto& kv : crow) kv.second = i++;
    i = 0;
    for (auto& kv : ccol) kv.second = i++;
  }
  const int64_t nbr = static_cast<int64_t>(crow.size()), nbc = static_cast<int64_t>(ccol.size());
  if (nbr * nbc > (1 << 24)) return false;
  const int64_t last_r = crow.rbegin()->first, last_c = ccol.rbegin()->first;
  if (last_r * static_cast<int64_t>(blkSize) >= M || last_c * static_cast<int64_t>(blkSize) >= N) return false;
  const int64_t Mc = (nbr - 1) * blkSize + std::min<int64_t>(blkSize, M - last_r * blkSize);
  const int64_t Nc = (nbc - 1) * blkSize + std::min<int64_t>(blkSize, N - last_c * blkSize);
  std::vector<double*> ctab(static_cast<size_t>(nbr * nbc), nullptr);
  std::map<const Block*, int> ia, ib;
  std::vector<OzakiOperand> va, vb;
  for (size_t i = 0; i < plans.size(); ++i) {
    const OutPlan& o = plans[i];
    if (o.gemm.empty()) continue;
    const int32_t cr = crow[o.rid], cc = ccol[o.cid];
    // every block row / column but the last must be a full blkSize tall / wide (the kernel finds blocks by division)
    if (o.m != ((o.rid == last_r) ? Mc - (nbr - 1) * blkSize : blkSize) || o.n != ((o.cid == last_c) ? Nc - (nbc - 1) * blkSize : blkSize))
      return false;
    if (!o.spmm.empty()) return false;  // mixed dense / sparse partial sums stay on the exact path
    ctab[static_cast<size_t>(cr) * nbc + cc] = cptr[i];
    for (const GemmSrc& g : o.src) {
      const int64_t k0 = outer ? 0 : static_cast<int64_t>(g.k) * blkSize;
      if (k0 + g.a->numCols > K || g.a->numCols != g.b->numRows) return false;
      if (!ia.count(g.a)) {
        ia[g.a] = 1;
        va.push_back(OzakiOperand{g.a->values.ptr<double>(), g.a->numRows, g.a->numCols, cr * blkSize, static_cast<int32_t>(k0),
                                  static_cast<uint8_t>(g.a->isT)});
      }
      if (!ib.count(g.b)) {
        ib[g.b] = 1;
        vb.push_back(OzakiOperand{g.b->values.ptr<double>(), g.b->numRows, g.b->numCols, static_cast<int32_t>(k0), cc * blkSize,
                                  static_cast<uint8_t>(g.b->isT)});
      }
    }
  }
  if (va.empty() || vb.empty()) return false;
  int launches = 0, nonfinite = 0;
  if (ctx->time_kernels) CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
  if (tf32)
    CUDA_CHECK(tf32x3_gemm(va.data(), static_cast<int>(va.size()), vb.data(), static_cast<int>(vb.size()), Mc, K, Nc, ctab.data(),
                           blkSize, static_cast<int>(nbr), static_cast<int>(nbc), &launches, ctx->stream));
  else if (crt)
    CUDA_CHECK(ozaki2_gemm_f64(va.data(), static_cast<int>(va.size()), vb.data(), static_cast<int>(vb.size()), Mc, K, Nc,
                               ctx->crt_moduli > 0 ? ctx->crt_moduli : 16, ctab.data(), blkSize, static_cast<int>(nbr),
                               static_cast<int>(nbc), &launches, &nonfinite, ctx->stream));
  else
    CUDA_CHECK(ozaki_gemm_f64(va.data(), static_cast<int>(va.size()), vb.data(), static_cast<int>(vb.size()), Mc, K, Nc,
                              ctx->ozaki_slices > 0 ? ctx->ozaki_slices : 7, ctab.data(), blkSize, static_cast<int>(nbr),
                              static_cast<int>(nbc), false, &launches, &nonfinite, ctx->stream));
  note_launch(ctx, launches);
  if (nonfinite) return false;
  ctx->stats.gemm_launches += 1;
  if (ctx->time_kernels) {
    CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.last_gemm_ms = ms;
    ctx->stats.gemm_ms_total += ms;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// dense window -> CSC compaction (DenseMatrix.toSparse, MLMatrix.scala:392-420), batched
// ------------------------------------------------------------------------------------------------
struct DenseWin {
  const double* p;  // column-major rows x cols
  int32_t rows, cols;
};

// Per-column counts of entries != 0.0 (NaN counts, as in the reference's `arr(i) != 0`): one launch + one sync.
std::vector<std::vector<int32_t>> column_counts(mr_context* ctx, const std::vector<DenseWin>& wins) {
  std::vector<std::vector<int32_t>> out(wins.size());
  if (wins.empty()) return out;
  size_t ncols_total = 0;
  int maxc = 0;
  for (auto& w : wins) {
    ncols_total += static_cast<size_t>(w.cols);
    maxc = std::max(maxc, w.cols);
  }
  Buf counts = std::make_shared<DevBuf>(ctx, std::max<size_t>(ncols_total * sizeof(int32_t), 16));
  std::vector<CscDesc> cd(wins.size());
  size_t off = 0;
  for (size_t i = 0; i < wins.size(); ++i) {
    cd[i] = CscDesc{wins[i].p, wins[i].rows, wins[i].cols, static_cast<int32_t*>(counts->p) + off, nullptr, nullptr, nullptr};
    off += static_cast<size_t>(wins[i].cols);
  }
  Buf dcd = upload(ctx, cd);
  CUDA_CHECK(launch_csc_count(static_cast<const CscDesc*>(dcd->p), static_cast<int>(cd.size()), maxc, ctx->stream));
  note_launch(ctx);
  std::vector<int32_t> h(ncols_total);
  if (ncols_total) {
    CUDA_CHECK(cudaMemcpyAsync(h.data(), counts->p, ncols_total * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += static_cast<int64_t>(ncols_total * sizeof(int32_t));
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  off = 0;
  for (size_t i = 0; i < wins.size(); ++i) {
    out[i].assign(h.begin() + static_cast<std::ptrdiff_t>(off), h.begin() + static_cast<std::ptrdiff_t>(off + wins[i].cols));
    off += static_cast<size_t>(wins[i].cols);
  }
  return out;
}

int64_t total_count(const std::vector<int32_t>& counts) {
  int64_t n = 0;
  for (int32_t c : counts) n += c;
  return n;
}

// CSC blocks (isTransposed = false) of the windows, given their column counts: one launch.
std::vector<Block> compact_csc(mr_context* ctx, const std::vector<DenseWin>& wins, const std::vector<std::vector<int32_t>>& counts) {
  std::vector<Block> out;
  std::vector<CscDesc> fill;
  int maxc = 0;
  for (size_t i = 0; i < wins.size(); ++i) {
    const int rows = wins[i].rows, cols = wins[i].cols;
    std::vector<int32_t> ptrs(static_cast<size_t>(cols) + 1, 0);
    for (int c = 0; c < cols; ++c) ptrs[c + 1] = ptrs[c] + counts[i][c];
    const int64_t nnz = ptrs[cols];
    Block sb;
    sb.type = 0;
    sb.numRows = rows;
    sb.numCols = cols;
    sb.isT = false;
    sb.valuesLen = nnz;
    sb.colPtrsLen = cols + 1;
    sb.colPtrs = upload_raw(ctx, ptrs.data(), ptrs.size() * sizeof(int32_t));
    sb.rowIndices = Span{std::make_shared<DevBuf>(ctx, std::max<size_t>(nnz * sizeof(int32_t), 16)), 0};
    sb.values = Span{std::make_shared<DevBuf>(ctx, std::max<size_t>(nnz * sizeof(double), 16)), 0};
    fill.push_back(CscDesc{wins[i].p, rows, cols, nullptr, sb.colPtrs.ptr<int32_t>(), sb.rowIndices.ptr<int32_t>(),
                           sb.values.ptr<double>()});
    maxc = std::max(maxc, cols);
    out.push_back(std::move(sb));
  }
  if (!fill.empty()) {
    Buf dfill = upload(ctx, fill);
    CUDA_CHECK(launch_csc_fill(static_cast<const CscDesc*>(dfill->p), static_cast<int>(fill.size()), maxc, ctx->stream));
    note_launch(ctx);
  }
  return out;
}

// ------------------------------------------------------------------------------------------------
// Output blocks whose partial products are ALL low-density sparse x sparse (LocalMatrix.multiplySparseSparse,
// LocalMatrix.scala:143-323).  Values: each partial is computed as sparse A x densified B (the same sums in a different
// order).  Storage format: the reference's four loop nests end in four different rules, and `reduceByKey(LocalMatrix.add)`
// re-decides the format at every sparse + sparse step, so the chain is replayed partial by partial in ascending k (Spark's
// own reduce order is arbitrary; ascending k is the deterministic choice):
//   CSC x CSC (:155-196): CSC iff rows*cols > 2 nnz + cols + 1, else dense      CSR x CSR (:198-239): CSR iff ... + rows + 1
//   CSR x CSC (:241-286): always CSC (both branches build a SparseMatrix)       CSC x CSR (:288-323): dense iff rows*cols <= 2 nnz + cols
//   sparse + sparse (:74-139): CSC iff rows*cols > 2 nnz + cols + 1, else dense; anything + dense: dense.
// ------------------------------------------------------------------------------------------------
enum ChainFmt { FMT_DENSE = 0, FMT_CSC = 1, FMT_CSR = 2 };

void run_sparse_chains(mr_context* ctx, std::vector<OutPlan>& plans, const std::vector<size_t>& chains,
                       const std::vector<double*>& cptr, MultiplyPlanner& planner, mr_matrix* result) {
  if (chains.empty()) return;
  size_t levels = 0;
  for (size_t i : chains) levels = std::max(levels, plans[i].spsp.size());
  std::vector<int> fmt(chains.size(), FMT_DENSE);
  std::vector<Buf> scratch(chains.size());
  for (size_t lv = 0; lv < levels; ++lv) {
    std::vector<size_t> act;  // chains that have a partial at this level
    std::vector<DenseWin> pw;
    for (size_t c = 0; c < chains.size(); ++c) {
      OutPlan& o = plans[chains[c]];
      if (lv >= o.spsp.size() || o.m == 0 || o.n == 0) continue;
      const Block& a = *o.spsp[lv].first;
      const Block& b = planner.dense_of(*o.spsp[lv].second);
      wait_ready(ctx, a);
      wait_ready(ctx, *o.spsp[lv].second);
      const size_t bytes = static_cast<size_t>(o.m) * o.n * sizeof(double);
      double* target = cptr[chains[c]];
      if (lv > 0) {
        if (!scratch[c]) scratch[c] = std::make_shared<DevBuf>(ctx, bytes);
        target = static_cast<double*>(scratch[c]->p);
      }
      if (!a.isT) CUDA_CHECK(cudaMemsetAsync(target, 0, bytes, ctx->stream));  // the CSC kernel scatters into zeros
      CUDA_CHECK(launch_spmm(a.colPtrs.ptr<int32_t>(), a.rowIndices.ptr<int32_t>(), a.values.ptr<double>(), a.isT,
                             b.values.ptr<double>(), b.isT, target, a.numRows, a.numCols, b.numCols, false, ctx->stream));
      note_launch(ctx);
      act.push_back(c);
      pw.push_back(DenseWin{target, o.m, o.n});
    }
    if (act.empty()) continue;
    const auto pcounts = column_counts(ctx, pw);
    std::vector<size_t> recount;  // chains whose running sum is sparse + sparse at this level
    std::vector<EwDesc> adds;
    int max_rows = 0, max_cols = 0;
    for (size_t t = 0; t < act.size(); ++t) {
      const size_t c = act[t];
      OutPlan& o = plans[chains[c]];
      const bool aT = o.spsp[lv].first->isT, bT = o.spsp[lv].second->isT;
      const int64_t cells = static_cast<int64_t>(o.m) * o.n, nnz = total_count(pcounts[t]);
      int pf;
      if (!aT && !bT) pf = cells > 2 * nnz + o.n + 1 ? FMT_CSC : FMT_DENSE;
      else if (aT && bT) pf = cells > 2 * nnz + o.m + 1 ? FMT_CSR : FMT_DENSE;
      else if (aT && !bT) pf = FMT_CSC;
      else pf = cells <= 2 * nnz + o.n ? FMT_DENSE : FMT_CSC;
      if (lv == 0) {
        fmt[c] = pf;
        continue;
      }
      EwDesc d{};
      d.A = cptr[chains[c]];
      d.B = pw[t].p;
      d.C = cptr[chains[c]];  // in place: every element is read and written by the same thread
      d.rows = o.m;
      d.cols = o.n;
      adds.push_back(d);
      max_rows = std::max(max_rows, o.m);
      max_cols = std::max(max_cols, o.n);
      if (fmt[c] == FMT_DENSE || pf == FMT_DENSE) fmt[c] = FMT_DENSE;
      else recount.push_back(c);
    }
    if (!adds.empty()) {
      Buf d = upload(ctx, adds);
      CUDA_CHECK(launch_ew_batched(EW_ADD, static_cast<const EwDesc*>(d->p), static_cast<int>(adds.size()), max_rows, max_cols,
                                   false, ctx->stream));
      note_launch(ctx);
    }
    if (!recount.empty()) {
      std::vector<DenseWin> sw;
      for (size_t c : recount) sw.push_back(DenseWin{cptr[chains[c]], plans[chains[c]].m, plans[chains[c]].n});
      const auto scounts = column_counts(ctx, sw);
      for (size_t t = 0; t < recount.size(); ++t) {
        const OutPlan& o = plans[chains[recount[t]]];
        fmt[recount[t]] = static_cast<int64_t>(o.m) * o.n > 2 * total_count(scounts[t]) + o.n + 1 ? FMT_CSC : FMT_DENSE;
      }
    }
  }
  // final storage: dense results stay in their slab window; CSC / CSR results are compacted out of it
  std::vector<DenseWin> cw;
  std::vector<size_t> cw_chain;
  std::vector<Buf> keep;
  for (size_t c = 0; c < chains.size(); ++c) {
    const OutPlan& o = plans[chains[c]];
    if (o.m == 0 || o.n == 0 || fmt[c] == FMT_DENSE) continue;
    if (fmt[c] == FMT_CSC) {
      cw.push_back(DenseWin{cptr[chains[c]], o.m, o.n});
    } else {  // CSR of C = CSC of C^T: materialise the row-major copy (= column-major n x m) first
      Buf t = std::make_shared<DevBuf>(ctx, static_cast<size_t>(o.m) * o.n * sizeof(double));
      EwDesc d{};
      d.A = cptr[chains[c]];
      d.C = static_cast<double*>(t->p);
      d.rows = o.n;
      d.cols = o.m;
      d.aT = 1;
      std::vector<EwDesc> one{d};
      Buf dd = upload(ctx, one);
      CUDA_CHECK(launch_ew_batched(EW_COPY, static_cast<const EwDesc*>(dd->p), 1, o.n, o.m, true, ctx->stream));
      note_launch(ctx);
      cw.push_back(DenseWin{static_cast<const double*>(t->p), o.n, o.m});
      keep.push_back(t);
    }
    cw_chain.push_back(c);
  }
  if (!cw.empty()) {
    const auto counts = column_counts(ctx, cw);
    std::vector<Block> blocks = compact_csc(ctx, cw, counts);
    for (size_t t = 0; t < blocks.size(); ++t) {
      const size_t c = cw_chain[t];
      const OutPlan& o = plans[chains[c]];
      Block& b = blocks[t];
      if (fmt[c] == FMT_CSR) {  // the CSC arrays of C^T are the CSR arrays of C
        b.numRows = o.m;
        b.numCols = o.n;
        b.isT = true;
      }
      result->blocks[{o.rid, o.cid}] = std::move(b);
    }
  }
}

void run_multiply(mr_context* ctx, std::vector<OutPlan>& plans, MultiplyPlanner& planner, int32_t blkSize,
                  mr_matrix* result, int64_t M, int64_t K, int64_t N, bool outer) {
  // Low-density sparse x sparse pairs: next to any dense partial the block sum is dense whatever the partial's own format
  // (LocalMatrix.add), so there they are ordinary sparse x dense products of the densified right operand; an output block
  // made of such pairs ONLY replays the reference's format rules (run_sparse_chains).
  std::vector<size_t> chains;
  for (size_t i = 0; i < plans.size(); ++i) {
    OutPlan& o = plans[i];
    if (o.spsp.empty()) continue;
    if (o.gemm.empty() && o.spmm.empty()) {
      chains.push_back(i);
    } else {
      for (auto& pr : o.spsp) o.spmm.emplace_back(pr.first, &planner.dense_of(*pr.second));
      o.spsp.clear();
    }
  }
  // allocate all output blocks from one slab
  size_t total = 0;
  for (auto& o : plans) total += align_up(static_cast<size_t>(o.m) * o.n * sizeof(double));
  Slab slab(ctx, total);
  std::vector<double*> cptr(plans.size());
  for (size_t i = 0; i < plans.size(); ++i) {
    auto& o = plans[i];
    Span s = slab.take(static_cast<size_t>(o.m) * o.n * sizeof(double));
    cptr[i] = s.ptr<double>();
    result->blocks[{o.rid, o.cid}] = dense_block(o.m, o.n, s, false);  // product is never transposed (MLMatrix.scala:101)
  }

  // ---- fused GEMM launch over every output block that has dense pairs
  std::vector<GemmOut> outs;
  std::vector<GemmPair> pairs;
  std::vector<size_t> out_plan;
  int64_t flops = 0;
  for (size_t i = 0; i < plans.size(); ++i) {
    auto& o = plans[i];
    if (o.gemm.empty() || o.m == 0 || o.n == 0) continue;
    GemmOut go{};
    go.C = cptr[i];
    go.m = o.m;
    go.n = o.n;
    go.pair_begin = static_cast<int32_t>(pairs.size());
    go.pair_count = static_cast<int32_t>(o.gemm.size());
    for (auto& p : o.gemm) {
      pairs.push_back(p);
      flops += 2ll * o.m * o.n * p.kdim;
    }
    outs.push_back(go);
    out_plan.push_back(i);
  }
  // which dense operands are still being produced on another stream (ingest copies)?
  bool pending = false;
  for (size_t i = 0; i < plans.size() && !pending; ++i)
    for (const GemmSrc& g : plans[i].src)
      if ((g.a->ready && !block_done(*g.a)) || (g.b->ready && !block_done(*g.b))) {
        pending = true;
        break;
      }
  (void)cudaGetLastError();
  auto wait_all_sources = [&] {
    for (auto& o : plans)
      for (const GemmSrc& g : o.src) {
        wait_ready(ctx, *g.a);
        wait_ready(ctx, *g.b);
      }
  };
  bool ozaki_done = false;
  if (!outs.empty() && (ctx->gemm_algo >= 2 && ctx->gemm_algo <= 4)) {
    wait_all_sources();
    ozaki_done = try_ozaki(ctx, plans, cptr, blkSize, M, K, N, outer);
    if (ozaki_done) ctx->stats.last_gemm_flops = flops;
  }
  if (!outs.empty() && !ozaki_done) {
    // tile shape: large tiles unless they cannot fill the 148 SMs
    int64_t tiles128 = 0;
    for (auto& go : outs) tiles128 += static_cast<int64_t>((go.m + 127) / 128) * ((go.n + 127) / 128);
    int variant = tiles128 >= 2 * 148 ? GEMM_128x128 : GEMM_64x64;
    if (ctx->force_variant >= 0) variant = ctx->force_variant;
    const int BM = gemm_tile_m(variant), BN = gemm_tile_n(variant);
    const int tpbm = std::max(1, (blkSize + BM - 1) / BM), tpbn = std::max(1, (blkSize + BN - 1) / BN);
    struct Keyed {
      int64_t band, gm, gn;
      GemmTile t;
    };
    std::vector<Keyed> keyed;
    constexpr int kBand = 12;  // ~12 x 12 tiles resident across 148 SMs share A row- and B column-panels in L2
    for (size_t oi = 0; oi < outs.size(); ++oi) {
      const OutPlan& o = plans[out_plan[oi]];
      const int tm = (outs[oi].m + BM - 1) / BM, tn = (outs[oi].n + BN - 1) / BN;
      for (int a = 0; a < tm; ++a)
        for (int b = 0; b < tn; ++b) {
          const int64_t gm = static_cast<int64_t>(o.rid) * tpbm + a, gn = static_cast<int64_t>(o.cid) * tpbn + b;
          keyed.push_back({gn / kBand, gm, gn, GemmTile{static_cast<int32_t>(oi), a, b, 0}});
        }
    }
    // Chunked launch when operands are still arriving from the host: one launch per block row of C, each waiting
    // only for the A row panel it reads (and all of B), so compute overlaps the remaining ingest and the egress of
    // finished rows.  With resident operands it is ONE launch.
    // Chunk key of an output block = ingest sequence number of the LAST operand block it needs that is still in flight
    // (the ingest stream is in order, so waiting for that one copy implies all earlier ones).  Output blocks are
    // launched in key order, in chunks of >= ~8 waves of tiles, so whatever has landed is multiplied while the rest of
    // A and B is still on the wire -- row panels, column panels or an interleaving of both, whatever order the caller
    // uploaded them in.
    const bool chunked = pending && ctx->pipeline != 0;
    std::vector<int> group_of(outs.size(), 0);
    int ngroups_dyn = 1;
    if (chunked) {
      std::vector<std::pair<uint64_t, size_t>> order(outs.size());
      for (size_t oi = 0; oi < outs.size(); ++oi) {
        uint64_t key = 0;
        for (const GemmSrc& g : plans[out_plan[oi]].src) {
          if (g.a->ready) key = std::max(key, g.a->seq);
          if (g.b->ready) key = std::max(key, g.b->seq);
        }
        order[oi] = {key, oi};
      }
      std::sort(order.begin(), order.end());
      const int64_t min_tiles = 148;  // >= one wave; the launches overlap on the side streams, so small chunks cost nothing
      int64_t acc_tiles = 0;
      int gcur = 0;
      for (size_t r = 0; r < order.size(); ++r) {
        const size_t oi = order[r].second;
        // close the chunk when it is big enough AND the next block waits for a later copy
        if (acc_tiles >= min_tiles && r > 0 && order[r].first != order[r - 1].first) {
          ++gcur;
          acc_tiles = 0;
        }
        group_of[oi] = gcur;
        acc_tiles += static_cast<int64_t>((outs[oi].m + BM - 1) / BM) * ((outs[oi].n + BN - 1) / BN);
      }
      ngroups_dyn = gcur + 1;
    }
    auto grp = [&](const Keyed& k) { return group_of[k.t.out]; };
    std::sort(keyed.begin(), keyed.end(), [&](const Keyed& x, const Keyed& y) {
      const int gx = grp(x), gy = grp(y);
      if (gx != gy) return gx < gy;
      if (x.band != y.band) return x.band < y.band;
      if (x.gm != y.gm) return x.gm < y.gm;
      return x.gn < y.gn;
    });
    std::vector<GemmTile> tiles(keyed.size());
    for (size_t i = 0; i < keyed.size(); ++i) tiles[i] = keyed[i].t;
    // TMA descriptors of the K-contiguous operand blocks (row-major A blocks, column-major B blocks)
    struct TmapBytes { unsigned char b[128]; };
    std::vector<TmapBytes> tmaps;
    std::map<std::tuple<const double*, int64_t, int64_t, int64_t>, int32_t> tmap_index;
    auto tmap_of = [&](const double* base, int64_t kdim, int64_t rows, int64_t ld, int tile) -> int32_t {
      auto key = std::make_tuple(base, kdim, rows, ld);
      auto it = tmap_index.find(key);
      if (it != tmap_index.end()) return it->second;
      TmapBytes t;
      int32_t idx = -1;
      if (encode_kcontig_tmap(t.b, base, kdim, rows, ld, tile)) {
        idx = static_cast<int32_t>(tmaps.size());
        tmaps.push_back(t);
      }
      tmap_index.emplace(key, idx);
      return idx;
    };
    for (auto& go : outs)
      for (int p = go.pair_begin; p < go.pair_begin + go.pair_count; ++p) {
        GemmPair& pr = pairs[p];
        if (pr.aT) pr.tmA = tmap_of(pr.A, pr.kdim, go.m, pr.lda, BM);
        if (!pr.bT) pr.tmB = tmap_of(pr.B, pr.kdim, go.n, pr.ldb, BN);
      }
    Buf d_outs = upload(ctx, outs), d_pairs = upload(ctx, pairs), d_tiles = upload(ctx, tiles), d_tmaps = upload(ctx, tmaps);
    if (ctx->time_kernels) CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    const int ngroups = ngroups_dyn;
    const bool side = chunked && ngroups > 1;
    if (side) {  // fork: the side streams start after everything enqueued so far (output slab, descriptor tables)
      CUDA_CHECK(cudaEventRecord(ctx->ev_order, ctx->stream));
      for (int i = 0; i < mr_context::kChunkStreams; ++i) CUDA_CHECK(cudaStreamWaitEvent(ctx->chunk_stream[i], ctx->ev_order, 0));
    }
    size_t t0 = 0;
    for (int gi = 0; gi < ngroups; ++gi) {
      size_t t1 = t0;
      while (t1 < keyed.size() && grp(keyed[t1]) == gi) ++t1;
      cudaStream_t cs = side ? ctx->chunk_stream[gi % mr_context::kChunkStreams] : ctx->stream;
      // wait for exactly the operand blocks this chunk reads
      std::vector<char> seen(outs.size(), 0);
      for (size_t t = t0; t < t1; ++t) {
        const int oi = keyed[t].t.out;
        if (seen[oi]) continue;
        seen[oi] = 1;
        for (const GemmSrc& g : plans[out_plan[oi]].src) {
          wait_ready_on(cs, *g.a);
          wait_ready_on(cs, *g.b);
        }
      }
      CUDA_CHECK(launch_gemm_f64(static_cast<const GemmOut*>(d_outs->p), static_cast<const GemmPair*>(d_pairs->p),
                                 static_cast<const GemmTile*>(d_tiles->p) + t0, static_cast<int>(t1 - t0), d_tmaps->p, variant,
                                 cs));
      note_launch(ctx);
      if (chunked) {  // consumers on the egress stream wait for this chunk only
        ReadyPtr r = std::make_shared<Ready>();
        CUDA_CHECK(cudaEventRecord(r->ev, cs));
        for (size_t oi = 0; oi < outs.size(); ++oi)
          if (seen[oi]) {
            const OutPlan& o = plans[out_plan[oi]];
            if (o.spmm.empty()) result->blocks[{o.rid, o.cid}].ready = r;  // blocks with sparse partials finish later
          }
      }
      t0 = t1;
    }
    if (side) {  // join: later work on the context stream (and the release of the descriptor tables) follows every chunk
      for (int i = 0; i < mr_context::kChunkStreams; ++i) {
        CUDA_CHECK(cudaEventRecord(ctx->chunk_join[i], ctx->chunk_stream[i]));
        CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->chunk_join[i], 0));
      }
    }
    ctx->stats.gemm_launches += 1;
    ctx->stats.last_gemm_flops = flops;
    if (ctx->time_kernels) {
      CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
      CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
      ctx->stats.last_gemm_ms = ms;
      ctx->stats.gemm_ms_total += ms;
    }
  }

  // ---- sparse x dense partial products accumulate onto the GEMM result (LocalMatrix.add of partials).
  // CSR pairs with block dims <= 1024 go through ONE fused launch (K loop over the pairs inside the kernel);
  // CSC pairs and oversized blocks use the per-pair kernels.
  std::vector<SpmmOut> fouts;
  std::vector<SpmmPair> fpairs;
  int fused_max_n = 0;
  for (size_t i = 0; i < plans.size(); ++i) {
    auto& o = plans[i];
    bool have = !o.gemm.empty();
    if (o.m == 0 || o.n == 0 || !o.spsp.empty()) continue;
    std::vector<std::pair<const Block*, const Block*>> slow;
    SpmmOut fo{};
    fo.C = cptr[i];
    fo.m = o.m;
    fo.n = o.n;
    fo.pair_begin = static_cast<int32_t>(fpairs.size());
    for (auto& sp : o.spmm) {
      const Block& s = *sp.first;
      const Block& b = *sp.second;
      wait_ready(ctx, s);
      wait_ready(ctx, b);
      if (s.isT && o.m <= kSpmmMaxDim && s.numCols <= kSpmmMaxDim) {
        SpmmPair pr{};
        pr.ptrs = s.colPtrs.ptr<int32_t>();
        pr.idx = s.rowIndices.ptr<int32_t>();
        pr.vals = s.values.ptr<double>();
        pr.B = b.values.ptr<double>();
        pr.kdim = s.numCols;
        pr.bT = b.isT;
        fpairs.push_back(pr);
      } else {
        slow.push_back(sp);
      }
    }
    fo.pair_count = static_cast<int32_t>(fpairs.size()) - fo.pair_begin;
    for (auto& sp : slow) {  // per-pair kernels first so the fused launch can simply accumulate on top
      const Block& s = *sp.first;
      const Block& b = *sp.second;
      CUDA_CHECK(launch_spmm(s.colPtrs.ptr<int32_t>(), s.rowIndices.ptr<int32_t>(), s.values.ptr<double>(), s.isT,
                             b.values.ptr<double>(), b.isT, cptr[i], s.numRows, s.numCols, b.numCols, have, ctx->stream));
      note_launch(ctx);
      have = true;
    }
    if (fo.pair_count > 0) {
      fo.accumulate = have ? 1 : 0;
      fouts.push_back(fo);
      fused_max_n = std::max(fused_max_n, o.n);
      have = true;
    }
    if (!have) CUDA_CHECK(cudaMemsetAsync(cptr[i], 0, static_cast<size_t>(o.m) * o.n * sizeof(double), ctx->stream));
  }
  if (!fouts.empty()) {
    Buf d_fo = upload(ctx, fouts), d_fp = upload(ctx, fpairs);
    if (ctx->time_kernels) CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    CUDA_CHECK(launch_spmm_fused(static_cast<const SpmmOut*>(d_fo->p), static_cast<int>(fouts.size()),
                                 static_cast<const SpmmPair*>(d_fp->p), fused_max_n, ctx->stream));
    note_launch(ctx);
    if (ctx->time_kernels) {
      CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
      CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
      ctx->stats.last_gemm_ms = ms;
      ctx->stats.gemm_ms_total += ms;
    }
  }
  run_sparse_chains(ctx, plans, chains, cptr, planner, result);
}


// ------------------------------------------------------------------------------------------------
// element-wise operator plumbing
// ------------------------------------------------------------------------------------------------
void check_same_dims(int64_t lr, int64_t lc, int64_t rr, int64_t rc) {
  // MatfastExecution.scala:584-587 (and :621-624, :658-661)
  MR_REQUIRE(lr == rr, MR_EDIM, "Row number not match, leftRowNum = %lld, rightRowNum = %lld", (long long)lr,
             (long long)rr);
  MR_REQUIRE(lc == rc, MR_EDIM, "Col number not match, leftColNum = %lld, rightColNum = %lld", (long long)lc,
             (long long)rc);
}

struct EwBatch {
  mr_context* ctx;
  int op;
  std::vector<EwDesc> descs;
  std::vector<Block> keep;  // densified temporaries
  std::vector<std::pair<std::pair<int32_t, int32_t>, std::pair<int32_t, int32_t>>> shapes;  // key -> (rows, cols)
  int max_rows = 0, max_cols = 0;
  bool any_T = false;
  size_t total = 0;
  std::vector<size_t> sparse_rule;  // indices of results that follow the sparse (op) sparse output-format rule

  void add(std::pair<int32_t, int32_t> key, const Block* a, const Block* b, const double* y, int rows, int cols) {
    EwDesc d{};
    d.A = a ? a->values.ptr<double>() : nullptr;
    d.B = b ? b->values.ptr<double>() : nullptr;
    d.Y = y;
    d.rows = rows;
    d.cols = cols;
    d.aT = a ? a->isT : 0;
    d.bT = (b && op != EW_RANK1 && op != EW_RANK1_COMPAT) ? b->isT : 0;
    any_T = any_T || d.aT || d.bT;
    max_rows = std::max(max_rows, rows);
    max_cols = std::max(max_cols, cols);
    total += align_up(static_cast<size_t>(rows) * cols * sizeof(double));
    descs.push_back(d);
    shapes.push_back({key, {rows, cols}});
  }

  void run(mr_matrix* result) {
    if (descs.empty()) return;
    Slab slab(ctx, total);
    for (size_t i = 0; i < descs.size(); ++i) {
      const int rows = shapes[i].second.first, cols = shapes[i].second.second;
      Span s = slab.take(static_cast<size_t>(rows) * cols * sizeof(double));
      descs[i].C = s.ptr<double>();
      result->blocks[shapes[i].first] = dense_block(rows, cols, s, false);  // always column-major (LocalMatrix.scala:62)
    }
    Buf d = upload(ctx, descs);
    CUDA_CHECK(launch_ew_batched(op, static_cast<const EwDesc*>(d->p), static_cast<int>(descs.size()), max_rows,
                                 max_cols, any_T, ctx->stream));
    note_launch(ctx);
    if (!sparse_rule.empty()) apply_sparse_rule(result);
  }

  // LocalMatrix.addSparseSparse / elementWiseOpSparseSparse output format (LocalMatrix.scala:74-139, 521-602): the
  // dense result is converted with toSparse (CSC, isTransposed = false) iff rows*cols > 2*nnz + cols + 1, where nnz
  // counts entries != 0.0 (NaN included).  Both the transposed and the native branch reduce to this rule.
  void apply_sparse_rule(mr_matrix* result) {
    std::vector<DenseWin> wins;
    for (size_t bi : sparse_rule) wins.push_back(DenseWin{descs[bi].C, shapes[bi].second.first, shapes[bi].second.second});
    const auto counts = column_counts(ctx, wins);
    std::vector<DenseWin> chosen;
    std::vector<std::vector<int32_t>> chosen_counts;
    std::vector<std::pair<int32_t, int32_t>> keys;
    for (size_t i = 0; i < wins.size(); ++i) {
      if (static_cast<int64_t>(wins[i].rows) * wins[i].cols > 2 * total_count(counts[i]) + wins[i].cols + 1) {
        chosen.push_back(wins[i]);
        chosen_counts.push_back(counts[i]);
        keys.push_back(shapes[sparse_rule[i]].first);
      }
    }
    std::vector<Block> fresh = compact_csc(ctx, chosen, chosen_counts);
    for (size_t i = 0; i < fresh.size(); ++i) result->blocks[keys[i]] = std::move(fresh[i]);  // replaces the dense window
  }
};

void check_block_dims_add(const Block& a, const Block& b) {
  // LocalMatrix.add (LocalMatrix.scala:36-41)
  MR_REQUIRE(a.numRows == b.numRows, MR_EDIM,
             "Matrix A and B must have the same number of rows. But found A.numRows = %d, B.numRows = %d", a.numRows,
             b.numRows);
  MR_REQUIRE(a.numCols == b.numCols, MR_EDIM,
             "Matrix A and B must have the same number of cols. But found A.numCols = %d, B.numCols = %d", a.numCols,
             b.numCols);
}

void check_block_dims_ew(const Block& a, const Block& b) {
  // LocalMatrix.elementWiseMultiply / Divide (LocalMatrix.scala:467-470, 480-483)
  MR_REQUIRE(a.numRows == b.numRows, MR_EDIM, "mat1.numRows = %d, mat2.numRows = %d", a.numRows, b.numRows);
  MR_REQUIRE(a.numCols == b.numCols, MR_EDIM, "mat1.numCols = %d, mat2.numCols = %d", a.numCols, b.numCols);
}

mr_matrix* new_matrix(mr_context* ctx) {
  auto* m = new mr_matrix;
  m->ctx = ctx;
  return m;
}

void elementwise_join(int op, mr_matrix* left, mr_matrix* right, mr_matrix* result) {
  mr_context* ctx = left->ctx;
  EwBatch batch{ctx, op};
  batch.keep.reserve(2 * (left->blocks.size() + right->blocks.size()) + 2);
  auto dense_view = [&](const Block& b) -> const Block* {
    if (b.dense()) return &b;
    batch.keep.push_back(densify(ctx, b));
    return &batch.keep.back();
  };
  for (auto& kv : left->blocks) {
    auto it = right->blocks.find(kv.first);
    if (it == right->blocks.end()) {
      if (op == EW_ADD) result->blocks[kv.first] = kv.second;  // outer join: one-sided blocks pass through
      continue;
    }
    const Block& a = kv.second;
    const Block& b = it->second;
    if (op == EW_ADD) check_block_dims_add(a, b);
    else check_block_dims_ew(a, b);
    const bool both_sparse = !a.dense() && !b.dense();
    const Block* x = dense_view(a);
    const Block* y = dense_view(b);
    // defect B4 (LocalMatrix.scala:474,487): (Sparse, Dense) swaps the operands; visible for divide only
    if (op == EW_DIV && !a.dense() && b.dense() && ctx->compat_bugs) std::swap(x, y);
    batch.add(kv.first, x, y, nullptr, a.numRows, a.numCols);
    if (both_sparse) batch.sparse_rule.push_back(batch.descs.size() - 1);
  }
  if (op == EW_ADD)
    for (auto& kv : right->blocks)
      if (!left->blocks.count(kv.first)) result->blocks[kv.first] = kv.second;
  batch.run(result);
}

void map_values(int op, mr_matrix* a, double alpha, mr_matrix* result) {
  mr_context* ctx = a->ctx;
  std::vector<MapDesc> descs;
  size_t total = 0;
  int64_t max_n = 0;
  for (auto& kv : a->blocks) total += align_up(static_cast<size_t>(kv.second.valuesLen) * sizeof(double));
  Slab slab(ctx, total);
  for (auto& kv : a->blocks) {
    const Block& b = kv.second;
    Block o = b;  // same type, dims, flag and (shared) index arrays: the map touches stored values only
    o.values = slab.take(static_cast<size_t>(b.valuesLen) * sizeof(double));
    if (b.valuesLen > 0) {
      descs.push_back(MapDesc{b.values.ptr<double>(), o.values.ptr<double>(), b.valuesLen});
      max_n = std::max(max_n, b.valuesLen);
    }
    result->blocks[kv.first] = std::move(o);
  }
  if (descs.empty()) return;
  Buf d = upload(ctx, descs);
  CUDA_CHECK(launch_map_batched(op, static_cast<const MapDesc*>(d->p), static_cast<int>(descs.size()), max_n, alpha,
                                ctx->stream));
  note_launch(ctx);
}

void validate_desc(const mr_block_desc* d) {
  MR_REQUIRE(d != nullptr, MR_EINVAL, "block descriptor is null");
  MR_REQUIRE(d->numRows >= 0 && d->numCols >= 0, MR_EINVAL, "negative block dimensions %d x %d", d->numRows, d->numCols);
  if (d->type == 1) {
    // DenseMatrix ctor (MLMatrix.scala:240)
    MR_REQUIRE(d->valuesLen == static_cast<int64_t>(d->numRows) * d->numCols, MR_EINVAL,
               "The number of values supplied doesn't match the size of the matrix! values.length: %lld, "
               "numRows * numCols: %lld",
               (long long)d->valuesLen, (long long)(static_cast<int64_t>(d->numRows) * d->numCols));
    MR_REQUIRE(d->valuesLen == 0 || d->values != nullptr, MR_EINVAL, "values is null");
  } else if (d->type == 0) {
    // SparseMatrix ctor (MLMatrix.scala:533-542)
    MR_REQUIRE(d->valuesLen == d->rowIndicesLen, MR_EINVAL,
               "The number of row indices and values don't match! values.length: %lld, rowIndices.length: %lld",
               (long long)d->valuesLen, (long long)d->rowIndicesLen);
    if (d->isTransposed)
      MR_REQUIRE(d->colPtrsLen == d->numRows + 1, MR_EINVAL, "Expecting %d colPtrs when numRows = %d but got %lld",
                 d->numRows + 1, d->numRows, (long long)d->colPtrsLen);
    else
      MR_REQUIRE(d->colPtrsLen == d->numCols + 1, MR_EINVAL, "Expecting %d colPtrs when numCols = %d but got %lld",
                 d->numCols + 1, d->numCols, (long long)d->colPtrsLen);
    MR_REQUIRE(d->colPtrs != nullptr, MR_EINVAL, "colPtrs is null");
    MR_REQUIRE(d->valuesLen == d->colPtrs[d->colPtrsLen - 1], MR_EINVAL,
               "The last value of colPtrs must equal the number of elements. values.length: %lld, colPtrs.last: %d",
               (long long)d->valuesLen, d->colPtrs[d->colPtrsLen - 1]);
    MR_REQUIRE(d->valuesLen == 0 || (d->values != nullptr && d->rowIndices != nullptr), MR_EINVAL,
               "values / rowIndices is null");
  } else {
    fail(MR_ENOTSUP, "Unsupported matrix type %d", static_cast<int>(d->type));
  }
}

Span upload_raw(mr_context* ctx, const void* host, size_t bytes) {
  Span s{std::make_shared<DevBuf>(ctx, std::max<size_t>(bytes, 16)), 0};
  if (bytes) {
    CUDA_CHECK(cudaMemcpyAsync(s.buf->p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += static_cast<int64_t>(bytes);
  }
  return s;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// ABI: lifetime
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* mr_last_error(void) { return g_last_error.c_str(); }
const char* mr_version(void) { return "matrel-b200 0.1 (sm_100a)"; }

mr_status mr_init(const mr_options* opts, mr_context** out) {
  return guarded([&] {
    MR_REQUIRE(out != nullptr, MR_EINVAL, "out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
      (void)cudaGetLastError();
      fail(MR_ECUDA, "no usable CUDA device (%s): the B200 engine has no CPU fallback",
           e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    auto ctx = std::unique_ptr<mr_context>(new mr_context);
    int dev = opts ? opts->device : -1;
    if (dev < 0) CUDA_CHECK(cudaGetDevice(&dev));
    MR_REQUIRE(dev < count, MR_EINVAL, "device ordinal %d out of range (have %d)", dev, count);
    CUDA_CHECK(cudaSetDevice(dev));
    ctx->device = dev;
    cudaDeviceProp prop{};
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
      fail(MR_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major, prop.minor);
    if (opts) {
      ctx->compat_bugs = opts->compat_bugs;
      ctx->gemm_algo = opts->gemm_algo;
      ctx->ozaki_slices = opts->ozaki_slices;
    }
    if (opts && opts->stream) {
      ctx->stream = static_cast<cudaStream_t>(opts->stream);
    } else {
      CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    }
    CUDA_CHECK(cudaEventCreate(&ctx->ev0));
    CUDA_CHECK(cudaEventCreate(&ctx->ev1));
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_alloc, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    for (int i = 0; i < mr_context::kChunkStreams; ++i) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->chunk_stream[i], cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&ctx->chunk_join[i], cudaEventDisableTiming));
    }
    {
      void* hp = nullptr;
      const size_t cap = 16u << 20;
      if (cudaHostAlloc(&hp, cap, cudaHostAllocMapped) == cudaSuccess) {
        void* dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, hp, 0) == cudaSuccess) {
          ctx->stage_host = static_cast<char*>(hp);
          ctx->stage_dev = static_cast<char*>(dp);
          ctx->stage_cap = cap;
        } else {
          cudaFreeHost(hp);
        }
      }
      (void)cudaGetLastError();
    }
    // keep freed blocks in the pool: operators allocate result slabs on every call
    cudaMemPool_t pool;
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thr = UINT64_MAX;
    CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    *out = ctx.release();
  });
}

mr_status mr_shutdown(mr_context* ctx) {
  return guarded([&] {
    if (!ctx) return;
    cudaStreamSynchronize(ctx->h2d_stream);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->d2h_stream);
    cudaStreamDestroy(ctx->h2d_stream);
    cudaStreamDestroy(ctx->d2h_stream);
    for (int i = 0; i < mr_context::kChunkStreams; ++i) {
      if (ctx->chunk_stream[i]) cudaStreamDestroy(ctx->chunk_stream[i]);
      if (ctx->chunk_join[i]) cudaEventDestroy(ctx->chunk_join[i]);
    }
    if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
    if (ctx->ev_alloc) cudaEventDestroy(ctx->ev_alloc);
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
  });
}

mr_status mr_set_stream(mr_context* ctx, void* cuda_stream) {
  return guarded([&] {
    MR_REQUIRE(ctx != nullptr, MR_EINVAL, "ctx is null");
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) {
      cudaStreamDestroy(ctx->stream);
      ctx->own_stream = false;
    }
    if (cuda_stream) {
      ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
      CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    }
  });
}

mr_status mr_set_option(mr_context* ctx, const char* key, int64_t value) {
  return guarded([&] {
    MR_REQUIRE(ctx != nullptr && key != nullptr, MR_EINVAL, "ctx/key is null");
    std::string k(key);
    if (k == "compat_bugs") ctx->compat_bugs = static_cast<int>(value);
    else if (k == "gemm_algo") ctx->gemm_algo = static_cast<int>(value);
    else if (k == "ozaki_slices") ctx->ozaki_slices = static_cast<int>(value);
    else if (k == "crt_moduli") ctx->crt_moduli = static_cast<int>(value);
    else if (k == "time_kernels") ctx->time_kernels = static_cast<int>(value);
    else if (k == "gemm_variant") ctx->force_variant = static_cast<int>(value);
    else if (k == "pipeline") ctx->pipeline = static_cast<int>(value);
    else fail(MR_EINVAL, "unknown option '%s'", key);
  });
}

mr_status mr_sync(mr_context* ctx) {
  return guarded([&] {
    MR_REQUIRE(ctx != nullptr, MR_EINVAL, "ctx is null");
    CUDA_CHECK(cudaStreamSynchronize(ctx->h2d_stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->d2h_stream));
  });
}

mr_status mr_get_stats(mr_context* ctx, mr_stats* out) {
  return guarded([&] {
    MR_REQUIRE(ctx != nullptr && out != nullptr, MR_EINVAL, "ctx/out is null");
    *out = ctx->stats;
  });
}

mr_status mr_reset_stats(mr_context* ctx) {
  return guarded([&] {
    MR_REQUIRE(ctx != nullptr, MR_EINVAL, "ctx is null");
    ctx->stats = mr_stats{};
  });
}

// ------------------------------------------------------------------------------------------------
// ABI: datasets
// ------------------------------------------------------------------------------------------------
mr_status mr_matrix_create(mr_context* ctx, mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(ctx != nullptr && out != nullptr, MR_EINVAL, "ctx/out is null");
    *out = new_matrix(ctx);
  });
}

mr_status mr_matrix_free(mr_matrix* m) {
  return guarded([&] { delete m; });
}

mr_status mr_matrix_put_block(mr_matrix* m, int32_t rid, int32_t cid, const mr_block_desc* d) {
  return guarded([&] {
    MR_REQUIRE(m != nullptr, MR_EINVAL, "matrix is null");
    validate_desc(d);
    mr_context* ctx = m->ctx;
    Block b;
    b.type = d->type;
    b.numRows = d->numRows;
    b.numCols = d->numCols;
    b.isT = d->isTransposed != 0;
    b.valuesLen = d->valuesLen;
    // Ingest on its own stream: the copy of block k+1 overlaps whatever the context stream is computing, and
    // operators wait per block (Block::ready), so a multiply can start on the row panels that have landed.
    cudaStream_t cs = ctx->pipeline ? ctx->h2d_stream : ctx->stream;
    auto put = [&](const void* host, size_t bytes) {
      Span sp{std::make_shared<DevBuf>(ctx, std::max<size_t>(bytes, 16)), 0};
      return std::make_pair(sp, bytes ? host : nullptr);
    };
    auto vals = put(d->values, static_cast<size_t>(d->valuesLen) * sizeof(double));
    std::pair<Span, const void*> cp{}, ri{};
    if (d->type == 0) {
      b.colPtrsLen = d->colPtrsLen;
      cp = put(d->colPtrs, static_cast<size_t>(d->colPtrsLen) * sizeof(int32_t));
      ri = put(d->rowIndices, static_cast<size_t>(d->rowIndicesLen) * sizeof(int32_t));
    }
    if (ctx->pipeline) {  // allocations are ordered on the context stream: the ingest stream must see them
      CUDA_CHECK(cudaEventRecord(ctx->ev_alloc, ctx->stream));
      CUDA_CHECK(cudaStreamWaitEvent(cs, ctx->ev_alloc, 0));
    }
    auto copy = [&](std::pair<Span, const void*>& x, size_t bytes) {
      if (x.second && bytes) {
        CUDA_CHECK(cudaMemcpyAsync(x.first.buf->p, x.second, bytes, cudaMemcpyHostToDevice, cs));
        ctx->stats.h2d_bytes += static_cast<int64_t>(bytes);
      }
    };
    copy(vals, static_cast<size_t>(d->valuesLen) * sizeof(double));
    b.values = vals.first;
    if (d->type == 0) {
      copy(cp, static_cast<size_t>(d->colPtrsLen) * sizeof(int32_t));
      copy(ri, static_cast<size_t>(d->rowIndicesLen) * sizeof(int32_t));
      b.colPtrs = cp.first;
      b.rowIndices = ri.first;
    }
    if (ctx->pipeline) {
      ReadyPtr r = std::make_shared<Ready>();
      CUDA_CHECK(cudaEventRecord(r->ev, cs));
      b.ready = r;
      b.seq = ++ctx->ingest_seq;
      b.values.buf->ready = r;
      if (d->type == 0) {
        b.colPtrs.buf->ready = r;
        b.rowIndices.buf->ready = r;
      }
    }
    m->blocks[{rid, cid}] = std::move(b);
  });
}

mr_status mr_matrix_put_blocks(mr_matrix* m, int64_t count, const int32_t* rids, const int32_t* cids, const mr_block_desc* blks) {
  if (count < 0 || (count > 0 && (!rids || !cids || !blks))) {
    g_last_error = "requirement failed: null argument";
    return MR_EINVAL;
  }
  for (int64_t i = 0; i < count; ++i) {
    const mr_status st = mr_matrix_put_block(m, rids[i], cids[i], &blks[i]);
    if (st != MR_OK) return st;
  }
  return MR_OK;
}

mr_status mr_matrix_put_block_device(mr_matrix* m, int32_t rid, int32_t cid, int32_t numRows, int32_t numCols,
                                     const double* dvalues, uint8_t isTransposed) {
  return guarded([&] {
    MR_REQUIRE(m != nullptr, MR_EINVAL, "matrix is null");
    MR_REQUIRE(numRows >= 0 && numCols >= 0, MR_EINVAL, "negative block dimensions %d x %d", numRows, numCols);
    MR_REQUIRE(dvalues != nullptr || static_cast<int64_t>(numRows) * numCols == 0, MR_EINVAL, "device pointer is null");
    Span s{std::make_shared<DevBuf>(m->ctx, const_cast<double*>(dvalues),
                                    static_cast<size_t>(numRows) * numCols * sizeof(double)),
           0};
    m->blocks[{rid, cid}] = dense_block(numRows, numCols, s, isTransposed != 0);
  });
}

mr_status mr_matrix_put_blocks_device(mr_matrix* m, int64_t count, const int32_t* rids, const int32_t* cids,
                                      const int32_t* numRows, const int32_t* numCols, const double* const* dvalues,
                                      const uint8_t* isTransposed) {
  return guarded([&] {
    MR_REQUIRE(m != nullptr, MR_EINVAL, "matrix is null");
    MR_REQUIRE(count == 0 || (rids && cids && numRows && numCols && dvalues), MR_EINVAL, "null argument");
    for (int64_t i = 0; i < count; ++i) {
      MR_REQUIRE(numRows[i] >= 0 && numCols[i] >= 0, MR_EINVAL, "negative block dimensions %d x %d", numRows[i], numCols[i]);
      Span s{std::make_shared<DevBuf>(m->ctx, const_cast<double*>(dvalues[i]),
                                      static_cast<size_t>(numRows[i]) * numCols[i] * sizeof(double)),
             0};
      m->blocks[{rids[i], cids[i]}] = dense_block(numRows[i], numCols[i], s, isTransposed ? isTransposed[i] != 0 : false);
    }
  });
}

mr_status mr_matrix_num_blocks(const mr_matrix* m, int64_t* out) {
  return guarded([&] {
    MR_REQUIRE(m != nullptr && out != nullptr, MR_EINVAL, "matrix/out is null");
    *out = static_cast<int64_t>(m->blocks.size());
  });
}

mr_status mr_matrix_has_block(const mr_matrix* m, int32_t rid, int32_t cid, int32_t* out) {
  return guarded([&] {
    MR_REQUIRE(m != nullptr && out != nullptr, MR_EINVAL, "matrix/out is null");
    *out = m->blocks.count({rid, cid}) ? 1 : 0;
  });
}

mr_status mr_matrix_block_ids(const mr_matrix* m, int32_t* rids, int32_t* cids, int64_t cap) {
  return guarded([&] {
    MR_REQUIRE(m != nullptr && rids != nullptr && cids != nullptr, MR_EINVAL, "null argument");
    MR_REQUIRE(cap >= static_cast<int64_t>(m->blocks.size()), MR_EINVAL, "capacity %lld < number of blocks %zu",
               (long long)cap, m->blocks.size());
    int64_t i = 0;
    for (auto& kv : m->blocks) {
      rids[i] = kv.first.first;
      cids[i] = kv.first.second;
      ++i;
    }
  });
}

mr_status mr_matrix_get_block(mr_matrix* m, int32_t rid, int32_t cid, mr_block_desc* io) {
  return guarded([&] {
    MR_REQUIRE(m != nullptr && io != nullptr, MR_EINVAL, "null argument");
    auto it = m->blocks.find({rid, cid});
    if (it == m->blocks.end()) fail(MR_ENOTFOUND, "no block (%d, %d) in this dataset", rid, cid);
    const Block& b = it->second;
    mr_context* ctx = m->ctx;
    ReadyPtr ready;
    bool settled;
    {
      std::lock_guard<std::mutex> lock(ctx->mu);
      ready = b.ready;  // snapshot: operators on other threads may drop completed events
      settled = b.settled;
    }
    const int64_t rowIndicesLen = b.dense() ? 0 : b.valuesLen;
    // Egress on its own stream, ordered after this block's producer only (one chunk of a chunked multiply, or
    // everything enqueued on the context stream so far when the block has no event of its own).
    cudaStream_t rs = ctx->pipeline ? ctx->d2h_stream : ctx->stream;
    if (ctx->pipeline && (io->values || io->colPtrs || io->rowIndices)) {
      if (ready) {
        CUDA_CHECK(cudaStreamWaitEvent(rs, ready->ev, 0));
      } else if (!settled) {
        CUDA_CHECK(cudaEventRecord(ctx->ev_order, ctx->stream));
        CUDA_CHECK(cudaStreamWaitEvent(rs, ctx->ev_order, 0));
      }
    }
    if (io->values != nullptr) {
      MR_REQUIRE(io->valuesLen >= b.valuesLen, MR_EINVAL, "values capacity %lld < %lld", (long long)io->valuesLen,
                 (long long)b.valuesLen);
      if (b.valuesLen)
        CUDA_CHECK(cudaMemcpyAsync(io->values, b.values.ptr<double>(), static_cast<size_t>(b.valuesLen) * sizeof(double),
                                   cudaMemcpyDeviceToHost, rs));
      ctx->stats.d2h_bytes += b.valuesLen * 8;
    }
    if (!b.dense() && io->colPtrs != nullptr) {
      MR_REQUIRE(io->colPtrsLen >= b.colPtrsLen, MR_EINVAL, "colPtrs capacity too small");
      CUDA_CHECK(cudaMemcpyAsync(io->colPtrs, b.colPtrs.ptr<int32_t>(), static_cast<size_t>(b.colPtrsLen) * 4,
                                 cudaMemcpyDeviceToHost, rs));
      ctx->stats.d2h_bytes += b.colPtrsLen * 4;
    }
    if (!b.dense() && io->rowIndices != nullptr) {
      MR_REQUIRE(io->rowIndicesLen >= rowIndicesLen, MR_EINVAL, "rowIndices capacity too small");
      if (rowIndicesLen)
        CUDA_CHECK(cudaMemcpyAsync(io->rowIndices, b.rowIndices.ptr<int32_t>(), static_cast<size_t>(rowIndicesLen) * 4,
                                   cudaMemcpyDeviceToHost, rs));
      ctx->stats.d2h_bytes += rowIndicesLen * 4;
    }
    if (io->values || io->colPtrs || io->rowIndices) CUDA_CHECK(cudaStreamSynchronize(rs));
    io->type = b.type;
    io->numRows = b.numRows;
    io->numCols = b.numCols;
    io->isTransposed = b.isT ? 1 : 0;
    io->valuesLen = b.valuesLen;
    io->colPtrsLen = b.dense() ? 0 : b.colPtrsLen;
    io->rowIndicesLen = rowIndicesLen;
  });
}

mr_status mr_matrix_block_device_ptr(mr_matrix* m, int32_t rid, int32_t cid, double** dptr) {
  return guarded([&] {
    MR_REQUIRE(m != nullptr && dptr != nullptr, MR_EINVAL, "null argument");
    auto it = m->blocks.find({rid, cid});
    if (it == m->blocks.end()) fail(MR_ENOTFOUND, "no block (%d, %d) in this dataset", rid, cid);
    *dptr = it->second.values.ptr<double>();
  });
}

mr_status mr_matrix_rand(mr_context* ctx, int64_t nrows, int64_t ncols, int32_t blkSize, int64_t seed0,
                         mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(ctx != nullptr && out != nullptr, MR_EINVAL, "ctx/out is null");
    MR_REQUIRE(nrows > 0 && ncols > 0 && blkSize > 0, MR_EINVAL, "nrows, ncols, blkSize must be positive");
    // DenseMatrix.rand `require` (MLMatrix.scala:454-455)
    MR_REQUIRE(static_cast<int64_t>(blkSize) * blkSize <= INT32_MAX, MR_EINVAL,
               "%d x %d dense matrix is too large to allocate", blkSize, blkSize);
    const int64_t nbr = ceil_div(nrows, blkSize), nbc = ceil_div(ncols, blkSize);
    std::unique_ptr<mr_matrix> m(new_matrix(ctx));
    size_t total = 0;
    for (int64_t i = 0; i < nbr; ++i)
      for (int64_t j = 0; j < nbc; ++j) {
        const int64_t r = std::min<int64_t>(blkSize, nrows - i * blkSize), c = std::min<int64_t>(blkSize, ncols - j * blkSize);
        total += align_up(static_cast<size_t>(r * c) * sizeof(double));
      }
    Slab slab(ctx, total);
    std::vector<RandDesc> descs;
    int64_t max_n = 0;
    for (int64_t i = 0; i < nbr; ++i)
      for (int64_t j = 0; j < nbc; ++j) {
        const int32_t r = static_cast<int32_t>(std::min<int64_t>(blkSize, nrows - i * blkSize));
        const int32_t c = static_cast<int32_t>(std::min<int64_t>(blkSize, ncols - j * blkSize));
        Span s = slab.take(static_cast<size_t>(r) * c * sizeof(double));
        descs.push_back(RandDesc{s.ptr<double>(), static_cast<int64_t>(r) * c, seed0 + i * nbc + j});
        max_n = std::max<int64_t>(max_n, static_cast<int64_t>(r) * c);
        m->blocks[{static_cast<int32_t>(i), static_cast<int32_t>(j)}] = dense_block(r, c, s, false);
      }
    // gridDim.y limit: launch in chunks of 65535 blocks
    for (size_t off = 0; off < descs.size(); off += 65535) {
      std::vector<RandDesc> chunk(descs.begin() + off, descs.begin() + std::min(descs.size(), off + 65535));
      Buf d = upload(ctx, chunk);
      CUDA_CHECK(launch_java_rand_batched(static_cast<const RandDesc*>(d->p), static_cast<int>(chunk.size()), max_n, ctx->stream));
      note_launch(ctx);
    }
    *out = m.release();
  });
}

mr_status mr_matrix_rand_partition(mr_context* ctx, int64_t nrows, int64_t ncols, int32_t blkSize, int64_t seed0,
                                   int32_t pr, int32_t pc, int32_t r, int32_t c, double* dslab, int64_t slotElems,
                                   mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(ctx != nullptr && out != nullptr && dslab != nullptr, MR_EINVAL, "null argument");
    MR_REQUIRE(nrows > 0 && ncols > 0 && blkSize > 0, MR_EINVAL, "nrows, ncols, blkSize must be positive");
    MR_REQUIRE(pr > 0 && pc > 0 && r >= 0 && r < pr && c >= 0 && c < pc, MR_EINVAL, "bad process grid %d x %d / (%d, %d)",
               pr, pc, r, c);
    MR_REQUIRE(slotElems >= static_cast<int64_t>(blkSize) * blkSize, MR_EINVAL, "slotElems %lld < blkSize^2",
               (long long)slotElems);
    const int64_t nbr = ceil_div(nrows, blkSize), nbc = ceil_div(ncols, blkSize);
    const int64_t slots_c = ceil_div(nbc, pc);
    std::unique_ptr<mr_matrix> m(new_matrix(ctx));
    std::vector<RandDesc> descs;
    int64_t max_n = 0;
    for (int64_t i = r; i < nbr; i += pr)
      for (int64_t j = c; j < nbc; j += pc) {
        const int32_t br = static_cast<int32_t>(std::min<int64_t>(blkSize, nrows - i * blkSize));
        const int32_t bc = static_cast<int32_t>(std::min<int64_t>(blkSize, ncols - j * blkSize));
        double* p = dslab + ((i / pr) * slots_c + (j / pc)) * slotElems;
        descs.push_back(RandDesc{p, static_cast<int64_t>(br) * bc, seed0 + i * nbc + j});
        max_n = std::max<int64_t>(max_n, static_cast<int64_t>(br) * bc);
        Span s{std::make_shared<DevBuf>(ctx, p, static_cast<size_t>(br) * bc * sizeof(double)), 0};
        m->blocks[{static_cast<int32_t>(i), static_cast<int32_t>(j)}] = dense_block(br, bc, s, false);
      }
    for (size_t off = 0; off < descs.size(); off += 65535) {
      std::vector<RandDesc> chunk(descs.begin() + off, descs.begin() + std::min(descs.size(), off + 65535));
      Buf d = upload(ctx, chunk);
      CUDA_CHECK(launch_java_rand_batched(static_cast<const RandDesc*>(d->p), static_cast<int>(chunk.size()), max_n, ctx->stream));
      note_launch(ctx);
    }
    *out = m.release();
  });
}

// ------------------------------------------------------------------------------------------------
// ABI: operators
// ------------------------------------------------------------------------------------------------
mr_status mr_matrix_multiply(mr_matrix* left, int64_t leftRowNum, int64_t leftColNum, mr_matrix* right,
                             int64_t rightRowNum, int64_t rightColNum, int32_t blkSize, mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(left && right && out, MR_EINVAL, "null argument");
    MR_REQUIRE(left->ctx == right->ctx, MR_EINVAL, "operands belong to different contexts");
    MR_REQUIRE(blkSize > 0, MR_EINVAL, "blkSize must be positive, got %d", blkSize);
    // MatfastExecution.scala:702-703
    MR_REQUIRE(leftColNum == rightRowNum, MR_EDIM, "Matrix dimension not match, leftColNum = %lld, rightRowNum = %lld",
               (long long)leftColNum, (long long)rightRowNum);
    (void)leftRowNum;
    (void)rightColNum;
    mr_context* ctx = left->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    for (auto& kv : left->blocks)
      if (!kv.second.dense()) wait_ready(ctx, kv.second);
    for (auto& kv : right->blocks)
      if (!kv.second.dense()) wait_ready(ctx, kv.second);
    MultiplyPlanner planner{ctx};
    std::vector<OutPlan> plans;
    const int64_t leftColBlkNum = ceil_div(leftColNum, blkSize);    // :712
    const int64_t rightRowBlkNum = ceil_div(rightRowNum, blkSize);  // :713
    if (leftColBlkNum == 1 && rightRowBlkNum == 1) {
      // outer-product paths (:714-721 -> helper :175-221): every left block x every right block, no
      // reduce.  Reference defect B1 (DuplicateLeft throws) is not reproduced; both branches return
      // what DuplicateRight returns.
      std::map<std::pair<int32_t, int32_t>, size_t> seen;
      for (auto& l : left->blocks)
        for (auto& r : right->blocks) {
          OutPlan o;
          o.rid = l.first.first;
          o.cid = r.first.second;
          planner.add_pair(o, l.second, r.second, 0);
          auto key = std::make_pair(o.rid, o.cid);
          auto it = seen.find(key);
          if (it == seen.end()) {
            seen[key] = plans.size();
            plans.push_back(std::move(o));
          } else {
            plans[it->second] = std::move(o);  // duplicate keys: the last row wins when collected into a map
          }
        }
    } else {
      // matrixMultiplyGeneral (helper :235-263): join on k, then reduce by (i, j) in ascending k.
      std::map<int32_t, std::vector<std::pair<int32_t, const Block*>>> rights;  // k -> (j, B(k,j))
      for (auto& r : right->blocks) rights[r.first.first].push_back({r.first.second, &r.second});
      std::map<std::pair<int32_t, int32_t>, size_t> index;
      // left->blocks is ordered by (i, k): iterating it visits k ascending within each i
      for (auto& l : left->blocks) {
        const int32_t i = l.first.first, k = l.first.second;
        auto rit = rights.find(k);
        if (rit == rights.end()) continue;
        for (auto& jb : rit->second) {
          auto key = std::make_pair(i, jb.first);
          auto it = index.find(key);
          if (it == index.end()) {
            OutPlan o;
            o.rid = i;
            o.cid = jb.first;
            it = index.emplace(key, plans.size()).first;
            plans.push_back(std::move(o));
          }
          planner.add_pair(plans[it->second], l.second, *jb.second, k);
        }
      }
    }
    std::unique_ptr<mr_matrix> result(new_matrix(ctx));
    run_multiply(ctx, plans, planner, blkSize, result.get(), leftRowNum, leftColNum, rightColNum,
                 leftColBlkNum == 1 && rightRowBlkNum == 1);
    *out = result.release();
  });
}

mr_status mr_transpose(mr_matrix* a, mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(a && out, MR_EINVAL, "null argument");
    std::unique_ptr<mr_matrix> r(new_matrix(a->ctx));
    for (auto& kv : a->blocks) {
      Block b = kv.second;  // shares the device arrays (DenseMatrix.transpose, MLMatrix.scala:312; Sparse :634-635)
      std::swap(b.numRows, b.numCols);
      b.isT = !b.isT;
      r->blocks[{kv.first.second, kv.first.first}] = std::move(b);  // (rid, cid) -> (cid, rid), MatfastExecution.scala:230-231
    }
    *out = r.release();
  });
}

static mr_status ew_operator(int op, mr_matrix* left, int64_t lr, int64_t lc, mr_matrix* right, int64_t rr, int64_t rc,
                             mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(left && right && out, MR_EINVAL, "null argument");
    MR_REQUIRE(left->ctx == right->ctx, MR_EINVAL, "operands belong to different contexts");
    check_same_dims(lr, lc, rr, rc);
    std::lock_guard<std::mutex> lock(left->ctx->mu);
    wait_ready_all(left->ctx, left);
    wait_ready_all(left->ctx, right);
    std::unique_ptr<mr_matrix> r(new_matrix(left->ctx));
    elementwise_join(op, left, right, r.get());
    *out = r.release();
  });
}

mr_status mr_add_element(mr_matrix* left, int64_t lr, int64_t lc, mr_matrix* right, int64_t rr, int64_t rc,
                         int32_t blkSize, mr_matrix** out) {
  (void)blkSize;
  return ew_operator(EW_ADD, left, lr, lc, right, rr, rc, out);
}
mr_status mr_multiply_element(mr_matrix* left, int64_t lr, int64_t lc, mr_matrix* right, int64_t rr, int64_t rc,
                              int32_t blkSize, mr_matrix** out) {
  (void)blkSize;
  return ew_operator(EW_MUL, left, lr, lc, right, rr, rc, out);
}
mr_status mr_divide_element(mr_matrix* left, int64_t lr, int64_t lc, mr_matrix* right, int64_t rr, int64_t rc,
                            int32_t blkSize, mr_matrix** out) {
  (void)blkSize;
  return ew_operator(EW_DIV, left, lr, lc, right, rr, rc, out);
}

static mr_status map_operator(int op, mr_matrix* a, double alpha, mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(a && out, MR_EINVAL, "null argument");
    std::lock_guard<std::mutex> lock(a->ctx->mu);
    wait_ready_all(a->ctx, a);
    std::unique_ptr<mr_matrix> r(new_matrix(a->ctx));
    map_values(op, a, alpha, r.get());
    *out = r.release();
  });
}
mr_status mr_add_scalar(mr_matrix* a, double alpha, mr_matrix** out) { return map_operator(MAP_ADD_SCALAR, a, alpha, out); }
mr_status mr_multiply_scalar(mr_matrix* a, double alpha, mr_matrix** out) { return map_operator(MAP_MUL_SCALAR, a, alpha, out); }
mr_status mr_power(mr_matrix* a, double alpha, mr_matrix** out) { return map_operator(MAP_POW, a, alpha, out); }

mr_status mr_rank_one_update(mr_matrix* left, int64_t lr, int64_t lc, mr_matrix* right, int64_t rr, int64_t rc,
                             int32_t blkSize, mr_matrix** out) {
  (void)blkSize;
  return guarded([&] {
    MR_REQUIRE(left && right && out, MR_EINVAL, "null argument");
    MR_REQUIRE(left->ctx == right->ctx, MR_EINVAL, "operands belong to different contexts");
    mr_context* ctx = left->ctx;
    if (ctx->compat_bugs) {
      // the reference's own `require`s (MatfastExecution.scala:741-744), which only admit 1-row matrices
      MR_REQUIRE(rr == 1, MR_EDIM, "Vector column size is not 1, but #cols = %lld", (long long)rr);
      MR_REQUIRE(lr == rr, MR_EDIM,
                 "Dimension not match for matrix addition, A.nrows = %lld, A.ncols = %lld, B.nrows = %lld, B.ncols = %lld",
                 (long long)lr, (long long)lc, (long long)rr, (long long)rc);
    } else {
      // intended semantics: A (n x n) + v v^T with v an n x 1 block column (MatrixOperator.scala:151-152)
      MR_REQUIRE(rc == 1, MR_EDIM, "Vector column size is not 1, but #cols = %lld", (long long)rc);
      MR_REQUIRE(lr == rr && lc == rr, MR_EDIM,
                 "Dimension not match for matrix addition, A.nrows = %lld, A.ncols = %lld, B.nrows = %lld, B.ncols = %lld",
                 (long long)lr, (long long)lc, (long long)rr, (long long)rc);
    }
    std::lock_guard<std::mutex> lock(ctx->mu);
    wait_ready_all(ctx, left);
    wait_ready_all(ctx, right);
    std::unique_ptr<mr_matrix> r(new_matrix(ctx));
    EwBatch batch{ctx, ctx->compat_bugs ? EW_RANK1_COMPAT : EW_RANK1};
    batch.keep.reserve(left->blocks.size() + 2 * right->blocks.size() + 2);
    std::map<const Block*, const Block*> dense_cache;
    auto dense_view = [&](const Block& b) -> const Block* {
      if (b.dense()) return &b;
      auto it = dense_cache.find(&b);
      if (it != dense_cache.end()) return it->second;
      batch.keep.push_back(densify(ctx, b));
      return dense_cache[&b] = &batch.keep.back();
    };
    for (auto& kv : left->blocks) {
      const int32_t i = kv.first.first, j = kv.first.second;
      const Block& a = kv.second;
      // helper :271-275: x2.rid == i, x3.rid == j (vector blocks are looked up by their row-block id)
      const Block *x = nullptr, *y = nullptr;
      for (auto& vb : right->blocks) {
        if (vb.first.first == i) x = &vb.second;
        if (vb.first.first == j) y = &vb.second;
      }
      if (!x || !y) continue;
      MR_REQUIRE(static_cast<int64_t>(x->numRows) * x->numCols >= a.numRows &&
                     static_cast<int64_t>(y->numRows) * y->numCols >= a.numCols,
                 MR_EDIM, "vector block shorter than matrix block (%d x %d)", a.numRows, a.numCols);
      if (ctx->compat_bugs && !a.dense())
        fail(MR_ENOTSUP, "rankOneAdd on a sparse block in compat mode (LocalMatrix.scala:1079-1081 indexes the "
                         "dense result with the sparse value index) is not reproduced");
      const Block* xd = dense_view(*x);
      const Block* yd = dense_view(*y);
      const Block* ad = ctx->compat_bugs ? &a : dense_view(a);
      // mat2(i, 0) of an n x 1 (or, transposed, 1 x n flagged) dense block is values[i] either way
      batch.add(kv.first, ad, xd, yd->values.ptr<double>(), a.numRows, a.numCols);
    }
    batch.run(r.get());
    *out = r.release();
  });
}

static mr_status aggregate_operator(int op, mr_matrix* a, int64_t nrows, int64_t ncols, mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(a && out, MR_EINVAL, "null argument");
    if (op == AGG_TRACE)  // Dataset.scala:80
      MR_REQUIRE(nrows == ncols, MR_EDIM, "Cannot perform trace() on a rectangle matrix");
    mr_context* ctx = a->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    wait_ready_all(ctx, a);
    std::unique_ptr<mr_matrix> r(new_matrix(ctx));
    // output blocks: one per block-row (rowSum), block-column (colSum), or a single scalar
    std::map<std::pair<int32_t, int32_t>, int32_t> out_len;
    for (auto& kv : a->blocks) {
      const Block& b = kv.second;
      if (op == AGG_TRACE && kv.first.first != kv.first.second) continue;
      if (op == AGG_TRACE && b.dense())  // MatfastExecution.scala:428
        MR_REQUIRE(b.numRows == b.numCols, MR_EDIM, "block is not square, row_num=%d, col_num=%d", b.numRows, b.numCols);
      std::pair<int32_t, int32_t> key = op == AGG_ROW_SUM ? std::make_pair(kv.first.first, 0)
                                        : op == AGG_COL_SUM ? std::make_pair(0, kv.first.second)
                                                            : std::make_pair(0, 0);
      const int32_t len = op == AGG_ROW_SUM ? b.numRows : op == AGG_COL_SUM ? b.numCols : 1;
      auto it = out_len.find(key);
      if (it == out_len.end()) out_len[key] = len;
      else if (op == AGG_ROW_SUM)  // LocalMatrix.add requires of the reduceByKey (LocalMatrix.scala:36-41)
        MR_REQUIRE(it->second == len, MR_EDIM,
                   "Matrix A and B must have the same number of rows. But found A.numRows = %d, B.numRows = %d", it->second, len);
      else if (op == AGG_COL_SUM)
        MR_REQUIRE(it->second == len, MR_EDIM,
                   "Matrix A and B must have the same number of cols. But found A.numCols = %d, B.numCols = %d", it->second, len);
    }
    if (out_len.empty()) {
      *out = r.release();
      return;
    }
    size_t total = 0;
    for (auto& kv : out_len) total += align_up(static_cast<size_t>(kv.second) * sizeof(double));
    Slab slab(ctx, total);
    CUDA_CHECK(cudaMemsetAsync(slab.buf->p, 0, std::max<size_t>(total, kAlign), ctx->stream));
    std::map<std::pair<int32_t, int32_t>, double*> out_ptr;
    for (auto& kv : out_len) {
      Span s = slab.take(static_cast<size_t>(kv.second) * sizeof(double));
      out_ptr[kv.first] = s.ptr<double>();
      r->blocks[kv.first] = op == AGG_ROW_SUM ? dense_block(kv.second, 1, s) : op == AGG_COL_SUM ? dense_block(1, kv.second, s)
                                                                                                  : dense_block(1, 1, s);
    }
    std::vector<AggDesc> descs;
    std::vector<Block> keep;
    keep.reserve(a->blocks.size() + 1);
    int max_r = 1, max_c = 1;
    for (auto& kv : a->blocks) {
      const Block& b = kv.second;
      if (op == AGG_TRACE && kv.first.first != kv.first.second) continue;
      AggDesc d{};
      if (op == AGG_SUM) {  // values.sum over the STORED values, dense or sparse (:381-384)
        d.v = b.values.ptr<double>();
        d.rows = static_cast<int32_t>(std::min<int64_t>(b.valuesLen, INT32_MAX));
        d.cols = 1;
        d.isT = 0;
        if (b.valuesLen == 0) continue;
      } else {
        const Block* src = &b;
        if (!b.dense()) {
          keep.push_back(densify(ctx, b));
          src = &keep.back();
        }
        d.v = src->values.ptr<double>();
        d.rows = src->numRows;
        d.cols = src->numCols;
        d.isT = src->isT;
      }
      d.out = out_ptr[op == AGG_ROW_SUM ? std::make_pair(kv.first.first, 0)
                      : op == AGG_COL_SUM ? std::make_pair(0, kv.first.second)
                                          : std::make_pair(0, 0)];
      max_r = std::max(max_r, d.rows);
      max_c = std::max(max_c, d.cols);
      descs.push_back(d);
    }
    if (!descs.empty()) {
      Buf dd = upload(ctx, descs);
      CUDA_CHECK(launch_aggregate(op, static_cast<const AggDesc*>(dd->p), static_cast<int>(descs.size()), max_r, max_c, ctx->stream));
      note_launch(ctx);
    }
    *out = r.release();
  });
}
mr_status mr_row_sum(mr_matrix* a, int64_t nrows, int64_t ncols, mr_matrix** out) { return aggregate_operator(AGG_ROW_SUM, a, nrows, ncols, out); }
mr_status mr_col_sum(mr_matrix* a, int64_t nrows, int64_t ncols, mr_matrix** out) { return aggregate_operator(AGG_COL_SUM, a, nrows, ncols, out); }
mr_status mr_sum(mr_matrix* a, int64_t nrows, int64_t ncols, mr_matrix** out) { return aggregate_operator(AGG_SUM, a, nrows, ncols, out); }
mr_status mr_trace(mr_matrix* a, int64_t nrows, int64_t ncols, mr_matrix** out) { return aggregate_operator(AGG_TRACE, a, nrows, ncols, out); }

static mr_status slice_operator(mr_matrix* a, int32_t blkSize, bool take_row, int64_t index, int64_t index2, mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(blkSize > 0, MR_EINVAL, "blkSize must be positive, got %d", blkSize);
    mr_context* ctx = a->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    wait_ready_all(ctx, a);
    std::unique_ptr<mr_matrix> r(new_matrix(ctx));
    const int32_t blkid = static_cast<int32_t>(index / blkSize), offset = static_cast<int32_t>(index % blkSize);  // :40-41
    const int32_t blkid2 = index2 >= 0 ? static_cast<int32_t>(index2 / blkSize) : -1;
    const int32_t offset2 = index2 >= 0 ? static_cast<int32_t>(index2 % blkSize) : -1;
    std::vector<LineDesc> descs;
    std::vector<std::pair<std::pair<int32_t, int32_t>, std::pair<int32_t, int32_t>>> outs;  // key -> (rows, cols)
    std::vector<Block> keep;
    keep.reserve(a->blocks.size() + 1);
    size_t total = 0;
    int max_len = 1;
    for (auto& kv : a->blocks) {
      const int32_t rid = kv.first.first, cid = kv.first.second;
      if ((take_row ? rid : cid) != blkid) continue;                       // filter(tuple => tuple._1 == rowblkID), :48
      if (index2 >= 0 && (take_row ? cid : rid) != blkid2) continue;
      const Block* src = &kv.second;
      if (!src->dense()) {
        keep.push_back(densify(ctx, *src));
        src = &keep.back();
      }
      if (offset >= (take_row ? src->numRows : src->numCols)) continue;
      if (index2 >= 0 && offset2 >= (take_row ? src->numCols : src->numRows)) continue;
      LineDesc d{};
      d.v = src->values.ptr<double>();
      d.rows = src->numRows;
      d.cols = src->numCols;
      d.offset = offset;
      d.offset2 = offset2;
      d.len = index2 >= 0 ? 1 : (take_row ? src->numCols : src->numRows);
      d.isT = src->isT;
      d.take_row = take_row;
      descs.push_back(d);
      if (index2 >= 0) outs.push_back({{0, 0}, {1, 1}});
      else if (take_row) outs.push_back({{0, cid}, {1, d.len}});
      else outs.push_back({{rid, 0}, {d.len, 1}});
      total += align_up(static_cast<size_t>(d.len) * sizeof(double));
      max_len = std::max(max_len, d.len);
    }
    if (!descs.empty()) {
      Slab slab(ctx, total);
      for (size_t i = 0; i < descs.size(); ++i) {
        Span sp = slab.take(static_cast<size_t>(descs[i].len) * sizeof(double));
        descs[i].out = sp.ptr<double>();
        r->blocks[outs[i].first] = dense_block(outs[i].second.first, outs[i].second.second, sp);
      }
      Buf dd = upload(ctx, descs);
      CUDA_CHECK(launch_extract_lines(static_cast<const LineDesc*>(dd->p), static_cast<int>(descs.size()), max_len, ctx->stream));
      note_launch(ctx);
    }
    *out = r.release();
  });
}

mr_status mr_project(mr_matrix* a, int64_t nrows, int64_t ncols, int32_t blkSize, int32_t rowOrCol, int64_t index, mr_matrix** out) {
  mr_status st = guarded([&] {
    MR_REQUIRE(a && out, MR_EINVAL, "null argument");
    if (rowOrCol)  // Dataset.scala:42,44
      MR_REQUIRE(index >= 0 && index < nrows, MR_EINVAL, "row index should be smaller than #rows, index=%lld, #rows=%lld",
                 (long long)index, (long long)nrows);
    else
      MR_REQUIRE(index >= 0 && index < ncols, MR_EINVAL, "col index should be smaller than #cols, index=%lld, #cols=%lld",
                 (long long)index, (long long)ncols);
  });
  if (st != MR_OK) return st;
  return slice_operator(a, blkSize, rowOrCol != 0, index, -1, out);
}

mr_status mr_selection(mr_matrix* a, int64_t nrows, int64_t ncols, int32_t blkSize, int64_t rowIdx, int64_t colIdx, mr_matrix** out) {
  mr_status st = guarded([&] {
    MR_REQUIRE(a && out, MR_EINVAL, "null argument");
    // Dataset.scala:52-53
    MR_REQUIRE(rowIdx >= 0 && rowIdx < nrows, MR_EINVAL, "row index should be smaller than #rows, rid=%lld, #rows=%lld",
               (long long)rowIdx, (long long)nrows);
    MR_REQUIRE(colIdx >= 0 && colIdx < ncols, MR_EINVAL, "col index should be smaller than #cols, cid=%lld, #cols=%lld",
               (long long)colIdx, (long long)ncols);
  });
  if (st != MR_OK) return st;
  return slice_operator(a, blkSize, true, rowIdx, colIdx, out);
}

mr_status mr_vec(mr_matrix* a, int64_t nrows, int64_t ncols, int32_t blkSize, mr_matrix** out) {
  (void)ncols;
  mr_matrix* canon = nullptr;
  mr_status st = guarded([&] {
    MR_REQUIRE(a && out, MR_EINVAL, "null argument");
    MR_REQUIRE(blkSize > 0 && nrows > 0, MR_EINVAL, "nrows and blkSize must be positive");
  });
  if (st != MR_OK) return st;
  st = mr_materialize(a, &canon);  // column-major, non-transposed dense blocks (shares already canonical ones)
  if (st != MR_OK) return st;
  st = guarded([&] {
    mr_context* ctx = a->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    std::unique_ptr<mr_matrix> r(new_matrix(ctx));
    const int64_t ROW_BLK_NUM = ceil_div(nrows, blkSize);  // MatfastExecution.scala:543
    for (auto& kv : canon->blocks) {
      const Block& b = kv.second;
      const int64_t i = kv.first.first, j = kv.first.second;
      for (int32_t t = 0; t < b.numCols; ++t) {
        const int64_t key = (j * blkSize + t) * ROW_BLK_NUM + i;  // :552 with the block-column offset in elements
        MR_REQUIRE(key <= INT32_MAX, MR_EINVAL, "vec(): block id %lld does not fit an Int", (long long)key);
        Span col{b.values.buf, b.values.off + static_cast<size_t>(t) * b.numRows * sizeof(double)};
        Block v = dense_block(b.numRows, 1, col, false);
        v.ready = b.ready;
        r->blocks[{static_cast<int32_t>(key), 0}] = std::move(v);
      }
    }
    *out = r.release();
  });
  mr_matrix_free(canon);
  return st;
}

mr_status mr_materialize(mr_matrix* a, mr_matrix** out) {
  return guarded([&] {
    MR_REQUIRE(a && out, MR_EINVAL, "null argument");
    mr_context* ctx = a->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    wait_ready_all(ctx, a);
    std::unique_ptr<mr_matrix> r(new_matrix(ctx));
    EwBatch batch{ctx, EW_COPY};
    for (auto& kv : a->blocks) {
      const Block& b = kv.second;
      if (!b.dense()) {
        r->blocks[kv.first] = densify(ctx, b);
      } else if (!b.isT) {
        r->blocks[kv.first] = b;  // already canonical: share
      } else {
        batch.add(kv.first, &b, nullptr, nullptr, b.numRows, b.numCols);
      }
    }
    batch.run(r.get());
    *out = r.release();
  });
}

// ------------------------------------------------------------------------------------------------
// ABI: placement (pure integer; bit-exact with M/partitioner/*.scala)
// ------------------------------------------------------------------------------------------------
mr_status mr_row_partition(int32_t rid, int32_t cid, int32_t partitions, int32_t* out) {
  return guarded([&] {
    (void)cid;
    MR_REQUIRE(out != nullptr, MR_EINVAL, "out is null");
    MR_REQUIRE(partitions >= 0, MR_EINVAL, "Number of partitions cannot be negative but found %d", partitions);
    if (partitions == 0) fail(MR_EINVAL, "/ by zero");  // java.lang.ArithmeticException
    *out = rid % partitions;                            // RowPartitioner.scala:34 (JVM % truncates like C)
  });
}

mr_status mr_column_partition(int32_t rid, int32_t cid, int32_t partitions, int32_t* out) {
  return guarded([&] {
    (void)rid;
    MR_REQUIRE(out != nullptr, MR_EINVAL, "out is null");
    MR_REQUIRE(partitions >= 0, MR_EINVAL, "Number of partitions cannot be negative but found %d", partitions);
    if (partitions == 0) fail(MR_EINVAL, "/ by zero");
    *out = cid % partitions;  // ColumnPartitioner.scala:34
  });
}

mr_status mr_index_partition(int32_t key, int32_t partitions, int32_t* out) {
  return guarded([&] {
    MR_REQUIRE(out != nullptr, MR_EINVAL, "out is null");
    MR_REQUIRE(partitions >= 0, MR_EINVAL, "Number of partitions cannot be negative but found %d", partitions);
    *out = key;  // IndexPartitioner.scala:31
  });
}

static int32_t java_round(double x) { return static_cast<int32_t>(std::floor(x + 0.5)); }  // math.round

mr_status mr_gen_block_cyclic(int64_t nrows, int64_t ncols, int32_t blkSize, int32_t out[4]) {
  return guarded([&] {
    MR_REQUIRE(out != nullptr, MR_EINVAL, "out is null");
    MR_REQUIRE(blkSize > 0, MR_EINVAL, "blkSize must be positive, got %d", blkSize);
    // MatfastExecutionHelper.scala:46-62
    const int32_t R = static_cast<int32_t>(std::ceil(nrows * 1.0 / blkSize));
    const int32_t C = static_cast<int32_t>(std::ceil(ncols * 1.0 / blkSize));
    const int numPartitions = 64;
    const double scale = 1.0 / std::sqrt(static_cast<double>(numPartitions));
    int32_t r = java_round(std::max(scale * R, 1.0));
    int32_t c = java_round(std::max(scale * C, 1.0));
    if (r == 1 || c == 1) {
      if (r != 1) r = java_round(std::max(r / 8.0, 1.0));
      if (c != 1) c = java_round(std::max(c / 8.0, 1.0));
    }
    out[0] = R;
    out[1] = C;
    out[2] = r;
    out[3] = c;
  });
}

static void block_cyclic_derive(const int32_t p[4], int32_t* rpn, int32_t* cpn, int32_t* nrp, int32_t* ncp) {
  // BlockCyclicPartitioner.scala:36-50
  MR_REQUIRE(p[0] > 0, MR_EINVAL, "Number of row blocks should be larger than 0, but found %d", p[0]);
  MR_REQUIRE(p[1] > 0, MR_EINVAL, "Number of col blocks should be larger than 0, but found %d", p[1]);
  MR_REQUIRE(p[2] > 0, MR_EINVAL, "Number of row blocks per partition should be larger than 0, but found %d", p[2]);
  MR_REQUIRE(p[3] > 0, MR_EINVAL, "Number of col blocks per partition should be larger than 0, but found %d", p[3]);
  *rpn = static_cast<int32_t>(std::ceil(p[0] * 1.0 / p[2]));
  *cpn = static_cast<int32_t>(std::ceil(p[1] * 1.0 / p[3]));
  *nrp = p[0] / *rpn;
  *ncp = p[1] / *cpn;
}

mr_status mr_block_cyclic_partition(const int32_t params[4], int32_t rid, int32_t cid, int32_t* out) {
  return guarded([&] {
    MR_REQUIRE(params != nullptr && out != nullptr, MR_EINVAL, "null argument");
    int32_t rpn, cpn, nrp, ncp;
    block_cyclic_derive(params, &rpn, &cpn, &nrp, &ncp);
    const int32_t n = rpn * cpn;
    *out = ((rid % nrp) * cpn + (cid % ncp)) % n;  // BlockCyclicPartitioner.scala:54-57 (defect B2 kept: bit-exact ids)
  });
}

mr_status mr_partition_id(int32_t scheme, const int32_t params[4], int32_t rid, int32_t cid, int32_t* out) {
  if (params == nullptr || out == nullptr) return guarded([&] { fail(MR_EINVAL, "null argument"); });
  switch (scheme) {
    case MR_PART_ROW: return mr_row_partition(rid, cid, params[0], out);
    case MR_PART_COLUMN: return mr_column_partition(rid, cid, params[0], out);
    case MR_PART_INDEX: return mr_index_partition(rid, params[0], out);
    case MR_PART_BLOCK_CYCLIC: return mr_block_cyclic_partition(params, rid, cid, out);
    default: return guarded([&] { fail(MR_EINVAL, "unknown partition scheme %d", scheme); });
  }
}

mr_status mr_block_cyclic_num_partitions(const int32_t params[4], int32_t* out) {
  return guarded([&] {
    MR_REQUIRE(params != nullptr && out != nullptr, MR_EINVAL, "null argument");
    int32_t rpn, cpn, nrp, ncp;
    block_cyclic_derive(params, &rpn, &cpn, &nrp, &ncp);
    *out = rpn * cpn;
  });
}

}  // extern "C"
