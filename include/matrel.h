/*
 * matrel.h -- C ABI of the B200-native block-matrix engine (drop-in for the MatRel/MatFast
 * block-multiply hot path).  Plain C: opaque handles, pointers and sizes; no C++/torch types.
 *
 * Every entry point names the reference interface it replaces.  Paths are relative to
 * /root/reference/src/main/scala/org/apache/spark/sql/matfast/ ("M/").
 *
 * Conventions
 *   - every function returns mr_status; on failure mr_last_error() (thread-local) holds the
 *     message, reproducing the reference's `require` text ("requirement failed: ...") so a JNI
 *     shim can ThrowNew(IllegalArgumentException, msg).  Nothing throws or aborts across the ABI.
 *   - a "dataset" (mr_matrix) is a bag of (rid, cid, block) rows exactly like the reference's
 *     Dataset rows (rid Int, cid Int, struct7); absent blocks are implicit zeros.  Matrix
 *     dimensions are passed per operator call, as in M/Dataset.scala.
 *   - blocks live in device memory (HBM) behind the handle; operators enqueue CUDA kernels on the
 *     context stream and return new handles (operators are pure).  There is NO CPU fallback:
 *     without a usable CUDA device mr_init fails with MR_ECUDA.
 */
#ifndef MATREL_H_
#define MATREL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MR_API __attribute__((visibility("default")))

typedef int32_t mr_status;
enum {
  MR_OK = 0,
  MR_EINVAL = 1,  /* IllegalArgumentException("requirement failed: ...") */
  MR_EDIM = 2,    /* dimension `require` of an operator failed (also IllegalArgumentException) */
  MR_ENOMEM = 3,
  MR_ECUDA = 4,
  MR_ENOTSUP = 5, /* SparkException("Unsupported matrix type ...") / out-of-scope dispatch row */
  MR_ENOTFOUND = 6
};

typedef struct mr_context mr_context; /* one per process per GPU (replaces MatfastSession) */
typedef struct mr_matrix mr_matrix;   /* a Dataset of MatrixBlock rows, device resident */

/* The 7-field block struct of MatrixUDT / MLMatrixSerializer
 * (M/matrix/MLMatrix.scala:176-184, M/util/MLMatrixSerializer.scala:26-48). */
typedef struct mr_block_desc {
  uint8_t type;          /* [0] 0 = sparse, 1 = dense */
  int32_t numRows;       /* [1] */
  int32_t numCols;       /* [2] */
  int32_t* colPtrs;      /* [3] NULL for dense; length (isTransposed ? numRows : numCols) + 1 */
  int32_t* rowIndices;   /* [4] NULL for dense; length nnz */
  double* values;        /* [5] dense: numRows*numCols, column-major (row-major if isTransposed) */
  uint8_t isTransposed;  /* [6] */
  int64_t colPtrsLen;    /* array lengths (JVM arrays carry theirs; C needs them spelled out) */
  int64_t rowIndicesLen;
  int64_t valuesLen;
} mr_block_desc;

typedef struct mr_options {
  int32_t device;       /* CUDA device ordinal; -1 = current device */
  int32_t compat_bugs;  /* 1 = reproduce reference defects B3/B4; 0 = intended math */
  int32_t gemm_algo;    /* 0 = auto: large regular dense products run on the tcgen05 tensor cores (Ozaki-II: int8 residue GEMMs +
                           CRT, fp64-exact to 1e-14 on data of ordinary range), guarded on the device: Inf / NaN operands, a
                           dynamic range of more than ~2^37 inside one row of A / column of B, K >= 2^17, irregular block grids
                           and small products take the exact DMMA kernel instead;
                           1 = DMMA fp64 tensor-core kernel (mma.sync m8n8k4 f64), always;
                           2 = Ozaki-I int8 tcgen05 kernel (digit slices); 3 = 3xTF32 tcgen05 kernel (fp32 results);
                           4 = Ozaki-II without the range guard (mr_set_option "crt_moduli" 6..16, default 16) */
  int32_t ozaki_slices; /* number of int8 slices for gemm_algo 2 (0 = default) */
  void* stream;         /* cudaStream_t to run on; NULL = a stream owned by the context */
} mr_options;

/* ---- lifetime: replaces MatfastSession.builder().getOrCreate() (M/MatfastSession.scala:177-234) */
MR_API mr_status mr_init(const mr_options* opts, mr_context** out);
MR_API mr_status mr_shutdown(mr_context* ctx);
MR_API mr_status mr_set_stream(mr_context* ctx, void* cuda_stream);
/* keys: "compat_bugs", "gemm_algo", "ozaki_slices", "crt_moduli", "ozaki_scratch_mb", "pipeline", "time_kernels", "gemm_variant" */
MR_API mr_status mr_set_option(mr_context* ctx, const char* key, int64_t value);
MR_API mr_status mr_sync(mr_context* ctx);
MR_API const char* mr_last_error(void);
MR_API const char* mr_version(void);

/* ---- datasets: replaces Seq(MatrixBlock(...)).toDS() (M/example/BasicMatrixOps.scala:115-116)
 *      and .collect() / .rdd.foreach on the result. */
MR_API mr_status mr_matrix_create(mr_context* ctx, mr_matrix** out);
MR_API mr_status mr_matrix_free(mr_matrix* m);
/* Copies the block to the device (MLMatrixSerializer.deserialize, :50-69, incl. the ctor
 * `require`s of DenseMatrix :240 and SparseMatrix :533-542).  Host arrays stay caller-owned. */
MR_API mr_status mr_matrix_put_block(mr_matrix* m, int32_t rid, int32_t cid, const mr_block_desc* blk);
/* Batched mr_matrix_put_block: `count` blocks in one call (one ABI crossing per Seq[MatrixBlock], not per row). */
MR_API mr_status mr_matrix_put_blocks(mr_matrix* m, int64_t count, const int32_t* rids, const int32_t* cids,
                                      const mr_block_desc* blks);
/* Adopts (borrows) a dense block already resident in device memory; not freed by the library. */
MR_API mr_status mr_matrix_put_block_device(mr_matrix* m, int32_t rid, int32_t cid, int32_t numRows,
                                            int32_t numCols, const double* dvalues, uint8_t isTransposed);
/* Batched form of mr_matrix_put_block_device (count blocks; arrays of length count). */
MR_API mr_status mr_matrix_put_blocks_device(mr_matrix* m, int64_t count, const int32_t* rids, const int32_t* cids,
                                             const int32_t* numRows, const int32_t* numCols,
                                             const double* const* dvalues, const uint8_t* isTransposed);
MR_API mr_status mr_matrix_num_blocks(const mr_matrix* m, int64_t* out);
/* *out = 1 iff block (rid, cid) is present (absent blocks are implicit zeros: the join semantics of
 * MatfastExecutionHelper.scala:64-263 depend on presence, not on values) */
MR_API mr_status mr_matrix_has_block(const mr_matrix* m, int32_t rid, int32_t cid, int32_t* out);
/* Fills rids/cids (capacity cap) in ascending (rid, cid) order. */
MR_API mr_status mr_matrix_block_ids(const mr_matrix* m, int32_t* rids, int32_t* cids, int64_t cap);
/* Two-call protocol (MLMatrixSerializer.serialize, :26-48): with NULL array pointers only the
 * scalar fields and *Len fields are filled; with non-NULL pointers the arrays are copied out
 * (capacity given by the *Len fields on entry). */
MR_API mr_status mr_matrix_get_block(mr_matrix* m, int32_t rid, int32_t cid, mr_block_desc* inout);
/* Device pointer of a dense block's values (for zero-copy consumers such as NCCL). */
MR_API mr_status mr_matrix_block_device_ptr(mr_matrix* m, int32_t rid, int32_t cid, double** dptr);
/* DenseMatrix.rand(numRows, numCols, new java.util.Random(seed)) generated on the device for
 * every block of an nrows x ncols matrix (M/matrix/MLMatrix.scala:453-457); per-block seed =
 * seed0 + rid*ceil(ncols/blk) + cid.  Used for synthetic benchmark inputs. */
MR_API mr_status mr_matrix_rand(mr_context* ctx, int64_t nrows, int64_t ncols, int32_t blkSize,
                                int64_t seed0, mr_matrix** out);

/* Same generator restricted to the blocks a rank of a pr x pc process grid owns
 * (rid % pr == r, RowPartitioner.scala:34; cid % pc == c, ColumnPartitioner.scala:34), written into a
 * caller-provided device slab: block (rid, cid) starts at dslab + ((rid / pr) * ceil(nbc / pc) + cid / pc) * slotElems.
 * The returned dataset borrows the slab. */
MR_API mr_status mr_matrix_rand_partition(mr_context* ctx, int64_t nrows, int64_t ncols, int32_t blkSize, int64_t seed0,
                                          int32_t pr, int32_t pc, int32_t r, int32_t c, double* dslab,
                                          int64_t slotElems, mr_matrix** out);

/* ---- operators: argument-for-argument with M/Dataset.scala:57-152 and the physical operators
 *      of M/execution/MatfastExecution.scala. */
/* Dataset.matrixMultiply :134-142 -> MatrixMatrixMultiplicationExecution :688-726 ->
 * MatfastExecutionHelper.matrixMultiplyGeneral :235-263 / multiplyOuterProductDuplicate* :175-221 */
MR_API mr_status mr_matrix_multiply(mr_matrix* left, int64_t leftRowNum, int64_t leftColNum,
                                    mr_matrix* right, int64_t rightRowNum, int64_t rightColNum,
                                    int32_t blkSize, mr_matrix** out);
/* Dataset.transpose / t :57-61 -> MatrixTransposeExecution :215-236 (flag flip + index swap) */
MR_API mr_status mr_transpose(mr_matrix* a, mr_matrix** out);
/* Dataset.addElement :105-112 -> MatrixElementAddExecution :571-607 (outer join) */
MR_API mr_status mr_add_element(mr_matrix* left, int64_t leftRowNum, int64_t leftColNum, mr_matrix* right,
                                int64_t rightRowNum, int64_t rightColNum, int32_t blkSize, mr_matrix** out);
/* Dataset.multiplyElement :114-122 -> MatrixElementMultiplyExecution :609-644 (inner join) */
MR_API mr_status mr_multiply_element(mr_matrix* left, int64_t leftRowNum, int64_t leftColNum, mr_matrix* right,
                                     int64_t rightRowNum, int64_t rightColNum, int32_t blkSize, mr_matrix** out);
/* Dataset.divideElement :124-132 -> MatrixElementDivideExecution :646-686 (inner join) */
MR_API mr_status mr_divide_element(mr_matrix* left, int64_t leftRowNum, int64_t leftColNum, mr_matrix* right,
                                   int64_t rightRowNum, int64_t rightColNum, int32_t blkSize, mr_matrix** out);
/* Dataset.addScalar :89-91 -> MatrixScalarAddExecution :465-486 -> LocalMatrix.addScalar */
MR_API mr_status mr_add_scalar(mr_matrix* a, double alpha, mr_matrix** out);
/* Dataset.multiplyScalar :93-97 -> MatrixScalarMultiplyExecution :488-509 */
MR_API mr_status mr_multiply_scalar(mr_matrix* a, double alpha, mr_matrix** out);
/* Dataset.power :99-103 -> MatrixPowerExecution :511-532 */
MR_API mr_status mr_power(mr_matrix* a, double alpha, mr_matrix** out);
/* Dataset.matrixRankOneUpdate :144-152 -> RankOneUpdateExecution :728-747 */
MR_API mr_status mr_rank_one_update(mr_matrix* left, int64_t leftRowNum, int64_t leftColNum, mr_matrix* right,
                                    int64_t rightRowNum, int64_t rightColNum, int32_t blkSize, mr_matrix** out);
/* ---- aggregates.  The mathematically intended reductions; the reference's index defects
 *      for non-square transposed / sparse blocks (MatfastExecution.scala:262-287, 338-357 = defect B5) are not reproduced. */
/* Dataset.rowSum :63-66 -> RowSumDirectExecution :239-300: blocks (rid, 0) of shape rows x 1 */
MR_API mr_status mr_row_sum(mr_matrix* a, int64_t nrows, int64_t ncols, mr_matrix** out);
/* Dataset.colSum :68-71 -> ColumnSumDirectExecution :303-366: blocks (0, cid) of shape 1 x cols */
MR_API mr_status mr_col_sum(mr_matrix* a, int64_t nrows, int64_t ncols, mr_matrix** out);
/* Dataset.sum :73-76 -> SumDirectExecution :369-398: one 1 x 1 block (0, 0) */
MR_API mr_status mr_sum(mr_matrix* a, int64_t nrows, int64_t ncols, mr_matrix** out);
/* Dataset.trace :78-82 -> TraceDirectExecution :401-463: one 1 x 1 block (0, 0) from the diagonal blocks */
MR_API mr_status mr_trace(mr_matrix* a, int64_t nrows, int64_t ncols, mr_matrix** out);
/* ---- slicing; intended semantics (the reference's transposed / sparse index defects are not reproduced) */
/* Dataset.project :38-47 -> Project{Row,Column}DirectExecution :31-150: rowOrCol != 0 -> row `index` as blocks (0, cid) of
 * shape 1 x cols; rowOrCol == 0 -> column `index` as blocks (rid, 0) of shape rows x 1 */
MR_API mr_status mr_project(mr_matrix* a, int64_t nrows, int64_t ncols, int32_t blkSize, int32_t rowOrCol, int64_t index,
                            mr_matrix** out);
/* Dataset.selection :49-55 -> SelectDirectExecution :152-213: the element (rowIdx, colIdx) as one 1 x 1 block (0, 0) */
MR_API mr_status mr_selection(mr_matrix* a, int64_t nrows, int64_t ncols, int32_t blkSize, int64_t rowIdx, int64_t colIdx,
                              mr_matrix** out);
/* Dataset.vec :84-87 -> VectorizeExecution :534-569: column t of block (i, j) becomes the rows x 1 block
 * ((j * blkSize + t) * ceil(nrows / blkSize) + i, 0).  Intended semantics (defect B6 -- the reference strides by numLocalCols --
 * is not reproduced); the output blocks are zero-copy views of the (column-major) input columns. */
MR_API mr_status mr_vec(mr_matrix* a, int64_t nrows, int64_t ncols, int32_t blkSize, mr_matrix** out);
/* Materialise every dense block as column-major, isTransposed = false (DenseMatrix.toArray,
 * M/matrix/MLMatrix.scala:55-61, as a device transpose kernel). */
MR_API mr_status mr_materialize(mr_matrix* a, mr_matrix** out);

/* ---- placement: bit-exact restatement of M/partitioner/{Row,Column,Index,BlockCyclic}Partitioner.scala (pure integer, host side) */
/* RowPartitioner.getPartition (RowPartitioner.scala:32-38) */
MR_API mr_status mr_row_partition(int32_t rid, int32_t cid, int32_t partitions, int32_t* out);
/* ColumnPartitioner.getPartition (ColumnPartitioner.scala:32-38) */
MR_API mr_status mr_column_partition(int32_t rid, int32_t cid, int32_t partitions, int32_t* out);
/* IndexPartitioner.getPartition (IndexPartitioner.scala:29-34) */
MR_API mr_status mr_index_partition(int32_t key, int32_t partitions, int32_t* out);
/* MatfastExecutionHelper.genBlockCyclicPartitioner (MatfastExecutionHelper.scala:46-62):
 * out = {ROW_BLK_NUM, COL_BLK_NUM, ROW_BLKS_PER_PARTITION, COL_BLKS_PER_PARTITION} */
MR_API mr_status mr_gen_block_cyclic(int64_t nrows, int64_t ncols, int32_t blkSize, int32_t out[4]);
/* BlockCyclicPartitioner.getPartition / numPartitions (BlockCyclicPartitioner.scala:31-62) */
MR_API mr_status mr_block_cyclic_partition(const int32_t params[4], int32_t rid, int32_t cid, int32_t* out);
MR_API mr_status mr_block_cyclic_num_partitions(const int32_t params[4], int32_t* out);
/* One entry point over the four schemes (Partitioner.getPartition of the four M/partitioner classes): params[0] = partitions for
 * MR_PART_ROW / MR_PART_COLUMN / MR_PART_INDEX (INDEX keys on rid), params[0..3] = the mr_gen_block_cyclic tuple for
 * MR_PART_BLOCK_CYCLIC. */
typedef enum mr_partition_scheme { MR_PART_ROW = 0, MR_PART_COLUMN = 1, MR_PART_INDEX = 2, MR_PART_BLOCK_CYCLIC = 3 } mr_partition_scheme;
MR_API mr_status mr_partition_id(int32_t scheme, const int32_t params[4], int32_t rid, int32_t cid, int32_t* out);

/* ---- introspection used by bench/tests (not part of the reference surface) */
typedef struct mr_stats {
  int64_t kernel_launches;  /* CUDA kernels launched by this library since mr_init / reset */
  int64_t gemm_launches;
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  double last_gemm_ms;      /* CUDA-event duration of the most recent GEMM kernel launch(es) */
  double gemm_ms_total;     /* sum of CUDA-event GEMM durations since reset (timing enabled only) */
  int64_t last_gemm_flops;
  double tc_gemm_ms_total;  /* CUDA-event time of the tcgen05 int8 GEMM launches since reset (timing enabled only) */
  int64_t tc_gemm_launches; /* multiplies that ran on the tcgen05 path since reset */
  int64_t tc_int8_ops;      /* int8 multiply-add operations (x 2) of the most recent tcgen05 multiply */
} mr_stats;
MR_API mr_status mr_get_stats(mr_context* ctx, mr_stats* out);
MR_API mr_status mr_reset_stats(mr_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MATREL_H_ */
