// Placement: bit-exact restatement of M/partitioner/*.scala (pure integer, host side).
#include "host.h"

using namespace matrel;
using namespace mrhost;

extern "C" {

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
