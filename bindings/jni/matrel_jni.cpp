// JNI shim between the Scala facade (bindings/scala/Dataset.scala) and the C ABI (include/matrel.h).
// NOT compiled in this image (no JDK: `javac`, `jni.h` absent); it is the binding a maintainer adds:
//   g++ -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude bindings/jni/matrel_jni.cpp \
//       -Lmatrel_b200 -lmatrel_b200 -o libmatrel_jni.so
#if __has_include(<jni.h>)
#include <jni.h>

#include "matrel.h"

namespace {
// `require` failures become IllegalArgumentException with the reference's message text.
bool throw_if(JNIEnv* env, mr_status st) {
  if (st == MR_OK) return false;
  const char* cls = (st == MR_EINVAL || st == MR_EDIM) ? "java/lang/IllegalArgumentException"
                    : (st == MR_ENOMEM)                ? "java/lang/OutOfMemoryError"
                    : (st == MR_ENOTSUP)               ? "java/lang/UnsupportedOperationException"
                                                       : "java/lang/RuntimeException";
  env->ThrowNew(env->FindClass(cls), mr_last_error());
  return true;
}
template <typename T>
T* ptr(jlong h) { return reinterpret_cast<T*>(static_cast<intptr_t>(h)); }
}  // namespace

extern "C" {

JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_matfast_b200_Native_init(JNIEnv* env, jclass, jint device, jboolean compat) {
  mr_options o{};
  o.device = device;
  o.compat_bugs = compat ? 1 : 0;
  mr_context* ctx = nullptr;
  if (throw_if(env, mr_init(&o, &ctx))) return 0;
  return reinterpret_cast<jlong>(ctx);
}
JNIEXPORT void JNICALL Java_org_apache_spark_sql_matfast_b200_Native_shutdown(JNIEnv* env, jclass, jlong ctx) {
  throw_if(env, mr_shutdown(ptr<mr_context>(ctx)));
}
JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_matfast_b200_Native_matrixCreate(JNIEnv* env, jclass, jlong ctx) {
  mr_matrix* m = nullptr;
  if (throw_if(env, mr_matrix_create(ptr<mr_context>(ctx), &m))) return 0;
  return reinterpret_cast<jlong>(m);
}
JNIEXPORT void JNICALL Java_org_apache_spark_sql_matfast_b200_Native_matrixFree(JNIEnv* env, jclass, jlong m) {
  throw_if(env, mr_matrix_free(ptr<mr_matrix>(m)));
}
// MLMatrixSerializer.serialize's 7 fields, arrays pinned with GetPrimitiveArrayCritical for the copy
JNIEXPORT void JNICALL Java_org_apache_spark_sql_matfast_b200_Native_putBlock(JNIEnv* env, jclass, jlong m, jint rid, jint cid,
                                                                              jbyte type, jint numRows, jint numCols,
                                                                              jintArray colPtrs, jintArray rowIndices,
                                                                              jdoubleArray values, jboolean isTransposed) {
  mr_block_desc d{};
  d.type = static_cast<uint8_t>(type);
  d.numRows = numRows;
  d.numCols = numCols;
  d.isTransposed = isTransposed ? 1 : 0;
  d.valuesLen = env->GetArrayLength(values);
  d.values = static_cast<double*>(env->GetPrimitiveArrayCritical(values, nullptr));
  if (colPtrs) {
    d.colPtrsLen = env->GetArrayLength(colPtrs);
    d.colPtrs = static_cast<int32_t*>(env->GetPrimitiveArrayCritical(colPtrs, nullptr));
    d.rowIndicesLen = env->GetArrayLength(rowIndices);
    d.rowIndices = static_cast<int32_t*>(env->GetPrimitiveArrayCritical(rowIndices, nullptr));
  }
  const mr_status st = mr_matrix_put_block(ptr<mr_matrix>(m), rid, cid, &d);
  if (colPtrs) {
    env->ReleasePrimitiveArrayCritical(rowIndices, d.rowIndices, JNI_ABORT);
    env->ReleasePrimitiveArrayCritical(colPtrs, d.colPtrs, JNI_ABORT);
  }
  env->ReleasePrimitiveArrayCritical(values, d.values, JNI_ABORT);
  throw_if(env, st);
}
#define MR_JNI_BINARY(JNAME, CNAME)                                                                                     \
  JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_matfast_b200_Native_##JNAME(                                        \
      JNIEnv* env, jclass, jlong left, jlong leftRowNum, jlong leftColNum, jlong right, jlong rightRowNum,              \
      jlong rightColNum, jint blkSize) {                                                                                \
    mr_matrix* out = nullptr;                                                                                           \
    if (throw_if(env, CNAME(ptr<mr_matrix>(left), leftRowNum, leftColNum, ptr<mr_matrix>(right), rightRowNum,           \
                            rightColNum, blkSize, &out)))                                                               \
      return 0;                                                                                                         \
    return reinterpret_cast<jlong>(out);                                                                                \
  }
MR_JNI_BINARY(matrixMultiply, mr_matrix_multiply)
MR_JNI_BINARY(addElement, mr_add_element)
MR_JNI_BINARY(multiplyElement, mr_multiply_element)
MR_JNI_BINARY(divideElement, mr_divide_element)
MR_JNI_BINARY(matrixRankOneUpdate, mr_rank_one_update)
JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_matfast_b200_Native_transpose(JNIEnv* env, jclass, jlong a) {
  mr_matrix* out = nullptr;
  if (throw_if(env, mr_transpose(ptr<mr_matrix>(a), &out))) return 0;
  return reinterpret_cast<jlong>(out);
}
#define MR_JNI_SCALAR(JNAME, CNAME)                                                                                   \
  JNIEXPORT jlong JNICALL Java_org_apache_spark_sql_matfast_b200_Native_##JNAME(JNIEnv* env, jclass, jlong a, jdouble alpha) { \
    mr_matrix* out = nullptr;                                                                                         \
    if (throw_if(env, CNAME(ptr<mr_matrix>(a), alpha, &out))) return 0;                                               \
    return reinterpret_cast<jlong>(out);                                                                              \
  }
MR_JNI_SCALAR(addScalar, mr_add_scalar)
MR_JNI_SCALAR(multiplyScalar, mr_multiply_scalar)
MR_JNI_SCALAR(power, mr_power)
// getBlock / blockIds / partitioner entry points follow the same pattern (two-call size query, then
// Set<Primitive>ArrayRegion into freshly allocated JVM arrays).
}  // extern "C"
#endif  // __has_include(<jni.h>)
