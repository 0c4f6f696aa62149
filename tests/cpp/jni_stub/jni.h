// Minimal stand-in for the JDK's <jni.h>: ONLY what bindings/jni/matrel_jni.cpp uses, so that the shim can be type-checked
// against include/matrel.h on a box without a JDK (tests/test_abi_cpu.py).  Test infrastructure; never shipped or linked.
#pragma once
#include <cstdint>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2

typedef int32_t jint;
typedef int64_t jlong;
typedef uint8_t jboolean;
typedef int8_t jbyte;
typedef double jdouble;
typedef jint jsize;

class _jobject {};
class _jclass : public _jobject {};
class _jarray : public _jobject {};
class _jintArray : public _jarray {};
class _jlongArray : public _jarray {};
class _jdoubleArray : public _jarray {};
typedef _jobject* jobject;
typedef _jclass* jclass;
typedef _jarray* jarray;
typedef _jintArray* jintArray;
typedef _jlongArray* jlongArray;
typedef _jdoubleArray* jdoubleArray;

struct JNIEnv {
  jclass FindClass(const char* name);
  jint ThrowNew(jclass cls, const char* msg);
  jboolean ExceptionCheck();
  jsize GetArrayLength(jarray a);
  jintArray NewIntArray(jsize n);
  jlongArray NewLongArray(jsize n);
  void GetIntArrayRegion(jintArray a, jsize start, jsize len, jint* buf);
  void GetDoubleArrayRegion(jdoubleArray a, jsize start, jsize len, jdouble* buf);
  void SetIntArrayRegion(jintArray a, jsize start, jsize len, const jint* buf);
  void SetLongArrayRegion(jlongArray a, jsize start, jsize len, const jlong* buf);
  void SetDoubleArrayRegion(jdoubleArray a, jsize start, jsize len, const jdouble* buf);
};
