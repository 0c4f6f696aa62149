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
class _jdoubleArray : public _jarray {};
typedef _jobject* jobject;
typedef _jclass* jclass;
typedef _jarray* jarray;
typedef _jintArray* jintArray;
typedef _jdoubleArray* jdoubleArray;

struct JNIEnv {
  jclass FindClass(const char* name);
  jint ThrowNew(jclass cls, const char* msg);
  jsize GetArrayLength(jarray a);
  void* GetPrimitiveArrayCritical(jarray a, jboolean* is_copy);
  void ReleasePrimitiveArrayCritical(jarray a, void* carray, jint mode);
};
