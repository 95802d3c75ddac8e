/* Minimal stand-in for the JDK's jni.h, ONLY for compile-checking java/librec_b200_jni.c in an image without a JDK
 * (tests/test_abi_symbols.py::test_jni_forwarder_compiles).  It declares the JNI types and the JNIEnv function-table
 * members the forwarder uses, with the signatures of the JNI specification (Java SE 8, chapter 4); the member ORDER of
 * the real table is irrelevant to a compile check.  Never ship or link this file. */
#ifndef LRK_STUB_JNI_H
#define LRK_STUB_JNI_H
#include <stdint.h>
#include <stddef.h>
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef uint8_t jboolean; typedef float jfloat; typedef double jdouble;
typedef jint jsize;
struct _jobject; typedef struct _jobject* jobject;
typedef jobject jclass; typedef jobject jstring; typedef jobject jarray; typedef jarray jintArray; typedef jarray jlongArray;
typedef jarray jdoubleArray; typedef jarray jbyteArray; typedef jarray jfloatArray;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    jstring (*NewStringUTF)(JNIEnv*, const char*);
    jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);
    void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
    jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
    jsize (*GetArrayLength)(JNIEnv*, jarray);
    void* (*GetPrimitiveArrayCritical)(JNIEnv*, jarray, jboolean*);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv*, jarray, void*, jint);
    void (*SetDoubleArrayRegion)(JNIEnv*, jdoubleArray, jsize, jsize, const jdouble*);
    void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
    void (*SetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, const jint*);
    void (*SetFloatArrayRegion)(JNIEnv*, jfloatArray, jsize, jsize, const jfloat*);
    void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
    void (*GetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*);
};
#endif
