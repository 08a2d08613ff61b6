/*
 * Minimal stand-in for <jni.h>, ONLY for building and testing lagassign_jni.c on a
 * machine without a JDK (this repository's build image has none). It declares the
 * handful of types the shim uses and a JNIEnv whose function table has the shim's
 * functions at the slot numbers of the JNI specification's interface function table
 * (FindClass 6, ThrowNew 14, GetArrayLength 171, GetIntArrayRegion 203,
 * NewDirectByteBuffer 229, GetDirectBufferAddress 230, GetDirectBufferCapacity 231). The slot numbers were written from the published
 * specification and have NOT been checked against a real JDK here; a production build
 * uses the JDK's own <jni.h> (see the build line in lagassign_jni.c).
 */
#ifndef LAGASSIGN_FAKE_JNI_H
#define LAGASSIGN_FAKE_JNI_H

#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

typedef int32_t jint;
typedef int64_t jlong;
typedef void *jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef uint8_t jboolean;
typedef jint jsize;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;

struct JNINativeInterface_ {
    void *slot_0_5[6];
    jclass (*FindClass)(JNIEnv *env, const char *name);                               /* 6 */
    void *slot_7_13[7];
    jint (*ThrowNew)(JNIEnv *env, jclass clazz, const char *msg);                     /* 14 */
    void *slot_15_170[156];
    jsize (*GetArrayLength)(JNIEnv *env, jarray array);                               /* 171 */
    void *slot_172_202[31];
    void (*GetIntArrayRegion)(JNIEnv *env, jintArray array, jsize start, jsize len, jint *buf);   /* 203 */
    void *slot_204_228[25];
    jobject (*NewDirectByteBuffer)(JNIEnv *env, void *address, jlong capacity);       /* 229 */
    void *(*GetDirectBufferAddress)(JNIEnv *env, jobject buf);                        /* 230 */
    jlong (*GetDirectBufferCapacity)(JNIEnv *env, jobject buf);                       /* 231 */
    void *slot_232_234[3];
};

#endif
