/*
 * lagassign_jni.c — JNI shim between NativeLagAssign.java and include/lagassign.h.
 * NOT BUILT HERE (no JDK / jni.h in the image). Build on a machine with a JDK:
 *
 *   gcc -O2 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       -o liblagassign_jni.so lagassign_jni.c -L../../kafka-lag-based-assignor_b200 -llagassign
 *
 * It replaces nothing in the reference by itself: the Java side calls it where
 * LagBasedPartitionAssignor.java:149 calls the static assign(Map, Map).
 */
#include <jni.h>
#include <stdint.h>

#include "lagassign.h"

static void throw_runtime(JNIEnv *env, const char *cls, const char *msg) {
    jclass c = (*env)->FindClass(env, cls);
    if (c) (*env)->ThrowNew(env, c, msg);
}

static void *addr(JNIEnv *env, jobject buf) {
    return buf ? (*env)->GetDirectBufferAddress(env, buf) : NULL;
}

JNIEXPORT jlong JNICALL Java_com_github_grantneale_kafka_NativeLagAssign_create(JNIEnv *env, jclass k, jint device) {
    (void)k;
    lagassign_ctx *ctx = NULL;
    int rc = lagassign_create(&ctx, device);
    if (rc != LAGASSIGN_OK) {
        throw_runtime(env, "java/lang/IllegalStateException", lagassign_last_error(NULL));
        return 0;
    }
    return (jlong)(intptr_t)ctx;
}

/* lagassign_create_multi: topics sharded over several devices of this process (SURVEY 8e through the plugin boundary) */
JNIEXPORT jlong JNICALL Java_com_github_grantneale_kafka_NativeLagAssign_createMulti(JNIEnv *env, jclass k, jintArray devices) {
    (void)k;
    jint dev[64];
    jsize n = devices ? (*env)->GetArrayLength(env, devices) : 0;
    if (n < 1 || n > 64) {
        throw_runtime(env, "java/lang/IllegalArgumentException", "createMulti: 1..64 devices expected");
        return 0;
    }
    (*env)->GetIntArrayRegion(env, devices, 0, n, dev);
    int d[64];
    for (jsize i = 0; i < n; i++) d[i] = (int)dev[i];
    lagassign_ctx *ctx = NULL;
    int rc = lagassign_create_multi(&ctx, d, (int)n);
    if (rc != LAGASSIGN_OK) {
        throw_runtime(env, "java/lang/IllegalStateException", lagassign_last_error(NULL));
        return 0;
    }
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL Java_com_github_grantneale_kafka_NativeLagAssign_destroy(JNIEnv *env, jclass k, jlong ctx) {
    (void)env; (void)k;
    lagassign_destroy((lagassign_ctx *)(intptr_t)ctx);
}

JNIEXPORT jobject JNICALL Java_com_github_grantneale_kafka_NativeLagAssign_allocPinned(JNIEnv *env, jclass k, jlong bytes) {
    (void)k;
    void *p = NULL;
    if (lagassign_host_alloc(&p, (size_t)bytes) != LAGASSIGN_OK) {
        throw_runtime(env, "java/lang/OutOfMemoryError", lagassign_last_error(NULL));
        return NULL;
    }
    return (*env)->NewDirectByteBuffer(env, p, bytes);
}

JNIEXPORT void JNICALL Java_com_github_grantneale_kafka_NativeLagAssign_freePinned(JNIEnv *env, jclass k, jobject buf) {
    (void)k;
    lagassign_host_free(addr(env, buf));
}

JNIEXPORT void JNICALL Java_com_github_grantneale_kafka_NativeLagAssign_assignBatch(
    JNIEnv *env, jclass k, jlong ctxh, jlong nTopics, jobject topicOff, jobject lag, jobject partition,
    jobject topicCset, jlong nCsets, jobject csetOff, jobject csetRank, jobject outPartition, jobject outConsumerRank) {
    (void)k;
    lagassign_ctx *ctx = (lagassign_ctx *)(intptr_t)ctxh;
    lagassign_batch in;
    in.n_topics = nTopics;
    in.topic_off = (const int64_t *)addr(env, topicOff);
    in.lag = (const int64_t *)addr(env, lag);
    in.partition = (const int32_t *)addr(env, partition);
    in.topic_cset = (const int32_t *)addr(env, topicCset);
    in.n_csets = nCsets;
    in.cset_off = (const int64_t *)addr(env, csetOff);
    in.cset_rank = (const int32_t *)addr(env, csetRank);
    lagassign_result out;
    out.out_partition = (int32_t *)addr(env, outPartition);
    out.out_consumer_rank = (int32_t *)addr(env, outConsumerRank);
    out.out_total = NULL;
    out.out_count = NULL;
    int rc = lagassign_assign_batch(ctx, &in, &out);
    if (rc != LAGASSIGN_OK) throw_runtime(env, "java/lang/RuntimeException", lagassign_last_error(ctx));
}

/* The same call fed with raw offsets: computePartitionLag (LagBasedPartitionAssignor.java:376-404, call site :344-356)
 * runs on the device, fused into the kernels' key build (lagassign_upload_offsets). */
JNIEXPORT void JNICALL Java_com_github_grantneale_kafka_NativeLagAssign_assignBatchOffsets(
    JNIEnv *env, jclass k, jlong ctxh, jlong nTopics, jobject topicOff, jobject partition,
    jobject topicCset, jlong nCsets, jobject csetOff, jobject csetRank,
    jobject beginOffset, jobject endOffset, jobject committed, jobject hasCommitted, jboolean resetLatest,
    jobject outPartition, jobject outConsumerRank) {
    (void)k;
    lagassign_ctx *ctx = (lagassign_ctx *)(intptr_t)ctxh;
    lagassign_batch in;
    in.n_topics = nTopics;
    in.topic_off = (const int64_t *)addr(env, topicOff);
    in.lag = NULL;
    in.partition = (const int32_t *)addr(env, partition);
    in.topic_cset = (const int32_t *)addr(env, topicCset);
    in.n_csets = nCsets;
    in.cset_off = (const int64_t *)addr(env, csetOff);
    in.cset_rank = (const int32_t *)addr(env, csetRank);
    lagassign_offsets offs;
    offs.begin_offset = (const int64_t *)addr(env, beginOffset);
    offs.end_offset = (const int64_t *)addr(env, endOffset);
    offs.committed = (const int64_t *)addr(env, committed);
    offs.has_committed = (const uint8_t *)addr(env, hasCommitted);
    offs.reset_latest = resetLatest ? 1 : 0;
    lagassign_result out;
    out.out_partition = (int32_t *)addr(env, outPartition);
    out.out_consumer_rank = (int32_t *)addr(env, outConsumerRank);
    out.out_total = NULL;
    out.out_count = NULL;
    int rc = lagassign_upload_offsets(ctx, &in, &offs);
    if (rc == LAGASSIGN_OK) rc = lagassign_execute(ctx);
    if (rc == LAGASSIGN_OK) rc = lagassign_download(ctx, &out);
    if (rc != LAGASSIGN_OK) throw_runtime(env, "java/lang/RuntimeException", lagassign_last_error(ctx));
}

/* lagassign_assign_batch_narrow: results as entries of part_width / cons_width bytes (include/lagassign.h); returns
 * part_width | cons_width << 8. The JVM reads bytes from a direct buffer as cheaply as ints, and at BASELINE config 4
 * the device-to-host copy shrinks from 8 bytes per partition to 1. */
JNIEXPORT jint JNICALL Java_com_github_grantneale_kafka_NativeLagAssign_assignBatchNarrow(
    JNIEnv *env, jclass k, jlong ctxh, jlong nTopics, jobject topicOff, jobject lag, jobject partition,
    jobject topicCset, jlong nCsets, jobject csetOff, jobject csetRank, jobject outPartition, jobject outConsumerRank) {
    (void)k;
    lagassign_ctx *ctx = (lagassign_ctx *)(intptr_t)ctxh;
    lagassign_batch in;
    in.n_topics = nTopics;
    in.topic_off = (const int64_t *)addr(env, topicOff);
    in.lag = (const int64_t *)addr(env, lag);
    in.partition = (const int32_t *)addr(env, partition);
    in.topic_cset = (const int32_t *)addr(env, topicCset);
    in.n_csets = nCsets;
    in.cset_off = (const int64_t *)addr(env, csetOff);
    in.cset_rank = (const int32_t *)addr(env, csetRank);
    lagassign_result_narrow out;
    out.out_partition = addr(env, outPartition);
    out.out_consumer_rank = addr(env, outConsumerRank);
    out.part_width = out.cons_width = 4;
    int rc = lagassign_assign_batch_narrow(ctx, &in, &out);
    if (rc != LAGASSIGN_OK) {
        throw_runtime(env, "java/lang/RuntimeException", lagassign_last_error(ctx));
        return 0;
    }
    return (jint)(out.part_width | (out.cons_width << 8));
}
