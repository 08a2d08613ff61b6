/*
 * Fake-JNIEnv harness for integration/jni/lagassign_jni.c (test infrastructure).
 * A "direct ByteBuffer" is a small struct; FindClass/ThrowNew record the pending
 * exception. Prints one line per check; exit code 0 on success.
 *
 *   harness nogpu   -> create() must throw IllegalStateException (no CPU fallback)
 *   harness gpu     -> the README example of the reference through assignBatch()
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jni.h"

typedef struct { void *addr; jlong cap; } fake_buffer;

static char pending_class[128];
static char pending_msg[512];

static jclass f_FindClass(JNIEnv *env, const char *name) { (void)env; return (jclass)strdup(name); }
static jint f_ThrowNew(JNIEnv *env, jclass c, const char *msg) {
    (void)env;
    snprintf(pending_class, sizeof pending_class, "%s", (const char *)c);
    snprintf(pending_msg, sizeof pending_msg, "%s", msg ? msg : "");
    return 0;
}
static jobject f_NewDirectByteBuffer(JNIEnv *env, void *a, jlong cap) {
    (void)env;
    fake_buffer *b = (fake_buffer *)malloc(sizeof *b);
    b->addr = a; b->cap = cap;
    return b;
}
static void *f_GetDirectBufferAddress(JNIEnv *env, jobject buf) { (void)env; return ((fake_buffer *)buf)->addr; }
static jlong f_GetDirectBufferCapacity(JNIEnv *env, jobject buf) { (void)env; return ((fake_buffer *)buf)->cap; }

jlong Java_com_github_grantneale_kafka_NativeLagAssign_create(JNIEnv *, jclass, jint);
void Java_com_github_grantneale_kafka_NativeLagAssign_destroy(JNIEnv *, jclass, jlong);
jobject Java_com_github_grantneale_kafka_NativeLagAssign_allocPinned(JNIEnv *, jclass, jlong);
void Java_com_github_grantneale_kafka_NativeLagAssign_freePinned(JNIEnv *, jclass, jobject);
void Java_com_github_grantneale_kafka_NativeLagAssign_assignBatch(JNIEnv *, jclass, jlong, jlong, jobject, jobject, jobject,
                                                                  jobject, jlong, jobject, jobject, jobject, jobject);

jlong Java_com_github_grantneale_kafka_NativeLagAssign_createMulti(JNIEnv *, jclass, jintArray);
void Java_com_github_grantneale_kafka_NativeLagAssign_assignBatchOffsets(JNIEnv *, jclass, jlong, jlong, jobject, jobject, jobject, jlong,
                                                                         jobject, jobject, jobject, jobject, jobject, jobject, jboolean,
                                                                         jobject, jobject);

jint Java_com_github_grantneale_kafka_NativeLagAssign_assignBatchNarrow(JNIEnv *, jclass, jlong, jlong, jobject, jobject, jobject, jobject,
                                                                        jlong, jobject, jobject, jobject, jobject);

/* a fake int[]: length, then the elements */
typedef struct { jsize n; jint v[8]; } fake_int_array;
static jsize f_GetArrayLength(JNIEnv *env, jarray a) { (void)env; return ((fake_int_array *)a)->n; }
static void f_GetIntArrayRegion(JNIEnv *env, jintArray a, jsize start, jsize len, jint *buf) {
    (void)env;
    memcpy(buf, ((fake_int_array *)a)->v + start, (size_t)len * sizeof(jint));
}

static jobject wrap(JNIEnv *env, void *p, size_t n) { return f_NewDirectByteBuffer(env, p, (jlong)n); }

int main(int argc, char **argv) {
    static struct JNINativeInterface_ table;
    table.FindClass = f_FindClass;
    table.ThrowNew = f_ThrowNew;
    table.NewDirectByteBuffer = f_NewDirectByteBuffer;
    table.GetDirectBufferAddress = f_GetDirectBufferAddress;
    table.GetDirectBufferCapacity = f_GetDirectBufferCapacity;
    table.GetArrayLength = f_GetArrayLength;
    table.GetIntArrayRegion = f_GetIntArrayRegion;
    JNIEnv env_obj = &table;
    JNIEnv *env = &env_obj;
    const char *mode = argc > 1 ? argv[1] : "nogpu";

    jlong ctx = Java_com_github_grantneale_kafka_NativeLagAssign_create(env, NULL, 0);
    if (strcmp(mode, "nogpu") == 0) {
        if (ctx != 0 || strcmp(pending_class, "java/lang/IllegalStateException") != 0) {
            printf("FAIL: expected IllegalStateException, got ctx=%lld class=%s\n", (long long)ctx, pending_class);
            return 1;
        }
        printf("ok: create() without a GPU throws %s: %s\n", pending_class, pending_msg);
        return 0;
    }
    if (ctx == 0) { printf("FAIL: create: %s: %s\n", pending_class, pending_msg); return 1; }

    /* README.md:42-57 of the reference: t0 lags 100000 / 50000 / 60000, members C0, C1 */
    int64_t topic_off[2] = {0, 3}, lag[3] = {100000, 50000, 60000}, cset_off[2] = {0, 2};
    int32_t topic_cset[1] = {0}, cset_rank[2] = {0, 1}, out_p[3] = {-9, -9, -9}, out_c[3] = {-9, -9, -9};
    Java_com_github_grantneale_kafka_NativeLagAssign_assignBatch(
        env, NULL, ctx, 1, wrap(env, topic_off, sizeof topic_off), wrap(env, lag, sizeof lag), NULL,
        wrap(env, topic_cset, sizeof topic_cset), 1, wrap(env, cset_off, sizeof cset_off),
        wrap(env, cset_rank, sizeof cset_rank), wrap(env, out_p, sizeof out_p), wrap(env, out_c, sizeof out_c));
    if (pending_class[0]) { printf("FAIL: assignBatch threw %s: %s\n", pending_class, pending_msg); return 1; }
    const int32_t exp_p[3] = {0, 2, 1}, exp_c[3] = {0, 1, 1};
    if (memcmp(out_p, exp_p, sizeof exp_p) || memcmp(out_c, exp_c, sizeof exp_c)) {
        printf("FAIL: got partitions %d %d %d consumers %d %d %d\n", out_p[0], out_p[1], out_p[2], out_c[0], out_c[1], out_c[2]);
        return 1;
    }
    printf("ok: README example through the JNI shim: C0=[p0] C1=[p2,p1]\n");

    /* the same example in the narrow result form: 3 partitions over 2 consumers is a multi-round fill -> 1-byte ids, 2-byte consumers */
    {
        uint8_t np[12] = {0}, nc[12] = {0};
        jint w = Java_com_github_grantneale_kafka_NativeLagAssign_assignBatchNarrow(
            env, NULL, ctx, 1, wrap(env, topic_off, sizeof topic_off), wrap(env, lag, sizeof lag), NULL,
            wrap(env, topic_cset, sizeof topic_cset), 1, wrap(env, cset_off, sizeof cset_off),
            wrap(env, cset_rank, sizeof cset_rank), wrap(env, np, sizeof np), wrap(env, nc, sizeof nc));
        if (pending_class[0]) { printf("FAIL: assignBatchNarrow threw %s: %s\n", pending_class, pending_msg); return 1; }
        const uint16_t *c16 = (const uint16_t *)nc;
        if (w != (1 | (2 << 8)) || np[0] != 0 || np[1] != 2 || np[2] != 1 || c16[0] != 0 || c16[1] != 1 || c16[2] != 1) {
            printf("FAIL: narrow form: widths %d, ids %d %d %d, consumers %d %d %d\n", (int)w, np[0], np[1], np[2], c16[0], c16[1], c16[2]);
            return 1;
        }
        printf("ok: README example in the narrow result form (1-byte ids, 2-byte consumers)\n");
    }

    /* error mapping: a malformed batch must surface as RuntimeException */
    int32_t bad_rank[2] = {1, 1};
    Java_com_github_grantneale_kafka_NativeLagAssign_assignBatch(
        env, NULL, ctx, 1, wrap(env, topic_off, sizeof topic_off), wrap(env, lag, sizeof lag), NULL,
        wrap(env, topic_cset, sizeof topic_cset), 1, wrap(env, cset_off, sizeof cset_off),
        wrap(env, bad_rank, sizeof bad_rank), wrap(env, out_p, sizeof out_p), wrap(env, out_c, sizeof out_c));
    if (strcmp(pending_class, "java/lang/RuntimeException") != 0) { printf("FAIL: expected RuntimeException, got '%s'\n", pending_class); return 1; }
    printf("ok: invalid input throws %s: %s\n", pending_class, pending_msg);

    /* createMulti over {0, 0} (two shards on one GPU) + raw offsets: two topics so that each shard gets one.
     * LagBasedPartitionAssignorTest.java:21-80 vectors as the offsets of topic A: lags 4444, 0, 0 (latest), and
     * topic B: begin 1111 end 9999 no commit, reset earliest would give 8888 but this call passes latest -> 0;
     * committed 5000 of end 9999 -> 4999; committed beyond the end -> 0 (the max() of Main:402). */
    {
        fake_int_array devs = {2, {0, 0}};
        pending_class[0] = 0;
        jlong mctx = Java_com_github_grantneale_kafka_NativeLagAssign_createMulti(env, NULL, &devs);
        if (mctx == 0) { printf("FAIL: createMulti: %s: %s\n", pending_class, pending_msg); return 1; }
        int64_t toff[3] = {0, 3, 6}, coff[2] = {0, 2};
        int32_t tcs[2] = {0, 0}, cr[2] = {0, 1}, op[6], oc[6];
        int64_t beg[6] = {1111, 0, 1111, 1111, 0, 0}, end[6] = {9999, 0, 9999, 9999, 9999, 100};
        int64_t com[6] = {5555, 5555, 0, 0, 5000, 200};
        uint8_t has[6] = {1, 1, 0, 0, 1, 1};
        Java_com_github_grantneale_kafka_NativeLagAssign_assignBatchOffsets(
            env, NULL, mctx, 2, wrap(env, toff, sizeof toff), NULL, wrap(env, tcs, sizeof tcs), 1, wrap(env, coff, sizeof coff),
            wrap(env, cr, sizeof cr), wrap(env, beg, sizeof beg), wrap(env, end, sizeof end), wrap(env, com, sizeof com),
            wrap(env, has, sizeof has), 1, wrap(env, op, sizeof op), wrap(env, oc, sizeof oc));
        if (pending_class[0]) { printf("FAIL: assignBatchOffsets threw %s: %s\n", pending_class, pending_msg); return 1; }
        /* topic A lags 4444, 0, 0 -> order p0, p1, p2 -> C0, C1, C1 (p2: equal counts, C1 has the smaller total); topic B lags 0, 4999, 0 -> order p1, p0, p2 -> C0, C1, C1 */
        const int32_t ep[6] = {0, 1, 2, 1, 0, 2}, ec[6] = {0, 1, 1, 0, 1, 1};
        if (memcmp(op, ep, sizeof ep) || memcmp(oc, ec, sizeof ec)) {
            printf("FAIL: offsets path got partitions %d %d %d %d %d %d consumers %d %d %d %d %d %d\n",
                   op[0], op[1], op[2], op[3], op[4], op[5], oc[0], oc[1], oc[2], oc[3], oc[4], oc[5]);
            return 1;
        }
        Java_com_github_grantneale_kafka_NativeLagAssign_destroy(env, NULL, mctx);
        printf("ok: createMulti + assignBatchOffsets (fused computePartitionLag) through the JNI shim\n");
    }

    jobject pin = Java_com_github_grantneale_kafka_NativeLagAssign_allocPinned(env, NULL, 4096);
    if (!pin || !f_GetDirectBufferAddress(env, pin)) { printf("FAIL: allocPinned\n"); return 1; }
    Java_com_github_grantneale_kafka_NativeLagAssign_freePinned(env, NULL, pin);
    Java_com_github_grantneale_kafka_NativeLagAssign_destroy(env, NULL, ctx);
    printf("ok: pinned buffer + destroy\n");
    return 0;
}
