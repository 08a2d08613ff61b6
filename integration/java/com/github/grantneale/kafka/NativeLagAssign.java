package com.github.grantneale.kafka;

import java.nio.ByteBuffer;

/**
 * JNI binding of include/lagassign.h. NOT COMPILED IN THIS REPOSITORY: the build
 * image has no JDK. It is the stub a maintainer of the reference adds next to
 * LagBasedPartitionAssignor.java; integration/jni/lagassign_jni.c implements it.
 *
 * All buffers are direct ByteBuffers in native byte order (ideally the pinned
 * ones returned by {@link #allocPinned}) laid out as include/lagassign.h says.
 */
final class NativeLagAssign {

    static {
        System.loadLibrary("lagassign_jni");   // links liblagassign.so
    }

    private NativeLagAssign() {
    }

    /** lagassign_create: returns the context handle; throws IllegalStateException when no GPU is usable. */
    static native long create(int device);

    static native void destroy(long ctx);

    /** lagassign_host_alloc wrapped by NewDirectByteBuffer. */
    static native ByteBuffer allocPinned(long bytes);

    static native void freePinned(ByteBuffer buffer);

    /**
     * lagassign_assign_batch. {@code partition} may be null (ids 0..P-1).
     * Throws RuntimeException carrying lagassign_last_error() on any failure, so a
     * failed GPU call fails the rebalance loudly; there is no CPU fallback.
     */
    static native void assignBatch(long ctx,
                                   long nTopics, ByteBuffer topicOff, ByteBuffer lag, ByteBuffer partition,
                                   ByteBuffer topicCset, long nCsets, ByteBuffer csetOff, ByteBuffer csetRank,
                                   ByteBuffer outPartition, ByteBuffer outConsumerRank);
}
