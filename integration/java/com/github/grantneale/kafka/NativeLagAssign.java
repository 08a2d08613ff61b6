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

    /** lagassign_create_multi: one context that shards the topics of every call over the listed devices. */
    static native long createMulti(int[] devices);

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

    /**
     * The same call fed with raw offsets instead of lags (lagassign_upload_offsets + execute + download):
     * computePartitionLag of LagBasedPartitionAssignor.java:376-404 runs on the device.
     * {@code hasCommitted} holds one byte per partition (committed metadata != null, :384);
     * {@code resetLatest} = auto.offset.reset equalsIgnoreCase "latest" (:391).
     */
    static native void assignBatchOffsets(long ctx,
                                          long nTopics, ByteBuffer topicOff, ByteBuffer partition,
                                          ByteBuffer topicCset, long nCsets, ByteBuffer csetOff, ByteBuffer csetRank,
                                          ByteBuffer beginOffset, ByteBuffer endOffset, ByteBuffer committed,
                                          ByteBuffer hasCommitted, boolean resetLatest,
                                          ByteBuffer outPartition, ByteBuffer outConsumerRank);

    /**
     * lagassign_assign_batch_narrow: like {@link #assignBatch}, but {@code outPartition} receives entries of
     * {@code w & 0xFF} bytes (1, 2 or 4, unsigned) and {@code outConsumerRank} entries of {@code w >> 8} bytes
     * (0: nothing written, every topic has P <= C and sorted partition k belongs to the k-th member of the topic's
     * consumer set; 2: unsigned, 0xFFFF = no subscriber; 4: int), where {@code w} is the return value. Both buffers
     * need 4 bytes per partition of capacity.
     */
    static native int assignBatchNarrow(long ctx,
                                        long nTopics, ByteBuffer topicOff, ByteBuffer lag, ByteBuffer partition,
                                        ByteBuffer topicCset, long nCsets, ByteBuffer csetOff, ByteBuffer csetRank,
                                        ByteBuffer outPartition, ByteBuffer outConsumerRank);
}
