package com.github.grantneale.kafka;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.Collections;
import java.util.HashMap;
import java.util.HashSet;
import java.util.List;
import java.util.Map;
import java.util.Properties;
import java.util.TreeSet;
import org.apache.kafka.clients.consumer.ConsumerConfig;
import org.apache.kafka.clients.consumer.ConsumerPartitionAssignor;
import org.apache.kafka.clients.consumer.KafkaConsumer;
import org.apache.kafka.clients.consumer.OffsetAndMetadata;
import org.apache.kafka.common.Cluster;
import org.apache.kafka.common.Configurable;
import org.apache.kafka.common.PartitionInfo;
import org.apache.kafka.common.TopicPartition;

/**
 * The plugin class of the GPU path (NOT COMPILED HERE: the build image has no JDK, no kafka-clients jar).
 *
 * <p>It is loaded exactly like the reference class, by
 * {@code partition.assignment.strategy=com.github.grantneale.kafka.GpuLagBasedPartitionAssignor}, and keeps the
 * reference's plugin surface: {@link #configure} (LagBasedPartitionAssignor.java:97-130, {@code group.id} required),
 * {@link #name()} = {@code "lag"} (:132-135; same protocol name, so groups may mix both classes during a rolling
 * upgrade: the results are identical), {@link #assign(Cluster, GroupSubscription)} (:137-157).
 *
 * <p>Differences behind that surface, none of them visible in the result:
 * <ul>
 * <li>the per-topic loop of the static {@code assign(Map, Map)} (:166-188, {@code assignTopic} :204-277) is ONE
 *     native call over all topics ({@link NativeLagAssign#assignBatch} / {@link NativeLagAssign#assignBatchOffsets});
 * <li>offsets are read with THREE broker calls for the whole group instead of three per topic (:327-342):
 *     {@code beginningOffsets}, {@code endOffsets}, {@code committed} each take the partitions of every subscribed
 *     topic at once (SURVEY 8f row 4);
 * <li>{@code computePartitionLag} (:376-404) runs on the device from the raw offsets (SURVEY 8f row 1): the Java side
 *     only writes begin / end / committed / has-committed into pinned buffers;
 * <li>partition ids are not shipped when a topic's ids are 0..P-1 in order (what {@code Cluster.partitionsForTopic}
 *     normally yields): the C ABI's {@code partition == NULL} fast path.
 * </ul>
 * The C++ class csrc/host/lag_based_partition_assignor.cpp performs the same flatten / unflatten and IS tested; this
 * file differs from it only in using a real {@code java.util.HashMap} for the topic order.
 */
public class GpuLagBasedPartitionAssignor implements ConsumerPartitionAssignor, Configurable, AutoCloseable {

    /** Devices of this box to shard topics over, e.g. "0" or "0,1,2,3,4,5,6,7" (lagassign_create_multi). */
    public static final String DEVICES_CONFIG = "lag.assignor.gpu.devices";

    private Properties consumerGroupProps;
    private Properties metadataConsumerProps;
    private KafkaConsumer<byte[], byte[]> metadataConsumer;
    private int[] devices = {0};
    private long ctx = 0;                                   // lagassign_ctx*, created on first use, freed by close()
    private final Map<String, ByteBuffer> pinned = new HashMap<>();   // reused across rebalances (grown on demand)

    @Override
    public void configure(Map<String, ?> configs) {
        consumerGroupProps = new Properties();
        for (final Map.Entry<String, ?> prop : configs.entrySet()) {
            if (prop.getValue() != null) {
                consumerGroupProps.put(prop.getKey(), prop.getValue());
            }
        }
        final String groupId = consumerGroupProps.getProperty(ConsumerConfig.GROUP_ID_CONFIG);
        if (groupId == null) {                                                               // Main:107-113
            throw new IllegalArgumentException(
                ConsumerConfig.GROUP_ID_CONFIG + " cannot be null when using "
                    + ConsumerConfig.PARTITION_ASSIGNMENT_STRATEGY_CONFIG + "=" + this.getClass().getName());
        }
        metadataConsumerProps = new Properties();                                            // Main:116-120
        metadataConsumerProps.putAll(consumerGroupProps);
        metadataConsumerProps.remove(DEVICES_CONFIG);
        metadataConsumerProps.put(ConsumerConfig.ENABLE_AUTO_COMMIT_CONFIG, "false");
        metadataConsumerProps.put(ConsumerConfig.CLIENT_ID_CONFIG, groupId + ".assignor");
        final Object dev = configs.get(DEVICES_CONFIG);
        if (dev != null) {
            final String[] parts = dev.toString().split(",");
            devices = new int[parts.length];
            for (int i = 0; i < parts.length; i++) {
                devices[i] = Integer.parseInt(parts[i].trim());
            }
        }
    }

    @Override
    public String name() {
        return "lag";                                                                        // Main:132-135
    }

    @Override
    public synchronized GroupAssignment assign(Cluster metadata, GroupSubscription subscriptions) {
        final Map<String, List<String>> topicSubscriptions = new HashMap<>();                // Main:140-147
        for (Map.Entry<String, Subscription> e : subscriptions.groupSubscription().entrySet()) {
            topicSubscriptions.put(e.getKey(), e.getValue().topics());
        }
        final Flat flat = flatten(topicSubscriptions, topic -> {
            final List<PartitionInfo> infos = metadata.partitionsForTopic(topic);            // Main:329-330
            if (infos == null || infos.isEmpty()) {
                return null;                                                                 // "Skipping assignment": Main:359-361
            }
            final int[] ids = new int[infos.size()];
            for (int i = 0; i < ids.length; i++) {
                ids[i] = infos.get(i).partition();
            }
            return ids;
        });
        final Map<String, List<TopicPartition>> raw = emptyAssignment(topicSubscriptions);
        if (flat.n > 0) {
            readOffsetsBatched(flat);
            runNative(flat, /*offsets=*/true);
            unflatten(flat, raw);
        }
        final Map<String, Assignment> assignments = new HashMap<>();                         // Main:152-156
        for (Map.Entry<String, List<TopicPartition>> e : raw.entrySet()) {
            assignments.put(e.getKey(), new Assignment(e.getValue()));
        }
        return new GroupAssignment(assignments);
    }

    /**
     * Same contract as the reference's package-private static seam (Main:166-169), as an instance method because it
     * owns a device context; the reference's unit tests (LagBasedPartitionAssignorTest.java:82-228) run against it
     * by calling it on an instance.
     */
    synchronized Map<String, List<TopicPartition>> assign(
        Map<String, List<LagBasedPartitionAssignor.TopicPartitionLag>> partitionLagPerTopic,
        Map<String, List<String>> subscriptions) {

        final Flat flat = flatten(subscriptions, topic -> {
            final List<LagBasedPartitionAssignor.TopicPartitionLag> l =
                partitionLagPerTopic.getOrDefault(topic, Collections.emptyList());           // Main:182
            final int[] ids = new int[l.size()];
            for (int i = 0; i < ids.length; i++) {
                ids[i] = l.get(i).getPartition();
            }
            return ids;
        });
        final Map<String, List<TopicPartition>> assignment = emptyAssignment(subscriptions);
        if (flat.n == 0) {
            return assignment;
        }
        final ByteBuffer lag = pinned("lag", 8L * flat.n);
        int k = 0;
        for (String t : flat.topics) {
            for (LagBasedPartitionAssignor.TopicPartitionLag p : partitionLagPerTopic.getOrDefault(t, Collections.emptyList())) {
                lag.putLong(8 * k++, p.getLag());
            }
        }
        runNative(flat, /*offsets=*/false);
        unflatten(flat, assignment);
        return assignment;
    }

    @Override
    public synchronized void close() {
        if (ctx != 0) {
            NativeLagAssign.destroy(ctx);
            ctx = 0;
        }
        for (ByteBuffer b : pinned.values()) {
            NativeLagAssign.freePinned(b);
        }
        pinned.clear();
        if (metadataConsumer != null) {
            metadataConsumer.close();
            metadataConsumer = null;
        }
    }

    // ---------------------------------------------------------------------------------------------------------

    private interface PartitionIds {
        int[] of(String topic);   // null: topic unknown, not assigned
    }

    /** The flat batch of include/lagassign.h plus what unflatten needs. */
    private static final class Flat {
        String[] byRank;                  // member ids in String.compareTo order (Main:258-259)
        List<String> topics;              // iteration order of consumersPerTopic's HashMap (Main:176-177)
        long[] topicOff;
        int[][] ids;                      // partition ids per topic, input order
        boolean explicitIds;              // some topic's ids are not 0..P-1
        int[] topicCset;
        List<int[]> sets;                 // de-duplicated consumer sets (ranks ascending)
        long n;
        int partWidth = 4;                // bytes per entry of the result buffers (NativeLagAssign.assignBatchNarrow)
        int consWidth = 4;
    }

    private static Map<String, List<TopicPartition>> emptyAssignment(Map<String, List<String>> subscriptions) {
        final Map<String, List<TopicPartition>> assignment = new HashMap<>();                // Main:171-174
        for (String memberId : subscriptions.keySet()) {
            assignment.put(memberId, new ArrayList<>());
        }
        return assignment;
    }

    private static Flat flatten(Map<String, List<String>> subscriptions, PartitionIds partitionIds) {
        final Flat f = new Flat();
        f.byRank = subscriptions.keySet().toArray(new String[0]);
        Arrays.sort(f.byRank);
        final Map<String, Integer> rank = new HashMap<>();
        for (int i = 0; i < f.byRank.length; i++) {
            rank.put(f.byRank[i], i);
        }
        // consumersPerTopic (Main:410-426) with a real HashMap: its key order is the order in which the reference
        // visits topics and therefore the order of every member's result list
        final Map<String, TreeSet<Integer>> consumersPerTopic = new HashMap<>();
        for (Map.Entry<String, List<String>> e : subscriptions.entrySet()) {
            for (String topic : e.getValue()) {
                consumersPerTopic.computeIfAbsent(topic, t -> new TreeSet<>()).add(rank.get(e.getKey()));
            }
        }
        f.topics = new ArrayList<>();
        final List<int[]> idsPerTopic = new ArrayList<>();
        final List<Integer> csetOfTopic = new ArrayList<>();
        final Map<List<Integer>, Integer> setIds = new HashMap<>();
        f.sets = new ArrayList<>();
        for (Map.Entry<String, TreeSet<Integer>> e : consumersPerTopic.entrySet()) {
            final int[] ids = partitionIds.of(e.getKey());
            if (ids == null) {
                continue;
            }
            f.topics.add(e.getKey());
            idsPerTopic.add(ids);
            for (int i = 0; i < ids.length && !f.explicitIds; i++) {
                f.explicitIds = ids[i] != i;
            }
            final List<Integer> set = new ArrayList<>(e.getValue());
            Integer id = setIds.get(set);
            if (id == null) {
                id = f.sets.size();
                setIds.put(set, id);
                final int[] s = new int[set.size()];
                for (int i = 0; i < s.length; i++) {
                    s[i] = set.get(i);
                }
                f.sets.add(s);
            }
            csetOfTopic.add(id);
            f.n += ids.length;
        }
        final int T = f.topics.size();
        f.ids = idsPerTopic.toArray(new int[0][]);
        f.topicOff = new long[T + 1];
        f.topicCset = new int[T];
        for (int t = 0; t < T; t++) {
            f.topicOff[t + 1] = f.topicOff[t] + f.ids[t].length;
            f.topicCset[t] = csetOfTopic.get(t);
        }
        return f;
    }

    /**
     * Three broker round trips for the whole group (the reference makes three per topic, Main:338-342), written
     * straight into the pinned offset buffers the device reads; a missing begin / end offset is 0 as at Main:350-351.
     */
    private void readOffsetsBatched(Flat f) {
        if (metadataConsumer == null) {
            metadataConsumer = new KafkaConsumer<>(metadataConsumerProps);                   // Main:322-324
        }
        final List<TopicPartition> all = new ArrayList<>((int) Math.min(f.n, Integer.MAX_VALUE - 8));
        for (int t = 0; t < f.topics.size(); t++) {
            for (int id : f.ids[t]) {
                all.add(new TopicPartition(f.topics.get(t), id));
            }
        }
        final Map<TopicPartition, Long> begin = metadataConsumer.beginningOffsets(all);
        final Map<TopicPartition, Long> end = metadataConsumer.endOffsets(all);
        final Map<TopicPartition, OffsetAndMetadata> committed = metadataConsumer.committed(new HashSet<>(all));
        final ByteBuffer b = pinned("begin", 8L * f.n), e = pinned("end", 8L * f.n);
        final ByteBuffer c = pinned("committed", 8L * f.n), h = pinned("has", f.n);
        int k = 0;
        for (TopicPartition tp : all) {
            final OffsetAndMetadata m = committed.get(tp);
            b.putLong(8 * k, begin.getOrDefault(tp, 0L));
            e.putLong(8 * k, end.getOrDefault(tp, 0L));
            c.putLong(8 * k, m != null ? m.offset() : 0L);
            h.put(k, (byte) (m != null ? 1 : 0));
            k++;
        }
    }

    /**
     * One native call for the whole group. Every array is one pinned direct buffer, so a batch is limited to
     * 2^28 - 8 partitions (a direct ByteBuffer's capacity is an int); {@link #pinned} refuses larger ones loudly.
     */
    private void runNative(Flat f, boolean offsets) {
        if (ctx == 0) {
            ctx = devices.length == 1 ? NativeLagAssign.create(devices[0]) : NativeLagAssign.createMulti(devices);
        }
        final int T = f.topics.size();
        long r = 0;
        for (int[] s : f.sets) {
            r += s.length;
        }
        final ByteBuffer csetOff = pinned("csetOff", 8L * (f.sets.size() + 1));
        final ByteBuffer csetRank = pinned("csetRank", 4L * r);
        long c = 0;
        int w = 0;
        csetOff.putLong(0, 0L);
        for (int s = 0; s < f.sets.size(); s++) {
            for (int m : f.sets.get(s)) {
                csetRank.putInt(4 * w++, m);
            }
            c += f.sets.get(s).length;
            csetOff.putLong(8 * (s + 1), c);
        }
        final ByteBuffer topicOff = pinned("topicOff", 8L * (T + 1));
        final ByteBuffer topicCset = pinned("topicCset", 4L * T);
        final ByteBuffer part = f.explicitIds ? pinned("partition", 4L * f.n) : null;
        final ByteBuffer outPartition = pinned("outPartition", 4L * f.n);
        final ByteBuffer outConsumer = pinned("outConsumer", 4L * f.n);
        if (part != null) {
            int k = 0;
            for (int[] ids : f.ids) {
                for (int id : ids) {
                    part.putInt(4 * k++, id);
                }
            }
        }
        for (int t = 0; t <= T; t++) {
            topicOff.putLong(8 * t, f.topicOff[t]);
        }
        for (int t = 0; t < T; t++) {
            topicCset.putInt(4 * t, f.topicCset[t]);
        }
        if (offsets) {
            final boolean latest = consumerGroupProps
                .getProperty(ConsumerConfig.AUTO_OFFSET_RESET_CONFIG, "latest").equalsIgnoreCase("latest");   // Main:346-347, :391
            NativeLagAssign.assignBatchOffsets(ctx, T, topicOff, part, topicCset, f.sets.size(), csetOff, csetRank,
                pinned.get("begin"), pinned.get("end"), pinned.get("committed"), pinned.get("has"), latest,
                outPartition, outConsumer);
            f.partWidth = 4;
            f.consWidth = 4;
        } else {
            final int w = NativeLagAssign.assignBatchNarrow(ctx, T, topicOff, pinned.get("lag"), part, topicCset,
                f.sets.size(), csetOff, csetRank, outPartition, outConsumer);                 // replaces Main:177-184
            f.partWidth = w & 0xFF;
            f.consWidth = w >> 8;
        }
    }

    /** Topic order, then assignment order: the appends of Main:264. */
    private void unflatten(Flat f, Map<String, List<TopicPartition>> assignment) {
        final ByteBuffer outPartition = pinned.get("outPartition"), outConsumer = pinned.get("outConsumer");
        final List<List<TopicPartition>> byRankList = new ArrayList<>(f.byRank.length);
        for (String m : f.byRank) {
            byRankList.add(assignment.get(m));
        }
        long k = 0;
        for (int t = 0; t < f.topics.size(); t++) {
            final String topic = f.topics.get(t);
            final int[] set = f.sets.get(f.topicCset[t]);
            final long start = f.topicOff[t];
            for (final long end = f.topicOff[t + 1]; k < end; k++) {
                final int i = (int) k;
                final int m;
                switch (f.consWidth) {
                    case 0:  m = set.length == 0 ? -1 : set[(int) (k - start)]; break;     // P <= C: k-th member of the set
                    case 2:  { final int v = outConsumer.getShort(2 * i) & 0xFFFF; m = v == 0xFFFF ? -1 : v; break; }
                    default: m = outConsumer.getInt(4 * i);
                }
                if (m >= 0) {
                    final int id = f.partWidth == 1 ? outPartition.get(i) & 0xFF
                                 : f.partWidth == 2 ? outPartition.getShort(2 * i) & 0xFFFF : outPartition.getInt(4 * i);
                    byRankList.get(m).add(new TopicPartition(topic, id));
                }
            }
        }
    }

    /** Pinned (page-locked) direct buffer from lagassign_host_alloc, kept and reused; grown when too small. */
    private ByteBuffer pinned(String name, long bytes) {
        if (bytes > Integer.MAX_VALUE - 64) {
            throw new IllegalArgumentException("batch too large for one direct buffer: " + name + " needs " + bytes + " bytes");
        }
        ByteBuffer b = pinned.get(name);
        if (b == null || b.capacity() < bytes) {
            if (b != null) {
                NativeLagAssign.freePinned(b);
            }
            b = NativeLagAssign.allocPinned(Math.max(bytes + bytes / 4, 4096)).order(ByteOrder.nativeOrder());
            pinned.put(name, b);
        }
        return b;
    }
}
