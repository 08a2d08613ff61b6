package com.github.grantneale.kafka;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.Collections;
import java.util.HashMap;
import java.util.List;
import java.util.Map;
import org.apache.kafka.common.TopicPartition;

/**
 * Sketch of the Java host side (NOT COMPILED HERE: no JDK in the build image).
 * Same static seam as LagBasedPartitionAssignor.assign(Map, Map) (Main:166-188):
 * flatten -> NativeLagAssign.assignBatch -> unflatten. The C++ class
 * csrc/host/lag_based_partition_assignor.cpp does exactly this and IS tested.
 */
final class GpuLagBasedPartitionAssignor {

    private static long ctx = 0;

    static synchronized Map<String, List<TopicPartition>> assign(
        Map<String, List<LagBasedPartitionAssignor.TopicPartitionLag>> partitionLagPerTopic,
        Map<String, List<String>> subscriptions) {

        final Map<String, List<TopicPartition>> assignment = new HashMap<>();      // Main:171-174
        for (String memberId : subscriptions.keySet()) {
            assignment.put(memberId, new ArrayList<>());
        }
        // member ranks in String.compareTo order (Main:258-259)
        final String[] byRank = subscriptions.keySet().toArray(new String[0]);
        Arrays.sort(byRank);
        final Map<String, Integer> rank = new HashMap<>();
        for (int i = 0; i < byRank.length; i++) {
            rank.put(byRank[i], i);
        }
        // consumersPerTopic with a real HashMap: its iteration order is the topic order (Main:176-177)
        final Map<String, List<Integer>> consumersPerTopic = new HashMap<>();
        for (Map.Entry<String, List<String>> e : subscriptions.entrySet()) {
            for (String topic : e.getValue()) {
                consumersPerTopic.computeIfAbsent(topic, t -> new ArrayList<>()).add(rank.get(e.getKey()));
            }
        }
        final List<String> topics = new ArrayList<>(consumersPerTopic.keySet());
        final int T = topics.size();
        long n = 0;
        for (String t : topics) {
            n += partitionLagPerTopic.getOrDefault(t, Collections.emptyList()).size();   // Main:182
        }
        if (T == 0 || n == 0) {
            return assignment;
        }
        // flat buffers (de-duplicated consumer sets)
        final ByteBuffer topicOff = direct(8L * (T + 1));
        final ByteBuffer lag = direct(8L * n);
        final ByteBuffer part = direct(4L * n);
        final ByteBuffer topicCset = direct(4L * T);
        final Map<List<Integer>, Integer> setIds = new HashMap<>();
        final List<List<Integer>> sets = new ArrayList<>();
        long off = 0;
        topicOff.putLong(0);
        for (String t : topics) {
            for (LagBasedPartitionAssignor.TopicPartitionLag p
                    : partitionLagPerTopic.getOrDefault(t, Collections.emptyList())) {
                lag.putLong(p.getLag());
                part.putInt(p.getPartition());
                off++;
            }
            topicOff.putLong(off);
            final List<Integer> set = new ArrayList<>(new java.util.TreeSet<>(consumersPerTopic.get(t)));
            Integer id = setIds.get(set);
            if (id == null) {
                id = sets.size();
                setIds.put(set, id);
                sets.add(set);
            }
            topicCset.putInt(id);
        }
        long r = 0;
        for (List<Integer> s : sets) {
            r += s.size();
        }
        final ByteBuffer csetOff = direct(8L * (sets.size() + 1));
        final ByteBuffer csetRank = direct(4L * r);
        long c = 0;
        csetOff.putLong(0);
        for (List<Integer> s : sets) {
            for (int m : s) {
                csetRank.putInt(m);
            }
            c += s.size();
            csetOff.putLong(c);
        }
        final ByteBuffer outPartition = direct(4L * n);
        final ByteBuffer outConsumer = direct(4L * n);
        if (ctx == 0) {
            ctx = NativeLagAssign.create(0);
        }
        NativeLagAssign.assignBatch(ctx, T, topicOff, lag, part, topicCset, sets.size(), csetOff, csetRank,
                                    outPartition, outConsumer);                       // replaces Main:178-183
        // unflatten: topic order, then assignment order (Main:264)
        int k = 0;
        for (int ti = 0; ti < T; ti++) {
            final long end = topicOff.getLong(8 * (ti + 1));
            for (; k < end; k++) {
                final int m = outConsumer.getInt(4 * k);
                if (m >= 0) {
                    assignment.get(byRank[m]).add(new TopicPartition(topics.get(ti), outPartition.getInt(4 * k)));
                }
            }
        }
        return assignment;
    }

    private static ByteBuffer direct(long bytes) {
        return ByteBuffer.allocateDirect((int) Math.max(bytes, 1)).order(ByteOrder.nativeOrder());
    }
}
