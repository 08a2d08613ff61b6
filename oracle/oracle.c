/*
 * oracle.c — CPU restatement of the reference algorithm (see oracle.h: test
 * infrastructure only, never on the product path).
 *
 * "Main:n" cites /root/reference/src/main/java/com/github/grantneale/kafka/
 * LagBasedPartitionAssignor.java line n.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>
#include <strings.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- sort of one topic's partitions: Main:228-235 ------------------------------- */

typedef struct {
    int64_t lag;
    int32_t part;
} plag_t;

/* The comparator lambda of Main:228-235. */
static inline int cmp_plag(const plag_t *p1, const plag_t *p2) {
    if (p1->lag == p2->lag) {                       /* Main:230 */
        return (p1->part > p2->part) - (p1->part < p2->part); /* Integer.compare, Main:231 */
    }
    return (p2->lag > p1->lag) - (p2->lag < p1->lag);          /* Long.compare(p2, p1), Main:234 */
}

/* List.sort is a stable merge sort (TimSort); any stable sort gives the same
 * sequence. Top-down merge sort with one scratch buffer. */
static void stable_sort_plag(plag_t *a, plag_t *tmp, int64_t n) {
    if (n < 2) return;
    if (n <= 16) { /* insertion sort, stable */
        for (int64_t i = 1; i < n; i++) {
            plag_t x = a[i];
            int64_t j = i;
            while (j > 0 && cmp_plag(&a[j - 1], &x) > 0) { a[j] = a[j - 1]; j--; }
            a[j] = x;
        }
        return;
    }
    int64_t h = n / 2;
    stable_sort_plag(a, tmp, h);
    stable_sort_plag(a + h, tmp, n - h);
    if (cmp_plag(&a[h - 1], &a[h]) <= 0) return;
    memcpy(tmp, a, (size_t)h * sizeof(plag_t));
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) {
        if (cmp_plag(&a[j], &tmp[i]) < 0) a[k++] = a[j++];   /* strict: left wins ties */
        else a[k++] = tmp[i++];
    }
    while (i < h) a[k++] = tmp[i++];
}

/* ---- consumer order for the round formulation ----------------------------------- */

typedef struct {
    int64_t total;
    int32_t c;     /* index into the topic's consumer list == member-id order */
} ctot_t;

static int cmp_ctot(const void *x, const void *y) {
    const ctot_t *a = (const ctot_t *)x, *b = (const ctot_t *)y;
    if (a->total != b->total) return (a->total > b->total) - (a->total < b->total);
    return (a->c > b->c) - (a->c < b->c);
}

static inline int64_t wrap_add(int64_t a, int64_t b) {  /* Java long '+' wraps, Main:265 */
    return (int64_t)((uint64_t)a + (uint64_t)b);
}

int oracle_assign_topic(int mode, int64_t P, const int64_t *lag, const int32_t *part,
                        int64_t C, const int32_t *crank,
                        int32_t *out_part, int32_t *out_cons,
                        int64_t *out_total, int32_t *out_count) {
    if (C <= 0) {
        /* Main:211-213 / Main:177: a topic without consumers is not assigned
         * (and its list is not sorted). */
        for (int64_t i = 0; i < P; i++) {
            out_part[i] = part ? part[i] : (int32_t)i;
            out_cons[i] = -1;
        }
        return 0;
    }
    /* Main:215-225: fresh totals and counts per topic. */
    int64_t *total = (int64_t *)calloc((size_t)C, sizeof(int64_t));
    int32_t *count = (int32_t *)calloc((size_t)C, sizeof(int32_t));
    plag_t *pl = (plag_t *)malloc((size_t)(P > 0 ? P : 1) * sizeof(plag_t));
    plag_t *tmp = (plag_t *)malloc((size_t)(P > 0 ? P : 1) * sizeof(plag_t));
    if (!total || !count || !pl || !tmp) { free(total); free(count); free(pl); free(tmp); return -1; }

    for (int64_t i = 0; i < P; i++) {
        pl[i].lag = lag[i];
        pl[i].part = part ? part[i] : (int32_t)i;
    }
    stable_sort_plag(pl, tmp, P);                    /* Main:228-235 */

    if (mode == ORACLE_LITERAL) {
        for (int64_t k = 0; k < P; k++) {            /* Main:237 */
            /* Collections.min with the 3-level comparator, Main:240-263.
             * Consumers are listed in member-id order, so "smaller id" is
             * "smaller index"; a strict '<' scan keeps the first minimum. */
            int64_t best = 0;
            for (int64_t c = 1; c < C; c++) {
                int r;
                if (count[c] != count[best]) r = (count[c] > count[best]) - (count[c] < count[best]); /* :246-250 */
                else if (total[c] != total[best]) r = (total[c] > total[best]) - (total[c] < total[best]); /* :253-256 */
                else r = (crank[c] > crank[best]) - (crank[c] < crank[best]); /* :259 */
                if (r < 0) best = c;
            }
            out_part[k] = pl[k].part;                /* Main:264 */
            out_cons[k] = crank[best];
            total[best] = wrap_add(total[best], pl[k].lag);  /* Main:265 */
            count[best] += 1;                        /* Main:266 */
        }
    } else {
        /* Round formulation: the count criterion
         * dominates, so partitions are consumed in rounds of C; inside a round
         * consumers are taken in ascending (total at round start, member id). */
        ctot_t *ord = (ctot_t *)malloc((size_t)C * sizeof(ctot_t));
        if (!ord) { free(total); free(count); free(pl); free(tmp); return -1; }
        for (int64_t base = 0; base < P; base += C) {
            for (int64_t c = 0; c < C; c++) { ord[c].total = total[c]; ord[c].c = (int32_t)c; }
            qsort(ord, (size_t)C, sizeof(ctot_t), cmp_ctot);
            int64_t m = P - base < C ? P - base : C;
            for (int64_t j = 0; j < m; j++) {
                int64_t c = ord[j].c;
                out_part[base + j] = pl[base + j].part;
                out_cons[base + j] = crank[c];
                total[c] = wrap_add(total[c], pl[base + j].lag);
                count[c] += 1;
            }
        }
        free(ord);
    }
    if (out_total) memcpy(out_total, total, (size_t)C * sizeof(int64_t));
    if (out_count) memcpy(out_count, count, (size_t)C * sizeof(int32_t));
    free(total); free(count); free(pl); free(tmp);
    return 0;
}

int64_t oracle_assign_batch(int mode, int threads,
                            int64_t n_topics, const int64_t *topic_off,
                            const int64_t *lag, const int32_t *partition,
                            const int32_t *topic_cset, int64_t n_csets,
                            const int64_t *cset_off, const int32_t *cset_rank,
                            int64_t topic_lo, int64_t topic_hi,
                            int32_t *out_partition, int32_t *out_consumer_rank,
                            int64_t *out_total, int32_t *out_count) {
    if (topic_lo < 0) topic_lo = 0;
    if (topic_hi > n_topics) topic_hi = n_topics;
    (void)n_csets;
    /* offsets of the per-topic totals/counts blocks */
    int64_t *tc_off = NULL;
    if (out_total || out_count) {
        tc_off = (int64_t *)malloc((size_t)(n_topics + 1) * sizeof(int64_t));
        if (!tc_off) return -1;
        tc_off[0] = 0;
        for (int64_t t = 0; t < n_topics; t++) {
            int32_t s = topic_cset[t];
            tc_off[t + 1] = tc_off[t] + (cset_off[s + 1] - cset_off[s]);
        }
    }
    int64_t assigned = 0;
    int failed = 0;
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads) reduction(+ : assigned) reduction(| : failed)
#endif
    for (int64_t t = topic_lo; t < topic_hi; t++) {      /* Main:177-184 */
        int64_t off = topic_off[t], P = topic_off[t + 1] - off;
        int32_t s = topic_cset[t];
        int64_t coff = cset_off[s], C = cset_off[s + 1] - coff;
        int rc = oracle_assign_topic(mode, P, lag + off, partition ? partition + off : NULL,
                                     C, cset_rank + coff,
                                     out_partition + off, out_consumer_rank + off,
                                     out_total ? out_total + tc_off[t] : NULL,
                                     out_count ? out_count + tc_off[t] : NULL);
        if (rc != 0) failed |= 1;
        if (C > 0) assigned += P;
    }
    free(tc_off);
    return failed ? -1 : assigned;
}

/* ---- Main:376-404 ---------------------------------------------------------------- */

int64_t oracle_compute_partition_lag(int has_committed, int64_t committed,
                                     int64_t begin_offset, int64_t end_offset,
                                     const char *reset_mode) {
    int64_t next;
    if (has_committed) {
        next = committed;                                    /* Main:386 */
    } else if (reset_mode && strcasecmp(reset_mode, "latest") == 0) {
        next = end_offset;                                   /* Main:391-392 */
    } else {
        next = begin_offset;                                 /* Main:395 */
    }
    int64_t d = (int64_t)((uint64_t)end_offset - (uint64_t)next);  /* Java long '-' wraps */
    return d > 0 ? d : 0;                                    /* Long.max(.., 0L), Main:402 */
}

/* ---- String.compareTo over UTF-16 code units (Main:258-259) ---------------------- */

/* Decode one UTF-8 sequence (lenient: malformed bytes decode as themselves). */
static const unsigned char *utf8_next(const unsigned char *s, uint32_t *cp) {
    unsigned char c = s[0];
    if (c < 0x80) { *cp = c; return s + 1; }
    if ((c & 0xE0) == 0xC0 && (s[1] & 0xC0) == 0x80) {
        *cp = ((uint32_t)(c & 0x1F) << 6) | (s[1] & 0x3F); return s + 2;
    }
    if ((c & 0xF0) == 0xE0 && (s[1] & 0xC0) == 0x80 && (s[2] & 0xC0) == 0x80) {
        *cp = ((uint32_t)(c & 0x0F) << 12) | ((uint32_t)(s[1] & 0x3F) << 6) | (s[2] & 0x3F); return s + 3;
    }
    if ((c & 0xF8) == 0xF0 && (s[1] & 0xC0) == 0x80 && (s[2] & 0xC0) == 0x80 && (s[3] & 0xC0) == 0x80) {
        *cp = ((uint32_t)(c & 0x07) << 18) | ((uint32_t)(s[1] & 0x3F) << 12) |
              ((uint32_t)(s[2] & 0x3F) << 6) | (s[3] & 0x3F);
        return s + 4;
    }
    *cp = c; return s + 1;
}

/* Pull the next UTF-16 code unit; *pending holds a low surrogate to emit next. */
static int next_unit(const unsigned char **s, uint32_t *pending, uint32_t *unit) {
    if (*pending) { *unit = *pending; *pending = 0; return 1; }
    if (**s == 0) return 0;
    uint32_t cp;
    *s = utf8_next(*s, &cp);
    if (cp >= 0x10000) {
        cp -= 0x10000;
        *unit = 0xD800 + (cp >> 10);
        *pending = 0xDC00 + (cp & 0x3FF);
    } else {
        *unit = cp;
    }
    return 1;
}

int oracle_compare_member_ids(const char *a_utf8, const char *b_utf8) {
    const unsigned char *a = (const unsigned char *)a_utf8, *b = (const unsigned char *)b_utf8;
    uint32_t pa = 0, pb = 0;
    for (;;) {
        uint32_t ua = 0, ub = 0;
        int ha = next_unit(&a, &pa, &ua), hb = next_unit(&b, &pb, &ub);
        if (!ha || !hb) return ha - hb;          /* shorter string first */
        if (ua != ub) return ua < ub ? -1 : 1;   /* code-unit difference */
    }
}

typedef struct { const char *id; int64_t i; } idref_t;
static int cmp_idref(const void *x, const void *y) {
    return oracle_compare_member_ids(((const idref_t *)x)->id, ((const idref_t *)y)->id);
}

int oracle_rank_members(int64_t n, const char *const *ids_utf8, int32_t *rank_out) {
    idref_t *r = (idref_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(idref_t));
    if (!r) return -1;
    for (int64_t i = 0; i < n; i++) { r[i].id = ids_utf8[i]; r[i].i = i; }
    qsort(r, (size_t)n, sizeof(idref_t), cmp_idref);
    for (int64_t k = 0; k < n; k++) rank_out[r[k].i] = (int32_t)k;
    free(r);
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
