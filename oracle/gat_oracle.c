/*
 * gat_oracle.c -- CPU restatement of the GAT simulation hot path.  TEST INFRASTRUCTURE ONLY
 * (see gat_oracle.h for the rules and the parity status: PINNED against the compiled reference).
 *
 * Written from the behaviour of AndreasHeger/gat 1.3.6; each function cites the file:line it restates
 * (paths relative to /root/reference).  Plain C99, no dependencies.
 */
#include "gat_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* signed helpers: the reference compares through PositionDifference = int32
 * (gat/SegmentList.pyx:64-75, gat/Engine.pyx:56-68) */
static inline int32_t i32min(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t i32max(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t seg_len(go_seg s) { return (int32_t)s.end - (int32_t)s.start; }

/* ------------------------------------------------------------------------------------------------
 * utils/gat_utils.c:8-32 searchsorted == lower_bound under the comparator.
 * cmpPosition (gat/SegmentList.pyx:131-132) returns (int)(a - b) on uint32; cmpSegments (:119-121)
 * compares starts as int32. */
long go_searchsorted_u32(const uint32_t *base, size_t n, uint32_t target)
{
    size_t imin = 0, imax = n;
    while (imin < imax) {
        size_t imid = imin + ((imax - imin) >> 1);
        if ((int)(base[imid] - target) < 0) imin = imid + 1; else imax = imid;
    }
    return (long)imin;
}

long go_searchsorted_seg(const go_seg *base, size_t n, go_seg target)
{
    size_t imin = 0, imax = n;
    while (imin < imax) {
        size_t imid = imin + ((imax - imin) >> 1);
        if (((int32_t)base[imid].start - (int32_t)target.start) < 0) imin = imid + 1; else imax = imid;
    }
    return (long)imin;
}

/* utils/gat_utils.c:37-61 with Engine.pyx's own cmpDouble (gat/Engine.pyx:122-127, a proper three-way
 * compare -- NOT the truncating cmpDouble of gat/SegmentList.pyx:135-136, which this path never uses) */
long go_searchargsorted_f64(const double *base, const int *sorted, size_t n, double target)
{
    size_t imin = 0, imax = n;
    while (imin < imax) {
        size_t imid = imin + ((imax - imin) >> 1);
        double v = sorted ? base[sorted[imid]] : base[imid];
        if (((v > target) - (v < target)) < 0) imin = imid + 1; else imax = imid;
    }
    return (long)imin;
}

/* ------------------------------------------------------------------------------------------------
 * gat/SegmentList.pyx:478-486 sort(): qsort by start only (ties unordered; no result depends on it) */
static int cmp_seg_start(const void *a, const void *b)
{
    return (int32_t)((const go_seg *)a)->start - (int32_t)((const go_seg *)b)->start;
}
void go_sort(go_seg *s, size_t n) { if (n) qsort(s, n, sizeof(go_seg), cmp_seg_start); }

/* gat/SegmentList.pyx:697-754 normalize(): merge overlapping, drop empty, keep adjacent apart */
size_t go_normalize(go_seg *s, size_t n)
{
    if (n == 0) return 0;
    go_sort(s, n);
    size_t ins = 0, idx = 0;
    while (idx < n && s[idx].start == s[idx].end) idx++;
    if (idx == n) return 0;
    s[ins].start = s[idx].start;
    uint32_t max_end = s[idx].end;
    while (idx < n) {
        if (s[idx].start == s[idx].end) { idx++; continue; }
        if (s[idx].start >= max_end) {
            s[ins].end = max_end;
            ins++;
            s[ins].start = s[idx].start;
        }
        max_end = (uint32_t)i32max((int32_t)s[idx].end, (int32_t)max_end);
        idx++;
    }
    s[ins].end = max_end;
    return ins + 1;
}

/* gat/SegmentList.pyx:756-816 merge(distance): as normalize but joins when start - distance <= max_end */
size_t go_merge(go_seg *s, size_t n, int32_t distance)
{
    if (n == 0) return 0;
    go_sort(s, n);
    size_t ins = 0, idx = 0;
    while (idx < n && s[idx].start == s[idx].end) idx++;
    if (idx == n) return 0;
    s[ins].start = s[idx].start;
    int32_t max_end = (int32_t)s[idx].end;
    while (idx < n) {
        if (s[idx].start == s[idx].end) { idx++; continue; }
        if ((int32_t)s[idx].start - distance > max_end) {
            s[ins].end = (uint32_t)max_end;
            ins++;
            s[ins].start = s[idx].start;
        }
        max_end = i32max((int32_t)s[idx].end, max_end);
        idx++;
    }
    s[ins].end = (uint32_t)max_end;
    return ins + 1;
}

/* gat/SegmentList.pyx:1401-1467 filter(other): keep self-segments overlapping other, untruncated */
size_t go_filter(const go_seg *self, size_t n, const go_seg *other, size_t m, go_seg *out)
{
    if (n == 0) return 0;
    size_t w = 0, ti = 0, oi = 0;
    uint32_t last_start = self[0].start - 1;
    while (ti < n && oi < m) {
        go_seg t = self[ti], o = other[oi];
        if (t.end <= o.start) ti++;
        else if (o.end <= t.start) oi++;
        else {
            if (last_start != t.start) { out[w++] = t; last_start = t.start; }
            if (t.end < o.end) ti++;
            else if (o.end < t.end) oi++;
            else { ti++; oi++; }
        }
    }
    return w;
}

/* gat/SegmentList.pyx:1469-1549 intersect(other): pairwise intersections, pieces not re-merged */
size_t go_intersect(const go_seg *self, size_t n, const go_seg *other, size_t m, go_seg *out)
{
    size_t w = 0, ti = 0, oi = 0;
    while (ti < n && oi < m) {
        go_seg t = self[ti], o = other[oi];
        if (t.end <= o.start) ti++;
        else if (o.end <= t.start) oi++;
        else {
            out[w].start = (uint32_t)i32max((int32_t)t.start, (int32_t)o.start);
            out[w].end = (uint32_t)i32min((int32_t)t.end, (int32_t)o.end);
            w++;
            if (t.end < o.end) ti++;
            else if (o.end < t.end) oi++;
            else { ti++; oi++; }
        }
    }
    return w;
}

/* gat/SegmentList.pyx:1607-1616 sum(): uint32 accumulation */
uint32_t go_sum(const go_seg *s, size_t n)
{
    uint32_t total = 0;
    for (size_t i = 0; i < n; i++) total += s[i].end - s[i].start;
    return total;
}

/* gat/SegmentList.pyx:1026-1076 overlapWithSegments(): shared bases of two normalized lists.
 * (the self-self shortcut :1036-1037 returns sum(), which equals the merge result) */
uint32_t go_overlap_with_segments(const go_seg *self, size_t n, const go_seg *other, size_t m)
{
    size_t ti = 0, oi = 0;
    uint32_t overlap = 0;
    while (ti < n && oi < m) {
        go_seg t = self[ti], o = other[oi];
        if (t.end <= o.start) ti++;
        else if (o.end <= t.start) oi++;
        else {
            overlap += (uint32_t)(i32min((int32_t)t.end, (int32_t)o.end) - i32max((int32_t)t.start, (int32_t)o.start));
            if (t.end < o.end) ti++;
            else if (o.end < t.end) oi++;
            else { ti++; oi++; }
        }
    }
    return overlap;
}

/* gat/SegmentList.pyx:1078-1146 intersectionWithSegments(mode): number of self-segments overlapping
 * other; midpoint mode tests only the FIRST overlapping other-segment (this_idx advances either way) */
uint32_t go_intersection_with_segments(const go_seg *self, size_t n, const go_seg *other, size_t m,
                                       int midpoint)
{
    size_t ti = 0, oi = 0;
    uint32_t noverlap = 0;
    while (ti < n && oi < m) {
        go_seg t = self[ti], o = other[oi];
        if (t.end <= o.start) ti++;
        else if (o.end <= t.start) oi++;
        else {
            if (midpoint) {
                uint32_t mid = t.start + (t.end - t.start) / 2;
                if (o.start <= mid && mid < o.end) noverlap++;
            } else noverlap++;
            ti++;
        }
    }
    return noverlap;
}

/* gat/SegmentList.pyx:853-887 _getInsertionPoint() */
int go_get_insertion_point(const go_seg *s, size_t n, go_seg other)
{
    if (n == 0) return -1;
    if (other.start >= s[n - 1].end) return (int)n;
    if (other.end <= s[0].start) return -1;
    long idx = go_searchsorted_seg(s, n, other);
    if (idx == (long)n) return (int)idx - 1;
    else if (s[idx].start != other.start) return (int)idx - 1;
    return (int)idx;
}

/* gat/SegmentList.pyx:545-597 trim_ends(pos, size, forward).  Returns 0, or -1 where the reference
 * asserts (sum() <= size). */
int go_trim_ends(go_seg *s, size_t n, uint32_t pos, uint32_t size, int forward)
{
    if (n == 0) return 0;
    int32_t sz = (int32_t)size;
    if (!((int64_t)go_sum(s, n) > (int64_t)sz)) return -1;
    go_seg other = { pos, pos + 1 };
    int idx = go_get_insertion_point(s, n, other);
    if (idx == (int)n) idx = 0;
    if (idx < 0) idx = (int)n - 1;
    if (forward) {
        while (sz > 0) {
            go_seg seg = s[idx];
            uint32_t l = (uint32_t)seg_len(seg);
            if (seg_len(seg) < sz) { s[idx].start = 0; s[idx].end = 0; sz -= (int32_t)l; }
            else { s[idx].start = seg.start + (uint32_t)sz; s[idx].end = seg.end; sz = 0; }
            idx += 1;
            if (idx == (int)n) idx = 0;
        }
    } else {
        while (sz > 0) {
            go_seg seg = s[idx];
            uint32_t l = (uint32_t)seg_len(seg);
            if (seg_len(seg) < sz) { s[idx].start = 0; s[idx].end = 0; sz -= (int32_t)l; }
            else { s[idx].start = seg.start; s[idx].end = (uint32_t)((int32_t)seg.end - sz); sz = 0; }
            idx -= 1;
            if (idx < 0) idx = (int)n - 1;
        }
    }
    return 0;
}

/* gat/SegmentList.pyx:1148-1184 getLengthDistribution(): histogram[ceil(len/bucket)]++ ;
 * bucket_size 0 -> ceil(largest/nbuckets).  Index >= nbuckets raises ValueError (-> return 0). */
uint32_t go_length_distribution(const go_seg *s, size_t n, uint32_t bucket_size, uint32_t nbuckets,
                                int64_t *histogram)
{
    memset(histogram, 0, sizeof(int64_t) * nbuckets);
    if (bucket_size == 0) {
        int32_t largest = 0;
        for (size_t i = 0; i < n; i++) if (seg_len(s[i]) > largest) largest = seg_len(s[i]);
        bucket_size = (uint32_t)ceil((double)largest / (double)nbuckets);
    }
    for (size_t i = 0; i < n; i++) {
        uint32_t l = (uint32_t)seg_len(s[i]);
        /* Python-object arithmetic in the reference: true division, then <int> truncation */
        int idx = (int)(((double)l + (double)bucket_size - 1.0) / (double)bucket_size);
        if (idx >= (int)nbuckets) return 0;
        histogram[idx] += 1;
    }
    return bucket_size;
}

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al. 2011), restated from the published round function:
 * multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85.
 * Counter/key layout and the bounded-integer map are the RNG contract in DESIGN.md; the CUDA
 * kernel implements the same bits. */
void go_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void go_philox_begin(go_philox_ctx *c, uint64_t seed, uint32_t track, uint32_t unit, uint32_t sample)
{
    c->seed = seed; c->track = track; c->unit = unit; c->sample = sample;
    c->turn = 0xFFFFFFFFu;   /* next_turn() makes the first turn 0 */
}

void go_philox_next_turn(void *ctx) { ((go_philox_ctx *)ctx)->turn += 1; }

/* draw `slot` of the current turn: counter = (turn, block | track<<8, unit, sample), key = seed.
 *   block 0: LEN = words 0,1   WS_R   = words 2,3
 *   block 1: WS_P = words 0,1  JITTER = words 2,3
 *   block 2: TRIM_R = words 0,1  TRIM_P = words 2,3
 *   block 3: TRIM_DIR = words 0,1
 * r64 = hi:lo of the two words (first word low); value = lo + floor(r64 * range / 2^64), range < 2^32 */
int64_t go_philox_randint(void *ctx, int slot, int64_t lo, int64_t hi)
{
    const go_philox_ctx *c = (const go_philox_ctx *)ctx;
    static const int block_of[7] = { 0, 1, 0, 1, 2, 2, 3 };
    static const int word_of[7] = { 0, 2, 2, 0, 0, 2, 0 };
    uint32_t ctr[4] = { c->turn, (uint32_t)block_of[slot] | (c->track << 8), c->unit, c->sample };
    uint32_t key[2] = { (uint32_t)c->seed, (uint32_t)(c->seed >> 32) };
    uint32_t out[4];
    go_philox4x32_10(ctr, key, out);
    uint32_t rlo = out[word_of[slot]], rhi = out[word_of[slot] + 1];
    uint32_t range = (uint32_t)(hi - lo);
    uint64_t v = ((uint64_t)rhi * range + (((uint64_t)rlo * range) >> 32)) >> 32;
    return lo + (int64_t)v;
}

/* ------------------------------------------------------------------------------------------------
 * gat/Engine.pyx:279-348 SegmentListSampler.sample(sample_length).
 * cdf[i] = cumulative length - 1 (:274-277); total = sum of lengths. */
int go_segmentlist_sample(const go_seg *ws, size_t m, const uint32_t *cdf, uint32_t total,
                          uint32_t sample_length, go_randint_fn rnd, void *ctx, int slot_r, int slot_p,
                          uint32_t *start, uint32_t *end, int32_t *overlap)
{
    uint32_t r = (uint32_t)rnd(ctx, slot_r, 0, (int64_t)total);
    long idx = go_searchsorted_u32(cdf, m, r);
    if (idx < 0 || idx >= (long)m) return -1;
    go_seg chosen = ws[idx];
    /* long arithmetic, then squeezed through lmax(PositionDifference, PositionDifference) (:318-325) */
    int64_t sampling_start = (int64_t)chosen.start - (int64_t)sample_length + 1;
    if (idx > 0) sampling_start = (int64_t)i32max((int32_t)ws[idx - 1].end, (int32_t)sampling_start);
    int64_t p = rnd(ctx, slot_p, sampling_start, (int64_t)chosen.end);
    *start = (uint32_t)i32max(0, (int32_t)p);
    *end = (uint32_t)(p + (int64_t)sample_length);
    *overlap = i32max(0, i32min((int32_t)chosen.end, (int32_t)*end) - i32max((int32_t)chosen.start, (int32_t)*start));
    return 0;
}

static uint32_t *build_ws_cdf(const go_seg *ws, size_t m, uint32_t *total)
{
    uint32_t *cdf = (uint32_t *)malloc(sizeof(uint32_t) * (m ? m : 1));
    if (!cdf) return NULL;
    uint32_t t = 0;
    for (size_t i = 0; i < m; i++) { t += (uint32_t)seg_len(ws[i]); cdf[i] = t - 1; }
    *total = t;
    return cdf;
}

/* gat/Engine.pyx:515-646 SamplerAnnotator.sample(segments, workspace) */
long go_sampler_annotator(const go_seg *segments, size_t n, const go_seg *workspace, size_t m,
                          uint32_t bucket_size, uint32_t nbuckets,
                          go_randint_fn rnd, go_turn_fn next_turn, void *ctx,
                          go_seg *out, size_t cap, go_sample_info *info)
{
    go_sample_info local;
    if (!info) info = &local;
    memset(info, 0, sizeof(*info));
    long result = -3;
    go_seg *working = NULL, *tmp = NULL, *U = NULL, *pending = NULL, *isect = NULL;
    int64_t *histogram = NULL;
    uint32_t *hcdf = NULL, *wcdf = NULL, *ucdf = NULL;

    /* :543-546 working = segments.filter(workspace) */
    working = (go_seg *)malloc(sizeof(go_seg) * (n ? n : 1));
    if (!working) goto done;
    size_t nw = go_filter(segments, n, workspace, m, working);
    if (nw == 0) { result = 0; goto done; }

    /* :550-552 ltotal = working.intersect(workspace).sum() */
    tmp = (go_seg *)malloc(sizeof(go_seg) * (nw + m));
    if (!tmp) goto done;
    size_t nt = go_intersect(working, nw, workspace, m, tmp);
    int32_t ltotal = (int32_t)go_sum(tmp, nt);
    info->ltotal = ltotal;

    /* :559-562 length histogram + HistogramSampler cdf (:395-411) */
    histogram = (int64_t *)malloc(sizeof(int64_t) * nbuckets);
    hcdf = (uint32_t *)malloc(sizeof(uint32_t) * nbuckets);
    if (!histogram || !hcdf) goto done;
    uint32_t bucket = go_length_distribution(working, nw, bucket_size, nbuckets, histogram);
    if (bucket == 0) { result = -2; goto done; }
    info->bucket_size = bucket;
    uint32_t htotal = 0;
    for (uint32_t i = 0; i < nbuckets; i++) { htotal += (uint32_t)histogram[i]; hcdf[i] = htotal; }

    /* :565 SegmentListSampler(workspace) */
    uint32_t wtotal = 0;
    wcdf = build_ws_cdf(workspace, m, &wtotal);
    if (!wcdf) goto done;

    /* internal lists grow on demand (the reference reallocs); only `out` is bounded by cap */
    size_t ucap = 2 * nw + 64, pcap = 2 * nw + 64;
    U = (go_seg *)malloc(sizeof(go_seg) * ucap);
    pending = (go_seg *)malloc(sizeof(go_seg) * pcap);
    isect = (go_seg *)malloc(sizeof(go_seg) * (ucap + m + 1));
    if (!U || !pending || !isect) goto done;
    size_t nu = 0, np = 0;

    int32_t remaining = ltotal, true_remaining = ltotal;
    int nunsuccessful = 0;
    const int max_unsuccessful = 20;

    while (true_remaining > 0 && nunsuccessful < max_unsuccessful) {
        if (next_turn) next_turn(ctx);
        info->nturns++;

        /* :576 length = hs.sample()  (:413-435) */
        uint32_t r = 1;
        if (htotal > 1) r = (uint32_t)rnd(ctx, GO_SLOT_LEN, 1, (int64_t)htotal);
        long hidx = go_searchsorted_u32(hcdf, nbuckets, r);
        uint32_t base = (uint32_t)hidx * bucket;
        if (bucket > 1) base += (uint32_t)rnd(ctx, GO_SLOT_JITTER, 0, (int64_t)bucket);
        int32_t length = (int32_t)base;

        /* :582-605 checkpoint */
        if (remaining <= length) {
            if (nu + np > ucap) {
                ucap = 2 * (nu + np);
                go_seg *nU = (go_seg *)realloc(U, sizeof(go_seg) * ucap);
                go_seg *nI = (go_seg *)realloc(isect, sizeof(go_seg) * (ucap + m + 1));
                if (nU) U = nU;
                if (nI) isect = nI;
                if (!nU || !nI) goto done;
            }
            memcpy(U + nu, pending, sizeof(go_seg) * np);
            nu = go_merge(U, nu + np, 0);
            np = 0;
            size_t ni = go_intersect(U, nu, workspace, m, isect);
            remaining = ltotal - (int32_t)go_sum(isect, ni);
            if (true_remaining == remaining) nunsuccessful++;
            else true_remaining = remaining;
            info->ncheckpoints++;
        }

        /* :608-625 overshoot */
        if (true_remaining < 0) {
            uint32_t utotal = 0, s0, e0;
            int32_t ov0;
            free(ucdf);
            ucdf = build_ws_cdf(U, nu, &utotal);
            if (!ucdf) goto done;
            if (go_segmentlist_sample(U, nu, ucdf, utotal, 1, rnd, ctx, GO_SLOT_TRIM_R, GO_SLOT_TRIM_P,
                                      &s0, &e0, &ov0) != 0) { result = -4; goto done; }
            int forward = (int)rnd(ctx, GO_SLOT_TRIM_DIR, 0, 2);
            if (go_trim_ends(U, nu, s0, (uint32_t)(-true_remaining), forward) != 0) { result = -5; goto done; }
            true_remaining = 1;
            info->ntrims++;
            continue;
        }

        /* :628 start, end, overlap = sls.sample(length) -- drawn even if it will not be kept */
        uint32_t start, end;
        int32_t overlap;
        if (go_segmentlist_sample(workspace, m, wcdf, wtotal, (uint32_t)length, rnd, ctx,
                                  GO_SLOT_WS_R, GO_SLOT_WS_P, &start, &end, &overlap) != 0) {
            result = -4; goto done;
        }
        /* :632-634 */
        if (true_remaining > 0) {
            if (np >= pcap) {
                pcap *= 2;
                go_seg *nP = (go_seg *)realloc(pending, sizeof(go_seg) * pcap);
                if (!nP) goto done;
                pending = nP;
            }
            pending[np].start = start; pending[np].end = end; np++;
            remaining -= overlap;
            info->nplaced++;
        }
    }
    info->true_remaining = true_remaining;
    info->nunsuccessful = nunsuccessful;

    /* :639-646 result = unintersected.merge(0).filter(workspace); pending placements are dropped */
    nu = go_merge(U, nu, 0);
    {
        size_t nr = go_filter(U, nu, workspace, m, isect);
        if (nr > cap) { result = -1; goto done; }
        memcpy(out, isect, sizeof(go_seg) * nr);
        result = (long)nr;
    }
done:
    free(working); free(tmp); free(U); free(pending); free(isect);
    free(histogram); free(hcdf); free(wcdf); free(ucdf);
    return result;
}

/* ------------------------------------------------------------------------------------------------
 * gat/Engine.pyx:1412-1472 Counter*.__call__ */
double go_counter(int counter, const go_seg *segments, size_t n, const go_seg *annotations, size_t m,
                  size_t workspace_nsegments)
{
    switch (counter) {
    case GO_NUCLEOTIDE_OVERLAP:     /* :1417-1425 annotations.overlapWithSegments(segments) */
        return (double)go_overlap_with_segments(annotations, m, segments, n);
    case GO_NUCLEOTIDE_DENSITY:     /* :1427-1441 divides by len(workspace) = number of segments */
        if (workspace_nsegments == 0) return 0.0;
        return (double)go_overlap_with_segments(annotations, m, segments, n) / (double)(uint32_t)workspace_nsegments;
    case GO_SEGMENT_OVERLAP:        /* :1443-1448 */
        return (double)go_intersection_with_segments(segments, n, annotations, m, 0);
    case GO_SEGMENT_MIDOVERLAP:     /* :1450-1456 */
        return (double)go_intersection_with_segments(segments, n, annotations, m, 1);
    case GO_ANNOTATION_OVERLAP:     /* :1458-1463 */
        return (double)go_intersection_with_segments(annotations, m, segments, n, 0);
    case GO_ANNOTATION_MIDOVERLAP:  /* :1465-1472 */
        return (double)go_intersection_with_segments(annotations, m, segments, n, 1);
    }
    return NAN;
}

/* gat/__init__.py:580-587: counts[counter][annotation] = sum(counter(sample[contig], annos[contig],
 * ws[contig]) for contig in sample.keys()) -- Python sum: left to right from int 0 */
void go_count_placed(int C, int A, const uint64_t *placed_off, const go_seg *placed,
                     const uint64_t *anno_off, const go_seg *anno, const uint32_t *cws_nseg,
                     int ncounters, const int32_t *counters, double *counts)
{
    for (int k = 0; k < ncounters; k++)
        for (int a = 0; a < A; a++) {
            /* Python sum(): ints add plainly, floats with Neumaier compensation (CPython >= 3.12,
             * Python/bltinmodule.c builtin_sum; the reference in oracle/_ref runs on such a Python) */
            double total = 0.0, comp = 0.0;
            for (int c = 0; c < C; c++) {
                /* contigs absent from the sample contribute nothing; present-but-empty contigs add 0 */
                const go_seg *s = placed + placed_off[c];
                size_t n = (size_t)(placed_off[c + 1] - placed_off[c]);
                const go_seg *an = anno + anno_off[(size_t)a * C + c];
                size_t m = (size_t)(anno_off[(size_t)a * C + c + 1] - anno_off[(size_t)a * C + c]);
                double x = go_counter(counters[k], s, n, an, m, cws_nseg ? cws_nseg[c] : 0);
                if (counters[k] == GO_NUCLEOTIDE_DENSITY && cws_nseg && cws_nseg[c] != 0) {
                    double t = total + x;
                    if (fabs(total) >= fabs(x)) comp += (total - t) + x; else comp += (x - t) + total;
                    total = t;
                } else total += x;
            }
            if (comp != 0.0 && isfinite(comp)) total += comp;
            counts[(size_t)k * A + a] = total;
        }
}

/* gat/Engine.pyx:653-737 SamplerSegments.sample(segments, workspace): exactly len(segments) placements
 * (lengths from the workspace-overlapping segments), returned in draw order -- unsorted, unmerged */
long go_sampler_segments(const go_seg *segments, size_t n, const go_seg *workspace, size_t m,
                         uint32_t bucket_size, uint32_t nbuckets,
                         go_randint_fn rnd, go_turn_fn next_turn, void *ctx, go_seg *out, size_t cap)
{
    long result = -3;
    go_seg *working = (go_seg *)malloc(sizeof(go_seg) * (n ? n : 1));
    int64_t *histogram = (int64_t *)malloc(sizeof(int64_t) * nbuckets);
    uint32_t *hcdf = (uint32_t *)malloc(sizeof(uint32_t) * nbuckets);
    uint32_t *wcdf = NULL;
    if (!working || !histogram || !hcdf) goto done;
    size_t nw = go_filter(segments, n, workspace, m, working);           /* :704-708 */
    if (nw == 0) { result = 0; goto done; }
    uint32_t bucket = go_length_distribution(working, nw, bucket_size, nbuckets, histogram);   /* :711-714 */
    if (bucket == 0) { result = -2; goto done; }
    uint32_t htotal = 0;
    for (uint32_t i = 0; i < nbuckets; i++) { htotal += (uint32_t)histogram[i]; hcdf[i] = htotal; }
    uint32_t wtotal = 0;
    wcdf = build_ws_cdf(workspace, m, &wtotal);                           /* :717 */
    if (!wcdf) goto done;
    if (n > cap) { result = -1; goto done; }
    for (size_t x = 0; x < n; x++) {                                      /* :719 for x in xrange(len(segments)) */
        if (next_turn) next_turn(ctx);
        uint32_t r = 1;
        if (htotal > 1) r = (uint32_t)rnd(ctx, GO_SLOT_LEN, 1, (int64_t)htotal);
        uint32_t base = (uint32_t)go_searchsorted_u32(hcdf, nbuckets, r) * bucket;
        if (bucket > 1) base += (uint32_t)rnd(ctx, GO_SLOT_JITTER, 0, (int64_t)bucket);
        uint32_t start, end;
        int32_t overlap;
        if (go_segmentlist_sample(workspace, m, wcdf, wtotal, base, rnd, ctx, GO_SLOT_WS_R, GO_SLOT_WS_P,
                                  &start, &end, &overlap) != 0) { result = -4; goto done; }
        out[x].start = start; out[x].end = end;
    }
    result = (long)n;
done:
    free(working); free(histogram); free(hcdf); free(wcdf);
    return result;
}

/* ------------------------------------------------------------------------------------------------
 * gat/Engine.pyx:998-1111 SamplerShift.sample(segments, workspace): every workspace-overlapping segment is
 * moved to a random position of the workspace within +-shift_area of its midpoint and wrapped around the
 * ends of that local workspace.  The C integer conversions of the Cython code (Position = uint32,
 * PositionDifference = int32) are kept as they are. */
typedef struct { go_seg *s; size_t n, cap; } seg_vec;
static int vec_push(seg_vec *v, go_seg x)
{
    if (v->n == v->cap) {
        size_t c = v->cap ? 2 * v->cap : 64;
        go_seg *t = (go_seg *)realloc(v->s, c * sizeof(go_seg));
        if (!t) return -1;
        v->s = t; v->cap = c;
    }
    v->s[v->n++] = x;
    return 0;
}

/* gat/SegmentList.pyx:1314-1355 getFilledSegmentsFromStart(start, remainder); `out` receives the
 * normalized result (:1354) */
static int shift_fill_from_start(const go_seg *w, size_t n, uint32_t start, int32_t remainder, seg_vec *out)
{
    size_t first = out->n;
    if ((uint32_t)remainder > go_sum(w, n)) {                             /* :1325-1326 clone of self */
        for (size_t i = 0; i < n; i++) if (vec_push(out, w[i])) return -1;
        return 0;
    }
    go_seg probe = { start, start + 1u };
    int idx = go_get_insertion_point(w, n, probe);                         /* :1331 */
    if (idx == (int)n) idx -= 1; else if (idx == -1) idx = 0;              /* :1336-1337 */
    while (remainder > 0) {                                                /* :1340-1352 */
        if (w[idx].end < start) {
        } else {
            start = (uint32_t)i32max((int32_t)w[idx].start, (int32_t)start);
            uint32_t end = (uint32_t)i32min((int32_t)w[idx].end, (int32_t)(start + (uint32_t)remainder));
            remainder -= (int32_t)(end - start);
            go_seg x = { start, end };
            if (vec_push(out, x)) return -1;
        }
        idx += 1;
        if (idx == (int)n) { idx = 0; start = w[idx].start; }
    }
    out->n = first + go_normalize(out->s + first, out->n - first);
    return 0;
}

/* gat/SegmentList.pyx:1357-1399 getFilledSegmentsFromEnd(end, remainder) */
static int shift_fill_from_end(const go_seg *w, size_t n, uint32_t end, int32_t remainder, seg_vec *out)
{
    size_t first = out->n;
    if ((uint32_t)remainder > go_sum(w, n)) {
        for (size_t i = 0; i < n; i++) if (vec_push(out, w[i])) return -1;
        return 0;
    }
    go_seg probe = { end, end + 1u };
    int idx = go_get_insertion_point(w, n, probe);
    if (idx == (int)n) idx -= 1; else if (idx == -1) idx = 0;
    while (remainder > 0) {
        if (w[idx].start > end) {
        } else {
            end = (uint32_t)i32min((int32_t)w[idx].end, (int32_t)end);
            uint32_t start = (uint32_t)i32max((int32_t)w[idx].start, (int32_t)(end - (uint32_t)remainder));
            remainder -= (int32_t)(end - start);
            go_seg x = { start, end };
            if (vec_push(out, x)) return -1;
        }
        idx -= 1;
        if (idx < 0) { idx = (int)n - 1; end = w[idx].end; }
    }
    out->n = first + go_normalize(out->s + first, out->n - first);
    return 0;
}

/* Returns the number of segments written to out (normalized), -1 capacity, -3 memory, -5 where the
 * reference would stop (unused now: an empty local workspace drops the segment, see below).
 * Draws per working segment x (turn x): GO_SLOT_SHIFT_POS = getRandomPosition (:1083),
 * GO_SLOT_SHIFT_DIR = direction (:1084). */
long go_sampler_shift(const go_seg *segments, size_t n, const go_seg *workspace, size_t m,
                      double radius, int32_t extension,
                      go_randint_fn rnd, go_turn_fn next_turn, void *ctx, go_seg *out, size_t cap)
{
    long result = -3;
    const double half_radius = radius / 2;                                /* :1054 */
    const int32_t half_extension = extension / 2;                         /* :1055 (extension >= 0) */
    go_seg *working = (go_seg *)malloc(sizeof(go_seg) * (n ? n : 1));
    go_seg *ws = (go_seg *)malloc(sizeof(go_seg) * (m ? m : 1));
    seg_vec sample = { NULL, 0, 0 };
    if (!working || !ws) goto done;
    size_t nw = go_filter(segments, n, workspace, m, working);            /* :1060-1062 */
    if (nw == 0) { result = 0; goto done; }                               /* :1067-1068 */
    for (size_t x = 0; x < nw; x++) {                                     /* :1070 */
        if (next_turn) next_turn(ctx);
        const go_seg segment = working[x];
        const uint32_t length = segment.end - segment.start;
        const uint32_t midpoint = segment.start + length / 2;
        int32_t shift_area;
        if (extension) shift_area = half_extension;                        /* :1074-1077 */
        else shift_area = (int32_t)(uint32_t)floor((double)length * half_radius);
        int32_t ws_start = i32max(0, (int32_t)(midpoint - (uint32_t)shift_area));      /* :1080-1081 */
        int32_t ws_end = i32max(0, (int32_t)(midpoint + (uint32_t)shift_area));
        /* :1082 workspace.getOverlappingSegmentsWithRange(ws_start, ws_end) = getOverlappingSegments
         * (gat/SegmentList.pyx:957-985): from the insertion point while start <= range end */
        go_seg other = { (uint32_t)ws_start, (uint32_t)ws_end };
        int idx = go_get_insertion_point(workspace, m, other);
        if (idx == (int)m) idx -= 1; else if (idx == -1) idx = 0;
        size_t k = 0;
        while (idx < (int)m && workspace[idx].start <= other.end) ws[k++] = workspace[idx++];
        /* :1083 ws.truncate(Segment(ws_start, ws_end)) (gat/SegmentList.pyx:1186-1202) */
        for (size_t i = 0; i < k; i++) {
            go_seg *s = &ws[i];
            if (s->end < other.start) s->start = s->end = 0;
            else if (s->start > other.end) s->start = s->end = 0;
            else {
                if (s->start < other.start) s->start = other.start;
                if (s->end > other.end) s->end = other.end;
            }
        }
        k = go_normalize(ws, k);
        /* :1085 ws.getRandomPosition() (gat/SegmentList.pyx:902-915) */
        const uint32_t total = go_sum(ws, k);
        if (total == 0) {
            /* numpy.random.randint(0, 0) raises ValueError inside the cpdef getRandomPosition, whose C return
             * type cannot carry it: the exception is printed and ignored, 0 comes back (:902-905), the direction
             * is still drawn (:1084) and every fill of the empty local workspace returns nothing -- the segment
             * silently drops out of the sample */
            (void)rnd(ctx, GO_SLOT_SHIFT_DIR, 0, 2);
            continue;
        }
        uint32_t pos = (uint32_t)rnd(ctx, GO_SLOT_SHIFT_POS, 0, (int64_t)total);
        int32_t start = 0, end;
        int found = 0;
        for (size_t i = 0; i < k; i++) {
            uint32_t l = ws[i].end - ws[i].start;
            if (pos > l) pos -= l;
            else { start = (int32_t)(ws[i].start + pos); found = 1; break; }
        }
        if (!found) { result = -6; goto done; }                           /* `assert False` :915 */
        if (rnd(ctx, GO_SLOT_SHIFT_DIR, 0, 2)) end = (int32_t)((uint32_t)start + length);      /* :1086-1090 */
        else { end = start; start = (int32_t)((uint32_t)end - length); }
        ws_start = (int32_t)ws[0].start; ws_end = (int32_t)ws[k - 1].end;  /* :1093 ws.min(), ws.max() */
        int32_t remainder;
        int rc;
        if (start < ws_start) {                                            /* :1096-1100 */
            remainder = i32min(ws_start - start, (int32_t)length);
            rc = shift_fill_from_start(ws, k, (uint32_t)start, (int32_t)(length - (uint32_t)remainder), &sample);
            if (!rc) rc = shift_fill_from_end(ws, k, (uint32_t)ws_end, remainder, &sample);
        } else if (end > ws_end) {                                         /* :1101-1105 */
            remainder = i32min(end - ws_end, (int32_t)length);
            rc = shift_fill_from_end(ws, k, (uint32_t)end, (int32_t)(length - (uint32_t)remainder), &sample);
            if (!rc) rc = shift_fill_from_start(ws, k, (uint32_t)ws_start, remainder, &sample);
        } else rc = shift_fill_from_start(ws, k, (uint32_t)start, (int32_t)length, &sample);   /* :1107 */
        if (rc) goto done;
    }
    sample.n = go_normalize(sample.s, sample.n);                          /* :1109 */
    if (sample.n > cap) { result = -1; goto done; }
    memcpy(out, sample.s, sizeof(go_seg) * sample.n);
    result = (long)sample.n;
done:
    free(working); free(ws); free(sample.s);
    return result;
}

/* which sampler go_compute_sample_philox places with: 0 = SamplerAnnotator, 1 = SamplerSegments,
 * 2 = SamplerShift (parameters from go_set_shift_params) */
static int g_sampler_kind = 0;
static double g_shift_radius = 2.0;
static int32_t g_shift_extension = 0;
void go_set_sampler_kind(int kind) { g_sampler_kind = kind; }
void go_set_shift_params(double radius, int32_t extension) { g_shift_radius = radius; g_shift_extension = extension; }

/* gat/__init__.py:494-591 computeSample, with the Philox stream of the CUDA kernel */
/* room per unit for the sampler's result (0: twice / four times the unit's segments + 64); tests of skewed units,
 * whose samples hold many more segments than the input, raise it */
static size_t g_unit_cap = 0;
void go_set_unit_cap(size_t cap) { g_unit_cap = cap; }

int go_compute_sample_philox(int U, int C, int A, const int32_t *unit_contig, int has_isochores,
                             const uint64_t *seg_off, const go_seg *seg,
                             const uint64_t *ws_off, const go_seg *ws,
                             const uint64_t *anno_off, const go_seg *anno,
                             const uint32_t *cws_nseg,
                             uint32_t bucket_size, uint32_t nbuckets,
                             uint64_t seed, uint32_t track, uint32_t sample,
                             int ncounters, const int32_t *counters, double *counts,
                             uint64_t *placed_off, go_seg *placed, size_t placed_cap)
{
    /* per-unit placement (:531-546); units with empty segments/workspace are skipped (:536-538) */
    size_t total_cap = 0;
    /* SamplerShift cuts a moved segment at workspace gaps: room for more pieces (exceeding it is reported) */
    const size_t cap_mul = g_sampler_kind == 2 ? 4 : 2;
    for (int u = 0; u < U; u++) {
        size_t cap = cap_mul * (size_t)(seg_off[u + 1] - seg_off[u]) + 64;
        total_cap += cap > g_unit_cap ? cap : g_unit_cap;
    }
    go_seg *unit_out = (go_seg *)malloc(sizeof(go_seg) * (total_cap ? total_cap : 1));
    size_t *unit_n = (size_t *)calloc((size_t)(U ? U : 1), sizeof(size_t));
    size_t *unit_o = (size_t *)calloc((size_t)(U ? U : 1), sizeof(size_t));
    go_seg *contig_buf = (go_seg *)malloc(sizeof(go_seg) * (total_cap ? total_cap : 1));
    uint64_t *coff = (uint64_t *)calloc((size_t)C + 1, sizeof(uint64_t));
    int rc = 0;
    if (!unit_out || !unit_n || !unit_o || !contig_buf || !coff) { rc = -3; goto done; }

    size_t o = 0;
    for (int u = 0; u < U; u++) {
        size_t n = (size_t)(seg_off[u + 1] - seg_off[u]), m = (size_t)(ws_off[u + 1] - ws_off[u]);
        size_t cap = cap_mul * n + 64;
        if (g_unit_cap > cap) cap = g_unit_cap;
        unit_o[u] = o;
        if (n == 0 || m == 0) { unit_n[u] = 0; o += cap; continue; }
        go_philox_ctx ctx;
        go_philox_begin(&ctx, seed, track, (uint32_t)u, sample);
        long r = g_sampler_kind == 2
            ? go_sampler_shift(seg + seg_off[u], n, ws + ws_off[u], m, g_shift_radius, g_shift_extension,
                               go_philox_randint, go_philox_next_turn, &ctx, unit_out + o, cap)
            : g_sampler_kind == 1
            ? go_sampler_segments(seg + seg_off[u], n, ws + ws_off[u], m, bucket_size, nbuckets,
                                  go_philox_randint, go_philox_next_turn, &ctx, unit_out + o, cap)
            : go_sampler_annotator(seg + seg_off[u], n, ws + ws_off[u], m, bucket_size, nbuckets,
                                   go_philox_randint, go_philox_next_turn, &ctx, unit_out + o, cap, NULL);
        if (r < 0) { rc = (int)r; goto done; }
        unit_n[u] = (size_t)r;
        o += cap;
    }
    /* :563 sample.fromIsochores(): extend per contig in unit order, then merge(0) when any key had
     * an isochore suffix (gat/Engine.pyx:2857-2876) */
    size_t w = 0;
    for (int c = 0; c < C; c++) {
        coff[c] = w;
        size_t b = w;
        for (int u = 0; u < U; u++)
            if (unit_contig[u] == c) { memcpy(contig_buf + w, unit_out + unit_o[u], sizeof(go_seg) * unit_n[u]); w += unit_n[u]; }
        if (has_isochores) w = b + go_merge(contig_buf + b, w - b, 0);
    }
    coff[C] = w;
    go_count_placed(C, A, coff, contig_buf, anno_off, anno, cws_nseg, ncounters, counters, counts);
    if (placed_off && placed) {
        if (w > placed_cap) { rc = -1; goto done; }
        memcpy(placed_off, coff, sizeof(uint64_t) * ((size_t)C + 1));
        memcpy(placed, contig_buf, sizeof(go_seg) * w);
    }
done:
    free(unit_out); free(unit_n); free(unit_o); free(contig_buf); free(coff);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * numpy's pairwise summation of a contiguous float64 array (numpy/_core/src/umath/loops_utils.h.src,
 * PW_BLOCKSIZE 128, 8 accumulators): numpy.mean / numpy.std reduce with it, so restating its
 * association order makes expected/stddev bit-identical to the reference (gat/Engine.pyx:1671,1684) */
static double np_pairwise_sum(const double *a, size_t n)
{
    if (n < 8) {
        double res = 0.;
        for (size_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        size_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        size_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* gat/Engine.pyx:1543-1576 getTwoSidedPValue on the sorted samples */
double go_two_sided_pvalue(const double *s, size_t l, double expected, double val)
{
    long idx = go_searchargsorted_f64(s, NULL, l, val);
    double min_pval = 1.0 / (double)l;
    if (idx == (long)l) idx = 1;
    else if (val > expected) {
        while (idx > 0 && s[idx] == val) idx -= 1;
        idx = (long)l - (idx + 1);
    } else {
        while (idx < (long)l && s[idx] == val) idx += 1;
    }
    double pval = (double)idx / (double)l;
    return pval > min_pval ? pval : min_pval;
}

/* gat/Engine.pyx:1635-1718 makeEnrichmentStatistics */
int go_enrichment_statistics(double observed, const double *samples, size_t l, int has_reference,
                             double ref_fold, double pseudo_count, go_stats *out)
{
    if (l < 1) return -1;
    double *sorted = (double *)malloc(sizeof(double) * l);
    double *dev = (double *)malloc(sizeof(double) * l);
    if (!sorted || !dev) { free(sorted); free(dev); return -3; }
    memcpy(sorted, samples, sizeof(double) * l);
    qsort(sorted, l, sizeof(double), cmp_double);          /* numpy.argsort (:1660); value order only */

    out->observed = observed;
    out->nsamples = (uint32_t)l;
    double mean = np_pairwise_sum(samples, l) / (double)l;   /* numpy.mean (:1671) */
    out->expected = mean;
    if (has_reference) out->expected *= ref_fold;            /* :1673-1676 */
    if (out->expected != 0) out->fold = (observed + pseudo_count) / (out->expected + pseudo_count);
    else out->fold = 1.0;                                    /* :1679-1682 */
    for (size_t i = 0; i < l; i++) { double d = samples[i] - mean; dev[i] = d * d; }
    out->stddev = sqrt(np_pairwise_sum(dev, l) / (double)l); /* numpy.std, ddof 0 (:1684) */

    uint32_t offset = (uint32_t)(int)(0.05 * (double)l);     /* :1689-1696 */
    if (offset > 0) {
        out->lower95 = sorted[i32min((int32_t)offset, (int32_t)l - 1)];
        out->upper95 = sorted[i32max((int32_t)l - (int32_t)offset, 0)];
    } else {
        out->lower95 = sorted[0];
        out->upper95 = sorted[l - 1];
    }
    if (!has_reference) out->pvalue = go_two_sided_pvalue(sorted, l, out->expected, observed);
    else {                                                    /* :1703-1714 */
        if (!(ref_fold > 0)) { free(sorted); free(dev); return -2; }
        out->pvalue = go_two_sided_pvalue(sorted, l, out->expected, observed / ref_fold);
        out->lower95 *= ref_fold;
        out->upper95 *= ref_fold;
    }
    out->qvalue = 1.0;
    free(sorted); free(dev);
    return 0;
}
