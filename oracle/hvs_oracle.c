/*
 * hvs_oracle.c -- CPU restatement of the reference's filtered k-NN solve step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the CUDA engine, the
 * C-ABI library, the vec_query shim) may include, link or call this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU baseline.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle_*.py) against
 *   - the reference's only known-answer vector, src/fp_inaccuracy_test.cpp:77-97
 *     (scalar 277762.34375, AVX2-order 277762.28125),
 *   - outputs of the reference itself (baseline.out / optimized.out built from
 *     /root/reference by oracle/Makefile into oracle/_ref/), committed as
 *     fixtures under tests/golden/ by oracle/gen_golden.py,
 *   - live differential runs against oracle/_ref/libref_*.so when present.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  Plain C, scalar, no FMA contraction (build with
 * -ffp-contract=off; the reference is built -O3 -mavx2 without -mfma,
 * CMakeLists.txt:8, so it has no fused multiply-add either).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define HVS_K 100        /* KNN_LIMIT, include/optimized_impl.h:26 */
#define HVS_DIM 100      /* VEC_DIM - 2, include/optimized_impl.h:28 */
#define HVS_DROW 102     /* data row: [C, T, x0..x99], README.md:32-44, src/test.cpp:66 */
#define HVS_QROW 104     /* query row: [type, v, l, r, q0..q99], src/test.cpp:76 */

/* include/baseline.hpp:53-64 (compare_with_id) and include/io.h:38-48 (calc_dist):
 * sequential fp32 sum over the 100 vector dims of (a-b)^2, sub/mul/add, no FMA. */
float hvs_oracle_dist_seq(const float *x /*100*/, const float *q /*100*/)
{
    float sum = 0.0f;
    for (int i = 0; i < HVS_DIM; ++i) {
        float diff = x[i] - q[i];
        sum += diff * diff;
    }
    return sum;
}

/* include/optimized_impl.h:96-125 + hsum256_ps_avx :37-47, emulated in scalar code.
 * Eight partial sums; lane j accumulates dims j, j+8, ..., j+88 (12 full 8-wide
 * steps over dims 0..95), then the masked tail load at row index VEC_DIM-8 = 94
 * (vector dim 92) keeps the upper four lanes, so dims 96..99 go to lanes 4..7.
 * Horizontal sum: (lo128 + hi128) -> t[0..3]; then (t0+t1) + (t2+t3). */
float hvs_oracle_dist_avx_order(const float *x /*100*/, const float *q /*100*/)
{
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 96; i += 8)
        for (int j = 0; j < 8; ++j) {
            float diff = x[i + j] - q[i + j];
            lane[j] += diff * diff;
        }
    for (int j = 4; j < 8; ++j) {
        float diff = x[92 + j] - q[92 + j];
        lane[j] += diff * diff;
    }
    float t0 = lane[0] + lane[4], t1 = lane[1] + lane[5];
    float t2 = lane[2] + lane[6], t3 = lane[3] + lane[7];
    /* movehdup: shuf = {t1,t1,t3,t3}; sums = t + shuf = {t0+t1, ., t2+t3, .};
     * movehl -> high half to low: sums[0] + sums[2] */
    float s01 = t0 + t1, s23 = t2 + t3;
    return s01 + s23;
}

/* float -> uint32_t / int32_t conversions as the reference's x86-64 build performs
 * them (include/baseline.hpp:90-91: `uint32_t query_type = queries[i][0]; int32_t v =
 * queries[i][1];`).  In range these truncate toward zero.  Out of range / NaN is UB
 * in C++; we pin the x86-64 behaviour (cvttss2si: 64-bit convert then low 32 bits for
 * the unsigned case, "integer indefinite" 0x80000000 for int32). */
static uint32_t f2u32_x86(float f)
{
    if (!(f > -9.2233720368547758e18f && f < 9.2233720368547758e18f)) return 0u; /* low 32 bits of 0x8000000000000000 */
    return (uint32_t)(uint64_t)(int64_t)f;
}
static int32_t f2i32_x86(float f)
{
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int32_t)f;
}

uint32_t hvs_oracle_query_type(float t) { return f2u32_x86(t); }
int32_t hvs_oracle_query_cat(float v) { return f2i32_x86(v); }

/* Predicate of include/baseline.hpp:107-136 for one data row (row -> [C, T, ...]).
 * `nodes[j][0] == v` compares float C with int32 v promoted to float. */
int hvs_oracle_match(const float *row, uint32_t type, int32_t v, float l, float r)
{
    switch (type) {
    case 0: return 1;
    case 1: return row[0] == (float)v;
    case 2: return row[1] >= l && row[1] <= r;
    case 3: return row[0] == (float)v && row[1] >= l && row[1] <= r;
    default: return 0; /* no branch taken: empty candidate set, baseline.hpp:107-136 */
    }
}

typedef struct { float d; uint32_t pos; uint32_t id; } cand_t;

/* strict weak order "closer first, earlier candidate position first on ties" */
static inline int cand_less(const cand_t *a, const cand_t *b)
{
    return (a->d < b->d) || (a->d == b->d && a->pos < b->pos);
}

/* max-heap on cand_less over h[0..n) */
static void heap_sift_down(cand_t *h, int n, int i)
{
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && cand_less(&h[m], &h[l])) m = l;
        if (r < n && cand_less(&h[m], &h[r])) m = r;
        if (m == i) return;
        cand_t t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}
static void heap_push(cand_t *h, int *n, cand_t c)
{
    int i = (*n)++;
    h[i] = c;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!cand_less(&h[p], &h[i])) break;
        cand_t t = h[i]; h[i] = h[p]; h[p] = t;
        i = p;
    }
}
static int cand_cmp_qsort(const void *a, const void *b)
{
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (cand_less(x, y)) return -1;
    if (cand_less(y, x)) return 1;
    return 0;
}

/*
 * include/baseline.hpp:68-190 (vec_query, IMPL=1 -- the oracle of record).
 *   nodes   : n  x 102 floats, row-major          (io.h:111-136 ReadBin layout)
 *   queries : nq x 104 floats, row-major
 *   out_ids : nq x 100 uint32 original row ids, ascending by distance
 *   out_dist: nq x 100 fp32 distances of those ids (may be NULL)
 *   out_nmatch: per query, number of rows that satisfied the predicate (may be NULL)
 *
 * The reference materialises every candidate, computes all distances and
 * std::sort()s them (baseline.hpp:149-172).  std::sort is unstable, so the
 * order among equal distances is unspecified there; this restatement makes it
 * deterministic (candidate position breaks ties) and keeps only a 100-entry
 * max-heap instead of sorting everything -- same result set, same order up to
 * exact-distance ties.
 *
 * Returns 0, or -1 if n < 100 (the reference's pad loop at baseline.hpp:138-147
 * would index nodes[n - s] out of bounds).
 */
int hvs_oracle_vec_query(const float *nodes, uint32_t n, const float *queries, uint32_t nq,
                         float sample_proportion, uint32_t *out_ids, float *out_dist,
                         uint32_t *out_nmatch)
{
    if (n < HVS_K) return -1;
    uint32_t sn = (uint32_t)(sample_proportion * n); /* baseline.hpp:74 */
    if (sn > n) sn = n;
    cand_t heap[HVS_K];
    for (uint32_t i = 0; i < nq; ++i) {
        const float *qrow = queries + (size_t)i * HVS_QROW;
        uint32_t type = f2u32_x86(qrow[0]);     /* baseline.hpp:90 */
        int32_t v = f2i32_x86(qrow[1]);         /* baseline.hpp:91 */
        float l = qrow[2], r = qrow[3];         /* baseline.hpp:92-93 */
        const float *q = qrow + 4;              /* baseline.hpp:96-100: query dims align with data dims 2.. */
        int hn = 0;
        uint32_t pos = 0, nmatch = 0;
        for (uint32_t j = 0; j < sn; ++j) {     /* baseline.hpp:107-136 */
            const float *row = nodes + (size_t)j * HVS_DROW;
            if (!hvs_oracle_match(row, type, v, l, r)) continue;
            cand_t c = { hvs_oracle_dist_seq(row + 2, q), pos++, j }; /* baseline.hpp:152-155 */
            ++nmatch;
            if (hn < HVS_K) heap_push(heap, &hn, c);
            else if (cand_less(&c, &heap[0])) { heap[0] = c; heap_sift_down(heap, hn, 0); }
        }
        /* baseline.hpp:138-147: pad with n-1, n-2, ... ignoring the predicate, no de-duplication,
         * until there are exactly 100 candidates; padded ids are ranked with the real matches. */
        if (nmatch < HVS_K) {
            uint32_t s = 1;
            while (pos < HVS_K) {
                uint32_t j = n - s;
                const float *row = nodes + (size_t)j * HVS_DROW;
                cand_t c = { hvs_oracle_dist_seq(row + 2, q), pos++, j };
                heap_push(heap, &hn, c);
                ++s;
            }
        }
        qsort(heap, (size_t)hn, sizeof(cand_t), cand_cmp_qsort); /* baseline.hpp:159-172 */
        for (int k = 0; k < HVS_K; ++k) {
            out_ids[(size_t)i * HVS_K + k] = heap[k].id;
            if (out_dist) out_dist[(size_t)i * HVS_K + k] = heap[k].d;
        }
        if (out_nmatch) out_nmatch[i] = nmatch;
    }
    return 0;
}

/* include/io.h:50-78 (SaveKNNFull): recompute, for given ids, the sequential fp32
 * distance between each result row and its query; this is what the reference's own
 * acceptance test compares (src/compare_data.cpp:40-60, abs tol 0.002). */
void hvs_oracle_rescore(const float *nodes, const float *queries, uint32_t nq,
                        const uint32_t *ids, float *out_dist)
{
    for (uint32_t i = 0; i < nq; ++i) {
        const float *q = queries + (size_t)i * HVS_QROW + 4;
        for (int k = 0; k < HVS_K; ++k) {
            uint32_t j = ids[(size_t)i * HVS_K + k];
            out_dist[(size_t)i * HVS_K + k] = hvs_oracle_dist_seq(nodes + (size_t)j * HVS_DROW + 2, q);
        }
    }
}

/*
 * include/optimized_impl.h:179-438 (class Knn) + include/optimized.hpp:54-146,
 * restated scalar: unsorted 100-slot array, strict `<` replacement of the current
 * worst (first-seen wins at the boundary, :301-310), arg-max rescan on replace
 * (:201-274), final sort of the 100 pairs (:392-437).  Distance in AVX2 summation
 * order.  Used to pin the optimized variants' behaviour (ids can legitimately differ
 * from the baseline on near-ties, include/optimized.hpp:33-41).
 */
int hvs_oracle_vec_query_optimized(const float *nodes, uint32_t n, const float *queries, uint32_t nq,
                                   float sample_proportion, uint32_t *out_ids, float *out_dist)
{
    if (n < HVS_K) return -1;
    uint32_t sn = (uint32_t)(sample_proportion * n);
    if (sn > n) sn = n;
    float dist_array[HVS_K];
    uint32_t idx_array[HVS_K];
    for (uint32_t i = 0; i < nq; ++i) {
        const float *qrow = queries + (size_t)i * HVS_QROW;
        uint32_t type = f2u32_x86(qrow[0]);
        int32_t v = f2i32_x86(qrow[1]);
        float l = qrow[2], r = qrow[3];
        const float *q = qrow + 4;
        uint32_t fill = 0, worst = 0;
        uint32_t s = 0; /* pad counter; 0 = still scanning */
        for (uint32_t jj = 0;; ++jj) {
            uint32_t j;
            if (jj < sn) {
                j = jj;
                if (!hvs_oracle_match(nodes + (size_t)j * HVS_DROW, type, v, l, r)) continue;
            } else {
                /* optimized.hpp:119-128 */
                if (fill >= HVS_K) break;
                ++s;
                j = n - s;
            }
            float d = hvs_oracle_dist_avx_order(nodes + (size_t)j * HVS_DROW + 2, q);
            if (fill < HVS_K) {                     /* optimized_impl.h:301-310, not_full branch */
                /* worst = better_than_worst ? worst : fill   (dist_array[worst] read before insert) */
                float worst_dist = fill ? dist_array[worst] : dist_array[0];
                int better = fill ? (d < worst_dist) : 0;
                dist_array[fill] = d;
                idx_array[fill] = j;
                if (!better) worst = fill;
                ++fill;
            } else if (d < dist_array[worst]) {
                dist_array[worst] = d;
                idx_array[worst] = j;
                /* find_worst: arg-max; on equal maxima the SIMD version (:201-274) returns the
                 * largest index among lanes holding the max; any choice yields the same multiset */
                uint32_t w = 0;
                for (uint32_t k = 1; k < HVS_K; ++k)
                    if (dist_array[k] > dist_array[w]) w = k;
                worst = w;
            }
        }
        cand_t tmp[HVS_K];
        for (uint32_t k = 0; k < HVS_K; ++k) { tmp[k].d = dist_array[k]; tmp[k].pos = k; tmp[k].id = idx_array[k]; }
        qsort(tmp, HVS_K, sizeof(cand_t), cand_cmp_qsort);
        for (int k = 0; k < HVS_K; ++k) {
            out_ids[(size_t)i * HVS_K + k] = tmp[k].id;
            if (out_dist) out_dist[(size_t)i * HVS_K + k] = tmp[k].d;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Restatement of the reference's input generators, seedable.  glibc rand() stream, same
 * call order and the same mixed float/double arithmetic as src/write_data.c:27-33 and
 * src/write_query.c:29-53, so that `hvs_oracle_refgen_*(seed, ...)` is byte-identical to
 * `HVS_SEED=seed LD_PRELOAD=time_shim.so write_data|write_query` (checked in
 * tests/test_oracle_refgen.py when oracle/_ref is present).  Buffers exclude the
 * leading uint32 row count of the file format. */
static float unit_rand(void) { return (float)rand() / RAND_MAX; } /* float / (float)RAND_MAX */

void hvs_oracle_refgen_data(unsigned seed, uint32_t n, float *out /* n x 102 */)
{
    srand(seed); /* write_data.c:16 with time() pinned */
    for (uint32_t i = 0; i < n; ++i) {
        float *b = out + (size_t)i * HVS_DROW;
        b[0] = (float)(((1.0 - -1.0) * unit_rand()) + -1.0);      /* C ~ U(-1,1)  write_data.c:28 */
        b[1] = (float)(((3.0 - -3.0) * unit_rand()) + -3.0);      /* T ~ U(-3,3)  write_data.c:30 */
        for (int j = 2; j < HVS_DROW; ++j)
            b[j] = (float)(((6.00 - -6.00) * unit_rand()) + -6.00); /* write_data.c:31-33 */
    }
}

/* returns number of rows written (stops early like the reference's `goto end` if the
 * type draw lands on 4, write_query.c:49) */
uint32_t hvs_oracle_refgen_query(unsigned seed, uint32_t m, float *out /* m x 104 */)
{
    srand(seed); /* write_query.c:18 */
    for (uint32_t i = 0; i < m; ++i) {
        float *b = out + (size_t)i * HVS_QROW;
        int qt = (int)(4 * unit_rand());                          /* write_query.c:30 */
        b[0] = (float)qt;
        switch (qt) {
        case 3:
            b[1] = (float)(((1.0 - -1.0) * unit_rand()) + -1.0);
            b[2] = (float)(((3.0 - -3.0) * unit_rand()) + -3.0);
            b[3] = (float)(((4.0 - b[2]) * unit_rand()) + b[2]);
            break;
        case 2:
            b[1] = -1.0f;
            b[2] = (float)(((3.0 - -3.0) * unit_rand()) + -3.0);
            b[3] = (float)(((4.0 - b[2]) * unit_rand()) + b[2]);
            break;
        case 1:
            b[1] = (float)(((1.0 - -1.0) * unit_rand()) + -1.0);
            b[2] = -1.0f; b[3] = -1.0f;
            break;
        case 0:
            b[1] = -1.0f; b[2] = -1.0f; b[3] = -1.0f;
            break;
        default:
            return i;
        }
        for (int j = 4; j < HVS_QROW; ++j)
            b[j] = (float)(((6.00 - -6.00) * unit_rand()) + -6.00); /* write_query.c:51-53 */
    }
    return m;
}
