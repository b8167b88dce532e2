/*
 * hvs.h -- C ABI of the B200-native filtered exact k-NN engine (libhvs_b200.so).
 *
 * This is the drop-in boundary for ONE path of atalantus/Project---Hybrid-Vector-Search-Queries:
 * the body of vec_query() -- predicate -> squared-L2 -> top-100 -- which the reference defines
 * three times behind one signature (include/baseline.hpp:68-69, include/optimized.hpp:54-55,
 * include/optimized_parallel.hpp:61-62) and calls once from src/test.cpp:85.  The reference has
 * no FFI; its "plugin API" is that one free function, chosen at compile time by -DIMPL
 * (src/test.cpp:6-13).  A 30-line C++ shim with the same signature
 * (include/hvs_vec_query.hpp, selected with -DIMPL=4) flattens the nested vectors and calls the
 * entry points below; INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes, no C++ / torch types; every function returns
 * HVS_OK (0) or a negative hvs_status; hvs_last_error() gives the message.  There is NO CPU
 * fallback: without a usable sm_100 device hvs_create() fails with HVS_ERR_NO_DEVICE.
 * Host buffers are caller-owned; the engine owns all device memory.
 *
 * Layouts (little-endian fp32, README.md:32-44 / include/io.h:111-136 of the reference):
 *   data row   : 102 floats [C, T, x0..x99]            (ReadBin(path, 102, ...), src/test.cpp:66-72)
 *   query row  : 104 floats [type, v, l, r, q0..q99]   (ReadBin(path, 104, ...), src/test.cpp:76-78)
 *   result row : 100 uint32 original row ids, ascending by distance (SaveKNN, include/io.h:23-36)
 */
#ifndef HVS_H_
#define HVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HVS_API __attribute__((visibility("default")))
#else
#define HVS_API
#endif

#define HVS_ABI_VERSION 1u
#define HVS_K 100u          /* KNN_LIMIT, include/optimized_impl.h:26 */
#define HVS_DIM 100u        /* VEC_DIM - 2, include/optimized_impl.h:28 */
#define HVS_DATA_ROW 102u   /* src/test.cpp:66 */
#define HVS_QUERY_ROW 104u  /* src/test.cpp:76 */

typedef enum hvs_status {
    HVS_OK = 0,
    HVS_ERR_INVALID = -1,     /* bad argument (NULL, n < 100: the reference's pad loop would read nodes[n-s] out of bounds, include/baseline.hpp:138-147) */
    HVS_ERR_NO_DEVICE = -2,   /* no CUDA device / not sm_100: no CPU fallback exists */
    HVS_ERR_CUDA = -3,        /* a CUDA runtime call or kernel failed */
    HVS_ERR_STATE = -4,       /* call order (solve before index_build, ...) */
    HVS_ERR_NOMEM = -5
} hvs_status;

/* Which kernels solve() may use for the distance + top-100 step. */
typedef enum hvs_mode {
    HVS_MODE_AUTO = 0,    /* slices that queries share: tcgen05 FP16 candidate sweep (K3) + exact FP32 re-rank (K5); sparse or tiny
                             slices and tiny jobs: direct scan (K4).  One tile-kernel family per solve.  Results are exact. */
    HVS_MODE_EXACT = 1,   /* FP32 only: FFMA tile kernel (K2) + direct scan (K4); never touches tensor cores */
    HVS_MODE_DIRECT = 2,  /* direct streaming scan only (K4): reference arithmetic order, no approximation anywhere */
    HVS_MODE_TENSOR = 3   /* as AUTO, but the tcgen05 FP16 sweep (K3) is used wherever a tile sweep is possible, however small the job */
} hvs_mode;

#define HVS_FLAG_USE_GIVEN_STREAM 1u   /* hvs_config.stream is used as given, even if it is NULL (= legacy default stream) */
#define HVS_FLAG_MARGIN_AUDIT 2u       /* K5 records, over every re-ranked survivor, the largest |approximate score - reference
                                          distance| relative to the bound the candidate margins assume (hvs_stats.margin_audit;
                                          must stay below 1).  Costs one atomic per survivor; results are unchanged. */

typedef struct hvs_engine hvs_engine;

typedef struct hvs_config {
    uint32_t struct_size;   /* sizeof(hvs_config); lets the struct grow */
    int32_t device;         /* CUDA device ordinal; -1 = current device */
    uint32_t mode;          /* hvs_mode */
    uint32_t flags;         /* HVS_FLAG_* */
    void *stream;           /* cudaStream_t to enqueue on.  NULL = the engine creates its own NON-BLOCKING stream (it does not
                               synchronise with the legacy default stream), unless HVS_FLAG_USE_GIVEN_STREAM is set, in which
                               case NULL means the legacy default stream itself.  The *_device entry points read and write
                               their device buffers in stream order on THIS stream: work the caller queued on another stream
                               (including an NCCL collective that produced an input) must be complete, or ordered before the
                               call with an event the engine's stream waits on, before the call is made. */
    uint32_t id_offset;     /* added to every returned row id (data-sharded variant: first global row of this shard) */
    uint32_t reserved;
} hvs_config;

/* Per-solve statistics (all times in milliseconds, CUDA events on the engine's stream). */
typedef struct hvs_stats {
    uint32_t struct_size;
    uint32_t n;                 /* rows indexed (sn of include/baseline.hpp:74) */
    uint32_t n_total;           /* rows given (n) */
    uint32_t m;                 /* queries of the last solve */
    uint64_t pairs;             /* sum over queries of max(slice length, pad) -- SURVEY 8d unit of work */
    uint64_t pairs_computed;    /* (query,row) distances the kernels actually evaluated (tile sweeps over-compute at slice edges) */
    uint64_t rows_union;        /* distinct arena rows touched */
    uint32_t n_direct;          /* queries solved by the direct scan kernel */
    uint32_t n_tile;            /* queries solved by a tile sweep (FFMA or tensor) + finalize */
    uint32_t n_items_ffma;      /* tile work items run on the FFMA kernel */
    uint32_t n_items_tensor;    /* tile work items run on the tcgen05 kernel */
    uint32_t n_fallback;        /* queries re-solved by the direct kernel because a candidate buffer overflowed its margin guarantee */
    uint32_t launches;          /* kernels launched by the last solve */
    float ms_index_build;       /* last hvs_index_build*, device time */
    float ms_h2d;               /* query upload (host entry points only) */
    float ms_plan;              /* slice search kernel + planner (device kernels and their sort; one 112-byte read-back) */
    float ms_direct;            /* K4 */
    float ms_tile;              /* K2 + K3 */
    float ms_tile_ffma;         /* K2 alone (sum of its launches) */
    float ms_tile_tensor;       /* K3 alone */
    float ms_finalize;          /* K5 */
    float ms_d2h;               /* result download (host entry points only) */
    float ms_solve_device;      /* plan .. finalize, device time */
    float ms_solve_wall;        /* whole call, host wall clock */
    uint64_t pairs_tile;        /* share of `pairs` that belongs to queries solved by tile sweeps */
    uint64_t pairs_direct;      /* share of `pairs` that belongs to queries solved by the direct scan */
    uint32_t n_outliers;        /* rows the index set aside as norm outliers (excluded from the approximate sweeps, scored exactly in K5) */
    float margin_audit;         /* HVS_FLAG_MARGIN_AUDIT: max over re-ranked survivors of |s~ + ||q||^2 - d_ref| / eps  (0 when off) */
} hvs_stats;

HVS_API uint32_t hvs_abi_version(void);

/* Message of the last error on this engine (or, with e == NULL, of the last failed hvs_create
 * on this thread).  Never NULL. */
HVS_API const char *hvs_last_error(const hvs_engine *e);

HVS_API int hvs_create(hvs_engine **out, const hvs_config *cfg);
HVS_API void hvs_destroy(hvs_engine *e);
/* Kernel family of the NEXT solves (hvs_mode); the index serves every mode, so one engine can answer the same
 * batch through two independent kernel families (what the parity checks of bench.py do). */
HVS_API int hvs_set_mode(hvs_engine *e, uint32_t mode);

/*
 * Indexing phase.  Replaces the per-query O(N) predicate scans of include/baseline.hpp:107-136 /
 * include/optimized.hpp:84-117: D is radix-sorted into a T-ordered arena and a (C,T)-ordered arena
 * (vectors re-laid out as 400-byte rows + ids + squared norms + sorted keys) so that every
 * predicate becomes one contiguous slice found by binary search.  Never sees query vectors
 * (contest rule, README.md:68).  `sample_proportion` is the reference's third vec_query argument:
 * sn = uint32_t(sample_proportion * n) rows are indexed (include/baseline.hpp:74) while the pad rule
 * keeps using n (include/baseline.hpp:138-147).  rows: n x 102 floats, row-major.
 */
HVS_API int hvs_index_build(hvs_engine *e, const float *rows_host, uint32_t n, float sample_proportion);
HVS_API int hvs_index_build_device(hvs_engine *e, const float *rows_dev, uint32_t n, float sample_proportion);
/*
 * The same from the layouts the reference's loader produces, without an intermediate copy.  D is streamed to the
 * device through two pinned staging buffers filled by host threads while the previous chunk is in flight.
 *   _rows      : n row pointers, each to 102 floats -- std::vector<std::vector<float>> as ReadBin builds it
 *                (include/io.h:123-133: one heap block per row) and as vec_query receives it (src/test.cpp:85)
 *   _from_file : the D file itself (uint32 N, then N x 102 float32; README.md:32-44, include/io.h:111-136),
 *                read with pread into the pinned buffers; *out_n (may be NULL) receives N
 */
HVS_API int hvs_index_build_rows(hvs_engine *e, const float *const *row_ptrs, uint32_t n, float sample_proportion);
HVS_API int hvs_index_build_from_file(hvs_engine *e, const char *path, float sample_proportion, uint32_t *out_n);

/*
 * The solve step: replaces the body of vec_query (include/baseline.hpp:88-177).  queries: m x 104
 * floats; out_ids: m x 100 uint32, row i = the ids the reference's knn_results[i] would hold.
 * Host variant copies queries in and ids out (pageable or pinned memory); device variant takes
 * device pointers and leaves the result on the device.
 */
HVS_API int hvs_solve(hvs_engine *e, const float *queries_host, uint32_t m, uint32_t *out_ids_host);
HVS_API int hvs_solve_device(hvs_engine *e, const float *queries_dev, uint32_t m, uint32_t *out_ids_dev);

/*
 * Query-sharded solve over `world` engines, one per GPU, every one holding the whole index (SURVEY 8e primary
 * variant; the reference's parallel variant partitions the work of one vec_query call across its thread pool,
 * include/optimized_parallel.hpp:100-157, include/threading.hpp:116-118 -- there the rows of D per query, here the
 * queries: D is 4 GB and fits every GPU).  Every rank passes the SAME m queries.  The engine resolves all m
 * predicates, assigns each query to a rank -- balanced by the rows the queries sweep, keeping queries that share rows
 * on one rank (hvs_shard_assign_host is the same pure function, exposed for tests) -- and solves the queries of
 * `rank`.  No communication happens here: the caller combines the ranks' rows (Python: sharding.solve_sharded, one
 * NCCL all-gather).
 *   out_order_host  : m query indices, rank-major; rank r owns out_order[sum(counts[0..r)) ...][counts[r]]
 *                     (may be NULL: the engine keeps the assignment on the device for hvs_shard_scatter_device)
 *   out_counts_host : world entries
 *   out_ids_dev     : counts[rank] x 100 uint32 -- row i = the answer of query out_order[offset(rank) + i]  (size it m x 100)
 */
HVS_API int hvs_solve_shard_device(hvs_engine *e, const float *queries_dev, uint32_t m, uint32_t rank, uint32_t world,
                                   uint32_t *out_ids_dev, uint32_t *out_order_host, uint32_t *out_counts_host);
/*
 * After the all-gather: rank r's rows arrive at gathered[r * cap ...][counts[r]] (cap = rows every rank contributed,
 * >= max(counts)); this places every row at its query's position: out_ids[order[p]] = row p of the rank-major sequence.
 * Uses the assignment of this engine's last hvs_solve_shard_device (kept on the device).  Device pointers;
 * out_ids_dev: m x 100.  Enqueued on the engine's stream, returns without synchronising.
 */
HVS_API int hvs_shard_scatter_device(hvs_engine *e, const uint32_t *gathered_dev, uint32_t cap, uint32_t *out_ids_dev);
/* The assignment alone, from slices (arena 0 = T-ordered, 1 = (C,T)-ordered; rows [begin,end)); CPU only. */
HVS_API int hvs_shard_assign_host(const uint32_t *arena, const uint32_t *begin, const uint32_t *end, uint32_t m,
                                  uint32_t world, uint32_t *out_order, uint32_t *out_counts);

/*
 * Data-sharded variant (mirrors include/optimized_parallel.hpp:100-157 at GPU scale):
 * each shard indexes its own rows (config.id_offset = first global row) and returns, per query,
 * its local best <= 100 (distance, global id) pairs ascending WITHOUT applying the pad rule, plus
 * the local match count.  Unused slots hold distance +inf and id 0xFFFFFFFF.  Device pointers.
 */
HVS_API int hvs_solve_partial_device(hvs_engine *e, const float *queries_dev, uint32_t m,
                             float *out_dist_dev /* m x 100 */, uint32_t *out_ids_dev /* m x 100 */,
                             uint32_t *out_count_dev /* m */);
/*
 * K5 merge (include/optimized_impl.h:337-385 Knn::merge + the pad rule): fold `g` partial lists per
 * query (as gathered from g shards, shard-major: [g][m][100]) into the final m x 100 ids.  The pad
 * rule is applied once, globally: `tail_rows_dev` = the last 100 rows of the GLOBAL data set
 * (100 x 102 floats, global rows n_total-100 .. n_total-1).  All device pointers.
 */
HVS_API int hvs_merge_partials_device(hvs_engine *e, const float *queries_dev, uint32_t m, uint32_t g,
                              const float *dist_dev, const uint32_t *ids_dev, const uint32_t *count_dev,
                              const float *tail_rows_dev, uint32_t n_total, uint32_t *out_ids_dev);

/*
 * include/io.h:50-78 (SaveKNNFull) on the device: sequential fp32 distance of every returned id
 * (what src/compare_data.cpp compares).  ids/out_dist: m x 100, host pointers.
 */
HVS_API int hvs_rescore(hvs_engine *e, const float *queries_host, uint32_t m, const uint32_t *ids_host, float *out_dist_host);

/*
 * Solve + the `.dist` side file in one call (src/test.cpp:85 followed by src/test.cpp:97-110, which copies the
 * 100 result rows of every query back into nested vectors only to hand them to SaveKNNFull): ids as hvs_solve,
 * out_dist_host[i][k] = the reference's calc_dist (include/io.h:38-48) between query i and row out_ids[i][k],
 * computed on the device from the rows already resident there.  Host pointers; m x 100 each.
 */
HVS_API int hvs_solve_full(hvs_engine *e, const float *queries_host, uint32_t m, uint32_t *out_ids_host, float *out_dist_host);

HVS_API int hvs_get_stats(const hvs_engine *e, hvs_stats *out);

/* Runs `iters` launches of an FFMA-only microkernel on the engine's device and returns the best
 * achieved FP32 TFLOP/s (the denominator of the FFMA roofline; MEASURED_PEAKS.json has none).
 * *out_sm_mhz (may be NULL) receives the SM clock that rate implies (TFLOP/s / (SMs x 128 lanes x 2)). */
HVS_API int hvs_measure_ffma_peak(hvs_engine *e, uint32_t iters, float *out_tflops, float *out_sm_mhz);

/*
 * Host planner, CPU only (no device needed; used by the CPU test-suite): given per-query slices
 * (arena 0 = T-ordered, 1 = (C,T)-ordered; [begin,end) rows) decide which queries go to the direct
 * scan and build the tile work items.  Returns the number of items, fills out_kind[m]
 * (0 = direct, 1 = tile) and, if non-NULL, up to max_items item descriptors of 4 uint32 each:
 * {arena, row_begin, row_end, n_queries}.
 */
HVS_API int hvs_plan_dryrun(const uint32_t *arena, const uint32_t *begin, const uint32_t *end, uint32_t m,
                    uint32_t mode, uint8_t *out_kind, uint32_t *out_items, uint32_t max_items,
                    uint64_t *out_pairs_computed);

#ifdef __cplusplus
}
#endif
#endif /* HVS_H_ */
