// hvs_engine.h -- internal (non-ABI) declarations shared by the engine's translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/hvs.h"
#include "hvs_common.cuh"

namespace hvs {

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostPinned {
    void *p = nullptr;
    size_t cap = 0;
    HostPinned() = default;
    HostPinned(const HostPinned &) = delete;
    HostPinned &operator=(const HostPinned &) = delete;
    ~HostPinned() { release(); }
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// The index: two re-ordered copies of D.
struct Index {
    uint32_t n_total = 0;      // rows given (n of baseline.hpp:71)
    uint32_t n = 0;            // rows indexed (sn of baseline.hpp:74)
    uint32_t id_offset = 0;
    DevBuf x[2];               // [n][100] fp32
    DevBuf ids[2];             // [n] u32
    DevBuf xnorm[2];           // [n] f32
    DevBuf xb[2];              // fp16 image for the tensor path, canonical K-major core-matrix layout, 224 B per row
    DevBuf keys_t;             // [n] u32  sorted ord(T)
    DevBuf keys_ct;            // [n] u64  sorted ord(C)<<32 | ord(T)
    DevBuf tail;               // [100][100] fp32: tail[s-1] = vector of row n_total - s  (pad rule)
    DevBuf inv_t;              // [n_total] u32: original row -> arena-T row (0xFFFFFFFF if not indexed); rescore only
    DevBuf outl[2];            // [n_outl] u32: arena positions (ascending) of the outlier rows -- excluded from K2/K3, scored exactly by K5
    uint32_t n_outl[2] = {0, 0};
    bool approx_ok = true;     // false: too many non-finite rows; no approximate sweep may run (everything takes K4)
    float xnorm_max = 0.f;     // max ||x||^2 over indexed non-outlier rows (margin bound)
    float img_scale = 1.f;     // sx: power of two applied to x before the fp16 image is written (K3)
    bool built = false;
    Arena arena(int a) const
    {
        Arena r;
        r.x = x[a].as<float>(); r.ids = ids[a].as<uint32_t>(); r.xnorm = xnorm[a].as<float>();
        r.xb = xb[a].p;
        r.outl = outl[a].as<uint32_t>(); r.n_outl = n_outl[a];
        return r;
    }
};

// ---- tile work items (K2 / K3) ----------------------------------------------------------------
constexpr int QT = 128;          // queries per FFMA tile item (K2)
constexpr int QT_TENSOR = 256;   // queries per tensor tile item (K3): two M=128 halves share every data stage
constexpr int KOUT = 256;        // candidates an item hands to finalize per query (<= this many)
#ifndef HVS_POOL
#define HVS_POOL 512
#endif
constexpr int TENSOR_POOL = HVS_POOL; // K3: survivor pool entries per (CTA, query) in global memory
#ifndef HVS_GBEST
#define HVS_GBEST 128
#endif
constexpr int TENSOR_GBEST = HVS_GBEST;
constexpr uint32_t SMALL_MAX = 255;  // K4s: slices of at most this many rows get a warp each (8 rounds of 32 rows; must stay below PlanParams::min_tile_len:
                                     // never tile queries).  Longer sparse slices take the CTA-per-query scan, which has the lower latency per query.
constexpr int OUTLIER_MAX = 256;  // K0: most rows that may be set aside as norm outliers // K3: per-query global list of the best scores seen by any CTA

struct TileItem {              // one CTA-sized unit of work: <= 128 queries sweep arena rows [row_begin,row_end)
    uint32_t arena;
    uint32_t row_begin;
    uint32_t row_end;
    uint32_t nq;               // 1..128 (FFMA) or 1..256 (tensor)
    uint32_t q_off;            // offset into the item-query list (item_q[q_off .. q_off+nq))
    uint32_t out_off;          // first candidate-list index of this item (list = out_off + slot)
    uint32_t kind;             // 0 = FFMA, 1 = tensor
    uint32_t pad;
};

struct Plan {
    std::vector<uint32_t> direct_q;       // queries for the direct scan kernel
    std::vector<TileItem> items;          // sorted by decreasing cost
    std::vector<uint32_t> item_q;         // query index per (item, slot)
    std::vector<uint32_t> tile_q;         // queries that go through finalize
    std::vector<uint32_t> q_list_off;     // CSR over tile_q: lists of each tile query  [tile_q.size()+1]
    std::vector<uint32_t> q_lists;        // candidate-list indices
    uint64_t pairs = 0, pairs_computed = 0, pairs_tile = 0;
    uint32_t n_lists = 0;
    uint32_t n_ffma = 0, n_tensor = 0;
    uint32_t n_small = 0;                 // queries left to K4s (not in direct_q)
    uint64_t pairs_small = 0;
    // scratch kept between solves (fresh multi-megabyte vectors cost more in page faults than the planning)
    struct Local { std::vector<TileItem> items; std::vector<uint32_t> item_q, cstart, cur, fill; uint64_t pairs_computed = 0; };
    std::vector<Local> locals;
    std::vector<std::pair<uint64_t, uint32_t>> sort_keys[2];   // (begin << 32 | end, query) of the tile queries of an arena
    std::vector<uint32_t> order[2], nlist, qpos;
    std::vector<uint8_t> is_tile;
    // incremental planning (plan_begin / plan_group / plan_finish): the sweep of group g runs on the GPU while
    // the host builds group g + 1
    struct Task { uint32_t arena, c0, c1; uint64_t incid; };
    std::vector<Task> tasks;
    std::vector<uint32_t> group_end;      // tasks [group_end[g-1], group_end[g]) form group g
    uint32_t R = 0, BQ = 0;
    std::vector<uint32_t> bnd[2];         // chunk boundaries of each arena: chunk c = rows [bnd[c], bnd[c+1]); <= R rows each
    std::vector<uint32_t> bnd_lut[2];     // non-uniform boundaries only: chunk that holds row (cell << 12)
    bool uniform_chunks = true;           // boundaries are the multiples of R
    uint32_t chunk_of(uint32_t a, uint32_t row) const   // the chunk that holds `row`
    {
        if (uniform_chunks) return row / R;
        uint32_t c = bnd_lut[a][row >> 12];
        while (row >= bnd[a][c + 1]) ++c;
        return c;
    }
    bool tensor = false;
    uint64_t incid = 0;                   // (chunk, query) incidences == candidate lists of the whole solve
    unsigned nthreads = 1;
    void reset()
    {
        direct_q.clear(); items.clear(); item_q.clear(); tile_q.clear(); q_list_off.clear(); q_lists.clear();
        pairs = pairs_computed = pairs_tile = pairs_small = 0;
        n_lists = n_ffma = n_tensor = n_small = 0;
        tasks.clear(); group_end.clear(); incid = 0; R = 0;
    }
};

struct PlanParams {
    uint32_t mode = HVS_MODE_AUTO;
    uint32_t chunk_rows = 1u << 17;       // max rows an item sweeps
    uint32_t min_tile_len = 1024;         // shorter slices always take the direct scan
    uint32_t small_max = 0;               // slices of at most this many rows were already solved by K4s (launch_small): neither tile nor direct_q
    double direct_cost_ratio = 12.0;      // FFMA tile pair-slot vs direct pair throughput ratio
    double tensor_min_depth = 0.75;       // tensor items: average queries per row needed (the sweep is bandwidth-, not slot-priced)
    uint64_t min_tile_pairs = 4000000;    // below this many (query,row) pairs a tile sweep's fixed cost (~0.5 ms: persistent launch, operand
                                          // build, finalize) exceeds the direct scan of all of them: everything goes direct
    bool tensor_available = false;
    bool approx_ok = true;                // false: the index allows no approximate sweep (Index::approx_ok)
};

// shard_assign's cost of a query, in (query,row) pairs: its rows plus a fixed part -- threshold warm-up, list merges,
// finalize -- that grows with the slice (SHARD_ROW_COST per row) up to SHARD_QUERY_COST: a long slice is swept in many
// chunks at once and warms a threshold up in each of them, a category's few chunks do it once.  Measured (every rank's
// share of the headline batch, tools/shard_all_ranks.py): K3 time per rank = 1.4-2.0 ms per 10^10 pairs + 2.3 us per
// T-range query (13-16 x 10^6 pairs' worth) + 0.18 us per category query (10^5 / 3x10^4 rows: ~10^6); a flat 6x10^6 per
// query left the ranks at 5.30-6.05 ms.  HVS_SHARD_QCOST / HVS_SHARD_ROWCOST override (ROWCOST 0: flat).
constexpr uint64_t SHARD_QUERY_COST = 14000000;
constexpr uint64_t SHARD_ROW_COST = 14;
uint64_t shard_query_cost();
uint64_t shard_row_cost();
__host__ __device__ inline unsigned long long shard_cost_of(uint32_t len, unsigned long long qcost, unsigned long long rowcost)
{
    const unsigned long long rows = len > (uint32_t)K ? len : (uint32_t)K;
    const unsigned long long grow = rows * rowcost;
    return rows + ((rowcost && grow < qcost) ? grow : qcost);
}
constexpr uint32_t SHARD_STRIPES = 16;          // shard_assign: segments per rank (T arena)
constexpr uint32_t SHARD_STRIPES_CT = 2;        // ... in the (C,T) arena: a category's queries share its rows and fill items together -- few cuts
void shard_assign(const QSlice *slices, uint32_t m, uint32_t world, uint32_t *order, uint32_t *counts);
void plan_build(const QSlice *slices, uint32_t m, const PlanParams &pp, Plan &out);     // begin + every group + finish

// Chunk size (rows an item sweeps), a power of two: ~items_per_sm items per SM over the whole job for balance -- but an
// item pays a fixed price at both ends (operand build, pipeline fill, global-list merge, barrier: ~250 kcycles), so short
// items are avoided while at least a quarter of that item count remains: 32768 rows (256 stages) if possible.
// Measured on one rank's share of 8 (1.7e10 pair job): 16384-row chunks 7.08 ms, 32768 6.66, 65536 6.64.
__host__ __device__ inline uint32_t plan_chunk_rows(unsigned long long tile_qrows, uint32_t bq, uint32_t items_per_sm, uint32_t sm_count)
{
    auto pow2_below = [](unsigned long long want) { uint32_t r = 8192; while ((unsigned long long)r * 2 <= want && r < (1u << 22)) r *= 2; return r; };
    const uint32_t r_fine = pow2_below(tile_qrows / ((unsigned long long)bq * items_per_sm * sm_count));
    const uint32_t r_coarse = pow2_below(tile_qrows / ((unsigned long long)bq * (items_per_sm >= 4 ? items_per_sm / 4 : 1) * sm_count));
    const uint32_t floor_rows = r_coarse < 32768u ? r_coarse : 32768u;
    return r_fine > floor_rows ? r_fine : floor_rows;
}

// ---- the device planner (hvs_plan_dev.cu) -------------------------------------------------------------------------
enum : uint32_t { PD_TILE = 0, PD_DIRECT = 1, PD_SMALL = 2 };   // query classes, in sort order
struct PlanCfg {                          // by value to the planner kernels
    double need;                          // average queries per row a tile query's slice must see
    unsigned long long min_tile_pairs;    // tiny-job rule
    uint32_t small_max, min_tile_len;
    uint32_t tile_allowed, force_tile;    // mode != DIRECT && index allows approximate sweeps; mode == TENSOR
    uint32_t bq, items_per_sm, sm_count, kind;
    uint32_t seed_phase;                  // 1: items that hold queries beginning in their chunk are swept in a first launch
    uint32_t ct_min_rows;                 // smallest chunk of the (C,T) arena (its slices are short: cutting them finer only repeats threshold warm-ups)
};
struct PlanHeader {                       // lives on the device; the host reads it back once per solve (and n_redo at the end)
    unsigned long long pairs, tile_qrows, pairs_tile, pairs_computed, incid;
    uint32_t maxend[2], nchunk[2], chunk_base[2];
    uint32_t nchunk_total, R, n_tile, n_tile_arena0, n_direct, n_small, n_items, tiny, n_redo, Ra[2];   // Ra: chunk rows per arena (R = Ra[0])
    uint32_t n_seed_items;                // items [0, n_seed_items) hold the queries whose slices BEGIN in the item's chunk: swept first (seed phase)
};
struct PlanDev {
    DevBuf header, diff, pref, cls, keys_in, keys, vals_in, vals, nch, qoff, cdiff, cbeg, cstart, ibase, ibase_rest, nrest, sort_tmp;
    DevBuf sa_owner_in, sa_owner, sa_order, sa_counts;     // query sharding
    uint32_t nb = 0;
    uint32_t launches = 0;                                  // kernels launched by the last plan_dev_begin / shard_assign_dev
};
void plan_begin(const QSlice *slices, uint32_t m, const PlanParams &pp, Plan &out);     // classify, order, cut into groups
void plan_group(const QSlice *slices, Plan &out, size_t g, uint32_t &item_begin, uint32_t &item_end);   // items of one group
void plan_finish(const QSlice *slices, uint32_t m, Plan &out);                          // candidate-list CSR per query

}  // namespace hvs

struct hvs_engine {
    int device = 0;
    uint32_t mode = HVS_MODE_AUTO;
    uint32_t id_offset = 0;
    uint32_t flags = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t stream2 = nullptr;   // second lane for tile launches: the tail of group g overlaps the head of group g + 1
    cudaStream_t stream_up = nullptr; // work-list uploads of the groups: never queued behind a running sweep
    cudaEvent_t ev_sync[4]{};
    cudaEvent_t ev_up[8]{};           // upload of group g done
    uint32_t work_slot = 0;           // which item counter (d_work_counter) the next K3 launch draws from
    uint32_t pool_slot = 0;           // which half of the survivor pools the next K3 launch uses
    int sm_count = 148;
    std::string err;
    hvs::Index index;
    hvs_stats stats{};
    // per-solve scratch (grow-only)
    hvs::DevBuf d_queries, d_out, d_slices, d_direct_q, d_items, d_item_q, d_tile_q, d_qoff, d_qlists;
    hvs::DevBuf d_cand, d_cand_cnt, d_scratch, d_flags, d_gthr, d_pool, d_gbest, d_glock, d_work_counter, d_rescore_ids, d_rescore_out, d_audit, d_shard_q, d_shard_sl, d_shard_own, d_split, d_k1acc;
    std::vector<hvs::QSlice> h_shard_sl;
    std::vector<uint32_t> h_shard_order, h_shard_counts;
    hvs::HostPinned h_slices, h_flags, h_stage_own, h_stage, h_ingest[2];
    cudaEvent_t ev[12]{};
    cudaEvent_t evg[16]{};     // start/end of each group's tile launch
    uint32_t shard_m = 0, shard_world = 0;        // the last hvs_solve_shard_device: its assignment stays in pdev.sa_order / sa_counts
    hvs::Plan plan;
    hvs::PlanDev pdev;
    hvs::HostPinned h_header;
};

namespace hvs {
// each returns cudaSuccess or the failing error (message left in e->err)
cudaError_t index_build_device(hvs_engine *e, const float *rows_dev, uint32_t n_total, float sample_proportion);
cudaError_t launch_plan_search(hvs_engine *e, const float *queries_dev, uint32_t m, QSlice *slices_dev);   // also zeroes and fills e->d_k1acc: {pairs, small queries}
// q_sub[i] = queries[own[i]], sl_sub[i] = slices[own[i]]  (the queries a rank owns, made contiguous)
cudaError_t launch_gather_queries(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev, const uint32_t *own_dev,
                                  uint32_t n_own, float *q_sub, QSlice *sl_sub);
// K4: solve `nq` queries listed in q_list_dev (or all 0..nq-1 if null) by direct scan.
// partial == false: writes out_ids[q][100] with the pad rule applied.
// partial == true : writes out_dist/out_ids (ascending, unused = +inf/0xFFFFFFFF) and out_count, no pad.
cudaError_t launch_direct(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev,
                          const uint32_t *q_list_dev, uint32_t nq, bool partial,
                          uint32_t *out_ids, float *out_dist, uint32_t *out_count, uint32_t skip_max = 0);
// K4s: every query of q_list (or 0..nq-1) whose slice has at most small_max rows, one warp each; longer ones are skipped.
cudaError_t launch_small(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev, const uint32_t *q_list_dev,
                         uint32_t nq, uint32_t small_max, bool partial, uint32_t *out_ids, float *out_dist, uint32_t *out_count);
cudaError_t launch_tile_ffma(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev,
                             const TileItem *items_dev, uint32_t item_begin, uint32_t n_items,
                             const uint32_t *item_q_dev, uint64_t *cand_dev, uint32_t *cand_cnt_dev,
                             uint32_t *gthr_dev, uint32_t *flags_dev, float margin_scale);
cudaError_t launch_tile_tensor(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev,
                               const TileItem *items_dev, uint32_t item_begin, uint32_t n_items,
                               const uint32_t *item_q_dev, uint64_t *cand_dev, uint32_t *cand_cnt_dev,
                               uint32_t *gthr_dev, uint32_t *flags_dev);
cudaError_t tile_tensor_begin(hvs_engine *e);
cudaError_t launch_fill_u32(hvs_engine *e, uint32_t *dst, uint32_t value, size_t n);
cudaError_t launch_finalize(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev,
                            const uint32_t *tile_q_dev, uint32_t n_tile_q, const uint32_t *qoff_dev,
                            const uint32_t *qlists_dev, const uint64_t *cand_dev, const uint32_t *cand_cnt_dev,
                            uint32_t *flags_dev, bool partial, bool tensor_lists, uint32_t *out_ids, float *out_dist, uint32_t *out_count);
cudaError_t launch_merge_partials(hvs_engine *e, const float *queries_dev, uint32_t m, uint32_t g,
                                  const float *dist_dev, const uint32_t *ids_dev, const uint32_t *count_dev,
                                  const float *tail_rows_dev, uint32_t n_total, uint32_t *out_ids_dev);
cudaError_t launch_rescore(hvs_engine *e, const float *queries_dev, uint32_t m, const uint32_t *ids_dev, float *out_dev,
                           const float *rows_unused);
cudaError_t measure_ffma_peak(hvs_engine *e, uint32_t iters, float *tflops, float *mhz);
// hvs_sort.cu: stable LSD radix sort of (key, u32 value) pairs by key bits [0, end_bit); (k0, v0) = input AND scratch, the
// result lands in (k1, v1); tmp = radix_sort_temp_bytes(n) bytes
size_t radix_sort_temp_bytes(uint32_t n);
uint32_t radix_sort_launches(uint32_t n, int end_bit);
template <class KeyT>
cudaError_t radix_sort_pairs(KeyT *k0, uint32_t *v0, KeyT *k1, uint32_t *v1, uint32_t n, int end_bit, void *tmp, cudaStream_t st);
// device planner: everything up to the header read-back (one stream sync); then the chunk lists, items and K5's list index
cudaError_t plan_dev_begin(hvs_engine *e, const QSlice *d_sl, uint32_t m, const PlanCfg &cfg, PlanHeader *h_out);
cudaError_t plan_dev_fill(hvs_engine *e, const QSlice *d_sl, const PlanHeader &h, const PlanCfg &cfg, uint32_t *item_q_dev,
                          TileItem *items_dev, uint32_t *qlists_dev);
cudaError_t plan_dev_redo(hvs_engine *e, const uint32_t *flags_dev, uint32_t n_tile, uint32_t *redo_dev);
cudaError_t shard_assign_dev(hvs_engine *e, const QSlice *d_sl, uint32_t m, uint32_t world, uint32_t stripes,
                             uint32_t **order_dev, uint32_t **counts_dev);
uint32_t shard_stripes(uint32_t m, uint32_t world);
cudaError_t launch_shard_scatter(hvs_engine *e, const uint32_t *gathered_dev, uint32_t cap, const uint32_t *order_dev,
                                 const uint32_t *counts_dev, uint32_t m, uint32_t world, uint32_t *out_ids_dev);
cudaError_t direct_init_attributes();
cudaError_t tile_ffma_init_attributes();
cudaError_t tile_tensor_init_attributes();
inline cudaError_t init_kernel_attributes()
{
    cudaError_t c = direct_init_attributes();
    if (c == cudaSuccess) c = tile_ffma_init_attributes();
    if (c == cudaSuccess) c = tile_tensor_init_attributes();
    return c;
}
}  // namespace hvs
