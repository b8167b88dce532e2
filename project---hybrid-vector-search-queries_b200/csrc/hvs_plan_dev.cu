// hvs_plan_dev.cu -- K1 on the device: the query planner as a handful of small kernels.
//
// Same job as the host planner of hvs_plan.cu (which stays as the CPU-testable statement of the rules and serves
// hvs_plan_dryrun): decide which queries sweep arena rows together in tile items and which are scanned alone, cut the
// tile queries' slices at chunk boundaries, batch the queries of every chunk, and hand K2/K3 their work items and K5
// its candidate-list index -- but without ever bringing the slices to the host.  The host reads back ONE 128-byte
// header (counts, so that it can size buffers and grids) and launches the rest.  Reference: the per-query decode and
// predicate scans this replaces are include/baseline.hpp:90-136; the reference has no planner (every query is a full
// scan of D).
//
//   k_pd_depth     queries -> depth histogram per arena over 1024-row cells (two atomics per query)
//   k_pd_cells     one CTA per arena: prefix sums (depth, cumulative depth)
//   k_pd_classify  per query: average depth over its slice >= need -> tile query; pair totals
//   k_pd_params    one thread: tiny-job rule, chunk size R
//   k_pd_keys      sort key per query: class | arena | begin | end   (tile queries first, then CTA-scan queries)
//   (radix sort of m 64-bit keys: hvs_sort.cu, the sort of K0)
//   k_pd_chunks    tile queries -> chunk occupancy (difference array), nchunks per query
//   k_pd_scan      one CTA: occupancy -> per-chunk list offsets and item offsets; per-query candidate-list offsets
//   [host: header D2H]
//   k_pd_fill      one CTA per chunk: the chunk's queries in (begin, end) order -> item_q, items, candidate-list CSR
//
// Item order: first the items in which slices BEGIN (their thresholds start cold: the slow items -- nearly all of the
// (C,T) arena's, the T arena's first chunk, and the last item or two of every other chunk's list), then the items made
// only of queries begun in earlier chunks, in chunk order ((C,T) arena, then T).  One persistent launch sweeps them all.
#include "hvs_engine.h"

namespace hvs {

namespace {
constexpr uint32_t CELL = 1024;        // depth histogram granularity (rows)
constexpr int SCAN_T = 1024;           // threads of the single-CTA scans

// block-wide exclusive scan of one value per thread (SCAN_T threads); returns the exclusive prefix, total in *total
template <class T>
__device__ __forceinline__ T block_excl_scan(T v, T *total, T *smem /* >= 33 entries */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const T y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        T w = lane < (int)(blockDim.x >> 5) ? smem[lane] : (T)0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const T y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        smem[lane] = w;                                 // inclusive over warps
    }
    __syncthreads();
    const T base = warp ? smem[warp - 1] : (T)0;
    if (total) *total = smem[(blockDim.x >> 5) - 1];
    __syncthreads();
    return base + x - v;
}
}  // namespace

// ---- classification ------------------------------------------------------------------------------------------------
__global__ void k_pd_depth(const QSlice *__restrict__ sl, uint32_t m, int *__restrict__ diff0, int *__restrict__ diff1,
                           PlanHeader *__restrict__ H)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const QSlice s = sl[i];
    if (s.end > s.begin) {
        int *d = s.arena == ARENA_T ? diff0 : diff1;
        atomicAdd(&d[s.begin / CELL], 1);
        atomicAdd(&d[(s.end - 1) / CELL + 1], -1);
        atomicMax(&H->maxend[s.arena & 1u], s.end);
    }
}

// depth[c] = queries covering cell c; pref[c] = sum of depth over cells < c.  One CTA per arena.
__global__ void __launch_bounds__(SCAN_T) k_pd_cells(const int *__restrict__ diff0, const int *__restrict__ diff1,
                                                     long long *__restrict__ pref0, long long *__restrict__ pref1, uint32_t ncell)
{
    __shared__ long long sm[33];
    const int *diff = blockIdx.x == 0 ? diff0 : diff1;
    long long *pref = blockIdx.x == 0 ? pref0 : pref1;
    // two passes: running depth, then its prefix.  Each thread owns a contiguous run of cells.
    const uint32_t per = (ncell + SCAN_T - 1) / SCAN_T;
    const uint32_t c0 = min(ncell, threadIdx.x * per), c1 = min(ncell, c0 + per);
    long long s = 0;
    for (uint32_t c = c0; c < c1; ++c) s += diff[c];
    long long tot;
    const long long dbase = block_excl_scan<long long>(s, &tot, sm);       // depth just before my first cell
    long long run = dbase, acc = 0;
    for (uint32_t c = c0; c < c1; ++c) { run += diff[c]; acc += run; }
    const long long pbase = block_excl_scan<long long>(acc, &tot, sm);
    run = dbase; acc = pbase;
    for (uint32_t c = c0; c < c1; ++c) { pref[c] = acc; run += diff[c]; acc += run; }
    if (c1 == ncell) pref[ncell] = acc;     // the owner of the last run and the idle threads behind it all hold the grand total here
}

__global__ void k_pd_classify(const QSlice *__restrict__ sl, uint32_t m, const long long *__restrict__ pref0,
                              const long long *__restrict__ pref1, PlanCfg cfg, uint8_t *__restrict__ cls, PlanHeader *__restrict__ H)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long pairs = 0, tile_rows = 0, tile_pairs = 0;
    uint32_t small = 0;
    if (i < m) {
        const QSlice s = sl[i];
        const uint32_t len = s.end - s.begin;
        pairs = len > (uint32_t)K ? len : (uint32_t)K;
        uint8_t c = PD_DIRECT;
        if (len <= cfg.small_max && cfg.small_max) { c = PD_SMALL; small = 1; }
        else if (cfg.tile_allowed && len >= cfg.min_tile_len) {
            const long long *pref = s.arena == ARENA_T ? pref0 : pref1;
            const uint32_t c0 = s.begin / CELL, c1 = (s.end - 1) / CELL + 1;
            const double avg = (double)(pref[c1] - pref[c0]) / (double)(c1 - c0);
            if (avg >= cfg.need) { c = PD_TILE; tile_rows = len; tile_pairs = pairs; }
        }
        cls[i] = c;
    }
    // block totals -> header
    __shared__ unsigned long long sp, sr, st;
    __shared__ uint32_t ss;
    if (threadIdx.x == 0) { sp = sr = st = 0; ss = 0; }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
        tile_rows += __shfl_xor_sync(0xffffffffu, tile_rows, o);
        tile_pairs += __shfl_xor_sync(0xffffffffu, tile_pairs, o);
        small += __shfl_xor_sync(0xffffffffu, small, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sp, pairs); atomicAdd(&sr, tile_rows); atomicAdd(&st, tile_pairs); atomicAdd(&ss, small); }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&H->pairs, sp); atomicAdd(&H->tile_qrows, sr); atomicAdd(&H->pairs_tile, st); atomicAdd(&H->n_small, ss);
    }
}

// tiny-job rule and chunk size: one thread
__global__ void k_pd_params(PlanCfg cfg, PlanHeader *__restrict__ H)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long rows = H->tile_qrows;
    if (rows && rows < cfg.min_tile_pairs && !cfg.force_tile) { rows = 0; H->tile_qrows = 0; H->pairs_tile = 0; H->tiny = 1; }
    uint32_t R = 8192;
    if (rows) R = plan_chunk_rows(rows, cfg.bq, cfg.items_per_sm, cfg.sm_count);
    H->R = R;
    H->Ra[0] = R;
    H->Ra[1] = R < cfg.ct_min_rows ? cfg.ct_min_rows : R;
    uint32_t base = 0;
    // item order: the (C,T) arena's chunks first, then the T arena's
    for (int a = 1; a >= 0; --a) {
        const uint32_t Rr = H->Ra[a];
        const uint32_t nc = rows && H->maxend[a] ? (H->maxend[a] + Rr - 1) / Rr : 0;
        H->nchunk[a] = nc;
        H->chunk_base[a] = base;
        base += nc;
    }
    H->nchunk_total = base;
}

// sort keys: [class:2][arena:1][begin:nb][end:nb+1]; tile queries first (by arena, begin, end), then the CTA-scan
// queries (same order: neighbours share rows in L2), then everything else
__global__ void k_pd_keys(const QSlice *__restrict__ sl, uint32_t m, const uint8_t *__restrict__ cls, uint32_t nb,
                          const PlanHeader *__restrict__ H, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const QSlice s = sl[i];
    uint32_t c = cls[i];
    if (c == PD_TILE && H->tiny) c = PD_DIRECT;
    const uint64_t k = ((uint64_t)c << (2 * nb + 2)) | ((uint64_t)(s.arena & 1u) << (2 * nb + 1)) | ((uint64_t)s.begin << (nb + 1)) | (uint64_t)s.end;
    keys[i] = k;
    vals[i] = i;
}

// After the sort: class counts, the chunk occupancy of the tile queries, chunks per tile query.
// sorted position t < n_tile  <=>  tile query.
__global__ void k_pd_chunks(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t m, uint32_t nb,
                            const QSlice *__restrict__ sl, PlanHeader *__restrict__ H, int *__restrict__ cdiff,
                            uint32_t *__restrict__ cbeg /* queries whose slice begins in the chunk */,
                            uint32_t *__restrict__ nch /* [m] chunks of the tile query at sorted position t */)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const uint32_t c = (uint32_t)(keys[t] >> (2 * nb + 2));
    uint32_t n = 0;
    if (c == PD_TILE) {
        const QSlice s = sl[vals[t]];
        const uint32_t a = s.arena & 1u, R = H->Ra[a];
        const uint32_t lo = s.begin / R, hi = (s.end - 1) / R;
        n = hi - lo + 1;
        int *d = cdiff + H->chunk_base[a];
        atomicAdd(&d[lo], 1);
        if (hi + 1 < H->nchunk[a]) atomicAdd(&d[hi + 1], -1);
        atomicAdd(&cbeg[H->chunk_base[a] + lo], 1u);
        atomicAdd(&H->n_tile, 1u);
        if (a == ARENA_T) atomicAdd(&H->n_tile_arena0, 1u);
    } else if (c == PD_DIRECT) {
        atomicAdd(&H->n_direct, 1u);
    }
    nch[t] = n;
}

// One CTA: chunk occupancy -> list offsets (cstart) and item offsets per chunk; candidate-list offsets per tile query.
__global__ void __launch_bounds__(SCAN_T) k_pd_scan(PlanHeader *__restrict__ H, const int *__restrict__ cdiff, uint32_t bq,
                                                    const uint32_t *__restrict__ cbeg, uint32_t seed_phase,
                                                    uint32_t *__restrict__ cstart /* [nchunk_total+1] */,
                                                    uint32_t *__restrict__ ibase /* [nchunk_total]: first SEED item of the chunk */,
                                                    uint32_t *__restrict__ ibase_rest /* [nchunk_total]: first other item, relative to the end of the seed items */,
                                                    uint32_t *__restrict__ nrest /* [nchunk_total]: items of the chunk that hold only queries begun earlier */,
                                                    const uint32_t *__restrict__ nch, uint32_t *__restrict__ qoff /* [n_tile+1] */)
{
    __shared__ unsigned long long sm64[33];
    __shared__ uint32_t sm32[33];
    // (a) per arena: running occupancy (the difference array restarts at every arena's first chunk)
    // The chunk's list is in (begin, end) order: the queries that began in an earlier chunk come first.  The items made
    // only of those ("rest") are swept in the second launch; the items that hold a query BEGINNING in this chunk ("seed":
    // its threshold starts cold) go first, so that no chunk of a slice is swept before the slice's first chunk has
    // produced a threshold (seed_phase == 0: every item counts as seed, one launch).
    unsigned long long incid_total = 0;
    uint32_t seed_total = 0, rest_total = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int a = pass == 0 ? 1 : 0;                         // same order as chunk_base
        const uint32_t nc = H->nchunk[a], base = H->chunk_base[a];
        const uint32_t per = (nc + SCAN_T - 1) / SCAN_T;
        const uint32_t c0 = min(nc, threadIdx.x * per), c1 = min(nc, c0 + per);
        int s = 0;
        for (uint32_t c = c0; c < c1; ++c) s += cdiff[base + c];
        uint32_t tot32;
        const uint32_t occ0 = block_excl_scan<uint32_t>((uint32_t)s, &tot32, sm32);   // occupancy just before my first chunk (mod 2^32 sums of +-1)
        unsigned long long li = 0;
        uint32_t its = 0, itr = 0, occ = occ0;
        for (uint32_t c = c0; c < c1; ++c) {
            occ += (uint32_t)cdiff[base + c];
            const uint32_t nit = (occ + bq - 1) / bq, nr = seed_phase ? (occ - min(occ, cbeg[base + c])) / bq : 0u;
            li += occ; its += nit - nr; itr += nr;
        }
        unsigned long long tot64;
        const unsigned long long lbase = block_excl_scan<unsigned long long>(li, &tot64, sm64) + incid_total;
        uint32_t stot, rtot;
        const uint32_t sb = block_excl_scan<uint32_t>(its, &stot, sm32) + seed_total;
        const uint32_t rb = block_excl_scan<uint32_t>(itr, &rtot, sm32) + rest_total;
        unsigned long long l = lbase;
        uint32_t is = sb, ir = rb;
        occ = occ0;
        for (uint32_t c = c0; c < c1; ++c) {
            occ += (uint32_t)cdiff[base + c];
            const uint32_t nit = (occ + bq - 1) / bq, nr = seed_phase ? (occ - min(occ, cbeg[base + c])) / bq : 0u;
            cstart[base + c] = (uint32_t)l; ibase[base + c] = is; ibase_rest[base + c] = ir; nrest[base + c] = nr;
            l += occ; is += nit - nr; ir += nr;
        }
        incid_total += tot64;
        seed_total += stot;
        rest_total += rtot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        cstart[H->nchunk_total] = (uint32_t)incid_total;
        H->incid = incid_total;
        H->n_items = seed_total + rest_total;
        H->n_seed_items = seed_total;
    }
    // (b) candidate-list offsets of the tile queries (sorted positions 0..n_tile)
    const uint32_t nt = H->n_tile;
    const uint32_t per = (nt + SCAN_T - 1) / SCAN_T;
    const uint32_t t0 = min(nt, threadIdx.x * per), t1 = min(nt, t0 + per);
    uint32_t s = 0;
    for (uint32_t t = t0; t < t1; ++t) s += nch[t];
    uint32_t tot;
    uint32_t off = block_excl_scan<uint32_t>(s, &tot, sm32);
    for (uint32_t t = t0; t < t1; ++t) { qoff[t] = off; off += nch[t]; }
    if (threadIdx.x == 0) qoff[nt] = tot;
}

// One CTA per chunk: walk the arena's tile queries in sorted order, keep those that overlap the chunk -> the chunk's
// query list (item_q), in order; every BQ of them form an item.  Also the per-query candidate-list index for K5:
// list id = position in item_q.
__global__ void __launch_bounds__(256) k_pd_fill(const PlanHeader *__restrict__ H, const uint32_t *__restrict__ vals,
                                                 const QSlice *__restrict__ sl, uint32_t bq, uint32_t kind,
                                                 const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ ibase,
                                                 const uint32_t *__restrict__ ibase_rest, const uint32_t *__restrict__ nrest,
                                                 const uint32_t *__restrict__ qoff, uint32_t *__restrict__ item_q,
                                                 TileItem *__restrict__ items, uint32_t *__restrict__ qlists,
                                                 unsigned long long *__restrict__ pairs_computed)
{
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t s_run;
    const uint32_t gc = blockIdx.x;                             // global chunk index (arena CT first)
    if (gc >= H->nchunk_total) return;
    const uint32_t a = gc >= H->chunk_base[0] && H->nchunk[0] && gc < H->chunk_base[0] + H->nchunk[0] ? 0u : 1u;
    const uint32_t c = gc - H->chunk_base[a];
    const uint32_t R = H->Ra[a];
    const uint64_t r0 = (uint64_t)c * R, r1 = r0 + R;
    // sorted tile queries of this arena: arena 0 occupies sorted positions [0, n0), arena 1 [n0, n_tile)
    const uint32_t n0 = H->n_tile_arena0, nt = H->n_tile;
    const uint32_t tb = a == 0 ? 0u : n0, te = a == 0 ? n0 : nt;
    const uint32_t base = cstart[gc];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (uint32_t t0 = tb; t0 < te; t0 += 256) {
        const uint32_t t = t0 + threadIdx.x;
        bool in = false;
        uint32_t q = 0, lo = 0;
        if (t < te) {
            q = vals[t];
            const QSlice s = sl[q];
            in = (uint64_t)s.begin < r1 && (uint64_t)s.end > r0;
            lo = s.begin / R;
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, in);
        if (lane == 0) wsum[warp] = __popc(mask);
        __syncthreads();
        uint32_t before = s_run;
        for (int w = 0; w < warp; ++w) before += wsum[w];
        if (in) {
            const uint32_t p = base + before + __popc(mask & ((1u << lane) - 1u));
            item_q[p] = q;
            qlists[qoff[t] + (c - lo)] = p;
        }
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t s = 0; for (int w = 0; w < 8; ++w) s += wsum[w]; s_run += s; }
        __syncthreads();
        // sorted by begin: once a whole tile of queries begins at or after the chunk's end, nothing further overlaps
        if (t0 + 255 < te) {
            const QSlice last = sl[vals[t0 + 255]];
            if ((uint64_t)last.begin >= r1) break;              // block-uniform: every thread reads the same entry
        }
    }
    __syncthreads();
    const uint32_t count = cstart[gc + 1] - base;
    const uint32_t nit = (count + bq - 1) / bq;
    // items of this chunk: one warp per item computes the union of its queries' rows inside the chunk
    for (uint32_t j = warp; j < nit; j += 8) {
        const uint32_t q0 = base + j * bq, nq = min(bq, count - j * bq);
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (uint32_t k = lane; k < nq; k += 32) {
            const QSlice s = sl[item_q[q0 + k]];
            lo = min(lo, (uint32_t)max((uint64_t)s.begin, r0));
            hi = max(hi, (uint32_t)min((uint64_t)s.end, r1));
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) {
            TileItem it;
            it.arena = a; it.row_begin = lo; it.row_end = hi; it.nq = nq; it.q_off = q0; it.out_off = q0; it.kind = kind; it.pad = 0;
            // the first nrest items of the list hold only queries begun in earlier chunks: second launch, behind all seed items
            const uint32_t nr = nrest[gc];
            items[j < nr ? H->n_seed_items + ibase_rest[gc] + j : ibase[gc] + (j - nr)] = it;
            atomicAdd(pairs_computed, (unsigned long long)(hi - lo) * nq);
        }
    }
}

// queries whose candidate lists overflowed their margin guarantee (flags) -> a list for the exact re-solve
__global__ void k_pd_redo(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ tile_q, uint32_t n_tile,
                          uint32_t *__restrict__ redo, PlanHeader *__restrict__ H)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tile) return;
    const uint32_t q = tile_q[t];
    if (flags[q]) redo[atomicAdd(&H->n_redo, 1u)] = q;
}

// ---- query sharding on the device (same rule as shard_assign in hvs_plan.cu) ---------------------------------------------
__global__ void k_sa_keys(const QSlice *__restrict__ sl, uint32_t m, uint32_t nb, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const QSlice s = sl[i];
    keys[i] = ((uint64_t)(s.arena & 1u) << (2 * nb + 1)) | ((uint64_t)s.begin << (nb + 1)) | (uint64_t)s.end;
    vals[i] = i;
}

// One CTA: cumulative cost along the (arena, begin, end) order -> segment -> owner rank; queries per rank.
__global__ void __launch_bounds__(SCAN_T) k_sa_assign(const QSlice *__restrict__ sl, const uint32_t *__restrict__ vals, uint32_t m,
                                                      uint32_t world, uint32_t stripes, unsigned long long query_cost, unsigned long long row_cost,
                                                      uint32_t *__restrict__ owner /* [m], by sorted position */,
                                                      uint32_t *__restrict__ counts /* [world] */)
{
    __shared__ unsigned long long sm64[33];
    __shared__ uint32_t scnt[256];
    if (threadIdx.x < 256) scnt[threadIdx.x] = 0;
    const uint32_t per = (m + SCAN_T - 1) / SCAN_T;
    const uint32_t t0 = min(m, threadIdx.x * per), t1 = min(m, t0 + per);
    auto cost_of = [&](uint32_t t, uint32_t *arena) {
        const QSlice s = sl[vals[t]];
        *arena = s.arena & 1u;
        return shard_cost_of(s.end - s.begin, query_cost, row_cost);
    };
    // the sorted order holds the T arena's queries, then the (C,T) arena's: each arena's cost is cut into segments of
    // its own, so every rank gets its share of BOTH (a category query costs a tenth of a range query)
    unsigned long long mine = 0, mine0 = 0;
    for (uint32_t t = t0; t < t1; ++t) {
        uint32_t a;
        const unsigned long long c = cost_of(t, &a);
        mine += c;
        if (a == 0) mine0 += c;
    }
    unsigned long long total, total0;
    unsigned long long cum = block_excl_scan<unsigned long long>(mine, &total, sm64);      // has the barriers scnt needs
    block_excl_scan<unsigned long long>(mine0, &total0, sm64);
    for (uint32_t t = t0; t < t1; ++t) {
        uint32_t a;
        const unsigned long long c = cost_of(t, &a);
        const unsigned long long nseg = (unsigned long long)world * (a ? min(stripes, SHARD_STRIPES_CT) : stripes);
        const unsigned long long cum_a = a ? cum - total0 : cum, total_a = a ? total - total0 : total0;
        unsigned long long seg = (cum_a + c / 2) * nseg / total_a;     // the segment that holds the midpoint of this query's cost interval
        if (seg >= nseg) seg = nseg - 1;
        const uint32_t o = (uint32_t)((seg + (a ? world / 2 : 0u)) % world);
        owner[t] = o;
        atomicAdd(&scnt[o], 1u);
        cum += c;
    }
    __syncthreads();
    if (threadIdx.x < world) counts[threadIdx.x] = scnt[threadIdx.x];
}

// one warp per row of the rank-major sequence: find its rank from the counts, copy its 400 bytes to the query's position
__global__ void k_shard_scatter(const uint32_t *__restrict__ gathered, uint32_t cap, const uint32_t *__restrict__ order,
                                const uint32_t *__restrict__ counts, uint32_t m, uint32_t world, uint32_t *__restrict__ out)
{
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= m) return;
    uint32_t r = 0, off = 0;
    while (r + 1 < world && p >= off + counts[r]) { off += counts[r]; ++r; }
    const uint4 *src = reinterpret_cast<const uint4 *>(gathered + ((size_t)r * cap + (p - off)) * K);
    uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)order[p] * K);
    if (lane < K / 4) dst[lane] = src[lane];
}

cudaError_t launch_shard_scatter(hvs_engine *e, const uint32_t *gathered_dev, uint32_t cap, const uint32_t *order_dev,
                                 const uint32_t *counts_dev, uint32_t m, uint32_t world, uint32_t *out_ids_dev)
{
    if (!m) return cudaSuccess;
    k_shard_scatter<<<(unsigned)(((size_t)m * 32 + 255) / 256), 256, 0, e->stream>>>(gathered_dev, cap, order_dev, counts_dev, m, world, out_ids_dev);
    return cudaGetLastError();
}

// ---- host side -------------------------------------------------------------------------------------------------------
#define PDCK(call) do { cudaError_t _c = (call); if (_c != cudaSuccess) return _c; } while (0)

// order_dev: all m query indices rank-major (each rank's in (arena, begin, end) order); counts_dev: [world].  Both stay on the
// device (the caller copies what it needs).  Precondition: (sum of costs) * world * stripes < 2^64 (checked by the caller).
cudaError_t shard_assign_dev(hvs_engine *e, const QSlice *d_sl, uint32_t m, uint32_t world, uint32_t stripes,
                             uint32_t **order_dev, uint32_t **counts_dev)
{
    cudaStream_t s = e->stream;
    PlanDev &P = e->pdev;
    const uint32_t n = e->index.n;
    uint32_t nb = 1;
    while ((1ull << nb) <= (uint64_t)n) ++nb;
    PDCK(P.keys_in.ensure((size_t)m * 8));
    PDCK(P.keys.ensure((size_t)m * 8));
    PDCK(P.vals_in.ensure((size_t)m * 4));
    PDCK(P.vals.ensure((size_t)m * 4));
    PDCK(P.sa_owner_in.ensure((size_t)m * 4));
    PDCK(P.sa_owner.ensure((size_t)m * 4));
    PDCK(P.sa_order.ensure((size_t)m * 4));
    PDCK(P.sa_counts.ensure(256 * 4));
    const unsigned gb = (m + 255) / 256;
    k_sa_keys<<<gb, 256, 0, s>>>(d_sl, m, nb, P.keys_in.as<uint64_t>(), P.vals_in.as<uint32_t>());
    const int end_bit = (int)(2 * nb + 2);
    PDCK(P.sort_tmp.ensure(radix_sort_temp_bytes(m)));
    PDCK(radix_sort_pairs<uint64_t>(P.keys_in.as<uint64_t>(), P.vals_in.as<uint32_t>(), P.keys.as<uint64_t>(), P.vals.as<uint32_t>(), m, end_bit,
                                    P.sort_tmp.p, s));
    k_sa_assign<<<1, SCAN_T, 0, s>>>(d_sl, P.vals.as<uint32_t>(), m, world, stripes, shard_query_cost(), shard_row_cost(), P.sa_owner_in.as<uint32_t>(),
                                     P.sa_counts.as<uint32_t>());
    // stable sort by owner (one 8-bit pass): rank-major, each rank's queries keep the (arena, begin, end) order
    P.launches = 2u + radix_sort_launches(m, end_bit) + radix_sort_launches(m, 8);
    PDCK(radix_sort_pairs<uint32_t>(P.sa_owner_in.as<uint32_t>(), P.vals.as<uint32_t>(), P.sa_owner.as<uint32_t>(), P.sa_order.as<uint32_t>(), m, 8,
                                    P.sort_tmp.p, s));
    PDCK(cudaGetLastError());
    *order_dev = P.sa_order.as<uint32_t>();
    *counts_dev = P.sa_counts.as<uint32_t>();
    return cudaSuccess;
}

cudaError_t plan_dev_begin(hvs_engine *e, const QSlice *d_sl, uint32_t m, const PlanCfg &cfg, PlanHeader *h_out)
{
    cudaStream_t s = e->stream;
    PlanDev &P = e->pdev;
    const uint32_t n = e->index.n;
    const uint32_t ncell = n / CELL + 2;
    uint32_t nb = 1;
    while ((1ull << nb) <= (uint64_t)n) ++nb;                   // begin < 2^nb, end <= n < 2^nb  (end field has nb + 1 bits)
    P.nb = nb;
    const uint32_t max_chunks = 2 * (n / 8192 + 2);          // R >= 8192 in both arenas
    PDCK(P.header.ensure(sizeof(PlanHeader)));
    PDCK(P.diff.ensure((size_t)2 * ncell * 4));
    PDCK(P.pref.ensure((size_t)2 * (ncell + 1) * 8));
    PDCK(P.cls.ensure(m));
    PDCK(P.keys_in.ensure((size_t)m * 8));
    PDCK(P.keys.ensure((size_t)m * 8));
    PDCK(P.vals_in.ensure((size_t)m * 4));
    PDCK(P.vals.ensure((size_t)m * 4));
    PDCK(P.nch.ensure((size_t)m * 4));
    PDCK(P.qoff.ensure((size_t)(m + 1) * 4));
    PDCK(P.cdiff.ensure((size_t)(max_chunks + 2) * 4));
    PDCK(P.cbeg.ensure((size_t)(max_chunks + 2) * 4));
    PDCK(P.ibase_rest.ensure((size_t)(max_chunks + 2) * 4));
    PDCK(P.nrest.ensure((size_t)(max_chunks + 2) * 4));
    PDCK(P.cstart.ensure((size_t)(max_chunks + 2) * 4));
    PDCK(P.ibase.ensure((size_t)(max_chunks + 2) * 4));
    PDCK(e->h_header.ensure(sizeof(PlanHeader)));
    PlanHeader *H = P.header.as<PlanHeader>();
    PDCK(cudaMemsetAsync(H, 0, sizeof(PlanHeader), s));
    PDCK(cudaMemsetAsync(P.diff.p, 0, (size_t)2 * ncell * 4, s));
    PDCK(cudaMemsetAsync(P.cdiff.p, 0, (size_t)(max_chunks + 2) * 4, s));
    PDCK(cudaMemsetAsync(P.cbeg.p, 0, (size_t)(max_chunks + 2) * 4, s));
    int *diff0 = P.diff.as<int>(), *diff1 = diff0 + ncell;
    long long *pref0 = P.pref.as<long long>(), *pref1 = pref0 + ncell + 1;
    const unsigned gb = (m + 255) / 256;
    if (cfg.tile_allowed) {
        k_pd_depth<<<gb, 256, 0, s>>>(d_sl, m, diff0, diff1, H);
        k_pd_cells<<<2, SCAN_T, 0, s>>>(diff0, diff1, pref0, pref1, ncell);
    }
    k_pd_classify<<<gb, 256, 0, s>>>(d_sl, m, pref0, pref1, cfg, P.cls.as<uint8_t>(), H);
    k_pd_params<<<1, 32, 0, s>>>(cfg, H);
    k_pd_keys<<<gb, 256, 0, s>>>(d_sl, m, P.cls.as<uint8_t>(), nb, H, P.keys_in.as<uint64_t>(), P.vals_in.as<uint32_t>());
    const int end_bit = (int)(2 * nb + 4);
    PDCK(P.sort_tmp.ensure(radix_sort_temp_bytes(m)));
    PDCK(radix_sort_pairs<uint64_t>(P.keys_in.as<uint64_t>(), P.vals_in.as<uint32_t>(), P.keys.as<uint64_t>(), P.vals.as<uint32_t>(), m, end_bit,
                                    P.sort_tmp.p, s));
    P.launches = (cfg.tile_allowed ? 7u : 5u) + radix_sort_launches(m, end_bit);    // depth, cells, classify, params, keys, the sort, chunks, scan
    k_pd_chunks<<<gb, 256, 0, s>>>(P.keys.as<uint64_t>(), P.vals.as<uint32_t>(), m, nb, d_sl, H, P.cdiff.as<int>(), P.cbeg.as<uint32_t>(),
                                   P.nch.as<uint32_t>());
    k_pd_scan<<<1, SCAN_T, 0, s>>>(H, P.cdiff.as<int>(), cfg.bq, P.cbeg.as<uint32_t>(), cfg.seed_phase, P.cstart.as<uint32_t>(), P.ibase.as<uint32_t>(),
                                   P.ibase_rest.as<uint32_t>(), P.nrest.as<uint32_t>(), P.nch.as<uint32_t>(), P.qoff.as<uint32_t>());
    PDCK(cudaGetLastError());
    PDCK(cudaMemcpyAsync(e->h_header.p, H, sizeof(PlanHeader), cudaMemcpyDeviceToHost, s));
    PDCK(cudaStreamSynchronize(s));
    *h_out = *e->h_header.as<PlanHeader>();
    return cudaSuccess;
}

cudaError_t plan_dev_fill(hvs_engine *e, const QSlice *d_sl, const PlanHeader &h, const PlanCfg &cfg, uint32_t *item_q_dev,
                          TileItem *items_dev, uint32_t *qlists_dev)
{
    if (!h.nchunk_total) return cudaSuccess;
    PlanDev &P = e->pdev;
    k_pd_fill<<<h.nchunk_total, 256, 0, e->stream>>>(P.header.as<PlanHeader>(), P.vals.as<uint32_t>(), d_sl, cfg.bq, cfg.kind, P.cstart.as<uint32_t>(),
                                                      P.ibase.as<uint32_t>(), P.ibase_rest.as<uint32_t>(), P.nrest.as<uint32_t>(), P.qoff.as<uint32_t>(), item_q_dev, items_dev, qlists_dev,
                                                      &P.header.as<PlanHeader>()->pairs_computed);
    return cudaGetLastError();
}

cudaError_t plan_dev_redo(hvs_engine *e, const uint32_t *flags_dev, uint32_t n_tile, uint32_t *redo_dev)
{
    if (!n_tile) return cudaSuccess;
    PlanDev &P = e->pdev;
    k_pd_redo<<<(n_tile + 255) / 256, 256, 0, e->stream>>>(flags_dev, P.vals.as<uint32_t>(), n_tile, redo_dev, P.header.as<PlanHeader>());
    return cudaGetLastError();
}

}  // namespace hvs
