// hvs_plan.cu -- K1: per-query slice lookup (device) + the query planner (host).
//
// Reference: the decode at include/baseline.hpp:90-93 and the four predicate scans at
// include/baseline.hpp:107-136.  Here each predicate becomes two binary searches over the sorted
// keys of one arena (hvs_index.cu); the planner then buckets queries that share rows into
// 128-query tile items, or sends sparse / tiny slices to the direct scan kernel.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <thread>

#include "hvs_engine.h"

namespace hvs {

template <class T>
__device__ __forceinline__ uint32_t lower_bound_dev(const T *__restrict__ a, uint32_t n, T key)
{   // first i with a[i] >= key
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
template <class T>
__device__ __forceinline__ uint32_t upper_bound_dev(const T *__restrict__ a, uint32_t n, T key)
{   // first i with a[i] > key
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One warp per query: lane 0 decodes and searches, the warp computes ||q||^2.
__global__ void k_plan_search(const float *__restrict__ queries, uint32_t m, const uint32_t *__restrict__ keys_t,
                              const uint64_t *__restrict__ keys_ct, uint32_t n, uint32_t small_max, QSlice *__restrict__ out,
                              unsigned long long *__restrict__ k1acc /* {sum of max(len, K), queries with len <= small_max} */)
{
    __shared__ unsigned long long s_pairs;
    __shared__ unsigned int s_small;
    if (threadIdx.x == 0) { s_pairs = 0; s_small = 0; }
    __syncthreads();
    uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q < m) {                                  // (whole warps: one query per warp)
    const float *row = queries + (size_t)q * QROW;
    float acc = 0.f;
    for (int i = lane; i < DIM; i += 32) { float v = row[4 + i]; acc = fmaf(v, v, acc); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
    uint32_t type = f2u32_x86(row[0]);           // baseline.hpp:90
    int32_t v = f2i32_x86(row[1]);               // baseline.hpp:91 (truncation toward zero)
    float l = row[2], r = row[3];                // baseline.hpp:92-93
    QSlice s;
    s.arena = ARENA_T; s.begin = 0; s.end = 0; s.qnorm = acc;
    const bool range_ok = !f32_isnan(l) && !f32_isnan(r);   // any comparison with NaN is false
    if (type == 0) {                              // baseline.hpp:107-113: every scanned row
        s.end = n;
    } else if (type == 2) {                       // baseline.hpp:123-129: l <= T <= r, inclusive
        if (range_ok) {
            s.begin = lower_bound_dev(keys_t, n, ord_key(l));
            s.end = upper_bound_dev(keys_t, n, ord_key(r));
        }
    } else if (type == 1 || type == 3) {          // baseline.hpp:114-122 / 130-136: C == (float)v
        s.arena = ARENA_CT;
        uint64_t kc = (uint64_t)ord_key((float)v) << 32;
        if (type == 1) {
            s.begin = lower_bound_dev(keys_ct, n, kc);
            s.end = lower_bound_dev(keys_ct, n, (uint64_t)(kc + (1ull << 32)));
        } else if (range_ok) {
            s.begin = lower_bound_dev(keys_ct, n, (uint64_t)(kc | ord_key(l)));
            s.end = upper_bound_dev(keys_ct, n, (uint64_t)(kc | ord_key(r)));
        }
    }                                             // any other type: no branch taken, empty set
    if (s.end < s.begin) s.end = s.begin;         // l > r
    out[q] = s;
    const uint32_t len = s.end - s.begin;
    atomicAdd(&s_pairs, (unsigned long long)(len > (uint32_t)K ? len : (uint32_t)K));     // per block first: one global atomic pair per 8 queries
    if (len <= small_max) atomicAdd(&s_small, 1u);
    }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&k1acc[0], s_pairs);
        if (s_small) atomicAdd(&k1acc[1], (unsigned long long)s_small);
    }
}

cudaError_t launch_plan_search(hvs_engine *e, const float *queries_dev, uint32_t m, QSlice *slices_dev)
{
    if (!m) return cudaSuccess;
    const Index &ix = e->index;
    cudaError_t c = e->d_k1acc.ensure(16);
    if (c == cudaSuccess) c = cudaMemsetAsync(e->d_k1acc.p, 0, 16, e->stream);
    if (c != cudaSuccess) return c;
    unsigned blocks = (unsigned)(((size_t)m * 32 + 255) / 256);
    k_plan_search<<<blocks, 256, 0, e->stream>>>(queries_dev, m, ix.keys_t.as<uint32_t>(), ix.keys_ct.as<uint64_t>(), ix.n, SMALL_MAX, slices_dev,
                                                 e->d_k1acc.as<unsigned long long>());
    return cudaGetLastError();
}

__global__ void k_gather_queries(const float *__restrict__ queries, const QSlice *__restrict__ slices,
                                 const uint32_t *__restrict__ own, uint32_t n_own, float *__restrict__ q_sub, QSlice *__restrict__ sl_sub)
{
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_own) return;
    const uint32_t q = own[i];
    const float4 *src = reinterpret_cast<const float4 *>(queries + (size_t)q * QROW);      // 416-byte rows: 16-byte aligned
    float4 *dst = reinterpret_cast<float4 *>(q_sub + (size_t)i * QROW);
    if (lane < QROW / 4) dst[lane] = src[lane];
    if (lane == 31) sl_sub[i] = slices[q];
}

cudaError_t launch_gather_queries(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev, const uint32_t *own_dev,
                                  uint32_t n_own, float *q_sub, QSlice *sl_sub)
{
    if (!n_own) return cudaSuccess;
    k_gather_queries<<<(unsigned)(((size_t)n_own * 32 + 255) / 256), 256, 0, e->stream>>>(queries_dev, slices_dev, own_dev, n_own, q_sub, sl_sub);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Host planner.
//
// A tile item makes a batch of queries sweep a run of arena rows together.  On the FFMA kernel its
// cost is rows x 128 pair-slots whatever the queries need, while a direct scan costs exactly the rows
// the query needs but runs at HBM speed (400 B per pair): a query takes the FFMA tile path when,
// averaged over its slice, at least 128 / direct_cost_ratio other queries want the same rows.  On the
// tensor kernel pair-slots are nearly free and the sweep is priced by the bytes of the fp16 image it
// streams (224 B per row per <= 256 queries), so sharing the rows with about one other query is enough.
// One solve never mixes the two tile kernels: their candidate thresholds are shared per query and
// carry different error margins.
//
// Everything is counting sorts over (chunk, query) incidences -- O(m log m + incidences), no
// comparator sorts over the incidence list.
namespace {
// stable LSD radix sort of (64-bit key, payload) pairs, 11 bits per pass; digits that are the same everywhere are skipped
void radix_sort_keys(std::vector<std::pair<uint64_t, uint32_t>> &keys)
{
    std::vector<std::pair<uint64_t, uint32_t>> tmp(keys.size());
    uint64_t all_or = 0, all_and = ~0ull;
    for (const auto &kv : keys) { all_or |= kv.first; all_and &= kv.first; }
    for (int shift = 0; shift < 64; shift += 11) {
        if ((((all_or ^ all_and) >> shift) & 0x7ffull) == 0) continue;
        uint32_t hist[2049] = {0};
        for (const auto &kv : keys) ++hist[((kv.first >> shift) & 0x7ffull) + 1];
        for (int d = 0; d < 2048; ++d) hist[d + 1] += hist[d];
        for (const auto &kv : keys) tmp[hist[(kv.first >> shift) & 0x7ffull]++] = kv;
        keys.swap(tmp);
    }
}

template <class F>
void plan_parallel(unsigned nthreads, size_t njobs, F &&job)
{
    if (nthreads <= 1 || njobs <= 1) { for (size_t j = 0; j < njobs; ++j) job(j); return; }
    std::atomic<size_t> next{0};
    auto worker = [&]() { for (size_t j; (j = next.fetch_add(1)) < njobs;) job(j); };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthreads && t < njobs; ++t) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
}
}  // namespace

void plan_begin(const QSlice *sl, uint32_t m, const PlanParams &pp, Plan &P)
{
    static const bool dbg = getenv("HVS_PLAN_DEBUG") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { if (dbg) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "  plan %-14s %.2f ms\n", what, std::chrono::duration<double, std::milli>(t - T0).count()); T0 = t; } };
    P.reset();
    lap("reset");
    const bool tensor = pp.tensor_available && (pp.mode == HVS_MODE_AUTO || pp.mode == HVS_MODE_TENSOR);
    const uint32_t BQ = tensor ? (uint32_t)QT_TENSOR : (uint32_t)QT;
    std::vector<uint8_t> &is_tile = P.is_tile;
    is_tile.assign(m, 0);
    uint64_t tile_qrows = 0;
    for (uint32_t i = 0; i < m; ++i) {
        uint32_t len = sl[i].end - sl[i].begin;
        P.pairs += len > (uint32_t)K ? len : (uint32_t)K;
    }
    if (pp.mode != HVS_MODE_DIRECT && pp.approx_ok) {
        constexpr uint32_t CELL = 1024;
        const double need = tensor ? pp.tensor_min_depth : (double)QT / pp.direct_cost_ratio;
        for (uint32_t a = 0; a < 2; ++a) {
            uint32_t maxend = 0;
            for (uint32_t i = 0; i < m; ++i)
                if (sl[i].arena == a && sl[i].end > sl[i].begin) maxend = std::max(maxend, sl[i].end);
            if (!maxend) continue;
            uint32_t ncell = (maxend + CELL - 1) / CELL;
            std::vector<int64_t> depth(ncell + 1, 0);
            for (uint32_t i = 0; i < m; ++i)
                if (sl[i].arena == a && sl[i].end > sl[i].begin) {
                    depth[sl[i].begin / CELL] += 1;
                    depth[(sl[i].end - 1) / CELL + 1] -= 1;
                }
            std::vector<int64_t> pref(ncell + 1, 0);   // pref[c] = sum of depth over cells < c
            int64_t run = 0;
            for (uint32_t c = 0; c < ncell; ++c) { run += depth[c]; pref[c + 1] = pref[c] + run; }
            for (uint32_t i = 0; i < m; ++i)
                if (sl[i].arena == a && sl[i].end - sl[i].begin >= pp.min_tile_len) {
                    uint32_t c0 = sl[i].begin / CELL, c1 = (sl[i].end - 1) / CELL + 1;
                    double avg = (double)(pref[c1] - pref[c0]) / (double)(c1 - c0);
                    if (avg >= need) {
                        is_tile[i] = 1;
                        tile_qrows += sl[i].end - sl[i].begin;
                        P.pairs_tile += std::max(sl[i].end - sl[i].begin, (uint32_t)K);
                    }
                }
        }
    }
    if (tile_qrows && tile_qrows < pp.min_tile_pairs && pp.mode != HVS_MODE_TENSOR) {   // tiny job: not worth a sweep
        std::fill(is_tile.begin(), is_tile.end(), (uint8_t)0);
        tile_qrows = 0;
        P.pairs_tile = 0;
    }
    lap("depth");
    for (uint32_t i = 0; i < m; ++i)
        if (!is_tile[i]) {
            const uint32_t len = sl[i].end - sl[i].begin;
            P.pairs_computed += len;
            if (len <= pp.small_max && pp.small_max) { ++P.n_small; P.pairs_small += std::max(len, (uint32_t)K); }   // K4s has it
            else P.direct_q.push_back(i);
        }
    // direct queries: neighbours in an arena share L2 lines.  A counting sort on (arena, begin >> 12) is all the
    // locality the scan needs and costs O(m) (a comparator sort of 4x10^4 slices took 3 ms of a 4 ms solve).
    if (P.direct_q.size() > 1) {
        constexpr uint32_t NB = 4096;                                    // buckets per arena
        uint32_t maxb = 0, SH = 0;
        for (uint32_t q : P.direct_q) maxb = std::max(maxb, sl[q].begin);
        while ((maxb >> SH) >= NB) ++SH;
        std::vector<uint32_t> &cnt = P.nlist;                            // scratch, re-assigned by plan_finish
        cnt.assign(2 * NB + 1, 0);
        auto bucket = [&](uint32_t q) { return sl[q].arena * NB + (sl[q].begin >> SH); };
        for (uint32_t q : P.direct_q) ++cnt[bucket(q) + 1];
        for (uint32_t b = 0; b < 2 * NB; ++b) cnt[b + 1] += cnt[b];
        std::vector<uint32_t> &out = P.qpos;                             // scratch as well
        out.resize(P.direct_q.size());
        for (uint32_t q : P.direct_q) out[cnt[bucket(q)]++] = q;
        P.direct_q.swap(out);
    }
    lap("classify");
    P.tensor = tensor;
    P.BQ = BQ;
    if (!tile_qrows) return;

    // chunk size: aim at ~16 items per SM over the whole job, power of two
    static const uint32_t ips = [] { const char *v = getenv("HVS_ITEMS_PER_SM"); int k = v ? atoi(v) : 0; return (uint32_t)(k > 0 ? k : 16); }();
    const uint32_t R = plan_chunk_rows(tile_qrows, BQ, ips, 148);
    P.R = R;

    lap("chunking");
    std::vector<uint32_t> *order = P.order;
    uint64_t incid = 0;
    P.nthreads = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 8u);
    uint32_t maxend[2] = {0, 0};
    for (uint32_t a = 0; a < 2; ++a) { order[a].clear(); P.sort_keys[a].clear(); }
    for (uint32_t i = 0; i < m; ++i)
        if (is_tile[i]) {
            const uint32_t a = sl[i].arena;
            order[a].push_back(i);
            maxend[a] = std::max(maxend[a], sl[i].end);
        }
    // Chunk boundaries: the multiples of R.  (Round 1 carried an experimental variant that moved boundaries to the rows
    // where many slices begin -- category starts; measured in round 2: no gain in the sweep, +1.4 ms of planning; removed.)
    for (uint32_t a = 0; a < 2; ++a) {
        std::vector<uint32_t> &b = P.bnd[a];
        b.clear();
        if (order[a].empty()) continue;
        const uint32_t nchunk = (maxend[a] + R - 1) / R;
        for (uint32_t c = 0; c <= nchunk; ++c) b.push_back((uint32_t)std::min<uint64_t>((uint64_t)c * R, 0xffffffffull));
    }
    P.uniform_chunks = true;
    for (uint32_t a = 0; a < 2; ++a)
        for (uint32_t i : order[a]) incid += P.chunk_of(a, sl[i].end - 1) - P.chunk_of(a, sl[i].begin) + 1;
    P.incid = incid;
    if (incid < 65536) P.nthreads = 1;
    lap("collect");
    // tile queries of each arena ordered by (begin, end, index): each chunk's batches then group queries
    // with similar slices, which keeps the union of rows an item sweeps tight
    plan_parallel(P.nthreads, 2, [&](size_t a) {
        std::vector<uint32_t> &ord = order[a];
        if (ord.empty()) return;
        bool same = true;                                            // all slices equal (e.g. unfiltered queries): already ordered
        for (size_t k = 1; k < ord.size() && same; ++k)
            same = sl[ord[k]].begin == sl[ord[0]].begin && sl[ord[k]].end == sl[ord[0]].end;
        if (same) return;
        // sort (begin, end, index) as plain integers: no slice look-ups inside the comparator
        std::vector<std::pair<uint64_t, uint32_t>> &keys = P.sort_keys[a];
        keys.resize(ord.size());
        for (size_t k = 0; k < ord.size(); ++k) keys[k] = {((uint64_t)sl[ord[k]].begin << 32) | sl[ord[k]].end, ord[k]};
        radix_sort_keys(keys);
        for (size_t k = 0; k < ord.size(); ++k) ord[k] = keys[k].second;
    });
    lap("sort");
    // tasks = (arena, block of chunks); the (C,T) arena first: its slices are short, so the GPU gets work early
    for (int a = 1; a >= 0; --a) {
        if (order[a].empty()) continue;
        const uint32_t nchunk = (uint32_t)P.bnd[a].size() - 1;
        const uint32_t per = std::max(1u, (nchunk + 7) / 8);
        for (uint32_t c = 0; c < nchunk; c += per) P.tasks.push_back({(uint32_t)a, c, std::min(nchunk, c + per), 0});
    }
    {   // incidences per task: every query adds its chunk range to the (few) tasks it overlaps
        size_t tbase[2] = {0, 0};
        uint32_t tper[2] = {1, 1};
        for (size_t ti = 0; ti < P.tasks.size(); ++ti)
            if (ti == 0 || P.tasks[ti].arena != P.tasks[ti - 1].arena) {
                tbase[P.tasks[ti].arena] = ti;
                tper[P.tasks[ti].arena] = P.tasks[ti].c1 - P.tasks[ti].c0;      // every task but the last of an arena spans `per` chunks
            }
        for (uint32_t a = 0; a < 2; ++a)
            for (uint32_t i : order[a]) {
                const uint32_t lo = P.chunk_of(a, sl[i].begin), hi = P.chunk_of(a, sl[i].end - 1) + 1;
                for (uint32_t t = lo / tper[a]; t <= (hi - 1) / tper[a]; ++t) {
                    Plan::Task &tk = P.tasks[tbase[a] + t];
                    const uint32_t l2 = std::max(lo, tk.c0), h2 = std::min(hi, tk.c1);
                    if (h2 > l2) tk.incid += h2 - l2;
                }
            }
    }
    // groups: ~20 % / 40 % / 40 % of the incidences, so that little planning stands before the first launch; when the
    // (C,T) arena's tasks are a small head of the list they form a group of their own -- planned in a fraction of a
    // millisecond, it gives the GPU work while the host is still planning the T arena
    const int ng = incid >= 200000 && P.tasks.size() >= 3 ? 3 : 1;
    const double cutf[3] = {ng == 1 ? 1.0 : 0.2, 0.6, 1.0};
    uint64_t run = 0;
    int g = 0;
    for (size_t ti = 0; ti < P.tasks.size(); ++ti) {
        run += P.tasks[ti].incid;
        const bool last = ti + 1 == P.tasks.size();
        if (ng > 1 && !last && P.group_end.empty() && P.tasks[ti + 1].arena != P.tasks[ti].arena &&
            (double)run < cutf[0] * (double)incid && run > 0) {
            P.group_end.push_back((uint32_t)ti + 1);                  // head group: the whole (C,T) arena
            continue;
        }
        if (g < ng - 1 && !last && (double)run >= cutf[g] * (double)incid) { P.group_end.push_back((uint32_t)ti + 1); ++g; }
    }
    P.group_end.push_back((uint32_t)P.tasks.size());
    if (P.locals.size() < P.tasks.size()) P.locals.resize(P.tasks.size());
    lap("order");
}

void plan_group(const QSlice *sl, Plan &P, size_t g, uint32_t &item_begin, uint32_t &item_end)
{
    static const bool dbg = getenv("HVS_PLAN_DEBUG") != nullptr;
    const auto T0 = std::chrono::steady_clock::now();
    using Local = Plan::Local;
    const uint32_t R = P.R, BQ = P.BQ;
    const bool tensor = P.tensor;
    std::vector<uint32_t> *order = P.order;
    std::vector<Local> &locals = P.locals;
    const size_t t0 = g ? P.group_end[g - 1] : 0, t1 = P.group_end[g];
    auto run_task = [&](size_t tj) {
        const size_t ti = t0 + tj;
        const Plan::Task tk = P.tasks[ti];
        Local &L = locals[ti];
        L.items.clear(); L.item_q.clear(); L.pairs_computed = 0;
        const std::vector<uint32_t> &ord = order[tk.arena];
        const uint32_t nc = tk.c1 - tk.c0;
        std::vector<uint32_t> &cstart = L.cstart, &fill = L.fill, &cur = L.cur;
        cstart.assign(nc + 1, 0);
        auto range = [&](uint32_t i, uint32_t &lo, uint32_t &hi) {     // chunks of query i inside this task, [lo, hi)
            lo = std::max(P.chunk_of(tk.arena, sl[i].begin), tk.c0);
            hi = std::min(P.chunk_of(tk.arena, sl[i].end - 1) + 1, tk.c1);
        };
        for (uint32_t i : ord) {
            uint32_t lo, hi;
            range(i, lo, hi);
            for (uint32_t c = lo; c < hi; ++c) ++cstart[c - tk.c0 + 1];
        }
        for (uint32_t c = 0; c < nc; ++c) cstart[c + 1] += cstart[c];
        fill.resize(cstart[nc]);
        cur.assign(cstart.begin(), cstart.end() - 1);
        for (uint32_t i : ord) {
            uint32_t lo, hi;
            range(i, lo, hi);
            for (uint32_t c = lo; c < hi; ++c) fill[cur[c - tk.c0]++] = i;
        }
        for (uint32_t c = 0; c < nc; ++c) {
            const uint32_t c0 = P.bnd[tk.arena][tk.c0 + c];
            const uint64_t c1 = P.bnd[tk.arena][tk.c0 + c + 1];
            for (uint32_t t = cstart[c]; t < cstart[c + 1]; t += BQ) {
                const uint32_t te = std::min(cstart[c + 1], t + BQ);
                TileItem it{};
                it.arena = tk.arena;
                uint32_t lo = 0xffffffffu, hi = 0;
                for (uint32_t k = t; k < te; ++k) {
                    lo = std::min(lo, std::max(sl[fill[k]].begin, c0));
                    hi = std::max(hi, (uint32_t)std::min<uint64_t>(sl[fill[k]].end, c1));
                }
                it.row_begin = lo; it.row_end = hi;
                it.nq = te - t;
                it.q_off = (uint32_t)L.item_q.size();                 // rebased below
                L.item_q.insert(L.item_q.end(), fill.begin() + t, fill.begin() + te);
                it.kind = tensor ? 1u : 0u;
                L.items.push_back(it);
                L.pairs_computed += (uint64_t)(hi - lo) * it.nq;
            }
        }
    };
    plan_parallel(P.nthreads, t1 - t0, run_task);
    item_begin = (uint32_t)P.items.size();
    for (size_t ti = t0; ti < t1; ++ti) {
        Local &L = locals[ti];
        const uint32_t base = (uint32_t)P.item_q.size();
        for (auto it : L.items) { it.q_off += base; P.items.push_back(it); }
        P.item_q.insert(P.item_q.end(), L.item_q.begin(), L.item_q.end());
        P.pairs_computed += L.pairs_computed;
    }
    item_end = (uint32_t)P.items.size();
    // longest first (LPT) inside the group; equal-cost items stay in (arena,row) order so that CTAs running
    // concurrently share the same rows in L2
    std::stable_sort(P.items.begin() + item_begin, P.items.end(), [](const TileItem &x, const TileItem &y) {
        return (x.row_end - x.row_begin) > (y.row_end - y.row_begin);
    });
    for (uint32_t k = item_begin; k < item_end; ++k) {
        TileItem &it = P.items[k];
        it.out_off = P.n_lists;
        P.n_lists += it.nq;
        if (it.kind) ++P.n_tensor; else ++P.n_ffma;
    }
    if (dbg) fprintf(stderr, "  plan group %zu      %.2f ms (%u items)\n", g, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count(), item_end - item_begin);
}

void plan_finish(const QSlice *sl, uint32_t m, Plan &P)
{
    (void)sl;
    if (P.items.empty()) return;
    const std::vector<uint8_t> &is_tile = P.is_tile;
    // candidate lists: CSR per tile query by counting (order inside a query is irrelevant)
    std::vector<uint32_t> &nlist = P.nlist, &qpos = P.qpos;
    nlist.assign(m, 0);
    // every thread owns a range of query indices and scans all (item, slot) pairs: no atomics, sequential reads
    const size_t nrange = P.nthreads > 1 ? P.nthreads : 1;
    auto qrange = [&](size_t r, uint32_t &q0, uint32_t &q1) { q0 = (uint32_t)((uint64_t)m * r / nrange); q1 = (uint32_t)((uint64_t)m * (r + 1) / nrange); };
    plan_parallel(P.nthreads, nrange, [&](size_t r) {
        uint32_t q0, q1;
        qrange(r, q0, q1);
        for (uint32_t q : P.item_q) if (q >= q0 && q < q1) ++nlist[q];
    });
    qpos.assign(m, 0);
    P.q_list_off.push_back(0);
    for (uint32_t i = 0; i < m; ++i)
        if (is_tile[i]) {
            qpos[i] = P.q_list_off.back();
            P.tile_q.push_back(i);
            P.q_list_off.push_back(qpos[i] + nlist[i]);
        }
    P.q_lists.resize(P.q_list_off.back());
    plan_parallel(P.nthreads, nrange, [&](size_t r) {
        uint32_t q0, q1;
        qrange(r, q0, q1);
        for (const TileItem &it : P.items) {
            const uint32_t *iq = P.item_q.data() + it.q_off;
            for (uint32_t s = 0; s < it.nq; ++s) {
                const uint32_t q = iq[s];
                if (q >= q0 && q < q1) P.q_lists[qpos[q]++] = it.out_off + s;
            }
        }
    });
}

// ------------------------------------------------------------------------------------------------
// Query sharding over `world` engines (primary variant: D replicated, queries split; the reference's own
// parallel variant splits D per query instead, include/optimized_parallel.hpp:100-146 / include/threading.hpp:116-118).
//
// What a rank's solve costs is the rows its queries sweep -- sum of slice lengths -- plus a fixed amount per query
// (threshold warm-up, finalize), and the tile kernels live on queries that share rows being batched together.  So the
// queries are ordered by (arena, begin, end) -- neighbours share rows -- and that sequence is cut, by cumulative cost,
// into world x STRIPES contiguous segments that are dealt to the ranks round-robin: every rank gets the same cost AND
// the same mix of query types (all-row sweeps, ranges, categories), while queries with equal or similar slices stay
// together in runs.  Pure function of the slices: every rank computes the same answer without talking to the others.
//   order  : all m query indices, rank-major (rank 0's queries first), each rank's in (arena, begin, end) order
//   counts : queries per rank
uint64_t shard_query_cost()
{
    static const uint64_t v = [] { const char *s = getenv("HVS_SHARD_QCOST"); long long k = s ? atoll(s) : 0; return k > 0 ? (uint64_t)k : SHARD_QUERY_COST; }();
    return v;
}

uint64_t shard_row_cost()
{
    static const uint64_t v = [] { const char *s = getenv("HVS_SHARD_ROWCOST"); long long k = s ? atoll(s) : -1; return k >= 0 ? (uint64_t)k : SHARD_ROW_COST; }();
    return v;
}

uint32_t shard_stripes(uint32_t m, uint32_t world)
{
    static const uint32_t env = [] { const char *s = getenv("HVS_SHARD_STRIPES"); int k = s ? atoi(s) : 0; return (uint32_t)(k > 0 ? k : 0); }();
    if (env) return env;
    uint32_t stripes = SHARD_STRIPES;
    while (stripes > 1 && (uint64_t)world * stripes * 64 > m) stripes >>= 1;      // small batches: fewer, longer runs
    return stripes;
}

void shard_assign(const QSlice *sl, uint32_t m, uint32_t world, uint32_t *order, uint32_t *counts)
{
    if (world == 0) return;
    for (uint32_t r = 0; r < world; ++r) counts[r] = 0;
    if (!m) return;
    std::vector<std::pair<uint64_t, uint32_t>> keys[2];
    for (uint32_t i = 0; i < m; ++i) keys[sl[i].arena & 1u].push_back({((uint64_t)sl[i].begin << 32) | sl[i].end, i});
    radix_sort_keys(keys[0]);
    radix_sort_keys(keys[1]);
    const uint64_t qcost = shard_query_cost(), rowcost = shard_row_cost();
    auto cost_of = [&](uint32_t i) { return (uint64_t)shard_cost_of(sl[i].end - sl[i].begin, qcost, rowcost); };
    // each arena's cost is cut into segments of its own: every rank gets its share of BOTH (a category query costs a
    // tenth of a range query; one cut over the whole order would hand all of them to a few ranks)
    const uint32_t stripes = shard_stripes(m, world);
    std::vector<uint8_t> owner(m);
    for (int a = 0; a < 2; ++a) {
        const uint64_t nseg = (uint64_t)world * (a ? std::min(stripes, SHARD_STRIPES_CT) : stripes);
        uint64_t total = 0;
        for (const auto &kv : keys[a]) total += cost_of(kv.second);
        const bool fits64 = (unsigned __int128)total * nseg < ((unsigned __int128)1 << 63);   // then plain 64-bit arithmetic (what the device kernel uses)
        uint64_t cum = 0;
        for (const auto &kv : keys[a]) {
            const uint64_t c = cost_of(kv.second);
            // the segment that holds the midpoint of this query's cost interval
            uint64_t seg = fits64 ? (cum + c / 2) * nseg / total : (uint64_t)((unsigned __int128)(cum + c / 2) * nseg / total);
            if (seg >= nseg) seg = nseg - 1;
            owner[kv.second] = (uint8_t)((seg + (a ? world / 2 : 0u)) % world);
            cum += c;
        }
    }
    for (uint32_t i = 0; i < m; ++i) ++counts[owner[i]];
    std::vector<uint32_t> off(world + 1, 0);
    for (uint32_t r = 0; r < world; ++r) off[r + 1] = off[r] + counts[r];
    for (int a = 0; a < 2; ++a)
        for (const auto &kv : keys[a]) order[off[owner[kv.second]]++] = kv.second;
}

void plan_build(const QSlice *sl, uint32_t m, const PlanParams &pp, Plan &P)
{
    plan_begin(sl, m, pp, P);
    for (size_t g = 0; g < P.group_end.size(); ++g) {
        uint32_t b, e;
        plan_group(sl, P, g, b, e);
    }
    plan_finish(sl, m, P);
}

}  // namespace hvs

extern "C" int hvs_plan_dryrun(const uint32_t *arena, const uint32_t *begin, const uint32_t *end, uint32_t m,
                               uint32_t mode, uint8_t *out_kind, uint32_t *out_items, uint32_t max_items,
                               uint64_t *out_pairs_computed)
{
    if ((!arena || !begin || !end || !out_kind) && m) return HVS_ERR_INVALID;
    std::vector<hvs::QSlice> sl(m);
    for (uint32_t i = 0; i < m; ++i) {
        if (arena[i] > 1 || end[i] < begin[i]) return HVS_ERR_INVALID;
        sl[i].arena = arena[i]; sl[i].begin = begin[i]; sl[i].end = end[i]; sl[i].qnorm = 0.f;
    }
    hvs::PlanParams pp;
    pp.mode = mode;
    pp.tensor_available = (mode == HVS_MODE_TENSOR || mode == HVS_MODE_AUTO);
    static thread_local hvs::Plan P;   // keeps its scratch between calls, like the engine's plan
    hvs::plan_build(sl.data(), m, pp, P);
    for (uint32_t i = 0; i < m; ++i) out_kind[i] = 1;
    for (uint32_t q : P.direct_q) out_kind[q] = 0;
    if (out_items)
        for (size_t k = 0; k < P.items.size() && k < max_items; ++k) {
            out_items[4 * k + 0] = P.items[k].arena;
            out_items[4 * k + 1] = P.items[k].row_begin;
            out_items[4 * k + 2] = P.items[k].row_end;
            out_items[4 * k + 3] = P.items[k].nq | (P.items[k].kind << 16);
        }
    if (out_pairs_computed) *out_pairs_computed = P.pairs_computed;
    return (int)P.items.size();
}

extern "C" int hvs_shard_assign_host(const uint32_t *arena, const uint32_t *begin, const uint32_t *end, uint32_t m,
                                     uint32_t world, uint32_t *out_order, uint32_t *out_counts)
{
    if (!world || world > 255 || !out_counts || (m && (!arena || !begin || !end || !out_order))) return HVS_ERR_INVALID;
    std::vector<hvs::QSlice> sl(m);
    for (uint32_t i = 0; i < m; ++i) {
        if (arena[i] > 1 || end[i] < begin[i]) return HVS_ERR_INVALID;
        sl[i].arena = arena[i]; sl[i].begin = begin[i]; sl[i].end = end[i]; sl[i].qnorm = 0.f;
    }
    hvs::shard_assign(sl.data(), m, world, out_order, out_counts);
    return HVS_OK;
}
