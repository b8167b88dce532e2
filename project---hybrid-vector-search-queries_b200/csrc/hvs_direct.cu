// hvs_direct.cu -- K4: direct streaming scan, one CTA per query.
//
// For slices that few other queries share (selective type-3 ranges, tiny categories) and as the
// engine's exact fallback.  Replaces, for one query, the reference's candidate loop + distance +
// top-100 (include/baseline.hpp:107-172; include/optimized.hpp:79-130 with Knn::check_add,
// include/optimized_impl.h:284-335).  The slice [begin,end) of an arena is contiguous, so tiles of
// 128 rows (51,200 B) are moved global->shared by the TMA engine as single 1-D bulk copies
// (cp.async.bulk + mbarrier, double buffered); each thread then owns one row and accumulates
// (x-q)^2 in the reference's order with unfused sub/mul/add, so distances are BIT-IDENTICAL to
// include/baseline.hpp:53-64.  Bound: HBM (400 algorithmic bytes per pair; ~300 FP32 ops per pair
// is 7x below the FP32 roof at that byte rate).
#include <algorithm>
#include <cstdlib>

#include "hvs_engine.h"
#include "hvs_topk.cuh"

namespace hvs {

constexpr int DT = 128;      // rows per tile == threads per CTA
constexpr int DCAP = 512;    // candidate buffer entries
constexpr int SPLIT_MAX = 5; // most parts a long slice is cut into when only a few queries are scanned (SPLIT_MAX * K <= DCAP)
constexpr int SPLIT_MIN_ROWS = 1024;

struct DirectSmem {
    alignas(128) float x[2][DT * DIM];
    alignas(16) float q[DIM];
    alignas(8) uint64_t bar[2];
    TopBuf<DCAP> top;
};

// split > 1 (few queries, long slices: the scan of ONE query is a serial chain of tiles, so a handful of queries leaves
// the GPU idle): blockIdx.y cuts the slice into `split` parts, every part leaves its best 100 (distance, row) keys in
// `scratch`, and the CTA that finishes last (a counter per query) merges them and completes the query.
__global__ void __launch_bounds__(DT, 2)
k_direct(const float *__restrict__ queries, const QSlice *__restrict__ slices, const uint32_t *__restrict__ q_list,
         Arena a0, Arena a1, const float *__restrict__ tail, uint32_t n_total, uint32_t id_offset, int partial,
         uint32_t split, uint64_t *__restrict__ scratch, uint32_t *__restrict__ counters, uint32_t skip_max,
         uint32_t *__restrict__ out_ids, float *__restrict__ out_dist, uint32_t *__restrict__ out_count)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DirectSmem &S = *reinterpret_cast<DirectSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const uint32_t q = q_list ? q_list[blockIdx.x] : blockIdx.x;
    const QSlice sl = slices[q];
    const Arena A = sl.arena == ARENA_T ? a0 : a1;
    const uint32_t len_all = sl.end - sl.begin;
    if (len_all <= skip_max && skip_max) return;             // K4s has this query (launches over ALL queries, no planner)
    const uint32_t parts = (split > 1 && len_all >= (uint32_t)SPLIT_MIN_ROWS) ? split : 1u;    // short slices are not worth cutting
    if (blockIdx.y >= parts) return;
    const uint32_t seg = parts > 1 ? (len_all + parts - 1) / parts : len_all;
    const uint32_t begin = min(sl.end, sl.begin + blockIdx.y * seg);
    const uint32_t len = min(sl.end - begin, seg);
    const uint32_t ntiles = (len + DT - 1) / DT;

    if (tid == 0) {
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
        mbar_fence_init();
    }
    S.top.init(tid);
    if (tid < DIM / 4)
        reinterpret_cast<float4 *>(S.q)[tid] = reinterpret_cast<const float4 *>(queries + (size_t)q * QROW + 4)[tid];
    __syncthreads();

    auto issue = [&](uint32_t it) {
        uint32_t rows = min((uint32_t)DT, len - it * DT);
        uint32_t bytes = rows * ROW_BYTES;
        mbar_expect_tx(&S.bar[it & 1], bytes);
        bulk_g2s(S.x[it & 1], A.x + (size_t)(begin + it * DT) * DIM, bytes, &S.bar[it & 1]);
    };
    if (tid == 0) {
        if (ntiles > 0) issue(0);
        if (ntiles > 1) issue(1);
    }
    for (uint32_t it = 0; it < ntiles; ++it) {
        mbar_wait(&S.bar[it & 1], (it >> 1) & 1);
        const uint32_t rows = min((uint32_t)DT, len - it * DT);
        bool high = false;                     // one of my pushes took a slot past the compaction mark
        if ((uint32_t)tid < rows) {
            float d = ref_dist_row(S.x[it & 1] + tid * DIM, S.q);
            if (d < S.top.thr) high = S.top.push(d, begin + it * DT + tid) >= (uint32_t)(DCAP - DT);
        }
        const bool need = __syncthreads_or(high);   // tile consumed, pushes visible; the decision is block-uniform
        if (tid == 0 && it + 2 < ntiles) issue(it + 2);
        if (need) S.top.compact(tid, DT, 0.f, K);
    }
    __syncthreads();
    if (parts > 1) {
        // this part's best K keys -> scratch; the last part to arrive merges all of them
        if ((int)S.top.cnt > K) S.top.compact(tid, DT, 0.f, K);
        const uint32_t c = min(S.top.cnt, (uint32_t)K);
        uint64_t *mine = scratch + ((size_t)blockIdx.x * split + blockIdx.y) * K;
        for (uint32_t i = tid; i < (uint32_t)K; i += DT) mine[i] = i < c ? S.top.cand[i] : KEY_INF;
        __threadfence();
        __syncthreads();
        __shared__ uint32_t s_last;
        if (tid == 0) s_last = atomicAdd(&counters[blockIdx.x], 1u) == parts - 1u ? 1u : 0u;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const uint64_t *all = scratch + (size_t)blockIdx.x * split * K;
        static_assert(SPLIT_MAX * K <= DCAP, "the merge loads every part's list into the candidate buffer");
        for (uint32_t i = tid; i < parts * (uint32_t)K; i += DT) S.top.cand[i] = __ldcg(all + i);
        __syncthreads();
        if (tid == 0) { S.top.cnt = parts * (uint32_t)K; S.top.thr = __int_as_float(0x7f800000); }
        __syncthreads();
        S.top.compact(tid, DT, 0.f, K);        // sorts: real keys first (len_all >= SPLIT_MIN_ROWS > K of them), KEY_INF fillers last; keeps K
    }
    finish_query(S.top, S.q, A, len_all, tail, n_total, id_offset, q, partial != 0, out_ids, out_dist, out_count, tid, DT);
}

// ---- K4s: tiny slices, one WARP per query ---------------------------------------------------------------------------
// Selective predicates (a category AND a narrow T range: SURVEY 2.2 "K4 small-slice", BASELINE configs[4]) leave 10^1..10^3
// rows per query.  A CTA per query, a 51 KB TMA stage and block-wide sorts are all fixed cost at that size, so these
// queries get a warp each (8 per CTA, no block-level synchronisation at all): lane = row, rows read straight from
// global memory (consecutive rows: the 32 lanes cover one contiguous 12.8 KB run), distances in the reference's
// arithmetic (ref_dist_row), candidates appended to a 256-entry list in shared memory by ballot, the list cut to the
// best 100 by a warp bitonic sort whenever it fills, the pad rule (include/baseline.hpp:138-147) applied from the
// resident tail buffer, one final warp sort, 100 ids written.  The kernel decides per query, on the device, whether the
// slice is small enough (len <= small_max) -- so it can be launched right behind the slice search, for ALL queries,
// before the host has seen a single slice; the host planner leaves those queries alone.
// Roof: latency (a dependent chain of 100 fp32 adds per row, two or three round trips to L2/HBM per query), not
// bandwidth: ~20 KB of rows per query.
constexpr int SW = 8;        // warps (queries) per CTA
constexpr int SCAP = 256;    // candidate list entries per warp

struct SmallSmem {
    uint64_t cand[SW][SCAP];
    alignas(16) float q[SW][DIM];
};

__global__ void __launch_bounds__(SW * 32)
k_small(const float *__restrict__ queries, const QSlice *__restrict__ slices, const uint32_t *__restrict__ q_list, uint32_t nq,
        uint32_t small_max, Arena a0, Arena a1, const float *__restrict__ tail, uint32_t n_total, uint32_t id_offset,
        int partial, uint32_t *__restrict__ out_ids, float *__restrict__ out_dist, uint32_t *__restrict__ out_count)
{
    __shared__ SmallSmem S;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * SW + w;
    if (qi >= nq) return;
    const uint32_t q = q_list ? q_list[qi] : qi;
    const QSlice sl = slices[q];
    const uint32_t len = sl.end - sl.begin;
    if (len > small_max) return;                                       // the planner routes this query (tile sweep or CTA scan)
    const Arena A = sl.arena == ARENA_T ? a0 : a1;
    uint64_t *cand = S.cand[w];
    float *qv = S.q[w];
    if (lane < DIM / 4)
        reinterpret_cast<float4 *>(qv)[lane] = reinterpret_cast<const float4 *>(queries + (size_t)q * QROW + 4)[lane];
    __syncwarp();
    const float INF = __int_as_float(0x7f800000);
    uint32_t cnt = 0;                                                  // warp-uniform
    float thr = INF;
    for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t row = sl.begin + base + lane;
        const bool valid = base + lane < len;
        // The whole row first -- 25 independent 16-byte loads in flight, ONE round trip to L2/HBM per 32 rows -- then the
        // reference's sequential sum.  Lanes past the end re-read the slice's last row (no divergence), and the warp
        // barrier keeps the compiler from sinking the loads into the dependent add chain (it interleaves them otherwise,
        // three in flight at a time).
        float d;
        {
            const uint32_t rr = valid ? row : sl.end - 1u;
            const float4 *x4 = reinterpret_cast<const float4 *>(A.x + (size_t)rr * DIM);
            const float4 *q4 = reinterpret_cast<const float4 *>(qv);
            float4 xr[DIM / 4];
#pragma unroll
            for (int i = 0; i < DIM / 4; ++i) xr[i] = __ldg(x4 + i);
            __syncwarp();
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < DIM / 4; ++i) acc = ref_accum4(acc, xr[i], q4[i]);
            d = valid ? acc : INF;
        }
        const bool pass = valid && d < thr;
        const uint32_t mask = __ballot_sync(0xffffffffu, pass);
        if (pass) cand[cnt + __popc(mask & ((1u << lane) - 1u))] = pack_key(d, row);
        cnt += __popc(mask);
        if (cnt > (uint32_t)(SCAP - 32)) {                             // keep the best K, tighten the threshold
            for (uint32_t i = cnt + lane; i < (uint32_t)SCAP; i += 32) cand[i] = KEY_INF;
            __syncwarp();
            warp_bitonic_sort(cand, SCAP, lane);
            cnt = K;
            thr = __uint_as_float((uint32_t)(cand[K - 1] >> 32));
        }
    }
    __syncwarp();
    if (cnt > (uint32_t)K) {
        const int n2 = next_pow2((int)cnt);
        for (uint32_t i = cnt + lane; i < (uint32_t)n2; i += 32) cand[i] = KEY_INF;
        __syncwarp();
        warp_bitonic_sort(cand, n2, lane);
        cnt = K;
    }
    // (dist, arena position) -> (dist, id): the output order is by (dist, id) like finish_query's
    for (uint32_t i = lane; i < cnt; i += 32) {
        const uint64_t k = cand[i];
        cand[i] = (k & 0xffffffff00000000ull) | A.ids[(uint32_t)k];
    }
    if (!partial && len < (uint32_t)K) {                               // include/baseline.hpp:138-147
        const uint32_t npad = (uint32_t)K - len;                       // cnt == len here: nothing was ever cut
        for (uint32_t s = lane; s < npad; s += 32) {
            const float d = ref_dist_row(tail + (size_t)s * DIM, qv);
            cand[cnt + s] = pack_key(d, n_total - 1u - s + id_offset);
        }
        cnt += npad;
    }
    for (uint32_t i = cnt + lane; i < 128u; i += 32) cand[i] = KEY_INF;
    __syncwarp();
    warp_bitonic_sort(cand, 128, lane);
    if (!partial) {
        for (int i = lane; i < K; i += 32) out_ids[(size_t)q * K + i] = (uint32_t)cand[i];
    } else {
        for (int i = lane; i < K; i += 32) {
            const uint64_t k = cand[i];
            const bool ok = (uint32_t)i < cnt;
            out_ids[(size_t)q * K + i] = ok ? (uint32_t)k : 0xffffffffu;
            out_dist[(size_t)q * K + i] = ok ? __uint_as_float((uint32_t)(k >> 32)) : INF;
        }
        if (lane == 0) out_count[q] = len;
    }
}

cudaError_t launch_small(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev, const uint32_t *q_list_dev,
                         uint32_t nq, uint32_t small_max, bool partial, uint32_t *out_ids, float *out_dist, uint32_t *out_count)
{
    if (!nq) return cudaSuccess;
    const Index &ix = e->index;
    k_small<<<(nq + SW - 1) / SW, SW * 32, 0, e->stream>>>(queries_dev, slices_dev, q_list_dev, nq, small_max, ix.arena(0), ix.arena(1),
                                                           ix.tail.as<float>(), ix.n_total, ix.id_offset, partial ? 1 : 0, out_ids,
                                                           out_dist, out_count);
    return cudaGetLastError();
}

cudaError_t direct_init_attributes()     // per device; called by hvs_create with the engine's device current
{
    return cudaFuncSetAttribute(k_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DirectSmem));
}

cudaError_t launch_direct(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev, const uint32_t *q_list_dev,
                          uint32_t nq, bool partial, uint32_t *out_ids, float *out_dist, uint32_t *out_count, uint32_t skip_max)
{
    if (!nq) return cudaSuccess;
    const int smem = (int)sizeof(DirectSmem);
    const Index &ix = e->index;
    // few queries: cut long slices so that the scan of a query is not one serial chain on one SM (2 CTAs fit an SM)
    uint32_t split = 1;
    static const bool split_ok = [] { const char *v = getenv("HVS_DIRECT_SPLIT"); return !(v && v[0] == '0'); }();
    if (split_ok && nq <= 2u * (uint32_t)e->sm_count) split = std::max<uint32_t>(1u, std::min<uint32_t>(SPLIT_MAX, (uint32_t)(4 * e->sm_count) / nq));
    uint64_t *scratch = nullptr;
    uint32_t *counters = nullptr;
    if (split > 1) {
        const size_t sb = (size_t)nq * split * K * 8;
        cudaError_t c = e->d_split.ensure(sb + (size_t)nq * 4);
        if (c != cudaSuccess) return c;
        scratch = e->d_split.as<uint64_t>();
        counters = reinterpret_cast<uint32_t *>(e->d_split.as<unsigned char>() + sb);
        c = cudaMemsetAsync(counters, 0, (size_t)nq * 4, e->stream);
        if (c != cudaSuccess) return c;
    }
    k_direct<<<dim3(nq, split), DT, smem, e->stream>>>(queries_dev, slices_dev, q_list_dev, ix.arena(0), ix.arena(1),
                                                       ix.tail.as<float>(), ix.n_total, ix.id_offset, partial ? 1 : 0, split, scratch, counters,
                                                       skip_max, out_ids, out_dist, out_count);
    return cudaGetLastError();
}

}  // namespace hvs
