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
#include "hvs_engine.h"
#include "hvs_topk.cuh"

namespace hvs {

constexpr int DT = 128;      // rows per tile == threads per CTA
constexpr int DCAP = 512;    // candidate buffer entries

struct DirectSmem {
    alignas(128) float x[2][DT * DIM];
    alignas(16) float q[DIM];
    alignas(8) uint64_t bar[2];
    TopBuf<DCAP> top;
};

__global__ void __launch_bounds__(DT, 2)
k_direct(const float *__restrict__ queries, const QSlice *__restrict__ slices, const uint32_t *__restrict__ q_list,
         Arena a0, Arena a1, const float *__restrict__ tail, uint32_t n_total, uint32_t id_offset, int partial,
         uint32_t *__restrict__ out_ids, float *__restrict__ out_dist, uint32_t *__restrict__ out_count)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DirectSmem &S = *reinterpret_cast<DirectSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const uint32_t q = q_list ? q_list[blockIdx.x] : blockIdx.x;
    const QSlice sl = slices[q];
    const Arena A = sl.arena == ARENA_T ? a0 : a1;
    const uint32_t len = sl.end - sl.begin;
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
        bulk_g2s(S.x[it & 1], A.x + (size_t)(sl.begin + it * DT) * DIM, bytes, &S.bar[it & 1]);
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
            if (d < S.top.thr) high = S.top.push(d, sl.begin + it * DT + tid) >= (uint32_t)(DCAP - DT);
        }
        const bool need = __syncthreads_or(high);   // tile consumed, pushes visible; the decision is block-uniform
        if (tid == 0 && it + 2 < ntiles) issue(it + 2);
        if (need) S.top.compact(tid, DT, 0.f, K);
    }
    __syncthreads();
    finish_query(S.top, S.q, A, len, tail, n_total, id_offset, q, partial != 0, out_ids, out_dist, out_count, tid, DT);
}

cudaError_t direct_init_attributes()     // per device; called by hvs_create with the engine's device current
{
    return cudaFuncSetAttribute(k_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DirectSmem));
}

cudaError_t launch_direct(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev, const uint32_t *q_list_dev,
                          uint32_t nq, bool partial, uint32_t *out_ids, float *out_dist, uint32_t *out_count)
{
    if (!nq) return cudaSuccess;
    const int smem = (int)sizeof(DirectSmem);
    const Index &ix = e->index;
    k_direct<<<nq, DT, smem, e->stream>>>(queries_dev, slices_dev, q_list_dev, ix.arena(0), ix.arena(1),
                                           ix.tail.as<float>(), ix.n_total, ix.id_offset, partial ? 1 : 0, out_ids, out_dist, out_count);
    return cudaGetLastError();
}

}  // namespace hvs
