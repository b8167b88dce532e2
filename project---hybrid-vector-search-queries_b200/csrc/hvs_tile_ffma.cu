// hvs_tile_ffma.cu -- K2: the exact-path distance sweep.  FP32 FFMA register tiles, TMA-staged
// data tiles, fused per-query top-K filter; the distance matrix never leaves the SM.
//
// Replaces, for up to 128 queries at a time, the reference's candidate loop + dist_to_query +
// Knn::check_add (include/optimized.hpp:84-117, include/optimized_impl.h:54-170, :284-335).  One CTA
// takes one work item of the planner: <= 128 queries that all want (part of) the arena rows
// [row_begin,row_end).  Rows of an arena are contiguous 400-byte records, so a tile of 128 rows is
// ONE 51,200-byte block that the TMA engine moves global->shared with a single 1-D bulk copy
// (cp.async.bulk + mbarrier complete_tx, double buffered).  256 threads hold an 8x8 register tile
// each (128 rows x 128 queries per CTA tile) and run the 100-deep contraction as FFMA on 128-bit
// shared-memory operands: 16 LDS.128 feed 256 FFMA.
//
// Scores are s = ||x||^2 - 2 q.x (the query-constant ||q||^2 is dropped).  A score survives when it
// is below the query's running threshold  s_(100) + margin ; survivors go to a 32-slot shared-memory
// buffer per query and are merged by one warp into the query's sorted candidate list (<= 256 entries,
// L2-resident) when the buffer fills.  `margin` bounds the rounding difference between this kernel's
// arithmetic and the reference's (hvs_margin.cuh), so the list provably contains the reference's
// top-100; K5 (hvs_finalize.cu) re-ranks it with the reference's own arithmetic.
//
// Roofline: FP32 FFMA.  200 algorithmic flop per (query,row) pair; HBM traffic is 400 B per row per
// 128 queries (~3 B per pair), i.e. ~64 flop/B -- far on the compute side of the machine balance.
#include "hvs_engine.h"
#include "hvs_margin.cuh"
#include "hvs_topk.cuh"

namespace hvs {

namespace {

constexpr int TQ = QT;       // queries per CTA tile
constexpr int TR = 128;      // rows per CTA tile
constexpr int NT = 256;      // threads
constexpr int NWARP = NT / 32;
constexpr int CB = 32;       // per-query survivor buffer (entries)

struct TileSmem {
    alignas(128) float x[2][TR * DIM];   // two row tiles, filled by the TMA engine
    alignas(16) float xn[2][TR];         // ||x||^2 of the tile rows
    alignas(16) float q[TQ * DIM];       // the item's query vectors, [query][dim]
    uint64_t buf[TQ][CB];                // survivors waiting to be merged
    float thr[TQ];                       // accept when s < thr
    float margin[TQ];
    uint32_t bcnt[TQ];                   // entries pushed to buf (may exceed CB: the excess is retried)
    uint32_t lcnt[TQ];                   // entries in the query's candidate list
    uint32_t qlo[TQ], qhi[TQ];           // rows of this item that belong to the query's slice
    uint32_t qid[TQ];
    alignas(8) uint64_t bar[2];
    uint32_t need[3];                    // "some buffer is at least half full", rotating per tile
    uint32_t lo_max, hi_min;             // rows [lo_max,hi_min) are wanted by every query of the item
};

// Merge the survivor buffer of query slot `qq` into its sorted candidate list.  One warp.
__device__ __forceinline__ void merge_slot(TileSmem &S, const TileItem &it, int qq, uint64_t *__restrict__ cand,
                                           uint32_t *__restrict__ gthr, uint32_t *__restrict__ flags, int lane)
{
    const uint32_t nb = min(S.bcnt[qq], (uint32_t)CB);
    const uint32_t nl = S.lcnt[qq];
    uint64_t *L = cand + (size_t)(it.out_off + qq) * KOUT;
    const uint64_t nk = (uint32_t)lane < nb ? S.buf[qq][lane] : KEY_INF;
    float lim;
    bool ovf;
    const uint32_t keep = warp_merge_list<KOUT>(L, nl, nk, nb, S.margin[qq], lim, ovf, lane);
    if (lane == 0) {
        S.lcnt[qq] = keep;
        S.bcnt[qq] = 0;
        const uint32_t qid = S.qid[qq];
        if (ovf) flags[qid] = 1u;          // more rows inside the margin than the list holds: K4 re-solves exactly
        // thresholds are shared by every CTA that works on this query (other row chunks)
        const float mine = nextafterf(lim, __int_as_float(0x7f800000));
        const float theirs = okey_inv(ld_relaxed_u32(&gthr[qid]));
        if (mine < theirs) atomicMin(&gthr[qid], okey(mine));
        S.thr[qq] = fminf(S.thr[qq], fminf(mine, theirs));
    }
    __syncwarp();
}

__device__ __forceinline__ void merge_pass(TileSmem &S, const TileItem &it, uint32_t min_fill, uint64_t *__restrict__ cand,
                                           uint32_t *__restrict__ gthr, uint32_t *__restrict__ flags, int warp, int lane)
{
    for (int qq = warp; qq < (int)it.nq; qq += NWARP)
        if (S.bcnt[qq] >= min_fill) merge_slot(S, it, qq, cand, gthr, flags, lane);
}

}  // namespace

__global__ void __launch_bounds__(NT, 1)
k_tile_ffma(const float *__restrict__ queries, const QSlice *__restrict__ slices, const TileItem *__restrict__ items,
            const uint32_t *__restrict__ item_q, Arena a0, Arena a1, uint32_t n_rows, float xnorm_max, float margin_scale,
            uint64_t *__restrict__ cand, uint32_t *__restrict__ cand_cnt, uint32_t *__restrict__ gthr,
            uint32_t *__restrict__ flags)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileSmem &S = *reinterpret_cast<TileSmem *>(smem_raw);
    const TileItem it = items[blockIdx.x];
    const Arena A = it.arena == ARENA_T ? a0 : a1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lx = lane & 7, ly = lane >> 3;     // 8 row lanes x 4 query lanes per warp
    const int wr = warp & 1, wq = warp >> 1;     // 2 x 4 warps: 64 rows x 32 queries each

    if (tid == 0) {
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
        mbar_fence_init();
        S.need[0] = S.need[1] = S.need[2] = 0;
        S.lo_max = 0; S.hi_min = 0xffffffffu;
    }
    __syncthreads();
    if (tid < TQ) {
        uint32_t lo = 1, hi = 0, id = 0xffffffffu;
        float mg = 0.f, th = __int_as_float(0xff800000);   // unused query slots never accept a row (nobody merges their buffers)
        if ((uint32_t)tid < it.nq) {
            id = item_q[it.q_off + tid];
            const QSlice sl = slices[id];
            lo = max(sl.begin, it.row_begin);
            hi = min(sl.end, it.row_end);
            mg = margin_ffma(sl.qnorm, xnorm_max) * margin_scale;
            th = okey_inv(ld_relaxed_u32(&gthr[id]));   // what other CTAs already know about this query
            atomicMax(&S.lo_max, lo);
            atomicMin(&S.hi_min, hi);
        }
        S.qid[tid] = id; S.qlo[tid] = lo; S.qhi[tid] = hi; S.margin[tid] = mg;
        S.thr[tid] = th;
        S.bcnt[tid] = 0; S.lcnt[tid] = 0;
    }
    for (int idx = tid; idx < TQ * (DIM / 4); idx += NT) {
        const int qq = idx / (DIM / 4), k4 = idx - qq * (DIM / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((uint32_t)qq < it.nq)
            v = reinterpret_cast<const float4 *>(queries + (size_t)item_q[it.q_off + qq] * QROW + 4)[k4];
        reinterpret_cast<float4 *>(S.q)[idx] = v;
    }

    // tiles start on a multiple of 4 rows so that the ||x||^2 copy is 16-byte aligned
    const uint32_t row0 = it.row_begin & ~3u;
    const uint32_t ntiles = (it.row_end - row0 + TR - 1) / TR;
    auto issue = [&](uint32_t t) {
        const uint32_t r = row0 + t * TR;
        const uint32_t rows = min((uint32_t)TR, n_rows - r);
        const uint32_t bx = rows * ROW_BYTES, bn = (rows * 4u + 15u) & ~15u;
        mbar_expect_tx(&S.bar[t & 1], bx + bn);
        bulk_g2s(S.x[t & 1], A.x + (size_t)r * DIM, bx, &S.bar[t & 1]);
        bulk_g2s(S.xn[t & 1], A.xnorm + r, bn, &S.bar[t & 1]);
    };
    if (tid == 0) {
        if (ntiles > 0) issue(0);
        if (ntiles > 1) issue(1);
    }
    __syncthreads();
    const uint32_t lo_max = S.lo_max, hi_min = S.hi_min;

    uint32_t qlo[8], qhi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { qlo[j] = S.qlo[wq * 32 + ly + 4 * j]; qhi[j] = S.qhi[wq * 32 + ly + 4 * j]; }

    for (uint32_t t = 0; t < ntiles; ++t) {
        const int st = t & 1;
        const uint32_t trow0 = row0 + t * TR;
        mbar_wait(&S.bar[st], (t >> 1) & 1);

        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        const float4 *xs = reinterpret_cast<const float4 *>(S.x[st] + (wr * 64 + lx) * DIM);
        const float4 *qs = reinterpret_cast<const float4 *>(S.q + (wq * 32 + ly) * DIM);
#pragma unroll 5
        for (int k4 = 0; k4 < DIM / 4; ++k4) {
            float4 xv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xv[i] = xs[i * 8 * (DIM / 4) + k4];      // rows lx + 8 i: conflict-free
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 qv = qs[j * 4 * (DIM / 4) + k4];                    // queries ly + 4 j
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i][j] = fmaf(xv[i].x, qv.x, acc[i][j]);
                    acc[i][j] = fmaf(xv[i].y, qv.y, acc[i][j]);
                    acc[i][j] = fmaf(xv[i].z, qv.z, acc[i][j]);
                    acc[i][j] = fmaf(xv[i].w, qv.w, acc[i][j]);
                }
            }
        }

        // ---- fused filter: s = ||x||^2 - 2 q.x against the query's threshold -------------------
        const bool interior = trow0 >= lo_max && trow0 + TR <= hi_min;
        uint64_t pend = 0;
        {
            float thr[8], xn[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) thr[j] = S.thr[wq * 32 + ly + 4 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i) xn[i] = S.xn[st][wr * 64 + lx + 8 * i];
            // scores first, branch-free; then ONE test per query (the minimum of its 8 rows in this thread): the
            // element-wise look happens only where that minimum beats the threshold (rare after the warm-up)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float m = __int_as_float(0x7f800000);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float s = fmaf(-2.f, acc[i][j], xn[i]);
                    acc[i][j] = s;
                    m = fminf(m, s);
                }
                if (m < thr[j]) {
                    const int qq = wq * 32 + ly + 4 * j;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float s = acc[i][j];
                        const uint32_t row = trow0 + wr * 64 + lx + 8 * i;
                        bool ok = s < thr[j];
                        if (!interior) ok = ok && row >= qlo[j] && row < qhi[j];
                        if (ok) {
                            const uint32_t slot = atomicAdd(&S.bcnt[qq], 1u);
                            if (slot < (uint32_t)CB) S.buf[qq][slot] = ((uint64_t)okey(s) << 32) | row;
                            else pend |= 1ull << (i * 8 + j);
                            if (slot >= (uint32_t)(CB / 2 - 1)) S.need[t % 3] = 1u;
                        }
                    }
                }
            }
        }
        __syncthreads();                                   // tile consumed; pushes visible
        if (tid == 0) {
            S.need[(t + 2) % 3] = 0;
            if (t + 2 < ntiles) issue(t + 2);
        }
        if (S.need[t % 3]) {                               // uniform: read after the barrier
            uint32_t min_fill = CB / 2;
            for (;;) {
                merge_pass(S, it, min_fill, cand, gthr, flags, warp, lane);
                __syncthreads();
                if (pend) {                                // retry the pushes that found the buffer full
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (pend & (1ull << (i * 8 + j))) {
                                const int qq = wq * 32 + ly + 4 * j;
                                const float s = acc[i][j];
                                if (s < S.thr[qq]) {
                                    const uint32_t slot = atomicAdd(&S.bcnt[qq], 1u);
                                    if (slot < (uint32_t)CB) {
                                        S.buf[qq][slot] = ((uint64_t)okey(s) << 32) | (trow0 + wr * 64 + lx + 8 * i);
                                        pend &= ~(1ull << (i * 8 + j));
                                    }
                                } else {
                                    pend &= ~(1ull << (i * 8 + j));
                                }
                            }
                }
                if (!__syncthreads_or(pend != 0)) break;
                min_fill = CB;                             // only the buffers that are full again
            }
        }
    }
    __syncthreads();
    merge_pass(S, it, 1u, cand, gthr, flags, warp, lane);
    __syncthreads();
    if ((uint32_t)tid < it.nq) cand_cnt[it.out_off + tid] = S.lcnt[tid];
}

cudaError_t tile_ffma_init_attributes()
{
    return cudaFuncSetAttribute(k_tile_ffma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
}

cudaError_t launch_tile_ffma(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev, const TileItem *items_dev,
                             uint32_t item_begin, uint32_t n_items, const uint32_t *item_q_dev, uint64_t *cand_dev,
                             uint32_t *cand_cnt_dev, uint32_t *gthr_dev, uint32_t *flags_dev, float margin_scale)
{
    if (!n_items) return cudaSuccess;
    const int smem = (int)sizeof(TileSmem);
    const Index &ix = e->index;
    k_tile_ffma<<<n_items, NT, smem, e->stream>>>(queries_dev, slices_dev, items_dev + item_begin, item_q_dev, ix.arena(0),
                                                   ix.arena(1), ix.n, ix.xnorm_max, margin_scale, cand_dev, cand_cnt_dev,
                                                   gthr_dev, flags_dev);
    return cudaGetLastError();
}

}  // namespace hvs
