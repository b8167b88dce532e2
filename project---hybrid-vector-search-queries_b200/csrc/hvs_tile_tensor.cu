// hvs_tile_tensor.cu -- K3: tcgen05 candidate pass for slices that queries share.
//
// Same job as K2 (hvs_tile_ffma.cu) -- replace the reference's candidate loop + dist_to_query +
// Knn::check_add (include/optimized.hpp:84-117, include/optimized_impl.h:54-170, :284-335) for a
// whole tile of queries at once -- but the 100-deep contraction runs on the 5th-generation tensor
// cores:   D[q][row] = sum_k A[q][k] * B[row][k]   with FP16 operands and FP32 accumulation in TMEM,
//   A[q]   = [ fp16(-2 sx q_0..99), 1, 1, 1, 0.. ]          (built in shared memory per work item)
//   B[row] = [ fp16(sx x_0..99), nh, nm, nl, 0.. ]           (nh+nm+nl = sx^2 ||x||^2 split into 3 fp16)
// so the accumulator IS the score  sx^2 (||x||^2 - 2 q.x)  and the epilogue is a pure min/compare.
// sx is a power of two chosen at index-build time so that every operand is inside fp16's range.
// B comes from an FP16 image of each arena written at index-build time in the tensor core's
// canonical K-major no-swizzle layout (8-row x 16-byte core matrices), so a stage of 128 rows is
// one contiguous 28,672-byte block moved by ONE 1-D TMA bulk copy -- no tensor map, no swizzle.
//
// Persistent kernel, one CTA per SM, work items handed out dynamically (one atomicAdd per item, longest
// first).  An item = up to 256 queries
// (two M=128 halves sharing every B stage) x a run of arena rows.  Warp roles: warp 0 TMA producer,
// warp 1 / warp 10 MMA issuers of query half 0 / 1 (independent pipelines; warp 1 owns TMEM),
// warps 2..9 epilogue (warp w reads TMEM lanes 32*(w%4)..; one thread = one query).  Four
// accumulators of 128 columns (2 halves x 2 buffers) use all 512 TMEM columns, so the epilogue of
// tile t overlaps the MMAs of tile t+1.
//
// The result is NOT approximate: a row survives when its fp16 score is within `margin_tensor`
// (hvs_margin.cuh: a rigorous bound on the rounding of both operands) of the running 100-th best,
// and K5 re-ranks all survivors with the reference's fp32 arithmetic.  Survivors are appended to a
// per-(CTA, query) pool in global memory (L2 resident, transposed so that the 32 lanes of a warp walk
// their 32 pools with coalesced loads) by the thread that owns the query.  When a pool fills, all 32
// queries of the warp are compacted together, lane-parallel: each lane brackets (select_probe: log-count
// interpolation on the counting function) a score with ~100 entries at or below it, keeps what is within
// the margin of it and tightens its threshold -- no sorting; the warps of a query half compact together
// (shared epoch), because they hand accumulators back together.  At the end of an item every lane folds its pool
// into its query's GLOBAL list of best scores (all row chunks, per-query try-lock), so a CTA that
// later sweeps another chunk of the slice starts with (nearly) the final threshold (`gthr`).
//
// Roofline: tensor pipe -- 2 x 7 MMAs (M128 N128 K16) = 896 tensor cycles per 128-row stage per SM;
// the epilogue must read the 128 KB of fp32 accumulators of a stage out of TMEM, and while the tensor
// core accumulates those reads take its TMEM port (measured: MMA + TMA alone 906 cycles per stage, with the
// reads 1030; tcgen05.ld alone moves 750-980 B/clk/SM, tools/tmem_bw.cu); the B stream is 28,672 B per
// stage per SM.  HVS_K3_STATS=1 prints the cycle budget per warp role, the hit statistics of the scan and
// the slowest CTAs; DESIGN.md section 9 has the numbers.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <cuda_fp16.h>

#include "hvs_engine.h"
#include "hvs_margin.cuh"
#include "hvs_topk.cuh"

namespace hvs {

namespace {

constexpr int KP = 112;               // padded contraction length (7 MMA k-steps of 16)
constexpr int KU = KP / 8;            // 16-byte units per row
constexpr int GROUP_B = KU * 128;     // bytes per 8-row group of the image (1792)
constexpr int ROW_B = KP * 2;         // bytes per row (224)
constexpr int TN = 128;               // data rows per stage == MMA N
constexpr int NST = 6;                // B stages in flight
constexpr int STAGE_B = TN * ROW_B;   // 28,672
constexpr int A_B = 128 * ROW_B;      // one query half
#ifndef HVS_K3_ACC64
#define HVS_K3_ACC64 0                // 1: accumulator hand-off in 64-row halves of a stage (four buffers per query half) instead of 128-row
                                      // stages (two).  Measured in round 2 (interleaved A/B, headline workload): correct, but the sweep takes
                                      // 36.7 ms instead of 33.9 -- twice the MMA instructions re-read the A operand twice as often and the
                                      // barrier traffic doubles, which costs more than the finer hand-off wins.  Kept as a build option.
#endif
#ifndef HVS_K3_LD2
#define HVS_K3_LD2 0
#endif
constexpr int NACC = HVS_K3_ACC64 ? 4 : 2;     // accumulator buffers per query half
constexpr int ACCN = HVS_K3_ACC64 ? 64 : 128;  // data rows (TMEM columns) per accumulator == MMA N
constexpr int NTHR = 352;              // warp 0 TMA, warp 1 MMA issuer of query half 0, warps 2..9 epilogue, warp 10 MMA issuer of half 1
constexpr int POOL = TENSOR_POOL;     // survivor pool entries per query (global memory)
constexpr uint32_t FULL = 0xffffffffu;
constexpr int GB = TENSOR_GBEST;      // per-query global list of best scores over all finished chunks
constexpr int GPL = GB / 32;          // list scores per lane when a warp holds one query's list in registers
static_assert(GB % 32 == 0 && GB >= 128 && GB <= 512, "global list size");
constexpr int SPARSE_LANES = 6;
constexpr int SPARSE_SEL = 8;         // merge_global: up to this many lanes whose global list needs a new K-th best are handled one query at a time       // compaction: up to this many participating lanes are handled one query at a time by the whole warp

static_assert(QT_TENSOR == 256, "two M=128 halves");
static_assert(POOL % 64 == 0 && POOL >= 512 && POOL <= 1024 && KOUT <= POOL - 32, "pool must take 32 more survivors after a compaction; passes read it in batches of 64");
constexpr int PPL = POOL / 32;        // pool entries per lane when a whole warp holds ONE query's pool in registers

struct TensorSmem {
    alignas(128) unsigned char b[NST][STAGE_B];
    alignas(128) unsigned char a[2][A_B];
    alignas(8) uint64_t full[NST], empty[NST];
    alignas(8) uint64_t tfull[2][NACC], tempty[2][NACC];
    uint32_t tmem_base;
    uint32_t next_item;     // dynamic work distribution: the item this CTA sweeps next
    uint32_t cepoch[2];     // per query half: bumped by an epilogue warp that starts a compaction (the others join it)
};

// ---- tcgen05 wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M=128, N=128, K=16, f16 x f16 -> f32
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes);
// LBO = byte distance between the two 16-byte k-units of one K=16 step, SBO = between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    uint64_t d = (uint64_t)((addr >> 4) & 0x3fffu);
    d |= (uint64_t)((128u >> 4) & 0x3fffu) << 16;              // LBO
    d |= (uint64_t)(((uint32_t)GROUP_B >> 4) & 0x3fffu) << 32;   // SBO
    d |= 1ull << 46;                                           // descriptor version (sm_100)
    return d;                                                  // layout_type 0 = SWIZZLE_NONE, base_offset 0
}
// c=F32 (bit 4), a=b=F16 (format 0 at bits 7,10), K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(ACCN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t p;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));     // one FMNMX3
    return d;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// per-query epilogue state, owned by one thread for the whole item
struct QState {
    float thr, margin;
    uint32_t cnt, qlo, qhi, qid;
};

// Pools are transposed: entry i of the query owned by lane l lives at pool_warp[i * 32 + l], so that the
// 32 lanes of a warp walk their 32 pools in lock step with fully coalesced loads.
//
// Lane-parallel compaction: every lane that holds at least K survivors searches, privately, for a score v
// with  K <= #{entries <= v} <= K + 8  (any v with at least K entries at or below it bounds the final K-th
// best; the search brackets the rank with select_probe, a bisection step every third pass), then
// rewrites its pool keeping only the entries within `margin` of v and tightens its threshold.
// One call serves up to 32 queries; no sorting, no shuffles, a handful of coalesced passes over the pools.
__device__ __noinline__ uint2 compact_warp(uint32_t cnt, float thr, float margin, uint32_t qid, bool valid,
                                           uint64_t *__restrict__ pool_warp, uint32_t *__restrict__ gthr,
                                           uint32_t *__restrict__ flags, uint32_t keep_cap, uint32_t part_min, int lane)
{
    // lanes that take part: those whose pool is at least part_min full, and those that have no threshold yet.  A lane
    // with a threshold and a half-empty pool gains little from a new cut, and leaving it out is what keeps most
    // compactions on the cheap one-query-at-a-time path below (a warp whose queries begin at different rows --
    // type-1/3 items -- would otherwise stream all 32 pools every time ONE of them fills).
    const bool part = valid && cnt >= (uint32_t)K && (cnt >= part_min || thr == __int_as_float(0x7f800000));
    const uint32_t maxc = __reduce_max_sync(FULL, part ? cnt : 0u);
    if (maxc == 0) return make_uint2(cnt, __float_as_uint(thr));
    // Few lanes take part (queries whose slices start at different rows warm up one after the other: type-1/3 items,
    // the first chunks of type-2 slices): the lane-parallel passes below would stream all 32 pools for them.  Instead
    // the WARP takes one such query at a time: its <= 512 entries go to registers (16 per lane, one round trip), the
    // rank is bracketed with warp-wide counts -- no further memory passes -- and the kept entries are written back.
    const uint32_t pmask = __ballot_sync(FULL, part);
    if (__popc(pmask) <= SPARSE_LANES) {
        uint32_t my_cnt = cnt;
        float my_thr = thr;
        for (uint32_t mm = pmask; mm; mm &= mm - 1u) {
            const int src = __ffs((int)mm) - 1;
            const uint32_t c = __shfl_sync(FULL, cnt, src), q_src = __shfl_sync(FULL, qid, src);
            const float thr_src = __shfl_sync(FULL, thr, src), margin_src = __shfl_sync(FULL, margin, src);
            uint64_t e[PPL];
            uint32_t klo = 0xffffffffu, khi = 0u;
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                const uint32_t idx = (uint32_t)lane + 32u * j;
                e[j] = idx < c ? __ldcg(pool_warp + (size_t)32 * idx + src) : ~0ull;
            }
#pragma unroll
            for (int j = 0; j < PPL; ++j)
                if ((uint32_t)lane + 32u * j < c) { klo = min(klo, (uint32_t)(e[j] >> 32)); khi = max(khi, (uint32_t)(e[j] >> 32)); }
            klo = __reduce_min_sync(FULL, klo);
            khi = __reduce_max_sync(FULL, khi);
            if (klo != khi) {                                         // #{<= klo} = clo < K <= chi = #{<= khi}
                uint32_t clo = 0u, chi = c;
                klo -= 1u;
                for (int it = 0; it < 48 && khi - klo > 1u; ++it) {
                    const uint32_t mid = select_probe(klo, khi, clo, chi, it);
                    uint32_t n = 0;
#pragma unroll
                    for (int j = 0; j < PPL; ++j) n += ((uint32_t)lane + 32u * j < c && (uint32_t)(e[j] >> 32) <= mid) ? 1u : 0u;
                    n = __reduce_add_sync(FULL, n);
                    if (n >= (uint32_t)K) { khi = mid; chi = n; if (n <= (uint32_t)K + 8u) break; }
                    else { klo = mid; clo = n; }
                }
            }
            const float lim = okey_inv(khi) + margin_src;
            const uint32_t limk = okey(lim);
            uint32_t kc = 0;
#pragma unroll
            for (int j = 0; j < PPL; ++j) kc += ((uint32_t)lane + 32u * j < c && (uint32_t)(e[j] >> 32) <= limk) ? 1u : 0u;
            uint32_t pos = kc;                                        // inclusive scan over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, pos, o); if (lane >= o) pos += v; }
            uint32_t w = __shfl_sync(FULL, pos, 31);
            pos -= kc;
            __syncwarp();                                             // every lane holds its share: the pool may be rewritten
#pragma unroll
            for (int j = 0; j < PPL; ++j)
                if ((uint32_t)lane + 32u * j < c && (uint32_t)(e[j] >> 32) <= limk) {
                    if (pos < keep_cap) pool_warp[(size_t)32 * pos + src] = e[j];
                    ++pos;
                }
            if (w > keep_cap) {                                       // more rows inside the margin than a list may hold
                if (lane == 0) flags[q_src] = 1u;
                w = keep_cap;
            }
            const float mine = nextafterf(lim, __int_as_float(0x7f800000));
            const float theirs = okey_inv(ld_relaxed_u32(&gthr[q_src]));
            if (lane == 0 && mine < theirs) atomicMin(&gthr[q_src], okey(mine));
            if (lane == src) { my_cnt = w; my_thr = fminf(thr_src, fminf(mine, theirs)); }
        }
        __syncwarp();
        return make_uint2(my_cnt, __float_as_uint(my_thr));
    }
    const uint32_t *sc = reinterpret_cast<const uint32_t *>(pool_warp) + 2 * lane + 1;   // score word of entry i: sc[64 i]
    // First pass: range of the scores, and the count at a first probe -- the bound the previous compaction left
    // (thr - margin still has at least K entries at or below it; with no threshold yet the probe counts nothing).
    const uint32_t g0 = thr < __int_as_float(0x7f800000) ? okey(thr - margin) : 0u;
    uint32_t klo = 0xffffffffu, khi = 0u, cg = 0u;
    for (uint32_t i0 = 0; i0 < maxc; i0 += 32) {                      // 32 independent loads in flight per lane
        uint32_t k[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) k[j] = __ldcg(sc + (size_t)64 * (i0 + j));
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (part && i0 + j < cnt) { klo = min(klo, k[j]); khi = max(khi, k[j]); cg += k[j] <= g0 ? 1u : 0u; }
    }
    // invariant: #{<= klo} = clo < K <= chi = #{<= khi}
    bool done = !part || klo == khi;
    uint32_t clo = 0, chi = cnt;
    if (!done) {
        klo -= 1;
        if (g0 > klo && g0 < khi) {
            if (cg >= (uint32_t)K) { khi = g0; chi = cg; done = cg <= (uint32_t)K + 8u; }
            else { klo = g0; clo = cg; }
        }
    }
    for (int it = 0; it < 40 && !__all_sync(FULL, done); ++it) {
        const uint32_t mid = select_probe(klo, khi, clo, chi, it);
        uint32_t c = 0;
        for (uint32_t i0 = 0; i0 < maxc; i0 += 64) {                  // 64 independent loads in flight per lane
            uint32_t k[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) k[j] = __ldcg(sc + (size_t)64 * (i0 + j));
#pragma unroll
            for (int j = 0; j < 64; ++j) c += (i0 + j < cnt && k[j] <= mid) ? 1u : 0u;
        }
        if (!done) {
            if (c >= (uint32_t)K) { khi = mid; chi = c; done = c <= (uint32_t)K + 8u; }
            else { klo = mid; clo = c; }
            if (khi - klo <= 1u) done = true;
        }
    }
    // keep what is within the margin of that bound
    float lim = __int_as_float(0x7f800000);
    uint32_t limk = 0xffffffffu;
    if (part) { lim = okey_inv(khi) + margin; limk = okey(lim); }
    uint32_t w = 0;
    for (uint32_t i0 = 0; i0 < maxc; i0 += 32) {
        uint64_t e[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) e[j] = __ldcg(pool_warp + (size_t)32 * (i0 + j) + lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint32_t kj = (uint32_t)(e[j] >> 32);
            if (part && i0 + j < cnt && kj <= limk) { pool_warp[(size_t)32 * w + lane] = e[j]; ++w; }
        }
    }
    if (part) {
        if (w > keep_cap) {                                           // more rows inside the margin than a list may hold:
            flags[qid] = 1u;                                          // K4 re-solves this query exactly
            w = keep_cap;
        }
        cnt = w;
        const float mine = nextafterf(lim, __int_as_float(0x7f800000));
        const float theirs = okey_inv(ld_relaxed_u32(&gthr[qid]));
        if (mine < theirs) atomicMin(&gthr[qid], okey(mine));
        thr = fminf(thr, fminf(mine, theirs));
    }
    __syncwarp();
    return make_uint2(cnt, __float_as_uint(thr));
}

// Lane-parallel, item end: every lane folds the scores of its pool (the survivors of THIS row chunk) into
// its query's global list of best scores -- at most GB scores of distinct rows, any order, guarded by a
// per-query try-lock -- and derives a threshold from the K-th best score over ALL chunks finished so far.
// This is what lets a CTA that sweeps one chunk of a slice filter with (nearly) the final threshold.
// Chunks are disjoint and every item contributes exactly once, so no row is ever counted twice.
// gcnt[q] = number of scores in the list, gcut[q] = key of its K-th best (0xffffffff while it holds < K).
__device__ __noinline__ float merge_global(uint32_t cnt, float thr, float margin, uint32_t qid, bool valid,
                                           const uint64_t *__restrict__ pool_warp, uint32_t *__restrict__ gbest,
                                           uint32_t *__restrict__ gcnt, uint32_t *__restrict__ gcut,
                                           uint32_t *__restrict__ glock, uint32_t *__restrict__ gthr, bool sparse_ok, int lane)
{
    const uint32_t *sc = reinterpret_cast<const uint32_t *>(pool_warp) + 2 * lane + 1;   // score word of pool entry i: sc[64 i]
    bool want = valid && cnt > 0;
    // nothing to tell if no survivor of this chunk beats the list's current K-th best
    {
        const uint32_t cutk = want ? ld_relaxed_u32(&gcut[qid]) : 0u;
        const uint32_t maxc0 = __reduce_max_sync(FULL, want ? cnt : 0u);
        bool better = false;
        for (uint32_t i0 = 0; i0 < maxc0; i0 += 32) {
            uint32_t k[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) k[j] = __ldcg(sc + (size_t)64 * (i0 + j));
#pragma unroll
            for (int j = 0; j < 32; ++j) better |= (i0 + j < cnt && k[j] < cutk);
        }
        want = want && better;
    }
    // Try-lock only: a lane that cannot get its query's lock skips the merge (the list only sharpens
    // thresholds, it never decides results).  Blocking here could deadlock: the lanes of a warp hold 32
    // different locks at once and reconverge before releasing them.
    bool part = false;
    for (int attempt = 0; attempt < 4 && __any_sync(FULL, want && !part); ++attempt) {
        if (want && !part) part = atomicCAS(&glock[qid], 0u, 1u) == 0u;
        if (attempt) __nanosleep(64);
    }
    if (!__any_sync(FULL, part)) return thr;
    uint32_t *G = gbest + (size_t)qid * GB;
    uint32_t gn = 0, cut_old = 0xffffffffu;
    if (part) {
        __threadfence();
        gn = ld_relaxed_u32(&gcnt[qid]);
        cut_old = ld_relaxed_u32(&gcut[qid]);
    }
    const uint32_t maxc = __reduce_max_sync(FULL, part ? cnt : 0u);
    // Cheap case: the survivors that beat the list's K-th best still fit into the list -> append them and keep
    // the bound (it then stands for the (K + a few)-th best, which is still a valid, slightly lazier bound).
    // Only when the list would overflow is a new K-th best selected.
    uint32_t nbeat = 0;
    for (uint32_t i0 = 0; i0 < maxc; i0 += 32) {
        uint32_t k[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) k[j] = __ldcg(sc + (size_t)64 * (i0 + j));
#pragma unroll
        for (int j = 0; j < 32; ++j) nbeat += (part && i0 + j < cnt && k[j] < cut_old) ? 1u : 0u;
    }
    const bool need_sel = part && (gn + nbeat > (uint32_t)GB || (cut_old == 0xffffffffu && gn + nbeat >= (uint32_t)K));
    const uint32_t selmask = __ballot_sync(FULL, need_sel);
    const bool sparse_sel = sparse_ok && selmask != 0u && __popc(selmask) <= SPARSE_SEL && maxc <= 256u;
    if (selmask == 0u || sparse_sel) {
        // cheap case for the lanes whose list still has room: append the survivors that beat the bound
        const bool app = part && !need_sel;
        if (__any_sync(FULL, app)) {
            uint32_t w = gn;
            for (uint32_t i0 = 0; i0 < maxc; i0 += 32) {
                uint32_t k[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) k[j] = __ldcg(sc + (size_t)64 * (i0 + j));
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (app && i0 + j < cnt && k[j] < cut_old) { G[w] = k[j]; ++w; }
            }
            if (app) {
                gcnt[qid] = w;
                __threadfence();
                atomicExch(&glock[qid], 0u);
            }
            __syncwarp();
        }
        // Few lanes need a new K-th best (their list would overflow): the WARP takes one such query at a time -- list
        // (<= 128 scores) and pool (<= 256) go to registers in one round trip, the rank is bracketed with warp-wide
        // counts, the scores at or below the new bound are written back.  The lane-parallel passes further down cost
        // the same for one lane as for 32 (every pass is a chain of L2 round trips over all 32 pools).
        float my_thr = thr;
        for (uint32_t mm = sparse_sel ? selmask : 0u; mm; mm &= mm - 1u) {
            const int src = __ffs((int)mm) - 1;
            const uint32_t c = __shfl_sync(FULL, cnt, src), g_n = __shfl_sync(FULL, gn, src), q_src = __shfl_sync(FULL, qid, src);
            const float margin_src = __shfl_sync(FULL, margin, src), thr_src = __shfl_sync(FULL, thr, src);
            uint32_t *Gs = gbest + (size_t)q_src * GB;
            const uint32_t *scs = reinterpret_cast<const uint32_t *>(pool_warp) + 2 * src + 1;
            uint32_t e[GPL + 8];                                             // GB/32 list scores + 8 pool scores per lane; 0xffffffff = none
#pragma unroll
            for (int j = 0; j < GPL; ++j) { const uint32_t i = (uint32_t)lane + 32u * j; e[j] = i < g_n ? __ldcg(Gs + i) : 0xffffffffu; }
#pragma unroll
            for (int j = 0; j < 8; ++j) { const uint32_t i = (uint32_t)lane + 32u * j; e[GPL + j] = i < c ? __ldcg(scs + (size_t)64 * i) : 0xffffffffu; }
            uint32_t klo = 0xffffffffu, khi = 0u;
#pragma unroll
            for (int j = 0; j < GPL + 8; ++j)
                if (e[j] != 0xffffffffu) { klo = min(klo, e[j]); khi = max(khi, e[j]); }
            klo = __reduce_min_sync(FULL, klo);
            khi = __reduce_max_sync(FULL, khi);
            if (klo != khi && klo != 0xffffffffu) {                          // #{<= klo} = clo < K <= chi = #{<= khi}
                uint32_t clo = 0u, chi = g_n + c;
                klo -= 1u;
                for (int it = 0; it < 48 && khi - klo > 1u; ++it) {
                    const uint32_t mid = select_probe(klo, khi, clo, chi, it);
                    uint32_t n = 0;
#pragma unroll
                    for (int j = 0; j < GPL + 8; ++j) n += (e[j] <= mid) ? 1u : 0u;      // 0xffffffff never counts: mid < khi <= 0xfffffffe
                    n = __reduce_add_sync(FULL, n);
                    if (n >= (uint32_t)K) { khi = mid; chi = n; if (n <= (uint32_t)K + 8u) break; }
                    else { klo = mid; clo = n; }
                }
            }
            uint32_t kc = 0;
#pragma unroll
            for (int j = 0; j < GPL + 8; ++j) kc += (e[j] != 0xffffffffu && e[j] <= khi) ? 1u : 0u;
            uint32_t pos = kc;                                               // inclusive scan over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, pos, o); if (lane >= o) pos += v; }
            const uint32_t total = __shfl_sync(FULL, pos, 31);
            pos -= kc;
            __syncwarp();                                                    // every lane holds its share: the list may be rewritten
#pragma unroll
            for (int j = 0; j < GPL + 8; ++j)
                if (e[j] != 0xffffffffu && e[j] <= khi) { if (pos < (uint32_t)GB) Gs[pos] = e[j]; ++pos; }
            __threadfence();
            __syncwarp();
            const float mine = nextafterf(okey_inv(khi) + margin_src, __int_as_float(0x7f800000));
            if (lane == 0) {
                gcnt[q_src] = min(total, (uint32_t)GB);
                gcut[q_src] = khi;
                __threadfence();
                atomicExch(&glock[q_src], 0u);
                atomicMin(&gthr[q_src], okey(mine));
            }
            if (lane == src) my_thr = fminf(thr_src, mine);
        }
        __syncwarp();
        return my_thr;
    }
    const uint32_t maxg = __reduce_max_sync(FULL, gn);                 // <= GB
    const bool sel = part && gn + cnt >= (uint32_t)K;
    uint32_t klo = 0xffffffffu, khi = 0u;
    for (uint32_t i0 = 0; i0 < maxg; i0 += 16) {
        uint32_t k[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) k[j] = (i0 + j < gn) ? __ldcg(G + i0 + j) : 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (i0 + j < gn) { klo = min(klo, k[j]); khi = max(khi, k[j]); }
    }
    for (uint32_t i0 = 0; i0 < maxc; i0 += 32) {
        uint32_t k[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) k[j] = __ldcg(sc + (size_t)64 * (i0 + j));
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (part && i0 + j < cnt) { klo = min(klo, k[j]); khi = max(khi, k[j]); }
    }
    bool done = !sel || klo == khi;
    uint32_t clo = 0, chi = gn + cnt;
    if (!done) klo -= 1;
    for (int it = 0; it < 40 && !__all_sync(FULL, done); ++it) {
        const uint32_t mid = select_probe(klo, khi, clo, chi, it);
        uint32_t c = 0;
        for (uint32_t i0 = 0; i0 < maxg; i0 += 16) {
            uint32_t k[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) k[j] = (i0 + j < gn) ? __ldcg(G + i0 + j) : 0xffffffffu;
#pragma unroll
            for (int j = 0; j < 16; ++j) c += (i0 + j < gn && k[j] <= mid) ? 1u : 0u;
        }
        for (uint32_t i0 = 0; i0 < maxc; i0 += 32) {
            uint32_t k[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) k[j] = __ldcg(sc + (size_t)64 * (i0 + j));
#pragma unroll
            for (int j = 0; j < 32; ++j) c += (i0 + j < cnt && k[j] <= mid) ? 1u : 0u;
        }
        if (!done) {
            if (c >= (uint32_t)K) { khi = mid; chi = c; done = c <= (uint32_t)K + 8u; }
            else { klo = mid; clo = c; }
            if (khi - klo <= 1u) done = true;
        }
    }
    // new global list: with a bound, the scores at or below it (at most GB of them); without, everything
    const uint32_t cut = sel ? khi : 0xffffffffu;
    uint32_t w = 0;
    for (uint32_t i0 = 0; i0 < maxg; i0 += 16) {                       // read a batch, then write: w never overtakes i0
        uint32_t k[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) k[j] = (i0 + j < gn) ? __ldcg(G + i0 + j) : 0xffffffffu;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (i0 + j < gn && k[j] <= cut) { G[w] = k[j]; ++w; }
    }
    for (uint32_t i0 = 0; i0 < maxc; i0 += 32) {
        uint32_t k[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) k[j] = __ldcg(sc + (size_t)64 * (i0 + j));
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (part && i0 + j < cnt && k[j] <= cut && w < (uint32_t)GB) { G[w] = k[j]; ++w; }
    }
    if (part) {
        gcnt[qid] = w;
        if (sel) gcut[qid] = khi;
        __threadfence();
        atomicExch(&glock[qid], 0u);
        if (sel) {
            const float mine = nextafterf(okey_inv(khi) + margin, __int_as_float(0x7f800000));
            const float theirs = okey_inv(ld_relaxed_u32(&gthr[qid]));
            if (mine < theirs) atomicMin(&gthr[qid], okey(mine));
            thr = fminf(thr, fminf(mine, theirs));
        }
    }
    __syncwarp();
    return thr;
}

}  // namespace

// ---- FP16 image of an arena ------------------------------------------------------------------------
// element (row, k) lives at  (row>>3)*GROUP_B + (k>>3)*128 + (row&7)*16 + (k&7)*2
__global__ void k_build_image(const float *__restrict__ x, const float *__restrict__ xnorm, uint32_t n, uint32_t n_img, float sx,
                              unsigned char *__restrict__ img)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_img * KU) return;
    const uint32_t row = (uint32_t)(i / KU), u = (uint32_t)(i % KU);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // norm outliers (K0 set their ||x||^2 to +inf): a zero vector with the largest norm term fp16 holds -- the score is
    // 60000 for every query, far above any inlier's (<= 32000 + 2 * 60000 / 2 ... see margin_tensor), so they never
    // crowd a pool; K5 scores them exactly whatever happens here
    const bool outlier = row < n && !(xnorm[row] < __int_as_float(0x7f800000));
    if (outlier) {
        if (u == 12) v[4] = 60000.f;
    } else if (row < n) {
        if (u < 12) {
            const float4 a = *reinterpret_cast<const float4 *>(x + (size_t)row * DIM + 8 * u);
            const float4 b = *reinterpret_cast<const float4 *>(x + (size_t)row * DIM + 8 * u + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else if (u == 12) {
            const float4 a = *reinterpret_cast<const float4 *>(x + (size_t)row * DIM + 96);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= sx;                         // power of two: exact
        if (u == 12) {
            const float xn = xnorm[row] * sx * sx;                      // <= 32000 by the choice of sx
            const float nh = __half2float(__float2half_rn(xn));
            const float r1 = xn - nh;                                   // exact
            const float nm = __half2float(__float2half_rn(r1));
            const float nl = r1 - nm;                                   // exact; its fp16 rounding is far below fp32 ulp of xn
            v[4] = nh; v[5] = nm; v[6] = nl;
        }
    }
    __half2 p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4 *>(img + (size_t)(row >> 3) * GROUP_B + u * 128 + (row & 7) * 16) = *reinterpret_cast<uint4 *>(p);
}

cudaError_t build_tensor_image(hvs_engine *e, int a)
{
    Index &ix = e->index;
    const uint32_t n_img = ((ix.n + 7u) & ~7u) + 2 * TN;        // every stage copy stays inside the image
    cudaError_t c = ix.xb[a].ensure((size_t)n_img * ROW_B);
    if (c != cudaSuccess) return c;
    const size_t total = (size_t)n_img * KU;
    k_build_image<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(ix.x[a].as<float>(), ix.xnorm[a].as<float>(), ix.n, n_img,
                                                                          ix.img_scale, ix.xb[a].as<unsigned char>());
    return cudaGetLastError();
}

// ---- the sweep --------------------------------------------------------------------------------------
// STATS: the instrumented build (HVS_K3_STATS / HVS_K3_DBG): cycle counters per warp role, hit statistics, the measurement-only
// modes.  The production instantiation carries none of it (the clock reads alone were ~10 of ~200 instructions per stage).
#define TICK() (STATS ? clock64() : 0ll)
template <bool STATS>
__global__ void __launch_bounds__(NTHR, 1)
k_tile_tensor(const float *__restrict__ queries, const QSlice *__restrict__ slices, const TileItem *__restrict__ items,
              uint32_t n_items, const uint32_t *__restrict__ item_q, const unsigned char *__restrict__ img0,
              const unsigned char *__restrict__ img1, float xnorm_max, float sx, uint64_t *__restrict__ pool,
              uint64_t *__restrict__ cand, uint32_t *__restrict__ cand_cnt, uint32_t *__restrict__ gthr,
              uint32_t *__restrict__ gbest, uint32_t *__restrict__ gcnt, uint32_t *__restrict__ gcut,
              uint32_t *__restrict__ glock, uint32_t *__restrict__ flags, uint32_t *__restrict__ work_counter, int dbg_arg,
              uint32_t knobs, unsigned long long *__restrict__ kstat_arg)
{
    constexpr bool PIPE = true;                           // the pipelined, unrolled stage scan (the only one kept)
    const int dbg = STATS ? dbg_arg : 0;
    unsigned long long *const kstat = STATS ? kstat_arg : nullptr;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TensorSmem &S = *reinterpret_cast<TensorSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t part_min = knobs & 0xffffu;            // compaction: pool fill from which a lane takes part
    const bool sparse_sel_ok = (knobs & 0x10000u) == 0u;  // merge_global: one-query-at-a-time selection when few lanes need one
    const bool half_mma = STATS && (knobs & 0x20000u) != 0u;
    const bool l2_hints = (knobs & 0x40000u) != 0u;       // image stages evict_last, candidate lists evict_first
    const uint32_t trig = (knobs >> 20) & 0x3ffu;         // pool fill that starts a compaction

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 2); }   // both MMA issuers release a stage
        for (int h = 0; h < 2; ++h)
            for (int b = 0; b < NACC; ++b) { mbar_init(&S.tfull[h][b], 1); mbar_init(&S.tempty[h][b], 4); }
        S.cepoch[0] = 0; S.cepoch[1] = 0;
        S.next_item = atomicAdd(work_counter, 1u);
        mbar_fence_init();
    }
    if (warp == 1) {                                                  // TMEM: all 512 columns, this warp owns them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    long long c_mergeonly = 0, c_abuild = 0;
    long long c_wait = 0, c_scan = 0, c_compact = 0, c_merge = 0, c_mma_full = 0, c_mma_tempty = 0, c_items = 0;
    unsigned n_items_done = 0, n_compact = 0, n_surv = 0, n_lhit = 0, n_whit = 0, n_infhit = 0, n_chunks = 0;
    const long long c_start = TICK();
    // running counters: the mbarrier phases continue across items
    uint32_t gt = 0;           // stages issued / consumed so far (producer, MMA)
    uint32_t ga[2] = {0, 0};   // accumulator uses so far, per half (MMA, epilogue)

    // Items are handed out dynamically, longest first (the planner sorts them): a CTA that draws expensive items
    // (queries still warming up their thresholds) simply takes fewer of them.
    for (;;) {
        const uint32_t item = S.next_item;
        if (item >= n_items) break;
        const TileItem it = items[item];
        const unsigned char *img = it.arena == ARENA_T ? img0 : img1;
        const int nhalf = it.nq > 128u ? 2 : 1;
        const uint32_t row0 = it.row_begin & ~7u;                     // stages start on an 8-row group
        const uint32_t ntiles = (it.row_end - row0 + TN - 1) / TN;

        const long long ta0 = TICK();
        // A operand: fp16(-2 sx q) | 1 1 1 | 0, written straight into the canonical layout.  All MMAs of the
        // previous item have retired (its epilogue consumed every accumulator before the barrier below).
        for (int idx = tid; idx < QT_TENSOR * KU; idx += NTHR) {
            const int qs = idx / KU, u = idx - qs * KU;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if ((uint32_t)qs < it.nq) {
                const float *qv = queries + (size_t)item_q[it.q_off + qs] * QROW + 4;
                if (u < 12) {
                    const float4 a = *reinterpret_cast<const float4 *>(qv + 8 * u);
                    const float4 b = *reinterpret_cast<const float4 *>(qv + 8 * u + 4);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else if (u == 12) {
                    const float4 a = *reinterpret_cast<const float4 *>(qv + 96);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] *= -2.f * sx;
                if (u == 12) { v[4] = 1.f; v[5] = 1.f; v[6] = 1.f; }
            }
            __half2 p[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) p[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            const int r = qs & 127;
            *reinterpret_cast<uint4 *>(S.a[qs >> 7] + (r >> 3) * GROUP_B + u * 128 + (r & 7) * 16) = *reinterpret_cast<uint4 *>(p);
        }
        fence_proxy_async();                                          // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        c_abuild += TICK() - ta0;
        uint32_t item_next = 0;
        if (tid == 0) item_next = atomicAdd(work_counter, 1u);        // everybody has read next_item; the answer is awaited at the end of the item

        if (warp == 0) {
            // ===== TMA producer (whole warp runs the loop, one elected lane talks to the TMA engine) =====
            const unsigned char *src = img + (size_t)(row0 >> 3) * GROUP_B;
            const uint64_t pol = l2_policy_evict_last();             // the other CTAs sweeping this chunk read the same stages shortly after
            for (uint32_t t = 0; t < ntiles; ++t) {
                const uint32_t g = gt + t;
                const int st = g % NST;
                mbar_wait(&S.empty[st], ((g / NST) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&S.full[st], STAGE_B);
                    if (l2_hints) bulk_g2s_hint(S.b[st], src + (size_t)t * STAGE_B, STAGE_B, &S.full[st], pol);
                    else bulk_g2s(S.b[st], src + (size_t)t * STAGE_B, STAGE_B, &S.full[st]);
                }
                __syncwarp();
            }
            if (tid == 0) S.next_item = item_next;
        } else if (warp == 1 || warp == 10) {
            // ===== MMA issuers: one warp per query half, warp-uniform control flow, one elected lane issues =====
            // Accumulator (half h, buffer b) = TMEM columns [128 (2h+b), +128): the epilogue drains buffer b of a
            // half while the tensor core fills buffer b^1 with the next stage.  The two halves are independent
            // pipelines (own issuer, own barriers): a slow epilogue warp in one half never delays the other.
            const int h = warp == 1 ? 0 : 1;
            const uint64_t adesc = smem_desc(smem_u32(S.a[h]));
            for (uint32_t t = 0; t < ntiles; ++t) {
                const uint32_t g = gt + t;
                const int st = g % NST;
                { const long long t0 = TICK(); mbar_wait(&S.full[st], (g / NST) & 1); c_mma_full += TICK() - t0; }
                tc_fence_after();
                if (h < nhalf) {
                    const uint64_t bdesc = smem_desc(smem_u32(S.b[st]));
#if HVS_K3_ACC64
                    // two 64-row accumulators per stage: the epilogue gets the first half of the stage while the tensor core
                    // works on the second, and with four buffers per query half a warp that stops for a hit or a compaction
                    // holds the others up two sub-stages later, not at the next one
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const uint32_t u = 2u * (ga[h] + t) + (uint32_t)sub;
                        const int b = u & 3;
                        { const long long t0 = TICK(); mbar_wait(&S.tempty[h][b], ((u >> 2) & 1) ^ 1); c_mma_tempty += TICK() - t0; }
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d = tmem + (uint32_t)(h * 4 + b) * ACCN;
                            const uint64_t bsub = bdesc + (uint64_t)(sub * (ACCN / 8) * GROUP_B / 16);   // rows 64.. of the stage: 8 row groups further
#pragma unroll
                            for (int j = 0; j < KP / 16; ++j)            // one k-step = two 16-byte units = 256 bytes
                                tc_mma(d, adesc + (uint64_t)(j * 16), bsub + (uint64_t)(j * 16), IDESC, j > 0);
                            tc_commit(&S.tfull[h][b]);
                        }
                        __syncwarp();
                    }
#else
                    const uint32_t u = ga[h] + t;
                    const int b = u & 1;
                    { const long long t0 = TICK(); mbar_wait(&S.tempty[h][b], ((u >> 1) & 1) ^ 1); c_mma_tempty += TICK() - t0; }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d = tmem + (uint32_t)(h * 2 + b) * TN;
#pragma unroll
                        for (int j = 0; j < KP / 16; ++j) {              // one k-step = two 16-byte units = 256 bytes
                            // HVS_K3_HALF_MMA=1 (measurement only, results are wrong): 4 of the 7 k-steps -- what a first pass at
                            // twice the tensor rate (kind::f8f6f4 operands) would cost the tensor pipe
                            if (half_mma && j >= 4) continue;
                            tc_mma(d, adesc + (uint64_t)(j * 16), bdesc + (uint64_t)(j * 16), IDESC, j > 0);
                        }
                        tc_commit(&S.tfull[h][b]);
                    }
                    __syncwarp();
#endif
                }
                if (elect_one()) tc_commit(&S.empty[st]);               // this half is done with the stage once its MMAs retire
                __syncwarp();
            }
        } else {
            // ===== epilogue: one thread = one query =====
            const int ew = warp - 2, h = ew >> 2, quad = warp & 3;
            if (h < nhalf) {
                const uint32_t qslot0 = (uint32_t)(h * 128 + quad * 32);
                const uint32_t qslot = qslot0 + lane;
                uint64_t *pool_warp = pool + ((size_t)blockIdx.x * QT_TENSOR + qslot0) * POOL;
                uint64_t *mypool = pool_warp + lane;                     // entry i at mypool[32 i]
                QState st;
                st.cnt = 0; st.qid = 0; st.qlo = 1; st.qhi = 0; st.margin = 0.f;
                st.thr = __int_as_float(0xff800000);                     // unused slot: nothing passes
                if (qslot < it.nq) {
                    st.qid = item_q[it.q_off + qslot];
                    const QSlice sl = slices[st.qid];
                    st.qlo = max(sl.begin, it.row_begin);
                    st.qhi = min(sl.end, it.row_end);
                    st.margin = margin_tensor(sl.qnorm, xnorm_max, sx) * sx * sx;
                    st.thr = okey_inv(ld_relaxed_u32(&gthr[st.qid]));
                    // -2 sx q must be representable in fp16: otherwise this query cannot use the tensor path
                    if (!(2.f * sx * sqrtf(sl.qnorm) < 60000.f)) { st.thr = __int_as_float(0xff800000); flags[st.qid] = 1u; }
                }
                if (dbg == 4 && qslot < it.nq) st.thr = 0.f;        // measurement only: (almost) nothing survives
                if (dbg == 5 && qslot < it.nq) st.thr = 3550.f;     // measurement only: steady-state-like survival rate ~1e-4
                const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
                // 32 columns (data rows) at a time: four 8-wide minima, one compare and one warp vote; only when some
                // lane of the warp has a group whose minimum beats its threshold is that group looked at element by
                // element -- under warp-uniform branches, so lanes with hits in different groups do not serialise.
                // thr_s = the lane's threshold for THIS stage: -inf while the stage lies outside the query's rows
                // (a query whose slice starts later in the chunk must not send the warp down the slow path).
                float thr_s = st.thr;
                auto scan = [&](const uint32_t (&r)[32], uint32_t rbase) {
                    float g[4];                                       // four independent chains of 3-input minima (FMNMX3): 18 ops per 32 columns
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        float mm = fmin3(__uint_as_float(r[8 * q4 + 0]), __uint_as_float(r[8 * q4 + 1]), __uint_as_float(r[8 * q4 + 2]));
                        mm = fmin3(mm, __uint_as_float(r[8 * q4 + 3]), __uint_as_float(r[8 * q4 + 4]));
                        mm = fmin3(mm, __uint_as_float(r[8 * q4 + 5]), __uint_as_float(r[8 * q4 + 6]));
                        g[q4] = fminf(mm, __uint_as_float(r[8 * q4 + 7]));
                    }
                    const bool hit = fminf(fmin3(g[0], g[1], g[2]), g[3]) < thr_s;
                    const bool whit = __any_sync(FULL, hit);
                    if (STATS && kstat) {                             // HVS_K3_STATS: how often the element-wise path runs
                        ++n_chunks;
                        n_lhit += hit ? 1u : 0u;
                        n_infhit += (hit && st.thr == __int_as_float(0x7f800000)) ? 1u : 0u;
                        n_whit += whit ? 1u : 0u;
                    }
                    if (whit) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            if (__any_sync(FULL, g[q4] < thr_s)) {
#pragma unroll
                                for (int c = 8 * q4; c < 8 * q4 + 8; ++c) {
                                    const float s = __uint_as_float(r[c]);
                                    const uint32_t row = rbase + c;
                                    if (s < thr_s && row >= st.qlo && row < st.qhi)
                                        mypool[(size_t)32 * st.cnt++] = ((uint64_t)okey(s) << 32) | row;
                                }
                            }
                    }
                };
                // room for 32 more survivors in every pool of this warp; when one pool is nearly full (or a query has
                // no threshold yet) all 32 queries of the warp are compacted together
                // The four warps of a half hand accumulators back together, so a warp that stops to compact stalls the
                // other three.  They therefore compact TOGETHER: the warp that has to (a pool nearly full, or a query
                // without a threshold yet) bumps the half's epoch, and the others join as soon as they see it if they
                // hold enough survivors for it to be worthwhile -- the stalls overlap instead of adding up.
                uint32_t my_ep = *reinterpret_cast<volatile uint32_t *>(&S.cepoch[h]);
                auto make_room = [&](uint32_t slack) {
                    if (dbg >= 3) { if (st.cnt > (uint32_t)POOL - slack) st.cnt = 0; return; }   // measurement only
                    const uint32_t ep = *reinterpret_cast<volatile uint32_t *>(&S.cepoch[h]);
                    // trigger: the pool must keep room for a stage (POOL - slack); a lower mark (`trig`) compacts earlier, which
                    // costs more (cheap, one-query-at-a-time) compactions but lets a warming-up threshold tighten in smaller
                    // steps: between compactions a lane admits D = trig - ~100 rows while the rows it has seen grow by the
                    // factor 1 + D/100, i.e. D / ln(1 + D/100) survivors per e-fold of rows (211 at D = 284, 144 at D = 100)
                    const bool full = st.cnt > min((uint32_t)POOL - slack, trig) || (st.cnt >= 192u && st.thr == __int_as_float(0x7f800000));
                    bool own, join = false;
                    if (ep == my_ep) {                                // the common case costs one vote
                        own = __any_sync(FULL, full);
                        if (!own) return;
                        if (lane == 0) atomicAdd(&S.cepoch[h], 1u);
                        my_ep = ep + 1;
                    } else {                                          // another warp of this half started a compaction: join if worthwhile
                        own = __any_sync(FULL, full);
                        join = __any_sync(FULL, st.cnt >= max(160u, part_min));
                        my_ep = ep;
                    }
                    if (own || join) {
                        const long long t0 = TICK();
                        if (STATS) n_surv += st.cnt;
                        const uint2 o = compact_warp(st.cnt, st.thr, st.margin, st.qid, qslot < it.nq, pool_warp, gthr, flags, KOUT, part_min, lane);
                        st.cnt = o.x; st.thr = __uint_as_float(o.y);
                        if (STATS) { n_surv -= st.cnt; ++n_compact; }
                        c_compact += TICK() - t0;
                    }
                };
                uint32_t gpre = 0xff800000u;                             // okey(+inf)
                for (uint32_t t = 0; t < ntiles; ++t) {
                    const uint32_t u = ga[h] + t;
                    // a look at what other CTAs found out about this query: the load is issued one stage before its use
                    if ((t & 7) == 7) st.thr = fminf(st.thr, okey_inv(gpre));
                    if ((t & 7) == 6 && qslot < it.nq && dbg < 3) gpre = ld_relaxed_u32(&gthr[st.qid]);
                    const uint32_t trow0 = row0 + t * TN;
#if HVS_K3_ACC64
                    {
                        const uint32_t u0 = 2u * u, u1 = 2u * u + 1u;
                        const int b0 = u0 & 3, b1 = u1 & 3;
                        const uint32_t tc0 = tlane + (uint32_t)(h * 4 + b0) * ACCN, tc1 = tlane + (uint32_t)(h * 4 + b1) * ACCN;
                        make_room(128u);
                        thr_s = (trow0 < st.qhi && trow0 + TN > st.qlo) ? st.thr : __int_as_float(0xff800000);
                        { const long long t0 = TICK(); mbar_wait(&S.tfull[h][b0], (u0 >> 2) & 1); c_wait += TICK() - t0; }
                        const long long ts0 = TICK();
                        tc_fence_after();
                        uint32_t ra[32], rb[32];
                        tmem_ld32(tc0, ra);
                        tmem_wait_ld();
                        tmem_ld32(tc0 + 32, rb);
                        scan(ra, trow0);
                        tmem_wait_ld();
                        // the second accumulator is usually complete by now (the tensor core runs ahead): start its first
                        // load before the last scan of the first; only if it is not, wait for it afterwards
                        const bool early = __all_sync(FULL, mbar_try_wait(&S.tfull[h][b1], (u1 >> 2) & 1));   // made warp-uniform: "not yet" is always safe
                        if (early) { tc_fence_after(); tmem_ld32(tc1, ra); }
                        scan(rb, trow0 + 32);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&S.tempty[h][b0]);
                        if (!early) {
                            const long long t0 = TICK(); mbar_wait(&S.tfull[h][b1], (u1 >> 2) & 1); c_wait += TICK() - t0;
                            tc_fence_after();
                            tmem_ld32(tc1, ra);
                        }
                        tmem_wait_ld();
                        tmem_ld32(tc1 + 32, rb);
                        scan(ra, trow0 + 64);
                        tmem_wait_ld();
                        scan(rb, trow0 + 96);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&S.tempty[h][b1]);
                        c_scan += TICK() - ts0;
                    }
                    if (false) {
                        const int b = 0; const uint32_t tcol = 0; const long long ts0 = 0;
#else
                    const int b = u & 1;
                    { const long long t0 = TICK(); mbar_wait(&S.tfull[h][b], (u >> 1) & 1); c_wait += TICK() - t0; }
                    const long long ts0 = TICK();
                    tc_fence_after();
                    const uint32_t tcol = tlane + (uint32_t)(h * 2 + b) * TN;
                    {
#endif
                    if (PIPE) {
                        // whole stage unrolled: column offsets are immediates, one room check per stage (128 slots),
                        // the TMEM load of the next 32 columns is in flight while these are scanned
                        make_room(128u);
                        thr_s = (trow0 < st.qhi && trow0 + TN > st.qlo) ? st.thr : __int_as_float(0xff800000);
                        uint32_t ra[32], rb[32];
#if HVS_K3_LD2
                        // two loads in flight, two waits per stage (tcgen05.wait::ld waits for every outstanding load)
                        tmem_ld32(tcol, ra);
                        tmem_ld32(tcol + 32, rb);
                        tmem_wait_ld();
                        scan(ra, trow0);
                        tmem_ld32(tcol + 64, ra);
                        scan(rb, trow0 + 32);
                        tmem_ld32(tcol + 96, rb);
                        tmem_wait_ld();
                        scan(ra, trow0 + 64);
                        scan(rb, trow0 + 96);
#else
                        tmem_ld32(tcol, ra);
                        tmem_wait_ld();
                        tmem_ld32(tcol + 32, rb);
                        scan(ra, trow0);
                        tmem_wait_ld();
                        tmem_ld32(tcol + 64, ra);
                        scan(rb, trow0 + 32);
                        tmem_wait_ld();
                        tmem_ld32(tcol + 96, rb);
                        scan(ra, trow0 + 64);
                        tmem_wait_ld();
                        scan(rb, trow0 + 96);
#endif
                    } else if (dbg == 0) {
#pragma unroll 1
                        for (int c4 = 0; c4 < TN / 32; ++c4) {
                            uint32_t r[32];
                            tmem_ld32(tcol + c4 * 32, r);
                            make_room(32u);
                            thr_s = (trow0 < st.qhi && trow0 + TN > st.qlo) ? st.thr : __int_as_float(0xff800000);
                            tmem_wait_ld();
                            scan(r, trow0 + c4 * 32);
                        }
                    } else if (dbg == 2) {          // measurement only: TMEM loads, no scan (results are wrong)
#pragma unroll 1
                        for (int c4 = 0; c4 < TN / 32; ++c4) {
                            uint32_t r[32];
                            tmem_ld32(tcol + c4 * 32, r);
                            tmem_wait_ld();
                            uint32_t x = 0;
#pragma unroll
                            for (int c = 0; c < 32; ++c) x ^= r[c];
                            if (x == 0x12345678u) st.cnt++;
                        }
                    }                               // dbg == 1: measurement only, accumulators are not even read
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.tempty[h][b]);
                    c_scan += TICK() - ts0;
                    }
                }
                // hand the pools to K5: lists need not be sorted, only short enough
                if (__any_sync(FULL, st.cnt > (uint32_t)KOUT)) {
                    const uint2 o = compact_warp(st.cnt, st.thr, st.margin, st.qid, qslot < it.nq, pool_warp, gthr, flags, KOUT, min(part_min, (uint32_t)KOUT), lane);
                    st.cnt = o.x; st.thr = __uint_as_float(o.y);
                }
                // tell the other chunks of these queries what this chunk found (worth it only for long sweeps)
                const long long tm0 = TICK();
                if (STATS) n_surv += st.cnt;
                if (ntiles >= 128u && dbg == 0)
                    st.thr = merge_global(st.cnt, st.thr, st.margin, st.qid, qslot < it.nq, pool_warp, gbest, gcnt, gcut, glock, gthr, sparse_sel_ok, lane);
                c_mergeonly += TICK() - tm0;
                {
                    const uint32_t maxc = __reduce_max_sync(FULL, st.cnt);
                    uint64_t *L = cand + (size_t)(it.out_off + qslot) * KOUT;
                    const uint64_t polw = l2_policy_evict_first();    // written once, read by K5 after the whole sweep: do not displace the image
                    // only what is still below the query's global bound goes to K5 (which starts from the same bound and would
                    // skip the rest anyway): the lists of a long slice's early chunks shrink, and so do K5's reads
                    const uint32_t gk = (qslot < it.nq && dbg == 0) ? ld_relaxed_u32(&gthr[st.qid]) : 0xffffffffu;
                    uint32_t w = 0;
                    for (uint32_t i0 = 0; i0 < maxc; i0 += 16) {      // batches of 16 loads in flight
                        uint64_t e[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) e[j] = __ldcg(mypool + (size_t)32 * (i0 + j));
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (i0 + j < st.cnt && (uint32_t)(e[j] >> 32) < gk) {
                                if (l2_hints) st_u64_hint(L + w, e[j], polw); else L[w] = e[j];
                                ++w;
                            }
                    }
                    if (qslot < it.nq) cand_cnt[it.out_off + qslot] = w;
                }
                c_merge += TICK() - tm0;
            }
        }
        { const long long t0 = TICK();
        gt += ntiles;
        ga[0] += ntiles;
        if (nhalf == 2) ga[1] += ntiles;
        tc_fence_before();
        __syncthreads();                                              // item boundary: A may be rebuilt
        tc_fence_after();
        c_items += TICK() - t0; }
        ++n_items_done;
    }
    if (STATS && kstat) {
        const long long c_total = TICK() - c_start;
        if (warp >= 2 && warp < 10) {
            n_surv = __reduce_add_sync(FULL, n_surv);
            n_lhit = __reduce_add_sync(FULL, n_lhit);
            n_infhit = __reduce_add_sync(FULL, n_infhit);
            if (lane == 0) {
                atomicMax(&kstat[20], (unsigned long long)c_total);
                if (warp == 2) {                                  // per-CTA record (first epilogue warp): who is slow, and why
                    unsigned long long *r = kstat + 32 + 8 * blockIdx.x;
                    r[0] = (unsigned long long)c_total; r[1] = n_items_done; r[2] = (unsigned long long)c_scan; r[3] = (unsigned long long)c_compact;
                    r[4] = n_compact; r[5] = (unsigned long long)c_wait; r[6] = (unsigned long long)c_merge; r[7] = (unsigned long long)c_items;
                }
                atomicAdd(&kstat[16], (unsigned long long)n_chunks); atomicAdd(&kstat[17], (unsigned long long)n_whit);
                atomicAdd(&kstat[18], (unsigned long long)n_lhit); atomicAdd(&kstat[19], (unsigned long long)n_infhit);
                atomicAdd(&kstat[0], (unsigned long long)c_total); atomicAdd(&kstat[1], (unsigned long long)c_wait);
                atomicAdd(&kstat[2], (unsigned long long)c_scan); atomicAdd(&kstat[3], (unsigned long long)c_compact);
                atomicAdd(&kstat[4], (unsigned long long)c_merge); atomicAdd(&kstat[5], (unsigned long long)c_items);
                atomicAdd(&kstat[6], (unsigned long long)n_compact); atomicAdd(&kstat[7], (unsigned long long)n_surv);
                atomicAdd(&kstat[8], 1ull); atomicAdd(&kstat[13], (unsigned long long)c_mergeonly); atomicAdd(&kstat[14], (unsigned long long)c_abuild);
            }
        } else if ((warp == 1 || warp == 10) && lane == 0) {
            atomicAdd(&kstat[9], (unsigned long long)c_total); atomicAdd(&kstat[10], (unsigned long long)c_mma_full);
            atomicAdd(&kstat[11], (unsigned long long)c_mma_tempty); atomicAdd(&kstat[12], 1ull);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// once per solve, before the first K3 launch: survivor pools and the per-query global lists
cudaError_t tile_tensor_begin(hvs_engine *e)
{
    cudaError_t c = e->d_pool.ensure((size_t)2 * e->sm_count * QT_TENSOR * POOL * 8);   // two pool sets: launches on the two lanes overlap
    if (c != cudaSuccess) return c;
    c = e->d_work_counter.ensure(64 * 4);                                                            // one item counter per launch of this solve
    if (c != cudaSuccess) return c;
    c = cudaMemsetAsync(e->d_work_counter.p, 0, 64 * 4, e->stream);
    if (c != cudaSuccess) return c;
    e->work_slot = 0;
    c = e->d_gbest.ensure((size_t)e->stats.m * GB * 4);
    if (c != cudaSuccess) return c;
    c = e->d_glock.ensure((size_t)e->stats.m * 12);                                                  // [m] locks, [m] counts, [m] K-th keys
    if (c != cudaSuccess) return c;
    c = cudaMemsetAsync(e->d_glock.p, 0, (size_t)e->stats.m * 8, e->stream);
    if (c != cudaSuccess) return c;
    c = cudaMemsetAsync(e->d_glock.as<uint32_t>() + 2 * (size_t)e->stats.m, 0xff, (size_t)e->stats.m * 4, e->stream);
    if (c != cudaSuccess) return c;
    return cudaSuccess;
}

cudaError_t tile_tensor_init_attributes()
{
    cudaError_t c = cudaFuncSetAttribute(k_tile_tensor<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TensorSmem));
    if (c == cudaSuccess) c = cudaFuncSetAttribute(k_tile_tensor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TensorSmem));
    return c;
}

cudaError_t launch_tile_tensor(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev, const TileItem *items_dev,
                               uint32_t item_begin, uint32_t n_items, const uint32_t *item_q_dev, uint64_t *cand_dev,
                               uint32_t *cand_cnt_dev, uint32_t *gthr_dev, uint32_t *flags_dev)
{
    if (!n_items) return cudaSuccess;
    const int smem = (int)sizeof(TensorSmem);
    const Index &ix = e->index;
    const uint32_t grid = n_items < (uint32_t)e->sm_count ? n_items : (uint32_t)e->sm_count;
    cudaError_t c = cudaSuccess;
    if (e->work_slot >= 63) return cudaErrorInvalidValue;
    static const bool want_stats = [] { const char *v = getenv("HVS_K3_STATS"); return v && v[0] != 0 && v[0] != '0'; }();
    unsigned long long *kstat = nullptr;
    if (want_stats) {
        c = e->d_scratch.ensure(256 + 64 * 256);
        if (c != cudaSuccess) return c;
        cudaMemsetAsync(e->d_scratch.p, 0, 256 + 64 * 256, e->stream);
        kstat = e->d_scratch.as<unsigned long long>();
    }
    static const int dbg = [] { const char *v = getenv("HVS_K3_DBG"); return v ? atoi(v) : 0; }();
    static const uint32_t knobs = [] {
        const char *v = getenv("HVS_K3_PARTMIN");
        int k = v ? atoi(v) : 0;
        uint32_t kn = (uint32_t)(k >= K && k <= POOL - 128 ? k : 256);     // measured: 256 takes the (C,T) head group from 8.9 to 5.8 Mcycles per warp
        const char *ss = getenv("HVS_K3_SPARSE_SEL");
        if (ss && ss[0] == '0') kn |= 0x10000u;
        const char *tg = getenv("HVS_K3_TRIG");
        int tv = tg ? atoi(tg) : 0;
        if (!(tv >= 128 && tv <= POOL - 128)) tv = POOL - 128;
        kn |= (uint32_t)tv << 20;
        if ((kn & 0xffffu) > (uint32_t)tv) kn = (kn & ~0xffffu) | (uint32_t)tv;       // lanes at the mark must take part
        const char *lh = getenv("HVS_K3_L2HINTS");
        if (lh && lh[0] == '1') kn |= 0x40000u;
        const char *hm = getenv("HVS_K3_HALF_MMA");
        if (hm && hm[0] == '1') kn |= 0x20000u;
        return kn;
    }();
    auto kern = (want_stats || dbg || (knobs & 0x20000u)) ? k_tile_tensor<true> : k_tile_tensor<false>;       // instrumented build only on request
    kern<<<grid, NTHR, smem, e->stream>>>(queries_dev, slices_dev, items_dev + item_begin, n_items, item_q_dev,
                                          ix.xb[0].as<unsigned char>(), ix.xb[1].as<unsigned char>(), ix.xnorm_max, ix.img_scale,
                                          e->d_pool.as<uint64_t>() + (size_t)e->pool_slot * e->sm_count * QT_TENSOR * POOL, cand_dev, cand_cnt_dev, gthr_dev, e->d_gbest.as<uint32_t>(),
                                          e->d_glock.as<uint32_t>() + e->stats.m, e->d_glock.as<uint32_t>() + 2 * (size_t)e->stats.m,
                                          e->d_glock.as<uint32_t>(), flags_dev, e->d_work_counter.as<uint32_t>() + e->work_slot, dbg, knobs, kstat);
    ++e->work_slot;
    if (kstat) {
        unsigned long long h[32 + 8 * 256];
        cudaStreamSynchronize(e->stream);
        cudaMemcpy(h, kstat, sizeof h, cudaMemcpyDeviceToHost);
        {   // the three slowest CTAs and the fastest one
            int idx[256];
            for (uint32_t i = 0; i < grid && i < 256; ++i) idx[i] = (int)i;
            std::sort(idx, idx + (grid < 256 ? grid : 256), [&](int a, int b) { return h[32 + 8 * a] > h[32 + 8 * b]; });
            for (uint32_t k = 0; k < grid && k < 256; ++k) {
                if (k >= 3 && k != grid - 1 && k != grid / 2) continue;
                const unsigned long long *r = h + 32 + 8 * idx[k];
                fprintf(stderr, "K3 CTA rank %u (block %d): total %.2f Mcycles, items %llu, scan %.2f (compact %.2f, n=%llu), wait_tfull %.2f, item_end %.2f, barrier %.2f\n",
                        k, idx[k], r[0] / 1e6, r[1], r[2] / 1e6, r[3] / 1e6, r[4], r[5] / 1e6, r[6] / 1e6, r[7] / 1e6);
            }
        }
        const double ew = (double)h[8], mw = (double)h[12];
        fprintf(stderr, "K3 hits: 32-column scans/warp %.0f, with a hit in the warp %.1f %%, lane hits per scan %.3f (with no threshold yet %.3f); slowest CTA %.2f Mcycles\n",
                h[16] / ew, 100.0 * h[17] / (double)(h[16] ? h[16] : 1), h[18] / (double)(h[16] ? h[16] : 1), h[19] / (double)(h[16] ? h[16] : 1), h[20] / 1e6);
        fprintf(stderr, "K3 stats per epilogue warp (Mcycles): total %.2f  wait_tfull %.2f  scan %.2f  (of which compact %.2f, n=%.1f)  item_end %.2f (merge_global %.2f)  A-build %.2f  barrier %.2f | survivors/warp %.0f | MMA warp: total %.2f wait_full %.2f wait_tempty %.2f\n",
                h[0] / ew / 1e6, h[1] / ew / 1e6, h[2] / ew / 1e6, h[3] / ew / 1e6, h[6] / ew, h[4] / ew / 1e6, h[13] / ew / 1e6, h[14] / ew / 1e6, h[5] / ew / 1e6, h[7] / ew,
                h[9] / mw / 1e6, h[10] / mw / 1e6, h[11] / mw / 1e6);
    }
    return cudaGetLastError();
}

}  // namespace hvs
