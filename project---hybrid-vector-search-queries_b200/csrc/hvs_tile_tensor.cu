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
// Persistent kernel, one CTA per SM, work items taken round-robin.  An item = up to 256 queries
// (two M=128 halves sharing every B stage) x a run of arena rows.  Warp roles: warp 0 TMA producer,
// warp 1 MMA issuer (+ TMEM owner), warps 2..9 epilogue (warp w reads TMEM lanes 32*(w%4)..; one
// thread = one query).  Four accumulators of 128 columns (2 halves x 2 buffers) use all 512 TMEM
// columns, so the epilogue of tile t overlaps the MMAs of tile t+1.
//
// The result is NOT approximate: a row survives when its fp16 score is within `margin_tensor`
// (hvs_margin.cuh: a rigorous bound on the rounding of both operands) of the running 100-th best,
// and K5 re-ranks all survivors with the reference's fp32 arithmetic.  Survivors are appended to a
// per-(CTA, query) pool in global memory (L2 resident) by the thread that owns the query; when a
// pool fills, one warp sorts it in registers, keeps what is still inside the margin and tightens
// the threshold.  Thresholds are shared between all CTAs working on the same query (`gthr`,
// atomicMin), so a CTA that starts late starts with a tight threshold.
//
// Roofline: tensor pipe -- 2 x 7 MMAs (M128 N128 K16) = 896 tensor cycles per 128-row stage per SM;
// the B stream is 28,672 B per stage per SM, read mostly from L2 (CTAs of one wave share the rows).
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
constexpr int NTHR = 320;
constexpr int POOL = TENSOR_POOL;     // survivor pool entries per query (global memory)
constexpr int PER = POOL / 32;        // pool entries per lane when a warp sorts it
constexpr uint32_t FULL = 0xffffffffu;
constexpr int GB = TENSOR_GBEST;      // per-query list of the best scores seen by ANY CTA (global memory)
constexpr uint32_t ROW_MASK = 0x7fffffffu, CONTRIB = 0x80000000u;   // top bit of a key's row word: "score already in gbest"
constexpr uint32_t NOKEY = 0xffffffffu;

static_assert(QT_TENSOR == 256, "two M=128 halves");
static_assert(POOL == 512 && KOUT <= POOL - 32, "pool must take 32 more survivors after a compaction");

struct TensorSmem {
    alignas(128) unsigned char b[NST][STAGE_B];
    alignas(128) unsigned char a[2][A_B];
    alignas(8) uint64_t full[NST], empty[NST];
    alignas(8) uint64_t tfull[2][2], tempty[2][2];
    uint32_t tmem_base;
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
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((128u >> 4) << 24);

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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// per-query epilogue state, owned by one thread for the whole item
struct QState {
    float thr, margin;
    uint32_t cnt, qlo, qhi, qid;
};

// Warp-cooperative: merge up to 32 scores (one per lane, NOKEY for none; distinct rows) into the query's
// global best-score list under the query's lock.  Returns the key of the list's K-th entry afterwards.
// This is what makes thresholds tight everywhere: the K-th best score over ALL rows any CTA has seen
// for this query bounds the final K-th best, whichever chunk of the slice a CTA is sweeping.
__device__ __noinline__ uint32_t contribute32(uint32_t *__restrict__ gbest_q, uint32_t *__restrict__ lock_q, uint32_t newkey, int lane)
{
    uint32_t key = newkey;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t other = __shfl_xor_sync(FULL, key, j);
            const bool take_min = ((lane & j) == 0) == ((lane & k) == 0);
            key = take_min ? min(key, other) : max(key, other);
        }
    const uint32_t nb = __popc(__ballot_sync(FULL, key != NOKEY));
    if (lane == 0) {
        while (atomicCAS(lock_q, 0u, 1u) != 0u) __nanosleep(64);
        __threadfence();
    }
    __syncwarp();
    uint32_t g[GB / 32], cl[GB / 32];
#pragma unroll
    for (int j = 0; j < GB / 32; ++j) { g[j] = ld_relaxed_u32(gbest_q + lane + 32 * j); cl[j] = 0; }
    uint32_t myrank = 0;
    for (uint32_t t = 0; t < nb; ++t) {
        const uint32_t nk = __shfl_sync(FULL, key, t);
        uint32_t c = 0;
#pragma unroll
        for (int j = 0; j < GB / 32; ++j) {
            const bool le = g[j] <= nk;       // equal scores: the list entry goes first
            c += le;
            cl[j] += !le;
        }
        c = __reduce_add_sync(FULL, c);
        if ((uint32_t)lane == t) myrank = c;
    }
    __syncwarp();
    uint32_t kth = 0;
#pragma unroll
    for (int j = 0; j < GB / 32; ++j) {
        const uint32_t pos = lane + 32 * j + cl[j];
        if (pos < (uint32_t)GB) {
            if (cl[j]) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(gbest_q + pos), "r"(g[j]) : "memory");
            if (pos == (uint32_t)(K - 1)) kth = g[j];
        }
    }
    if ((uint32_t)lane < nb) {
        const uint32_t pos = lane + myrank;
        if (pos < (uint32_t)GB) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(gbest_q + pos), "r"(key) : "memory");
            if (pos == (uint32_t)(K - 1)) kth = key;
        }
    }
    kth = __reduce_max_sync(FULL, kth);
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(lock_q, 0u);
    return kth;
}

// Warp-cooperative: sort lane `l`'s pool in registers (bitonic over 32 lanes x PER registers,
// element e = j*32 + lane), keep what is within the margin of the K-th best, hand the best scores that
// are news to the query's global list, tighten thresholds.
struct CompactOut { uint32_t cnt; float thr; };
__device__ __noinline__ CompactOut compact_lane(int l, uint32_t my_cnt, float my_thr, float my_margin, uint32_t my_qid,
                                                uint64_t *__restrict__ pool_warp, uint32_t *__restrict__ gthr,
                                                uint32_t *__restrict__ gbest, uint32_t *__restrict__ glock,
                                                uint32_t *__restrict__ flags, int lane)
{
    const uint32_t n = __shfl_sync(FULL, my_cnt, l);
    const float margin = __shfl_sync(FULL, my_margin, l);
    const uint32_t qid = __shfl_sync(FULL, my_qid, l);
    uint64_t *P = pool_warp + (size_t)l * POOL;
    uint64_t k[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t e = j * 32 + lane;
        k[j] = e < n ? P[e] : KEY_INF;
    }
#pragma unroll
    for (int size = 2; size <= POOL; size <<= 1) {
#pragma unroll
        for (int d = size >> 1; d > 0; d >>= 1) {
            if (d >= 32) {                                   // partner is another register of this lane
                const int dj = d >> 5;
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    if ((j & dj) == 0) {
                        const bool up = (((j * 32) & size) == 0);   // lane bits are below `size` here (size >= 64)
                        const uint64_t a = k[j], b = k[j | dj];
                        const bool sw = (a > b) == up;
                        k[j] = sw ? b : a;
                        k[j | dj] = sw ? a : b;
                    }
                }
            } else {                                         // partner is the same register of lane ^ d
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const uint64_t o = __shfl_xor_sync(FULL, k[j], d);
                    const uint32_t e = j * 32 + lane;
                    const bool up = (e & size) == 0;
                    const bool lower = (lane & d) == 0;
                    const bool take_min = lower == up;
                    k[j] = take_min ? (k[j] < o ? k[j] : o) : (k[j] < o ? o : k[j]);
                }
            }
        }
    }
    // sorted ascending by e = j*32 + lane
    uint32_t keep = n;
    float lim = __int_as_float(0x7f800000);
    bool ovf = false;
    if (n >= (uint32_t)K) {
        const uint64_t kk = __shfl_sync(FULL, k[(K - 1) >> 5], (K - 1) & 31);
        lim = okey_inv((uint32_t)(kk >> 32)) + margin;
        uint32_t nin = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if ((uint32_t)(j * 32 + lane) < n && okey_inv((uint32_t)(k[j] >> 32)) <= lim) ++nin;
        keep = __reduce_add_sync(FULL, nin);
        if (keep > (uint32_t)KOUT) { keep = KOUT; ovf = true; }   // more rows inside the margin than a list may hold
    }
    // the best GB entries that are not in the global list yet and would enter it
    uint32_t *gbest_q = gbest + (size_t)qid * GB;
    uint32_t gk = ld_relaxed_u32(gbest_q + GB - 1);               // worst score the global list still holds
    uint32_t kth = ld_relaxed_u32(gbest_q + K - 1);
#pragma unroll
    for (int j = 0; j < GB / 32; ++j) {
        const uint32_t sc = (uint32_t)(k[j] >> 32);
        const bool want = (uint32_t)(j * 32 + lane) < keep && !((uint32_t)k[j] & CONTRIB) && sc < gk;
        if (__any_sync(FULL, want)) {
            kth = contribute32(gbest_q, glock + qid, want ? sc : NOKEY, lane);
            if (want) k[j] |= (uint64_t)CONTRIB;
            gk = ld_relaxed_u32(gbest_q + GB - 1);
        }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t e = j * 32 + lane;
        if (e < keep) P[e] = k[j];
    }
    CompactOut out{my_cnt, my_thr};
    if (lane == l) {
        out.cnt = keep;
        if (ovf) flags[my_qid] = 1u;                          // K4 re-solves this query exactly
        const float glim = okey_inv(kth) + margin;            // +inf while the global list holds fewer than K scores
        const float mine = nextafterf(fminf(lim, glim), __int_as_float(0x7f800000));
        const float theirs = okey_inv(ld_relaxed_u32(&gthr[my_qid]));
        if (mine < theirs) atomicMin(&gthr[my_qid], okey(mine));
        out.thr = fminf(my_thr, fminf(mine, theirs));
    }
    __syncwarp();
    return out;
}

// Warp-cooperative, item end: tell the query's global list about pool entries it has not seen (no sort).
// Returns the (possibly tightened) threshold for lane `l`.
__device__ __noinline__ float contribute_pool(int l, uint32_t my_cnt, float my_thr, float my_margin, uint32_t my_qid,
                                              uint64_t *__restrict__ pool_warp, uint32_t *__restrict__ gthr,
                                              uint32_t *__restrict__ gbest, uint32_t *__restrict__ glock, int lane)
{
    const uint32_t n = __shfl_sync(FULL, my_cnt, l);
    const float margin = __shfl_sync(FULL, my_margin, l);
    const uint32_t qid = __shfl_sync(FULL, my_qid, l);
    uint64_t *P = pool_warp + (size_t)l * POOL;
    uint32_t *gbest_q = gbest + (size_t)qid * GB;
    uint32_t gk = ld_relaxed_u32(gbest_q + GB - 1);
    uint32_t kth = NOKEY;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t e = base + lane;
        const uint64_t kk = e < n ? P[e] : KEY_INF;
        const uint32_t sc = (uint32_t)(kk >> 32);
        const bool want = e < n && !((uint32_t)kk & CONTRIB) && sc < gk;
        if (__any_sync(FULL, want)) {
            kth = contribute32(gbest_q, glock + qid, want ? sc : NOKEY, lane);
            if (want) P[e] = kk | (uint64_t)CONTRIB;
            gk = ld_relaxed_u32(gbest_q + GB - 1);
        }
    }
    float thr = my_thr;
    if (lane == l && kth != NOKEY) {
        const float mine = nextafterf(okey_inv(kth) + margin, __int_as_float(0x7f800000));
        const float theirs = okey_inv(ld_relaxed_u32(&gthr[my_qid]));
        if (mine < theirs) atomicMin(&gthr[my_qid], okey(mine));
        thr = fminf(my_thr, fminf(mine, theirs));
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
    if (row < n) {
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

void build_tensor_image(hvs_engine *e, int a)
{
    Index &ix = e->index;
    const uint32_t n_img = ((ix.n + 7u) & ~7u) + 2 * TN;        // every stage copy stays inside the image
    if (ix.xb[a].ensure((size_t)n_img * ROW_B) != cudaSuccess) { cudaGetLastError(); ix.xb[a].release(); return; }
    const size_t total = (size_t)n_img * KU;
    k_build_image<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(ix.x[a].as<float>(), ix.xnorm[a].as<float>(), ix.n, n_img,
                                                                          ix.img_scale, ix.xb[a].as<unsigned char>());
}

bool tensor_path_available() { return true; }

// ---- the sweep --------------------------------------------------------------------------------------
template <bool PIPE>
__global__ void __launch_bounds__(NTHR, 1)
k_tile_tensor(const float *__restrict__ queries, const QSlice *__restrict__ slices, const TileItem *__restrict__ items,
              uint32_t n_items, const uint32_t *__restrict__ item_q, const unsigned char *__restrict__ img0,
              const unsigned char *__restrict__ img1, float xnorm_max, float sx, uint64_t *__restrict__ pool,
              uint64_t *__restrict__ cand, uint32_t *__restrict__ cand_cnt, uint32_t *__restrict__ gthr,
              uint32_t *__restrict__ gbest, uint32_t *__restrict__ glock, uint32_t *__restrict__ flags, int dbg)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TensorSmem &S = *reinterpret_cast<TensorSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
        for (int h = 0; h < 2; ++h)
            for (int b = 0; b < 2; ++b) { mbar_init(&S.tfull[h][b], 1); mbar_init(&S.tempty[h][b], 4); }
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

    // running counters: the mbarrier phases continue across items
    uint32_t gt = 0;           // stages issued / consumed so far (producer, MMA)
    uint32_t ga[2] = {0, 0};   // accumulator uses so far, per half (MMA, epilogue)

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const TileItem it = items[item];
        const unsigned char *img = it.arena == ARENA_T ? img0 : img1;
        const int nhalf = it.nq > 128u ? 2 : 1;
        const uint32_t row0 = it.row_begin & ~7u;                     // stages start on an 8-row group
        const uint32_t ntiles = (it.row_end - row0 + TN - 1) / TN;

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

        if (warp == 0) {
            // ===== TMA producer (whole warp runs the loop, one elected lane talks to the TMA engine) =====
            const unsigned char *src = img + (size_t)(row0 >> 3) * GROUP_B;
            for (uint32_t t = 0; t < ntiles; ++t) {
                const uint32_t g = gt + t;
                const int st = g % NST;
                mbar_wait(&S.empty[st], ((g / NST) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&S.full[st], STAGE_B);
                    bulk_g2s(S.b[st], src + (size_t)t * STAGE_B, STAGE_B, &S.full[st]);
                }
                __syncwarp();
            }
        } else if (warp == 1) {
            // ===== MMA issuer: warp-uniform control flow, one elected lane issues =====
            // Accumulator (half h, buffer b) = TMEM columns [128 (2h+b), +128): the epilogue drains buffer b of
            // a half while the tensor core fills buffer b^1 with the next stage.
            const uint64_t adesc0 = smem_desc(smem_u32(S.a[0])), adesc1 = smem_desc(smem_u32(S.a[1]));
            for (uint32_t t = 0; t < ntiles; ++t) {
                const uint32_t g = gt + t;
                const int st = g % NST;
                mbar_wait(&S.full[st], (g / NST) & 1);
                const uint64_t bdesc = smem_desc(smem_u32(S.b[st]));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h < nhalf) {
                        const uint32_t u = ga[h] + t;
                        const int b = u & 1;
                        mbar_wait(&S.tempty[h][b], ((u >> 1) & 1) ^ 1);  // epilogue drained this accumulator
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d = tmem + (uint32_t)(h * 2 + b) * TN;
                            const uint64_t ad = h ? adesc1 : adesc0;
#pragma unroll
                            for (int j = 0; j < KP / 16; ++j)            // one k-step = two 16-byte units = 256 bytes
                                tc_mma(d, ad + (uint64_t)(j * 16), bdesc + (uint64_t)(j * 16), IDESC, j > 0);
                            tc_commit(&S.tfull[h][b]);
                        }
                        __syncwarp();
                    }
                }
                if (elect_one()) tc_commit(&S.empty[st]);               // stage may be refilled once these MMAs retire
                __syncwarp();
            }
        } else {
            // ===== epilogue: one thread = one query =====
            const int ew = warp - 2, h = ew >> 2, quad = warp & 3;
            if (h < nhalf) {
                const uint32_t qslot0 = (uint32_t)(h * 128 + quad * 32);
                const uint32_t qslot = qslot0 + lane;
                uint64_t *pool_warp = pool + ((size_t)blockIdx.x * QT_TENSOR + qslot0) * POOL;
                uint64_t *mypool = pool_warp + (size_t)lane * POOL;
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
                const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
                // 32 columns (data rows) at a time: four 8-wide minima, one compare; only a group whose minimum
                // beats the threshold is looked at element by element
                auto scan = [&](const uint32_t (&r)[32], uint32_t rbase) {
                    float g[4];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const float a0 = fminf(__uint_as_float(r[8 * q4 + 0]), __uint_as_float(r[8 * q4 + 1]));
                        const float a1 = fminf(__uint_as_float(r[8 * q4 + 2]), __uint_as_float(r[8 * q4 + 3]));
                        const float a2 = fminf(__uint_as_float(r[8 * q4 + 4]), __uint_as_float(r[8 * q4 + 5]));
                        const float a3 = fminf(__uint_as_float(r[8 * q4 + 6]), __uint_as_float(r[8 * q4 + 7]));
                        g[q4] = fminf(fminf(a0, a1), fminf(a2, a3));
                    }
                    if (fminf(fminf(g[0], g[1]), fminf(g[2], g[3])) < st.thr) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            if (g[q4] < st.thr) {
#pragma unroll
                                for (int c = 8 * q4; c < 8 * q4 + 8; ++c) {
                                    const float s = __uint_as_float(r[c]);
                                    const uint32_t row = rbase + c;
                                    if (s < st.thr && row >= st.qlo && row < st.qhi)
                                        mypool[st.cnt++] = ((uint64_t)okey(s) << 32) | row;
                                }
                            }
                    }
                };
                // room for 32 more survivors in every pool of this warp (else: sort, keep, tighten)
                auto make_room = [&]() {
                    uint32_t need = __ballot_sync(FULL, st.cnt > (uint32_t)(POOL - 32) ||
                                                            (st.cnt >= 128u && st.thr == __int_as_float(0x7f800000)));
                    while (need) {
                        const int l = __ffs(need) - 1;
                        need &= need - 1;
                        const CompactOut o = compact_lane(l, st.cnt, st.thr, st.margin, st.qid, pool_warp, gthr, gbest, glock, flags, lane);
                        st.cnt = o.cnt; st.thr = o.thr;
                    }
                };
                for (uint32_t t = 0; t < ntiles; ++t) {
                    const uint32_t u = ga[h] + t;
                    // a look at what other CTAs found out about this query
                    if ((t & 7) == 7 && qslot < it.nq) st.thr = fminf(st.thr, okey_inv(ld_relaxed_u32(&gthr[st.qid])));
                    const int b = u & 1;
                    mbar_wait(&S.tfull[h][b], (u >> 1) & 1);
                    tc_fence_after();
                    const uint32_t tcol = tlane + (uint32_t)(h * 2 + b) * TN;
                    const uint32_t trow0 = row0 + t * TN;
                    if (PIPE) {
                        // two register sets: the TMEM load of the next 32 columns is in flight while these are scanned
                        uint32_t ra[32], rb[32];
                        tmem_ld32(tcol, ra);
#pragma unroll 1
                        for (int c4 = 0; c4 < TN / 32; c4 += 2) {
                            make_room();
                            tmem_wait_ld();
                            tmem_ld32(tcol + (c4 + 1) * 32, rb);
                            scan(ra, trow0 + c4 * 32);
                            make_room();
                            tmem_wait_ld();
                            if (c4 + 2 < TN / 32) tmem_ld32(tcol + (c4 + 2) * 32, ra);
                            scan(rb, trow0 + (c4 + 1) * 32);
                        }
                    } else if (dbg == 0) {
#pragma unroll 1
                        for (int c4 = 0; c4 < TN / 32; ++c4) {
                            uint32_t r[32];
                            tmem_ld32(tcol + c4 * 32, r);
                            make_room();
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
                }
                // hand the pool to K5: lists need not be sorted, only short enough
                uint32_t need = __ballot_sync(FULL, st.cnt > (uint32_t)KOUT);
                while (need) {
                    const int l = __ffs(need) - 1;
                    need &= need - 1;
                    const CompactOut o = compact_lane(l, st.cnt, st.thr, st.margin, st.qid, pool_warp, gthr, gbest, glock, flags, lane);
                    st.cnt = o.cnt; st.thr = o.thr;
                }
                // pools that hold scores the global list has not seen pass them on, so that the next items of
                // these queries (other row chunks) start with tight thresholds
                bool news = false;
                if (st.cnt) {
                    const uint32_t gk = ld_relaxed_u32(gbest + (size_t)st.qid * GB + GB - 1);
                    for (uint32_t e = 0; e < st.cnt && !news; ++e) {
                        const uint64_t kk = mypool[e];
                        news = !((uint32_t)kk & CONTRIB) && (uint32_t)(kk >> 32) < gk;
                    }
                }
                need = __ballot_sync(FULL, news);
                while (need) {
                    const int l = __ffs(need) - 1;
                    need &= need - 1;
                    st.thr = contribute_pool(l, st.cnt, st.thr, st.margin, st.qid, pool_warp, gthr, gbest, glock, lane);
                }
                for (int l = 0; l < 32; ++l) {
                    const uint32_t c = __shfl_sync(FULL, st.cnt, l);
                    if (qslot0 + l >= it.nq) break;
                    uint64_t *L = cand + (size_t)(it.out_off + qslot0 + l) * KOUT;
                    const uint64_t *P = pool_warp + (size_t)l * POOL;
                    for (uint32_t e = lane; e < c; e += 32) L[e] = P[e];
                }
                if (qslot < it.nq) cand_cnt[it.out_off + qslot] = st.cnt;
            }
        }
        gt += ntiles;
        ga[0] += ntiles;
        if (nhalf == 2) ga[1] += ntiles;
        tc_fence_before();
        __syncthreads();                                              // item boundary: A may be rebuilt
        tc_fence_after();
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

cudaError_t launch_tile_tensor(hvs_engine *e, const float *queries_dev, const QSlice *slices_dev, const TileItem *items_dev,
                               uint32_t item_begin, uint32_t n_items, const uint32_t *item_q_dev, uint64_t *cand_dev,
                               uint32_t *cand_cnt_dev, uint32_t *gthr_dev, uint32_t *flags_dev)
{
    if (!n_items) return cudaSuccess;
    static bool attr_done = false;
    const int smem = (int)sizeof(TensorSmem);
    if (!attr_done) {
        cudaError_t c = cudaFuncSetAttribute(k_tile_tensor<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (c == cudaSuccess) c = cudaFuncSetAttribute(k_tile_tensor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (c != cudaSuccess) return c;
        attr_done = true;
    }
    const Index &ix = e->index;
    const uint32_t grid = n_items < (uint32_t)e->sm_count ? n_items : (uint32_t)e->sm_count;
    cudaError_t c = e->d_pool.ensure((size_t)e->sm_count * QT_TENSOR * POOL * 8);
    if (c != cudaSuccess) return c;
    c = e->d_gbest.ensure((size_t)e->stats.m * GB * 4);
    if (c != cudaSuccess) return c;
    c = e->d_glock.ensure((size_t)e->stats.m * 4);
    if (c != cudaSuccess) return c;
    c = launch_fill_u32(e, e->d_gbest.as<uint32_t>(), 0xff800000u /* okey(+inf) */, (size_t)e->stats.m * GB);
    if (c != cudaSuccess) return c;
    c = cudaMemsetAsync(e->d_glock.p, 0, (size_t)e->stats.m * 4, e->stream);
    if (c != cudaSuccess) return c;
    static const bool pipe = [] { const char *v = getenv("HVS_K3_PIPE"); return v && v[0] == '1'; }();
    static const int dbg = [] { const char *v = getenv("HVS_K3_DBG"); return v ? atoi(v) : 0; }();
    auto kern = pipe ? k_tile_tensor<true> : k_tile_tensor<false>;
    kern<<<grid, NTHR, smem, e->stream>>>(queries_dev, slices_dev, items_dev + item_begin, n_items, item_q_dev,
                                          ix.xb[0].as<unsigned char>(), ix.xb[1].as<unsigned char>(), ix.xnorm_max, ix.img_scale,
                                          e->d_pool.as<uint64_t>(), cand_dev, cand_cnt_dev, gthr_dev, e->d_gbest.as<uint32_t>(),
                                          e->d_glock.as<uint32_t>(), flags_dev, dbg);
    return cudaGetLastError();
}

}  // namespace hvs
