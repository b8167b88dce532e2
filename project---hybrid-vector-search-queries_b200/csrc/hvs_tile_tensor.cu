// hvs_tile_tensor.cu -- K3: tcgen05 candidate pass for slices that many queries share.
//
// Same job as K2 (hvs_tile_ffma.cu) -- replace the reference's candidate loop + dist_to_query +
// Knn::check_add (include/optimized.hpp:84-117, include/optimized_impl.h:54-170, :284-335) for a
// whole tile of queries at once -- but the 100-deep contraction runs on the 5th-generation tensor
// cores:   D[q][row] = sum_k A[q][k] * B[row][k]   with BF16 operands and FP32 accumulation in TMEM,
//   A[q]   = [ bf16(-2 q_0..99), 1, 1, 1, 0.. ]           (built in shared memory per work item)
//   B[row] = [ bf16(x_0..99), nh, nm, nl, 0.. ]            (nh+nm+nl = ||x||^2 split into 3 bf16)
// so the accumulator IS the score s~ = ||x||^2 - 2 q.x and the epilogue is a pure min/compare.
// B comes from a BF16 image of each arena written at index-build time in the tensor core's
// canonical K-major no-swizzle layout (8-row x 16-byte core matrices), so a stage of 128 rows is
// one contiguous 28,672-byte block moved by ONE 1-D TMA bulk copy -- no tensor map, no swizzle.
//
// One CTA = one work item = up to 256 queries (two M=128 halves sharing every B stage) x a run
// of arena rows.  Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 2..9
// epilogue (warp w reads TMEM lanes 32*(w%4)..; one thread = one query).  Four accumulators of
// 128 columns (2 halves x 2 buffers) use all 512 TMEM columns, so the epilogue of tile t overlaps
// the MMAs of tile t+1.
//
// The result is NOT approximate: a row survives when its bf16 score is within `margin_tensor`
// (hvs_margin.cuh, a rigorous bound on the bf16 rounding of both operands) of the running 100-th
// best, and K5 re-ranks all survivors with the reference's fp32 arithmetic.  Thresholds are shared
// between all CTAs working on the same query through `gthr` (atomicMin), so a CTA that starts late
// starts with a tight threshold.
//
// Roofline: tensor pipe.  224 flop x 256 queries per row-byte... the kernel is bounded by the MMA
// issue rate (7 x M128 N128 K16 per half per stage = 896 cycles per stage) as long as the B stream
// (28,672 B per stage per SM) stays inside L2 bandwidth; see DESIGN.md.
#include <cuda_bf16.h>

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
constexpr int NST = 4;                // B stages in flight
constexpr int STAGE_B = TN * ROW_B;   // 28,672
constexpr int A_B = 128 * ROW_B;      // one query half
constexpr int CBT = 16;               // survivor buffer per query
constexpr int NTHR = 320;
constexpr uint32_t FULL = 0xffffffffu;

static_assert(QT_TENSOR == 256, "two M=128 halves");

struct TensorSmem {
    alignas(128) unsigned char b[NST][STAGE_B];
    alignas(128) unsigned char a[2][A_B];
    uint64_t buf[QT_TENSOR][CBT];
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
// D[tmem] (+)= A[smem] * B[smem]^T, M=128, N=128, K=16, bf16 x bf16 -> f32
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
// c=F32 (bit 4), a=b=BF16 (bits 7,10), K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((128u >> 4) << 24);

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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// per-query epilogue state, owned by one thread for the whole item
struct QState {
    float thr, margin;
    uint32_t cnt, lcnt, qlo, qhi, qid;
};

// Warp-cooperative: merge lane `l`'s survivor buffer into its candidate list.
__device__ __forceinline__ void merge_lane(TensorSmem &S, int l, uint32_t qslot0, QState &st, uint64_t *__restrict__ cand_item,
                                           uint32_t *__restrict__ gthr, uint32_t *__restrict__ flags, int lane)
{
    const uint32_t nb = __shfl_sync(FULL, st.cnt, l);
    const uint32_t nl = __shfl_sync(FULL, st.lcnt, l);
    const float margin = __shfl_sync(FULL, st.margin, l);
    const uint32_t qslot = qslot0 + l;
    uint64_t *L = cand_item + (size_t)qslot * KOUT;
    const uint64_t nk = (uint32_t)lane < nb ? S.buf[qslot][lane] : KEY_INF;
    float lim;
    bool ovf;
    const uint32_t keep = warp_merge_list<KOUT>(L, nl, nk, nb, margin, lim, ovf, lane);
    if (lane == l) {
        st.cnt = 0;
        st.lcnt = keep;
        if (ovf) flags[st.qid] = 1u;                         // K4 re-solves this query exactly
        const float mine = nextafterf(lim, __int_as_float(0x7f800000));
        const float theirs = okey_inv(ld_relaxed_u32(&gthr[st.qid]));
        if (mine < theirs) atomicMin(&gthr[st.qid], okey(mine));
        st.thr = fminf(st.thr, fminf(mine, theirs));
    }
    __syncwarp();
}

}  // namespace

// ---- BF16 image of an arena -----------------------------------------------------------------------
// element (row, k) lives at  (row>>3)*GROUP_B + (k>>3)*128 + (row&7)*16 + (k&7)*2
__global__ void k_build_image(const float *__restrict__ x, const float *__restrict__ xnorm, uint32_t n, uint32_t n_img,
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
            const float xn = xnorm[row];
            const float nh = __bfloat162float(__float2bfloat16_rn(xn));
            const float r1 = xn - nh;                                   // exact
            const float nm = __bfloat162float(__float2bfloat16_rn(r1));
            const float nl = r1 - nm;                                   // exact; its bf16 rounding is below fp32 ulp of xn
            v[4] = nh; v[5] = nm; v[6] = nl;
        }
    }
    __nv_bfloat162 p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4 *>(img + (size_t)(row >> 3) * GROUP_B + u * 128 + (row & 7) * 16) = *reinterpret_cast<uint4 *>(p);
}

void build_bf16_image(hvs_engine *e, int a)
{
    Index &ix = e->index;
    const uint32_t n_img = ((ix.n + 7u) & ~7u) + 2 * TN;        // every stage copy stays inside the image
    if (ix.xb[a].ensure((size_t)n_img * ROW_B) != cudaSuccess) { cudaGetLastError(); ix.xb[a].release(); return; }
    const size_t total = (size_t)n_img * KU;
    k_build_image<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(ix.x[a].as<float>(), ix.xnorm[a].as<float>(), ix.n, n_img,
                                                                          ix.xb[a].as<unsigned char>());
}

bool tensor_path_available() { return true; }

// ---- the sweep --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHR, 1)
k_tile_tensor(const float *__restrict__ queries, const QSlice *__restrict__ slices, const TileItem *__restrict__ items,
              const uint32_t *__restrict__ item_q, const unsigned char *__restrict__ img0, const unsigned char *__restrict__ img1,
              float xnorm_max, uint64_t *__restrict__ cand, uint32_t *__restrict__ cand_cnt, uint32_t *__restrict__ gthr,
              uint32_t *__restrict__ flags)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TensorSmem &S = *reinterpret_cast<TensorSmem *>(smem_raw);
    const TileItem it = items[blockIdx.x];
    const unsigned char *img = it.arena == ARENA_T ? img0 : img1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nhalf = it.nq > 128u ? 2 : 1;
    const uint32_t row0 = it.row_begin & ~7u;                         // stages start on an 8-row group
    const uint32_t ntiles = (it.row_end - row0 + TN - 1) / TN;

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
    // A operand: bf16(-2 q) | 1 1 1 | 0, written straight into the canonical layout
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
            for (int j = 0; j < 8; ++j) v[j] *= -2.f;
            if (u == 12) { v[4] = 1.f; v[5] = 1.f; v[6] = 1.f; }
        }
        __nv_bfloat162 p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        const int r = qs & 127;
        *reinterpret_cast<uint4 *>(S.a[qs >> 7] + (r >> 3) * GROUP_B + u * 128 + (r & 7) * 16) = *reinterpret_cast<uint4 *>(p);
    }
    fence_proxy_async();                                              // generic-proxy writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const unsigned char *src = img + (size_t)(row0 >> 3) * GROUP_B;
            for (uint32_t t = 0; t < ntiles; ++t) {
                const int st = t % NST;
                mbar_wait(&S.empty[st], ((t / NST) & 1) ^ 1);
                mbar_expect_tx(&S.full[st], STAGE_B);
                bulk_g2s(S.b[st], src + (size_t)t * STAGE_B, STAGE_B, &S.full[st]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint64_t adesc[2] = {smem_desc(smem_u32(S.a[0])), smem_desc(smem_u32(S.a[1]))};
            for (uint32_t t = 0; t < ntiles; ++t) {
                const int st = t % NST, b = t & 1;
                mbar_wait(&S.full[st], (t / NST) & 1);
                tc_fence_after();
                const uint64_t bdesc = smem_desc(smem_u32(S.b[st]));
                for (int h = 0; h < nhalf; ++h) {
                    mbar_wait(&S.tempty[h][b], ((t >> 1) & 1) ^ 1);      // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t d = tmem + (uint32_t)(h * 2 + b) * TN;
#pragma unroll
                    for (int j = 0; j < KP / 16; ++j)                    // one k-step = two 16-byte units = 256 bytes
                        tc_mma(d, adesc[h] + (uint64_t)(j * 16), bdesc + (uint64_t)(j * 16), IDESC, j > 0);
                    tc_commit(&S.tfull[h][b]);
                }
                tc_commit(&S.empty[st]);                                // stage may be refilled once these MMAs retire
            }
        }
    } else {
        // ===== epilogue: one thread = one query =====
        const int ew = warp - 2, h = ew >> 2, quad = warp & 3;
        if (h < nhalf) {
            const uint32_t qslot0 = (uint32_t)(h * 128 + quad * 32);
            const uint32_t qslot = qslot0 + lane;
            uint64_t *cand_item = cand + (size_t)it.out_off * KOUT;
            QState st;
            st.cnt = 0; st.lcnt = 0; st.qid = 0; st.qlo = 1; st.qhi = 0; st.margin = 0.f;
            st.thr = __int_as_float(0xff800000);                         // unused slot: nothing passes
            if (qslot < it.nq) {
                st.qid = item_q[it.q_off + qslot];
                const QSlice sl = slices[st.qid];
                st.qlo = max(sl.begin, it.row_begin);
                st.qhi = min(sl.end, it.row_end);
                st.margin = margin_tensor(sl.qnorm, xnorm_max);
                st.thr = okey_inv(ld_relaxed_u32(&gthr[st.qid]));
            }
            const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
            for (uint32_t t = 0; t < ntiles; ++t) {
                const int b = t & 1;
                if ((t & 7) == 7 && qslot < it.nq) st.thr = fminf(st.thr, okey_inv(ld_relaxed_u32(&gthr[st.qid])));
                mbar_wait(&S.tfull[h][b], (t >> 1) & 1);
                tc_fence_after();
                const uint32_t tcol = tlane + (uint32_t)(h * 2 + b) * TN;
                const uint32_t trow0 = row0 + t * TN;
#pragma unroll 1
                for (int c4 = 0; c4 < TN / 32; ++c4) {
                    uint32_t r[32];
                    tmem_ld32(tcol + c4 * 32, r);
                    tmem_wait_ld();
                    float m0 = fminf(__uint_as_float(r[0]), __uint_as_float(r[1]));
                    float m1 = fminf(__uint_as_float(r[2]), __uint_as_float(r[3]));
#pragma unroll
                    for (int c = 4; c < 32; c += 4) {
                        m0 = fminf(m0, fminf(__uint_as_float(r[c]), __uint_as_float(r[c + 1])));
                        m1 = fminf(m1, fminf(__uint_as_float(r[c + 2]), __uint_as_float(r[c + 3])));
                    }
                    const bool hit = fminf(m0, m1) < st.thr;
                    if (__any_sync(FULL, hit)) {
                        // slow path: some lane has a survivor among these 32 rows
                        uint32_t pend = 0;
                        const uint32_t rbase = trow0 + c4 * 32;
                        if (hit) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) {
                                const float s = __uint_as_float(r[c]);
                                const uint32_t row = rbase + c;
                                if (s < st.thr && row >= st.qlo && row < st.qhi) {
                                    if (st.cnt < (uint32_t)CBT) S.buf[qslot][st.cnt++] = ((uint64_t)okey(s) << 32) | row;
                                    else pend |= 1u << c;
                                }
                            }
                        }
                        for (;;) {
                            uint32_t need = __ballot_sync(FULL, pend != 0 || st.cnt >= (uint32_t)(CBT - 2));
                            if (!need) break;
                            while (need) {
                                const int l = __ffs(need) - 1;
                                need &= need - 1;
                                merge_lane(S, l, qslot0, st, cand_item, gthr, flags, lane);
                            }
                            if (pend) {
#pragma unroll
                                for (int c = 0; c < 32; ++c)
                                    if (pend & (1u << c)) {
                                        const float s = __uint_as_float(r[c]);
                                        if (s < st.thr) {
                                            if (st.cnt < (uint32_t)CBT) {
                                                S.buf[qslot][st.cnt++] = ((uint64_t)okey(s) << 32) | (rbase + c);
                                                pend &= ~(1u << c);
                                            }
                                        } else {
                                            pend &= ~(1u << c);
                                        }
                                    }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.tempty[h][b]);
            }
            // flush what is still buffered, publish the list lengths
            uint32_t need = __ballot_sync(FULL, st.cnt > 0);
            while (need) {
                const int l = __ffs(need) - 1;
                need &= need - 1;
                merge_lane(S, l, qslot0, st, cand_item, gthr, flags, lane);
            }
            if (qslot < it.nq) cand_cnt[it.out_off + qslot] = st.lcnt;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
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
        cudaError_t c = cudaFuncSetAttribute(k_tile_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (c != cudaSuccess) return c;
        attr_done = true;
    }
    const Index &ix = e->index;
    k_tile_tensor<<<n_items, NTHR, smem, e->stream>>>(queries_dev, slices_dev, items_dev + item_begin, item_q_dev,
                                                       ix.xb[0].as<unsigned char>(), ix.xb[1].as<unsigned char>(), ix.xnorm_max,
                                                       cand_dev, cand_cnt_dev, gthr_dev, flags_dev);
    return cudaGetLastError();
}

}  // namespace hvs
