// hvs_topk.cuh -- block-level selection primitives shared by the direct scan (K4), the finalize
// kernel (K5) and the merge kernel.  Candidates are 64-bit keys (distance bits << 32 | payload):
// distances are sums of squares (>= 0) so the unsigned integer order of the key is the order
// "closer first, smaller payload first on ties".
//
// Reference: this replaces class Knn of include/optimized_impl.h:179-438 (unsorted 100-slot array,
// replace-the-worst in check_add :284-335, arg-max rescan in find_worst :201-274, final std::sort
// in get_knn_sorted :392-437).  Instead of rescanning 100 slots per accepted candidate, accepted
// candidates are appended to a buffer that is compacted (bitonic sort, keep the best 100) only when
// it fills; between compactions the acceptance test is one compare against a stale threshold.
#pragma once
#include "hvs_common.cuh"

namespace hvs {

constexpr uint64_t KEY_INF = 0xffffffffffffffffull;

// In-place ascending bitonic sort of a[0..n) in shared memory, n a power of two, all `nthreads` (a multiple of 32)
// threads of the block participate.  Ends with a __syncthreads().
// Every thread takes PAIRS (t -> elements i and i|j), so all lanes work and the exchange is branch-free; 32 consecutive
// pairs with j <= 32 are one aligned run of 64 elements, the same run for every such j, so a step is separated from the
// next by a warp barrier only -- a block barrier is needed just around the steps with j > 32 (512 keys: 9 instead of 45).
__device__ __forceinline__ void block_bitonic_sort(uint64_t *a, int n, int tid, int nthreads)
{
    const int half = n >> 1;
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < half; t += nthreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const uint64_t x = a[i], y = a[ixj];
                const bool up = (i & k) == 0;
                const uint64_t lo = x < y ? x : y, hi = x < y ? y : x;
                a[i] = up ? lo : hi;
                a[ixj] = up ? hi : lo;
            }
            const int next_j = j > 1 ? (j >> 1) : (k < n ? k : 64);    // the step after this one (64: the end, block barrier)
            if (j > 32 || next_j > 32) __syncthreads();
            else __syncwarp();
        }
    }
}

// In-place ascending bitonic sort of a[0..n) in shared memory by ONE warp, n a power of two >= 2.
__device__ __forceinline__ void warp_bitonic_sort(uint64_t *a, int n, int lane)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (n >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const uint64_t x = a[i], y = a[ixj];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { a[i] = y; a[ixj] = x; }
            }
            __syncwarp();
        }
    }
}

// Order-preserving float <-> uint32 for scores that may be negative (s = ||x||^2 - 2 q.x).
__device__ __forceinline__ uint32_t okey(float s)
{
    const uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Next probe of the rank search shared by K3's compact_warp / merge_global and K5's compact_select: a score v between the bracket ends
// (#{<= klo} = clo < K <= chi = #{<= khi}).  The survivors sit in the lower tail of the score distribution, where the
// count grows roughly exponentially with the score, so the probe interpolates log(count) linearly -- in the FLOAT
// domain (order-preserving keys are very non-linear around zero) -- and every third step bisects the key interval,
// which bounds the worst case.  Measured on pools of 192..480 tail scores: ~3 probes instead of 9-18 with linear
// interpolation on the keys.
__device__ __forceinline__ uint32_t select_probe(uint32_t klo, uint32_t khi, uint32_t clo, uint32_t chi, int it)
{
    uint32_t mid;
    if (it % 3 == 2) mid = klo + ((khi - klo) >> 1);
    else {
        const float flo = okey_inv(klo), fhi = okey_inv(khi);
        const float a = __logf(fmaxf((float)clo, 0.5f)), b = __logf((float)chi);
        const float t = (__logf((float)(K + 4)) - a) / fmaxf(b - a, 1e-6f);
        mid = okey(flo + (fhi - flo) * fminf(fmaxf(t, 0.f), 1.f));
    }
    return min(max(mid, klo + 1u), khi - 1u);
}

// Warp-cooperative merge of up to 32 new keys (lane i passes key i, KEY_INF when it has none) into a
// sorted candidate list L[0..nl) in global memory (nl <= KOUT_), in place and without scratch:
// the new keys are sorted across lanes with shuffles, every element's merged position is its own
// index plus the number of elements of the other sequence that precede it, and everything is read
// into registers before anything is written.  Keeps the entries whose score is within `margin` of
// the K-th best (all of them while fewer than K are known).  Returns the new length; lim_out is
// (K-th score + margin) or +inf; overflow is set when more than KOUT_ entries sit inside the margin.
// Keys must be distinct (they carry the row index).
template <int KOUT_>
__device__ __forceinline__ uint32_t warp_merge_list(uint64_t *__restrict__ L, uint32_t nl, uint64_t newkey, uint32_t nb,
                                                    float margin, float &lim_out, bool &overflow, int lane)
{
    constexpr int PER = KOUT_ / 32;
    constexpr uint32_t FULLMASK = 0xffffffffu;
    uint64_t key = newkey;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint64_t other = __shfl_xor_sync(FULLMASK, key, j);
            const bool take_min = ((lane & j) == 0) == ((lane & k) == 0);
            key = take_min ? (key < other ? key : other) : (key < other ? other : key);
        }
    uint64_t lk[PER];
    uint32_t cl[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t i = lane + 32 * j;
        lk[j] = i < nl ? L[i] : KEY_INF;
        cl[j] = 0;
    }
    uint32_t myrank = 0;
    for (uint32_t t = 0; t < nb; ++t) {
        const uint64_t nk = __shfl_sync(FULLMASK, key, t);
        uint32_t c = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const bool lt = lk[j] < nk;
            c += lt;
            cl[j] += !lt;
        }
        c = __reduce_add_sync(FULLMASK, c);
        if ((uint32_t)lane == t) myrank = c;
    }
    const uint32_t total = nl + nb;
    uint32_t keep = total;
    lim_out = __int_as_float(0x7f800000);
    overflow = false;
    if (total >= (uint32_t)K) {
        uint32_t hi = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const uint32_t i = lane + 32 * j;
            if (i < nl && i + cl[j] == (uint32_t)(K - 1)) hi = (uint32_t)(lk[j] >> 32);
        }
        if ((uint32_t)lane < nb && lane + myrank == (uint32_t)(K - 1)) hi = (uint32_t)(key >> 32);
        hi = __reduce_max_sync(FULLMASK, hi);
        const float lim = okey_inv(hi) + margin;
        uint32_t nin = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if ((uint32_t)(lane + 32 * j) < nl && okey_inv((uint32_t)(lk[j] >> 32)) <= lim) ++nin;
        if ((uint32_t)lane < nb && okey_inv((uint32_t)(key >> 32)) <= lim) ++nin;
        nin = __reduce_add_sync(FULLMASK, nin);
        keep = nin;
        if (keep > (uint32_t)KOUT_) { keep = KOUT_; overflow = true; }
        lim_out = lim;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t i = lane + 32 * j, pos = i + cl[j];
        if (i < nl && pos < keep) L[pos] = lk[j];
    }
    if ((uint32_t)lane < nb && lane + myrank < keep) L[lane + myrank] = key;
    __syncwarp();
    return keep;
}

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Streaming top-K buffer in shared memory, one per CTA (one query per CTA).
//   accept test : key_dist < thr           (thr = +inf until the first compaction)
//   compaction  : sort, keep the K best plus every entry within `margin` of the K-th
// With margin == 0 and exact distances this is an exact top-K; with margin = 2*eps and approximate
// distances it keeps a superset that provably contains the exact top-K (DESIGN.md "margins").
//   ORD == false : key = raw bits of a non-negative distance   (pack_key)
//   ORD == true  : key = okey(score), scores of either sign
template <int CAP, bool ORD = false>
struct TopBuf {
    uint64_t cand[CAP];
    uint32_t cnt;
    float thr;
    uint32_t overflow;   // set when the margin set alone exceeds the buffer

    __device__ __forceinline__ void init(int tid)
    {
        if (tid == 0) { cnt = 0; thr = __int_as_float(0x7f800000); overflow = 0; }
    }
    static __device__ __forceinline__ float key_val(uint64_t k)
    {
        return ORD ? okey_inv((uint32_t)(k >> 32)) : __uint_as_float((uint32_t)(k >> 32));
    }
    // Returns the slot the entry took.  Callers decide about a compaction from the slots their own pushes got
    // (cnt > L  <=>  some push of this round got slot >= L), combined with __syncthreads_or: reading `cnt` after
    // the barrier would race with threads that already push the next round.
    __device__ __forceinline__ uint32_t push(float d, uint32_t payload)
    {
        uint32_t slot = atomicAdd(&cnt, 1u);
        if (slot < (uint32_t)CAP) cand[slot] = ORD ? (((uint64_t)okey(d) << 32) | payload) : pack_key(d, payload);
        return slot;
    }
    // Call from ALL threads after a __syncthreads().  Keeps <= keep_max entries.
    __device__ __forceinline__ void compact(int tid, int nthreads, float margin, int keep_max)
    {
        static_assert((CAP & (CAP - 1)) == 0, "the bitonic sort pads to the next power of two, which must fit the buffer");
        int c = min((int)cnt, CAP);
        int n = next_pow2(c < 2 ? 2 : c);
        for (int i = c + tid; i < n; i += nthreads) cand[i] = KEY_INF;
        __syncthreads();
        block_bitonic_sort(cand, n, tid, nthreads);
        if (c > K) {
            float dk = key_val(cand[K - 1]);
            float lim = dk + margin;
            // entries are sorted: count those with dist <= lim (all of the first K qualify)
            int keep = K;
            if (margin > 0.f) {
                // every thread scans a strided share; tiny (CAP <= 1024)
                __shared__ int s_keep;
                if (tid == 0) s_keep = K;
                __syncthreads();
                int local = 0;
                for (int i = K + tid; i < c; i += nthreads)
                    if (key_val(cand[i]) <= lim) local = i + 1;
                if (local) atomicMax(&s_keep, local);
                __syncthreads();
                keep = s_keep;
            }
            if (keep > keep_max) { keep = keep_max; if (tid == 0) overflow = 1; }
            __syncthreads();
            if (tid == 0) { cnt = keep; thr = (margin > 0.f) ? nextafterf(lim, __int_as_float(0x7f800000)) : dk; }
        } else if (tid == 0) {
            cnt = c;
        }
        __syncthreads();
    }
    // Same contract as compact() for margin > 0, ORD keys and a block of NT threads -- without sorting: the entries
    // go to registers, the block brackets a score v with K <= #{<= v} <= K + 8 by counting (select_probe: ~3-4
    // probes, one barrier each), and writes back what is within `margin` of v.  K5 spent most of its time in the
    // 55 barrier-separated steps of a 1024-key bitonic sort here; order was never needed, only the cut.
    template <int NT>
    __device__ __forceinline__ void compact_select(int tid, float margin, int keep_max)
    {
        static_assert(ORD, "scores of either sign");
        constexpr int PER = (CAP + NT - 1) / NT;
        __shared__ uint32_t s_lo, s_hi, s_cnt[3];
        const int c = min((int)cnt, CAP);
        if (c <= K) {                                                 // nothing to cut (cnt is block-uniform here)
            if (tid == 0) cnt = c;
            __syncthreads();
            return;
        }
        uint64_t e[PER];
        uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = tid + j * NT;
            e[j] = i < c ? cand[i] : KEY_INF;
            if (i < c) { lo = min(lo, (uint32_t)(e[j] >> 32)); hi = max(hi, (uint32_t)(e[j] >> 32)); }
        }
        if (tid == 0) { s_lo = 0xffffffffu; s_hi = 0u; s_cnt[0] = s_cnt[1] = s_cnt[2] = 0u; }
        __syncthreads();
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if ((tid & 31) == 0) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
        __syncthreads();
        uint32_t klo = s_lo, khi = s_hi, clo = 0u, chi = (uint32_t)c;   // #{<= klo} = clo < K <= chi = #{<= khi}
        if (klo != khi) {
            klo -= 1u;
            for (int it = 0; it < 48 && khi - klo > 1u; ++it) {
                const uint32_t mid = select_probe(klo, khi, clo, chi, it);
                uint32_t n = 0;
#pragma unroll
                for (int j = 0; j < PER; ++j) n += (tid + j * NT < c && (uint32_t)(e[j] >> 32) <= mid) ? 1u : 0u;
                n = __reduce_add_sync(0xffffffffu, n);
                if ((tid & 31) == 0 && n) atomicAdd(&s_cnt[it % 3], n);
                __syncthreads();
                n = s_cnt[it % 3];
                if (tid == 0) s_cnt[(it + 2) % 3] = 0u;                // used two probes from now: behind the next barrier
                if (n >= (uint32_t)K) { khi = mid; chi = n; if (n <= (uint32_t)K + 8u) break; }
                else { klo = mid; clo = n; }
            }
        }
        const float lim = okey_inv(khi) + margin;
        const uint32_t limk = okey(lim);
        __syncthreads();
        if (tid == 0) cnt = 0u;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if (tid + j * NT < c && (uint32_t)(e[j] >> 32) <= limk) {
                const uint32_t slot = atomicAdd(&cnt, 1u);
                if (slot < (uint32_t)keep_max) cand[slot] = e[j];
            }
        __syncthreads();
        if (tid == 0) {
            if (cnt > (uint32_t)keep_max) { cnt = (uint32_t)keep_max; overflow = 1; }
            thr = nextafterf(lim, __int_as_float(0x7f800000));
        }
        __syncthreads();
    }
};

// Final stage shared with K5: `top.cand[0..cnt)` hold (dist, arena_pos); convert to (dist, id),
// apply the pad rule, sort, write.
template <int CAP, bool ORD>
__device__ __forceinline__ void finish_query(TopBuf<CAP, ORD> &top, const float *q_smem, const Arena &A, uint32_t len,
                                             const float *__restrict__ tail, uint32_t n_total, uint32_t id_offset, uint32_t q, bool partial,
                                             uint32_t *__restrict__ out_ids, float *__restrict__ out_dist,
                                             uint32_t *__restrict__ out_count, int tid, int nthreads)
{
    // keep the best K by (dist, arena position), then re-key by original id: the output is ordered by (dist, id);
    // WHICH rows survive among exactly equal distances at the 100-th place is decided by arena position (and, in
    // the streaming kernels, by when a compaction ran) -- like the reference, whose tie order is unspecified
    // (unstable std::sort, include/baseline.hpp:159-166)
    if ((int)top.cnt > K) top.compact(tid, nthreads, 0.f, K);
    int c = min((int)top.cnt, K);
    for (int i = tid; i < c; i += nthreads) {
        uint64_t k = top.cand[i];
        top.cand[i] = (k & 0xffffffff00000000ull) | A.ids[(uint32_t)k];
    }
    __syncthreads();
    if (!partial && len < (uint32_t)K) {
        // include/baseline.hpp:138-147: append ids n-1, n-2, ... (no predicate, no de-duplication)
        // until there are K candidates; they are ranked together with the real matches.
        int npad = K - (int)len;
        for (int s = tid; s < npad; s += nthreads) {
            float d = ref_dist_row(tail + (size_t)s * DIM, q_smem);
            top.cand[c + s] = pack_key(d, n_total - 1u - (uint32_t)s + id_offset);   // same id space as A.ids (k_gather adds id_offset)
        }
        c += npad;
    }
    for (int i = c + tid; i < 128; i += nthreads) top.cand[i] = KEY_INF;
    __syncthreads();
    block_bitonic_sort(top.cand, 128, tid, nthreads);
    if (!partial) {
        for (int i = tid; i < K; i += nthreads) out_ids[(size_t)q * K + i] = (uint32_t)top.cand[i];
    } else {
        for (int i = tid; i < K; i += nthreads) {
            uint64_t k = top.cand[i];
            bool valid = i < c;
            out_ids[(size_t)q * K + i] = valid ? (uint32_t)k : 0xffffffffu;
            out_dist[(size_t)q * K + i] = valid ? __uint_as_float((uint32_t)(k >> 32)) : __int_as_float(0x7f800000);
        }
        if (tid == 0) out_count[q] = len;
    }
}

}  // namespace hvs
