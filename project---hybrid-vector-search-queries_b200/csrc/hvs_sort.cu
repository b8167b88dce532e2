// hvs_sort.cu -- stable LSD radix sort of (key, 32-bit payload) pairs, 8 bits per pass: the sort behind K0 (two orders of
// D: by T, by (C,T); the per-query O(N) predicate scans of include/baseline.hpp:107-136 become binary searches over the
// result) and behind the device planner / shard assignment of K1 (queries by class, arena, begin, end).
// Hand-written; round 1 used cub::DeviceRadixSort here.
//
// One pass = three launches:
//   k_rs_hist     every CTA counts the digits of its tile (4096 keys)                    -> hist[digit][cta]
//   k_rs_scan     one CTA per digit: exclusive scan of the digit's row of CTA counts, row total -> tot[digit]
//                 (the scatter kernel scans the 256 totals itself: base of (digit, cta) = digits below + CTAs before)
//   k_rs_scatter  every CTA re-reads its tile; each WARP owns 512 consecutive keys of it (16 rounds of 32): per-warp digit
//                 counts -> bases per (warp, digit) in warp order -> each round ranks its keys inside the warp with
//                 __match_any_sync (equal digits keep their lane order) and writes them out.  Tile order = CTA order, warp
//                 order, round order, lane order = the input order: the sort is stable, which LSD needs.
// HBM-bound (K0: 8 passes x ~32 B per key; 10^7 keys: a few ms, once per index build).  Up to 65536 keys (the planner's
// 4x10^4) all passes run in ONE launch of an 8-CTA cluster (k_rs_cluster below) instead of 7 x 3 launches.
#include "hvs_engine.h"
#include <cooperative_groups.h>
#include <cstdlib>

namespace hvs {

namespace {
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;                              // keys per lane
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;            // 4096 keys per CTA

template <class KeyT>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const KeyT *__restrict__ keys, uint32_t n, int shift, uint32_t nblocks,
                                                        uint32_t *__restrict__ hist)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of x over the CTA's NW warps; *total = the sum (sm: NW + 1 words)
template <int NW>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t x, uint32_t *sm, uint32_t *total)
{
    static_assert(NW <= 32 && (NW & (NW - 1)) == 0, "warp count: power of two, at most 32");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    __syncthreads();                                          // sm may still be read from the previous call
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < NW ? sm[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < NW; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
        if (lane < NW) sm[lane] = wi - w;
        if (lane == NW - 1) sm[NW] = wi;
    }
    __syncthreads();
    *total = sm[NW];
    return sm[warp] + inc - x;
}

// one CTA per digit: exclusive scan of the digit's row of per-CTA counts in place (coalesced, 256 at a time); the row's
// total goes to tot[digit] -- the scatter kernel turns the 256 totals into digit bases itself
__global__ void __launch_bounds__(RS_THREADS) k_rs_scan(uint32_t *__restrict__ hist, uint32_t nblocks, uint32_t *__restrict__ tot)
{
    __shared__ uint32_t sm[RS_WARPS + 1];
    uint32_t *row = hist + (size_t)blockIdx.x * nblocks;
    uint32_t carry = 0;
    for (uint32_t i0 = 0; i0 < nblocks; i0 += RS_THREADS) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t c = i < nblocks ? row[i] : 0u;
        uint32_t sum;
        const uint32_t ex = block_excl_scan<RS_WARPS>(c, sm, &sum);
        if (i < nblocks) row[i] = carry + ex;
        carry += sum;
    }
    if (threadIdx.x == 0) tot[blockIdx.x] = carry;
}

template <class KeyT>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const KeyT *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                           uint32_t n, int shift, uint32_t nblocks, const uint32_t *__restrict__ gbase,
                                                           const uint32_t *__restrict__ tot,
                                                           KeyT *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
    __shared__ uint32_t hw[RS_WARPS][256];                  // per-warp digit counts, then per-warp running output positions
    __shared__ uint32_t sm[RS_WARPS + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&hw[0][0])[i] = 0;
    __syncthreads();
    const uint32_t wbase = blockIdx.x * RS_TILE + warp * (32 * RS_ROUNDS);     // this warp's 512 consecutive keys
    KeyT k[RS_ROUNDS];
    uint32_t v[RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        k[r] = i < n ? keys_in[i] : (KeyT)0;
        v[r] = i < n ? vals_in[i] : 0u;
    }
    // (1) this warp's digit counts
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const bool valid = wbase + r * 32 + lane < n;
        const uint32_t d = valid ? ((uint32_t)(k[r] >> shift) & 0xffu) : 256u + (uint32_t)lane;     // invalid lanes match nobody
        const uint32_t mask = __match_any_sync(0xffffffffu, d);
        if (valid && lane == __ffs((int)mask) - 1) hw[warp][d] += __popc(mask);                  // one leader per digit group; rows are per warp
        __syncwarp();
    }
    __syncthreads();
    // (2) output position of every (warp, digit): the CTA's global base for the digit, then the warps in order
    {
        const uint32_t d = threadIdx.x;                       // RS_THREADS == 256 digits
        uint32_t all;
        const uint32_t dbase = block_excl_scan<RS_WARPS>(tot[d], sm, &all);            // keys with a smaller digit, over the whole input
        uint32_t run = dbase + gbase[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { const uint32_t c = hw[w][d]; hw[w][d] = run; run += c; }
    }
    __syncthreads();
    // (3) rank inside the warp, round by round (equal digits keep their lane order), and write
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const bool valid = wbase + r * 32 + lane < n;
        const uint32_t d = valid ? ((uint32_t)(k[r] >> shift) & 0xffu) : 256u + (uint32_t)lane;
        const uint32_t mask = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs((int)mask) - 1;
        uint32_t base = 0;
        if (valid && lane == leader) { base = hw[warp][d]; hw[warp][d] = base + __popc(mask); }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) {
            const uint32_t pos = base + __popc(mask & ((1u << lane) - 1u));
            keys_out[pos] = k[r];
            vals_out[pos] = v[r];
        }
        __syncwarp();
    }
}

// Small inputs (the planner's 4x10^4 query keys): ALL passes in one launch.  One cluster of 8 CTAs, each owning 8192
// consecutive keys (a warp: 512); per pass the CTAs publish their digit totals in shared memory, read each other's over
// DSMEM between two cluster barriers, and scatter through global memory (L2: loads bypass L1, the barrier's
// release/acquire orders the stores).  Same order rule as the three-launch pass, so equally stable.
constexpr int RC_THREADS = 512;
constexpr int RC_WARPS = RC_THREADS / 32;
constexpr int RC_ROUNDS = 16;
constexpr int RC_CTAS = 8;
constexpr uint32_t RC_TILE = RC_THREADS * RC_ROUNDS;
constexpr uint32_t RC_MAX = RC_TILE * RC_CTAS;             // 65536 keys

template <class KeyT>
__global__ void __cluster_dims__(RC_CTAS, 1, 1) __launch_bounds__(RC_THREADS)
k_rs_cluster(KeyT *ka, uint32_t *va, KeyT *kb, uint32_t *vb, uint32_t n, int passes)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ uint32_t hw[RC_WARPS][256];
    __shared__ uint32_t ctot[256];                           // this CTA's digit totals, read by the whole cluster
    __shared__ uint32_t sm[RC_WARPS + 1];
    const uint32_t me = cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t wbase = me * RC_TILE + warp * (32 * RC_ROUNDS);
    for (int p = 0; p < passes; ++p) {
        const KeyT *ki = (p & 1) ? kb : ka;
        const uint32_t *vi = (p & 1) ? vb : va;
        KeyT *ko = (p & 1) ? ka : kb;
        uint32_t *vo = (p & 1) ? va : vb;
        const int shift = 8 * p;
        for (int i = tid; i < RC_WARPS * 256; i += RC_THREADS) (&hw[0][0])[i] = 0;
        __syncthreads();
        KeyT k[RC_ROUNDS];
        uint32_t v[RC_ROUNDS];
#pragma unroll
        for (int r = 0; r < RC_ROUNDS; ++r) {
            const uint32_t i = wbase + r * 32 + lane;
            k[r] = i < n ? __ldcg(ki + i) : (KeyT)0;
            v[r] = i < n ? __ldcg(vi + i) : 0u;
        }
#pragma unroll
        for (int r = 0; r < RC_ROUNDS; ++r) {
            const bool valid = wbase + r * 32 + lane < n;
            const uint32_t d = valid ? ((uint32_t)(k[r] >> shift) & 0xffu) : 256u + (uint32_t)lane;
            const uint32_t mask = __match_any_sync(0xffffffffu, d);
            if (valid && lane == __ffs((int)mask) - 1) hw[warp][d] += __popc(mask);
            __syncwarp();
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t c = 0;
#pragma unroll
            for (int w = 0; w < RC_WARPS; ++w) c += hw[w][tid];
            ctot[tid] = c;
        }
        cluster.sync();
        uint32_t before = 0, total = 0;
        if (tid < 256) {
#pragma unroll
            for (uint32_t r = 0; r < (uint32_t)RC_CTAS; ++r) {
                const uint32_t x = cluster.map_shared_rank(ctot, r)[tid];
                total += x;
                if (r < me) before += x;
            }
        }
        uint32_t all;
        const uint32_t dbase = block_excl_scan<RC_WARPS>(total, sm, &all);
        if (tid < 256) {
            uint32_t run = dbase + before;
#pragma unroll
            for (int w = 0; w < RC_WARPS; ++w) { const uint32_t c = hw[w][tid]; hw[w][tid] = run; run += c; }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RC_ROUNDS; ++r) {
            const bool valid = wbase + r * 32 + lane < n;
            const uint32_t d = valid ? ((uint32_t)(k[r] >> shift) & 0xffu) : 256u + (uint32_t)lane;
            const uint32_t mask = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs((int)mask) - 1;
            uint32_t base = 0;
            if (valid && lane == leader) { base = hw[warp][d]; hw[warp][d] = base + __popc(mask); }
            base = __shfl_sync(0xffffffffu, base, leader);
            if (valid) {
                const uint32_t pos = base + __popc(mask & ((1u << lane) - 1u));
                ko[pos] = k[r];
                vo[pos] = v[r];
            }
            __syncwarp();
        }
        cluster.sync();                                       // every CTA has read ctot; this pass's stores are visible to the next
    }
}
}  // namespace

// kernel launches radix_sort_pairs makes for n keys of end_bit bits (the engine's launch count is a claim the bench reports)
uint32_t radix_sort_launches(uint32_t n, int end_bit)
{
    if (!n) return 0;
    const int passes = end_bit > 8 ? (end_bit + 7) / 8 : 1;
    static const bool one_launch = [] { const char *e = std::getenv("HVS_SORT_CLUSTER"); return !(e && e[0] == '0'); }();
    return (one_launch && n <= RC_MAX) ? 1u : 3u * (uint32_t)passes;
}

size_t radix_sort_temp_bytes(uint32_t n)
{
    const uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
    return (size_t)256 * (nblocks ? nblocks : 1) * 4 + 256 * 4;
}

// Sorts n (key, value) pairs by bits [0, end_bit) of the key.  (k0, v0) holds the input and is used as scratch; the result
// lands in (k1, v1).  tmp: radix_sort_temp_bytes(n) bytes.
template <class KeyT>
cudaError_t radix_sort_pairs(KeyT *k0, uint32_t *v0, KeyT *k1, uint32_t *v1, uint32_t n, int end_bit, void *tmp, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    const uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
    int passes = (end_bit + 7) / 8;
    if (passes < 1) passes = 1;
    uint32_t *hist = reinterpret_cast<uint32_t *>(tmp);
    uint32_t *tot = hist + (size_t)256 * nblocks;
    KeyT *ki = k0, *ko = k1;
    uint32_t *vi = v0, *vo = v1;
    if ((passes & 1) == 0) {                                  // an even number of passes ends where it began: start from the other side
        cudaError_t c = cudaMemcpyAsync(k1, k0, (size_t)n * sizeof(KeyT), cudaMemcpyDeviceToDevice, st);
        if (c == cudaSuccess) c = cudaMemcpyAsync(v1, v0, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
        if (c != cudaSuccess) return c;
        ki = k1; ko = k0; vi = v1; vo = v0;
    }
    static const bool one_launch = [] { const char *e = std::getenv("HVS_SORT_CLUSTER"); return !(e && e[0] == '0'); }();
    if (one_launch && n <= RC_MAX) {
        k_rs_cluster<KeyT><<<RC_CTAS, RC_THREADS, 0, st>>>(ki, vi, ko, vo, n, passes);
        return cudaGetLastError();
    }
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        k_rs_hist<KeyT><<<nblocks, RS_THREADS, 0, st>>>(ki, n, shift, nblocks, hist);
        k_rs_scan<<<256, RS_THREADS, 0, st>>>(hist, nblocks, tot);
        k_rs_scatter<KeyT><<<nblocks, RS_THREADS, 0, st>>>(ki, vi, n, shift, nblocks, hist, tot, ko, vo);
        KeyT *tk = ki; ki = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
    }
    return cudaGetLastError();                                // after an odd number of swaps from the chosen start the result is in (k1, v1)
}

template cudaError_t radix_sort_pairs<uint32_t>(uint32_t *, uint32_t *, uint32_t *, uint32_t *, uint32_t, int, void *, cudaStream_t);
template cudaError_t radix_sort_pairs<uint64_t>(uint64_t *, uint32_t *, uint64_t *, uint32_t *, uint32_t, int, void *, cudaStream_t);

}  // namespace hvs
