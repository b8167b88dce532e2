// hvs_sort.cu -- stable LSD radix sort of (key, 32-bit payload) pairs, 8 bits per pass: the sort behind K0 (two orders of
// D: by T, by (C,T); the per-query O(N) predicate scans of include/baseline.hpp:107-136 become binary searches over the
// result) and behind the device planner / shard assignment of K1 (queries by class, arena, begin, end).
// Hand-written; round 1 used cub::DeviceRadixSort here.
//
// One pass = three launches:
//   k_rs_hist     every CTA counts the digits of its tile (4096 keys)                    -> hist[digit][cta]
//   k_rs_scan     one CTA: exclusive scan over hist in (digit, cta) order                -> global base of every (digit, cta)
//   k_rs_scatter  every CTA re-reads its tile; each WARP owns 512 consecutive keys of it (16 rounds of 32): per-warp digit
//                 counts -> bases per (warp, digit) in warp order -> each round ranks its keys inside the warp with
//                 __match_any_sync (equal digits keep their lane order) and writes them out.  Tile order = CTA order, warp
//                 order, round order, lane order = the input order: the sort is stable, which LSD needs.
// HBM-bound (K0: 8 passes x ~32 B per key; 10^7 keys: a few ms, once per index build); launch-bound for the planner's
// 4x10^4 keys (7 passes x 3 launches).
#include "hvs_engine.h"

namespace hvs {

namespace {
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;                              // keys per lane
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;            // 4096 keys per CTA
constexpr int RS_SCAN_T = 1024;

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

// exclusive scan of `total` counters in place (one CTA; each thread owns a contiguous run)
__global__ void __launch_bounds__(RS_SCAN_T) k_rs_scan(uint32_t *__restrict__ hist, uint32_t total)
{
    __shared__ uint32_t sm[33];
    const uint32_t per = (total + RS_SCAN_T - 1) / RS_SCAN_T;
    const uint32_t i0 = min(total, threadIdx.x * per), i1 = min(total, i0 + per);
    uint32_t s = 0;
    for (uint32_t i = i0; i < i1; ++i) s += hist[i];
    // block exclusive scan of s
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) sm[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = sm[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        sm[lane] = w;
    }
    __syncthreads();
    uint32_t run = (warp ? sm[warp - 1] : 0u) + x - s;
    for (uint32_t i = i0; i < i1; ++i) { const uint32_t c = hist[i]; hist[i] = run; run += c; }
}

template <class KeyT>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const KeyT *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                           uint32_t n, int shift, uint32_t nblocks, const uint32_t *__restrict__ gbase,
                                                           KeyT *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
    __shared__ uint32_t hw[RS_WARPS][256];                  // per-warp digit counts, then per-warp running output positions
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
        uint32_t run = gbase[(size_t)d * nblocks + blockIdx.x];
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
}  // namespace

size_t radix_sort_temp_bytes(uint32_t n)
{
    const uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
    return (size_t)256 * (nblocks ? nblocks : 1) * 4;
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
    KeyT *ki = k0, *ko = k1;
    uint32_t *vi = v0, *vo = v1;
    if ((passes & 1) == 0) {                                  // an even number of passes ends where it began: start from the other side
        cudaError_t c = cudaMemcpyAsync(k1, k0, (size_t)n * sizeof(KeyT), cudaMemcpyDeviceToDevice, st);
        if (c == cudaSuccess) c = cudaMemcpyAsync(v1, v0, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
        if (c != cudaSuccess) return c;
        ki = k1; ko = k0; vi = v1; vo = v0;
    }
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        k_rs_hist<KeyT><<<nblocks, RS_THREADS, 0, st>>>(ki, n, shift, nblocks, hist);
        k_rs_scan<<<1, RS_SCAN_T, 0, st>>>(hist, 256u * nblocks);
        k_rs_scatter<KeyT><<<nblocks, RS_THREADS, 0, st>>>(ki, vi, n, shift, nblocks, hist, ko, vo);
        KeyT *tk = ki; ki = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
    }
    return cudaGetLastError();                                // after an odd number of swaps from the chosen start the result is in (k1, v1)
}

template cudaError_t radix_sort_pairs<uint32_t>(uint32_t *, uint32_t *, uint32_t *, uint32_t *, uint32_t, int, void *, cudaStream_t);
template cudaError_t radix_sort_pairs<uint64_t>(uint64_t *, uint32_t *, uint64_t *, uint32_t *, uint32_t, int, void *, cudaStream_t);

}  // namespace hvs
