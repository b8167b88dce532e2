// hvs_index.cu -- K0: the indexing phase.
//
// Replaces the per-query O(N) predicate scans of the reference (include/baseline.hpp:107-136,
// include/optimized.hpp:84-117, include/optimized_parallel.hpp:105-138) with two radix-sorted,
// re-laid-out copies of D so that every predicate is one contiguous row range:
//   arena T  : rows ordered by ord(T)               -> types 0 (all rows) and 2 (l <= T <= r)
//   arena CT : rows ordered by ord(C)<<32 | ord(T)  -> types 1 (C == v) and 3 (C == v, l <= T <= r)
// The sorts are the hand-written LSD radix sort of hvs_sort.cu.
// Each arena holds the 100-d vectors as contiguous 400-byte rows (16-byte aligned, so a tile of
// consecutive rows is ONE contiguous block that the TMA engine moves with a single 1-D bulk copy),
// the original row ids, ||x||^2, and the sorted keys for binary search.  Never sees queries
// (contest rule, README.md:68).  HBM-bound: ~2 sorts + 2 gathers of 400 B/row.
#include <algorithm>
#include <cmath>

#include "hvs_engine.h"

namespace hvs {

__global__ void k_make_keys(const float *__restrict__ rows, uint32_t n, uint32_t *__restrict__ key_t,
                            uint64_t *__restrict__ key_ct, uint32_t *__restrict__ perm)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float2 ct = *reinterpret_cast<const float2 *>(rows + (size_t)j * DROW);   // rows are 408 B: 8-byte aligned
    uint32_t kc = ord_key(ct.x), kt = ord_key(ct.y);
    key_t[j] = kt;
    key_ct[j] = ((uint64_t)kc << 32) | kt;
    perm[j] = j;
}

__global__ void k_iota(uint32_t *__restrict__ p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// One warp per destination row: gather the vector of source row perm[p] into arena row p,
// record its id and squared norm.  Source rows are only 8-byte aligned (408-byte pitch, vector at +8).
__global__ void k_gather(const float *__restrict__ rows, const uint32_t *__restrict__ perm, uint32_t n,
                         uint32_t id_offset, float *__restrict__ x, uint32_t *__restrict__ ids,
                         float *__restrict__ xnorm, uint32_t *__restrict__ inv /* may be null */)
{
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    uint32_t src = perm[warp];
    const float2 *s = reinterpret_cast<const float2 *>(rows + (size_t)src * DROW + 2);
    float2 *d = reinterpret_cast<float2 *>(x + (size_t)warp * DIM);
    float acc = 0.f;
    float2 v = s[lane];
    d[lane] = v;
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
    if (lane < DIM / 2 - 32) {
        v = s[32 + lane];
        d[32 + lane] = v;
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        ids[warp] = src + id_offset;
        xnorm[warp] = acc;
        if (inv) inv[src] = warp;
    }
}

// tail[s-1] = vector of row n_total - s, s = 1..100  (pad rule, include/baseline.hpp:138-147)
__global__ void k_tail(const float *__restrict__ rows, uint32_t n_total, float *__restrict__ tail)
{
    uint32_t s = blockIdx.x + 1;
    const float *src = rows + (size_t)(n_total - s) * DROW + 2;
    for (int i = threadIdx.x; i < DIM; i += blockDim.x) tail[(s - 1) * DIM + i] = src[i];
}

// Histogram of the exponent field of ||x||^2 (bin 255 = inf / NaN): the host picks the norm cutoff above which
// rows are "outliers" (at most OUTLIER_MAX of them).
__global__ void k_norm_hist(const float *__restrict__ v, uint32_t n, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;                                    // blockDim.x == 256
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&h[(__float_as_uint(v[i]) >> 23) & 0xffu], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// Rows whose ||x||^2 is not below `cutoff` (a power of two; NaN and inf included): their arena position goes to
// out_pos (any order; the host sorts the few of them) and their ||x||^2 becomes +inf, which makes the approximate
// sweeps (K2: score = ||x||^2 - 2 q.x = +inf; K3: see k_build_image) skip them.  K5 scores them exactly.
__global__ void k_mark_outliers(float *__restrict__ xnorm, uint32_t n, float cutoff, uint32_t *__restrict__ out_pos,
                                uint32_t *__restrict__ out_cnt, uint32_t cap)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = xnorm[i];
    if (!(a < cutoff)) {
        const uint32_t slot = atomicAdd(out_cnt, 1u);
        if (slot < cap) out_pos[slot] = i;
        xnorm[i] = __int_as_float(0x7f800000);
    }
}

__global__ void k_max_f32(const float *__restrict__ v, uint32_t n, uint32_t *__restrict__ out_bits)
{
    float m = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float a = v[i];
        if (a < __int_as_float(0x7f800000) && a > m) m = a;    // outliers (+inf) and NaN do not count
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order like their bits
}

cudaError_t build_tensor_image(hvs_engine *e, int a);   // hvs_tile_tensor.cu

#define CK(call)                                                                     \
    do {                                                                             \
        cudaError_t _e = (call);                                                     \
        if (_e != cudaSuccess) {                                                     \
            e->err = std::string(#call) + ": " + cudaGetErrorString(_e);             \
            return _e;                                                               \
        }                                                                            \
    } while (0)

cudaError_t index_build_device(hvs_engine *e, const float *rows, uint32_t n_total, float sample_proportion)
{
    Index &ix = e->index;
    ix.built = false;
    const float fn = sample_proportion * (float)n_total;           // include/baseline.hpp:74 (float multiply, then truncation)
    const uint32_t n = fn >= (float)n_total ? n_total : (uint32_t)fn;   // clamped BEFORE the cast: float -> uint32 of >= 2^32 is undefined
    ix.n_total = n_total;
    ix.n = n;
    ix.id_offset = e->id_offset;
    cudaStream_t st = e->stream;
    const size_t n1 = n ? n : 1;

    DevBuf key_t_in, key_ct_in, perm_in, perm_out, tmp, maxbits, hist, ol_tmp;   // released by their destructors on every exit path
    CK(key_t_in.ensure(n1 * 4));
    CK(key_ct_in.ensure(n1 * 8));
    CK(perm_in.ensure(n1 * 4));
    CK(perm_out.ensure(n1 * 4));
    CK(ix.keys_t.ensure(n1 * 4));
    CK(ix.keys_ct.ensure(n1 * 8));
    CK(ix.tail.ensure((size_t)K * DIM * 4));
    CK(ix.inv_t.ensure((size_t)n_total * 4));
    CK(cudaMemsetAsync(ix.inv_t.p, 0xff, (size_t)n_total * 4, st));
    CK(maxbits.ensure(4));
    CK(hist.ensure(256 * 4));
    CK(ol_tmp.ensure((size_t)(OUTLIER_MAX + 1) * 4));
    for (int a = 0; a < 2; ++a) { CK(ix.outl[a].ensure((size_t)OUTLIER_MAX * 4)); ix.n_outl[a] = 0; }
    for (int a = 0; a < 2; ++a) {
        CK(ix.x[a].ensure(n1 * ROW_BYTES + 64 * ROW_BYTES));   // slack rows: tile loads may over-read nothing, but keep TMA boxes in bounds
        CK(ix.ids[a].ensure(n1 * 4));
        CK(ix.xnorm[a].ensure(n1 * 4));
    }
    cudaError_t rc = cudaSuccess;
    bool approx_ok = true;
    auto fail = [&](cudaError_t c, const char *what) { e->err = std::string(what) + ": " + cudaGetErrorString(c); rc = c; };

    if (n) {
        k_make_keys<<<(n + 255) / 256, 256, 0, st>>>(rows, n, key_t_in.as<uint32_t>(), key_ct_in.as<uint64_t>(), perm_in.as<uint32_t>());
        cudaError_t c = tmp.ensure(radix_sort_temp_bytes(n));
        if (c != cudaSuccess) fail(c, "sort temp alloc");
        const unsigned gather_blocks = (unsigned)(((size_t)n * 32 + 255) / 256);
        if (rc == cudaSuccess) {
            // (key_t_in, perm_in) are the sort's scratch afterwards: the row permutation is written afresh for the second sort
            c = radix_sort_pairs<uint32_t>(key_t_in.as<uint32_t>(), perm_in.as<uint32_t>(), ix.keys_t.as<uint32_t>(), perm_out.as<uint32_t>(), n, 32, tmp.p, st);
            if (c != cudaSuccess) fail(c, "radix sort (T)");
            k_gather<<<gather_blocks, 256, 0, st>>>(rows, perm_out.as<uint32_t>(), n, ix.id_offset, ix.x[ARENA_T].as<float>(),
                                                     ix.ids[ARENA_T].as<uint32_t>(), ix.xnorm[ARENA_T].as<float>(), ix.inv_t.as<uint32_t>());
        }
        if (rc == cudaSuccess) {
            k_iota<<<(n + 255) / 256, 256, 0, st>>>(perm_in.as<uint32_t>(), n);
            c = radix_sort_pairs<uint64_t>(key_ct_in.as<uint64_t>(), perm_in.as<uint32_t>(), ix.keys_ct.as<uint64_t>(), perm_out.as<uint32_t>(), n, 64, tmp.p, st);
            if (c != cudaSuccess) fail(c, "radix sort (C,T)");
            k_gather<<<gather_blocks, 256, 0, st>>>(rows, perm_out.as<uint32_t>(), n, ix.id_offset, ix.x[ARENA_CT].as<float>(),
                                                     ix.ids[ARENA_CT].as<uint32_t>(), ix.xnorm[ARENA_CT].as<float>(), nullptr);
        }
        // Outliers: the few rows (<= OUTLIER_MAX) whose norm lies above a power-of-two cutoff that all other rows stay
        // below.  They are taken out of the approximate sweeps and scored exactly in K5, so that ONE huge row cannot set
        // the fp16 scale and the candidate margins of the whole data set (hvs_margin.cuh).  Non-finite rows always count.
        float cutoff = __builtin_inff();
        if (rc == cudaSuccess) {
            cudaMemsetAsync(hist.p, 0, 256 * 4, st);
            k_norm_hist<<<296, 256, 0, st>>>(ix.xnorm[ARENA_T].as<float>(), n, hist.as<uint32_t>());
            uint32_t h[256];
            c = cudaMemcpyAsync(h, hist.p, sizeof h, cudaMemcpyDeviceToHost, st);
            if (c == cudaSuccess) c = cudaStreamSynchronize(st);
            if (c != cudaSuccess) fail(c, "norm histogram");
            else {
                uint64_t above = 0;                       // rows with exponent field > E
                int E = 254;
                while (E > 0 && above + h[E] <= (uint64_t)OUTLIER_MAX - h[255]) { above += h[E]; --E; }
                if (h[255] > (uint32_t)OUTLIER_MAX) E = -1;     // too many non-finite rows: no approximate sweeps at all (see below)
                else if (above + h[255] > 0 && E < 254) cutoff = std::ldexp(1.0f, E + 1 - 127);   // rows with exponent <= E are inliers
                if (E == -1) cutoff = 0.f;
            }
        }
        if (rc == cudaSuccess && cutoff > 0.f && cutoff < __builtin_inff()) {
            for (int a = 0; a < 2 && rc == cudaSuccess; ++a) {
                cudaMemsetAsync(ol_tmp.p, 0, 4, st);
                k_mark_outliers<<<(n + 255) / 256, 256, 0, st>>>(ix.xnorm[a].as<float>(), n, cutoff, ol_tmp.as<uint32_t>() + 1,
                                                                 ol_tmp.as<uint32_t>(), OUTLIER_MAX);
                uint32_t buf[OUTLIER_MAX + 1];
                c = cudaMemcpyAsync(buf, ol_tmp.p, sizeof buf, cudaMemcpyDeviceToHost, st);
                if (c == cudaSuccess) c = cudaStreamSynchronize(st);
                if (c != cudaSuccess) { fail(c, "outlier list"); break; }
                const uint32_t cnt = buf[0] < (uint32_t)OUTLIER_MAX ? buf[0] : (uint32_t)OUTLIER_MAX;
                std::sort(buf + 1, buf + 1 + cnt);
                ix.n_outl[a] = cnt;
                c = cudaMemcpyAsync(ix.outl[a].p, buf + 1, (size_t)cnt * 4, cudaMemcpyHostToDevice, st);
                if (c == cudaSuccess) c = cudaStreamSynchronize(st);        // buf is on the stack
                if (c != cudaSuccess) fail(c, "outlier upload");
            }
        }
        approx_ok = cutoff > 0.f;
        if (rc == cudaSuccess) {
            cudaMemsetAsync(maxbits.p, 0, 4, st);
            k_max_f32<<<296, 256, 0, st>>>(ix.xnorm[ARENA_T].as<float>(), n, maxbits.as<uint32_t>());
        }
    }
    if (rc == cudaSuccess) {
        k_tail<<<K, 128, 0, st>>>(rows, n_total, ix.tail.as<float>());
        uint32_t bits = 0;
        if (n) {
            cudaError_t c = cudaMemcpyAsync(&bits, maxbits.p, 4, cudaMemcpyDeviceToHost, st);
            if (c != cudaSuccess) fail(c, "memcpy xnorm_max");
        }
        cudaError_t c = cudaStreamSynchronize(st);
        if (c != cudaSuccess) fail(c, "index build sync");
        union { uint32_t u; float f; } cv; cv.u = bits;
        ix.xnorm_max = cv.f;
    }
    if (rc == cudaSuccess) {
        cudaError_t c = cudaGetLastError();
        if (c != cudaSuccess) fail(c, "index build kernels");
    }
    if (rc != cudaSuccess) return rc;
    ix.xb[0].release(); ix.xb[1].release();
    ix.img_scale = 1.f;
    ix.approx_ok = approx_ok && std::isfinite(ix.xnorm_max);
    if (ix.approx_ok && n && n < 0x80000000u) {
        // sx = 2^e with sx^2 max||x||^2 <= 32000: every image element and every split-norm term fits fp16.  An exponent
        // beyond +-40 (norms around 1e-20 or 1e28) would break that invariant if clamped: such data gets no image and
        // AUTO mode sweeps it with the FP32 tile kernel instead.
        int ex = 0;
        if (ix.xnorm_max > 0.f) ex = (int)std::floor(0.5 * std::log2(32000.0 / (double)ix.xnorm_max));
        if (ex >= -40 && ex <= 40) {
            ix.img_scale = std::ldexp(1.0f, ex);
            cudaError_t c = build_tensor_image(e, ARENA_T);
            if (c == cudaSuccess) c = build_tensor_image(e, ARENA_CT);
            if (c == cudaSuccess) c = cudaStreamSynchronize(st);
            if (c == cudaErrorMemoryAllocation) {             // no room for the images: the engine works without them (K2 / K4)
                cudaGetLastError();
                ix.xb[0].release(); ix.xb[1].release();
                ix.img_scale = 1.f;
            } else if (c != cudaSuccess) { e->err = std::string("fp16 image: ") + cudaGetErrorString(c); return c; }
        }
    }
    ix.built = true;
    return cudaSuccess;
}

}  // namespace hvs
