// hvs_common.cuh -- shared device/host definitions for the B200 filtered k-NN engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hvs {

constexpr int K = 100;          // KNN_LIMIT (reference include/optimized_impl.h:26)
constexpr int DIM = 100;        // vector dims (VEC_DIM - 2, include/optimized_impl.h:28)
constexpr int DROW = 102;       // [C, T, x0..x99]
constexpr int QROW = 104;       // [type, v, l, r, q0..q99]
constexpr int ROW_BYTES = DIM * 4;   // one arena row: 400 B, 16-byte aligned, contiguous

// ---- slices -------------------------------------------------------------------------------
// Every predicate of include/baseline.hpp:107-136 resolves to one contiguous row range of one
// of two re-ordered copies ("arenas") of D.
constexpr uint32_t ARENA_T = 0;   // rows ordered by T            (types 0 and 2)
constexpr uint32_t ARENA_CT = 1;  // rows ordered by (C, T)       (types 1 and 3)

struct QSlice {          // one per query, produced by k_plan_search
    uint32_t arena;
    uint32_t begin;      // first arena row
    uint32_t end;        // one past the last arena row (end >= begin)
    float qnorm;         // ||q||^2 (fp32)
};

// Arena: a re-ordered, re-laid-out copy of the indexed rows.
struct Arena {
    const float *x;          // [n][100] fp32, 400-byte rows
    const uint32_t *ids;     // [n] original row id (+ id_offset)
    const float *xnorm;      // [n] ||x||^2
    const void *xb;          // fp16 image for the tensor path (may be null)
    const uint32_t *outl;    // ascending arena positions of the norm-outlier rows (xnorm = +inf there): K5 scores them exactly
    uint32_t n_outl;
};

// ---- order-preserving float keys ------------------------------------------------------------
// -0.0 is folded onto +0.0 because the reference's predicates are IEEE comparisons
// (`nodes[j][0] == v`, `nodes[j][1] >= l`), under which the two zeros are equal.
__host__ __device__ inline uint32_t f32_bits(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
__host__ __device__ inline uint32_t ord_key(float f)
{
    uint32_t u = f32_bits(f);
    if (u == 0x80000000u) u = 0u;                       // -0.0 -> +0.0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // NaNs land outside [ord(-inf), ord(+inf)]
}
__host__ __device__ inline bool f32_isnan(float f) { return (f32_bits(f) & 0x7fffffffu) > 0x7f800000u; }

// float -> uint32 / int32 as the reference's x86-64 build converts them (baseline.hpp:90-91);
// out-of-range / NaN pinned to cvttss2si behaviour (64-bit convert, low 32 bits; int32 "indefinite").
__host__ __device__ inline uint32_t f2u32_x86(float f)
{
    if (!(f > -9.2233720368547758e18f && f < 9.2233720368547758e18f)) return 0u;
    return (uint32_t)(unsigned long long)(long long)f;
}
__host__ __device__ inline int32_t f2i32_x86(float f)
{
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return (int32_t)0x80000000;
    return (int32_t)f;
}

#ifdef __CUDACC__
// ---- reference arithmetic -------------------------------------------------------------------
// include/baseline.hpp:53-64 / include/io.h:38-48: sum_{i=0..99} (x_i - q_i)^2, sequential fp32,
// separate sub / mul / add (the reference build has no FMA: CMakeLists.txt:8 has -mavx2 only).
// The _rn intrinsics are never contracted by nvcc, so this is bit-identical to the CPU result.
__device__ __forceinline__ float ref_accum4(float s, float4 x, float4 q)
{
    float d;
    d = __fsub_rn(x.x, q.x); s = __fadd_rn(s, __fmul_rn(d, d));
    d = __fsub_rn(x.y, q.y); s = __fadd_rn(s, __fmul_rn(d, d));
    d = __fsub_rn(x.z, q.z); s = __fadd_rn(s, __fmul_rn(d, d));
    d = __fsub_rn(x.w, q.w); s = __fadd_rn(s, __fmul_rn(d, d));
    return s;
}
// x: 16-byte aligned 100-float row (global or shared), q: 16-byte aligned 100 floats
__device__ __forceinline__ float ref_dist_row(const float *__restrict__ x, const float *__restrict__ q)
{
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const float4 *q4 = reinterpret_cast<const float4 *>(q);
    float s = 0.0f;
#pragma unroll 5
    for (int i = 0; i < DIM / 4; ++i) s = ref_accum4(s, x4[i], q4[i]);
    return s;
}

// ---- mbarrier + bulk async copy (TMA engine, 1-D form: SASS UBLKCP) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
// global -> shared bulk copy, completion signalled on an mbarrier (complete_tx::bytes).
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same with an L2 cache policy (createpolicy): evict_last keeps tiles that other CTAs are about to read resident
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_u64_hint(uint64_t *p, uint64_t v, uint64_t policy)
{
    asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ uint64_t pack_key(float d, uint32_t v)
{
    // distances are sums of squares (>= 0), so the raw bit pattern orders like the value;
    // NaN sorts after +inf.
    return ((uint64_t)__float_as_uint(d) << 32) | v;
}
#endif  // __CUDACC__

}  // namespace hvs
