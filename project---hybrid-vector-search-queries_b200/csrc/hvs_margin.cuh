// hvs_margin.cuh -- error bounds that make the approximate candidate passes EXACT after re-rank.
//
// The tile kernels rank rows by an approximation s~ of s = ||x||^2 - 2 q.x (the query-constant
// ||q||^2 is dropped).  The final answer is ranked by the reference's own arithmetic d_ref
// (include/baseline.hpp:53-64, recomputed bit-identically in K5).  If |s~ + ||q||^2 - d_ref| <= eps
// for every row, then every row of the exact top-100 satisfies  s~ <= s~_(100) + 2*eps, where
// s~_(100) is the 100-th smallest approximation over any superset of candidates seen so far.
// Keeping everything below that line and re-ranking it is therefore exact (DESIGN.md "margins").
#pragma once
#include "hvs_common.cuh"

namespace hvs {

// fp32 norm expansion with FMA accumulation (K2) + the reference's sequential sum:
// both are within gamma_104 * (||x|| + ||q||)^2 of the real value.
__host__ __device__ inline float margin_ffma(float qnorm, float xnorm_max)
{
    float r = sqrtf(qnorm) + sqrtf(xnorm_max);
    float eps = 2.0f * 104.0f * 5.9604645e-8f * r * r;
    return 2.0f * eps * 1.5f + 1e-30f;
}
// FP16 operands in the tcgen05 pass (K3): a_k = fp16(-2 sx q_k), b_k = fp16(sx x_k), sx a power of two.
// fp16 keeps 11 significant bits (unit roundoff 2^-11) and flushes to multiples of 2^-24 below 2^-14,
// so |fl(v) - v| <= 2^-11 |v| + 2^-25.  Summed over the 100 products (Cauchy-Schwarz, ||.||_1 <= 10 ||.||_2)
// and divided by sx^2 to come back to the units of d:
//   |sum fl(a_k) fl(b_k) - a.b| / sx^2  <=  (2^-10 + 2^-22) 2 ||q|| ||x||  +  2^-25 * 10.1 (2||q|| + ||x||) / sx
// The split ||x||^2 (three fp16 terms, exact to below fp32 ulp) and the tensor core's fp32 accumulation
// add only fp32-level terms, covered generously by 4 x the K2 margin.  Result in the units of d.
__host__ __device__ inline float margin_tensor(float qnorm, float xnorm_max, float sx)
{
    const float nq = sqrtf(qnorm), nx = sqrtf(xnorm_max);
    float eps = (1.0f / 1024.0f + 1.0f / 4194304.0f) * 2.0f * nq * nx + 2.98023224e-8f * 10.1f * (2.0f * nq + nx) / sx;
    return 2.0f * eps * 1.02f + margin_ffma(qnorm, xnorm_max) * 4.0f;
}

}  // namespace hvs
