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
// BF16 operands in the tcgen05 pass (K3).  bf16 keeps 8 significant bits, round-to-nearest: unit
// roundoff 2^-8 per operand, so each product q_i x_i is off by at most (2^-7 + 2^-16)|q_i x_i| and
// |sum - q.x| <= (2^-7 + 2^-16) ||q|| ||x|| (Cauchy-Schwarz); the score carries -2 q.x.  The split
// ||x||^2 (three bf16 terms) and the tensor core's fp32 accumulation add only fp32-level terms,
// covered generously by 4 x the K2 margin.
__host__ __device__ inline float margin_tensor(float qnorm, float xnorm_max)
{
    float eps = 2.0f * (1.0f / 128.0f + 1.0f / 65536.0f) * sqrtf(qnorm) * sqrtf(xnorm_max);
    return 2.0f * eps * 1.02f + margin_ffma(qnorm, xnorm_max) * 4.0f;
}

}  // namespace hvs
