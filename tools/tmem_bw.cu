// tmem_bw.cu -- microbenchmark: TMEM -> register bandwidth of tcgen05.ld on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tools/tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0: 32x32b.x32   1: 32x32b.x32.pack::16b   2: 32x32b.x64   3: 32x32b.x128
__global__ void k(int iters, long long *out, uint32_t *sink)
{
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const uint32_t col = (uint32_t)((i * 128 + (warp >> 2) * 256) & 511) & ~127u;
        if (MODE == 0 || MODE == 1) {
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (MODE == 0)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),"=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                                 : "r"(t + col + c * 32));
                else
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),"=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                                 : "r"(t + col + (c & 1) * 64));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= r[j];
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345u) sink[0] = acc;
    if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

int main()
{
    long long *out; uint32_t *sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 64);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            cudaMemset(out, 0, 64 * 8);
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            if (mode == 0) k<0><<<1, warps * 32>>>(iters, out, sink); else k<1><<<1, warps * 32>>>(iters, out, sink);
            cudaEventRecord(b);
            cudaError_t e = cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, a, b);
            long long h[64]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
            // per iteration per warp: 4 loads; mode 0: 4 x 4 KB of fp32 cells; mode 1: 4 x 64 columns (8 KB of cells, 4 KB of registers)
            const double regbytes = (double)iters * 4 * 4096 * warps;
            const double cells = mode == 0 ? regbytes : 2 * regbytes;
            printf("mode %d (%s) warps %2d: %s  %.3f ms  clock64 cycles/warp %lld  -> %.1f B/clk/SM of TMEM cells, %.1f B/clk/SM into registers (at 1.965 GHz: %.1f cells-B/clk)\n",
                   mode, mode ? "x32.pack::16b" : "x32", warps, cudaGetErrorString(e), ms, h[0], cells / (double)h[0], regbytes / (double)h[0],
                   cells / (ms * 1e-3 * 1.965e9));
        }
    return 0;
}
