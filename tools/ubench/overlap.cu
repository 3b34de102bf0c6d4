// Do tcgen05.mma streams and the softmax warps' TMEM loads + MUFU.EX2 overlap on one SM (sm_100a)?
// One CTA per SM, 18 warps: warp 1 issues groups of [4 SS MMAs M128 N128 (S = Q K^T) + 8 TS MMAs M128 N64 (P V)] back to back,
// warps 2..17 run `iters` rounds of tcgen05.ld x16 -> wait -> 16 MUFU.EX2 (+ a tcgen05.st x8 of the packed result, as the
// kernel does).  Three runs: MMA stream alone, softmax loop alone, both together.  If they overlapped perfectly the combined
// time would be the maximum of the two; if the TMEM port serialises them, the sum.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../whisperlivekit_b200/csrc -o overlap overlap.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace wlk;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__global__ void __launch_bounds__(576, 1) bench(int mma_groups, int sm_iters, int sm_warps, long long* out, float* sink) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot;
    __shared__ uint64_t bar_storage;
    const uint32_t bar = ptx::smem_u32(&bar_storage);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
    if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - ptx::smem_u32(raw)))[i] = 0x3c003c00u;
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t*)&slot;
    long long t0 = clock64(), t1 = t0;
    if (warp == 1) {
        if (lane == 0 && mma_groups > 0) {
            const uint32_t sA = base, sB = base + 16384, sV = base + 32768;
            constexpr uint32_t id_s = ptx::umma_idesc_bf16(128, 128, 0, 0), id_o = ptx::umma_idesc_bf16(128, 64, 0, 1);
            const uint64_t dA = ptx::umma_desc_kmajor_sw128(sA), dB = ptx::umma_desc_kmajor_sw128(sB);
            const uint64_t dV = ptx::umma_desc_mnmajor_sw128(sV, 128 * 128);
            for (int g = 0; g < mma_groups; ++g) {
                const uint32_t tb = tmem + (g & 1) * 256;                 // alternate the two query tiles' TMEM halves
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_bf16_ss(tb, dA + 2 * k, dB + 2 * k, id_s, k > 0);
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ts(tb + 128, tb + 192 + 8 * k, dV + 128 * k, id_o, 1);
            }
            ptx::umma_commit(bar);
            ptx::mbar_wait(bar, 0);
            t1 = clock64();
        }
    } else if (warp >= 2 && warp < 2 + sm_warps && sm_iters > 0) {
        // softmax-like loop on the OTHER columns than the MMAs of the same moment would be ideal; here: tile (warp-2)/8,
        // S columns [0,128) of that tile, P columns [192,256): the same TMEM regions the kernel's softmax warps use
        const int t = (warp - 2) >> 3, half = ((warp - 2) & 7) >> 2;
        const uint32_t tb = tmem + t * 256 + ((uint32_t)((warp & 3) * 32) << 16);
        float acc = 0.f;
        uint32_t r[16], pk[8];
        for (int i = 0; i < sm_iters; ++i) {
            ptx::tmem_ld_32x16(tb + half * 64 + (i & 3) * 16, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float y0, y1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(__uint_as_float(r[2 * k]) * 1e-9f));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(__uint_as_float(r[2 * k + 1]) * 1e-9f));
                acc += y0 + y1;
                pk[k] = __float_as_uint(y0) ^ __float_as_uint(y1);
            }
            st8(tb + 192 + half * 32 + (i & 3) * 8, pk);
        }
        ptx::tmem_st_wait();
        t1 = clock64();
        sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    }
    if (lane == 0) out[(blockIdx.x * 18 + warp)] = t1 - t0;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

static void run(const char* what, int groups, int iters, int sm_warps) {
    long long* d; float* sink;
    cudaMalloc(&d, 148 * 18 * 8); cudaMalloc(&sink, 148 * 576 * 4);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; ++rep) bench<<<148, 576, 64 * 1024>>>(groups, iters, sm_warps, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    static long long h[148 * 18]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double mma = 0, sm = 0;
    for (int b = 0; b < 148; ++b) { mma += h[b * 18 + 1]; double mx = 0; for (int w = 2; w < 2 + sm_warps; ++w) mx = mx > h[b * 18 + w] ? mx : h[b * 18 + w]; sm += mx; }
    printf("%-44s mma stream %9.0f clk (%6.1f per S+PV group)   softmax warps %9.0f clk (%5.2f exps/clk/SM)  %s\n", what, mma / 148,
           groups ? mma / 148 / groups : 0.0, sm / 148, iters && sm > 0 ? (double)iters * sm_warps * 32 * 16 / (sm / 148) : 0.0,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(sink);
}

int main() {
    const int G = 256, I = 4 * 2 * 256 / 2;    // 256 (S + PV) groups = 128 key tiles for two query tiles; 16 warps x I x 16 x 32 exps
    run("MMA stream alone", G, 0, 16);
    run("softmax loop alone, 16 warps", 0, 2 * I, 16);
    run("both, 16 softmax warps", G, 2 * I, 16);
    run("softmax loop alone, 8 warps", 0, 2 * I, 8);
    run("both, 8 softmax warps", G, 2 * I, 8);
    return 0;
}
