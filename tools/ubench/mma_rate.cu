// Issue-to-completion cost of the tcgen05.mma shapes the attention kernel uses (sm_100a), one CTA per SM, one issuing thread:
//   SS  M128 N128 K16, A and B K-major 128B-swizzled smem tiles          (S = Q K^T)
//   TS  M128 N64  K16, A from TMEM, B MN-major 128B-swizzled smem tile   (O += P V as the kernel does it)
//   TS  M128 N64  K16, A from TMEM, B K-major                            (what a transposed V would allow)
//   SS  M128 N64  K16, B MN-major                                        (P through shared memory instead of TMEM)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../whisperlivekit_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace wlk;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS N128 (4 MMAs per group, K = 64)   1: TS N64 B MN-major (8 per group, K = 128)
//      2: TS N64 B K-major (8 per group)       3: SS N64 B MN-major (8 per group)
__global__ void __launch_bounds__(128, 1) bench(int mode, int groups, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot;
    __shared__ uint64_t bar_storage;
    const uint32_t bar = ptx::smem_u32(&bar_storage);
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - ptx::smem_u32(raw)))[i] = 0x3c003c00u;
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t*)&slot;
    if (threadIdx.x == 0) {
        const uint32_t sA = base, sB = base + 16384, sV = base + 32768;
        constexpr uint32_t id_s = ptx::umma_idesc_bf16(128, 128, 0, 0), id_o_mn = ptx::umma_idesc_bf16(128, 64, 0, 1),
                           id_o_k = ptx::umma_idesc_bf16(128, 64, 0, 0);
        const uint64_t dA = ptx::umma_desc_kmajor_sw128(sA), dB = ptx::umma_desc_kmajor_sw128(sB);
        const uint64_t dVmn = ptx::umma_desc_mnmajor_sw128(sV, 128 * 128), dVk = ptx::umma_desc_kmajor_sw128(sV);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            if (mode == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_bf16_ss(tmem, dA + 2 * k, dB + 2 * k, id_s, k > 0);
            } else if (mode == 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ts(tmem + 128, tmem + 192 + 8 * k, dVmn + 128 * k, id_o_mn, 1);
            } else if (mode == 2) {
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ts(tmem + 128, tmem + 192 + 8 * k, dVk + 2 * (k & 3), id_o_k, 1);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) ptx::umma_bf16_ss(tmem + 128, dA + 2 * (k & 3), dVmn + 128 * k, id_o_mn, 1);
            }
        }
        const long long t1 = clock64();                      // all issued
        ptx::umma_commit(bar);
        ptx::mbar_wait(bar, 0);
        const long long t2 = clock64();                      // all complete
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
    const char* names[4] = {"SS M128 N128 K-major/K-major x4 (S)", "TS M128 N64 B MN-major x8 (P V)", "TS M128 N64 B K-major x8",
                            "SS M128 N64 B MN-major x8"};
    long long* d; cudaMalloc(&d, 148 * 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mode = 0; mode < 4; ++mode) {
        for (int groups : {1, 64}) {
            bench<<<148, 128, 64 * 1024>>>(mode, groups, d);
            bench<<<148, 128, 64 * 1024>>>(mode, groups, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double a = 0, b = 0; for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
            const int n = groups * (mode == 0 ? 4 : 8);
            printf("%-40s groups %3d: issue %8.1f clk, complete %8.1f clk = %6.1f clk per MMA  %s\n", names[mode], groups, a / 148, b / 148,
                   b / 148 / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
