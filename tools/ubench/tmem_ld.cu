// tcgen05.ld / tcgen05.st throughput per SM (sm_100a).  148 CTAs, NW warps each (warp w reads the TMEM lane quadrant w % 4,
// column block (w / 4) * 128), every warp issues `iters` loads of the given width back to back (one tcgen05.wait::ld per
// `batch` loads).  Prints bytes per clock per SM.    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W>   // W = 16 or 32 columns per load
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                   "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

template <int W, int MODE>   // MODE 0: loads, 1: stores (x16), 2: loads + 16 MUFU.EX2 per 16 columns
__global__ void bench(int iters, int batch, long long* clocks, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t*)&slot;
    const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    st16(addr, r); st16(addr + 16, r); st16(addr + 32, r + 16); st16(addr + 48, r + 16);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += batch) {
        for (int b = 0; b < batch; ++b) {
            const uint32_t a = addr + (uint32_t)(((i + b) & 1) * 32);
            if (MODE == 1) { st16(a, r); }
            else ld<W>(a, r);
        }
        if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < W; ++k) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(r[k]))); acc += y; }
        } else acc += __uint_as_float(r[0] ^ r[W - 1]);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

template <int W, int MODE>
void run(const char* name, int nw, int batch) {
    const int iters = 4096, blocks = 148;
    long long* clk; float* sink;
    cudaMalloc(&clk, blocks * 8); cudaMalloc(&sink, blocks * nw * 32 * 4);
    bench<W, MODE><<<blocks, nw * 32>>>(64, batch, clk, sink);
    bench<W, MODE><<<blocks, nw * 32>>>(iters, batch, clk, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, blocks * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += (double)h[i]; mean /= blocks;
    const double bytes = (double)iters * nw * 32 * (MODE == 1 ? 16 : W) * 4;
    printf("%-28s warps %2d batch %d : %8.1f clk  %7.1f B/clk/SM  %s\n", name, nw, batch, mean, bytes / mean, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(clk); cudaFree(sink);
}

int main() {
    for (int nw : {4, 8, 16}) {
        run<16, 0>("tcgen05.ld x16", nw, 1);
        run<16, 0>("tcgen05.ld x16", nw, 4);
        run<32, 0>("tcgen05.ld x32", nw, 1);
        run<32, 0>("tcgen05.ld x32", nw, 2);
        run<16, 1>("tcgen05.st x16", nw, 4);
        run<16, 2>("tcgen05.ld x16 + 16 ex2", nw, 1);
        run<32, 2>("tcgen05.ld x32 + 32 ex2", nw, 1);
    }
    return 0;
}
