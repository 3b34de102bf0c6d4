// Microbenchmark: MUFU.EX2 / FFMA / FMNMX / F2FP issue rates per SM on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mufu tools/ubench/mufu.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
            if (OP == 2) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
            if (OP == 3) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(r); }
            if (OP == 4) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
            if (OP == 5) { unsigned r = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
            if (OP == 6) { unsigned r = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
            if (OP == 7) { unsigned r = __float_as_uint(a[i]); asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
            if (OP == 8) { unsigned r = __float_as_uint(a[i]), q = __float_as_uint(a[(i + 1) & 7]); asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(r) : "r"(q)); a[i] = __uint_as_float(r); }
            if (OP == 9) { unsigned r = __float_as_uint(a[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;
}

template <int OP>
void run(const char* name, int threads) {
    float* out; cudaMalloc(&out, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<p.multiProcessorCount, threads>>>(out, 16, 0.5f);
    cudaEventRecord(e0);
    k<OP><<<p.multiProcessorCount, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ops = (double)p.multiProcessorCount * threads * iters * 8;
    printf("%-8s threads/SM=%4d  %.3f ms  %.1f Gop/s  %.2f ops/clk/SM (at nominal %d MHz)\n", name, threads, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000);
}

int main() {
    for (int t : {128, 256, 512}) {
        run<0>("ex2", t); run<1>("ffma", t); run<2>("fmnmx", t); run<3>("f2fp", t); run<4>("fadd", t);
        run<5>("ex2bf16x2", t); run<6>("ex2f16x2", t); run<7>("hfma2bf16", t); run<8>("hmnmx2bf16", t); run<9>("tanhbf16x2", t);
    }
    return 0;
}
