// fp32-accumulate SIMT GEMM:  C[M,N] = epilogue(A[M,K] * W[N,K]^T).
// This is the exact-arithmetic path of the engine (WLK_PREC_FP32: fp32 operands, the
// parity mode that holds 1e-3 on logits against the reference CPU backend) and the
// small-M path of the bf16 mode.  Classic smem-tiled register-blocked kernel; operands
// are converted to fp32 on the way into shared memory.
#include "common.cuh"

namespace wlk {

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float* o) {
        float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ void load(const bf16* p, float* o) {
        uint2 u = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
        o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
    }
};

// 256 threads; thread tile TM x TN; block tile BM x BN; BK-wide k slabs.
template <typename TA, typename TW, int BM, int BN, int BK, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, int64_t lda,
                                                        const TW* __restrict__ W, int64_t ldw,
                                                        int M, int N, int K, Epilogue epi) {
    static_assert((BM / TM) * (BN / TN) == 256, "thread layout");
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        if (VEC) {
            constexpr int KV = BK / 4;
            for (int idx = tid; idx < BM * KV; idx += 256) {
                int r = idx / KV, kq = (idx % KV) * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                int m = m0 + r, k = k0 + kq;
                if (m < M && k < K) Vec4<TA>::load(A + (int64_t)m * lda + k, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) As[kq + q][r] = v[q];
            }
            for (int idx = tid; idx < BN * KV; idx += 256) {
                int r = idx / KV, kq = (idx % KV) * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                int n = n0 + r, k = k0 + kq;
                if (n < N && k < K) Vec4<TW>::load(W + (int64_t)n * ldw + k, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) Bs[kq + q][r] = v[q];
            }
        } else {
            for (int idx = tid; idx < BM * BK; idx += 256) {
                int r = idx / BK, kk = idx % BK;
                int m = m0 + r, k = k0 + kk;
                As[kk][r] = (m < M && k < K) ? to_f32(A[(int64_t)m * lda + k]) : 0.f;
            }
            for (int idx = tid; idx < BN * BK; idx += 256) {
                int r = idx / BK, kk = idx % BK;
                int n = n0 + r, k = k0 + kk;
                Bs[kk][r] = (n < N && k < K) ? to_f32(W[(int64_t)n * ldw + k]) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= M) continue;
        int nb = n0 + tx * TN;
        if (TN == 8 && nb + 8 <= N && (N % 8) == 0) {
            epi_store8(epi, m, nb, acc[i]);
        } else {
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (nb + j < N) epi_store1(epi, m, nb + j, acc[i][j]);
        }
    }
}

template <typename TA, typename TW>
static void launch(const GemmArgs& g, cudaStream_t st) {
    const TA* A = reinterpret_cast<const TA*>(g.A);
    const TW* W = reinterpret_cast<const TW*>(g.W);
    const size_t va = sizeof(TA) * 4;
    const size_t vw = sizeof(TW) * 4;
    bool vec = (g.K % 4 == 0) && (g.lda % 4 == 0) && (g.ldw % 4 == 0) &&
               (reinterpret_cast<uintptr_t>(A) % va == 0) && (reinterpret_cast<uintptr_t>(W) % vw == 0);
    if (g.M > 48) {
        dim3 grid((g.N + 127) / 128, (g.M + 127) / 128);
        if (vec) gemm_simt_kernel<TA, TW, 128, 128, 16, 8, 8, true><<<grid, 256, 0, st>>>(A, g.lda, W, g.ldw, g.M, g.N, g.K, g.epi);
        else     gemm_simt_kernel<TA, TW, 128, 128, 16, 8, 8, false><<<grid, 256, 0, st>>>(A, g.lda, W, g.ldw, g.M, g.N, g.K, g.epi);
    } else {
        dim3 grid((g.N + 63) / 64, (g.M + 15) / 16);
        if (vec) gemm_simt_kernel<TA, TW, 16, 64, 32, 1, 4, true><<<grid, 256, 0, st>>>(A, g.lda, W, g.ldw, g.M, g.N, g.K, g.epi);
        else     gemm_simt_kernel<TA, TW, 16, 64, 32, 1, 4, false><<<grid, 256, 0, st>>>(A, g.lda, W, g.ldw, g.M, g.N, g.K, g.epi);
    }
    CUDA_CHECK(cudaGetLastError());
}

void gemm_simt(const GemmArgs& g, cudaStream_t st) {
    WLK_CHECK(g.M > 0 && g.N > 0 && g.K > 0, "gemm_simt: empty problem %d %d %d", g.M, g.N, g.K);
    if (g.a_type == DT_F32 && g.w_type == DT_F32) launch<float, float>(g, st);
    else if (g.a_type == DT_BF16 && g.w_type == DT_BF16) launch<bf16, bf16>(g, st);
    else if (g.a_type == DT_F32 && g.w_type == DT_BF16) launch<float, bf16>(g, st);
    else launch<bf16, float>(g, st);
}

}  // namespace wlk
