// tcgen05 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T), bf16 operands, fp32 accumulate.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer  : cp.async.bulk.tensor 2D loads of the A (128x64) and W (BNx64) k-slabs,
//                               128B-swizzled, into a STAGES-deep shared-memory ring (mbarrier full/empty)
//   warp 1      MMA issuer    : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN,
//                               K=16) x4 per slab; accumulators live in TMEM, double-buffered (2 x BN cols)
//   warps 2..9  epilogue      : tcgen05.ld the finished accumulator (lane = row), fused bias / GELU /
//                               column scale / fp32 residual, then bf16 or fp32 stores (plain or scattered
//                               into the head-major KV layouts), overlapping the next tile's MMAs.
// Tiles are walked m-fastest so the CTAs of a wave share one W panel (L2-resident) while A streams.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "ptx.cuh"

namespace wlk {

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        WLK_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    return g_encode;
}

// 2D bf16 tensor map over a row-major [rows, cols] matrix with row pitch ld (elements); box = [box_rows, 64].
bool make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols, std::string* err) {
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
        return false;
    }
    return true;
}

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;   // TMA warp, MMA warp, 8 epilogue warps
constexpr int STG_PITCH = 17;

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr uint32_t A_BYTES = BM * BK * 2;
    static constexpr uint32_t B_BYTES = BN * BK * 2;
    static constexpr uint32_t STG_OFF = STAGES * (A_BYTES + B_BYTES);     // epilogue transpose staging
    static constexpr uint32_t STG_BYTES = NUM_EPI_WARPS * 32 * STG_PITCH * 4;   // per warp [32 rows][16 cols (+1)] fp32
    static constexpr uint32_t ROW_OFF = STG_OFF + STG_BYTES;                     // per warp EpiRow[32]
    static constexpr uint32_t ROW_BYTES = NUM_EPI_WARPS * 32 * 24;
    static constexpr uint32_t BAR_OFF = ROW_OFF + ROW_BYTES;
    static constexpr uint32_t TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16;
    static constexpr uint32_t DYN = TOTAL + 1024;   // slack for manual 1024-byte alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int M, int N, int K,
               Epilogue epi) {
    using L = SmemLayout<BN, STAGES>;
    constexpr uint32_t TMEM_COLS = 2 * BN;           // 128 / 256 / 512: powers of two >= 32
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const uint32_t sA = smem_base;
    const uint32_t sB = smem_base + STAGES * L::A_BYTES;
    const uint32_t bar_full = smem_base + L::BAR_OFF;              // [STAGES]
    const uint32_t bar_empty = bar_full + STAGES * 8;              // [STAGES]
    const uint32_t bar_tfull = bar_empty + STAGES * 8;             // [2]
    const uint32_t bar_tempty = bar_tfull + 16;                    // [2]
    const uint32_t tmem_slot = bar_tempty + 16;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_k = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmW);
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(bar_full + 8 * i, 1);
            ptx::mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar_tfull + 8 * i, 1);
            ptx::mbar_init(bar_tempty + 8 * i, 32 * NUM_EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % num_m, n_blk = tile / num_m;
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, L::A_BYTES + L::B_BYTES);
                    ptx::tma_load_2d(sA + stage * L::A_BYTES, &tmA, bar_full + 8 * stage, kb * BK, m_blk * BM);
                    ptx::tma_load_2d(sB + stage * L::B_BYTES, &tmW, bar_full + 8 * stage, kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, 0, 0);
        uint32_t stage = 0, phase = 0, it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const uint32_t as = it & 1, ap = (it >> 1) & 1;
            ptx::mbar_wait(bar_tempty + 8 * as, ap ^ 1);
            ptx::tc_fence_after();
            for (int kb = 0; kb < num_k; ++kb) {
                ptx::mbar_wait(bar_full + 8 * stage, phase);
                ptx::tc_fence_after();
                if (lane == 0) {
                    const uint64_t da = ptx::umma_desc_kmajor_sw128(sA + stage * L::A_BYTES);
                    const uint64_t db = ptx::umma_desc_kmajor_sw128(sB + stage * L::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing K by 16 bf16 = 32 bytes = +2 in the (addr >> 4) field
                        ptx::umma_bf16_ss(tmem_base + as * BN, da + 2 * k, db + 2 * k, idesc,
                                          (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(bar_empty + 8 * stage);              // frees the smem slot when MMAs retire
                    if (kb == num_k - 1) ptx::umma_commit(bar_tfull + 8 * as);   // accumulator ready
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // Two warps per TMEM lane quadrant, each owning half of the tile's columns.  tcgen05.ld hands a
        // thread one accumulator ROW (16 columns at a time); the warp transposes that 32x16 block through
        // a private padded smem tile so that lanes walk consecutive COLUMNS: every global access (output
        // row, fp32 residual row, scattered KV row) is a contiguous 32-64 B run per half-warp, the bias /
        // scale of a column live in registers, and all div/mod addressing is hoisted to once per row per
        // tile (EpiRow) and once per chunk (epi_col).
        const int ew = warp - 2;
        const int q = warp & 3;                       // TMEM lane quadrant this warp may access
        const int hc = ew >> 2;                       // which half of the tile's columns
        float* stg = reinterpret_cast<float*>(smem_gen + L::STG_OFF) + ew * (32 * STG_PITCH);
        EpiRow* rowinfo = reinterpret_cast<EpiRow*>(smem_gen + L::ROW_OFF) + ew * 32;
        const int rsel = lane >> 4, col = lane & 15;
        const int es = (epi.c_type == DT_F32) ? 4 : 2;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            const uint32_t as = it & 1, ap = (it >> 1) & 1;
            const int row_base = m_blk * BM + q * 32;
            rowinfo[lane] = epi_row(epi, row_base + lane, M);
            __syncwarp();
            ptx::mbar_wait(bar_tfull + 8 * as, ap);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int c0 = hc * (BN / 2) + c * 16;            // first column of this chunk inside the tile
                uint32_t r[16];
                ptx::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c0, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[lane * STG_PITCH + j] = __uint_as_float(r[j]);
                __syncwarp();
                const int n = n_blk * BN + c0 + col;
                if (n < N) {
                    int variant;
                    const int64_t coff = epi_col(epi, n, &variant) * es;
                    const float bias_v = epi.bias ? __ldg(epi.bias + n) : 0.f;
                    const bool scaled = (epi.scale_period ? (n % epi.scale_period) : n) < epi.scale_cols;
                    const float scale_v = scaled ? epi.col_scale : 1.f;
#pragma unroll 4
                    for (int i2 = 0; i2 < 16; ++i2) {
                        const int i = 2 * i2 + rsel;
                        const EpiRow ri = rowinfo[i];
                        char* p = variant ? ri.ptr1 : ri.ptr0;
                        if (p == nullptr) continue;
                        float v = stg[i * STG_PITCH + col] + bias_v;
                        if (epi.gelu) v = gelu_erf(v);
                        v *= scale_v;
                        if (ri.res) v += ri.res[n];
                        if (es == 4) *reinterpret_cast<float*>(p + coff) = v;
                        else *reinterpret_cast<bf16*>(p + coff) = __float2bfloat16_rn(v);
                    }
                }
                __syncwarp();
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar_tempty + 8 * as);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES>
void launch(const GemmArgs& g, cudaStream_t st, int num_sms) {
    using L = SmemLayout<BN, STAGES>;
    CUtensorMap tmA, tmW;
    std::string err;
    WLK_CHECK(make_tmap_bf16_2d(&tmA, g.A, g.M, g.K, g.lda, BM, BK, &err), "A tensor map: %s", err.c_str());
    WLK_CHECK(make_tmap_bf16_2d(&tmW, g.W, g.N, g.K, g.ldw, BN, BK, &err), "W tensor map: %s", err.c_str());
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)L::DYN));
        attr_set = true;
    }
    const int num_tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
    const int grid = num_tiles < num_sms ? num_tiles : num_sms;
    gemm_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, L::DYN, st>>>(tmA, tmW, g.M, g.N, g.K, g.epi);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace

bool gemm_tcgen05_supported(const GemmArgs& g, std::string* why) {
    auto fail = [&](const char* m) { if (why) *why = m; return false; };
    if (g.a_type != DT_BF16 || g.w_type != DT_BF16) return fail("operands must be bf16");
    if (g.lda % 8 || g.ldw % 8) return fail("row pitch must be a multiple of 8 elements (16 bytes)");
    if (reinterpret_cast<uintptr_t>(g.A) % 16 || reinterpret_cast<uintptr_t>(g.W) % 16) return fail("operands must be 16-byte aligned");
    if (g.K % 8) return fail("K must be a multiple of 8");
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return fail("empty problem");
    return true;
}

void gemm_tcgen05(const GemmArgs& g, cudaStream_t st, int num_sms) {
    std::string why;
    WLK_CHECK(gemm_tcgen05_supported(g, &why), "gemm_tcgen05: %s", why.c_str());
    // BN=256 keeps the tensor pipe busiest per smem byte; narrow outputs use smaller tiles so the
    // grid still covers the SMs.
    const int tiles256 = ((g.M + BM - 1) / BM) * ((g.N + 255) / 256);
    if (g.N >= 256 && tiles256 >= num_sms) launch<256, 4>(g, st, num_sms);
    else if (g.N >= 128 && ((g.M + BM - 1) / BM) * ((g.N + 127) / 128) >= num_sms / 2) launch<128, 6>(g, st, num_sms);
    else launch<64, 8>(g, st, num_sms);
}

}  // namespace wlk
