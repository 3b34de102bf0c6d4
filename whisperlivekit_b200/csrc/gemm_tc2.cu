// CTA-pair tcgen05 GEMM for sm_100a (cta_group::2): one 256 x 256 output tile per 2-CTA cluster.
//
// Why pairs: with one CTA per tile every k-slab costs 48 KB of L2->smem traffic per 128x256x64 MACs and the
// smem port carries both the TMA writes and the MMA reads (~190 B/clk, above its 128 B/clk), which held the
// 1-CTA kernel near a third of the tensor peak.  In a pair each CTA stages only its own 128 rows of A and its
// own 128 rows of W per slab (32 KB), the UMMA (M=256, N=256, K=16, issued by the leader CTA) reads both
// halves of W through the pair, and each CTA accumulates its 128 output rows in its own TMEM.
//
//   warp 0  TMA producer (both CTAs)  : cp.async.bulk.tensor...cta_group::2, completion bytes of BOTH CTAs land
//                                       on the leader's "full" barrier
//   warp 1  MMA issuer (leader only)  : tcgen05.mma.cta_group::2.kind::f16; tcgen05.commit...multicast frees
//                                       the smem slot in both CTAs and publishes the accumulator to both
//   warps 2..17 epilogue (both CTAs)   : same fused epilogue as the 1-CTA kernel on the CTA's own 128 rows;
//                                       "accumulator drained" arrives (remotely for the peer) on the leader
#include <cudaTypedefs.h>

#include "gemm_epi.cuh"

namespace wlk {
namespace {

constexpr int BM2 = 256;          // rows per cluster tile (128 per CTA)
constexpr int BN2 = 256;          // columns per cluster tile (each CTA stages 128 rows of W)
constexpr int BK2 = 64;
constexpr int STAGES2 = 6;
constexpr int STAGES2_X3 = 3;           // X3: four tiles per slab (A_hi, A_lo, W_hi, W_lo), see gemm_tc.cu
constexpr int NUM_EPI_WARPS2 = 16;      // 4 per TMEM lane quadrant, each draining a quarter of the tile's columns
constexpr int NUM_THREADS2 = 64 + 32 * NUM_EPI_WARPS2;
constexpr uint32_t A_BYTES2 = 128 * BK2 * 2;     // 16 KB
constexpr uint32_t B_BYTES2 = 128 * BK2 * 2;     // 16 KB
constexpr uint32_t STG_OFF2 = STAGES2 * (A_BYTES2 + B_BYTES2);      // == STAGES2_X3 * 2 * (A_BYTES2 + B_BYTES2)
static_assert(STAGES2 == 2 * STAGES2_X3, "both instantiations share one smem layout");
constexpr uint32_t STG_BYTES2 = NUM_EPI_WARPS2 * EPI_BIAS_FLOATS * 4;
constexpr uint32_t BAR_OFF2 = STG_OFF2 + STG_BYTES2;
constexpr uint32_t SMEM2 = BAR_OFF2 + (2 * STAGES2 + 4) * 8 + 16 + 1024;
constexpr uint32_t TMEM_COLS2 = 512;             // two accumulator stages of 256 fp32 columns
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address (-> CTA 0)

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar, uint32_t cta) {   // arrive on CTA `cta`'s copy of bar
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap* tm, uint32_t bar, int32_t c0,
                                                 int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on `bar` in both CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(static_cast<uint16_t>(3))
        : "memory");
}

template <bool X3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmWlo, int M, int N, int K,
                Epilogue epi, int band_arg) {
    constexpr int NST = X3 ? STAGES2_X3 : STAGES2;
    constexpr uint32_t STAGE_TX = (X3 ? 4u : 2u) * (A_BYTES2 + B_BYTES2);       // bytes landing on a full barrier (both CTAs)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const uint32_t sA = smem_base;
    const uint32_t sB = smem_base + NST * A_BYTES2;
    const uint32_t sAlo = sB + NST * B_BYTES2;                     // X3 only
    const uint32_t sBlo = sAlo + NST * A_BYTES2;                   // X3 only
    const uint32_t bar_full = smem_base + BAR_OFF2;                // [STAGES2]  (used in the leader)
    const uint32_t bar_empty = bar_full + STAGES2 * 8;             // [STAGES2]  (each CTA its own)
    const uint32_t bar_tfull = bar_empty + STAGES2 * 8;            // [2]        (each CTA its own)
    const uint32_t bar_tempty = bar_tfull + 16;                    // [2]        (used in the leader)
    const uint32_t tmem_slot = bar_tempty + 16;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int num_m = (M + BM2 - 1) / BM2, num_n = (N + BN2 - 1) / BN2;
    const int num_tiles = num_m * num_n;
    const int num_k = (K + BK2 - 1) / BK2;
    // Rasterisation.  band_arg >= 1 (host heuristic): tiles run n-fastest inside bands of that many m-blocks.  When the
    // whole weight matrix is L2-resident (every GEMM of a layer: <= 13 MB) the band is ONE m-block: the clusters that
    // run concurrently take all n-blocks of the same few m-blocks and stream the same A slabs at the same time, so each A
    // row block comes from HBM once (ncu, fc2 at 96 streams: 4.4 GB read for a 1.5 GB operand with 19-block bands).
    // 0: two n-blocks of a band in flight at a time; the band's A rows (256 x K bf16 each) should stay L2-resident
    const int band = band_arg >= 1 ? band_arg
                                   : max(1, min(num_clusters / 2, (int)((48u << 20) / (512u * (uint32_t)K))));

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmW);
        if (X3) { ptx::prefetch_tensormap(&tmAlo); ptx::prefetch_tensormap(&tmWlo); }
        for (int i = 0; i < NST; ++i) {
            ptx::mbar_init(bar_full + 8 * i, 2);                   // one arrive per CTA's producer
            ptx::mbar_init(bar_empty + 8 * i, 1);                  // one multicast commit
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar_tfull + 8 * i, 1);
            ptx::mbar_init(bar_tempty + 8 * i, 2 * 32 * NUM_EPI_WARPS2);   // epilogue threads of both CTAs
        }
        ptx::fence_barrier_init();
    }
    cluster_sync();                                                // barriers of both CTAs exist before any remote arrive
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, TMEM_COLS2);
        tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    cluster_sync();                                                // both halves of the paired TMEM allocation are visible
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                int m_blk, n_blk;
                tile_coords(tile, num_m, num_n, band, &m_blk, &n_blk);
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (leader) ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, STAGE_TX);
                    else mbar_arrive_cta(bar_full + 8 * stage, 0);
                    tma_load_2d_pair(sA + stage * A_BYTES2, &tmA, bar_full + 8 * stage, kb * BK2, m_blk * BM2 + rank * 128);
                    tma_load_2d_pair(sB + stage * B_BYTES2, &tmW, bar_full + 8 * stage, kb * BK2, n_blk * BN2 + rank * 128);
                    if (X3) {
                        tma_load_2d_pair(sAlo + stage * A_BYTES2, &tmAlo, bar_full + 8 * stage, kb * BK2, m_blk * BM2 + rank * 128);
                        tma_load_2d_pair(sBlo + stage * B_BYTES2, &tmWlo, bar_full + 8 * stage, kb * BK2, n_blk * BN2 + rank * 128);
                    }
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM2, BN2, 0, 0);
            uint32_t stage = 0, phase = 0, it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const uint32_t as = it & 1, ap = (it >> 1) & 1;
                ptx::mbar_wait(bar_tempty + 8 * as, ap ^ 1);
                ptx::tc_fence_after();
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(bar_full + 8 * stage, phase);
                    ptx::tc_fence_after();
                    if (lane == 0) {
                        const uint64_t da = ptx::umma_desc_kmajor_sw128(sA + stage * A_BYTES2);
                        const uint64_t db = ptx::umma_desc_kmajor_sw128(sB + stage * B_BYTES2);
                        const uint64_t dal = ptx::umma_desc_kmajor_sw128(sAlo + stage * A_BYTES2);
                        const uint64_t dbl = ptx::umma_desc_kmajor_sw128(sBlo + stage * B_BYTES2);
#pragma unroll
                        for (int k = 0; k < BK2 / 16; ++k) {
                            umma_bf16_ss_pair(tmem_base + as * BN2, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            if (X3) {       // + A_lo W_hi + A_hi W_lo into the same accumulator
                                umma_bf16_ss_pair(tmem_base + as * BN2, dal + 2 * k, db + 2 * k, idesc, 1u);
                                umma_bf16_ss_pair(tmem_base + as * BN2, da + 2 * k, dbl + 2 * k, idesc, 1u);
                            }
                        }
                        umma_commit_pair(bar_empty + 8 * stage);
                        if (kb == num_k - 1) umma_commit_pair(bar_tfull + 8 * as);
                    }
                    __syncwarp();
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue (both CTAs, own 128 rows) =====================
        const int ew = warp - 2;
        const int q = warp & 3;
        const int hc = ew >> 2;                                       // which quarter of the columns
        constexpr int COLS_PER_WARP = BN2 / (NUM_EPI_WARPS2 / 4);
        float* sbias = reinterpret_cast<float*>(smem_gen + STG_OFF2) + ew * EPI_BIAS_FLOATS;
        uint32_t it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            int m_blk, n_blk;
            tile_coords(tile, num_m, num_n, band, &m_blk, &n_blk);
            const uint32_t as = it & 1, ap = (it >> 1) & 1;
            const EpiRow row = epi_row(epi, m_blk * BM2 + rank * 128 + q * 32 + lane, M);
            epilogue_prefetch_residual(row, n_blk * BN2 + hc * COLS_PER_WARP, COLS_PER_WARP, N);
            ptx::mbar_wait(bar_tfull + 8 * as, ap);
            ptx::tc_fence_after();
            epilogue_warp_tile(epi, sbias, row, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN2,
                               n_blk * BN2, hc * COLS_PER_WARP, COLS_PER_WARP / 16, N, lane);
            ptx::tc_fence_before();
            mbar_arrive_cta(bar_tempty + 8 * as, 0);
        }
    }

    ptx::tc_fence_before();
    cluster_sync();                                                // neither CTA frees TMEM / exits while the peer still uses it
    if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS2);
}

}  // namespace

template <bool X3>
static void launch_pair(const GemmArgs& g, const void* A_hi, const void* A_lo, cudaStream_t st, int num_sms) {
    CUtensorMap tmA, tmW, tmAlo, tmWlo;
    std::string err;
    WLK_CHECK(make_tmap_bf16_2d(&tmA, A_hi, g.M, g.K, g.lda, 128, BK2, &err), "A tensor map: %s", err.c_str());
    WLK_CHECK(make_tmap_bf16_2d(&tmW, g.W, g.N, g.K, g.ldw, 128, BK2, &err), "W tensor map: %s", err.c_str());
    if (X3) {
        WLK_CHECK(make_tmap_bf16_2d(&tmAlo, A_lo, g.M, g.K, g.lda, 128, BK2, &err), "A_lo tensor map: %s", err.c_str());
        WLK_CHECK(make_tmap_bf16_2d(&tmWlo, g.W_lo, g.N, g.K, g.ldw, 128, BK2, &err), "W_lo tensor map: %s", err.c_str());
    } else { tmAlo = tmA; tmWlo = tmW; }
    static bool seen[64] = {};
    if (first_on_device(seen))
        CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2));
    const int num_tiles = ((g.M + BM2 - 1) / BM2) * ((g.N + BN2 - 1) / BN2);
    int clusters = num_sms / 2;
    if (num_tiles < clusters) clusters = num_tiles;
    static const int forced_band = [] { const char* v = getenv("WLK_GEMM_BAND"); return v ? atoi(v) : -1; }();
    const double w_bytes = (double)g.N * g.K * 2.0 * (X3 ? 2 : 1);
    int band = forced_band >= 0 ? forced_band : (w_bytes <= 24.0 * (1 << 20) ? 1 : 0);
    gemm_tc2_kernel<X3><<<2 * clusters, NUM_THREADS2, SMEM2, st>>>(tmA, tmW, tmAlo, tmWlo, g.M, g.N, g.K, g.epi, band);
    CUDA_CHECK(cudaGetLastError());
}

void gemm_tcgen05_pair(const GemmArgs& g, const void* A_hi, const void* A_lo, cudaStream_t st, int num_sms) {
    if (g.w_type == DT_BF16X2) launch_pair<true>(g, A_hi, A_lo, st, num_sms);
    else launch_pair<false>(g, A_hi, A_lo, st, num_sms);
}

}  // namespace wlk
