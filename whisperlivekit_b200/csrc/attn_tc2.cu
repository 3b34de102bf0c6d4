// Encoder self-attention, two query tiles per CTA (sm_100a) -- the successor of attn_tc.cu's MODE_ENC for the bf16 serving mode.
//   O = softmax(Q K^T) V over all 1500 positions, no mask (reference whisper/model.py:148-173).
//
// Why: tools/attn_trace.py on attn_tc.cu (two independent CTAs per SM, one 128-query tile each) showed a serial chain per CTA
// -- S = Q K^T, exponentials, P V -- with blocking MMA issue (a group of MMAs that finds the pipe idle costs ~500-700 clk to
// issue and ~180 more to retire, tools/ubench/mma_rate.cu) and the two co-resident CTAs falling into phase with each other, so
// the MUFU pipe sat idle while both were in their MMA phases (52 % busy).  Here ONE CTA per SM owns two 128-query tiles (A, B)
// of the same head and all 512 TMEM columns; one MMA warp serves both tiles in a fixed order, which puts them in anti-phase by
// construction: while tile A's warps run their exponentials, the tensor pipe works for tile B, and vice versa.  K and V tiles
// are loaded once for both query tiles (half the TMA traffic) through a 3-stage ring.
//   warp 0       TMA producer: Q_A, Q_B once; 128-key K and V tiles (128B-swizzled) out of the fused [rows, 3d] qkv buffer
//   warp 1       MMA issuer, per key tile j:  S_A(j+1), P_A(j) V(j)  then  S_B(j+1), P_B(j) V(j)
//                (S of the next tile goes first: the exponentials only need S; P is single-buffered per query tile, so a tile's
//                softmax warps wait for their own P V before the first P store of the next key tile)
//   warps 2-9    softmax of tile A, warps 10-17 softmax of tile B: two threads per query row (64 keys each), one pass per key
//                tile against a per-row reference maximum that only moves when exceeded by 2^8 (attn_tc.cu explains)
// TMEM: per query tile S [0,128) fp32, O [128,192) fp32, P [192,256) bf16 pairs; tile B at +256.
#include <cudaTypedefs.h>

#include <type_traits>

#include "kernels.cuh"
#include "ptx.cuh"

namespace wlk {

bool make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols, std::string* err);

namespace {

constexpr int BQ = 128, BKV = 128, DH = 64, NTILE = 2, NST = 3;
constexpr int ATT2_THREADS = 64 + NTILE * 8 * 32;                 // 576
constexpr uint32_t TILE_BYTES = BQ * DH * 2;                       // 16 KB
constexpr uint32_t SM_Q = 0;                                       // [2] query tiles
constexpr uint32_t SM_K = NTILE * TILE_BYTES;                      // [NST]
constexpr uint32_t SM_V = SM_K + NST * TILE_BYTES;                 // [NST]
constexpr uint32_t SM_BAR = SM_V + NST * TILE_BYTES;
constexpr uint32_t SM_XCH = SM_BAR + 256;                          // [2 tiles][2 parities][2 halves][128] floats
constexpr uint32_t SMEM2 = SM_XCH + NTILE * 2 * 2 * BQ * 4 + 1024;
constexpr uint32_t TM_S = 0, TM_O = 128, TM_P = 192, TM_TILE = 256;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void named_bar_64(int id) {
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tm, int n_head, int d_model, bf16* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - ptx::smem_u32(smem_raw));
    // barriers (8 bytes each)
    const uint32_t bar_q = sbase + SM_BAR;
    const uint32_t bar_kv_full = bar_q + 8;                  // [NST]
    const uint32_t bar_kv_empty = bar_kv_full + 8 * NST;     // [NST]
    const uint32_t bar_tile = bar_kv_empty + 8 * NST;        // per query tile: s_full, s_free, p_full, o_full (32 bytes)
    const uint32_t tmem_slot = bar_tile + 32 * NTILE;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + (tmem_slot - sbase));
    auto bar_s_full = [&](int t) { return bar_tile + 32 * t; };
    auto bar_s_free = [&](int t) { return bar_tile + 32 * t + 8; };
    auto bar_p_full = [&](int t) { return bar_tile + 32 * t + 16; };
    auto bar_o_full = [&](int t) { return bar_tile + 32 * t + 24; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (NTILE * BQ), h = blockIdx.y, b = blockIdx.z;
    constexpr int NT = (N_CTX + BKV - 1) / BKV;              // 12 key tiles
    const int q_row = b * N_CTX + q0, q_col = h * DH;
    const int k_row = b * N_CTX, k_col = d_model + h * DH, v_col = 2 * d_model + h * DH;
    // a second query tile that starts past the sequence end does not exist (the last CTA of a stream: rows 1280..1499)
    const int n_tiles = (q0 + BQ < N_CTX) ? 2 : 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm);
        ptx::mbar_init(bar_q, 1);
        for (int i = 0; i < NST; ++i) { ptx::mbar_init(bar_kv_full + 8 * i, 1); ptx::mbar_init(bar_kv_empty + 8 * i, 1); }
        for (int t = 0; t < NTILE; ++t) {
            ptx::mbar_init(bar_s_full(t), 1);
            ptx::mbar_init(bar_s_free(t), 256);
            ptx::mbar_init(bar_p_full(t), 256);
            ptx::mbar_init(bar_o_full(t), 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(bar_q, n_tiles * TILE_BYTES);
            for (int t = 0; t < n_tiles; ++t)
                ptx::tma_load_2d(sbase + SM_Q + t * TILE_BYTES, &tm, bar_q, q_col, q_row + t * BQ);
            for (int j = 0; j < NT; ++j) {
                const uint32_t s = j % NST, ph = (j / NST) & 1;
                ptx::mbar_wait(bar_kv_empty + 8 * s, ph ^ 1);
                ptx::mbar_arrive_expect_tx(bar_kv_full + 8 * s, 2 * TILE_BYTES);
                ptx::tma_load_2d(sbase + SM_K + s * TILE_BYTES, &tm, bar_kv_full + 8 * s, k_col, k_row + j * BKV);
                ptx::tma_load_2d(sbase + SM_V + s * TILE_BYTES, &tm, bar_kv_full + 8 * s, v_col, k_row + j * BKV);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BKV, 0, 0);   // A = Q K-major, B = K K-major
        constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, DH, 0, 1);    // A = P (TMEM), B = V MN-major
        auto issue_s = [&](int t, int j) {
            if (lane == 0) {
                const uint64_t dq = ptx::umma_desc_kmajor_sw128(sbase + SM_Q + t * TILE_BYTES);
                const uint64_t dk = ptx::umma_desc_kmajor_sw128(sbase + SM_K + (j % NST) * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    ptx::umma_bf16_ss(tmem + t * TM_TILE + TM_S, dq + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                ptx::umma_commit(bar_s_full(t));
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t, int j, bool release_kv) {
            if (lane == 0) {
                const uint64_t dv = ptx::umma_desc_mnmajor_sw128(sbase + SM_V + (j % NST) * TILE_BYTES, BKV * 128);
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)
                    umma_bf16_ts(tmem + t * TM_TILE + TM_O, tmem + t * TM_TILE + TM_P + 8 * k, dv + 128 * k, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                ptx::umma_commit(bar_o_full(t));
                if (release_kv) ptx::umma_commit(bar_kv_empty + 8 * (j % NST));
            }
            __syncwarp();
        };
        ptx::mbar_wait(bar_q, 0);
        ptx::mbar_wait(bar_kv_full, 0);
        ptx::tc_fence_after();
        for (int t = 0; t < n_tiles; ++t) issue_s(t, 0);
        for (int j = 0; j < NT; ++j) {
            for (int t = 0; t < n_tiles; ++t) {
                if (j + 1 < NT) {
                    if (t == 0) ptx::mbar_wait(bar_kv_full + 8 * ((j + 1) % NST), ((j + 1) / NST) & 1);
                    ptx::mbar_wait(bar_s_free(t), j & 1);     // this tile's softmax warps have read S(j)
                    ptx::tc_fence_after();
                    issue_s(t, j + 1);
                }
                ptx::mbar_wait(bar_p_full(t), j & 1);         // P(j) of this tile is in TMEM
                ptx::tc_fence_after();
                issue_pv(t, j, t == n_tiles - 1);             // the K/V stage is free once the last tile's P V has retired
            }
        }
    } else {
        const int t = (warp - 2) >> 3;                        // query tile of this warp
        if (t < n_tiles) {
            const int w = (warp - 2) & 7;                     // softmax warp within the tile
            const int qd = warp & 3;                          // TMEM lane quadrant = warp id % 4 (hardware rule)
            const int half = w >> 2;
            const int r = qd * 32 + lane;                     // query row within the tile == TMEM lane
            const uint32_t tbase = tmem + t * TM_TILE + (static_cast<uint32_t>(qd * 32) << 16);
            const uint32_t s_addr = tbase + TM_S + half * (BKV / 2);
            const uint32_t p_addr = tbase + TM_P + half * (BKV / 4);
            const uint32_t o_addr = tbase + TM_O + half * (DH / 2);
            float* xch = reinterpret_cast<float*>(sgen + SM_XCH) + t * 2 * 2 * BQ;
            const int bar_id = 1 + t * 4 + qd;                // the two warps of a row quadrant
            uint32_t xn = 0;
            auto exchange = [&](float v) -> float {
                float* slot = xch + (xn & 1) * 2 * BQ;
                ++xn;
                slot[half * BQ + r] = v;
                named_bar_64(bar_id);
                return slot[(half ^ 1) * BQ + r];
            };
            float m = -INFINITY, l = 0.f;
            auto tile = [&](int j, auto masked_tag) {
                constexpr bool MASKED = decltype(masked_tag)::value;
                const int n_valid = N_CTX - j * BKV - half * (BKV / 2);
                ptx::mbar_wait(bar_s_full(t), j & 1);
                ptx::tc_fence_after();
                uint32_t va[16], vb[16];
                if (j == 0) {                                 // first tile: a true row maximum seeds the reference
                    float mx = -INFINITY;
#pragma unroll 1
                    for (int c = 0; c < 4; c += 2) {
                        ptx::tmem_ld_32x16(s_addr + c * 16, va);
                        ptx::tmem_ld_32x16(s_addr + c * 16 + 16, vb);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (!MASKED || c * 16 + i < n_valid) mx = fmaxf(mx, __uint_as_float(va[i]));
                            if (!MASKED || c * 16 + 16 + i < n_valid) mx = fmaxf(mx, __uint_as_float(vb[i]));
                        }
                    }
                    m = fmaxf(mx, exchange(mx)) * LOG2E;
                }
                bool pv_done = j == 0;                        // P(j-1) V(j-1) of this tile retired: P and O may be touched
#pragma unroll 1
                for (;;) {
                    float rs = 0.f, mx = -INFINITY;
                    auto emit = [&](const uint32_t* v, int c) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
                            float p0 = fast_exp2(fmaf(s0, LOG2E, -m));
                            float p1 = fast_exp2(fmaf(s1, LOG2E, -m));
                            if (MASKED) {
                                if (c * 16 + 2 * i >= n_valid) p0 = 0.f; else mx = fmaxf(mx, s0);
                                if (c * 16 + 2 * i + 1 >= n_valid) p1 = 0.f; else mx = fmaxf(mx, s1);
                            } else {
                                mx = fmaxf(mx, fmaxf(s0, s1));
                            }
                            rs += p0 + p1;
                            __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
                            pk[i] = *reinterpret_cast<uint32_t*>(&hb);
                        }
                        if (!pv_done) {                       // first store of this key tile: the previous P V must have read P
                            ptx::mbar_wait(bar_o_full(t), (j - 1) & 1);
                            ptx::tc_fence_after();
                            pv_done = true;
                        }
                        tmem_st_32x8(p_addr + c * 8, pk);
                    };
                    ptx::tmem_ld_32x16(s_addr, va);
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x16(s_addr + 16, vb);
                    emit(va, 0);
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x16(s_addr + 32, va);
                    emit(vb, 1);
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x16(s_addr + 48, vb);
                    emit(va, 2);
                    ptx::tmem_ld_wait();
                    emit(vb, 3);
                    const float mx2 = fmaxf(mx, exchange(mx)) * LOG2E;
                    const bool need = mx2 > m + 8.0f;
                    if (!__any_sync(0xffffffffu, need)) { l += rs; break; }
                    // rare: move the reference, rescale this half of O and l, then redo the tile's probabilities
                    const float alpha = need ? fast_exp2(m - mx2) : 1.0f;
                    if (j > 0) {
                        uint32_t o[32];
                        ptx::tmem_ld_32x32(o_addr, o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32(o_addr, o);
                    }
                    l *= alpha;
                    if (need) m = mx2;
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(bar_s_free(t));              // S fully consumed: the next Q K^T may overwrite it
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(bar_p_full(t));              // P written, O rescaled if needed: P V may run
            };
#pragma unroll 1
            for (int j = 0; j < NT - 1; ++j) tile(j, std::false_type{});
            tile(NT - 1, std::true_type{});
            ptx::mbar_wait(bar_o_full(t), (NT - 1) & 1);
            ptx::tc_fence_after();
            const float inv = 1.0f / (l + exchange(l));
            uint32_t v[32];
            ptx::tmem_ld_32x32(o_addr, v);
            ptx::tmem_ld_wait();
            const int row = q0 + t * BQ + r;
            if (row < N_CTX) {
                bf16* o = out + (int64_t)(b * N_CTX + row) * d_model + h * DH + half * (DH / 2);
#pragma unroll
                for (int e8 = 0; e8 < 4; ++e8) {
                    uint4 u;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[e8 * 8 + 0]) * inv, __uint_as_float(v[e8 * 8 + 1]) * inv);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[e8 * 8 + 2]) * inv, __uint_as_float(v[e8 * 8 + 3]) * inv);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[e8 * 8 + 4]) * inv, __uint_as_float(v[e8 * 8 + 5]) * inv);
                    __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(v[e8 * 8 + 6]) * inv, __uint_as_float(v[e8 * 8 + 7]) * inv);
                    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
                    reinterpret_cast<uint4*>(o)[e8] = u;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace

void enc_attention_tcgen05_two_tile(const void* qkv, int batch, int n_head, int d_model, void* out, cudaStream_t st) {
    CUtensorMap tm;
    std::string err;
    WLK_CHECK(make_tmap_bf16_2d(&tm, qkv, (uint64_t)batch * N_CTX, (uint64_t)3 * d_model, (uint64_t)3 * d_model, BQ, DH, &err),
              "qkv tensor map: %s", err.c_str());
    static bool seen[64] = {};
    if (first_on_device(seen))
        CUDA_CHECK(cudaFuncSetAttribute(attn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2));
    dim3 grid((N_CTX + NTILE * BQ - 1) / (NTILE * BQ), n_head, batch);
    attn_tc2_kernel<<<grid, ATT2_THREADS, SMEM2, st>>>(tm, n_head, d_model, reinterpret_cast<bf16*>(out));
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace wlk
