// C[M][N] (fp32) = A[rows][M]^T (bf16) * B[rows][N] (bf16): the weight gradient of a Linear layer, dW[out][in] = sum_r dY[r][out] X[r][in],
// on tcgen05 -- the reduction runs over the ROWS of both operands, i.e. both are "MN-major" in UMMA terms (the M / N index is the contiguous
// one in memory), so no transposed copy of either matrix is ever made:
//   TMA boxes of 64 columns x 64 rows, SWIZZLE_128B (one box = eight 8-row x 128 B swizzle atoms stacked along K)
//   -> 4-stage shared-memory ring (A: 2 boxes = M 128, B: 4 boxes = N 256 per stage)
//   -> tcgen05.mma M128 x N256 x K16, a_major = b_major = MN, descriptors with LBO = 8 KB (next 64-column box) and SBO = 1 KB (next 8 rows);
//      one K step = 16 rows = 2 KB further into every box
//   -> fp32 accumulator in TMEM (256 columns) -> four epilogue warps store their 32 rows x 256 columns of the split's partial tile.
// The row range is split over the grid (split-K: (M/128)(N/256) tiles x S row ranges ~ one CTA per SM); gemm_tn_reduce_kernel adds the S
// partial tiles in a fixed order (deterministic).  Rows past the end of the matrices are zero-filled by TMA and contribute nothing.
// Used by the stage-2 head trainer (src/head_main.py:464-480: wgrad of pred.0 / pred.3, 512 x 512 over R = N H/4 W/4 rows).
#pragma once
#include "gemm_tc.cuh"

namespace ptta {

struct GemmTnParams {
    float* part;                  // [splits][M][N]
    long long rows; int M, N;
    int m_tiles, n_tiles, splits, kb_per_split, kb_total;
};

struct GemmTnCfg {
    static const int BM = 128, BN = 256, BK = 64, STAGES = 4;
    static const int BOX_BYTES = 64 * BK * 2;      // 8 KB: 64 rows (K) x 64 columns (M or N) of bf16
    static const int A_BYTES = 2 * BOX_BYTES, B_BYTES = 4 * BOX_BYTES;
    static const int STAGE_BYTES = A_BYTES + B_BYTES;     // 48 KB
    static const int SMEM = 1024 + STAGES * STAGE_BYTES + 256;
    static const int THREADS = 192;                // warp 0 TMA | warp 1 MMA + TMEM alloc | warps 2-5 epilogue
};

namespace tc {
// MN-major SWIZZLE_128B operand descriptor: 64-element (128 B) runs along M/N, 8 K-rows per swizzle atom (1024 B);
// lbo = bytes between consecutive 64-element blocks along M/N, sbo = bytes between consecutive groups of 8 K-rows
__device__ __forceinline__ uint64_t make_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
}  // namespace tc

__global__ void __launch_bounds__(GemmTnCfg::THREADS, 1) gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                          const __grid_constant__ CUtensorMap tmap_b, const GemmTnParams p) {
    typedef GemmTnCfg C;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_s = smem_base + C::STAGES * C::STAGE_BYTES;
    const uint32_t full = bar_s, empty = bar_s + 8 * C::STAGES, acc_full = bar_s + 16 * C::STAGES;
    const uint32_t tmem_slot = acc_full + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int tile = blockIdx.x % (p.m_tiles * p.n_tiles), split = blockIdx.x / (p.m_tiles * p.n_tiles);
    const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
    const int KB = kb1 - kb0;                      // >= 1 by construction of `splits`

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_a);
        tc::prefetch_tmap(&tmap_b);
        for (int i = 0; i < C::STAGES; ++i) { tc::mbar_init(full + 8 * i, 1); tc::mbar_init(empty + 8 * i, 1); }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 256);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    PDL_SYNC();

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < KB; ++it) {
                const uint32_t s = it % C::STAGES, par = (it / C::STAGES) & 1;
                tc::mbar_wait(empty + 8 * s, par ^ 1);
                tc::mbar_arrive_expect_tx(full + 8 * s, C::STAGE_BYTES);
                const uint32_t dst = smem_base + s * C::STAGE_BYTES;
                const int r0 = (kb0 + it) * C::BK;
#pragma unroll
                for (int j = 0; j < 2; ++j) tc::tma_load_2d(dst + j * C::BOX_BYTES, &tmap_a, full + 8 * s, mt * C::BM + j * 64, r0);
#pragma unroll
                for (int j = 0; j < 4; ++j) tc::tma_load_2d(dst + C::A_BYTES + j * C::BOX_BYTES, &tmap_b, full + 8 * s, nt * C::BN + j * 64, r0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(128, 256) | (1u << 15) | (1u << 16);      // A and B MN-major
            const uint64_t d0 = tc::make_desc_sw128_mn(0, C::BOX_BYTES, 1024);
            const uint32_t hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
            for (int it = 0; it < KB; ++it) {
                const uint32_t s = it % C::STAGES;
                tc::mbar_wait(full + 8 * s, (it / C::STAGES) & 1);
                tc::tc_fence_after();
                const uint32_t a_lo = lo0 + ((smem_base + s * C::STAGE_BYTES) >> 4);
                const uint32_t b_lo = a_lo + (C::A_BYTES >> 4);
                if (it == 0) tc::umma_f16_split<false>(tmem_base, a_lo, hi, b_lo, hi, idesc);
                else tc::umma_f16_split<true>(tmem_base, a_lo, hi, b_lo, hi, idesc);
#pragma unroll
                for (int k = 1; k < 4; ++k) tc::umma_f16_split<true>(tmem_base, a_lo + k * (2048 >> 4), hi, b_lo + k * (2048 >> 4), hi, idesc);
                tc::umma_commit(empty + 8 * s);
            }
            tc::umma_commit(acc_full);
        }
    } else {
        const int q = warp & 3;
        tc::mbar_wait(acc_full, 0);
        tc::tc_fence_after();
        const int row = mt * C::BM + q * 32 + lane;
        float* dst = p.part + ((size_t)split * p.M + row) * p.N + nt * C::BN;
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {
            uint32_t v[32];
            tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cb * 32, v);
#pragma unroll
            for (int g = 0; g < 8; ++g)
                *reinterpret_cast<float4*>(dst + cb * 32 + g * 4) = make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]),
                                                                                __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 256);
    }
}

// out[i] = sum_s part[s][i]  (fixed order)
__global__ void __launch_bounds__(256) gemm_tn_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int splits, long long n4) {
    PDL_SYNC();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = __ldg(reinterpret_cast<const float4*>(part) + i);
    for (int s = 1; s < splits; ++s) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(part) + (size_t)s * n4 + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
}

inline bool gemm_tn_supported(long long rows, int M, int N) { return rows > 0 && M % 128 == 0 && N % 256 == 0; }

struct GemmTnPlan { int m_tiles, n_tiles, kb_total, kb_per_split, splits; };
inline GemmTnPlan gemm_tn_plan(long long rows, int M, int N, int sms) {
    GemmTnPlan g;
    g.m_tiles = M / GemmTnCfg::BM; g.n_tiles = N / GemmTnCfg::BN;
    g.kb_total = (int)cdiv(rows, GemmTnCfg::BK);
    int want = std::max(1, sms / (g.m_tiles * g.n_tiles));
    if (want > g.kb_total) want = g.kb_total;
    g.kb_per_split = cdiv(g.kb_total, want);
    g.splits = cdiv(g.kb_total, g.kb_per_split);
    return g;
}
inline int gemm_tn_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    }
    return sms;
}
// bytes of the split-K partial buffer; planned for 148 SMs or the device's count, whichever needs more
inline size_t gemm_tn_workspace_bytes(long long rows, int M, int N) {
    GemmTnPlan g = gemm_tn_plan(rows, M, N, 160);
    return (size_t)g.splits * M * N * sizeof(float);
}

inline int launch_gemm_tn_tc(const bf16* A, const bf16* B, float* out, float* workspace, long long rows, int M, int N, cudaStream_t st) {
    typedef GemmTnCfg C;
    PTTA_CHECK(gemm_tn_supported(rows, M, N), "gemm_tn_tc: unsupported shape rows=%lld M=%d N=%d", rows, M, N);
    static bool attr = false;
    if (!attr) {
        PTTA_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr = true;
    }
    const GemmTnPlan g = gemm_tn_plan(rows, M, N, std::min(gemm_tn_sms(), 160));
    CUtensorMap ta, tb;
    PTTA_TRY(make_tmap_2d(&ta, A, rows, M, C::BK));
    PTTA_TRY(make_tmap_2d(&tb, B, rows, N, C::BK));
    GemmTnParams p; p.part = workspace; p.rows = rows; p.M = M; p.N = N;
    p.m_tiles = g.m_tiles; p.n_tiles = g.n_tiles; p.splits = g.splits; p.kb_per_split = g.kb_per_split; p.kb_total = g.kb_total;
    launch_k(gemm_tn_tc_kernel, g.m_tiles * g.n_tiles * g.splits, C::THREADS, C::SMEM, st, ta, tb, p);
    PTTA_TRY(check_launch("gemm_tn_tc"));
    const long long n4 = (long long)M * N / 4;
    launch_k(gemm_tn_reduce_kernel, cdiv(n4, 256), 256, 0, st, (const float*)workspace, out, g.splits, n4);
    return check_launch("gemm_tn_reduce");
}

}  // namespace ptta
