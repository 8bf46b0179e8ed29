// C[M][N] (bf16) = A[M][K] (bf16, row-major) * B[N][K]^T (bf16, nn.Linear weight layout) + bias[N] on tcgen05:
// TMA (SWIZZLE_128B boxes of 64 bf16 = 128 B rows) -> 4-stage shared-memory ring -> tcgen05.mma M128 x N256 x K16 with the
// accumulator in TMEM (2 x 256 columns, double buffered) -> four epilogue warps (tcgen05.ld, bias, bf16 into a SWIZZLE_128B staging
// tile of 32 rows x 64 columns per warp, double buffered) -> TMA stores (whole 128 B lines; the tensor map clips the last row block).
// Persistent CTAs, tiles ordered so that the two N-halves of one row block run back to back (A stays in L2).
// Used for the Linear layers of the proxy heads (network_exp_msg_chn_adapt.py:1089-1098: 32 -> 512, 512 -> 512) and their data
// gradient; N must be a multiple of 256 and K a multiple of 64 -- or K = 32: the 64-wide boxes then read past the 32-column rows and TMA
// zero-fills the upper half of every operand row, so the same kernel runs one K block whose second half contributes nothing.
#pragma once
#include "conv_tc.cuh"

namespace ptta {

struct GemmTcParams {
    bf16* C; const float* bias;
    long long M; int N, K;
    int m_tiles, n_tiles;
};

struct GemmTcCfg {
    static const int BM = 128, BN = 256, BK = 64, STAGES = 4;
    static const int A_BYTES = BM * BK * 2;        // 16 KB
    static const int B_BYTES = BN * BK * 2;        // 32 KB
    static const int STAGE_BYTES = A_BYTES + B_BYTES;
    static const int OUT_TILE = 32 * 128;          // one epilogue warp's staging tile: 32 rows x 64 columns of bf16
    static const int SMEM = 1024 + STAGES * STAGE_BYTES + 4 * 2 * OUT_TILE + 256;
    static const int THREADS = 192;               // warp 0 TMA | warp 1 MMA + TMEM alloc | warps 2-5 epilogue
};

__global__ void __launch_bounds__(GemmTcCfg::THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                       const __grid_constant__ CUtensorMap tmap_b,
                                                                       const __grid_constant__ CUtensorMap tmap_c, const GemmTcParams p) {
    typedef GemmTcCfg C;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t out_s = smem_base + C::STAGES * C::STAGE_BYTES;
    const uint32_t bar_s = out_s + 4 * 2 * C::OUT_TILE;
    const uint32_t full = bar_s, empty = bar_s + 8 * C::STAGES, acc_full = bar_s + 16 * C::STAGES, acc_empty = acc_full + 16;
    const uint32_t tmem_slot = acc_empty + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int num_tiles = p.m_tiles * p.n_tiles, KB = (p.K + C::BK - 1) / C::BK;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_a);
        tc::prefetch_tmap(&tmap_b);
        tc::prefetch_tmap(&tmap_c);
        for (int i = 0; i < C::STAGES; ++i) { tc::mbar_init(full + 8 * i, 1); tc::mbar_init(empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(acc_full + 8 * i, 1); tc::mbar_init(acc_empty + 8 * i, 128); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    PDL_SYNC();      // the set-up above (barriers, TMEM, descriptor prefetch) overlaps the previous kernel; nothing before this line touches global data

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % C::STAGES, par = (it / C::STAGES) & 1;
                    tc::mbar_wait(empty + 8 * s, par ^ 1);
                    tc::mbar_arrive_expect_tx(full + 8 * s, C::STAGE_BYTES);
                    const uint32_t a_dst = smem_base + s * C::STAGE_BYTES;
                    tc::tma_load_2d(a_dst, &tmap_a, full + 8 * s, kb * C::BK, mt * C::BM);
                    tc::tma_load_2d(a_dst + C::A_BYTES, &tmap_b, full + 8 * s, kb * C::BK, nt * C::BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(128, 256);
            const uint64_t d0 = make_desc_sw128(0);
            const uint32_t hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
            uint32_t it = 0, t = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                const uint32_t as = t & 1;
                tc::mbar_wait(acc_empty + 8 * as, ((t >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % C::STAGES;
                    tc::mbar_wait(full + 8 * s, (it / C::STAGES) & 1);
                    tc::tc_fence_after();
                    const uint32_t a_lo = lo0 + ((smem_base + s * C::STAGE_BYTES) >> 4);
                    const uint32_t b_lo = a_lo + (C::A_BYTES >> 4);
                    if (kb == 0) tc::umma_f16_split<false>(d_tmem, a_lo, hi, b_lo, hi, idesc);
                    else tc::umma_f16_split<true>(d_tmem, a_lo, hi, b_lo, hi, idesc);
#pragma unroll
                    for (int k = 1; k < 4; ++k) tc::umma_f16_split<true>(d_tmem, a_lo + k * 2, hi, b_lo + k * 2, hi, idesc);
                    tc::umma_commit(empty + 8 * s);          // stage may be refilled once these MMAs have read it
                }
                tc::umma_commit(acc_full + 8 * as);
            }
        }
    } else {
        const int q = warp & 3;
        unsigned char* stage = smem + (out_s - smem_base) + q * 2 * C::OUT_TILE;
        const uint32_t stage_s = out_s + q * 2 * C::OUT_TILE;
        uint32_t t = 0, nst = 0;             // nst: stores issued by this warp (staging buffer = nst & 1)
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
            const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
            const uint32_t as = t & 1;
            tc::mbar_wait(acc_full + 8 * as, (t >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll 1
            for (int cp = 0; cp < 4; ++cp, ++nst) {                     // 64 output columns per round: one 128 B staging row per lane
                const int col0 = nt * C::BN + cp * 64;
                unsigned char* srow = stage + (nst & 1) * C::OUT_TILE + lane * 128;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");      // the store that used this buffer two rounds ago has read it
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * 256 + cp * 64 + h * 32, v);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]) + (p.bias ? __ldg(p.bias + col0 + h * 32 + g * 8 + j) : 0.f);
                        uint4 ov;
                        ov.x = pack_bf162(f[0], f[1]); ov.y = pack_bf162(f[2], f[3]);
                        ov.z = pack_bf162(f[4], f[5]); ov.w = pack_bf162(f[6], f[7]);
                        *reinterpret_cast<uint4*>(srow + (((h * 4 + g) ^ (lane & 7)) << 4)) = ov;
                    }
                }
                tc::fence_proxy_async();                                // generic-proxy writes -> visible to the TMA store
                __syncwarp();
                if (lane == 0 && (long long)mt * C::BM + q * 32 < p.M) {     // (a quarter that lies entirely below the matrix stores nothing)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&tmap_c), "r"(stage_s + (nst & 1) * C::OUT_TILE), "r"(col0), "r"(mt * C::BM + q * 32) : "memory");
                    tc::bulk_store_commit();
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(acc_empty + 8 * as);
        }
        if (lane == 0) tc::bulk_store_wait_read_all();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// row-major [rows][cols] bf16 matrix, box {64, box_rows}, SWIZZLE_128B (cols < 64: the box reaches past the row, TMA zero-fills)
inline int make_tmap_2d(CUtensorMap* map, const void* ptr, long long rows, int cols, int box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    PTTA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTTA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%d)", (int)r, rows, cols);
    return 0;
}

inline bool gemm_tc_supported(long long M, int N, int K) { return M > 0 && N % 256 == 0 && ((K % 64 == 0 && K >= 64) || K == 32); }

inline int launch_gemm_tc(const bf16* A, const bf16* B, bf16* Cout, const float* bias, long long M, int N, int K, cudaStream_t st) {
    typedef GemmTcCfg C;
    PTTA_CHECK(gemm_tc_supported(M, N, K), "gemm_tc: unsupported shape M=%lld N=%d K=%d", M, N, K);
    static int sms = 0;
    if (!sms) {
        PTTA_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int dev = 0;
        PTTA_CUDA(cudaGetDevice(&dev));
        PTTA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    CUtensorMap ta, tb, tcm;
    PTTA_TRY(make_tmap_2d(&ta, A, M, K, C::BM));
    PTTA_TRY(make_tmap_2d(&tb, B, N, K, C::BN));
    PTTA_TRY(make_tmap_2d(&tcm, Cout, M, N, 32));
    GemmTcParams p; p.C = Cout; p.bias = bias; p.M = M; p.N = N; p.K = K;
    p.m_tiles = cdiv(M, C::BM); p.n_tiles = N / C::BN;
    int tiles = p.m_tiles * p.n_tiles;
    int grid = tiles < sms ? tiles : sms;
    launch_k(gemm_tc_kernel, grid, C::THREADS, C::SMEM, st, ta, tb, tcm, p);
    return check_launch("gemm_tc");
}

}  // namespace ptta
