// C[M][N] (bf16) = A[M][K] (bf16, row-major) * B[N][K]^T (bf16, row-major = nn.Linear weight layout) + bias[N]
// Legacy tensor path (mma.sync m16n8k16, fp32 accumulate), 3-stage cp.async pipeline, K step 32.
// Used for the proxy-head MLPs (network_exp_msg_chn_adapt.py:1089-1098: Linear 32->512, 512->512)
// and their data gradients (B = pre-transposed weight).  M is arbitrary (rows are predicated), N must
// be a multiple of the tile width, K a multiple of 32.
#pragma once
#include "common.cuh"
#include "conv_mma.cuh"   // swz_off

namespace ptta {

struct GemmParams {
    const bf16* A; const bf16* B; bf16* C; const float* bias;
    long long M; int N, K;
};

template <int WARPS_M, int WARPS_N, int MT>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32) gemm_mma_kernel(const GemmParams p) {
    PDL_SYNC();
    const int BM = WARPS_M * MT * 16, BN = WARPS_N * 32, BK = 32, STAGES = 3;
    const int NT = WARPS_M * WARPS_N * 32;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;                              // [STAGES][BM][32] bf16
    unsigned char* sB = smem + STAGES * BM * 64;           // [STAGES][BN][32] bf16
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int KT = p.K / BK;

    auto load_stage = [&](int stage, int kt) {
        const uint32_t a_base = smem_u32(sA + stage * BM * 64), b_base = smem_u32(sB + stage * BN * 64);
        for (int i = tid; i < BM * 4; i += NT) {
            int row = i >> 2, c = i & 3;
            long long gm = m0 + row;
            bool ok = gm < p.M;
            const bf16* src = p.A + (size_t)(ok ? gm : 0) * p.K + kt * BK + c * 8;
            cp_async16(a_base + swz_off<32>(row, c), src, ok);
        }
        for (int i = tid; i < BN * 4; i += NT) {
            int row = i >> 2, c = i & 3;
            const bf16* src = p.B + (size_t)(n0 + row) * p.K + kt * BK + c * 8;
            cp_async16(b_base + swz_off<32>(row, c), src, true);
        }
    };

    float acc[MT][4][4];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kc = lane >> 4;
    const int b_row = (lane & 7) + (lane >> 4) * 8, b_kc = (lane >> 3) & 1;

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // prefetch tile kt + STAGES - 1 into the slot freed at the previous iteration
            int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const int stage = kt % STAGES;
        const uint32_t a_base = smem_u32(sA + stage * BM * 64), b_base = smem_u32(sB + stage * BN * 64);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t b[8];
            ldmatrix_x4(b[0], b[1], b[2], b[3], b_base + swz_off<32>(wn * 32 + b_row, ks * 2 + b_kc));
            ldmatrix_x4(b[4], b[5], b[6], b[7], b_base + swz_off<32>(wn * 32 + 16 + b_row, ks * 2 + b_kc));
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                uint32_t a[4];
                ldmatrix_x4(a[0], a[1], a[2], a[3], a_base + swz_off<32>((wm * MT + mt) * 16 + a_row, ks * 2 + a_kc));
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], a, b[nt * 2], b[nt * 2 + 1]);
            }
        }
    }
    cp_async_wait<0>();

    const int c_row = lane >> 2, c_col = (lane & 3) * 2;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            long long gm = m0 + (wm * MT + mt) * 16 + c_row + half * 8;
            if (gm >= p.M) continue;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                int gn = n0 + wn * 32 + nt * 8 + c_col;
                float v0 = acc[mt][nt][half * 2], v1 = acc[mt][nt][half * 2 + 1];
                if (p.bias) { v0 += p.bias[gn]; v1 += p.bias[gn + 1]; }
                *reinterpret_cast<uint32_t*>(p.C + (size_t)gm * p.N + gn) = pack_bf162(v0, v1);
            }
        }
}

template <int WARPS_M, int WARPS_N, int MT>
int launch_gemm_t(const GemmParams& p, cudaStream_t st) {
    const int BM = WARPS_M * MT * 16, BN = WARPS_N * 32;
    const int smem = 3 * (BM + BN) * 64;
    static bool attr_set = false;
    if (!attr_set) {
        PTTA_CUDA(cudaFuncSetAttribute(gemm_mma_kernel<WARPS_M, WARPS_N, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    dim3 grid(cdiv(p.M, BM), p.N / BN);
    launch_k(gemm_mma_kernel<WARPS_M, WARPS_N, MT>, grid, WARPS_M * WARPS_N * 32, smem, st, p);
    return check_launch("gemm_mma");
}

inline int launch_gemm(const GemmParams& p, cudaStream_t st) {
    PTTA_CHECK(p.K % 32 == 0 && p.K >= 32, "gemm: K=%d must be a positive multiple of 32", p.K);
    PTTA_CHECK(p.N % 32 == 0, "gemm: N=%d must be a multiple of 32", p.N);
    if (p.M <= 0) return 0;
    if (p.N % 128 == 0) return launch_gemm_t<2, 4, 4>(p, st);
    return launch_gemm_t<8, 1, 1>(p, st);
}

}  // namespace ptta
