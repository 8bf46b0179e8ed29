// {1,2,3} -> 32 channel 3x3 s1 p1 convolution ("stem": init.0 of every MSG-CHN encoder, network_exp_msg_chn_adapt.py:166-186; and,
// with flipped weights + a ReLU-derivative mask, the data gradient of the 32 -> 1 prediction layers prdct.3, :289) on tcgen05.
//
// The layer reads 4..12 B and writes 64 B per pixel: it is bound by the write of the 32-channel map.  The CUDA-core form spends
// 288..864 FMAs per pixel (26..40 us at 352x1216, FMA-pipe bound); here the contraction (K = 9 CIN <= 27, padded to 32) runs on the
// tensor cores and the CUDA cores only build the operand and pack the result:
//
//   * tile = 128 consecutive pixels of the flattened [N*H*W] sequence; thread t of warps 0-3 owns pixel t: it gathers its 9 CIN
//     fp32 inputs (normalisation scale/shift folded in, zero outside the image), splits every value into a bf16 head and a bf16
//     remainder (v = hi + lo to 16 mantissa bits) and writes its 64 B rows of the two K-major SWIZZLE_64B operand tiles A_hi, A_lo;
//   * warp 4 issues six tcgen05.mma M128 x N32 x K16 (A_hi B_hi + A_lo B_hi + A_hi B_lo over two K steps; the weights are split the
//     same way at pack time), accumulators in 32 TMEM columns;
//   * the same thread t reads TMEM lane t (its pixel's 32 outputs), adds the bias, applies the mask / ReLU, packs to bf16 into a
//     its warp's staging rows, and the warp stores its 32 pixels (2 KB contiguous) with 512 B coalesced stores.
//
// Products carry 16 mantissa bits per factor and accumulate in fp32: the result equals the fp32 FMA kernel up to rare 1-ulp flips of
// the bf16 output.  Three CTAs are resident per SM (46 KB of shared memory, 64 TMEM columns each) and every gather warp runs a two-stage
// software pipeline, so the global-load latency and the MMA round trip are covered by the epilogue of the previous tile.
#pragma once
#include "conv_tc.cuh"

namespace ptta {

struct StemTcParams {
    const float* plane[3];
    long long batch_stride[3];
    float scale[3], shift[3];
    const bf16* w;       // weight image of pack_stem_weight_tc_kernel (4 KB: B_hi then B_lo)
    const float* bias;   // [32] or null
    const bf16* mask;    // relu mask (NHWC 32) or null
    bf16* out;
    int N, H, W;
    int relu_out;
    long long total;     // N*H*W
    int tiles;
};

struct StemTcCfg {
    static const int A_BYTES = 128 * 64;            // one operand tile: 128 pixels x 32 bf16
    static const int W_BYTES = 2 * 32 * 64;         // B_hi, B_lo
    static const int OUT_BYTES = 128 * 64;
    static const int SMEM = 1024 /*align slack*/ + 4 * A_BYTES + W_BYTES + OUT_BYTES + 128 /*bias*/ + 64;
    static const int THREADS = 160;                 // warps 0-3: gather + epilogue (TMEM lane quarter = warp) | warp 4: MMA issuer
};

// SWIZZLE_64B position of 16 B chunk c of 64 B row r (rows 64 B apart, pattern repeats every 512 B; base 512 B aligned)
__device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4); }

// fp32 [32][CIN][3][3] (stem weight) -> B_hi | B_lo, each [32 cout rows][32 k] bf16, k = ci*9 + ky*3 + kx (zero for k >= 9 CIN), SWIZZLE_64B
__global__ void pack_stem_weight_tc_kernel(const float* __restrict__ w, bf16* __restrict__ image, int cin) {
    PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * 32 * 4) return;
    const int part = i >> 7, row = (i >> 2) & 31, c = i & 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    uint32_t* u = reinterpret_cast<uint32_t*>(&v);
    const int K = 9 * cin;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = 0.f, b = 0.f;
        const int k0 = c * 8 + 2 * j;
        if (k0 < K) a = w[row * K + k0];
        if (k0 + 1 < K) b = w[row * K + k0 + 1];
        const float ah = __bfloat162float(__float2bfloat16_rn(a)), bh = __bfloat162float(__float2bfloat16_rn(b));
        if (part) { a -= ah; b -= bh; } else { a = ah; b = bh; }
        u[j] = pack_bf162(a, b);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(image) + part * 2048 + sw64_off(row, c)) = v;
}

template <int CIN>
__global__ void __launch_bounds__(StemTcCfg::THREADS, 3) stem_tc_kernel(const StemTcParams p) {
    typedef StemTcCfg C;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    // two operand stages (A_hi, A_lo each), weights, output staging, bias, barriers
    const uint32_t a_s = smem_base, w_s = a_s + 4 * C::A_BYTES, out_s = w_s + C::W_BYTES;
    const uint32_t bias_s = out_s + C::OUT_BYTES, bar_s = bias_s + 128;
    const uint32_t a_full = bar_s, d_full = bar_s + 16, tmem_slot = bar_s + 32;      // [2] each
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- set-up: nothing here depends on the previous kernel --------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(a_full + 8 * i, 4);        // one arrival per gather warp
            tc::mbar_init(d_full + 8 * i, 1);        // tcgen05.commit
        }
        tc::fence_barrier_init();
    }
    if (warp == 4) tc::tmem_alloc(tmem_slot, 64);
    {   // weight image -> shared memory (frozen layer: written at pack time, long before this launch)
        const uint4* src = reinterpret_cast<const uint4*>(p.w);
        for (int i = tid; i < C::W_BYTES / 16; i += C::THREADS) *reinterpret_cast<uint4*>(smem + (w_s - smem_base) + i * 16) = __ldg(src + i);
    }
    if (tid < 32) reinterpret_cast<float*>(smem + (bias_s - smem_base))[tid] = p.bias ? __ldg(p.bias + tid) : 0.f;
    tc::fence_proxy_async();             // generic-proxy writes of the weights -> visible to the tensor core (async proxy)
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    PDL_SYNC();

    if (warp == 4) {
        // =========================== MMA issuer ===========================
        const uint32_t idesc = tc::make_idesc_bf16(128, 32);
        const uint64_t d0 = tc::make_desc_sw64(0, 512, 0);
        const uint32_t hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
        const uint32_t bh = lo0 + (w_s >> 4), bl = bh + (2048 >> 4);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            const uint32_t s = it & 1;
            const uint32_t ah = lo0 + ((a_s + s * 2 * C::A_BYTES) >> 4), al = ah + (C::A_BYTES >> 4);
            const uint32_t d_tmem = tmem_base + s * 32;
            tc::mbar_wait(a_full + 8 * s, (it >> 1) & 1);
            tc::tc_fence_after();
            if (elect_one()) {
                tc::umma_f16_split<false>(d_tmem, ah, hi, bh, hi, idesc);
                if (CIN > 1) tc::umma_f16_split<true>(d_tmem, ah + 2, hi, bh + 2, hi, idesc);      // K step 1: k = 16..31 (all zero for CIN = 1)
                tc::umma_f16_split<true>(d_tmem, al, hi, bh, hi, idesc);
                if (CIN > 1) tc::umma_f16_split<true>(d_tmem, al + 2, hi, bh + 2, hi, idesc);
                tc::umma_f16_split<true>(d_tmem, ah, hi, bl, hi, idesc);
                if (CIN > 1) tc::umma_f16_split<true>(d_tmem, ah + 2, hi, bl + 2, hi, idesc);
                tc::umma_commit(d_full + 8 * s);
            }
            __syncwarp();
        }
    } else {
        // =========================== gather -> (MMA) -> epilogue: thread tid = pixel tid of a tile = TMEM lane tid ===========================
        // Software pipeline per warp: operand rows of tile i are written and handed to the MMA warp, the loads of tile i+1's window are
        // issued, and only then the accumulators of tile i-1 are drained -- the global-load latency and the MMA round trip both run
        // under the epilogue of the previous tile (two operand stages, two 32-column accumulators).
        constexpr int KV = 9 * CIN;
        const float4* bias4 = reinterpret_cast<const float4*>(smem + (bias_s - smem_base));
        unsigned char* stage = smem + (out_s - smem_base);
        float vn[KV];
        int nx = 0, ny = -4;             // pixel coordinates of the gathered tile's pixel (ny = -4: outside every image -> all taps zero)
        const uint32_t total32 = (uint32_t)p.total, hw32 = (uint32_t)(p.H * p.W);      // launch_stem_tc checks N*H*W < 2^31: 32-bit index math
        auto gather = [&](int tile) {
#pragma unroll
            for (int k = 0; k < KV; ++k) vn[k] = 0.f;
            const uint32_t q = (uint32_t)tile * 128u + (uint32_t)tid;
            nx = 0; ny = -4;
            if (q < total32) {
                const uint32_t n = q / hw32, rem = q - n * hw32;
                const int y = (int)(rem / (uint32_t)p.W);
                const int x = (int)(rem - (uint32_t)y * (uint32_t)p.W);
                nx = x; ny = y;
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    const float* pl = p.plane[ci] + n * p.batch_stride[ci];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int gy = y + ky - 1;
                        if (gy < 0 || gy >= p.H) continue;
                        const float* row = pl + (size_t)gy * p.W + x;
                        // raw values; scale / shift are applied when the operand rows are built (zero stays zero: padding of the NORMALISED input)
                        if (x > 0) vn[ci * 9 + ky * 3 + 0] = __ldg(row - 1);
                        vn[ci * 9 + ky * 3 + 1] = __ldg(row);
                        if (x + 1 < p.W) vn[ci * 9 + ky * 3 + 2] = __ldg(row + 1);
                    }
                }
            }
        };
        auto epilogue = [&](int tile, uint32_t ite) {
            const uint32_t se = ite & 1;
            const uint32_t q = (uint32_t)tile * 128u + (uint32_t)tid;
            uint4 mk[4];
            if (p.mask && q < total32) {
#pragma unroll
                for (int g = 0; g < 4; ++g) mk[g] = __ldg(reinterpret_cast<const uint4*>(p.mask + (size_t)q * 32) + g);
            }
            tc::mbar_wait(d_full + 8 * se, (ite >> 1) & 1);
            tc::tc_fence_after();
            uint32_t acc[32];
            tc::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + se * 32, acc);
            tc::tc_fence_before();
            float f[32];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 b4 = bias4[c4];                               // broadcast read
                f[c4 * 4 + 0] = __uint_as_float(acc[c4 * 4 + 0]) + b4.x; f[c4 * 4 + 1] = __uint_as_float(acc[c4 * 4 + 1]) + b4.y;
                f[c4 * 4 + 2] = __uint_as_float(acc[c4 * 4 + 2]) + b4.z; f[c4 * 4 + 3] = __uint_as_float(acc[c4 * 4 + 3]) + b4.w;
            }
            if (p.mask && q < total32) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t b = positive_bits(mk[g]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[g * 8 + j] = (b >> j) & 1u ? f[g * 8 + j] : 0.f;
                }
            }
            if (p.relu_out) {
#pragma unroll
                for (int c = 0; c < 32; ++c) f[c] = fmaxf(f[c], 0.f);
            }
            // ---- bf16 -> this warp's 2 KB staging rows (generic proxy only: no fence, no CTA barrier) -> 512 B coalesced global stores ----
            unsigned char* wst = stage + warp * 2048;
            __syncwarp();                                              // the previous tile's reads of the staging rows are done
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 ov;
                ov.x = pack_bf162(f[g * 8 + 0], f[g * 8 + 1]); ov.y = pack_bf162(f[g * 8 + 2], f[g * 8 + 3]);
                ov.z = pack_bf162(f[g * 8 + 4], f[g * 8 + 5]); ov.w = pack_bf162(f[g * 8 + 6], f[g * 8 + 7]);
                *reinterpret_cast<uint4*>(wst + sw64_off(lane, g)) = ov;
            }
            __syncwarp();
            const long long q0 = (long long)tile * 128 + warp * 32;   // first pixel of this warp
            const int nvalid = (int)max((long long)0, min((long long)32, p.total - q0));
            uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)q0 * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = j * 32 + lane, r = i >> 2, c = i & 3;
                if (r < nvalid) dst[i] = *reinterpret_cast<const uint4*>(wst + sw64_off(r, c));
            }
        };

        int tile = blockIdx.x, prev_tile = -1;
        uint32_t it = 0;
        if (tile < p.tiles) gather(tile);
        for (; tile < p.tiles; tile += gridDim.x, ++it) {
            const uint32_t s = it & 1;
            unsigned char* a_hi = smem + (a_s - smem_base) + s * 2 * C::A_BYTES;
            unsigned char* a_lo = a_hi + C::A_BYTES;
            // ---- normalise, split into bf16 head + remainder, write this pixel's operand rows (stage s was last read by the MMAs of tile
            //      it-2, whose completion the epilogue of tile it-2 has already waited for) ----
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = 0.f;
            {
                const int x = nx, y = ny;                              // of THIS tile (set by the gather one iteration ago)
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    const float sc = p.scale[ci], sh = p.shift[ci];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int gy = y + ky - 1;
                        const bool rin = gy >= 0 && gy < p.H;
                        v[ci * 9 + ky * 3 + 0] = (rin && x > 0) ? fmaf(vn[ci * 9 + ky * 3 + 0], sc, sh) : 0.f;
                        v[ci * 9 + ky * 3 + 1] = rin ? fmaf(vn[ci * 9 + ky * 3 + 1], sc, sh) : 0.f;
                        v[ci * 9 + ky * 3 + 2] = (rin && x + 1 < p.W) ? fmaf(vn[ci * 9 + ky * 3 + 2], sc, sh) : 0.f;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < (CIN == 1 ? 2 : 4); ++c) {                 // CIN = 1: K step 1 (chunks 2, 3) is never read
                uint4 h4, l4;
                uint32_t* hu = reinterpret_cast<uint32_t*>(&h4);
                uint32_t* lu = reinterpret_cast<uint32_t*>(&l4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float a = v[c * 8 + 2 * j], b = v[c * 8 + 2 * j + 1];
                    const uint32_t hp = pack_bf162(a, b);
                    const float2 hf = unpack_bf162(hp);
                    hu[j] = hp;
                    lu[j] = pack_bf162(a - hf.x, b - hf.y);
                }
                *reinterpret_cast<uint4*>(a_hi + sw64_off(tid, c)) = h4;
                *reinterpret_cast<uint4*>(a_lo + sw64_off(tid, c)) = l4;
            }
            tc::fence_proxy_async();                                   // operand rows -> visible to the tensor core
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(a_full + 8 * s);
            const int next = tile + (int)gridDim.x;
            if (next < p.tiles) gather(next);                          // in flight during the epilogue below
            if (prev_tile >= 0) epilogue(prev_tile, it - 1);
            prev_tile = tile;
        }
        if (prev_tile >= 0) epilogue(prev_tile, it - 1);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 64);
    }
}

inline int launch_stem_tc(StemTcParams p, int cin, cudaStream_t st) {
    typedef StemTcCfg C;
    static int sms = 0;
    if (!sms) {
        PTTA_CUDA(cudaFuncSetAttribute(stem_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        PTTA_CUDA(cudaFuncSetAttribute(stem_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        PTTA_CUDA(cudaFuncSetAttribute(stem_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int dev = 0;
        PTTA_CUDA(cudaGetDevice(&dev));
        PTTA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    PTTA_CHECK(cin >= 1 && cin <= 3 && p.w && p.out, "stem_tc: bad arguments (cin=%d)", cin);
    p.total = (long long)p.N * p.H * p.W;
    PTTA_CHECK(p.total < (1LL << 31) - 65536, "stem_tc: %lld pixels exceed the 32-bit index range of the kernel", p.total);
    p.tiles = cdiv(p.total, 128);
    const int grid = p.tiles < 3 * sms ? p.tiles : 3 * sms;
    if (cin == 1) launch_k(stem_tc_kernel<1>, grid, C::THREADS, C::SMEM, st, p);
    else if (cin == 2) launch_k(stem_tc_kernel<2>, grid, C::THREADS, C::SMEM, st, p);
    else launch_k(stem_tc_kernel<3>, grid, C::THREADS, C::SMEM, st, p);
    return check_launch("stem_tc");
}

}  // namespace ptta
