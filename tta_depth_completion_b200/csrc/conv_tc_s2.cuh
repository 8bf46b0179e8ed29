// 3x3 STRIDE-2 32->32 NHWC bf16 convolution on tcgen05 -- the down-sampling layers of every MSG-CHN encoder
// (network_exp_msg_chn_adapt.py:175-185,223-243: enc{1..4}.1) and, with the matching packed weights, the data gradient of
// the ConvTranspose2d layers of the decoders (:276-283).  Same machinery as conv_tc.cuh (TMA-fed SWIZZLE_128B pixel-pair
// rows, scatter-form accumulation in TMEM, elect.sync issue, coalesced epilogue); what changes is the geometry:
//
//   output pixel x reads input pixels 2x-1, 2x, 2x+1 = odd(pair x-1), even(pair x), odd(pair x): the A operand rows are the
//   INPUT pixel pairs, so M = 128 output pixels <-> 128 (+1 halo) input pairs and the three horizontal taps are three
//   descriptor start offsets (64 B, 128 B, 192 B) into the staged row -- the stride costs nothing.
//   output row y reads input rows 2y-1, 2y, 2y+1: an odd input row 2y+1 feeds output rows y (ky = 2) and y+1 (ky = 0) with
//   ONE N = 64 MMA group (weights stacked [ky=2 | ky=0]), an even input row 2y feeds output row y (ky = 1, N = 32).
//   Every input row is read from HBM/L2 once; output row j is complete after input row 2j+1.
#pragma once
#include "conv_tc.cuh"

namespace ptta {

struct ConvTcS2Cfg {
    static const int RB = 8;                      // input-row slots in the shared-memory ring
    static const int NSLOT = 16;                  // output-row accumulator slots (16 x 32 columns = all 512)
    static const int BOXP = 130;                  // pixel pairs per staged row (box shared with the stride-1 kernel's tensor map)
    static const int ROW_BYTES = BOXP * 128;
    static const int SLOT_BYTES = 17408;
    static const int W_BYTES = 9 * 32 * 64;
    static const int STAGE_BYTES = 2048;          // one epilogue warp's 32 output pixels x 64 B
    static const int MSTAGE_BYTES = 128;
    static const int EPI_BYTES = 2 * STAGE_BYTES + MSTAGE_BYTES;
    static const int BAR_BYTES = 1024;
    static const int SMEM = 1024 + RB * SLOT_BYTES + W_BYTES + 4 * EPI_BYTES + BAR_BYTES;
    static const int THREADS = 192;               // warp 0 TMA producer | warp 1 MMA issuer | warps 2-5 epilogue
};

// [tap][cout][cin] bf16 -> shared-memory weight image of the stride-2 kernel: row = kx*96 + blk*32 + cout with
// blk 0 = ky 2, blk 1 = ky 0 (the pair an odd input row feeds), blk 2 = ky 1; 64 B rows, SWIZZLE_64B chunk order
__global__ void pack_conv_weight_tc_s2_kernel(const bf16* __restrict__ pack, bf16* __restrict__ image) {
    PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * 32 * 4) return;
    const int row = i >> 2, c = i & 3;
    const int kx = row / 96, rem = row - kx * 96;
    const int blk = rem / 32, co = rem & 31;
    const int ky = blk == 0 ? 2 : (blk == 1 ? 0 : 1);
    uint4 v = *reinterpret_cast<const uint4*>(pack + (size_t)((ky * 3 + kx) * 32 + co) * 32 + c * 8);
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(image) + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = v;
}

__device__ __forceinline__ uint4 relu_bf16x8(uint4 v) {
    const bf162 z = __floats2bfloat162_rn(0.f, 0.f);
    bf162* h = reinterpret_cast<bf162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __hmax2(h[j], z);
    return v;
}

// p.H, p.W: INPUT size (both even); output (H/2) x (W/2); p.rows_per_seg / p.segs_y count OUTPUT rows
__global__ void __launch_bounds__(ConvTcS2Cfg::THREADS, 1) conv3x3_tc_s2_kernel(const __grid_constant__ CUtensorMap tmap_in, const ConvTcParams p) {
    typedef ConvTcS2Cfg C;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t rows_s = smem_base;
    const uint32_t w_s = smem_base + C::RB * C::SLOT_BYTES;
    const uint32_t stage_s = w_s + C::W_BYTES;
    const uint32_t bar_s = stage_s + 4 * C::EPI_BYTES;
    const uint32_t row_full = bar_s;
    const uint32_t row_free = bar_s + 8 * C::RB;
    const uint32_t slot_full = bar_s + 16 * C::RB;
    const uint32_t slot_empty = slot_full + 8 * C::NSLOT;
    const uint32_t w_full = slot_empty + 8 * C::NSLOT;
    const uint32_t tmem_ready = w_full + 8;
    const uint32_t tmem_slot = tmem_ready + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Ho = p.H >> 1, Wo = p.W >> 1;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_in);
        for (int i = 0; i < C::RB; ++i) {
            tc::mbar_init(row_full + 8 * i, 1);
            tc::mbar_init(row_free + 8 * i, 1);
        }
        for (int i = 0; i < C::NSLOT; ++i) {
            tc::mbar_init(slot_full + 8 * i, 1);
            tc::mbar_init(slot_empty + 8 * i, 4);
        }
        tc::mbar_init(w_full, 1);
        tc::mbar_init(tmem_ready, 4);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    PDL_SYNC();      // the set-up above (barriers, TMEM, descriptor prefetch) overlaps the previous kernel; nothing before this line touches global data

    const int seg_stride = gridDim.x;
    const int segs_per_image = p.strips * p.segs_y;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (elect_one()) {
            tc::mbar_arrive_expect_tx(w_full, C::W_BYTES);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(w_s), "l"(p.w), "r"((uint32_t)C::W_BYTES), "r"(w_full) : "memory");
        }
        __syncwarp();
        uint32_t r = 0;
        for (int seg = blockIdx.x; seg < p.total_segs; seg += seg_stride) {
            const int n = seg / segs_per_image;
            const int rem = seg - n * segs_per_image;
            const int sx = rem / p.segs_y, sy = rem - sx * p.segs_y;
            const int y0 = sy * p.rows_per_seg;
            const int nrows = min(y0 + p.rows_per_seg, Ho) - y0;
            for (int i = 0; i <= 2 * nrows; ++i, ++r) {
                const uint32_t slot = r % C::RB;
                tc::mbar_wait(row_free + 8 * slot, ((r / C::RB) & 1) ^ 1);
                if (elect_one()) {
                    tc::mbar_arrive_expect_tx(row_full + 8 * slot, C::ROW_BYTES);
                    tc::tma_load_4d(rows_s + slot * C::SLOT_BYTES, &tmap_in, row_full + 8 * slot, 0, sx * 128 - 1, 2 * y0 - 1 + i, n);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        const uint32_t idesc32 = tc::make_idesc_bf16(128, 32), idesc64 = tc::make_idesc_bf16(128, 64);
        const uint64_t da0 = make_desc_sw128(0), db0 = tc::make_desc_sw64(0, 512, 0);
        const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32);
        const uint32_t a_lo0 = (uint32_t)da0 + (rows_s >> 4), b_lo0 = (uint32_t)db0 + (w_s >> 4);
        uint32_t r = 0, t_base = 0;
        tc::mbar_wait(w_full, 0);
        tc::mbar_wait(tmem_ready, 0);
        tc::tc_fence_after();
        for (int seg = blockIdx.x; seg < p.total_segs; seg += seg_stride) {
            const int rem = seg % segs_per_image;
            const int sy = rem % p.segs_y;
            const int y0 = sy * p.rows_per_seg;
            const int nrows = min(y0 + p.rows_per_seg, Ho) - y0;
            for (int i = 0; i <= 2 * nrows; ++i, ++r) {
                const uint32_t rs = r % C::RB;
                const bool odd_row = (i & 1) == 0;                 // i even <-> image row 2*y0-1+i is odd: feeds output rows jm (ky 2), jp (ky 0)
                const int jp = i >> 1, jm = jp - 1;
                if (odd_row && jp < nrows) {                       // first touch of output row jp's slot in this round
                    const uint32_t tn = t_base + jp;
                    tc::mbar_wait(slot_empty + 8 * (tn % C::NSLOT), ((tn / C::NSLOT) & 1) ^ 1);
                }
                tc::mbar_wait(row_full + 8 * rs, (r / C::RB) & 1);
                tc::tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo = a_lo0 + rs * (C::SLOT_BYTES >> 4);
                    // up to two MMA groups: {tmem slot, first stacked-weight row, N}
                    uint32_t g_slot[2], g_brow[2], g_idesc[2];
                    int ng = 0;
                    if (!odd_row) {
                        g_slot[0] = (t_base + jp) % C::NSLOT; g_brow[0] = 64; g_idesc[0] = idesc32; ng = 1;          // ky = 1 -> output row (i-1)/2 == jp
                    } else {
                        const bool vm = jm >= 0, vp = jp < nrows;
                        const uint32_t sm = (t_base + (uint32_t)max(jm, 0)) % C::NSLOT, sp = (t_base + jp) % C::NSLOT;
                        if (vm && vp && sm + 1 == sp) { g_slot[0] = sm; g_brow[0] = 0; g_idesc[0] = idesc64; ng = 1; }
                        else {
                            if (vm) { g_slot[ng] = sm; g_brow[ng] = 0; g_idesc[ng] = idesc32; ++ng; }
                            if (vp) { g_slot[ng] = sp; g_brow[ng] = 32; g_idesc[ng] = idesc32; ++ng; }
                        }
                    }
                    for (int g = 0; g < ng; ++g) {
                        const uint32_t d = tmem_base + g_slot[g] * 32;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const uint32_t ao = kx == 0 ? 4u : (kx == 1 ? 8u : 12u);     // odd(x-1) | even(x) | odd(x), 16 B units
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks)
                                tc::umma_f16_split<true>(d, a_lo + ao + ks * 2, a_hi, b_lo0 + (kx * 96 + g_brow[g]) * 4 + ks * 2, b_hi, g_idesc[g]);
                        }
                    }
                    tc::umma_commit(row_free + 8 * rs);
                    if (odd_row && jm >= 0) tc::umma_commit(slot_full + 8 * ((t_base + jm) % C::NSLOT));    // output row jm is complete
                }
                __syncwarp();
            }
            t_base += nrows;
        }
    } else {
        // =========================== epilogue ===========================
        const int q = warp & 3;
        unsigned char* stage = smem + (stage_s - smem_base) + q * C::EPI_BYTES;
        unsigned char* astage = stage + C::STAGE_BYTES;
        unsigned char* mstage = astage + C::STAGE_BYTES;
        {
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
            for (int s2 = 0; s2 < C::NSLOT; ++s2) tmem_st32_zero(lane_base + s2 * 32);
            tmem_wait_st();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tmem_ready);
        }
        float bias[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) bias[c] = p.bias ? __ldg(p.bias + c) : 0.f;
        const int swz = (lane >> 1) & 3;
        uint32_t t = 0;
        for (int seg = blockIdx.x; seg < p.total_segs; seg += seg_stride) {
            const int n = seg / segs_per_image;
            const int rem = seg - n * segs_per_image;
            const int sx = rem / p.segs_y, sy = rem - sx * p.segs_y;
            const int xw = sx * 128 + q * 32, y0 = sy * p.rows_per_seg;
            const int y1 = min(y0 + p.rows_per_seg, Ho);
            const int vx = min(32, Wo - xw);                                 // valid output pixels of this warp (may be <= 0)
            for (int y = y0; y < y1; ++y, ++t) {
                const uint32_t sl = t % C::NSLOT;
                const size_t off0 = (((size_t)n * Ho + y) * Wo + xw) * 32;
                uint4 mk[4], ad[4];
                if (p.mask) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = k * 32 + lane;
                        mk[k] = (idx >> 2) < vx ? __ldg(reinterpret_cast<const uint4*>(p.mask + off0 + (size_t)idx * 8)) : make_uint4(0, 0, 0, 0);
                    }
                }
                if (p.add) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = k * 32 + lane;
                        ad[k] = (idx >> 2) < vx ? *reinterpret_cast<const uint4*>(p.add + off0 + (size_t)idx * 8) : make_uint4(0, 0, 0, 0);
                    }
                }
                tc::mbar_wait(slot_full + 8 * sl, (t / C::NSLOT) & 1);
                tc::tc_fence_after();
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + sl * 32;
                tmem_ld32_nowait(taddr, v);
                tmem_wait_ld();
                tmem_st32_zero(taddr);
                tmem_wait_st();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(slot_empty + 8 * sl);
                if (vx <= 0) continue;
                if (p.mask) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) mstage[k * 32 + lane] = (unsigned char)positive_bits(mk[k]);
                }
                if (p.add) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = k * 32 + lane, px = idx >> 2, c = idx & 3;
                        *reinterpret_cast<uint4*>(astage + px * 64 + ((c ^ ((px >> 1) & 3)) << 4)) = ad[k];
                    }
                }
                if (p.mask || p.add) __syncwarp();
                uint4 ov[4];
                if (lane < vx) {
                    const uint32_t mb = p.mask ? *reinterpret_cast<const uint32_t*>(mstage + lane * 4) : 0xffffffffu;
                    conv_tc_finish_pixel(v, bias, mb, p.add ? astage + lane * 64 : nullptr, 0, swz, p.relu_out, ov);
#pragma unroll
                    for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(stage + lane * 64 + ((g ^ swz) << 4)) = ov[g];
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; ++k) {        // 512 contiguous bytes per warp instruction
                    const int idx = k * 32 + lane, px = idx >> 2, c = idx & 3;
                    if (px < vx) {
                        const uint4 val = *reinterpret_cast<const uint4*>(stage + px * 64 + ((c ^ ((px >> 1) & 3)) << 4));
                        *reinterpret_cast<uint4*>(p.out + off0 + (size_t)idx * 8) = val;
                        if (p.out2) *reinterpret_cast<uint4*>(p.out2 + off0 + (size_t)idx * 8) = relu_bf16x8(val);
                    }
                }
                __syncwarp();
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

inline bool conv_tc_s2_supported(int N, int H, int W) { return N >= 1 && H >= 2 && W >= 2 && (H % 2) == 0 && (W % 2) == 0; }

// p.H, p.W = input size; `in` is the [N,H,W,32] input, p.out the [N,H/2,W/2,32] output
inline int launch_conv_tc_s2(const bf16* in, ConvTcParams p, cudaStream_t st) {
    typedef ConvTcS2Cfg C;
    static int sms = 0;
    if (!sms) {
        PTTA_CUDA(cudaFuncSetAttribute(conv3x3_tc_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int dev = 0;
        PTTA_CUDA(cudaGetDevice(&dev));
        PTTA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    PTTA_CHECK(conv_tc_s2_supported(p.N, p.H, p.W), "conv3x3_tc_s2: H=%d, W=%d must be even", p.H, p.W);
    PTTA_CHECK(!p.relu_in, "conv3x3_tc_s2: ReLU-on-load is not supported (producers store ReLU(x))");
    const int Ho = p.H / 2, Wo = p.W / 2;
    p.strips = cdiv(Wo, 128);
    int best_rows = Ho; long long best_cost = -1;
    for (int segs = 1; segs <= Ho; ++segs) {
        const int rows = cdiv(Ho, segs);
        const long long total = (long long)p.N * p.strips * cdiv(Ho, rows);
        const long long waves = (total + sms - 1) / sms;
        const long long cost = waves * (2 * rows + 1 + 4);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_rows = rows; }
    }
    p.rows_per_seg = best_rows;
    p.segs_y = cdiv(Ho, p.rows_per_seg);
    p.total_segs = p.N * p.strips * p.segs_y;
    int grid = p.total_segs < sms ? p.total_segs : sms;
    const CUtensorMap* map = nullptr;
    PTTA_TRY(conv_tc_tmap(in, p.N, p.H, p.W, &map));
    launch_k(conv3x3_tc_s2_kernel, grid, C::THREADS, C::SMEM, st, *map, p);
    return check_launch("conv3x3_tc_s2");
}

}  // namespace ptta
