// 3x3 stride-1 32->32 NHWC bf16 convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators
// in TMEM, operands staged by TMA) -- the layer shape that carries most of MSG-CHN's FLOPs and bytes
// (network_exp_msg_chn_adapt.py:166-311: every `conv(ReLU(x))` of the encoders / decoders, and their
// data gradients, which are the same convolution with flipped taps).
//
// Implicit GEMM per output row segment: M = 128 consecutive pixels of one image row, N = 32 output channels,
// K = 9 taps x 32 input channels = 18 UMMA steps of K = 16.  A persistent CTA walks down a vertical strip
// (128 columns wide): every new output row needs ONE new input row from HBM/L2 (TMA box of 130 pixels x 64 B,
// zero-filled outside the image = the conv's padding); the previous two rows are still resident in a ring of
// row buffers.  The three horizontal taps of a row are three UMMA descriptors into the same buffer, offset by
// one pixel (64 B) each; SWIZZLE_64B is a function of the shared-memory address bits, so TMA (writer) and UMMA
// (reader) agree for any start offset.  (SHIFT_TMA=true is the conservative variant: three TMA copies of each row,
// shifted by -1/0/+1 pixel, every descriptor start 1024 B aligned.)
//
// Warp roles (320 threads): warp 0 TMA producer | warp 1 TMEM allocator + MMA issuer | warps 2-5 epilogue
// (TMEM -> registers -> bias / derivative mask / add -> bf16 -> 64 B per pixel, fully coalesced) | warps 6-9
// prologue transform (ReLU in place on freshly landed rows, then fence.proxy.async).  Double-buffered
// accumulators (2 x 32 TMEM columns) overlap the epilogue of row y with the MMAs of row y+1.
#pragma once
#include <cuda.h>
#include <vector>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "conv_mma.cuh"

namespace ptta {

struct ConvTcParams {
    const bf16* w;       // 18 KB shared-memory image of the weights (pack_conv_weight_tc_kernel)
    const float* bias;   // [32] or null
    bf16* out;
    bf16* out2;          // optional second output: ReLU(out [+ add2]) (for consumers that have no ReLU-on-load), or null
    const bf16* add2;    // optional addend of out2 only (decoder sums s = ReLU(conv + skip)); mutually exclusive with `add`
    const bf16* mask;    // relu-derivative mask source (same shape as out) or null
    const bf16* add;     // out = add + mask * (conv + bias), or null
    int N, H, W;
    int relu_in;         // apply ReLU to the input while it sits in shared memory
    int relu_out;        // store ReLU(result) (producer-side activation for consumers that only read ReLU(x))
    int strips, segs_y, rows_per_seg, total_segs;
};


__device__ long long g_tc_ts[3 * 2048];   // timing experiments (dbg & 64): per-row clock64() stamps of CTA 0 (MMA | epilogue warp | producer)
#define TC_TS(role, row, k) do { if ((dbg & 64) && blockIdx.x == 0 && (row) < 256) g_tc_ts[(role) * 2048 + (row) * 8 + (k)] = clock64(); } while (0)
__device__ int g_tc_dbg = 0;   // timing experiments only (ptta_debug_set): 1 one MMA pair per row, 2 no epilogue stores, 4 no loads, 64 stamps

struct ConvTcCfg {
    static const int RB = 8;                      // input-row slots in the shared-memory ring
    static const int NSLOT = 8;                   // output-row accumulator slots per pixel parity (2 x 8 x 32 columns = all 512)
    static const int BOXP = 130;                  // pixel PAIRS per staged row (128 + one halo pair each side)
    static const int ROW_BYTES = BOXP * 128;      // 16640: what one TMA box delivers
    static const int SLOT_BYTES = 17408;          // ROW_BYTES rounded up to the 1024 B swizzle-pattern alignment
    static const int W_BYTES = 9 * 32 * 64;       // 18432
    static const int STAGE_BYTES = 4096;          // one epilogue warp's 32 pixel pairs x 128 B (output / `add` staging)
    static const int MSTAGE_BYTES = 256;          // one epilogue warp's mask bits: one byte per 16 B chunk
    static const int EPI_BYTES = 2 * STAGE_BYTES + MSTAGE_BYTES;
    static const int BAR_BYTES = 1024;
    static const int SMEM = 1024 /*align slack*/ + RB * SLOT_BYTES + W_BYTES + 4 * EPI_BYTES + BAR_BYTES;
    static const int THREADS = 192;               // warp 0 TMA producer | warp 1 MMA issuer | warps 2-5 epilogue
};

// one bit per bf16 of a 16 B chunk: value > 0
__device__ __forceinline__ uint32_t positive_bits(const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t b = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        b |= ((short)(w[j] & 0xffffu) > 0 ? 1u : 0u) << (2 * j);
        b |= ((int)w[j] >= 0x10000 ? 1u : 0u) << (2 * j + 1);
    }
    return b;
}

// bias / derivative mask / add / ReLU on one pixel's 32 accumulators -> four 16 B chunks of bf16
// mbits: bit c set = keep channel c (all ones without a mask); addp: this pixel's four staged `add` chunks or null
__device__ __forceinline__ void conv_tc_finish_pixel(const uint32_t (&v)[32], const float (&bias)[32], uint32_t mbits, const unsigned char* addrow,
                                                     int chunk0, int swz, int relu_out, uint4 (&ov)[4]) {
    float f[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) f[c] = (mbits >> c) & 1u ? __uint_as_float(v[c]) + bias[c] : 0.f;
    if (addrow) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint4 av = *reinterpret_cast<const uint4*>(addrow + (((chunk0 + g) ^ swz) << 4));
            const uint32_t* au = reinterpret_cast<const uint32_t*>(&av);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 a = unpack_bf162(au[j]);
                f[g * 8 + j * 2] += a.x;
                f[g * 8 + j * 2 + 1] += a.y;
            }
        }
    }
    if (relu_out) {
#pragma unroll
        for (int c = 0; c < 32; ++c) f[c] = fmaxf(f[c], 0.f);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        ov[g].x = pack_bf162(f[g * 8 + 0], f[g * 8 + 1]);
        ov[g].y = pack_bf162(f[g * 8 + 2], f[g * 8 + 3]);
        ov[g].z = pack_bf162(f[g * 8 + 4], f[g * 8 + 5]);
        ov[g].w = pack_bf162(f[g * 8 + 6], f[g * 8 + 7]);
    }
}

// [tap][cout][cin] bf16 (pack_conv_weight_kernel) -> the kernel's shared-memory weight image: row = kx*96 + (2-ky)*32 + cout,
// 64 B per row, 16 B chunks XOR-swizzled exactly as SWIZZLE_64B lays them out at a 512 B aligned base.  One bulk copy stages it.
__global__ void pack_conv_weight_tc_kernel(const bf16* __restrict__ pack, bf16* __restrict__ image) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * 32 * 4) return;
    const int row = i >> 2, c = i & 3;
    const int kx = row / 96, rem = row - kx * 96;
    const int ky = 2 - rem / 32, co = rem & 31;
    uint4 v = *reinterpret_cast<const uint4*>(pack + (size_t)((ky * 3 + kx) * 32 + co) * 32 + c * 8);
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(image) + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = v;
}

// Scatter-form implicit GEMM over PIXEL PAIRS.
//
// Shared-memory / TMA view of the NHWC bf16 input: [N][H][W/2][64] -- one 128 B row per pixel pair, SWIZZLE_128B, the layout
// TMA moves at full speed (64 B rows do not).  A strip is 128 pairs = 256 output pixels wide; a staged input row holds 130
// pairs (one halo pair each side, zero-filled by TMA outside the image = the convolution's padding).  The UMMA A operand is
// "128 rows of 128 B"; a K = 16 slice at byte offset 0/32 of a row is the EVEN pixel of each pair, at 64/96 the ODD pixel,
// and a descriptor may start at any pair row (the swizzle is a function of the shared-memory address).  Hence
//     even output pixel 2j   = W[kx=0] . odd(j-1) + W[kx=1] . even(j) + W[kx=2] . odd(j)
//     odd  output pixel 2j+1 = W[kx=0] . even(j)  + W[kx=1] . odd(j)  + W[kx=2] . even(j+1)
// with no wasted FLOPs, and the even / odd outputs of one lane are 128 contiguous bytes of the output row.
//
// Vertical taps: input row i of a segment (image row y0-1+i) contributes to output rows j = i-2, i-1, i with ky = 2, 1, 0.
// The three taps are stacked along N (weights pre-arranged as [kx][(2-ky)*32+cout][cin]), so 12 MMAs (2 parities x 3 kx x 2 K
// steps, M128 x N96 x K16) consume an input row exactly once and accumulate into three neighbouring 32-column TMEM slots of
// each parity bank; slot(t) = t mod 8 of the running output-row counter.  A slot is complete after input row j+2, is
// drained and re-zeroed by the epilogue warps and reused 8 rows later.
//
// Roles: warp 0 TMA producer (one box per input row) | warp 1 MMA issuer (warp stays converged, elect.sync issues) |
// warps 2-5 epilogue: TMEM -> registers -> bias / mask / add / ReLU -> bf16 -> warp-private transpose through shared
// memory -> 512 B contiguous global stores.  Every mbarrier has one arrival per phase (slot_empty: one per epilogue warp).
__global__ void __launch_bounds__(ConvTcCfg::THREADS, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const ConvTcParams p) {
    typedef ConvTcCfg C;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t rows_s = smem_base;                                   // RB row slots
    const uint32_t w_s = smem_base + C::RB * C::SLOT_BYTES;              // weights, SW64 canonical, row = kx*96 + (2-ky)*32 + cout
    const uint32_t stage_s = w_s + C::W_BYTES;                           // 4 epilogue warps x (out 4 KB | add 4 KB | mask bits 256 B)
    const uint32_t bar_s = stage_s + 4 * C::EPI_BYTES;
    const uint32_t row_full = bar_s;                         // [RB]    TMA     -> MMA       (expect_tx + complete_tx)
    const uint32_t row_free = bar_s + 8 * C::RB;             // [RB]    MMA     -> producer  (tcgen05.commit)
    const uint32_t slot_full = bar_s + 16 * C::RB;           // [NSLOT] MMA     -> epilogue  (tcgen05.commit)
    const uint32_t slot_empty = slot_full + 8 * C::NSLOT;    // [NSLOT] epilogue -> MMA      (4 arrivals: one per epilogue warp)
    const uint32_t w_full = slot_empty + 8 * C::NSLOT;       //         bulk copy of the weight image landed
    const uint32_t tmem_ready = w_full + 8;                  //         accumulator slots zeroed (4 arrivals)
    const uint32_t tmem_slot = tmem_ready + 8;               // u32
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dbg = g_tc_dbg;
    if ((dbg & 64) && blockIdx.x == 0 && tid == 0) g_tc_ts[3 * 2048 - 8] = clock64();

    // ---- one-time setup: one block-wide barrier, everything else overlaps with the first row loads ------------------
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
    if ((dbg & 64) && blockIdx.x == 0 && tid == 0) g_tc_ts[3 * 2048 - 7] = clock64();

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
            const int y1 = min(y0 + p.rows_per_seg, p.H);
            for (int yy = y0 - 1; yy <= y1; ++yy, ++r) {
                const uint32_t slot = r % C::RB;
                if (lane == 0) TC_TS(2, r, 0);
                tc::mbar_wait(row_free + 8 * slot, ((r / C::RB) & 1) ^ 1);
                if (lane == 0) TC_TS(2, r, 1);
                if (elect_one()) {
                    if (!(dbg & 4)) {
                        tc::mbar_arrive_expect_tx(row_full + 8 * slot, C::ROW_BYTES);
                        tc::tma_load_4d(rows_s + slot * C::SLOT_BYTES, &tmap_in, row_full + 8 * slot, 0, sx * 128 - 1, yy, n);
                    } else {
                        tc::mbar_arrive(row_full + 8 * slot);
                    }
                }
                __syncwarp();
                if (lane == 0) TC_TS(2, r, 2);
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer (converged warp, one elected lane issues) ===========================
        const uint32_t idesc32 = tc::make_idesc_bf16(128, 32), idesc64 = tc::make_idesc_bf16(128, 64), idesc96 = tc::make_idesc_bf16(128, 96);
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
            const int nrows = min(y0 + p.rows_per_seg, p.H) - y0;
            for (int i = 0; i < nrows + 2; ++i, ++r) {
                const uint32_t rs = r % C::RB;
                if (lane == 0) TC_TS(0, r, 0);
                if (i < nrows) {                                   // first touch of the slot of output row i in this round
                    const uint32_t tn = t_base + i;
                    tc::mbar_wait(slot_empty + 8 * (tn % C::NSLOT), ((tn / C::NSLOT) & 1) ^ 1);
                }
                if (lane == 0) TC_TS(0, r, 1);
                tc::mbar_wait(row_full + 8 * rs, (r / C::RB) & 1);
                tc::tc_fence_after();
                if (lane == 0) TC_TS(0, r, 2);
                if (elect_one()) {
                    const int j_lo = max(i - 2, 0), j_hi = min(i, nrows - 1);
                    int cnt = j_hi - j_lo + 1;                         // 1..3 output rows receive this input row
                    int b_row = (2 - i + j_lo) * 32;                   // first stacked-weight row: ky = i - j_lo
                    uint32_t s0 = (t_base + j_lo) % C::NSLOT;
                    const uint32_t a_lo = a_lo0 + rs * (C::SLOT_BYTES >> 4);
                    while (cnt > 0) {
                        const int c1 = min(cnt, C::NSLOT - (int)s0);   // contiguous TMEM slots before the ring wraps
                        const uint32_t idesc = c1 == 3 ? idesc96 : (c1 == 2 ? idesc64 : idesc32);
                        const uint32_t d_even = tmem_base + s0 * 32, d_odd = d_even + 32 * C::NSLOT;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            // 16 B units inside the slot: pair row = 8 units, odd pixel = +4, K step = +2
                            const uint32_t ae = kx == 0 ? 4u : (kx == 1 ? 8u : 12u);      // odd(j-1) | even(j) | odd(j)
                            const uint32_t ao = kx == 0 ? 8u : (kx == 1 ? 12u : 16u);     // even(j)  | odd(j)  | even(j+1)
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                if ((dbg & 1) && (kx | ks)) continue;
                                const uint32_t b_lo = b_lo0 + (kx * 96 + b_row) * 4 + ks * 2;
                                tc::umma_f16_split<true>(d_even, a_lo + ae + ks * 2, a_hi, b_lo, b_hi, idesc);
                                tc::umma_f16_split<true>(d_odd, a_lo + ao + ks * 2, a_hi, b_lo, b_hi, idesc);
                            }
                        }
                        cnt -= c1; b_row += c1 * 32; s0 = 0;
                    }
                    tc::umma_commit(row_free + 8 * rs);                                    // this input row is never read again
                    if (i >= 2) tc::umma_commit(slot_full + 8 * ((t_base + i - 2) % C::NSLOT));   // output row i-2 is complete
                }
                __syncwarp();
                if (lane == 0) TC_TS(0, r, 3);
            }
            t_base += nrows;
        }
    } else {
        // =========================== epilogue ===========================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        unsigned char* stage = smem + (stage_s - smem_base) + q * C::EPI_BYTES;       // output transpose
        unsigned char* astage = stage + C::STAGE_BYTES;                               // `add` transpose
        unsigned char* mstage = astage + C::STAGE_BYTES;                              // mask bits
        {   // all accumulator slots start at zero: every MMA accumulates
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
            for (int s2 = 0; s2 < 2 * C::NSLOT; ++s2) tmem_st32_zero(lane_base + s2 * 32);
            tmem_wait_st();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tmem_ready);
        }
        float bias[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) bias[c] = p.bias ? __ldg(p.bias + c) : 0.f;
        const int swz = lane & 7;
        uint32_t t = 0;
        for (int seg = blockIdx.x; seg < p.total_segs; seg += seg_stride) {
            const int n = seg / segs_per_image;
            const int rem = seg - n * segs_per_image;
            const int sx = rem / p.segs_y, sy = rem - sx * p.segs_y;
            const int xw = sx * 256 + q * 64, y0 = sy * p.rows_per_seg;      // first pixel of this warp's 64
            const int y1 = min(y0 + p.rows_per_seg, p.H);
            const int vp = min(32, (p.W - xw) / 2);                          // valid pixel pairs of this warp (may be <= 0)
            for (int y = y0; y < y1; ++y, ++t) {
                const uint32_t sl = t % C::NSLOT;
                const size_t off0 = (((size_t)n * p.H + y) * p.W + xw) * 32;     // first element of the warp's 64 pixels
                // mask / add rows are fetched coalesced (512 B per warp instruction) BEFORE waiting for the accumulators
                uint4 mk[8], ad[8];
                if (p.mask) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int idx = k * 32 + lane;
                        mk[k] = (idx >> 3) < vp ? __ldg(reinterpret_cast<const uint4*>(p.mask + off0 + (size_t)idx * 8)) : make_uint4(0, 0, 0, 0);
                    }
                }
                const bf16* addsrc = p.add ? p.add : p.add2;             // same registers: the two are never used together
                if (addsrc) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int idx = k * 32 + lane;
                        ad[k] = (idx >> 3) < vp ? *reinterpret_cast<const uint4*>(addsrc + off0 + (size_t)idx * 8) : make_uint4(0, 0, 0, 0);   // may alias `out`
                    }
                }
                if (q == 2 && lane == 0) TC_TS(1, t, 0);
                tc::mbar_wait(slot_full + 8 * sl, (t / C::NSLOT) & 1);
                tc::tc_fence_after();
                if (q == 2 && lane == 0) TC_TS(1, t, 1);
                uint32_t ve[32], vo[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + sl * 32;
                tmem_ld32_nowait(taddr, ve);
                tmem_ld32_nowait(taddr + 32 * C::NSLOT, vo);
                tmem_wait_ld();
                tmem_st32_zero(taddr);               // hand the slots back zeroed
                tmem_st32_zero(taddr + 32 * C::NSLOT);
                tmem_wait_st();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(slot_empty + 8 * sl);
                if (q == 2 && lane == 0) TC_TS(1, t, 2);
                if (vp <= 0 || (dbg & 2)) continue;
                // transpose mask bits / add chunks to "lane = pixel pair" through the warp's staging buffers
                if (p.mask) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) mstage[k * 32 + lane] = (unsigned char)positive_bits(mk[k]);
                }
                if (p.add) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int idx = k * 32 + lane, pp = idx >> 3, c = idx & 7;
                        *reinterpret_cast<uint4*>(astage + pp * 128 + ((c ^ (pp & 7)) << 4)) = ad[k];
                    }
                }
                if (p.mask || p.add) __syncwarp();
                if (lane < vp) {
                    uint32_t me = 0xffffffffu, mo = 0xffffffffu;
                    if (p.mask) {
                        const uint2 mb = *reinterpret_cast<const uint2*>(mstage + lane * 8);
                        me = mb.x; mo = mb.y;
                    }
                    const unsigned char* addrow = p.add ? astage + lane * 128 : nullptr;
                    uint4 ov[4];
                    conv_tc_finish_pixel(ve, bias, me, addrow, 0, swz, p.relu_out, ov);
#pragma unroll
                    for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(stage + lane * 128 + ((g ^ swz) << 4)) = ov[g];
                    conv_tc_finish_pixel(vo, bias, mo, addrow, 4, swz, p.relu_out, ov);
#pragma unroll
                    for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(stage + lane * 128 + (((g + 4) ^ swz) << 4)) = ov[g];
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; ++k) {        // 512 contiguous bytes per warp instruction
                    const int idx = k * 32 + lane, pp = idx >> 3, c = idx & 7;
                    if (pp < vp) {
                        uint4 val = *reinterpret_cast<const uint4*>(stage + pp * 128 + ((c ^ (pp & 7)) << 4));
                        *reinterpret_cast<uint4*>(p.out + off0 + (size_t)idx * 8) = val;
                        if (p.out2) {
                            if (p.add2) {        // out2 = ReLU(bf16(out) + add2): the chunk prefetched for this lane is the one it stores
                                uint32_t* vu = reinterpret_cast<uint32_t*>(&val);
                                const uint32_t* au = reinterpret_cast<const uint32_t*>(&ad[k]);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 a = unpack_bf162(vu[j]), b = unpack_bf162(au[j]);
                                    vu[j] = pack_bf162(a.x + b.x, a.y + b.y);
                                }
                            }
                            const bf162 z2 = __floats2bfloat162_rn(0.f, 0.f);
                            bf162* h2 = reinterpret_cast<bf162*>(&val);
#pragma unroll
                            for (int j = 0; j < 4; ++j) h2[j] = __hmax2(h2[j], z2);
                            *reinterpret_cast<uint4*>(p.out2 + off0 + (size_t)idx * 8) = val;
                        }
                    }
                }
                __syncwarp();                        // staging buffers are rewritten next row
                if (q == 2 && lane == 0) TC_TS(1, t, 3);
            }
        }
    }

    // ---- teardown -----------------------------------------------------------------------------------------
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
    if ((dbg & 64) && blockIdx.x == 0 && tid == 0) g_tc_ts[3 * 2048 - 6] = clock64();
}

// ---- host side ----------------------------------------------------------------------------------------------
// NHWC bf16 [N,H,W,C] activation, box = {C, box_w, 1, 1}, SWIZZLE_64B (C = 32) / 128B (C = 64), zero fill outside
inline int make_tmap_nhwc(CUtensorMap* map, const void* ptr, int N, int H, int W, int Cc, int box_w) {
    PFN_encodeTiled enc = get_encode_tiled();
    PTTA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2};
    cuuint32_t box[4] = {(cuuint32_t)Cc, (cuuint32_t)box_w, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMapSwizzle sw = Cc * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTTA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (N=%d H=%d W=%d C=%d box_w=%d)", (int)r, N, H, W, Cc, box_w);
    return 0;
}

// NHWC bf16 [N,H,W,32] activation viewed as [N][H][W/2][64]: box = {64, 130 pairs, 1, 1}, SWIZZLE_128B, zero fill outside
inline int make_tmap_pairs(CUtensorMap* map, const void* ptr, int N, int H, int W) {
    PFN_encodeTiled enc = get_encode_tiled();
    PTTA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {64, (cuuint64_t)(W / 2), (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 64, (cuuint64_t)H * W * 64};
    cuuint32_t box[4] = {64, (cuuint32_t)ConvTcCfg::BOXP, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTTA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (N=%d H=%d W=%d)", (int)r, N, H, W);
    return 0;
}

inline bool conv_tc_supported(int N, int H, int W) { return N >= 1 && H >= 1 && W >= 2 && (W % 2) == 0; }

// tensor maps are cached per (pointer, shape): the engine's arena addresses are fixed, so a step encodes nothing
inline int conv_tc_tmap(const bf16* in, int N, int H, int W, const CUtensorMap** out) {
    struct Entry { const void* ptr; int n, h, w; CUtensorMap map; };
    static thread_local std::vector<Entry> cache;
    for (const Entry& e : cache)
        if (e.ptr == in && e.n == N && e.h == H && e.w == W) { *out = &e.map; return 0; }
    if (cache.size() >= 512) cache.clear();
    Entry e; e.ptr = in; e.n = N; e.h = H; e.w = W;
    PTTA_TRY(make_tmap_pairs(&e.map, in, N, H, W));
    cache.push_back(e);
    *out = &cache.back().map;
    return 0;
}

inline int launch_conv_tc(const bf16* in, ConvTcParams p, cudaStream_t st) {
    typedef ConvTcCfg C;
    static int sms = 0;
    if (!sms) {
        PTTA_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int dev = 0;
        PTTA_CUDA(cudaGetDevice(&dev));
        PTTA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    PTTA_CHECK(conv_tc_supported(p.N, p.H, p.W), "conv3x3_tc: W=%d must be even", p.W);
    PTTA_CHECK(!p.relu_in, "conv3x3_tc: ReLU-on-load is not supported (producers store ReLU(x): relu_out)");
    p.strips = cdiv(p.W, 256);
    // rows per segment: minimise waves x (rows + halo + fixed cost) over the segment counts that fill the machine
    int best_rows = p.H; long long best_cost = -1;
    for (int segs = 1; segs <= p.H; ++segs) {
        const int rows = cdiv(p.H, segs);
        const long long total = (long long)p.N * p.strips * cdiv(p.H, rows);
        const long long waves = (total + sms - 1) / sms;
        const long long cost = waves * (rows + 2 + 4);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_rows = rows; }
    }
    p.rows_per_seg = best_rows;
    p.segs_y = cdiv(p.H, p.rows_per_seg);
    p.total_segs = p.N * p.strips * p.segs_y;
    int grid = p.total_segs < sms ? p.total_segs : sms;
    const CUtensorMap* map = nullptr;
    PTTA_TRY(conv_tc_tmap(in, p.N, p.H, p.W, &map));
    conv3x3_tc_kernel<<<grid, C::THREADS, C::SMEM, st>>>(*map, p);
    return check_launch("conv3x3_tc");
}

}  // namespace ptta
