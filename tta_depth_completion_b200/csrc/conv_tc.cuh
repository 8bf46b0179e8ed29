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
#include "common.cuh"
#include "conv_mma.cuh"

namespace ptta {

struct ConvTcParams {
    const bf16* w;       // [9][32][32] (tap, cout, cin)
    const float* bias;   // [32] or null
    bf16* out;
    const bf16* mask;    // relu-derivative mask source (same shape as out) or null
    const bf16* add;     // out = add + mask * (conv + bias), or null
    int N, H, W;
    int relu_in;         // apply ReLU to the input while it sits in shared memory
    int relu_out;        // store ReLU(result) (producer-side activation for consumers that only read ReLU(x))
    int strips, segs_y, rows_per_seg, total_segs;
};

namespace tc {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// same, descriptors passed as (lo, hi) halves: the issuing thread only ever adds to `lo` (start-address field)
template <bool ACC>
__device__ __forceinline__ void umma_f16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if (ACC)
        asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.eq.b32 p, 0, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                     ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, 0, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                     ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
        "%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_64B operand descriptor (cute::UMMA::SmemDescriptor, sm100 "version 1"):
// rows of 64 B, 8-row groups `sbo_bytes` apart, swizzle = XOR of address bits [4,6) with bits [7,9)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr, uint32_t sbo_bytes, int base_offset_mode) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major): 1
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    if (base_offset_mode) d |= (uint64_t)((saddr >> 7) & 7) << 49;
    d |= (uint64_t)4 << 61;                               // SWIZZLE_64B
    return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, N = n, M = m
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace tc

__device__ int g_tc_dbg = 0;   // timing experiments only (ptta_debug_set): 1 one MMA per row, 2 no epilogue stores, 4 no loads

struct ConvTcCfg {
    static const int RB = 16;                     // input-row slots in the shared-memory ring (power of two: cheap index math)
    static const int NSLOT = 16;                  // output-row accumulator slots in TMEM (16 x 32 columns = all 512)
    static const int BOXW = 130;                  // pixels per staged row (128 + one halo pixel each side)
    static const int SLOT_BYTES = 9216;           // 130 x 64 B rounded up to the 1024 B swizzle-pattern alignment
    static const int W_BYTES = 9 * 32 * 64;       // 18432
    static const int BAR_BYTES = 1024;
    static const int SMEM = 1024 /*align slack*/ + RB * SLOT_BYTES + W_BYTES + BAR_BYTES;
    static const int THREADS = 416;               // warp 0 MMA | warps 1-4 epilogue | warps 5-8 loaders | warps 9-12 transform
};

// the mbarrier receives one arrival from this thread once all of its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
        ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Scatter-form implicit GEMM.  Input row i of a segment (image row y0-1+i) contributes to output rows j = i-2, i-1, i with
// the vertical taps ky = 2, 1, 0.  The three taps are stacked along N, so one group of 6 MMAs (3 horizontal taps x 2 K
// steps, M128 x N96 x K16) consumes an input row exactly once and accumulates into three neighbouring 32-column TMEM
// slots; slot(t) = t mod 16 for the running output-row counter t.  A slot is complete after input row j+2, is drained
// and re-zeroed by the epilogue warps, and is reused 16 rows later.
__global__ void __launch_bounds__(ConvTcCfg::THREADS, 1) conv3x3_tc_kernel(const bf16* __restrict__ in, const ConvTcParams p) {
    typedef ConvTcCfg C;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t rows_s = smem_base;                                   // RB row slots
    const uint32_t w_s = smem_base + C::RB * C::SLOT_BYTES;              // weights, SW64 canonical, row = kx*96 + (2-ky)*32 + cout
    const uint32_t bar_s = w_s + C::W_BYTES;
    const uint32_t row_landed = bar_s;                       // [RB]    loaders -> transform     (128 async arrivals, cp.async completion)
    const uint32_t row_ready = bar_s + 8 * C::RB;            // [RB]    transform -> MMA        (128 arrivals, after ReLU + proxy fence)
    const uint32_t row_free = bar_s + 16 * C::RB;            // [RB]    MMA -> loaders          (tcgen05.commit)
    const uint32_t slot_full = bar_s + 24 * C::RB;           // [NSLOT] MMA -> epilogue         (tcgen05.commit)
    const uint32_t slot_empty = slot_full + 8 * C::NSLOT;    // [NSLOT] epilogue -> MMA         (128 arrivals, slot re-zeroed)
    const uint32_t tmem_slot = slot_empty + 8 * C::NSLOT;    // u32
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dbg = g_tc_dbg;

    // ---- one-time setup ----------------------------------------------------------------------------------
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::RB; ++i) {
            tc::mbar_init(row_landed + 8 * i, 128);
            tc::mbar_init(row_ready + 8 * i, 128);
            tc::mbar_init(row_free + 8 * i, 1);
        }
        for (int i = 0; i < C::NSLOT; ++i) {
            tc::mbar_init(slot_full + 8 * i, 1);
            tc::mbar_init(slot_empty + 8 * i, 128);
        }
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
    {   // weights -> shared memory: K-major SWIZZLE_64B, vertical taps stacked along N per horizontal tap
        unsigned char* wdst = smem + (w_s - smem_base);
        for (int i = tid; i < 9 * 32 * 4; i += C::THREADS) {
            const int row = i >> 2, c = i & 3;
            const int kx = row / 96, rem = row - kx * 96;
            const int ky = 2 - rem / 32, co = rem & 31;
            uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)((ky * 3 + kx) * 32 + co) * 32 + c * 8));
            *reinterpret_cast<uint4*>(wdst + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = v;
        }
        tc::fence_proxy_async();      // generic-proxy writes -> visible to the tensor core's async proxy
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (warp >= 1 && warp < 5) {      // all accumulator slots start at zero: every MMA accumulates
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int s2 = 0; s2 < C::NSLOT; ++s2) tmem_st32_zero(lane_base + s2 * 32);
        tmem_wait_st();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    const int seg_stride = gridDim.x;

    if (warp == 0) {
        // =========================== MMA issuer (one thread) ===========================
        if (lane == 0) {
            const uint32_t idesc32 = tc::make_idesc_bf16(128, 32), idesc64 = tc::make_idesc_bf16(128, 64), idesc96 = tc::make_idesc_bf16(128, 96);
            const uint64_t d0 = tc::make_desc_sw64(0, 512, 0);
            const uint32_t hi = (uint32_t)(d0 >> 32);
            const uint32_t lo0 = (uint32_t)d0;
            const uint32_t a_lo0 = lo0 + (rows_s >> 4), b_lo0 = lo0 + (w_s >> 4);
            uint32_t r = 0, t_base = 0;
            for (int seg = blockIdx.x; seg < p.total_segs; seg += seg_stride) {
                const int rem = seg % (p.strips * p.segs_y);
                const int sy = rem % p.segs_y;
                const int y0 = sy * p.rows_per_seg;
                const int nrows = min(y0 + p.rows_per_seg, p.H) - y0;
                for (int i = 0; i < nrows + 2; ++i, ++r) {
                    const uint32_t rs = r % C::RB;
                    if (i < nrows) {                                   // first touch of the slot of output row i in this round
                        const uint32_t tn = t_base + i;
                        tc::mbar_wait(slot_empty + 8 * (tn % C::NSLOT), ((tn / C::NSLOT) & 1) ^ 1);
                    }
                    tc::mbar_wait(row_ready + 8 * rs, (r / C::RB) & 1);
                    tc::tc_fence_after();
                    const int j_lo = max(i - 2, 0), j_hi = min(i, nrows - 1);
                    int cnt = j_hi - j_lo + 1;                         // 1..3 output rows receive this input row
                    int b_row = (2 - i + j_lo) * 32;                   // first stacked-weight row: ky = i - j_lo
                    uint32_t s0 = (t_base + j_lo) % C::NSLOT;
                    const uint32_t a_lo = a_lo0 + rs * (C::SLOT_BYTES >> 4);
                    while (cnt > 0) {
                        const int c1 = min(cnt, C::NSLOT - (int)s0);   // contiguous TMEM slots before the ring wraps
                        const uint32_t idesc = c1 == 3 ? idesc96 : (c1 == 2 ? idesc64 : idesc32);
                        const uint32_t d_tmem = tmem_base + s0 * 32;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                if ((dbg & 1) && (kx | ks)) continue;
                                tc::umma_f16_split<true>(d_tmem, a_lo + kx * 4 + ks * 2, hi, b_lo0 + (kx * 96 + b_row) * 4 + ks * 2, hi, idesc);
                            }
                        }
                        cnt -= c1; b_row += c1 * 32; s0 = 0;
                    }
                    tc::umma_commit(row_free + 8 * rs);                                    // this input row is never read again
                    if (i >= 2) tc::umma_commit(slot_full + 8 * ((t_base + i - 2) % C::NSLOT));   // output row i-2 is complete
                }
                t_base += nrows;
            }
        }
    } else if (warp < 5) {
        // =========================== epilogue ===========================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        float bias[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) bias[c] = p.bias ? __ldg(p.bias + c) : 0.f;
        uint32_t t = 0;
        for (int seg = blockIdx.x; seg < p.total_segs; seg += seg_stride) {
            const int n = seg / (p.strips * p.segs_y);
            const int rem = seg - n * p.strips * p.segs_y;
            const int sx = rem / p.segs_y, sy = rem - sx * p.segs_y;
            const int x = sx * 128 + q * 32 + lane, y0 = sy * p.rows_per_seg;
            const int y1 = min(y0 + p.rows_per_seg, p.H);
            for (int y = y0; y < y1; ++y, ++t) {
                const uint32_t sl = t % C::NSLOT;
                tc::mbar_wait(slot_full + 8 * sl, (t / C::NSLOT) & 1);
                tc::tc_fence_after();
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + sl * 32;
                if (!(dbg & 8)) {
                    tc::tmem_ld32(taddr, v);
                    tmem_st32_zero(taddr);               // hand the slot back zeroed
                    tmem_wait_st();
                }
                tc::tc_fence_before();
                tc::mbar_arrive(slot_empty + 8 * sl);
                if (x < p.W && !(dbg & 2)) {
                    const size_t off = (((size_t)n * p.H + y) * p.W + x) * 32;
                    float f[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) f[c] = __uint_as_float(v[c]) + bias[c];
                    if (p.mask) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 mv = __ldg(reinterpret_cast<const uint4*>(p.mask + off + g * 8));
                            const uint32_t* mu = reinterpret_cast<const uint32_t*>(&mv);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 m = unpack_bf162(mu[j]);
                                if (!(m.x > 0.f)) f[g * 8 + j * 2] = 0.f;
                                if (!(m.y > 0.f)) f[g * 8 + j * 2 + 1] = 0.f;
                            }
                        }
                    }
                    if (p.add) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 av = *reinterpret_cast<const uint4*>(p.add + off + g * 8);
                            const uint32_t* au = reinterpret_cast<const uint32_t*>(&av);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 a = unpack_bf162(au[j]);
                                f[g * 8 + j * 2] += a.x;
                                f[g * 8 + j * 2 + 1] += a.y;
                            }
                        }
                    }
                    if (p.relu_out) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) f[c] = fmaxf(f[c], 0.f);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 ov;
                        ov.x = pack_bf162(f[g * 8 + 0], f[g * 8 + 1]);
                        ov.y = pack_bf162(f[g * 8 + 2], f[g * 8 + 3]);
                        ov.z = pack_bf162(f[g * 8 + 4], f[g * 8 + 5]);
                        ov.w = pack_bf162(f[g * 8 + 6], f[g * 8 + 7]);
                        *reinterpret_cast<uint4*>(p.out + off + g * 8) = ov;
                    }
                }
            }
        }
    } else if (warp < 9) {
        // =========================== loaders: global -> shared (cp.async, SW64 swizzle) ===========================
        // Thread tt owns 16 B chunks tt, tt+128, ... of every row (pixel = chunk/4): coalesced 512 B per warp instruction.
        // Loaders never wait for data: completion is signalled to `row_landed` by cp.async.mbarrier.arrive.noinc, so up to
        // RB-1 rows stay in flight.
        const int tt = tid - 160;                        // 0..127
        uint32_t r = 0;
        for (int seg = blockIdx.x; seg < p.total_segs; seg += seg_stride) {
            const int n = seg / (p.strips * p.segs_y);
            const int rem = seg - n * p.strips * p.segs_y;
            const int sx = rem / p.segs_y, sy = rem - sx * p.segs_y;
            const int x0 = sx * 128, y0 = sy * p.rows_per_seg;
            const int y1 = min(y0 + p.rows_per_seg, p.H);
            for (int yy = y0 - 1; yy <= y1; ++yy, ++r) {
                const uint32_t slot = r % C::RB;
                tc::mbar_wait(row_free + 8 * slot, ((r / C::RB) & 1) ^ 1);
                const uint32_t dst = rows_s + slot * C::SLOT_BYTES;
                const bool row_ok = yy >= 0 && yy < p.H;
                const bf16* row_src = in + ((size_t)n * p.H + (row_ok ? yy : 0)) * p.W * 32;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int i = tt + k * 128;
                    if (i < C::BOXW * 4 && !(dbg & 4)) {
                        const int px = i >> 2, c = i & 3;
                        const int x = x0 - 1 + px;
                        const bool ok = row_ok && x >= 0 && x < p.W;
                        cp_async16_zfill(dst + px * 64 + ((c ^ ((px >> 1) & 3)) << 4), row_src + (size_t)(ok ? x : 0) * 32 + c * 8, ok ? 16u : 0u);
                    }
                }
                cp_async_mbar_arrive_noinc(row_landed + 8 * slot);
            }
        }
        cp_async_wait_all();
    } else {
        // =========================== transform: optional ReLU in place, then publish the row to the tensor core ===========================
        const int tt = tid - 288;                        // 0..127
        const bf162 z = __floats2bfloat162_rn(0.f, 0.f);
        uint32_t r = 0;
        for (int seg = blockIdx.x; seg < p.total_segs; seg += seg_stride) {
            const int rem = seg % (p.strips * p.segs_y);
            const int sy = rem % p.segs_y;
            const int y0 = sy * p.rows_per_seg;
            const int nload = min(y0 + p.rows_per_seg, p.H) - y0 + 2;
            for (int k2 = 0; k2 < nload; ++k2, ++r) {
                const uint32_t slot = r % C::RB;
                tc::mbar_wait(row_landed + 8 * slot, (r / C::RB) & 1);
                if (p.relu_in) {
                    uint4* b4 = reinterpret_cast<uint4*>(smem + (rows_s - smem_base) + slot * C::SLOT_BYTES);
                    // swizzling permutes 16 B chunks inside a pixel's 64 B only and ReLU is elementwise: walk linearly
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int i = tt + k * 128;
                        if (i < C::BOXW * 4) {
                            uint4 v = b4[i];
                            bf162* h = reinterpret_cast<bf162*>(&v);
#pragma unroll
                            for (int j = 0; j < 4; ++j) h[j] = __hmax2(h[j], z);
                            b4[i] = v;
                        }
                    }
                }
                // the row was written through the generic proxy (cp.async, ReLU stores) and observed complete through the
                // mbarrier: order it before the tensor core's async-proxy reads, then hand it to the MMA thread
                if (!(dbg & 16)) tc::fence_proxy_async();
                tc::mbar_arrive(row_ready + 8 * slot);
            }
        }
    }

    // ---- teardown -----------------------------------------------------------------------------------------
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)ptr;
    }
    return fn;
}

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

inline int launch_conv_tc(const bf16* in, ConvTcParams p, cudaStream_t st) {
    typedef ConvTcCfg C;
    static int max_ctas = 0;
    if (!max_ctas) {
        PTTA_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int dev = 0, sms = 0, occ = 0;
        PTTA_CUDA(cudaGetDevice(&dev));
        PTTA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PTTA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv3x3_tc_kernel, C::THREADS, C::SMEM));
        PTTA_CHECK(occ >= 1, "conv3x3_tc does not fit on an SM");
        max_ctas = sms * occ;
    }
    p.strips = cdiv(p.W, 128);
    int want = 2 * 148;
    int segs = cdiv(want, p.N * p.strips);
    if (segs < 1) segs = 1;
    if (segs > p.H) segs = p.H;
    p.rows_per_seg = cdiv(p.H, segs);
    p.segs_y = cdiv(p.H, p.rows_per_seg);
    p.total_segs = p.N * p.strips * p.segs_y;
    int grid = p.total_segs < max_ctas ? p.total_segs : max_ctas;
    conv3x3_tc_kernel<<<grid, C::THREADS, C::SMEM, st>>>(in, p);
    return check_launch("conv3x3_tc");
}

}  // namespace ptta
