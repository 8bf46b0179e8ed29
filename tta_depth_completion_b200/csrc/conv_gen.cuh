// General-channel NHWC bf16 convolution family on the 5th-generation tensor cores, for the NLSPN network
// (external_src/NLSPN/src/model/nlspnmodel_adapt.py:384-448: ResNet34 layers 64..512 channels, the transposed-conv decoder
// with skip concatenation, the three full-resolution 128->64 heads) and every data gradient of those layers.
//
// One kernel, implicit GEMM:   out[pixel, n] = sum over K-items  A_src[pixel + (dy, dx), c0 .. c0+63] . Wpk[item][n][0..63]
//   * M tile = 128 output positions = a th x tw patch (8x16, 4x32, 2x64 or 1x128) of ONE image, brought by ONE 5-D TMA box per
//     K-item, {64 channels, tw, 1, th, 1}, SWIZZLE_128B, zero-filled outside the image (= the conv's padding).  The 5-D view
//     {PX*C, W/PX, PY, H/PY, N} addresses plain maps (PX = PY = 1) and the pixel-parity view (PX = PY = 2) a stride-2 conv
//     reads, so the stride costs nothing; the same view on the OUTPUT side places the four parity classes of a transposed
//     conv (and of the data gradient of a stride-2 conv).
//   * a K-item = (source tensor 0/1, channel offset, dx, dy, parity, 64-wide weight chunk): 3x3 / 1x1 taps, channel
//     chunks, the two halves of a skip concatenation (two tensor maps -- no concat copy), and the 1x1/s2 shortcut of a
//     ResNet down-sampling block folded into the data gradient of its 3x3/s2 conv are all just items.  The list sits in
//     the kernel parameters (constant bank).
//   * N tile = BN in {64, 128, 256} output channels; weights pre-packed [item][Cout][64] bf16 (pack_convg_kernel), one 3-D
//     TMA box {64, BN, 1} per item.  tcgen05.mma M128 x BN x K16, 4 per item, fp32 accumulators in TMEM, 2 x 256 columns
//     double buffered so the epilogue of tile t overlaps the MMAs of tile t+1.
//   * epilogue (4 warps): tcgen05.ld -> + bias -> bf16 -> SWIZZLE_128B staging tile in shared memory -> one TMA store per 64
//     channels (the store clips partial tiles, and addresses parity classes through the 5-D output view).
// Roles: warp 0 TMA producer | warp 1 TMEM allocator + MMA issuer, warp 2 second issuer (resident-weight layers: the two threads
// take alternate tiles, so one keeps the tensor pipe fed while the other waits / commits) | warps 3-6 epilogue.  Persistent CTAs.
#pragma once
#include <cuda.h>
#include <vector>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ptta {

struct ConvGItem {
    int c_inner;   // inner coordinate in the source view: px * C + channel offset
    int dx;        // added to the tile's grid x
    int dyps;      // dy (low 16 bits, signed) | py << 16 | src << 20 | new_a << 24 | last_of_a << 25
    int wk;        // K-chunk index in the packed weights
    int a_off16;   // start of this item's A operand inside the current A tile, in 16-byte units (halo mode: (ky*10 + kx)*8)
    int pad;
};

struct ConvGParams {
    ConvGItem items[128];
    int cls_start[4], cls_count[4];
    int cls_out_c[4];    // inner coordinate of the class in the output view (qx * Cout)
    int cls_out_py[4];   // qy
    int n_classes;
    int N, tiles_y, tiles_x, n_tiles, BN;
    int th, tw;
    int total_tiles;
    // operand rings: A tiles (16 KB boxes, or 18x10-pixel halo tiles shared by the 9 taps of a 3x3/s1 conv) and B tiles
    // (BN x 128 B); b_resident: every B tile of the layer is loaded once and stays in shared memory
    int n_a, a_slot_bytes, a_tx_bytes, n_b, b_slot_bytes, b_resident, halo;
    int halo_rev;        // data-gradient role: tap (ky, kx) reads the halo at (2-ky, 2-kx)
    int mcast;           // 1: launched as clusters of 2 CTAs (adjacent tiles, same weights): each CTA loads half of every weight tile and
                         // multicasts it to both, halving the L2 -> SM weight traffic that bounds the streamed-weight layers
    int b_group;         // streamed weights in halo mode: taps per B stage (3 = one kernel row per barrier round, 1 otherwise)
    // tile index -> (n-tile, tile x, tile y, image, class) without integer division: q = umulhi(x, mul) >> shr (d > 1)
    unsigned div_mul[4], div_shr[4];
    int per_class;
    const float* bias;   // [Cout] or null
    // thin epilogue (BN == 16): fp32 planar outputs, one plane pointer per channel (image 0), activation per channel
    int thin, thin_n, H, W;
    float* thin_ptr[16];
    long long thin_ns[16];   // elements between consecutive images of that plane
    int thin_act[16];        // 0 none, 1 LeakyReLU(0.2), 2 sigmoid
};

struct ConvGCfg {
    static const int MAX_STAGES = 8;
    static const int A_BYTES = 128 * 128;            // 128 positions x 64 channels bf16
    static const int RING_BYTES = 4 * (A_BYTES + 256 * 128);   // 192 KB of operand stages: 4 at BN = 256 ... 8 at BN <= 64
    static const int OUT_BYTES = 128 * 128;          // one 64-channel output group
    static const int HALO_W = 10, HALO_H = 18;        // halo of a 16-row x 8-column output tile
    static const int HALO_BYTES = HALO_W * HALO_H * 128;          // 23040
    static const int HALO_SLOT = 23 * 1024;
    static const int SMEM = 1024 + RING_BYTES + 2 * OUT_BYTES + 512;
    static const int THREADS = 224;                 // warp 0 TMA producer | warps 1-2 MMA issue | warps 3-6 epilogue
};

namespace tc {
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// weight tile multicast: the box lands at the same CTA-relative offset (and signals the same barrier offset) in every CTA of `mask`
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
// tcgen05.commit that arrives on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
}  // namespace tc

__device__ __forceinline__ void fast_divmod(int x, int d, unsigned mul, unsigned shr, int& q, int& r) {
    q = d == 1 ? x : (int)(__umulhi((unsigned)x, mul) >> shr);
    r = x - q * d;
}
// tile -> coordinates; the class index (transposed-conv parity classes) is the slowest and rare: one real division only there
#define CONVG_DECODE_TILE(tile, nt, tx, ty, n, cls)                                   \
    int nt, tx, ty, n, cls;                                                           \
    {                                                                                 \
        int r_ = (tile);                                                              \
        cls = p.n_classes == 1 ? 0 : r_ / p.per_class;                                \
        r_ -= cls * p.per_class;                                                      \
        fast_divmod(r_, p.n_tiles, p.div_mul[0], p.div_shr[0], r_, nt);               \
        fast_divmod(r_, p.tiles_x, p.div_mul[1], p.div_shr[1], r_, tx);               \
        fast_divmod(r_, p.tiles_y, p.div_mul[2], p.div_shr[2], n, ty);                \
    }

// Timing experiments (tools/convg_experiment.py, tools/convg_trace.py) exist only in the experiments build (-DPTTA_EXPERIMENTS ->
// lib/libptta_b200_experiments.so, never loaded by the package): switches that skip work (1 one MMA per K-item, 2 no epilogue work,
// 4 no epilogue fence / store, 8 / 16 no A / B loads -- results are then wrong) and cycle stamps (64).  In the product build `dbg` is the
// constant 0: every branch below folds away and the kernel reads no switch.
#ifdef PTTA_EXPERIMENTS
__device__ long long g_convg_ts[64 * 16];   // dbg & 64: clock64() stamps of CTA 0, [tile][event]
#define CONVG_TS(tl, k) do { if ((dbg & 64) && blockIdx.x == 0 && (tl) < 64) g_convg_ts[(tl) * 16 + (k)] = clock64(); } while (0)
__device__ unsigned long long g_convg_cta[256 * 2];   // dbg & 64: %globaltimer (ns) at entry / exit of every CTA
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ int g_convg_dbg = 0;
#define CONVG_DBG_LOAD() g_convg_dbg
#define CONVG_CTA_STAMP(slot) do { if ((dbg & 64) && tid == 0 && blockIdx.x < 256) g_convg_cta[blockIdx.x * 2 + (slot)] = globaltimer_ns(); } while (0)
#else
#define CONVG_TS(tl, k) do { } while (0)
#define CONVG_DBG_LOAD() 0
#define CONVG_CTA_STAMP(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(ConvGCfg::THREADS, 1)
convg_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
             const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_out,
             const __grid_constant__ ConvGParams p) {
    typedef ConvGCfg C;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t out_s = smem_base + C::RING_BYTES;
    const uint32_t bar_s = out_s + 2 * C::OUT_BYTES;
    const uint32_t a_full = bar_s, a_empty = bar_s + 64, b_full = bar_s + 128, b_empty = bar_s + 192, acc_full = bar_s + 256,
                   acc_empty = bar_s + 320, w_full = bar_s + 384;
    const uint32_t tmem_slot = bar_s + 392;
    // accumulator ring in TMEM: 512 columns / (64 | 128 | 256 columns per tile) = 8 | 4 | 2 buffers (tile t uses buffer t mod NACC):
    // with two issuing threads the epilogue's latency per tile has to be covered by more than one spare accumulator
    const uint32_t ACC_LOG = p.BN <= 64 ? 3u : (p.BN <= 128 ? 2u : 1u), NACC = 1u << ACC_LOG, ACC_STRIDE = 512u >> ACC_LOG;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dbg = CONVG_DBG_LOAD();
    if (tid == 0) CONVG_TS(63, 0);                   // kernel entry
    CONVG_CTA_STAMP(0);
    const uint32_t NA = p.n_a, NB = p.n_b, A_SLOT = p.a_slot_bytes, B_SLOT = p.b_slot_bytes;
    const uint32_t b_base = smem_base + NA * A_SLOT;            // B ring, or the resident weights
    const uint32_t b_tx = (uint32_t)p.BN * 128u;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_a0);
        tc::prefetch_tmap(&tmap_a1);
        tc::prefetch_tmap(&tmap_b);
        tc::prefetch_tmap(&tmap_out);
        for (int i = 0; i < C::MAX_STAGES; ++i) {
            tc::mbar_init(a_full + 8 * i, 1); tc::mbar_init(a_empty + 8 * i, 1);
            tc::mbar_init(b_full + 8 * i, 1); tc::mbar_init(b_empty + 8 * i, p.mcast ? 2 : 1);
        }
        for (int i = 0; i < 8; ++i) { tc::mbar_init(acc_full + 8 * i, 1); tc::mbar_init(acc_empty + 8 * i, 128); }
        tc::mbar_init(w_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    if (p.mcast) tc::cluster_sync();        // the peer's barriers are initialised before anything is multicast to them
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    PDL_SYNC();      // the set-up above (barriers, TMEM, descriptor prefetch) overlaps the previous kernel; nothing before this line touches global data
    if (tid == 0) CONVG_TS(63, 1);                   // set-up done (barriers, TMEM allocation)
    const uint32_t crank = p.mcast ? tc::cluster_ctarank() : 0u;
    // multicast mode: tile indices run to an even count; a ghost tile (image index == N) loads zeros and stores nothing (TMA clips)
    const int tiles_end = p.mcast ? ((p.total_tiles + 1) & ~1) : p.total_tiles;

    if (warp == 0) {
        // TMA producer: the warp stays converged, one elected lane issues (a divergent `if (lane == 0)` makes the compiler wrap
        // every UTMALDG / UTCHMMA in an ELECT + BRA.U.ANY loop)
        if (p.b_resident) {
            if (elect_one()) {
                int total = 0;
                for (int c = 0; c < p.n_classes; ++c) total += p.cls_count[c];
                tc::mbar_arrive_expect_tx(w_full, (uint32_t)total * b_tx);
                for (int i = 0; i < total; ++i) tc::tma_load_3d(b_base + i * B_SLOT, &tmap_b, w_full, 0, 0, p.items[i].wk);
            }
            __syncwarp();
        }
        // ring positions are kept as (slot, phase) pairs: no division in the per-item path.  ONE elected lane runs the whole producer
        // loop (waits included): no per-item warp synchronisation; the other lanes wait at the final barrier
        if (elect_one()) {
            uint32_t sa = 0, pa = 1, sb = 0, pb = 1;
            int tlp = 0;
            for (int tile = blockIdx.x; tile < tiles_end; tile += gridDim.x, ++tlp) {
                CONVG_TS(tlp, 8);
                CONVG_DECODE_TILE(tile, nt, tx, ty, n, cls)
                const int gx0 = tx * p.tw, gy0 = ty * p.th;
                const int i0 = p.cls_start[cls], cnt = p.cls_count[cls];
                if (p.halo) {
                    for (int i = 0; i < cnt; i += 9) {              // one halo tile per 64-channel chunk, then its nine weight tiles
                        const int c_inner = p.items[i0 + i].c_inner, src = (p.items[i0 + i].dyps >> 20) & 1;
                        tc::mbar_wait(a_empty + 8 * sa, pa);
                        if (dbg & 8) {
                            tc::mbar_arrive(a_full + 8 * sa);
                        } else {
                            tc::mbar_arrive_expect_tx(a_full + 8 * sa, (uint32_t)p.a_tx_bytes);
                            tc::tma_load_5d(smem_base + sa * A_SLOT, src ? &tmap_a1 : &tmap_a0, a_full + 8 * sa, c_inner, gx0 - 1, 0, gy0 - 1, n);
                        }
                        CONVG_TS(tlp, 9);
                        if (++sa == NA) { sa = 0; pa ^= 1; }
                        if (!p.b_resident) {
                            const int bg = p.b_group;                    // taps per stage: B_SLOT holds bg weight tiles
                            for (int tap = 0; tap < 9; tap += bg) {
                                tc::mbar_wait(b_empty + 8 * sb, pb);
                                if (dbg & 16) {
                                    tc::mbar_arrive(b_full + 8 * sb);
                                } else {
                                    tc::mbar_arrive_expect_tx(b_full + 8 * sb, b_tx * bg);
                                    for (int u = 0; u < bg; ++u) {
                                        const uint32_t dst = b_base + sb * B_SLOT + u * (B_SLOT / bg);
                                        if (p.mcast)     // my half of the rows, to both CTAs of the pair
                                            tc::tma_load_3d_mc(dst + crank * (b_tx >> 1), &tmap_b, b_full + 8 * sb, 0, nt * p.BN + crank * (p.BN >> 1),
                                                               i0 + i + tap + u, (uint16_t)3);
                                        else
                                            tc::tma_load_3d(dst, &tmap_b, b_full + 8 * sb, 0, nt * p.BN, i0 + i + tap + u);
                                    }
                                }
                                if (++sb == NB) { sb = 0; pb ^= 1; }
                            }
                        }
                    }
                } else {
                    for (int i = 0; i < cnt; ++i) {                 // one stage = A box + B tile, one barrier
                        const ConvGItem item = p.items[i0 + i];
                        tc::mbar_wait(a_empty + 8 * sa, pa);
                        tc::mbar_arrive_expect_tx(a_full + 8 * sa, (uint32_t)p.a_tx_bytes + b_tx);
                        const int dy = (int)(short)(item.dyps & 0xffff), py = (item.dyps >> 16) & 3, src = (item.dyps >> 20) & 1;
                        tc::tma_load_5d(smem_base + sa * A_SLOT, src ? &tmap_a1 : &tmap_a0, a_full + 8 * sa, item.c_inner, gx0 + item.dx, py,
                                        gy0 + dy, n);
                        tc::tma_load_3d(b_base + sa * B_SLOT, &tmap_b, a_full + 8 * sa, 0, nt * p.BN, i0 + i);
                        if (++sa == NA) { sa = 0; pa ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 2) {
        // MMA issuers (one elected lane each).  Layers with resident weights use BOTH: thread `me` takes the tiles t = me, me + 2, ...
        // (accumulator buffer t & 1 = me), so the ~750 cycles a thread spends per tile outside the MMA issue (barrier waits, fences,
        // commits, loop overhead: measured with tools/convg_trace.py) overlap with the other thread's 36 MMAs.
        const int nissue = (p.halo && p.b_resident) ? 2 : 1, me = warp - 1;
        const uint32_t idesc = tc::make_idesc_bf16(128, p.BN);
        // A: 8-row groups are 1024 B apart in a plain tile, HALO_W pixels (1280 B) apart in a halo tile (tile rows of 8 pixels);
        // SWIZZLE_128B is a function of the shared-memory address, so any 128 B-aligned start inside the halo is a valid operand
        const uint64_t da0 = make_desc_sw128_sbo(0, p.halo ? C::HALO_W * 128 : 1024), db0 = make_desc_sw128_sbo(0, 1024);
        const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32), lo0 = (uint32_t)da0;
        const int rev = p.halo_rev;
        // ONE elected lane runs the whole issue loop, waits included (no per-item elect / __syncwarp)
        if (me < nissue && elect_one()) {
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, t = me;
            if (p.b_resident) { tc::mbar_wait(w_full, 0); tc::tc_fence_after(); }
            const int skip = (nissue - 1) * (p.cls_count[0] / 9);      // A-ring slots the other issuer consumes between two of my tiles
            for (int k = 0; k < me * (p.cls_count[0] / 9); ++k) { if (++sa == NA) { sa = 0; pa ^= 1; } }
            for (int tile = blockIdx.x + me * gridDim.x; tile < tiles_end; tile += nissue * gridDim.x, t += nissue) {
                const int cls = p.n_classes == 1 ? 0 : tile / p.per_class;
                const int i0 = p.cls_start[cls], cnt = p.cls_count[cls];
                const uint32_t as = t & (NACC - 1);
                CONVG_TS(t, 0);
                tc::mbar_wait(acc_empty + 8 * as, ((t >> ACC_LOG) & 1) ^ 1);
                tc::tc_fence_after();
                CONVG_TS(t, 1);
                const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
                if (p.halo) {
                    for (int i = 0; i < cnt; i += 9) {
                        tc::mbar_wait(a_full + 8 * sa, pa);
                        tc::tc_fence_after();
                        CONVG_TS(t, 2);
                        const uint32_t a_tile = lo0 + ((smem_base + sa * A_SLOT) >> 4);
                        const uint32_t b_res = lo0 + ((b_base + (uint32_t)(i0 + i) * B_SLOT) >> 4);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            uint32_t b_lo = b_res + (uint32_t)tap * (B_SLOT >> 4);
                            if (!p.b_resident) {
                                const int bg = p.b_group, u = bg == 3 ? tap % 3 : 0;
                                if (u == 0) {
                                    if (i == 0 && tap % 3 == 0) CONVG_TS(t, 10 + 2 * (tap / 3));        // chunk 0: before the wait of taps 0 / 3 / 6
                                    tc::mbar_wait(b_full + 8 * sb, pb);
                                    tc::tc_fence_after();
                                    if (i == 0 && tap % 3 == 0) CONVG_TS(t, 11 + 2 * (tap / 3));        // ... and after it
                                }
                                b_lo = lo0 + ((b_base + sb * B_SLOT + u * (B_SLOT / bg)) >> 4);
                            }
                            const int hy = tap / 3, hx = tap - hy * 3;
                            const uint32_t off_f = (uint32_t)(hy * C::HALO_W + hx) * 8, off_r = (uint32_t)((2 - hy) * C::HALO_W + (2 - hx)) * 8;
                            const uint32_t a_lo = a_tile + (rev ? off_r : off_f);
                            if (tap == 0 && i == 0) tc::umma_f16_split<false>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc);
                            else tc::umma_f16_split<true>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc);
                            if (!(dbg & 1)) {
#pragma unroll
                                for (int k = 1; k < 4; ++k) tc::umma_f16_split<true>(d_tmem, a_lo + k * 2, a_hi, b_lo + k * 2, b_hi, idesc);
                            }
                            if (!p.b_resident && (p.b_group == 1 || tap % 3 == 2)) {
                                if (p.mcast) tc::umma_commit_mc(b_empty + 8 * sb, (uint16_t)3);     // both CTAs' producers wait for both consumers
                                else tc::umma_commit(b_empty + 8 * sb);
                                if (++sb == NB) { sb = 0; pb ^= 1; }
                            }
                        }
                        tc::umma_commit(a_empty + 8 * sa);
                        if (i + 9 >= cnt) tc::umma_commit(acc_full + 8 * as);
                        CONVG_TS(t, 3);
                        if (++sa == NA) { sa = 0; pa ^= 1; }
                    }
                    for (int k = 0; k < skip; ++k) { if (++sa == NA) { sa = 0; pa ^= 1; } }
                } else {
                    for (int i = 0; i < cnt; ++i) {
                        tc::mbar_wait(a_full + 8 * sa, pa);
                        tc::tc_fence_after();
                        const uint32_t a_lo = lo0 + ((smem_base + sa * A_SLOT) >> 4);
                        const uint32_t b_lo = lo0 + ((b_base + sa * B_SLOT) >> 4);
                        if (i == 0) tc::umma_f16_split<false>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc);
                        else tc::umma_f16_split<true>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc);
                        if (!(dbg & 1)) {
#pragma unroll
                            for (int k = 1; k < 4; ++k) tc::umma_f16_split<true>(d_tmem, a_lo + k * 2, a_hi, b_lo + k * 2, b_hi, idesc);
                        }
                        tc::umma_commit(a_empty + 8 * sa);
                        if (i == cnt - 1) tc::umma_commit(acc_full + 8 * as);
                        if (++sa == NA) { sa = 0; pa ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3, et = (warp - 3) * 32 + lane;     // et: 0..127, thread 0 issues the stores
        const int row = q * 32 + lane;                           // position inside the tile == TMEM lane
        const int groups = p.BN >> 6;      // 0 in thin mode
        uint32_t t = 0, sg = 0;
        for (int tile = blockIdx.x; tile < tiles_end; tile += gridDim.x, ++t) {
            CONVG_DECODE_TILE(tile, nt, tx, ty, n, cls)
            const uint32_t as = t & (NACC - 1);
            if (et == 0) CONVG_TS(t, 4);
            tc::mbar_wait(acc_full + 8 * as, (t >> ACC_LOG) & 1);
            tc::tc_fence_after();
            if (et == 0) CONVG_TS(t, 5);
            if (p.thin) {
                uint32_t v[32];
                tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE, v);
                const int gy = ty * p.th + row / p.tw, gx = tx * p.tw + row % p.tw;
                if (gy < p.H && gx < p.W) {
                    const long long pix = (long long)gy * p.W + gx;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        if (c < p.thin_n) {
                            float f = __uint_as_float(v[c]) + (p.bias ? __ldg(p.bias + c) : 0.f);
                            const int a = p.thin_act[c];
                            f = a == 1 ? (f > 0.f ? f : 0.2f * f) : (a == 2 ? 1.f / (1.f + __expf(-f)) : f);
                            p.thin_ptr[c][(long long)n * p.thin_ns[c] + pix] = f;
                        }
                    }
                }
            }
            for (int g = 0; g < ((dbg & 2) ? 0 : groups); ++g, ++sg) {
                const uint32_t buf = out_s + (sg & 1) * C::OUT_BYTES;
                if (et == 0) tc::bulk_wait_read<1>();            // the store that used this buffer two groups ago has read it
                tc::epi_bar();
                const int ch0 = nt * p.BN + g * 64;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE + g * 64 + h * 32, v);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[c * 8 + j]) + (p.bias ? __ldg(p.bias + ch0 + h * 32 + c * 8 + j) : 0.f);
                        uint4 ov;
                        ov.x = pack_bf162(f[0], f[1]); ov.y = pack_bf162(f[2], f[3]);
                        ov.z = pack_bf162(f[4], f[5]); ov.w = pack_bf162(f[6], f[7]);
                        const int chunk = h * 4 + c;
                        *reinterpret_cast<uint4*>(smem + (buf - smem_base) + row * 128 + ((chunk ^ (row & 7)) << 4)) = ov;
                    }
                }
                if (!(dbg & 4)) tc::fence_proxy_async();
                tc::epi_bar();
                if (et == 0 && g == 0) CONVG_TS(t, 6);
                if (et == 0 && !(dbg & 4)) {
                    tc::tma_store_5d(&tmap_out, buf, p.cls_out_c[cls] + ch0, tx * p.tw, p.cls_out_py[cls], ty * p.th, n);
                    tc::bulk_commit();
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(acc_empty + 8 * as);
            if (et == 0) CONVG_TS(t, 7);
        }
        if (et == 0) tc::bulk_wait_all();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) CONVG_TS(63, 2);                   // all roles finished (stores drained)
    CONVG_CTA_STAMP(1);
    if (p.mcast) tc::cluster_sync();        // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ---- weight packing: Wpk[wk][n][j] = W[n * sn + (k0 + j) * sk + tap] (0 outside the real extents) ---------------------------
struct ConvGPackEntry { int wsel_tap; int k0; };      // wsel_tap = which weight tensor (bit 8) | tap index (low 8 bits)
struct ConvGPackParams {
    ConvGPackEntry e[128];
    const float* w[2];
    long long sn[2], sk[2];
    int n_real[2], k_real[2];
    int n_pad;            // rows per chunk in the packed buffer
    int ident_from;       // >= 0: rows/cols [ident_from, n_pad) of the centre tap (tap 4) of tensor 0 form an identity
};

__global__ void __launch_bounds__(256) pack_convg_kernel(const __grid_constant__ ConvGPackParams p, bf16* __restrict__ out) {
    PDL_SYNC();
    // grid (chunk, 16-row group): the adapted meta conv is re-packed after every Adam step, so this sits on the step's critical path
    const int wk = blockIdx.x;
    const ConvGPackEntry e = p.e[wk];
    const int sel = (e.wsel_tap >> 8) & 1, tap = e.wsel_tap & 255;
    const float* w = p.w[sel];
    const int row0 = blockIdx.y * 16, rows = min(16, p.n_pad - row0);
    for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) {
        const int n = row0 + (i >> 6), j = i & 63, k = e.k0 + j;
        float v = 0.f;
        if (n < p.n_real[sel] && k < p.k_real[sel]) v = w[(long long)n * p.sn[sel] + (long long)k * p.sk[sel] + tap];
        else if (p.ident_from >= 0 && sel == 0 && tap == 4 && n >= p.ident_from && n == k) v = 1.f;
        out[((size_t)wk * p.n_pad + n) * 64 + j] = __float2bfloat16(v);
    }
}

inline int& convg_multicast_flag() { static int f = 0; return f; }
inline bool convg_multicast_enabled() { return convg_multicast_flag() != 0; }

// ---- host side: plans ------------------------------------------------------------------------------------------------------
enum { CONVG_S1 = 0, CONVG_S2 = 1, CONVG_T2 = 2, CONVG_P1S2 = 3 };

struct ConvGPlan {
    ConvGParams p;
    ConvGPackEntry pack[128];
    int n_items;          // == number of K-chunks in the packed weights
    int in_parity;        // source view: 1 plain, 2 pixel-parity
    int out_parity;       // output view
    int k_src[2];         // stored channels of source 0 / 1 (0 = absent)
    int n_out;            // stored output channels
    int grid_h, grid_w;   // tile grid extents (output positions per class)
    int in_h, in_w, out_h, out_w;
};

// kind/role/shapes -> item list.  h, w: spatial size of the LAYER's input (forward sense); cin0 + cin1 = layer input channels
// (two sources only in the forward role), cout = layer output channels; all stored channel counts are multiples of 64.
// has_short: role 1 of CONVG_S2 additionally folds the data gradient of a 1x1/s2 shortcut conv (second source = its dY).
inline int convg_make_plan(ConvGPlan& pl, int kind, int role, int n, int h, int w, int cin0, int cin1, int cout, int has_short) {
    PTTA_CHECK(kind >= 0 && kind <= 3 && (role == 0 || role == 1), "convg: bad kind %d / role %d", kind, role);
    const bool thin = cout == 16 && kind == CONVG_S1 && role == 0;      // thin head: 16 output channels, fp32 planar epilogue
    PTTA_CHECK(cin0 > 0 && cin0 % 64 == 0 && cin1 % 64 == 0 && cout > 0 && (cout % 64 == 0 || thin), "convg: channels must be multiples of 64 (%d+%d -> %d)", cin0, cin1, cout);
    PTTA_CHECK(role == 0 || cin1 == 0, "convg: two sources only in the forward role");
    PTTA_CHECK(!has_short || (kind == CONVG_S2 && role == 1), "convg: shortcut folding only for the data gradient of a stride-2 conv");
    PTTA_CHECK(kind == CONVG_S1 || kind == CONVG_T2 || (h % 2 == 0 && w % 2 == 0), "convg: stride-2 layers need even H, W (%d x %d)", h, w);
    PTTA_CHECK(!(kind == CONVG_P1S2 && role == 1), "convg: the 1x1/s2 data gradient only exists folded into its block's 3x3/s2 one (has_short)");
    memset(&pl, 0, sizeof(pl));
    const int cin = cin0 + cin1;
    // the GEMM's K sources and N extent
    int ksrc[2] = {cin0, cin1}, nout = cout;
    if (role == 1) { ksrc[0] = cout; ksrc[1] = has_short ? cout : 0; nout = cin; }
    pl.k_src[0] = ksrc[0]; pl.k_src[1] = ksrc[1]; pl.n_out = nout;
    // geometry family
    //   gather-s1 : plain in, plain out, same grid                       (S1 fwd, S1 dgrad)
    //   gather-s2 : parity in, plain out, grid = in/2                    (S2 fwd, P1S2 fwd, T2 dgrad)
    //   scatter-s2: plain in, parity out (4 classes), grid = in          (T2 fwd, S2 dgrad [+ P1S2 dgrad])
    int family;
    if (kind == CONVG_S1) family = 0;
    else if ((kind == CONVG_T2) == (role == 0)) family = 2;
    else family = 1;
    // spatial sizes of the GEMM's source / destination tensors
    int lin_h = h, lin_w = w, lout_h, lout_w;              // layer input / output
    if (kind == CONVG_S1) { lout_h = h; lout_w = w; }
    else if (kind == CONVG_T2) { lout_h = 2 * h; lout_w = 2 * w; }
    else { lout_h = h / 2; lout_w = w / 2; }
    pl.in_h = role == 0 ? lin_h : lout_h; pl.in_w = role == 0 ? lin_w : lout_w;
    pl.out_h = role == 0 ? lout_h : lin_h; pl.out_w = role == 0 ? lout_w : lin_w;
    pl.in_parity = family == 1 ? 2 : 1;
    pl.out_parity = family == 2 ? 2 : 1;
    pl.grid_h = family == 1 ? pl.in_h / 2 : pl.in_h;
    pl.grid_w = family == 1 ? pl.in_w / 2 : pl.in_w;
    const int ntap = kind == CONVG_P1S2 ? 1 : 9;
    int ni = 0;
    int overflow = 0;
    auto add_item = [&](int src, int csrc_total, int c, int px, int dx, int py, int dy, int wsel, int tap, int k0, int new_a, int last_a,
                        int a_off16) {
        ConvGItem& it = pl.p.items[ni];
        it.c_inner = px * csrc_total + c;
        it.dx = dx;
        it.dyps = (dy & 0xffff) | (py << 16) | (src << 20) | (new_a << 24) | (last_a << 25);
        it.wk = ni;
        it.a_off16 = a_off16;
        pl.pack[ni].wsel_tap = (wsel << 8) | tap;
        pl.pack[ni].k0 = k0;
        ++ni;
    };
    // one tap of one source: all its 64-channel chunks, each with its own A box
    auto add_tap = [&](int src, int px, int dx, int py, int dy, int wsel, int tap, int kbase) -> int {
        for (int c = 0; c < ksrc[src]; c += 64) {
            if (ni >= 128) return 1;
            add_item(src, ksrc[src], c, px, dx, py, dy, wsel, tap, kbase + c, 1, 1, 0);
        }
        return 0;
    };
    if (family == 0) {
        // halo mode: ONE 18x10-pixel A tile per 64-channel chunk serves the nine taps (descriptor start offsets into the halo)
        pl.p.n_classes = 1;
        pl.p.cls_start[0] = 0;
        pl.p.halo = 1;
        pl.p.halo_rev = role;
        for (int src = 0; src < 2; ++src)
            for (int c = 0; c < ksrc[src]; c += 64)
                for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx) {
                        if (ni >= 128) { overflow = 1; break; }
                        const int hy = role == 0 ? ky : 2 - ky, hx = role == 0 ? kx : 2 - kx;     // position of the tap inside the halo
                        const int tap = ky * 3 + kx;
                        add_item(src, ksrc[src], c, 0, -1, 0, -1, 0, tap, (src ? cin0 : 0) + c, tap == 0, tap == 8,
                                 (hy * ConvGCfg::HALO_W + hx) * 8);
                    }
        pl.p.cls_count[0] = ni;
    } else if (family == 1) {
        pl.p.n_classes = 1;
        pl.p.cls_start[0] = 0;
        if (ntap == 1) {
            overflow |= add_tap(0, 0, 0, 0, 0, 0, 0, 0);
        } else {
            // input row 2*oy + ky - 1:  ky 0 -> parity 1 of pair oy-1, ky 1 -> parity 0 of pair oy, ky 2 -> parity 1 of pair oy
            const int par[3] = {1, 0, 1}, off[3] = {-1, 0, 0};
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    overflow |= add_tap(0, par[kx], off[kx], par[ky], off[ky], 0, ky * 3 + kx, 0);
                    if (ksrc[1]) overflow |= add_tap(1, par[kx], off[kx], par[ky], off[ky], 0, ky * 3 + kx, cin0);
                }
        }
        pl.p.cls_count[0] = ni;
    } else {
        // output position 2*g + q receives: q = 0: tap 1 from g;  q = 1: tap 0 from g+1 and tap 2 from g
        pl.p.n_classes = 4;
        for (int qy = 0; qy < 2; ++qy)
            for (int qx = 0; qx < 2; ++qx) {
                const int cls = qy * 2 + qx;
                pl.p.cls_start[cls] = ni;
                pl.p.cls_out_c[cls] = qx * nout;
                pl.p.cls_out_py[cls] = qy;
                const int nky = qy ? 2 : 1, nkx = qx ? 2 : 1;
                const int kys[2] = {qy ? 0 : 1, 2}, dys[2] = {qy ? 1 : 0, 0};
                const int kxs[2] = {qx ? 0 : 1, 2}, dxs[2] = {qx ? 1 : 0, 0};
                for (int a = 0; a < nky; ++a)
                    for (int b = 0; b < nkx; ++b) {
                        overflow |= add_tap(0, 0, dxs[b], 0, dys[a], 0, kys[a] * 3 + kxs[b], 0);
                        if (ksrc[1] && !has_short) overflow |= add_tap(1, 0, dxs[b], 0, dys[a], 0, kys[a] * 3 + kxs[b], cin0);
                    }
                if (has_short && cls == 0) overflow |= add_tap(1, 0, 0, 0, 0, 1, 0, 0);
                pl.p.cls_count[cls] = ni - pl.p.cls_start[cls];
            }
    }
    PTTA_CHECK(!overflow, "convg: more than 128 K-items (kind %d role %d, %d+%d -> %d)", kind, role, cin0, cin1, cout);
    pl.n_items = ni;
    // tiles
    int BN = 16;                       // largest multiple of 64 that is <= 256 and divides the output channels
    if (!thin) for (BN = 256; nout % BN; BN -= 64) {}
    pl.p.thin = thin ? 1 : 0;
    static const int shapes[4][2] = {{8, 16}, {4, 32}, {2, 64}, {1, 128}};
    long long best = -1; int bi = 0;
    for (int i = 0; i < 4; ++i) {
        long long tiles = (long long)cdiv(pl.grid_h, shapes[i][0]) * cdiv(pl.grid_w, shapes[i][1]);
        if (best < 0 || tiles < best) { best = tiles; bi = i; }
    }
    pl.p.th = shapes[bi][0]; pl.p.tw = shapes[bi][1];
    if (pl.p.halo) { pl.p.th = 16; pl.p.tw = 8; }
    pl.p.tiles_y = cdiv(pl.grid_h, pl.p.th); pl.p.tiles_x = cdiv(pl.grid_w, pl.p.tw);
    pl.p.N = n; pl.p.BN = BN; pl.p.n_tiles = nout / BN;
    pl.p.total_tiles = pl.p.n_classes * n * pl.p.tiles_y * pl.p.tiles_x * pl.p.n_tiles;
    pl.p.per_class = n * pl.p.tiles_y * pl.p.tiles_x * pl.p.n_tiles;
    {
        const int divs[4] = {pl.p.n_tiles, pl.p.tiles_x, pl.p.tiles_y, n};
        for (int i = 0; i < 4; ++i) {
            const unsigned d = (unsigned)divs[i];
            unsigned lg = 0;
            while ((1u << lg) < d) ++lg;                 // ceil(log2 d)
            const unsigned pw = 31 + lg;
            pl.p.div_mul[i] = d <= 1 ? 0u : (unsigned)(((1ull << pw) + d - 1) / d);
            pl.p.div_shr[i] = d <= 1 ? 0u : pw - 32;
        }
    }
    // operand rings
    const int b_tile = ((BN * 128 + 1023) / 1024) * 1024;
    pl.p.b_slot_bytes = b_tile;
    pl.p.b_group = 1;
    if (pl.p.halo) {
        pl.p.a_slot_bytes = ConvGCfg::HALO_SLOT; pl.p.a_tx_bytes = ConvGCfg::HALO_BYTES;
        if (pl.p.n_tiles == 1 && ni * b_tile <= 96 * 1024) {
            pl.p.b_resident = 1; pl.p.n_b = 0;
            pl.p.n_a = (ConvGCfg::RING_BYTES - ni * b_tile) / ConvGCfg::HALO_SLOT;
        } else if (2 * (3 * b_tile) + 2 * ConvGCfg::HALO_SLOT <= ConvGCfg::RING_BYTES) {
            // one kernel row (3 taps) of weights per stage: one barrier round per 12 MMAs instead of per 4
            pl.p.b_group = 3;
            pl.p.b_slot_bytes = 3 * b_tile;
            pl.p.n_a = 2;
            pl.p.n_b = (ConvGCfg::RING_BYTES - pl.p.n_a * ConvGCfg::HALO_SLOT) / (3 * b_tile);
        } else {
            pl.p.n_a = BN >= 256 ? 2 : 3;
            pl.p.n_b = (ConvGCfg::RING_BYTES - pl.p.n_a * ConvGCfg::HALO_SLOT) / b_tile;
        }
    } else {
        pl.p.a_slot_bytes = pl.p.a_tx_bytes = ConvGCfg::A_BYTES;
        pl.p.n_a = pl.p.n_b = ConvGCfg::RING_BYTES / (ConvGCfg::A_BYTES + b_tile);
    }
    // weight-tile multicast over clusters of two CTAs: implemented and parity-tested, but it does not shorten the streamed-weight layers
    // (128->128: 29.3 us with and without: the per-tile trace shows them MMA-bound at ~75 cycles per N = 128 MMA with 8 us of
    // prologue / tail per launch, not L2-bound), so it stays a tested option, off by default: convg_multicast_enabled() <- ptta_convg_debug_set(32)
    pl.p.mcast = (convg_multicast_enabled() && pl.p.halo && !pl.p.b_resident && pl.p.n_tiles == 1 && BN % 16 == 0 && pl.p.total_tiles >= 2) ? 1 : 0;
    if (pl.p.n_a > ConvGCfg::MAX_STAGES) pl.p.n_a = ConvGCfg::MAX_STAGES;
    if (pl.p.n_b > ConvGCfg::MAX_STAGES) pl.p.n_b = ConvGCfg::MAX_STAGES;
    return 0;
}

// 5-D view {P*C, W/P, P, H/P, N} of an NHWC bf16 tensor, box {64, tw, 1, th, 1}, SWIZZLE_128B
inline int make_tmap_view5(CUtensorMap* map, const void* ptr, int N, int H, int W, int Cc, int P, int th, int tw) {
    PFN_encodeTiled enc = get_encode_tiled();
    PTTA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[5] = {(cuuint64_t)P * Cc, (cuuint64_t)(W / P), (cuuint64_t)P, (cuuint64_t)(H / P), (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)P * Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)P * W * Cc * 2, (cuuint64_t)H * W * Cc * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)tw, 1, (cuuint32_t)th, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTTA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(view5) failed with %d (N=%d H=%d W=%d C=%d P=%d box %dx%d)", (int)r, N, H, W, Cc, P, th, tw);
    return 0;
}

// packed weights [chunks][n_pad][64] bf16, box {64, BN, 1}
inline int make_tmap_wpk(CUtensorMap* map, const void* ptr, int chunks, int n_pad, int BN) {
    PFN_encodeTiled enc = get_encode_tiled();
    PTTA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {64, (cuuint64_t)n_pad, (cuuint64_t)chunks};
    cuuint64_t strides[2] = {128, (cuuint64_t)n_pad * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTTA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wpk) failed with %d (chunks=%d n=%d BN=%d)", (int)r, chunks, n_pad, BN);
    return 0;
}

inline long long convg_packed_elems(const ConvGPlan& pl) { return (long long)pl.n_items * pl.n_out * 64; }

// w: the layer's weight (Conv2d [cout_w][cin_w][k][k] or ConvTranspose2d [cin_w][cout_w][3][3]); w_short: the 1x1 shortcut [cout_w][cin_w]
inline int launch_convg_pack(const ConvGPlan& pl, int kind, int role, const float* w, const float* w_short, int cin_w, int cout_w,
                             int ident_from, bf16* packed, cudaStream_t st) {
    ConvGPackParams pp;
    memset(&pp, 0, sizeof(pp));
    memcpy(pp.e, pl.pack, sizeof(pp.e));
    const int T = kind == CONVG_P1S2 ? 1 : 9;
    pp.w[0] = w; pp.w[1] = w_short;
    const bool conv_layout = kind != CONVG_T2;     // [cout][cin][T] vs [cin][cout][T]
    if (role == 0) {          // n = cout, k = cin
        pp.sn[0] = conv_layout ? (long long)cin_w * T : T;
        pp.sk[0] = conv_layout ? T : (long long)cout_w * T;
        pp.n_real[0] = cout_w; pp.k_real[0] = cin_w;
    } else {                  // n = cin, k = cout
        pp.sn[0] = conv_layout ? T : (long long)cout_w * T;
        pp.sk[0] = conv_layout ? (long long)cin_w * T : T;
        pp.n_real[0] = cin_w; pp.k_real[0] = cout_w;
    }
    // shortcut (role 1 only): Conv2d [cout][cin][1][1] -> n = cin, k = cout
    pp.sn[1] = 1; pp.sk[1] = cin_w; pp.n_real[1] = cin_w; pp.k_real[1] = cout_w;
    pp.n_pad = pl.n_out;
    pp.ident_from = ident_from;
    launch_k(pack_convg_kernel, dim3(pl.n_items, cdiv(pl.n_out, 16)), 256, 0, st, pp, packed);
    return check_launch("pack_convg");
}

inline int launch_convg(const ConvGPlan& pl, const bf16* x0, const bf16* x1, const bf16* packed, const float* bias, bf16* out, cudaStream_t st) {
    typedef ConvGCfg C;
    static int sms = 0;
    if (!sms) {
        PTTA_CUDA(cudaFuncSetAttribute(convg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int dev = 0;
        PTTA_CUDA(cudaGetDevice(&dev));
        PTTA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    PTTA_CHECK(x0 && packed && (out || pl.p.thin) && (x1 || !pl.k_src[1]), "convg: null operand");
    CUtensorMap ta0, ta1, tb, to;
    const ConvGParams& p = pl.p;
    const int ath = p.halo ? ConvGCfg::HALO_H : p.th, atw = p.halo ? ConvGCfg::HALO_W : p.tw;
    PTTA_TRY(make_tmap_view5(&ta0, x0, p.N, pl.in_h, pl.in_w, pl.k_src[0], pl.in_parity, ath, atw));
    if (pl.k_src[1]) PTTA_TRY(make_tmap_view5(&ta1, x1, p.N, pl.in_h, pl.in_w, pl.k_src[1], pl.in_parity, ath, atw));
    else ta1 = ta0;
    PTTA_TRY(make_tmap_wpk(&tb, packed, pl.n_items, pl.n_out, p.mcast ? p.BN / 2 : p.BN));
    if (p.thin) to = ta0;
    else PTTA_TRY(make_tmap_view5(&to, out, p.N, pl.out_h, pl.out_w, pl.n_out, pl.out_parity, p.th, p.tw));
    ConvGParams pr = p;
    pr.bias = bias;
    pr.H = pl.out_h; pr.W = pl.out_w;
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    if (p.mcast) {
        int g2 = grid & ~1;                      // clusters of two CTAs
        if (g2 < 2) g2 = 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(g2); cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PTTA_CUDA(cudaLaunchKernelEx(&cfg, convg_kernel, ta0, ta1, tb, to, pr));
        return check_launch("convg(mc)");
    }
    launch_k(convg_kernel, grid, C::THREADS, C::SMEM, st, ta0, ta1, tb, to, pr);
    return check_launch("convg");
}

}  // namespace ptta
