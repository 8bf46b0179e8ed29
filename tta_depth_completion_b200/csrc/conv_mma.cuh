// Generic NHWC bf16 3x3 implicit-GEMM convolution on the legacy tensor path (mma.sync m16n8k16,
// fp32 accumulate).  One kernel template covers the three geometries MSG-CHN needs -- and, by
// choice of the packed weight, both the forward and the data-gradient role of each:
//
//   MODE_S1  3x3 stride 1 pad 1            fwd of Conv2d(s1)          | dgrad of Conv2d(s1) (flipped taps)
//   MODE_S2  3x3 stride 2 pad 1            fwd of Conv2d(s2)          | dgrad of ConvTranspose2d(s2)
//   MODE_T2  3x3 transposed s2 p1 op1      fwd of ConvTranspose2d(s2) | dgrad of Conv2d(s2)
//
// (reference layers: external_src/MSG_CHN/workspace/exp_msg_chn/network_exp_msg_chn_adapt.py:166-311).
// Weights arrive pre-packed as bf16 [tap][COUT][CIN] (see pack_conv_weight in small_kernels.cuh).
// Fusions: prologue on the input (ReLU, or BatchNorm-affine + LeakyReLU) applied while the halo
// tile is staged in shared memory; epilogue bias, activation-derivative mask (ReLU mask of a saved
// pre-activation, or LeakyReLU'(BN(x))) and accumulate-into / add-from a second tensor.
//
// This is the general-shape kernel of the engine (strided / transposed / 128-channel layers and
// every data-gradient); the dominant 32->32 stride-1 forward layers have a tcgen05 kernel of their own.
#pragma once
#include "common.cuh"

namespace ptta {

enum { MODE_S1 = 0, MODE_S2 = 1, MODE_T2 = 2 };
enum { PRO_NONE = 0, PRO_RELU = 1, PRO_BN_LEAKY = 2 };
enum { MASK_NONE = 0, MASK_RELU = 1, MASK_BN_LEAKY = 2 };

struct ConvParams {
    const bf16* in;
    bf16* out;
    const bf16* w;       // [9][COUT][CIN]
    const float* bias;   // [COUT] or null
    int N, Hin, Win, Hout, Wout;
    int pro;             // PRO_*
    const float* pro_scale;
    const float* pro_shift;
    float slope;         // LeakyReLU slope for PRO_BN_LEAKY / MASK_BN_LEAKY
    const bf16* mask;    // [N,Hout,Wout,COUT] or null
    int mask_mode;       // MASK_*
    const float* mask_scale;
    const float* mask_shift;
    const bf16* add;     // [N,Hout,Wout,COUT] or null; out = add + mask*(acc + bias)
    int relu_out;        // store ReLU(out) (outputs that are only ever consumed through a ReLU)
    bf16* out2;          // optional second output: ReLU(out [+ add2])
    const bf16* add2;    // optional addend of out2 only
    int tiles_x, tiles_y;
};

template <int CIN>
__device__ __forceinline__ int swz_off(int row, int chunk) {
    if (CIN == 32) return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
    return row * (CIN * 2) + ((chunk ^ (row & 7)) << 4);
}

template <int MODE> struct ConvTile;
template <> struct ConvTile<MODE_S1> { static const int TH = 16, TW = 16, HH = 18, HW = 18; };
template <> struct ConvTile<MODE_S2> { static const int TH = 16, TW = 16, HH = 33, HW = 33; };
template <> struct ConvTile<MODE_T2> { static const int TH = 8, TW = 16, HH = 9, HW = 17; };   // tile over the INPUT

template <int CIN, int COUT, int MODE>
struct ConvSmem {
    static const int HALO_BYTES = ((ConvTile<MODE>::HH * ConvTile<MODE>::HW * CIN * 2 + 127) / 128) * 128;
    static const int W_BYTES = 9 * COUT * CIN * 2;
    static const int AUX_FLOATS = 2 * CIN + 3 * COUT;
    static const int TOTAL = HALO_BYTES + W_BYTES + AUX_FLOATS * 4;
};

template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(256) conv3x3_mma_kernel(const ConvParams p) {
    PDL_SYNC();
    typedef ConvTile<MODE> T;
    typedef ConvSmem<CIN, COUT, MODE> S;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_halo = smem;
    unsigned char* s_w = smem + S::HALO_BYTES;
    float* s_aux = reinterpret_cast<float*>(smem + S::HALO_BYTES + S::W_BYTES);
    float* s_pro_scale = s_aux;
    float* s_pro_shift = s_aux + CIN;
    float* s_bias = s_aux + 2 * CIN;
    float* s_mscale = s_bias + COUT;
    float* s_mshift = s_mscale + COUT;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.y;
    const int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
    const int t_y0 = ty * T::TH, t_x0 = tx * T::TW;     // tile origin (output coords for S1/S2, input for T2)

    // ---- stage weights (cp.async) and small per-channel vectors -----------------------------------
    {
        const int CH = CIN / 8;
        const int total = 9 * COUT * CH;
        const uint32_t wb = smem_u32(s_w);
        for (int i = tid; i < total; i += 256) {
            int row = i / CH, c = i - row * CH;
            cp_async16(wb + swz_off<CIN>(row, c), p.w + (size_t)row * CIN + c * 8, true);
        }
        cp_async_commit();
        for (int i = tid; i < CIN; i += 256) {
            s_pro_scale[i] = p.pro == PRO_BN_LEAKY ? p.pro_scale[i] : 1.f;
            s_pro_shift[i] = p.pro == PRO_BN_LEAKY ? p.pro_shift[i] : 0.f;
        }
        for (int i = tid; i < COUT; i += 256) {
            s_bias[i] = p.bias ? p.bias[i] : 0.f;
            s_mscale[i] = p.mask_mode == MASK_BN_LEAKY ? p.mask_scale[i] : 1.f;
            s_mshift[i] = p.mask_mode == MASK_BN_LEAKY ? p.mask_shift[i] : 0.f;
        }
    }
    __syncthreads();   // aux vectors visible before the halo prologue uses them

    // ---- stage the input halo tile with the fused prologue ---------------------------------------
    {
        const int CH = CIN / 8;
        const int total = T::HH * T::HW * CH;
        int gy0, gx0;
        if (MODE == MODE_S1) { gy0 = t_y0 - 1; gx0 = t_x0 - 1; }
        else if (MODE == MODE_S2) { gy0 = 2 * t_y0 - 1; gx0 = 2 * t_x0 - 1; }
        else { gy0 = t_y0; gx0 = t_x0; }
        const bf16* in_n = p.in + (size_t)n * p.Hin * p.Win * CIN;
        for (int i = tid; i < total; i += 256) {
            int pix = i / CH, c = i - pix * CH;
            int hy = pix / T::HW, hx = pix - hy * T::HW;
            int gy = gy0 + hy, gx = gx0 + hx;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (gy >= 0 && gy < p.Hin && gx >= 0 && gx < p.Win) {
                v = __ldg(reinterpret_cast<const uint4*>(in_n + ((size_t)gy * p.Win + gx) * CIN + c * 8));
                if (p.pro == PRO_RELU) {
                    const bf162 z = __floats2bfloat162_rn(0.f, 0.f);
                    bf162* h = reinterpret_cast<bf162*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) h[j] = __hmax2(h[j], z);
                } else if (p.pro == PRO_BN_LEAKY) {
                    uint32_t* u = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = unpack_bf162(u[j]);
                        int ch = c * 8 + j * 2;
                        f.x = f.x * s_pro_scale[ch] + s_pro_shift[ch];
                        f.y = f.y * s_pro_scale[ch + 1] + s_pro_shift[ch + 1];
                        f.x = f.x > 0.f ? f.x : f.x * p.slope;
                        f.y = f.y > 0.f ? f.y : f.y * p.slope;
                        u[j] = pack_bf162(f.x, f.y);
                    }
                }
            }
            *reinterpret_cast<uint4*>(s_halo + swz_off<CIN>(pix, c)) = v;
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    const uint32_t hb = smem_u32(s_halo), wb = smem_u32(s_w);
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;   // pixel (row of the 16xk A tile) this lane addresses
    const int a_kc = lane >> 4;                             // which 8-wide k chunk of the k16 step
    const int b_row = (lane & 7) + (lane >> 4) * 8;         // cout (row of the [n][k] weight tile)
    const int b_kc = (lane >> 3) & 1;
    const int c_row = lane >> 2, c_col = (lane & 3) * 2;    // accumulator fragment coordinates

    const size_t out_n = (size_t)n * p.Hout * p.Wout * COUT;

#pragma unroll 1
    for (int nc = 0; nc < COUT / 32; ++nc) {
        float acc[4][4][4];   // [slot][ntile][frag]; slot = m-tile (S1/S2: 2 used) or output parity class (T2: 4)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

        if (MODE != MODE_T2) {
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
                for (int ks = 0; ks < CIN / 16; ++ks) {
                    uint32_t b[8];
                    const int wrow = tap * COUT + nc * 32 + b_row;
                    ldmatrix_x4(b[0], b[1], b[2], b[3], wb + swz_off<CIN>(wrow, ks * 2 + b_kc));
                    ldmatrix_x4(b[4], b[5], b[6], b[7], wb + swz_off<CIN>(wrow + 16, ks * 2 + b_kc));
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        int prow;
                        if (MODE == MODE_S1) prow = (2 * warp + mt + ky) * T::HW + (a_row + kx);
                        else prow = (2 * (2 * warp + mt) + ky) * T::HW + (2 * a_row + kx);
                        uint32_t a[4];
                        ldmatrix_x4(a[0], a[1], a[2], a[3], hb + swz_off<CIN>(prow, ks * 2 + a_kc));
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], a, b[nt * 2], b[nt * 2 + 1]);
                    }
                }
            }
        } else {
            // transposed stride 2: input pixel (i,j) of this warp's row feeds the four output parities
            //   (2i  ,2j  ) += in(i  ,j  ) W11
            //   (2i  ,2j+1) += in(i  ,j+1) W10 + in(i,j) W12
            //   (2i+1,2j  ) += in(i+1,j  ) W01 + in(i,j) W21
            //   (2i+1,2j+1) += in(i+1,j+1) W00 + in(i+1,j) W02 + in(i,j+1) W20 + in(i,j) W22
            // combos: {class, dy, dx, tap}
            const int combo_cls[9] = {0, 1, 1, 2, 2, 3, 3, 3, 3};
            const int combo_a[9] = {0, 1, 0, 2, 0, 3, 2, 1, 0};      // A index = dy*2+dx
            const int combo_tap[9] = {4, 3, 5, 1, 7, 0, 2, 6, 8};
#pragma unroll
            for (int ks = 0; ks < CIN / 16; ++ks) {
                uint32_t a[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int prow = (warp + (q >> 1)) * T::HW + (a_row + (q & 1));
                    ldmatrix_x4(a[q][0], a[q][1], a[q][2], a[q][3], hb + swz_off<CIN>(prow, ks * 2 + a_kc));
                }
#pragma unroll
                for (int cb = 0; cb < 9; ++cb) {
                    uint32_t b[8];
                    const int wrow = combo_tap[cb] * COUT + nc * 32 + b_row;
                    ldmatrix_x4(b[0], b[1], b[2], b[3], wb + swz_off<CIN>(wrow, ks * 2 + b_kc));
                    ldmatrix_x4(b[4], b[5], b[6], b[7], wb + swz_off<CIN>(wrow + 16, ks * 2 + b_kc));
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
                        mma_bf16_16816(acc[combo_cls[cb]][nt], a[combo_a[cb]], b[nt * 2], b[nt * 2 + 1]);
                }
            }
        }

        // ---- epilogue -----------------------------------------------------------------------------
        const int NSLOT = (MODE == MODE_T2) ? 4 : 2;
#pragma unroll
        for (int slot = 0; slot < NSLOT; ++slot) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                int oy, ox;
                if (MODE == MODE_T2) {
                    int iy = t_y0 + warp, ix = t_x0 + c_row + half * 8;
                    if (iy >= p.Hin || ix >= p.Win) continue;
                    oy = 2 * iy + (slot >> 1);
                    ox = 2 * ix + (slot & 1);
                } else {
                    oy = t_y0 + 2 * warp + slot;
                    ox = t_x0 + c_row + half * 8;
                }
                if (oy >= p.Hout || ox >= p.Wout) continue;
                const size_t pix_off = out_n + ((size_t)oy * p.Wout + ox) * COUT;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int co = nc * 32 + nt * 8 + c_col;
                    float v0 = acc[slot][nt][half * 2 + 0] + s_bias[co];
                    float v1 = acc[slot][nt][half * 2 + 1] + s_bias[co + 1];
                    if (p.mask_mode != MASK_NONE) {
                        float2 m = unpack_bf162(*reinterpret_cast<const uint32_t*>(p.mask + pix_off + co));
                        if (p.mask_mode == MASK_RELU) {
                            v0 = m.x > 0.f ? v0 : 0.f;
                            v1 = m.y > 0.f ? v1 : 0.f;
                        } else {
                            float y0 = m.x * s_mscale[co] + s_mshift[co];
                            float y1 = m.y * s_mscale[co + 1] + s_mshift[co + 1];
                            v0 = y0 > 0.f ? v0 : v0 * p.slope;
                            v1 = y1 > 0.f ? v1 : v1 * p.slope;
                        }
                    }
                    if (p.add) {
                        float2 a2 = unpack_bf162(*reinterpret_cast<const uint32_t*>(p.add + pix_off + co));
                        v0 += a2.x;
                        v1 += a2.y;
                    }
                    if (p.relu_out) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                    *reinterpret_cast<uint32_t*>(p.out + pix_off + co) = pack_bf162(v0, v1);
                    if (p.out2) {
                        if (p.add2) {   // ReLU(bf16(out) + add2), bit-identical to a separate add kernel reading the stored output
                            const float2 o2 = unpack_bf162(pack_bf162(v0, v1));
                            const float2 a2 = unpack_bf162(*reinterpret_cast<const uint32_t*>(p.add2 + pix_off + co));
                            v0 = o2.x + a2.x; v1 = o2.y + a2.y;
                        }
                        *reinterpret_cast<uint32_t*>(p.out2 + pix_off + co) = pack_bf162(fmaxf(v0, 0.f), fmaxf(v1, 0.f));
                    }
                }
            }
        }
    }
}

template <int CIN, int COUT, int MODE>
int launch_conv3x3_t(ConvParams p, cudaStream_t st) {
    typedef ConvTile<MODE> T;
    typedef ConvSmem<CIN, COUT, MODE> S;
    static bool attr_set = false;
    if (!attr_set) {
        PTTA_CUDA(cudaFuncSetAttribute(conv3x3_mma_kernel<CIN, COUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        attr_set = true;
    }
    if (MODE == MODE_T2) {
        p.tiles_x = cdiv(p.Win, T::TW);
        p.tiles_y = cdiv(p.Hin, T::TH);
    } else {
        p.tiles_x = cdiv(p.Wout, T::TW);
        p.tiles_y = cdiv(p.Hout, T::TH);
    }
    dim3 grid(p.tiles_x * p.tiles_y, p.N);
    launch_k(conv3x3_mma_kernel<CIN, COUT, MODE>, grid, 256, S::TOTAL, st, p);
    return check_launch("conv3x3_mma");
}

// Hout/Wout are derived here: S1 same, S2 ceil(H/2), T2 2*H.
inline int launch_conv3x3(ConvParams p, int cin, int cout, int mode, cudaStream_t st) {
    if (mode == MODE_S1) { p.Hout = p.Hin; p.Wout = p.Win; }
    else if (mode == MODE_S2) { p.Hout = (p.Hin + 1) / 2; p.Wout = (p.Win + 1) / 2; }
    else if (mode == MODE_T2) { p.Hout = 2 * p.Hin; p.Wout = 2 * p.Win; }
    else { set_error("conv3x3: bad mode %d", mode); return 1; }
    if (cin == 32 && cout == 32) {
        if (mode == MODE_S1) return launch_conv3x3_t<32, 32, MODE_S1>(p, st);
        if (mode == MODE_S2) return launch_conv3x3_t<32, 32, MODE_S2>(p, st);
        return launch_conv3x3_t<32, 32, MODE_T2>(p, st);
    }
    if (cin == 32 && cout == 128 && mode == MODE_S1) return launch_conv3x3_t<32, 128, MODE_S1>(p, st);
    if (cin == 128 && cout == 32 && mode == MODE_S1) return launch_conv3x3_t<128, 32, MODE_S1>(p, st);
    set_error("conv3x3: unsupported shape cin=%d cout=%d mode=%d", cin, cout, mode);
    return 1;
}

// =================================================================================================
// Weight gradient of a 3x3 stride-1 conv (only the adapted "meta" layers need it:
// network_exp_msg_chn_adapt.py:28-36).  dW[co][ci][ky][kx] = sum_pixels g[p][co] * X[p + tap][ci],
// X = prologue(in).  Each CTA owns one 16x16 pixel tile and emits a partial [9][COUT][CIN] fp32
// block (pixels are the MMA K dimension; both operands are read through ldmatrix.trans).  A second
// kernel sums the partials in a fixed order -> deterministic, no atomics.
// =================================================================================================
struct WgradParams {
    const bf16* in;      // [N,H,W,CIN]   forward input of the conv (before the prologue)
    const bf16* gout;    // [N,H,W,COUT]  gradient wrt the conv output
    float* partial;      // [N*tiles][9][COUT][CIN]
    int N, H, W;
    int pro; const float* pro_scale; const float* pro_shift; float slope;
    int tiles_x, tiles_y;
};

template <int CIN, int COUT>
struct WgradSmem {
    static const int HALO_BYTES = ((18 * 18 * CIN * 2 + 127) / 128) * 128;
    static const int G_BYTES = 256 * COUT * 2;
    static const int TOTAL = HALO_BYTES + G_BYTES + 2 * CIN * 4;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(const WgradParams p) {
    PDL_SYNC();
    typedef WgradSmem<CIN, COUT> S;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_halo = smem;
    unsigned char* s_g = smem + S::HALO_BYTES;
    float* s_pro_scale = reinterpret_cast<float*>(smem + S::HALO_BYTES + S::G_BYTES);
    float* s_pro_shift = s_pro_scale + CIN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.y;
    const int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
    const int y0 = ty * 16, x0 = tx * 16;

    for (int i = tid; i < CIN; i += 256) {
        s_pro_scale[i] = p.pro == PRO_BN_LEAKY ? p.pro_scale[i] : 1.f;
        s_pro_shift[i] = p.pro == PRO_BN_LEAKY ? p.pro_shift[i] : 0.f;
    }
    __syncthreads();
    {   // input halo with prologue
        const int CH = CIN / 8;
        const bf16* in_n = p.in + (size_t)n * p.H * p.W * CIN;
        for (int i = tid; i < 18 * 18 * CH; i += 256) {
            int pix = i / CH, c = i - pix * CH;
            int hy = pix / 18, hx = pix - hy * 18;
            int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
                v = __ldg(reinterpret_cast<const uint4*>(in_n + ((size_t)gy * p.W + gx) * CIN + c * 8));
                if (p.pro == PRO_RELU) {
                    const bf162 z = __floats2bfloat162_rn(0.f, 0.f);
                    bf162* h = reinterpret_cast<bf162*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) h[j] = __hmax2(h[j], z);
                } else if (p.pro == PRO_BN_LEAKY) {
                    uint32_t* u = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = unpack_bf162(u[j]);
                        int ch = c * 8 + j * 2;
                        f.x = f.x * s_pro_scale[ch] + s_pro_shift[ch];
                        f.y = f.y * s_pro_scale[ch + 1] + s_pro_shift[ch + 1];
                        f.x = f.x > 0.f ? f.x : f.x * p.slope;
                        f.y = f.y > 0.f ? f.y : f.y * p.slope;
                        u[j] = pack_bf162(f.x, f.y);
                    }
                }
            }
            *reinterpret_cast<uint4*>(s_halo + swz_off<CIN>(pix, c)) = v;
        }
    }
    {   // output-gradient tile (zero outside the image)
        const int CH = COUT / 8;
        const bf16* g_n = p.gout + (size_t)n * p.H * p.W * COUT;
        for (int i = tid; i < 256 * CH; i += 256) {
            int pix = i / CH, c = i - pix * CH;
            int py = pix >> 4, px = pix & 15;
            int gy = y0 + py, gx = x0 + px;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (gy < p.H && gx < p.W)
                v = __ldg(reinterpret_cast<const uint4*>(g_n + ((size_t)gy * p.W + gx) * COUT + c * 8));
            *reinterpret_cast<uint4*>(s_g + swz_off<COUT>(pix, c)) = v;
        }
    }
    __syncthreads();

    const uint32_t hb = smem_u32(s_halo), gb = smem_u32(s_g);
    const int NCO = COUT / 32, NCI = CIN / 32;
    const int ntasks = 9 * NCO * NCI;
    // ldmatrix.trans lane roles
    const int a_k = (lane & 7) + (lane >> 4) * 8;   // A: stored row (pixel) within the k16 step
    const int a_mc = (lane >> 3) & 1;               // A: which 8-wide co chunk of the 16-row m-tile
    const int b_k = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int b_nc = lane >> 4;                     // B: which of two 8-wide ci chunks
    const int c_row = lane >> 2, c_col = (lane & 3) * 2;
    float* part = p.partial + ((size_t)(n * gridDim.x + blockIdx.x)) * 9 * COUT * CIN;

    for (int task = warp; task < ntasks; task += 8) {
        const int tap = task / (NCO * NCI);
        const int rem = task - tap * (NCO * NCI);
        const int cob = rem / NCI, cib = rem - cob * NCI;
        const int ky = tap / 3, kx = tap - ky * 3;
        float acc[2][4][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {   // one tile row of 16 pixels per k16 step
            uint32_t a[2][4], b[8];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                int grow = ks * 16 + a_k;
                int chunk = cob * 4 + mt * 2 + a_mc;
                ldmatrix_x4_trans(a[mt][0], a[mt][1], a[mt][2], a[mt][3], gb + swz_off<COUT>(grow, chunk));
            }
            const int prow = (ks + ky) * 18 + (b_k + kx);
            ldmatrix_x4_trans(b[0], b[1], b[2], b[3], hb + swz_off<CIN>(prow, cib * 4 + b_nc));
            ldmatrix_x4_trans(b[4], b[5], b[6], b[7], hb + swz_off<CIN>(prow, cib * 4 + 2 + b_nc));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], a[mt], b[nt * 2], b[nt * 2 + 1]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    int co = cob * 32 + mt * 16 + c_row + half * 8;
                    int ci = cib * 32 + nt * 8 + c_col;
                    float2 v = make_float2(acc[mt][nt][half * 2], acc[mt][nt][half * 2 + 1]);
                    *reinterpret_cast<float2*>(part + ((size_t)tap * COUT + co) * CIN + ci) = v;
                }
    }
}

// dW[co][ci][ky][kx] (PyTorch Conv2d layout) = sum over partial blocks.  Block = 32 outputs x 8 slices of the partial
// list; slices are combined through shared memory in a fixed order (deterministic, coalesced 128 B reads).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int nblocks, int cout, int cin) {
    PDL_SYNC();
    __shared__ float sh[8][32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int total = 9 * cout * cin;
    const int i = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < total)
        for (int b = slice; b < nblocks; b += 8) s += partial[(size_t)b * total + i];
    sh[slice][lane] = s;
    __syncthreads();
    if (slice != 0 || i >= total) return;
#pragma unroll
    for (int k = 1; k < 8; ++k) s += sh[k][lane];
    int tap = i / (cout * cin);
    int rem = i - tap * cout * cin;
    int co = rem / cin, ci = rem - co * cin;
    dw[((size_t)co * cin + ci) * 9 + tap] = s;
}

template <int CIN, int COUT>
int launch_wgrad_t(WgradParams p, float* dw, cudaStream_t st) {
    typedef WgradSmem<CIN, COUT> S;
    static bool attr_set = false;
    if (!attr_set) {
        PTTA_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        attr_set = true;
    }
    p.tiles_x = cdiv(p.W, 16);
    p.tiles_y = cdiv(p.H, 16);
    dim3 grid(p.tiles_x * p.tiles_y, p.N);
    launch_k(conv3x3_wgrad_kernel<CIN, COUT>, grid, 256, S::TOTAL, st, p);
    PTTA_TRY(check_launch("conv3x3_wgrad"));
    int total = 9 * COUT * CIN;
    launch_k(wgrad_reduce_kernel, cdiv(total, 32), 256, 0, st, p.partial, dw, p.tiles_x * p.tiles_y * p.N, COUT, CIN);
    return check_launch("wgrad_reduce");
}

inline size_t wgrad_partial_bytes(int N, int H, int W, int cin, int cout) {
    return (size_t)N * cdiv(W, 16) * cdiv(H, 16) * 9 * cout * cin * sizeof(float);
}

inline int launch_wgrad(WgradParams p, float* dw, int cin, int cout, cudaStream_t st) {
    if (cin == 32 && cout == 32) return launch_wgrad_t<32, 32>(p, dw, st);
    if (cin == 32 && cout == 128) return launch_wgrad_t<32, 128>(p, dw, st);
    if (cin == 128 && cout == 32) return launch_wgrad_t<128, 32>(p, dw, st);
    set_error("wgrad: unsupported shape cin=%d cout=%d", cin, cout);
    return 1;
}

}  // namespace ptta
