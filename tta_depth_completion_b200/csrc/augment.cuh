// On-device augmentations of the adaptation / preparation loops (the reference: src/transforms.py).
//   photometric (src/transforms.py:236-333 + torchvision.transforms.functional_tensor 0.10: `_blend`, `rgb_to_grayscale`, applied per
//   sample with python-float factors): the fp32 [0,255] image is truncated to uint8 (`images.to(torch.uint8)`, :240), then
//       brightness  u <- trunc(clamp(f u))
//       contrast    u <- trunc(clamp(f u + (1 - f) mean(gray(u))))        gray = trunc(0.2989 r + 0.587 g + 0.114 b)
//       gamma       u <- trunc(255.999 clamp((u / 255) ^ g, 0, 1))          (torchvision adjust_gamma: uint8 -> float -> pow -> uint8)
//       hue         (r, g, b) / 255 -> HSV -> h <- (h + f) mod 1 -> RGB -> trunc(255.999 x)  (torchvision adjust_hue: `_rgb2hsv`, `_hsv2rgb`)
//       saturation  u <- trunc(clamp(f u + (1 - f) gray(u)))
//   and, on the float image, additive noise (:839-875): x + spread * n with n drawn by torch (gaussian) or spread * (u - 0.5) (uniform)
//   each only for the samples whose flag is set, each on the uint8 result of the previous one (the reference runs them as separate
//   tensor ops: every product and sum below is rounded on its own, no fused multiply-add), then `.float()` and the image
//   normalisation of :669-712.  One reduction pass (only for samples with the contrast flag: exact integer sum of the grey values)
//   and ONE elementwise pass instead of ~15 tensor ops and three host syncs per sample (`float(factors[b])`).
//   flips (src/transforms.py:386-407, 990-1034): per-sample horizontal / vertical mirror of an N x C x H x W map.
//   rotation (:406-423) and resize-and-crop (:425-502): per-sample resampling, nearest or bilinear (see the kernels).
//   random crop to a common shape (:337-383): per-sample window copy.
// The random draws stay where the reference makes them (torch.rand on the device, same order): tta_depth_completion_b200/transforms.py.
#pragma once
#include "common.cuh"

namespace ptta {

struct PhotoParams {
    const float* in; float* out;
    const unsigned char *do_b, *do_c, *do_s;      // [N] flags (nullptr = transform not configured)
    const float *f_b, *f_c, *f_s;                 // [N] factors, fp32 as drawn
    const unsigned char* do_g; const float* f_g;  // gamma jitter (between contrast and saturation, as in the reference)
    const unsigned char* do_h; const float* f_h;  // hue jitter (after gamma, before saturation)
    const unsigned char* do_n; const float* noise; float noise_spread; int noise_uniform;   // additive noise on the float image (before normalisation)
    unsigned long long* gray_sum;                 // [N] (contrast only)
    int N, HW;
    int quantize;                                 // images.to(uint8) happens whenever any photometric transform is configured
    int norm_mode;                                // 0 none ([0,255]), 1 [0,1], 2 [-1,1], 3 (x/255 - mean) / std
    float mean[3], std[3];
};

__device__ __forceinline__ float photo_trunc(float v) { return truncf(fminf(fmaxf(v, 0.f), 255.f)); }
__device__ __forceinline__ float photo_gray(float r, float g, float b) {
    return truncf(__fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b)));
}
// ratio u + (1 - ratio) other, the two products and the sum rounded separately (torchvision `_blend`); omr = 1 - ratio, formed in double
__device__ __forceinline__ float photo_one_minus(float ratio) { return (float)(1.0 - (double)ratio); }      // per sample, outside the pixel loops
__device__ __forceinline__ float photo_blend(float u, float other, float ratio, float omr) {
    return photo_trunc(__fadd_rn(__fmul_rn(ratio, u), __fmul_rn(omr, other)));
}
__device__ __forceinline__ float photo_quant(float x) { return truncf(fminf(fmaxf(x, 0.f), 255.f)); }     // float -> uint8 cast of an in-range value

// torchvision.transforms.functional_tensor.adjust_hue on one uint8 pixel, operation by operation (every elementwise tensor op of
// `_rgb2hsv` / `_hsv2rgb` is one separately rounded fp32 operation here; the one-hot einsum of `_hsv2rgb` selects one value exactly)
__device__ __forceinline__ float photo_remainder1(float x) {       // torch.remainder(x, 1.0)
    float m = fmodf(x, 1.f);
    if (m != 0.f && m < 0.f) m = __fadd_rn(m, 1.f);
    return m;
}
__device__ __forceinline__ void photo_hue(float& r8, float& g8, float& b8, float hue) {
    const float r = __fdiv_rn(r8, 255.f), g = __fdiv_rn(g8, 255.f), b = __fdiv_rn(b8, 255.f);
    const float maxc = fmaxf(fmaxf(r, g), b), minc = fminf(fminf(r, g), b);
    const bool eqc = maxc == minc;
    const float cr = __fsub_rn(maxc, minc);
    const float s = __fdiv_rn(cr, eqc ? 1.f : maxc);
    const float div = eqc ? 1.f : cr;
    const float rc = __fdiv_rn(__fsub_rn(maxc, r), div), gc = __fdiv_rn(__fsub_rn(maxc, g), div), bc = __fdiv_rn(__fsub_rn(maxc, b), div);
    const float hr = (maxc == r ? 1.f : 0.f) * __fsub_rn(bc, gc);
    const float hg = ((maxc == g && maxc != r) ? 1.f : 0.f) * __fsub_rn(__fadd_rn(2.f, rc), bc);
    const float hb = ((maxc != g && maxc != r) ? 1.f : 0.f) * __fsub_rn(__fadd_rn(4.f, gc), rc);
    float h = __fadd_rn(__fadd_rn(hr, hg), hb);
    h = fmodf(__fadd_rn(__fdiv_rn(h, 6.f), 1.f), 1.f);
    h = photo_remainder1(__fadd_rn(h, hue));
    const float v = maxc;
    const float h6 = __fmul_rn(h, 6.f);
    const float fi = floorf(h6);
    const float f = __fsub_rn(h6, fi);
    int i = (int)fi % 6;
    if (i < 0) i += 6;
    const float p = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, s)), 0.f), 1.f);
    const float q = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(f, s))), 0.f), 1.f);
    const float t = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(__fsub_rn(1.f, f), s))), 0.f), 1.f);
    float ro, go, bo;
    switch (i) {
        case 0: ro = v; go = t; bo = p; break;
        case 1: ro = q; go = v; bo = p; break;
        case 2: ro = p; go = v; bo = t; break;
        case 3: ro = p; go = q; bo = v; break;
        case 4: ro = t; go = p; bo = v; break;
        default: ro = v; go = p; bo = q; break;
    }
    // back to uint8 as torchvision's convert_image_dtype does it: trunc(x * (255 + 1 - 1e-3)).  (The release the reference pins, 0.10.1, is
    // believed to multiply by 255.0 at this one place; parity here is pinned on the installed release, 0.26, by running the reference's class.)
    r8 = truncf(__fmul_rn(ro, 255.999f)); g8 = truncf(__fmul_rn(go, 255.999f)); b8 = truncf(__fmul_rn(bo, 255.999f));
}

// exact sum of the grey values of every sample whose contrast flag is set, taken AFTER its brightness step.
// VEC = 4: four consecutive pixels per thread through 16-byte loads of each colour plane (HW % 4 == 0, 16-byte aligned planes)
template <int VEC>
__global__ void __launch_bounds__(256) photo_gray_sum_kernel(const PhotoParams p) {
    PDL_SYNC();
    const int n = blockIdx.y;
    if (!p.do_c || !p.do_c[n]) return;
    const float* base = p.in + (size_t)n * 3 * p.HW;
    const bool bright = p.do_b && p.do_b[n];
    const float fb = bright ? p.f_b[n] : 1.f, ob_ = photo_one_minus(fb);
    unsigned int local = 0;
    for (int i = (blockIdx.x * 256 + threadIdx.x) * VEC; i < p.HW; i += gridDim.x * 256 * VEC) {
        float c[3][VEC];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (VEC == 4) *reinterpret_cast<float4*>(c[k]) = __ldg(reinterpret_cast<const float4*>(base + (size_t)k * p.HW + i));
            else c[k][0] = base[(size_t)k * p.HW + i];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float r = photo_quant(c[0][v]), g = photo_quant(c[1][v]), b = photo_quant(c[2][v]);
            if (bright) { r = photo_blend(r, 0.f, fb, ob_); g = photo_blend(g, 0.f, fb, ob_); b = photo_blend(b, 0.f, fb, ob_); }
            local += (unsigned int)photo_gray(r, g, b);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ unsigned int sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        for (int k = 0; k < 8; ++k) tot += sh[k];
        atomicAdd(p.gray_sum + n, tot);             // integer: the total does not depend on the order
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) photo_apply_kernel(const PhotoParams p) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const float* base = p.in + (size_t)n * 3 * p.HW;
    float* ob = p.out + (size_t)n * 3 * p.HW;
    const bool bright = p.do_b && p.do_b[n], contrast = p.do_c && p.do_c[n], sat = p.do_s && p.do_s[n];
    const float fb = bright ? p.f_b[n] : 1.f, fc = contrast ? p.f_c[n] : 1.f, fs = sat ? p.f_s[n] : 1.f;
    const float mean = contrast ? (float)((double)p.gray_sum[n] / (double)p.HW) : 0.f;
    const float omb = photo_one_minus(fb), omc = photo_one_minus(fc), oms = photo_one_minus(fs);
    // after the uint8 truncation a channel holds one of 256 values: the normalisation (IEEE divisions, as torch's `/ 255.0` and
    // `normalize`) is tabulated once per block instead of being evaluated 12 times per thread and iteration
    // gamma acts on one of 256 channel values: tabulated per block (one powf per entry instead of three per pixel)
    __shared__ float glut[256];
    const bool hue_on = p.do_h && p.do_h[n];
    const float fh = hue_on ? p.f_h[n] : 0.f;
    const bool noisy = p.do_n && p.do_n[n];
    const float* nz = p.noise ? p.noise + (size_t)n * 3 * p.HW : nullptr;
    const bool gamma = p.do_g && p.do_g[n];
    if (gamma) {
        const float x = __fdiv_rn((float)threadIdx.x, 255.f);
        const float y = fminf(fmaxf(powf(x, p.f_g[n]), 0.f), 1.f);
        glut[threadIdx.x] = truncf(__fmul_rn(y, 255.999f));
    }
    __shared__ float lut[3][256];
    const bool use_lut = p.quantize && p.norm_mode != 0;
    if (use_lut) {
        const float v = (float)threadIdx.x;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float x = __fdiv_rn(v, 255.f);
            if (p.norm_mode == 2) x = __fsub_rn(__fmul_rn(2.f, x), 1.f);
            else if (p.norm_mode == 3) x = __fdiv_rn(__fsub_rn(x, p.mean[k]), p.std[k]);
            lut[k][threadIdx.x] = x;
        }
    }
    __syncthreads();
    for (int i = (blockIdx.x * 256 + threadIdx.x) * VEC; i < p.HW; i += gridDim.x * 256 * VEC) {
        float c[3][VEC];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (VEC == 4) *reinterpret_cast<float4*>(c[k]) = __ldg(reinterpret_cast<const float4*>(base + (size_t)k * p.HW + i));
            else c[k][0] = base[(size_t)k * p.HW + i];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (p.quantize) {
#pragma unroll
                for (int k = 0; k < 3; ++k) c[k][v] = photo_quant(c[k][v]);
                if (bright) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) c[k][v] = photo_blend(c[k][v], 0.f, fb, omb);
                }
                if (contrast) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) c[k][v] = photo_blend(c[k][v], mean, fc, omc);
                }
                if (gamma) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) c[k][v] = glut[(int)c[k][v]];
                }
                if (hue_on) photo_hue(c[0][v], c[1][v], c[2][v], fh);
                if (sat) {
                    const float gr = photo_gray(c[0][v], c[1][v], c[2][v]);
#pragma unroll
                    for (int k = 0; k < 3; ++k) c[k][v] = photo_blend(c[k][v], gr, fs, oms);
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float x = c[k][v];
                if (noisy) {
                    float nv = nz[(size_t)k * p.HW + i + v];
                    if (p.noise_uniform) nv = __fsub_rn(nv, 0.5f);
                    x = __fadd_rn(x, __fmul_rn(p.noise_spread, nv));
                }
                if (use_lut && !noisy) x = lut[k][(int)x];
                else if (p.norm_mode == 1) x = __fdiv_rn(x, 255.f);
                else if (p.norm_mode == 2) x = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(x, 255.f)), 1.f);
                else if (p.norm_mode == 3) x = __fdiv_rn(__fsub_rn(__fdiv_rn(x, 255.f), p.mean[k]), p.std[k]);
                c[k][v] = x;
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (VEC == 4) *reinterpret_cast<float4*>(ob + (size_t)k * p.HW + i) = *reinterpret_cast<const float4*>(c[k]);
            else ob[(size_t)k * p.HW + i] = c[k][0];
        }
    }
}

// out[n][c][y][x] = in[n][c][vflip ? H-1-y : y][hflip ? W-1-x : x]
// VEC = 4: four consecutive output pixels per thread (W % 4 == 0, 16-byte aligned rows): one 16-byte load (of the mirrored quad when
// hflip) and one 16-byte store
template <int VEC>
__global__ void __launch_bounds__(256) flip_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W,
                                                   const unsigned char* __restrict__ do_h, const unsigned char* __restrict__ do_v) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const bool fh = do_h && do_h[n], fv = do_v && do_v[n];
    const int plane = H * W;
    const size_t off = (size_t)n * C * plane;
    for (int i = (blockIdx.x * 256 + threadIdx.x) * VEC; i < C * plane; i += gridDim.x * 256 * VEC) {
        const int c = i / plane, r = i - c * plane, y = r / W, x = r - y * W;
        const int sy = fv ? H - 1 - y : y;
        const float* row = in + off + (size_t)c * plane + (size_t)sy * W;
        if (VEC == 4) {
            float4 v = __ldg(reinterpret_cast<const float4*>(row + (fh ? W - 4 - x : x)));
            if (fh) v = make_float4(v.w, v.z, v.y, v.x);
            *reinterpret_cast<float4*>(out + off + i) = v;
        } else {
            out[off + i] = row[fh ? W - 1 - x : x];
        }
    }
}

// Per-sample rotation about the image centre (src/transforms.py:406-423, 1036-1070 -> torchvision functional.rotate, expand=False,
// fill=None -> affine grid + grid_sample(padding_mode='zeros', align_corners=False)).  theta: [N][6] = the 3 x 2 matrix
// `theta^T / (0.5 w, 0.5 h)` torchvision multiplies the base grid with (row-major: x-column then y-column), prepared on the host exactly
// as torchvision does (double-precision matrix -> fp32 -> fp32 division).  Base grid = pixel centres relative to the image centre.
// mode 0: nearest (round half to even, as grid_sample), 1: bilinear.  Samples whose flag is clear are copied.
__global__ void __launch_bounds__(256) rotate_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W,
                                                     const unsigned char* __restrict__ do_rot, const float* __restrict__ theta, int mode) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const int plane = H * W;
    const float* ib = in + (size_t)n * C * plane;
    float* ob = out + (size_t)n * C * plane;
    if (!do_rot[n]) {
        for (int i = blockIdx.x * 256 + threadIdx.x; i < C * plane; i += gridDim.x * 256) ob[i] = ib[i];
        return;
    }
    const float t0 = theta[n * 6 + 0], t1 = theta[n * 6 + 1], t2 = theta[n * 6 + 2], t3 = theta[n * 6 + 3], t4 = theta[n * 6 + 4], t5 = theta[n * 6 + 5];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < plane; i += gridDim.x * 256) {
        const int y = i / W, x = i - y * W;
        const float xb = (float)x - 0.5f * (float)W + 0.5f, yb = (float)y - 0.5f * (float)H + 0.5f;
        const float gx = __fadd_rn(__fadd_rn(__fmul_rn(xb, t0), __fmul_rn(yb, t1)), t2);
        const float gy = __fadd_rn(__fadd_rn(__fmul_rn(xb, t3), __fmul_rn(yb, t4)), t5);
        const float ix = ((gx + 1.f) * (float)W - 1.f) / 2.f, iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
        if (mode == 0) {
            const float fx = nearbyintf(ix), fy = nearbyintf(iy);
            const bool inb = fx >= 0.f && fx <= (float)(W - 1) && fy >= 0.f && fy <= (float)(H - 1);
            const int src = inb ? (int)fy * W + (int)fx : 0;
            for (int c = 0; c < C; ++c) ob[c * plane + i] = inb ? ib[c * plane + src] : 0.f;
        } else {
            const float x0 = floorf(ix), y0 = floorf(iy);
            const float wx1 = ix - x0, wx0 = (x0 + 1.f) - ix, wy1 = iy - y0, wy0 = (y0 + 1.f) - iy;
            const int xi = (int)x0, yi = (int)y0;
            const bool vx0 = xi >= 0 && xi < W, vx1 = xi + 1 >= 0 && xi + 1 < W, vy0 = yi >= 0 && yi < H, vy1 = yi + 1 >= 0 && yi + 1 < H;
            for (int c = 0; c < C; ++c) {
                const float* pc = ib + c * plane;
                float v = 0.f;
                if (vy0 && vx0) v += pc[yi * W + xi] * (wx0 * wy0);
                if (vy0 && vx1) v += pc[yi * W + xi + 1] * (wx1 * wy0);
                if (vy1 && vx0) v += pc[(yi + 1) * W + xi] * (wx0 * wy1);
                if (vy1 && vx1) v += pc[(yi + 1) * W + xi + 1] * (wx1 * wy1);
                ob[c * plane + i] = v;
            }
        }
    }
}

// Per-sample enlargement to (rh, rw) >= (H, W) followed by the crop [sy, sy + H) x [sx, sx + W) (src/transforms.py:425-502, 1222-1283 ->
// torchvision functional.resize -> F.interpolate, align_corners=False): only the H x W pixels that survive the crop are computed.
// mode 0: nearest (source index floor(dst * in / out)), 1: bilinear (source coordinate max(in / out * (dst + 0.5) - 0.5, 0)).
// src/transforms.py:1274-1275 (`resize_scaling_depth`): after resize-and-crop every tensor but the first is divided, per sample the transform
// touched, by float32(r_width / n_width) -- an IEEE fp32 division per element, in place, as `image /= (r_width / n_width)` does
__global__ void __launch_bounds__(256) divide_samples_kernel(float* __restrict__ data, long long per_sample, const unsigned char* __restrict__ flags,
                                                             const float* __restrict__ divisor) {
    PDL_SYNC();
    const int n = blockIdx.y;
    if (!flags[n]) return;
    const float d = divisor[n];
    float* p = data + (size_t)n * per_sample;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < per_sample; i += gridDim.x * 256ll) p[i] = __fdiv_rn(p[i], d);
}

__global__ void __launch_bounds__(256) resize_crop_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W,
                                                          const unsigned char* __restrict__ do_rs, const int* __restrict__ rh,
                                                          const int* __restrict__ rw, const int* __restrict__ sy, const int* __restrict__ sx,
                                                          int mode) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const int plane = H * W;
    const float* ib = in + (size_t)n * C * plane;
    float* ob = out + (size_t)n * C * plane;
    if (!do_rs[n]) {
        for (int i = blockIdx.x * 256 + threadIdx.x; i < C * plane; i += gridDim.x * 256) ob[i] = ib[i];
        return;
    }
    const float sh = (float)H / (float)rh[n], sw = (float)W / (float)rw[n];
    const int oy = sy[n], ox = sx[n];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < plane; i += gridDim.x * 256) {
        const int y = i / W, x = i - y * W;
        const int Y = y + oy, X = x + ox;                  // position in the enlarged image
        if (mode == 0) {
            const int ys = min((int)floorf((float)Y * sh), H - 1), xs = min((int)floorf((float)X * sw), W - 1);
            for (int c = 0; c < C; ++c) ob[c * plane + i] = ib[c * plane + ys * W + xs];
        } else {
            const float yr = fmaxf(sh * ((float)Y + 0.5f) - 0.5f, 0.f), xr = fmaxf(sw * ((float)X + 0.5f) - 0.5f, 0.f);
            const int y1 = (int)yr, x1 = (int)xr;
            const int yp = y1 < H - 1 ? 1 : 0, xp = x1 < W - 1 ? 1 : 0;
            const float ly1 = yr - (float)y1, ly0 = 1.f - ly1, lx1 = xr - (float)x1, lx0 = 1.f - lx1;
            for (int c = 0; c < C; ++c) {
                const float* pc = ib + c * plane;
                ob[c * plane + i] = ly0 * (lx0 * pc[y1 * W + x1] + lx1 * pc[y1 * W + x1 + xp]) +
                                    ly1 * (lx0 * pc[(y1 + yp) * W + x1] + lx1 * pc[(y1 + yp) * W + x1 + xp]);
            }
        }
    }
}

// Per-sample crop to a common (ch, cw) window (src/transforms.py:337-383, 955-988): out[n][c][y][x] = in[n][c][sy[n] + y][sx[n] + x]
__global__ void __launch_bounds__(256) crop_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W, int ch, int cw,
                                                   const int* __restrict__ sy, const int* __restrict__ sx) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const int oplane = ch * cw, oy = sy[n], ox = sx[n];
    const float* ib = in + (size_t)n * C * H * W;
    float* ob = out + (size_t)n * C * oplane;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < C * oplane; i += gridDim.x * 256) {
        const int c = i / oplane, r = i - c * oplane, y = r / cw, x = r - y * cw;
        ob[i] = ib[(size_t)c * H * W + (size_t)(oy + y) * W + ox + x];
    }
}

// Per-sample crop-and-pad (src/transforms.py:508-566, 1072-1135, constant padding): the window [sy, sy + ch) x [sx, sx + cw) of the sample is
// placed at (pad_top, pad_left) of an H x W map of zeros.  win: [N][6] = {sy, sx, ch, cw, pad_top, pad_left}.  Flag clear: copy.
__global__ void __launch_bounds__(256) crop_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W,
                                                       const unsigned char* __restrict__ do_cp, const int* __restrict__ win) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const int plane = H * W;
    const float* ib = in + (size_t)n * C * plane;
    float* ob = out + (size_t)n * C * plane;
    if (!do_cp[n]) {
        for (int i = blockIdx.x * 256 + threadIdx.x; i < C * plane; i += gridDim.x * 256) ob[i] = ib[i];
        return;
    }
    const int sy = win[n * 6], sx = win[n * 6 + 1], ch = win[n * 6 + 2], cw = win[n * 6 + 3], pt = win[n * 6 + 4], pl = win[n * 6 + 5];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < C * plane; i += gridDim.x * 256) {
        const int c = i / plane, r = i - c * plane, y = r / W, x = r - y * W;
        const int yy = y - pt, xx = x - pl;
        ob[i] = (yy >= 0 && yy < ch && xx >= 0 && xx < cw) ? ib[(size_t)c * plane + (size_t)(sy + yy) * W + sx + xx] : 0.f;
    }
}

// Random point removal (src/transforms.py:625-652, 878-953): every SELECTED pixel of a sample (uint8 map `sel`, chosen by the caller from the
// sample's non-zero pixels with torch.randperm, as the reference does) erases the ph x pw patch around it (the reference marks the point
// with inf, max-pools with the patch and zeroes what became inf).  out = 0 inside a patch, the input elsewhere.  patch: [N][2] = {ph, pw} (odd).
__global__ void __launch_bounds__(256) remove_patches_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W,
                                                             const unsigned char* __restrict__ do_rm, const unsigned char* __restrict__ sel,
                                                             const int* __restrict__ patch) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const int plane = H * W;
    const float* ib = in + (size_t)n * C * plane;
    float* ob = out + (size_t)n * C * plane;
    if (!do_rm[n]) {
        for (int i = blockIdx.x * 256 + threadIdx.x; i < C * plane; i += gridDim.x * 256) ob[i] = ib[i];
        return;
    }
    const unsigned char* sb = sel + (size_t)n * plane;
    const int ry = patch[n * 2] / 2, rx = patch[n * 2 + 1] / 2;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < plane; i += gridDim.x * 256) {
        const int y = i / W, x = i - y * W;
        bool hit = false;
        for (int yy = max(y - ry, 0); yy <= min(y + ry, H - 1) && !hit; ++yy)
            for (int xx = max(x - rx, 0); xx <= min(x + rx, W - 1); ++xx)
                if (sb[yy * W + xx]) { hit = true; break; }
        for (int c = 0; c < C; ++c) ob[c * plane + i] = hit ? 0.f : ib[c * plane + i];
    }
}

// Per-sample reduction to (rh, rw) <= (H, W) placed at (pad_top, pad_left) of an H x W map of zeros (src/transforms.py:578-622, 1137-1220:
// functional.resize + functional.pad, constant padding).  Source coordinates as F.interpolate computes them (align_corners = False, NO
// anti-aliasing: the behaviour of the torchvision release the reference pins, 0.10.1; later releases low-pass filter a bilinear reduction
// by default).  mode 0 nearest, 1 bilinear.  geo: [N][4] = {rh, rw, pad_top, pad_left}.
__global__ void __launch_bounds__(256) resize_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W,
                                                         const unsigned char* __restrict__ do_rp, const int* __restrict__ geo, int mode) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const int plane = H * W;
    const float* ib = in + (size_t)n * C * plane;
    float* ob = out + (size_t)n * C * plane;
    if (!do_rp[n]) {
        for (int i = blockIdx.x * 256 + threadIdx.x; i < C * plane; i += gridDim.x * 256) ob[i] = ib[i];
        return;
    }
    const int rh = geo[n * 4], rw = geo[n * 4 + 1], pt = geo[n * 4 + 2], pl = geo[n * 4 + 3];
    const float sh = (float)H / (float)rh, sw = (float)W / (float)rw;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < plane; i += gridDim.x * 256) {
        const int y = i / W, x = i - y * W;
        const int Y = y - pt, X = x - pl;                  // position in the reduced image
        if (Y < 0 || Y >= rh || X < 0 || X >= rw) {
            for (int c = 0; c < C; ++c) ob[c * plane + i] = 0.f;
        } else if (mode == 0) {
            const int ys = min((int)floorf((float)Y * sh), H - 1), xs = min((int)floorf((float)X * sw), W - 1);
            for (int c = 0; c < C; ++c) ob[c * plane + i] = ib[c * plane + ys * W + xs];
        } else {
            const float yr = fmaxf(sh * ((float)Y + 0.5f) - 0.5f, 0.f), xr = fmaxf(sw * ((float)X + 0.5f) - 0.5f, 0.f);
            const int y1 = (int)yr, x1 = (int)xr;
            const int yp = y1 < H - 1 ? 1 : 0, xp = x1 < W - 1 ? 1 : 0;
            const float ly1 = yr - (float)y1, ly0 = 1.f - ly1, lx1 = xr - (float)x1, lx0 = 1.f - lx1;
            for (int c = 0; c < C; ++c) {
                const float* pc = ib + c * plane;
                ob[c * plane + i] = ly0 * (lx0 * pc[y1 * W + x1] + lx1 * pc[y1 * W + x1 + xp]) +
                                    ly1 * (lx0 * pc[(y1 + yp) * W + x1] + lx1 * pc[(y1 + yp) * W + x1 + xp]);
            }
        }
    }
}

}  // namespace ptta
