// On-device augmentations of the adaptation / preparation loops (the reference: src/transforms.py).
//   photometric (src/transforms.py:236-333 + torchvision.transforms.functional_tensor 0.10: `_blend`, `rgb_to_grayscale`, applied per
//   sample with python-float factors): the fp32 [0,255] image is truncated to uint8 (`images.to(torch.uint8)`, :240), then
//       brightness  u <- trunc(clamp(f u))
//       contrast    u <- trunc(clamp(f u + (1 - f) mean(gray(u))))        gray = trunc(0.2989 r + 0.587 g + 0.114 b)
//       saturation  u <- trunc(clamp(f u + (1 - f) gray(u)))
//   each only for the samples whose flag is set, each on the uint8 result of the previous one (the reference runs them as separate
//   tensor ops: every product and sum below is rounded on its own, no fused multiply-add), then `.float()` and the image
//   normalisation of :669-712.  One reduction pass (only for samples with the contrast flag: exact integer sum of the grey values)
//   and ONE elementwise pass instead of ~15 tensor ops and three host syncs per sample (`float(factors[b])`).
//   flips (src/transforms.py:386-407, 990-1034): per-sample horizontal / vertical mirror of an N x C x H x W map.
// The random draws stay where the reference makes them (torch.rand on the device, same order): tta_depth_completion_b200/transforms.py.
#pragma once
#include "common.cuh"

namespace ptta {

struct PhotoParams {
    const float* in; float* out;
    const unsigned char *do_b, *do_c, *do_s;      // [N] flags (nullptr = transform not configured)
    const float *f_b, *f_c, *f_s;                 // [N] factors, fp32 as drawn
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
// ratio u + (1 - ratio) other, the two products and the sum rounded separately (torchvision `_blend`); 1 - ratio is formed in double
__device__ __forceinline__ float photo_blend(float u, float other, float ratio) {
    const float omr = (float)(1.0 - (double)ratio);
    return photo_trunc(__fadd_rn(__fmul_rn(ratio, u), __fmul_rn(omr, other)));
}
__device__ __forceinline__ float photo_quant(float x) { return truncf(fminf(fmaxf(x, 0.f), 255.f)); }     // float -> uint8 cast of an in-range value

// exact sum of the grey values of every sample whose contrast flag is set, taken AFTER its brightness step
__global__ void __launch_bounds__(256) photo_gray_sum_kernel(const PhotoParams p) {
    PDL_SYNC();
    const int n = blockIdx.y;
    if (!p.do_c || !p.do_c[n]) return;
    const float* base = p.in + (size_t)n * 3 * p.HW;
    const bool bright = p.do_b && p.do_b[n];
    const float fb = bright ? p.f_b[n] : 1.f;
    unsigned int local = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < p.HW; i += gridDim.x * 256) {
        float r = photo_quant(base[i]), g = photo_quant(base[p.HW + i]), b = photo_quant(base[2 * p.HW + i]);
        if (bright) { r = photo_blend(r, 0.f, fb); g = photo_blend(g, 0.f, fb); b = photo_blend(b, 0.f, fb); }
        local += (unsigned int)photo_gray(r, g, b);
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

__global__ void __launch_bounds__(256) photo_apply_kernel(const PhotoParams p) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const float* base = p.in + (size_t)n * 3 * p.HW;
    float* ob = p.out + (size_t)n * 3 * p.HW;
    const bool bright = p.do_b && p.do_b[n], contrast = p.do_c && p.do_c[n], sat = p.do_s && p.do_s[n];
    const float fb = bright ? p.f_b[n] : 1.f, fc = contrast ? p.f_c[n] : 1.f, fs = sat ? p.f_s[n] : 1.f;
    const float mean = contrast ? (float)((double)p.gray_sum[n] / (double)p.HW) : 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < p.HW; i += gridDim.x * 256) {
        float c[3] = {base[i], base[p.HW + i], base[2 * p.HW + i]};
        if (p.quantize) {
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = photo_quant(c[k]);
            if (bright) {
#pragma unroll
                for (int k = 0; k < 3; ++k) c[k] = photo_blend(c[k], 0.f, fb);
            }
            if (contrast) {
#pragma unroll
                for (int k = 0; k < 3; ++k) c[k] = photo_blend(c[k], mean, fc);
            }
            if (sat) {
                const float gr = photo_gray(c[0], c[1], c[2]);
#pragma unroll
                for (int k = 0; k < 3; ++k) c[k] = photo_blend(c[k], gr, fs);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = c[k];
            if (p.norm_mode == 1) v = __fdiv_rn(v, 255.f);
            else if (p.norm_mode == 2) v = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(v, 255.f)), 1.f);
            else if (p.norm_mode == 3) v = __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.f), p.mean[k]), p.std[k]);
            ob[k * p.HW + i] = v;
        }
    }
}

// out[n][c][y][x] = in[n][c][vflip ? H-1-y : y][hflip ? W-1-x : x]
__global__ void __launch_bounds__(256) flip_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W,
                                                   const unsigned char* __restrict__ do_h, const unsigned char* __restrict__ do_v) {
    PDL_SYNC();
    const int n = blockIdx.y;
    const bool fh = do_h && do_h[n], fv = do_v && do_v[n];
    const int plane = H * W;
    const size_t off = (size_t)n * C * plane;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < C * plane; i += gridDim.x * 256) {
        const int c = i / plane, r = i - c * plane, y = r / W, x = r - y * W;
        const int sy = fv ? H - 1 - y : y, sx = fh ? W - 1 - x : x;
        out[off + i] = in[off + (size_t)c * plane + sy * W + sx];
    }
}

}  // namespace ptta
