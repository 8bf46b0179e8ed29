// NLSPN non-local spatial propagation on B200.
//
// Reference: external_src/NLSPN/src/model/nlspnmodel_adapt.py:340-373 runs `prop_time` (18) modulated deformable
// convolutions (DCNv2, deformconv/src/cuda/modulated_deform_im2col_cuda.cuh) with a 1-channel feature map, an all-ones
// 3x3 weight and the 9 affinities as the modulation mask; every call materialises a 9 x H x W `columns` buffer, runs a
// GEMV over it and permutes the result (modulated_deform_conv_cuda.cu:78-113).  Here one thread owns one pixel: it reads its
// 18 offsets + 9 affinities (planar, coalesced), gathers 9 x 4 corners of the feature map (L1/L2 hits: offsets are a
// few pixels) and writes one float -- no columns buffer, no GEMV, no permute; the input-preserving blend
// `(1-m)*feat + m*sparse` (:364-366) is folded into the store of the previous iteration.
//
// The arithmetic follows the reference kernels term by term (sampling window (-1,H)x(-1,W), floor corners, corner validity,
// value * mask, coordinate weights) so that results agree to fp32 rounding; only the order of the 9-term sum differs
// (the reference's GEMV order is cuBLAS-internal).
//
// Memory-bound: per pixel and iteration 27 floats of offsets/affinities + 1 read + 1 write = 116 B;
// the 46 MB of offsets + affinities of a 352x1216 frame stay resident in the 126 MB L2 across the 18 iterations.
#pragma once
#include "common.cuh"

namespace ptta {

struct BilinearTap {
    int h_low, w_low;
    float hh, hw, lh, lw;
    bool inside, ok1, ok2, ok3, ok4;
};

// modulated_deform_im2col_cuda.cuh:24-54 (corner rules) and :177 (window test)
__device__ __forceinline__ BilinearTap make_tap(float h_im, float w_im, int H, int W) {
    BilinearTap t;
    t.inside = h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;
    const float hf = floorf(h_im), wf = floorf(w_im);
    t.h_low = (int)hf; t.w_low = (int)wf;
    t.lh = h_im - hf; t.lw = w_im - wf;
    t.hh = 1.f - t.lh; t.hw = 1.f - t.lw;
    const int h_high = t.h_low + 1, w_high = t.w_low + 1;
    t.ok1 = t.inside && t.h_low >= 0 && t.w_low >= 0;
    t.ok2 = t.inside && t.h_low >= 0 && w_high <= W - 1;
    t.ok3 = t.inside && h_high <= H - 1 && t.w_low >= 0;
    t.ok4 = t.inside && h_high <= H - 1 && w_high <= W - 1;
    return t;
}

__device__ __forceinline__ void tap_corners(const BilinearTap& t, const float* __restrict__ img, int W, float& v1, float& v2, float& v3, float& v4) {
    const float* p = img + (long long)t.h_low * W + t.w_low;
    v1 = t.ok1 ? __ldg(p) : 0.f;
    v2 = t.ok2 ? __ldg(p + 1) : 0.f;
    v3 = t.ok3 ? __ldg(p + W) : 0.f;
    v4 = t.ok4 ? __ldg(p + W + 1) : 0.f;
}

__device__ __forceinline__ float tap_value(const BilinearTap& t, float v1, float v2, float v3, float v4) {
    return (t.hh * t.hw) * v1 + (t.hh * t.lw) * v2 + (t.lh * t.hw) * v3 + (t.lh * t.lw) * v4;
}

#define PROP_TX 32
#define PROP_TY 8

// ---------------------------------------------------------------------------------------------------------------------
// a21: single-channel modulated deformable convolution, kernel KS x KS (3x3 pad 1 / 1x1 pad 0 are what NLSPN calls;
// any odd KS <= 7 with stride 1, dilation 1 works).  out = bias + sum_k w[k] * mask_k * bilinear(in, p + k + offset_k)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PROP_TX * PROP_TY) mdconv1_forward_kernel(
    const float* __restrict__ in, const float* __restrict__ weight, const float* __restrict__ bias, const float* __restrict__ offset,
    const float* __restrict__ mask, float* __restrict__ out, int H, int W, int Ho, int Wo, int KS, int pad) {
    PDL_SYNC();
    const int x = blockIdx.x * PROP_TX + threadIdx.x, y = blockIdx.y * PROP_TY + threadIdx.y, n = blockIdx.z;
    if (x >= Wo || y >= Ho) return;
    const int K = KS * KS;
    const long long plane = (long long)Ho * Wo, pix = (long long)y * Wo + x;
    const float* img = in + (long long)n * H * W;
    const float* off = offset + (long long)n * 2 * K * plane + pix;
    const float* msk = mask + (long long)n * K * plane + pix;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        const int i = k / KS, j = k - i * KS;
        const float h_im = (float)(y - pad + i) + __ldg(off + (long long)(2 * k) * plane);
        const float w_im = (float)(x - pad + j) + __ldg(off + (long long)(2 * k + 1) * plane);
        const BilinearTap t = make_tap(h_im, w_im, H, W);
        float v1, v2, v3, v4;
        tap_corners(t, img, W, v1, v2, v3, v4);
        acc = fmaf(__ldg(weight + k), tap_value(t, v1, v2, v3, v4) * __ldg(msk + (long long)k * plane), acc);
    }
    out[(long long)n * plane + pix] = acc + (bias ? __ldg(bias) : 0.f);
}

// backward of the above: grad_offset / grad_mask written, grad_input accumulated with fp32 reductions (as the reference's
// col2im kernel does, cuh:196-252), grad_weight / grad_bias accumulated per block then one atomic per block and tap
__global__ void __launch_bounds__(PROP_TX * PROP_TY) mdconv1_backward_kernel(
    const float* __restrict__ in, const float* __restrict__ weight, const float* __restrict__ offset, const float* __restrict__ mask,
    const float* __restrict__ gout, float* __restrict__ gin, float* __restrict__ goffset, float* __restrict__ gmask,
    float* __restrict__ gweight, float* __restrict__ gbias, int H, int W, int Ho, int Wo, int KS, int pad) {
    PDL_SYNC();
    __shared__ float s_red[50][PROP_TY];
    const int x = blockIdx.x * PROP_TX + threadIdx.x, y = blockIdx.y * PROP_TY + threadIdx.y, n = blockIdx.z;
    const bool active = x < Wo && y < Ho;
    const int K = KS * KS;
    const long long plane = (long long)Ho * Wo, pix = (long long)y * Wo + x;
    const float* img = in + (long long)n * H * W;
    float* gimg = gin ? gin + (long long)n * H * W : nullptr;
    const float go = active ? __ldg(gout + (long long)n * plane + pix) : 0.f;
    for (int k = 0; k < K; ++k) {
        float gw_k = 0.f;
        if (active) {
            const int i = k / KS, j = k - i * KS;
            const float* off = offset + (long long)n * 2 * K * plane + pix;
            const float h_im = (float)(y - pad + i) + __ldg(off + (long long)(2 * k) * plane);
            const float w_im = (float)(x - pad + j) + __ldg(off + (long long)(2 * k + 1) * plane);
            const float m = __ldg(mask + ((long long)n * K + k) * plane + pix);
            const BilinearTap t = make_tap(h_im, w_im, H, W);
            float v1, v2, v3, v4;
            tap_corners(t, img, W, v1, v2, v3, v4);
            const float val = tap_value(t, v1, v2, v3, v4);
            const float gcol = __ldg(weight + k) * go;                    // columns = W^T grad_out (modulated_deform_conv_cuda.cu:216-221)
            // cuh:254-328: grad_mask = gcol * sample; grad_offset = coordinate weight * gcol * mask (0 outside the window)
            const float wgt_h = -t.hw * v1 - t.lw * v2 + t.hw * v3 + t.lw * v4;
            const float wgt_w = -t.hh * v1 + t.hh * v2 - t.lh * v3 + t.lh * v4;
            if (gmask) gmask[((long long)n * K + k) * plane + pix] = t.inside ? gcol * val : 0.f;
            if (goffset) {
                goffset[((long long)n * 2 * K + 2 * k) * plane + pix] = t.inside ? wgt_h * gcol * m : 0.f;
                goffset[((long long)n * 2 * K + 2 * k + 1) * plane + pix] = t.inside ? wgt_w * gcol * m : 0.f;
            }
            if (gimg) {
                const float top = gcol * m;
                float* p = gimg + (long long)t.h_low * W + t.w_low;
                if (t.ok1) atomicAdd(p, t.hh * t.hw * top);
                if (t.ok2) atomicAdd(p + 1, t.hh * t.lw * top);
                if (t.ok3) atomicAdd(p + W, t.lh * t.hw * top);
                if (t.ok4) atomicAdd(p + W + 1, t.lh * t.lw * top);
            }
            gw_k = go * val * m;
        }
        if (gweight) {
            gw_k = warp_sum(gw_k);
            if (threadIdx.x == 0) s_red[k][threadIdx.y] = gw_k;
        }
    }
    if (gbias) {
        const float gb = warp_sum(go);
        if (threadIdx.x == 0) s_red[49][threadIdx.y] = gb;
    }
    __syncthreads();
    const int tid = threadIdx.y * PROP_TX + threadIdx.x;
    if (gweight && tid < K) {
        float s = 0.f;
        for (int r = 0; r < PROP_TY; ++r) s += s_red[tid][r];
        atomicAdd(gweight + tid, s);
    }
    if (gbias && tid == 49) {
        float s = 0.f;
        for (int r = 0; r < PROP_TY; ++r) s += s_red[49][r];
        atomicAdd(gbias, s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// a20: the propagation loop
// ---------------------------------------------------------------------------------------------------------------------
// out = fix > 0 ? fix : in   (nlspnmodel_adapt.py:356-358, 364-366: mask_fix = (feat_fix > 0), feat = (1-m)*feat + m*fix)
__global__ void prop_blend_kernel(const float* __restrict__ in, const float* __restrict__ fix, float* __restrict__ out, long long total) {
    PDL_SYNC();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float f = fix ? __ldg(fix + i) : 0.f;
    out[i] = f > 0.f ? f : __ldg(in + i);
}

// one propagation step: raw = sum_k aff_k * bilinear(fb, p + grid_k + offset_k);
//   out_raw (optional)     = raw
//   out_blend (optional)   = fix > 0 ? fix : raw      (input of the next step)
__global__ void __launch_bounds__(PROP_TX * PROP_TY) prop_step_kernel(const float* __restrict__ fb, const float* __restrict__ offset,
                                                                      const float* __restrict__ aff, const float* __restrict__ fix,
                                                                      float* __restrict__ out_raw, float* __restrict__ out_blend, int H, int W) {
    PDL_SYNC();
    const int x = blockIdx.x * PROP_TX + threadIdx.x, y = blockIdx.y * PROP_TY + threadIdx.y, n = blockIdx.z;
    if (x >= W || y >= H) return;
    const long long plane = (long long)H * W, pix = (long long)y * W + x;
    const float* img = fb + (long long)n * plane;
    const float* off = offset + (long long)n * 18 * plane + pix;
    const float* af = aff + (long long)n * 9 * plane + pix;
    float o[18], a[9];
#pragma unroll
    for (int k = 0; k < 18; ++k) o[k] = __ldg(off + (long long)k * plane);
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = __ldg(af + (long long)k * plane);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float h_im = (float)(y - 1 + k / 3) + o[2 * k];
        const float w_im = (float)(x - 1 + k % 3) + o[2 * k + 1];
        const BilinearTap t = make_tap(h_im, w_im, H, W);
        float v1, v2, v3, v4;
        tap_corners(t, img, W, v1, v2, v3, v4);
        acc += tap_value(t, v1, v2, v3, v4) * a[k];           // weight == 1, bias == 0 (nlspnmodel_adapt.py:232-236)
    }
    const long long idx = (long long)n * plane + pix;
    if (out_raw) out_raw[idx] = acc;
    if (out_blend) {
        const float f = fix ? __ldg(fix + idx) : 0.f;
        out_blend[idx] = f > 0.f ? f : acc;
    }
}

// backward of one propagation step.
//   g = grad wrt this step's raw output: g_in[p] (masked by [fix <= 0] when `mask_g`: the blend that followed this step passes
//       the gradient only where the sparse input is absent); g_in[p] is cleared after it is read (the buffer is the scatter
//       target two steps later)
//   grad_offset / grad_aff (+)= this step's contribution (first call: accumulate == 0 -> plain store, no zero-fill needed)
//   g_scatter += grad wrt this step's (blended) input, fp32 reductions
__global__ void __launch_bounds__(PROP_TX * PROP_TY) prop_step_backward_kernel(const float* __restrict__ fb, const float* __restrict__ offset,
                                                                               const float* __restrict__ aff, const float* __restrict__ fix,
                                                                               float* __restrict__ g_in, float* __restrict__ g_scatter,
                                                                               float* __restrict__ goffset, float* __restrict__ gaff, int H, int W,
                                                                               int mask_g, int accumulate) {
    PDL_SYNC();
    const int x = blockIdx.x * PROP_TX + threadIdx.x, y = blockIdx.y * PROP_TY + threadIdx.y, n = blockIdx.z;
    if (x >= W || y >= H) return;
    const long long plane = (long long)H * W, pix = (long long)y * W + x, idx = (long long)n * plane + pix;
    const float* img = fb + (long long)n * plane;
    float* gimg = g_scatter + (long long)n * plane;
    float go = g_in[idx];
    g_in[idx] = 0.f;
    if (mask_g && fix && __ldg(fix + idx) > 0.f) go = 0.f;
    const float* off = offset + (long long)n * 18 * plane + pix;
    const float* af = aff + (long long)n * 9 * plane + pix;
    float* gof = goffset + (long long)n * 18 * plane + pix;
    float* gaf = gaff + (long long)n * 9 * plane + pix;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float h_im = (float)(y - 1 + k / 3) + __ldg(off + (long long)(2 * k) * plane);
        const float w_im = (float)(x - 1 + k % 3) + __ldg(off + (long long)(2 * k + 1) * plane);
        const float m = __ldg(af + (long long)k * plane);
        const BilinearTap t = make_tap(h_im, w_im, H, W);
        float v1, v2, v3, v4;
        tap_corners(t, img, W, v1, v2, v3, v4);
        const float val = tap_value(t, v1, v2, v3, v4);
        const float wgt_h = -t.hw * v1 - t.lw * v2 + t.hw * v3 + t.lw * v4;
        const float wgt_w = -t.hh * v1 + t.hh * v2 - t.lh * v3 + t.lh * v4;
        float gm = t.inside ? go * val : 0.f;
        float gh = t.inside ? wgt_h * go * m : 0.f;
        float gw = t.inside ? wgt_w * go * m : 0.f;
        if (accumulate) {
            gm += gaf[(long long)k * plane];
            gh += gof[(long long)(2 * k) * plane];
            gw += gof[(long long)(2 * k + 1) * plane];
        }
        gaf[(long long)k * plane] = gm;
        gof[(long long)(2 * k) * plane] = gh;
        gof[(long long)(2 * k + 1) * plane] = gw;
        if (go != 0.f) {
            const float top = go * m;
            float* p = gimg + (long long)t.h_low * W + t.w_low;
            if (t.ok1) atomicAdd(p, t.hh * t.hw * top);
            if (t.ok2) atomicAdd(p + 1, t.hh * t.lw * top);
            if (t.ok3) atomicAdd(p + W, t.lh * t.hw * top);
            if (t.ok4) atomicAdd(p + W + 1, t.lh * t.lw * top);
        }
    }
}

// grad wrt feat_init = [fix <= 0] * g   (backward of the first blend)
__global__ void prop_mask_grad_kernel(const float* __restrict__ g, const float* __restrict__ fix, float* __restrict__ out, long long total) {
    PDL_SYNC();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    out[i] = (fix && __ldg(fix + i) > 0.f) ? 0.f : g[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// a19: NLSPN._get_offset_affinity (nlspnmodel_adapt.py:255-330), affinity 'TGASS', fused into one pass per direction.
//   offset_aff [N,24,H,W] = conv_offset_aff(guidance): channels 0-15 are the 8 neighbours' (dh, dw) pairs in the order the
//   reference's cat(o1,o2).view(B,8,2,H,W) produces (neighbour j = channels 2j, 2j+1), channels 16-23 the raw affinities.
//   b_j = tanh(x_j) / (scale + 1e-8); u_j = b_j * conf_j with conf_j = the confidence sampled (1x1 DCN, pad 0, window
//   (-1,H)x(-1,W)) at p + offset_j [+ grid_j when `legacy`]; S = max(sum|u| + 1e-4, 1); a_j = u_j / S; a_ref = 1 - sum a_j.
//   Outputs: offset [N,18,H,W] (zero pair inserted at tap 4), aff [N,9,H,W] (a_ref at tap 4).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void offaff_conf_position(int j, int y, int x, float dh, float dw, int legacy, float& h_im, float& w_im) {
    const int tap = j < 4 ? j : j + 1;
    h_im = (float)y + dh + (legacy ? (float)(tap / 3 - 1) : 0.f);
    w_im = (float)x + dw + (legacy ? (float)(tap % 3 - 1) : 0.f);
}

__global__ void __launch_bounds__(PROP_TX * PROP_TY) offset_affinity_forward_kernel(const float* __restrict__ offset_aff,
                                                                                    const float* __restrict__ confidence, float inv_scale,
                                                                                    int legacy, float* __restrict__ offset, float* __restrict__ aff,
                                                                                    int H, int W) {
    PDL_SYNC();
    const int x = blockIdx.x * PROP_TX + threadIdx.x, y = blockIdx.y * PROP_TY + threadIdx.y, n = blockIdx.z;
    if (x >= W || y >= H) return;
    const long long plane = (long long)H * W, pix = (long long)y * W + x;
    const float* oa = offset_aff + (long long)n * 24 * plane + pix;
    const float* conf = confidence ? confidence + (long long)n * plane : nullptr;
    float* of = offset + (long long)n * 18 * plane + pix;
    float* af = aff + (long long)n * 9 * plane + pix;
    float u[8], sum_abs = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int tap = j < 4 ? j : j + 1;
        const float dh = __ldg(oa + (long long)(2 * j) * plane), dw = __ldg(oa + (long long)(2 * j + 1) * plane);
        of[(long long)(2 * tap) * plane] = dh;
        of[(long long)(2 * tap + 1) * plane] = dw;
        float v = tanhf(__ldg(oa + (long long)(16 + j) * plane)) * inv_scale;
        if (conf) {
            float h_im, w_im;
            offaff_conf_position(j, y, x, dh, dw, legacy, h_im, w_im);
            const BilinearTap t = make_tap(h_im, w_im, H, W);
            float v1, v2, v3, v4;
            tap_corners(t, conf, W, v1, v2, v3, v4);
            v *= tap_value(t, v1, v2, v3, v4);
        }
        u[j] = v;
        sum_abs += fabsf(v);
    }
    of[8 * plane] = 0.f;
    of[9 * plane] = 0.f;
    float S = sum_abs + 1e-4f;
    if (S < 1.f) S = 1.f;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float a = u[j] / S;
        af[(long long)(j < 4 ? j : j + 1) * plane] = a;
        sum += a;
    }
    af[4 * plane] = 1.f - sum;
}

// g_offset_aff [N,24,H,W] written; g_confidence [N,H,W] (zero-filled by the caller) accumulated with fp32 reductions.
// The offsets used by the confidence gather are detached in the reference (:293), so they receive g_offset only.
__global__ void __launch_bounds__(PROP_TX * PROP_TY) offset_affinity_backward_kernel(const float* __restrict__ offset_aff,
                                                                                     const float* __restrict__ confidence, float inv_scale,
                                                                                     int legacy, const float* __restrict__ g_offset,
                                                                                     const float* __restrict__ g_aff, float* __restrict__ g_offset_aff,
                                                                                     float* __restrict__ g_confidence, int H, int W) {
    PDL_SYNC();
    const int x = blockIdx.x * PROP_TX + threadIdx.x, y = blockIdx.y * PROP_TY + threadIdx.y, n = blockIdx.z;
    if (x >= W || y >= H) return;
    const long long plane = (long long)H * W, pix = (long long)y * W + x;
    const float* oa = offset_aff + (long long)n * 24 * plane + pix;
    const float* conf = confidence ? confidence + (long long)n * plane : nullptr;
    const float* go = g_offset + (long long)n * 18 * plane + pix;
    const float* ga = g_aff + (long long)n * 9 * plane + pix;
    float* goa = g_offset_aff + (long long)n * 24 * plane + pix;
    float* gconf = g_confidence ? g_confidence + (long long)n * plane : nullptr;
    float b[8], c[8], th[8], u[8], sum_abs = 0.f;
    BilinearTap taps[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int tap = j < 4 ? j : j + 1;
        const float dh = __ldg(oa + (long long)(2 * j) * plane), dw = __ldg(oa + (long long)(2 * j + 1) * plane);
        goa[(long long)(2 * j) * plane] = __ldg(go + (long long)(2 * tap) * plane);
        goa[(long long)(2 * j + 1) * plane] = __ldg(go + (long long)(2 * tap + 1) * plane);
        th[j] = tanhf(__ldg(oa + (long long)(16 + j) * plane));
        b[j] = th[j] * inv_scale;
        c[j] = 1.f;
        if (conf) {
            float h_im, w_im;
            offaff_conf_position(j, y, x, dh, dw, legacy, h_im, w_im);
            taps[j] = make_tap(h_im, w_im, H, W);
            float v1, v2, v3, v4;
            tap_corners(taps[j], conf, W, v1, v2, v3, v4);
            c[j] = tap_value(taps[j], v1, v2, v3, v4);
        }
        u[j] = b[j] * c[j];
        sum_abs += fabsf(u[j]);
    }
    const float S0 = sum_abs + 1e-4f;
    const bool clamped = S0 < 1.f;
    const float S = clamped ? 1.f : S0;
    const float g_ref = __ldg(ga + 4 * plane);
    float gA[8], dot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        gA[j] = __ldg(ga + (long long)(j < 4 ? j : j + 1) * plane) - g_ref;      // a_ref = 1 - sum a_j
        dot += gA[j] * u[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float gu = gA[j] / S;
        if (!clamped) {
            const float sgn = u[j] > 0.f ? 1.f : (u[j] < 0.f ? -1.f : 0.f);
            gu -= sgn * dot / (S * S);
        }
        goa[(long long)(16 + j) * plane] = gu * c[j] * (1.f - th[j] * th[j]) * inv_scale;
        if (gconf) {
            const float top = gu * b[j];
            const BilinearTap& t = taps[j];
            float* p = gconf + (long long)t.h_low * W + t.w_low;
            if (t.ok1) atomicAdd(p, t.hh * t.hw * top);
            if (t.ok2) atomicAdd(p + 1, t.hh * t.lw * top);
            if (t.ok3) atomicAdd(p + W, t.lh * t.hw * top);
            if (t.ok4) atomicAdd(p + W + 1, t.lh * t.lw * top);
        }
    }
}

}  // namespace ptta
