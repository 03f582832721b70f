// fsgs_kernels_loss.cuh -- the image loss Free-SurGS applies to every rendered frame, fused:
//   rgb_loss_func(img, gt, lambda, mask) = (1 - lambda) * mean|x - y| + lambda * (1 - mean SSIM(x, y)),
//   x = img * mask, y = gt * mask          (reference utils/loss_utils.py:47-96)
// with the reference's SSIM: 11x11 Gaussian window (sigma 1.5, outer product of the normalised 1-D window), zero
// padding, C1 = 0.01^2, C2 = 0.03^2, mean over all C*H*W elements.
//
// The PyTorch formulation runs five depthwise 11x11 convolutions forward (and their transposes backward) plus a
// dozen element-wise kernels per call -- several milliseconds at 1280x1024, i.e. more than the whole render.
// Here one forward kernel and one backward kernel do it, both tiled 16x16 with a 5-pixel halo staged in shared
// memory and the window applied separably (11 + 11 taps instead of 121):
//
//   forward : a = G*x, b = G*y, p = G*x^2, q = G*y^2, r = G*xy per pixel ->
//               A1 = 2ab + C1, A2 = 2(r - ab) + C2, B1 = a^2 + b^2 + C1, B2 = (p - a^2) + (q - b^2) + C2,
//               S = A1 A2 / (B1 B2);
//             block sums of S and |x - y| (double) -> per-block partials -> k_rgb_loss_reduce (deterministic);
//             when a gradient is wanted, the three per-pixel partial derivatives everything else is linear in:
//               Da = dS/da, Dp = dS/dp, Dr = dS/dr.
//   backward: dL/dx = (1-lambda)/N sign(x - y) - lambda/N [ G*Da + 2x G*Dp + y G*Dr ]   (G symmetric, zero padding
//             on both sides, so the transpose of the forward convolution is the same convolution of the maps),
//             dL/dimg = mask * dL/dx * upstream.
//
// HBM traffic per call at C = 3, 1280x1024: forward reads 2 planes-sets (+ mask) and writes 3 map sets = 79 MB,
// backward reads 5 and writes 1 = 94 MB.
#pragma once

#include "fsgs_device.cuh"

namespace fsgs {

constexpr int LOSS_T = 16;             // output tile edge
constexpr int LOSS_R = 5;              // window radius (11 taps)
constexpr int LOSS_H = LOSS_T + 2 * LOSS_R;   // halo tile edge (26)
constexpr int LOSS_HS = 48;            // row stride of the halo tiles: two consecutive rows 16 banks apart
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

struct LossWindow {
    float g[2 * LOSS_R + 1];
};

struct LossMask {
    const unsigned char *u8;    // bool mask (one byte per element) or NULL
    const float *f32;           // float mask or NULL
    long long cstride;          // 0: one [H,W] plane for all channels; H*W: a [C,H,W] mask
};
__device__ __forceinline__ float loss_mask_at(const LossMask &m, int c, size_t pix) {
    const size_t k = (size_t)c * (size_t)m.cstride + pix;
    if (m.u8) return m.u8[k] ? 1.f : 0.f;
    if (m.f32) return m.f32[k];
    return 1.f;
}

// Forward.  grid (ceil(W/16), ceil(H/16), C), 256 threads.  partial[2 * block] = (sum S, sum |x-y|) of the block.
__global__ void __launch_bounds__(CTA)
k_rgb_loss_fwd(int H, int W, const float *__restrict__ img, const float *__restrict__ gt, LossMask mask, LossWindow win,
               float *__restrict__ maps, double *__restrict__ partial) {
    __shared__ float sx[LOSS_H][LOSS_HS], sy[LOSS_H][LOSS_HS];
    __shared__ float sh[5][LOSS_H][LOSS_T + 1];
    __shared__ double s_red[2][CTA / 32];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * LOSS_T, y0 = blockIdx.y * LOSS_T;
    const size_t HW = (size_t)H * W;
    const float *ic = img + (size_t)c * HW, *gc = gt + (size_t)c * HW;
    for (int e = threadIdx.x; e < LOSS_H * LOSS_H; e += CTA) {
        const int ry = e / LOSS_H, rx = e - ry * LOSS_H;
        const int py = y0 + ry - LOSS_R, px = x0 + rx - LOSS_R;
        float vx = 0.f, vy = 0.f;
        if (py >= 0 && py < H && px >= 0 && px < W) {
            const size_t pix = (size_t)py * W + px;
            const float m = loss_mask_at(mask, c, pix);
            vx = ic[pix] * m; vy = gc[pix] * m;
        }
        sx[ry][rx] = vx; sy[ry][rx] = vy;
    }
    __syncthreads();
    // horizontal pass: LOSS_H rows x LOSS_T columns x 5 quantities
    for (int e = threadIdx.x; e < LOSS_H * LOSS_T; e += CTA) {
        const int ry = e / LOSS_T, cx = e - ry * LOSS_T;
        float a = 0.f, b = 0.f, p = 0.f, q = 0.f, r = 0.f;
#pragma unroll
        for (int t = 0; t < 2 * LOSS_R + 1; ++t) {
            const float w = win.g[t], vx = sx[ry][cx + t], vy = sy[ry][cx + t];
            const float wx = w * vx, wy = w * vy;
            a += wx; b += wy; p = fmaf(wx, vx, p); q = fmaf(wy, vy, q); r = fmaf(wx, vy, r);
        }
        sh[0][ry][cx] = a; sh[1][ry][cx] = b; sh[2][ry][cx] = p; sh[3][ry][cx] = q; sh[4][ry][cx] = r;
    }
    __syncthreads();
    const int tx = threadIdx.x & (LOSS_T - 1), ty = threadIdx.x >> 4;
    const int px = x0 + tx, py = y0 + ty;
    double sumS = 0.0, sumL1 = 0.0;
    if (px < W && py < H) {
        float a = 0.f, b = 0.f, p = 0.f, q = 0.f, r = 0.f;
#pragma unroll
        for (int t = 0; t < 2 * LOSS_R + 1; ++t) {
            const float w = win.g[t];
            a = fmaf(w, sh[0][ty + t][tx], a); b = fmaf(w, sh[1][ty + t][tx], b); p = fmaf(w, sh[2][ty + t][tx], p);
            q = fmaf(w, sh[3][ty + t][tx], q); r = fmaf(w, sh[4][ty + t][tx], r);
        }
        const float ab = a * b, aa = a * a, bb = b * b;
        const float A1 = 2.f * ab + SSIM_C1, A2 = 2.f * (r - ab) + SSIM_C2;
        const float B1 = aa + bb + SSIM_C1, B2 = (p - aa) + (q - bb) + SSIM_C2;
        const float iB1 = 1.f / B1, iB2 = 1.f / B2;
        const float S = A1 * A2 * iB1 * iB2;
        sumS = (double)S;
        sumL1 = (double)fabsf(sx[ty + LOSS_R][tx + LOSS_R] - sy[ty + LOSS_R][tx + LOSS_R]);
        if (maps) {
            // dS/da with sigma's expanded: d(A1 A2)/da = 2b (A2 - A1), d(B1 B2)/da = 2a (B2 - B1)
            const float Da = (2.f * b * (A2 - A1) - S * 2.f * a * (B2 - B1)) * iB1 * iB2;
            const float Dp = -S * iB2;
            const float Dr = 2.f * A1 * iB1 * iB2;
            const size_t CHW = (size_t)gridDim.z * HW, o = (size_t)c * HW + (size_t)py * W + px;
            maps[o] = Da; maps[CHW + o] = Dp; maps[2 * CHW + o] = Dr;
        }
    }
    // block sums (fixed order: lanes by shuffle tree, warps serially -> deterministic)
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        sumS += __shfl_xor_sync(FULL, sumS, d);
        sumL1 += __shfl_xor_sync(FULL, sumL1, d);
    }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = sumS; s_red[1][threadIdx.x >> 5] = sumL1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0, l = 0.0;
        for (int w = 0; w < CTA / 32; ++w) { s += s_red[0][w]; l += s_red[1][w]; }
        const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partial[2 * blk] = s; partial[2 * blk + 1] = l;
    }
}

// out[0] = loss, out[1] = mean |x - y|, out[2] = mean SSIM.  One CTA, fixed summation order.
__global__ void __launch_bounds__(CTA)
k_rgb_loss_reduce(long long n_blocks, const double *__restrict__ partial, double inv_n, float lambda_dssim,
                  float *__restrict__ out) {
    __shared__ double s_red[2][CTA / 32];
    double s = 0.0, l = 0.0;
    for (long long k = threadIdx.x; k < n_blocks; k += CTA) { s += partial[2 * k]; l += partial[2 * k + 1]; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        s += __shfl_xor_sync(FULL, s, d);
        l += __shfl_xor_sync(FULL, l, d);
    }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = s; s_red[1][threadIdx.x >> 5] = l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = 0.0; l = 0.0;
        for (int w = 0; w < CTA / 32; ++w) { s += s_red[0][w]; l += s_red[1][w]; }
        const double ssim = s * inv_n, l1 = l * inv_n;
        out[0] = (float)((1.0 - (double)lambda_dssim) * l1 + (double)lambda_dssim * (1.0 - ssim));
        out[1] = (float)l1;
        out[2] = (float)ssim;
    }
}

// Backward.  Same grid.  upstream = d(total loss)/d(this loss), a device scalar (NULL = 1).
__global__ void __launch_bounds__(CTA)
k_rgb_loss_bwd(int H, int W, const float *__restrict__ img, const float *__restrict__ gt, LossMask mask, LossWindow win,
               const float *__restrict__ maps, const float *__restrict__ upstream, float lambda_dssim, float inv_n,
               float *__restrict__ dimg) {
    __shared__ float sm[3][LOSS_H][LOSS_HS];
    __shared__ float sh[3][LOSS_H][LOSS_T + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * LOSS_T, y0 = blockIdx.y * LOSS_T;
    const size_t HW = (size_t)H * W, CHW = (size_t)gridDim.z * HW;
    const float *mc = maps + (size_t)c * HW;
    for (int e = threadIdx.x; e < LOSS_H * LOSS_H; e += CTA) {
        const int ry = e / LOSS_H, rx = e - ry * LOSS_H;
        const int py = y0 + ry - LOSS_R, px = x0 + rx - LOSS_R;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (py >= 0 && py < H && px >= 0 && px < W) {
            const size_t pix = (size_t)py * W + px;
            v0 = mc[pix]; v1 = mc[CHW + pix]; v2 = mc[2 * CHW + pix];
        }
        sm[0][ry][rx] = v0; sm[1][ry][rx] = v1; sm[2][ry][rx] = v2;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < LOSS_H * LOSS_T; e += CTA) {
        const int ry = e / LOSS_T, cx = e - ry * LOSS_T;
        float a = 0.f, p = 0.f, r = 0.f;
#pragma unroll
        for (int t = 0; t < 2 * LOSS_R + 1; ++t) {
            const float w = win.g[t];
            a = fmaf(w, sm[0][ry][cx + t], a); p = fmaf(w, sm[1][ry][cx + t], p); r = fmaf(w, sm[2][ry][cx + t], r);
        }
        sh[0][ry][cx] = a; sh[1][ry][cx] = p; sh[2][ry][cx] = r;
    }
    __syncthreads();
    const int tx = threadIdx.x & (LOSS_T - 1), ty = threadIdx.x >> 4;
    const int px = x0 + tx, py = y0 + ty;
    if (px >= W || py >= H) return;
    float ga = 0.f, gp = 0.f, gr = 0.f;
#pragma unroll
    for (int t = 0; t < 2 * LOSS_R + 1; ++t) {
        const float w = win.g[t];
        ga = fmaf(w, sh[0][ty + t][tx], ga); gp = fmaf(w, sh[1][ty + t][tx], gp); gr = fmaf(w, sh[2][ty + t][tx], gr);
    }
    const size_t pix = (size_t)py * W + px, o = (size_t)c * HW + pix;
    const float m = loss_mask_at(mask, c, pix);
    const float x = img[o] * m, y = gt[o] * m;
    const float d = x - y;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    const float up = upstream ? __ldg(upstream) : 1.f;
    const float gx = (1.f - lambda_dssim) * inv_n * sgn - lambda_dssim * inv_n * (ga + 2.f * x * gp + y * gr);
    dimg[o] = m * gx * up;
}

}  // namespace fsgs
