// fsgs_kernels_loss.cuh -- the image loss Free-SurGS applies to every rendered frame, fused:
//   rgb_loss_func(img, gt, lambda, mask) = (1 - lambda) * mean|x - y| + lambda * (1 - mean SSIM(x, y)),
//   x = img * mask, y = gt * mask          (reference utils/loss_utils.py:47-96)
// with the reference's SSIM: 11x11 Gaussian window (sigma 1.5, outer product of the normalised 1-D window), zero
// padding, C1 = 0.01^2, C2 = 0.03^2, mean over all C*H*W elements.
//
// The PyTorch formulation runs five depthwise 11x11 convolutions forward (and their transposes backward) plus a
// dozen element-wise kernels per call -- several milliseconds at 1280x1024, i.e. more than the whole render.
// Here one forward kernel and one backward kernel do it, both tiled 32x16 with a 5-pixel halo staged in shared
// memory and the window applied separably (11 + 11 taps instead of 121), each thread producing 4 adjacent outputs of
// the horizontal pass (128-bit shared-memory loads, products formed once per input) and 2 of the vertical pass:
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

constexpr int LOSS_TW = 32, LOSS_TH = 16;      // output tile
constexpr int LOSS_R = 5;                      // window radius (11 taps)
constexpr int LOSS_K = 2 * LOSS_R + 1;
constexpr int LOSS_HW = LOSS_TW + 2 * LOSS_R;  // halo tile width  (42)
constexpr int LOSS_HH = LOSS_TH + 2 * LOSS_R;  // halo tile height (26)
constexpr int LOSS_HS = 48;                    // row stride of the halo tiles (16-byte aligned rows, >= 4*7+16 = 44)
constexpr int LOSS_ITEMS = LOSS_HH * (LOSS_TW / 4);   // horizontal-pass work items: (halo row, group of 4 columns)
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;
static_assert(LOSS_ITEMS <= CTA && LOSS_TW * LOSS_TH == 2 * CTA, "thread mapping of the loss kernels");

struct LossWindow {
    float g[LOSS_K];
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

// 16 consecutive floats of a halo row starting at a multiple-of-4 column (four 128-bit loads)
__device__ __forceinline__ void loss_load16(const float *row, float *v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 q = *reinterpret_cast<const float4 *>(row + 4 * k);
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
}

// Forward.  grid (ceil(W/32), ceil(H/16), C), 256 threads.  partial[2 * block] = (sum S, sum |x-y|) of the block.
__global__ void __launch_bounds__(CTA)
k_rgb_loss_fwd(int H, int W, const float *__restrict__ img, const float *__restrict__ gt, LossMask mask, LossWindow win,
               float *__restrict__ maps, double *__restrict__ partial) {
    __shared__ __align__(16) float sx[LOSS_HH][LOSS_HS], sy[LOSS_HH][LOSS_HS];
    __shared__ __align__(16) float sh[5][LOSS_HH][LOSS_TW];
    __shared__ double s_red[2][CTA / 32];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * LOSS_TW, y0 = blockIdx.y * LOSS_TH;
    const size_t HW = (size_t)H * W;
    const float *ic = img + (size_t)c * HW, *gc = gt + (size_t)c * HW;
    for (int e = threadIdx.x; e < LOSS_HH * LOSS_HS; e += CTA) {
        const int ry = e / LOSS_HS, rx = e - ry * LOSS_HS;
        const int py = y0 + ry - LOSS_R, px = x0 + rx - LOSS_R;
        float vx = 0.f, vy = 0.f;
        if (rx < LOSS_HW && py >= 0 && py < H && px >= 0 && px < W) {
            const size_t pix = (size_t)py * W + px;
            const float m = loss_mask_at(mask, c, pix);
            vx = ic[pix] * m; vy = gc[pix] * m;
        }
        sx[ry][rx] = vx; sy[ry][rx] = vy;      // (columns 42..47 are zero padding the 128-bit loads may touch)
    }
    __syncthreads();
    // horizontal pass: one (halo row, 4 adjacent columns) item per thread
    if (threadIdx.x < LOSS_ITEMS) {
        const int ry = threadIdx.x >> 3, cg = (threadIdx.x & 7) * 4;
        float vx[16], vy[16];
        loss_load16(&sx[ry][cg], vx);
        loss_load16(&sy[ry][cg], vy);
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, p[4] = {0.f, 0.f, 0.f, 0.f},
              q[4] = {0.f, 0.f, 0.f, 0.f}, r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < LOSS_K + 3; ++i) {
            const float xx = vx[i] * vx[i], yy = vy[i] * vy[i], xy = vx[i] * vy[i];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = i - j;                  // tap index of input i for output j
                if (t >= 0 && t < LOSS_K) {
                    const float w = win.g[t];
                    a[j] = fmaf(w, vx[i], a[j]); b[j] = fmaf(w, vy[i], b[j]); p[j] = fmaf(w, xx, p[j]);
                    q[j] = fmaf(w, yy, q[j]); r[j] = fmaf(w, xy, r[j]);
                }
            }
        }
        *reinterpret_cast<float4 *>(&sh[0][ry][cg]) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4 *>(&sh[1][ry][cg]) = make_float4(b[0], b[1], b[2], b[3]);
        *reinterpret_cast<float4 *>(&sh[2][ry][cg]) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4 *>(&sh[3][ry][cg]) = make_float4(q[0], q[1], q[2], q[3]);
        *reinterpret_cast<float4 *>(&sh[4][ry][cg]) = make_float4(r[0], r[1], r[2], r[3]);
    }
    __syncthreads();
    // vertical pass: column tx, output rows 2*tg and 2*tg + 1
    const int tx = threadIdx.x & 31, tg = threadIdx.x >> 5;
    const int px = x0 + tx;
    float acc[2][5];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[j][k] = 0.f;
#pragma unroll
    for (int i = 0; i < LOSS_K + 1; ++i) {
        float v[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = sh[k][2 * tg + i][tx];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int t = i - j;
            if (t >= 0 && t < LOSS_K) {
                const float w = win.g[t];
#pragma unroll
                for (int k = 0; k < 5; ++k) acc[j][k] = fmaf(w, v[k], acc[j][k]);
            }
        }
    }
    double sumS = 0.0, sumL1 = 0.0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int ty = 2 * tg + j, py = y0 + ty;
        if (px < W && py < H) {
            const float a = acc[j][0], b = acc[j][1], p = acc[j][2], q = acc[j][3], r = acc[j][4];
            const float ab = a * b, aa = a * a, bb = b * b;
            const float A1 = 2.f * ab + SSIM_C1, A2 = 2.f * (r - ab) + SSIM_C2;
            const float B1 = aa + bb + SSIM_C1, B2 = (p - aa) + (q - bb) + SSIM_C2;
            const float iB1 = 1.f / B1, iB2 = 1.f / B2;
            const float S = A1 * A2 * iB1 * iB2;
            sumS += (double)S;
            sumL1 += (double)fabsf(sx[ty + LOSS_R][tx + LOSS_R] - sy[ty + LOSS_R][tx + LOSS_R]);
            if (maps) {
                // dS/da with sigma's expanded: d(A1 A2)/da = 2b (A2 - A1), d(B1 B2)/da = 2a (B2 - B1)
                const float Da = (2.f * b * (A2 - A1) - S * 2.f * a * (B2 - B1)) * iB1 * iB2;
                const float Dp = -S * iB2;
                const float Dr = 2.f * A1 * iB1 * iB2;
                const size_t CHW = (size_t)gridDim.z * HW, o = (size_t)c * HW + (size_t)py * W + px;
                maps[o] = Da; maps[CHW + o] = Dp; maps[2 * CHW + o] = Dr;
            }
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
    __shared__ __align__(16) float sm[3][LOSS_HH][LOSS_HS];
    __shared__ __align__(16) float sh[3][LOSS_HH][LOSS_TW];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * LOSS_TW, y0 = blockIdx.y * LOSS_TH;
    const size_t HW = (size_t)H * W, CHW = (size_t)gridDim.z * HW;
    const float *mc = maps + (size_t)c * HW;
    for (int e = threadIdx.x; e < LOSS_HH * LOSS_HS; e += CTA) {
        const int ry = e / LOSS_HS, rx = e - ry * LOSS_HS;
        const int py = y0 + ry - LOSS_R, px = x0 + rx - LOSS_R;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (rx < LOSS_HW && py >= 0 && py < H && px >= 0 && px < W) {
            const size_t pix = (size_t)py * W + px;
            v0 = mc[pix]; v1 = mc[CHW + pix]; v2 = mc[2 * CHW + pix];
        }
        sm[0][ry][rx] = v0; sm[1][ry][rx] = v1; sm[2][ry][rx] = v2;
    }
    __syncthreads();
    if (threadIdx.x < LOSS_ITEMS) {
        const int ry = threadIdx.x >> 3, cg = (threadIdx.x & 7) * 4;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v[16], o[4] = {0.f, 0.f, 0.f, 0.f};
            loss_load16(&sm[k][ry][cg], v);
#pragma unroll
            for (int i = 0; i < LOSS_K + 3; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = i - j;
                    if (t >= 0 && t < LOSS_K) o[j] = fmaf(win.g[t], v[i], o[j]);
                }
            *reinterpret_cast<float4 *>(&sh[k][ry][cg]) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, tg = threadIdx.x >> 5;
    const int px = x0 + tx;
    float acc[2][3];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[j][k] = 0.f;
#pragma unroll
    for (int i = 0; i < LOSS_K + 1; ++i) {
        float v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = sh[k][2 * tg + i][tx];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int t = i - j;
            if (t >= 0 && t < LOSS_K) {
                const float w = win.g[t];
#pragma unroll
                for (int k = 0; k < 3; ++k) acc[j][k] = fmaf(w, v[k], acc[j][k]);
            }
        }
    }
    const float up = upstream ? __ldg(upstream) : 1.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int py = y0 + 2 * tg + j;
        if (px >= W || py >= H) continue;
        const size_t pix = (size_t)py * W + px, o = (size_t)c * HW + pix;
        const float m = loss_mask_at(mask, c, pix);
        const float x = img[o] * m, y = gt[o] * m;
        const float d = x - y;
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        const float gx = (1.f - lambda_dssim) * inv_n * sgn -
                         lambda_dssim * inv_n * (acc[j][0] + 2.f * x * acc[j][1] + y * acc[j][2]);
        dimg[o] = m * gx * up;
    }
}

// ---- Pearson depth loss (reference utils/loss_utils.py:98-109) ------------------------------------------
//   loss = 1 - mean( (x - mean x) / (std x + 1e-6) * (y - mean y) / (std y + 1e-6) ),  std = unbiased (n - 1)
// The PyTorch formulation is ~12 element-wise / reduction launches forward and ~20 backward over the depth plane.
// Here: one pass accumulates the five raw sums (double), a one-CTA kernel turns them into the statistics
//   stats = (mean x, mean y, sqrt var x, sqrt var y, c = sum (x - mx)(y - my), n)  and the loss,
// and the backward is one element-wise kernel:
//   dL/dy_i = -[ (x_i - mx) / (n sx sy) - c / (n sx sy^2) * (y_i - my) / ((n - 1) sqrt var y) ],  s = sqrt var + 1e-6
// (symmetric for x).  Fixed grid and summation order: deterministic.
constexpr int PEARSON_MAX_BLOCKS = 1184;     // 148 SMs x 8

__global__ void __launch_bounds__(CTA)
k_pearson_sums(long long n, const float *__restrict__ x, const float *__restrict__ y, double *__restrict__ partial) {
    __shared__ double s_red[5][CTA / 32];
    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * CTA + threadIdx.x; i < n; i += (long long)gridDim.x * CTA) {
        const double xv = (double)x[i], yv = (double)y[i];
        a[0] += xv; a[1] += yv; a[2] += xv * xv; a[3] += yv * yv; a[4] += xv * yv;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) a[k] += __shfl_xor_sync(FULL, a[k], d);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = a[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int w = 0; w < CTA / 32; ++w) t += s_red[threadIdx.x][w];
        partial[5 * (size_t)blockIdx.x + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32)
k_pearson_finish(long long n, int n_blocks, const double *__restrict__ partial, double *__restrict__ stats,
                 float *__restrict__ out) {
    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < n_blocks; b += 32)
#pragma unroll
        for (int k = 0; k < 5; ++k) a[k] += partial[5 * (size_t)b + k];
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) a[k] += __shfl_xor_sync(FULL, a[k], d);
    if (threadIdx.x == 0) {
        const double N = (double)n, mx = a[0] / N, my = a[1] / N;
        const double den = N > 1.0 ? N - 1.0 : 1.0;
        const double vx = fmax(a[2] - N * mx * mx, 0.0) / den, vy = fmax(a[3] - N * my * my, 0.0) / den;
        const double qx = sqrt(vx), qy = sqrt(vy);
        const double c = a[4] - N * mx * my;
        stats[0] = mx; stats[1] = my; stats[2] = qx; stats[3] = qy; stats[4] = c; stats[5] = N;
        out[0] = (float)(1.0 - c / (N * (qx + 1e-6) * (qy + 1e-6)));
    }
}

__global__ void __launch_bounds__(CTA)
k_pearson_bwd(long long n, const float *__restrict__ x, const float *__restrict__ y, const double *__restrict__ stats,
              const float *__restrict__ upstream, float *__restrict__ dx, float *__restrict__ dy) {
    const double mx = stats[0], my = stats[1], qx = stats[2], qy = stats[3], c = stats[4], N = stats[5];
    const double sx = qx + 1e-6, sy = qy + 1e-6, den = N > 1.0 ? N - 1.0 : 1.0;
    const double up = upstream ? (double)__ldg(upstream) : 1.0;
    // dL/dy_i = -(k1 (x_i - mx) - ky (y_i - my)),  dL/dx_i = -(k1 (y_i - my) - kx (x_i - mx))
    const float k1 = (float)(up / (N * sx * sy));
    const float ky = qy > 0.0 ? (float)(up * c / (N * sx * sy * sy * den * qy)) : 0.f;
    const float kx = qx > 0.0 ? (float)(up * c / (N * sx * sx * sy * den * qx)) : 0.f;
    for (long long i = (long long)blockIdx.x * CTA + threadIdx.x; i < n; i += (long long)gridDim.x * CTA) {
        // (centre in double: x - mean cancels leading digits)
        const float xc = (float)((double)x[i] - mx), yc = (float)((double)y[i] - my);
        if (dy) dy[i] = -(k1 * xc - ky * yc);
        if (dx) dx[i] = -(k1 * yc - kx * xc);
    }
}

// ---- local Pearson loss (reference utils/loss_utils.py:112-127) ------------------------------------------------
//   n random box x box patches (top-left corners (x0[i], y0[i]) = (row, column), drawn by the caller exactly as the
//   reference draws them); loss = (1/n) sum_i pearson_depth_loss(src[patch_i], target[patch_i]).
// The reference loops over the patches in Python, slicing with device-tensor indices (2 host syncs per patch, ~25
// launches each).  Here: one launch accumulates the five raw sums of every patch (LP_BLOCKS CTAs per patch, double),
// one small launch turns them into per-patch statistics + the loss, and the backward is one launch whose threads
// add the per-patch Pearson derivative into the (overlapping) patches with float atomics.
constexpr int LP_BLOCKS = 8;

__global__ void __launch_bounds__(CTA)
k_local_pearson_sums(int W, int box, const long long *__restrict__ x0, const long long *__restrict__ y0,
                     const float *__restrict__ x, const float *__restrict__ y, double *__restrict__ partial) {
    __shared__ double s_red[5][CTA / 32];
    const int patch = blockIdx.y;
    const long long r0 = x0[patch], c0 = y0[patch];
    const int n = box * box;
    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int e = blockIdx.x * CTA + threadIdx.x; e < n; e += LP_BLOCKS * CTA) {
        const size_t p = (size_t)(r0 + e / box) * W + (size_t)(c0 + e % box);
        const double xv = (double)x[p], yv = (double)y[p];
        a[0] += xv; a[1] += yv; a[2] += xv * xv; a[3] += yv * yv; a[4] += xv * yv;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) a[k] += __shfl_xor_sync(FULL, a[k], d);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = a[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int w = 0; w < CTA / 32; ++w) t += s_red[threadIdx.x][w];
        partial[5 * ((size_t)patch * LP_BLOCKS + blockIdx.x) + threadIdx.x] = t;
    }
}

// stats[i] = (mean x, mean y, sqrt var x, sqrt var y, c, n) per patch; out = mean over the patches of 1 - corr_i
__global__ void __launch_bounds__(CTA)
k_local_pearson_finish(int n_patches, int box, const double *__restrict__ partial, double *__restrict__ stats,
                       float *__restrict__ out) {
    __shared__ double s_loss[CTA];
    double local = 0.0;
    for (int i = threadIdx.x; i < n_patches; i += CTA) {
        double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        for (int b = 0; b < LP_BLOCKS; ++b)
#pragma unroll
            for (int k = 0; k < 5; ++k) a[k] += partial[5 * ((size_t)i * LP_BLOCKS + b) + k];
        const double N = (double)box * box, mx = a[0] / N, my = a[1] / N;
        const double den = N > 1.0 ? N - 1.0 : 1.0;
        const double vx = fmax(a[2] - N * mx * mx, 0.0) / den, vy = fmax(a[3] - N * my * my, 0.0) / den;
        const double qx = sqrt(vx), qy = sqrt(vy), c = a[4] - N * mx * my;
        double *st = stats + 6 * (size_t)i;
        st[0] = mx; st[1] = my; st[2] = qx; st[3] = qy; st[4] = c; st[5] = N;
        local += 1.0 - c / (N * (qx + 1e-6) * (qy + 1e-6));
    }
    s_loss[threadIdx.x] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < CTA; ++k) t += s_loss[k];          // fixed order
        out[0] = (float)(t / (double)n_patches);
    }
}

__global__ void __launch_bounds__(CTA)
k_local_pearson_bwd(int W, int box, int n_patches, const long long *__restrict__ x0, const long long *__restrict__ y0,
                    const float *__restrict__ x, const float *__restrict__ y, const double *__restrict__ stats,
                    const float *__restrict__ upstream, float *__restrict__ dx, float *__restrict__ dy) {
    const int patch = blockIdx.y;
    const double *st = stats + 6 * (size_t)patch;
    const double mx = st[0], my = st[1], qx = st[2], qy = st[3], c = st[4], N = st[5];
    const double sx = qx + 1e-6, sy = qy + 1e-6, den = N > 1.0 ? N - 1.0 : 1.0;
    const double up = (upstream ? (double)__ldg(upstream) : 1.0) / (double)n_patches;
    const float k1 = (float)(up / (N * sx * sy));
    const float ky = qy > 0.0 ? (float)(up * c / (N * sx * sy * sy * den * qy)) : 0.f;
    const float kx = qx > 0.0 ? (float)(up * c / (N * sx * sx * sy * den * qx)) : 0.f;
    const long long r0 = x0[patch], c0 = y0[patch];
    const int n = box * box;
    for (int e = blockIdx.x * CTA + threadIdx.x; e < n; e += LP_BLOCKS * CTA) {
        const size_t p = (size_t)(r0 + e / box) * W + (size_t)(c0 + e % box);
        const float xc = (float)((double)x[p] - mx), yc = (float)((double)y[p] - my);
        if (dy) atomicAdd(dy + p, -(k1 * xc - ky * yc));
        if (dx) atomicAdd(dx + p, -(k1 * yc - kx * xc));
    }
}

}  // namespace fsgs
