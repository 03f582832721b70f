// fsgs_kernels_refstyle.cuh -- BASELINE, not the product: the two compositors in the kernel STRUCTURE of the published
// 3D-Gaussian-splatting rasteriser that Free-SurGS installs (requirements.txt:26; SURVEY.md Appendix A, kernels K6 / K7),
// restated from its description -- the package's source is not on disk:
//   * one thread per pixel, the whole tile list walked by every pixel (no block masks, no per-warp compaction);
//   * records fetched cooperatively, 256 per round, with plain loads and two block barriers per round;
//   * backward: every contributing (pixel, Gaussian) pair adds its ten gradient terms to global memory with ten
//     scalar float atomics; IEEE division for T / (1 - alpha).
// Selected with FSGS_FLAG_UPSTREAM_STYLE on the API flavour (one GaussianRasterizer pass), together with
// FSGS_FLAG_NO_TILE_CULL (the reference's full 3-sigma rectangles).  It exists so that bench.py can time "the generic
// kernel design compiled for sm_100a"  beside the library's own compositors on the same lists; binning and
// the per-Gaussian kernels stay the library's, which can only flatter this baseline.  The arithmetic per pair is the
// library's (same records, same pinned alpha expression), so the forward is bit-identical and the tests can hold the
// baseline to the same oracle.
#pragma once

#include "fsgs_device.cuh"

namespace fsgs {

__global__ void __launch_bounds__(CTA)
k_composite_fwd_ref(CamConst cc, const unsigned int *__restrict__ tile_offset, const float4 *__restrict__ sorted_rec,
                    const float *__restrict__ bg, float *__restrict__ out_planes, float *__restrict__ out_depth,
                    float *__restrict__ final_T, unsigned int *__restrict__ n_contrib,
                    const unsigned long long *__restrict__ counters, unsigned long long capacity, unsigned int bin_cap) {
    __shared__ float4 s_rec[BATCH * REC_F4];
    if (counters[CNT_R] > capacity || (bin_cap && counters[CNT_MAXLIST] > bin_cap)) return;
    const int tile = blockIdx.x, tid = threadIdx.x;
    const unsigned int start = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - start);
    const int px = (tile % cc.gx) * TILE + (tid & 15), py = (tile / cc.gx) * TILE + (tid >> 4);
    const bool inside = px < cc.W && py < cc.H;
    const float pxf = (float)px, pyf = (float)py;
    bool done = !inside;
    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    unsigned int contributor = 0, last = 0;
    const int rounds = (n + BATCH - 1) / BATCH;
    for (int r = 0; r < rounds; ++r) {
        if (__syncthreads_count(done ? 1 : 0) == CTA) break;
        const int idx = r * BATCH + tid;
        if (idx < n) {
            const float4 *src = sorted_rec + ((size_t)start + idx) * REC_F4;
            s_rec[tid * 3] = ldg4(src); s_rec[tid * 3 + 1] = ldg4(src + 1); s_rec[tid * 3 + 2] = ldg4(src + 2);
        }
        __syncthreads();
        const int cnt = min(BATCH, n - r * BATCH);
        for (int j = 0; !done && j < cnt; ++j) {
            ++contributor;
            const float4 q0 = s_rec[j * 3], q1 = s_rec[j * 3 + 1];
            const float2 q2 = *reinterpret_cast<const float2 *>(&s_rec[j * 3 + 2]);
            const float dx = q0.x - pxf, dy = q0.y - pyf;
            const float p2 = gauss_power2(q0.z, q0.w, q1.x, dx, dy);
            if (p2 > 0.f) continue;
            const float alpha = fminf(ALPHA_MAX, q1.y * fast_exp2(p2));
            if (alpha < ALPHA_MIN) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < T_MIN) { done = true; continue; }
            const float w = alpha * T;
            C0 = fmaf(q1.z, w, C0); C1 = fmaf(q1.w, w, C1); C2 = fmaf(q2.x, w, C2); D = fmaf(q2.y, w, D);
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t HW = (size_t)cc.W * cc.H, p = (size_t)py * cc.W + px;
        final_T[p] = T;
        n_contrib[p] = last;
        out_planes[p] = C0 + T * __ldg(bg);
        out_planes[HW + p] = C1 + T * __ldg(bg + 1);
        out_planes[2 * HW + p] = C2 + T * __ldg(bg + 2);
        out_depth[p] = D;
    }
}

__global__ void __launch_bounds__(CTA)
k_composite_bwd_ref(CamConst cc, const unsigned int *__restrict__ tile_offset, const float4 *__restrict__ sorted_rec,
                    const float *__restrict__ bg, const float *__restrict__ final_T,
                    const unsigned int *__restrict__ n_contrib, const float *__restrict__ g_rgb,
                    const float *__restrict__ g_depth, float *__restrict__ grad_acc,
                    const unsigned long long *__restrict__ counters, unsigned long long capacity, unsigned int bin_cap) {
    __shared__ float4 s_rec[BATCH * REC_F4];
    if (counters[CNT_R] > capacity || (bin_cap && counters[CNT_MAXLIST] > bin_cap)) return;
    const int tile = blockIdx.x, tid = threadIdx.x;
    const unsigned int start = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - start);
    const int px = (tile % cc.gx) * TILE + (tid & 15), py = (tile / cc.gx) * TILE + (tid >> 4);
    const bool inside = px < cc.W && py < cc.H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)cc.W * cc.H, p = (size_t)py * cc.W + px;
    const float T_final = inside ? final_T[p] : 0.f;
    const int last = inside ? (int)n_contrib[p] : 0;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (inside) {
        if (g_rgb) { g[0] = g_rgb[p]; g[1] = g_rgb[HW + p]; g[2] = g_rgb[2 * HW + p]; }
        if (g_depth) g[3] = g_depth[p];
    }
    const float bgdot = __ldg(bg) * g[0] + __ldg(bg + 1) * g[1] + __ldg(bg + 2) * g[2];
    const float kx = 0.5f * cc.W, ky = 0.5f * cc.H;
    float T = T_final, last_alpha = 0.f;
    float accum[4] = {0.f, 0.f, 0.f, 0.f}, last_c[4] = {0.f, 0.f, 0.f, 0.f};
    const int rounds = (n + BATCH - 1) / BATCH;
    for (int r = 0; r < rounds; ++r) {
        __syncthreads();
        const int idx = n - 1 - (r * BATCH + tid);            // back to front
        if (idx >= 0) {
            const float4 *src = sorted_rec + ((size_t)start + idx) * REC_F4;
            s_rec[tid * 3] = ldg4(src); s_rec[tid * 3 + 1] = ldg4(src + 1); s_rec[tid * 3 + 2] = ldg4(src + 2);
        }
        __syncthreads();
        const int cnt = min(BATCH, n - r * BATCH);
        for (int j = 0; j < cnt; ++j) {
            const int contributor = n - (r * BATCH + j);        // 1-based position of this entry in the list
            if (contributor > last) continue;
            const float4 q0 = s_rec[j * 3], q1 = s_rec[j * 3 + 1], q2 = s_rec[j * 3 + 2];
            const float dx = q0.x - pxf, dy = q0.y - pyf;
            const float p2 = gauss_power2(q0.z, q0.w, q1.x, dx, dy);
            if (p2 > 0.f) continue;
            const float G = fast_exp2(p2);
            const float alpha = fminf(ALPHA_MAX, q1.y * G);
            if (alpha < ALPHA_MIN) continue;
            T = T / (1.f - alpha);
            const float w = alpha * T;
            const float c[4] = {q1.z, q1.w, q2.x, q2.y};
            float *row = grad_acc + (size_t)__float_as_uint(q2.w) * ACC_F;
            float dL_dalpha = 0.f;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                accum[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * accum[ch];
                last_c[ch] = c[ch];
                dL_dalpha += (c[ch] - accum[ch]) * g[ch];
                atomicAdd(row + 6 + ch, w * g[ch]);
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
            const float dL_dG = q1.y * dL_dalpha;
            float A, B, C;
            unscale_conic(q0.z, q0.w, q1.x, A, B, C);
            const float gdx = G * dx, gdy = G * dy;
            atomicAdd(row + 0, dL_dG * (-gdx * A - gdy * B) * kx);
            atomicAdd(row + 1, dL_dG * (-gdy * C - gdx * B) * ky);
            atomicAdd(row + 2, -0.5f * gdx * dx * dL_dG);
            atomicAdd(row + 3, -0.5f * gdx * dy * dL_dG);
            atomicAdd(row + 4, -0.5f * gdy * dy * dL_dG);
            atomicAdd(row + 5, G * dL_dalpha);
        }
    }
}

}  // namespace fsgs
