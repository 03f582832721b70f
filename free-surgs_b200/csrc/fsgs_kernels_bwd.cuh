// fsgs_kernels_bwd.cuh -- per-Gaussian backward (K8 + K9): conic/mean2D/colour/depth gradients
// from the compositor's accumulator -> gradients of the rasteriser inputs; the fused flavour
// continues through the activations, the SH evaluation and transform_to_frame and reduces the
// pose gradient dL/dRt = sum_i g_i [p_i;1]^T on the fly (warp reduce-scatter -> shared memory ->
// 12 atomics per CTA).
#pragma once

#include "fsgs_device.cuh"
#include "fsgs_kernels_pre.cuh"

namespace fsgs {

__device__ __forceinline__ void load_acc(const float *__restrict__ grad_acc, int i, float *a) {
    const float4 *r = reinterpret_cast<const float4 *>(grad_acc + (size_t)i * ACC_F);
    const float4 x = r[0], y = r[1], z = r[2];
    a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
    a[8] = z.x; a[9] = z.y; a[10] = z.z; a[11] = z.w;
}

__global__ void __launch_bounds__(CTA)
k_preprocess_api_bwd(CamConst cc, int P, const float *__restrict__ means3D, const float *__restrict__ shs,
                     const float *__restrict__ scales, const float *__restrict__ rotations,
                     const float *__restrict__ cov3D_precomp, const float *__restrict__ viewmatrix,
                     const float *__restrict__ projmatrix, const float *__restrict__ campos,
                     const float4 *__restrict__ records, const uint8_t *__restrict__ clamped,
                     const float *__restrict__ grad_acc, float *__restrict__ dL_dmeans2D,
                     float *__restrict__ dL_dcolors, float *__restrict__ dL_dopacity, float *__restrict__ dL_dmeans3D,
                     float *__restrict__ dL_dcov3D, float *__restrict__ dL_dsh, float *__restrict__ dL_dscales,
                     float *__restrict__ dL_drot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const size_t n = (size_t)i;
    const int radius = __float_as_int(records[n * 3 + 2].z);
    float a[ACC_F];
    float dmean[3] = {0.f, 0.f, 0.f}, dc6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < ACC_F; ++k) a[k] = 0.f;
    const int nsh = cc.n_coeffs * 3;
    float *dsh_i = (shs && dL_dsh) ? dL_dsh + n * nsh : nullptr;
    if (radius > 0) {
        load_acc(grad_acc, i, a);
        float V[16], PM[16], cp[3] = {0.f, 0.f, 0.f};
        load16(viewmatrix, V);
        load16(projmatrix, PM);
        if (shs) { cp[0] = __ldg(campos); cp[1] = __ldg(campos + 1); cp[2] = __ldg(campos + 2); }
        const float mean[3] = {means3D[3 * n], means3D[3 * n + 1], means3D[3 * n + 2]};
        api_backward_one(cc, V, PM, cp, mean, shs ? shs + n * nsh : nullptr, scales ? scales + 3 * n : nullptr,
                         rotations ? rotations + 4 * n : nullptr, cov3D_precomp ? cov3D_precomp + 6 * n : nullptr,
                         clamped[i], a, dmean, dc6, dsh_i, ds, dq);
    } else if (dsh_i) {
        for (int k = 0; k < nsh; ++k) dsh_i[k] = 0.f;
    }
    if (dL_dmeans2D) { dL_dmeans2D[3 * n] = a[0]; dL_dmeans2D[3 * n + 1] = a[1]; dL_dmeans2D[3 * n + 2] = 0.f; }
    if (dL_dcolors) { dL_dcolors[3 * n] = a[6]; dL_dcolors[3 * n + 1] = a[7]; dL_dcolors[3 * n + 2] = a[8]; }
    if (dL_dopacity) dL_dopacity[i] = a[5];
    if (dL_dmeans3D) { dL_dmeans3D[3 * n] = dmean[0]; dL_dmeans3D[3 * n + 1] = dmean[1]; dL_dmeans3D[3 * n + 2] = dmean[2]; }
    if (dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dL_dcov3D[6 * n + k] = dc6[k];
    }
    if (dL_dscales) { dL_dscales[3 * n] = ds[0]; dL_dscales[3 * n + 1] = ds[1]; dL_dscales[3 * n + 2] = ds[2]; }
    if (dL_drot) *reinterpret_cast<float4 *>(dL_drot + 4 * n) = make_float4(dq[0], dq[1], dq[2], dq[3]);
}

#ifndef FSGS_PREBWD_MINB
#define FSGS_PREBWD_MINB 1
#endif
#ifndef FSGS_PREBWD_CTA
#define FSGS_PREBWD_CTA 64
#endif
// threads per CTA of the fused per-Gaussian backward.  A/B on B200: 256 -> 0.074 / 0.273 ms, 128 -> 0.072 / 0.268,
// 64 -> 0.070 / 0.256 (P = 500k / 2M): at 128 registers only 512 threads fit an SM, and smaller CTAs interleave
// their load -> compute -> store phases better.  (A persistent variant with a two-stage bulk-TMA ring per CTA --
// next chunk's coefficients in flight while the current one is computed -- was tried and was slower: 0.072 / 0.275 ms;
// 140 registers, 7 CTAs/SM and a store-read wait per chunk.)
constexpr int PREBWD_CTA = FSGS_PREBWD_CTA;
__global__ void __launch_bounds__(PREBWD_CTA, FSGS_PREBWD_MINB)
k_preprocess_fused_bwd(CamConst cc, int P, const float *__restrict__ xyz, const float *__restrict__ f_dc,
                       const float *__restrict__ f_rest, const float *__restrict__ opacity_raw,
                       const float *__restrict__ scaling_raw, const float *__restrict__ rotation_raw,
                       const float *__restrict__ pose, const float *__restrict__ cam_center,
                       const float *__restrict__ viewmatrix, const float *__restrict__ projmatrix,
                       const float4 *__restrict__ records, const uint8_t *__restrict__ clamped,
                       const float *__restrict__ grad_acc, int gs_grad, int cam_grad, float *__restrict__ dL_dxyz,
                       float *__restrict__ dL_dfdc, float *__restrict__ dL_dfrest, float *__restrict__ dL_dopacity_raw,
                       float *__restrict__ dL_dscaling_raw, float *__restrict__ dL_drotation_raw,
                       float *__restrict__ dL_dpose, float *__restrict__ dL_dmeans2D, int use_tma,
                       unsigned long long *__restrict__ err, float *__restrict__ dL_dsh_rgb, int first, int end,
                       float *__restrict__ stat_accum, float *__restrict__ stat_denom, float *__restrict__ compact) {
    // [first, end): the Gaussians this launch covers (the frame-parallel exchange launches the kernel range by range
    //   and reduces range k over the ranks while range k + 1 is computed); first % 4 == 0 keeps the bulk copies aligned.
    // stat_accum / stat_denom [P,1] (optional): GaussianModel.add_densification_stats (scene/gaussian_model.py:678-681)
    //   folded in -- accum[i] += ||dL/dmeans2D_i||, denom[i] += 1 for every visible Gaussian.
    // compact [P,14] (optional): rotation | xyz | scaling | opacity | masked colour gradient as ONE 56-byte row per
    //   Gaussian (the exchange unit of the frame-parallel mode: a Gaussian range is then one contiguous buffer);
    //   replaces the separate xyz / opacity / scaling / rotation / dL_dsh_rgb outputs.
    // SH coefficients in, SH gradients out through ONE shared-memory buffer: bulk TMA load of the CTA's
    // 256 x 180 B slice, each thread turns its 45 coefficients into their gradients in place, bulk TMA
    // store to dL/dfeatures_rest (the plain path does 45 scalar loads + 45 scalar stores at a 180 B stride).
    __shared__ __align__(128) float s_rest[PREBWD_CTA * 45];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ float s_pose[16];
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int base = first + blockIdx.x * blockDim.x, count = min((int)blockDim.x, end - base);
    P = end;
    // (without dL_dfrest -- the frame-parallel compact mode -- the coefficients are still staged: the SH
    // view-direction gradient needs them; nothing is stored back)
    const bool staged = (reinterpret_cast<uintptr_t>(f_rest) & 15u) == 0 &&
                        (reinterpret_cast<uintptr_t>(dL_dfrest) & 15u) == 0 && use_tma;
    if (threadIdx.x < 16) s_pose[threadIdx.x] = 0.f;
    if (staged) stage_rows_issue<45>(s_rest, f_rest, base, count, &s_bar);
    float pg[16];   // pose-gradient contributions of this Gaussian: g_r * [x y z 1]_c at [4r + c]
#pragma unroll
    for (int k = 0; k < 16; ++k) pg[k] = 0.f;
    // this thread's own inputs (radius word, then the 48 B accumulator row + 56 B of parameters) go in flight while
    // the bulk copy of the SH rows is under way: the wait for the copy, the radius load and the loads it gates were
    // three serialised DRAM latencies (ncu: 37 % of the kernel's stall samples on those three instructions)
    int radius = 0;
    float a[ACC_F], w[3] = {0.f, 0.f, 0.f}, sc[3] = {0.f, 0.f, 0.f}, q[4] = {1.f, 0.f, 0.f, 0.f}, dcv[3] = {0.f, 0.f, 0.f};
    float op_raw = 0.f;
    uint8_t cl = 0;
#pragma unroll
    for (int k = 0; k < ACC_F; ++k) a[k] = 0.f;
    if (i < P) {
        const size_t n = (size_t)i;
        radius = __float_as_int(records[n * 3 + 2].z);
        if (radius > 0) {
            load_acc(grad_acc, i, a);
            w[0] = xyz[3 * n]; w[1] = xyz[3 * n + 1]; w[2] = xyz[3 * n + 2];
            sc[0] = scaling_raw[3 * n]; sc[1] = scaling_raw[3 * n + 1]; sc[2] = scaling_raw[3 * n + 2];
            const float4 q4 = *reinterpret_cast<const float4 *>(rotation_raw + 4 * n);
            q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
            dcv[0] = f_dc[3 * n]; dcv[1] = f_dc[3 * n + 1]; dcv[2] = f_dc[3 * n + 2];
            op_raw = opacity_raw[i];
            cl = clamped[i];
        }
    }
    if (staged) stage_rows_wait(count, &s_bar, err);
    else __syncthreads();
    if (i < P) {
        const size_t n = (size_t)i;
        float dxyz[3] = {0.f, 0.f, 0.f}, ds_raw[3] = {0.f, 0.f, 0.f}, dq_raw[4] = {0.f, 0.f, 0.f, 0.f};
        float dop_raw = 0.f, dfdc[3] = {0.f, 0.f, 0.f}, m2d[2] = {0.f, 0.f}, gc[3] = {0.f, 0.f, 0.f};
        float *drest = !dL_dfrest ? nullptr : (staged ? s_rest + 45 * threadIdx.x : dL_dfrest + 45 * n);
        const float *rest = staged ? s_rest + 45 * threadIdx.x : f_rest + 45 * n;
        if (radius > 0) {
            float V[16], PM[16], Rt[12], cp[3];
            load16(viewmatrix, V);
            load16(projmatrix, PM);
#pragma unroll
            for (int k = 0; k < 12; ++k) Rt[k] = __ldg(pose + k);
            cp[0] = __ldg(cam_center); cp[1] = __ldg(cam_center + 1); cp[2] = __ldg(cam_center + 2);
            fused_backward_one(cc, V, PM, Rt, cp, w, dcv, rest, op_raw, sc, q, cl, a,
                               gs_grad, cam_grad, dxyz, dfdc, drest, dop_raw, ds_raw, dq_raw, pg, m2d, gc);
        } else if (drest) {
            for (int k = 0; k < 45; ++k) drest[k] = 0.f;
        }
        if (dL_dxyz) { dL_dxyz[3 * n] = dxyz[0]; dL_dxyz[3 * n + 1] = dxyz[1]; dL_dxyz[3 * n + 2] = dxyz[2]; }
        if (dL_dfdc) { dL_dfdc[3 * n] = dfdc[0]; dL_dfdc[3 * n + 1] = dfdc[1]; dL_dfdc[3 * n + 2] = dfdc[2]; }
        if (dL_dsh_rgb) { dL_dsh_rgb[3 * n] = gc[0]; dL_dsh_rgb[3 * n + 1] = gc[1]; dL_dsh_rgb[3 * n + 2] = gc[2]; }
        if (dL_dopacity_raw) dL_dopacity_raw[i] = dop_raw;
        if (dL_dscaling_raw) { dL_dscaling_raw[3 * n] = ds_raw[0]; dL_dscaling_raw[3 * n + 1] = ds_raw[1]; dL_dscaling_raw[3 * n + 2] = ds_raw[2]; }
        if (dL_drotation_raw) *reinterpret_cast<float4 *>(dL_drotation_raw + 4 * n) = make_float4(dq_raw[0], dq_raw[1], dq_raw[2], dq_raw[3]);
        if (dL_dmeans2D) { dL_dmeans2D[3 * n] = m2d[0]; dL_dmeans2D[3 * n + 1] = m2d[1]; dL_dmeans2D[3 * n + 2] = 0.f; }
        if (stat_accum && radius > 0) {
            stat_accum[i] += sqrtf(m2d[0] * m2d[0] + m2d[1] * m2d[1]);
            stat_denom[i] += 1.0f;
        }
        if (compact) {
            float2 *row = reinterpret_cast<float2 *>(compact + 14 * n);          // 56-byte rows: 8-byte aligned
            row[0] = make_float2(dq_raw[0], dq_raw[1]); row[1] = make_float2(dq_raw[2], dq_raw[3]);
            row[2] = make_float2(dxyz[0], dxyz[1]);     row[3] = make_float2(dxyz[2], ds_raw[0]);
            row[4] = make_float2(ds_raw[1], ds_raw[2]); row[5] = make_float2(dop_raw, gc[0]);
            row[6] = make_float2(gc[1], gc[2]);
        }
    }
    if (staged && dL_dfrest) {
        // gradients -> global: bulk store for the 16-byte-multiple prefix, plain stores for <= 3 rows
        fence_proxy_async_smem();
        __syncthreads();
        const int rows_tma = count & ~3;
        float *gdst = dL_dfrest + (size_t)base * 45;
        if (threadIdx.x == 0 && rows_tma > 0) tma_store_1d(gdst, s_rest, (uint32_t)rows_tma * 180u);
        for (int q = rows_tma * 45 + threadIdx.x; q < count * 45; q += blockDim.x) gdst[q] = s_rest[q];
        if (threadIdx.x == 0 && rows_tma > 0) tma_store_commit_and_wait();   // smem must outlive the read
    }
    if (cam_grad && dL_dpose) {
        warp_reduce_scatter16(pg, lane);
        const int idx = lane >> 1;
        if ((lane & 1) == 0 && idx < 12 && pg[0] != 0.f) atomicAdd(&s_pose[idx], pg[0]);
        __syncthreads();
        if (threadIdx.x < 12 && s_pose[threadIdx.x] != 0.f) atomicAdd(&dL_dpose[threadIdx.x], s_pose[threadIdx.x]);
    }
}

// Pose-only flavour of the fused per-Gaussian backward (tracking with a frozen Gaussian model: the caller
// wants dL/dpose and nothing else).  Reads 48 B accumulator row + 12 B xyz + 28 B scale / rotation + the radius
// word per Gaussian, writes nothing per Gaussian: no SH staging, no 236 B/Gaussian of parameter gradients.
// The block reduction of dL/dRt is the one of k_preprocess_fused_bwd.
__global__ void __launch_bounds__(CTA)
k_preprocess_pose_bwd(CamConst cc, int P, const float *__restrict__ xyz, const float *__restrict__ scaling_raw,
                      const float *__restrict__ rotation_raw, const float *__restrict__ pose,
                      const float *__restrict__ viewmatrix, const float *__restrict__ projmatrix,
                      const float4 *__restrict__ records, const float *__restrict__ grad_acc,
                      float *__restrict__ dL_dpose) {
    __shared__ float s_pose[16];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 16) s_pose[threadIdx.x] = 0.f;
    __syncthreads();
    float pg[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) pg[k] = 0.f;
    if (i < P) {
        const size_t n = (size_t)i;
        const int radius = __float_as_int(records[n * 3 + 2].z);
        if (radius > 0) {
            float a[ACC_F];
            load_acc(grad_acc, i, a);
            float V[16], PM[16], Rt[12];
            load16(viewmatrix, V);
            load16(projmatrix, PM);
#pragma unroll
            for (int k = 0; k < 12; ++k) Rt[k] = __ldg(pose + k);
            const float w[3] = {xyz[3 * n], xyz[3 * n + 1], xyz[3 * n + 2]};
            const float sc[3] = {scaling_raw[3 * n], scaling_raw[3 * n + 1], scaling_raw[3 * n + 2]};
            const float4 q4 = *reinterpret_cast<const float4 *>(rotation_raw + 4 * n);
            const float q[4] = {q4.x, q4.y, q4.z, q4.w};
            fused_backward_pose_one(cc, V, PM, Rt, w, sc, q, a, pg);
        }
    }
    warp_reduce_scatter16(pg, lane);
    const int idx = lane >> 1;
    if ((lane & 1) == 0 && idx < 12 && pg[0] != 0.f) atomicAdd(&s_pose[idx], pg[0]);
    __syncthreads();
    if (threadIdx.x < 12 && s_pose[threadIdx.x] != 0.f) atomicAdd(&dL_dpose[threadIdx.x], s_pose[threadIdx.x]);
}

// Frame-parallel exchange, second half: after the masked colour gradients gc[P,3] have been summed over the
// ranks, every SH-coefficient gradient is basis_k(dir) * gc (zero beyond the active degree), with
// dir = normalize(xyz - cam_center) identical on all ranks.  Pure write kernel: 192 B/Gaussian, staged in
// shared memory and written with one bulk TMA store per CTA like the backward above.
// One-shot variant of the exchange for SMALL rank counts (fsgs_compact_grad_expand_peers): the sum over the ranks is
// folded into this kernel -- collective + the compute that consumes it in ONE kernel over peer memory.  The CTA pulls
// its 256 rows (14 336 contiguous bytes) from every rank's buffer with coalesced 128-bit loads (the peers' through
// NVLink peer pointers), adds them in rank order (every rank adds in the same order: bit-identical sums), parks them in
// shared memory and continues as below.  Per rank (N - 1) x the payload crosses the links, against 2 (N - 1) / N x for
// the two-shot kernel -- equal at N = 2, where it saves the separate exchange kernel and its launch.
// Pull-gather variant (owner_slices, 4+ ranks): the exchange kernel stopped after the reduce-scatter -- rank g holds the
// sums of the g-th 1/N slice, in a second region of its symmetric buffer -- and this kernel fetches every 16-byte word
// from its owner while it writes the SH gradients: the all-gather half of the two-shot exchange rides on the expansion.
struct PeerRows {
    const float4 *p[8];    // row (one-shot) / sum (pull-gather) buffers of ranks 0 .. world-1, 16-byte aligned
    int world;
    int owner_slices;      // 0: add the rows of all ranks;  1: take each word from the rank that owns its slice
    int slice_first, slice_per;   // slice geometry of the reduce-scatter, in float4 (fsgs_exchange_rows)
};

__global__ void __launch_bounds__(CTA)
k_sh_grad_expand(int P, int sh_deg, const float *__restrict__ xyz, const float *__restrict__ cam_center,
                 const float *__restrict__ gc, float *__restrict__ dL_dfdc, float *__restrict__ dL_dfrest, int use_tma,
                 int first, const float *__restrict__ compact, float *__restrict__ dL_dxyz,
                 float *__restrict__ dL_dopacity_raw, float *__restrict__ dL_dscaling_raw,
                 float *__restrict__ dL_drotation_raw, PeerRows peers) {
    // [first, P): the Gaussian range of this launch (first % 4 == 0).  With `compact` [P,14] (the rank-summed rows of
    // k_preprocess_fused_bwd) the colour gradient is taken from the row and the row's other 11 floats are unpacked
    // into the per-parameter gradient tensors on the way.
    __shared__ __align__(128) float s_rest[CTA * 45];
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    const int base = first + blockIdx.x * blockDim.x, count = min((int)blockDim.x, P - base);
    const bool staged = (reinterpret_cast<uintptr_t>(dL_dfrest) & 15u) == 0 && use_tma;
    if (peers.world > 1) {
        // rows [base, base + count) of every rank, summed in rank order into the (not yet used) staging buffer
        const int n4 = (count * 14 + 3) / 4;                       // base % 4 == 0: the slice starts on a float4
        float4 *s_rows = reinterpret_cast<float4 *>(s_rest);
        // 256 rows = 896 float4: up to four per thread, all of one rank's loads in flight before they are added (the
        // peers' cross NVLink: microseconds of latency each)
        constexpr int U = (CTA * 14 / 4 + CTA - 1) / CTA;
        const size_t at = (size_t)base * 14 / 4;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 acc[U];
        if (peers.owner_slices) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = u * CTA + (int)threadIdx.x;
                const int w = (int)at + q;
                const int owner = min(peers.world - 1, (w - peers.slice_first) / peers.slice_per);
                acc[u] = q < n4 ? peers.p[owner][w] : zero4;
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = u * CTA + (int)threadIdx.x;
                acc[u] = q < n4 ? peers.p[0][at + q] : zero4;
            }
        }
        for (int r = 1; r < (peers.owner_slices ? 1 : peers.world); ++r) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = u * CTA + (int)threadIdx.x;
                v[u] = q < n4 ? peers.p[r][at + q] : zero4;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = u * CTA + (int)threadIdx.x;
            if (q < n4) s_rows[q] = acc[u];
        }
        __syncthreads();
    }
    float2 mine[7];
    if (peers.world > 1 && i < P) {
        const float2 *row = reinterpret_cast<const float2 *>(s_rest) + 7 * threadIdx.x;
#pragma unroll
        for (int k = 0; k < 7; ++k) mine[k] = row[k];
    }
    if (peers.world > 1) __syncthreads();                          // the staging buffer is reused for the SH rows below
    if (i < P) {
        const size_t n = (size_t)i;
        float g0, g1, g2;
        if (compact || peers.world > 1) {
            const float2 *row = peers.world > 1 ? mine : reinterpret_cast<const float2 *>(compact + 14 * n);
            const float2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4], r5 = row[5], r6 = row[6];
            *reinterpret_cast<float4 *>(dL_drotation_raw + 4 * n) = make_float4(r0.x, r0.y, r1.x, r1.y);
            dL_dxyz[3 * n] = r2.x; dL_dxyz[3 * n + 1] = r2.y; dL_dxyz[3 * n + 2] = r3.x;
            dL_dscaling_raw[3 * n] = r3.y; dL_dscaling_raw[3 * n + 1] = r4.x; dL_dscaling_raw[3 * n + 2] = r4.y;
            dL_dopacity_raw[i] = r5.x;
            g0 = r5.y; g1 = r6.x; g2 = r6.y;
        } else {
            g0 = gc[3 * n]; g1 = gc[3 * n + 1]; g2 = gc[3 * n + 2];
        }
        float d[3] = {xyz[3 * n] - __ldg(cam_center), xyz[3 * n + 1] - __ldg(cam_center + 1),
                      xyz[3 * n + 2] - __ldg(cam_center + 2)};
        const float inv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        d[0] *= inv; d[1] *= inv; d[2] *= inv;
        float B[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) B[k] = 0.f;
        sh_basis(sh_deg, d[0], d[1], d[2], B);
        const bool any = (g0 != 0.f) | (g1 != 0.f) | (g2 != 0.f);     // untouched Gaussians: exact zeros, also for NaN-free dir
        dL_dfdc[3 * n] = any ? B[0] * g0 : 0.f; dL_dfdc[3 * n + 1] = any ? B[0] * g1 : 0.f; dL_dfdc[3 * n + 2] = any ? B[0] * g2 : 0.f;
        float *o = staged ? s_rest + 45 * threadIdx.x : dL_dfrest + 45 * n;
        const int nb = (sh_deg + 1) * (sh_deg + 1);
#pragma unroll
        for (int k = 1; k < 16; ++k) {
            const float b = (any && k < nb) ? B[k] : 0.f;
            o[3 * (k - 1)] = b * g0; o[3 * (k - 1) + 1] = b * g1; o[3 * (k - 1) + 2] = b * g2;
        }
    }
    if (staged) {
        fence_proxy_async_smem();
        __syncthreads();
        const int rows_tma = count & ~3;
        float *gdst = dL_dfrest + (size_t)base * 45;
        if (threadIdx.x == 0 && rows_tma > 0) tma_store_1d(gdst, s_rest, (uint32_t)rows_tma * 180u);
        for (int q = rows_tma * 45 + threadIdx.x; q < count * 45; q += blockDim.x) gdst[q] = s_rest[q];
        if (threadIdx.x == 0 && rows_tma > 0) tma_store_commit_and_wait();
    }
}

// ---- frame-parallel exchange over NVLink / NVSwitch ------------------------------------------
// Two-shot all-reduce of the 56-byte gradient rows, hand-written instead of ncclAllReduce.  Every rank holds the
// rows of ITS frame in a symmetric buffer (same offset on every GPU, mapped into every peer's address space).
// After a cross-GPU barrier (all rows written), rank g owns the g-th 1/N slice of the buffer:
//   multicast path (NVSwitch, NVLS): ONE multimem.ld_reduce per 16 bytes makes the switch fetch the N copies and
//     add them on the way, ONE multimem.st broadcasts the sum back into all N buffers -- per GPU 1/N of the
//     payload crosses its links in each direction, whatever N is;
//   peer path (no multicast object): the owner loads its slice from the N buffers through peer pointers, adds,
//     and stores the sum to each of them.
// After a second barrier every buffer holds the sum of all frames' rows, bit-identical on every rank (each 16-byte
// word was summed exactly once, by its owner).  The barriers are the caller's (stream-ordered signal kernels).
struct PeerPtrs {
    float4 *p[8];
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4 *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4 *mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(CTA)
k_exchange_rows(float4 *__restrict__ mc, PeerPtrs peers, int world, int rank, long long begin, long long end,
                long long scatter_off) {
    // scatter_off == 0: all-reduce (the sums go back into every rank's rows).  scatter_off > 0: reduce-scatter only --
    // the sums of this rank's slice go to ITS OWN buffer, scatter_off float4 behind the rows (a separate region: peers
    // may still be pulling the previous step's sums while the next step's rows are being written).
    // [begin, end): this rank's slice, in float4 units from the start of the symmetric buffer.  Four independent
    // 16-byte words per thread and iteration: the round trip through the switch is microseconds long, the links
    // only fill up with a few MB in flight.
    constexpr int U = 4;
    const long long stride = (long long)gridDim.x * CTA * U;
    const long long t0 = begin + (long long)blockIdx.x * CTA * U + threadIdx.x;
    if (mc) {
        for (long long i = t0; i < end; i += stride) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * CTA < end) v[u] = multimem_ld_reduce_add(mc + i + u * CTA);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * CTA < end) {
                    if (scatter_off) peers.p[rank][scatter_off + i + u * CTA] = v[u];
                    else multimem_st(mc + i + u * CTA, v[u]);
                }
        }
    } else {
        for (long long i = t0; i < end; i += stride) {
            float4 s[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * CTA < end) s[u] = peers.p[rank][i + u * CTA];
            for (int d = 1; d < world; ++d) {               // peers in ring order from this rank: spreads the links
                const float4 *src = peers.p[(rank + d) % world];
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (i + u * CTA < end) v[u] = src[i + u * CTA];
#pragma unroll
                for (int u = 0; u < U; ++u) { s[u].x += v[u].x; s[u].y += v[u].y; s[u].z += v[u].z; s[u].w += v[u].w; }
            }
            for (int d = 0; d < (scatter_off ? 1 : world); ++d) {
                float4 *dst = peers.p[(rank + d) % world] + scatter_off;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (i + u * CTA < end) dst[i + u * CTA] = s[u];
            }
        }
    }
}

}  // namespace fsgs
