// fsgs_kernels_pre.cuh -- per-Gaussian kernels: projection (API + fused), per-tile counting,
// tile scan, instance scatter, per-tile sort.  sm_100a.
#pragma once

#include "fsgs_device.cuh"

namespace fsgs {

// One compiled copy of the tile test, so the counting pass (inside the projection kernels) and
// the scatter pass take bit-identical keep/drop decisions.
__device__ __noinline__ bool tile_hit_ni(float px, float py, float A, float B, float C, float tau, int tx, int ty) {
    CullEllipse e;
    e.px = px; e.py = py; e.A = A; e.B = B; e.C = C; e.tau = tau; e.hx = 0.f; e.hy = 0.f;
    e.nBC = -B * fast_rcp(C); e.nBA = -B * fast_rcp(A);     // two MUFU.RCP instead of four IEEE divisions
    return tile_hit(e, tx, ty);
}

// Tile range + padded threshold of one Gaussian; also a single compiled copy (it contains
// mul+add chains that the compiler could otherwise contract differently per call site).
struct TileRange {
    int x0, y0, x1, y1;
    float tau;
};
__device__ __noinline__ TileRange tile_range_ni(int gx, int gy, float px, float py, float A, float B, float C,
                                                float opacity, int radius, int no_cull) {
    const float rad = (float)radius;
    const int rminx = clampi((int)((px - rad) / TILE), 0, gx);
    const int rminy = clampi((int)((py - rad) / TILE), 0, gy);
    const int rmaxx = clampi((int)((px + rad + TILE - 1) / TILE), 0, gx);
    const int rmaxy = clampi((int)((py + rad + TILE - 1) / TILE), 0, gy);
    TileRange r;
    if (no_cull) {
        r.x0 = rminx; r.y0 = rminy; r.x1 = rmaxx; r.y1 = rmaxy; r.tau = 3.0e38f;
        return r;
    }
    const CullEllipse e = make_cull_ellipse(px, py, A, B, C, opacity);
    cull_rect(e, rminx, rminy, rmaxx, rmaxy, r.x0, r.y0, r.x1, r.y1);
    r.tau = e.tau;
    return r;
}

// Visit every kept tile of one Gaussian.  f(tile_index) is called for each; returns the count.
template <typename F>
__device__ __forceinline__ int for_each_tile(const CamConst &cc, float px, float py, float A, float B, float C,
                                             float opacity, int radius, bool no_cull, F &&f) {
    const TileRange tr = tile_range_ni(cc.gx, cc.gy, px, py, A, B, C, opacity, radius, no_cull ? 1 : 0);
    int n = 0;
    for (int ty = tr.y0; ty < tr.y1; ++ty)
        for (int tx = tr.x0; tx < tr.x1; ++tx)
            if (tile_hit_ni(px, py, A, B, C, tr.tau, tx, ty)) { f(ty * cc.gx + tx); ++n; }
    return n;
}

// Warp-cooperative version: the 32 lanes of a warp each own one Gaussian, but the candidate
// (Gaussian, tile) pairs of the whole warp are enumerated together, 32 per round, so a splat
// covering 40 tiles no longer leaves 31 lanes idle (ncu: 8-12 active lanes per instruction with the
// per-thread loops).  Every lane of the warp must call this (active = false when it has no splat).
// f(owner_lane, tile_index) is called, by an arbitrary lane, once for every kept tile.
template <typename F>
__device__ __forceinline__ void warp_for_each_tile(int gx, int gy, bool active, float px, float py, float A, float B,
                                                   float C, float opacity, int radius, bool no_cull, int lane, F &&f) {
    TileRange tr;
    tr.x0 = tr.y0 = tr.x1 = tr.y1 = 0; tr.tau = -1.f;
    if (active) tr = tile_range_ni(gx, gy, px, py, A, B, C, opacity, radius, no_cull ? 1 : 0);
    const int w = tr.x1 - tr.x0;
    const unsigned int c = (unsigned int)(w * (tr.y1 - tr.y0));
    unsigned int incl = c;                                   // inclusive prefix sum of candidate counts
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int nb = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += nb;
    }
    const unsigned int total = __shfl_sync(FULL, incl, 31);
    const unsigned int excl = incl - c;
    const float inv_w = w > 0 ? 1.0f / (float)w : 0.f;
    for (unsigned int base = 0; base < total; base += 32) {
        const unsigned int q = base + lane;
        // owner = first lane whose inclusive count exceeds q (5-step search over the warp's prefix sums)
        int owner = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const unsigned int v = __shfl_sync(FULL, incl, owner + step - 1);
            if (v <= q) owner += step;
        }
        owner = min(owner, 31);
        const float opx = __shfl_sync(FULL, px, owner), opy = __shfl_sync(FULL, py, owner);
        const float oA = __shfl_sync(FULL, A, owner), oB = __shfl_sync(FULL, B, owner), oC = __shfl_sync(FULL, C, owner);
        const float otau = __shfl_sync(FULL, tr.tau, owner), oinv = __shfl_sync(FULL, inv_w, owner);
        const int ox0 = __shfl_sync(FULL, tr.x0, owner), oy0 = __shfl_sync(FULL, tr.y0, owner);
        const int ow = __shfl_sync(FULL, w, owner);
        const unsigned int oexcl = __shfl_sync(FULL, excl, owner);
        if (q < total) {
            const int local = (int)(q - oexcl);
            int row = (int)(((float)local + 0.5f) * oinv);       // local / ow for the small integers involved
            row -= (row * ow > local) ? 1 : 0;
            row += ((row + 1) * ow <= local) ? 1 : 0;
            const int tx = ox0 + (local - row * ow), ty = oy0 + row;
            if (tile_hit_ni(opx, opy, oA, oB, oC, otau, tx, ty)) f(owner, ty * gx + tx);
        }
    }
}

__device__ __forceinline__ void load16(const float *__restrict__ p, float *o) {
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __ldg(p + i);
}

__device__ __forceinline__ void store_record(float4 *rec, int i, const Splat &s, float opacity, float r, float g,
                                             float b, int tiles) {
    rec[(size_t)i * 3 + 0] = make_float4(s.px, s.py, s.conx, s.cony);
    rec[(size_t)i * 3 + 1] = make_float4(s.conz, opacity, r, g);
    rec[(size_t)i * 3 + 2] = make_float4(b, s.depth, __int_as_float(s.radius), __int_as_float(tiles));
}
__device__ __forceinline__ void store_empty_record(float4 *rec, int i) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    rec[(size_t)i * 3 + 0] = z; rec[(size_t)i * 3 + 1] = z; rec[(size_t)i * 3 + 2] = z;
}

// block-wide sum of a small integer, one atomic per CTA
__device__ __forceinline__ void block_add_u64(unsigned long long *dst, unsigned int v) {
    __shared__ unsigned int s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const unsigned int w = __reduce_add_sync(FULL, v);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_sum, w);
    __syncthreads();
    if (threadIdx.x == 0 && s_sum) atomicAdd(dst, (unsigned long long)s_sum);
}

// The counting pass of the three projection kernels: one atomic per (Gaussian, tile) instance on the tile's counter.
// With bins (BinLayout) the value the atomic returns is the instance's slot in its tile's bin and its (depth, id) key
// is written there at once -- k_scatter (a second enumeration of the same tiles, a second atomic and an offset load
// per instance) is then not needed.  s_depth: the CTA's view depths (the callback runs on an arbitrary lane).
// (Deferring the key store by one enumeration round, so that the warp does not wait for the atomic's return, was
// measured: 0.0709 vs 0.0725 ms -- the cost is the 2.3 M scattered 8-byte stores themselves, as in k_scatter.)
struct BinSink {
    unsigned long long *bins;     // NULL = count only
    unsigned int cap;             // keys per tile
};
__device__ __forceinline__ void count_instance(unsigned int *tile_count, const BinSink &bs, const unsigned int *s_depth,
                                               unsigned int warp_first, int owner, int t) {
    if (bs.bins) {
        const unsigned int slot = atomicAdd(&tile_count[t], 1u);
        if (slot < bs.cap)
            bs.bins[(size_t)t * bs.cap + slot] =
                ((unsigned long long)s_depth[(threadIdx.x & ~31) + owner] << 32) | (warp_first + owner);
    } else {
        atomicAdd(&tile_count[t], 1u);
    }
}

// ---- K1, API flavour: one GaussianRasterizer call -----------------------------------------------
__global__ void __launch_bounds__(CTA)
k_preprocess_api(CamConst cc, int P, const float *__restrict__ means3D, const float *__restrict__ colors_precomp,
                 const float *__restrict__ shs, const float *__restrict__ opacities,
                 const float *__restrict__ scales, const float *__restrict__ rotations,
                 const float *__restrict__ cov3D_precomp, const float *__restrict__ viewmatrix,
                 const float *__restrict__ projmatrix, const float *__restrict__ campos, float4 *__restrict__ records,
                 uint8_t *__restrict__ clamped, int *__restrict__ radii, unsigned int *__restrict__ tile_count,
                 unsigned long long *__restrict__ counters, unsigned int flags, BinSink bs) {
    __shared__ unsigned int s_depth[CTA];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    unsigned int rect_tiles = 0;
    Splat sp;
    sp.px = sp.py = sp.conx = sp.cony = sp.conz = sp.depth = 0.f; sp.radius = 0;
    float opacity = 0.f;
    bool vis = false;
    if (i < P) {
        float V[16], PM[16], cp[3] = {0.f, 0.f, 0.f};
        load16(viewmatrix, V);
        load16(projmatrix, PM);
        if (shs) { cp[0] = __ldg(campos); cp[1] = __ldg(campos + 1); cp[2] = __ldg(campos + 2); }
        const size_t n = (size_t)i;
        const float mean[3] = {means3D[3 * n], means3D[3 * n + 1], means3D[3 * n + 2]};
        float rgb[3];
        uint8_t cl = 0;
        vis = api_forward_one(cc, V, PM, cp, mean, colors_precomp ? colors_precomp + 3 * n : nullptr,
                              shs ? shs + n * cc.n_coeffs * 3 : nullptr, scales ? scales + 3 * n : nullptr,
                              rotations ? rotations + 4 * n : nullptr, cov3D_precomp ? cov3D_precomp + 6 * n : nullptr, sp,
                              rgb, cl);
        if (vis) {
            opacity = opacities[i];
            rect_tiles = (unsigned int)((sp.rmaxx - sp.rminx) * (sp.rmaxy - sp.rminy));
            store_record(records, i, sp, opacity, rgb[0], rgb[1], rgb[2], 1);
        } else {
            store_empty_record(records, i);
        }
        clamped[i] = cl;
        radii[i] = vis ? sp.radius : 0;
    }
    s_depth[threadIdx.x] = __float_as_uint(sp.depth);
    __syncwarp();
    const unsigned int warp_first = (unsigned int)(i - lane);
    warp_for_each_tile(cc.gx, cc.gy, vis, sp.px, sp.py, sp.conx, sp.cony, sp.conz, opacity, sp.radius, (flags & 2u) != 0,
                       lane, [&](int owner, int t) { count_instance(tile_count, bs, s_depth, warp_first, owner, t); });
    block_add_u64(&counters[CNT_RECT], rect_tiles);
}

// ---- K1, fused flavour: gaussian_renderer.render's pre-processing folded in -----------------------
// pose is ROW-major [4,4] (LearnPose.forward output); features_dc [P,1,3], features_rest [P,15,3].
// (4 resident CTAs = 64 registers: with the 46 KB SH staging buffer that is also the shared-memory limit; an
// unconstrained build takes 74 registers, drops to 3 CTAs/SM and runs 10 % slower)
#ifndef FSGS_PRE_CTA
#define FSGS_PRE_CTA 128
#endif
// 1: all six SoA attribute arrays of the CTA's Gaussians through bulk-TMA staging (16-byte vectorised, no stride-3
// scalar loads); 0: only the SH rows (76 % of the bytes) are staged and every thread puts its own 56 B of loads in
// flight while that copy is under way.  A/B on B200 (round 2, tools/ab_variants.py, P = 500k / 2M):
//   staged-all, 128 threads 0.0605 / 0.1435 ms   staged-all, 64 threads 0.0584 / 0.1411 ms   SH rows only 0.0548 / 0.1366 ms
// -- the per-thread loads are already sector-efficient (a warp's stride-3 loads cover 384 contiguous bytes) and
// overlapping them with the bulk copy beats waiting for six copies behind one barrier.  Default 0.
#ifndef FSGS_PRE_STAGE_ALL
#define FSGS_PRE_STAGE_ALL 0
#endif
constexpr int PRE_CTA = FSGS_PRE_CTA;   // A/B on B200: 256 -> 0.063 / 0.157 ms, 128 -> 0.060 / 0.150, 64 -> 0.061 / 0.150 (500k / 2M)
__global__ void __launch_bounds__(PRE_CTA, 1024 / PRE_CTA)
k_preprocess_fused(CamConst cc, int P, const float *__restrict__ xyz, const float *__restrict__ f_dc,
                   const float *__restrict__ f_rest, const float *__restrict__ opacity_raw,
                   const float *__restrict__ scaling_raw, const float *__restrict__ rotation_raw,
                   const float *__restrict__ pose, const float *__restrict__ cam_center,
                   const float *__restrict__ viewmatrix, const float *__restrict__ projmatrix,
                   float4 *__restrict__ records, uint8_t *__restrict__ clamped, int *__restrict__ radii,
                   unsigned int *__restrict__ tile_count, unsigned long long *__restrict__ counters,
                   unsigned int flags, unsigned char *__restrict__ visibility, float *__restrict__ max_radii2D,
                   unsigned long long *__restrict__ err, BinSink bs) {
    __shared__ unsigned int s_depth[PRE_CTA];
    // The 180 B/Gaussian of higher-order SH coefficients (76 % of the input bytes) are contiguous per
    // CTA: one bulk TMA copy stages them; threads then read their own 45 floats at a conflict-free
    // stride.  FSGS_FLAG_NO_TMA (or a mis-aligned tensor) reads them straight from global memory.
    __shared__ __align__(128) float s_rest[PRE_CTA * 45];
    __shared__ __align__(16) float s_xyz[PRE_CTA * 3], s_sc[PRE_CTA * 3], s_dc[PRE_CTA * 3], s_rot[PRE_CTA * 4], s_op[PRE_CTA];
    __shared__ __align__(8) uint64_t s_bar;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int base = blockIdx.x * blockDim.x;
    // All six SoA attribute arrays of the CTA's Gaussians (236 B each) are contiguous slices: six bulk-TMA copies
    // behind one mbarrier bring them into shared memory (16-byte vectorised, fully coalesced, no per-thread
    // stride-3 scalar loads); threads then read their own rows at conflict-free strides (3, 4, 45 words).
    const bool aligned = ((reinterpret_cast<uintptr_t>(f_rest) | reinterpret_cast<uintptr_t>(xyz) |
                           reinterpret_cast<uintptr_t>(scaling_raw) | reinterpret_cast<uintptr_t>(f_dc) |
                           reinterpret_cast<uintptr_t>(rotation_raw) | reinterpret_cast<uintptr_t>(opacity_raw)) & 15u) == 0;
    const bool staged = (flags & 1u) == 0 && aligned;
    const int count = min((int)blockDim.x, P - base);
    constexpr bool ALL = FSGS_PRE_STAGE_ALL != 0;
    StageMulti sm{&s_bar, 0u};
    if (staged) {
        sm = stage_multi_begin(&s_bar);
        if (ALL) {
            stage_multi_add<3>(sm, s_xyz, xyz, base, count);
            stage_multi_add<3>(sm, s_sc, scaling_raw, base, count);
            stage_multi_add<4>(sm, s_rot, rotation_raw, base, count);
            stage_multi_add<3>(sm, s_dc, f_dc, base, count);
            stage_multi_add<1>(sm, s_op, opacity_raw, base, count);
        }
        stage_multi_add<45>(sm, s_rest, f_rest, base, count);
        if (ALL) stage_multi_wait(sm, err);
    }
    const int lane = threadIdx.x & 31;
    unsigned int rect_tiles = 0;
    Splat sp;
    sp.px = sp.py = sp.conx = sp.cony = sp.conz = sp.depth = 0.f; sp.radius = 0;
    float opacity = 0.f;
    bool vis = false;
    float w[3] = {0.f, 0.f, 0.f}, sc[3] = {0.f, 0.f, 0.f}, q[4] = {1.f, 0.f, 0.f, 0.f}, dcv[3] = {0.f, 0.f, 0.f};
    float op_raw = 0.f;
    if (i < P) {
        if (staged && ALL) {
            const int tl = threadIdx.x;
            w[0] = s_xyz[3 * tl]; w[1] = s_xyz[3 * tl + 1]; w[2] = s_xyz[3 * tl + 2];
            sc[0] = s_sc[3 * tl]; sc[1] = s_sc[3 * tl + 1]; sc[2] = s_sc[3 * tl + 2];
            const float4 q4 = *reinterpret_cast<const float4 *>(s_rot + 4 * tl);
            q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
            dcv[0] = s_dc[3 * tl]; dcv[1] = s_dc[3 * tl + 1]; dcv[2] = s_dc[3 * tl + 2];
            op_raw = s_op[tl];
        } else {
            const size_t n = (size_t)i;
            w[0] = xyz[3 * n]; w[1] = xyz[3 * n + 1]; w[2] = xyz[3 * n + 2];
            sc[0] = scaling_raw[3 * n]; sc[1] = scaling_raw[3 * n + 1]; sc[2] = scaling_raw[3 * n + 2];
            const float4 q4 = *reinterpret_cast<const float4 *>(rotation_raw + 4 * n);
            q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
            dcv[0] = f_dc[3 * n]; dcv[1] = f_dc[3 * n + 1]; dcv[2] = f_dc[3 * n + 2];
            op_raw = opacity_raw[i];
        }
    }
    if (staged && !ALL) stage_multi_wait(sm, err);      // own loads were put in flight while the SH rows were under way
    if (i < P) {
        float V[16], PM[16], Rt[12], cp[3];
        load16(viewmatrix, V);
        load16(projmatrix, PM);
#pragma unroll
        for (int k = 0; k < 12; ++k) Rt[k] = __ldg(pose + k);
        cp[0] = __ldg(cam_center); cp[1] = __ldg(cam_center + 1); cp[2] = __ldg(cam_center + 2);
        const size_t n = (size_t)i;
        float rgb[3];
        uint8_t cl = 0;
        const float *rest = staged ? s_rest + 45 * threadIdx.x : f_rest + 45 * n;
        vis = fused_forward_one(cc, V, PM, Rt, cp, w, dcv, rest, op_raw, sc, q, sp, opacity, rgb, cl);
        if (vis) {
            rect_tiles = (unsigned int)((sp.rmaxx - sp.rminx) * (sp.rmaxy - sp.rminy));
            store_record(records, i, sp, opacity, rgb[0], rgb[1], rgb[2], 1);
        } else {
            store_empty_record(records, i);
        }
        clamped[i] = cl;
        radii[i] = vis ? sp.radius : 0;
        // render()'s bookkeeping (gaussian_renderer/__init__.py:77-80,88): seen / visibility_filter and
        // max_radii2D[seen] = max(radius[seen], max_radii2D[seen])
        if (visibility) visibility[i] = vis ? 1 : 0;
        if (max_radii2D && vis) max_radii2D[i] = fmaxf(max_radii2D[i], (float)sp.radius);
    }
    s_depth[threadIdx.x] = __float_as_uint(sp.depth);
    __syncwarp();
    const unsigned int warp_first = (unsigned int)(i - lane);
    warp_for_each_tile(cc.gx, cc.gy, vis, sp.px, sp.py, sp.conx, sp.cony, sp.conz, opacity, sp.radius, (flags & 2u) != 0,
                       lane, [&](int owner, int t) { count_instance(tile_count, bs, s_depth, warp_first, owner, t); });
    block_add_u64(&counters[CNT_RECT], rect_tiles);
}

// ---- K1, frozen-model flavour -------------------------------------------------------------------------
// Free-SurGS' tracking loop renders the SAME Gaussian model from 50 successive pose estimates per frame
// (train.py:154-210), and with the reference's quirks everything about a Gaussian except its
// camera-frame mean is independent of the pose: the quaternions are not rotated into the camera frame, so
// Sigma_3D = R S^2 R^T is a model constant, and the SH view direction is WORLD position minus a frozen camera
// centre, so the colour and its clamp mask are too.  k_freeze_model evaluates those once into one 64-byte row per
// Gaussian,
//   f0 = (x, y, z, sigmoid(opacity))   f1 = (S00, S01, S02, S11)   f2 = (S12, S22, r, g)   f3 = (b, clamp bits, -, -)
// and k_preprocess_frozen is k_preprocess_fused reading that row with four 128-bit loads (64 B instead of 236 B per
// Gaussian, no SH evaluation, no exp / normalise / R S^2 R^T).  Both use the very functions the fused kernel uses, so
// the records -- and with them the image and every gradient -- are bit-identical (tests/test_gpu_tracking.py).
__global__ void __launch_bounds__(CTA)
k_freeze_model(CamConst cc, int P, const float *__restrict__ xyz, const float *__restrict__ f_dc,
               const float *__restrict__ f_rest, const float *__restrict__ opacity_raw,
               const float *__restrict__ scaling_raw, const float *__restrict__ rotation_raw,
               const float *__restrict__ cam_center, float4 *__restrict__ frozen) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const size_t n = (size_t)i;
    const float w[3] = {xyz[3 * n], xyz[3 * n + 1], xyz[3 * n + 2]};
    const float sc[3] = {scaling_raw[3 * n], scaling_raw[3 * n + 1], scaling_raw[3 * n + 2]};
    const float4 q4 = *reinterpret_cast<const float4 *>(rotation_raw + 4 * n);
    const float rot[4] = {q4.x, q4.y, q4.z, q4.w};
    const float dcv[3] = {f_dc[3 * n], f_dc[3 * n + 1], f_dc[3 * n + 2]};
    const float cp[3] = {__ldg(cam_center), __ldg(cam_center + 1), __ldg(cam_center + 2)};
    float c6[6], rgb[3], opacity;
    uint8_t cl = 0;
    fused_frozen_one(cc, cp, w, dcv, f_rest + 45 * n, opacity_raw[i], sc, rot, c6, opacity, rgb, cl);
    frozen[n * 4 + 0] = make_float4(w[0], w[1], w[2], opacity);
    frozen[n * 4 + 1] = make_float4(c6[0], c6[1], c6[2], c6[3]);
    frozen[n * 4 + 2] = make_float4(c6[4], c6[5], rgb[0], rgb[1]);
    frozen[n * 4 + 3] = make_float4(rgb[2], __uint_as_float((unsigned int)cl), 0.f, 0.f);
}

__global__ void __launch_bounds__(CTA)
k_preprocess_frozen(CamConst cc, int P, const float4 *__restrict__ frozen, const float *__restrict__ pose,
                    const float *__restrict__ viewmatrix, const float *__restrict__ projmatrix,
                    float4 *__restrict__ records, uint8_t *__restrict__ clamped, int *__restrict__ radii,
                    unsigned int *__restrict__ tile_count, unsigned long long *__restrict__ counters,
                    unsigned int flags, unsigned char *__restrict__ visibility, float *__restrict__ max_radii2D,
                    BinSink bs) {
    __shared__ unsigned int s_depth[CTA];
    // the CTA's rows are one contiguous 16 KB slice: fully coalesced 128-bit loads into shared memory (row stride
    // 5 float4 against bank conflicts), then every thread picks up its own row
    __shared__ float4 s_rows[CTA * 5];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    {
        const int base = blockIdx.x * blockDim.x;
        const int n4 = min((int)blockDim.x, P - base) * 4;
        const float4 *src = frozen + (size_t)base * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = k * CTA + (int)threadIdx.x;
            if (idx < n4) s_rows[(idx >> 2) * 5 + (idx & 3)] = ldg4(src + idx);
        }
        __syncthreads();
    }
    unsigned int rect_tiles = 0;
    Splat sp;
    sp.px = sp.py = sp.conx = sp.cony = sp.conz = sp.depth = 0.f; sp.radius = 0;
    float opacity = 0.f;
    bool vis = false;
    if (i < P) {
        const float4 *row = s_rows + threadIdx.x * 5;
        const float4 f0 = row[0], f1 = row[1], f2 = row[2], f3 = row[3];
        float V[16], PM[16], Rt[12];
        load16(viewmatrix, V);
        load16(projmatrix, PM);
#pragma unroll
        for (int k = 0; k < 12; ++k) Rt[k] = __ldg(pose + k);
        const float w[3] = {f0.x, f0.y, f0.z};
        const float c6[6] = {f1.x, f1.y, f1.z, f1.w, f2.x, f2.y};
        vis = fused_forward_frozen_one(cc, V, PM, Rt, w, c6, sp);
        if (vis) {
            opacity = f0.w;
            rect_tiles = (unsigned int)((sp.rmaxx - sp.rminx) * (sp.rmaxy - sp.rminy));
            store_record(records, i, sp, opacity, f2.z, f2.w, f3.x, 1);
        } else {
            store_empty_record(records, i);
        }
        clamped[i] = vis ? (uint8_t)__float_as_uint(f3.y) : (uint8_t)0;
        radii[i] = vis ? sp.radius : 0;
        if (visibility) visibility[i] = vis ? 1 : 0;
        if (max_radii2D && vis) max_radii2D[i] = fmaxf(max_radii2D[i], (float)sp.radius);
    }
    s_depth[threadIdx.x] = __float_as_uint(sp.depth);
    __syncwarp();
    const unsigned int warp_first = (unsigned int)(i - lane);
    warp_for_each_tile(cc.gx, cc.gy, vis, sp.px, sp.py, sp.conx, sp.cony, sp.conz, opacity, sp.radius, (flags & 2u) != 0,
                       lane, [&](int owner, int t) { count_instance(tile_count, bs, s_depth, warp_first, owner, t); });
    block_add_u64(&counters[CNT_RECT], rect_tiles);
}

// ---- K2: exclusive scan of the per-tile counts (single CTA; a few thousand tiles) ------------------
__global__ void __launch_bounds__(1024)
k_tile_scan(int tiles, const unsigned int *__restrict__ tile_count, unsigned int *__restrict__ tile_offset,
            unsigned int *__restrict__ cursor, unsigned long long *__restrict__ counters) {
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_max[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (tiles + 1023) / 1024;
    const int b = tid * per, e = min(tiles, b + per);
    unsigned int sum = 0, mx = 0;
    for (int t = b; t < e; ++t) { const unsigned int c = tile_count[t]; sum += c; mx = max(mx, c); }
    // inclusive scan of per-thread sums
    unsigned int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int n = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += n;
    }
    mx = __reduce_max_sync(FULL, mx);
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_max[warp] = mx;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int n = __shfl_up_sync(FULL, w, d);
            if (lane >= d) w += n;
        }
        s_warp[lane] = w;   // inclusive over warps
        const unsigned int m = __reduce_max_sync(FULL, s_max[lane]);
        if (lane == 0) s_max[0] = m;
    }
    __syncthreads();
    unsigned int run = incl - sum + (warp > 0 ? s_warp[warp - 1] : 0u);
    for (int t = b; t < e; ++t) {
        tile_offset[t] = run;
        cursor[t] = 0;
        run += tile_count[t];
    }
    if (tid == 1023) {
        tile_offset[tiles] = s_warp[31];
        counters[CNT_R] = s_warp[31];
        counters[CNT_MAXLIST] = s_max[0];
    }
}

// ---- K3: scatter (depth, id) keys into the per-tile segments ----------------------------------------
// (ncu shows 2/3 of this kernel's stall samples on the instruction that consumes the slot returned by the cursor
// atomic.  Putting 2 or 4 rounds of atomics in flight before the first key store was tried and changed nothing
// (0.053 ms either way): per instance the kernel issues one L2 atomic, one scattered 8-byte store and one offset
// load, ~7 M sector operations in ~50 us -- it sits on the L2's sector-operation rate, not on latency.  The way
// down is fewer scattered operations per instance, e.g. taking the slot from the counting pass's atomic.)
__global__ void __launch_bounds__(CTA)
k_scatter(CamConst cc, int P, const float4 *__restrict__ records, const unsigned int *__restrict__ tile_offset,
          unsigned int *__restrict__ cursor, unsigned long long *__restrict__ keys, unsigned int flags,
          const unsigned long long *__restrict__ counters, unsigned long long capacity) {
    if (counters[CNT_R] > capacity) return;   // optimistic launch into a too-small buffer: the host relaunches
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0;
    if (i < P) {
        q2 = records[(size_t)i * 3 + 2];
        if (__float_as_int(q2.z) > 0) { q0 = records[(size_t)i * 3 + 0]; q1 = records[(size_t)i * 3 + 1]; }
    }
    const bool vis = __float_as_int(q2.z) > 0;
    __shared__ unsigned int s_depth[CTA];                 // depth keys of the CTA's splats, read by other lanes
    s_depth[threadIdx.x] = __float_as_uint(q2.y);
    __syncwarp();
    const unsigned int warp_first = (unsigned int)(i - lane);
    warp_for_each_tile(cc.gx, cc.gy, vis, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, __float_as_int(q2.z), (flags & 2u) != 0, lane,
                       [&](int owner, int t) {
                           // called by an arbitrary lane: fetch nothing else from the owner but its depth key
                           const unsigned int slot = atomicAdd(&cursor[t], 1u);
                           keys[(size_t)tile_offset[t] + slot] =
                               ((unsigned long long)s_depth[(threadIdx.x & ~31) + owner] << 32) | (warp_first + owner);
                       });
}

// ---- K4+K5: per-tile sort by (depth bits, Gaussian id) and gather of the splat records -------------
// A data-independent compare-exchange network whose comparators all point the same way, so a list
// of arbitrary length n behaves as if padded with +inf to the next power of two.
// Key storage accessors: the shared-memory one names the dynamic shared array directly, so the
// compiler emits LDS/STS (a plain pointer shared between the two instantiations would force generic
// loads with run-time address-space resolution -- ncu showed 47 instructions per compare-exchange).
extern __shared__ __align__(16) unsigned long long fsgs_sort_smem[];
struct SmemKeys {
    __device__ __forceinline__ unsigned long long ld(int i) const { return fsgs_sort_smem[i]; }
    __device__ __forceinline__ void st(int i, unsigned long long v) const { fsgs_sort_smem[i] = v; }
};
struct GmemKeys {
    unsigned long long *p;
    __device__ __forceinline__ unsigned long long ld(int i) const { return p[i]; }
    __device__ __forceinline__ void st(int i, unsigned long long v) const { p[i] = v; }
};
template <typename Keys>
__device__ __forceinline__ void tile_sort_ce(const Keys &a, int lo, int hi, int n) {
    if (hi < n) {
        const unsigned long long x = a.ld(lo), y = a.ld(hi);
        if (x > y) { a.st(lo, y); a.st(hi, x); }
    }
}
// Pair p = tid + 256*i: the 32 pairs a warp handles per i cover 64 contiguous elements, so every
// step whose comparator blocks are <= 64 elements wide only needs a warp barrier; block barriers
// are reserved for the few wide steps (6 instead of 45 for a 512-entry list).
__device__ __forceinline__ void tile_sort_sync(int this_block, int next_block) {
    if (this_block > 64 || next_block > 64) __syncthreads();
    else __syncwarp();
}
template <typename Keys>
__device__ __forceinline__ void tile_sort_network(const Keys &a, int n) {
    int lgN = 0;
    while ((1 << lgN) < n) ++lgN;
    const int half = (1 << lgN) >> 1;
    for (int m = 1; m <= lgN; ++m) {
        const int k = 1 << m, hk = k >> 1;
        for (int p = threadIdx.x; p < half; p += CTA) {                  // "flip": i <-> block_end - i
            const int blk = p >> (m - 1), r = p & (hk - 1);
            tile_sort_ce(a, (blk << m) + r, (blk << m) + (k - 1 - r), n);
        }
        tile_sort_sync(k, m >= 2 ? (k >> 1) : 4);
        for (int s = m - 2; s >= 0; --s) {                                // half-cleaners, distance j = 2^s
            const int j = 1 << s;
            for (int p = threadIdx.x; p < half; p += CTA) {
                const int lo = ((p >> s) << (s + 1)) + (p & (j - 1));
                tile_sort_ce(a, lo, lo + j, n);
            }
            tile_sort_sync(2 * j, s > 0 ? j : 2 * k);
        }
    }
    __syncthreads();
}

// Shared-memory window of k_tile_sort, in 8-byte keys, chosen per launch (`win`):
//   SORT_SMEM_KEYS (64 KB, 3 CTAs/SM) when nothing is known about the list lengths or the longest list of the previous
//   frame was long; SORT_SMEM_KEYS_SMALL (32 KB, 4 CTAs/SM -- then limited by registers) when the previous frame's
//   longest list leaves 25 % headroom below win/2.  A/B on B200 (round 2, config 2): 0.093 -> 0.080 ms.
// Lists up to win/2 take the bucket path (3/4 of the window), of which lists up to win/4 also keep their unsorted keys
// in shared memory; lists up to win take the compare-exchange network in shared memory, longer ones in global memory --
// so a wrong guess costs time on the affected tiles, never correctness.
constexpr int SORT_SMEM_KEYS = 8192;
constexpr int SORT_SMEM_KEYS_SMALL = 4096;
constexpr int BUCKET_MAX_FILL = 24;    // fullest bucket the rank pass accepts before falling back to the network

// Gather one sorted instance: per-Gaussian record -> compositor record (conic pre-scaled for the
// base-2 exponent, 8x4-block reach mask of this tile in q2.z).
__device__ __forceinline__ void emit_sorted_record(const float4 *__restrict__ records, unsigned int id,
                                                   float4 *__restrict__ dst, int tile_x0, int tile_y0, bool no_cull) {
    const float4 a = ldg4(records + (size_t)id * 3), b = ldg4(records + (size_t)id * 3 + 1),
                 c = ldg4(records + (size_t)id * 3 + 2);
    float a2, b2, c2;
    scale_conic(a.z, a.w, b.x, a2, b2, c2);
    const unsigned mask = no_cull ? 0xffu : block_mask(a.x, a.y, a.z, a.w, b.x, b.y, tile_x0, tile_y0);
    dst[0] = make_float4(a.x, a.y, a2, b2);
    dst[1] = make_float4(c2, b.y, b.z, b.w);
    dst[2] = make_float4(c.x, c.y, __uint_as_float(mask), __uint_as_float(id));
}

// Bucket of a depth: monotone non-decreasing in z (float subtraction, multiplication by a positive
// constant and truncation all are), so buckets are ordered like the keys.
__device__ __forceinline__ int depth_bucket(unsigned long long key, float zmin, float inv, int nb) {
    const float z = __uint_as_float((unsigned int)(key >> 32));
    return min(nb - 1, (int)((z - zmin) * inv));
}

// Per-tile sort, bucket path (lists of up to BUCKET_MAX_KEYS entries; the usual case).  The depths of
// one tile's splats are spread roughly evenly between the tile's nearest and farthest, so a linear map
// of depth onto ~n buckets leaves about one key per bucket:
//   1. read the keys (staged in shared memory up to BUCKET_SMEM_IN entries, else re-read from global memory --
//      the three passes over them hit L1/L2), block-reduce min / max depth;
//   2. histogram (shared-memory integer atomics), block exclusive scan of the bucket counts;
//   3. scatter the keys to their bucket's segment (order inside a bucket arbitrary);
//   4. every key ranks itself inside its bucket by comparing full 64-bit (depth, id) keys -- O(fill) with
//      fill ~ 1 -- which gives its final position in the tile list; the compositor record is emitted
//      straight to that position (the sorted keys themselves are never materialised).
// ~45 instructions per key instead of the ~500 of the 45..55-step compare-exchange network.  The result is
// the same total order on unique keys, hence bit-identical.  A tile whose fullest bucket exceeds
// BUCKET_MAX_FILL (many equal depths, strongly clustered depths) returns false and the caller runs the
// network instead.
__device__ __forceinline__ bool tile_sort_bucket(int n, int win, const unsigned long long *__restrict__ g,
                                                 const float4 *__restrict__ records, float4 *__restrict__ dst,
                                                 int tile_x0, int tile_y0, bool no_cull) {
    const int bucket_max = win >> 1;                                                         // n <= bucket_max here
    unsigned long long *s_out = fsgs_sort_smem;                                              // [bucket_max]
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(fsgs_sort_smem + bucket_max);    // [nb <= bucket_max]
    unsigned long long *s_in = fsgs_sort_smem + bucket_max + (bucket_max >> 1);              // [win / 4]
    const bool in_smem = n <= (win >> 2);
    auto key_at = [&](int p) { return in_smem ? s_in[p] : __ldg(g + p); };
    __shared__ unsigned int s_red[3][CTA / 32];
    __shared__ unsigned int s_warp_sum[CTA / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + CTA - 1) / CTA;          // buckets per thread in the scan (<= 16)
    const int nb = per * CTA;                     // number of buckets: n rounded up to a multiple of the CTA

    unsigned int dmin = 0xffffffffu, dmax = 0u;
    for (int p = tid; p < n; p += CTA) {
        const unsigned long long k = __ldg(g + p);
        if (in_smem) s_in[p] = k;
        const unsigned int d = (unsigned int)(k >> 32);   // depth > 0.2: bit order == numeric order
        dmin = min(dmin, d); dmax = max(dmax, d);
    }
    for (int b = tid; b < nb; b += CTA) s_hist[b] = 0u;
    dmin = __reduce_min_sync(FULL, dmin); dmax = __reduce_max_sync(FULL, dmax);
    if (lane == 0) { s_red[0][warp] = dmin; s_red[1][warp] = dmax; }
    __syncthreads();
    dmin = s_red[0][lane & 7]; dmax = s_red[1][lane & 7];
    dmin = __reduce_min_sync(FULL, dmin); dmax = __reduce_max_sync(FULL, dmax);
    const float zmin = __uint_as_float(dmin), range = __uint_as_float(dmax) - zmin;
    const float inv = range > 0.f ? fminf((float)nb / range, 3.0e38f) : 0.f;

    for (int p = tid; p < n; p += CTA) atomicAdd(&s_hist[depth_bucket(key_at(p), zmin, inv, nb)], 1u);
    __syncthreads();

    // exclusive scan of s_hist (thread t owns buckets [t*per, (t+1)*per)) + the fullest bucket
    constexpr int PER_MAX = SORT_SMEM_KEYS / 2 / CTA;
    unsigned int cnt[PER_MAX], sum = 0, fill = 0;
#pragma unroll
    for (int i = 0; i < PER_MAX; ++i) {
        cnt[i] = i < per ? s_hist[tid * per + i] : 0u;
        sum += cnt[i]; fill = max(fill, cnt[i]);
    }
    unsigned int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
    }
    fill = __reduce_max_sync(FULL, fill);
    if (lane == 31) s_warp_sum[warp] = incl;
    if (lane == 0) s_red[2][warp] = fill;
    __syncthreads();
    unsigned int run = incl - sum;
#pragma unroll
    for (int w = 0; w < CTA / 32; ++w) {
        if (w < warp) run += s_warp_sum[w];
        fill = max(fill, s_red[2][w]);
    }
    if (fill > (unsigned int)BUCKET_MAX_FILL) return false;   // uniform: every thread sees the same maximum
#pragma unroll
    for (int i = 0; i < PER_MAX; ++i)
        if (i < per) { s_hist[tid * per + i] = run; run += cnt[i]; }
    __syncthreads();

    // scatter: afterwards s_hist[b] = END of bucket b (= start of bucket b + 1)
    for (int p = tid; p < n; p += CTA) {
        const unsigned long long k = key_at(p);
        s_out[atomicAdd(&s_hist[depth_bucket(k, zmin, inv, nb)], 1u)] = k;
    }
    __syncthreads();

    // (gathering the 48-byte records of 2 or 4 entries per thread before using the first was tried: 0.095 ms either
    // way with 2, slower with 4 -- the gather is not latency-bound at 24 resident warps per SM)
    for (int p = tid; p < n; p += CTA) {
        const unsigned long long k = s_out[p];
        const int b = depth_bucket(k, zmin, inv, nb);
        const int lo = b > 0 ? (int)s_hist[b - 1] : 0, hi = (int)s_hist[b];
        int rank = lo;
        for (int q = lo; q < hi; ++q) rank += s_out[q] < k ? 1 : 0;
        emit_sorted_record(records, (unsigned int)k, dst + (size_t)rank * 3, tile_x0, tile_y0, no_cull);
    }
    return true;
}

// 4 resident CTAs (64 registers) match the small shared-memory window; 5 (48 registers) measured 0.0805 vs 0.0823 ms,
// 6 (40 registers, spills) 0.105; without a bound the compiler takes 116 registers (2 CTAs, 0.123 ms).
#ifndef FSGS_SORT_MINB
#define FSGS_SORT_MINB 4
#endif
__global__ void __launch_bounds__(CTA, FSGS_SORT_MINB)
k_tile_sort(int gx, const unsigned int *__restrict__ tile_offset, unsigned long long *__restrict__ keys,
            const float4 *__restrict__ records, float4 *__restrict__ sorted_rec, unsigned int flags,
            const unsigned long long *__restrict__ counters, unsigned long long capacity, int win,
            unsigned long long *__restrict__ bins, unsigned int bin_cap) {
    // (bins: the counting pass already dropped the keys into fixed-stride per-tile bins; a list longer than a bin means
    // the bins are incomplete -- leave, the host relaunches the scatter path)
    if (counters[CNT_R] > capacity || (bins && counters[CNT_MAXLIST] > bin_cap)) return;
    unsigned long long *s_keys = fsgs_sort_smem;
    const unsigned int start = tile_offset[blockIdx.x];
    const int n = (int)(tile_offset[blockIdx.x + 1] - start);
    if (n == 0) return;
    const int tile_x0 = (int)(blockIdx.x % gx) * TILE, tile_y0 = (int)(blockIdx.x / gx) * TILE;
    const bool no_cull = (flags & 2u) != 0;
    unsigned long long *g = bins ? bins + (size_t)blockIdx.x * bin_cap : keys + start;
    float4 *dst = sorted_rec + (size_t)start * 3;
    if (n <= (win >> 1) && !(flags & 16u)) {                // FSGS_FLAG_SORT_NETWORK forces the network (A/B, tests)
        if (tile_sort_bucket(n, win, g, records, dst, tile_x0, tile_y0, no_cull)) return;
        __syncthreads();                                    // bucket path declined: the network, in shared memory
        for (int p = threadIdx.x; p < n; p += blockDim.x) s_keys[p] = g[p];
        __syncthreads();
        if (n > 1) tile_sort_network(SmemKeys{}, n);
        for (int p = threadIdx.x; p < n; p += blockDim.x)
            emit_sorted_record(records, (unsigned int)s_keys[p], dst + (size_t)p * 3, tile_x0, tile_y0, no_cull);
    } else if (n <= win) {
        for (int p = threadIdx.x; p < n; p += blockDim.x) s_keys[p] = g[p];
        __syncthreads();
        if (n > 1) tile_sort_network(SmemKeys{}, n);
        for (int p = threadIdx.x; p < n; p += blockDim.x) {
            const unsigned long long k = s_keys[p];   // (the sorted keys themselves are not needed again)
            emit_sorted_record(records, (unsigned int)k, dst + (size_t)p * 3, tile_x0, tile_y0, no_cull);
        }
    } else {
        // rare: list longer than the shared-memory window -> same network in global memory (L2)
        tile_sort_network(GmemKeys{g}, n);
        for (int p = threadIdx.x; p < n; p += blockDim.x)
            emit_sorted_record(records, (unsigned int)g[p], dst + (size_t)p * 3, tile_x0, tile_y0, no_cull);
    }
}

// LearnPose.forward / its backward for one frame: a single thread each (7 parameters).
// r is the [1,4,N] quaternion tensor, t the [3,N] translation tensor of the reference; `cam`
// selects the column, `n_cams` is N.
__global__ void k_pose_forward(const float *__restrict__ r, const float *__restrict__ t, int cam, int n_cams,
                               float *__restrict__ Rt) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float q[4] = {r[cam], r[n_cams + cam], r[2 * n_cams + cam], r[3 * n_cams + cam]};
    const float tt[3] = {t[cam], t[n_cams + cam], t[2 * n_cams + cam]};
    float out[16];
    pose_forward(q, tt, out);
    for (int k = 0; k < 16; ++k) Rt[k] = out[k];
}
__global__ void k_pose_backward(const float *__restrict__ r, int cam, int n_cams, const float *__restrict__ dRt,
                                float *__restrict__ dr, float *__restrict__ dt) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float q[4] = {r[cam], r[n_cams + cam], r[2 * n_cams + cam], r[3 * n_cams + cam]};
    float g[16], gq[4], gt[3];
    for (int k = 0; k < 16; ++k) g[k] = dRt[k];
    pose_backward(q, g, gq, gt);
    // dr [1,4,N], dt [3,N]: only column `cam` is non-zero (the caller zero-fills)
    for (int k = 0; k < 4; ++k) dr[k * n_cams + cam] = gq[k];
    for (int k = 0; k < 3; ++k) dt[k * n_cams + cam] = gt[k];
}

__global__ void __launch_bounds__(CTA)
k_mark_visible(int P, const float *__restrict__ means3D, const float *__restrict__ viewmatrix,
               uint8_t *__restrict__ visible) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float V[16];
    load16(viewmatrix, V);
    float x, y, z;
    xf43(V, means3D[3 * (size_t)i], means3D[3 * (size_t)i + 1], means3D[3 * (size_t)i + 2], x, y, z);
    visible[i] = z > NEAR_CULL ? 1 : 0;
}

}  // namespace fsgs
