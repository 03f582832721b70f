// fsgs_kernels_composite.cuh -- tile-order alpha compositing, forward and backward (K6, K7).
//
// One CTA per 16x16 tile, 8 warps, each warp owning an 8x4 pixel block.  The tile's depth-sorted
// splat records (48 B each, written contiguously by k_tile_sort) are streamed into shared memory
// in batches of 256 by 1-D bulk TMA copies (cp.async.bulk + mbarrier, double buffered).
//
// Sorted record: q0 = (x, y, a2, b2)  q1 = (c2, opacity, r, g)  q2 = (b, depth, block mask, -)
// with the conic pre-scaled for a base-2 exponent (fsgs_math.cuh) and an 8-bit mask saying which
// of the tile's eight 8x4 blocks the splat can reach with alpha >= 1/255.  For every staged batch
// each warp first compacts, with ballots, the indices of the entries whose mask names its block
// into a private shared-memory list and then walks only that list -- exact, because for the
// skipped entries every lane would have taken the reference's `alpha < 1/255 -> continue`.
//
//   FUSED = false : one GaussianRasterizer pass -- 3 colour planes + the package's depth plane.
//   FUSED = true  : Free-SurGS' two passes at once -- RGB | depth, silhouette, depth^2, all six
//                   planes with "+ T_final * bg" exactly as the reference's second pass produces
//                   them (gaussian_renderer/__init__.py:68-74, bg = 1 quirk iv in SURVEY.md 8a).
#pragma once

#include "fsgs_device.cuh"

// Launch bounds of the two compositors, A/B-tested on B200 (round 1).  Forward: the plain bound
// (48 registers) is the best; 6 resident CTAs spill.  Backward (two-phase): 4 resident CTAs
// (64 registers, no spills, 4 x 48 KB shared memory) beat the unconstrained 80-register build
// (0.657 vs 0.679 ms at config 2).
// (-DFSGS_FWD_MINB=n / -DFSGS_BWD_MINB=n set the min-resident-CTAs operand for A/B builds; nvcc splits
// commas inside -D values, hence the two-macro form.)
#ifdef FSGS_FWD_MINB
#define FSGS_FWD_LB CTA, FSGS_FWD_MINB
#else
#define FSGS_FWD_LB CTA
#endif
#ifndef FSGS_BWD_MINB
#define FSGS_BWD_MINB 4
#endif
#define FSGS_BWD_LB CTA, FSGS_BWD_MINB

namespace fsgs {

#ifdef FSGS_PAIR_STATS
// Instrumented A/B build only (tools/pair_stats.py): how full are the compositors' (warp, entry) visits?
//   [0] backward visits  [1] valid (pixel, entry) pairs  [2] visits with no valid pixel
//   [3] visits with exactly one of the two 4x4 halves of the 8x4 block populated
//   [4] forward visits   [5] forward contributing pairs
__device__ unsigned long long g_pair_stats[8];
#endif

// Derived per-pixel / per-Gaussian outputs of the reference's render() (fused flavour only; NULL = skip).
struct RenderExtras {
    float *uncertainty;
    unsigned char *presence_mask, *nan_mask, *visibility;
    float *max_radii2D;
};

struct TilePix {
    int px, py;
    bool inside;
};
__device__ __forceinline__ TilePix tile_pixel(const CamConst &cc, int tile) {
    const int tx = tile % cc.gx, ty = tile / cc.gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TilePix p;
    p.px = tx * TILE + (warp & 1) * 8 + (lane & 7);
    p.py = ty * TILE + (warp >> 1) * 4 + (lane >> 3);
    p.inside = p.px < cc.W && p.py < cc.H;
    return p;
}

// Stage `cnt` records starting at `src` into `dst`: bulk TMA (thread 0) or cooperative loads.
__device__ __forceinline__ void stage_issue_tma(float4 *dst, const float4 *src, int cnt, uint64_t *bar) {
    const uint32_t bytes = (uint32_t)cnt * 48u;
    mbar_arrive_expect_tx(bar, bytes);
    tma_load_1d(dst, src, bytes, bar);
}
__device__ __forceinline__ void stage_plain(float4 *dst, const float4 *src, int cnt) {
    for (int p = threadIdx.x; p < cnt * REC_F4; p += CTA) dst[p] = ldg4(src + p);
}

// Ballot-compact the entries j in [0, limit) of the staged batch whose block mask contains `warp_bit` into `list`
// (ascending).  Returns the count.  One call per warp per batch.
// The list holds the entries' BYTE OFFSETS inside the staged batch (j * 48, 16 bits): the compositors' inner loops
// then address a record as [offset + batch base] straight from the unpacked list word -- with 8-bit indices every
// visit paid a constant materialisation and an integer multiply-add for j * 48 (ncu source page, round 2: 2 of the
// forward's 33 instructions per visit).
typedef unsigned short list_t;
constexpr int REC_BYTES = REC_F4 * 16;
__device__ __forceinline__ int compact_entries(const float4 *sb, int limit, unsigned int warp_bit, int lane,
                                               list_t *list) {
    int n = 0;
    for (int c = 0; c < limit; c += 32) {
        const int j = c + lane;
        const bool rel = j < limit && (__float_as_uint(sb[j * 3 + 2].z) & warp_bit) != 0;
        const unsigned int b = __ballot_sync(FULL, rel);
        if (rel) list[n + __popc(b & ((1u << lane) - 1u))] = (list_t)(j * REC_BYTES);
        n += __popc(b);
    }
    __syncwarp();
    return n;
}

// entry s (0..7) of a chunk of 8 list entries loaded as one 128-bit word
__device__ __forceinline__ int list_off(uint4 packed, int s) {
    const unsigned int w = s < 2 ? packed.x : s < 4 ? packed.y : s < 6 ? packed.z : packed.w;
    return (int)((s & 1) ? (w >> 16) : (w & 0xffffu));
}
// the record at byte offset `off` of a staged batch
__device__ __forceinline__ const float4 *rec_at(const float4 *sb, int off) {
    return reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(sb) + off);
}

template <bool FUSED>
__global__ void __launch_bounds__(FSGS_FWD_LB)
k_composite_fwd(CamConst cc, const unsigned int *__restrict__ tile_offset, const float4 *__restrict__ sorted_rec,
                const float *__restrict__ bg, float *__restrict__ out_planes, float *__restrict__ out_depth,
                float *__restrict__ final_T, unsigned int *__restrict__ n_contrib, unsigned int flags,
                unsigned long long *__restrict__ err, const unsigned long long *__restrict__ counters,
                unsigned long long capacity, RenderExtras ex, unsigned int bin_cap) {
    __shared__ __align__(128) float4 s_rec[2][BATCH * REC_F4];
    __shared__ __align__(8) uint64_t s_full[2];
    __shared__ __align__(16) list_t s_list[CTA / 32][BATCH];   // rows read 8 entries (16 B) at a time
    // optimistic launch into a too-small buffer (or from incomplete bins, bin_cap > 0): the host relaunches
    if (counters[CNT_R] > capacity || (bin_cap && counters[CNT_MAXLIST] > bin_cap)) return;
    const int tile = blockIdx.x;
    const unsigned int start = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - start);
    const int nb = (n + BATCH - 1) / BATCH;
    const bool use_tma = (flags & 1u) == 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int warp_bit = 1u << warp;
    const TilePix pix = tile_pixel(cc, tile);
    const float pxf = (float)pix.px, pyf = (float)pix.py;
    const float4 *src = sorted_rec + (size_t)start * REC_F4;

    if (use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(&s_full[0], 1);
            mbar_init(&s_full[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0 && nb > 0) stage_issue_tma(s_rec[0], src, min(BATCH, n), &s_full[0]);
    }

    // A finished pixel (T would drop below 1e-4, or outside the image) is marked by the SIGN of T: the next
    // entry's test_T = T * (1 - alpha) is then negative, fails `test_T >= T_MIN` and re-marks the pixel --
    // no separate flag, no per-entry "done" branch.  The body is branch-free (selects), entries are taken in
    // aligned chunks of 8 list positions (one 64-bit load of the list bytes per chunk).
    float T = pix.inside ? 1.f : -1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, S = 0.f, D2 = 0.f;
    unsigned int last = 0;

    for (int k = 0; k < nb; ++k) {
        const int all_done = __syncthreads_and(T < 0.f ? 1 : 0);   // everyone is also past batch k-1
        const int buf = k & 1;
        const int cnt = min(BATCH, n - k * BATCH);
        if (use_tma) {
            if (threadIdx.x == 0 && !all_done && k + 1 < nb)
                stage_issue_tma(s_rec[buf ^ 1], src + (size_t)(k + 1) * BATCH * REC_F4, min(BATCH, n - (k + 1) * BATCH),
                                &s_full[buf ^ 1]);
            mbar_wait(&s_full[buf], (uint32_t)(k >> 1) & 1u, err);   // drain even when leaving
            if (all_done) break;
        } else {
            if (all_done) break;
            stage_plain(s_rec[buf], src + (size_t)k * BATCH * REC_F4, cnt);
            __syncthreads();
        }
        if (__all_sync(FULL, T < 0.f)) continue;                    // this warp's pixels are finished
        const float4 *sb = s_rec[buf];
        const int nrel = compact_entries(sb, cnt, warp_bit, lane, s_list[warp]);
        int lj = -1;                                                // last contributing entry of this batch (byte offset)
        auto entry = [&](int j) __attribute__((always_inline)) {    // j = byte offset of the record in the batch
            const float4 *r = rec_at(sb, j);
            const float4 q0 = r[0], q1 = r[1];
            const float2 q2 = *reinterpret_cast<const float2 *>(r + 2);   // (b, depth)
            const float dx = q0.x - pxf, dy = q0.y - pyf;
            const float p2 = gauss_power2(q0.z, q0.w, q1.x, dx, dy);
            const float alpha = fminf(ALPHA_MAX, q1.y * fast_exp2(p2));
            const float test_T = T * (1.f - alpha);
            const bool pa = (p2 <= 0.f) & (alpha >= ALPHA_MIN);
            const bool ok = pa & (test_T >= T_MIN);
#ifdef FSGS_PAIR_STATS
            {
                const unsigned int bv = __ballot_sync(FULL, ok);
                if (lane == 0) { atomicAdd(&g_pair_stats[4], 1ull); atomicAdd(&g_pair_stats[5], (unsigned long long)__popc(bv)); }
            }
#endif
            const float w = ok ? alpha * T : 0.f;
            C0 = fmaf(q1.z, w, C0); C1 = fmaf(q1.w, w, C1); C2 = fmaf(q2.x, w, C2); D = fmaf(q2.y, w, D);
            if (FUSED) { S += w; D2 = fmaf(q2.y * q2.y, w, D2); }
            if (pa & !ok) T = -fabsf(T);
            T = ok ? test_T : T;
            lj = ok ? j : lj;
        };
        for (int c0 = 0; c0 < nrel; c0 += 8) {
            if (__all_sync(FULL, T < 0.f)) break;
            const uint4 packed = *reinterpret_cast<const uint4 *>(&s_list[warp][c0]);
            if (c0 + 8 <= nrel) {
#pragma unroll
                for (int s = 0; s < 8; ++s) entry(list_off(packed, s));
            } else {
                for (int s = 0; s < nrel - c0; ++s) entry((int)s_list[warp][c0 + s]);
            }
        }
        if (lj >= 0) last = (unsigned int)(k * BATCH + lj / REC_BYTES + 1);
    }
    T = fabsf(T);

    if (pix.inside) {
        const size_t HW = (size_t)cc.W * cc.H, p = (size_t)pix.py * cc.W + pix.px;
        const float b0 = __ldg(bg), b1 = __ldg(bg + 1), b2 = __ldg(bg + 2);
        final_T[p] = T;
        n_contrib[p] = last;
        out_planes[p] = C0 + T * b0;
        out_planes[HW + p] = C1 + T * b1;
        out_planes[2 * HW + p] = C2 + T * b2;
        if (FUSED) {
            out_planes[3 * HW + p] = D + T * b0;
            out_planes[4 * HW + p] = S + T * b1;
            out_planes[5 * HW + p] = D2 + T * b2;
            // render()'s derived maps, bit-identical to the element-wise torch formulation
            const float dep = D + T * b0, sil = S + T * b1, dsq = D2 + T * b2;
            const float unc = __fsub_rn(dsq, __fmul_rn(dep, dep));
            if (ex.uncertainty) ex.uncertainty[p] = unc;
            if (ex.presence_mask) ex.presence_mask[p] = sil > 0.3f ? 1 : 0;
            if (ex.nan_mask) ex.nan_mask[p] = (dep == dep && unc == unc) ? 1 : 0;
        } else {
            out_depth[p] = D;
        }
    }
}

// ---- backward: transposed two-phase formulation ----------------------------------------------------
// Back-to-front replay over the first max(n_contrib) entries of the tile list.  Each contributing
// (pixel, Gaussian) pair produces 12 moments; a first formulation (round-1 history, removed) combined the
// 32 lanes' moments with a 13-shuffle reduce-scatter per (warp, entry) and spent ~50 of its ~150
// instructions per visit there.  Here the reduction is replaced by a change of thread layout through
// shared memory:
//
//   phase A (lane = pixel of the warp's 8x4 block, as in the forward): for each relevant entry,
//     evaluate alpha, advance the pixel's replay state and write the three scalars the moments are
//     linear in -- q = G*o*dL/dalpha, w = alpha*T, q_rgb -- to s_pair[component][slot][pixel].
//     Non-contributing lanes go through the same arithmetic with alpha = G*o = 0 (an exact no-op on
//     the state, see bwd_pair_weights), so the phase is branch-free.  Entries are taken in aligned
//     chunks of PCHUNK = 8 list positions (one 64-bit load of the 8 list bytes per chunk).
//   phase B (once per chunk; lane = (slot, pixel row)): each lane sums ITS entry's moments over the
//     8 pixels of ITS row in registers -- dy is constant along the row, so only S0, Sx, Sxx (+ the
//     RGB-only pair and the colour sums) are accumulated and Sy, Sxy, Syy follow from dy -- then a
//     2-step reduce-scatter over the 4 rows (9 shuffles per 8 entries instead of 13 per entry), the
//     moments -> gradient-row map (linear, so it can be applied to the warp's partial sums), and ONE
//     red.global.add.v4.f32 per lane (3 lanes x 16 B = the entry's 48-byte accumulator row).
//     There is no shared-memory accumulator: float atomics on shared memory are CAS loops
//     (ATOMS.CAST.SPIN, ~10 wavefronts each), the L2 does them natively.
//
// Shared memory (dynamic, ~53 KB -> 4 CTAs/SM): records 2 x 192 x 48 B (bulk-TMA ring),
// s_pair 8 warps x 3 x 8 x 32 floats (pixel index XOR-swizzled by the slot's low bit so that both
// the phase-A scalar stores and the phase-B 128-bit row loads are bank-conflict free), per-pixel
// upstream gradients 8 warps x 2 x 4 x 9 float4 (row stride 9 for the same reason).
// Record staging ring: A/B on B200 (round 2, config 2): 128 x 3 stages 0.4753 ms, 192 x 2 stages 0.4719 (fewer batches:
// less per-batch compaction and fewer partly filled phase-B rounds; same 18 KB), 160 x 2 0.4911*, 144 x 3 0.4928*
// (* before the phase-B address clean-up, against 0.4963 for 128 x 3).
#ifndef FSGS_BWD_BATCH
#define FSGS_BWD_BATCH 192
#endif
#ifndef FSGS_BWD_STAGES
#define FSGS_BWD_STAGES 2
#endif
constexpr int BWD_BATCH = FSGS_BWD_BATCH;
constexpr int PCHUNK = 8;
constexpr int PAIR_COMP = 3;
constexpr int NWARP = CTA / 32;
constexpr int SG_ROW = 9;   // float4 per pixel row of s_g (8 used)

constexpr int BWD_STAGES = FSGS_BWD_STAGES;   // record staging ring (2 x 9 KB): warps may drift one batch apart
struct BwdSmem {
    float4 rec[BWD_STAGES][BWD_BATCH * REC_F4];
    float pair[NWARP][PAIR_COMP][PCHUNK][32];
    float4 g[NWARP][2][4 * SG_ROW];
    uint64_t full[BWD_STAGES];
    alignas(16) list_t list[NWARP][BWD_BATCH];   // byte offsets; 16-byte aligned rows (read 8 entries at a time)
    unsigned int done_cnt[BWD_STAGES];      // warps finished with the batch currently in each stage
    unsigned int maxlast;
};

// Phase B for the first `cn` slots of this warp's chunk.  All 32 lanes call it.
//   1. lane (entry e, pixel row r) accumulates its row's sums in registers, with the pixel column i = 0..7
//      as a compile-time constant:  S0 = sum q, S1 = sum q i, S2 = sum q i^2, R0, R1 (RGB-only q), colour sums;
//      the moments about the entry's centre follow from dx = dx0 - i:  Sx = dx0 S0 - S1,
//      Sxx = dx0^2 S0 - 2 dx0 S1 + S2, and Sy, Sxy, Syy from the row's constant dy.
//   2. the moments -> accumulator-row map (bwd_finalize; linear, so it may be applied to a row's partial
//      sums) is applied by EVERY lane to its own 12 values -- uniform code, no per-row branches;
//   3. the four rows of an entry are combined through shared memory: each lane stores its 12 values as
//      3 x float4 into the (now consumed) pair buffer of its warp, then lane (e, k < 3) adds the four rows
//      of float4 k and issues ONE red.global.add.v4.f32 -- 48 B per entry in 3 lanes.
//      (A shuffle reduce-scatter + regroup of 12 values over 4 lanes cost ~60 instructions here; this is ~25.)
//
// POSE_ONLY (fused flavour, tracking with a frozen Gaussian model: only dL/dpose is wanted): the colour, opacity
// and RGB-only screen-space columns of the accumulator row are not needed -- the pose reaches a splat only through
// its mean (2-D mean, conic via the Jacobian, view depth) -- so their sums, the w / q_rgb planes of the pair buffer
// (w is still needed for the depth column when a depth-side gradient is present) and the third flush are dropped.
// Phase B addressed through C++ pointers into the shared array: the form the POSE-ONLY kernel keeps (its phase A is
// the tighter one; with the explicit-address form below it spilled and ran 0.350 instead of 0.341 ms).
template <bool FUSED, int LEVEL, bool POSE_ONLY>
__device__ __forceinline__ void bwd_phase_b_ptr(BwdSmem &sm, const float4 *sbatch, int warp, int lane,
                                            int cn, uint4 packed, float bx, float by, float kx, float ky,
                                            float *__restrict__ grad_acc) {
    const int e = lane >> 2, row = lane & 3;
    const bool act = e < cn;
    const float4 *sb = rec_at(sbatch, list_off(packed, e));   // this lane's entry record
    constexpr int j = 0;
    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0, o2 = o0;
    if (act) {
        const float4 q0 = sb[j * 3];
        const float dx0 = q0.x - bx, dy = q0.y - (by + (float)row);
        float S0 = 0.f, S1 = 0.f, S2 = 0.f, R0 = 0.f, R1 = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, cz = 0.f, cz2 = 0.f;
        const int sw = (e & 1);   // swizzle of the 16-byte chunk index
        const float4 *pq = reinterpret_cast<const float4 *>(&sm.pair[warp][0][e][0]);
        const float4 *pw = reinterpret_cast<const float4 *>(&sm.pair[warp][1][e][0]);
        const float4 *pr = reinterpret_cast<const float4 *>(&sm.pair[warp][2][e][0]);
        const float4 *pg = &sm.g[warp][0][row * SG_ROW];
        const float4 *pg2 = &sm.g[warp][1][row * SG_ROW];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = (row * 2 + h) ^ sw;
            const float4 qv = pq[c];
            float4 wv = make_float4(0.f, 0.f, 0.f, 0.f), rv = wv;
            if (!POSE_ONLY || LEVEL >= 1) wv = pw[c];
            if (FUSED && LEVEL >= 1 && !POSE_ONLY) rv = pr[c];
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, wa[4] = {wv.x, wv.y, wv.z, wv.w},
                        ra[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = h * 4 + k;
                const float fi = (float)i, fi2 = (float)(i * i);
                S0 += qa[k];
                if (i > 0) { S1 = fmaf(qa[k], fi, S1); S2 = fmaf(qa[k], fi2, S2); }
                if (POSE_ONLY) {
                    if (LEVEL >= 1) cz = fmaf(wa[k], pg[i].w, cz);
                } else {
                    const float4 g4 = pg[i];
                    if (FUSED && LEVEL >= 1) { R0 += ra[k]; if (i > 0) R1 = fmaf(ra[k], fi, R1); }
                    cr = fmaf(wa[k], g4.x, cr); cg = fmaf(wa[k], g4.y, cg); cb = fmaf(wa[k], g4.z, cb);
                    if (LEVEL >= 1) cz = fmaf(wa[k], g4.w, cz);
                }
                if (FUSED && LEVEL >= 2) cz2 = fmaf(wa[k], pg2[i].y, cz2);
            }
        }
        if (!FUSED || LEVEL == 0) { R0 = S0; R1 = S1; }   // no depth-side gradient (or API flavour): q_rgb == q
        const float Sx = fmaf(dx0, S0, -S1), Rx = fmaf(dx0, R0, -R1);
        const float Sxx = fmaf(dx0, fmaf(dx0, S0, -2.f * S1), S2);
        const float Sy = dy * S0, Sxy = dy * Sx, Syy = dy * Sy, Ry = dy * R0;
        if (FUSED && LEVEL >= 2) cz = fmaf(2.f * sb[j * 3 + 2].y, cz2, cz);
        // moments -> accumulator row (bwd_finalize in fsgs_math.cuh) on this row's partial sums
        const float2 q1 = *reinterpret_cast<const float2 *>(&sb[j * 3 + 1]);   // (c2, opacity)
        float A, B, C;
        unscale_conic(q0.z, q0.w, q1.x, A, B, C);
        o0 = make_float4(-kx * (A * Sx + B * Sy), -ky * (C * Sy + B * Sx), -0.5f * Sxx, -0.5f * Sxy);
        if (POSE_ONLY) {
            o1 = make_float4(-0.5f * Syy, 0.f, 0.f, 0.f);
            o2 = make_float4(0.f, cz, 0.f, 0.f);
        } else {
            o1 = make_float4(-0.5f * Syy, S0 * fast_rcp(q1.y), cr, cg);
            o2 = make_float4(cb, cz, -kx * (A * Rx + B * Ry), -ky * (C * Ry + B * Rx));
        }
    }
    __syncwarp();                                   // every lane is done reading the pair buffer
    // [entry][row][3] float4 = 1536 B of 3072.  (The four-row loads below put the eight entries of a quarter-warp on two
    // 16-byte bank groups -- ncu: 5.4 M shared-memory bank conflicts per launch.  Padding the entry stride to 13 float4
    // removes them and was measured SLOWER, 0.511 vs 0.494 ms: the kernel is issue bound and the padded indexing costs
    // more instructions than the conflicts cost cycles.  Round 2 A/B, tools/ab_variants.py.)
    float4 *tp = reinterpret_cast<float4 *>(&sm.pair[warp][0][0][0]);
    if (act) {
        tp[lane * 3] = o0; tp[lane * 3 + 1] = o1;
        if (!POSE_ONLY || LEVEL >= 1) tp[lane * 3 + 2] = o2;
    }
    __syncwarp();
    if (act && row < ((POSE_ONLY && LEVEL == 0) ? 2 : 3)) {
        const float4 r0 = tp[(e * 4) * 3 + row], r1 = tp[(e * 4 + 1) * 3 + row], r2 = tp[(e * 4 + 2) * 3 + row],
                     r3 = tp[(e * 4 + 3) * 3 + row];
        float4 o;
        o.x = (r0.x + r1.x) + (r2.x + r3.x); o.y = (r0.y + r1.y) + (r2.y + r3.y);
        o.z = (r0.z + r1.z) + (r2.z + r3.z); o.w = (r0.w + r1.w) + (r2.w + r3.w);
        if ((o.x != 0.f) | (o.y != 0.f) | (o.z != 0.f) | (o.w != 0.f)) {
            const unsigned int gid = __float_as_uint(sb[j * 3 + 2].w);
            red_add_v4(reinterpret_cast<float4 *>(grad_acc + (size_t)gid * ACC_F) + row, o);
        }
    }
}

// Per-lane shared-window addresses of phase B, computed once per kernel (lane = (entry e, pixel row)):
//   pair  : s_pair[warp][0][e][0] + the row's first 16-byte chunk, swizzled (the second chunk is pair ^ 16;
//           components w / q_rgb at +1024 / +2048)
//   g     : the row's upstream gradients, 8 x float4 (plane 1 -- silhouette / depth^2 -- at +4 * SG_ROW * 16)
//   tp_st : this lane's 3 float4 of the row-combine buffer;  tp_ld : float4 `row` of the entry's first row
//   list  : this lane's entry of a chunk, s_list[warp][e] (+ the chunk's first list position)
struct BwdLaneAddr {
    uint32_t pair, g, tp_st, tp_ld, list;
};
// KEEP: pin the addresses in registers (the general kernel); the pose-only kernel, whose phase A is the tighter one,
// was measured faster when the compiler stays free to re-derive them (0.341 vs 0.351 ms).
template <bool KEEP>
__device__ __forceinline__ uint32_t keep_if(uint32_t v) { return KEEP ? keep_reg(v) : v; }
template <bool KEEP>
__device__ __forceinline__ BwdLaneAddr bwd_lane_addr(BwdSmem &sm, int warp, int lane) {
    const int e = lane >> 2, row = lane & 3;
    BwdLaneAddr a;
    a.pair = keep_if<KEEP>(smem_u32(&sm.pair[warp][0][e][0]) + (uint32_t)(((row * 2) ^ (e & 1)) * 16));
    a.g = keep_if<KEEP>(smem_u32(&sm.g[warp][0][row * SG_ROW]));
    const uint32_t tp = smem_u32(&sm.pair[warp][0][0][0]);
    a.tp_st = keep_if<KEEP>(tp + (uint32_t)(lane * 48));
    a.tp_ld = keep_if<KEEP>(tp + (uint32_t)((e * 4 * 3 + row) * 16));
    a.list = keep_if<KEEP>(smem_u32(&sm.list[warp][e]));
    return a;
}
constexpr int PAIR_COMP_BYTES = PCHUNK * 32 * 4;          // one component plane of a warp's pair buffer
constexpr int SG_PLANE_BYTES = 4 * SG_ROW * 16;           // one plane of a warp's upstream-gradient buffer

template <bool FUSED, int LEVEL, bool POSE_ONLY>
__device__ __forceinline__ void bwd_phase_b(const BwdLaneAddr &la, uint32_t sb_a, int lane,
                                            int cn, int c0, float bx, float by, float kx, float ky,
                                            float *__restrict__ grad_acc) {
    const int e = lane >> 2, row = lane & 3;
    // lanes of the first cn entries; they alone touch the pair buffer below, so they alone synchronise
    const unsigned int amask = cn >= PCHUNK ? FULL : ((1u << (4 * cn)) - 1u);
    if (e < cn) {
        const uint32_t rec = sb_a + lds_u16(la.list + (uint32_t)(2 * c0));      // this lane's entry record
        float4 o0, o1, o2;
        const float4 q0 = lds128<0>(rec);
        const float dx0 = q0.x - bx, dy = q0.y - (by + (float)row);
        float S0 = 0.f, S1 = 0.f, S2 = 0.f, R0 = 0.f, R1 = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, cz = 0.f, cz2 = 0.f;
        const uint32_t p0 = la.pair, p1 = la.pair ^ 16u;
        auto half = [&](const float4 qv, const float4 wv, const float4 rv, const float4 g0, const float4 g1, const float4 g2,
                        const float4 g3, const float4 h01, const float4 h23, int hh) __attribute__((always_inline)) {
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, wa[4] = {wv.x, wv.y, wv.z, wv.w},
                        ra[4] = {rv.x, rv.y, rv.z, rv.w};
            const float4 ga[4] = {g0, g1, g2, g3};
            const float g5[4] = {h01.x, h01.y, h23.x, h23.y};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = hh * 4 + k;
                const float fi = (float)i, fi2 = (float)(i * i);
                S0 += qa[k];
                if (i > 0) { S1 = fmaf(qa[k], fi, S1); S2 = fmaf(qa[k], fi2, S2); }
                if (POSE_ONLY) {
                    if (LEVEL >= 1) cz = fmaf(wa[k], ga[k].w, cz);
                } else {
                    if (FUSED && LEVEL >= 1) { R0 += ra[k]; if (i > 0) R1 = fmaf(ra[k], fi, R1); }
                    cr = fmaf(wa[k], ga[k].x, cr); cg = fmaf(wa[k], ga[k].y, cg); cb = fmaf(wa[k], ga[k].z, cb);
                    if (LEVEL >= 1) cz = fmaf(wa[k], ga[k].w, cz);
                }
                if (FUSED && LEVEL >= 2) cz2 = fmaf(wa[k], g5[k], cz2);
            }
        };
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        constexpr bool NEED_W = !POSE_ONLY || LEVEL >= 1, NEED_R = FUSED && LEVEL >= 1 && !POSE_ONLY;
        constexpr bool NEED_G = !POSE_ONLY || LEVEL >= 1, NEED_G5 = FUSED && LEVEL >= 2;
        // (depth^2-plane gradients: .y of plane 1; two pixels per 128-bit load would need a different layout -- this
        //  level only occurs when a caller differentiates the detached depth^2 output)
        {
            const float4 qv = lds128<0>(p0);
            const float4 wv = NEED_W ? lds128<PAIR_COMP_BYTES>(p0) : z4;
            const float4 rv = NEED_R ? lds128<2 * PAIR_COMP_BYTES>(p0) : z4;
            const float4 g0 = NEED_G ? lds128<0>(la.g) : z4, g1 = NEED_G ? lds128<16>(la.g) : z4,
                         g2 = NEED_G ? lds128<32>(la.g) : z4, g3 = NEED_G ? lds128<48>(la.g) : z4;
            float4 h01 = z4, h23 = z4;
            if (NEED_G5) {
                h01 = make_float4(lds128<SG_PLANE_BYTES>(la.g).y, lds128<SG_PLANE_BYTES + 16>(la.g).y, 0.f, 0.f);
                h23 = make_float4(lds128<SG_PLANE_BYTES + 32>(la.g).y, lds128<SG_PLANE_BYTES + 48>(la.g).y, 0.f, 0.f);
            }
            half(qv, wv, rv, g0, g1, g2, g3, h01, h23, 0);
        }
        {
            const float4 qv = lds128<0>(p1);
            const float4 wv = NEED_W ? lds128<PAIR_COMP_BYTES>(p1) : z4;
            const float4 rv = NEED_R ? lds128<2 * PAIR_COMP_BYTES>(p1) : z4;
            const float4 g0 = NEED_G ? lds128<64>(la.g) : z4, g1 = NEED_G ? lds128<80>(la.g) : z4,
                         g2 = NEED_G ? lds128<96>(la.g) : z4, g3 = NEED_G ? lds128<112>(la.g) : z4;
            float4 h01 = z4, h23 = z4;
            if (NEED_G5) {
                h01 = make_float4(lds128<SG_PLANE_BYTES + 64>(la.g).y, lds128<SG_PLANE_BYTES + 80>(la.g).y, 0.f, 0.f);
                h23 = make_float4(lds128<SG_PLANE_BYTES + 96>(la.g).y, lds128<SG_PLANE_BYTES + 112>(la.g).y, 0.f, 0.f);
            }
            half(qv, wv, rv, g0, g1, g2, g3, h01, h23, 1);
        }
        if (!FUSED || LEVEL == 0) { R0 = S0; R1 = S1; }   // no depth-side gradient (or API flavour): q_rgb == q
        const float Sx = fmaf(dx0, S0, -S1), Rx = fmaf(dx0, R0, -R1);
        const float Sxx = fmaf(dx0, fmaf(dx0, S0, -2.f * S1), S2);
        const float Sy = dy * S0, Sxy = dy * Sx, Syy = dy * Sy, Ry = dy * R0;
        if (FUSED && LEVEL >= 2) cz = fmaf(2.f * lds32<36>(rec), cz2, cz);
        // moments -> accumulator row (bwd_finalize in fsgs_math.cuh) on this row's partial sums
        const float2 q1 = lds64<16>(rec);   // (c2, opacity)
        float A, B, C;
        unscale_conic(q0.z, q0.w, q1.x, A, B, C);
        o0 = make_float4(-kx * (A * Sx + B * Sy), -ky * (C * Sy + B * Sx), -0.5f * Sxx, -0.5f * Sxy);
        if (POSE_ONLY) {
            o1 = make_float4(-0.5f * Syy, 0.f, 0.f, 0.f);
            o2 = make_float4(0.f, cz, 0.f, 0.f);
        } else {
            o1 = make_float4(-0.5f * Syy, S0 * fast_rcp(q1.y), cr, cg);
            o2 = make_float4(cb, cz, -kx * (A * Rx + B * Ry), -ky * (C * Ry + B * Rx));
        }
        __syncwarp(amask);                              // every lane is done reading the pair buffer
    // [entry][row][3] float4 = 1536 B of 3072.  (The four-row loads below put the eight entries of a quarter-warp on two
    // 16-byte bank groups -- ncu: 5.4 M shared-memory bank conflicts per launch.  Padding the entry stride to 13 float4
    // removes them and was measured SLOWER, 0.511 vs 0.494 ms: the kernel is issue bound and the padded indexing costs
    // more instructions than the conflicts cost cycles.  Round 2 A/B, tools/ab_variants.py.)
        sts128<0>(la.tp_st, o0); sts128<16>(la.tp_st, o1);
        if (!POSE_ONLY || LEVEL >= 1) sts128<32>(la.tp_st, o2);
        __syncwarp(amask);
        if (row < ((POSE_ONLY && LEVEL == 0) ? 2 : 3)) {
            const float4 r0 = lds128<0>(la.tp_ld), r1 = lds128<48>(la.tp_ld), r2 = lds128<96>(la.tp_ld),
                         r3 = lds128<144>(la.tp_ld);
            float4 o;
            o.x = (r0.x + r1.x) + (r2.x + r3.x); o.y = (r0.y + r1.y) + (r2.y + r3.y);
            o.z = (r0.z + r1.z) + (r2.z + r3.z); o.w = (r0.w + r1.w) + (r2.w + r3.w);
            if ((o.x != 0.f) | (o.y != 0.f) | (o.z != 0.f) | (o.w != 0.f)) {
                const unsigned int gid = __float_as_uint(lds32<44>(rec));
                red_add_v4(reinterpret_cast<float4 *>(grad_acc + (size_t)gid * ACC_F) + row, o);
            }
        }
    }
}

// One staged batch, back to front, for one warp (phase A + embedded phase B).  last_rel: this pixel's contributor
// count relative to the batch, as a byte offset (entries at offsets >= last_rel lie behind the pixel's last one).
template <bool FUSED, int LEVEL, bool POSE_ONLY>
__device__ __forceinline__ void bwd_batch(BwdSmem &sm, const float4 *sb, int warp, int lane,
                                          int nrel, int last_rel, float pxf, float pyf, float bx, float by,
                                          float kx, float ky, const float *g, float T_final, float bgdot_rgb,
                                          float bgdot_dep, BwdPixel &ps, float *__restrict__ grad_acc,
                                          const BwdLaneAddr &la) {
    const uint32_t sb_a = smem_u32(sb);
    for (int c0 = ((nrel - 1) / PCHUNK) * PCHUNK; c0 >= 0; c0 -= PCHUNK) {
        const int cn = min(PCHUNK, nrel - c0);                 // this chunk: list positions c0 .. c0+cn-1
        const uint4 packed = *reinterpret_cast<const uint4 *>(&sm.list[warp][c0]);
        auto slot = [&](int s) __attribute__((always_inline)) {
            const int j = list_off(packed, s);                                     // byte offset of the record
            const float4 *r = rec_at(sb, j);
            const float4 q0 = r[0], q1 = r[1];
            const float2 q2 = *reinterpret_cast<const float2 *>(r + 2);            // (b, depth)
            const float dx = q0.x - pxf, dy = q0.y - pyf;
            const float p2 = gauss_power2(q0.z, q0.w, q1.x, dx, dy);
            const float G = fast_exp2(p2);
            const float alpha = fminf(ALPHA_MAX, q1.y * G);
            const bool valid = (j < last_rel) && (p2 <= 0.f) && (alpha >= ALPHA_MIN);
#ifdef FSGS_PAIR_STATS
            {
                const unsigned int bv = __ballot_sync(FULL, valid);
                if (lane == 0) {
                    atomicAdd(&g_pair_stats[0], 1ull);
                    atomicAdd(&g_pair_stats[1], (unsigned long long)__popc(bv));
                    if (!bv) atomicAdd(&g_pair_stats[2], 1ull);
                    else if (!(bv & 0x0f0f0f0fu) || !(bv & 0xf0f0f0f0u)) atomicAdd(&g_pair_stats[3], 1ull);
                }
            }
#endif
            float q, w, q_rgb;
            bwd_pair_weights<FUSED, LEVEL>(ps, valid ? G * q1.y : 0.f, valid ? alpha : 0.f, q1.z, q1.w, q2.x, q2.y,
                                           g, T_final, bgdot_rgb, bgdot_dep, q, w, q_rgb);
            const int px = lane ^ ((s & 1) << 2);
            sm.pair[warp][0][s][px] = q;
            if (!POSE_ONLY || LEVEL >= 1) sm.pair[warp][1][s][px] = w;
            if (FUSED && LEVEL >= 1 && !POSE_ONLY) sm.pair[warp][2][s][px] = q_rgb;
        };
        if (cn == PCHUNK) {                                       // all but the first-visited chunk of a batch
#pragma unroll
            for (int s = PCHUNK - 1; s >= 0; --s) slot(s);        // back to front
        } else {
#pragma unroll
            for (int s = PCHUNK - 1; s >= 0; --s)
                if (s < cn) slot(s);
        }
        __syncwarp();
        if (POSE_ONLY) bwd_phase_b_ptr<FUSED, LEVEL, POSE_ONLY>(sm, sb, warp, lane, cn, packed, bx, by, kx, ky, grad_acc);
        else bwd_phase_b<FUSED, LEVEL, POSE_ONLY>(la, sb_a, lane, cn, c0, bx, by, kx, ky, grad_acc);
        __syncwarp();
    }
}

template <bool FUSED, bool POSE_ONLY = false>
__global__ void __launch_bounds__(FSGS_BWD_LB)
k_composite_bwd(CamConst cc, const unsigned int *__restrict__ tile_offset, const float4 *__restrict__ sorted_rec,
                const float *__restrict__ bg, const float *__restrict__ final_T,
                const unsigned int *__restrict__ n_contrib, const float *__restrict__ g_rgb,
                const float *__restrict__ g_depth, const float *__restrict__ g_sil, const float *__restrict__ g_dsq,
                float *__restrict__ grad_acc, unsigned int flags, unsigned long long *__restrict__ err,
                const unsigned long long *__restrict__ counters, unsigned long long capacity, unsigned int bin_cap) {
    // (the forward's tail did not run if the frame had more instances than the buffer it was given, or a longer tile
    // list than the bins it was given: see FSGS_FLAG_FIXED_CAPACITY; err = the device's sticky watchdog word)
    if (counters[CNT_R] > capacity || (bin_cap && counters[CNT_MAXLIST] > bin_cap)) return;
    // upstream gradients: g_rgb[3,H,W] and one [H,W] plane each for depth | silhouette | depth^2 (fused
    // flavour; the API flavour has the package's depth output in g_depth).  A NULL plane is all zeros.
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw);
    const int tile = blockIdx.x;
    const unsigned int start = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - start);
    if (n == 0) return;
    const bool use_tma = (flags & 1u) == 0;
    const int warp = threadIdx.x >> 5, lane = (int)keep_if<!POSE_ONLY>(threadIdx.x & 31);   // (kept: not re-read from %tid where used)
    const unsigned int warp_bit = 1u << warp;
    const TilePix pix = tile_pixel(cc, tile);
    const float pxf = (float)pix.px, pyf = (float)pix.py;
    const float bx = (float)(pix.px - (lane & 7)), by = (float)(pix.py - (lane >> 3));   // block corner
    const size_t HW = (size_t)cc.W * cc.H, p = (size_t)pix.py * cc.W + pix.px;

    if (threadIdx.x == 0) {
        sm.maxlast = 0;
        for (int st = 0; st < BWD_STAGES; ++st) {
            sm.done_cnt[st] = 0u;
            if (use_tma) mbar_init(&sm.full[st], 1);
        }
        if (use_tma) mbar_fence_init();
    }
    __syncthreads();

    const int last = pix.inside ? (int)n_contrib[p] : 0;
    const int warp_last = (int)__reduce_max_sync(FULL, (unsigned int)last);   // this warp's deepest contributor
    if (lane == 0 && warp_last) atomicMax(&sm.maxlast, (unsigned int)warp_last);
    __syncthreads();
    const int maxlast = min((int)sm.maxlast, n);
    if (maxlast == 0) return;
    const int nb = (maxlast + BWD_BATCH - 1) / BWD_BATCH;

    constexpr int NG = FUSED ? 6 : 4;   // pixel gradients: planes (+ depth plane for the API flavour)
    float g[NG];
    const float T_final = pix.inside ? final_T[p] : 0.f;
    float bgdot_rgb = 0.f, bgdot_dep = 0.f;
    {
        const float b0 = __ldg(bg), b1 = __ldg(bg + 1), b2 = __ldg(bg + 2);
#pragma unroll
        for (int ch = 0; ch < NG; ++ch) g[ch] = 0.f;
        if (pix.inside) {
            if (g_rgb) { g[0] = g_rgb[p]; g[1] = g_rgb[HW + p]; g[2] = g_rgb[2 * HW + p]; }
            if (g_depth) g[3] = g_depth[p];
            if (FUSED) {
                if (g_sil) g[4] = g_sil[p];
                if (g_dsq) g[5] = g_dsq[p];
                bgdot_dep = b0 * g[3] + b1 * g[4] + b2 * g[5];
            }
            bgdot_rgb = b0 * g[0] + b1 * g[1] + b2 * g[2];
        }
        // phase B reads the block's upstream gradients by (row, column)
        sm.g[warp][0][(lane >> 3) * SG_ROW + (lane & 7)] = make_float4(g[0], g[1], g[2], g[3]);
        if (FUSED) sm.g[warp][1][(lane >> 3) * SG_ROW + (lane & 7)] = make_float4(g[4], g[5], 0.f, 0.f);
    }
    const float kx = 0.5f * cc.W, ky = 0.5f * cc.H;
    const BwdLaneAddr la = bwd_lane_addr<!POSE_ONLY>(sm, warp, lane);
    // which upstream planes are non-zero anywhere in this warp's block (uniform per warp):
    // 0 = colour only (pose tracking), 1 = + depth (mapping), 2 = + silhouette / depth^2
    int level = __any_sync(FULL, g[3] != 0.f) ? 1 : 0;
    if (FUSED && __any_sync(FULL, g[4] != 0.f || g[5] != 0.f)) level = 2;

    BwdPixel ps;
    ps.T = T_final;
    ps.ag_rgb = ps.ag_dep = 0.f;

    const float4 *src = sorted_rec + (size_t)start * REC_F4;
    auto batch_cnt = [&](int k) { return min(BWD_BATCH, maxlast - k * BWD_BATCH); };

    // one staged batch for this warp: entries that can matter = mask hit AND not deeper than the warp's
    // deepest contributor
    auto process = [&](int k, const float4 *sb) __attribute__((always_inline)) {
        const int limit = min(batch_cnt(k), warp_last - k * BWD_BATCH);
        const int nrel = limit > 0 ? compact_entries(sb, limit, warp_bit, lane, sm.list[warp]) : 0;
        if (nrel > 0) {
            if (level == 0)
                bwd_batch<FUSED, 0, POSE_ONLY>(sm, sb, warp, lane, nrel, (last - k * BWD_BATCH) * REC_BYTES, pxf, pyf, bx, by, kx, ky, g,
                                    T_final, bgdot_rgb, bgdot_dep, ps, grad_acc, la);
            else if (!FUSED || level == 1)
                bwd_batch<FUSED, 1, POSE_ONLY>(sm, sb, warp, lane, nrel, (last - k * BWD_BATCH) * REC_BYTES, pxf, pyf, bx, by, kx, ky, g,
                                    T_final, bgdot_rgb, bgdot_dep, ps, grad_acc, la);
            else
                bwd_batch<FUSED, 2, POSE_ONLY>(sm, sb, warp, lane, nrel, (last - k * BWD_BATCH) * REC_BYTES, pxf, pyf, bx, by, kx, ky, g,
                                    T_final, bgdot_rgb, bgdot_dep, ps, grad_acc, la);
        }
    };

    if (!use_tma) {
        // FSGS_FLAG_NO_TMA: cooperative loads, two block barriers per batch
        for (int it = 0; it < nb; ++it) {
            const int k = nb - 1 - it;
            __syncthreads();
            stage_plain(sm.rec[0], src + (size_t)k * BWD_BATCH * REC_F4, batch_cnt(k));
            __syncthreads();
            process(k, sm.rec[0]);
        }
        return;
    }

    // Bulk-TMA ring without block barriers.  The warps of a tile have very different amounts of work per batch
    // (their blocks see different splats, and a block whose pixels saturate early skips the deep batches), so a
    // __syncthreads per batch left warps parked at the barrier (ncu: top stall reason).  Instead every warp
    // counts itself out of a stage when it is done with the batch in it; the LAST warp out re-arms that stage's
    // mbarrier and issues the bulk copy of the batch BWD_STAGES further on.  A warp only ever waits for data.
    if (threadIdx.x == 0)
        for (int it = 0; it < min(BWD_STAGES, nb); ++it) {
            const int k = nb - 1 - it;
            stage_issue_tma(sm.rec[it], src + (size_t)k * BWD_BATCH * REC_F4, batch_cnt(k), &sm.full[it]);
        }
    __syncwarp();          // this warp's rows of sm.g are written
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nb; ++it) {
        const int k = nb - 1 - it;
        mbar_wait(&sm.full[stage], phase, err);
        process(k, sm.rec[stage]);
        __syncwarp();      // all lanes are done reading the stage
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&sm.done_cnt[stage], 1u) == (unsigned int)(NWARP - 1)) {
                sm.done_cnt[stage] = 0u;          // next arrivals come only after the copy below has landed
                __threadfence_block();
                const int itn = it + BWD_STAGES;
                if (itn < nb) {
                    const int kn = nb - 1 - itn;
                    stage_issue_tma(sm.rec[stage], src + (size_t)kn * BWD_BATCH * REC_F4, batch_cnt(kn), &sm.full[stage]);
                }
            }
        }
        if (++stage == BWD_STAGES) { stage = 0; phase ^= 1u; }
    }
}

}  // namespace fsgs
