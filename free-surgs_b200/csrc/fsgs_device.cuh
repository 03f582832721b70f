// fsgs_device.cuh -- sm_100a device helpers: mbarrier + 1-D bulk TMA (cp.async.bulk), warp
// reduce-scatter, the pinned-order alpha evaluation shared by the forward and backward
// compositors, and the buffer layouts shared between kernels and host code.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fsgs_math.cuh"

namespace fsgs {

constexpr int CTA = 256;              // threads per compositor / sort CTA (one 16x16 tile)
constexpr int BATCH = 256;            // tile-list entries staged per bulk copy
constexpr int REC_F4 = 3;             // float4 per splat record (48 B)
constexpr int ACC_F = 12;             // floats per Gaussian in the gradient accumulator (48 B)
constexpr uint32_t FULL = 0xffffffffu;

// ---- splat record (48 B, 16-B aligned; one per Gaussian, and one per sorted tile instance) ----
//   q0 = (x, y, conic.x, conic.y)   q1 = (conic.z, opacity, r, g)   q2 = (b, depth, radius, tiles)
// radius / tiles are int bit patterns.  The sorted per-instance copy carries the conic pre-scaled for a
// base-2 exponent, the 8x4-block reach mask in q2.z and the Gaussian id in q2.w.
//
// ---- gradient accumulator row (48 B) written by the backward compositor ----
//   [0..3]  = d(mean2D.x), d(mean2D.y), d(conic.x), d(conic.y)
//   [4..7]  = d(conic.z), d(opacity), d(r), d(g)
//   [8..11] = d(b), d(depth), d(mean2D.x | RGB planes only), d(mean2D.y | RGB planes only)

struct ImgLayout {
    size_t final_T, n_contrib, tile_count, tile_offset, cursor, counters, total;
    int tiles;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ inline ImgLayout img_layout(int W, int H) {
    ImgLayout L;
    const size_t HW = (size_t)W * H;
    L.tiles = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    size_t o = 0;
    L.final_T = o; o = align_up(o + HW * 4, 256);
    L.n_contrib = o; o = align_up(o + HW * 4, 256);
    L.tile_count = o; o = align_up(o + (size_t)L.tiles * 4, 256);
    L.tile_offset = o; o = align_up(o + ((size_t)L.tiles + 1) * 4, 256);
    L.cursor = o; o = align_up(o + (size_t)L.tiles * 4, 256);
    L.counters = o; o = align_up(o + 64, 256);
    L.total = o;
    return L;
}

// counters[] (uint64): 0 = instances after culling (R), 1 = instances of the reference's
// rectangles, 2 = longest tile list, 3 = device error / watchdog flag
enum { CNT_R = 0, CNT_RECT = 1, CNT_MAXLIST = 2, CNT_ERR = 3 };

struct GeomLayout {
    size_t records, clamped, total;
};
__host__ inline GeomLayout geom_layout(int P) {
    GeomLayout L;
    const size_t n = P > 0 ? P : 1;
    size_t o = 0;
    L.records = o; o = align_up(o + n * 48, 256);
    L.clamped = o; o = align_up(o + n, 256);
    L.total = o;
    return L;
}

struct BinLayout {
    size_t keys, records, bins, total;
};
// bins (optional): tiles x bin_cap keys -- the counting pass drops every instance's (depth, id) key straight into its
// tile's fixed-stride bin (slot = the value its counting atomic returns), so the separate scatter pass disappears;
// bin_cap comes from the previous frame's longest tile list (forward_tail; the scatter path is the fallback).
__host__ inline BinLayout bin_layout(int64_t R, int tiles = 0, unsigned int bin_cap = 0) {
    BinLayout L;
    const size_t n = R > 0 ? (size_t)R : 1;
    size_t o = 0;
    // records first: the backward only needs them, so their offset must not depend on the capacity the
    // forward allocated (which may exceed the instance count, see forward_tail)
    L.records = o; o = align_up(o + n * 48, 256);
    L.keys = o; o = align_up(o + n * 8, 256);
    L.bins = o; o = align_up(o + (size_t)tiles * bin_cap * 8, 256);
    L.total = o;
    return L;
}

#if defined(__CUDACC__)

// ---- mbarrier / bulk-copy PTX (sm_90+; SASS: SYNCS.*, UBLKCP) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the initialised barrier visible to the async proxy before the first bulk copy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a stuck barrier sets the error flag instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, unsigned long long *err) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            if (err) atomicExch(err, 1ull);
            break;
        }
    }
}
// global -> shared 1-D bulk copy through the TMA engine; bytes % 16 == 0, both sides 16-B aligned
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared -> global 1-D bulk copy (TMA store); completion tracked with bulk async-groups
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// make this thread's generic-proxy shared-memory writes visible to the async proxy (before a TMA store)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Stage the CTA's slice of a [P, row_floats] float array into shared memory: one bulk TMA copy for the
// largest 16-byte-multiple prefix, plain loads for the (at most 3-row) remainder.  `count` rows starting at
// row `base`.  Requires src 16-B aligned.  In two halves, so that a kernel can put its own (independent) global
// loads in flight between issuing the bulk copy and waiting for it instead of serialising two DRAM latencies:
//   stage_rows_issue: thread 0 arms the barrier and issues the copy; every thread loads its share of the remainder.
//   stage_rows_wait : block barrier (mbarrier initialised + remainder visible), then the wait for the bulk copy.
// Both must be called by all threads of the CTA, outside divergent code.
template <int ROW_FLOATS>
__device__ __forceinline__ void stage_rows_issue(float *s_dst, const float *__restrict__ src, int base, int count,
                                                 uint64_t *bar) {
    const int rows_tma = count & ~3;
    const float *g = src + (size_t)base * ROW_FLOATS;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        if (rows_tma > 0) {
            const uint32_t bytes = (uint32_t)rows_tma * ROW_FLOATS * 4u;
            mbar_arrive_expect_tx(bar, bytes);
            tma_load_1d(s_dst, g, bytes, bar);
        }
    }
    for (int q = rows_tma * ROW_FLOATS + threadIdx.x; q < count * ROW_FLOATS; q += blockDim.x) s_dst[q] = __ldg(g + q);
}
// Several arrays staged behind ONE mbarrier (the per-Gaussian kernels stage every SoA attribute of their CTA's
// Gaussians this way: xyz | scaling | f_dc | opacity | rotation | f_rest = six contiguous slices, six bulk copies,
// one wait).  Call stage_multi_begin once (thread 0 initialises the barrier), stage_multi_add per array (thread 0
// issues the 16-byte-multiple prefix, all threads load the remainder), stage_multi_wait once.
struct StageMulti {
    uint64_t *bar;
    uint32_t bytes;     // thread 0: bytes expected so far
};
__device__ __forceinline__ StageMulti stage_multi_begin(uint64_t *bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    return StageMulti{bar, 0u};
}
template <int ROW_FLOATS>
__device__ __forceinline__ void stage_multi_add(StageMulti &sm, float *s_dst, const float *__restrict__ src, int base,
                                                int count) {
    // rows whose byte length is a multiple of 16 from a 16-byte aligned start: count rounded down to a multiple of 4
    const int rows_tma = count & ~3;
    const float *g = src + (size_t)base * ROW_FLOATS;
    if (threadIdx.x == 0 && rows_tma > 0) {
        const uint32_t bytes = (uint32_t)rows_tma * ROW_FLOATS * 4u;
        sm.bytes += bytes;
        // (expect_tx may be raised in several steps before the single arrive: use the no-arrive form)
        asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(sm.bar)), "r"(bytes) : "memory");
        tma_load_1d(s_dst, g, bytes, sm.bar);
    }
    for (int q = rows_tma * ROW_FLOATS + threadIdx.x; q < count * ROW_FLOATS; q += blockDim.x) s_dst[q] = __ldg(g + q);
}
__device__ __forceinline__ void stage_multi_wait(StageMulti &sm, unsigned long long *err) {
    if (threadIdx.x == 0) {
        uint64_t st;
        asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(smem_u32(sm.bar)) : "memory");
        (void)st;
    }
    __syncthreads();                 // barrier initialised + arrived, remainders visible
    mbar_wait(sm.bar, 0u, err);
}

__device__ __forceinline__ void stage_rows_wait(int count, uint64_t *bar, unsigned long long *err) {
    __syncthreads();
    if ((count & ~3) > 0) mbar_wait(bar, 0u, err);
}

// ---- warp reduce-scatter of 16 values ----------------------------------------------------------
// After the call lane L holds in v[0] the warp-wide sum of input index (L >> 1) & 15.
// 16 shuffles instead of 16 * 5.
__device__ __forceinline__ void warp_reduce_scatter16(float (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, mask = 16; half >= 1; half >>= 1, mask >>= 1) {
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = upper ? v[i + half] : v[i];
            const float send = upper ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(FULL, send, mask);
        }
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
}

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// Shared-memory accesses by explicit 32-bit shared-window address + compile-time offset.  The backward compositor keeps
// a handful of per-lane base addresses in registers for the whole kernel and addresses everything relative to them:
// with C++ pointers into the dynamic shared array the compiler re-derived those bases from %tid / the shared-window
// base inside every phase-B round (ncu source page: ~40 of 220 instructions per round were S2R / IMAD / LEA / LOP3
// address re-materialisation under the 64-register cap).
template <int OFF>
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int lds_u16(uint32_t a) {
    unsigned int v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// make a kernel-constant register value opaque, so that the compiler keeps it instead of re-deriving it where used
__device__ __forceinline__ uint32_t keep_reg(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}

// 16-byte vector reduction into global memory (sm_90+): one L2 operation, no value returned.
// (atomicAdd(float4*) compiles to ATOMG with a discarded result; this is the plain RED form.)
__device__ __forceinline__ void red_add_v4(float4 *addr, float4 v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

#endif  // __CUDACC__

}  // namespace fsgs
