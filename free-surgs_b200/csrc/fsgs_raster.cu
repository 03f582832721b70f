// fsgs_raster.cu -- host side of the C ABI declared in include/fsgs_raster.h.
// Builds into libfsgs_raster.so with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// No torch, no Python: raw pointers, a stream, allocation callbacks.
#include "../../include/fsgs_raster.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "fsgs_kernels_bwd.cuh"
#include "fsgs_kernels_loss.cuh"
#include "fsgs_kernels_composite.cuh"
#include "fsgs_kernels_refstyle.cuh"
#include "fsgs_kernels_pre.cuh"

using namespace fsgs;

namespace {

thread_local char g_cuda_msg[256] = {0};

#define FSGS_CUDA(call)                                                                             \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            snprintf(g_cuda_msg, sizeof(g_cuda_msg), "%s failed: %s (%s:%d)", #call,                \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                                  \
            return FSGS_E_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

inline int check_launch(const fsgs_settings *st, cudaStream_t s, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && st->debug) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        snprintf(g_cuda_msg, sizeof(g_cuda_msg), "kernel %s failed: %s", what, cudaGetErrorString(e));
        return FSGS_E_CUDA;
    }
    return FSGS_OK;
}
#define FSGS_LAUNCH_OK(what)                                  \
    do {                                                      \
        int r__ = check_launch(st, stream, what);             \
        if (r__ != FSGS_OK) return r__;                       \
    } while (0)

int make_cam(const fsgs_settings *st, CamConst &cc) {
    if (!st || st->image_width <= 0 || st->image_height <= 0 || st->tanfovx <= 0.f || st->tanfovy <= 0.f)
        return FSGS_E_INVALID;
    if (st->sh_degree < 0 || st->sh_degree > 3) return FSGS_E_INVALID;
    cc.W = st->image_width; cc.H = st->image_height;
    cc.gx = (cc.W + TILE - 1) / TILE; cc.gy = (cc.H + TILE - 1) / TILE;
    cc.fx = cc.W / (2.0f * st->tanfovx); cc.fy = cc.H / (2.0f * st->tanfovy);
    cc.limx = 1.3f * st->tanfovx; cc.limy = 1.3f * st->tanfovy;
    cc.mod = st->scale_modifier;
    cc.sh_deg = st->sh_degree; cc.n_coeffs = st->n_coeffs;
    return FSGS_OK;
}

int check_arch() {
    static signed char cached[64] = {0};   // per device: 0 unknown, 1 ok, -1 bad
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return FSGS_E_CUDA;
    if (dev >= 0 && dev < 64 && cached[dev] != 0) return cached[dev] == 1 ? FSGS_OK : FSGS_E_ARCH;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return FSGS_E_CUDA;
    const signed char v = (major == 10) ? 1 : -1;
    if (dev >= 0 && dev < 64) cached[dev] = v;
    return v == 1 ? FSGS_OK : FSGS_E_ARCH;
}

inline int blocks(int n) { return (n + CTA - 1) / CTA; }

// ---- optional per-kernel timing (CUDA events on the launching stream) ---------------------------
// Off by default.  bench.py switches it on for a separate profiling pass (never for the timed
// steps) to obtain the per-launch duration of each kernel for the roofline line.
enum KernelId { K_PRE_API = 0, K_PRE_FUSED, K_SCAN, K_SCATTER, K_SORT, K_COMP_FWD, K_COMP_BWD, K_PRE_API_BWD,
                K_PRE_FUSED_BWD, K_MARK_VISIBLE, K_POSE_FWD, K_POSE_BWD, K_SH_EXPAND, K_PRE_POSE_BWD, K_LOSS_FWD, K_LOSS_BWD, K_PEARSON_FWD, K_PEARSON_BWD,
                K_LOCAL_PEARSON_FWD, K_LOCAL_PEARSON_BWD, K_EXCHANGE, K_FREEZE, K_PRE_FROZEN, K_COUNT };
struct Profiler {
    bool on = false;
    static constexpr int MAXREC = 8192;
    cudaEvent_t ev[MAXREC][2];
    int kid[MAXREC];
    int created = 0, used = 0;
};
Profiler g_prof;

inline void prof_begin(int id, cudaStream_t s) {
    if (!g_prof.on || g_prof.used >= Profiler::MAXREC) return;
    if (g_prof.used >= g_prof.created) {
        cudaEventCreate(&g_prof.ev[g_prof.created][0]);
        cudaEventCreate(&g_prof.ev[g_prof.created][1]);
        g_prof.created++;
    }
    g_prof.kid[g_prof.used] = id;
    cudaEventRecord(g_prof.ev[g_prof.used][0], s);
}
inline void prof_end(int id, cudaStream_t s) {
    if (!g_prof.on || g_prof.used >= Profiler::MAXREC) return;
    (void)id;
    cudaEventRecord(g_prof.ev[g_prof.used][1], s);
    g_prof.used++;
}

struct Buffers {
    char *geom, *img, *bin;
    GeomLayout gl;
    ImgLayout il;
    BinLayout bl;
    unsigned long long capacity = 0;   // instances the binning buffer was allocated for BEFORE the projection kernel (0: not yet)
    unsigned int bin_cap = 0;          // keys per tile bin in it (0: no bins, the scatter path)
};

// ---- instance-count read-back without idling the GPU ----------------------------------------------
// The binning buffer is sized by R (tile instances), which only the device knows after k_tile_scan.
// Blocking on that number leaves the GPU idle while the host wakes up, runs the allocation callback and
// launches the rest of the forward.  Instead, when a previous forward on this device left a hint,
// the tail (scatter -> sort -> composite) is launched OPTIMISTICALLY into a buffer sized from the hint
// before the host waits for R; every tail kernel compares the device-side R with that capacity and
// returns at once if it does not fit, in which case the host allocates the exact size and launches
// the tail again.  Results never depend on the hint -- it only decides whether the GPU waits.
struct HostMailbox {
    unsigned long long *pinned = nullptr;   // 4 counters, page-locked so the D2H copy is truly asynchronous
    cudaEvent_t ev[64] = {};                // per device: "counters have landed"
    bool have_ev[64] = {};
};
thread_local HostMailbox g_mail;
unsigned long long g_hint[64] = {};         // last R per device (benign race: performance hint only)
unsigned long long g_hint_maxlist[64] = {}; // longest tile list of the last frame per device (likewise; sizes k_tile_sort's window
                                            // and the per-tile key bins) ...
unsigned long long g_hint_maxlist2[64] = {};// ... and of the frame before it: loops that alternate between views (mapping over
                                            // key frames, the viewer thread) size for the larger of the two
inline unsigned long long hint_maxlist(int dev) {
    return g_hint_maxlist[dev] > g_hint_maxlist2[dev] ? g_hint_maxlist[dev] : g_hint_maxlist2[dev];
}
// Device watchdog: one sticky 64-bit word per device (allocated once, never freed).  A bounded device-side wait that
// gives up (mbar_wait) sets it; every forward reads it back together with the instance count -- the read-back it
// performs anyway -- and fails with FSGS_E_WATCHDOG, so a stuck barrier is reported by the NEXT forward on that
// device (like an asynchronous CUDA error), in release mode too.  fsgs_watchdog_flag() reads it on demand.
unsigned long long *g_sticky[64] = {};
unsigned long long g_fixed_cap[64] = {};    // FSGS_FLAG_FIXED_CAPACITY: instance capacity per device (fsgs_set_instance_capacity)
unsigned int g_fixed_bin_cap[64] = {};      // ... and the per-tile bin capacity the last fixed-capacity forward on the device ran with
                                            // (0 = scatter path); its backward compares the longest list with it (device guard)

// Binning buffer ahead of the projection kernel.  When the device left hints (instance count and longest tile list of
// the previous frame) -- or the capacity is fixed (graph capture) -- the buffer is allocated before the projection
// kernel runs, with one fixed-stride bin of keys per tile, and that kernel's counting pass writes every instance's key
// into its bin (BinSink): no scatter pass.  A frame that outgrows the bins (or the capacity) is caught by the device
// guards of the tail kernels and relaunched through the scatter path with exact sizes.
int prepare_binning(const fsgs_settings *st, int tiles, fsgs_alloc_fn binning_alloc, void *binning_user, Buffers &B) {
    int dev = 0;
    FSGS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return FSGS_E_INVALID;
    const bool fixed = (st->flags & FSGS_FLAG_FIXED_CAPACITY) != 0;
    unsigned long long capacity = 0;
    if (fixed) {
        capacity = g_fixed_cap[dev];
        if (capacity == 0 || st->debug) return FSGS_E_INVALID;
    } else if (g_hint[dev] > 0 && !(st->flags & FSGS_FLAG_NO_OPTIMISTIC)) {
        capacity = g_hint[dev] + g_hint[dev] / 8 + 4096;
    }
    if (capacity == 0) return FSGS_OK;               // first frame on this device: the tail waits for the counts
    unsigned int bin_cap = 0;
    const unsigned long long ml = hint_maxlist(dev);
    if (ml > 0 && ml < (1ull << 20) && !(st->flags & FSGS_FLAG_NO_BINS))
        bin_cap = (unsigned int)((ml + ml / 2 + 64 + 63) / 64 * 64);
    // (a pathological frame -- one enormous list -- would make tiles x bin_cap explode: fall back to the scatter path)
    if ((unsigned long long)tiles * bin_cap > 4 * capacity + (1u << 20)) bin_cap = 0;
    B.bl = bin_layout((int64_t)capacity, tiles, bin_cap);
    B.bin = static_cast<char *>(binning_alloc(binning_user, B.bl.total));
    if (!B.bin) return FSGS_E_ALLOC;
    B.capacity = capacity;
    B.bin_cap = bin_cap;
    if (fixed) g_fixed_bin_cap[dev] = bin_cap;
    return FSGS_OK;
}
// Backward of a fixed-capacity (captured) forward: the bin capacity that forward ran with, for the device guard.
// (Outside capture the host has already checked the counts before a backward can be issued: 0 = no guard.)
inline unsigned int fixed_bin_cap(const fsgs_settings *st) {
    if (!(st->flags & FSGS_FLAG_FIXED_CAPACITY)) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    return g_fixed_bin_cap[dev];
}
inline BinSink bin_sink(const Buffers &B) {
    return BinSink{B.bin_cap ? reinterpret_cast<unsigned long long *>(B.bin + B.bl.bins) : nullptr, B.bin_cap};
}

int mailbox(int dev, unsigned long long **pinned, cudaEvent_t *ev) {
    if (!g_mail.pinned) FSGS_CUDA(cudaHostAlloc((void **)&g_mail.pinned, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    if (!g_mail.have_ev[dev]) {
        FSGS_CUDA(cudaEventCreateWithFlags(&g_mail.ev[dev], cudaEventDisableTiming));
        g_mail.have_ev[dev] = true;
    }
    *pinned = g_mail.pinned;
    *ev = g_mail.ev[dev];
    return FSGS_OK;
}

// Shared tail of both forward flavours: scan -> (R read-back) -> scatter -> sort -> composite.
template <bool FUSED>
int forward_tail(const fsgs_settings *st, const CamConst &cc, int P, const float *bg, Buffers &B,
                 fsgs_alloc_fn binning_alloc, void *binning_user, float *out_planes, float *out_depth,
                 int64_t *num_rendered_host, int64_t *num_rect_host, const RenderExtras &ex, cudaStream_t stream) {
    const int tiles = B.il.tiles;
    unsigned int *tile_count = reinterpret_cast<unsigned int *>(B.img + B.il.tile_count);
    unsigned int *tile_offset = reinterpret_cast<unsigned int *>(B.img + B.il.tile_offset);
    unsigned int *cursor = reinterpret_cast<unsigned int *>(B.img + B.il.cursor);
    unsigned long long *counters = reinterpret_cast<unsigned long long *>(B.img + B.il.counters);
    float4 *records = reinterpret_cast<float4 *>(B.geom + B.gl.records);
    int dev = 0;
    FSGS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return FSGS_E_INVALID;
    const bool fixed = (st->flags & FSGS_FLAG_FIXED_CAPACITY) != 0;
    unsigned long long *h_cnt = nullptr;
    cudaEvent_t landed = nullptr;
    int rc = FSGS_OK;
    if (!fixed && (rc = mailbox(dev, &h_cnt, &landed)) != FSGS_OK) return rc;

    prof_begin(K_SCAN, stream);
    k_tile_scan<<<1, 1024, 0, stream>>>(tiles, tile_count, tile_offset, cursor, counters);
    prof_end(K_SCAN, stream);
    FSGS_LAUNCH_OK("k_tile_scan");
    unsigned long long *sticky = g_sticky[dev];
    if (!fixed) {
        FSGS_CUDA(cudaMemcpyAsync(h_cnt, counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        FSGS_CUDA(cudaMemcpyAsync(h_cnt + 4, sticky, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        FSGS_CUDA(cudaEventRecord(landed, stream));
    }

    // launches (scatter ->) sort -> composite into a buffer of `capacity` instances; bin_cap > 0: the keys already sit
    // in the per-tile bins the counting pass filled
    auto launch_tail = [&](unsigned long long capacity, bool have_instances, unsigned int bin_cap) -> int {
        unsigned long long *keys = reinterpret_cast<unsigned long long *>(B.bin + B.bl.keys);
        unsigned long long *bins = bin_cap ? reinterpret_cast<unsigned long long *>(B.bin + B.bl.bins) : nullptr;
        float4 *sorted_rec = reinterpret_cast<float4 *>(B.bin + B.bl.records);
        if (have_instances) {
            if (!bins) {
                prof_begin(K_SCATTER, stream);
                k_scatter<<<blocks(P), CTA, 0, stream>>>(cc, P, records, tile_offset, cursor, keys, (unsigned)st->flags,
                                                         counters, capacity);
                prof_end(K_SCATTER, stream);
                FSGS_LAUNCH_OK("k_scatter");
            }
            prof_begin(K_SORT, stream);
            // shared-memory window: the small one (4 resident CTAs) when the previous frame's longest list fits its
            // bucket path with 25 % headroom; results do not depend on the choice (fsgs_kernels_pre.cuh)
            const unsigned long long ml = hint_maxlist(dev);
            const int win = (ml > 0 && ml + ml / 4 <= (unsigned long long)(SORT_SMEM_KEYS_SMALL / 2) &&
                             !(st->flags & FSGS_FLAG_SORT_WINDOW_LARGE))
                                ? SORT_SMEM_KEYS_SMALL : SORT_SMEM_KEYS;
            k_tile_sort<<<tiles, CTA, (size_t)win * sizeof(unsigned long long), stream>>>(
                cc.gx, tile_offset, keys, records, sorted_rec, (unsigned)st->flags, counters, capacity, win, bins, bin_cap);
            prof_end(K_SORT, stream);
            FSGS_LAUNCH_OK("k_tile_sort");
        }
        prof_begin(K_COMP_FWD, stream);
        if (!FUSED && (st->flags & FSGS_FLAG_UPSTREAM_STYLE)) {
            // baseline, not the product (fsgs_kernels_refstyle.cuh): the published rasteriser's kernel structure
            k_composite_fwd_ref<<<tiles, CTA, 0, stream>>>(
                cc, tile_offset, sorted_rec, bg, out_planes, out_depth, reinterpret_cast<float *>(B.img + B.il.final_T),
                reinterpret_cast<unsigned int *>(B.img + B.il.n_contrib), counters, capacity, bin_cap);
            prof_end(K_COMP_FWD, stream);
            FSGS_LAUNCH_OK("k_composite_fwd_ref");
            return FSGS_OK;
        }
        k_composite_fwd<FUSED><<<tiles, CTA, 0, stream>>>(
            cc, tile_offset, sorted_rec, bg, out_planes, out_depth, reinterpret_cast<float *>(B.img + B.il.final_T),
            reinterpret_cast<unsigned int *>(B.img + B.il.n_contrib), (unsigned)st->flags, sticky, counters,
            capacity, ex, bin_cap);
        prof_end(K_COMP_FWD, stream);
        FSGS_LAUNCH_OK("k_composite_fwd");
        return FSGS_OK;
    };

    if (fixed) {
        // Stream-capture mode (CUDA graphs): nothing here may touch the host.  The binning buffer is sized for the
        // capacity the caller declared; if the frame has more instances, every tail kernel (and the backward
        // compositor) returns at once on the device-side guard and the caller finds counters[0] > capacity.
        const unsigned long long cap = B.capacity;          // allocated by prepare_binning
        if (cap == 0 || !B.bin) return FSGS_E_INVALID;
        if ((rc = launch_tail(cap, true, B.bin_cap)) != FSGS_OK) return rc;
        if (num_rendered_host) *num_rendered_host = (int64_t)cap;     // an upper bound; the backward only needs > 0
        if (num_rect_host) *num_rect_host = 0;
        return FSGS_OK;
    }
    unsigned long long capacity = 0;
    bool launched = false;
    if (B.capacity > 0 && B.bin) {                          // prepare_binning found hints: optimistic launch
        capacity = B.capacity;
        if ((rc = launch_tail(capacity, true, B.bin_cap)) != FSGS_OK) return rc;
        launched = true;
    }
    FSGS_CUDA(cudaEventSynchronize(landed));
    if (h_cnt[4]) {
        // a device-side wait of an EARLIER launch on this device gave up: its results are not to be trusted
        FSGS_CUDA(cudaMemsetAsync(sticky, 0, sizeof(unsigned long long), stream));
        return FSGS_E_WATCHDOG;
    }
    const int64_t R = (int64_t)h_cnt[CNT_R];
    if (num_rendered_host) *num_rendered_host = R;
    if (num_rect_host) *num_rect_host = (int64_t)h_cnt[CNT_RECT];
    if (R >= (int64_t)1 << 32) return FSGS_E_INVALID;
    g_hint[dev] = (unsigned long long)R;
    g_hint_maxlist2[dev] = g_hint_maxlist[dev];
    g_hint_maxlist[dev] = h_cnt[CNT_MAXLIST];

    if (!launched || (unsigned long long)R > capacity || (B.bin_cap && h_cnt[CNT_MAXLIST] > B.bin_cap)) {
        // no hint, or the frame outgrew the buffer / a tile outgrew its bin (the optimistic kernels left on their
        // device guards): exact size, scatter path
        B.bl = bin_layout(R);
        B.bin = static_cast<char *>(binning_alloc(binning_user, B.bl.total));
        if (!B.bin) return FSGS_E_ALLOC;
        B.bin_cap = 0;
        if ((rc = launch_tail((unsigned long long)(R > 0 ? R : 1), R > 0, 0)) != FSGS_OK) return rc;
    }
    if (st->debug) {
        FSGS_CUDA(cudaMemcpyAsync(h_cnt + 4, sticky, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        FSGS_CUDA(cudaStreamSynchronize(stream));
        if (h_cnt[4]) {
            FSGS_CUDA(cudaMemsetAsync(sticky, 0, sizeof(unsigned long long), stream));
            return FSGS_E_WATCHDOG;
        }
    }
    return FSGS_OK;
}

int alloc_fixed(int P, const CamConst &cc, fsgs_alloc_fn geom_alloc, void *geom_user, fsgs_alloc_fn img_alloc,
                void *img_user, Buffers &B, cudaStream_t stream) {
    B.gl = geom_layout(P);
    B.il = img_layout(cc.W, cc.H);
    B.geom = static_cast<char *>(geom_alloc(geom_user, B.gl.total));
    B.img = static_cast<char *>(img_alloc(img_user, B.il.total));
    B.bin = nullptr;
    if (!B.geom || !B.img) return FSGS_E_ALLOC;
    // tile_count .. counters are contiguous: one memset
    FSGS_CUDA(cudaMemsetAsync(B.img + B.il.tile_count, 0, B.il.total - B.il.tile_count, stream));
    return FSGS_OK;
}

int one_time_setup() {
    // function attributes are per device: remember which devices of this process have been set up
    static bool done[64] = {false};
    int dev = 0;
    FSGS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
        FSGS_CUDA(cudaFuncSetAttribute(k_tile_sort, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(SORT_SMEM_KEYS * sizeof(unsigned long long))));
        FSGS_CUDA(cudaFuncSetAttribute(k_composite_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(BwdSmem)));
        FSGS_CUDA(cudaFuncSetAttribute(k_composite_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(BwdSmem)));
        FSGS_CUDA((cudaFuncSetAttribute(k_composite_bwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(BwdSmem))));
        FSGS_CUDA((cudaFuncSetAttribute(k_composite_bwd<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        (int)cudaSharedmemCarveoutMaxShared)));
        // 4 resident CTAs x ~48 KB: ask for the large shared-memory carve-out
        FSGS_CUDA(cudaFuncSetAttribute(k_composite_bwd<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
        FSGS_CUDA(cudaFuncSetAttribute(k_composite_bwd<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
        if (dev >= 0 && dev < 64 && !g_sticky[dev]) {
            FSGS_CUDA(cudaMalloc((void **)&g_sticky[dev], 256));
            FSGS_CUDA(cudaMemset(g_sticky[dev], 0, 256));
        }
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return FSGS_OK;
}

unsigned long long *sticky_flag() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    return g_sticky[dev];
}

// fills the whole image with the background (P == 0 or nothing visible is handled by the kernels,
// this is only for P == 0 where no kernel runs)
__global__ void k_fill_bg(int HW, int planes, const float *__restrict__ bg, float *__restrict__ out,
                          float *__restrict__ depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    for (int c = 0; c < planes; ++c) out[(size_t)c * HW + i] = __ldg(bg + (c % 3));
    if (depth) depth[i] = 0.f;
}

// the derived maps of an empty scene: every plane equals the background
__global__ void k_fill_extras(int HW, const float *__restrict__ bg, RenderExtras ex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const float dep = __ldg(bg), sil = __ldg(bg + 1), dsq = __ldg(bg + 2);
    const float unc = __fsub_rn(dsq, __fmul_rn(dep, dep));
    if (ex.uncertainty) ex.uncertainty[i] = unc;
    if (ex.presence_mask) ex.presence_mask[i] = sil > 0.3f ? 1 : 0;
    if (ex.nan_mask) ex.nan_mask[i] = (dep == dep && unc == unc) ? 1 : 0;
}

}  // namespace

extern "C" {

int fsgs_abi_version(void) { return FSGS_ABI_VERSION; }

int fsgs_profile_enable(int on) {
    g_prof.on = on != 0;
    g_prof.used = 0;
    return FSGS_OK;
}

int fsgs_profile_collect(double *ms_sum, int64_t *count, int n) {
    if (!ms_sum || !count || n < K_COUNT) return FSGS_E_INVALID;
    for (int k = 0; k < n; ++k) { ms_sum[k] = 0.0; count[k] = 0; }
    FSGS_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < g_prof.used; ++i) {
        float ms = 0.f;
        FSGS_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[i][0], g_prof.ev[i][1]));
        ms_sum[g_prof.kid[i]] += ms;
        count[g_prof.kid[i]] += 1;
    }
    g_prof.used = 0;
    return FSGS_OK;
}

const char *fsgs_error_string(int code) {
    switch (code) {
        case FSGS_OK: return "ok";
        case FSGS_E_INVALID: return "invalid argument";
        case FSGS_E_CUDA: return g_cuda_msg[0] ? g_cuda_msg : "CUDA error";
        case FSGS_E_ALLOC: return "allocation callback returned NULL";
        case FSGS_E_ARCH: return "device is not compute capability 10.x (this library is sm_100a only)";
        case FSGS_E_WATCHDOG: return "device-side wait exceeded its spin budget (mbarrier never completed)";
        default: return "unknown error";
    }
}

const char *fsgs_kernel_names(void) {
    return "k_preprocess_api,k_preprocess_fused,k_tile_scan,k_scatter,k_tile_sort,k_composite_fwd,"
           "k_composite_bwd,k_preprocess_api_bwd,k_preprocess_fused_bwd,k_mark_visible,k_pose_forward,k_pose_backward,"
           "k_sh_grad_expand,k_preprocess_pose_bwd,k_rgb_loss_fwd,k_rgb_loss_bwd,k_pearson_sums,k_pearson_bwd,"
           "k_local_pearson_sums,k_local_pearson_bwd,k_exchange_rows,k_freeze_model,k_preprocess_frozen";
}

size_t fsgs_geom_bytes(int32_t P) { return geom_layout(P).total; }
size_t fsgs_img_bytes(int32_t W, int32_t H) { return img_layout(W, H).total; }
size_t fsgs_binning_bytes(int64_t R) { return bin_layout(R).total; }
size_t fsgs_grad_scratch_bytes(int32_t P) { return align_up((size_t)(P > 0 ? P : 1) * ACC_F * 4, 256); }
size_t fsgs_geom_record_offset(int32_t P) { return geom_layout(P).records; }
void fsgs_img_offsets(int32_t W, int32_t H, size_t *out6) {
    const ImgLayout L = img_layout(W, H);
    out6[0] = L.final_T; out6[1] = L.n_contrib; out6[2] = L.tile_count; out6[3] = L.tile_offset; out6[4] = L.cursor;
    out6[5] = L.counters;
}
void fsgs_binning_offsets(int64_t R, size_t *out2) {
    const BinLayout L = bin_layout(R);
    out2[0] = L.keys; out2[1] = L.records;
}

int fsgs_rasterize_forward(const fsgs_settings *st, int32_t P, const float *bg, const float *means3D,
                           const float *colors_precomp, const float *shs, const float *opacities,
                           const float *scales, const float *rotations, const float *cov3D_precomp,
                           const float *viewmatrix, const float *projmatrix, const float *campos,
                           fsgs_alloc_fn geom_alloc, void *geom_user, fsgs_alloc_fn binning_alloc,
                           void *binning_user, fsgs_alloc_fn img_alloc, void *img_user, float *out_color,
                           float *out_depth, int32_t *radii, int64_t *num_rendered_host, int64_t *num_rect_host,
                           void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CamConst cc;
    int rc = make_cam(st, cc);
    if (rc) return rc;
    if (P < 0 || !bg || !out_color || !out_depth || !viewmatrix || !projmatrix || !geom_alloc || !binning_alloc ||
        !img_alloc)
        return FSGS_E_INVALID;
    if ((colors_precomp == nullptr) == (shs == nullptr) && P > 0) return FSGS_E_INVALID;
    if (P > 0 && (!means3D || !opacities || !radii)) return FSGS_E_INVALID;
    if (P > 0 && ((cov3D_precomp != nullptr) == (scales != nullptr && rotations != nullptr))) return FSGS_E_INVALID;
    if (shs && (!campos || st->n_coeffs < (st->sh_degree + 1) * (st->sh_degree + 1) || st->n_coeffs > 16))
        return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    if ((rc = one_time_setup())) return rc;
    if (num_rendered_host) *num_rendered_host = 0;
    if (num_rect_host) *num_rect_host = 0;
    const int HW = cc.W * cc.H;
    if (P == 0) {
        k_fill_bg<<<blocks(HW), CTA, 0, stream>>>(HW, 3, bg, out_color, out_depth);
        FSGS_LAUNCH_OK("k_fill_bg");
        return FSGS_OK;
    }
    Buffers B;
    if ((rc = alloc_fixed(P, cc, geom_alloc, geom_user, img_alloc, img_user, B, stream))) return rc;
    prof_begin(K_PRE_API, stream);
    if ((rc = prepare_binning(st, B.il.tiles, binning_alloc, binning_user, B))) return rc;
    k_preprocess_api<<<blocks(P), CTA, 0, stream>>>(
        cc, P, means3D, colors_precomp, shs, opacities, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
        reinterpret_cast<float4 *>(B.geom + B.gl.records), reinterpret_cast<uint8_t *>(B.geom + B.gl.clamped), radii,
        reinterpret_cast<unsigned int *>(B.img + B.il.tile_count),
        reinterpret_cast<unsigned long long *>(B.img + B.il.counters), (unsigned)st->flags, bin_sink(B));
    prof_end(K_PRE_API, stream);
    FSGS_LAUNCH_OK("k_preprocess_api");
    return forward_tail<false>(st, cc, P, bg, B, binning_alloc, binning_user, out_color, out_depth, num_rendered_host,
                               num_rect_host, RenderExtras{}, stream);
}

int fsgs_rasterize_backward(const fsgs_settings *st, int32_t P, int64_t num_rendered, const float *bg,
                            const float *means3D, const float *colors_precomp, const float *shs,
                            const float *opacities, const float *scales, const float *rotations,
                            const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
                            const float *campos, const void *geom, const void *binning, const void *img,
                            const float *dL_dout_color, const float *dL_dout_depth, void *grad_scratch,
                            float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity, float *dL_dmeans3D,
                            float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CamConst cc;
    int rc = make_cam(st, cc);
    if (rc) return rc;
    if (P <= 0) return P == 0 ? FSGS_OK : FSGS_E_INVALID;
    if (!geom || !img || !binning || !dL_dout_color || !grad_scratch || !means3D || !viewmatrix || !projmatrix || !bg)
        return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    if ((rc = one_time_setup())) return rc;
    const GeomLayout gl = geom_layout(P);
    const ImgLayout il = img_layout(cc.W, cc.H);
    const BinLayout bl = bin_layout(num_rendered);
    const char *g = static_cast<const char *>(geom), *im = static_cast<const char *>(img),
               *bn = static_cast<const char *>(binning);
    float *acc = static_cast<float *>(grad_scratch);
    FSGS_CUDA(cudaMemsetAsync(acc, 0, (size_t)P * ACC_F * 4, stream));
    if (num_rendered > 0) {
        prof_begin(K_COMP_BWD, stream);
        if (st->flags & FSGS_FLAG_UPSTREAM_STYLE) {
            k_composite_bwd_ref<<<il.tiles, CTA, 0, stream>>>(
                cc, reinterpret_cast<const unsigned int *>(im + il.tile_offset),
                reinterpret_cast<const float4 *>(bn + bl.records), bg, reinterpret_cast<const float *>(im + il.final_T),
                reinterpret_cast<const unsigned int *>(im + il.n_contrib), dL_dout_color, dL_dout_depth, acc,
                reinterpret_cast<const unsigned long long *>(im + il.counters), (unsigned long long)num_rendered,
                fixed_bin_cap(st));
        } else
        k_composite_bwd<false><<<il.tiles, CTA, sizeof(BwdSmem), stream>>>(
            cc, reinterpret_cast<const unsigned int *>(im + il.tile_offset), reinterpret_cast<const float4 *>(bn + bl.records),
            bg, reinterpret_cast<const float *>(im + il.final_T), reinterpret_cast<const unsigned int *>(im + il.n_contrib),
            dL_dout_color, dL_dout_depth, nullptr, nullptr, acc, (unsigned)st->flags, sticky_flag(),
            reinterpret_cast<const unsigned long long *>(im + il.counters), (unsigned long long)num_rendered,
            fixed_bin_cap(st));
        prof_end(K_COMP_BWD, stream);
        FSGS_LAUNCH_OK("k_composite_bwd");
    }
    prof_begin(K_PRE_API_BWD, stream);
    k_preprocess_api_bwd<<<blocks(P), CTA, 0, stream>>>(
        cc, P, means3D, shs, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
        reinterpret_cast<const float4 *>(g + gl.records), reinterpret_cast<const uint8_t *>(g + gl.clamped), acc,
        dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations);
    prof_end(K_PRE_API_BWD, stream);
    FSGS_LAUNCH_OK("k_preprocess_api_bwd");
    (void)colors_precomp; (void)opacities;
    return FSGS_OK;
}

int fsgs_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                      uint8_t *visible, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !visible))) return FSGS_E_INVALID;
    if (P == 0) return FSGS_OK;
    int rc = check_arch();
    if (rc) return rc;
    prof_begin(K_MARK_VISIBLE, stream);
    k_mark_visible<<<blocks(P), CTA, 0, stream>>>(P, means3D, viewmatrix, visible);
    prof_end(K_MARK_VISIBLE, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_cuda_msg, sizeof(g_cuda_msg), "kernel k_mark_visible failed: %s", cudaGetErrorString(e));
        return FSGS_E_CUDA;
    }
    return FSGS_OK;
}

int fsgs_pose_forward(const float *r, const float *t, int32_t cam, int32_t n_cams, float *Rt, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!r || !t || !Rt || n_cams <= 0 || cam < 0 || cam >= n_cams) return FSGS_E_INVALID;
    k_pose_forward<<<1, 32, 0, stream>>>(r, t, cam, n_cams, Rt);
    if (cudaGetLastError() != cudaSuccess) return FSGS_E_CUDA;
    return FSGS_OK;
}

int fsgs_pose_backward(const float *r, int32_t cam, int32_t n_cams, const float *dRt, float *dr, float *dt,
                       void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!r || !dRt || !dr || !dt || n_cams <= 0 || cam < 0 || cam >= n_cams) return FSGS_E_INVALID;
    FSGS_CUDA(cudaMemsetAsync(dr, 0, sizeof(float) * 4 * n_cams, stream));
    FSGS_CUDA(cudaMemsetAsync(dt, 0, sizeof(float) * 3 * n_cams, stream));
    k_pose_backward<<<1, 32, 0, stream>>>(r, cam, n_cams, dRt, dr, dt);
    if (cudaGetLastError() != cudaSuccess) return FSGS_E_CUDA;
    return FSGS_OK;
}

int fsgs_render_forward(const fsgs_settings *st, int32_t P, const float *bg, const float *xyz,
                        const float *features_dc, const float *features_rest, const float *opacity_raw,
                        const float *scaling_raw, const float *rotation_raw, const float *pose,
                        const float *cam_center, const float *viewmatrix, const float *projmatrix,
                        fsgs_alloc_fn geom_alloc, void *geom_user, fsgs_alloc_fn binning_alloc, void *binning_user,
                        fsgs_alloc_fn img_alloc, void *img_user, float *out_planes, int32_t *radii,
                        int64_t *num_rendered_host, int64_t *num_rect_host, void *stream_) {
    return fsgs_render_forward_ex(st, P, bg, xyz, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw, pose,
                                  cam_center, viewmatrix, projmatrix, geom_alloc, geom_user, binning_alloc, binning_user,
                                  img_alloc, img_user, out_planes, radii, num_rendered_host, num_rect_host, nullptr,
                                  stream_);
}

int fsgs_render_forward_ex(const fsgs_settings *st, int32_t P, const float *bg, const float *xyz,
                           const float *features_dc, const float *features_rest, const float *opacity_raw,
                           const float *scaling_raw, const float *rotation_raw, const float *pose,
                           const float *cam_center, const float *viewmatrix, const float *projmatrix,
                           fsgs_alloc_fn geom_alloc, void *geom_user, fsgs_alloc_fn binning_alloc, void *binning_user,
                           fsgs_alloc_fn img_alloc, void *img_user, float *out_planes, int32_t *radii,
                           int64_t *num_rendered_host, int64_t *num_rect_host, const fsgs_render_extras *extras,
                           void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RenderExtras ex{};
    if (extras) {
        ex.uncertainty = extras->uncertainty; ex.presence_mask = extras->presence_mask; ex.nan_mask = extras->nan_mask;
        ex.visibility = extras->visibility; ex.max_radii2D = extras->max_radii2D;
    }
    CamConst cc;
    int rc = make_cam(st, cc);
    if (rc) return rc;
    if (P < 0 || !bg || !out_planes || !viewmatrix || !projmatrix || !pose || !cam_center || !geom_alloc ||
        !binning_alloc || !img_alloc)
        return FSGS_E_INVALID;
    if (P > 0 && (!xyz || !features_dc || !features_rest || !opacity_raw || !scaling_raw || !rotation_raw || !radii))
        return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    if ((rc = one_time_setup())) return rc;
    cc.n_coeffs = 16;
    if (num_rendered_host) *num_rendered_host = 0;
    if (num_rect_host) *num_rect_host = 0;
    const int HW = cc.W * cc.H;
    if (P == 0) {
        k_fill_bg<<<blocks(HW), CTA, 0, stream>>>(HW, 6, bg, out_planes, nullptr);
        FSGS_LAUNCH_OK("k_fill_bg");
        k_fill_extras<<<blocks(HW), CTA, 0, stream>>>(HW, bg, ex);
        FSGS_LAUNCH_OK("k_fill_extras");
        return FSGS_OK;
    }
    Buffers B;
    if ((rc = alloc_fixed(P, cc, geom_alloc, geom_user, img_alloc, img_user, B, stream))) return rc;
    if ((rc = prepare_binning(st, B.il.tiles, binning_alloc, binning_user, B))) return rc;
    prof_begin(K_PRE_FUSED, stream);
    k_preprocess_fused<<<(P + PRE_CTA - 1) / PRE_CTA, PRE_CTA, 0, stream>>>(
        cc, P, xyz, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw, pose, cam_center, viewmatrix,
        projmatrix, reinterpret_cast<float4 *>(B.geom + B.gl.records), reinterpret_cast<uint8_t *>(B.geom + B.gl.clamped),
        radii, reinterpret_cast<unsigned int *>(B.img + B.il.tile_count),
        reinterpret_cast<unsigned long long *>(B.img + B.il.counters), (unsigned)st->flags, ex.visibility, ex.max_radii2D,
        sticky_flag(), bin_sink(B));
    prof_end(K_PRE_FUSED, stream);
    FSGS_LAUNCH_OK("k_preprocess_fused");
    return forward_tail<true>(st, cc, P, bg, B, binning_alloc, binning_user, out_planes, nullptr, num_rendered_host,
                              num_rect_host, ex, stream);
}

// ---- frozen-model forward (tracking loops: many poses, one Gaussian model) -------------------------------
size_t fsgs_frozen_bytes(int32_t P) { return (size_t)(P > 0 ? P : 1) * 64; }

int fsgs_freeze_model(const fsgs_settings *st, int32_t P, const float *xyz, const float *features_dc,
                      const float *features_rest, const float *opacity_raw, const float *scaling_raw,
                      const float *rotation_raw, const float *cam_center, void *frozen, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CamConst cc;
    int rc = make_cam(st, cc);
    if (rc) return rc;
    if (P < 0 || !cam_center || (P > 0 && (!xyz || !features_dc || !features_rest || !opacity_raw || !scaling_raw ||
                                           !rotation_raw || !frozen)))
        return FSGS_E_INVALID;
    if ((reinterpret_cast<uintptr_t>(frozen) | reinterpret_cast<uintptr_t>(rotation_raw)) & 15u) return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    if (P == 0) return FSGS_OK;
    cc.n_coeffs = 16;
    prof_begin(K_FREEZE, stream);
    k_freeze_model<<<blocks(P), CTA, 0, stream>>>(cc, P, xyz, features_dc, features_rest, opacity_raw, scaling_raw,
                                                  rotation_raw, cam_center, static_cast<float4 *>(frozen));
    prof_end(K_FREEZE, stream);
    FSGS_LAUNCH_OK("k_freeze_model");
    return FSGS_OK;
}

int fsgs_render_forward_frozen(const fsgs_settings *st, int32_t P, const float *bg, const void *frozen, const float *pose,
                               const float *viewmatrix, const float *projmatrix, fsgs_alloc_fn geom_alloc,
                               void *geom_user, fsgs_alloc_fn binning_alloc, void *binning_user, fsgs_alloc_fn img_alloc,
                               void *img_user, float *out_planes, int32_t *radii, int64_t *num_rendered_host,
                               int64_t *num_rect_host, const fsgs_render_extras *extras, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RenderExtras ex{};
    if (extras) {
        ex.uncertainty = extras->uncertainty; ex.presence_mask = extras->presence_mask; ex.nan_mask = extras->nan_mask;
        ex.visibility = extras->visibility; ex.max_radii2D = extras->max_radii2D;
    }
    CamConst cc;
    int rc = make_cam(st, cc);
    if (rc) return rc;
    if (P < 0 || !bg || !out_planes || !viewmatrix || !projmatrix || !pose || !geom_alloc || !binning_alloc || !img_alloc)
        return FSGS_E_INVALID;
    if (P > 0 && (!frozen || !radii || (reinterpret_cast<uintptr_t>(frozen) & 15u))) return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    if ((rc = one_time_setup())) return rc;
    cc.n_coeffs = 16;
    if (num_rendered_host) *num_rendered_host = 0;
    if (num_rect_host) *num_rect_host = 0;
    const int HW = cc.W * cc.H;
    if (P == 0) {
        k_fill_bg<<<blocks(HW), CTA, 0, stream>>>(HW, 6, bg, out_planes, nullptr);
        FSGS_LAUNCH_OK("k_fill_bg");
        k_fill_extras<<<blocks(HW), CTA, 0, stream>>>(HW, bg, ex);
        FSGS_LAUNCH_OK("k_fill_extras");
        return FSGS_OK;
    }
    Buffers B;
    if ((rc = alloc_fixed(P, cc, geom_alloc, geom_user, img_alloc, img_user, B, stream))) return rc;
    if ((rc = prepare_binning(st, B.il.tiles, binning_alloc, binning_user, B))) return rc;
    prof_begin(K_PRE_FROZEN, stream);
    k_preprocess_frozen<<<blocks(P), CTA, 0, stream>>>(
        cc, P, static_cast<const float4 *>(frozen), pose, viewmatrix, projmatrix,
        reinterpret_cast<float4 *>(B.geom + B.gl.records), reinterpret_cast<uint8_t *>(B.geom + B.gl.clamped), radii,
        reinterpret_cast<unsigned int *>(B.img + B.il.tile_count),
        reinterpret_cast<unsigned long long *>(B.img + B.il.counters), (unsigned)st->flags, ex.visibility, ex.max_radii2D,
        bin_sink(B));
    prof_end(K_PRE_FROZEN, stream);
    FSGS_LAUNCH_OK("k_preprocess_frozen");
    return forward_tail<true>(st, cc, P, bg, B, binning_alloc, binning_user, out_planes, nullptr, num_rendered_host,
                              num_rect_host, ex, stream);
}

int fsgs_render_backward(const fsgs_settings *st, int32_t P, int64_t num_rendered, const float *bg, const float *xyz,
                         const float *features_dc, const float *features_rest, const float *opacity_raw,
                         const float *scaling_raw, const float *rotation_raw, const float *pose,
                         const float *cam_center, const float *viewmatrix, const float *projmatrix, const void *geom,
                         const void *binning, const void *img, const float *dL_dplanes, void *grad_scratch,
                         int32_t gs_grad, int32_t cam_grad, float *dL_dxyz, float *dL_dfeatures_dc,
                         float *dL_dfeatures_rest, float *dL_dopacity_raw, float *dL_dscaling_raw,
                         float *dL_drotation_raw, float *dL_dpose, float *dL_dmeans2D, void *stream_) {
    if (!dL_dplanes || !st) return FSGS_E_INVALID;
    const size_t HW = (size_t)st->image_width * (size_t)st->image_height;
    return fsgs_render_backward_ex(st, P, num_rendered, bg, xyz, features_dc, features_rest, opacity_raw, scaling_raw,
                                   rotation_raw, pose, cam_center, viewmatrix, projmatrix, geom, binning, img, dL_dplanes,
                                   dL_dplanes + 3 * HW, dL_dplanes + 4 * HW, dL_dplanes + 5 * HW, grad_scratch, gs_grad,
                                   cam_grad, dL_dxyz, dL_dfeatures_dc, dL_dfeatures_rest, dL_dopacity_raw, dL_dscaling_raw,
                                   dL_drotation_raw, dL_dpose, dL_dmeans2D, nullptr, stream_);
}

int fsgs_render_backward_ex(const fsgs_settings *st, int32_t P, int64_t num_rendered, const float *bg, const float *xyz,
                            const float *features_dc, const float *features_rest, const float *opacity_raw,
                            const float *scaling_raw, const float *rotation_raw, const float *pose,
                            const float *cam_center, const float *viewmatrix, const float *projmatrix, const void *geom,
                            const void *binning, const void *img, const float *dL_drgb, const float *dL_ddepth,
                            const float *dL_dsil, const float *dL_ddepth_sq, void *grad_scratch, int32_t gs_grad,
                            int32_t cam_grad, float *dL_dxyz, float *dL_dfeatures_dc, float *dL_dfeatures_rest,
                            float *dL_dopacity_raw, float *dL_dscaling_raw, float *dL_drotation_raw, float *dL_dpose,
                            float *dL_dmeans2D, float *dL_dsh_rgb, void *stream_) {
    return fsgs_render_backward_v2(st, P, num_rendered, bg, xyz, features_dc, features_rest, opacity_raw, scaling_raw,
                                   rotation_raw, pose, cam_center, viewmatrix, projmatrix, geom, binning, img, dL_drgb,
                                   dL_ddepth, dL_dsil, dL_ddepth_sq, grad_scratch, gs_grad, cam_grad, dL_dxyz,
                                   dL_dfeatures_dc, dL_dfeatures_rest, dL_dopacity_raw, dL_dscaling_raw, dL_drotation_raw,
                                   dL_dpose, dL_dmeans2D, dL_dsh_rgb, nullptr, stream_);
}

int fsgs_render_backward_v2(const fsgs_settings *st, int32_t P, int64_t num_rendered, const float *bg, const float *xyz,
                            const float *features_dc, const float *features_rest, const float *opacity_raw,
                            const float *scaling_raw, const float *rotation_raw, const float *pose,
                            const float *cam_center, const float *viewmatrix, const float *projmatrix, const void *geom,
                            const void *binning, const void *img, const float *dL_drgb, const float *dL_ddepth,
                            const float *dL_dsil, const float *dL_ddepth_sq, void *grad_scratch, int32_t gs_grad,
                            int32_t cam_grad, float *dL_dxyz, float *dL_dfeatures_dc, float *dL_dfeatures_rest,
                            float *dL_dopacity_raw, float *dL_dscaling_raw, float *dL_drotation_raw, float *dL_dpose,
                            float *dL_dmeans2D, float *dL_dsh_rgb, const fsgs_backward_opts *opts, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CamConst cc;
    int rc = make_cam(st, cc);
    if (rc) return rc;
    cc.n_coeffs = 16;
    fsgs_backward_opts o{};
    if (opts) o = *opts;
    const bool ranged = o.count > 0;
    if (ranged && (o.first < 0 || (o.first & 3) || (int64_t)o.first + o.count > P)) return FSGS_E_INVALID;
    if ((o.xyz_gradient_accum == nullptr) != (o.denom == nullptr)) return FSGS_E_INVALID;
    if (o.compact && (dL_dxyz || dL_dopacity_raw || dL_dscaling_raw || dL_drotation_raw || dL_dsh_rgb)) return FSGS_E_INVALID;
    const int first = ranged ? o.first : 0, end = ranged ? o.first + o.count : P;
    if (dL_dpose && !o.skip_composite) FSGS_CUDA(cudaMemsetAsync(dL_dpose, 0, 16 * sizeof(float), stream));
    if (P <= 0) return P == 0 ? FSGS_OK : FSGS_E_INVALID;
    if (!geom || !img || !binning || !grad_scratch || !xyz || !features_dc || !features_rest ||
        !opacity_raw || !scaling_raw || !rotation_raw || !pose || !cam_center || !viewmatrix || !projmatrix || !bg)
        return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    if ((rc = one_time_setup())) return rc;
    const GeomLayout gl = geom_layout(P);
    const ImgLayout il = img_layout(cc.W, cc.H);
    const BinLayout bl = bin_layout(num_rendered);
    const char *g = static_cast<const char *>(geom), *im = static_cast<const char *>(img),
               *bn = static_cast<const char *>(binning);
    float *acc = static_cast<float *>(grad_scratch);
    if (!o.skip_composite) FSGS_CUDA(cudaMemsetAsync(acc, 0, (size_t)P * ACC_F * 4, stream));
    // Pose-only request (tracking against a frozen Gaussian model): dL/dpose is the only output asked for.
    // The compositor then skips the colour / opacity / RGB-only columns and a lean per-Gaussian kernel
    // reduces dL/dRt without touching the SH coefficients or writing per-Gaussian gradients.
    const bool pose_only = cam_grad && dL_dpose && !dL_dxyz && !dL_dfeatures_dc && !dL_dfeatures_rest &&
                           !dL_dopacity_raw && !dL_dscaling_raw && !dL_drotation_raw && !dL_dmeans2D && !dL_dsh_rgb &&
                           !o.compact && !o.xyz_gradient_accum && !ranged && !(st->flags & FSGS_FLAG_NO_POSE_ONLY);
    if (num_rendered > 0 && !o.skip_composite) {
        prof_begin(K_COMP_BWD, stream);
        auto kern = pose_only ? k_composite_bwd<true, true> : k_composite_bwd<true, false>;
        kern<<<il.tiles, CTA, sizeof(BwdSmem), stream>>>(
            cc, reinterpret_cast<const unsigned int *>(im + il.tile_offset), reinterpret_cast<const float4 *>(bn + bl.records),
            bg, reinterpret_cast<const float *>(im + il.final_T), reinterpret_cast<const unsigned int *>(im + il.n_contrib),
            dL_drgb, dL_ddepth, dL_dsil, dL_ddepth_sq, acc, (unsigned)st->flags, sticky_flag(),
            reinterpret_cast<const unsigned long long *>(im + il.counters), (unsigned long long)num_rendered,
            fixed_bin_cap(st));
        prof_end(K_COMP_BWD, stream);
        FSGS_LAUNCH_OK("k_composite_bwd");
    }
    if (pose_only) {
        prof_begin(K_PRE_POSE_BWD, stream);
        k_preprocess_pose_bwd<<<blocks(P), CTA, 0, stream>>>(
            cc, P, xyz, scaling_raw, rotation_raw, pose, viewmatrix, projmatrix,
            reinterpret_cast<const float4 *>(g + gl.records), acc, dL_dpose);
        prof_end(K_PRE_POSE_BWD, stream);
        FSGS_LAUNCH_OK("k_preprocess_pose_bwd");
        return FSGS_OK;
    }
    prof_begin(K_PRE_FUSED_BWD, stream);
    k_preprocess_fused_bwd<<<(end - first + PREBWD_CTA - 1) / PREBWD_CTA, PREBWD_CTA, 0, stream>>>(
        cc, P, xyz, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw, pose, cam_center, viewmatrix,
        projmatrix, reinterpret_cast<const float4 *>(g + gl.records), reinterpret_cast<const uint8_t *>(g + gl.clamped),
        acc, gs_grad, cam_grad, dL_dxyz, dL_dfeatures_dc, dL_dfeatures_rest, dL_dopacity_raw, dL_dscaling_raw,
        dL_drotation_raw, dL_dpose, dL_dmeans2D, (st->flags & FSGS_FLAG_NO_TMA) ? 0 : 1, sticky_flag(), dL_dsh_rgb,
        first, end, o.xyz_gradient_accum, o.denom, o.compact);
    prof_end(K_PRE_FUSED_BWD, stream);
    FSGS_LAUNCH_OK("k_preprocess_fused_bwd");
    return FSGS_OK;
}

int fsgs_watchdog_flag(int32_t device, int32_t reset) {
    if (device < 0 || device >= 64) return FSGS_E_INVALID;
    if (!g_sticky[device]) return 0;      // no launch on this device yet
    int cur = 0;
    FSGS_CUDA(cudaGetDevice(&cur));
    if (cur != device) FSGS_CUDA(cudaSetDevice(device));
    unsigned long long v = 0;
    cudaError_t e = cudaMemcpy(&v, g_sticky[device], sizeof(v), cudaMemcpyDeviceToHost);     // synchronises
    if (e == cudaSuccess && v && reset) e = cudaMemset(g_sticky[device], 0, sizeof(v));
    if (cur != device) cudaSetDevice(cur);
    if (e != cudaSuccess) {
        snprintf(g_cuda_msg, sizeof(g_cuda_msg), "fsgs_watchdog_flag: %s", cudaGetErrorString(e));
        return FSGS_E_CUDA;
    }
    return v ? 1 : 0;
}

#ifdef FSGS_PAIR_STATS
// instrumented A/B build only (tools/pair_stats.py): read and clear the compositors' pair statistics
extern "C" int fsgs_debug_pair_stats(unsigned long long *out8) {
    FSGS_CUDA(cudaDeviceSynchronize());
    FSGS_CUDA(cudaMemcpyFromSymbol(out8, g_pair_stats, 8 * sizeof(unsigned long long)));
    unsigned long long z[8] = {};
    FSGS_CUDA(cudaMemcpyToSymbol(g_pair_stats, z, sizeof(z)));
    return FSGS_OK;
}
#endif

int64_t fsgs_fixed_bin_capacity(int32_t device) {
    if (device < 0 || device >= 64) return FSGS_E_INVALID;
    return (int64_t)g_fixed_bin_cap[device];
}

int fsgs_set_instance_capacity(int32_t device, int64_t capacity) {
    if (device < 0 || device >= 64 || capacity < 0) return FSGS_E_INVALID;
    g_fixed_cap[device] = (unsigned long long)capacity;
    return FSGS_OK;
}

int fsgs_sh_grad_expand(const fsgs_settings *st, int32_t P, const float *xyz, const float *cam_center,
                        const float *dL_dsh_rgb, float *dL_dfeatures_dc, float *dL_dfeatures_rest, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!st || st->sh_degree < 0 || st->sh_degree > 3 || P < 0) return FSGS_E_INVALID;
    if (P == 0) return FSGS_OK;
    if (!xyz || !cam_center || !dL_dsh_rgb || !dL_dfeatures_dc || !dL_dfeatures_rest) return FSGS_E_INVALID;
    int rc = check_arch();
    if (rc) return rc;
    prof_begin(K_SH_EXPAND, stream);
    k_sh_grad_expand<<<blocks(P), CTA, 0, stream>>>(P, st->sh_degree, xyz, cam_center, dL_dsh_rgb, dL_dfeatures_dc,
                                                    dL_dfeatures_rest, (st->flags & FSGS_FLAG_NO_TMA) ? 0 : 1, 0, nullptr,
                                                    nullptr, nullptr, nullptr, nullptr, PeerRows{});
    prof_end(K_SH_EXPAND, stream);
    FSGS_LAUNCH_OK("k_sh_grad_expand");
    return FSGS_OK;
}

int fsgs_exchange_rows(void *multicast_ptr, void *const *peer_ptrs_host, int32_t world, int32_t rank, int64_t first_vec4,
                       int64_t n_vec4, void *stream_) {
    return fsgs_exchange_rows_scatter(multicast_ptr, peer_ptrs_host, world, rank, first_vec4, n_vec4, 0, stream_);
}

int fsgs_exchange_rows_scatter(void *multicast_ptr, void *const *peer_ptrs_host, int32_t world, int32_t rank,
                               int64_t first_vec4, int64_t n_vec4, int64_t scatter_offset_vec4, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (world < 1 || world > 8 || rank < 0 || rank >= world || first_vec4 < 0 || n_vec4 < 0 || scatter_offset_vec4 < 0)
        return FSGS_E_INVALID;
    if (!multicast_ptr && !peer_ptrs_host) return FSGS_E_INVALID;
    if (scatter_offset_vec4 > 0 && !peer_ptrs_host) return FSGS_E_INVALID;      // the own buffer's address is needed
    if (n_vec4 == 0 || world == 1) return FSGS_OK;
    int rc = check_arch();
    if (rc) return rc;
    PeerPtrs pp{};
    if (peer_ptrs_host)
        for (int r = 0; r < world; ++r) {
            if (!multicast_ptr && !peer_ptrs_host[r]) return FSGS_E_INVALID;
            pp.p[r] = static_cast<float4 *>(peer_ptrs_host[r]);
        }
    // this rank's slice of [first, first + n): equal parts, the remainder to the last rank
    const int64_t per = n_vec4 / world;
    const int64_t begin = first_vec4 + per * rank;
    const int64_t end = (rank == world - 1) ? first_vec4 + n_vec4 : begin + per;
    if (end <= begin) return FSGS_OK;
    int64_t nb = (end - begin + CTA * 4 - 1) / (CTA * 4);
    if (nb > 148 * 8) nb = 148 * 8;
    prof_begin(K_EXCHANGE, stream);
    k_exchange_rows<<<(int)nb, CTA, 0, stream>>>(static_cast<float4 *>(multicast_ptr), pp, world, rank, (long long)begin,
                                                  (long long)end, (long long)scatter_offset_vec4);
    prof_end(K_EXCHANGE, stream);
    if (cudaGetLastError() != cudaSuccess) return FSGS_E_CUDA;
    return FSGS_OK;
}

int fsgs_compact_grad_expand_peers(const fsgs_settings *st, int32_t P, int32_t first, int32_t count, const float *xyz,
                                   const float *cam_center, void *const *row_ptrs_host, int32_t world,
                                   int32_t owner_slices, int64_t slice_first_vec4, int64_t slice_n_vec4, float *dL_dxyz,
                                   float *dL_dfeatures_dc, float *dL_dfeatures_rest, float *dL_dopacity_raw,
                                   float *dL_dscaling_raw, float *dL_drotation_raw, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!st || st->sh_degree < 0 || st->sh_degree > 3 || P < 0 || first < 0 || (first & 255) || count < 0 ||
        (int64_t)first + count > P || world < 2 || world > 8 || !row_ptrs_host)
        return FSGS_E_INVALID;
    if (owner_slices && (slice_first_vec4 < 0 || slice_n_vec4 < world || slice_first_vec4 + slice_n_vec4 >= (1ll << 31)))
        return FSGS_E_INVALID;
    if (count == 0) return FSGS_OK;
    if (!xyz || !cam_center || !dL_dxyz || !dL_dfeatures_dc || !dL_dfeatures_rest || !dL_dopacity_raw ||
        !dL_dscaling_raw || !dL_drotation_raw)
        return FSGS_E_INVALID;
    PeerRows pr{};
    pr.world = world;
    pr.owner_slices = owner_slices ? 1 : 0;
    pr.slice_first = (int)slice_first_vec4;
    pr.slice_per = owner_slices ? (int)(slice_n_vec4 / world) : 1;      // the slices of fsgs_exchange_rows: equal parts, rest to the last
    for (int r = 0; r < world; ++r) {
        if (!row_ptrs_host[r] || (reinterpret_cast<uintptr_t>(row_ptrs_host[r]) & 15u)) return FSGS_E_INVALID;
        pr.p[r] = static_cast<const float4 *>(row_ptrs_host[r]);
    }
    int rc = check_arch();
    if (rc) return rc;
    prof_begin(K_SH_EXPAND, stream);
    k_sh_grad_expand<<<blocks(count), CTA, 0, stream>>>(first + count, st->sh_degree, xyz, cam_center, nullptr,
                                                        dL_dfeatures_dc, dL_dfeatures_rest,
                                                        (st->flags & FSGS_FLAG_NO_TMA) ? 0 : 1, first, nullptr, dL_dxyz,
                                                        dL_dopacity_raw, dL_dscaling_raw, dL_drotation_raw, pr);
    prof_end(K_SH_EXPAND, stream);
    FSGS_LAUNCH_OK("k_sh_grad_expand");
    return FSGS_OK;
}

int fsgs_compact_grad_expand(const fsgs_settings *st, int32_t P, int32_t first, int32_t count, const float *xyz,
                             const float *cam_center, const float *compact, float *dL_dxyz, float *dL_dfeatures_dc,
                             float *dL_dfeatures_rest, float *dL_dopacity_raw, float *dL_dscaling_raw,
                             float *dL_drotation_raw, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!st || st->sh_degree < 0 || st->sh_degree > 3 || P < 0 || first < 0 || (first & 3) || count < 0 ||
        (int64_t)first + count > P)
        return FSGS_E_INVALID;
    if (count == 0) return FSGS_OK;
    if (!xyz || !cam_center || !compact || !dL_dxyz || !dL_dfeatures_dc || !dL_dfeatures_rest || !dL_dopacity_raw ||
        !dL_dscaling_raw || !dL_drotation_raw)
        return FSGS_E_INVALID;
    int rc = check_arch();
    if (rc) return rc;
    prof_begin(K_SH_EXPAND, stream);
    k_sh_grad_expand<<<blocks(count), CTA, 0, stream>>>(first + count, st->sh_degree, xyz, cam_center, nullptr,
                                                        dL_dfeatures_dc, dL_dfeatures_rest,
                                                        (st->flags & FSGS_FLAG_NO_TMA) ? 0 : 1, first, compact, dL_dxyz,
                                                        dL_dopacity_raw, dL_dscaling_raw, dL_drotation_raw, PeerRows{});
    prof_end(K_SH_EXPAND, stream);
    FSGS_LAUNCH_OK("k_sh_grad_expand");
    return FSGS_OK;
}

// ---- fused image loss --------------------------------------------------------------
static LossWindow make_loss_window() {
    // the reference's 1-D window: exp(-(i - 5)^2 / (2 * 1.5^2)) normalised, float32 (utils/loss_utils.py:56-58)
    LossWindow w;
    float sum = 0.f;
    for (int i = 0; i < 2 * LOSS_R + 1; ++i) {
        const float d = (float)(i - LOSS_R);
        w.g[i] = expf(-(d * d) / (2.0f * 1.5f * 1.5f));
        sum += w.g[i];
    }
    for (int i = 0; i < 2 * LOSS_R + 1; ++i) w.g[i] /= sum;
    return w;
}
static inline size_t loss_blocks(int C, int H, int W) {
    return (size_t)C * (size_t)((H + LOSS_TH - 1) / LOSS_TH) * (size_t)((W + LOSS_TW - 1) / LOSS_TW);
}

size_t fsgs_rgb_loss_scratch_bytes(int32_t C, int32_t H, int32_t W) {
    if (C <= 0 || H <= 0 || W <= 0) return 256;
    return align_up(loss_blocks(C, H, W) * 2 * sizeof(double), 256);
}

static int loss_args_ok(int32_t C, int32_t H, int32_t W, const float *img, const float *gt, const unsigned char *mask_u8,
                        const float *mask_f32, int64_t mask_cstride) {
    if (C <= 0 || C > 65535 || H <= 0 || W <= 0 || !img || !gt) return FSGS_E_INVALID;
    if (mask_u8 && mask_f32) return FSGS_E_INVALID;
    if (mask_cstride != 0 && mask_cstride != (int64_t)H * W) return FSGS_E_INVALID;
    if ((H + LOSS_TH - 1) / LOSS_TH > 65535) return FSGS_E_INVALID;
    return FSGS_OK;
}

int fsgs_rgb_loss_forward(int32_t C, int32_t H, int32_t W, const float *img, const float *gt, const unsigned char *mask_u8,
                          const float *mask_f32, int64_t mask_cstride, float lambda_dssim, float *maps, void *scratch,
                          float *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = loss_args_ok(C, H, W, img, gt, mask_u8, mask_f32, mask_cstride);
    if (rc) return rc;
    if (!scratch || !out) return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    const LossMask mask{mask_u8, mask_f32, (long long)mask_cstride};
    const LossWindow win = make_loss_window();
    const dim3 grid((W + LOSS_TW - 1) / LOSS_TW, (H + LOSS_TH - 1) / LOSS_TH, C);
    double *partial = static_cast<double *>(scratch);
    prof_begin(K_LOSS_FWD, stream);
    k_rgb_loss_fwd<<<grid, CTA, 0, stream>>>(H, W, img, gt, mask, win, maps, partial);
    k_rgb_loss_reduce<<<1, CTA, 0, stream>>>((long long)loss_blocks(C, H, W), partial, 1.0 / ((double)C * H * W),
                                             lambda_dssim, out);
    prof_end(K_LOSS_FWD, stream);
    FSGS_CUDA(cudaGetLastError());
    return FSGS_OK;
}

int fsgs_rgb_loss_backward(int32_t C, int32_t H, int32_t W, const float *img, const float *gt, const unsigned char *mask_u8,
                           const float *mask_f32, int64_t mask_cstride, float lambda_dssim, const float *maps,
                           const float *upstream, float *dimg, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = loss_args_ok(C, H, W, img, gt, mask_u8, mask_f32, mask_cstride);
    if (rc) return rc;
    if (!maps || !dimg) return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    const LossMask mask{mask_u8, mask_f32, (long long)mask_cstride};
    const LossWindow win = make_loss_window();
    const dim3 grid((W + LOSS_TW - 1) / LOSS_TW, (H + LOSS_TH - 1) / LOSS_TH, C);
    prof_begin(K_LOSS_BWD, stream);
    k_rgb_loss_bwd<<<grid, CTA, 0, stream>>>(H, W, img, gt, mask, win, maps, upstream, lambda_dssim,
                                             (float)(1.0 / ((double)C * H * W)), dimg);
    prof_end(K_LOSS_BWD, stream);
    FSGS_CUDA(cudaGetLastError());
    return FSGS_OK;
}

// ---- fused Pearson depth loss ---------------------------------------------------------------------------
static inline int pearson_blocks(int64_t n) {
    const int64_t b = (n + (int64_t)CTA * 4 - 1) / ((int64_t)CTA * 4);
    return (int)(b < 1 ? 1 : (b > PEARSON_MAX_BLOCKS ? PEARSON_MAX_BLOCKS : b));
}

size_t fsgs_pearson_scratch_bytes(void) { return align_up((size_t)PEARSON_MAX_BLOCKS * 5 * sizeof(double), 256); }

int fsgs_pearson_forward(int64_t n, const float *src, const float *target, void *scratch, double *stats, float *out,
                         void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0 || !src || !target || !scratch || !stats || !out) return FSGS_E_INVALID;
    int rc = check_arch();
    if (rc) return rc;
    const int nb = pearson_blocks(n);
    double *partial = static_cast<double *>(scratch);
    prof_begin(K_PEARSON_FWD, stream);
    k_pearson_sums<<<nb, CTA, 0, stream>>>((long long)n, src, target, partial);
    k_pearson_finish<<<1, 32, 0, stream>>>((long long)n, nb, partial, stats, out);
    prof_end(K_PEARSON_FWD, stream);
    FSGS_CUDA(cudaGetLastError());
    return FSGS_OK;
}

int fsgs_pearson_backward(int64_t n, const float *src, const float *target, const double *stats, const float *upstream,
                          float *dsrc, float *dtarget, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0 || !src || !target || !stats || (!dsrc && !dtarget)) return FSGS_E_INVALID;
    int rc = check_arch();
    if (rc) return rc;
    prof_begin(K_PEARSON_BWD, stream);
    k_pearson_bwd<<<pearson_blocks(n), CTA, 0, stream>>>((long long)n, src, target, stats, upstream, dsrc, dtarget);
    prof_end(K_PEARSON_BWD, stream);
    FSGS_CUDA(cudaGetLastError());
    return FSGS_OK;
}

// ---- fused local Pearson loss ------------------------------------------------------------------------------
size_t fsgs_local_pearson_scratch_bytes(int32_t n_patches) {
    return align_up((size_t)(n_patches > 0 ? n_patches : 1) * LP_BLOCKS * 5 * sizeof(double), 256);
}

static int local_pearson_args_ok(int32_t H, int32_t W, int32_t box, int32_t n, const void *x0, const void *y0,
                                 const float *src, const float *target) {
    if (H <= 0 || W <= 0 || box <= 0 || box > H || box > W || n <= 0 || n > 65535 || !x0 || !y0 || !src || !target)
        return FSGS_E_INVALID;
    return FSGS_OK;
}

int fsgs_local_pearson_forward(int32_t H, int32_t W, int32_t box, int32_t n_patches, const int64_t *x0, const int64_t *y0,
                               const float *src, const float *target, void *scratch, double *stats, float *out,
                               void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = local_pearson_args_ok(H, W, box, n_patches, x0, y0, src, target);
    if (rc) return rc;
    if (!scratch || !stats || !out) return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    prof_begin(K_LOCAL_PEARSON_FWD, stream);
    k_local_pearson_sums<<<dim3(LP_BLOCKS, n_patches), CTA, 0, stream>>>(
        W, box, reinterpret_cast<const long long *>(x0), reinterpret_cast<const long long *>(y0), src, target,
        static_cast<double *>(scratch));
    k_local_pearson_finish<<<1, CTA, 0, stream>>>(n_patches, box, static_cast<const double *>(scratch), stats, out);
    prof_end(K_LOCAL_PEARSON_FWD, stream);
    FSGS_CUDA(cudaGetLastError());
    return FSGS_OK;
}

int fsgs_local_pearson_backward(int32_t H, int32_t W, int32_t box, int32_t n_patches, const int64_t *x0, const int64_t *y0,
                                const float *src, const float *target, const double *stats, const float *upstream,
                                float *dsrc, float *dtarget, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = local_pearson_args_ok(H, W, box, n_patches, x0, y0, src, target);
    if (rc) return rc;
    if (!stats || (!dsrc && !dtarget)) return FSGS_E_INVALID;
    if ((rc = check_arch())) return rc;
    const size_t bytes = (size_t)H * W * sizeof(float);
    if (dsrc) FSGS_CUDA(cudaMemsetAsync(dsrc, 0, bytes, stream));
    if (dtarget) FSGS_CUDA(cudaMemsetAsync(dtarget, 0, bytes, stream));
    prof_begin(K_LOCAL_PEARSON_BWD, stream);
    k_local_pearson_bwd<<<dim3(LP_BLOCKS, n_patches), CTA, 0, stream>>>(
        W, box, n_patches, reinterpret_cast<const long long *>(x0), reinterpret_cast<const long long *>(y0), src, target,
        stats, upstream, dsrc, dtarget);
    prof_end(K_LOCAL_PEARSON_BWD, stream);
    FSGS_CUDA(cudaGetLastError());
    return FSGS_OK;
}

}  // extern "C"
