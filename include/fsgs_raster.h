/*
 * fsgs_raster.h -- C ABI of the B200-native differentiable Gaussian rasteriser for Free-SurGS.
 *
 * Plain `extern "C"`, raw device pointers + sizes + a CUDA stream handle; no torch types.
 * Every entry point returns 0 on success or a negative FSGS_E_* code (fsgs_error_string()).
 * All pointers are DEVICE pointers unless the parameter name ends in `_host`; tensors are
 * contiguous float32 unless stated.  Work is enqueued on `stream` (a cudaStream_t passed as
 * void*); the only host synchronisation is the one read-back of the tile-instance count in the
 * forward calls (the reference rasteriser has the same one).
 *
 * What each entry point replaces in the reference (wrld/Free-SurGS):
 *
 *   fsgs_rasterize_forward / _backward / fsgs_mark_visible
 *       the pybind functions `_C.rasterize_gaussians`, `_C.rasterize_gaussians_backward`,
 *       `_C.mark_visible` of the un-vendored package `diff_gaussian_rasterization`
 *       (requirements.txt:26, .gitmodules:4-6) that `GaussianRasterizer.forward` calls; reference
 *       call sites gaussian_renderer/__init__.py:68,69,131 and scene/pose_optimizer.py:619-632.
 *
 *   fsgs_render_forward / _backward
 *       the whole per-frame body of `gaussian_renderer.render` (gaussian_renderer/__init__.py:49-92):
 *       transform_to_frame (scene/pose_optimizer.py:960-989), the activations
 *       (scene/gaussian_model.py:118-138), SH->RGB (scene/gaussian_model.py:308-333 +
 *       utils/sh_utils.py:57-112), the depth/silhouette colours (scene/gaussian_model.py:260-291)
 *       and BOTH rasteriser passes (one projection / binning / sort, six composited planes), with
 *       the backward additionally reducing dL/d(pose) = sum_i g_i [p_i;1]^T, i.e. the backward of
 *       the matmul at scene/pose_optimizer.py:987.
 *
 * Scratch memory is requested through caller-supplied allocation callbacks (the analogue of the
 * `resizeFunctional` lambdas the reference's binding passes to its C++ rasteriser) so that the
 * host framework's caching allocator owns every byte; the three buffers must be kept alive by
 * the caller until the matching backward has run.
 */
#ifndef FSGS_RASTER_H_
#define FSGS_RASTER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSGS_ABI_VERSION 1

enum {
    FSGS_OK = 0,
    FSGS_E_INVALID = -1,   /* bad argument (null pointer, negative size, unsupported degree ...) */
    FSGS_E_CUDA = -2,      /* a CUDA runtime call failed; see fsgs_error_string(code)            */
    FSGS_E_ALLOC = -3,     /* an allocation callback returned NULL                               */
    FSGS_E_ARCH = -4,      /* device is not sm_100                                               */
    FSGS_E_WATCHDOG = -5   /* a device-side wait of an earlier launch on this device exceeded its spin
                              budget (reported by the next forward, or by fsgs_watchdog_flag)          */
};

/* Returns a device pointer to at least `bytes` bytes, 256-byte aligned, valid until the caller
 * releases it (after backward).  Called on the calling host thread, synchronously. */
typedef void *(*fsgs_alloc_fn)(void *user, size_t bytes);

/* Static per-call configuration: the scalar fields of GaussianRasterizationSettings
 * (scene/pose_optimizer.py:619-632) plus the SH layout. */
typedef struct fsgs_settings {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    float scale_modifier;
    int32_t sh_degree;   /* active SH degree 0..3 (only read when SH coefficients are given)     */
    int32_t n_coeffs;    /* coefficients per channel stored in `shs` (M of [P,M,3]); 16 for fused */
    int32_t debug;       /* !=0: synchronise + check after every kernel, enable device watchdog   */
    int32_t flags;       /* FSGS_FLAG_*                                                           */
} fsgs_settings;

#define FSGS_FLAG_NO_TMA 1u        /* stage tile batches with plain loads instead of bulk TMA     */
#define FSGS_FLAG_NO_TILE_CULL 2u  /* keep every tile of the reference's 3-sigma rectangle        */
#define FSGS_FLAG_RESERVED_4 4u    /* (was: first backward formulation, removed; ignored)          */
#define FSGS_FLAG_NO_OPTIMISTIC 8u /* forward: always wait for the instance count before binning   */
#define FSGS_FLAG_SORT_NETWORK 16u /* per-tile sort: always the compare-exchange network (A/B, tests) */
#define FSGS_FLAG_UPSTREAM_STYLE 512u /* BASELINE for bench.py, API flavour only: compositors in the published rasteriser's
                                         kernel structure (one thread per pixel over the whole list, per-pair scalar atomics);
                                         set FSGS_FLAG_NO_TILE_CULL with it for the reference's full 3-sigma rectangles */
#define FSGS_FLAG_NO_BINS 256u     /* forward: never bin the keys in the counting pass, always the scatter pass (A/B, tests) */
#define FSGS_FLAG_SORT_WINDOW_LARGE 128u /* per-tile sort: always the 64 KB shared-memory window (A/B, tests)  */
#define FSGS_FLAG_NO_POSE_ONLY 64u /* fused backward: never take the pose-only specialisation (A/B, tests) */
#define FSGS_FLAG_FIXED_CAPACITY 32u /* forward: no host read-back of the instance count (CUDA-graph capture);
                                        the binning buffer is sized by fsgs_set_instance_capacity()     */

/* ------------------------------------------------------------------------------------------
 * API-level rasteriser (one GaussianRasterizer call).
 *   bg[3], viewmatrix[16], projmatrix[16] (column-major, i.e. the transposed row-major matrices the
 *   reference passes), campos[3]: device pointers.
 *   Exactly one of {colors_precomp[P,3], shs[P,n_coeffs,3]} and exactly one of
 *   {(scales[P,3], rotations[P,4]), cov3D_precomp[P,6]} must be non-NULL.
 *   out_color[3,H,W], out_depth[1,H,W], radii[P] (int32).  *num_rendered_host receives the
 *   number of (tile, Gaussian) instances actually composited (after exact tile culling);
 *   *num_rect_host (optional) the reference's 3-sigma-rectangle count.
 * ------------------------------------------------------------------------------------------ */
int fsgs_rasterize_forward(const fsgs_settings *st, int32_t P, const float *bg, const float *means3D,
                           const float *colors_precomp, const float *shs, const float *opacities,
                           const float *scales, const float *rotations, const float *cov3D_precomp,
                           const float *viewmatrix, const float *projmatrix, const float *campos,
                           fsgs_alloc_fn geom_alloc, void *geom_user, fsgs_alloc_fn binning_alloc,
                           void *binning_user, fsgs_alloc_fn img_alloc, void *img_user, float *out_color,
                           float *out_depth, int32_t *radii, int64_t *num_rendered_host,
                           int64_t *num_rect_host, void *stream);

/* Backward of the call above.  geom/binning/img are the buffers the forward obtained from the
 * callbacks.  dL_dout_depth may be NULL (treated as zero).  Outputs (all overwritten; a NULL
 * output is skipped): dL_dmeans2D[P,3] (z = 0), dL_dcolors[P,3], dL_dopacity[P,1],
 * dL_dmeans3D[P,3], dL_dcov3D[P,6], dL_dsh[P,n_coeffs,3], dL_dscales[P,3], dL_drotations[P,4].
 * `grad_scratch` is a caller-provided device buffer of fsgs_grad_scratch_bytes(P) bytes. */
int fsgs_rasterize_backward(const fsgs_settings *st, int32_t P, int64_t num_rendered, const float *bg,
                            const float *means3D, const float *colors_precomp, const float *shs,
                            const float *opacities, const float *scales, const float *rotations,
                            const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
                            const float *campos, const void *geom, const void *binning, const void *img,
                            const float *dL_dout_color, const float *dL_dout_depth, void *grad_scratch,
                            float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity, float *dL_dmeans3D,
                            float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations,
                            void *stream);

/* visible[P] (uint8): view-space z > 0.2 (the reference's frustum test). */
int fsgs_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                      uint8_t *visible, void *stream);

/* ------------------------------------------------------------------------------------------
 * Fused per-frame render (gaussian_renderer.render).  Raw (pre-activation) parameters:
 *   xyz[P,3] world, features_dc[P,1,3], features_rest[P,15,3], opacity_raw[P,1] (sigmoid),
 *   scaling_raw[P,3] (exp), rotation_raw[P,4] (L2-normalised), pose[4,4] ROW-major world->camera
 *   (LearnPose.forward output), cam_center[3] (SH view origin; the reference keeps it at the
 *   first frame's centre).  viewmatrix/projmatrix/bg as above (identity view in Free-SurGS).
 *   out_planes[6,H,W] = RGB | depth, silhouette, depth^2 (each + T_final * bg as the reference's
 *   second pass does).  radii[P] int32.
 * ------------------------------------------------------------------------------------------ */
int fsgs_render_forward(const fsgs_settings *st, int32_t P, const float *bg, const float *xyz,
                        const float *features_dc, const float *features_rest, const float *opacity_raw,
                        const float *scaling_raw, const float *rotation_raw, const float *pose,
                        const float *cam_center, const float *viewmatrix, const float *projmatrix,
                        fsgs_alloc_fn geom_alloc, void *geom_user, fsgs_alloc_fn binning_alloc,
                        void *binning_user, fsgs_alloc_fn img_alloc, void *img_user, float *out_planes,
                        int32_t *radii, int64_t *num_rendered_host, int64_t *num_rect_host, void *stream);

/* The same forward, additionally producing the derived outputs of the reference's render()
 * (gaussian_renderer/__init__.py:70-88) from inside the kernels instead of a dozen element-wise passes:
 *   uncertainty[1,H,W]   = depth^2 plane - (depth plane)^2                       (:73-75, detached)
 *   presence_mask[H,W]   = silhouette plane > 0.3                                (:72)      uint8 0/1
 *   nan_mask[1,H,W]      = !isnan(depth plane) & !isnan(uncertainty)             (:81)      uint8 0/1
 *   visibility[P]        = radii > 0                                             (:77,:88)  uint8 0/1
 *   max_radii2D[P]       = max(max_radii2D, radii), updated IN PLACE             (:78)      float32
 * Any pointer may be NULL (that output is skipped); extras == NULL is fsgs_render_forward. */
typedef struct fsgs_render_extras {
    float *uncertainty;
    uint8_t *presence_mask;
    uint8_t *nan_mask;
    uint8_t *visibility;
    float *max_radii2D;
} fsgs_render_extras;
int fsgs_render_forward_ex(const fsgs_settings *st, int32_t P, const float *bg, const float *xyz,
                           const float *features_dc, const float *features_rest, const float *opacity_raw,
                           const float *scaling_raw, const float *rotation_raw, const float *pose,
                           const float *cam_center, const float *viewmatrix, const float *projmatrix,
                           fsgs_alloc_fn geom_alloc, void *geom_user, fsgs_alloc_fn binning_alloc,
                           void *binning_user, fsgs_alloc_fn img_alloc, void *img_user, float *out_planes,
                           int32_t *radii, int64_t *num_rendered_host, int64_t *num_rect_host,
                           const fsgs_render_extras *extras, void *stream);

/* Frozen-model forward: the fused render for loops that render ONE Gaussian model from many poses -- Free-SurGS'
 * tracking (train.py:154-210: 50 iterations per frame that only move the pose; gaussian_renderer/__init__.py:49-92 is
 * called with the same parameters every time).  With the reference's quirks (identity rasteriser view, quaternions not
 * rotated into the camera frame, SH view direction = world position - a frozen camera centre; pose_optimizer.py:603,
 * gaussian_model.py:308-333) everything about a Gaussian except its camera-frame mean is independent of the pose.
 *   fsgs_freeze_model           evaluates sigmoid(opacity), Sigma_3D = R S^2 R^T and the SH colour + clamp mask once
 *                               into `frozen` (fsgs_frozen_bytes(P) = 64 B per Gaussian, 16-byte aligned device memory);
 *   fsgs_render_forward_frozen  = fsgs_render_forward_ex reading 64 B instead of 236 B per Gaussian.  Same outputs, bit
 *                               for bit; the buffers it hands out feed the same fsgs_render_backward* calls (which
 *                               still take the raw parameters).
 * The caller owns the invalidation: re-freeze after any change of the parameters, cam_center, st->sh_degree or
 * st->scale_modifier. */
size_t fsgs_frozen_bytes(int32_t P);
int fsgs_freeze_model(const fsgs_settings *st, int32_t P, const float *xyz, const float *features_dc,
                      const float *features_rest, const float *opacity_raw, const float *scaling_raw,
                      const float *rotation_raw, const float *cam_center, void *frozen, void *stream);
int fsgs_render_forward_frozen(const fsgs_settings *st, int32_t P, const float *bg, const void *frozen,
                               const float *pose, const float *viewmatrix, const float *projmatrix,
                               fsgs_alloc_fn geom_alloc, void *geom_user, fsgs_alloc_fn binning_alloc,
                               void *binning_user, fsgs_alloc_fn img_alloc, void *img_user, float *out_planes,
                               int32_t *radii, int64_t *num_rendered_host, int64_t *num_rect_host,
                               const fsgs_render_extras *extras, void *stream);

/* Backward of the fused render.  dL_dplanes[6,H,W].  Outputs (overwritten; NULL = skip):
 * dL_dxyz[P,3], dL_dfeatures_dc[P,1,3], dL_dfeatures_rest[P,15,3], dL_dopacity_raw[P,1],
 * dL_dscaling_raw[P,3], dL_drotation_raw[P,4], dL_dpose[4,4] (row 3 = 0),
 * dL_dmeans2D[P,3] (screen-space gradient of the RGB planes only, as the reference's
 * `viewspace_points.grad`).  gs_grad / cam_grad mirror transform_to_frame's detach flags:
 * gs_grad=0 drops the pose path from dL_dxyz (the SH view-direction path stays, as in the
 * reference); cam_grad=0 leaves dL_dpose zero.
 * Pose-only request: when every per-Gaussian output pointer is NULL and only dL_dpose is asked for (pose tracking
 * against a frozen Gaussian model) the backward runs a specialisation that skips the colour / opacity columns in the
 * compositor and neither reads the SH coefficients nor writes per-Gaussian gradients; dL_dpose is the same sum
 * (different float summation order only).  FSGS_FLAG_NO_POSE_ONLY forces the general path. */
int fsgs_render_backward(const fsgs_settings *st, int32_t P, int64_t num_rendered, const float *bg,
                         const float *xyz, const float *features_dc, const float *features_rest,
                         const float *opacity_raw, const float *scaling_raw, const float *rotation_raw,
                         const float *pose, const float *cam_center, const float *viewmatrix,
                         const float *projmatrix, const void *geom, const void *binning, const void *img,
                         const float *dL_dplanes, void *grad_scratch, int32_t gs_grad, int32_t cam_grad,
                         float *dL_dxyz, float *dL_dfeatures_dc, float *dL_dfeatures_rest,
                         float *dL_dopacity_raw, float *dL_dscaling_raw, float *dL_drotation_raw,
                         float *dL_dpose, float *dL_dmeans2D, void *stream);

/* The same backward with the upstream gradient given per output of render() instead of as one packed
 * [6,H,W] array: dL_drgb[3,H,W] ("render"), dL_ddepth[H,W] ("render_dep"), dL_dsil[H,W]
 * ("render_opacity"), dL_ddepth_sq[H,W] (the depth^2 plane; detached in the reference, normally NULL).
 * A NULL plane is all zeros and is neither read nor materialised. */
int fsgs_render_backward_ex(const fsgs_settings *st, int32_t P, int64_t num_rendered, const float *bg,
                            const float *xyz, const float *features_dc, const float *features_rest,
                            const float *opacity_raw, const float *scaling_raw, const float *rotation_raw,
                            const float *pose, const float *cam_center, const float *viewmatrix,
                            const float *projmatrix, const void *geom, const void *binning, const void *img,
                            const float *dL_drgb, const float *dL_ddepth, const float *dL_dsil,
                            const float *dL_ddepth_sq, void *grad_scratch, int32_t gs_grad, int32_t cam_grad,
                            float *dL_dxyz, float *dL_dfeatures_dc, float *dL_dfeatures_rest,
                            float *dL_dopacity_raw, float *dL_dscaling_raw, float *dL_drotation_raw,
                            float *dL_dpose, float *dL_dmeans2D, float *dL_dsh_rgb, void *stream);

/* Options of fsgs_render_backward_v2 (all optional; a zeroed struct or NULL = fsgs_render_backward_ex).
 *   xyz_gradient_accum, denom [P,1]: GaussianModel.add_densification_stats (scene/gaussian_model.py:678-681, called
 *       from train.py:298-303 after every mapping backward) folded into the per-Gaussian backward kernel:
 *       accum[i] += ||dL/dmeans2D_i||_2 and denom[i] += 1 for every Gaussian the frame saw (radius > 0).  Both or none.
 *   first, count: restrict the per-Gaussian kernel to Gaussians [first, first+count) (first % 4 == 0; count <= 0 = all).
 *   skip_composite: the accumulator in grad_scratch already holds this frame's compositor output (an earlier call
 *       with the same scratch buffer ran it): only the per-Gaussian kernel runs, dL_dpose is accumulated into, not
 *       zeroed.  Lets a caller run the backward range by range and exchange range k over the GPUs while range k+1
 *       is computed (frame-parallel mode).
 *   compact [P,14]: rotation(4) | xyz(3) | scaling(3) | opacity(1) | clamp-masked colour gradient(3) written as one
 *       56-byte row per Gaussian instead of the separate dL_dxyz / dL_dopacity_raw / dL_dscaling_raw /
 *       dL_drotation_raw / dL_dsh_rgb outputs (which must then be NULL): the exchange unit of the frame-parallel
 *       mode; fsgs_compact_grad_expand turns summed rows into the per-parameter gradients. */
typedef struct fsgs_backward_opts {
    float *xyz_gradient_accum;
    float *denom;
    float *compact;
    int32_t first;
    int32_t count;
    int32_t skip_composite;
    int32_t reserved;
} fsgs_backward_opts;

int fsgs_render_backward_v2(const fsgs_settings *st, int32_t P, int64_t num_rendered, const float *bg,
                            const float *xyz, const float *features_dc, const float *features_rest,
                            const float *opacity_raw, const float *scaling_raw, const float *rotation_raw,
                            const float *pose, const float *cam_center, const float *viewmatrix,
                            const float *projmatrix, const void *geom, const void *binning, const void *img,
                            const float *dL_drgb, const float *dL_ddepth, const float *dL_dsil,
                            const float *dL_ddepth_sq, void *grad_scratch, int32_t gs_grad, int32_t cam_grad,
                            float *dL_dxyz, float *dL_dfeatures_dc, float *dL_dfeatures_rest,
                            float *dL_dopacity_raw, float *dL_dscaling_raw, float *dL_drotation_raw,
                            float *dL_dpose, float *dL_dmeans2D, float *dL_dsh_rgb,
                            const fsgs_backward_opts *opts, void *stream);

/* Device watchdog.  Every device-side wait in the library is bounded; one that gives up sets a sticky per-device
 * word instead of hanging the GPU.  Each (non-captured) forward reads the word back together with the instance
 * count and returns FSGS_E_WATCHDOG if an EARLIER launch on the device set it.  This call reads it on demand
 * (synchronises the device): returns 1 if set, 0 if clear, a negative FSGS_E_* code on failure; `reset` clears it. */
int fsgs_watchdog_flag(int32_t device, int32_t reset);

/* CUDA-graph capture support.  A forward normally reads the number of (tile, Gaussian) instances back to size
 * the binning buffer -- a host synchronisation that cannot be captured.  With FSGS_FLAG_FIXED_CAPACITY in
 * st->flags the forward sizes that buffer for `capacity` instances (declared here, per device), touches
 * the host nowhere (st->debug must be 0) and reports *num_rendered_host = capacity; pass that value on to the
 * backward.  If a frame has more instances than the capacity, the binning / compositing kernels of both
 * directions return immediately (outputs undefined) and counters[0] (see fsgs_img_offsets) > capacity tells
 * the caller to re-capture with a larger capacity. */
int fsgs_set_instance_capacity(int32_t device, int64_t capacity);
/* Per-tile bin capacity (keys) the last FSGS_FLAG_FIXED_CAPACITY forward on the device ran with; 0 = it took the scatter
 * path.  A replayed frame whose longest tile list (counters[2], see fsgs_img_offsets) exceeds it skipped its binning /
 * compositing kernels exactly as when counters[0] > capacity: re-capture after an eager frame has refreshed the hints. */
int64_t fsgs_fixed_bin_capacity(int32_t device);

/* Frame-parallel gradient exchange (one frame per GPU, shared Gaussian model).
 * Every SH-coefficient gradient of a Gaussian is  basis_k(dir) * gc[ch]  with gc = the colour gradient after
 * the clamp mask, and in Free-SurGS neither dir = normalize(xyz - cam_center) nor the mask depends on the frame
 * (the SH view origin is frozen, gaussian_model.py:317 / pose_optimizer.py:603).  So instead of all-reducing
 * 48 SH-gradient floats per Gaussian the ranks reduce the 3 floats of gc:
 *   1. fsgs_render_backward_ex with dL_dfeatures_dc = dL_dfeatures_rest = NULL and dL_dsh_rgb[P,3] set:
 *      the other gradients as usual (dL_dxyz includes the SH view-direction term, which is linear in gc);
 *   2. the caller sum-all-reduces xyz | opacity | scaling | rotation | gc  (14 floats = 56 B per Gaussian
 *      instead of 59 floats = 236 B);
 *   3. fsgs_sh_grad_expand turns the reduced gc into dL_dfeatures_dc[P,1,3] and dL_dfeatures_rest[P,15,3]
 *      (st->sh_degree = active degree; coefficients above it get zeros). */
int fsgs_sh_grad_expand(const fsgs_settings *st, int32_t P, const float *xyz, const float *cam_center,
                        const float *dL_dsh_rgb, float *dL_dfeatures_dc, float *dL_dfeatures_rest, void *stream);

/* Frame-parallel exchange over NVLink / NVSwitch, hand-written (two-shot all-reduce; replaces ncclAllReduce on this
 * path).  The caller keeps the 56-byte rows of fsgs_backward_opts.compact in a SYMMETRIC buffer: the same allocation
 * on every GPU, every copy mapped into every process (peer_ptrs_host[r] = rank r's copy as seen from this process,
 * r = 0..world-1, host array of device pointers) and, where the fabric supports it, bound to one multicast address
 * (multicast_ptr, else NULL).  Between two cross-GPU barriers of the caller's (1: every rank's rows are written,
 * 2: every slice is summed), this call sums THIS rank's 1/world slice of float4 words [first_vec4, first_vec4 + n_vec4)
 * over all copies and writes the sum back into all of them -- multimem.ld_reduce / multimem.st through the switch
 * when multicast_ptr is given, peer loads and stores otherwise.  Every word is summed once, by its owner: all ranks end
 * up with bit-identical sums.  world <= 8. */
int fsgs_exchange_rows(void *multicast_ptr, void *const *peer_ptrs_host, int32_t world, int32_t rank, int64_t first_vec4,
                       int64_t n_vec4, void *stream);

/* The same for the 56-byte rows of fsgs_backward_opts.compact, one Gaussian range [first, first+count) at a time
 * (first % 4 == 0): unpacks rotation / xyz / scaling / opacity from the (rank-summed) rows into their gradient tensors
 * and expands the colour gradient into the SH-coefficient gradients. */
/* One-shot form of the exchange for small rank counts: the rank sum is folded into the expansion kernel -- collective
 * and consumer in one kernel over peer memory.  row_ptrs_host[world]: every rank's [P,14] row buffer as mapped into
 * THIS process (this rank's own among them, in rank order, 16-byte aligned; e.g. the buffer_ptrs of a symmetric
 * allocation).  Rows [first, first+count) (first % 256 == 0) of all ranks are summed in rank order -- bit-identical on
 * every rank -- and expanded like fsgs_compact_grad_expand.  The caller orders it: all rows written (a cross-GPU
 * barrier) before, all ranks done reading before any row buffer is written again.  Per rank (N-1) x the rows cross the
 * links (two-shot: 2 (N-1)/N x): the better choice at N = 2. */
int fsgs_compact_grad_expand_peers(const fsgs_settings *st, int32_t P, int32_t first, int32_t count, const float *xyz,
                                   const float *cam_center, void *const *row_ptrs_host, int32_t world,
                                   int32_t owner_slices, int64_t slice_first_vec4, int64_t slice_n_vec4, float *dL_dxyz,
                                   float *dL_dfeatures_dc, float *dL_dfeatures_rest, float *dL_dopacity_raw,
                                   float *dL_dscaling_raw, float *dL_drotation_raw, void *stream);
/* Pull-gather form for 4+ ranks = two-shot with its second half riding on the expansion:
 *   fsgs_exchange_rows_scatter(..., scatter_offset_vec4 > 0)  reduce-scatter only: rank g sums the g-th 1/N slice of the
 *       rows over all ranks and stores the sums into ITS OWN buffer, scatter_offset_vec4 float4 behind the rows
 *       (peer_ptrs_host is required; scatter_offset_vec4 == 0 is fsgs_exchange_rows);
 *   fsgs_compact_grad_expand_peers(..., owner_slices = 1, slice_first_vec4, slice_n_vec4) with row_ptrs_host[r] = rank
 *       r's SUM region: every 16-byte word is fetched from the rank that owns its slice (same slice arithmetic) while
 *       the SH gradients are written.
 * Order: barrier (rows written) - scatter - barrier (sums written) - expand.  No trailing barrier: the next step
 * writes rows, not sums, and reaches its own scatter only behind the next first barrier. */
int fsgs_exchange_rows_scatter(void *multicast_ptr, void *const *peer_ptrs_host, int32_t world, int32_t rank,
                               int64_t first_vec4, int64_t n_vec4, int64_t scatter_offset_vec4, void *stream);
int fsgs_compact_grad_expand(const fsgs_settings *st, int32_t P, int32_t first, int32_t count, const float *xyz,
                             const float *cam_center, const float *compact, float *dL_dxyz, float *dL_dfeatures_dc,
                             float *dL_dfeatures_rest, float *dL_dopacity_raw, float *dL_dscaling_raw,
                             float *dL_drotation_raw, void *stream);

/* Per-frame pose, LearnPose.forward (scene/pose_optimizer.py:822-877): r[1,4,N] raw quaternion
 * (w,x,y,z) and t[3,N] as the reference stores them; column `cam` -> Rt[4,4] row-major
 * (F.normalize, q2rot with its own normalisation, [[R,t],[0,0,0,1]]).  The backward turns dL/dRt
 * into dL/dr[1,4,N], dL/dt[3,N] (zero outside column `cam`). */
int fsgs_pose_forward(const float *r, const float *t, int32_t cam, int32_t n_cams, float *Rt, void *stream);
int fsgs_pose_backward(const float *r, int32_t cam, int32_t n_cams, const float *dRt, float *dr, float *dt,
                       void *stream);

/* Sizes / layout helpers (host only, no CUDA calls). */
size_t fsgs_geom_bytes(int32_t P);
size_t fsgs_img_bytes(int32_t image_width, int32_t image_height);
size_t fsgs_binning_bytes(int64_t num_rendered);
size_t fsgs_grad_scratch_bytes(int32_t P);
/* Byte offset of the packed per-Gaussian splat records (48 B each: x, y, conic.xyz, opacity,
 * r, g, b, depth, radius-as-int, tiles-touched-as-int) inside the geometry buffer. */
size_t fsgs_geom_record_offset(int32_t P);
/* Documented buffer layouts (for debugging / tests that read the scratch buffers back):
 *   img     : out[0] final_T f32[HW], out[1] n_contrib u32[HW], out[2] tile_count u32[T],
 *             out[3] tile_offset u32[T+1] (exclusive scan; [T] = instance count), out[4] cursor u32[T],
 *             out[5] counters u64[4] (instances, rect instances, longest list, error flag)
 *   binning : out[1] = 0: depth-sorted per-instance splat records, 48 B x R, instance i of tile t at
 *             tile_offset[t] + i: (x, y, a2, b2 | c2, opacity, r, g | b, depth, 8x4-block mask, Gaussian id)
 *             with the conic pre-scaled for a base-2 exponent; inside each tile's segment the records
 *             ascend by (depth float bits, Gaussian id).  out[0]: sort-key scratch ((depth bits << 32) | id),
 *             contents unspecified after the forward.  The forward may request MORE than
 *             fsgs_binning_bytes(R) from the allocation callback (it sizes the buffer before R is
 *             known on the host, see FSGS_FLAG_NO_OPTIMISTIC) and may call the callback twice; the
 *             buffer returned LAST is the one to keep for the backward. */
void fsgs_img_offsets(int32_t image_width, int32_t image_height, size_t *out6);
void fsgs_binning_offsets(int64_t num_rendered, size_t *out2);

/* Fused image loss of Free-SurGS (reference utils/loss_utils.py:47-96, rgb_loss_func):
 *   loss = (1 - lambda_dssim) * mean|x - y| + lambda_dssim * (1 - mean SSIM(x, y)),  x = img * mask, y = gt * mask,
 * SSIM with the reference's 11x11 Gaussian window (sigma 1.5), zero padding, C1 = 0.01^2, C2 = 0.03^2.
 * img, gt: float32 [C,H,W] device pointers.  Mask (optional): mask_u8 (one byte per element, non-zero = 1) or
 * mask_f32, at most one non-NULL; mask_cstride = 0 for one [H,W] plane shared by the channels, H*W for [C,H,W].
 * forward : out[3] = (loss, mean|x-y|, mean SSIM) (device); maps[3,C,H,W] (device, NULL = no gradient wanted)
 *           receives the per-pixel partial derivatives the backward needs; scratch = fsgs_rgb_loss_scratch_bytes().
 * backward: dimg[C,H,W] = d(upstream * loss)/d(img); `upstream` is a DEVICE scalar (NULL = 1) so that no host
 *           synchronisation is needed between the caller's loss arithmetic and this call.
 * One forward kernel + a one-CTA reduction, one backward kernel; the sums are taken in a fixed order
 * (deterministic). */
size_t fsgs_rgb_loss_scratch_bytes(int32_t C, int32_t H, int32_t W);
int fsgs_rgb_loss_forward(int32_t C, int32_t H, int32_t W, const float *img, const float *gt,
                          const unsigned char *mask_u8, const float *mask_f32, int64_t mask_cstride,
                          float lambda_dssim, float *maps, void *scratch, float *out, void *stream);
int fsgs_rgb_loss_backward(int32_t C, int32_t H, int32_t W, const float *img, const float *gt,
                           const unsigned char *mask_u8, const float *mask_f32, int64_t mask_cstride,
                           float lambda_dssim, const float *maps, const float *upstream, float *dimg, void *stream);

/* Fused Pearson depth loss (reference utils/loss_utils.py:98-109, pearson_depth_loss):
 *   loss = 1 - mean( (x - mean x)/(std x + 1e-6) * (y - mean y)/(std y + 1e-6) ),  unbiased std, over n elements.
 * forward : out[1] (device float) = loss; stats[8] (device doubles) = the statistics the backward needs;
 *           scratch = fsgs_pearson_scratch_bytes().
 * backward: dsrc / dtarget [n] (either may be NULL) = d(upstream * loss)/d(src | target); upstream is a DEVICE
 *           scalar (NULL = 1).  Sums are taken in double in a fixed order (deterministic). */
size_t fsgs_pearson_scratch_bytes(void);
int fsgs_pearson_forward(int64_t n, const float *src, const float *target, void *scratch, double *stats, float *out,
                         void *stream);
int fsgs_pearson_backward(int64_t n, const float *src, const float *target, const double *stats, const float *upstream,
                          float *dsrc, float *dtarget, void *stream);

/* local_pearson_loss (utils/loss_utils.py:112-127; train.py:257 calls it with box 128, p_corr 0.5 on every mapping
 * view): the mean over n_patches box x box patches of pearson_depth_loss(src[patch], target[patch]).  x0 / y0
 * [n_patches] (int64, device): top-left ROW / COLUMN of each patch -- drawn by the caller with the reference's two
 * torch.randint calls so that the patches are the reference's.  src / target [H,W]; scratch of
 * fsgs_local_pearson_scratch_bytes(n_patches) bytes; stats [6 * n_patches] doubles (kept for the backward); out [1].
 * The backward zero-fills and accumulates dsrc / dtarget [H,W] (either may be NULL); upstream [1] or NULL (= 1). */
size_t fsgs_local_pearson_scratch_bytes(int32_t n_patches);
int fsgs_local_pearson_forward(int32_t H, int32_t W, int32_t box, int32_t n_patches, const int64_t *x0,
                               const int64_t *y0, const float *src, const float *target, void *scratch,
                               double *stats, float *out, void *stream);
int fsgs_local_pearson_backward(int32_t H, int32_t W, int32_t box, int32_t n_patches, const int64_t *x0,
                                const int64_t *y0, const float *src, const float *target, const double *stats,
                                const float *upstream, float *dsrc, float *dtarget, void *stream);

/* Optional per-kernel timing with CUDA events on the launching stream (single-threaded use; off
 * by default).  fsgs_profile_collect synchronises the device and returns, per kernel in the
 * order of fsgs_kernel_names(), the summed duration in ms and the launch count since enable. */
int fsgs_profile_enable(int on);
int fsgs_profile_collect(double *ms_sum, int64_t *count, int n);

int fsgs_abi_version(void);
const char *fsgs_error_string(int code);
/* Name of the kernels a forward+backward launches, comma separated (for launch accounting). */
const char *fsgs_kernel_names(void);

#ifdef __cplusplus
}
#endif
#endif /* FSGS_RASTER_H_ */
