"""Fused, drop-in replacement for ``gaussian_renderer.render`` (reference
``gaussian_renderer/__init__.py:49-92``).

``render(viewpoint_camera, index, pc, gs_grad=True, cam_grad=True)`` takes the reference's own
objects (``PoseModel`` / ``GaussianModel``, or the light-weight mirrors in ``fsgs_b200.model``) and
returns the same dict with the same side effects, but performs the whole per-frame path in one
library call: pose transform + activations + SH + projection, ONE tile binning / sort, a six-plane
composite (RGB | depth, silhouette, depth^2) and a backward that emits the parameter gradients
and dL/d(pose) directly (``fsgs_render_forward`` / ``fsgs_render_backward`` in
``include/fsgs_raster.h``).

``render_two_pass`` is the literal two-rasteriser-call formulation of the reference on top of
``GaussianRasterizer`` (same library, un-fused) -- it is what an unmodified
``gaussian_renderer.render`` executes when it imports our ``diff_gaussian_rasterization``.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Dict

import torch

from . import _lib
from .rasterizer import GaussianRasterizer, _Arena, _f32, _on_device, _ptr, _require_cuda, make_settings

_IDENTITY_OK: Dict[tuple, bool] = {}
_TLS = threading.local()
_CAPTURED = []      # (image-state buffer, capacity, W, H, bin capacity) of every fused forward recorded during a graph capture

# Frame-parallel gradient exchange hook (set through fsgs_b200.dist.enable_frame_parallel): a callable that
# sum-all-reduces a flat float32 CUDA tensor in place, ordered on the current stream; None = single GPU.
_GRAD_REDUCER = {"fn": None, "chunks": 1, "alloc": None, "expand": None}
_XCHG_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def set_grad_reducer(fn, chunks: int = 1, alloc=None, expand=None) -> None:
    """``fn(flat)`` must SUM a flat float32 CUDA tensor over the ranks in place, ordered on the current stream.
    ``chunks`` > 1: the fused backward runs its per-Gaussian kernel in that many Gaussian ranges and hands each
    range's 56-byte rows to ``fn`` on a side stream while the next range is computed (``fn`` is then called
    ``chunks`` times per backward, each time under ``torch.cuda.stream(side)``).
    ``alloc(n_floats, device) -> flat float32 tensor``: where the rows live (the NVLink exchange keeps them in a
    symmetric buffer mapped into every rank); default: a fresh tensor per backward.
    ``expand(st, P, first, count, xyz, cam_center, rows, grads, stream)``: replaces ``fn`` + the library's expansion
    by ONE call that sums the rows over the ranks and expands them (the one-shot exchange,
    ``fsgs_compact_grad_expand_peers``); ``grads`` = the dict of gradient tensors to fill."""
    _GRAD_REDUCER["fn"] = fn
    _GRAD_REDUCER["chunks"] = max(1, int(chunks))
    _GRAD_REDUCER["alloc"] = alloc
    _GRAD_REDUCER["expand"] = expand


class FrozenModel:
    """Pose-independent per-Gaussian rows of one Gaussian model (``fsgs_freeze_model``: sigmoid(opacity), Sigma_3D,
    SH colour + clamp mask; 64 B per Gaussian) for the tracking loop, which renders the same model from 50 pose
    estimates per frame (reference train.py:154-210).  ``rows(...)`` returns the buffer, re-evaluating it IN PLACE
    when any of the seven source tensors was replaced or written to since (data pointer + version counter), so a
    captured CUDA graph that reads the buffer sees the refreshed rows (``GraphedStep.replay`` calls ``refresh``)."""

    def __init__(self):
        self.key, self.buf, self.src, self.st, self.ready = None, None, None, None, None
        self.lock = threading.Lock()

    @staticmethod
    def _key(src, st, dev):
        return (tuple((x.data_ptr(), x._version, tuple(x.shape)) for x in src), int(st.sh_degree),
                float(st.scale_modifier), int(st.flags) & ~_lib.FLAG_FIXED_CAPACITY, dev.index)

    def rows(self, src, st, dev):
        """src = (xyz, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, cam_center): float32 contiguous CUDA
        tensors.  Not called while a stream is capturing unless the rows are current."""
        with self.lock:
            key = self._key(src, st, dev)
            capturing = torch.cuda.is_current_stream_capturing()
            if key != self.key:
                if capturing:
                    return None                     # never evaluate inside a capture: the caller takes the plain forward
                P = src[0].shape[0]
                nbytes = int(_lib.lib().fsgs_frozen_bytes(P))
                if self.buf is None or self.buf.numel() != nbytes or self.buf.device != dev:
                    self.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                stream = torch.cuda.current_stream(dev)
                if self.ready is not None:
                    stream.wait_event(self.ready)   # earlier readers / writers of the buffer on other streams
                with _on_device(dev):
                    rc = _lib.lib().fsgs_freeze_model(ctypes.byref(st), P, *[_ptr(x) for x in src], _ptr(self.buf),
                                                      ctypes.c_void_p(stream.cuda_stream))
                _lib.check(rc)
                self.ready = torch.cuda.Event()
                self.ready.record(stream)
                self.key, self.src, self.st = key, tuple(src), st
            elif not capturing and self.ready is not None:
                torch.cuda.current_stream(dev).wait_event(self.ready)
            return self.buf

    def refresh(self):
        """Re-evaluate the rows if their sources changed (eager; for captured graphs that read the buffer)."""
        if self.src is not None:
            self.rows(self.src, self.st, self.src[0].device)


USE_FROZEN_MODEL = True          # tracking against a frozen model renders from FrozenModel rows (bit-identical)
_FROZEN: Dict[int, FrozenModel] = {}
_CAPTURED_FROZEN = []            # FrozenModel objects read by fused forwards recorded during a graph capture


def _frozen_model(dev) -> FrozenModel:
    fm = _FROZEN.get(dev.index)
    if fm is None:
        fm = _FROZEN[dev.index] = FrozenModel()
    return fm


def _exchange_stream(dev) -> "torch.cuda.Stream":
    s = _XCHG_STREAMS.get(dev.index)
    if s is None:
        s = _XCHG_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return s


def _chunk_bounds(P: int, n: int):
    """Split [0, P) into <= n ranges whose starts are multiples of 256 (CTA-aligned for both kernels)."""
    step = -(-P // n)
    step = -(-step // 256) * 256
    return [(a, min(P, a + step)) for a in range(0, P, step)]


class keep_geometry:
    """Test / debugging hook: inside ``with keep_geometry() as g:`` the per-Gaussian projection records of every fused
    forward on this thread are copied out (``g.records``: list of float32 [P,12] tensors, layout in
    csrc/fsgs_device.cuh -- x, y, conic xyz, opacity, r, g, b, view depth, radius bits, tiles).  The parity tests
    feed the view depths to the oracle as sort keys, so that both sides composite near-equal depths in the same
    (float32-rounding dependent) order."""

    def __enter__(self):
        self.records = []
        self._prev = getattr(_TLS, "keep_geom", None)
        _TLS.keep_geom = self.records
        return self

    def __exit__(self, *a):
        _TLS.keep_geom = self._prev
        return False


def _check_identity_view(rs) -> None:
    """The fused path composites the depth planes from the view-space z of the (single) projection,
    which equals the reference's ``get_depth_and_silhouette`` only for the identity rasteriser view
    Free-SurGS always uses (train.py:41, gaussian_model.py:245-246).  Checked once per tensor."""
    vm = rs.viewmatrix
    key = (vm.data_ptr(), vm._version, vm.device.index)
    ok = _IDENTITY_OK.get(key)
    if ok is None:
        ok = bool(torch.equal(vm.reshape(4, 4).float().cpu(), torch.eye(4)))
        _IDENTITY_OK[key] = ok
    if not ok:
        raise NotImplementedError("fused render requires the identity rasteriser view Free-SurGS uses; "
                                  "use render_two_pass / GaussianRasterizer for a general view matrix")


def _check_fused_shapes(P, xyz, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, pose, cam_center) -> None:
    """The fused kernels address the parameter tensors with fixed row widths (3 / 3 / 45 / 1 / 3 / 4 floats per
    Gaussian: SH degree 3 storage as Free-SurGS allocates it, gaussian_model.py:66-67 with max_sh_degree = 3) and
    write gradients of the same widths -- a tensor of another shape would be read and written out of bounds."""
    want = (("_xyz", xyz, (P, 3)), ("_features_dc", f_dc, (P, 1, 3)), ("_features_rest", f_rest, (P, 15, 3)),
            ("_opacity", opacity_raw, (P, 1)), ("_scaling", scaling_raw, (P, 3)), ("_rotation", rotation_raw, (P, 4)))
    for name, t, shape in want:
        if tuple(t.shape) != shape:
            hint = (" (the fused render needs max_sh_degree = 3 storage; use render_two_pass / GaussianRasterizer "
                    "for other SH layouts)") if name == "_features_rest" else ""
            raise ValueError(f"fused render: {name} has shape {tuple(t.shape)}, expected {shape}{hint}")
    if pose.numel() != 16:
        raise ValueError(f"fused render: the pose must be a 4x4 matrix, got shape {tuple(pose.shape)}")
    if cam_center.numel() != 3:
        raise ValueError(f"fused render: cam_center must have 3 elements, got shape {tuple(cam_center.shape)}")


class _RenderFused(torch.autograd.Function):

    @staticmethod
    def forward(ctx, xyz, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, pose, means2D, rs, cam_center,
                active_sh_degree, gs_grad, cam_grad, max_radii2D=None, want_extras=False, stats=None,
                use_frozen=False):
        _require_cuda(xyz)
        dev = xyz.device
        P = xyz.shape[0]
        _check_fused_shapes(P, xyz, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, pose, cam_center)
        H, W = int(rs.image_height), int(rs.image_width)
        st = make_settings(rs, n_coeffs=16, sh_degree=int(active_sh_degree))
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            # CUDA-graph capture (fsgs_b200.graphs): no host read-back of the instance count; the binning buffer is
            # sized by the capacity fsgs_b200.graphs declared for this device
            st.flags |= _lib.FLAG_FIXED_CAPACITY
            st.debug = 0
        t = [_f32(x, dev) for x in (rs.bg, xyz, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, pose, cam_center,
                                    rs.viewmatrix, rs.projmatrix)]
        planes = torch.empty(6, H, W, dtype=torch.float32, device=dev)
        radii = torch.empty(P, dtype=torch.int32, device=dev)     # the projection kernel writes every entry
        stream = torch.cuda.current_stream(dev).cuda_stream
        arena = _Arena(dev, stream)
        nr, nrect = ctypes.c_int64(0), ctypes.c_int64(0)
        # render()'s derived maps come out of the same kernels (fsgs_render_extras): one byte buffer for the
        # three masks, the uncertainty map, and max_radii2D updated in place when it is a float32 tensor
        extras, ex_ptr = (), None
        if want_extras:
            unc = torch.empty(1, H, W, dtype=torch.float32, device=dev)
            masks = torch.empty(2 * H * W + P, dtype=torch.uint8, device=dev)
            ex = _lib.RenderExtras(unc.data_ptr(), masks.data_ptr(), masks.data_ptr() + H * W,
                                   masks.data_ptr() + 2 * H * W if P else None,
                                   max_radii2D.data_ptr() if (max_radii2D is not None and P) else None)
            ex_ptr = ctypes.cast(ctypes.pointer(ex), ctypes.c_void_p)
            mb = masks.view(torch.bool)
            extras = (unc, mb[:H * W].view(H, W), mb[H * W:2 * H * W].view(1, H, W), mb[2 * H * W:])
        rows = None
        if use_frozen and P > 0:
            # pose-independent quantities evaluated once per model state (FrozenModel), 64 B instead of 236 B per Gaussian
            fm = _frozen_model(dev)
            rows = fm.rows((*t[1:7], t[8]), st, dev)
            if rows is not None and capturing:
                _CAPTURED_FROZEN.append(fm)
        with _on_device(dev):
            if rows is not None:
                rc = _lib.lib().fsgs_render_forward_frozen(
                    ctypes.byref(st), P, _ptr(t[0]), _ptr(rows), _ptr(t[7]), _ptr(t[9]), _ptr(t[10]),
                    arena.callback("geom"), None, arena.callback("binning"), None, arena.callback("img"), None,
                    _ptr(planes), _ptr(radii), ctypes.byref(nr), ctypes.byref(nrect), ex_ptr, ctypes.c_void_p(stream))
            else:
                rc = _lib.lib().fsgs_render_forward_ex(
                    ctypes.byref(st), P, *[_ptr(x) for x in t], arena.callback("geom"), None, arena.callback("binning"),
                    None, arena.callback("img"), None, _ptr(planes), _ptr(radii), ctypes.byref(nr), ctypes.byref(nrect),
                    ex_ptr, ctypes.c_void_p(stream))
        _lib.check(rc)
        empty = torch.empty(0, dtype=torch.uint8, device=dev)
        keep = getattr(_TLS, "keep_geom", None)
        if keep is not None and P > 0:
            off = int(_lib.lib().fsgs_geom_record_offset(P))
            keep.append(arena.tensors["geom"][off:off + 48 * P].view(torch.float32).view(P, 12).clone())
        ctx.save_for_backward(*t, *[arena.tensors.get(k, empty) for k in ("geom", "binning", "img")])
        ctx.lease = arena.finish()             # scratch goes back to the workspace pool when this node dies
        ctx.st, ctx.P, ctx.num_rendered, ctx.num_rect, ctx.dev = st, P, int(nr.value), int(nrect.value), dev
        ctx.flags = (bool(gs_grad), bool(cam_grad))
        # densification statistics folded into the backward (xyz_gradient_accum [P,1], denom [P,1]; both float32,
        # contiguous, on this device): updated in place by k_preprocess_fused_bwd
        ctx.stats = None
        if stats is not None and P > 0:
            acc, den = stats
            ok = all(torch.is_tensor(x) and x.dtype == torch.float32 and x.is_contiguous() and x.device == dev
                     and x.numel() == P for x in (acc, den))
            if not ok:
                raise ValueError("densification statistics must be float32 contiguous tensors with one entry per "
                                 "Gaussian on the model's device")
            ctx.stats = (acc, den)
        if capturing:
            # for GraphedStep.overflowed(): image state, instance capacity, per-tile bin capacity of this forward
            _CAPTURED.append((arena.tensors["img"], int(nr.value), W, H,
                              int(_lib.lib().fsgs_fixed_bin_capacity(dev.index))))
            del _CAPTURED[:-64]                                                 # (bounded, whoever does the capturing)
        ctx.mark_non_differentiable(radii, *extras)
        ctx.set_materialize_grads(False)        # an output the loss does not use arrives as None, not as zeros
        _TLS.last_stats = (int(nr.value), int(nrect.value))     # per thread: the viewer thread renders too
        # the largest instance count of ANY fused forward since the caller last reset it (GraphedStep sizes the
        # fixed binning capacity of a captured step from it: a step may render several frames)
        _TLS.max_instances = max(int(nr.value), int(getattr(_TLS, "max_instances", 0)))
        # one output per render() product, all views of the one [6,H,W] buffer the compositor writes: the upstream
        # gradients then arrive per product and go to the library as separate planes (fsgs_render_backward_ex) --
        # no zero-filled [6,H,W] gradient is assembled from slices by autograd
        return (planes[0:3], planes[3], planes[4], planes[5], radii, *extras)

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_sil, g_dsq, _g_radii=None, *_g_extras):
        *t, geom, binning, img = ctx.saved_tensors
        dev = ctx.dev          # (an empty model saves None for its parameter tensors)
        P = ctx.P
        gs_grad, cam_grad = ctx.flags
        # every output is fully overwritten by the library (zeros where a Gaussian is not visible)
        z = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        # The six Gaussian-parameter gradients are carved out of ONE flat buffer (59 floats/Gaussian): autograd
        # adopts the views as .grad, and fsgs_b200.dist.allreduce_gaussian_grads can then sum-all-reduce the whole
        # model gradient with a single collective and no packing copies.
        # (rotation and f_rest first: their float4 / bulk-TMA stores need 16-byte alignment for any P)
        reducer = _GRAD_REDUCER["fn"]
        g = {}
        # Pose tracking against a frozen Gaussian model (no Gaussian parameter and not means2D requires grad):
        # ask the library for dL/dpose only -- it then runs its pose-only backward (no colour / opacity sums in the
        # compositor, no SH reads, no 236 B/Gaussian of gradient writes).
        need = ctx.needs_input_grad
        if cam_grad and need[6] and not any(need[k] for k in (0, 1, 2, 3, 4, 5, 7)) and P > 0 and ctx.stats is None:
            gp = [None if x is None else _f32(x, dev) for x in (g_rgb, g_depth, g_sil, g_dsq)]
            stream = torch.cuda.current_stream(dev).cuda_stream
            arena = _Arena(dev, stream)
            scratch = arena.take("grad_scratch", _lib.lib().fsgs_grad_scratch_bytes(P))
            g_pose = z(4, 4)
            with _on_device(dev):
                rc = _lib.lib().fsgs_render_backward_ex(
                    ctypes.byref(ctx.st), P, ctx.num_rendered, *[_ptr(x) for x in t], _ptr(geom), _ptr(binning),
                    _ptr(img), *[None if x is None else _ptr(x) for x in gp], _ptr(scratch), int(gs_grad), 1,
                    None, None, None, None, None, None, _ptr(g_pose), None, None, ctypes.c_void_p(stream))
            arena.finish().release()
            _lib.check(rc)
            return (None, None, None, None, None, None, g_pose, None, None, None, None, None, None, None, None, None,
                    None)

        def carve(flat, layout):
            off = 0
            for name, shape in layout:
                n = 1
                for s_ in shape:
                    n *= s_
                g[name] = flat[off:off + n].view(*shape)
                off += n

        carve(z(P * 59), (("rotation", (P, 4)), ("f_rest", (P, 15, 3)), ("xyz", (P, 3)), ("f_dc", (P, 1, 3)),
                         ("scaling", (P, 3)), ("opacity", (P, 1))))
        g["pose"], g["means2D"] = z(4, 4), z(P, 3)
        if P == 0:
            g["pose"].zero_()
        if P > 0:
            L = _lib.lib()
            gp = [None if x is None else _f32(x, dev) for x in (g_rgb, g_depth, g_sil, g_dsq)]
            main = torch.cuda.current_stream(dev)
            stream = main.cuda_stream
            arena = _Arena(dev, stream)
            scratch = arena.take("grad_scratch", L.fsgs_grad_scratch_bytes(P))
            acc_p, den_p = (None, None) if ctx.stats is None else (ctx.stats[0].data_ptr(), ctx.stats[1].data_ptr())

            def launch(opts, outs):
                with _on_device(dev):
                    rc = L.fsgs_render_backward_v2(
                        ctypes.byref(ctx.st), P, ctx.num_rendered, *[_ptr(x) for x in t], _ptr(geom), _ptr(binning),
                        _ptr(img), *[None if x is None else _ptr(x) for x in gp], _ptr(scratch), int(gs_grad),
                        int(cam_grad), *outs, _ptr(g["pose"]), _ptr(g["means2D"]), None, ctypes.byref(opts),
                        ctypes.c_void_p(stream))
                _lib.check(rc)

            if reducer is None:
                launch(_lib.BackwardOpts(acc_p, den_p, None, 0, 0, 0, 0),
                       (_ptr(g["xyz"]), _ptr(g["f_dc"]), _ptr(g["f_rest"]), _ptr(g["opacity"]), _ptr(g["scaling"]),
                        _ptr(g["rotation"])))
                arena.finish().release()           # kernels are enqueued; reuse is ordered on this stream
            else:
                # frame-parallel mode (fsgs_b200.dist.enable_frame_parallel): per Gaussian the library emits ONE 56-byte
                # row -- rotation | xyz | scaling | opacity | clamp-masked colour gradient -- instead of the 59 gradient
                # floats; the rows are summed over the ranks and fsgs_compact_grad_expand unpacks them and expands
                # the SH-coefficient gradients from the summed colour gradient.  With chunks > 1 the per-Gaussian
                # kernel runs range by range and range k is exchanged + expanded on a side stream while range k+1
                # is computed on this one.
                compact = z(P * 14) if _GRAD_REDUCER["alloc"] is None else _GRAD_REDUCER["alloc"](P * 14, dev)
                bounds = _chunk_bounds(P, _GRAD_REDUCER["chunks"])
                side = _exchange_stream(dev) if len(bounds) > 1 else None
                none6 = (None,) * 6

                def exchange(a, b, on_stream):
                    if _GRAD_REDUCER["expand"] is not None:    # one-shot: rank sum folded into the expansion kernel
                        _GRAD_REDUCER["expand"](ctx.st, P, a, b - a, t[1], t[8], compact, g, on_stream)
                        return
                    reducer(compact[a * 14:b * 14])            # SUM over the ranks, in place, ordered on the current stream
                    with _on_device(dev):
                        rc = L.fsgs_compact_grad_expand(
                            ctypes.byref(ctx.st), P, a, b - a, _ptr(t[1]), _ptr(t[8]), _ptr(compact), _ptr(g["xyz"]),
                            _ptr(g["f_dc"]), _ptr(g["f_rest"]), _ptr(g["opacity"]), _ptr(g["scaling"]),
                            _ptr(g["rotation"]), ctypes.c_void_p(on_stream))
                    _lib.check(rc)

                for k, (a, b) in enumerate(bounds):
                    launch(_lib.BackwardOpts(acc_p, den_p, compact.data_ptr(), a, b - a, 1 if k else 0, 0), none6)
                    if side is None:
                        exchange(a, b, stream)
                    else:
                        ready = torch.cuda.Event()
                        ready.record(main)
                        with torch.cuda.stream(side):
                            side.wait_event(ready)
                            exchange(a, b, side.cuda_stream)
                arena.finish().release()
                if side is not None:
                    main.wait_stream(side)             # every gradient tensor is complete before autograd hands it on
        return (g["xyz"], g["f_dc"], g["f_rest"], g["opacity"], g["scaling"], g["rotation"],
                g["pose"] if cam_grad else None, g["means2D"], None, None, None, None, None, None, None, None, None)


def render_planes(xyz, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, pose, means2D, raster_settings,
                  cam_center, active_sh_degree, gs_grad=True, cam_grad=True, max_radii2D=None, want_extras=False,
                  densification_stats=None, frozen_model=False):
    """Tensor-level entry: -> ((rgb[3,H,W], depth[H,W], silhouette[H,W], depth_sq[H,W]), radii[P] int32,
    (instances, reference-rectangle instances)); the four images are views of one [6,H,W] buffer.
    With ``want_extras`` a 4th element: (uncertainty[1,H,W], presence_mask[H,W], nan_mask[1,H,W],
    visibility[P], max_radii2D_updated_in_place: bool), produced by the forward kernels.
    ``frozen_model``: the Gaussian parameters are not being optimised (pose tracking): the forward reads the
    pose-independent rows of ``FrozenModel`` instead of the raw parameters -- same result, bit for bit."""
    _check_identity_view(raster_settings)
    mr = max_radii2D
    fuse_mr = (want_extras and mr is not None and mr.dtype == torch.float32 and mr.is_contiguous()
               and mr.device == xyz.device and mr.numel() == xyz.shape[0])
    rgb, depth, sil, dsq, radii, *extras = _RenderFused.apply(xyz, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, pose, means2D,
                                                raster_settings, cam_center, active_sh_degree, gs_grad, cam_grad,
                                                mr if fuse_mr else None, want_extras, densification_stats,
                                                bool(frozen_model) and USE_FROZEN_MODEL)
    if want_extras:
        return (rgb, depth, sil, dsq), radii, getattr(_TLS, "last_stats", (0, 0)), (*extras, fuse_mr)
    return (rgb, depth, sil, dsq), radii, getattr(_TLS, "last_stats", (0, 0))


def _pack_fused(pc, viewmatrix_cur, planes, radius, means2D, extras):
    """render()'s return dict and side effects when the derived maps came out of the kernels."""
    uncertainty, presence, nan_mask, visible, mr_done = extras
    pc.variables['means2D'] = means2D
    if not mr_done:
        mr = pc.variables['max_radii2D']
        torch.maximum(mr, radius.to(mr.dtype), out=mr)
    pc.variables['seen'] = visible
    return {"render": planes[0], "render_dep": planes[1], "render_w2c": viewmatrix_cur, "render_opacity": planes[2],
            "nan_mask": nan_mask, "presence_mask": presence, "uncertainty": uncertainty,
            "viewspace_points": means2D, "visibility_filter": visible, "radii": radius}


def _pack(pc, viewmatrix_cur, im, depth_sil, radius, means2D):
    depth = depth_sil[0, :, :]
    silhouette = depth_sil[1, :, :]
    presence_sil_mask = (silhouette > 0.3)
    depth_sq = depth_sil[2, :, :].unsqueeze(0)
    uncertainty = (depth_sq - depth ** 2).detach()
    pc.variables['means2D'] = means2D
    seen = radius > 0
    # reference: max_radii2D[seen] = max(radius[seen], max_radii2D[seen]) (__init__.py:77-80).  radii are 0
    # where not seen and max_radii2D >= 0, so an in-place element-wise maximum is the same update without
    # the boolean-index gather/scatter (and its host sync).
    mr = pc.variables['max_radii2D']
    torch.maximum(mr, radius.to(mr.dtype), out=mr)
    pc.variables['seen'] = seen
    nan_mask = (~torch.isnan(depth)) & (~torch.isnan(uncertainty))
    return {"render": im, "render_dep": depth, "render_w2c": viewmatrix_cur, "render_opacity": silhouette,
            "nan_mask": nan_mask, "presence_mask": presence_sil_mask, "uncertainty": uncertainty,
            "viewspace_points": means2D, "visibility_filter": radius > 0, "radii": radius}


def _folded_stats(pc, gs_grad):
    """``pc.fold_densification_stats = True`` (opt-in; the reference's objects do not have the attribute): the
    backward of this render adds ||dL/dmeans2D|| and 1 to ``pc.variables['xyz_gradient_accum']`` / ``['denom']`` for
    every visible Gaussian -- GaussianModel.add_densification_stats(viewspace_points, visibility_filter)
    (scene/gaussian_model.py:678-681, train.py:298-303) without the [P,3] gradient round trip.  A caller that
    sets it must not call add_densification_stats for the same render as well."""
    if not (gs_grad and getattr(pc, "fold_densification_stats", False)):
        return None
    return pc.variables['xyz_gradient_accum'], pc.variables['denom']


def render(viewpoint_camera, index, pc, gs_grad=True, cam_grad=True):
    """Same signature, return dict and side effects as the reference's ``render``."""
    xyz = pc.params['_xyz']
    frozen = (not gs_grad) and not any(v.requires_grad for v in pc.params.values())
    if frozen:
        # Pose tracking against a frozen model: the reference's screen-space tensor could only hand its gradient to
        # a temporary nobody holds (it is retained under gs_grad only), so it is created without grad here and the
        # backward takes the library's pose-only path.  With trainable parameters nothing changes.
        means2D = torch.zeros_like(xyz)
    else:
        means2D = torch.zeros_like(xyz, requires_grad=True, device=xyz.device) + 0
    if gs_grad and means2D.requires_grad:        # (under torch.no_grad() there is nothing to retain)
        means2D.retain_grad()
    viewmatrix_cur = viewpoint_camera.get_pose(index)
    pose = viewmatrix_cur if cam_grad else viewmatrix_cur.detach()
    planes, radius, stats, extras = render_planes(
        xyz, pc.params['_features_dc'], pc.params['_features_rest'], pc.params['_opacity'], pc.params['_scaling'],
        pc.params['_rotation'], pose, means2D, pc.cam, viewpoint_camera.cam_center, pc.active_sh_degree,
        gs_grad=gs_grad, cam_grad=cam_grad, max_radii2D=pc.variables.get('max_radii2D'), want_extras=True,
        densification_stats=_folded_stats(pc, gs_grad), frozen_model=frozen)
    out = _pack_fused(pc, viewmatrix_cur, planes, radius, means2D, extras)
    out["num_rendered"] = stats
    return out


def render_two_pass(viewpoint_camera, index, pc, gs_grad=True, cam_grad=True):
    """The reference's own formulation (PyTorch pre-processing + two GaussianRasterizer calls),
    restated here so that tests/bench can run the un-fused drop-in path without /root/reference."""
    from .model import eval_sh, transform_to_frame
    xyz = pc.params['_xyz']
    means2D = torch.zeros_like(xyz, requires_grad=True, device=xyz.device) + 0
    if gs_grad:
        means2D.retain_grad()
    viewmatrix_cur = viewpoint_camera.get_pose(index)
    means_cam = transform_to_frame(xyz, viewmatrix_cur, gs_grad, cam_grad)
    opacity, scales, rots = pc.get_opacity, pc.get_scaling, pc.get_rotation
    feats = pc.get_features
    shs_view = feats.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
    dir_pp = xyz - viewpoint_camera.cam_center.repeat(feats.shape[0], 1)
    dir_pp = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    colors = torch.clamp_min(eval_sh(pc.active_sh_degree, shs_view, dir_pp) + 0.5, 0.0)
    pts4 = torch.cat((means_cam, torch.ones_like(means_cam[:, :1])), dim=-1)
    z = (pc.cam.viewmatrix[0] @ pts4.transpose(0, 1)).transpose(0, 1)[:, 2]
    dcol = torch.stack([z, torch.ones_like(z), z * z], dim=1)
    means2D_b = torch.zeros_like(xyz, requires_grad=True, device=xyz.device) + 0
    im, radius, _ = GaussianRasterizer(raster_settings=pc.cam)(
        means3D=means_cam, shs=None, colors_precomp=colors, rotations=rots, opacities=opacity, scales=scales,
        cov3D_precomp=None, means2D=means2D)
    depth_sil, _, _ = GaussianRasterizer(raster_settings=pc.cam)(
        means3D=means_cam, colors_precomp=dcol, rotations=rots, opacities=opacity, scales=scales, means2D=means2D_b)
    return _pack(pc, viewmatrix_cur, im, depth_sil, radius, means2D)
