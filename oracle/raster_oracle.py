"""ORACLE (test infrastructure, never a product path).

CPU restatement of the tile rasteriser that Free-SurGS calls through
``diff_gaussian_rasterization.GaussianRasterizer`` (call sites: reference
``gaussian_renderer/__init__.py:68,69,131``; settings built at
``scene/pose_optimizer.py:619-632``).

PARITY UNPINNED: the rasteriser is a third-party dependency that is absent from
/root/reference (``requirements.txt:26`` -> ingra14m/depth-diff-gaussian-rasterization, HEAD,
no pin; ``.gitmodules:4-6`` submodule directory missing).  This file restates the *published*
algorithm of that package.  The Python half of the path
(SH evaluation, pose, camera, depth/silhouette colours) IS pinned against the reference's own
code through ``tests/golden/ref_python_half.npz`` (see ``oracle/make_golden_ref_python.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.

Formulation: generic in dtype (float64 = truth, float32 = same-precision restatement).
Decisions (culling, integer radii, tile rectangles, per-tile order, alpha / transmittance
thresholds) are taken outside autograd; every continuous quantity is a torch expression, so the
backward is obtained from ``torch.autograd`` instead of being transcribed.  Three places where
the upstream closed-form backward is deliberately *not* the true derivative are mirrored with
``detach`` so that the oracle states the reference's semantics:
  (1) alpha = min(0.99, o*G): gradient passes as if un-clamped (straight-through);
  (2) the +-1.3*tanfov clamp of t.x/t.z: a clamped coordinate is treated as a constant;
  (3) integer radius / tile rect / thresholds carry no gradient.
(The upstream ``1/(det^2+1e-7)`` regulariser in the conic backward is a <=1.2e-5 relative
deviation from autograd and is restated exactly in ``oracle/raster_oracle.c``.)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

BLOCK = 16
ALPHA_MIN = 1.0 / 255.0
T_MIN = 1e-4
NEAR_CULL = 0.2

# --- real SH constants, restating reference utils/sh_utils.py:26-54 -----------------------
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """sh [..., C, K], dirs [..., 3] -> [..., C].  Restates utils/sh_utils.py:57-112 (deg<=3)."""
    assert 0 <= deg <= 3 and sh.shape[-1] >= (deg + 1) ** 2
    result = SH_C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - SH_C1 * y * sh[..., 1] + SH_C1 * z * sh[..., 2] - SH_C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + SH_C2[0] * xy * sh[..., 4] + SH_C2[1] * yz * sh[..., 5]
                      + SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                      + SH_C2[3] * xz * sh[..., 7] + SH_C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + SH_C3[0] * y * (3 * xx - yy) * sh[..., 9]
                          + SH_C3[1] * xy * z * sh[..., 10]
                          + SH_C3[2] * y * (4 * zz - xx - yy) * sh[..., 11]
                          + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + SH_C3[4] * x * (4 * zz - xx - yy) * sh[..., 13]
                          + SH_C3[5] * z * (xx - yy) * sh[..., 14]
                          + SH_C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """[P,4] (w,x,y,z) used AS GIVEN (no normalisation) -> [P,3,3].
    Same matrix as reference utils/general_utils.py:216-224 (which normalises first)."""
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(-1, 3, 3)


def build_cov3d(scales: torch.Tensor, rots: torch.Tensor, mod: float) -> torch.Tensor:
    """Sigma = R diag((mod*s)^2) R^T -> [P,3,3] (Appendix A, K1)."""
    R = quat_to_rotmat(rots)
    S2 = (mod * scales) ** 2
    return (R * S2[:, None, :]) @ R.transpose(1, 2)


@dataclass
class RasterSettings:
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor     # any shape with 16 elements, column-major (= w2c transposed)
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    debug: bool = False

    @staticmethod
    def from_any(s, dtype) -> "RasterSettings":
        return RasterSettings(
            int(s.image_height), int(s.image_width), float(s.tanfovx), float(s.tanfovy),
            torch.as_tensor(s.bg).detach().cpu().to(dtype), float(s.scale_modifier),
            torch.as_tensor(s.viewmatrix).detach().cpu().to(dtype).reshape(16),
            torch.as_tensor(s.projmatrix).detach().cpu().to(dtype).reshape(16),
            int(s.sh_degree), torch.as_tensor(s.campos).detach().cpu().to(dtype).reshape(3),
            bool(s.prefiltered), bool(s.debug))


def _xf43(p, M):
    return torch.stack([M[0] * p[:, 0] + M[4] * p[:, 1] + M[8] * p[:, 2] + M[12],
                        M[1] * p[:, 0] + M[5] * p[:, 1] + M[9] * p[:, 2] + M[13],
                        M[2] * p[:, 0] + M[6] * p[:, 1] + M[10] * p[:, 2] + M[14]], dim=1)


def _xf44(p, M):
    return torch.stack([M[0] * p[:, 0] + M[4] * p[:, 1] + M[8] * p[:, 2] + M[12],
                        M[1] * p[:, 0] + M[5] * p[:, 1] + M[9] * p[:, 2] + M[13],
                        M[2] * p[:, 0] + M[6] * p[:, 1] + M[10] * p[:, 2] + M[14],
                        M[3] * p[:, 0] + M[7] * p[:, 1] + M[11] * p[:, 2] + M[15]], dim=1)


def preprocess(means3D, means2D, opacities, scales, rotations, cov3D_precomp, st: RasterSettings,
               colors_precomp=None, shs=None) -> Dict[str, torch.Tensor]:
    """Appendix A, K1.  All per-Gaussian; differentiable where continuous."""
    dt = means3D.dtype
    P = means3D.shape[0]
    H, W = st.image_height, st.image_width
    V, PM = st.viewmatrix.to(dt), st.projmatrix.to(dt)
    focal_x = W / (2.0 * st.tanfovx)
    focal_y = H / (2.0 * st.tanfovy)
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK

    p_view = _xf43(means3D, V)
    p_hom = _xf44(means3D, PM)
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    p_proj = p_hom[:, :3] * p_w[:, None]

    if cov3D_precomp is not None:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4],
                             c[:, 2], c[:, 4], c[:, 5]], dim=1).view(P, 3, 3)
    else:
        Sigma = build_cov3d(scales, rotations, st.scale_modifier)

    # cov2D with the frustum clamp; a clamped coordinate is a constant for the backward
    tz = p_view[:, 2]
    limx, limy = 1.3 * st.tanfovx, 1.3 * st.tanfovy
    txtz, tytz = p_view[:, 0] / tz, p_view[:, 1] / tz
    cx_ = (txtz < -limx) | (txtz > limx)
    cy_ = (tytz < -limy) | (tytz > limy)
    tx = torch.where(cx_, (txtz.clamp(-limx, limx) * tz).detach(), p_view[:, 0])
    ty = torch.where(cy_, (tytz.clamp(-limy, limy) * tz).detach(), p_view[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([focal_x / tz, zero, -(focal_x * tx) / (tz * tz),
                     zero, focal_y / tz, -(focal_y * ty) / (tz * tz)], dim=1).view(P, 2, 3)
    W3 = torch.stack([V[0], V[4], V[8], V[1], V[5], V[9], V[2], V[6], V[10]]).view(3, 3)
    M = J @ W3
    cov = M @ Sigma @ M.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c2 = cov[:, 1, 1] + 0.3
    det = a * c2 - b * b
    det_safe = torch.where(det == 0, torch.ones_like(det), det)
    conic = torch.stack([c2 / det_safe, -b / det_safe, a / det_safe], dim=1)
    with torch.no_grad():
        mid = 0.5 * (a + c2)
        lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
        rad_f = 3.0 * torch.sqrt(lam)
        radius = torch.ceil(rad_f)
    xy = torch.stack([((p_proj[:, 0] + means2D[:, 0] + 1.0) * W - 1.0) * 0.5,
                      ((p_proj[:, 1] + means2D[:, 1] + 1.0) * H - 1.0) * 0.5], dim=1)
    with torch.no_grad():
        xyd = xy.detach()
        rmin_x = torch.clamp(torch.trunc((xyd[:, 0] - radius) / BLOCK), 0, gx).long()
        rmin_y = torch.clamp(torch.trunc((xyd[:, 1] - radius) / BLOCK), 0, gy).long()
        rmax_x = torch.clamp(torch.trunc((xyd[:, 0] + radius + BLOCK - 1) / BLOCK), 0, gx).long()
        rmax_y = torch.clamp(torch.trunc((xyd[:, 1] + radius + BLOCK - 1) / BLOCK), 0, gy).long()
        area = (rmax_x - rmin_x) * (rmax_y - rmin_y)
        visible = (p_view[:, 2] > NEAR_CULL) & (det != 0) & (area > 0)
        radii = torch.where(visible, radius, torch.zeros_like(radius)).to(torch.int32)
        # fragility of integer decisions (used by tests to exclude decision-flip pixels)
        frac = rad_f - torch.floor(rad_f)
        m_rad = torch.minimum(frac, 1 - frac) / rad_f.clamp(min=1.0)
        def _edge_px(vv):      # distance (pixels) of an edge coordinate to the nearest tile boundary
            q = vv / BLOCK
            f = q - torch.floor(q)
            return torch.minimum(f, 1 - f) * BLOCK

        def _edge(vv):         # the same, relative to the coordinate's magnitude
            return _edge_px(vv) / (vv.abs() + 1.0)
        m_edges = torch.stack([_edge(xyd[:, 0] - radius), _edge(xyd[:, 1] - radius),
                               _edge(xyd[:, 0] + radius + BLOCK - 1), _edge(xyd[:, 1] + radius + BLOCK - 1)], dim=1)
        # a radius that is about to round the other way moves all four edges by one pixel; that only
        # matters where the moved edge is within one pixel of a tile boundary
        near_px = torch.stack([_edge_px(xyd[:, 0] - radius), _edge_px(xyd[:, 1] - radius),
                               _edge_px(xyd[:, 0] + radius + BLOCK - 1), _edge_px(xyd[:, 1] + radius + BLOCK - 1)], dim=1)
        m_rad_edges = torch.where(near_px <= 1.0 + 1e-3, m_rad[:, None].expand(-1, 4), torch.full_like(m_edges, float("inf")))
        m_edges = torch.minimum(m_edges, m_rad_edges)
        m_z = (p_view[:, 2] - NEAR_CULL).abs() / NEAR_CULL
        margin = torch.minimum(m_edges.min(dim=1).values, m_z)

    out = dict(p_view=p_view, depth=p_view[:, 2], xy=xy, conic=conic, cov2d=torch.stack([a, b, c2], 1),
               radii=radii, rect=torch.stack([rmin_x, rmin_y, rmax_x, rmax_y], 1), visible=visible,
               tiles_touched=torch.where(visible, area, torch.zeros_like(area)), margin=margin,
               margin_edges=m_edges, margin_z=m_z, margin_radius=m_rad, rad_f=rad_f)
    if opacities is not None:
        out["opacity"] = opacities.reshape(-1).detach()

    if shs is not None:
        # shs [P,K,3]; dir from campos (Appendix A K1 colour branch)
        dirs = means3D - st.campos.to(dt)[None, :]
        dirs = dirs / dirs.norm(dim=1, keepdim=True)
        rgb = eval_sh(st.sh_degree, shs.transpose(1, 2), dirs) + 0.5
        out["colors"] = torch.clamp_min(rgb, 0.0)
        out["clamped"] = (rgb < 0).detach()
    else:
        out["colors"] = colors_precomp
    return out


def build_tile_lists(pre: Dict[str, torch.Tensor], H: int, W: int, sort_depth=None) -> Tuple[List[torch.Tensor], int]:
    """Appendix A, K2-K5: per tile, Gaussian ids ordered by (depth float32 bits, id).  ``sort_depth`` (optional,
    float32 [P]): take the sort keys from these depths instead (see raster_oracle.c, fsgs_oracle_forward)."""
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    vis = pre["visible"].nonzero().flatten()
    rect = pre["rect"][vis]
    depth32 = pre["depth"].detach().to(torch.float32)
    if sort_depth is not None:        # (a NaN entry = no override for that Gaussian)
        sd = sort_depth.detach().to(torch.float32)
        depth32 = torch.where(torch.isnan(sd), depth32, sd)
    ids_l, tiles_l = [], []
    # expand rect -> (tile, id) instances, row-major inside the rect like duplicateWithKeys
    w = (rect[:, 2] - rect[:, 0])
    h = (rect[:, 3] - rect[:, 1])
    n = w * h
    R = int(n.sum())
    rep = torch.repeat_interleave(torch.arange(vis.numel()), n)
    start = torch.cumsum(n, 0) - n
    local = torch.arange(R) - start[rep]
    ty = rect[rep, 1] + local // w[rep]
    tx = rect[rep, 0] + local % w[rep]
    tile = ty * gx + tx
    gid = vis[rep]
    # stable sort: key = (tile, depth bits), ties keep ascending id (emission order is id order)
    dbits = depth32[gid].view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    key = (tile.to(torch.int64) << 32) | dbits
    order = torch.argsort(key, stable=True)
    tile_s, gid_s = tile[order], gid[order]
    counts = torch.bincount(tile_s, minlength=gx * gy)
    lists = list(torch.split(gid_s, counts.tolist()))
    return lists, R


def composite(pre: Dict[str, torch.Tensor], lists: List[torch.Tensor], st: RasterSettings,
              extra_channels: Optional[torch.Tensor] = None, want_aux: bool = False, margin_kappa: float = 0.0):
    """Appendix A, K6, vectorised per tile.  Returns color [C,H,W], depth [1,H,W], aux."""
    xy, conic, colors, depth = pre["xy"], pre["conic"], pre["colors"], pre["depth"]
    dt = xy.dtype
    opac = pre["opacity"]
    H, W = st.image_height, st.image_width
    gx = (W + BLOCK - 1) // BLOCK
    C = colors.shape[1]
    bg = st.bg.to(dt)
    out_color = torch.zeros(C, H, W, dtype=dt)
    out_depth = torch.zeros(1, H, W, dtype=dt)
    out_color = out_color + bg[:, None, None]  # tiles with empty lists: T=1 -> bg
    n_contrib = torch.zeros(H, W, dtype=torch.int32)
    final_T = torch.ones(H, W, dtype=dt)
    pix_margin = torch.full((H, W), float("inf"), dtype=torch.float64)
    order_margin = torch.full((H, W), float("inf"), dtype=torch.float64)
    color_tiles, depth_tiles = {}, {}
    for t, ids in enumerate(lists):
        if ids.numel() == 0:
            continue
        ty0, tx0 = (t // gx) * BLOCK, (t % gx) * BLOCK
        ph, pw = min(BLOCK, H - ty0), min(BLOCK, W - tx0)
        py, px = torch.meshgrid(torch.arange(ph), torch.arange(pw), indexing="ij")
        pxf = (px.reshape(-1) + tx0).to(dt)
        pyf = (py.reshape(-1) + ty0).to(dt)
        dx = xy[ids, 0:1] - pxf[None, :]
        dy = xy[ids, 1:2] - pyf[None, :]
        con = conic[ids]
        power = -0.5 * (con[:, 0:1] * dx * dx + con[:, 2:3] * dy * dy) - con[:, 1:2] * dx * dy
        G = torch.exp(power)
        oG = opac[ids][:, None] * G
        alpha = oG + (torch.clamp(oG, max=0.99) - oG).detach()      # straight-through clamp
        with torch.no_grad():
            valid = (power <= 0) & (alpha >= ALPHA_MIN)
        a = torch.where(valid, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a
        T_incl = torch.cumprod(one_m, dim=0)
        T_before = torch.cat([torch.ones_like(T_incl[:1]), T_incl[:-1]], dim=0)
        with torch.no_grad():
            test_T = T_before * one_m
            stop = valid & (test_T < T_MIN)
            stopped = torch.cumsum(stop.to(torch.int32), dim=0) > 0      # includes the stop entry
            keep = valid & ~stopped
            last = torch.where(keep, torch.arange(1, ids.numel() + 1)[:, None], 0).max(dim=0).values
        wgt = torch.where(keep, a * T_before, torch.zeros_like(a))
        Tfin = torch.where(keep, one_m, torch.ones_like(one_m)).prod(dim=0)
        col = wgt.transpose(0, 1) @ colors[ids]                     # [npix, C]
        dep = wgt.transpose(0, 1) @ depth[ids]                      # [npix]
        col = col + Tfin[:, None] * bg[None, :]
        color_tiles[t] = (ty0, tx0, ph, pw, col)
        depth_tiles[t] = dep
        with torch.no_grad():
            n_contrib[ty0:ty0 + ph, tx0:tx0 + pw] = last.view(ph, pw).to(torch.int32)
            final_T[ty0:ty0 + ph, tx0:tx0 + pw] = Tfin.view(ph, pw)
            if want_aux:
                # entries the kernel actually evaluates: everything up to and incl. the stop entry
                first_stop = stop & (torch.cumsum(stop.to(torch.int32), 0) == 1)
                live = (~stopped) | first_stop
                # margin_kappa: relative distances are divided by (1 + kappa * g), g = |d power / d centre| per pixel
                # of centre displacement (T test: alpha-weighted sum over the entries in front) -- see raster_oracle.c
                gpx = (con[:, 0:1] * dx + con[:, 1:2] * dy).abs() + (con[:, 2:3] * dy + con[:, 1:2] * dx).abs()
                gsum = torch.cumsum(gpx * a / (1.0 - a), dim=0)
                m_a = torch.where(live & (power <= 0), (alpha - ALPHA_MIN).abs() / ALPHA_MIN / (1 + margin_kappa * gpx),
                                  torch.full_like(alpha, float("inf")))
                m_t = torch.where(live & valid, (test_T - T_MIN).abs() / T_MIN / (1 + margin_kappa * gsum),
                                  torch.full_like(alpha, float("inf")))
                m_p = torch.where(live, power.abs() < 1e-12, torch.zeros_like(valid))
                m = torch.minimum(m_a, m_t).min(dim=0).values.double()
                m = torch.where(m_p.any(dim=0), torch.zeros_like(m), m)
                pix_margin[ty0:ty0 + ph, tx0:tx0 + pw] = m.view(ph, pw)
                # depth-order margin: smallest relative depth gap between consecutive entries that both reach
                # alpha >= 1/255 at the pixel (entries evaluated up to and including the stop entry)
                reach = live & valid
                dcol = depth[ids].detach()[:, None].expand_as(alpha)
                seen_d = torch.where(reach, dcol, torch.full_like(dcol, -1.0))
                prev = torch.cat([torch.full_like(seen_d[:1], -1.0), torch.cummax(seen_d, dim=0).values[:-1]], dim=0)
                gap = torch.where(reach & (prev >= 0), (dcol - prev) / dcol, torch.full_like(dcol, float("inf")))
                order_margin[ty0:ty0 + ph, tx0:tx0 + pw] = gap.min(dim=0).values.double().view(ph, pw)
    # assemble with autograd-friendly ops (index_put on views keeps the graph)
    if color_tiles:
        rows_c, rows_d, idx = [], [], []
        for t, (ty0, tx0, ph, pw, col) in color_tiles.items():
            py, px = torch.meshgrid(torch.arange(ph), torch.arange(pw), indexing="ij")
            idx.append(((py.reshape(-1) + ty0) * W + px.reshape(-1) + tx0))
            rows_c.append(col)
            rows_d.append(depth_tiles[t])
        idx = torch.cat(idx)
        flat_c = out_color.reshape(C, H * W).clone()
        flat_c = flat_c.index_copy(1, idx, torch.cat(rows_c, 0).transpose(0, 1))
        flat_d = out_depth.reshape(1, H * W).clone()
        flat_d = flat_d.index_copy(1, idx, torch.cat(rows_d, 0)[None, :])
        out_color = flat_c.view(C, H, W)
        out_depth = flat_d.view(1, H, W)
    aux = dict(n_contrib=n_contrib, final_T=final_T, pix_margin=pix_margin, order_margin=order_margin)
    return out_color, out_depth, aux


def rasterize(means3D, means2D, opacities, st, colors_precomp=None, shs=None, scales=None,
              rotations=None, cov3D_precomp=None, want_aux: bool = False, sort_depth=None, margin_kappa: float = 0.0):
    """One ``GaussianRasterizer(raster_settings)(...)`` call.  Returns (color, radii, depth, aux)."""
    if (shs is None) == (colors_precomp is None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
    dt = means3D.dtype
    if not isinstance(st, RasterSettings):
        st = RasterSettings.from_any(st, dt)
    H, W = st.image_height, st.image_width
    P = means3D.shape[0]
    if P == 0:
        C = 3
        return (st.bg.to(dt)[:, None, None].expand(C, H, W).clone(), torch.zeros(0, dtype=torch.int32),
                torch.zeros(1, H, W, dtype=dt), {})
    pre = preprocess(means3D, means2D, opacities, scales, rotations, cov3D_precomp, st,
                     colors_precomp=colors_precomp, shs=shs)
    pre["opacity"] = opacities.reshape(-1)
    lists, R = build_tile_lists(pre, H, W, sort_depth=sort_depth)
    color, depth, aux = composite(pre, lists, st, want_aux=want_aux, margin_kappa=margin_kappa)
    aux.update(num_rendered=R, lists=lists, pre=pre)
    return color, pre["radii"], depth, aux


def fragile_pixel_mask(aux, H: int, W: int, eps_pix: float = 1e-4, eps_gauss: float = 2e-6,
                       refine: bool = True, eps_order: float = 1e-6) -> torch.Tensor:
    """Pixels where a float32 implementation may legitimately flip a threshold decision relative to
    this oracle.  Needs ``want_aux=True``.  ``eps_*`` are relative distances to the threshold
    (float32 epsilon is 6e-8).
      * per pixel: some evaluated entry has alpha within eps_pix of 1/255, or T within eps_pix of 1e-4;
      * per pixel: two consecutive entries that both reach alpha >= 1/255 there have view depths within a relative
        eps_order of each other -- the list is sorted by FLOAT32 depth bits, so an implementation whose depths
        round differently composites the pair in the other order (an O(alpha^2) change over their overlap);
      * per Gaussian: a tile-rectangle edge within eps_gauss of a tile boundary (directly, or through
        an integer radius about to round the other way) can add/remove one line of tiles at that
        edge -- the two tile lines either side of the edge are candidates; a near-plane decision within
        eps_gauss makes the whole rectangle a candidate.  With ``refine`` (default) a candidate pixel is
        only marked if that Gaussian could change it at all, i.e. its alpha there reaches 1/255 (a tile
        line beyond the 3-sigma radius mostly holds pixels the splat skips anyway); ``refine=False`` marks
        every candidate pixel (the round-1 behaviour, ~8 % of config 1).
    """
    mask = aux["pix_margin"] < eps_pix
    if "order_margin" in aux and eps_order > 0:
        mask = mask | (aux["order_margin"] < eps_order)
    pre = aux["pre"]
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    tile_mask = torch.zeros(gy, gx, dtype=torch.bool)
    me, mz = pre["margin_edges"], pre["margin_z"]
    frag = ((me.min(dim=1).values < eps_gauss) | (pre["visible"] & (mz < eps_gauss)) |
            (~pre["visible"] & (mz < eps_gauss))).nonzero().flatten()
    refine = refine and ("opacity" in pre) and ("rad_f" in pre)
    ys = torch.arange(H, dtype=torch.float64)[:, None]
    xs = torch.arange(W, dtype=torch.float64)[None, :]

    def rect_of(x, y, r, d):
        # the reference's tile rectangle with every edge coordinate moved by d * (|edge| + 1)
        e = [x - r, y - r, x + r + BLOCK - 1, y + r + BLOCK - 1]
        e = [v + d * (abs(v) + 1.0) for v in e]
        lim = [gx, gy, gx, gy]
        return [min(max(int(v / BLOCK), 0), lim[k]) for k, v in enumerate(e)]    # int(): truncation toward zero

    for i in frag.tolist():
        x0, y0, x1, y1 = pre["rect"][i].tolist()
        ox0, oy0, ox1, oy1 = max(0, x0 - 1), max(0, y0 - 1), min(gx, x1 + 1), min(gy, y1 + 1)
        if not refine:
            if mz[i] < eps_gauss:
                tile_mask[oy0:oy1, ox0:ox1] = True
                continue
            e = me[i] < eps_gauss
            if e[0]:
                tile_mask[oy0:oy1, max(0, x0 - 1):min(gx, x0 + 1)] = True
            if e[1]:
                tile_mask[max(0, y0 - 1):min(gy, y0 + 1), ox0:ox1] = True
            if e[2]:
                tile_mask[oy0:oy1, max(0, x1 - 1):min(gx, x1 + 1)] = True
            if e[3]:
                tile_mask[max(0, y1 - 1):min(gy, y1 + 1), ox0:ox1] = True
            continue
        # candidate tiles: those a float32 implementation may add to / drop from this Gaussian's rectangle, i.e.
        # the tiles not common to every variant of the rectangle under the perturbations that are in question
        # (edge coordinates moved by +-eps, the integer radius rounding the other way); a near-plane flip
        # puts the whole rectangle in question
        tm = torch.zeros(gy, gx, dtype=torch.bool)
        if mz[i] < eps_gauss:
            tm[oy0:oy1, ox0:ox1] = True
        else:
            x, y = float(pre["xy"][i, 0]), float(pre["xy"][i, 1])
            rf = float(pre["rad_f"][i])
            rad = float(math.ceil(rf))
            rads = [rad]
            if float(pre["margin_radius"][i]) < eps_gauss:
                rads.append(rad - 1.0 if rf - math.floor(rf) < 0.5 else rad + 1.0)
            union = torch.zeros(gy, gx, dtype=torch.bool)
            inter = torch.ones(gy, gx, dtype=torch.bool)
            for r_ in rads:
                for d in (0.0, -eps_gauss, eps_gauss):
                    a0, b0, a1, b1 = rect_of(x, y, r_, d)
                    v = torch.zeros(gy, gx, dtype=torch.bool)
                    v[b0:b1, a0:a1] = True
                    union |= v
                    inter &= v
            tm = union & ~inter
        if tm.any():
            rows = tm.any(dim=1).nonzero().flatten()
            cols = tm.any(dim=0).nonzero().flatten()
            ya, yb = int(rows[0]) * BLOCK, min(H, (int(rows[-1]) + 1) * BLOCK)
            xa, xb = int(cols[0]) * BLOCK, min(W, (int(cols[-1]) + 1) * BLOCK)
            dx = float(pre["xy"][i, 0]) - xs[:, xa:xb]
            dy = float(pre["xy"][i, 1]) - ys[ya:yb]
            con = pre["conic"][i].double()
            power = -0.5 * (con[0] * dx * dx + con[2] * dy * dy) - con[1] * dx * dy
            alpha = float(pre["opacity"][i]) * torch.exp(power)
            hit = (power <= 1e-9) & (alpha >= ALPHA_MIN * (1.0 - 1e-3))
            cand = tm.repeat_interleave(BLOCK, 0).repeat_interleave(BLOCK, 1)[ya:yb, xa:xb]
            mask[ya:yb, xa:xb] |= hit & cand
    if frag.numel() and not refine:
        mask = mask | tile_mask.repeat_interleave(BLOCK, 0).repeat_interleave(BLOCK, 1)[:H, :W]
    return mask


def fragile_radii(aux, eps: float = 1e-5) -> torch.Tensor:
    """Gaussians whose integer screen radius ceil(3 sqrt(lambda)) or near-plane decision sits within a relative
    ``eps`` of rounding the other way in float32 -- the only ones whose ``radii`` entry may differ from this
    oracle's.  Needs ``want_aux=True``."""
    pre = aux["pre"]
    return (pre["margin_radius"] < eps) | (pre["margin_z"] < eps)
