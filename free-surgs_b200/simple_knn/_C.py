"""``distCUDA2(points[P,3]) -> [P]``: mean squared distance to the 3 nearest OTHER points.

Replaces reference ``submodules/simple-knn`` (``spatial.cu:15-26`` -> ``SimpleKNN::knn``
``simple_knn.cu:185-219``; per-point result ``simple_knn.cu:147-183``).  The reference's
Morton-order + box-pruning search is exact, so an exact search returns the same values up to float
rounding.  It runs once at initialisation (``scene/gaussian_model.py:346`` on ~10 % of the first frame's
pixels, ~131 k points at 1280x1024) and stays PyTorch on-device per the north-star; it is not part of the
per-frame hot path.

Memory-bounded: the [rows, columns] distance tile is sized from ``budget_bytes`` (default 256 MiB for the few
temporaries alive at once) and both axes are tiled, with a running best-3 per row merged tile by tile -- the
round-1 version materialised a [2048, P, 3] difference tensor (3.2 GB at 131 k points, 49 GB at 2 M).
"""
import torch


def distCUDA2(points: torch.Tensor, budget_bytes: int = 256 << 20) -> torch.Tensor:
    pts = points.detach().float()
    P = pts.shape[0]
    out = torch.empty(P, dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    x, y, z = (pts[:, k].contiguous() for k in range(3))
    cols = min(P, 1 << 16)
    rows = max(1, min(P, budget_bytes // (cols * 4 * 4)))             # ~4 float tiles alive at once
    inf = float("inf")
    for s in range(0, P, rows):
        e = min(P, s + rows)
        qx, qy, qz = x[s:e, None], y[s:e, None], z[s:e, None]
        best = torch.full((e - s, 3), inf, device=pts.device)
        for c in range(0, P, cols):
            ce = min(P, c + cols)
            d2 = (qx - x[None, c:ce]) ** 2                              # exact differences, no matmul trick
            d2 += (qy - y[None, c:ce]) ** 2
            d2 += (qz - z[None, c:ce]) ** 2
            lo, hi = max(s, c), min(e, ce)
            if lo < hi:                                                 # exclude self by INDEX, as the reference does
                idx = torch.arange(lo, hi, device=pts.device)
                d2[idx - s, idx - c] = inf
            k = min(3, ce - c)
            cand = torch.topk(d2, k, dim=1, largest=False).values
            best = torch.topk(torch.cat([best, cand], dim=1), 3, dim=1, largest=False).values
        # fewer than 3 other points: the reference leaves FLT_MAX in the unused slots (-> an overflowing mean); inf here
        out[s:e] = best.sum(dim=1) / 3.0
    return out
