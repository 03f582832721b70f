"""``structural_similarity`` with scikit-image's defaults for float images: uniform 7x7 window, K1 = 0.01,
K2 = 0.03, sample covariance (N / (N - 1)), mean over the interior (window fully inside), channels averaged."""
import numpy as np
from scipy.ndimage import uniform_filter


def structural_similarity(im1, im2, *, win_size=None, data_range=None, channel_axis=None, multichannel=False,
                          gaussian_weights=False, full=False, **kwargs):
    im1 = np.asarray(im1, dtype=np.float64)
    im2 = np.asarray(im2, dtype=np.float64)
    if channel_axis is not None or multichannel:
        ax = -1 if channel_axis is None else channel_axis
        vals = [structural_similarity(np.take(im1, c, axis=ax), np.take(im2, c, axis=ax), win_size=win_size,
                                      data_range=data_range) for c in range(im1.shape[ax])]
        return float(np.mean(vals))
    win = 7 if win_size is None else int(win_size)
    if data_range is None:
        data_range = 2.0 if im1.min() < 0 else 1.0
    K1, K2 = 0.01, 0.03
    npx = win ** im1.ndim
    cov_norm = npx / (npx - 1.0)
    ux, uy = uniform_filter(im1, size=win), uniform_filter(im2, size=win)
    uxx, uyy, uxy = uniform_filter(im1 * im1, size=win), uniform_filter(im2 * im2, size=win), uniform_filter(im1 * im2, size=win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win - 1) // 2
    sl = tuple(slice(pad, s - pad) for s in S.shape)
    return float(S[sl].mean())
