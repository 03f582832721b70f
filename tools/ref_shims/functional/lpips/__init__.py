"""lpips stand-in (utils/general_utils.py:31-35).  The metric needs pretrained AlexNet weights that are a
network download; offline it cannot be computed, so the stand-in returns NaN for every pair -- visibly not a
number -- and PSNR / SSIM, which rgb_evaluation computes next to it, are unaffected."""
import torch


class LPIPS(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, a, b, *args, **kw):
        return torch.full((a.shape[0], 1, 1, 1), float("nan"))
