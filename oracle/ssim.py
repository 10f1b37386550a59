"""Oracle: SSIM map (reference ``ssim.py``).  TEST INFRASTRUCTURE.

13x13 Gaussian (sigma 1.5), zero 'same' padding, depthwise."""
from math import exp
import torch
import torch.nn.functional as F


def gaussian(window_size, sigma):
    """Reference ssim.py:9-11 (fp32 taps, normalised in fp32)."""
    g = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2))
                      for x in range(window_size)])
    return g / g.sum()


def create_window(window_size, channel):
    """Reference ssim.py:13-17: 2-D window = fp32 outer product of the 1-D taps."""
    g = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=13, size_average=True):
    """Returns the SSIM *map* [B,C,H,W].  Reference ssim.py:19-36,68-76."""
    C = img1.size(1)
    win = create_window(window_size, C).to(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, win, padding=pad, groups=C)
    mu2 = F.conv2d(img2, win, padding=pad, groups=C)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(img1 * img1, win, padding=pad, groups=C) - mu1_sq
    s2 = F.conv2d(img2 * img2, win, padding=pad, groups=C) - mu2_sq
    s12 = F.conv2d(img1 * img2, win, padding=pad, groups=C) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
