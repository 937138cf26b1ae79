"""Image metrics and colour maps used by ``UMHSModel.get_image_metrics_and_images`` (ref umhsnerf/umhs_model.py:407-527).

The reference takes them from third-party packages that are not part of the hot path: torchmetrics 1.5.2
(``PeakSignalNoiseRatio(data_range=1.0)``, ``structural_similarity_index_measure``, ``SpectralAngleMapper(reduction=
"none")``, ``LearnedPerceptualImagePatchSimilarity``) through nerfstudio's ``NGPModel``, and nerfstudio's
``colormaps``.  Restated here in plain torch (runs once per evaluation image, on whatever device the images are on) so
that the model's evaluation surface works without those packages; when torchmetrics IS importable the model uses its
LPIPS (a pretrained network: there is no way to restate its weights) and reports ``lpips`` = NaN otherwise.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import Tensor


def psnr(preds: Tensor, target: Tensor, data_range: float = 1.0) -> Tensor:
    """torchmetrics PeakSignalNoiseRatio(data_range): 10 log10(data_range^2 / mse) over all elements."""
    mse = torch.mean((preds - target) ** 2)
    return 10.0 * torch.log10(torch.as_tensor(data_range ** 2, dtype=mse.dtype, device=mse.device) / mse)


def _gaussian_kernel(size: int, sigma: float, dtype, device) -> Tensor:
    dist = torch.arange((1 - size) / 2, (1 + size) / 2, 1, dtype=dtype, device=device)
    g = torch.exp(-((dist / sigma) ** 2) / 2)
    return (g / g.sum()).unsqueeze(0)


def ssim(preds: Tensor, target: Tensor, kernel_size: int = 11, sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03,
         data_range=None) -> Tensor:
    """torchmetrics ``structural_similarity_index_measure`` defaults ([N, C, H, W] inputs): gaussian window 11 / 1.5,
    data range = max(range of preds, range of target), reflect padding by the half window, SSIM map cropped back to
    the un-padded image, mean over everything."""
    if data_range is None:
        data_range = torch.max(preds.max() - preds.min(), target.max() - target.min())
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    C = preds.shape[1]
    pad = (kernel_size - 1) // 2
    k1d = _gaussian_kernel(kernel_size, sigma, preds.dtype, preds.device)
    kernel = torch.matmul(k1d.t(), k1d).expand(C, 1, kernel_size, kernel_size)
    p = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    stack = torch.cat((p, t, p * p, t * t, p * t))  # 5N x C x H x W
    out = F.conv2d(stack, kernel, groups=C)
    n = preds.shape[0]
    mu_p, mu_t, pp, tt, pt = (out[i * n:(i + 1) * n] for i in range(5))
    mu_pp, mu_tt, mu_pt = mu_p * mu_p, mu_t * mu_t, mu_p * mu_t
    s_pp, s_tt, s_pt = pp - mu_pp, tt - mu_tt, pt - mu_pt
    upper, lower = 2 * s_pt + c2, s_pp + s_tt + c2
    full = ((2 * mu_pt + c1) * upper) / ((mu_pp + mu_tt + c1) * lower)
    full = full[..., pad:-pad, pad:-pad] if full.shape[-1] > 2 * pad and full.shape[-2] > 2 * pad else full
    return full.reshape(full.shape[0], -1).mean(-1).mean()


def spectral_angle(preds: Tensor, target: Tensor) -> Tensor:
    """torchmetrics SpectralAngleMapper(reduction="none") on [N, C, H, W]: per-pixel angle (radians) between the
    spectra, ``acos(clamp(<p, t> / (|p| |t|), -1, 1))`` -> [N, H, W] (0 / 0 gives NaN, which the caller nan-means)."""
    dot = (preds * target).sum(dim=1)
    return torch.clamp(dot / (preds.norm(dim=1) * target.norm(dim=1)), -1, 1).acos()


def compute_sam(pred: Tensor, gt: Tensor, eps: float = 1e-8) -> Tensor:
    """ref umhs_model.py:514-527 (channels last): mean spectral angle."""
    dot = (pred * gt).sum(dim=-1)
    cos = dot / (torch.norm(pred, dim=-1) * torch.norm(gt, dim=-1) + eps)
    return torch.acos(torch.clamp(cos, -1, 1)).mean()


# ---------------------------------------------------------------------------------------------- colour maps
# "turbo" (nerfstudio's default for single-channel images), 5th-order polynomial fit published with the colour map
_TURBO = ((0.13572138, 4.61539260, -42.66032258, 132.13108234, -152.94239396, 59.28637943),
          (0.09140261, 2.19418839, 4.84296658, -14.18503333, 4.27729857, 2.82956604),
          (0.10667330, 12.64194608, -60.58204836, 110.36276771, -89.90310912, 27.34824973))


def apply_colormap(image: Tensor) -> Tensor:
    """nerfstudio colormaps.apply_colormap for a [..., 1] image in [0, 1]: turbo -> [..., 3]."""
    x = torch.nan_to_num(image, 0.0).clamp(0, 1)
    ch = []
    for c in _TURBO:
        v = torch.zeros_like(x)
        for k in reversed(c):
            v = v * x + k
        ch.append(v)
    return torch.cat(ch, dim=-1).clamp(0, 1)


def apply_depth_colormap(depth: Tensor, accumulation: Tensor = None, near_plane=None, far_plane=None) -> Tensor:
    """nerfstudio colormaps.apply_depth_colormap: normalise to [near, far] (the image's own range by default), turbo,
    then composite over white with the accumulation."""
    near = float(depth.min()) if near_plane is None else near_plane
    far = float(depth.max()) if far_plane is None else far_plane
    d = ((depth - near) / (far - near + 1e-10)).clamp(0, 1)
    col = apply_colormap(d)
    if accumulation is not None:
        col = col * accumulation + (1 - accumulation)
    return col


def lpips_or_nan(module, a: Tensor, b: Tensor) -> float:
    if module is None:
        return math.nan
    with torch.no_grad():
        return float(module(a, b))
