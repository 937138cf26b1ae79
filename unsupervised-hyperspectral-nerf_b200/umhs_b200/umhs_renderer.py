"""SpectralRenderer (a.k.a. UMHSRenderer): per-ray accumulation of per-sample values, and the dense weight function.

Mirrors ref umhsnerf/umhs_renderer.py:10-139: ``SpectralRenderer.forward(spectral, weights, ray_indices, num_rays)``
(classmethod; dense ``sum(dim=-2)`` when ``ray_indices`` is None) and ``get_weights_spectral(deltas, densities)``.
Packed accumulation and the weight scan run in libumhs_sm100.so (csrc/unfused.cu).  UMHSModel does not go through
this module on its hot path: compositing is fused with the unmixing there (csrc/composite.cu).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from . import ops


class SpectralRenderer(nn.Module):
    background_color: Optional[str] = "random"

    @classmethod
    def forward(cls, spectral: Tensor, weights: Tensor, ray_indices: Optional[Tensor] = None,
                num_rays: Optional[int] = None) -> Tensor:
        if spectral.dim() == 3:
            spectral = spectral.squeeze(0)
        elif spectral.dim() == 1:
            spectral = spectral.unsqueeze(0)
        if ray_indices is not None and num_rays is not None:
            return ops.AccumulateFunction.apply(weights[..., 0], spectral, ray_indices, int(num_rays))
        # dense layout [..., S, C]: a ray is a contiguous run of S samples
        S, C = spectral.shape[-2], spectral.shape[-1]
        lead = spectral.shape[:-2]
        R = int(torch.tensor(lead).prod().item()) if len(lead) else 1
        idx = torch.arange(R, device=spectral.device).repeat_interleave(S)
        out = ops.AccumulateFunction.apply(weights.reshape(-1), spectral.reshape(-1, C), idx, R)
        return out.view(*lead, C)

    def __call__(self, *args, **kwargs):  # nn.Module.__call__ would pass `self` to the classmethod twice
        return type(self).forward(*args, **kwargs)


UMHSRenderer = SpectralRenderer


def get_weights_spectral(deltas: Tensor, densities: Tensor) -> Tensor:
    """Dense-layout weights [R,S,1] (ref umhs_renderer.py:117-139): alpha_i * exp(-sum_{j<i} delta_j sigma_j), with
    NaNs mapped to 0 as the reference's nan_to_num does."""
    R, S = deltas.shape[0], deltas.shape[1]
    dev = deltas.device
    packed = torch.stack([torch.arange(R, device=dev, dtype=torch.int64) * S,
                          torch.full((R,), S, device=dev, dtype=torch.int64)], dim=-1).contiguous()
    zeros = torch.zeros((R * S,), device=dev, dtype=torch.float32)
    w, _ = ops.RenderWeightsFunction.apply(zeros, deltas.reshape(-1), densities.reshape(-1), packed)
    return torch.nan_to_num(w).view(R, S, 1)
