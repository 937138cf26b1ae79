"""Spectrum -> CIE XYZ -> linear sRGB -> gamma, the fixed colour projection of UMHS.

Mirrors ``ColourSystem`` of ref umhsnerf/utils/spec_to_rgb.py:61-127 (same constructor arguments, the
``transform_matrix`` buffer name and forward semantics).  Inside UMHSModel the projection runs fused in the
compositing kernel epilogue (csrc/composite.cu); ``forward`` here serves standalone use of the module.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor, nn

# piecewise-Gaussian fit of the CIE 1931 colour matching functions (wavelength in Angstrom):
# rows of (amplitude, centre, sigma_left, sigma_right) per lobe -- ref spec_to_rgb.py:6-21
_CMF_LOBES = {
    "x": ((1.056, 5998.0, 379.0, 310.0), (0.362, 4420.0, 160.0, 267.0), (-0.065, 5011.0, 204.0, 262.0)),
    "y": ((0.821, 5688.0, 469.0, 405.0), (0.286, 5309.0, 163.0, 311.0)),
    "z": ((1.217, 4370.0, 118.0, 360.0), (0.681, 4590.0, 260.0, 138.0)),
}
# chromaticities (x, y) of primaries + white point -- ref spec_to_rgb.py:34-59
COLOR_SPACE = {
    "sRGB": {"red": (0.64, 0.33), "green": (0.30, 0.60), "blue": (0.15, 0.06), "white": (0.3127, 0.3291)},
}


def _cmf(angstrom: np.ndarray) -> np.ndarray:
    rows = []
    for axis in ("x", "y", "z"):
        acc = np.zeros_like(angstrom, dtype=np.float64)
        for amp, mu, s_lo, s_hi in _CMF_LOBES[axis]:
            sig = np.maximum(np.where(angstrom < mu, s_lo, s_hi), 1e-6)
            acc = acc + amp * np.exp(-((angstrom - mu) ** 2) / (2.0 * sig ** 2))
        rows.append(acc)
    return np.stack(rows, axis=0)  # [3, B]


def build_transform_matrix(bands, cs: str = "sRGB") -> np.ndarray:
    """[B, 3] matrix, columns normalised to sum to one (ref spec_to_rgb.py:73-90).  ``bands`` in nm."""
    c = COLOR_SPACE[cs]
    xyz = lambda xy: np.array([xy[0], xy[1], 1.0 - xy[0] - xy[1]])  # noqa: E731
    prim = np.stack([xyz(c["red"]), xyz(c["green"]), xyz(c["blue"])], axis=1)
    inv = np.linalg.inv(prim)
    scale = inv @ xyz(c["white"])
    xyz_to_rgb = inv / scale[:, None]
    m = _cmf(np.asarray(bands, dtype=np.float64) * 10.0).T @ xyz_to_rgb.T
    return m / m.sum(axis=0, keepdims=True)


class ColourSystem(nn.Module):
    def __init__(self, bands, cs: str = "sRGB", device=None):
        super().__init__()
        self.bands = list(bands)
        self.register_buffer("transform_matrix", torch.from_numpy(build_transform_matrix(bands, cs)).float())
        if device is not None:
            self.to(device)

    @staticmethod
    def gamma_correction(rgb: Tensor) -> Tensor:
        lo = 12.92 * rgb
        hi = 1.055 * rgb.clamp(min=1e-6).pow(1.0 / 2.4) - 0.055
        return torch.where(rgb < 0.0031308, lo, hi)

    def forward(self, spec: Tensor) -> Tensor:
        rgb = spec.to(self.transform_matrix.dtype) @ self.transform_matrix
        return self.gamma_correction(rgb).clamp(0.0, 1.0)
