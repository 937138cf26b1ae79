"""UMHSField: hash-grid density field with endmember-abundance heads (spectral unmixing), B200-native.

Drop-in for ``UMHSField`` of ref umhsnerf/umhs_field.py:28-329 on its spectral branch: same constructor
arguments, the same parameter / buffer names and shapes as the reference's torch implementation (so reference
checkpoints load: SURVEY.md A.9), and the same ``get_density`` / ``get_outputs`` / ``forward`` / ``density_fn``
surface.  All per-sample arithmetic runs in libumhs_sm100.so (csrc/hash_grid.cu, csrc/field_mlp.cu,
csrc/unfused.cu); there is no PyTorch fallback.

The model's fused path calls ``forward_packed`` (sigma / scalar / abundances only: per-sample spectra are never
materialised, the unmixing happens after compositing, see csrc/composite.cu).  ``get_outputs`` is the reference's
module boundary and does return per-sample spectra.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Literal, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import ops
from .ops import MLP_NAMES, PointSet
from .rays import Frustums, RaySamples

GEO_FEAT_DIM = 15
PE_DIM = 12
SH_DIM = 16
HIDDEN = 64


class _Layers(nn.Module):
    """Parameter container with nerfstudio MLP's state-dict names: ``layers.{i}.{weight,bias}`` (nn.Linear init)."""

    def __init__(self, dims):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])


class _HashEncoder(nn.Module):
    """Parameter container of nerfstudio HashEncoding (torch path): ``hash_table`` [L*T, F], ``scalings`` [L]."""

    def __init__(self, num_levels, min_res, max_res, log2_hashmap_size, features_per_level, hash_init_scale=1e-3):
        super().__init__()
        self.num_levels, self.log2_hashmap_size = num_levels, log2_hashmap_size
        table = (torch.rand((num_levels << log2_hashmap_size, features_per_level)) * 2 - 1) * hash_init_scale
        self.hash_table = nn.Parameter(table)
        growth = math.exp((math.log(max_res) - math.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
        # fp32 pow, as nerfstudio evaluates it (the buffer travels in the state dict; kernels take it as data)
        self.register_buffer("scalings", torch.floor(min_res * growth ** torch.arange(num_levels)).to(torch.float32))

    def forward(self, x01: Tensor) -> Tensor:
        return ops.HashEncodeFunction.apply(x01, self.hash_table, self.scalings, self.log2_hashmap_size)


class _MLPWithHashEncoding(nn.Module):
    def __init__(self, num_levels, min_res, max_res, log2_hashmap_size, features_per_level, out_dim):
        super().__init__()
        self.encoder = _HashEncoder(num_levels, min_res, max_res, log2_hashmap_size, features_per_level)
        self.mlp = _Layers([num_levels * features_per_level, HIDDEN, out_dim])


class UMHSField(nn.Module):
    """See module docstring.  ``implementation`` is accepted for signature compatibility; the only
    implementation is the sm_100a kernel library."""

    aabb: Tensor

    def __init__(
        self,
        aabb: Tensor,
        num_images: int = 0,
        implementation: Literal["tcnn", "torch", "b200"] = "b200",
        num_layers_color: int = 3,
        hidden_dim_color: int = 64,
        wavelengths: int = 128,
        method: Literal["rgb", "spectral", "rgb+spectral"] = "rgb+spectral",
        num_classes: int = 4,
        feature_dim: int = 256,
        temperature: float = 0.5,
        converter=None,
        pred_dino: bool = False,
        pred_specular: bool = False,
        load_vca: bool = False,
        # NerfactoField keyword arguments reached through **kwargs in the reference
        num_levels: int = 16,
        base_res: int = 16,
        max_res: int = 2048,
        log2_hashmap_size: int = 19,
        features_per_level: int = 2,
        appearance_embedding_dim: int = 0,
        spatial_distortion="contraction",
        **unused,
    ) -> None:
        super().__init__()
        if "spectral" not in method:
            raise NotImplementedError("method='rgb' is broken in the reference (ref umhs_model.py:266) and not built")
        if pred_dino:
            raise NotImplementedError("pred_dino crashes in the reference (UMHSField._get_encoding is undefined)")
        if num_layers_color != 3 or hidden_dim_color != HIDDEN:
            raise NotImplementedError("the fused MLP kernel is built for num_layers_color=3, hidden_dim_color=64")
        if num_levels != 16 or features_per_level != 2:
            raise NotImplementedError("the fused kernels are built for 16 levels x 2 features")
        if appearance_embedding_dim != 0:
            raise NotImplementedError("appearance embedding is 0-wide in every reference configuration (quirk 1)")
        if not (1 <= num_classes <= 16) or not (1 <= wavelengths <= 512):
            raise ValueError("num_classes in [1,16], wavelengths in [1,512]")
        self.method, self.num_classes, self.wavelengths = method, num_classes, wavelengths
        self.feature_dim, self.pred_specular, self.temperature = feature_dim, pred_specular, float(temperature)
        self.geo_feat_dim, self.average_init_density = GEO_FEAT_DIM, 1
        self.spatial_distortion = spatial_distortion
        self.converter = converter
        self.use_scalar = True
        self.register_buffer("aabb", torch.as_tensor(aabb, dtype=torch.float32))
        self._aabb_host = None  # host copy of aabb for kernel arguments (refreshed when the buffer object changes)
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self._log2_T = int(log2_hashmap_size)

        self.mlp_base = _MLPWithHashEncoding(num_levels, base_res, max_res, log2_hashmap_size, features_per_level,
                                             1 + GEO_FEAT_DIM)
        in_dim = PE_DIM + GEO_FEAT_DIM
        self.feature_mlp = _Layers([in_dim, HIDDEN, HIDDEN, num_classes + (1 if pred_specular else 0)])
        if load_vca and os.path.exists("vca.npy"):  # ref umhs_field.py:77-81
            self.endmembers = nn.Parameter(torch.tensor(np.load("vca.npy"), dtype=torch.float32))
        else:
            self.endmembers = nn.Parameter(torch.randn(num_classes, wavelengths))
        self.mlp_head = _Layers([in_dim, HIDDEN, HIDDEN, num_classes])
        self.mlp_directional = _Layers([SH_DIM + PE_DIM, 16, wavelengths])
        self._sample_locations = None
        self._density_before_activation = None

    # ------------------------------------------------------------------ parameter plumbing
    def _mlp_tensors(self):
        b, h, f, d = self.mlp_base.mlp.layers, self.mlp_head.layers, self.feature_mlp.layers, \
            self.mlp_directional.layers
        t = [b[0].weight, b[0].bias, b[1].weight, b[1].bias,
             h[0].weight, h[0].bias, h[1].weight, h[1].bias, h[2].weight, h[2].bias,
             f[0].weight, f[0].bias, f[1].weight, f[1].bias, f[2].weight, f[2].bias]
        t += [d[0].weight, d[0].bias] if self.pred_specular else [None, None]
        return t

    def _weights_dict(self) -> Dict[str, Optional[Tensor]]:
        return {n: (w.detach() if w is not None else None) for n, w in zip(MLP_NAMES, self._mlp_tensors())}

    def _cfg(self) -> dict:
        return {"log2_T": self._log2_T, "K": self.num_classes, "specular": self.pred_specular,
                "temperature": self.temperature}

    @property
    def contraction(self) -> bool:
        return self.spatial_distortion is not None

    def _aabb6(self):
        """aabb as six Python floats.  Cached: reading a CUDA buffer back is a device sync, and this is called on
        every step; the cache is keyed on the buffer's identity/version so load_state_dict / .to() refresh it."""
        key = (self.aabb.data_ptr(), self.aabb._version)
        if self._aabb_host is None or self._aabb_host[0] != key:
            self._aabb_host = (key, tuple(float(v) for v in self.aabb.detach().flatten().cpu().tolist()))
        return self._aabb_host[1]

    def _points(self, ray_samples: RaySamples) -> PointSet:
        packed = getattr(ray_samples, "packed", None)
        if isinstance(packed, PointSet):
            return packed
        fr = ray_samples.frustums
        o = fr.origins.reshape(-1, 3)
        d = fr.directions.reshape(-1, 3)
        # positions exactly as Frustums.get_positions forms them
        pos = o + d * (fr.starts.reshape(-1, 1) + fr.ends.reshape(-1, 1)) / 2
        return PointSet.from_positions(pos, d, contraction=self.contraction, aabb=self._aabb6())

    # ------------------------------------------------------------------ fused path used by UMHSModel
    def forward_packed(self, pts: PointSet, want_embedding: bool = True) -> Dict[str, Tensor]:
        """sigma [N], embedding [N,15], scalar [N,K], abund [N,K], (s1 [N], spec_hidden [N,16]) with autograd.
        want_embedding=False (the model's fused path: both heads run inside the kernel and nothing downstream reads
        the geometry features) skips the 60 B/sample embedding output."""
        enc = self.mlp_base.encoder
        cfg = self._cfg()
        cfg["want_embedding"] = want_embedding
        # training steps only: the forward keeps its activation tiles for the backward (ops.SAVE_ACTIVATIONS); an
        # evaluation image (24 M samples) must not write 17 GB of them
        cfg["save_activations"] = bool(self.training and torch.is_grad_enabled())
        sigma, emb, scalar, abund, s1, sh = ops.FieldFunction.apply(pts, cfg, enc.hash_table, enc.scalings,
                                                                    *self._mlp_tensors())
        out = {"sigma": sigma, "embedding": emb, "scalar": scalar, "abund": abund}
        if self.pred_specular:
            out["s1"], out["spec_hidden"] = s1, sh
        return out

    @torch.no_grad()
    def density_packed(self, pts: PointSet, return_enc: bool = False, zero_tail: bool = True):
        """Density only, no autograd: the sampler's sigma_fn and the occupancy-grid update."""
        enc = self.mlp_base.encoder
        e = ops.hash_encode_fwd(pts, enc.hash_table.detach(), enc.scalings, self._log2_T, tiles=ops._tiles_on())
        sigma = ops.density_fwd(pts, e, self._weights_dict(), zero_tail=zero_tail)
        return (sigma, e) if return_enc else sigma

    # ------------------------------------------------------------------ reference module surface
    def get_density(self, ray_samples: RaySamples) -> Tuple[Tensor, Tensor]:
        """ref umhs_field.py:300-329 -> (density [*,1], base_mlp_out [*,15])."""
        pts = self._points(ray_samples)
        out = self.forward_packed(pts)
        self._cached = (pts, out)
        shape = tuple(ray_samples.frustums.shape)
        return out["sigma"].view(*shape, 1), out["embedding"].view(*shape, GEO_FEAT_DIM)

    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """ref umhs_field.py:151-261 (spectral branch): per-sample spectra and abundances.

        The heads are evaluated together with the base MLP in one fused kernel, so this reuses the evaluation
        cached by ``get_density`` on the same samples (the reference always calls the two back to back through
        ``Field.forward``); ``density_embedding`` is accepted for signature compatibility."""
        cached = getattr(self, "_cached", None)
        pts = self._points(ray_samples)
        if cached is not None and cached[0].n == pts.n and (cached[0] is pts or getattr(ray_samples, "packed", None) is None):
            out = cached[1]
        else:
            out = self.forward_packed(pts)
        self._cached = None
        d = self.mlp_directional.layers
        spec, spc = ops.UnmixFunction.apply(out["scalar"], out["abund"], out.get("s1"), out.get("spec_hidden"),
                                            self.endmembers, d[1].weight if self.pred_specular else None,
                                            d[1].bias if self.pred_specular else None)
        res: Dict[str, Tensor] = {}
        if self.pred_specular:
            res["spectral"] = (spec + spc).unsqueeze(0)  # [1,N,B] (quirk 7: broadcast of s1[1,N,1])
            res["spectral2"] = spec
            res["specular"] = spc.detach().unsqueeze(0)
        else:
            res["spectral"] = spec
        res["abundances"] = out["abund"].unsqueeze(0)  # [1,N,K]
        return res

    def forward(self, ray_samples: RaySamples, compute_normals: bool = False) -> Dict[str, Tensor]:
        density, emb = self.get_density(ray_samples)
        out = self.get_outputs(ray_samples, density_embedding=emb)
        out["density"] = density
        return out

    def density_fn(self, positions: Tensor, times: Optional[Tensor] = None, n_dev: Optional[Tensor] = None) -> Tensor:
        """nerfstudio Field.density_fn: positions [..., 3] -> density [..., 1].  ``n_dev`` (extension): device int64
        count; only the first n_dev positions are evaluated, the rest read as zero density (no host sync)."""
        pts = PointSet.from_positions(positions, None, contraction=self.contraction, aabb=self._aabb6(), n_dev=n_dev)
        return self.density_packed(pts).view(*positions.shape[:-1], 1)
