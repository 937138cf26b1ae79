"""CPU oracle: restatement of the reference's PyTorch hot path (fp32, CPU, autograd).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module,
and only as the checker.  The product package never imports ``oracle``.

What it restates (each function cites the reference lines it follows):

* ``OracleField``          <- ref umhsnerf/umhs_field.py:33-148 (ctor), :151-261
                              (get_outputs, spectral branch), :300-329 (get_density)
* ``spectral_render``      <- ref umhsnerf/umhs_renderer.py:15-30
* ``get_weights_spectral`` <- ref umhsnerf/umhs_renderer.py:117-139
* ``ColourSystem``         <- ref umhsnerf/utils/spec_to_rgb.py:6-127
* ``cluster_lookup``       <- ref umhsnerf/utils/clusterprobe.py:17-38
* ``OracleModel``          <- ref umhsnerf/umhs_model.py:127-218 (modules), :225-327
                              (get_outputs), :329-383 (get_loss_dict), :385-405
                              (metrics), :542-591 (callbacks)

Pinning status.  The reference has no tests and no golden vectors.  What CAN be
pinned is pinned by ``tests/golden/make_golden.py``: it executes the UNMODIFIED
reference classes (``UMHSField``, ``SpectralRenderer``, ``get_weights_spectral``,
``ColourSystem``, ``ClusterLookup``, ``UMHSModel.get_outputs/get_loss_dict``) from
``/root/reference`` on top of a shim that provides the absent third-party modules
(nerfstudio / nerfacc) from ``oracle/nerf3p.py``, and freezes inputs+outputs into
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle against
them.  The third-party arithmetic itself (``nerf3p.py``, ``pack_oracle.c``) is a
restatement from recollection: PARITY UNPINNED for those parts.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import nerf3p
from .nerf3p import Frustums, RayBundle, RaySamples

# ----------------------------------------------------------------------------
# Colour system  (ref umhsnerf/utils/spec_to_rgb.py)
# ----------------------------------------------------------------------------


def _g(x, alpha, mu, sigma1, sigma2):
    sigma = np.clip((x < mu) * sigma1 + (x >= mu) * sigma2, a_min=1e-6, a_max=None)
    return alpha * np.exp((x - mu) ** 2 / (-2 * (sigma ** 2)))


def _cmf(bands_angstrom):
    x = bands_angstrom
    cx = _g(x, 1.056, 5998, 379, 310) + _g(x, 0.362, 4420, 160, 267) + _g(x, -0.065, 5011, 204, 262)
    cy = _g(x, 0.821, 5688, 469, 405) + _g(x, 0.286, 5309, 163, 311)
    cz = _g(x, 1.217, 4370, 118, 360) + _g(x, 0.681, 4590, 260, 138)
    return np.array([cx, cy, cz])


def spec_to_rgb_matrix(bands) -> np.ndarray:
    """[B,3] spectrum -> linear sRGB matrix, columns normalised to sum 1
    (ref spec_to_rgb.py:73-90; sRGB primaries + D65 white, :34-38)."""
    bands = np.array(bands) * 10
    cmf = _cmf(bands)

    def xyz(x, y):
        return np.array((x, y, 1 - x - y))

    red, green, blue, white = xyz(0.64, 0.33), xyz(0.30, 0.60), xyz(0.15, 0.06), xyz(0.3127, 0.3291)
    M = np.vstack((red, green, blue)).T
    MI = np.linalg.inv(M)
    wscale = MI.dot(white)
    A = MI / wscale[:, np.newaxis]
    RGB = cmf.T @ A.T
    RGB = RGB / np.sum(RGB, axis=0, keepdims=True)
    return RGB


class ColourSystem(nn.Module):
    def __init__(self, bands):
        super().__init__()
        self.register_buffer("transform_matrix", torch.from_numpy(spec_to_rgb_matrix(bands)).float())

    @staticmethod
    def gamma_correction(x: Tensor) -> Tensor:
        return torch.where(x < 0.0031308, 12.92 * x, 1.055 * (x.clamp(min=1e-6).pow(1 / 2.4)) - 0.055)

    def forward(self, spec: Tensor) -> Tensor:
        rgb = torch.matmul(spec, self.transform_matrix)
        rgb = self.gamma_correction(rgb)
        return rgb.clamp(0, 1)


# ----------------------------------------------------------------------------
# Cluster probe  (ref umhsnerf/utils/clusterprobe.py:17-38)
# ----------------------------------------------------------------------------


def cluster_lookup(x: Tensor, alpha: Optional[float], clusters: Tensor):
    normed_clusters = F.normalize(clusters, dim=1)
    normed_features = F.normalize(x, dim=1)
    inner = torch.matmul(normed_features, normed_clusters.t())
    if alpha is None:
        probs = F.one_hot(torch.argmax(inner, dim=1), clusters.shape[0]).to(torch.float32)
    else:
        probs = F.softmax(inner * alpha, dim=1)
    return inner, probs


CLASS_COLORS = torch.tensor([  # ref umhs_model.py:146-162
    [0.49, 0.29, 0.95], [0.29, 0.95, 0.30], [0.95, 0.29, 0.47], [0.29, 0.66, 0.95], [0.86, 0.95, 0.29],
    [0.85, 0.29, 0.95], [0.29, 0.95, 0.66], [0.95, 0.46, 0.29], [0.29, 0.30, 0.95], [0.50, 0.95, 0.29],
    [0.95, 0.29, 0.69], [0.29, 0.88, 0.95], [0.95, 0.82, 0.29], [0.63, 0.29, 0.95], [0.29, 0.95, 0.43]])


# ----------------------------------------------------------------------------
# Renderer  (ref umhsnerf/umhs_renderer.py)
# ----------------------------------------------------------------------------


def spectral_render(spectral: Tensor, weights: Tensor, ray_indices: Optional[Tensor] = None,
                    num_rays: Optional[int] = None) -> Tensor:
    """SpectralRenderer.forward (ref umhs_renderer.py:15-30)."""
    if spectral.dim() == 3:
        spectral = spectral.squeeze(0)
    elif spectral.dim() == 1:
        spectral = spectral.unsqueeze(0)
    return nerf3p.accumulate_along_rays(weights[..., 0], spectral, ray_indices, num_rays)


def get_weights_spectral(deltas: Tensor, densities: Tensor) -> Tensor:
    """Dense-layout weights (ref umhs_renderer.py:117-139)."""
    delta_density = deltas * densities
    alphas = 1 - torch.exp(-delta_density)
    transmittance = torch.cumsum(delta_density[..., :-1, :], dim=-2)
    transmittance = torch.cat(
        [torch.zeros((*transmittance.shape[:1], 1, transmittance.shape[-1])), transmittance], dim=-2)
    transmittance = torch.exp(-transmittance)
    weights = alphas * transmittance
    return torch.nan_to_num(weights)


# ----------------------------------------------------------------------------
# Field  (ref umhsnerf/umhs_field.py)
# ----------------------------------------------------------------------------


class OracleField(nn.Module):
    """UMHSField on the torch path with the NerfactoField defaults it inherits (A.0):
    hash grid 16 levels x 2 feat, base MLP 32->64->16, geo_feat_dim 15, NeRF position
    encoding (2 freqs -> 12), SH degree 3 (16), appearance embedding width 0
    (quirk: ref umhs_model.py:135,181)."""

    def __init__(self, aabb: Tensor, wavelengths: int = 128, num_classes: int = 4, temperature: float = 0.5,
                 pred_specular: bool = False, log2_hashmap_size: int = 19, max_res: int = 2048,
                 scene_contraction: bool = True, num_layers_color: int = 3, hidden_dim_color: int = 64,
                 num_levels: int = 16, base_res: int = 16, features_per_level: int = 2, geo_feat_dim: int = 15):
        super().__init__()
        self.register_buffer("aabb", aabb)
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.geo_feat_dim = geo_feat_dim
        self.num_classes, self.wavelengths = num_classes, wavelengths
        self.temperature, self.pred_specular = temperature, pred_specular
        self.scene_contraction = scene_contraction
        self.average_init_density = 1
        self.pe_dim, self.sh_dim = 12, 16
        self.mlp_base = nerf3p.MLPWithHashEncoding(num_levels, base_res, max_res, log2_hashmap_size,
                                                   features_per_level, 2, 64, 1 + geo_feat_dim)
        in_dim = self.pe_dim + geo_feat_dim  # ref umhs_field.py:64
        out_feat = num_classes + 1 if pred_specular else num_classes  # :65
        self.feature_mlp = nerf3p.MLP(in_dim, 3, hidden_dim_color, out_feat, nn.ReLU(), None)  # :66-74
        self.endmembers = nn.Parameter(torch.randn(num_classes, wavelengths))  # :77-85
        self.mlp_head = nerf3p.MLP(in_dim, num_layers_color, hidden_dim_color, num_classes, nn.ReLU(), None)  # :95
        self.mlp_directional = nerf3p.MLP(self.sh_dim + self.pe_dim, 2, 16, wavelengths, nn.ReLU(), nn.Sigmoid())

    # ref umhs_field.py:300-329
    def get_density(self, ray_samples: RaySamples):
        if self.scene_contraction:
            positions = ray_samples.frustums.get_positions()
            positions = nerf3p.scene_contraction_inf(positions)
            positions = (positions + 2.0) / 4.0
        else:
            positions = nerf3p.scene_box_normalize(ray_samples.frustums.get_positions(), self.aabb)
        selector = ((positions > 0.0) & (positions < 1.0)).all(dim=-1)
        positions = positions * selector[..., None]
        positions_flat = positions.view(-1, 3)
        h = self.mlp_base(positions_flat).view(*ray_samples.frustums.shape, -1)
        density_before_activation, base_mlp_out = torch.split(h, [1, self.geo_feat_dim], dim=-1)
        density = self.average_init_density * nerf3p.trunc_exp(density_before_activation.to(positions))
        density = density * selector[..., None]
        return density, base_mlp_out

    # ref umhs_field.py:151-261 (spectral branch; embedded_appearance is None since A=0)
    def get_outputs(self, ray_samples: RaySamples, density_embedding: Tensor) -> Dict[str, Tensor]:
        outputs = {}
        directions = nerf3p.get_normalized_directions(ray_samples.frustums.directions)
        directions_flat = directions.view(-1, 3)
        d = nerf3p.sh_encoding_deg3(directions_flat)
        outputs_shape = ray_samples.frustums.directions.shape[:-1]

        positions = ray_samples.frustums.get_positions()  # raw world positions (quirk 3)
        positions_flat = nerf3p.nerf_encoding(positions.view(-1, 3))
        positions_flat = positions_flat.unsqueeze(0)  # always 2-D at this point (:187-188)
        if len(density_embedding.shape) == 2:
            density_embedding = density_embedding.unsqueeze(0)
        h1 = torch.cat([positions_flat.view(-1, self.pe_dim), density_embedding.view(-1, self.geo_feat_dim)], dim=-1)
        scalar = self.mlp_head(h1).view(*outputs_shape, -1, self.num_classes)
        scalar = torch.sigmoid(scalar)
        features_input = torch.cat([positions_flat, density_embedding], dim=-1)
        size = features_input.size()
        features_input = features_input.view(-1, features_input.size(-1))
        features = self.feature_mlp(features_input)
        logits = features.view(*size[:-1], -1)
        if self.pred_specular:
            logits, s1 = torch.split(logits, [self.num_classes, 1], dim=-1)
            s1 = torch.sigmoid(s1)
        abundances = F.softmax(logits / self.temperature, dim=-1)
        endmembers = self.endmembers.unsqueeze(0).unsqueeze(0)
        endmembers = endmembers.expand(abundances.shape[0], abundances.shape[1], -1, -1).transpose(2, 3).squeeze(0)
        adapted = scalar * endmembers  # [N, B, K]
        spec = (adapted @ abundances.unsqueeze(-1)).squeeze()
        if self.pred_specular:
            input_spec = torch.cat([d, positions_flat.view(-1, self.pe_dim)], dim=-1)
            specular = self.mlp_directional(input_spec).view(*outputs_shape, self.wavelengths)
            spec2 = spec + (s1 * specular)
            outputs["spectral"] = spec2
            outputs["spectral2"] = spec
            with torch.no_grad():
                outputs["specular"] = (s1 * specular)
        else:
            outputs["spectral"] = spec
        outputs["abundances"] = abundances
        return outputs

    def forward(self, ray_samples: RaySamples) -> Dict[str, Tensor]:
        density, emb = self.get_density(ray_samples)
        out = self.get_outputs(ray_samples, density_embedding=emb)
        out["density"] = density
        return out

    def density_fn(self, positions: Tensor) -> Tensor:
        rs = RaySamples(frustums=Frustums(origins=positions, directions=torch.ones_like(positions),
                                          starts=torch.zeros_like(positions[..., :1]),
                                          ends=torch.zeros_like(positions[..., :1]),
                                          pixel_area=torch.ones_like(positions[..., :1])))
        density, _ = self.get_density(rs)
        return density


# ----------------------------------------------------------------------------
# Model  (ref umhsnerf/umhs_model.py)
# ----------------------------------------------------------------------------


@dataclass
class OracleConfig:
    """The UMHSConfig fields the hot path reads (ref umhs_model.py:61-119)."""
    grid_resolution: int = 128
    grid_levels: int = 4
    max_res: int = 2048
    log2_hashmap_size: int = 19
    alpha_thre: float = 0.01
    cone_angle: float = 0.004
    render_step_size: Optional[float] = None
    near_plane: float = 0.05
    far_plane: float = 1e3
    use_gradient_scaling: bool = True
    background_color: str = "random"
    disable_scene_contraction: bool = False
    method: str = "rgb+spectral"
    rgb_loss_weight: float = 1.0
    temperature: float = 0.2
    pred_specular: bool = False


class OracleModel(nn.Module):
    def __init__(self, config: OracleConfig, scene_aabb: Tensor, wavelengths: List[int], num_classes: int):
        super().__init__()
        self.config = config
        self.converter = ColourSystem(wavelengths)
        self.num_classes = num_classes
        self.field = OracleField(aabb=scene_aabb, wavelengths=len(wavelengths), num_classes=num_classes,
                                 temperature=config.temperature, pred_specular=config.pred_specular,
                                 log2_hashmap_size=config.log2_hashmap_size, max_res=config.max_res,
                                 scene_contraction=not config.disable_scene_contraction)
        self.scene_aabb = nn.Parameter(scene_aabb.flatten(), requires_grad=False)
        if config.render_step_size is None:  # ref umhs_model.py:199-200
            config.render_step_size = ((self.scene_aabb[3:] - self.scene_aabb[:3]) ** 2).sum().sqrt().item() / 1000
        self.occupancy_grid = nerf3p.OccGridEstimator(self.scene_aabb, config.grid_resolution, config.grid_levels)
        self.sampler = nerf3p.VolumetricSampler(self.occupancy_grid, self.field.density_fn)

    def train(self, mode: bool = True):
        super().train(mode)
        self.sampler.training = mode
        return self

    # ref umhs_model.py:225-313
    def get_outputs(self, ray_bundle: RayBundle, jitter: Optional[Tensor] = None) -> Dict[str, Tensor]:
        cfg = self.config
        num_rays = len(ray_bundle)
        ray_bundle = nerf3p.near_far_collider(ray_bundle, cfg.near_plane, cfg.far_plane, self.training)
        with torch.no_grad():
            ray_samples, ray_indices = self.sampler(
                ray_bundle=ray_bundle, near_plane=cfg.near_plane, far_plane=cfg.far_plane,
                render_step_size=cfg.render_step_size, alpha_thre=cfg.alpha_thre, cone_angle=cfg.cone_angle,
                jitter=jitter)
        return self.render_samples(ray_samples, ray_indices, num_rays)

    def render_samples(self, ray_samples: RaySamples, ray_indices: Tensor, num_rays: int) -> Dict[str, Tensor]:
        cfg = self.config
        field_outputs = self.field(ray_samples)
        if cfg.use_gradient_scaling:
            field_outputs = nerf3p.scale_gradients_by_distance_squared(field_outputs, ray_samples)
        packed_info = nerf3p.pack_info(ray_indices, num_rays)
        weights = nerf3p.render_weight_from_density(
            t_starts=ray_samples.frustums.starts[..., 0], t_ends=ray_samples.frustums.ends[..., 0],
            sigmas=field_outputs["density"][..., 0], packed_info=packed_info)[0]
        weights = weights[..., None]
        depth = nerf3p.render_depth_expected(weights, ray_samples, ray_indices, num_rays)
        accumulation = nerf3p.render_accumulation(weights, ray_indices, num_rays)
        outputs = {"accumulation": accumulation, "depth": depth, "weights": weights, "ray_indices": ray_indices,
                   "t_starts": ray_samples.frustums.starts[..., 0], "t_ends": ray_samples.frustums.ends[..., 0]}
        spectral = spectral_render(field_outputs["spectral"], weights, ray_indices, num_rays)
        outputs["spectral"] = spectral
        if cfg.pred_specular:
            outputs["spectral2"] = spectral_render(field_outputs["spectral2"], weights, ray_indices, num_rays)
            with torch.no_grad():
                outputs["specular"] = spectral_render(field_outputs["specular"], weights, ray_indices, num_rays)
        if cfg.method == "spectral":
            with torch.no_grad():
                outputs["rgb"] = self.converter(spectral)
        else:
            outputs["rgb"] = self.converter(spectral)
        outputs["num_samples_per_ray"] = packed_info[:, 1]
        with torch.no_grad():
            outputs["abundances"] = spectral_render(field_outputs["abundances"], weights, ray_indices, num_rays)
        inner, probs = cluster_lookup(spectral, alpha=0.2, clusters=self.field.endmembers)
        acc_if = torch.where(accumulation > 0.5, torch.tensor(1.0), torch.tensor(0.0))
        outputs["seg_probs"] = probs
        with torch.no_grad():
            outputs["seg_raw"] = probs.argmax(1) * acc_if.squeeze()
            outputs["seg_pred"] = CLASS_COLORS[probs.argmax(1).long()] * acc_if
        return outputs

    # ref umhs_model.py:329-383 (method "rgb+spectral" / "spectral")
    def get_loss_dict(self, outputs, batch, bg: Optional[Tensor] = None) -> Dict[str, Tensor]:
        cfg = self.config
        image = batch["image"]
        gt_spectral = batch["hs_image"]
        pred_rgb, gt_rgb = nerf3p.blend_background_for_loss(outputs["rgb"], outputs["accumulation"], image,
                                                            cfg.background_color, bg)
        loss = {}
        if cfg.method == "spectral":
            loss["spectral_loss"] = F.mse_loss(outputs["spectral"], gt_spectral)
        elif cfg.method == "rgb+spectral":
            loss["spectral_loss"] = 5 * F.mse_loss(outputs["spectral"], gt_spectral)  # hard-coded 5 (quirk 4)
            loss["rgb_loss"] = cfg.rgb_loss_weight * F.mse_loss(pred_rgb, gt_rgb)
        return loss

    # ref umhs_model.py:385-405
    def get_metrics_dict(self, outputs, batch) -> Dict[str, Tensor]:
        m = {}
        gt_rgb = batch["image"][..., :3]
        m["psnr"] = nerf3p.psnr(outputs["rgb"], gt_rgb)
        m["rmse"] = torch.sqrt(F.mse_loss(outputs["rgb"], gt_rgb)).item()
        m["psnr_spectral"] = nerf3p.psnr(outputs["spectral"], batch["hs_image"])
        m["rmse_spectral"] = torch.sqrt(F.mse_loss(outputs["spectral"], batch["hs_image"])).item()
        m["num_samples_per_batch"] = outputs["num_samples_per_ray"].sum()
        return m

    # ref umhs_model.py:549-554
    def update_occupancy_grid(self, step: int, **kw):
        self.occupancy_grid.update_every_n_steps(
            step=step, occ_eval_fn=lambda x: self.field.density_fn(x) * self.config.render_step_size, **kw)

    # ref umhs_model.py:568-570
    @torch.no_grad()
    def clamp_endmembers(self):
        self.field.endmembers[:] = self.field.endmembers.clamp(0, 1)


# ----------------------------------------------------------------------------
# BASELINE config 1: dense [R,S] layout (the reference's CPU-capable functions)
# ----------------------------------------------------------------------------


def dense_forward(model: OracleModel, origins: Tensor, directions: Tensor, t_starts: Tensor, t_ends: Tensor):
    """Field on the flattened [R*S] samples (the reference field only supports the packed
    layout: ref umhs_field.py:187-217 concatenates a [1,N,12] with the embedding), then the
    reference's dense weight function (umhs_renderer.py:117-139) and the dense
    ``sum(dim=-2)`` accumulate (umhs_renderer.py:28 with ray_indices=None)."""
    R, S = t_starts.shape
    rs = RaySamples(frustums=Frustums(origins=origins[:, None, :].expand(R, S, 3).reshape(-1, 3),
                                      directions=directions[:, None, :].expand(R, S, 3).reshape(-1, 3),
                                      starts=t_starts.reshape(-1, 1), ends=t_ends.reshape(-1, 1)))
    fo = model.field(rs)
    if model.config.use_gradient_scaling:
        fo = nerf3p.scale_gradients_by_distance_squared(fo, rs)
    deltas = (t_ends - t_starts)[..., None]
    weights = get_weights_spectral(deltas, fo["density"].view(R, S, 1))
    spec = fo["spectral"]
    spec = spec.squeeze(0) if spec.dim() == 3 else spec
    spectral = spectral_render(spec.view(R, S, -1), weights)
    accumulation = torch.sum(weights, dim=-2)
    rgb = model.converter(spectral)
    return {"spectral": spectral, "accumulation": accumulation, "rgb": rgb, "weights": weights}
