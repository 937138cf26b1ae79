"""Import shim that lets the UNMODIFIED reference modules under /root/reference run in
a container where their third-party dependencies (nerfstudio, nerfacc, tinycudann,
torchmetrics, matplotlib) are absent.

Used ONLY by ``make_golden.py`` (run once, in the build container, to freeze golden
vectors).  It never runs on the GPU box and is not imported by tests or the product.

The shim provides the third-party symbols the reference imports, implemented with
``oracle/nerf3p.py`` (the documented restatement of nerfstudio 1.1.5 / nerfacc 0.5.2).
The reference's OWN code -- ``UMHSField.get_density/get_outputs``, ``SpectralRenderer``,
``get_weights_spectral``, ``ColourSystem``, ``ClusterLookup``, ``UMHSModel.get_outputs``
and ``get_loss_dict`` -- then executes unmodified, which is what pins ``oracle/umhs_oracle.py``.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, Optional, Type

import torch
from torch import nn

sys.path.insert(0, "/root/repo")
from oracle import nerf3p  # noqa: E402

_STUB_ROOTS = ("nerfstudio", "nerfacc", "tinycudann", "torchmetrics", "matplotlib", "wandb")


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None


class _AutoModule(types.ModuleType):
    """Any attribute that was not set explicitly resolves to a throw-away class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _AutoModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        _populate(module)


# ---------------------------------------------------------------- nerfstudio stand-ins


class FieldHeadNames(Enum):
    RGB = "rgb"
    DENSITY = "density"


class NeRFEncoding(nn.Module):
    def __init__(self, in_dim, num_frequencies, min_freq_exp, max_freq_exp, include_input=False, implementation="torch"):
        super().__init__()
        self.a = (in_dim, num_frequencies, float(min_freq_exp), float(max_freq_exp))

    def get_out_dim(self):
        return self.a[0] * self.a[1] * 2

    def forward(self, x):
        return nerf3p.nerf_encoding(x, self.a[1], self.a[2], self.a[3])


class SHEncoding(nn.Module):
    def __init__(self, levels=4, implementation="torch"):
        super().__init__()
        assert levels == 4

    def get_out_dim(self):
        return 16

    @torch.no_grad()
    def forward(self, x):
        return nerf3p.sh_encoding_deg3(x)


class MLP(nerf3p.MLP):
    def __init__(self, in_dim, num_layers, layer_width, out_dim, activation=nn.ReLU(), out_activation=None,
                 implementation="torch", **kw):
        super().__init__(in_dim, num_layers, layer_width, out_dim, activation, out_activation)


class SceneContraction(nn.Module):
    def __init__(self, order=None):
        super().__init__()
        assert order == float("inf")

    def forward(self, x):
        return nerf3p.scene_contraction_inf(x)


class SceneBox:
    def __init__(self, aabb):
        self.aabb = aabb

    @staticmethod
    def get_normalized_positions(positions, aabb):
        return nerf3p.scene_box_normalize(positions, aabb)


class Field(nn.Module):
    def density_fn(self, positions):
        rs = nerf3p.RaySamples(frustums=nerf3p.Frustums(
            origins=positions, directions=torch.ones_like(positions), starts=torch.zeros_like(positions[..., :1]),
            ends=torch.zeros_like(positions[..., :1]), pixel_area=torch.ones_like(positions[..., :1])))
        density, _ = self.get_density(rs)
        return density

    def forward(self, ray_samples, compute_normals=False):
        density, emb = self.get_density(ray_samples)
        out = self.get_outputs(ray_samples, density_embedding=emb)
        out[FieldHeadNames.DENSITY] = density
        return out


class NerfactoField(Field):
    """A.0: the NerfactoField constructor state UMHSField relies on."""

    def __init__(self, aabb, num_images, num_layers=2, hidden_dim=64, geo_feat_dim=15, num_levels=16, base_res=16,
                 max_res=2048, log2_hashmap_size=19, features_per_level=2, appearance_embedding_dim=32,
                 spatial_distortion=None, implementation="tcnn", average_init_density=1.0, **kw):
        super().__init__()
        assert implementation == "torch", "the parity target is the reference's torch path"
        self.register_buffer("aabb", aabb)
        self.geo_feat_dim = geo_feat_dim
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.spatial_distortion = spatial_distortion
        self.num_images = num_images
        self.appearance_embedding_dim = appearance_embedding_dim
        self.embedding_appearance = None
        assert appearance_embedding_dim == 0
        self.use_average_appearance_embedding = True
        self.average_init_density = average_init_density
        self.direction_encoding = SHEncoding(levels=4)
        self.position_encoding = NeRFEncoding(in_dim=3, num_frequencies=2, min_freq_exp=0, max_freq_exp=2 - 1)
        self.mlp_base = nerf3p.MLPWithHashEncoding(num_levels, base_res, max_res, log2_hashmap_size,
                                                   features_per_level, num_layers, hidden_dim, 1 + geo_feat_dim)


class NearFarCollider(nn.Module):
    def __init__(self, near_plane, far_plane, reset_near_plane=True):
        super().__init__()
        self.near_plane, self.far_plane = near_plane, far_plane

    def forward(self, ray_bundle):
        return nerf3p.near_far_collider(ray_bundle, self.near_plane, self.far_plane, self.training)


class VolumetricSampler(nn.Module):
    def __init__(self, occupancy_grid, density_fn=None):
        super().__init__()
        self.inner = nerf3p.VolumetricSampler(occupancy_grid, density_fn)
        self.jitter_override = None

    def forward(self, ray_bundle, render_step_size, near_plane=0.0, far_plane=None, alpha_thre=0.01, cone_angle=0.0):
        self.inner.training = self.training
        return self.inner(ray_bundle, render_step_size, near_plane, far_plane, alpha_thre, cone_angle,
                          jitter=self.jitter_override)


class RGBRenderer(nn.Module):
    def __init__(self, background_color="random"):
        super().__init__()
        self.background_color = background_color
        self.bg_override = None

    def blend_background_for_loss_computation(self, pred_image, pred_accumulation, gt_image):
        return nerf3p.blend_background_for_loss(pred_image, pred_accumulation, gt_image, self.background_color,
                                                self.bg_override)

    def blend_background(self, image, background_color=None):
        return image[..., :3]


class AccumulationRenderer(nn.Module):
    def forward(self, weights, ray_indices=None, num_rays=None):
        return nerf3p.render_accumulation(weights, ray_indices, num_rays)


class DepthRenderer(nn.Module):
    def __init__(self, method="expected"):
        super().__init__()

    def forward(self, weights, ray_samples, ray_indices=None, num_rays=None):
        return nerf3p.render_depth_expected(weights, ray_samples, ray_indices, num_rays)


class SemanticRenderer(nn.Module):
    pass


@dataclass
class InstantNGPModelConfig:
    _target: Type = None
    eval_num_rays_per_chunk: int = 512

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


class NGPModel(nn.Module):
    def __init__(self, config, scene_box, num_train_data, **kwargs):
        super().__init__()
        self.config = config
        self.scene_box = scene_box
        self.num_train_data = num_train_data
        self.kwargs = kwargs
        self.collider = None
        self.populate_modules()

    def populate_modules(self):
        self.renderer_rgb = RGBRenderer(background_color=self.config.background_color)
        self.renderer_accumulation = AccumulationRenderer()
        self.renderer_depth = DepthRenderer(method="expected")
        self.psnr = nerf3p.psnr

    @property
    def device(self):
        return torch.device("cpu")


class _PsnrLike(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


def to_immutable_dict(d: Dict[str, Any]):
    return tuple(d.items())


def _populate(m: types.ModuleType) -> None:
    n = m.__name__
    table = {
        "nerfstudio.fields.nerfacto_field": dict(NerfactoField=NerfactoField),
        "nerfstudio.cameras.rays": dict(RaySamples=nerf3p.RaySamples, RayBundle=nerf3p.RayBundle,
                                         Frustums=nerf3p.Frustums),
        "nerfstudio.field_components.mlp": dict(MLP=MLP, MLPWithHashEncoding=nerf3p.MLPWithHashEncoding),
        "nerfstudio.fields.base_field": dict(get_normalized_directions=nerf3p.get_normalized_directions, Field=Field),
        "nerfstudio.field_components.field_heads": dict(FieldHeadNames=FieldHeadNames),
        "nerfstudio.field_components.activations": dict(trunc_exp=nerf3p.trunc_exp),
        "nerfstudio.field_components.encodings": dict(NeRFEncoding=NeRFEncoding, SHEncoding=SHEncoding),
        "nerfstudio.field_components.spatial_distortions": dict(SceneContraction=SceneContraction),
        "nerfstudio.data.scene_box": dict(SceneBox=SceneBox),
        "nerfstudio.model_components.renderers": dict(SemanticRenderer=SemanticRenderer),
        "nerfstudio.models.instant_ngp": dict(NGPModel=NGPModel, InstantNGPModelConfig=InstantNGPModelConfig),
        "nerfstudio.model_components.scene_colliders": dict(NearFarCollider=NearFarCollider),
        "nerfstudio.model_components.ray_samplers": dict(VolumetricSampler=VolumetricSampler),
        "nerfstudio.model_components.losses": dict(
            MSELoss=nn.MSELoss, scale_gradients_by_distance_squared=nerf3p.scale_gradients_by_distance_squared),
        "nerfstudio.configs.config_utils": dict(to_immutable_dict=to_immutable_dict),
        "nerfacc": dict(OccGridEstimator=nerf3p.OccGridEstimator, pack_info=nerf3p.pack_info,
                        render_weight_from_density=nerf3p.render_weight_from_density,
                        accumulate_along_rays=nerf3p.accumulate_along_rays),
        "torchmetrics.image": dict(SpectralAngleMapper=_PsnrLike),
    }
    for k, v in table.get(n, {}).items():
        setattr(m, k, v)


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
