"""UMHSModel: the UMHS nerfstudio model, B200-native hot path.

Drop-in for ``UMHSModel`` / ``UMHSConfig`` of ref umhsnerf/umhs_model.py:61-620 for methods "rgb+spectral" and
"spectral": same config fields, the same ``get_outputs`` output dictionary, ``get_loss_dict``, ``get_metrics_dict``,
``get_training_callbacks`` and chunk-free ``get_outputs_for_camera_ray_bundle``.  One training step launches:

  march_count -> pack_scan -> march_write                      (csrc/packing.cu)
  hash_encode_fwd + density_fwd -> visibility -> compaction    (sampler's sigma_fn pass, no grad)
  hash_encode_fwd + field_mlp_fwd                              (csrc/hash_grid.cu, csrc/field_mlp*.cu)
  composite_fwd (+ depth clip)                                 (csrc/composite.cu)
  [loss on the [R,B] / [R,3] outputs]
  composite_bwd -> field_mlp_bwd -> hash_encode_bwd

When nerfstudio is installed the class registers with it through ``umhs_config.py``; the arithmetic never depends on it.
"""
from __future__ import annotations

from dataclasses import dataclass, field as dc_field
from typing import Dict, List, Literal, Optional, Tuple, Type, Union

import os

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import ops
from .occ_grid import OccGridEstimator, VolumetricSampler
from .rays import RayBundle
from .spec_to_rgb import ColourSystem
from .clusterprobe import ClusterLookup
from .umhs_field import UMHSField
from .umhs_renderer import SpectralRenderer

# ref umhs_model.py:146-162
CLASS_COLORS = [
    [0.49, 0.29, 0.95], [0.29, 0.95, 0.30], [0.95, 0.29, 0.47], [0.29, 0.66, 0.95], [0.86, 0.95, 0.29],
    [0.85, 0.29, 0.95], [0.29, 0.95, 0.66], [0.95, 0.46, 0.29], [0.29, 0.30, 0.95], [0.50, 0.95, 0.29],
    [0.95, 0.29, 0.69], [0.29, 0.88, 0.95], [0.95, 0.82, 0.29], [0.63, 0.29, 0.95], [0.29, 0.95, 0.43],
]


# nerfstudio's own base classes when it is installed (the plugin is then a real ``InstantNGPModelConfig`` /
# ``NGPModel`` pair: tyro's ``--pipeline.model.*`` flags and the pipeline's isinstance paths see what they expect, ref
# umhs_model.py:62,122); plain stand-ins otherwise (nerfstudio is absent from the build container).
try:  # pragma: no cover - exercised with the stub package of tests/test_plugin_cpu.py
    from nerfstudio.models.instant_ngp import InstantNGPModelConfig as _NSConfigBase, NGPModel as _NSModelBase

    HAVE_NERFSTUDIO = True
except Exception:  # ModuleNotFoundError or a broken install
    _NSConfigBase, _NSModelBase, HAVE_NERFSTUDIO = None, None, False


@dataclass
class _UMHSFields:
    """Fields of ref umhs_model.py:61-119 (+ the InstantNGPModelConfig / ModelConfig fields the hot path reads)."""
    _target: Type = dc_field(default_factory=lambda: UMHSModel)
    enable_collider: bool = False
    collider_params: Optional[Dict[str, float]] = None
    grid_resolution: Union[int, List[int]] = 128
    grid_levels: int = 4
    max_res: int = 2048
    log2_hashmap_size: int = 19
    alpha_thre: float = 0.01
    cone_angle: float = 0.004
    render_step_size: Optional[float] = None
    near_plane: float = 0.05
    far_plane: float = 1e3
    use_gradient_scaling: bool = True
    use_appearance_embedding: bool = True
    background_color: Literal["random", "black", "white"] = "random"
    disable_scene_contraction: bool = False
    implementation: Literal["tcnn", "torch", "b200"] = "b200"
    method: Literal["rgb", "spectral", "rgb+spectral"] = "rgb+spectral"
    rgb_loss_weight: float = 1.0
    spectral_loss_weight: float = 1.0  # ignored, as in the reference (hard-coded 5: ref umhs_model.py:369)
    temperature: float = 0.2
    pred_dino: bool = False
    pred_specular: bool = False
    load_vca: bool = False
    eval_num_rays_per_chunk: int = 512  # ref umhs_config.py:50 (the fused eval path renders whole tiles instead)
    # rays per fused eval pass: a whole 800x600 image is one pass (~25 M samples, ~6 GB of transients); measured
    # 12.7 ms per image at 2^16 rays per pass, 10.7 ms at 2^18
    eval_tile_rays: int = 1 << 19


if HAVE_NERFSTUDIO:  # pragma: no cover
    @dataclass
    class UMHSConfig(_UMHSFields, _NSConfigBase):  # type: ignore[misc, valid-type]
        """UMHS model config (a nerfstudio ``InstantNGPModelConfig``)."""

    _ModelBase = _NSModelBase
else:
    @dataclass
    class UMHSConfig(_UMHSFields):  # type: ignore[no-redef]
        """UMHS model config (stand-alone: nerfstudio is not installed)."""

        def setup(self, **kwargs) -> "UMHSModel":
            return self._target(self, **kwargs)

    _ModelBase = nn.Module


class _SceneBox:
    def __init__(self, aabb: Tensor):
        self.aabb = aabb


class NearFarCollider(nn.Module):
    """nerfstudio NearFarCollider (A.4): nears = near_plane while training, 0 in eval (reset_near_plane); fars =
    far_plane.  The constant planes are cached per ray-batch shape (read-only downstream): no fill launches per step."""

    def __init__(self, near_plane: float, far_plane: float):
        super().__init__()
        self.near_plane, self.far_plane = float(near_plane), float(far_plane)
        self._cache: Dict = {}

    def forward(self, ray_bundle: RayBundle) -> RayBundle:
        o = ray_bundle.origins
        near = self.near_plane if self.training else 0.0
        key = (tuple(o.shape[:-1]), o.device, near, self.far_plane)
        if key not in self._cache:
            if len(self._cache) > 8:
                self._cache.clear()
            ones = torch.ones_like(o[..., 0:1])
            self._cache[key] = (ones * near, ones * self.far_plane)
        ray_bundle.nears, ray_bundle.fars = self._cache[key]
        return ray_bundle


class UMHSModel(_ModelBase):  # type: ignore[misc, valid-type]
    config: UMHSConfig

    def __init__(self, config: UMHSConfig, scene_box=None, num_train_data: int = 0, metadata: Optional[dict] = None,
                 **kwargs):
        if scene_box is None:
            scene_box = _SceneBox(torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]]))
        elif isinstance(scene_box, Tensor):
            scene_box = _SceneBox(scene_box)
        if HAVE_NERFSTUDIO:  # pragma: no cover - nerfstudio's Model.__init__ stores these and calls populate_modules
            super().__init__(config=config, scene_box=scene_box, num_train_data=num_train_data,
                             metadata=metadata or {}, **kwargs)
            return
        super().__init__()
        self.config = config
        self.scene_box = scene_box
        self.num_train_data = num_train_data
        self.kwargs = {"metadata": metadata or {}, **kwargs}
        self.device_indicator_param = nn.Parameter(torch.empty(0))
        self.populate_modules()

    if not HAVE_NERFSTUDIO:
        @property
        def device(self):
            return self.device_indicator_param.device

    # ref umhs_model.py:127-218
    def populate_modules(self):
        cfg = self.config
        if "spectral" not in cfg.method:
            raise NotImplementedError("method='rgb' is broken in the reference (ref umhs_model.py:266) and not built")
        self.step = 0
        self.i = 0
        self.kwargs = self.kwargs["metadata"]
        wl = list(self.kwargs["wavelengths"])
        K = int(self.kwargs["num_classes"])
        self.converter = ColourSystem(bands=wl, cs="sRGB")
        self.renderer_spectral = SpectralRenderer()
        self.register_buffer("class_colors", torch.tensor(CLASS_COLORS, dtype=torch.float32), persistent=False)
        self.field = UMHSField(
            aabb=self.scene_box.aabb,
            appearance_embedding_dim=0 if cfg.use_appearance_embedding else 32,  # quirk 1 (ref :135,181)
            num_images=self.num_train_data, log2_hashmap_size=cfg.log2_hashmap_size, max_res=cfg.max_res,
            spatial_distortion=None if cfg.disable_scene_contraction else "contraction",
            implementation=cfg.implementation, method=cfg.method, wavelengths=len(wl), num_classes=K,
            temperature=cfg.temperature, converter=self.converter, pred_dino=cfg.pred_dino,
            pred_specular=cfg.pred_specular, load_vca=cfg.load_vca)
        self.scene_aabb = nn.Parameter(torch.as_tensor(self.scene_box.aabb, dtype=torch.float32).flatten(),
                                       requires_grad=False)
        if cfg.render_step_size is None:
            cfg.render_step_size = ((self.scene_aabb[3:] - self.scene_aabb[:3]) ** 2).sum().sqrt().item() / 1000
        res = cfg.grid_resolution if isinstance(cfg.grid_resolution, int) else int(cfg.grid_resolution[0])
        self.occupancy_grid = OccGridEstimator(roi_aabb=self.scene_aabb, resolution=res, levels=cfg.grid_levels)
        self.sampler = VolumetricSampler(occupancy_grid=self.occupancy_grid, field=self.field)
        self.cluster_probe = ClusterLookup(len(wl), K)
        self.collider = NearFarCollider(near_plane=cfg.near_plane, far_plane=cfg.far_plane)  # ref umhs_model.py:212
        try:  # the pretrained LPIPS network lives in torchmetrics (ref: NGPModel.populate_modules); optional here
            from torchmetrics.image.lpip import LearnedPerceptualImagePatchSimilarity  # type: ignore

            self.lpips = LearnedPerceptualImagePatchSimilarity(normalize=True)
        except Exception:
            self.lpips = None
        self._bg_override: Optional[Tensor] = None  # test hook: fixes the "random" background draw
        # test hook: fixes the stratified near-plane draw of the sampler ([R] in [0,1); nerfacc: rand_like(near_planes));
        # a persistent device tensor, so a captured step graph sees what a test writes into it between replays
        self._jitter_override: Optional[Tensor] = None
        # None: exact-sized sample buffers (host sync on the sample count, reference behaviour).
        # int: sync-free steps on capacity-sized buffers (set by UMHSTrainer; see VolumetricSampler).
        self.sample_capacity: Optional[int] = None

    def forward(self, ray_bundle: RayBundle) -> Dict[str, Tensor]:
        return self.get_outputs(self.collider(ray_bundle))

    def label_to_rgb(self, labels: Tensor) -> Tensor:
        return self.class_colors[labels.long().squeeze(-1)]

    # ------------------------------------------------------------------ ref umhs_model.py:225-327
    @torch.no_grad()
    def premarch(self, ray_bundle: RayBundle, jitter: Optional[Tensor] = None):
        """Occupancy march of a (collided) ray bundle: the parameter-independent first stage of ``get_outputs``, for
        trainers that run it one batch ahead (pass the result back as ``get_outputs(..., premarched=...)``)."""
        cfg = self.config
        if ray_bundle.nears is None:
            ray_bundle = self.collider(ray_bundle)
        if jitter is None:
            jitter = self._jitter_override
        return self.sampler.march_only(
            ray_bundle, render_step_size=cfg.render_step_size, near_plane=cfg.near_plane, far_plane=cfg.far_plane,
            cone_angle=cfg.cone_angle, jitter=jitter, capacity=self.sample_capacity if self.training else None)

    def get_outputs(self, ray_bundle: RayBundle, jitter: Optional[Tensor] = None,
                    per_band_views: bool = True, premarched=None, lean: bool = False,
                    depth_chunks=None) -> Dict[str, Tensor]:
        """lean=True (the standalone trainer): no "depth" entry and no per-sample weights by-product -- nothing in
        the training losses reads them (ref umhs_model.py:329-383) -- which saves three launches per step.
        depth_chunks=(chunk_rays, ray_ids | None, n_image_rays): the full-image drivers render many of the reference's
        512-ray forward calls in one pass; the depth clip then uses each call's own bounds (ops.CompositeFunction)."""
        cfg = self.config
        num_rays = int(ray_bundle.origins.reshape(-1, 3).shape[0])
        if ray_bundle.nears is None:
            ray_bundle = self.collider(ray_bundle)
        if jitter is None:
            jitter = self._jitter_override
        pts, ray_indices, t0, t1, packed_info = self.sampler(
            ray_bundle=ray_bundle, near_plane=cfg.near_plane, far_plane=cfg.far_plane,
            render_step_size=cfg.render_step_size, alpha_thre=cfg.alpha_thre, cone_angle=cfg.cone_angle,
            jitter=jitter, capacity=self.sample_capacity if self.training else None,  # eval: exact-sized packing
            premarched=premarched)
        fo = self.field.forward_packed(pts, want_embedding=False)
        K, B, spec = self.field.num_classes, self.field.wavelengths, cfg.pred_specular
        d = self.field.mlp_directional.layers
        ccfg = {"K": K, "B": B, "specular": spec, "grad_scaling": cfg.use_gradient_scaling, "want_weights": not lean,
                "want_depth": not lean, "depth_chunks": depth_chunks,
                # for the tensor-core specular kernels (sample -> ray map, device-side sample count)
                "ray_indices": ray_indices, "n_dev": pts.n_dev}
        (spectral, spectral2, specular, abund, acc, depth, rgb, seg_probs, seg_raw, seg_pred,
         weights) = ops.CompositeFunction.apply(
            ccfg, packed_info, t0, t1, fo["sigma"], fo["scalar"], fo["abund"], fo.get("s1"), fo.get("spec_hidden"),
            self.field.endmembers, d[1].weight if spec else None, d[1].bias if spec else None,
            self.converter.transform_matrix, self.class_colors)
        outputs: Dict[str, Tensor] = {"accumulation": acc.unsqueeze(-1), "spectral": spectral}
        if not lean:
            outputs["depth"] = depth.unsqueeze(-1)
        if per_band_views:
            for i in range(B):
                outputs[f"wv_{i}"] = spectral[..., i]
        if spec:
            outputs["spectral2"] = spectral2
            outputs["specular"] = specular
            if per_band_views:
                for i in range(B):
                    outputs[f"residual_{i}"] = specular[..., i]
        outputs["rgb"] = rgb.detach() if cfg.method == "spectral" else rgb
        outputs["num_samples_per_ray"] = packed_info[:, 1]
        outputs["abundances"] = abund
        if per_band_views:
            for i in range(K):
                outputs[f"abundances_{i}"] = abund[:, i]
        outputs["seg_probs"] = seg_probs
        outputs["seg_raw"] = seg_raw
        outputs["seg_pred"] = seg_pred
        # sample-level by-products (not in the reference dict; used by tests and the bench)
        self._last = {"ray_indices": ray_indices, "t_starts": t0, "t_ends": t1, "weights": weights,
                      "packed_info": packed_info}
        return outputs

    # ------------------------------------------------------------------ ref umhs_model.py:329-383
    def _blend_background(self, pred: Tensor, acc: Tensor, gt: Tensor) -> Tuple[Tensor, Tensor]:
        color = self.config.background_color
        if color == "random":
            bg = self._bg_override if self._bg_override is not None else torch.rand_like(pred)
            pred = pred + bg * (1.0 - acc)
        else:
            bg = torch.full_like(pred, {"black": 0.0, "white": 1.0, "last_sample": 0.0}[color])
        if gt.shape[-1] == 4:
            gt = gt[..., :3] * gt[..., 3:] + bg * (1.0 - gt[..., 3:])
        return pred, gt

    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, Tensor]:
        cfg = self.config
        image = batch["image"].to(self.device)
        gt_spectral = batch["hs_image"].to(self.device)
        pred_rgb, gt_rgb = self._blend_background(outputs["rgb"], outputs["accumulation"], image)
        loss = {}
        if cfg.method == "spectral":
            loss["spectral_loss"] = F.mse_loss(outputs["spectral"], gt_spectral)
        else:
            loss["spectral_loss"] = 5 * F.mse_loss(outputs["spectral"], gt_spectral)  # hard-coded 5 (quirk 4)
            loss["rgb_loss"] = cfg.rgb_loss_weight * F.mse_loss(pred_rgb, gt_rgb)
        return loss

    # ------------------------------------------------------------------ ref umhs_model.py:385-405
    def get_metrics_dict(self, outputs, batch) -> Dict[str, Tensor]:
        m = {}
        gt_rgb = batch["image"].to(self.device)
        if gt_rgb.shape[-1] == 4:
            gt_rgb = gt_rgb[..., :3] * gt_rgb[..., 3:]
        mse_rgb = F.mse_loss(outputs["rgb"], gt_rgb)
        m["psnr"] = 10.0 * torch.log10(1.0 / mse_rgb)
        m["rmse"] = torch.sqrt(mse_rgb)
        gt_s = batch["hs_image"].to(self.device)
        mse_s = F.mse_loss(outputs["spectral"], gt_s)
        m["psnr_spectral"] = 10.0 * torch.log10(1.0 / mse_s)
        m["rmse_spectral"] = torch.sqrt(mse_s)
        m["num_samples_per_batch"] = outputs["num_samples_per_ray"].sum()
        return m

    # ------------------------------------------------------------------ ref umhs_model.py:407-527
    def _blend_gt_background(self, image: Tensor) -> Tensor:
        """nerfstudio RGBRenderer.blend_background: an RGBA ground truth over the configured background colour
        ("random" / "last_sample" count as black here, as in nerfstudio)."""
        if image.shape[-1] < 4:
            return image
        bg = 1.0 if self.config.background_color == "white" else 0.0
        return image[..., :3] * image[..., 3:] + bg * (1.0 - image[..., 3:])

    @torch.no_grad()
    def get_image_metrics_and_images(self, outputs: Dict[str, Tensor], batch: Dict[str, Tensor]
                                     ) -> Tuple[Dict[str, float], Dict[str, Tensor]]:
        """ref umhs_model.py:407-512: PSNR / SSIM / LPIPS of the rendered sRGB image, PSNR / SSIM / SAM / RMSE of the
        rendered spectrum, and the viewer's image dictionary.  The reference also writes ``seg_pred`` PNGs into a
        hard-coded ``./eval/ajar_final`` (ref :495-501): kept behind ``UMHS_EVAL_DUMP_DIR``."""
        from . import metrics as M

        gt_rgb = self._blend_gt_background(batch["image"].to(self.device))
        predicted_rgb = outputs["rgb"]
        acc = M.apply_colormap(outputs["accumulation"])
        depth = M.apply_depth_colormap(outputs["depth"], accumulation=outputs["accumulation"])
        combined_rgb = torch.cat([gt_rgb, predicted_rgb], dim=1)
        gt_c = torch.moveaxis(gt_rgb, -1, 0)[None, ...]         # [H, W, C] -> [1, C, H, W]
        pr_c = torch.moveaxis(predicted_rgb, -1, 0)[None, ...]
        se_per_pixel = ((gt_c - pr_c) ** 2).squeeze(0).permute(1, 2, 0).mean(dim=-1).unsqueeze(-1)
        metrics_dict = {"psnr": float(M.psnr(gt_c, pr_c)), "ssim": float(M.ssim(gt_c, pr_c)),
                        "lpips": M.lpips_or_nan(self.lpips, gt_c, pr_c)}
        if "spectral" in self.config.method:
            gt_s = torch.moveaxis(batch["hs_image"].to(self.device), -1, 0)[None, ...]
            pr_s = torch.moveaxis(outputs["spectral"], -1, 0)[None, ...]
            metrics_dict["psnr_spectral"] = float(M.psnr(gt_s, pr_s))
            metrics_dict["ssim_spectral"] = float(M.ssim(gt_s, pr_s))
            metrics_dict["sam_spectral"] = float(torch.nanmean(M.spectral_angle(pr_s, gt_s)))
            metrics_dict["rmse_spectral"] = float(torch.sqrt(F.mse_loss(pr_s, gt_s)))
            self.i = self.i + 1
            dump = os.environ.get("UMHS_EVAL_DUMP_DIR")
            if dump:  # the reference's side effect, opt-in: raw class ids + colour-coded segmentation
                os.makedirs(os.path.join(dump, "color"), exist_ok=True)
                np.save(os.path.join(dump, f"seg_pred_{self.i}.npy"),
                        outputs["seg_raw"].cpu().numpy().astype(np.uint8).squeeze())
                np.save(os.path.join(dump, "color", f"{self.i}.npy"),
                        (outputs["seg_pred"].cpu().numpy() * 255).astype(np.uint8))
        images_dict = {"img": combined_rgb, "accumulation": acc, "depth": depth, "se_per_pixel": se_per_pixel}
        return metrics_dict, images_dict

    def compute_sam(self, pred: Tensor, gt: Tensor, eps: float = 1e-8) -> Tensor:
        """ref umhs_model.py:514-527"""
        from .metrics import compute_sam

        return compute_sam(pred, gt, eps)

    # ------------------------------------------------------------------ ref umhs_model.py:529-540
    @torch.no_grad()
    def get_outputs_for_camera(self, camera, obb_box=None) -> Dict[str, Tensor]:
        """Camera -> ray bundle -> full-image outputs (the viewer's and ``ns-eval``'s entry point).  The reference
        wraps this in fp16 autocast; the kernels here choose their own precision (bf16 tensor-core operands)."""
        return self.get_outputs_for_camera_ray_bundle(
            camera.generate_rays(camera_indices=0, keep_shape=True, obb_box=obb_box))

    # ------------------------------------------------------------------ ref umhs_model.py:542-591
    def update_occupancy_grid(self, step: int, **kw) -> None:
        self.step = step
        self.occupancy_grid.update_every_n_steps(
            step=step, occ_eval_fn=lambda x, **k: self.field.density_fn(x, **k) * self.config.render_step_size, **kw)

    @torch.no_grad()
    def clamp_endmembers(self, step: int = 0) -> None:
        self.field.endmembers.clamp_(0.0, 1.0)

    def get_training_callbacks(self, training_callback_attributes=None) -> List:
        try:  # pragma: no cover - needs nerfstudio
            from nerfstudio.engine.callbacks import TrainingCallback, TrainingCallbackLocation
        except Exception:
            # plain (where, every, fn) triples when nerfstudio is absent
            return [("after_train_iteration", 1, self.clamp_endmembers),
                    ("before_train_iteration", 1, self.update_occupancy_grid)]
        return [TrainingCallback(where_to_run=[TrainingCallbackLocation.AFTER_TRAIN_ITERATION],
                                 update_every_num_iters=1, func=self.clamp_endmembers),
                TrainingCallback(where_to_run=[TrainingCallbackLocation.BEFORE_TRAIN_ITERATION],
                                 update_every_num_iters=1, func=self.update_occupancy_grid)]

    def get_param_groups(self) -> Dict[str, List[nn.Parameter]]:
        return {"fields": list(self.field.parameters())}

    # ------------------------------------------------------------------ ref umhs_model.py:593-620
    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle: RayBundle, per_band_views: bool = True
                                          ) -> Dict[str, Tensor]:
        """Full-image render.  The reference loops over 512-ray chunks; here a tile is as many rays as fit a launch
        (``eval_tile_rays``), each tile one fused pass, outputs written straight into [H, W, C] tensors."""
        input_device = camera_ray_bundle.directions.device
        H, W = camera_ray_bundle.origins.shape[:2]
        num_rays = H * W
        tile = int(self.config.eval_tile_rays)
        chunk = int(self.config.eval_num_rays_per_chunk)  # the reference's forward calls: depth is clipped per call
        chunks: Dict[str, List[Tensor]] = {}
        for s in range(0, num_rays, tile):
            e = min(s + tile, num_rays)
            rb = camera_ray_bundle.get_row_major_sliced_ray_bundle(s, e)
            rb = RayBundle(origins=rb.origins.to(self.device), directions=rb.directions.to(self.device),
                           nears=None, fars=None) if not hasattr(rb, "to") else rb.to(self.device)
            ids = None if s % chunk == 0 else torch.arange(s, e, device=self.device, dtype=torch.int64)
            out = self.get_outputs(self.collider(rb), per_band_views=False,
                                   depth_chunks=(chunk, ids, e if ids is not None else e - s))
            for k, v in out.items():
                if isinstance(v, Tensor):
                    chunks.setdefault(k, []).append(v.to(input_device))
        outputs = {k: torch.cat(v).view(H, W, -1) for k, v in chunks.items()}
        if per_band_views:
            for i in range(outputs["spectral"].shape[-1]):
                outputs[f"wv_{i}"] = outputs["spectral"][..., i:i + 1]
            for i in range(outputs["abundances"].shape[-1]):
                outputs[f"abundances_{i}"] = outputs["abundances"][..., i:i + 1]
            if "specular" in outputs:
                for i in range(outputs["specular"].shape[-1]):
                    outputs[f"residual_{i}"] = outputs["specular"][..., i:i + 1]
        return outputs
