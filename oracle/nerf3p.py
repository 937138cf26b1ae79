"""CPU restatement of the THIRD-PARTY semantics the UMHS hot path relies on.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker.

The reference (``/root/reference``, Factral/unsupervised-hyperspectral-nerf) is a
nerfstudio plugin; the arithmetic of its hot path lives partly in third-party
packages that are NOT vendored under ``/root/reference`` and are NOT installed
here (no network):

* ``nerfstudio == 1.1.5``  (pinned: ref ``nerfstudioa100_environment.yml:189``)
* ``nerfacc    == 0.5.2``  (pinned: ref ``nerfstudioa100_environment.yml:188``)

This module restates their *published* algorithms from recollection of the
public sources.  PARITY UNPINNED for this file: no
golden vector of the third-party code exists in the reference (it has no tests),
so every function here is an explicit, documented choice anchored on the
reference's own call sites, which are cited per function.

Everything is plain fp32 torch on CPU, autograd-capable.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

# ----------------------------------------------------------------------------
# Ray containers (nerfstudio.cameras.rays) -- only the fields the hot path reads
# call sites: ref umhsnerf/umhs_field.py:159-183,300-304; umhs_model.py:225-258
# ----------------------------------------------------------------------------


@dataclass
class Frustums:
    origins: Tensor  # [..., 3]
    directions: Tensor  # [..., 3]
    starts: Tensor  # [..., 1]
    ends: Tensor  # [..., 1]
    pixel_area: Optional[Tensor] = None

    def get_positions(self) -> Tensor:
        # nerfstudio Frustums.get_positions: origins + directions * (starts + ends) / 2
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        return pos

    @property
    def shape(self):
        return self.origins.shape[:-1]


@dataclass
class RaySamples:
    frustums: Frustums
    camera_indices: Optional[Tensor] = None


@dataclass
class RayBundle:
    origins: Tensor  # [R, 3]
    directions: Tensor  # [R, 3]
    pixel_area: Optional[Tensor] = None
    camera_indices: Optional[Tensor] = None
    nears: Optional[Tensor] = None
    fars: Optional[Tensor] = None

    def __len__(self) -> int:
        return self.origins.shape[0]


# ----------------------------------------------------------------------------
# A.1  HashEncoding (torch path) -- call site ref umhs_field.py:320 (mlp_base)
# ----------------------------------------------------------------------------


def hash_scalings(num_levels: int = 16, min_res: int = 16, max_res: int = 2048) -> Tensor:
    """``floor(min_res * growth**levels)`` evaluated the way nerfstudio does:
    ``growth`` is a numpy float64 scalar, ``levels`` an int64 tensor, so the power is
    ``Tensor.__rpow__`` -> fp32 tensor.  (fp32 pow makes the last level 2047 or
    2048 depending on rounding; it is a buffer in the state dict, so the kernels
    take it as data, never recompute it.)"""
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
    return torch.floor(min_res * float(growth) ** levels).to(torch.float32)


_HASH_PRIMES = (1, 2654435761, 805459861)


def hash_fn(coords_i32: Tensor, table_size: int, hash_offset: Tensor) -> Tensor:
    """int32 coords * int64 primes -> xor -> mod T -> + level*T   (all int64)."""
    c = coords_i32 * torch.tensor(_HASH_PRIMES, dtype=torch.int64)
    x = torch.bitwise_xor(c[..., 0], c[..., 1])
    x = torch.bitwise_xor(x, c[..., 2])
    x = x % table_size
    return x + hash_offset


def hash_encode(x01: Tensor, table: Tensor, scalings: Tensor, log2_T: int) -> Tensor:
    """x01 [N,3] in [0,1] -> [N, L*F].  Eight ceil/floor corners, trilinear blend in
    the exact association order of HashEncoding.pytorch_fwd (A.1)."""
    L = scalings.shape[0]
    T = 1 << log2_T
    hash_offset = torch.arange(L, dtype=torch.int64) * T
    scaled = x01[..., None, :] * scalings.view(-1, 1)  # [N, L, 3]
    sc = torch.ceil(scaled).type(torch.int32)
    sf = torch.floor(scaled).type(torch.int32)
    off = scaled - sf

    def h(a, b, c):
        return hash_fn(torch.cat([a, b, c], dim=-1), T, hash_offset)

    cx, cy, cz = sc[..., 0:1], sc[..., 1:2], sc[..., 2:3]
    fx, fy, fz = sf[..., 0:1], sf[..., 1:2], sf[..., 2:3]
    f0 = table[h(cx, cy, cz)]
    f1 = table[h(cx, fy, cz)]
    f2 = table[h(fx, fy, cz)]
    f3 = table[h(fx, cy, cz)]
    f4 = table[h(cx, cy, fz)]
    f5 = table[h(cx, fy, fz)]
    f6 = table[h(fx, fy, fz)]
    f7 = table[h(fx, cy, fz)]
    ox, oy, oz = off[..., 0:1], off[..., 1:2], off[..., 2:3]
    f03 = f0 * ox + f3 * (1 - ox)
    f12 = f1 * ox + f2 * (1 - ox)
    f56 = f5 * ox + f6 * (1 - ox)
    f47 = f4 * ox + f7 * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    enc = f0312 * oz + f4756 * (1 - oz)  # [N, L, F]
    return torch.flatten(enc, start_dim=-2, end_dim=-1)


# ----------------------------------------------------------------------------
# A.2  MLP (torch path) -- call sites ref umhs_field.py:67,95,105,211,219,250
# ----------------------------------------------------------------------------


class MLP(nn.Module):
    """``num_layers`` nn.Linear layers with bias, activation after all but the last,
    then ``out_activation``.  State-dict names ``layers.{i}.{weight,bias}`` (A.9)."""

    def __init__(self, in_dim, num_layers, layer_width, out_dim, activation=nn.ReLU(), out_activation=None):
        super().__init__()
        dims = [in_dim] + [layer_width] * (num_layers - 1) + [out_dim]
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers)])
        self.activation = activation
        self.out_activation = out_activation

    def forward(self, x: Tensor) -> Tensor:
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if self.activation is not None and i < len(self.layers) - 1:
                x = self.activation(x)
        if self.out_activation is not None:
            x = self.out_activation(x)
        return x


class HashEncoding(nn.Module):
    def __init__(self, num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=19, features_per_level=2,
                 hash_init_scale=1e-3):
        super().__init__()
        self.num_levels, self.log2_T, self.F = num_levels, log2_hashmap_size, features_per_level
        T = 1 << log2_hashmap_size
        tab = torch.rand(size=(T * num_levels, features_per_level)) * 2 - 1
        tab *= hash_init_scale
        self.hash_table = nn.Parameter(tab)
        self.register_buffer("scalings", hash_scalings(num_levels, min_res, max_res))

    def forward(self, x01: Tensor) -> Tensor:
        return hash_encode(x01, self.hash_table, self.scalings, self.log2_T)


class MLPWithHashEncoding(nn.Module):
    """nerfstudio MLPWithHashEncoding: ``mlp(encoder(x))``; names ``encoder.*``, ``mlp.*``."""

    def __init__(self, num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=19, features_per_level=2,
                 num_layers=2, layer_width=64, out_dim=16):
        super().__init__()
        self.encoder = HashEncoding(num_levels, min_res, max_res, log2_hashmap_size, features_per_level)
        self.mlp = MLP(num_levels * features_per_level, num_layers, layer_width, out_dim, nn.ReLU(), None)

    def forward(self, x01: Tensor) -> Tensor:
        return self.mlp(self.encoder(x01))


# ----------------------------------------------------------------------------
# A.3  Encodings -- call sites ref umhs_field.py:162 (SH), :184 (NeRF position enc.)
# ----------------------------------------------------------------------------


def nerf_encoding(x: Tensor, num_frequencies: int = 2, min_freq: float = 0.0, max_freq: float = 1.0) -> Tensor:
    """NeRFEncoding.pytorch_fwd, include_input=False: sin(cat([s, s + pi/2])) with
    s = (2*pi*x)[..., None] * 2**linspace(min,max,n) flattened dim-major/freq-minor."""
    s = 2 * torch.pi * x
    freqs = 2 ** torch.linspace(min_freq, max_freq, num_frequencies)
    si = s[..., None] * freqs
    si = si.view(*si.shape[:-2], -1)
    return torch.sin(torch.cat([si, si + torch.pi / 2.0], dim=-1))


def sh_encoding_deg3(d: Tensor) -> Tensor:
    """SHEncoding(levels=4).pytorch_fwd: 16 real-SH components of the *given* vector
    (the field passes (d+1)/2, ref umhs_field.py:160, so inputs are in [0,1]^3)."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz = x * x, y * y, z * z
    c = torch.zeros((*d.shape[:-1], 16), dtype=d.dtype)
    c[..., 0] = 0.28209479177387814
    c[..., 1] = 0.4886025119029199 * y
    c[..., 2] = 0.4886025119029199 * z
    c[..., 3] = 0.4886025119029199 * x
    c[..., 4] = 1.0925484305920792 * x * y
    c[..., 5] = 1.0925484305920792 * y * z
    c[..., 6] = 0.9461746957575601 * zz - 0.31539156525251999
    c[..., 7] = 1.0925484305920792 * x * z
    c[..., 8] = 0.5462742152960396 * (xx - yy)
    c[..., 9] = 0.5900435899266435 * y * (3 * xx - yy)
    c[..., 10] = 2.890611442640554 * x * y * z
    c[..., 11] = 0.4570457994644658 * y * (5 * zz - 1)
    c[..., 12] = 0.3731763325901154 * z * (5 * zz - 3)
    c[..., 13] = 0.4570457994644658 * x * (5 * zz - 1)
    c[..., 14] = 1.445305721320277 * z * (xx - yy)
    c[..., 15] = 0.5900435899266435 * x * (xx - 3 * yy)
    return c


def get_normalized_directions(directions: Tensor) -> Tensor:
    return (directions + 1.0) / 2.0


# ----------------------------------------------------------------------------
# A.8  SceneContraction(order=inf), trunc_exp -- ref umhs_field.py:304-305,327
# ----------------------------------------------------------------------------


def scene_contraction_inf(x: Tensor) -> Tensor:
    mag = torch.linalg.norm(x, ord=float("inf"), dim=-1)[..., None]
    return torch.where(mag < 1, x, (2 - (1 / mag)) * (x / mag))


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


def scene_box_normalize(positions: Tensor, aabb: Tensor) -> Tensor:
    """SceneBox.get_normalized_positions (ref umhs_field.py:307)."""
    return (positions - aabb[0]) / (aabb[1] - aabb[0])


# ----------------------------------------------------------------------------
# A.6  scale_gradients_by_distance_squared -- call site ref umhs_model.py:241-242
# ----------------------------------------------------------------------------


class _GradientScaler(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, scaling):
        ctx.save_for_backward(scaling)
        return value, scaling

    @staticmethod
    def backward(ctx, output_grad, grad_scaling):
        (scaling,) = ctx.saved_tensors
        return output_grad * scaling, grad_scaling


def scale_gradients_by_distance_squared(field_outputs: Dict, ray_samples: RaySamples) -> Dict:
    out = {}
    ray_dist = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
    scaling = torch.square(ray_dist).clamp(0, 1)
    for key, value in field_outputs.items():
        out[key], _ = _GradientScaler.apply(value, scaling)
    return out


# ----------------------------------------------------------------------------
# A.5  nerfacc rendering -- call sites ref umhs_model.py:245-252, umhs_renderer.py:28
# ----------------------------------------------------------------------------


def pack_info(ray_indices: Tensor, n_rays: int) -> Tensor:
    """[n_rays, 2] (start, count), dtype of ray_indices (int64)."""
    cnt = torch.zeros((n_rays,), dtype=ray_indices.dtype)
    cnt.index_add_(0, ray_indices, torch.ones_like(ray_indices))
    start = cnt.cumsum(0) - cnt
    return torch.stack([start, cnt], dim=-1)


def exclusive_sum_packed(x: Tensor, packed_info: Tensor) -> Tensor:
    """Per-ray exclusive prefix sum of a packed [N] tensor, summed sequentially in
    fp32 inside each segment (dense scatter -> cumsum -> gather keeps autograd)."""
    n_rays = packed_info.shape[0]
    cnt = packed_info[:, 1]
    smax = int(cnt.max().item()) if n_rays > 0 and x.numel() > 0 else 0
    if smax == 0:
        return torch.zeros_like(x)
    ray_of = torch.repeat_interleave(torch.arange(n_rays), cnt)
    pos_in = torch.arange(x.shape[0]) - packed_info[:, 0][ray_of]
    dense = torch.zeros((n_rays, smax), dtype=x.dtype)
    dense = dense.index_put((ray_of, pos_in), x)
    incl = torch.cumsum(dense, dim=1)
    excl = incl - dense
    return excl[ray_of, pos_in]


def render_weight_from_density(t_starts: Tensor, t_ends: Tensor, sigmas: Tensor, packed_info: Tensor):
    sd = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sd)
    trans = torch.exp(-exclusive_sum_packed(sd, packed_info))
    return trans * alphas, trans, alphas


def accumulate_along_rays(weights: Tensor, values: Optional[Tensor], ray_indices: Optional[Tensor], n_rays: Optional[int]):
    if values is None:
        src = weights[..., None]
    else:
        src = weights[..., None] * values
    if ray_indices is not None:
        out = torch.zeros((n_rays, src.shape[-1]), dtype=src.dtype)
        out = out.index_add(0, ray_indices, src)
        return out
    return torch.sum(src, dim=-2)


def render_visibility_from_density(t_starts, t_ends, sigmas, packed_info, early_stop_eps=1e-4, alpha_thre=0.0):
    _, trans, alphas = render_weight_from_density(t_starts, t_ends, sigmas, packed_info)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


# ----------------------------------------------------------------------------
# A.7  nerfstudio renderers -- call sites ref umhs_model.py:254-258, 358-362
# ----------------------------------------------------------------------------


def render_accumulation(weights: Tensor, ray_indices: Tensor, num_rays: int) -> Tensor:
    return accumulate_along_rays(weights[..., 0], None, ray_indices, num_rays)


def render_depth_expected(weights: Tensor, ray_samples: RaySamples, ray_indices: Tensor, num_rays: int) -> Tensor:
    eps = 1e-10
    steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
    depth = accumulate_along_rays(weights[..., 0], steps, ray_indices, num_rays)
    acc = accumulate_along_rays(weights[..., 0], None, ray_indices, num_rays)
    depth = depth / (acc + eps)
    return torch.clip(depth, steps.min(), steps.max())


def blend_background_for_loss(pred_image: Tensor, pred_accumulation: Tensor, gt_image: Tensor,
                              background_color: str = "random", bg: Optional[Tensor] = None):
    """RGBRenderer.blend_background_for_loss_computation.  ``bg`` is the random
    background drawn by ``torch.rand_like(pred_image)``; passed in so that both sides
    of a parity test see the same draw."""
    if background_color == "random":
        if bg is None:
            bg = torch.rand_like(pred_image)
        pred_image = pred_image + bg * (1.0 - pred_accumulation)
        bg_col = bg
    else:
        v = {"black": 0.0, "last_sample": 0.0, "white": 1.0}[background_color]
        bg_col = torch.full_like(pred_image, v)
    if gt_image.size(-1) == 4:
        rgb, opacity = gt_image[..., :3], gt_image[..., 3:]
        gt_image = rgb * opacity + bg_col * (1 - opacity)
    return pred_image, gt_image


# ----------------------------------------------------------------------------
# A.4  Packing: nerfacc OccGridEstimator + traverse_grids (C oracle does the march)
# ----------------------------------------------------------------------------


def occ_grid_aabbs(roi_aabb: Tensor, levels: int) -> Tensor:
    """aabbs[l] = roi enlarged x 2^l about its centre."""
    lo, hi = roi_aabb[:3], roi_aabb[3:]
    c = (lo + hi) / 2
    half = (hi - lo) / 2
    return torch.stack([torch.cat([c - half * (2 ** l), c + half * (2 ** l)]) for l in range(levels)], dim=0)


class OccGridEstimator(nn.Module):
    """State of nerfacc.OccGridEstimator (names as in A.9) + the sampling logic.
    The march itself is ``oracle.pack_oracle`` (C, -ffp-contract=off)."""

    def __init__(self, roi_aabb: Tensor, resolution: int = 128, levels: int = 4):
        super().__init__()
        self.levels = levels
        self.register_buffer("resolution", torch.tensor([resolution] * 3, dtype=torch.int32))
        self.register_buffer("aabbs", occ_grid_aabbs(roi_aabb.float(), levels))
        self.register_buffer("occs", torch.zeros(levels * resolution ** 3))
        self.register_buffer("binaries", torch.zeros((levels, resolution, resolution, resolution), dtype=torch.bool))

    @torch.no_grad()
    def sampling(self, rays_o, rays_d, sigma_fn=None, near_plane=0.0, far_plane=1e10, t_min=None, t_max=None,
                 render_step_size=1e-3, early_stop_eps=1e-4, alpha_thre=0.0, stratified=False, cone_angle=0.0,
                 jitter: Optional[Tensor] = None):
        from . import pack_oracle

        near_planes = torch.full_like(rays_o[..., 0], fill_value=near_plane)
        far_planes = torch.full_like(rays_o[..., 0], fill_value=far_plane)
        if t_min is not None:
            near_planes = torch.clamp(near_planes, min=t_min)
        if t_max is not None:
            far_planes = torch.clamp(far_planes, max=t_max)
        if stratified:
            if jitter is None:
                jitter = torch.rand_like(near_planes)
            near_planes = near_planes + jitter * render_step_size
        ray_indices, t_starts, t_ends, packed_info = pack_oracle.traverse_grids(
            rays_o, rays_d, self.binaries, self.aabbs, near_planes, far_planes, render_step_size, cone_angle)
        if (alpha_thre > 0.0 or early_stop_eps > 0.0) and sigma_fn is not None:
            alpha_thre = min(alpha_thre, self.occs.mean().item())
            if t_starts.shape[0] != 0:
                sigmas = sigma_fn(t_starts, t_ends, ray_indices)
            else:
                sigmas = torch.empty((0,))
            masks = render_visibility_from_density(t_starts, t_ends, sigmas, packed_info, early_stop_eps, alpha_thre)
            ray_indices, t_starts, t_ends = ray_indices[masks], t_starts[masks], t_ends[masks]
        return ray_indices, t_starts, t_ends

    @torch.no_grad()
    def update_every_n_steps(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16,
                             cell_ids: Optional[List[Tensor]] = None, cell_jitter: Optional[List[Tensor]] = None,
                             cell_draws=None):
        """A.4 ``update_every_n_steps``.  The random cell choice / in-cell jitter can be
        injected (``cell_ids`` / ``cell_jitter`` per level) for parity tests; ``cell_draws`` =
        (uniform draws, pick draws) per level, float32 in [0, 1), stands for the two ``torch.randint``
        calls of the steady-state selection (randint(n) == floor(u * n) in distribution)."""
        if not (self.training and step % n == 0):
            return
        res = int(self.resolution[0])
        n_cells = res ** 3
        for lvl in range(self.levels):
            if cell_ids is not None:
                idx = cell_ids[lvl]
            elif step < warmup_steps:
                idx = torch.arange(n_cells)
            else:
                n_u = n_cells // 4
                occ_idx = torch.nonzero(self.binaries[lvl].flatten())[:, 0]
                if cell_draws is None:
                    uni = torch.randint(n_cells, (n_u,))
                    if n_u < len(occ_idx):
                        occ_idx = occ_idx[torch.randint(len(occ_idx), (n_u,))]
                else:
                    uu, up = cell_draws[0][lvl].float(), cell_draws[1][lvl].float()
                    uni = (uu * float(n_cells)).long().clamp(max=n_cells - 1)
                    if n_u < len(occ_idx):
                        occ_idx = occ_idx[(up * float(len(occ_idx))).long().clamp(max=len(occ_idx) - 1)]
                idx = torch.cat([uni, occ_idx], dim=0)
            coords = torch.stack([idx // (res * res), (idx // res) % res, idx % res], dim=-1).float()
            jit = cell_jitter[lvl] if cell_jitter is not None else torch.rand_like(coords)
            x = (coords + jit) / res
            aabb = self.aabbs[lvl]
            x = aabb[:3] + x * (aabb[3:] - aabb[:3])
            occ = occ_eval_fn(x).squeeze(-1)
            gidx = idx + lvl * n_cells
            self.occs[gidx] = torch.maximum(self.occs[gidx] * ema_decay, occ)
        thre = torch.clamp(self.occs[self.occs >= 0].mean(), max=occ_thre)
        self.binaries = (self.occs > thre).view(self.binaries.shape)


def near_far_collider(ray_bundle: RayBundle, near_plane: float, far_plane: float, training: bool) -> RayBundle:
    """NearFarCollider: nears = near_plane while training, 0 in eval (reset_near_plane)."""
    ones = torch.ones_like(ray_bundle.origins[..., 0:1])
    ray_bundle.nears = ones * (near_plane if training else 0.0)
    ray_bundle.fars = ones * far_plane
    return ray_bundle


class VolumetricSampler:
    """nerfstudio VolumetricSampler.forward over an OccGridEstimator (A.4)."""

    def __init__(self, occupancy_grid: OccGridEstimator, density_fn):
        self.occupancy_grid = occupancy_grid
        self.density_fn = density_fn
        self.training = True

    def get_sigma_fn(self, origins, directions):
        if self.density_fn is None or not self.training:
            return None
        density_fn = self.density_fn

        def sigma_fn(t_starts, t_ends, ray_indices):
            t_origins = origins[ray_indices]
            t_dirs = directions[ray_indices]
            positions = t_origins + t_dirs * (t_starts + t_ends)[:, None] / 2.0
            return density_fn(positions).squeeze(-1)

        return sigma_fn

    def __call__(self, ray_bundle: RayBundle, render_step_size, near_plane=0.0, far_plane=None, alpha_thre=0.01,
                 cone_angle=0.0, jitter=None):
        rays_o = ray_bundle.origins.contiguous()
        rays_d = ray_bundle.directions.contiguous()
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            t_min = ray_bundle.nears.contiguous().reshape(-1)
            t_max = ray_bundle.fars.contiguous().reshape(-1)
        else:
            t_min = t_max = None
        if far_plane is None:
            far_plane = 1e10
        ray_indices, starts, ends = self.occupancy_grid.sampling(
            rays_o=rays_o, rays_d=rays_d, t_min=t_min, t_max=t_max, sigma_fn=self.get_sigma_fn(rays_o, rays_d),
            render_step_size=render_step_size, near_plane=near_plane, far_plane=far_plane,
            stratified=self.training, cone_angle=cone_angle, alpha_thre=alpha_thre, jitter=jitter)
        if starts.shape[0] == 0:
            ray_indices = torch.zeros((1,), dtype=torch.long)
            starts = torch.ones((1,), dtype=rays_o.dtype)
            ends = torch.ones((1,), dtype=rays_o.dtype)
        origins = rays_o[ray_indices]
        dirs = rays_d[ray_indices]
        cam = ray_bundle.camera_indices[ray_indices] if ray_bundle.camera_indices is not None else None
        rs = RaySamples(frustums=Frustums(origins=origins, directions=dirs, starts=starts[..., None],
                                          ends=ends[..., None], pixel_area=torch.zeros_like(origins[:, :1])),
                        camera_indices=cam)
        return rs, ray_indices


def psnr(pred: Tensor, target: Tensor) -> Tensor:
    """torchmetrics PeakSignalNoiseRatio(data_range=1.0)."""
    mse = torch.mean((pred - target) ** 2)
    return 10.0 * torch.log10(1.0 / mse)
