"""Occupancy grid + packed ray-sample generation (B200-native twin of nerfacc 0.5.2 OccGridEstimator and
nerfstudio VolumetricSampler, reached from ref umhsnerf/umhs_model.py:198-212, 229-237, 549-554).

State-dict names and shapes follow nerfacc (``resolution``, ``aabbs``, ``occs``, ``binaries``) so reference
checkpoints load.  Marching, the visibility filter, compaction and the grid update run in
libumhs_sm100.so (csrc/packing.cu).
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib, ops
from ._lib import check, ptr, stream_ptr
from .ops import PointSet


class OccGridEstimator(nn.Module):
    def __init__(self, roi_aabb: Tensor, resolution: int = 128, levels: int = 4):
        super().__init__()
        roi = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        self.levels = int(levels)
        res = int(resolution)
        c, half = (roi[:3] + roi[3:]) / 2, (roi[3:] - roi[:3]) / 2
        aabbs = torch.stack([torch.cat([c - half * 2 ** l, c + half * 2 ** l]) for l in range(levels)], dim=0)
        self.register_buffer("resolution", torch.tensor([res] * 3, dtype=torch.int32))
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(levels * res ** 3))
        self.register_buffer("binaries", torch.zeros((levels, res, res, res), dtype=torch.bool))
        self._res = res
        self._occ_mean: Optional[Tensor] = None  # device scalar cache of occs.mean()
        # bit-packed copy of ``binaries`` + per-level "any occupied" flags for the march kernels: persistent buffers,
        # refreshed when the grid changes (every 16 steps) instead of once per march
        self._bits: Optional[Tensor] = None
        self._level_any: Optional[Tensor] = None
        self._bits_key = None

    def mark_dirty(self) -> None:
        """Call after writing ``occs`` / ``binaries`` by hand (tests, checkpoint surgery)."""
        self._occ_mean = None
        self._bits_key = None

    def _load_from_state_dict(self, *a, **k):
        self._occ_mean = None
        self._bits_key = None
        return super()._load_from_state_dict(*a, **k)

    def _repack(self) -> None:
        b = self.binaries
        if self._bits is None or self._bits.device != b.device:
            self._bits = self._level_any = None
        self._bits, self._level_any = ops.occ_pack_bits(b, self._bits, self._level_any)
        self._bits_key = (b.data_ptr(), b._version)

    def packed_grid(self):
        """(bits, level_any) of the current ``binaries``.  In-place torch writes are seen through the tensor version;
        kernels that write ``binaries`` through its address (the occupancy update) repack explicitly."""
        b = self.binaries
        if not b.is_cuda:
            return None
        if self._bits_key != (b.data_ptr(), b._version):
            self._repack()
        return self._bits, self._level_any

    def occ_mean(self) -> Tensor:
        if self._occ_mean is None or self._occ_mean.device != self.occs.device:
            self._occ_mean = self.occs.mean().reshape(1).contiguous()
        return self._occ_mean

    # ------------------------------------------------------------------ sampling (A.4)
    @torch.no_grad()
    def sampling(self, rays_o: Tensor, rays_d: Tensor, sigma_fn: Optional[Callable] = None, near_plane: float = 0.0,
                 far_plane: float = 1e10, t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None,
                 render_step_size: float = 1e-3, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                 stratified: bool = False, cone_angle: float = 0.0, jitter: Optional[Tensor] = None,
                 return_packed_info: bool = False, capacity: Optional[int] = None, premarched=None):
        """-> (ray_indices int64 [N], t_starts [N], t_ends [N]) (+ packed_info [R,2]).  ``sigma_fn(t_starts, t_ends,
        ray_indices) -> sigma [N]`` enables the visibility filter, as in nerfacc.  ``jitter`` injects the stratified
        draw (nerfacc: ``torch.rand_like(near_planes)``) so parity tests see identical near planes.
        ``premarched``: the tuple ``march_candidates`` returned for these rays (the march itself is then skipped)."""
        if premarched is None:
            premarched = self.march_candidates(rays_o, rays_d, near_plane, far_plane, t_min, t_max, render_step_size,
                                               stratified, cone_angle, jitter, capacity)
        ri, t0, t1, packed, total = premarched
        return self._filter(ri, t0, t1, packed, total, sigma_fn, early_stop_eps, alpha_thre, return_packed_info,
                            capacity)

    @torch.no_grad()
    def march_candidates(self, rays_o: Tensor, rays_d: Tensor, near_plane: float = 0.0, far_plane: float = 1e10,
                         t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None, render_step_size: float = 1e-3,
                         stratified: bool = False, cone_angle: float = 0.0, jitter: Optional[Tensor] = None,
                         capacity: Optional[int] = None):
        """The parameter-independent half of ``sampling``: near / far planes, stratified jitter, occupancy march.
        -> (ray_indices, t_starts, t_ends, packed_info, total).  It depends only on the rays and on ``binaries``, so
        a trainer may run it for the NEXT batch while the current step is still computing."""
        # the planes before jitter are the same every step when the collider hands over the same constant tensors
        # (UMHSModel.collider caches them): keep them instead of four fill / clamp launches per march
        c = getattr(self, "_plane_cache", None)
        if (c is not None and c[0] is t_min and c[1] is t_max and c[2] == (near_plane, far_plane, tuple(rays_o.shape))
                and c[3].device == rays_o.device):
            near, far = c[3], c[4]
        else:
            near = torch.full_like(rays_o[..., 0], fill_value=near_plane)
            far = torch.full_like(rays_o[..., 0], fill_value=far_plane)
            if t_min is not None:
                near = torch.clamp(near, min=t_min)
            if t_max is not None:
                far = torch.clamp(far, max=t_max)
            self._plane_cache = (t_min, t_max, (near_plane, far_plane, tuple(rays_o.shape)), near, far)
        if stratified:
            if jitter is None:
                jitter = torch.rand_like(near)
            near = near + jitter * render_step_size
        return ops.march(rays_o, rays_d, near, far, self.binaries, self.aabbs, render_step_size, cone_angle,
                         capacity=capacity, packed_grid=self.packed_grid())

    def _filter(self, ri, t0, t1, packed, total, sigma_fn, early_stop_eps, alpha_thre, return_packed_info, capacity):
        march_total = total
        self.last_src_index = None
        if (alpha_thre > 0.0 or early_stop_eps > 0.0) and sigma_fn is not None and t0.shape[0] > 0:
            # nerfacc's sigma_fn(t_starts, t_ends, ray_indices); the sync-free path also hands over the device count
            sigma = sigma_fn(t0, t1, ri) if capacity is None else sigma_fn(t0, t1, ri, total)
            march_total = total
            ri, t0, t1, packed, total, self.last_src_index = ops.visibility_compact(
                packed, t0, t1, sigma, early_stop_eps, alpha_thre, self.occ_mean() if alpha_thre > 0.0 else None,
                capacity=capacity)
        # device int64[2] each: (packed, unclamped) candidates / kept.  Training calls only: the trainer's capacity
        # check reads the LAST TRAINING step's counts (graph replays do not come through here), an evaluation render in
        # between must not replace them with a whole image's sample count
        if self.training or not hasattr(self, "last_totals"):
            self.last_totals = (march_total, total)
        self.last_eval_totals = (march_total, total)
        if return_packed_info:
            return ri, t0, t1, packed, total
        return ri, t0, t1

    # ------------------------------------------------------------------ update (A.4)
    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                             warmup_steps: int = 256, n: int = 16, cell_ids: Optional[List[Tensor]] = None,
                             cell_jitter: Optional[List[Tensor]] = None, freeze_binaries: bool = False,
                             cell_draws: Optional[Tuple[List[Tensor], List[Tensor]]] = None) -> None:
        """occs[id] = max(occs[id]*decay, occ_eval_fn(x)); binaries = occs > min(mean(occs), occ_thre).

        Cells: all of them while ``step < warmup_steps``; afterwards nerfacc's n/4 uniform + up to n/4 occupied cells
        per level, chosen by ``umhs_occ_select_cells`` from the bit-packed grid (no nonzero / sort / host sync: the
        update can be captured as a CUDA graph).  The random numbers come from a counter-based generator inside the
        kernels, seeded by (grid seed, level) and a device-side epoch that advances with every update -- identical on
        all data-parallel ranks.  Test hooks: ``cell_ids`` (per level) replaces the selection, ``cell_draws`` =
        (uniform draws, pick draws) per level feeds it, ``cell_jitter`` fixes the in-cell offsets.  ``freeze_binaries``
        (bench only) does all the work but writes occs / binaries of a scratch copy, so a synthetic scene stays fixed."""
        if not (self.training and step % n == 0):
            return
        _lib.load()
        dev = self.occs.device
        res, n_cells = self._res, self._res ** 3
        st = stream_ptr(dev)
        occs, binaries = self.occs, self.binaries
        if freeze_binaries:
            if getattr(self, "_scratch", None) is None:
                self._scratch = (self.occs.clone(), self.binaries.clone())
            occs, binaries = self._scratch
        steady = cell_ids is None and step >= warmup_steps
        if getattr(self, "_epoch", None) is None or self._epoch.device != dev:
            self._epoch = torch.zeros((1,), dtype=torch.int64, device=dev)   # device-side epoch of the draws
            self._prefix = torch.empty(((n_cells + 31) // 32 + 2 + (n_cells + 32767) // 32768,), dtype=torch.int32, device=dev)
        self._epoch.add_(1)  # enqueued (and captured) with the update: every replay draws fresh numbers
        if steady:
            bits = self.packed_grid()[0] if not freeze_binaries else ops.occ_pack_bits(binaries)[0]
        for lvl in range(self.levels):
            n_valid = None
            seed = (int(getattr(self, "seed", 0)) * 1000003 + 7919 * lvl + 1) & 0xFFFFFFFFFFFFFFFF
            if cell_ids is not None:
                idx = cell_ids[lvl].to(torch.int64).contiguous()
            elif not steady:
                idx = None  # all cells, in order
            else:
                n_u = n_cells // 4
                idx = torch.empty((2 * n_u,), dtype=torch.int64, device=dev)
                n_valid = torch.empty((1,), dtype=torch.int64, device=dev)
                uu = cell_draws[0][lvl].to(dev, torch.float32).contiguous() if cell_draws is not None else None
                up = cell_draws[1][lvl].to(dev, torch.float32).contiguous() if cell_draws is not None else None
                ops._call("umhs_occ_select_cells", 2 * n_u, ptr(bits[lvl]), n_cells, n_u, ptr(uu), ptr(up), seed,
                          ptr(self._epoch), ptr(self._prefix), ptr(idx), ptr(n_valid), st)
            m = n_cells if idx is None else int(idx.shape[0])
            if m == 0:
                continue
            jit_all = cell_jitter[lvl].to(dev, torch.float32).contiguous() if cell_jitter is not None else None
            aabb = self.aabbs[lvl].contiguous()
            sigma = torch.empty((m,), dtype=torch.float32, device=dev)
            # the density pass runs in fixed-size chunks: its temporaries (positions, hash encodings) keep one size
            # from call to call, so the caching allocator serves them from the same blocks.  One chunk holds a whole
            # level of the steady-state update (2 x 128^3 / 4 cells): four launches of 262 k points took 4 x 42 us in
            # the hash gather alone where one of 1 M takes 100 us
            CH = 1 << 20
            for c0 in range(0, m, CH):
                mc = min(CH, m - c0)
                idx_c = None if idx is None else idx[c0:c0 + mc]
                if idx is None:  # all cells in order: explicit ids for this chunk
                    idx_c = torch.arange(c0, c0 + mc, device=dev, dtype=torch.int64)
                jit = jit_all[c0:c0 + mc].contiguous() if jit_all is not None else None
                pos = torch.empty((mc, 3), dtype=torch.float32, device=dev)
                ops._call("umhs_occ_cell_points", mc, ptr(idx_c), ptr(jit), (seed + 0x51ED27 * (c0 + 1)) & (2 ** 64 - 1),
                          ptr(self._epoch), mc, res, ptr(aabb), ptr(pos), st)
                if n_valid is not None:  # the unused slots (-1) sit behind n_valid: they are not evaluated
                    cnt = (n_valid - c0).clamp(min=0, max=mc).to(torch.int64).reshape(1)
                    sigma[c0:c0 + mc] = occ_eval_fn(pos, n_dev=cnt).reshape(-1)
                else:
                    sigma[c0:c0 + mc] = occ_eval_fn(pos).reshape(-1)
            # ONE read phase over all of the level's cells, then one write phase (torch index-assign semantics: every
            # right-hand side sees the OLD occs, however often a cell was drawn)
            tmp = torch.empty((m,), dtype=torch.float32, device=dev)
            ops._call("umhs_occ_ema_update", m, ptr(occs), ptr(idx), lvl * n_cells, ptr(sigma), 1.0,
                      float(ema_decay), m, ptr(tmp), st)
        scratch = torch.empty((2,), dtype=torch.float64, device=dev)
        mean = torch.empty((1,), dtype=torch.float32, device=dev)
        bn = binaries.view(torch.uint8)
        ops._call("umhs_occ_binarize", occs.numel(), ptr(occs), occs.numel(), float(occ_thre), ptr(scratch), ptr(bn),
                  ptr(mean), st)
        if not freeze_binaries:
            self._occ_mean = mean
        if self._bits is not None:  # binaries were (re)written through their address: refresh the packed copy
            self._repack()


class VolumetricSampler(nn.Module):
    """nerfstudio VolumetricSampler over the grid above.  ``__call__`` returns packed samples of a ray batch:
    (PointSet, ray_indices, t_starts, t_ends, packed_info)."""

    def __init__(self, occupancy_grid: OccGridEstimator, field):
        super().__init__()
        self.occupancy_grid = occupancy_grid
        self._field = [field]  # not a submodule: the model owns the field

    @torch.no_grad()
    def march_only(self, ray_bundle, render_step_size: float, near_plane: float = 0.0,
                   far_plane: Optional[float] = None, cone_angle: float = 0.0, jitter: Optional[Tensor] = None,
                   capacity: Optional[int] = None):
        """Candidate samples of a ray bundle (``OccGridEstimator.march_candidates``) with this sampler's conventions."""
        rays_o = ray_bundle.origins.reshape(-1, 3).to(torch.float32).contiguous()
        rays_d = ray_bundle.directions.reshape(-1, 3).to(torch.float32).contiguous()
        t_min = t_max = None
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            t_min = ray_bundle.nears.contiguous().reshape(-1)
            t_max = ray_bundle.fars.contiguous().reshape(-1)
        return self.occupancy_grid.march_candidates(
            rays_o, rays_d, near_plane, 1e10 if far_plane is None else far_plane, t_min, t_max, render_step_size,
            self.training, cone_angle, jitter, capacity)

    @torch.no_grad()
    def forward(self, ray_bundle, render_step_size: float, near_plane: float = 0.0, far_plane: Optional[float] = None,
                alpha_thre: float = 0.01, cone_angle: float = 0.0, jitter: Optional[Tensor] = None,
                capacity: Optional[int] = None, premarched=None):
        """capacity=None: exact-sized packed samples (one host sync per march / compaction, like the reference).
        capacity=C: sync-free; buffers hold C samples, ``pts.n_dev`` carries the valid count on the device.
        premarched: result of ``march_only`` on the same ray bundle."""
        field = self._field[0]
        rays_o = ray_bundle.origins.reshape(-1, 3).to(torch.float32).contiguous()
        rays_d = ray_bundle.directions.reshape(-1, 3).to(torch.float32).contiguous()
        t_min = t_max = None
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            t_min = ray_bundle.nears.contiguous().reshape(-1)
            t_max = ray_bundle.fars.contiguous().reshape(-1)
        if far_plane is None:
            far_plane = 1e10
        sigma_fn = None
        cand_enc = [None]  # hash encodings of the candidate samples (kept samples reuse them)
        if self.training:  # nerfstudio builds sigma_fn only while training
            def sigma_fn(t0, t1, ri, n_dev=None):
                pts = PointSet.from_packed(rays_o, rays_d, ri, t0, t1, contraction=field.contraction,
                                           aabb=field._aabb6(), n_dev=n_dev)
                sigma, cand_enc[0] = field.density_packed(pts, return_enc=True, zero_tail=False)
                return sigma
        ri, t0, t1, packed, total = self.occupancy_grid.sampling(
            rays_o, rays_d, sigma_fn=sigma_fn, near_plane=near_plane, far_plane=far_plane, t_min=t_min, t_max=t_max,
            render_step_size=render_step_size, stratified=self.training, cone_angle=cone_angle, alpha_thre=alpha_thre,
            jitter=jitter, return_packed_info=True, capacity=capacity, premarched=premarched)
        if capacity is None and t0.shape[0] == 0:  # nerfstudio fabricates one sample so shapes are never empty
            dev = rays_o.device
            ri = torch.zeros((1,), dtype=torch.int64, device=dev)
            t0 = torch.ones((1,), dtype=torch.float32, device=dev)
            t1 = torch.ones((1,), dtype=torch.float32, device=dev)
            packed = ops.pack_info(ri, rays_o.shape[0])
        pts = PointSet.from_packed(rays_o, rays_d, ri, t0, t1, contraction=field.contraction, aabb=field._aabb6(),
                                   n_dev=total if capacity is not None else None)
        src = self.occupancy_grid.last_src_index
        if cand_enc[0] is not None and src is not None and src.shape[0] == t0.shape[0]:
            pts.enc_src, pts.enc_index = cand_enc[0], src
        return pts, ri, t0, t1, packed
