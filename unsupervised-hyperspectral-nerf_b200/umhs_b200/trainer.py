"""Data-parallel training step for UMHSModel: one process per GPU, rays sharded across ranks.

What nerfstudio's Trainer does around ``get_train_loss_dict`` for this method (ref
umhsnerf/umhs_config.py:34-69: Adam lr 2e-2 eps 1e-15, exponential decay to 1e-5 over 30 k steps; ref
umhsnerf/umhs_model.py:542-591: occupancy update before, endmember clamp after each iteration), restated for the
fused path:

  * all field parameters live in ONE flat fp32 buffer (the nn.Parameters are views into it), with flat grad /
    exp_avg / exp_avg_sq twins, so the optimiser is a single fused kernel (csrc/optim.cu: Adam + endmember clamp +
    grad zeroing) and the gradient exchange is a single collective;
  * every rank renders its own shard of the ray batch; one NCCL all-reduce (sum) of the flat gradient per step over
    NVLink / NVSwitch, averaged (DDP semantics: losses are means over the local shard), then identical Adam on every
    rank.  Evaluation shards image tiles with no collective.

The reference itself never syncs gradients (its pipeline forces world_size = 1: ref umhsnerf/umhs_pipeline.py:86,
108-113); this is the data-parallel path the north star asks for.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor, nn

from . import ops
from .rays import RayBundle
from .umhs_model import UMHSModel


def tile_pixel_indices(H: int, W: int, tile: int, rank: int, world_size: int, device=None) -> Tensor:
    """Flat pixel ids (row-major) of the tile x tile image tiles dealt round-robin to ``rank``.  The union over ranks
    covers every pixel exactly once: evaluation needs no collective."""
    ty, tx = (H + tile - 1) // tile, (W + tile - 1) // tile
    ids = []
    for t in range(rank, ty * tx, world_size):
        y0, x0 = (t // tx) * tile, (t % tx) * tile
        ys = torch.arange(y0, min(y0 + tile, H), device=device)
        xs = torch.arange(x0, min(x0 + tile, W), device=device)
        ids.append((ys[:, None] * W + xs[None, :]).reshape(-1))
    return torch.cat(ids) if ids else torch.zeros((0,), dtype=torch.int64, device=device)


def average_gradients(flat_grad: Tensor, world_size: int) -> None:
    """The one exchange step of data-parallel training: all-reduce(sum) of the flat gradient.  The 1/world_size
    factor (DDP semantics: each rank's loss is a mean over its own ray shard) is applied inside the fused Adam."""
    if world_size > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)


class FlatParams:
    """Re-home a list of parameters into one contiguous fp32 buffer (16-byte aligned segments)."""

    def __init__(self, params: List[nn.Parameter]):
        dev = params[0].device
        self.params = params
        self.offsets: List[int] = []
        n = 0
        for p in params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.data = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(params, self.offsets):
            view = self.data[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def range_of(self, p: nn.Parameter):
        for q, o in zip(self.params, self.offsets):
            if q is p:
                return o, o + q.numel()
        raise KeyError("parameter not in the flat buffer")


class UMHSTrainer:
    def __init__(self, model: UMHSModel, lr: float = 2e-2, eps: float = 1e-15, betas=(0.9, 0.999),
                 lr_final: float = 1e-5, max_steps: int = 30000, world_size: int = 1, rank: int = 0,
                 occ_update: bool = True, freeze_binaries: bool = False, start_step: int = 0):
        self.model = model
        self.world_size, self.rank = world_size, rank
        self.lr_init, self.lr_final, self.max_steps = lr, lr_final, max_steps
        self.eps, self.betas = eps, betas
        self.flat = FlatParams([p for p in model.field.parameters() if p.requires_grad])
        self.clamp_range = self.flat.range_of(model.field.endmembers)
        self.step = start_step
        self.opt_steps = 0
        self.occ_update, self.freeze_binaries = occ_update, freeze_binaries
        ops.DIRECT_GRAD = True  # backward kernels accumulate into the flat gradient views
        self._inv_world = torch.full((1,), 1.0 / world_size, dtype=torch.float32, device=self.flat.data.device)
        if world_size > 1:  # identical replicas
            dist.broadcast(self.flat.data, src=0)

    def lr_at(self, step: int) -> float:
        """nerfstudio ExponentialDecayScheduler without warm-up: lr_init * (lr_final/lr_init)^(step/max_steps)."""
        t = min(max(step / self.max_steps, 0.0), 1.0)
        return self.lr_init * (self.lr_final / self.lr_init) ** t

    def train_step(self, ray_bundle: RayBundle, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        m = self.model
        m.train()
        if self.occ_update:
            m.update_occupancy_grid(self.step, freeze_binaries=self.freeze_binaries)
        outputs = m.get_outputs(m.collider(ray_bundle), per_band_views=False)
        loss_dict = m.get_loss_dict(outputs, batch)
        loss = sum(loss_dict.values())
        loss.backward()  # kernels accumulate straight into the flat gradient buffer's views
        average_gradients(self.flat.grad, self.world_size)
        self.opt_steps += 1
        ops.adam_step(self.flat.data, self.flat.grad, self.flat.exp_avg, self.flat.exp_avg_sq, self.lr_at(self.step),
                      self.betas[0], self.betas[1], self.eps, self.opt_steps,
                      inv_scale=self._inv_world if self.world_size > 1 else None, clamp_range=self.clamp_range,
                      zero_grad=True)
        self.step += 1
        loss_dict["loss"] = loss.detach()
        loss_dict["num_samples"] = outputs["num_samples_per_ray"].sum()
        return loss_dict

    @torch.no_grad()
    def render_tiles(self, camera_ray_bundle: RayBundle, tile: int = 32) -> Dict[str, Tensor]:
        """Evaluation rendering of an [H, W] ray bundle: tiles of tile x tile pixels dealt round-robin to the ranks,
        no collective.  Returns this rank's pixels: {"index": flat pixel ids, <output name>: [n, C]}."""
        m = self.model
        m.eval()
        H, W = camera_ray_bundle.origins.shape[:2]
        dev = m.device
        index = tile_pixel_indices(H, W, tile, self.rank, self.world_size, dev)
        o = camera_ray_bundle.origins.reshape(-1, 3).to(dev)[index]
        d = camera_ray_bundle.directions.reshape(-1, 3).to(dev)[index]
        out: Dict[str, List[Tensor]] = {}
        step = int(m.config.eval_tile_rays)
        for s in range(0, index.numel(), step):
            rb = RayBundle(origins=o[s:s + step].contiguous(), directions=d[s:s + step].contiguous())
            res = m.get_outputs(m.collider(rb), per_band_views=False)
            for k, v in res.items():
                out.setdefault(k, []).append(v.reshape(v.shape[0], -1))
        result = {k: torch.cat(v) for k, v in out.items()}
        result["index"] = index
        return result
