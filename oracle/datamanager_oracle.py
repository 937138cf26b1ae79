"""CPU restatement of the step in front of the hot path: ``UMHSDataManager.next_train``
(ref umhsnerf/data/umhs_datamanager.py:95-108) = pixel sampler -> per-pixel gather of every
batch entry (``image``, ``hs_image``) -> ray generator.  SURVEY.md 8(f) row 2.

TEST INFRASTRUCTURE ONLY (see oracle/nerf3p.py): only ``tests/`` may import this, as the checker
of ``umhs_sample_batch`` (csrc/batch.cu).

The reference's own code on this path is three lines (:103-106); the arithmetic is nerfstudio
1.1.5's (pinned ref nerfstudioa100_environment.yml:189, not vendored, not installed here).
PARITY UNPINNED: restated from recollection of the public sources --

* ``PixelSampler.sample_method``: ``indices = floor(rand(R, 3) * [num_images, H, W]).long()``
  (fp32 product, truncation).  The uniform draws are an INPUT here (``u``), so the random stream
  is not part of the comparison.  An fp32 product that rounds up to the bound would index out of
  range in nerfstudio; this restatement (and the kernel) clamp to bound - 1.
* ``PixelSampler.collate_image_dataset_batch``: ``{key: value[c, y, x]}`` for every image-like
  entry; ``indices[:, 0] = batch["image_idx"][c]`` (position in the resident batch -> dataset /
  camera index).  ``hs_image`` is the ``H x W x B`` cube of ref data/utils/hs_dataloader.py:46-50
  (``np.load(...)``, float, clamped to [0, 1]).
* ``RayGenerator.forward``: ``coords = image_coords[y, x]`` (pixel centres, +0.5) ->
  ``Cameras.generate_rays``: perspective, no distortion:
  ``dir = ((x - cx) / fx, -(y - cy) / fy, -1)``; rotated by ``c2w[:3, :3]``; normalised
  (``normalize_with_norm``, eps 1e-? irrelevant for finite inputs); ``origins = c2w[:3, 3]``;
  ``pixel_area = |d - d(x+1)| * |d - d(y+1)|``; ``metadata["directions_norm"]``.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor


def pixel_sample(u: Tensor, num_images: int, height: int, width: int) -> Tensor:
    """u [R,3] fp32 in [0,1) -> (image slot, row, column) int64 [R,3]."""
    bound = torch.tensor([num_images, height, width], dtype=torch.float32)
    idx = (u.to(torch.float32) * bound).long()
    return torch.minimum(idx, (bound.long() - 1))


def collate(batch: Dict[str, Tensor], indices: Tensor) -> Dict[str, Tensor]:
    c, y, x = indices[:, 0], indices[:, 1], indices[:, 2]
    out = {k: v[c, y, x] for k, v in batch.items() if k != "image_idx" and v is not None}
    ind = indices.clone()
    if batch.get("image_idx") is not None:
        ind[:, 0] = batch["image_idx"][c]
    out["indices"] = ind
    return out


def generate_rays(c2w: Tensor, fx: Tensor, fy: Tensor, cx: Tensor, cy: Tensor, indices: Tensor
                  ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """indices [R,3] = (camera, row, column).  -> origins [R,3], directions [R,3], pixel_area [R,1],
    directions_norm [R,1]; all fp32."""
    cam, yi, xi = indices[:, 0], indices[:, 1], indices[:, 2]
    y = yi.to(torch.float32) + 0.5
    x = xi.to(torch.float32) + 0.5
    fx, fy, cx, cy = fx[cam], fy[cam], cx[cam], cy[cam]
    coord = torch.stack([(x - cx) / fx, -(y - cy) / fy], -1)
    coord_x = torch.stack([(x - cx + 1) / fx, -(y - cy) / fy], -1)
    coord_y = torch.stack([(x - cx) / fx, -(y - cy + 1) / fy], -1)
    stack = torch.stack([coord, coord_x, coord_y], 0)                                   # [3,R,2]
    dirs = torch.cat([stack, -torch.ones_like(stack[..., :1])], -1)                      # [3,R,3]
    rot = c2w[cam][:, :3, :3]                                                            # [R,3,3]
    dirs = torch.sum(dirs[..., None, :] * rot[None], dim=-1)                             # [3,R,3]
    norm = torch.linalg.vector_norm(dirs, dim=-1, keepdim=True)
    dirs = dirs / torch.maximum(norm, torch.tensor([1e-10]))
    origins = c2w[cam][:, :3, 3]
    d = dirs[0]
    dx = torch.sqrt(torch.sum((d - dirs[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((d - dirs[2]) ** 2, dim=-1))
    return origins, d, (dx * dy)[:, None], norm[0]


def next_train(u: Tensor, images: Tensor, hs_images: Tensor, image_idx: Optional[Tensor], c2w: Tensor, fx: Tensor,
               fy: Tensor, cx: Tensor, cy: Tensor) -> Dict[str, Tensor]:
    """The whole step: draws u [R,3] -> {"indices", "image", "hs_image", "origins", "directions", "pixel_area",
    "directions_norm"}.  images [N,H,W,3|4] fp32, hs_images [N,H,W,B] (any float dtype, returned as fp32)."""
    n, h, w = images.shape[:3]
    idx = pixel_sample(u, n, h, w)
    out = collate({"image": images, "hs_image": hs_images, "image_idx": image_idx}, idx)
    out["hs_image"] = out["hs_image"].to(torch.float32)
    if out["image"].dtype == torch.uint8:
        out["image"] = out["image"].to(torch.float32) / 255.0
    o, d, area, norm = generate_rays(c2w, fx, fy, cx, cy, out["indices"])
    out.update(origins=o, directions=d, pixel_area=area, directions_norm=norm)
    return out
