"""Device-resident training data: the step in front of the hot path.

Mirror of ``UMHSDataManager.next_train`` (ref umhsnerf/data/umhs_datamanager.py:95-108): pixel sampler -> per-pixel
gather of ``image`` / ``hs_image`` -> ray generator, returning ``(RayBundle, batch)`` with the reference's batch keys
(``image``, ``hs_image``, ``indices``).  The reference draws pixel indices on the image batch's device, gathers on the
CPU (nerfstudio moves ``c, y, x`` to the CPU) and copies R x B floats to the GPU every step (8 MB at 16 384 rays x 128
bands); here the RGB images, the ``H x W x B`` cubes (ref umhsnerf/data/utils/hs_dataloader.py:46-50: float, clamped
to [0, 1]) and the cameras live in HBM and one kernel (csrc/batch.cu) produces the whole batch.  100 images of
800 x 800 x 128 bands are 32.8 GB in fp32, 16.4 GB in fp16 / bf16 -- both fit the 180 GB of one B200.

Reading the ``.npy`` cubes / ``transforms.json`` stays with nerfstudio's dataparser (out of scope).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import ptr, stream_ptr
from .ops import _call
from .rays import RayBundle

_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.uint8: 3}


def sample_batch(u: Tensor, images: Optional[Tensor], hs_images: Optional[Tensor], image_idx: Optional[Tensor],
                 c2w: Optional[Tensor], intrinsics: Optional[Tensor]) -> Dict[str, Tensor]:
    """One launch of ``umhs_sample_batch``.  u [R,3] fp32 uniform draws; images [N,H,W,3|4] fp32 / uint8;
    hs_images [N,H,W,B] fp32 / fp16 / bf16; image_idx [N] int64 or None; c2w [cams,3,4]; intrinsics [cams,4]
    (fx, fy, cx, cy).  -> indices, image, hs_image, origins, directions, pixel_area, directions_norm."""
    _lib.load()
    ref = hs_images if hs_images is not None else images
    if ref is None:
        raise _lib.UmhsError("sample_batch needs images and / or hs_images")
    dev = u.device
    R = int(u.shape[0])
    N, H, W = (int(s) for s in ref.shape[:3])
    bi, bo = _lib.BatchIn(), _lib.BatchOut()
    bi.u, bi.n_rays, bi.num_images, bi.height, bi.width = ptr(u), R, N, H, W
    out: Dict[str, Tensor] = {"indices": torch.empty((R, 3), dtype=torch.int64, device=dev)}
    bo.indices = ptr(out["indices"])
    if hs_images is not None:
        if hs_images.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise _lib.UmhsError("hs_images must be fp32, fp16 or bf16")
        bi.hs, bi.num_bands, bi.hs_dtype = ptr(hs_images), int(hs_images.shape[-1]), _DTYPES[hs_images.dtype]
        out["hs_image"] = torch.empty((R, int(hs_images.shape[-1])), dtype=torch.float32, device=dev)
        bo.hs_image = ptr(out["hs_image"])
    if images is not None:
        if images.dtype not in (torch.float32, torch.uint8) or tuple(images.shape[:3]) != (N, H, W):
            raise _lib.UmhsError("images must be fp32 or uint8 with the cubes' [N,H,W]")
        bi.image, bi.image_channels, bi.image_dtype = ptr(images), int(images.shape[-1]), _DTYPES[images.dtype]
        out["image"] = torch.empty((R, int(images.shape[-1])), dtype=torch.float32, device=dev)
        bo.image = ptr(out["image"])
    bi.image_idx = ptr(image_idx)
    if c2w is not None:
        bi.c2w, bi.intrinsics = ptr(c2w), ptr(intrinsics)
        for k, shape in (("origins", (R, 3)), ("directions", (R, 3)), ("pixel_area", (R, 1)),
                         ("directions_norm", (R, 1))):
            out[k] = torch.empty(shape, dtype=torch.float32, device=dev)
            setattr(bo, k, ptr(out[k]))
    _call("umhs_sample_batch", R, ctypes.byref(bi), ctypes.byref(bo), stream_ptr(dev))
    return out


class UMHSDataManager:
    """``next_train(step) -> (RayBundle, batch)`` over device-resident data.

    images [N,H,W,3|4] (fp32 in [0,1] or uint8), hs_images [N,H,W,B] (stored as given: fp32 / fp16 / bf16),
    camera_to_worlds [cams,3,4], fx / fy / cx / cy [cams] or scalars, image_idx [N] (camera of each image; None =
    identity).  ``world_size`` / ``rank`` only decorrelate the random streams of data-parallel ranks: every rank
    draws its own ``train_num_rays_per_batch`` rays (ray-sharded data parallelism, no collective)."""

    def __init__(self, images: Tensor, hs_images: Tensor, camera_to_worlds: Tensor, fx, fy, cx, cy,
                 image_idx: Optional[Tensor] = None, train_num_rays_per_batch: int = 9216 * 4, device=None,
                 seed: int = 0, rank: int = 0, world_size: int = 1):
        dev = torch.device(device) if device is not None else images.device
        self.device = dev
        self.images = images.to(dev).contiguous()
        self.hs_images = hs_images.to(dev).contiguous()
        cams = int(camera_to_worlds.shape[0])
        self.camera_to_worlds = camera_to_worlds.to(dev, torch.float32)[:, :3, :4].contiguous()

        def per_cam(v):
            return torch.as_tensor(v, dtype=torch.float32).reshape(-1).expand(cams) if torch.as_tensor(v).numel() == 1 \
                else torch.as_tensor(v, dtype=torch.float32).reshape(cams)

        self.intrinsics = torch.stack([per_cam(fx), per_cam(fy), per_cam(cx), per_cam(cy)], -1).to(dev).contiguous()
        self.image_idx = image_idx.to(dev, torch.int64).contiguous() if image_idx is not None else None
        self.train_num_rays_per_batch = int(train_num_rays_per_batch)  # ref umhs_config.py:46 (9216 * 4)
        self.train_count = 0
        self.rank, self.world_size = rank, world_size
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(seed * 7919 + rank)

    def next_train(self, step: int = 0, u: Optional[Tensor] = None) -> Tuple[RayBundle, Dict[str, Tensor]]:
        """u: optional [R,3] uniform draws (tests); default: this manager's own generator."""
        self.train_count += 1
        if u is None:
            u = torch.rand((self.train_num_rays_per_batch, 3), generator=self._gen, device=self.device)
        o = sample_batch(u.to(self.device, torch.float32).contiguous(), self.images, self.hs_images, self.image_idx,
                         self.camera_to_worlds, self.intrinsics)
        bundle = RayBundle(origins=o["origins"], directions=o["directions"], pixel_area=o["pixel_area"],
                           camera_indices=o["indices"][:, 0:1], metadata={"directions_norm": o["directions_norm"]})
        batch = {"image": o["image"], "hs_image": o["hs_image"], "indices": o["indices"]}
        return bundle, batch
