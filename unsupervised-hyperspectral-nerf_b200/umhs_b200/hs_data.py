"""On-disk dataset side of SURVEY 8f row 4: nerfstudio ``transforms.json`` with the reference's hyperspectral extension
-> the device-resident tensors ``UMHSDataManager`` samples from.

What is restated, and from where:

* ``transforms.json``: per frame ``file_path``, ``transform_matrix`` and ``hyperspectral_file_path`` (ref
  umhsnerf/data/umhs_dataparser.py:222-228), top level ``wavelengths`` (ref :318-320, ``int`` per band), intrinsics
  ``fl_x fl_y cx cy w h`` at top level or per frame (ref :334-339); train / eval split by the words ``train`` / ``eval``
  in the file name (``eval_mode="filename"``, ref :43-65, :96); poses oriented ("up") and centred ("poses") as
  nerfstudio's ``auto_orient_and_center_poses`` does [3P], then scaled into the +-1 box (``auto_scale_poses``, ref
  :302-308); scene box ``[-scene_scale, scene_scale]^3`` (ref :327-332).
* hyperspectral cubes: ``np.load`` of an ``H x W x B`` array, float, clamped to [0, 1] (ref
  umhsnerf/data/utils/hs_dataloader.py:46-50).
* ``vca.npy``: Vertex Component Analysis of the first cube, saved as ``[K, B]`` (ref hs_dataloader.py:52-58 ->
  ``data/utils/vca.py``); read back by ``UMHSField(load_vca=True)`` (ref umhs_field.py:77-81).  The algorithm is
  restated from its publication (Nascimento & Dias, IEEE TGRS 43(4), 2005); the reference's copy leaves the projection
  unassigned when ``SNR < SNR_th`` and ``verbose=False`` and swallows the exception -- here both branches
  work.

This is host code that runs once per scene (file I/O + one SVD); nothing here is on the per-step path.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor


# ---------------------------------------------------------------------------------------------- transforms.json
def split_by_filename(image_filenames: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """ref umhs_dataparser.py:43-65: indices of frames whose basename contains ``train`` / ``eval``."""
    i_train, i_eval = [], []
    for idx, name in enumerate(image_filenames):
        base = os.path.basename(str(name))
        if "train" in base:
            i_train.append(idx)
        elif "eval" in base:
            i_eval.append(idx)
        else:
            raise ValueError("frame should contain train/eval in its name to use this eval-frame-index eval mode")
    return np.array(i_train, dtype=np.int64), np.array(i_eval, dtype=np.int64)


def _rotation_between(a: Tensor, b: Tensor) -> Tensor:
    """nerfstudio camera_utils.rotation_matrix_between: the rotation that takes unit vector a to unit vector b."""
    a, b = a / a.norm(), b / b.norm()
    v = torch.linalg.cross(a, b)
    eps = 1e-6
    if float(v.abs().sum()) < eps:  # parallel: perturb, as nerfstudio does
        x = torch.tensor([1.0, 0.0, 0.0]) if abs(float(a[0])) < eps else torch.tensor([0.0, 1.0, 0.0])
        v = torch.linalg.cross(a, x)
    v = v / v.norm()
    skew = torch.tensor([[0.0, -float(v[2]), float(v[1])], [float(v[2]), 0.0, -float(v[0])],
                         [-float(v[1]), float(v[0]), 0.0]])
    theta = torch.acos(torch.clamp(torch.dot(a, b), -1, 1))
    return torch.eye(3) + torch.sin(theta) * skew + (1 - torch.cos(theta)) * (skew @ skew)


def orient_and_center(poses: Tensor, method: str = "up", center_method: str = "poses") -> Tuple[Tensor, Tensor]:
    """nerfstudio auto_orient_and_center_poses for methods "up" / "none" and centring "poses" / "none".
    poses [N, 3|4, 4] camera-to-world -> (oriented poses [N, 3, 4], the applied [3, 4] transform)."""
    poses = poses[:, :3, :4].to(torch.float32)
    origins = poses[:, :3, 3]
    translation = origins.mean(dim=0) if center_method == "poses" else torch.zeros(3)
    if method == "up":
        up = poses[:, :3, 1].mean(dim=0)
        rot = _rotation_between(up / up.norm(), torch.tensor([0.0, 0.0, 1.0]))
    elif method == "none":
        rot = torch.eye(3)
    else:
        raise NotImplementedError(f"orientation_method={method!r} (built: 'up', 'none')")
    transform = torch.cat([rot, rot @ -translation[:, None]], dim=-1)  # [3, 4]
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]]).expand(poses.shape[0], 1, 4)
    return transform @ torch.cat([poses, bottom], dim=1), transform


def load_hs_cube(path: str) -> Tensor:
    """ref hs_dataloader.py:49-50: ``H x W x B`` ``.npy`` -> float32 in [0, 1]."""
    cube = np.load(path)
    if cube.ndim != 3:
        raise ValueError(f"{path}: expected an H x W x B cube, got shape {cube.shape}")
    return torch.from_numpy(np.ascontiguousarray(cube)).float().clamp(0, 1)


def load_rgb(path: str) -> Tensor:
    """An RGB(A) image as float32 [H, W, 3|4] in [0, 1]: ``.npy`` arrays, or anything PIL opens when PIL is there."""
    if path.endswith(".npy"):
        a = np.load(path)
    else:
        from PIL import Image  # the reference's own dependency for image files

        a = np.array(Image.open(path))
    a = torch.from_numpy(np.ascontiguousarray(a))
    if a.dtype == torch.uint8:
        a = a.float() / 255.0
    a = a.float()
    return a[..., None].expand(-1, -1, 3) if a.ndim == 2 else a


@dataclass
class HSScene:
    """One split of a hyperspectral scene, ready for ``UMHSDataManager`` / ``UMHSModel``."""
    images: Tensor               # [N, H, W, 3|4] float32
    hs_images: Tensor            # [N, H, W, B] float32 (or the dtype asked for)
    camera_to_worlds: Tensor     # [N, 3, 4]
    fx: Tensor
    fy: Tensor
    cx: Tensor
    cy: Tensor
    wavelengths: List[int]
    scene_box: Tensor            # [2, 3]
    image_filenames: List[str] = field(default_factory=list)
    hs_filenames: List[str] = field(default_factory=list)
    dataparser_scale: float = 1.0
    dataparser_transform: Optional[Tensor] = None

    @property
    def metadata(self) -> Dict:
        """What the reference's dataparser hands the model (ref umhs_dataparser.py:478-496)."""
        return {"wavelengths": self.wavelengths, "hs_filenames": self.hs_filenames}


def load_scene(data: str, split: str = "train", scale_factor: float = 1.0, scene_scale: float = 1.0,
               orientation_method: str = "up", center_method: str = "poses", auto_scale_poses: bool = True,
               eval_mode: str = "filename", hs_dtype: torch.dtype = torch.float32) -> HSScene:
    """Parse ``<data>/transforms.json`` (or an explicit json path) and load the frames of ``split`` ("train" / "val" /
    "test"; "all" = every frame).  Poses are oriented / centred / scaled over ALL frames before the split, as the
    reference does (ref umhs_dataparser.py:296-312)."""
    json_path = data if data.endswith(".json") else os.path.join(data, "transforms.json")
    data_dir = os.path.dirname(os.path.abspath(json_path))
    with open(json_path) as fh:
        meta = json.load(fh)
    frames = meta["frames"]
    # nerfstudio sorts the frames by file name
    frames = sorted(frames, key=lambda f: f["file_path"])
    if not all("hyperspectral_file_path" in f for f in frames):
        raise KeyError("every frame needs 'hyperspectral_file_path' (ref umhs_dataparser.py:222-225)")
    if "wavelengths" not in meta:
        raise KeyError("Wavelengths not specified in metadata")  # ref umhs_dataparser.py:319
    wavelengths = [int(x) for x in meta["wavelengths"]]
    image_filenames = [os.path.join(data_dir, f["file_path"]) for f in frames]
    hs_filenames = [os.path.join(data_dir, f["hyperspectral_file_path"]) for f in frames]
    poses = torch.tensor(np.array([f["transform_matrix"] for f in frames], dtype=np.float32))
    poses, transform = orient_and_center(poses, orientation_method, center_method)
    scale = 1.0
    if auto_scale_poses:
        scale /= float(torch.max(torch.abs(poses[:, :3, 3])))
    scale *= scale_factor
    poses[:, :3, 3] *= scale

    if split == "all" or eval_mode == "all":
        idx = np.arange(len(frames))
    elif eval_mode == "filename":
        i_train, i_eval = split_by_filename(image_filenames)
        idx = i_train if split == "train" else i_eval
    else:
        raise NotImplementedError(f"eval_mode={eval_mode!r} (built: 'filename', 'all')")

    def intr(key):
        if key in meta:
            return torch.full((len(idx),), float(meta[key]))
        return torch.tensor([float(frames[i][key]) for i in idx])

    images = torch.stack([load_rgb(image_filenames[i]) for i in idx])
    cubes = torch.stack([load_hs_cube(hs_filenames[i]) for i in idx]).to(hs_dtype)
    if cubes.shape[-1] != len(wavelengths):
        raise ValueError(f"cubes hold {cubes.shape[-1]} bands, transforms.json lists {len(wavelengths)} wavelengths")
    if tuple(cubes.shape[1:3]) != tuple(images.shape[1:3]):
        raise ValueError(f"cube size {tuple(cubes.shape[1:3])} != image size {tuple(images.shape[1:3])}")
    s = float(scene_scale)
    return HSScene(images=images, hs_images=cubes, camera_to_worlds=poses[torch.as_tensor(idx)],
                   fx=intr("fl_x"), fy=intr("fl_y"), cx=intr("cx"), cy=intr("cy"), wavelengths=wavelengths,
                   scene_box=torch.tensor([[-s, -s, -s], [s, s, s]]), image_filenames=[image_filenames[i] for i in idx],
                   hs_filenames=[hs_filenames[i] for i in idx], dataparser_scale=scale, dataparser_transform=transform)


def datamanager_from_scene(scene: HSScene, device, train_num_rays_per_batch: int = 9216 * 4, **kw):
    """The scene's tensors moved into HBM behind ``UMHSDataManager.next_train``."""
    from .umhs_datamanager import UMHSDataManager

    return UMHSDataManager(scene.images, scene.hs_images, scene.camera_to_worlds, scene.fx, scene.fy, scene.cx,
                           scene.cy, train_num_rays_per_batch=train_num_rays_per_batch, device=device, **kw)


# ---------------------------------------------------------------------------------------------- VCA
def vca(Y: np.ndarray, R: int, snr_input: float = 0.0, seed: Optional[int] = None
        ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Vertex Component Analysis.  Y [B, N] (bands x pixels), R endmembers -> (Ae [B, R] endmember signatures,
    indices [R] of the chosen pixels, Yp [B, N] the data projected onto the signal subspace)."""
    Y = np.asarray(Y, dtype=np.float64)
    B, N = Y.shape
    if not (0 < R <= B):
        raise ValueError("R must be in (0, B]")
    rng = np.random.default_rng(seed)
    y_m = Y.mean(axis=1, keepdims=True)
    Y_o = Y - y_m
    if snr_input == 0:
        Ud = np.linalg.svd(Y_o @ Y_o.T / N)[0][:, :R]
        x_p = Ud.T @ Y_o
        P_y, P_x = np.sum(Y ** 2) / N, np.sum(x_p ** 2) / N + float((y_m.T @ y_m)[0, 0])
        snr = 10 * np.log10(max((P_x - R / B * P_y) / max(P_y - P_x, 1e-300), 1e-300))
    else:
        snr = snr_input
    snr_th = 15 + 10 * np.log10(R)
    if snr < snr_th:  # projective projection onto R - 1 dimensions
        d = R - 1
        Ud = np.linalg.svd(Y_o @ Y_o.T / N)[0][:, :d]
        x_p = Ud.T @ Y_o
        Yp = Ud @ x_p + y_m
        x = x_p
        c = np.sqrt(np.max(np.sum(x ** 2, axis=0)))
        y = np.vstack([x, c * np.ones((1, N))])
    else:  # projection onto R dimensions followed by a perspective projection
        d = R
        Ud = np.linalg.svd(Y @ Y.T / N)[0][:, :d]
        x_p = Ud.T @ Y
        Yp = Ud @ x_p
        u = x_p.mean(axis=1, keepdims=True)
        y = x_p / np.maximum(np.sum(x_p * u, axis=0, keepdims=True), 1e-300)
    indices = np.zeros(R, dtype=np.int64)
    A = np.zeros((R, R))
    A[-1, 0] = 1
    for i in range(R):
        w = rng.random((R, 1))
        f = w - A @ (np.linalg.pinv(A) @ w)
        f = f / np.linalg.norm(f)
        v = f.T @ y
        indices[i] = int(np.argmax(np.abs(v)))
        A[:, i] = y[:, indices[i]]
    return Yp[:, indices], indices, Yp


def write_vca_npy(hs_image: Tensor, num_classes: int, path: str = "vca.npy", seed: Optional[int] = None) -> np.ndarray:
    """ref hs_dataloader.py:52-58: VCA of one ``H x W x B`` cube, saved as ``[K, B]`` for ``UMHSField(load_vca=True)``."""
    Y = hs_image.permute(2, 0, 1).reshape(hs_image.shape[2], -1).cpu().numpy()
    Ae, _, _ = vca(Y, num_classes, seed=seed)
    np.save(path, Ae.T.astype(np.float32))
    return Ae.T.astype(np.float32)
