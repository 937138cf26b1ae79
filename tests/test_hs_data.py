"""On-disk dataset side: transforms.json with the hyperspectral extension + H x W x B .npy cubes ->
tensors for UMHSDataManager (ref umhsnerf/data/umhs_dataparser.py:43-65,222-228,302-332; ref
umhsnerf/data/utils/hs_dataloader.py:46-58), and the VCA endmember initialiser behind ``vca.npy``."""
import json
import os

import numpy as np
import pytest
import torch

from umhs_b200 import hs_data


def _write_scene(root, n=6, H=5, W=7, B=21, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "images"))
    os.makedirs(os.path.join(root, "hs"))
    frames, cubes, rgbs = [], {}, {}
    for i in range(n):
        name = f"frame_{'train' if i % 3 else 'eval'}_{i:03d}"
        cube = rng.uniform(-0.2, 1.3, size=(H, W, B)).astype(np.float32)  # out-of-range values exercise the clamp
        rgb = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        np.save(os.path.join(root, "hs", name + ".npy"), cube)
        np.save(os.path.join(root, "images", name + ".npy"), rgb)
        c2w = np.eye(4, dtype=np.float32)
        c2w[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        c2w[:3, 3] = rng.normal(size=3) * 3 + np.array([5.0, 0, 0])
        frames.append({"file_path": f"images/{name}.npy", "hyperspectral_file_path": f"hs/{name}.npy",
                       "transform_matrix": c2w.tolist()})
        cubes[name], rgbs[name] = cube, rgb
    wl = np.linspace(450, 650, B).astype(int).tolist()
    meta = {"fl_x": 30.0, "fl_y": 31.0, "cx": W / 2, "cy": H / 2, "w": W, "h": H, "wavelengths": wl,
            "frames": frames[::-1]}  # unsorted on disk
    with open(os.path.join(root, "transforms.json"), "w") as fh:
        json.dump(meta, fh)
    return cubes, rgbs, wl


def test_scene_round_trip(tmp_path):
    cubes, rgbs, wl = _write_scene(str(tmp_path))
    tr = hs_data.load_scene(str(tmp_path), split="train")
    ev = hs_data.load_scene(os.path.join(tmp_path, "transforms.json"), split="val")
    assert tr.wavelengths == wl and tr.metadata["wavelengths"] == wl
    assert tr.hs_images.shape == (4, 5, 7, 21) and ev.hs_images.shape == (2, 5, 7, 21)
    assert all("train" in os.path.basename(f) for f in tr.image_filenames)
    assert all("eval" in os.path.basename(f) for f in ev.image_filenames)
    for k, f in enumerate(tr.hs_filenames):  # cubes: float, clamped to [0, 1]; images: uint8 / 255
        name = os.path.basename(f)[:-4]
        assert torch.equal(tr.hs_images[k], torch.from_numpy(cubes[name]).clamp(0, 1))
        assert torch.equal(tr.images[k], torch.from_numpy(rgbs[name]).float() / 255.0)
    assert float(tr.hs_images.min()) == 0.0 and float(tr.hs_images.max()) == 1.0
    assert torch.equal(tr.fx, torch.full((4,), 30.0)) and torch.equal(tr.cy, torch.full((4,), 2.5))
    assert torch.equal(tr.scene_box, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]))
    # poses: centred + scaled over ALL frames so that the largest |translation| is 1, "up" = mean camera y-axis -> +z
    allp = hs_data.load_scene(str(tmp_path), split="all").camera_to_worlds
    assert abs(float(allp[:, :3, 3].abs().max()) - 1.0) < 1e-5
    up = allp[:, :3, 1].mean(0)
    assert float(up[:2].abs().max()) < 1e-5 and float(up[2]) > 0
    R = allp[:, :3, :3]
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand_as(R), atol=1e-5)  # still rotations
    assert abs(tr.dataparser_scale - ev.dataparser_scale) < 1e-12


def test_scene_errors(tmp_path):
    _write_scene(str(tmp_path))
    p = os.path.join(tmp_path, "transforms.json")
    meta = json.load(open(p))
    bad = dict(meta)
    del bad["wavelengths"]
    json.dump(bad, open(p, "w"))
    with pytest.raises(KeyError, match="Wavelengths"):
        hs_data.load_scene(str(tmp_path))
    bad = dict(meta, wavelengths=meta["wavelengths"][:-1])
    json.dump(bad, open(p, "w"))
    with pytest.raises(ValueError, match="bands"):
        hs_data.load_scene(str(tmp_path))
    bad = dict(meta, frames=[dict(f, file_path=f["file_path"].replace("train", "foo")) for f in meta["frames"]])
    json.dump(bad, open(p, "w"))
    with pytest.raises((ValueError, FileNotFoundError)):
        hs_data.load_scene(str(tmp_path))


def test_vca_recovers_pure_pixels_and_writes_vca_npy(tmp_path):
    """Linear mixtures of K spectra with pure pixels present: VCA must pick the pure pixels; ``vca.npy`` is [K, B] and is
    what ``UMHSField(load_vca=True)`` loads (ref umhs_field.py:77-81)."""
    rng = np.random.default_rng(3)
    K, B, H, W = 4, 30, 16, 16
    E = rng.uniform(0.05, 0.95, size=(K, B))
    A = rng.dirichlet(np.ones(K) * 0.7, size=H * W)
    A[:K] = np.eye(K)                       # pure pixels
    Y = (A @ E).reshape(H, W, B).astype(np.float32)
    for snr in (0.0, 5.0):                  # both projection branches (the reference's copy breaks on the low-SNR one)
        Ae, idx, Yp = hs_data.vca(Y.reshape(-1, B).T, K, snr_input=snr, seed=0)
        assert Ae.shape == (B, K) and Yp.shape == (B, H * W)
        assert sorted(idx.tolist()) == list(range(K))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        out = hs_data.write_vca_npy(torch.from_numpy(Y), K, seed=1)
        saved = np.load("vca.npy")
        assert saved.shape == (K, B) and np.array_equal(saved, out)
        err = np.abs(saved[:, None, :] - E[None]).max(-1).min(1)  # every recovered spectrum matches one endmember
        assert err.max() < 1e-4
        import umhs_b200

        f = umhs_b200.UMHSField(torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), wavelengths=B, num_classes=K, load_vca=True,
                                log2_hashmap_size=8)
        assert torch.equal(f.endmembers.detach(), torch.from_numpy(saved))
    finally:
        os.chdir(cwd)
