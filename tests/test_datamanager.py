"""next_train on device (csrc/batch.cu, umhs_b200.UMHSDataManager) against oracle/datamanager_oracle.py:
pixel indices and gathered ground truth bit-exact, camera rays within fp32 rounding."""
import numpy as np
import pytest
import torch

from oracle import datamanager_oracle as DO

DEV = "cuda"


def _scene(n_img, H, W, B, C, seed, cams=None, u8=False):
    g = torch.Generator().manual_seed(seed)
    cams = cams or n_img
    hs = torch.rand(n_img, H, W, B, generator=g)
    img = torch.rand(n_img, H, W, C, generator=g)
    if u8:
        img = (img * 255).to(torch.uint8)
    # random rigid poses looking roughly at the origin, nerfstudio c2w [cams,3,4]
    q = torch.linalg.qr(torch.randn(cams, 3, 3, generator=g))[0]
    t = torch.randn(cams, 3, 1, generator=g)
    c2w = torch.cat([q, t], -1).contiguous()
    fx = 500 + 100 * torch.rand(cams, generator=g)
    fy = 500 + 100 * torch.rand(cams, generator=g)
    cx = torch.full((cams,), W / 2.0) + torch.rand(cams, generator=g)
    cy = torch.full((cams,), H / 2.0) + torch.rand(cams, generator=g)
    return hs, img, c2w, fx, fy, cx, cy


# ------------------------------------------------------------------ oracle only (CPU)
def test_oracle_pixel_sample_and_collate():
    u = torch.tensor([[0.0, 0.0, 0.0], [0.999999, 0.999999, 0.999999], [0.5, 0.25, 0.75],
                      [float(np.nextafter(np.float32(1), np.float32(0)))] * 3])
    idx = DO.pixel_sample(u, 7, 10, 12)
    assert idx.dtype == torch.int64
    assert idx.tolist() == [[0, 0, 0], [6, 9, 11], [3, 2, 9], [6, 9, 11]]
    hs, img, *_ = _scene(7, 10, 12, 5, 3, 0)
    image_idx = torch.arange(7) * 3 + 1
    b = DO.collate({"image": img, "hs_image": hs, "image_idx": image_idx}, idx)
    assert torch.equal(b["hs_image"][2], hs[3, 2, 9]) and torch.equal(b["image"][1], img[6, 9, 11])
    assert b["indices"][:, 0].tolist() == [1, 19, 10, 19] and torch.equal(b["indices"][:, 1:], idx[:, 1:])


def test_oracle_rays_are_unit_and_hit_the_principal_axis():
    _, _, c2w, fx, fy, cx, cy = _scene(2, 8, 8, 1, 3, 1)
    cx[:] = 4.5  # pixel (4, 4) has its centre on the principal point
    cy[:] = 4.5
    idx = torch.tensor([[0, 4, 4], [1, 0, 7]])
    o, d, area, norm = DO.generate_rays(c2w, fx, fy, cx, cy, idx)
    assert torch.allclose(d.norm(dim=-1), torch.ones(2), atol=1e-6)
    assert torch.allclose(d[0], -c2w[0, :, 2], atol=1e-6)            # looks down -z of camera 0
    assert torch.allclose(o, c2w[idx[:, 0], :, 3])
    assert torch.allclose(norm[0], torch.ones(1), atol=1e-6) and float(norm[1]) > 1.0
    assert torch.allclose(area[0, 0], 1.0 / (fx[0] * fy[0]), rtol=1e-3)


# ------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("n_img,H,W,B,C,dtype,u8,R", [
    (5, 40, 50, 128, 3, torch.float32, False, 4096),
    (3, 33, 17, 141, 4, torch.float32, False, 1000),    # the hotdog fixture's 141 bands: rows not 16-byte aligned
    (4, 20, 30, 128, 3, torch.float16, False, 777),
    (4, 20, 30, 64, 3, torch.bfloat16, True, 513),
    (1, 9, 9, 21, 3, torch.float16, False, 100),        # NeSpoF-style 21 bands, odd row pitch for 16-bit storage
])
def test_next_train_matches_oracle(n_img, H, W, B, C, dtype, u8, R):
    from umhs_b200 import UMHSDataManager

    hs, img, c2w, fx, fy, cx, cy = _scene(n_img, H, W, B, C, seed=R, cams=2 * n_img, u8=u8)
    hs = hs.to(dtype)
    image_idx = torch.randperm(2 * n_img, generator=torch.Generator().manual_seed(1))[:n_img]
    u = torch.rand(R, 3, generator=torch.Generator().manual_seed(2))
    u[0] = 0.0
    u[1] = float(np.nextafter(np.float32(1), np.float32(0)))  # the largest draw: clamps to the last pixel
    want = DO.next_train(u, img, hs, image_idx, c2w, fx, fy, cx, cy)

    dm = UMHSDataManager(img, hs, c2w, fx, fy, cx, cy, image_idx=image_idx, train_num_rays_per_batch=R, device=DEV)
    bundle, batch = dm.next_train(0, u=u.to(DEV))
    assert batch["indices"].dtype == torch.int64
    assert np.array_equal(batch["indices"].cpu().numpy(), want["indices"].numpy())
    assert np.array_equal(batch["hs_image"].cpu().numpy(), want["hs_image"].numpy())           # copies: bit-exact
    if u8:
        assert torch.allclose(batch["image"].cpu(), want["image"], atol=1e-7)
    else:
        assert np.array_equal(batch["image"].cpu().numpy(), want["image"].numpy())
    assert np.array_equal(bundle.origins.cpu().numpy(), want["origins"].numpy())
    assert torch.allclose(bundle.directions.cpu(), want["directions"], atol=2e-6)
    assert torch.allclose(bundle.metadata["directions_norm"].cpu(), want["directions_norm"], rtol=1e-6)
    assert torch.allclose(bundle.pixel_area.cpu(), want["pixel_area"], rtol=2e-3)  # difference of nearby unit vectors
    assert np.array_equal(bundle.camera_indices.cpu().numpy(), want["indices"][:, :1].numpy())


@pytest.mark.gpu
def test_next_train_own_draws_and_empty_batch():
    from umhs_b200 import UMHSDataManager
    from umhs_b200.umhs_datamanager import sample_batch

    hs, img, c2w, fx, fy, cx, cy = _scene(3, 16, 16, 128, 3, seed=5)
    a = UMHSDataManager(img, hs, c2w, fx, fy, cx, cy, train_num_rays_per_batch=2048, device=DEV, seed=3, rank=0)
    b = UMHSDataManager(img, hs, c2w, fx, fy, cx, cy, train_num_rays_per_batch=2048, device=DEV, seed=3, rank=1)
    (ra, ba), (rb, bb) = a.next_train(0), b.next_train(0)
    assert ba["hs_image"].shape == (2048, 128) and ra.origins.shape == (2048, 3)
    assert not torch.equal(ba["indices"], bb["indices"])                 # ranks draw different rays
    i = ba["indices"].cpu()
    assert int(i[:, 0].max()) < 3 and int(i[:, 1].max()) < 16 and int(i[:, 2].max()) < 16 and int(i.min()) >= 0
    assert torch.equal(ba["hs_image"].cpu(), hs[i[:, 0], i[:, 1], i[:, 2]])
    out = sample_batch(torch.empty((0, 3), device=DEV), a.images, a.hs_images, None, a.camera_to_worlds, a.intrinsics)
    assert out["hs_image"].shape == (0, 128)


@pytest.mark.gpu
def test_scene_from_disk_feeds_next_train_and_a_training_step(tmp_path):
    """SURVEY 8f rows 4 -> 2 -> hot path: transforms.json + .npy cubes on disk -> hs_data.load_scene -> device-resident
    UMHSDataManager.next_train -> UMHSTrainer.train_step.  The gathered hs / rgb rows are exactly the pixels the batch
    indices name."""
    import umhs_b200
    from test_hs_data import _write_scene
    from umhs_b200 import hs_data
    from umhs_b200.trainer import UMHSTrainer

    _write_scene(str(tmp_path), n=6, H=12, W=16, B=32)
    scene = hs_data.load_scene(str(tmp_path), split="train")
    dm = hs_data.datamanager_from_scene(scene, "cuda", train_num_rays_per_batch=256, seed=3)
    bundle, batch = dm.next_train(0)
    idx = batch["indices"].cpu()
    want_hs = scene.hs_images[idx[:, 0], idx[:, 1], idx[:, 2]]
    want_rgb = scene.images[idx[:, 0], idx[:, 1], idx[:, 2]]
    assert torch.equal(batch["hs_image"].cpu(), want_hs) and torch.equal(batch["image"].cpu(), want_rgb)
    assert bundle.origins.shape == (256, 3) and torch.allclose(bundle.directions.norm(dim=-1).cpu(), torch.ones(256), atol=1e-5)
    cfg = umhs_b200.UMHSConfig(log2_hashmap_size=10, grid_resolution=32, temperature=0.4)
    m = cfg.setup(scene_box=scene.scene_box, num_train_data=len(scene.image_filenames),
                  metadata={**scene.metadata, "num_classes": 4}).to("cuda")
    tr = UMHSTrainer(m, occ_update=True)
    losses = []
    for step in range(3):
        bundle, batch = dm.next_train(step)
        losses.append(float(tr.train_step(bundle, batch)["loss"]))
    assert all(np.isfinite(losses))
