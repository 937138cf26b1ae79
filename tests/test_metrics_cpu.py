"""Evaluation-image metrics of ``UMHSModel.get_image_metrics_and_images`` (ref umhs_model.py:407-527; torchmetrics /
nerfstudio colormaps restated in umhs_b200/metrics.py): known answers + an independent direct-convolution SSIM."""
import math

import numpy as np
import torch

import umhs_b200
from umhs_b200 import metrics as M


def _ssim_direct(a, b, k=11, sigma=1.5):
    """Loop implementation of torchmetrics' default SSIM for one [C, H, W] pair (reflect pad, gaussian window, crop)."""
    a, b = a.double().numpy(), b.double().numpy()
    L = max(a.max() - a.min(), b.max() - b.min())
    c1, c2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    g = np.exp(-(np.arange(k) - (k - 1) / 2) ** 2 / (2 * sigma ** 2))
    g /= g.sum()
    w2 = np.outer(g, g)
    pad = (k - 1) // 2
    vals = []
    for c in range(a.shape[0]):
        pa, pb = np.pad(a[c], pad, mode="reflect"), np.pad(b[c], pad, mode="reflect")
        H, W = pa.shape
        for y in range(pad, H - k + 1 - pad):      # positions whose windows survive the final crop
            for x in range(pad, W - k + 1 - pad):
                A, Bw = pa[y:y + k, x:x + k], pb[y:y + k, x:x + k]
                ma, mb = (w2 * A).sum(), (w2 * Bw).sum()
                va, vb, cab = (w2 * A * A).sum() - ma * ma, (w2 * Bw * Bw).sum() - mb * mb, (w2 * A * Bw).sum() - ma * mb
                vals.append(((2 * ma * mb + c1) * (2 * cab + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2)))
    return float(np.mean(vals))


def test_psnr_ssim_sam_known_answers():
    g = torch.Generator().manual_seed(0)
    a = torch.rand(1, 3, 24, 20, generator=g)
    b = (a + 0.1 * torch.randn(1, 3, 24, 20, generator=g)).clamp(0, 1)
    assert abs(float(M.psnr(a, a + 0.1)) - 20.0) < 1e-4                  # mse = 0.01 -> 20 dB
    assert abs(float(M.ssim(a, a)) - 1.0) < 1e-6
    assert abs(float(M.ssim(a, b)) - _ssim_direct(a[0], b[0])) < 1e-5
    p = torch.tensor([[1.0, 0, 0], [1, 1, 0], [2, 0, 0]]).t().reshape(1, 3, 3, 1)   # [1, C=3, H=3, W=1]
    t = torch.tensor([[0.0, 1, 0], [1, 0, 0], [1, 0, 0]]).t().reshape(1, 3, 3, 1)
    ang = M.spectral_angle(p, t).reshape(-1)
    assert torch.allclose(ang, torch.tensor([math.pi / 2, math.pi / 4, 0.0]), atol=1e-6)
    assert abs(float(M.compute_sam(p[0].permute(1, 2, 0), t[0].permute(1, 2, 0))) - math.pi / 4) < 1e-5
    assert math.isnan(float(M.spectral_angle(torch.zeros(1, 3, 1, 1), torch.ones(1, 3, 1, 1))))  # nan-meaned by the model
    cm = M.apply_colormap(torch.linspace(0, 1, 11).reshape(11, 1))
    assert cm.shape == (11, 3) and float(cm.min()) >= 0 and float(cm.max()) <= 1
    assert float(cm[1, 2]) > float(cm[1, 0]) and float(cm[-2, 0]) > float(cm[-2, 2])  # turbo: blue -> red


def test_get_image_metrics_and_images_surface():
    """Keys, shapes and values of the reference's evaluation surface on a CPU model (no kernels are involved)."""
    cfg = umhs_b200.UMHSConfig(log2_hashmap_size=8, grid_resolution=16, background_color="white")
    m = cfg.setup(metadata={"wavelengths": list(range(400, 528)), "num_classes": 4})
    g = torch.Generator().manual_seed(1)
    H, W, B = 16, 18, 128
    outputs = {"rgb": torch.rand(H, W, 3, generator=g), "accumulation": torch.rand(H, W, 1, generator=g),
               "depth": torch.rand(H, W, 1, generator=g) + 0.5, "spectral": torch.rand(H, W, B, generator=g),
               "seg_raw": torch.zeros(H, W, 1), "seg_pred": torch.zeros(H, W, 3)}
    rgba = torch.rand(H, W, 4, generator=g)
    batch = {"image": rgba, "hs_image": torch.rand(H, W, B, generator=g)}
    md, im = m.get_image_metrics_and_images(outputs, batch)
    assert set(md) == {"psnr", "ssim", "lpips", "psnr_spectral", "ssim_spectral", "sam_spectral", "rmse_spectral"}
    assert set(im) == {"img", "accumulation", "depth", "se_per_pixel"}
    assert im["img"].shape == (H, 2 * W, 3) and im["depth"].shape == (H, W, 3) and im["se_per_pixel"].shape == (H, W, 1)
    gt = rgba[..., :3] * rgba[..., 3:] + (1 - rgba[..., 3:])            # RGBA over the white background
    assert abs(md["psnr"] - float(10 * torch.log10(1 / torch.mean((gt - outputs["rgb"]) ** 2)))) < 1e-4
    assert abs(md["rmse_spectral"] - float(torch.sqrt(torch.mean((outputs["spectral"] - batch["hs_image"]) ** 2)))) < 1e-6
    assert abs(md["sam_spectral"] - float(M.compute_sam(outputs["spectral"], batch["hs_image"]))) < 1e-4
    assert m.i == 1 and (math.isnan(md["lpips"]) or md["lpips"] >= 0)


def test_get_outputs_for_camera_goes_through_generate_rays():
    """ref umhs_model.py:529-540: camera.generate_rays(camera_indices=0, keep_shape=True, obb_box=...) -> ray bundle ->
    get_outputs_for_camera_ray_bundle."""
    cfg = umhs_b200.UMHSConfig(log2_hashmap_size=8, grid_resolution=16)
    m = cfg.setup(metadata={"wavelengths": list(range(400, 528)), "num_classes": 4})
    seen = {}

    class Cam:
        def generate_rays(self, camera_indices, keep_shape, obb_box=None):
            seen["args"] = (camera_indices, keep_shape, obb_box)
            return "bundle"

    m.get_outputs_for_camera_ray_bundle = lambda rb: {"got": rb}
    assert m.get_outputs_for_camera(Cam(), obb_box="box") == {"got": "bundle"}
    assert seen["args"] == (0, True, "box")
