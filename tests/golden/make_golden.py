"""Freeze golden vectors by running the UNMODIFIED reference code (/root/reference).

Run once, in the build container:   python tests/golden/make_golden.py
(needs /root/reference; the GPU box and the test-suite only read the committed .npz).

How: ``ref_shim.install()`` provides the absent third-party modules (nerfstudio, nerfacc,
...) from ``oracle/nerf3p.py``; the reference's own classes then run as shipped:

  colour.npz       ColourSystem (ref utils/spec_to_rgb.py), ClusterLookup (ref
                   utils/clusterprobe.py), get_weights_spectral + dense
                   SpectralRenderer.forward (ref umhs_renderer.py)
  field_*.npz      UMHSField.get_density / get_outputs  (ref umhs_field.py:151-329)
  model_*.npz      UMHSModel.get_outputs / get_loss_dict (ref umhs_model.py:225-383)
                   incl. sampler + packing, and gradients of the summed loss.

Every file stores inputs, all parameters (small hash table: log2_hashmap_size=10) and the
reference's outputs, so the tests need neither /root/reference nor an RNG replay.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from umhsnerf.umhs_model import UMHSConfig  # noqa: E402
from umhsnerf.umhs_renderer import SpectralRenderer, get_weights_spectral  # noqa: E402
from umhsnerf.utils.clusterprobe import ClusterLookup  # noqa: E402
from umhsnerf.utils.spec_to_rgb import ColourSystem  # noqa: E402

from oracle import nerf3p  # noqa: E402

WAVELENGTHS = np.linspace(400, 1000, 128).astype(int).tolist()
E_HOTDOG = np.load("/root/reference/endmembers_hotdog.npy")[:, :128].astype(np.float32)  # 4x141 -> first 128 bands


def npify(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.detach().cpu().numpy()
        else:
            out[k] = np.asarray(v)
    return out


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **npify(arrs))
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(arrs)} arrays")


def make_colour():
    torch.manual_seed(10)
    cs = ColourSystem(bands=WAVELENGTHS, cs="sRGB")
    spec = torch.rand(64, 128) * torch.linspace(0.0, 1.5, 64)[:, None]
    spec[0] = 0.0
    spec[1] = 1e-5
    rgb = cs(spec)
    probe = ClusterLookup(128, 4)
    E = torch.from_numpy(E_HOTDOG)
    inner, probs = probe(spec, alpha=0.2, clusters=E)
    deltas = torch.rand(16, 24, 1) * 0.05
    dens = torch.exp(torch.randn(16, 24, 1) * 2)
    w = get_weights_spectral(deltas, dens)
    vals = torch.rand(16, 24, 128)
    dense_acc = SpectralRenderer.forward(vals, w)
    save("colour.npz", wavelengths=np.array(WAVELENGTHS), transform_matrix=cs.transform_matrix, spec=spec, rgb=rgb,
         endmembers=E, inner=inner, probs=probs, deltas=deltas, densities=dens, weights=w, vals=vals,
         dense_acc=dense_acc)


def build_model(K, temperature, specular, log2_T=10, grid_res=32, seed=0):
    torch.manual_seed(seed)
    cfg = UMHSConfig(method="rgb+spectral", temperature=temperature, pred_specular=specular,
                     log2_hashmap_size=log2_T, grid_resolution=grid_res, implementation="torch")
    sb = ref_shim.SceneBox(torch.tensor([[-1.0, -1, -1], [1, 1, 1]]))
    model = cfg.setup(scene_box=sb, num_train_data=10, metadata={"wavelengths": WAVELENGTHS, "num_classes": K},
                      num_classes=K)
    with torch.no_grad():
        f = model.field
        f.mlp_base.encoder.hash_table.mul_(1000.0)  # U(-1,1): non-degenerate features
        f.mlp_base.mlp.layers[1].bias[0] = 3.0  # sigma ~ e^3: visible, partially terminating rays
        if K == 4:
            f.endmembers.copy_(torch.from_numpy(E_HOTDOG))
        else:
            f.endmembers.copy_(torch.rand(K, 128))
    return cfg, model


def params_of(model, prefix="param."):
    return {prefix + k: v for k, v in model.state_dict().items()
            if k.startswith("field.") and "converter" not in k}


def make_field(name, K, temperature, specular, N=384, seed=1):
    cfg, model = build_model(K, temperature, specular, seed=seed)
    f = model.field
    g = torch.Generator().manual_seed(seed + 100)
    origins = (torch.rand(N, 3, generator=g) * 2 - 1) * 1.2
    dirs = torch.randn(N, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    starts = torch.rand(N, 1, generator=g) * 1.5
    ends = starts + torch.rand(N, 1, generator=g) * 0.05
    origins[:4] = 0.0  # contraction / selector edge cases
    starts[:4] = 0.0
    ends[:4] = 0.0
    origins[4:8] *= 40.0  # far outside: contraction active
    rs = nerf3p.RaySamples(frustums=nerf3p.Frustums(origins, dirs, starts, ends),
                           camera_indices=torch.zeros(N, 1, dtype=torch.long))
    density, emb = f.get_density(rs)
    out = f.get_outputs(rs, density_embedding=emb)
    gw = torch.randn(N, 128, generator=g)
    gd = torch.randn(N, 1, generator=g) * 0.01
    spec = out["spectral"]
    loss = (spec.reshape(N, 128) * gw).sum() + (density * gd).sum()
    if specular:
        loss = loss + 0.5 * (out["spectral2"] * gw.flip(0)).sum()
    loss.backward()
    grads = {"grad." + k: p.grad for k, p in f.named_parameters() if p.grad is not None}
    arrs = dict(meta=json.dumps(dict(K=K, temperature=temperature, specular=specular, log2_T=10, N=N)),
                origins=origins, directions=dirs, starts=starts, ends=ends, gw=gw, gd=gd,
                density=density, embedding=emb, spectral=spec, abundances=out["abundances"], loss=loss)
    if specular:
        arrs.update(spectral2=out["spectral2"], specular=out["specular"])
    arrs.update({k.replace("param.field.", "param."): v for k, v in params_of(model).items()})
    arrs.update(grads)
    save(name, **arrs)


def make_model(name, K, temperature, specular, training, R=96, grid_res=32, seed=2):
    cfg, model = build_model(K, temperature, specular, grid_res=grid_res, seed=seed)
    model.train(training)
    g = torch.Generator().manual_seed(seed + 200)
    # occupancy: a ball in level 0 and a wider blob in level 1, so rays cross two levels
    gridc = (torch.arange(grid_res) + 0.5) / grid_res
    X, Y, Z = torch.meshgrid(gridc, gridc, gridc, indexing="ij")
    occ = model.occupancy_grid
    for lvl, rad in ((0, 0.55), (1, 1.4)):
        ab = occ.aabbs[lvl]
        x = ab[0] + X * (ab[3] - ab[0])
        y = ab[1] + Y * (ab[4] - ab[1])
        z = ab[2] + Z * (ab[5] - ab[2])
        occ.binaries[lvl] = (x ** 2 + y ** 2 + z ** 2) < rad ** 2
    occ.occs.copy_(occ.binaries.flatten().float() * 0.05)
    rad = 0.7 + 2.0 * torch.rand(R, 1, generator=g)
    o = torch.randn(R, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * rad
    d = -o / o.norm(dim=-1, keepdim=True) + 0.15 * torch.randn(R, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    d[0] = torch.tensor([1.0, 0.0, 0.0])  # axis-aligned ray: zero direction components
    o[0] = torch.tensor([-1.5, 0.01, 0.02])
    d[1] = torch.tensor([0.0, 0.0, 1.0])  # misses everything occupied? still marches empty cells
    o[1] = torch.tensor([5.0, 5.0, -3.0])
    jitter = torch.rand(R, generator=g)
    bg = torch.rand(R, 3, generator=g)
    image = torch.rand(R, 3, generator=g)
    hs = torch.rand(R, 128, generator=g)
    rb = nerf3p.RayBundle(origins=o, directions=d, camera_indices=torch.zeros(R, 1, dtype=torch.long))
    rb = model.collider(rb)
    model.renderer_rgb.bg_override = bg
    # capture what the sampler produced (get_outputs does not return it)
    real = nerf3p.VolumetricSampler(model.occupancy_grid, model.field.density_fn)
    real.training = training
    captured = {}

    def fwd(ray_bundle, render_step_size, near_plane=0.0, far_plane=None, alpha_thre=0.01, cone_angle=0.0):
        rs, ri = real(ray_bundle, render_step_size, near_plane, far_plane, alpha_thre, cone_angle, jitter=jitter)
        captured["ray_indices"] = ri
        captured["t_starts"] = rs.frustums.starts[:, 0]
        captured["t_ends"] = rs.frustums.ends[:, 0]
        return rs, ri

    model.sampler.forward = fwd

    outputs = model.get_outputs(rb)
    batch = {"image": image, "hs_image": hs}
    loss_dict = model.get_loss_dict(outputs, batch)
    arrs = dict(meta=json.dumps(dict(K=K, temperature=temperature, specular=specular, log2_T=10, R=R,
                                     grid_res=grid_res, training=training,
                                     render_step_size=float(cfg.render_step_size))),
                origins=o, directions=d, jitter=jitter, bg=bg, image=image, hs_image=hs,
                binaries=np.packbits(occ.binaries.numpy().astype(np.uint8).ravel()), occs_scale=np.float32(0.05),
                aabbs=occ.aabbs, **captured)
    for k in ("accumulation", "depth", "spectral", "rgb", "num_samples_per_ray", "abundances", "seg_probs",
              "seg_raw", "seg_pred", "spectral2", "specular"):
        if k in outputs:
            arrs["out." + k] = outputs[k]
    for k, v in loss_dict.items():
        arrs["loss." + k] = v
    if training:
        total = sum(loss_dict.values())
        total.backward()
        arrs.update({"grad." + k: p.grad for k, p in model.field.named_parameters() if p.grad is not None})
    arrs.update({k.replace("param.field.", "param."): v for k, v in params_of(model).items()})
    print(name, "samples:", int(captured["ray_indices"].numel()), "mean/ray:",
          float(outputs["num_samples_per_ray"].float().mean()), "acc mean:", float(outputs["accumulation"].mean()))
    save(name, **arrs)


if __name__ == "__main__":
    make_colour()
    make_field("field_k4.npz", K=4, temperature=0.4, specular=False)
    make_field("field_k6_spec.npz", K=6, temperature=0.3, specular=True, seed=3)
    make_model("model_train_k4.npz", K=4, temperature=0.4, specular=False, training=True)
    make_model("model_train_k4_spec.npz", K=4, temperature=0.5, specular=True, training=True, seed=5)
    make_model("model_eval_k9_spec.npz", K=9, temperature=0.3, specular=True, training=False, seed=7)
