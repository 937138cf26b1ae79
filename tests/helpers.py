"""Shared helpers for the test-suite: load golden fixtures, build the CPU oracle from them."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle import nerf3p, umhs_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WAVELENGTHS = np.linspace(400, 1000, 128).astype(int).tolist()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    meta = json.loads(str(d.pop("meta"))) if "meta" in d else {}
    return d, meta


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def oracle_model_from_golden(d, meta, training=True):
    cfg = O.OracleConfig(method="rgb+spectral", temperature=meta["temperature"], pred_specular=meta["specular"],
                         log2_hashmap_size=meta["log2_T"], grid_resolution=meta.get("grid_res", 32))
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    m = O.OracleModel(cfg, aabb, WAVELENGTHS, meta["K"])
    sd = {k[len("param."):]: t(v) for k, v in d.items() if k.startswith("param.")}
    missing, unexpected = m.field.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    if "binaries" in d:
        res = meta["grid_res"]
        bits = np.unpackbits(d["binaries"])[: 4 * res ** 3].reshape(4, res, res, res).astype(bool)
        m.occupancy_grid.binaries.copy_(t(bits))
        m.occupancy_grid.occs.copy_(m.occupancy_grid.binaries.flatten().float() * float(d["occs_scale"]))
    m.train(training)
    return m


# ------------------------------------------------------------------ B200 side
def b200_model_from_golden(d, meta, training=True, device="cuda"):
    """UMHSModel (the product) carrying the golden parameters / occupancy grid."""
    import umhs_b200

    cfg = umhs_b200.UMHSConfig(method="rgb+spectral", temperature=meta["temperature"],
                               pred_specular=meta["specular"], log2_hashmap_size=meta["log2_T"],
                               grid_resolution=meta.get("grid_res", 32))
    m = cfg.setup(scene_box=torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_train_data=10,
                  metadata={"wavelengths": WAVELENGTHS, "num_classes": meta["K"]})
    sd = {k[len("param."):]: t(v) for k, v in d.items() if k.startswith("param.")}
    missing, unexpected = m.field.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("converter" in k for k in missing), missing
    if "binaries" in d:
        res = meta["grid_res"]
        bits = np.unpackbits(d["binaries"])[: 4 * res ** 3].reshape(4, res, res, res).astype(bool)
        m.occupancy_grid.binaries.copy_(t(bits))
        m.occupancy_grid.occs.copy_(m.occupancy_grid.binaries.flatten().float() * float(d["occs_scale"]))
        m.occupancy_grid.mark_dirty()
    m = m.to(device)
    m.train(training)
    return m


def assert_close(a, b, rtol, atol, name=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    a = a.reshape(b.shape).astype(np.float64)
    b = b.astype(np.float64)
    err = np.abs(a - b)
    bad = err > atol + rtol * np.abs(b)
    if bad.any():
        i = np.unravel_index(np.argmax(err - (atol + rtol * np.abs(b))), b.shape) if b.ndim else ()
        raise AssertionError(
            f"{name}: {int(bad.sum())}/{bad.size} elements differ (rtol={rtol}, atol={atol:.3g}); worst at {i}: "
            f"got {a[i]:.9g} want {b[i]:.9g}; max|err|={err.max():.3g}, max|want|={np.abs(b).max():.3g}, "
            f"rel L2 err={np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30):.3g}")


import contextlib


@contextlib.contextmanager
def mlp_precision(mode):
    """Select the MLP kernel family for a test: "fp32" (SIMT, tight parity) or "bf16" (tcgen05 tensor cores)."""
    from umhs_b200 import ops
    old = ops.MLP_PRECISION
    ops.MLP_PRECISION = mode
    try:
        yield
    finally:
        ops.MLP_PRECISION = old


def rel_l2(a, b):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a.reshape(b.shape) - b) / (np.linalg.norm(b) + 1e-30))
