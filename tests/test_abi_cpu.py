"""CPU: the C-ABI library loads and exports every symbol include/umhs.h declares; host-side mirrors import and
keep the reference's names; the product fails loudly without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    from umhs_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol():
    lib_path = _build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "umhs.h")).read()
    declared = set(re.findall(r"\b(umhs_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in umhs.h but not exported"
    from umhs_b200 import _lib
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert _lib.load().umhs_version() >= 100


def test_argument_validation_without_gpu():
    """Entry points validate arguments before touching the device: error codes + message, no crash."""
    from umhs_b200 import _lib
    lib = _lib.load()
    rc = lib.umhs_hash_encode_fwd(None, 10, None, None, None)
    assert rc == -1 and b"umhs_hash_encode_fwd" in lib.umhs_last_error()
    assert lib.umhs_adam_step(None, None, None, None, 4, 0.1, 0.9, 0.999, 1e-8, 1, None, None, 0, 0, 1, None) == -1
    assert lib.umhs_composite_fwd(None, None, None) == -1


def test_no_cpu_fallback():
    from umhs_b200 import ops, _lib
    with pytest.raises(_lib.UmhsError):
        ops.hash_encode_x01_fwd(torch.rand(4, 3), torch.rand(16 << 4, 2), torch.ones(16), 4)


def test_host_mirror_names_and_state_dict():
    import umhs_b200
    cfg = umhs_b200.UMHSConfig(log2_hashmap_size=8, grid_resolution=16, pred_specular=True)
    m = cfg.setup(metadata={"wavelengths": list(range(400, 528)), "num_classes": 4})
    keys = set(m.state_dict().keys())
    for k in ("field.mlp_base.encoder.hash_table", "field.mlp_base.encoder.scalings",
              "field.mlp_base.mlp.layers.1.bias", "field.mlp_head.layers.2.weight", "field.feature_mlp.layers.2.weight",
              "field.mlp_directional.layers.1.weight", "field.endmembers", "field.aabb", "field.max_res",
              "field.num_levels", "field.log2_hashmap_size", "field.converter.transform_matrix",
              "converter.transform_matrix", "scene_aabb", "occupancy_grid.occs", "occupancy_grid.binaries",
              "occupancy_grid.aabbs", "occupancy_grid.resolution"):
        assert k in keys, k
    assert m.field.feature_mlp.layers[2].weight.shape == (5, 64)
    assert m.field.mlp_base.encoder.hash_table.shape == (16 << 8, 2)
    assert abs(m.config.render_step_size - (12 ** 0.5) / 1000) < 1e-9
    assert umhs_b200.UMHSRenderer is umhs_b200.SpectralRenderer
    assert set(m.get_param_groups()) == {"fields"}


def test_colour_and_probe_mirrors_match_golden():
    from helpers import load_golden, t
    import numpy as np
    import umhs_b200
    d, _ = load_golden("colour.npz")
    cs = umhs_b200.ColourSystem(d["wavelengths"].tolist())
    np.testing.assert_allclose(cs.transform_matrix.numpy(), d["transform_matrix"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(cs(t(d["spec"])).numpy(), d["rgb"], rtol=1e-5, atol=1e-6)
    inner, probs = umhs_b200.ClusterLookup(128, 4)(t(d["spec"]), alpha=0.2, clusters=t(d["endmembers"]))
    np.testing.assert_allclose(inner.numpy(), d["inner"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(probs.numpy(), d["probs"], rtol=1e-5, atol=1e-6)


def test_activation_tile_size_agrees_across_header_and_host():
    """UMHS_ACT_TILE_BYTES (include/umhs.h) is what ops.field_mlp_fwd allocates per 128-sample tile: five [128 x 64] and
    one [128 x 32] bf16 operand tiles."""
    import re

    import umhs_b200.ops as ops

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "umhs.h")).read()
    m = re.search(r"#define\s+UMHS_ACT_TILE_BYTES\s+(\d+)", hdr)
    assert m and int(m.group(1)) == ops.ACT_TILE_BYTES == 5 * 128 * 64 * 2 + 128 * 32 * 2

