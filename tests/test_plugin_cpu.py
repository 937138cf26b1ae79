"""The drop-in boundary against a stub of nerfstudio: ``umhs_b200.umhs_config.umhs_method`` builds (its
``_build()`` executes), ``UMHSConfig`` / ``UMHSModel`` become subclasses of nerfstudio's ``InstantNGPModelConfig`` /
``NGPModel`` when nerfstudio is importable (ref umhs_model.py:62,122), and the packaging names the reference's entry
point (ref pyproject.toml:12-13).  nerfstudio itself cannot be installed here: ``tests/ns_stub`` carries its base
classes, reduced; the import runs in a subprocess so that the stub never leaks into the other tests."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "unsupervised-hyperspectral-nerf_b200")
STUB = os.path.join(ROOT, "tests", "ns_stub")

_SCRIPT = r"""
import dataclasses, json, sys
import torch
from nerfstudio.models.instant_ngp import InstantNGPModelConfig, NGPModel
from nerfstudio.plugins.types import MethodSpecification
import umhs_b200
from umhs_b200 import umhs_model
from umhs_b200.umhs_config import umhs_method          # module __getattr__ -> _build()

assert umhs_model.HAVE_NERFSTUDIO
assert isinstance(umhs_method, MethodSpecification)
tc = umhs_method.config
cfg = tc.pipeline.model
assert isinstance(cfg, umhs_b200.UMHSConfig) and isinstance(cfg, InstantNGPModelConfig)
assert issubclass(umhs_b200.UMHSModel, NGPModel) and cfg._target is umhs_b200.UMHSModel
# the reference's trainer settings (ref umhs_config.py:34-69)
assert tc.method_name == "umhsnerf" and tc.max_num_iterations == 30000 and tc.mixed_precision is True
assert tc.pipeline.datamanager.train_num_rays_per_batch == 9216 * 4 and cfg.eval_num_rays_per_chunk == 512
opt = tc.optimizers["fields"]
assert opt["optimizer"].lr == 2e-2 and opt["optimizer"].eps == 1e-15 and opt["scheduler"].lr_final == 1e-5
# the reference's own defaults survive the multiple inheritance (ref umhs_model.py:61-119), not InstantNGP's
assert cfg.use_gradient_scaling is True and cfg.use_appearance_embedding is True and cfg.temperature == 0.2
# the model is constructed the way the reference's pipeline does it (ref umhs_pipeline.py:98-105) ...
cfg2 = dataclasses.replace(cfg, log2_hashmap_size=8, grid_resolution=16)
class Box:
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
m = cfg2.setup(scene_box=Box(), num_train_data=3, metadata={"wavelengths": list(range(400, 528)), "num_classes": 4},
               device="cpu", grad_scaler=None, num_classes=4, wavelengths=list(range(400, 528)))
# ... through nerfstudio's Model.__init__ (collider reset to None there, then populate_modules installs ours)
assert isinstance(m, NGPModel) and m.collider is not None and m.callbacks is None
assert m.field.endmembers.shape == (4, 128) and m.device.type == "cpu"
cbs = m.get_training_callbacks(None)
assert len(cbs) == 2 and {type(c).__name__ for c in cbs} == {"TrainingCallback"}
assert list(m.get_param_groups()) == ["fields"]
print(json.dumps({"ok": True, "bases": [c.__name__ for c in type(m).__mro__[:4]]}))
"""


def test_plugin_builds_against_nerfstudio_stub():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([STUB, PKG, ROOT]))
    r = subprocess.run([sys.executable, "-c", _SCRIPT], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ok"] and out["bases"][:2] == ["UMHSModel", "NGPModel"]


def test_without_nerfstudio_the_stand_ins_are_used_and_umhs_method_says_why():
    import pytest

    import umhs_b200
    from umhs_b200 import umhs_config, umhs_model

    assert not umhs_model.HAVE_NERFSTUDIO and issubclass(umhs_b200.UMHSModel, __import__("torch").nn.Module)
    with pytest.raises(ModuleNotFoundError, match="nerfstudio"):
        umhs_config.umhs_method  # noqa: B018


def test_packaging_names_the_reference_entry_point():
    import re

    txt = open(os.path.join(ROOT, "pyproject.toml")).read()
    assert "[project.entry-points.'nerfstudio.method_configs']" in txt
    assert re.search(r"umhsnerf\s*=\s*'umhs_b200\.umhs_config:umhs_method'", txt)
    assert 'where = ["unsupervised-hyperspectral-nerf_b200"]' in txt
