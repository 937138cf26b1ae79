"""nerfstudio checkpoint format for the UMHS model.

nerfstudio's Trainer writes ``<dir>/step-{step:09d}.ckpt`` = ``torch.save({"step", "pipeline", "optimizers",
"schedulers", "scalers"})`` where ``pipeline`` is ``pipeline.state_dict()`` (every key under ``_model.``, plus
``module.`` when DDP-wrapped) and ``optimizers["fields"]`` is the ``torch.optim.Adam`` state dict of
``list(field.parameters())`` (ref umhsnerf/umhs_config.py:59-64).  The reference's own hook on that path is
``UMHSPipeline.load_pipeline`` (ref umhsnerf/umhs_pipeline.py:157-175): strip ``module.``, ``update_to_step``,
``load_state_dict``, dump the endmembers to ``endmembers_hotdog.npy``.

The model mirror keeps the reference's state-dict names and shapes, so loading is a key translation; entries of
modules that are not part of the hot path (torchmetrics / LPIPS buffers, a datamanager's camera optimiser) are
reported, not loaded.  Optimiser state goes straight into the trainer's flat Adam buffers.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

# order of ``field.parameters()`` in the REFERENCE (direct parameters first, then sub-modules in registration order:
# NerfactoField creates mlp_base and mlp_head, UMHSField then adds feature_mlp and mlp_directional -- ref
# umhsnerf/umhs_field.py:51,67,81-85,95,105); Adam's per-parameter state is indexed in this order
_REF_MODULE_ORDER = ("endmembers", "mlp_base.", "mlp_head.", "feature_mlp.", "mlp_directional.")
_IGNORED_PREFIXES = ("lpips.", "psnr.", "ssim.", "sam.", "field.embedding_appearance.", "camera_optimizer.")


def checkpoint_name(step: int) -> str:
    return f"step-{step:09d}.ckpt"


def reference_param_order(field) -> List[str]:
    names = [n for n, _ in field.named_parameters()]
    out: List[str] = []
    for prefix in _REF_MODULE_ORDER:
        out += [n for n in names if (n == prefix or n.startswith(prefix)) and n not in out]
    return out + [n for n in names if n not in out]


def model_state_from_pipeline(pipeline_state: Dict[str, Tensor]) -> Tuple[Dict[str, Tensor], List[str]]:
    """``pipeline.state_dict()`` keys -> ``model.state_dict()`` keys.  -> (state, skipped keys)."""
    state, skipped = {}, []
    for key, value in pipeline_state.items():
        k = key[len("module."):] if key.startswith("module.") else key  # ref umhs_pipeline.py:165-167
        if not k.startswith("_model."):
            skipped.append(key)  # datamanager entries (camera optimiser)
            continue
        k = k[len("_model."):]
        if k.startswith(_IGNORED_PREFIXES):
            skipped.append(key)
            continue
        state[k] = value
    return state, skipped


def load_pipeline(model, loaded_state: Dict[str, Tensor], step: int, endmembers_npy: Optional[str] = None
                  ) -> List[str]:
    """Mirror of ``UMHSPipeline.load_pipeline`` for the model part.  Shapes must match (``K``, ``B``, table size and
    grid resolution are fixed at construction, as in the reference); returns the keys that were not loaded.
    endmembers_npy: where to dump the endmembers like ref umhs_pipeline.py:175 (None: do not write a file)."""
    state, skipped = model_state_from_pipeline(loaded_state)
    tcnn = [k for k in state if "tcnn_encoding" in k or k.endswith(".params")]
    if tcnn:
        raise ValueError(
            "this checkpoint was trained with implementation='tcnn' (flat tiny-cuda-nn parameter blobs such as "
            f"{tcnn[0]!r}): tiny-cuda-nn's grid (dense coarse levels, no biases, fp16 MLP layout) is a different "
            "function from nerfstudio's torch path, which is what this package reproduces (north star: the PyTorch "
            "path).  Re-train or export with --pipeline.model.implementation torch.")
    own = model.state_dict()
    # the sampler holds the same occupancy grid module: reference checkpoints carry it under both names
    for k in list(own):
        if k not in state and k.startswith("sampler.occupancy_grid.") and k[len("sampler."):] in state:
            state[k] = state[k[len("sampler."):]]
    if "device_indicator_param" in own and "device_indicator_param" not in state:
        state["device_indicator_param"] = own["device_indicator_param"]
    unknown = [k for k in state if k not in own]
    for k in unknown:
        skipped.append("_model." + k)
        del state[k]
    missing = [k for k in own if k not in state]
    if missing:
        raise KeyError(f"checkpoint lacks {missing[:6]}{' ...' if len(missing) > 6 else ''}")
    for k, v in state.items():
        if tuple(v.shape) != tuple(own[k].shape):
            raise ValueError(f"{k}: checkpoint shape {tuple(v.shape)} != model shape {tuple(own[k].shape)}")
    model.step = step  # nerfstudio Model.update_to_step is a no-op for NGP models; keep the counter in sync
    with torch.no_grad():  # in place: the trainer's flat buffer aliases the parameters
        for k, v in state.items():
            own[k].copy_(v.to(own[k].device, own[k].dtype))
    if hasattr(model.occupancy_grid, "mark_dirty"):
        model.occupancy_grid.mark_dirty()
    if endmembers_npy is not None:
        np.save(endmembers_npy, model.field.endmembers.detach().cpu().numpy())
    return skipped


def _adam_to_flat(trainer, opt_state: Dict[str, Any]) -> int:
    """torch.optim.Adam state dict (reference parameter order) -> the trainer's flat exp_avg / exp_avg_sq."""
    field = trainer.model.field
    ref_names = reference_param_order(field)
    ids = opt_state["param_groups"][0]["params"]
    by_name = dict(field.named_parameters())
    steps = 0
    trainer.flat.exp_avg.zero_()
    trainer.flat.exp_avg_sq.zero_()
    for pid, name in zip(ids, ref_names):
        st = opt_state["state"].get(pid)
        if st is None:  # a parameter that never received a gradient (e.g. mlp_directional without specular)
            continue
        a, b = trainer.flat.range_of(by_name[name])
        trainer.flat.exp_avg[a:b].copy_(st["exp_avg"].reshape(-1))
        trainer.flat.exp_avg_sq[a:b].copy_(st["exp_avg_sq"].reshape(-1))
        steps = max(steps, int(st["step"]))
    return steps


def _adam_from_flat(trainer) -> Dict[str, Any]:
    field = trainer.model.field
    by_name = dict(field.named_parameters())
    state = {}
    names = reference_param_order(field)
    # fused data-parallel step: every rank holds the moments of its own shard only -> gather them (a collective: every
    # rank must call save_checkpoint; single-GPU / all-reduce mode: the local buffers are complete)
    exp_avg, exp_avg_sq = trainer.full_optimizer_state()
    for i, name in enumerate(names):
        p = by_name[name]
        a, b = trainer.flat.range_of(p)
        state[i] = {"step": torch.tensor(float(trainer.opt_steps)),
                    "exp_avg": exp_avg[a:b].reshape(p.shape).detach().cpu().clone(),
                    "exp_avg_sq": exp_avg_sq[a:b].reshape(p.shape).detach().cpu().clone()}
    group = {"lr": trainer.lr_at(trainer.step), "betas": tuple(trainer.betas), "eps": trainer.eps, "weight_decay": 0,
             "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
             "fused": None, "params": list(range(len(names)))}
    return {"state": state, "param_groups": [group]}


def save_checkpoint(directory: str, model, trainer=None, step: Optional[int] = None,
                    extras_from: Optional[str] = None, write: bool = True) -> str:
    """Write ``step-{step:09d}.ckpt`` in nerfstudio's layout.

    What a nerfstudio ``Trainer._load_checkpoint`` + the reference's ``UMHSPipeline.load_pipeline`` (a STRICT
    ``load_state_dict``, ref umhs_pipeline.py:157-175) find in it: every ``_model.*`` entry of the hot path, the
    occupancy grid under both of its names (``_model.occupancy_grid.*`` and ``_model.sampler.occupancy_grid.*``), a
    complete ``torch.optim.Adam`` state for ``"fields"`` in the reference's parameter order, a ``LambdaLR`` scheduler
    state and an enabled-``GradScaler`` state (the reference trains with ``mixed_precision=True``: ref
    umhs_config.py:40; an empty dict makes ``GradScaler.load_state_dict`` raise).  What it can NOT contain by itself:
    the buffers of modules that are not part of this package -- the pretrained LPIPS network of
    ``NGPModel.lpips`` (``_model.lpips.*``, torchmetrics) and a datamanager's camera optimiser.  ``extras_from`` = a
    checkpoint written by the reference: its entries that this model does not own are copied over, which makes the
    file pass the reference's strict load.  With a fused data-parallel trainer EVERY rank must call this (the Adam
    moments are gathered); ``write=False`` on the ranks that should not write the file."""
    step = int(step if step is not None else (trainer.step if trainer is not None else getattr(model, "step", 0)))
    pipeline = {"_model." + k: v.detach().cpu() for k, v in model.state_dict().items()}
    if extras_from is not None:
        ref = torch.load(extras_from, map_location="cpu", weights_only=False)["pipeline"]
        for key, value in ref.items():
            k = key[len("module."):] if key.startswith("module.") else key
            if k not in pipeline:
                pipeline[k] = value
    ckpt: Dict[str, Any] = {"step": step, "pipeline": pipeline, "optimizers": {}, "schedulers": {},
                            # torch.amp.GradScaler.state_dict() of an enabled scaler (defaults; bf16/f32 kernels here
                            # never overflow, so the scale is never consulted by this package)
                            "scalers": {"scale": 65536.0, "growth_factor": 2.0, "backoff_factor": 0.5,
                                        "growth_interval": 2000, "_growth_tracker": 0}}
    if trainer is not None:
        ckpt["optimizers"]["fields"] = _adam_from_flat(trainer)
        lr = trainer.lr_at(step + 1)
        # torch.optim.lr_scheduler.LambdaLR.state_dict() (nerfstudio ExponentialDecayScheduler): plain lambdas are
        # stored as None and kept by the loading side
        ckpt["schedulers"]["fields"] = {"base_lrs": [trainer.lr_init], "last_epoch": step + 1, "verbose": False,
                                        "_step_count": step + 2, "_get_lr_called_within_step": False,
                                        "_last_lr": [lr], "lr_lambdas": [None]}
    path = os.path.join(directory, checkpoint_name(step))
    if write:
        os.makedirs(directory, exist_ok=True)
        torch.save(ckpt, path)
    return path


def load_checkpoint(path: str, model, trainer=None, endmembers_npy: Optional[str] = None) -> Dict[str, Any]:
    """Load a nerfstudio checkpoint (a file, or a directory: its latest ``step-*.ckpt``).  With a trainer the Adam
    moments and the step counters are restored too.  -> {"step", "skipped"}."""
    if os.path.isdir(path):
        files = sorted(f for f in os.listdir(path) if f.startswith("step-") and f.endswith(".ckpt"))
        if not files:
            raise FileNotFoundError(f"no step-*.ckpt in {path}")
        path = os.path.join(path, files[-1])
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    step = int(ckpt.get("step", 0))
    skipped = load_pipeline(model, ckpt["pipeline"], step, endmembers_npy)
    if trainer is not None:
        trainer.step = step + 1  # nerfstudio resumes at the step after the saved one
        opt = ckpt.get("optimizers", {}).get("fields")
        if opt is not None:
            trainer.opt_steps = _adam_to_flat(trainer, opt)
    return {"step": step, "skipped": skipped}
