"""nerfstudio checkpoint layout: a checkpoint written the way the reference's
trainer writes it (``module._model.`` keys, Adam state in the reference's parameter order, metric buffers mixed in)
loads into the mirror + the trainer's flat Adam buffers; what ``save_checkpoint`` writes has the same layout."""
import os

import numpy as np
import pytest
import torch

import umhs_b200
from umhs_b200 import checkpoint as ck
from umhs_b200.trainer import UMHSTrainer


def _model(seed, specular=False):
    torch.manual_seed(seed)
    cfg = umhs_b200.UMHSConfig(log2_hashmap_size=8, grid_resolution=16, pred_specular=specular)
    m = cfg.setup(metadata={"wavelengths": list(range(400, 528)), "num_classes": 4})
    with torch.no_grad():
        for p in m.field.parameters():
            p.copy_(torch.randn_like(p))
        m.occupancy_grid.occs.copy_(torch.rand_like(m.occupancy_grid.occs))
        m.occupancy_grid.binaries.copy_(torch.rand(m.occupancy_grid.binaries.shape) > 0.5)
    return m


def _reference_style_checkpoint(m, step):
    """What nerfstudio's Trainer.save_checkpoint produces for this model (DDP-wrapped pipeline)."""
    names = ck.reference_param_order(m.field)
    assert names[0] == "endmembers" and names[1].startswith("mlp_base.")
    assert names.index("mlp_head.layers.0.weight") < names.index("feature_mlp.layers.0.weight")  # parent's slot first
    by_name = dict(m.field.named_parameters())
    params = [by_name[n] for n in names]
    opt = torch.optim.Adam(params, lr=2e-2, eps=1e-15)
    for it in range(3):
        for i, p in enumerate(params):
            # mlp_directional receives no gradient without the specular head: Adam keeps no state for it
            p.grad = None if names[i].startswith("mlp_directional.") else torch.randn_like(p) * (it + 1)
        opt.step()
    pipeline = {"module._model." + k: v.clone() for k, v in m.state_dict().items() if not k.startswith("sampler.")}
    pipeline["module._model.lpips.net.scaling_layer.shift"] = torch.zeros(1, 3, 1, 1)
    pipeline["module.datamanager.train_camera_optimizer.pose_adjustment"] = torch.zeros(4, 6)
    return {"step": step, "pipeline": pipeline, "optimizers": {"fields": opt.state_dict()},
            "schedulers": {"fields": {}}, "scalers": {}}, opt, names, params


def test_reference_checkpoint_loads_into_mirror_and_flat_adam(tmp_path):
    src = _model(0)
    ckpt, opt, names, params = _reference_style_checkpoint(src, 1234)
    path = os.path.join(tmp_path, ck.checkpoint_name(1234))
    assert path.endswith("step-000001234.ckpt")
    torch.save(ckpt, path)

    dst = _model(1)
    tr = UMHSTrainer(dst, occ_update=False)
    table_view = dst.field.mlp_base.encoder.hash_table.data_ptr()
    info = ck.load_checkpoint(str(tmp_path), dst, tr, endmembers_npy=os.path.join(tmp_path, "endmembers_hotdog.npy"))
    assert info["step"] == 1234 and tr.step == 1235 and tr.opt_steps == 3
    assert sorted(info["skipped"]) == ["module._model.lpips.net.scaling_layer.shift",
                                       "module.datamanager.train_camera_optimizer.pose_adjustment"]
    # parameters landed in place (the trainer's flat buffer still aliases them)
    assert dst.field.mlp_base.encoder.hash_table.data_ptr() == table_view
    for (k, a), (_, b) in zip(sorted(src.state_dict().items()), sorted(dst.state_dict().items())):
        assert torch.equal(a, b), k
    by_name = dict(dst.field.named_parameters())
    for i, n in enumerate(names):
        a, b = tr.flat.range_of(by_name[n])
        assert torch.equal(tr.flat.data[a:b].reshape(by_name[n].shape), by_name[n].data)
        st = opt.state_dict()["state"].get(i)
        if n.startswith("mlp_directional."):
            assert st is None and float(tr.flat.exp_avg[a:b].abs().sum()) == 0.0
        else:
            assert torch.equal(tr.flat.exp_avg[a:b], st["exp_avg"].reshape(-1)), n
            assert torch.equal(tr.flat.exp_avg_sq[a:b], st["exp_avg_sq"].reshape(-1)), n
    e = np.load(os.path.join(tmp_path, "endmembers_hotdog.npy"))  # ref umhs_pipeline.py:175
    assert e.shape == (4, 128) and np.array_equal(e, src.field.endmembers.detach().numpy())


def test_save_checkpoint_layout_round_trip(tmp_path):
    m = _model(2, specular=True)
    tr = UMHSTrainer(m, occ_update=False, start_step=77)
    tr.opt_steps = 5
    tr.flat.exp_avg.copy_(torch.randn_like(tr.flat.exp_avg))
    tr.flat.exp_avg_sq.copy_(torch.rand_like(tr.flat.exp_avg_sq))
    path = ck.save_checkpoint(str(tmp_path), m, tr)
    assert os.path.basename(path) == "step-000000077.ckpt"
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert set(raw) == {"step", "pipeline", "optimizers", "schedulers", "scalers"}
    assert all(k.startswith("_model.") for k in raw["pipeline"])
    for k in ("_model.field.mlp_base.encoder.hash_table", "_model.field.endmembers", "_model.occupancy_grid.binaries",
              "_model.field.converter.transform_matrix", "_model.converter.transform_matrix", "_model.scene_aabb"):
        assert k in raw["pipeline"], k
    # torch's own Adam accepts the optimiser entry for the reference's parameter order
    by_name = dict(m.field.named_parameters())
    opt = torch.optim.Adam([by_name[n] for n in ck.reference_param_order(m.field)], lr=2e-2, eps=1e-15)
    opt.load_state_dict(raw["optimizers"]["fields"])

    m2 = _model(3, specular=True)
    tr2 = UMHSTrainer(m2, occ_update=False)
    ck.load_checkpoint(path, m2, tr2)
    assert tr2.step == 78 and tr2.opt_steps == 5
    for n, p in m.field.named_parameters():  # per parameter: the flat buffers pad every segment to 4 entries
        (a, b), (a2, b2) = tr.flat.range_of(p), tr2.flat.range_of(dict(m2.field.named_parameters())[n])
        assert torch.equal(tr2.flat.data[a2:b2], tr.flat.data[a:b]), n
        assert torch.equal(tr2.flat.exp_avg[a2:b2], tr.flat.exp_avg[a:b]), n
        assert torch.equal(tr2.flat.exp_avg_sq[a2:b2], tr.flat.exp_avg_sq[a:b]), n
    assert torch.equal(m2.occupancy_grid.binaries, m.occupancy_grid.binaries)


def test_shape_mismatch_is_an_error(tmp_path):
    src = _model(0)
    ckpt, *_ = _reference_style_checkpoint(src, 1)
    ckpt["pipeline"]["module._model.field.endmembers"] = torch.zeros(6, 128)  # a K = 6 checkpoint into a K = 4 model
    with pytest.raises(ValueError, match="endmembers"):
        ck.load_pipeline(_model(1), ckpt["pipeline"], 1)
    del ckpt["pipeline"]["module._model.field.endmembers"]
    with pytest.raises(KeyError):
        ck.load_pipeline(_model(1), ckpt["pipeline"], 1)


def test_saved_checkpoint_passes_a_strict_reference_style_load(tmp_path):
    """what the REFERENCE does with a checkpoint -- nerfstudio's Trainer._load_checkpoint feeds
    optimizers / schedulers / scalers to torch's own load_state_dict, and UMHSPipeline.load_pipeline (ref
    umhs_pipeline.py:157-175) does a STRICT load_state_dict.  A stub pipeline with the reference's module tree (this
    model + the LPIPS buffers of NGPModel) must accept the file when the entries this package does not own are
    carried over from a reference checkpoint (``extras_from``), and must name exactly those entries otherwise."""
    m = _model(4)
    tr = UMHSTrainer(m, occ_update=False, start_step=10)
    tr.opt_steps = 2
    tr.flat.exp_avg.copy_(torch.randn_like(tr.flat.exp_avg))
    tr.flat.exp_avg_sq.copy_(torch.rand_like(tr.flat.exp_avg_sq))

    class _Lpips(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.register_buffer("shift", torch.zeros(1, 3, 1, 1))

    class _RefPipeline(torch.nn.Module):  # the reference's tree: pipeline._model = NGPModel subclass with .lpips
        def __init__(self, model):
            super().__init__()
            self._model = model
            self._model.lpips = _Lpips()

    # a checkpoint "written by the reference" (supplies _model.lpips.*)
    ref_like = _RefPipeline(_model(5))
    ref_path = os.path.join(tmp_path, "ref.ckpt")
    torch.save({"step": 1, "pipeline": {"module." + k: v for k, v in ref_like.state_dict().items()}}, ref_path)

    bare = torch.load(ck.save_checkpoint(str(tmp_path), m, tr), map_location="cpu", weights_only=False)
    target = _RefPipeline(_model(6))
    missing = target.load_state_dict(bare["pipeline"], strict=False).missing_keys
    assert missing == ["_model.lpips.shift"], missing  # only the module this package does not own
    assert "_model.sampler.occupancy_grid.binaries" in bare["pipeline"]

    full = torch.load(ck.save_checkpoint(str(tmp_path), m, tr, extras_from=ref_path), map_location="cpu",
                      weights_only=False)
    target.load_state_dict(full["pipeline"], strict=True)  # what ref umhs_pipeline.py:167 does
    assert torch.equal(target._model.field.endmembers, m.field.endmembers)
    # torch's own loaders accept the trainer-side entries
    by_name = dict(m.field.named_parameters())
    opt = torch.optim.Adam([by_name[n] for n in ck.reference_param_order(m.field)], lr=2e-2, eps=1e-15)
    opt.load_state_dict(full["optimizers"]["fields"])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda s: 1.0)
    sched.load_state_dict(full["schedulers"]["fields"])
    assert sched.last_epoch == 11
    scaler = torch.amp.GradScaler("cpu", enabled=True)
    scaler.load_state_dict(full["scalers"])  # an empty dict raises here


def test_tcnn_checkpoints_are_rejected_with_a_clear_message():
    src = _model(0)
    ckpt, *_ = _reference_style_checkpoint(src, 1)
    pipe = {k: v for k, v in ckpt["pipeline"].items() if "hash_table" not in k}
    pipe["module._model.field.mlp_base.encoder.tcnn_encoding.params"] = torch.zeros(1000)
    with pytest.raises(ValueError, match="tcnn"):
        ck.load_pipeline(_model(1), pipe, 1)


def test_sharded_adam_moments_are_gathered_for_saving():
    """Fused data-parallel step: rank r keeps Adam moments for shard r only.  ``full_optimizer_state`` (used by
    save_checkpoint) must return the concatenation; single-process form of the logic (the 2-rank gloo run is in
    tests/test_dp_cpu.py)."""
    m = _model(7)
    tr = UMHSTrainer(m, occ_update=False)
    n = tr.flat.numel
    assert tr.dp_shard() == (0, n)
    tr.world_size = 4
    spans = [tr.dp_shard(r) for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
    assert all(lo % 4 == 0 for lo, _ in spans)


def test_shard_ownership_with_a_split_exchange():
    """``dp_spans`` / ``dp_shards``: with the exchange cut in two slices (UMHS_DP_SPLIT_LEVEL; csrc/optim.cu::
    umhs_dp_adam_step_range) every rank owns one part of EACH slice; together the ranks' shards tile the flat buffer
    exactly once, on multiples of 4."""
    m = _model(3)
    tr = UMHSTrainer(m, occ_update=False)
    n = tr.flat.numel
    assert tr.dp_spans() == [(0, n)] and tr.dp_shards() == [(0, n)]
    tr.world_size = 8
    t_lo, t_hi = tr.flat.range_of(m.field.mlp_base.encoder.hash_table)
    tr._dp_cut = t_lo + (t_hi - t_lo) // 16 * 8  # first element of level 8
    assert tr._dp_cut % 4 == 0
    assert tr.dp_spans() == [(0, tr._dp_cut), (tr._dp_cut, n)]
    owned = torch.zeros(n, dtype=torch.int32)
    for r in range(8):
        shards = tr.dp_shards(r)
        assert len(shards) == 2
        for lo, hi in shards:
            assert lo % 4 == 0 and (hi % 4 == 0 or hi == n)
            owned[lo:hi] += 1
    assert int(owned.min()) == 1 and int(owned.max()) == 1

