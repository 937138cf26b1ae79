"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path.

The reference never synchronises gradients (world_size forced to 1: ref umhsnerf/umhs_pipeline.py:86,108-113), so the
contract is stated against the single-process result: all-reduce(sum) of per-shard gradients / world == gradient of
the concatenated batch (losses are means over the local shard).  The compute on each rank is the CPU oracle (the
checker; kernels need a GPU) -- what is under test is FlatParams aliasing, the exchange step and tile sharding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import WAVELENGTHS, load_golden, oracle_model_from_golden, t


def _shard_grads(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import nerf3p
        from umhs_b200.trainer import FlatParams, average_gradients

        torch.set_num_threads(2)
        d, meta = load_golden("model_train_k4.npz")
        m = oracle_model_from_golden(d, meta, training=True)
        flat = FlatParams([p for p in m.field.parameters() if p.requires_grad])
        R = d["origins"].shape[0]
        sl = slice(rank * R // world, (rank + 1) * R // world)
        rb = nerf3p.RayBundle(t(d["origins"][sl]), t(d["directions"][sl]))
        out = m.get_outputs(rb, jitter=t(d["jitter"][sl]))
        loss = sum(m.get_loss_dict(out, {"image": t(d["image"][sl]), "hs_image": t(d["hs_image"][sl])},
                                   bg=t(d["bg"][sl])).values())
        loss.backward()  # accumulates into the flat gradient through the parameter views
        assert flat.grad.abs().sum() > 0
        average_gradients(flat.grad, world)
        flat.grad.div_(world)
        if rank == 0:
            np.save(os.path.join(out_dir, "dp_grad.npy"), flat.grad.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp_gradient_equals_full_batch(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_shard_grads, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    dp = np.load(tmp_path / "dp_grad.npy")
    # single-process gradient of the concatenated batch
    from oracle import nerf3p
    from umhs_b200.trainer import FlatParams

    d, meta = load_golden("model_train_k4.npz")
    m = oracle_model_from_golden(d, meta, training=True)
    flat = FlatParams([p for p in m.field.parameters() if p.requires_grad])
    out = m.get_outputs(nerf3p.RayBundle(t(d["origins"]), t(d["directions"])), jitter=t(d["jitter"]))
    sum(m.get_loss_dict(out, {"image": t(d["image"]), "hs_image": t(d["hs_image"])}, bg=t(d["bg"])).values()).backward()
    ref = flat.grad.numpy()
    np.testing.assert_allclose(dp, ref, rtol=2e-4, atol=2e-5 * np.abs(ref).max())
    # and the flat buffer really is the parameters' storage
    p0 = flat.params[0]
    flat.data[flat.offsets[0]] += 1.0
    assert float(p0.data.flatten()[0]) == float(flat.data[flat.offsets[0]])


def test_tile_sharding_partitions_the_image():
    from umhs_b200.trainer import tile_pixel_indices

    for H, W, tile, world in ((600, 800, 32, 8), (37, 53, 16, 3), (5, 5, 32, 2)):
        parts = [tile_pixel_indices(H, W, tile, r, world) for r in range(world)]
        allidx = torch.cat(parts)
        assert allidx.numel() == H * W
        assert torch.equal(torch.sort(allidx).values, torch.arange(H * W))
        if H * W > tile * tile * world:
            assert max(p.numel() for p in parts) - min(p.numel() for p in parts) <= 2 * tile * tile


def _sharded_moments(rank, world, port, out_dir):
    """Each rank owns the Adam moments of one shard (what the fused NVLink exchange + Adam kernel leaves behind):
    save_checkpoint on every rank must gather them, and the file written by rank 0 must hold ALL of them."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import umhs_b200
        from umhs_b200 import checkpoint as ck
        from umhs_b200.trainer import UMHSTrainer

        torch.manual_seed(0)
        cfg = umhs_b200.UMHSConfig(log2_hashmap_size=8, grid_resolution=16)
        m = cfg.setup(metadata={"wavelengths": list(range(400, 528)), "num_classes": 4})
        tr = UMHSTrainer(m, occ_update=False, world_size=world, rank=rank, start_step=5)
        tr.opt_steps = 3
        g = torch.Generator().manual_seed(1)
        full_m, full_v = torch.randn(tr.flat.numel, generator=g), torch.rand(tr.flat.numel, generator=g)
        lo, hi = tr.dp_shard()
        tr.flat.exp_avg.zero_()
        tr.flat.exp_avg_sq.zero_()
        tr.flat.exp_avg[lo:hi] = full_m[lo:hi]      # this rank's shard only, as the fused kernel keeps it
        tr.flat.exp_avg_sq[lo:hi] = full_v[lo:hi]
        tr.fused_dp = True                         # (the CUDA-only fused mode; here only its bookkeeping)
        path = ck.save_checkpoint(out_dir, m, tr, write=(rank == 0))
        dist.barrier()
        if rank == 1:  # a fresh single-process trainer resumes from the file rank 0 wrote
            m2 = cfg.setup(metadata={"wavelengths": list(range(400, 528)), "num_classes": 4})
            tr2 = UMHSTrainer(m2, occ_update=False)
            ck.load_checkpoint(path, m2, tr2)
            by2 = dict(m2.field.named_parameters())
            for n, p in m.field.named_parameters():
                a, b = tr.flat.range_of(p)
                a2, b2 = tr2.flat.range_of(by2[n])
                assert torch.equal(tr2.flat.exp_avg[a2:b2], full_m[a:b]), n
                assert torch.equal(tr2.flat.exp_avg_sq[a2:b2], full_v[a:b]), n
            assert tr2.opt_steps == 3
            open(os.path.join(out_dir, "ok"), "w").write("1")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_checkpoint_gathers_sharded_adam_state(tmp_path):
    world, port = 2, 31500 + (os.getpid() % 2000)
    mp.spawn(_sharded_moments, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "ok")
