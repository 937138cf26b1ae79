#!/usr/bin/env python
"""bench.py -- UMHS hot-path benchmark (BASELINE.json metric: train rays/s at 128 bands on 1-8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload = "hotdog-synthetic-train-step"): BASELINE.json configs[1] -- full training step of the
hotdog-shaped synthetic scene, 16384 rays per GPU per step, 128 bands, K=4 endmembers (endmembers_hotdog[:, :128],
frozen in tests/golden/colour.npz), temperature 0.4, hash grid 16x2 @ 2^19, occupancy grid 128^3 x 4 with a
spherical-shell surface (radius 0.5, thickness 0.08 -> ~50 kept samples per ray).  One step = occupancy-grid update
(steady-state cadence) + march/pack + sampler density pass + visibility/compaction + field forward + fused
compositing/unmixing/sRGB + loss + full backward + [NCCL all-reduce of the flat gradient] + fused Adam.

`value`     : whole-job rays/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`       : same metric through the public API (UMHSTrainer.train_step) with the batch copied from pinned host
              memory every step and the loss read back every step.
`roofline`  : dominant kernel, CUDA events on the launching stream in a second timed pass of the same K steps.
`cpu_baseline` / `--impl reference`: the CPU oracle (oracle/, the restatement of the reference's PyTorch path
              pinned by tests/golden) running the SAME training step on a bounded ray sample on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-hyperspectral-nerf_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

RAYS_PER_GPU = 16384
BANDS, K_CLASSES, TEMPERATURE, LOG2_T, GRID_RES = 128, 4, 0.4, 19, 128
SHELL_R, SHELL_T = 0.5, 0.08
WAVELENGTHS = np.linspace(400, 1000, BANDS).astype(int).tolist()
CPU_SAMPLE_RAYS = 1024
CONDITIONING_STEPS = 34
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
# (profiles/r1_ncu_full_summary.md, N = 825k samples, R = 16384 rays); None = not captured
NCU_TRAFFIC = {"umhs_hash_encode_fwd": 120.8e6, "umhs_hash_encode_bwd": 150.6e6, "umhs_composite_fwd": 36.8e6,
               "umhs_composite_bwd": 60.6e6, "umhs_adam_step": 283.1e6, "umhs_adam_step_dev": 283.1e6,
               "umhs_field_mlp_tc_bwd": 225.1e6, "umhs_field_mlp_tc_fwd": 142.3e6, "umhs_density_tc_fwd": 126.4e6,
               "umhs_train_loss": 17.5e6}


# ------------------------------------------------------------------------------------------ synthetic scene
def make_rays(R, seed):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(R, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * (0.8 + 0.2 * torch.rand(R, 1, generator=g))
    d = -o / o.norm(dim=-1, keepdim=True) + 0.25 * torch.randn(R, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    image = torch.rand(R, 3, generator=g)
    hs = torch.rand(R, BANDS, generator=g)
    return o.contiguous(), d.contiguous(), image, hs


def shell_binaries():
    c = (torch.arange(GRID_RES) + 0.5) / GRID_RES * 2 - 1
    X, Y, Z = torch.meshgrid(c, c, c, indexing="ij")
    rr = (X ** 2 + Y ** 2 + Z ** 2).sqrt()
    b = torch.zeros(4, GRID_RES, GRID_RES, GRID_RES, dtype=torch.bool)
    b[0] = (rr > SHELL_R - SHELL_T / 2) & (rr < SHELL_R + SHELL_T / 2)
    return b


def endmembers():
    z = np.load(os.path.join(ROOT, "tests", "golden", "colour.npz"))
    return torch.from_numpy(z["endmembers"].astype(np.float32))  # endmembers_hotdog.npy[:, :128]


def init_field_(field, seed=0):
    """Deterministic random-init weights of the reference architecture (nn.Linear default init, hash table
    U(-1e-3, 1e-3)), hotdog endmembers, density bias so that sigma ~ e^3 (partially terminating rays)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in sorted(field.named_parameters()):  # same draw order for the product and the oracle
            if p.dim() == 2 and p.shape[1] == 2 and p.shape[0] > 4096:  # hash table
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 1e-3)
            else:
                bound = 1.0 / max(p.shape[-1], 1) ** 0.5 if p.dim() == 2 else 0.1
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
        field.endmembers.copy_(endmembers())
        field.mlp_base.mlp.layers[1].bias[0] = 3.0


# ------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions.  In-process NVML (nvidia_ml_py): a
    `nvidia-smi -lms` subprocess holds driver locks long enough to stall kernel launches of a step that syncs every
    iteration (measured: +1.3 ms/step on the end-to-end pass), NVML queries cost tens of microseconds."""

    def __init__(self, index, period_s=0.02):
        self.index, self.period, self.rows, self._stop, self._thr, self._nvml = index, period_s, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self._max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None
            return
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def _loop(self):
        nv = self._nvml
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.rows.append((sm, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        if self._nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self._stop = True
        self._thr.join(timeout=1.0)
        nv = self._nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted({n for _, rs in self.rows for n, bit in names.items() if rs & bit})
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self._max), "reasons": reasons,
                "samples": len(sm), "source": "nvml, 20 ms period, sampled over the two timed passes"}


# ------------------------------------------------------------------------------------------ CPU oracle arm
def oracle_trainer(seed=0):
    from oracle import umhs_oracle as O

    cfg = O.OracleConfig(method="rgb+spectral", temperature=TEMPERATURE, pred_specular=False, log2_hashmap_size=LOG2_T,
                         grid_resolution=GRID_RES)
    m = O.OracleModel(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), WAVELENGTHS, K_CLASSES)
    init_field_(m.field, seed)
    m.occupancy_grid.binaries.copy_(shell_binaries())
    m.occupancy_grid.occs.copy_(m.occupancy_grid.binaries.flatten().float() * 0.05)
    m.train()
    opt = torch.optim.Adam(m.field.parameters(), lr=2e-2, eps=1e-15)
    return m, opt


def oracle_step(m, opt, R, seed):
    from oracle import nerf3p

    o, d, image, hs = make_rays(R, seed)
    opt.zero_grad(set_to_none=True)
    out = m.get_outputs(nerf3p.RayBundle(o, d))
    loss = sum(m.get_loss_dict(out, {"image": image, "hs_image": hs}).values())
    loss.backward()
    opt.step()
    m.clamp_endmembers()
    return float(loss.detach())


def time_oracle(steps, warmup, R=CPU_SAMPLE_RAYS):
    torch.set_num_threads(os.cpu_count() or 1)
    m, opt = oracle_trainer()
    for i in range(warmup):
        oracle_step(m, opt, R, 1000 + i)
    t0 = time.perf_counter()
    for i in range(steps):
        oracle_step(m, opt, R, 2000 + i)
    dt = time.perf_counter() - t0
    return {"value": R * steps / dt, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} full training steps of {R} rays of the same synthetic scene (oracle/ fp32 torch CPU, "
                      f"C packing march, torch.optim.Adam), {warmup} warm-up", "ms_per_step": dt / steps * 1e3}


# ------------------------------------------------------------------------------------------ roofline bookkeeping
def algorithmic_bytes(name, units, samples_per_ray):
    """Algorithmic bytes per launch (DESIGN.md 'Kernels'); units = what the wrapper passed (samples, rays, params)."""
    L = 16
    per = {
        "umhs_hash_encode_fwd": 8 * L * 8 + 16 + 8 * L,            # gathers + (ray_idx,t0,t1) + enc out
        "umhs_hash_encode_bwd": 8 * L + 16 + 2 * 8 * L * 8,        # d_enc in + points + table read-modify-write
        "umhs_density_fwd": 8 * L + 16 + 4,                        # enc in + points + sigma out
        "umhs_field_mlp_fwd": 8 * L + 16 + 4 + 60 + 8 * K_CLASSES,  # enc + points + sigma + emb + scalar/abund
        "umhs_field_mlp_bwd": 8 * L + 16 + 4 + 8 * K_CLASSES + 8 * L,  # enc + points + grads in + d_enc out
        "umhs_composite_fwd": None,
        "umhs_composite_bwd": None,
        "umhs_adam_step": 32,                                      # p,g,m,v read; p,m,v,g written
        "umhs_adam_step_dev": 32,
    }.get(name, None)
    if name == "umhs_composite_fwd":
        S = samples_per_ray
        return units * (S * (12 + 8 * K_CLASSES) + 16 + 4 * BANDS + 4 * K_CLASSES + 20 + 4 * S)  # + weights out
    if name == "umhs_composite_bwd":
        S = samples_per_ray
        return units * (2 * S * (12 + 8 * K_CLASSES) + 8 * BANDS)
    return None if per is None else units * per


def algorithmic_flops(name, units):
    """Algorithmic FLOPs per launch of the tensor-core MLP kernels (2 x MACs/sample, K = 4, no specular;
    the backward recomputes the forward and adds one dgrad + one wgrad GEMM per layer: 3 x forward)."""
    base = 2 * (32 * 64 + 64 * 16)
    heads = 2 * 2 * (27 * 64 + 64 * 64 + 64 * K_CLASSES)
    per = {"umhs_density_tc_fwd": base, "umhs_field_mlp_tc_fwd": base + heads,
           "umhs_field_mlp_tc_bwd": 3 * (base + heads)}.get(name)
    return None if per is None else units * per


# ------------------------------------------------------------------------------------------ main arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: ~8k rays of CPU work in total whatever K is (a 1024-ray oracle step takes seconds)
    R = int(min(1024, max(128, 8192 // max(args.steps, 1))))
    r = time_oracle(max(args.steps, 1), max(args.warmup, 1), R)
    line = {"impl": "reference", "metric": "train rays/s", "value": r["value"], "unit": "rays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {**workload_config(args.gpus, cpu=True), "rays_per_gpu_per_step": R},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, cpu=False, samples_per_ray=None):
    c = {"workload": "hotdog-synthetic-train-step", "rays_per_gpu_per_step": RAYS_PER_GPU if not cpu else CPU_SAMPLE_RAYS,
         "bands": BANDS, "endmembers": K_CLASSES, "temperature": TEMPERATURE, "hash": "16x2@2^19",
         "occupancy": f"128^3x4 shell r={SHELL_R} t={SHELL_T}", "method": "rgb+spectral", "specular": False,
         "parallelism": (f"dp{n_gpus} ray-sharded; gradient exchange + Adam fused in one kernel over NVLink peer "
                         f"memory (reduce-scatter reads, all-gather writes; NVLS multimem when the fabric offers a "
                         f"multicast address), 2 device barriers/step") if n_gpus > 1
         else "single",
         "conditioning_steps": CONDITIONING_STEPS,
         "launch": "whole step replayed as one CUDA graph (march+fwd+loss+bwd+exchange+Adam); occupancy update = "
                   "its own graph every 16 steps",
         "l2_policy": "working set > L2: Adam state 4 x 64 MiB streamed every step + rotating batches"}
    if samples_per_ray is not None:
        c["kept_samples_per_ray"] = samples_per_ray
    return c


def time_next_train(dev, R, iters=20):
    """The step in front of the hot path: UMHSDataManager.next_train on device-resident images --
    pixel draws + RGB / 128-band gather + camera rays as one kernel.  Reported beside the train step, not inside it
    (the train step's e2e number copies its batches from pinned host memory, as the contract asks)."""
    from umhs_b200 import UMHSDataManager

    n_img, H, W = 24, 400, 400  # 1.97 GB of fp32 cubes: far beyond L2, random 512-byte rows
    g = torch.Generator(device=dev).manual_seed(11)
    hs = torch.rand((n_img, H, W, BANDS), device=dev, generator=g)
    img = torch.rand((n_img, H, W, 3), device=dev, generator=g)
    c2w = torch.cat([torch.linalg.qr(torch.randn(n_img, 3, 3))[0], torch.randn(n_img, 3, 1)], -1)
    dm = UMHSDataManager(img, hs, c2w, 600.0, 600.0, W / 2.0, H / 2.0, train_num_rays_per_batch=R, device=dev)
    from umhs_b200 import ops

    def timed_calls(rays):
        dm.train_num_rays_per_batch = rays
        for _ in range(3):
            dm.next_train(0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            dm.next_train(0)
        b.record()
        torch.cuda.synchronize()
        ops.Profiler.reset(enabled=True, timing=True)
        for _ in range(iters):
            dm.next_train(0)
        torch.cuda.synchronize()
        k = ops.Profiler.summary()["umhs_sample_batch"]
        ops.Profiler.reset(enabled=False)
        return a.elapsed_time(b) / iters, k["ms"] / k["calls"]

    per_ray = 12 + 24 + 32 + 2 * 4 * BANDS + 2 * 4 * 3  # draws, indices, ray, spectrum in + out, rgb in + out
    call_ms, kernel_ms = timed_calls(R)
    _, big_ms = timed_calls(16 * R)
    del dm, hs, img
    return {"rays": R, "ms_per_call": call_ms, "kernel_ms": kernel_ms, "algorithmic_bytes_per_ray": per_ray,
            "achieved_gbs": 16 * R * per_ray / (big_ms * 1e-3) / 1e9, "unit": "GB/s",
            "achieved_at": f"{16 * R} rays per launch ({big_ms * 1e3:.1f} us); {R} rays: "
                           f"{R * per_ray / (kernel_ms * 1e-3) / 1e9:.0f} GB/s (latency-sized launch)",
            "ms_per_call_includes": "torch.rand draws, output allocation, umhs_sample_batch (host-bound at this size)",
            "dataset": f"{n_img} x {H}x{W}x{BANDS} fp32 cubes + RGB resident in HBM"}


def run_b200(args):
    import torch.distributed as dist

    import umhs_b200
    from umhs_b200 import ops
    from umhs_b200.rays import RayBundle
    from umhs_b200.trainer import UMHSTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the contract is ONE JSON line on stdout: NCCL_DEBUG=VERSION (set in this image) makes NCCL print its version
        # there; INFO / TRACE asked for by a user are left alone
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    cfg = umhs_b200.UMHSConfig(method="rgb+spectral", temperature=TEMPERATURE, pred_specular=False,
                               log2_hashmap_size=LOG2_T, grid_resolution=GRID_RES)
    model = cfg.setup(metadata={"wavelengths": WAVELENGTHS, "num_classes": K_CLASSES})
    init_field_(model.field, 0)
    model.occupancy_grid.binaries.copy_(shell_binaries())
    model.occupancy_grid.occs.copy_(model.occupancy_grid.binaries.flatten().float() * 0.05)
    model.occupancy_grid.mark_dirty()
    model = model.to(dev)
    # steady-state occupancy update cadence (step >= warm-up 256), binaries frozen so the synthetic scene stays fixed;
    # late-training learning rate (step 29000 of the 30000-step exponential decay, lr ~ 1.3e-5) keeps the synthetic
    # workload stationary across timed steps: same Adam arithmetic, the density field barely moves
    trainer = UMHSTrainer(model, world_size=world, rank=rank, occ_update=True, freeze_binaries=True, start_step=29008,
                          cuda_graph=not args.no_graph)

    R, POOL = RAYS_PER_GPU, 8
    host = [make_rays(R, 7919 * rank + i) for i in range(POOL)]
    pinned = [tuple(t.pin_memory() for t in b) for b in host]
    resident = [tuple(t.to(dev) for t in b) for b in host]

    # UMHS_BENCH_PIPELINE=1: announce the next batch so that its occupancy march runs beside this step (experimental,
    # measured gain 1-2 %: see trainer._capture); default: every step marches its own batch first
    PIPE = bool(os.environ.get("UMHS_BENCH_PIPELINE"))

    def step_resident(i):
        o, d, image, hs = resident[i % POOL]
        nxt = RayBundle(*resident[(i + 1) % POOL][:2]) if PIPE else None
        return trainer.train_step(RayBundle(o, d), {"image": image, "hs_image": hs}, next_rays=nxt)

    e2e_state = {"tokens": [], "prev_loss": None, "losses": []}

    def stage_e2e(j):
        o, d, image, hs = pinned[j % POOL]
        return trainer.stage(RayBundle(o, d), {"image": image, "hs_image": hs})

    def step_e2e(i):
        """End-to-end step through the public API: every batch is copied from pinned host memory (batch i+2 while
        step i computes; batch i+1 is already on the device and its occupancy march runs beside step i), every
        step's loss is read back to the host (the read of step i-1 is issued after step i is enqueued)."""
        st = e2e_state
        while len(st["tokens"]) < 2:
            st["tokens"].append(stage_e2e(i + len(st["tokens"])))
        out = trainer.train_step(st["tokens"][0], next_rays=st["tokens"][1] if PIPE else None)
        loss = trainer.read_async(out["loss"])  # device -> host copy of this step's result, on a side stream
        st["tokens"].pop(0)
        st["tokens"].append(stage_e2e(i + 2))
        if st["prev_loss"] is not None:
            st["losses"].append(st["prev_loss"].result())  # the host reads step i-1's loss while step i runs
        st["prev_loss"] = loss

    def finish_e2e():
        st = e2e_state
        if st["prev_loss"] is not None:
            st["losses"].append(st["prev_loss"].result())
            st["prev_loss"] = None
        st["tokens"].clear()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        stamps = [time.perf_counter()]
        for i in range(steps):
            fn(i)
            stamps.append(time.perf_counter())
        if os.environ.get("UMHS_BENCH_DEBUG"):
            print("cpu ms per iteration:", [round(1e3 * (y - x), 2) for x, y in zip(stamps[:-1], stamps[1:])],
                  file=sys.stderr)
        if finish is not None:
            finish()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # conditioning (untimed, on top of the W warm-up steps): graph capture, capacity learning and the first two
    # occupancy-update periods allocate their buffers here, not inside the timed region
    for i in range(args.warmup + CONDITIONING_STEPS):
        step_resident(i)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(step_resident, args.steps)
    # warm the end-to-end path too (staging buffers, copy stream), through one whole occupancy-update period
    for i in range(max(3, args.warmup) + 16):
        step_e2e(i)
    finish_e2e()
    e2e_state["losses"].clear()
    ms_e2e = timed(step_e2e, args.steps, finish_e2e)
    assert len(e2e_state["losses"]) == args.steps and all(np.isfinite(e2e_state["losses"]))
    clk = clocks.stop() if rank == 0 else None

    # second pass of the same K steps with per-entry-point CUDA events (roofline shares); eager launches so that
    # every entry point can be bracketed (the timed passes above replay the whole step as one CUDA graph)
    trainer.cuda_graph = False
    ops.Profiler.reset(enabled=True, timing=True)
    n_samples = 0
    barrier()
    for i in range(args.steps):
        n_samples += int(step_resident(i)["num_samples"].item())
    torch.cuda.synchronize()
    summ = ops.Profiler.summary()
    # kernels of this repo launched per timed pass of K steps (the graph replays the same launches)
    launches = ops.Profiler.launches
    ops.Profiler.reset(enabled=False)
    spr = n_samples / (args.steps * R)

    # evaluation rendering (BASELINE config 4 shape: 800 x 600 image, 32 x 32 tiles dealt round-robin to the ranks, no
    # collective, eval mode: no visibility filter, no jitter): Msamples/s = packed samples evaluated per second
    H_EVAL, W_EVAL = 600, 800
    eo, ed, _, _ = make_rays(H_EVAL * W_EVAL, 4242)
    cam = RayBundle(eo.view(H_EVAL, W_EVAL, 3).to(dev), ed.view(H_EVAL, W_EVAL, 3).to(dev))
    res = None
    for _ in range(3):  # warm-up with the timed loop's own allocation pattern (previous image released first)
        res = None
        res = trainer.render_tiles(cam)
    barrier()
    N_EVAL_IMAGES = 8
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(N_EVAL_IMAGES + 1)]
    eval_clocks = ClockSampler(local, period_s=0.005)
    if rank == 0:
        eval_clocks.start()
    mem0 = torch.cuda.memory_stats(dev)
    marks[0].record()
    n_eval = 0
    for it in range(N_EVAL_IMAGES):
        res = None  # release the previous image first
        res = trainer.render_tiles(cam)
        n_eval = res["num_samples_per_ray"].sum()
        marks[it + 1].record()
    barrier()
    mem1 = torch.cuda.memory_stats(dev)
    eval_clk = eval_clocks.stop() if rank == 0 else None
    per_image = [marks[i].elapsed_time(marks[i + 1]) for i in range(N_EVAL_IMAGES)]
    # the mean of all images is the reported figure; the list and the allocator counters explain outliers (a
    # cudaMalloc / cudaFree of the multi-GB transients inside the region shows up here)
    eval_alloc = {"cuda_malloc": int(mem1.get("segment.all.allocated", 0) - mem0.get("segment.all.allocated", 0)),
                  "cuda_free": int(mem1.get("segment.all.freed", 0) - mem0.get("segment.all.freed", 0)),
                  "alloc_retries": int(mem1.get("num_alloc_retries", 0) - mem0.get("num_alloc_retries", 0))}
    ea, eb = marks[0], marks[-1]
    eval_ms = torch.tensor([ea.elapsed_time(eb) / N_EVAL_IMAGES], device=dev)
    eval_samples = n_eval.to(torch.float32).reshape(1).clone()
    if world > 1:
        dist.all_reduce(eval_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(eval_samples, op=dist.ReduceOp.SUM)
    model.train()
    next_train = time_next_train(dev, R) if rank == 0 else None

    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
        # kernels timed inside a long step: the sustained tensor figure when the driver measured one
        tf_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1500.0)))
        tf_src = ("MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else
                  "MEASURED_PEAKS.json bf16_tflops" if "bf16_tflops" in peaks else "1500 TFLOP/s dense bf16 (fallback)")
        kernels = {}
        for name, d in summ.items():
            per_launch_ms = d["ms"] / d["calls"]
            ab = algorithmic_bytes(name, d["units"] / d["calls"], spr)
            af = algorithmic_flops(name, d["units"] / d["calls"])
            kernels[name] = {"calls_per_step": d["calls"] / args.steps, "ms_per_step": d["ms"] / args.steps,
                             "ms_per_launch": per_launch_ms,
                             "achieved_gbs": None if ab is None else ab / (per_launch_ms * 1e-3) / 1e9,
                             "achieved_tflops": None if af is None else af / (per_launch_ms * 1e-3) / 1e12}
        total_k = sum(v["ms_per_step"] for v in kernels.values())
        top = max(kernels, key=lambda n: kernels[n]["ms_per_step"])
        share = {"share_of_kernel_time": kernels[top]["ms_per_step"] / max(total_k, 1e-9),
                 "kernel_ms_per_step_sum": total_k, "traffic": NCU_TRAFFIC.get(top)}
        if kernels[top]["achieved_tflops"] is not None:  # the fused MLP: the only dense contraction on the path
            roof = {"kernel": top, "bound": "tensor", "achieved": kernels[top]["achieved_tflops"], "peak": tf_peak,
                    "unit": "TFLOP/s", "frac": kernels[top]["achieved_tflops"] / tf_peak, "peak_source": tf_src, **share}
        else:
            roof = {"kernel": top, "bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": (kernels[top]["achieved_gbs"] or 0.0) / hbm_peak, "peak_source": peak_src,
                    **share}
        # the HBM-bound kernels the north star asks about (gather, scatter, compositing, optimiser), same definition
        notes = {"umhs_hash_encode_fwd": "gathers are served by L2 (64 MiB table of a 126 MB L2): algorithmic bytes >> DRAM "
                                         "traffic; the bound is L2 sector throughput + gather latency",
                 "umhs_hash_encode_bwd": "reductions resolve in L2: algorithmic read-modify-write bytes >> DRAM traffic; "
                                         "the bound is the L2 reduction rate",
                 "umhs_adam_step_dev": "entries with g = m = v = 0 are skipped (12 of 32 B): algorithmic bytes count "
                                       "32 B for every parameter, hence frac > 1",
                 "umhs_adam_step": "see umhs_adam_step_dev"}
        roof["hbm_kernels"] = {n: {"achieved": v["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                                   "frac": v["achieved_gbs"] / hbm_peak, "traffic": NCU_TRAFFIC.get(n),
                                   **({"note": notes[n]} if n in notes else {})}
                               for n, v in kernels.items() if v["achieved_gbs"] is not None}
        cpu = time_oracle(steps=4, warmup=1) if world == 1 else None
        h2d = sum(t.numel() * t.element_size() for t in pinned[0])
        line = {"metric": "train rays/s", "value": world * R * args.steps / (ms * 1e-3), "unit": "rays/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {**workload_config(world, samples_per_ray=spr),
                           **({"dp_exchange": "nvls multimem" if trainer.flat.mc_grad else "peer loads/stores"}
                              if world > 1 else {})},
                "e2e": {"value": world * R * args.steps / (ms_e2e * 1e-3), "unit": "rays/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                        "pipeline": "pinned-host batches copied two steps ahead on a copy stream (3 staging slots) "
                                    "while earlier steps run; loss of every step read back (one step late); all "
                                    "inside the timed region"},
                "gpu_launches": launches, "clocks": clk, "roofline": roof,
                "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None,
                "train_msamples_per_s": world * n_samples / args.steps / (ms / args.steps * 1e-3) / 1e6,
                "render": {"msamples_per_s": float(eval_samples.item()) / (float(eval_ms.item()) * 1e-3) / 1e6,
                           "image": f"{W_EVAL}x{H_EVAL}", "ms_per_image": float(eval_ms.item()),
                           "samples_per_image": float(eval_samples.item()),
                           "rays_per_s": H_EVAL * W_EVAL / (float(eval_ms.item()) * 1e-3),
                           "sharding": "32x32 tiles round-robin over ranks, no collective; one fused pass per rank",
                           "images_timed": N_EVAL_IMAGES, "ms_per_image_each": [round(x, 3) for x in per_image],
                           "allocator_in_region": eval_alloc,
                           "clocks": {k: eval_clk[k] for k in ("sm_mhz", "reasons", "samples")} if eval_clk else None},
                "next_train": {**next_train, "frac": next_train["achieved_gbs"] / hbm_peak, "peak": hbm_peak},
                "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        # data-parallel invariant: after any number of steps every rank holds bit-identical parameters
        mine = trainer.flat.data.clone()
        dist.broadcast(mine, src=0)
        same = torch.tensor([1.0 if torch.equal(mine, trainer.flat.data) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        assert float(same.item()) == 1.0, "data-parallel replicas diverged"
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)  # three occupancy-update periods
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the whole-step CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
