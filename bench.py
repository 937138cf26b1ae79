#!/usr/bin/env python
"""bench.py -- UMHS hot-path benchmark (BASELINE.json metric: train rays/s at 128 bands on 1-8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload hotdog|cbox_dragon|
                    eval_specular] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (config.workload names the headline one; the others ride along under config.other_workloads):

  hotdog        BASELINE.json configs[1] (the default headline): full training step of the hotdog-shaped synthetic
                scene, 16384 rays per GPU per step, 128 bands, K = 4 endmembers (endmembers_hotdog[:, :128], frozen in
                tests/golden/colour.npz), temperature 0.4, hash grid 16x2 @ 2^19, occupancy grid 128^3 x 4 with a
                spherical-shell surface (radius 0.5, thickness 0.08 -> ~50 kept samples per ray).
  cbox_dragon   configs[2] shape (ref scripts/cbox_dragon.sh:3-9): K = 9, T = 0.3, specular head, 4096 rays per GPU
                (weak scaling) and, with more than one GPU, a fixed global batch of 32768 rays (strong scaling).
  eval_specular configs[3] shape: 800 x 600 evaluation render, K = 4, specular head, T in {0.4, 0.5, 0.7}
                (ref scripts/anacampseros.sh:10, pinecone.sh:10, caladium.sh:4), 32 x 32 tiles over the ranks.

One training step = occupancy-grid update (steady-state cadence) + march/pack + sampler density pass +
visibility/compaction + field forward + fused compositing/unmixing/sRGB + loss + full backward + gradient exchange
+ fused Adam, replayed as one CUDA graph.

`value`     : whole-job rays/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`       : same metric through the public API (UMHSTrainer.stage / train_step) with the batch copied from pinned
              host memory every step and the loss read back every step.
`roofline`  : dominant kernel, CUDA events on the launching stream in a second pass of the same K steps.
`cpu_baseline` / `--impl reference`: the CPU oracle (oracle/, the restatement of the reference's PyTorch path pinned by
              tests/golden) running the SAME training step on the host cores.
`dp_check`  : (N > 1, untimed preamble) the fused NVLink exchange + Adam kernel against NCCL all-reduce + the single-GPU
              Adam on the same shards, and against ONE GPU training on the concatenated batch.
"""
from __future__ import annotations

import argparse
import json
import os
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
CONDITIONING_STEPS = 34
STRONG_GLOBAL_RAYS = 32768
# UMHS_BENCH_PIPELINE=1: announce the next batch so that its occupancy march (parameter-independent) runs as a parallel
# branch of this step's graph (UMHSTrainer.train_step(next_rays=...)); same results
PIPELINE = bool(int(os.environ.get("UMHS_BENCH_PIPELINE", "1")))  # measured: 1.349 -> 1.306 ms per step
CPU_RAYS_PER_STEP = 4096   # the CPU arm's rays per step (2 warm-up + 5 timed, median)
WORKLOADS = {
    "hotdog": {"name": "hotdog-synthetic-train-step", "rays": RAYS_PER_GPU, "K": K_CLASSES, "T": TEMPERATURE,
               "specular": False, "ref": "BASELINE.json configs[1]"},
    "cbox_dragon": {"name": "cbox_dragon-synthetic-train-step", "rays": 4096, "K": 9, "T": 0.3, "specular": True,
                    "ref": "BASELINE.json configs[2]; ref scripts/cbox_dragon.sh:3-9"},
    "eval_specular": {"name": "eval-800x600-specular-render", "rays": 480000, "K": 4, "T": 0.4, "specular": True,
                      "temperatures": (0.4, 0.5, 0.7), "ref": "BASELINE.json configs[3]"},
}
# L2 (LTS) throughput cap of the guide's microarchitecture notes (B300_MICROARCH.md: ~6300 B/cycle full chip, measured
# at locked clocks; same L2 on B200): the roof of the hash-grid kernels, whose table is L2-resident
L2_BYTES_PER_CYCLE = 6300.0
# Per-launch traffic from the committed `ncu --set full` capture of the config-2 step (profiles/r2_ncu_full_summary.md):
# name -> (dram__bytes_read.sum + dram__bytes_write.sum, lts__t_bytes.sum); None = not captured
SAVE_ACTIVATIONS_ON = os.environ.get("UMHS_SAVE_ACTIVATIONS", "1") != "0"
NCU_TRAFFIC = {}
try:
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as _fh:
        NCU_TRAFFIC = json.load(_fh)
    # the capture is of the bf16-tile form of the hash kernels (what the config-2 step launches)
    for _k in ("umhs_hash_encode_fwd", "umhs_hash_encode_bwd"):
        if _k in NCU_TRAFFIC:
            NCU_TRAFFIC.setdefault(_k + "_tiles", NCU_TRAFFIC[_k])
except Exception:  # noqa: BLE001
    NCU_TRAFFIC = {}


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


def endmembers(K=K_CLASSES):
    z = np.load(os.path.join(ROOT, "tests", "golden", "colour.npz"))
    E = torch.from_numpy(z["endmembers"].astype(np.float32))  # endmembers_hotdog.npy[:, :128]
    if K <= E.shape[0]:
        return E[:K]
    # more classes than the shipped fixture holds (cbox_dragon: 9): smooth seeded spectra behind the four real ones
    g = torch.Generator().manual_seed(K)
    x = torch.linspace(0, 1, BANDS)
    extra = [0.5 + 0.4 * torch.sin(2 * np.pi * (float(torch.rand(1, generator=g)) * 2 + 0.5) * x
                                   + 6.28 * float(torch.rand(1, generator=g))) for _ in range(K - E.shape[0])]
    return torch.cat([E, torch.stack(extra)], 0)


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
        field.endmembers.copy_(endmembers(field.endmembers.shape[0]))
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
def oracle_trainer(seed=0, workload="hotdog"):
    from oracle import umhs_oracle as O

    wl = WORKLOADS[workload]
    cfg = O.OracleConfig(method="rgb+spectral", temperature=wl["T"], pred_specular=wl["specular"],
                         log2_hashmap_size=LOG2_T, grid_resolution=GRID_RES)
    m = O.OracleModel(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), WAVELENGTHS, wl["K"])
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


def time_oracle(steps, warmup, R, workload="hotdog", budget_s=None):
    """`steps` timed full training steps of R rays on all host cores (median and mean).  budget_s: shrink R (in
    multiples of 1024) when the first warm-up step says that the timed steps would not fit the budget."""
    torch.set_num_threads(os.cpu_count() or 1)
    m, opt = oracle_trainer(workload=workload)
    t0 = time.perf_counter()
    oracle_step(m, opt, R, 1000)
    first = time.perf_counter() - t0
    if budget_s is not None and first * (steps + max(warmup - 1, 0)) > budget_s:
        R = int(max(1024, (R * budget_s / (first * (steps + max(warmup - 1, 0)))) // 1024 * 1024))
    for i in range(1, max(warmup, 1)):
        oracle_step(m, opt, R, 1000 + i)
    ts = []
    for i in range(steps):
        t0 = time.perf_counter()
        oracle_step(m, opt, R, 2000 + i)
        ts.append(time.perf_counter() - t0)
    med, mean = float(np.median(ts)), float(np.mean(ts))
    return {"value": R / med, "value_mean": R / mean, "unit": "rays/s", "cores": torch.get_num_threads(),
            "kind": "port", "rays_per_step": R, "ms_per_step": med * 1e3,
            "sample": f"{steps} timed full training steps of {R} rays of the same synthetic scene after {max(warmup, 1)} "
                      f"warm-up (oracle/: fp32 torch on the host cores, C packing march, torch.optim.Adam); median"}


def time_oracle_config1():
    """BASELINE.md section 2 / SURVEY 8(d) "config 1": the reference's CPU-capable DENSE functions (4096 rays x 48
    samples x 128 bands, K = 4): field forward + dense weights + accumulate + sRGB + loss, then backward; 2 + 5, median."""
    from oracle import umhs_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    m, _ = oracle_trainer()
    R, S = 4096, 48
    o, d, image, hs = make_rays(R, 5)
    t0 = (0.3 + 0.0125 * torch.arange(S))[None, :].expand(R, S).contiguous()
    t1 = t0 + 0.0125
    ts = []
    for i in range(7):
        a = time.perf_counter()
        m.zero_grad(set_to_none=True)
        out = O.dense_forward(m, o, d, t0, t1)
        loss = 5 * torch.nn.functional.mse_loss(out["spectral"], hs) + torch.nn.functional.mse_loss(out["rgb"], image)
        loss.backward()
        if i >= 2:
            ts.append(time.perf_counter() - a)
    return {"rays_per_s": R / float(np.median(ts)), "ms": float(np.median(ts)) * 1e3,
            "shape": "4096 rays x 48 samples x 128 bands, dense layout, fwd + bwd, 2 warm-up + 5 timed, median"}


# ------------------------------------------------------------------------------------------ roofline bookkeeping
def algorithmic_bytes(name, units, samples_per_ray, K, adam_touched_frac=None):
    """Algorithmic bytes per launch (DESIGN.md 'Kernels'); units = what the wrapper passed (samples, rays, params)."""
    L = 16
    per = {
        "umhs_hash_encode_fwd": 8 * L * 8 + 16 + 8 * L,            # gathers + (ray_idx,t0,t1) + enc out (fp32 rows)
        "umhs_hash_encode_bwd": 8 * L + 16 + 2 * 8 * L * 8,        # d_enc in + points + table read-modify-write
        "umhs_hash_encode_fwd_tiles": 8 * L * 8 + 16 + 4 * L,      # ... with bf16 tile buffers in between
        "umhs_hash_encode_bwd_tiles": 4 * L + 16 + 2 * 8 * L * 8,
        "umhs_train_loss": None,
    }.get(name, None)
    S = samples_per_ray
    if name == "umhs_composite_fwd":
        return units * (S * (12 + 8 * K) + 16 + 4 * BANDS + 4 * K + 20)
    if name == "umhs_composite_bwd":
        return units * (2 * S * (12 + 8 * K) + 8 * BANDS)
    if name == "umhs_train_loss":
        return units * 3 * 4 * BANDS
    if name == "umhs_field_mlp_tc_fwd" and SAVE_ACTIVATIONS_ON:
        # training forward that keeps its activation tiles: enc tile row 64 + ray data 28 in; sigma, scalar, abund out;
        # 704 B of bf16 activation tiles out -- an HBM WRITE stream (3.9 TB/s is what a pure write reaches on this part)
        return units * (64 + 28 + 4 + 8 * K + 704)
    if name == "umhs_field_mlp_tc_bwd_saved":
        # activation tiles 704 + enc row 64 + forward outputs and their gradients 2 (4 + 8 K) in; d_enc row 64 out
        return units * (704 + 64 + 2 * (4 + 8 * K) + 28 + 64)
    if name in ("umhs_adam_step", "umhs_adam_step_dev"):
        # touched entries: p, g, m, v read + p, m, v, g written = 32 B; entries with g = m = v = 0 are skipped after
        # reading g, m, v = 12 B
        f = 1.0 if adam_touched_frac is None else adam_touched_frac
        return units * (32 * f + 12 * (1.0 - f))
    return None if per is None else units * per


def mlp_flops(name, units, K, specular):
    """(algorithmic, executed) FLOPs per launch of the tensor-core MLP kernels.  Algorithmic: forward =
    2 x MACs/sample; backward = dgrad + wgrad = 2 x forward.  Executed: the recomputing backward (umhs_field_mlp_tc_bwd)
    also runs the forward again in shared memory: 3 x forward; the backward from saved activations (the default,
    umhs_field_mlp_tc_bwd_saved) executes what it is credited with."""
    base = 2 * (32 * 64 + 64 * 16)
    heads = 2 * ((27 * 64 + 64 * 64 + 64 * K) + (27 * 64 + 64 * 64 + 64 * (K + (1 if specular else 0))))
    spec = 2 * (28 * 16) if specular else 0  # layer 0 of mlp_directional (layer 1 lives in the compositing kernels)
    fwd = base + heads + spec
    per = {"umhs_density_tc_fwd": (base, base), "umhs_field_mlp_tc_fwd": (fwd, fwd),
           "umhs_field_mlp_tc_bwd": (2 * fwd, 3 * fwd), "umhs_field_mlp_tc_bwd_saved": (2 * fwd, 2 * fwd)}.get(name)
    return (None, None) if per is None else (units * per[0], units * per[1])


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl_key = args.workload if args.workload in ("hotdog", "cbox_dragon") else "hotdog"
    wl = WORKLOADS[wl_key]
    steps = max(args.steps, 1)
    # Each step is a bounded sample of the workload's batch: 4096 of its rays (the reference's own batch size, ref
    # scripts/hotdog.sh:9 / cbox_dragon.sh:9, and BASELINE.md section 2's CPU protocol).  A full 16384-ray oracle step
    # holds the reference's [N,B,K] product (1.7 GB + autograd copies) and takes ~10-40 s of mostly page-fault time.
    # The count shrinks further (and says so) if K steps would not fit ~150 s of host time.
    r = time_oracle(steps, min(max(args.warmup, 1), 2), min(wl["rays"], CPU_RAYS_PER_STEP), workload=wl_key,
                    budget_s=150.0)
    R = r["rays_per_step"]
    value = r["value_mean"]  # K timed steps: total rays / total time, like the GPU arm
    line = {"impl": "reference", "metric": "train rays/s", "value": value, "unit": "rays/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": R / value * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {**workload_config(wl_key, args.gpus), "cpu_rays_per_step": R},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"].replace("; median", "; mean of the timed steps")},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(wl_key, n_gpus, samples_per_ray=None, rays=None):
    wl = WORKLOADS[wl_key]
    c = {"workload": wl["name"], "workload_ref": wl["ref"], "rays_per_gpu_per_step": rays or wl["rays"],
         "bands": BANDS, "endmembers": wl["K"], "temperature": wl["T"], "hash": "16x2@2^19",
         "occupancy": f"128^3x4 shell r={SHELL_R} t={SHELL_T}", "method": "rgb+spectral", "specular": wl["specular"],
         "parallelism": (f"dp{n_gpus} ray-sharded; gradient exchange + Adam fused in one kernel over NVLink peer "
                         f"memory (reduce-scatter reads, all-gather writes; NVLS multimem when the fabric offers a "
                         f"multicast address), 2 device barriers/step") if n_gpus > 1
         else "single",
         "conditioning_steps": CONDITIONING_STEPS,
         "launch": "whole step replayed as one CUDA graph (march+fwd+loss+bwd+exchange+Adam); occupancy update = "
                   "its own graph every 16 steps" + ("; the NEXT batch's occupancy march (parameter-independent) runs "
                                                      "as a parallel branch beside the hash scatter and the optimiser "
                                                      "(train_step(next_rays=...))" if PIPELINE else ""),
         "l2_policy": "working set > L2: Adam state 4 x 64 MiB streamed every step + rotating batches"}
    if samples_per_ray is not None:
        c["kept_samples_per_ray"] = samples_per_ray
    return c


# ------------------------------------------------------------------------------------------ B200 arm
def build_model(wl_key, dev, temperature=None):
    import umhs_b200

    wl = WORKLOADS[wl_key]
    cfg = umhs_b200.UMHSConfig(method="rgb+spectral", temperature=temperature or wl["T"], pred_specular=wl["specular"],
                               log2_hashmap_size=LOG2_T, grid_resolution=GRID_RES)
    model = cfg.setup(metadata={"wavelengths": WAVELENGTHS, "num_classes": wl["K"]})
    init_field_(model.field, 0)
    model.occupancy_grid.binaries.copy_(shell_binaries())
    model.occupancy_grid.occs.copy_(model.occupancy_grid.binaries.flatten().float() * 0.05)
    model.occupancy_grid.mark_dirty()
    return model.to(dev)


class Dist:
    """rank / world / barrier / max-over-ranks helpers (NCCL when world > 1)."""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # the contract is ONE JSON line on stdout: NCCL_DEBUG=VERSION (set in this image) makes NCCL print its
            # version banner there (WARN prints it too), so NCCL's log goes to stderr; INFO / TRACE asked for by a
            # user are left alone
            if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
                os.environ["NCCL_DEBUG"] = "WARN"
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms):
        t = torch.tensor([ms], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, finish=None):
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        b.record()
        self.barrier()
        return self.max_ms(a.elapsed_time(b))


class TrainRun:
    """One training workload on this rank: model + trainer + rotating batches, resident and pinned-host."""

    def __init__(self, D, wl_key, rays, graph=True, pool=8):
        from umhs_b200.trainer import UMHSTrainer

        self.D, self.wl_key, self.R, self.POOL = D, wl_key, rays, pool
        self.model = build_model(wl_key, D.dev)
        # steady-state occupancy update cadence (step >= warm-up 256), binaries frozen so the synthetic scene stays
        # fixed; late-training learning rate (step 29000 of the 30000-step exponential decay, lr ~ 1.3e-5) keeps the
        # synthetic workload stationary across timed steps: same Adam arithmetic, the density field barely moves
        self.trainer = UMHSTrainer(self.model, world_size=D.world, rank=D.rank, occ_update=True, freeze_binaries=True,
                                   start_step=29008, cuda_graph=graph)
        host = [make_rays(rays, 7919 * D.rank + i) for i in range(pool)]
        self.pinned = [tuple(t.pin_memory() for t in b) for b in host]
        self.resident = [tuple(t.to(D.dev) for t in b) for b in host]
        self.e2e_state = {"tokens": [], "prev_loss": None, "losses": []}

    def step_resident(self, i):
        from umhs_b200.rays import RayBundle

        o, d, image, hs = self.resident[i % self.POOL]
        nxt = RayBundle(*self.resident[(i + 1) % self.POOL][:2]) if PIPELINE else None
        return self.trainer.train_step(RayBundle(o, d), {"image": image, "hs_image": hs}, next_rays=nxt)

    def _stage(self, j):
        from umhs_b200.rays import RayBundle

        o, d, image, hs = self.pinned[j % self.POOL]
        return self.trainer.stage(RayBundle(o, d), {"image": image, "hs_image": hs})

    def step_e2e(self, i):
        """End-to-end step through the public API: every batch is copied from pinned host memory (batch i+2 while
        step i computes), every step's loss is read back to the host (the read of step i-1 is issued after step i is
        enqueued)."""
        st = self.e2e_state
        while len(st["tokens"]) < 2:
            st["tokens"].append(self._stage(i + len(st["tokens"])))
        out = self.trainer.train_step(st["tokens"][0], next_rays=st["tokens"][1] if PIPELINE else None)
        loss = self.trainer.read_async(out["loss"])  # device -> host copy of this step's result, on a side stream
        st["tokens"].pop(0)
        st["tokens"].append(self._stage(i + 2))
        if st["prev_loss"] is not None:
            st["losses"].append(st["prev_loss"].result())  # the host reads step i-1's loss while step i runs
        st["prev_loss"] = loss

    def finish_e2e(self):
        st = self.e2e_state
        if st["prev_loss"] is not None:
            st["losses"].append(st["prev_loss"].result())
            st["prev_loss"] = None
        st["tokens"].clear()

    def measure(self, steps, warmup, e2e=True):
        D = self.D
        for i in range(warmup + CONDITIONING_STEPS):
            self.step_resident(i)
        ms = D.timed(self.step_resident, steps)
        res = {"ms": ms, "rays_per_s": D.world * self.R * steps / (ms * 1e-3), "ms_per_step": ms / steps}
        if e2e:
            for i in range(max(3, warmup) + 16):
                self.step_e2e(i)
            self.finish_e2e()
            self.e2e_state["losses"].clear()
            ms_e2e = D.timed(self.step_e2e, steps, self.finish_e2e)
            assert len(self.e2e_state["losses"]) == steps and all(np.isfinite(self.e2e_state["losses"]))
            res.update({"ms_e2e": ms_e2e, "e2e_rays_per_s": D.world * self.R * steps / (ms_e2e * 1e-3)})
        # a truncated step (more candidates than the learnt capacity) does less work: there must be none
        res["truncated_steps"] = int(self.trainer.truncated_steps)
        assert res["truncated_steps"] == 0, "sample capacity overflowed inside the benchmark"
        return res

    def profile(self, steps):
        """second pass of the same steps with per-entry-point CUDA events (eager launches, so that every entry point
        can be bracketed; the timed passes replay the whole step as one CUDA graph)"""
        from umhs_b200 import ops

        graph = self.trainer.cuda_graph
        self.trainer.cuda_graph = False
        ops.Profiler.reset(enabled=True, timing=True)
        n_samples = 0
        self.D.barrier()
        for i in range(steps):
            n_samples += int(self.step_resident(i)["num_samples"].item())
        torch.cuda.synchronize()
        summ = ops.Profiler.summary()
        launches = ops.Profiler.launches
        ops.Profiler.reset(enabled=False)
        self.trainer.cuda_graph = graph
        return summ, launches, n_samples

    def replicas_identical(self):
        D = self.D
        if D.world == 1:
            return True
        mine = self.trainer.flat.data.clone()
        D.dist.broadcast(mine, src=0)
        same = torch.tensor([1.0 if torch.equal(mine, self.trainer.flat.data) else 0.0], device=D.dev)
        D.dist.all_reduce(same, op=D.dist.ReduceOp.MIN)
        return float(same.item()) == 1.0


def dp_check(D, R=4096, steps=2, lr=1e-3):
    """Untimed preamble at N > 1: the fused exchange + Adam kernel over NVLink peer memory (what the timed steps run)
    against (a) NCCL all-reduce + the single-GPU fused Adam on the same shards and (b) ONE GPU training on the
    concatenated batch (world_size = 1 trainer on every rank's rays), with the stratified jitter and the random
    background pinned so that all three see the same samples.  Adam with eps = 1e-15 turns a gradient that is zero up
    to summation order into a +-lr step, so parameters are compared through the entries that moved consistently and
    the first moments (linear in the gradients) through their maximum."""
    from umhs_b200.rays import RayBundle
    from umhs_b200.trainer import UMHSTrainer

    dist, dev, world, rank = D.dist, D.dev, D.world, D.rank
    g = torch.Generator().manual_seed(4242)
    all_batches = [[make_rays(R, 555 + 17 * s + r) for r in range(world)] for s in range(steps)]
    jit = [torch.rand(world, R, generator=g) for _ in range(steps)]
    bgs = [torch.rand(world, R, 3, generator=g) for _ in range(steps)]

    def run(world_size, fused):
        m = build_model("hotdog", dev)
        n = R if world_size > 1 else R * world
        m._jitter_override = torch.zeros(n, device=dev)
        m._bg_override = torch.zeros(n, 3, device=dev)
        tr = UMHSTrainer(m, lr=lr, lr_final=lr, world_size=world_size, rank=rank if world_size > 1 else 0,
                         occ_update=False, cuda_graph=False, fused_dp=fused)
        for s in range(steps):
            if world_size > 1:
                o, d, image, hs = all_batches[s][rank]
                m._jitter_override.copy_(jit[s][rank])
                m._bg_override.copy_(bgs[s][rank])
            else:
                o, d, image, hs = (torch.cat([all_batches[s][r][k] for r in range(world)]) for k in range(4))
                m._jitter_override.copy_(jit[s].reshape(-1))
                m._bg_override.copy_(bgs[s].reshape(-1, 3))
            out = tr.train_step(RayBundle(o.to(dev), d.to(dev)), {"image": image.to(dev), "hs_image": hs.to(dev)})
        torch.cuda.synchronize()
        return tr, float(out["loss"])

    ta, loss_a = run(world, True)     # fused NVLink exchange + Adam
    tb, loss_b = run(world, False)    # NCCL all-reduce + single-GPU Adam
    tc, loss_c = run(1, False)        # one GPU, concatenated batch
    n = ta.flat.numel
    own = torch.zeros(n, dtype=torch.bool, device=ta.flat.data.device)
    for lo, hi in ta.dp_shards():     # this rank's shards of the optimiser state in the fused path
        own[lo:hi] = True
    pa, pb, pc = ta.flat.data, tb.flat.data, tc.flat.data
    d_ab, d_ac = (pa - pb).abs(), (pa - pc).abs()
    m_a, m_b, m_c = ta.flat.exp_avg[own], tb.flat.exp_avg[own], tc.flat.exp_avg[own]
    scale = float(m_c.abs().max()) + 1e-30
    res = {"steps": steps, "rays_per_rank": R, "lr": lr,
           "fused_vs_allreduce": {"max_abs_diff": float(d_ab.max()), "mean_abs_diff": float(d_ab.mean()),
                                  "entries_off_by_more_than_5pct_lr": int((d_ab > 0.05 * lr * steps).sum()),
                                  "exp_avg_max_rel": float((m_a - m_b).abs().max()) / scale},
           "fused_vs_single_gpu_concat": {"max_abs_diff": float(d_ac.max()), "mean_abs_diff": float(d_ac.mean()),
                                          "entries_off_by_more_than_5pct_lr": int((d_ac > 0.05 * lr * steps).sum()),
                                          "exp_avg_max_rel": float((m_a - m_c).abs().max()) / scale},
           "losses": {"fused": loss_a, "allreduce": loss_b, "single_gpu_concat_last_step": loss_c},
           "params": n}
    ok = (res["fused_vs_allreduce"]["exp_avg_max_rel"] < 1e-3 and res["fused_vs_single_gpu_concat"]["exp_avg_max_rel"] < 5e-3
          and res["fused_vs_allreduce"]["entries_off_by_more_than_5pct_lr"] <= 1e-4 * n
          and res["fused_vs_single_gpu_concat"]["entries_off_by_more_than_5pct_lr"] <= 1e-3 * n
          and res["fused_vs_allreduce"]["mean_abs_diff"] < 1e-3 * lr)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["ok"] = bool(flag.item() == 1.0)
    del ta, tb, tc
    torch.cuda.empty_cache()
    return res


def time_next_train(dev, R, iters=20):
    """The step in front of the hot path: UMHSDataManager.next_train on device-resident images --
    pixel draws + RGB / 128-band gather + camera rays as one kernel.  Reported beside the train step, not inside it
    (the train step's e2e number copies its batches from pinned host memory, as the contract asks)."""
    from umhs_b200 import UMHSDataManager, ops

    n_img, H, W = 24, 400, 400  # 1.97 GB of fp32 cubes: far beyond L2, random 512-byte rows
    g = torch.Generator(device=dev).manual_seed(11)
    hs = torch.rand((n_img, H, W, BANDS), device=dev, generator=g)
    img = torch.rand((n_img, H, W, 3), device=dev, generator=g)
    c2w = torch.cat([torch.linalg.qr(torch.randn(n_img, 3, 3))[0], torch.randn(n_img, 3, 1)], -1)
    dm = UMHSDataManager(img, hs, c2w, 600.0, 600.0, W / 2.0, H / 2.0, train_num_rays_per_batch=R, device=dev)

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
            "dataset": f"{n_img} x {H}x{W}x{BANDS} fp32 cubes + RGB resident in HBM"}


def time_render(D, trainer, n_images=8, seed=4242):
    """Evaluation rendering (800 x 600 image, 32 x 32 tiles dealt round-robin to the ranks, no collective, eval mode:
    no visibility filter, no jitter): Msamples/s = packed samples evaluated per second."""
    from umhs_b200.rays import RayBundle

    H_EVAL, W_EVAL = 600, 800
    dev = D.dev
    eo, ed, _, _ = make_rays(H_EVAL * W_EVAL, seed)
    cam = RayBundle(eo.view(H_EVAL, W_EVAL, 3).to(dev), ed.view(H_EVAL, W_EVAL, 3).to(dev))
    res = None
    # the sampler thread starts (NVML initialisation: ~15 ms of host time) before the warm-up, not inside the timed region
    clocks = ClockSampler(D.local, period_s=0.005)
    if D.rank == 0:
        clocks.start()
    for _ in range(3):  # warm-up with the timed loop's own allocation pattern (previous image released first)
        res = None
        res = trainer.render_tiles(cam)
    D.barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(n_images + 1)]
    marks[0].record()
    n_eval = 0
    for it in range(n_images):
        res = None  # release the previous image first
        res = trainer.render_tiles(cam)
        n_eval = res["num_samples_per_ray"].sum()
        marks[it + 1].record()
    D.barrier()
    clk = clocks.stop() if D.rank == 0 else None
    per_image = [marks[i].elapsed_time(marks[i + 1]) for i in range(n_images)]
    # median over the timed images (max over ranks): on one GPU the first timed image of the loop is regularly ~15 ms
    # late (host side: the allocator re-growing its multi-GB blocks after the warm-up images were released); the mean
    # over all timed images is printed beside it
    ms = D.max_ms(float(np.median(per_image)))
    ms_mean = D.max_ms(marks[0].elapsed_time(marks[-1]) / n_images)
    samples = n_eval.to(torch.float32).reshape(1).clone()
    if D.world > 1:
        D.dist.all_reduce(samples, op=D.dist.ReduceOp.SUM)
    trainer.model.train()
    s = float(samples.item())
    return {"msamples_per_s": s / (ms * 1e-3) / 1e6, "image": f"{W_EVAL}x{H_EVAL}", "ms_per_image": ms,
            "samples_per_image": s, "rays_per_s": H_EVAL * W_EVAL / (ms * 1e-3),
            "sharding": "32x32 tiles round-robin over ranks, no collective; one fused pass per rank",
            "images_timed": n_images, "ms_per_image_mean": ms_mean, "ms_per_image_each": [round(x, 3) for x in per_image],
            "clocks": {k: clk[k] for k in ("sm_mhz", "reasons", "samples")} if clk else None}


def roofline_line(summ, steps, spr, wl, peaks, sm_mhz, adam_touched_frac):
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "6650 GB/s (fallback)"
    # kernels timed inside a long step: the sustained tensor figure when the driver measured one
    tf_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1500.0)))
    tf_src = ("MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else
              "MEASURED_PEAKS.json bf16_tflops" if "bf16_tflops" in peaks else "1500 TFLOP/s dense bf16 (fallback)")
    l2_peak = L2_BYTES_PER_CYCLE * (sm_mhz or 1965.0) * 1e6 / 1e9
    K, spec = wl["K"], wl["specular"]
    kernels = {}
    for name, d in summ.items():
        per_launch_ms = d["ms"] / d["calls"]
        units = d["units"] / d["calls"]
        ab = algorithmic_bytes(name, units, spr, K, adam_touched_frac)
        af, ef = mlp_flops(name, units, K, spec)
        kernels[name] = {"calls_per_step": d["calls"] / steps, "ms_per_step": d["ms"] / steps,
                         "ms_per_launch": per_launch_ms,
                         "achieved_gbs": None if ab is None else ab / (per_launch_ms * 1e-3) / 1e9,
                         "achieved_tflops": None if af is None else af / (per_launch_ms * 1e-3) / 1e12,
                         "executed_tflops": None if ef is None else ef / (per_launch_ms * 1e-3) / 1e12}
    total_k = sum(v["ms_per_step"] for v in kernels.values())
    top = max(kernels, key=lambda n: kernels[n]["ms_per_step"])
    tr = NCU_TRAFFIC.get(top) or {}
    share = {"share_of_kernel_time": kernels[top]["ms_per_step"] / max(total_k, 1e-9),
             "kernel_ms_per_step_sum": total_k, "traffic": tr.get("dram_bytes")}
    if kernels[top]["achieved_tflops"] is not None:  # the fused MLP: the only dense contraction on the path
        roof = {"kernel": top, "bound": "tensor", "achieved": kernels[top]["achieved_tflops"], "peak": tf_peak,
                "unit": "TFLOP/s", "frac": kernels[top]["achieved_tflops"] / tf_peak, "peak_source": tf_src,
                "algorithmic_flops": "backward = dgrad + wgrad = 2 x forward MACs x 2 (SURVEY 8d); forward = 2 x MACs",
                "executed_tflops": kernels[top]["executed_tflops"],
                "executed_note": "umhs_field_mlp_tc_bwd recomputes the forward in shared memory (3 x forward FLOPs "
                                 "executed); umhs_field_mlp_tc_bwd_saved starts from saved activation tiles (2 x)",
                **share}
    elif top.startswith("umhs_hash_encode"):
        roof = {"kernel": top, "bound": "l2", "achieved": kernels[top]["achieved_gbs"], "peak": l2_peak, "unit": "GB/s",
                "frac": (kernels[top]["achieved_gbs"] or 0.0) / l2_peak,
                "peak_source": f"L2 LTS cap {L2_BYTES_PER_CYCLE:.0f} B/cycle (B300_MICROARCH.md) x SM clock", **share}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": hbm_peak,
                "unit": "GB/s", "frac": (kernels[top]["achieved_gbs"] or 0.0) / hbm_peak, "peak_source": peak_src,
                **share}
    # the other kernels the north star asks about, each against the roof that really bounds it
    roof["tensor_kernels"] = {n: {"achieved": v["achieved_tflops"], "executed": v["executed_tflops"], "peak": tf_peak,
                                  "unit": "TFLOP/s", "frac": v["achieved_tflops"] / tf_peak,
                                  "traffic": (NCU_TRAFFIC.get(n) or {}).get("dram_bytes")}
                              for n, v in kernels.items() if v["achieved_tflops"] is not None}
    roof["hbm_kernels"] = {n: {"achieved": v["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                               "frac": v["achieved_gbs"] / hbm_peak, "traffic": (NCU_TRAFFIC.get(n) or {}).get("dram_bytes")}
                           for n, v in kernels.items()
                           if v["achieved_gbs"] is not None and not n.startswith("umhs_hash_encode")}
    roof["l2_kernels"] = {n: {"achieved": v["achieved_gbs"], "peak": l2_peak, "unit": "GB/s",
                              "frac": v["achieved_gbs"] / l2_peak,
                              "traffic_l2": (NCU_TRAFFIC.get(n) or {}).get("lts_bytes"),
                              "traffic_dram": (NCU_TRAFFIC.get(n) or {}).get("dram_bytes"),
                              "note": "the 64 MiB table is L2-resident (126 MB L2): gathers / reductions are served by "
                                      "L2; algorithmic bytes count 8 B per gathered entry, the L2 moves 32-byte sectors"}
                          for n, v in kernels.items() if v["achieved_gbs"] is not None and n.startswith("umhs_hash_encode")}
    return roof, kernels


def run_b200(args):
    from umhs_b200 import ops  # noqa: F401
    from umhs_b200.trainer import UMHSTrainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    D = Dist()
    world, rank, dev = D.world, D.rank, D.dev
    head_key = args.workload
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))

    dpc = dp_check(D) if world > 1 and not args.no_dp_check else None

    # ---------------------------------------------------------------- eval-only headline
    if head_key == "eval_specular":
        line = {"metric": "render Msamples/s", "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16 tcgen05 MLP operands, f32 accumulate; f32 elsewhere", "data": "synthetic"}
        per_T = {}
        for T in WORKLOADS["eval_specular"]["temperatures"]:
            m = build_model("eval_specular", dev, temperature=T)
            tr = UMHSTrainer(m, world_size=1, rank=0, occ_update=False)
            tr.world_size, tr.rank = world, rank  # tile sharding only; no gradient exchange is ever set up
            per_T[str(T)] = time_render(D, tr, n_images=max(2, min(args.steps, 8)))
            del m, tr
        first = per_T[str(WORKLOADS["eval_specular"]["temperatures"][0])]
        line.update({"value": first["msamples_per_s"], "ms_per_step": first["ms_per_image"],
                     "config": {**workload_config("eval_specular", world), "per_temperature": per_T},
                     "e2e": {"value": first["msamples_per_s"], "unit": "Msamples/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0, "note": "camera rays resident; outputs stay on the device"},
                     "gpu_launches": None, "dp_check": dpc})
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            D.dist.barrier()
            D.dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- training headline
    wl = WORKLOADS[head_key]
    R = wl["rays"] if args.scaling == "weak" else max(1024, STRONG_GLOBAL_RAYS // world)
    run = TrainRun(D, head_key, R, graph=not args.no_graph)
    clocks = ClockSampler(D.local)
    if rank == 0:
        clocks.start()
    res = run.measure(args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None
    summ, launches, n_samples = run.profile(args.steps)
    spr = n_samples / (args.steps * R)
    f = run.trainer.flat
    touched = float(((f.exp_avg != 0) | (f.exp_avg_sq != 0)).float().mean())
    render = time_render(D, run.trainer)
    identical = run.replicas_identical()
    h2d = sum(t.numel() * t.element_size() for t in run.pinned[0])
    exchange = ("nvls multimem" if f.mc_grad else "peer loads/stores") if world > 1 else None

    # ---------------------------------------------------------------- the other workloads ride along
    others = {}
    if head_key == "hotdog" and not args.headline_only:
        del run.resident, run.pinned
        cb = TrainRun(D, "cbox_dragon", WORKLOADS["cbox_dragon"]["rays"], graph=not args.no_graph, pool=4)
        r_w = cb.measure(args.steps, args.warmup, e2e=False)
        others["cbox_dragon"] = {"workload": WORKLOADS["cbox_dragon"]["name"], "endmembers": 9, "temperature": 0.3,
                                 "specular": True, "metric": "train rays/s",
                                 "weak": {"rays_per_gpu_per_step": cb.R, "rays_per_s": r_w["rays_per_s"],
                                          "ms_per_step": r_w["ms_per_step"], "truncated_steps": r_w["truncated_steps"]},
                                 "replicas_identical": cb.replicas_identical()}
        del cb
        torch.cuda.empty_cache()
        if world > 1:
            cs = TrainRun(D, "cbox_dragon", STRONG_GLOBAL_RAYS // world, graph=not args.no_graph, pool=4)
            r_s = cs.measure(args.steps, args.warmup, e2e=False)
            others["cbox_dragon"]["strong"] = {"global_rays_per_step": STRONG_GLOBAL_RAYS,
                                               "rays_per_gpu_per_step": cs.R, "rays_per_s": r_s["rays_per_s"],
                                               "ms_per_step": r_s["ms_per_step"]}
            del cs
            torch.cuda.empty_cache()
        ev = {}
        for T in WORKLOADS["eval_specular"]["temperatures"]:
            m = build_model("eval_specular", dev, temperature=T)
            tr = UMHSTrainer(m, world_size=1, rank=0, occ_update=False)
            tr.world_size, tr.rank = world, rank
            r = time_render(D, tr, n_images=3)
            ev[str(T)] = {k: r[k] for k in ("msamples_per_s", "ms_per_image", "samples_per_image", "rays_per_s")}
            del m, tr
            torch.cuda.empty_cache()
        others["eval_specular"] = {"workload": WORKLOADS["eval_specular"]["name"], "endmembers": 4, "specular": True,
                                   "metric": "render Msamples/s", "image": "800x600", "per_temperature": ev}
    next_train = time_next_train(dev, R) if rank == 0 else None

    if rank == 0:
        roof, kernels = roofline_line(summ, args.steps, spr, wl, peaks, (clk or {}).get("sm_mhz"), touched)
        cpu = None
        if world == 1 and not args.no_cpu:
            # BASELINE.md section 2 protocol: 4096 rays per step, 2 warm-up + 5 timed, median, all host cores
            c = time_oracle(steps=5, warmup=2, R=CPU_RAYS_PER_STEP, workload=head_key)
            cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu["ms_per_step"] = c["ms_per_step"]
            if head_key == "hotdog":
                cpu["config1_dense"] = time_oracle_config1()
        line = {"metric": "train rays/s", "value": res["rays_per_s"], "unit": "rays/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "bf16 tcgen05 MLP operands, f32 accumulate; f32 elsewhere", "data": "synthetic",
                "config": {**workload_config(head_key, world, samples_per_ray=spr, rays=R),
                           **({"dp_exchange": exchange} if world > 1 else {}),
                           "other_workloads": others},
                "e2e": {"value": res["e2e_rays_per_s"], "unit": "rays/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": res["ms_e2e"] / args.steps,
                        "pipeline": "pinned-host batches copied two steps ahead on a copy stream (3 staging slots) "
                                    "while earlier steps run; loss of every step read back (one step late); all "
                                    "inside the timed region"},
                "gpu_launches": launches, "clocks": clk, "roofline": roof,
                "cpu_baseline": cpu,
                "truncated_steps": res["truncated_steps"], "replicas_identical": identical, "dp_check": dpc,
                "adam_touched_fraction": touched,
                "train_msamples_per_s": world * n_samples / args.steps / (res["ms_per_step"] * 1e-3) / 1e6,
                "render": render,
                "other_workloads": others,
                "next_train": ({**next_train, "frac": next_train["achieved_gbs"] / float(peaks.get("hbm_gbs", 6650.0)),
                                "peak": float(peaks.get("hbm_gbs", 6650.0))} if next_train else None),
                "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        assert identical, "data-parallel replicas diverged"
        assert dpc is None or dpc["ok"], f"dp_check failed: {dpc}"
        D.dist.barrier()
        D.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)  # three occupancy-update periods
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hotdog", choices=list(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the whole-step CUDA graph")
    ap.add_argument("--headline-only", action="store_true", help="skip the ride-along workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-dp-check", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
