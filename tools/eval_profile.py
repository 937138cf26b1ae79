"""Where an 800x600 evaluation render spends its time: wall clock per image against the sum of the kernels' own
times (per-entry-point CUDA events), on the bench's synthetic scene.  `python tools/eval_profile.py [tile_rays]`."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-hyperspectral-nerf_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
import umhs_b200  # noqa: E402
from umhs_b200 import ops  # noqa: E402
from umhs_b200.rays import RayBundle  # noqa: E402
from umhs_b200.trainer import UMHSTrainer  # noqa: E402

dev = torch.device("cuda", 0)
cfg = umhs_b200.UMHSConfig(method="rgb+spectral", temperature=bench.TEMPERATURE, pred_specular=False,
                           log2_hashmap_size=bench.LOG2_T, grid_resolution=bench.GRID_RES)
if len(sys.argv) > 1:
    cfg.eval_tile_rays = int(sys.argv[1])
model = cfg.setup(metadata={"wavelengths": bench.WAVELENGTHS, "num_classes": bench.K_CLASSES})
bench.init_field_(model.field, 0)
model.occupancy_grid.binaries.copy_(bench.shell_binaries())
model.occupancy_grid.occs.copy_(model.occupancy_grid.binaries.flatten().float() * 0.05)
model.occupancy_grid.mark_dirty()
model = model.to(dev)
trainer = UMHSTrainer(model, occ_update=False)
H, W = 600, 800
eo, ed, _, _ = bench.make_rays(H * W, 4242)
cam = RayBundle(eo.view(H, W, 3).to(dev), ed.view(H, W, 3).to(dev))
for _ in range(2):
    res = trainer.render_tiles(cam)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    res = trainer.render_tiles(cam)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 3
n = int(res["num_samples_per_ray"].sum())
ops.Profiler.reset(enabled=True, timing=True)
res = trainer.render_tiles(cam)
torch.cuda.synchronize()
summ = ops.Profiler.summary()
ops.Profiler.reset(enabled=False)
print(f"tile_rays {model.config.eval_tile_rays}  wall {wall * 1e3:.2f} ms/image  samples {n}  "
      f"{n / wall / 1e6:.0f} Msamples/s")
tot = 0.0
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    tot += v["ms"]
    print(f"  {k:28s} calls {v['calls']:4d}  {v['ms']:.3f} ms")
print(f"  kernels sum {tot:.3f} ms")
