"""Host-side overhead of a training step: CPU enqueue time per step (no sync) and a cProfile of the Python path."""
import cProfile
import pstats
import sys
import time

sys.path.insert(0, ".")
import bench  # noqa: E402
import torch  # noqa: E402

import umhs_b200  # noqa: E402
from umhs_b200.rays import RayBundle  # noqa: E402
from umhs_b200.trainer import UMHSTrainer  # noqa: E402

dev = torch.device("cuda", 0)
cfg = umhs_b200.UMHSConfig(method="rgb+spectral", temperature=bench.TEMPERATURE, log2_hashmap_size=bench.LOG2_T,
                           grid_resolution=bench.GRID_RES)
model = cfg.setup(metadata={"wavelengths": bench.WAVELENGTHS, "num_classes": bench.K_CLASSES})
bench.init_field_(model.field, 0)
model.occupancy_grid.binaries.copy_(bench.shell_binaries())
model.occupancy_grid.occs.copy_(model.occupancy_grid.binaries.flatten().float() * 0.05)
model.occupancy_grid.mark_dirty()
model = model.to(dev)
tr = UMHSTrainer(model, occ_update=True, freeze_binaries=True, start_step=29008)
batches = [tuple(t.to(dev) for t in bench.make_rays(bench.RAYS_PER_GPU, i)) for i in range(4)]


def step(i):
    o, d, im, hs = batches[i % 4]
    return tr.train_step(RayBundle(o, d), {"image": im, "hs_image": hs})


for i in range(5):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(32):
    step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"CPU enqueue {1e3 * (t1 - t0) / 32:.3f} ms/step ; wall incl. drain {1e3 * (t2 - t0) / 32:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(32):
    step(i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
