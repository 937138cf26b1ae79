"""Which kernels does one training step launch (the bench's step, eager), by name and count: finds the torch-side
fill / copy / elementwise launches that ride along with this repo's kernels inside the step graph."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-hyperspectral-nerf_b200"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import umhs_b200  # noqa: E402
from umhs_b200.rays import RayBundle  # noqa: E402
from umhs_b200.trainer import UMHSTrainer  # noqa: E402

dev = torch.device("cuda", 0)
cfg = umhs_b200.UMHSConfig(method="rgb+spectral", temperature=bench.TEMPERATURE, pred_specular=False,
                           log2_hashmap_size=bench.LOG2_T, grid_resolution=bench.GRID_RES)
model = cfg.setup(metadata={"wavelengths": bench.WAVELENGTHS, "num_classes": bench.K_CLASSES})
bench.init_field_(model.field, 0)
model.occupancy_grid.binaries.copy_(bench.shell_binaries())
model.occupancy_grid.occs.copy_(model.occupancy_grid.binaries.flatten().float() * 0.05)
model.occupancy_grid.mark_dirty()
model = model.to(dev)
trainer = UMHSTrainer(model, occ_update=False, start_step=29008, cuda_graph=False)
o, d, image, hs = (t.to(dev) for t in bench.make_rays(bench.RAYS_PER_GPU, 1))
for _ in range(5):
    trainer.train_step(RayBundle(o, d), {"image": image, "hs_image": hs})
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    trainer.train_step(RayBundle(o, d), {"image": image, "hs_image": hs})
    torch.cuda.synchronize()
cnt, dur = collections.Counter(), collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        cnt[ev.name[:90]] += 1
        dur[ev.name[:90]] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
ours = sum(c for n, c in cnt.items() if "umhs::" in n)
print(f"kernels + memcpy/memset in one step: {sum(cnt.values())} (umhs: {ours})")
for n, c in sorted(cnt.items(), key=lambda kv: -dur[kv[0]]):
    print(f"  {c:3d} x {dur[n]:9.1f} us  {n}")
