"""Training step at other BASELINE configurations than the bench's headline one (config 3: cbox_dragon-shaped, K = 9,
T = 0.3, specular head on, 4096 rays per GPU -- ref scripts/cbox_dragon.sh:3-9; config 4 shape: K = 4, specular,
800x600 eval).  Prints ms/step (whole-step CUDA graph) and the per-kernel CUDA-event times of an eager pass.
  python tools/config_probe.py [K] [rays] [specular 0/1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-hyperspectral-nerf_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
import umhs_b200  # noqa: E402
from umhs_b200 import ops  # noqa: E402
from umhs_b200.rays import RayBundle  # noqa: E402
from umhs_b200.trainer import UMHSTrainer  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 9
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
SPEC = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
dev = torch.device("cuda", 0)
cfg = umhs_b200.UMHSConfig(method="rgb+spectral", temperature=0.3, pred_specular=SPEC, log2_hashmap_size=bench.LOG2_T,
                           grid_resolution=bench.GRID_RES)
model = cfg.setup(metadata={"wavelengths": bench.WAVELENGTHS, "num_classes": K})
g = torch.Generator().manual_seed(0)
with torch.no_grad():
    for _, p in sorted(model.field.named_parameters()):
        if p.dim() == 2 and p.shape[1] == 2 and p.shape[0] > 4096:
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 1e-3)
        else:
            bound = 1.0 / max(p.shape[-1], 1) ** 0.5 if p.dim() == 2 else 0.1
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
    model.field.endmembers.copy_(torch.rand(model.field.endmembers.shape, generator=g))
    model.field.mlp_base.mlp.layers[1].bias[0] = 3.0
model.occupancy_grid.binaries.copy_(bench.shell_binaries())
model.occupancy_grid.occs.copy_(model.occupancy_grid.binaries.flatten().float() * 0.05)
model.occupancy_grid.mark_dirty()
model = model.to(dev)
trainer = UMHSTrainer(model, occ_update=True, freeze_binaries=True, start_step=29008, cuda_graph=True)
batches = [tuple(t.to(dev) for t in bench.make_rays(R, 100 + i)) for i in range(4)]


def step(i):
    o, d, image, hs = batches[i % 4]
    return trainer.train_step(RayBundle(o, d), {"image": image, "hs_image": hs})


for i in range(40):
    step(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
N = 32
for i in range(N):
    out = step(i)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / N
trainer.cuda_graph = False
ops.Profiler.reset(enabled=True, timing=True)
ns = 0
for i in range(16):
    ns += int(step(i)["num_samples"].item())
torch.cuda.synchronize()
summ = ops.Profiler.summary()
ops.Profiler.reset(enabled=False)
print(f"K={K} rays={R} specular={SPEC}: {ms:.3f} ms/step, {R / ms / 1e3:.2f} M rays/s, {ns / 16 / R:.1f} samples/ray")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:28s} {v['ms'] / 16:.4f} ms/step ({v['calls'] / 16:.2f} calls)")
