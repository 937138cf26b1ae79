"""Where does the end-to-end step time go?  (H2D alone, graph step alone, staged pipeline variants)"""
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
tr = UMHSTrainer(model, occ_update=True, freeze_binaries=True, start_step=29008, cuda_graph=True)
host = [bench.make_rays(bench.RAYS_PER_GPU, i) for i in range(4)]
pinned = [tuple(t.pin_memory() for t in b) for b in host]
res = [tuple(t.to(dev) for t in b) for b in host]


def sync():
    torch.cuda.synchronize()


def timeit(name, fn, n=32):
    sync()
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    sync()
    print(f"{name:50s} {1e3 * (time.perf_counter() - t0) / n:.3f} ms/iter")


for i in range(6):
    o, d, im, hs = res[i % 4]
    tr.train_step(RayBundle(o, d), {"image": im, "hs_image": hs})
dst = [torch.empty_like(t) for t in res[0]]


def h2d(i):
    for a, b in zip(dst, pinned[i % 4]):
        a.copy_(b, non_blocking=True)


def step_res(i):
    o, d, im, hs = res[i % 4]
    return tr.train_step(RayBundle(o, d), {"image": im, "hs_image": hs})


def step_pinned(i):
    o, d, im, hs = pinned[i % 4]
    return tr.train_step(RayBundle(o, d), {"image": im, "hs_image": hs})


def step_pinned_item(i):
    return float(step_pinned(i)["loss"].item())


tok = [None]


def step_staged(i, item=False):
    if tok[0] is None:
        o, d, im, hs = pinned[i % 4]
        tok[0] = tr.stage(RayBundle(o, d), {"image": im, "hs_image": hs})
    out = tr.train_step(tok[0])
    o, d, im, hs = pinned[(i + 1) % 4]
    tok[0] = tr.stage(RayBundle(o, d), {"image": im, "hs_image": hs})
    if item:
        return float(out["loss"].item())


for rep in range(2):
    print("--- repetition", rep)
    timeit("H2D of one batch (4 tensors, pinned)", h2d)
    timeit("graph step, resident inputs", step_res)
    timeit("graph step, pinned inputs (H2D on main stream)", step_pinned)
    timeit("  + loss.item() every step", step_pinned_item)
    timeit("staged (copy stream) no item", step_staged)
    timeit("staged + item() of the same step", lambda i: step_staged(i, True))
tr.occ_update = False
for rep in range(2):
    timeit("graph step resident, no occupancy update", step_res)
    timeit("staged, no occupancy update", step_staged)
    timeit("staged + item, no occupancy update", lambda i: step_staged(i, True))

# ---- the bench's own e2e closure (deferred loss read) ----
tr.occ_update = True
state = {"token": None, "prev": None}


def bench_e2e(i):
    if state["token"] is None:
        o, d, im, hs = pinned[i % 4]
        state["token"] = tr.stage(RayBundle(o, d), {"image": im, "hs_image": hs})
    out = tr.train_step(state["token"])
    loss = out["loss"].clone()
    o, d, im, hs = pinned[(i + 1) % 4]
    state["token"] = tr.stage(RayBundle(o, d), {"image": im, "hs_image": hs})
    if state["prev"] is not None:
        float(state["prev"].item())
    state["prev"] = loss


for rep in range(2):
    timeit("bench-style e2e (deferred item, clone)", bench_e2e)
    timeit("staged + item same step (occ update on)", lambda i: step_staged(i, True))
