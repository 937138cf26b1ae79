import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "unsupervised-hyperspectral-nerf_b200")
import bench
from umhs_b200 import ops
from umhs_b200.ops import PointSet
dev = torch.device("cuda")
m = bench.build_model("hotdog", dev)
f = m.field
res, n_cells = 128, 128 ** 3
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for lvl in range(4):
    aabb = m.occupancy_grid.aabbs[lvl]
    idx = torch.randint(n_cells, (1 << 18,), device=dev)
    for name, ids in (("random", idx), ("sorted", torch.sort(idx).values), ("bucket>>10", idx[torch.sort(idx >> 10, stable=True).indices]), ("bucket>>14", idx[torch.sort(idx >> 14, stable=True).indices])):
        c = torch.stack([ids // (res * res), (ids // res) % res, ids % res], -1).float()
        x = (c + torch.rand_like(c)) / res
        pos = (aabb[:3] + x * (aabb[3:] - aabb[:3])).contiguous()
        pts = PointSet.from_positions(pos, None, contraction=True, aabb=f._aabb6())
        enc = f.mlp_base.encoder
        t_h = timed(lambda: ops.hash_encode_fwd(pts, enc.hash_table.detach(), enc.scalings, f._log2_T))
        e = ops.hash_encode_fwd(pts, enc.hash_table.detach(), enc.scalings, f._log2_T)
        t_d = timed(lambda: ops.density_fwd(pts, e, f._weights_dict()))
        print(f"level {lvl} {name:12s} hash {t_h*1e3:7.1f} us  density {t_d*1e3:6.1f} us")
