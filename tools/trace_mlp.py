"""Cycle trace of one CTA of the tensor-core MLP backward kernel: where does a round's latency go?"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "unsupervised-hyperspectral-nerf_b200")
import torch  # noqa: E402

from umhs_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
g = torch.Generator().manual_seed(0)
K, N = 4, 128 * 148 * 4
enc = torch.randn(N, 32, generator=g).to(DEV)
pos = ((torch.rand(N, 3, generator=g) * 2 - 1)).to(DEV)
mk = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5).to(DEV)  # noqa: E731
w = {}
for name, (o, i) in {"base_w0": (64, 32), "base_w1": (16, 64), "head_w0": (64, 27), "head_w1": (64, 64),
                     "head_w2": (K, 64), "feat_w0": (64, 27), "feat_w1": (64, 64), "feat_w2": (K, 64)}.items():
    w[name] = mk(o, i).contiguous()
    w[name.replace("_w", "_b")] = torch.zeros(o, device=DEV)
w["dir_w0"] = w["dir_b0"] = None
grads = {"d_sigma": torch.randn(N, device=DEV), "d_scalar": torch.randn(N, K, device=DEV),
         "d_abund": torch.randn(N, K, device=DEV), "d_embedding": None}
pts = ops.PointSet.from_positions(pos)
ops.field_mlp_bwd(pts, enc, w, K, False, 0.4, grads)
cap = 4096
buf = torch.zeros(4 * cap, dtype=torch.int64, device=DEV)
lib = _lib.load()
lib.umhs_debug_trace(buf.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
ops.field_mlp_bwd(pts, enc, w, K, False, 0.4, grads)
torch.cuda.synchronize()
lib.umhs_debug_trace(None, 0, torch.cuda.current_stream().cuda_stream)
t = buf.view(cap, 2, 2).cpu()
# thread 0 = issuer, thread 96 = a plain epilogue thread; print deltas for the second tile (16 rounds x 4 tags)
for who in (0, 1):
    ev = [(int(t[i, who, 0]), int(t[i, who, 1])) for i in range(cap) if int(t[i, who, 1]) != 0]
    per_tile = 15 * 4
    seg = ev[per_tile:2 * per_tile + 1]
    print("thread", 0 if who == 0 else 96, "events:", len(ev))
    names = {1: "epilogue done", 2: "barrier passed", 3: "issued/committed", 4: "accumulators ready"}
    line = []
    for (ta, ca), (tb, cb) in zip(seg[:-1], seg[1:]):
        line.append(f"{names[ta][:9]}->{names[tb][:9]}:{cb - ca}")
    for i in range(0, len(line), 4):
        print("   ", " | ".join(line[i:i + 4]))
    tot = seg[-1][1] - seg[0][1]
    print("   cycles for the tile:", tot)
