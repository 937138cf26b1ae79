"""Time the tensor-core MLP kernels alone (CUDA events, 825 k samples by default): fwd, density, bwd."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "unsupervised-hyperspectral-nerf_b200")
import torch  # noqa: E402

from umhs_b200 import ops  # noqa: E402

DEV = "cuda"
N = int(os.environ.get("N", 825_000))
K = int(os.environ.get("K", 4))
SPEC = bool(int(os.environ.get("SPEC", "0")))
g = torch.Generator().manual_seed(0)
enc = (torch.randn(N, 32, generator=g) * 0.1).to(DEV)
pos = ((torch.rand(N, 3, generator=g) * 2 - 1)).to(DEV)
dirs = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1).to(DEV)
mk = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5).to(DEV).contiguous()  # noqa: E731
w = {}
for name, (o, i) in {"base_w0": (64, 32), "base_w1": (16, 64), "head_w0": (64, 27), "head_w1": (64, 64),
                     "head_w2": (K, 64), "feat_w0": (64, 27), "feat_w1": (64, 64), "feat_w2": (K + SPEC, 64),
                     "dir_w0": (16, 28)}.items():
    w[name] = mk(o, i)
    w[name.replace("_w", "_b")] = torch.zeros(o, device=DEV)
if not SPEC:
    w["dir_w0"] = w["dir_b0"] = None
grads = {"d_sigma": torch.randn(N, device=DEV), "d_scalar": torch.randn(N, K, device=DEV),
         "d_abund": torch.randn(N, K, device=DEV), "d_embedding": None}
if SPEC:
    grads["d_s1"] = torch.randn(N, device=DEV)
    grads["d_spec_hidden"] = torch.randn(N, 16, device=DEV)
pts = ops.PointSet.from_positions(pos, dirs)
big = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)  # L2 flush between timed launches


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        big.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


res = {"fwd": timed(lambda: ops.field_mlp_fwd(pts, enc, w, K, SPEC, 0.4, want_embedding=False)),
       "density": timed(lambda: ops.density_fwd(pts, enc, w)),
       "bwd": timed(lambda: ops.field_mlp_bwd(pts, enc, w, K, SPEC, 0.4, grads))}
print(os.environ.get("UMHS_LIB_PATH", "default"), f"N={N} K={K} spec={SPEC}",
      " ".join(f"{k}={v * 1e3:.1f}us" for k, v in res.items()))
