"""Time the tensor-core MLP backward + hash-grid scatter, as two kernels and as the fused kernel
(umhs_field_mlp_tc_bwd_scatter), on ray-ordered samples; check that both give the same table gradient."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "unsupervised-hyperspectral-nerf_b200")
import torch  # noqa: E402

from umhs_b200 import ops  # noqa: E402

DEV = "cuda"
R = int(os.environ.get("R", 16384))
S = int(os.environ.get("S", 50))
K = int(os.environ.get("K", 4))
SPEC = bool(int(os.environ.get("SPEC", "0")))
N = R * S
L, LOG2T = 16, 19
g = torch.Generator().manual_seed(0)
# rays through the unit box, samples at a fixed step along each ray (consecutive samples of a ray are neighbours)
o = (torch.rand(R, 3, generator=g) * 1.6 - 0.8)
d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
t = (torch.arange(S).float() * (1.6 / 1024 * 3 ** 0.5) - 0.2)
pos = (o[:, None, :] + t[None, :, None] * d[:, None, :]).reshape(N, 3).to(DEV)
dirs = d[:, None, :].expand(R, S, 3).reshape(N, 3).contiguous().to(DEV)
pts = ops.PointSet.from_positions(pos, dirs, contraction=False, aabb=(-1, -1, -1, 1, 1, 1))
table = ((torch.rand(L << LOG2T, 2, generator=g) * 2 - 1) * 1e-1).to(DEV)
growth = (4096 / 16) ** (1 / (L - 1))
scalings = torch.tensor([float(int(16 * growth ** i)) for i in range(L)], device=DEV)
mk = lambda o_, i_: ((torch.rand(o_, i_, generator=g) * 2 - 1) / i_ ** 0.5).to(DEV).contiguous()  # noqa: E731
w = {}
for name, (o_, i_) in {"base_w0": (64, 32), "base_w1": (16, 64), "head_w0": (64, 27), "head_w1": (64, 64),
                       "head_w2": (K, 64), "feat_w0": (64, 27), "feat_w1": (64, 64), "feat_w2": (K + SPEC, 64),
                       "dir_w0": (16, 28)}.items():
    w[name] = mk(o_, i_)
    w[name.replace("_w", "_b")] = torch.zeros(o_, device=DEV)
if not SPEC:
    w["dir_w0"] = w["dir_b0"] = None
grads = {"d_sigma": torch.randn(N, device=DEV), "d_scalar": torch.randn(N, K, device=DEV),
         "d_abund": torch.randn(N, K, device=DEV), "d_embedding": None}
if SPEC:
    grads["d_s1"] = torch.randn(N, device=DEV)
    grads["d_spec_hidden"] = torch.randn(N, 16, device=DEV)
enc = ops.hash_encode_fwd(pts, table, scalings, LOG2T, tiles=True)
big = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)  # L2 flush between timed launches
d_table = torch.zeros_like(table)


def two():
    d_enc, dw = ops.field_mlp_bwd(pts, enc, w, K, SPEC, 0.4, grads)
    ops.hash_encode_bwd(pts, table, scalings, LOG2T, d_enc, d_table)
    return dw


def mlp_only():
    return ops.field_mlp_bwd(pts, enc, w, K, SPEC, 0.4, grads)[1]


def fused():
    return ops.field_mlp_bwd(pts, enc, w, K, SPEC, 0.4, grads, scatter=(table, scalings, LOG2T, d_table))[1]


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


if os.environ.get("ONLY") == "fused":
    for _ in range(3):
        fused()
    torch.cuda.synchronize()
    sys.exit(0)
d_table.zero_()
dw_a = two()
ga = d_table.clone()
d_table.zero_()
dw_b = fused()
torch.cuda.synchronize()
gb = d_table.clone()
scale = float(ga.abs().max())
print("table grad: max |a| %.4g, max |a - b| %.4g, nonzero a %d b %d" % (scale, float((ga - gb).abs().max()),
                                                                      int((ga != 0).sum()), int((gb != 0).sum())))
for k in dw_a:
    print("  %-8s max |a - b| / max |a| = %.3g" % (k, float((dw_a[k] - dw_b[k]).abs().max()) / (float(dw_a[k].abs().max()) + 1e-30)))
res = {"mlp_bwd": timed(mlp_only), "two_kernels": timed(two), "fused": timed(fused)}
print(f"N={N} K={K} spec={SPEC}", " ".join(f"{k}={v * 1e3:.1f}us" for k, v in res.items()))
