"""Tensor-core MLP backward from saved activations (umhs_field_mlp_tc_bwd_saved) against the recomputing backward:
parity of d_enc and of the weight gradients, and the times of forward (with / without saving) and backward."""
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
L, LOG2T = 16, 19
g = torch.Generator().manual_seed(0)
pos = ((torch.rand(N, 3, generator=g) * 2 - 1)).to(DEV)
dirs = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1).to(DEV)
pts = ops.PointSet.from_positions(pos, dirs)
table = ((torch.rand(L << LOG2T, 2, generator=g) * 2 - 1) * 1e-1).to(DEV)
growth = (4096 / 16) ** (1 / (L - 1))
scalings = torch.tensor([float(int(16 * growth ** i)) for i in range(L)], device=DEV)
mk = lambda o_, i_: ((torch.rand(o_, i_, generator=g) * 2 - 1) / i_ ** 0.5).to(DEV).contiguous()  # noqa: E731
w = {}
for name, (o_, i_) in {"base_w0": (64, 32), "base_w1": (16, 64), "head_w0": (64, 27), "head_w1": (64, 64),
                       "head_w2": (K, 64), "feat_w0": (64, 27), "feat_w1": (64, 64), "feat_w2": (K + SPEC, 64),
                       "dir_w0": (16, 28)}.items():
    w[name] = mk(o_, i_)
    w[name.replace("_w", "_b")] = (torch.rand(o_, generator=g) * 0.1).to(DEV)
if not SPEC:
    w["dir_w0"] = w["dir_b0"] = None
grads = {"d_sigma": torch.randn(N, device=DEV), "d_scalar": torch.randn(N, K, device=DEV),
         "d_abund": torch.randn(N, K, device=DEV), "d_embedding": torch.randn(N, 15, device=DEV)}
if SPEC:
    grads["d_s1"] = torch.randn(N, device=DEV)
    grads["d_spec_hidden"] = torch.randn(N, 16, device=DEV)
enc = ops.hash_encode_fwd(pts, table, scalings, LOG2T, tiles=True)
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


plain = ops.field_mlp_fwd(pts, enc, w, K, SPEC, 0.4)
saved = ops.field_mlp_fwd(pts, enc, w, K, SPEC, 0.4, save_activations=True)
torch.cuda.synchronize()
for k in ("sigma", "scalar", "abund"):
    assert torch.equal(plain[k], saved[k]), k
d_a, gw_a = ops.field_mlp_bwd(pts, enc, w, K, SPEC, 0.4, grads)
d_b, gw_b = ops.field_mlp_bwd(pts, enc, w, K, SPEC, 0.4, grads, saved=saved)
torch.cuda.synchronize()
da, db = d_a.view(-1)[: (N // 128) * 4096].float(), d_b.view(-1)[: (N // 128) * 4096].float()
print("d_enc (whole tiles): max |a| %.4g  max |a - b| %.4g  rel L2 %.3g" %
      (float(da.abs().max()), float((da - db).abs().max()), float((da - db).norm() / da.norm())))
for k in gw_a:
    print("  %-8s rel L2 %.3g" % (k, float((gw_a[k] - gw_b[k]).norm() / (gw_a[k].norm() + 1e-30))))
res = {"fwd": timed(lambda: ops.field_mlp_fwd(pts, enc, w, K, SPEC, 0.4, want_embedding=False)),
       "fwd_save": timed(lambda: ops.field_mlp_fwd(pts, enc, w, K, SPEC, 0.4, want_embedding=False, save_activations=True)),
       "bwd": timed(lambda: ops.field_mlp_bwd(pts, enc, w, K, SPEC, 0.4, grads)),
       "bwd_saved": timed(lambda: ops.field_mlp_bwd(pts, enc, w, K, SPEC, 0.4, grads, saved=saved))}
print(f"N={N} K={K} spec={SPEC}", " ".join(f"{k}={v * 1e3:.1f}us" for k, v in res.items()))
