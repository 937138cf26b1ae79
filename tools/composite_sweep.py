"""BASELINE config 5: band / endmember sweep of the compositing x unmixing kernel (forward + backward), roofline study.
R rays x S samples, B in {32..512}, K in {4, 8, 16}; achieved GB/s of ALGORITHMIC bytes against
the measured HBM copy bandwidth.  Writes a markdown table to stdout (committed as profiles/r1_composite_sweep.md)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-hyperspectral-nerf_b200"))
import torch  # noqa: E402

from umhs_b200 import ops  # noqa: E402

DEV = "cuda"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def run(R, S, B, K, iters=20):
    g = torch.Generator().manual_seed(B * 100 + K)
    N = R * S
    packed = torch.stack([torch.arange(R) * S, torch.full((R,), S)], -1).to(torch.int64).to(DEV)
    t0 = (0.05 + 0.0035 * torch.arange(S).repeat(R)).to(DEV)
    t1 = t0 + 0.0035
    sigma = (torch.rand(N, generator=g) * 20).to(DEV).requires_grad_()
    scalar = torch.rand(N, K, generator=g).to(DEV).requires_grad_()
    abund = torch.softmax(torch.randn(N, K, generator=g), -1).to(DEV).requires_grad_()
    E = torch.rand(K, B, generator=g).to(DEV).requires_grad_()
    M = (torch.rand(B, 3, generator=g) / B).to(DEV)
    palette = torch.rand(max(K, 16), 3, generator=g).to(DEV)
    cfg = {"K": K, "B": B, "specular": False, "grad_scaling": True, "want_weights": False}
    d_spec = torch.rand(R, B, generator=g).to(DEV)
    d_rgb = torch.rand(R, 3, generator=g).to(DEV)

    def fwd():
        return ops.CompositeFunction.apply(cfg, packed, t0, t1, sigma, scalar, abund, None, None, E, None, None, M, palette)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.Profiler.reset(enabled=True, timing=True)
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        summ = ops.Profiler.summary()
        ops.Profiler.reset(enabled=False)
        return {k: v["ms"] / v["calls"] for k, v in summ.items()}

    def step():
        for t in (sigma, scalar, abund, E):
            t.grad = None
        out = fwd()
        torch.autograd.backward([out[0], out[6]], [d_spec, d_rgb])

    ms = timed(step)
    fb = R * (S * (12 + 8 * K) + 16 + 4 * B + 4 * K + 20)
    bb = R * (2 * S * (12 + 8 * K) + 8 * B)
    return ms.get("umhs_composite_fwd", float("nan")), ms.get("umhs_composite_bwd", float("nan")), fb, bb


print("# r1 compositing x unmixing sweep (BASELINE config 5)\n")
print(f"`python tools/composite_sweep.py` on 1 x B200; per-launch CUDA-event times; GB/s of algorithmic bytes; peak = "
      f"{peak:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs).  The reference's unfused chain moves about S*B*4*3 bytes per ray.\n")
for R, S in ((4096, 48), (65536, 48)):
    print(f"\n## R = {R} rays x S = {S} samples (N = {R * S})\n")
    print("| B | K | fwd us | fwd GB/s | fwd % of peak | bwd us | bwd GB/s | bwd % of peak | reference-chain bytes / ours (fwd) |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for B in (32, 64, 128, 256, 512):
        for K in (4, 8, 16):
            f, b, fb, bb = run(R, S, B, K)
            ref_bytes = R * S * B * 4 * 3
            print(f"| {B} | {K} | {f * 1e3:.1f} | {fb / f / 1e6:.0f} | {100 * fb / f / 1e6 / peak:.1f} | {b * 1e3:.1f} | "
                  f"{bb / b / 1e6:.0f} | {100 * bb / b / 1e6 / peak:.1f} | {ref_bytes / fb:.0f}x |")
