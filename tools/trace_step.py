"""Cycle trace of the tensor-core MLP kernels INSIDE bench.py's config-2 training step (eager mode): the kept samples
are a gathered subset of the candidates' tiles and the points come in the ray form, unlike tools/trace_mlp.py.
Needs a -DUMHS_TRACE=1 build (tools/build_variant.py) selected with UMHS_LIB_PATH.  MODE = bwd | fwd picks which kernel's
stamps are read back (both write the same buffer; the other kernel's trace is overwritten by launching order, so the
hook below installs the buffer right before the wanted kernel)."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "unsupervised-hyperspectral-nerf_b200")
import torch  # noqa: E402

import bench  # noqa: E402
from umhs_b200 import _lib, ops  # noqa: E402

MODE = sys.argv[1] if len(sys.argv) > 1 else "bwd"
ROUNDS = {"bwd": 8 if os.environ.get("UMHS_SAVE_ACTIVATIONS", "1") != "0" else 16, "fwd": 5}[MODE]
D = bench.Dist()
run = bench.TrainRun(D, "hotdog", 16384, graph=False)
for i in range(bench.CONDITIONING_STEPS if hasattr(bench, "CONDITIONING_STEPS") else 8):
    run.step_resident(i)
torch.cuda.synchronize()
cap = 4096
buf = torch.zeros(4 * cap, dtype=torch.int64, device="cuda")
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
real = ops.field_mlp_bwd if MODE == "bwd" else ops.field_mlp_fwd


def traced(*a, **k):
    lib.umhs_debug_trace(buf.data_ptr(), cap, st)
    out = real(*a, **k)
    lib.umhs_debug_trace(None, 0, st)
    return out


if MODE == "bwd":
    ops.field_mlp_bwd = traced
else:
    ops.field_mlp_fwd = traced
run.step_resident(100)
torch.cuda.synchronize()
t = buf.view(cap, 2, 2).cpu()
ev = [(int(t[i, 0, 0]), int(t[i, 0, 1])) for i in range(cap) if int(t[i, 0, 1]) != 0]
iv = [(int(t[i, 1, 0]), int(t[i, 1, 1])) for i in range(cap) if int(t[i, 1, 1]) != 0]
print("data thread events:", len(ev))
seg = ev[:2 * ROUNDS + 1]
tot_e = tot_w = 0
for (ta, ca), (tb, cb) in zip(seg[:-1], seg[1:]):
    kind = "wait" if tb % 2 == 1 else "epilogue"
    dt = (cb - ca) & 0xFFFFFFFF
    tot_w += dt if tb % 2 == 1 else 0
    tot_e += dt if tb % 2 == 0 else 0
    print(f"  round {tb // 2:2d} {kind:9s} {dt:6d}")
print(f"  tile: {(seg[-1][1] - seg[0][1]) & 0xFFFFFFFF} cycles (epilogues {tot_e}, waits {tot_w})")
it_ev, d_ev = {}, {}
for tag, c in iv:
    it_ev.setdefault(tag, []).append(c)
for tag, c in ev:
    d_ev.setdefault(tag, []).append(c)
print("round: arrive->seen  seen->commit  commit->wake")
for r in range(ROUNDS):
    try:
        a = d_ev[2 * r][1]
        seen = min(x for x in it_ev[3000 + r] if (x - a) & 0xFFFFFFFF < 1 << 20)
        com = min(x for x in it_ev[4000 + r] if (x - seen) & 0xFFFFFFFF < 1 << 20)
        w = min(x for x in d_ev[2 * r + 1] if (x - com) & 0xFFFFFFFF < 1 << 20)
        print(f"  {r:2d}: {(seen - a) & 0xFFFFFFFF:6d} {(com - seen) & 0xFFFFFFFF:6d} {(w - com) & 0xFFFFFFFF:6d}")
    except Exception as e:  # noqa: BLE001
        print(f"  {r:2d}: incomplete ({e})")
