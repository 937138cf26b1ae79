"""Cycle trace of CTA 0 of the tensor-core MLP backward kernel: where does a round's latency go?
who = 0: thread 0 of data group 0 (tag 2r = operands of round r written / arrive, 2r+1 = accumulators of round r ready);
who = 1: the issuer (tag 1000 + 100*group + round, stamped after the commit)."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "unsupervised-hyperspectral-nerf_b200")
import torch  # noqa: E402

from umhs_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
g = torch.Generator().manual_seed(0)
K, N = 4, 128 * 148 * 8
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
MODE = sys.argv[1] if len(sys.argv) > 1 else "bwd"   # bwd | fwd | density
ROUNDS = {"bwd": 16, "fwd": 5, "density": 2}[MODE]


def run():
    if MODE == "bwd":
        ops.field_mlp_bwd(pts, enc, w, K, False, 0.4, grads, want_d_enc=os.environ.get("NO_DENC") is None)
    elif MODE == "fwd":
        ops.field_mlp_fwd(pts, enc, w, K, False, 0.4, want_embedding=False)
    else:
        ops.density_fwd(pts, enc, w)


run()
cap = 4096
buf = torch.zeros(4 * cap, dtype=torch.int64, device=DEV)
lib = _lib.load()
lib.umhs_debug_trace(buf.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
run()
torch.cuda.synchronize()
lib.umhs_debug_trace(None, 0, torch.cuda.current_stream().cuda_stream)
t = buf.view(cap, 2, 2).cpu()
ev = [(int(t[i, 0, 0]), int(t[i, 0, 1])) for i in range(cap) if int(t[i, 0, 1]) != 0]
print("data thread events:", len(ev), "(the first tile is skipped)")
if os.environ.get("RAW"):  # trace level 2 (-DUMHS_TRACE=2): extra stamps 200 + r (fence of round r done), 300 / 301 (outputs)
    for (ta, ca), (tb, cb) in list(zip(ev[:-1], ev[1:]))[:int(os.environ["RAW"])]:
        print(f"  {ta:4d} -> {tb:4d} {(cb - ca) & 0xFFFFFFFF:6d}")
    sys.exit(0)
seg = ev[:2 * ROUNDS + 1]  # second tile of group 0
tot_e = tot_w = 0
for (ta, ca), (tb, cb) in zip(seg[:-1], seg[1:]):
    kind = "wait (issue + tensor core)" if tb % 2 == 1 else "epilogue"
    dt = (cb - ca) & 0xFFFFFFFF
    tot_w += dt if tb % 2 == 1 else 0
    tot_e += dt if tb % 2 == 0 else 0
    print(f"  round {tb // 2:2d} {kind:28s} {dt:6d} cycles")
print(f"  cycles for the tile: {(seg[-1][1] - seg[0][1]) & 0xFFFFFFFF}  (epilogues {tot_e}, waits {tot_w})")
iv = [(int(t[i, 1, 0]), int(t[i, 1, 1])) for i in range(cap) if int(t[i, 1, 1]) != 0]
if os.environ.get("ABS"):
    # one round of group 0 on a common clock: operands written (data thread) -> ready seen by the issuer -> commit
    # issued -> accumulators ready (data thread woken)
    print("round: arrive->seen  seen->commit  commit->wake   (cycles, group 0, second traced tile)")
    it_ev = {}
    for tag, c in iv:
        it_ev.setdefault(tag, []).append(c)
    d_ev = {}
    for tag, c in ev:
        d_ev.setdefault(tag, []).append(c)
    for r in range(ROUNDS):
        try:
            a = d_ev[2 * r][1]
            seen = min(x for x in it_ev[3000 + r] if (x - a) & 0xFFFFFFFF < 1 << 20)
            com = min(x for x in it_ev[4000 + r] if (x - seen) & 0xFFFFFFFF < 1 << 20)
            w = min(x for x in d_ev[2 * r + 1] if (x - com) & 0xFFFFFFFF < 1 << 20)
            print(f"  {r:2d}: {(seen - a) & 0xFFFFFFFF:6d} {(com - seen) & 0xFFFFFFFF:6d} {(w - com) & 0xFFFFFFFF:6d}")
        except Exception as e:  # noqa: BLE001
            print(f"  {r:2d}: incomplete ({e})")
    sys.exit(0)
print("issuer events:", len(iv), "(tag, delta): 3xxx = ready signal seen, 2xxx = + tcgen05 fence, 4xxx = critical MMAs + commit issued, "
      "5xxx = deferred MMAs issued, 1xxx = warp reconverged")
rows = []
for (ta, ca), (tb, cb) in zip(iv[:-1], iv[1:]):
    rows.append(f"{tb:5d} {(cb - ca) & 0xFFFFFFFF:5d}")
for i in range(0, min(len(rows), 204), 6):
    print("   " + "   ".join(rows[i:i + 6]))
