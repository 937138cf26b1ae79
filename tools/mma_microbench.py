"""Issue / execution cycles of the tcgen05 GEMM forms the MLP kernels are built from (one CTA, zeroed operands)."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "unsupervised-hyperspectral-nerf_b200")
import torch  # noqa: E402

from umhs_b200 import _lib  # noqa: E402

lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
names = {0: "kmajor  D[128xN] = A[128xK] B[NxK]^T", 1: "dgrad   D[128xN] = A[128xK] W[KxN]", 2: "wgrad   D[64xN] += P^T Q (K=128)",
         3: "bias    one K=16 step"}
for mode, N, K in [(0, 64, 32), (0, 64, 64), (0, 16, 64), (0, 128, 32), (0, 32, 64), (1, 64, 64), (1, 64, 16), (1, 32, 64),
                   (2, 72, 128), (2, 32, 128), (2, 16, 128), (2, 40, 128), (3, 64, 16), (3, 16, 16)]:
    for reps in (1, 8, 32):
        lib.umhs_tc05_microbench(mode, N, K, 1, out.data_ptr(), st)  # warm the instruction cache
        rc = lib.umhs_tc05_microbench(mode, N, K, reps, out.data_ptr(), st)
        torch.cuda.synchronize()
        assert rc == 0
        n_mma = reps * (1 if mode == 3 else (8 if mode == 2 else K // 16))
        i, t = out.tolist()
        print(f"{names[mode]:40s} N={N:3d} K={K:3d} reps={reps}: {n_mma:3d} MMAs  issue {i:6d} cyc ({i / n_mma:6.1f}/MMA)  "
              f"done {t:6d} cyc ({t / n_mma:6.1f}/MMA)")
