"""Build a variant of libumhs_sm100.so with extra -D flags for some translation units (A/B kernel experiments):

    python tools/build_variant.py <tag> <file.cu>:-DNAME=VALUE[,-D...] [...]   ->   gpurun_variants/libumhs_<tag>.so

Run with UMHS_LIB_PATH=gpurun_variants/libumhs_<tag>.so.  The objects of all other files come from the regular build."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unsupervised-hyperspectral-nerf_b200"))
from umhs_b200 import build as B  # noqa: E402

B.build()
tag = sys.argv[1]
over = dict(a.split(":", 1) for a in sys.argv[2:])
out_dir = os.path.join(ROOT, "gpurun_variants")
os.makedirs(out_dir, exist_ok=True)
objs = []
for src in B._sources():
    obj = os.path.join(B.OBJ_DIR, src.replace(".cu", ".o"))
    if src in over:
        obj = os.path.join(out_dir, f"{tag}_{src.replace('.cu', '.o')}")
        cmd = [B.NVCC, *B.ARCH, *B.COMMON, *B.EXTRA.get(src, []), *over[src].split(","), "-c", os.path.join(B.CSRC, src), "-o", obj]
        subprocess.run(cmd, check=True)
    objs.append(obj)
lib = os.path.join(out_dir, f"libumhs_{tag}.so")
subprocess.run([B.NVCC, *B.ARCH, "-shared", "-o", lib, *objs, "-lcudart"], check=True)
print(lib)
