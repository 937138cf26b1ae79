"""ctypes front-end of the C packing oracle (``pack_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of ``pack_oracle.c``.  Compiles the C file
with ``gcc -O2 -ffp-contract=off`` into ``oracle/libpack_oracle.so`` on first use
(``__graft_entry__.build()`` does the same ahead of time).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Tuple

import numpy as np
import torch
from torch import Tensor

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pack_oracle.c")
_SO = os.path.join(_HERE, "libpack_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def traverse_grids(rays_o: Tensor, rays_d: Tensor, binaries: Tensor, aabbs: Tensor, near_planes: Tensor,
                   far_planes: Tensor, step_size: float, cone_angle: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (ray_indices int64 [N], t_starts f32 [N], t_ends f32 [N], packed_info int64 [R,2])."""
    lib = _load()
    ro = np.ascontiguousarray(rays_o.detach().numpy(), dtype=np.float32)
    rd = np.ascontiguousarray(rays_d.detach().numpy(), dtype=np.float32)
    bn = np.ascontiguousarray(binaries.detach().numpy().astype(np.uint8))
    ab = np.ascontiguousarray(aabbs.detach().numpy(), dtype=np.float32)
    nr = np.ascontiguousarray(near_planes.detach().numpy(), dtype=np.float32)
    fr = np.ascontiguousarray(far_planes.detach().numpy(), dtype=np.float32)
    levels, res = int(bn.shape[0]), int(bn.shape[1])
    R = ro.shape[0]
    counts = np.zeros(R, dtype=np.int64)
    f = ctypes.c_float
    lib.umhs_oracle_march_count(_p(ro), _p(rd), _p(bn), _p(ab), ctypes.c_int(levels), ctypes.c_int(res), _p(nr),
                                _p(fr), f(step_size), f(cone_angle), ctypes.c_int64(R), _p(counts))
    offsets = np.cumsum(counts) - counts
    N = int(counts.sum())
    ray_idx = np.zeros(N, dtype=np.int64)
    t0 = np.zeros(N, dtype=np.float32)
    t1 = np.zeros(N, dtype=np.float32)
    lib.umhs_oracle_march_write(_p(ro), _p(rd), _p(bn), _p(ab), ctypes.c_int(levels), ctypes.c_int(res), _p(nr),
                                _p(fr), f(step_size), f(cone_angle), ctypes.c_int64(R), _p(offsets), _p(ray_idx),
                                _p(t0), _p(t1))
    packed = np.stack([offsets, counts], axis=-1)
    return torch.from_numpy(ray_idx), torch.from_numpy(t0), torch.from_numpy(t1), torch.from_numpy(packed)
