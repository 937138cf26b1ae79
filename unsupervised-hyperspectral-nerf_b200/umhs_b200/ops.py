"""Host-side operators: thin wrappers + torch.autograd.Functions over the C-ABI kernels.

Every function here enqueues hand-written sm_100a kernels from libumhs_sm100.so on torch's
current CUDA stream.  torch is used for device memory, streams and autograd plumbing only.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, f32, ptr, stream_ptr

ENC_DIM = 32  # 16 levels x 2 features

# kernels launched per C-ABI entry point (for the bench's gpu_launches count)
_LAUNCHES = {"umhs_hash_encode_fwd": 1, "umhs_hash_encode_bwd": 1, "umhs_hash_encode_fwd_tiles": 1,
             "umhs_hash_encode_bwd_tiles": 1, "umhs_hash_encode_bwd_tiles_levels": 1, "umhs_hash_encode_x01_fwd": 1,
             "umhs_hash_encode_x01_bwd": 1, "umhs_density_fwd": 1, "umhs_field_mlp_fwd": 1, "umhs_field_mlp_bwd": 1,
             "umhs_composite_fwd": 1, "umhs_specular_tc_fwd": 1, "umhs_specular_tc_bwd": 1, "umhs_spectral_grad": 1, "umhs_depth_clip": 1, "umhs_depth_clip_chunks": 2, "umhs_composite_bwd": 1, "umhs_unmix_fwd": 1,
             "umhs_unmix_bwd": 1, "umhs_render_weights_fwd": 1, "umhs_render_weights_bwd": 1,
             "umhs_accumulate_fwd": 1, "umhs_accumulate_bwd": 1, "umhs_march_count": 1, "umhs_march_level_any": 1, "umhs_occ_pack_bits": 1, "umhs_pack_scan": 1,
             "umhs_march_write": 1, "umhs_visibility_mask": 1, "umhs_compact_samples": 1, "umhs_pack_info": 2,
             "umhs_occ_cell_points": 1, "umhs_occ_select_cells": 2, "umhs_occ_ema_update": 2, "umhs_occ_binarize": 2, "umhs_adam_step": 1,
             "umhs_reduce_slabs": 1, "umhs_train_loss": 1, "umhs_dp_barrier": 1, "umhs_dp_adam_step": 1, "umhs_adam_step_dev": 1, "umhs_density_tc_fwd": 1, "umhs_field_mlp_tc_fwd": 1, "umhs_field_mlp_tc_bwd": 1, "umhs_field_mlp_tc_bwd_scatter": 1, "umhs_field_mlp_tc_bwd_saved": 1, "umhs_debug_trace": 1,
             "umhs_tc05_selftest": 1, "umhs_sample_batch": 1}

# "bf16": tensor-core MLP kernels (tcgen05 + TMEM, bf16 operands, fp32 accumulate) -- the default.
# "fp32": fp32 SIMT MLP kernels (precise mode for tight parity tests).
MLP_PRECISION = "bf16"
# True: the specular head's last layer ([N x 16] . [16 x B] + sigmoid + per-ray compositing) runs on the tensor cores
# (csrc/specular_tc.cu, bf16 operands) whenever MLP_PRECISION is "bf16"; False / fp32 mode: the fp32 warp-per-ray
# compositing kernel (csrc/composite.cu), the tight-parity bridge.
SPECULAR_TC = True
# True: between the hash-grid kernels and the tensor-core MLP kernels the encodings and their gradients travel as bf16
# TILES (8 KB per 128 samples, already in the tensor core's operand order: csrc/common.cuh enc_tile_piece) instead of
# fp32 rows -- half the bytes on five passes and no conversion / transposition in the MLP kernels.  Bit-identical
# forward (the tensor-core kernels round the encodings to bf16 on load anyway); fp32 mode always uses fp32 rows.
ENC_TILES = True
# The training forward keeps its hidden activations as bf16 operand tiles (704 B per sample) and the backward starts
# from them instead of recomputing the forward: 8 tensor-core rounds per tile instead of 16 (UMHS_SAVE_ACTIVATIONS=0:
# the recomputing backward).
SAVE_ACTIVATIONS = os.environ.get("UMHS_SAVE_ACTIVATIONS", "1") != "0"
# opt-in (UMHS_FUSED_SCATTER=1): the hash-grid scatter runs on 7 extra warps of the tensor-core MLP backward (one kernel
# instead of two back to back).  Measured and NOT the default: 0.66 ms against 0.63 ms for the two kernels at 819 k
# samples -- the reductions of the scatter warps and the shared-memory stores of the MLP epilogues share the SM's
# load/store path, and the MLP's chain of dependent rounds pays for every queued reduction.
FUSED_SCATTER = os.environ.get("UMHS_FUSED_SCATTER", "0") == "1"


def _tiles_on() -> bool:
    return ENC_TILES and MLP_PRECISION == "bf16"


def is_tile_buffer(enc: Tensor) -> bool:
    return enc.dtype == torch.bfloat16


class Profiler:
    """Optional per-entry-point CUDA-event timing on the launching stream + launch counting (bench.py)."""
    enabled = False
    timing = False
    launches = 0
    records: List[Tuple[str, "torch.cuda.Event", "torch.cuda.Event", int]] = []

    @classmethod
    def reset(cls, enabled=True, timing=False):
        cls.enabled, cls.timing, cls.launches, cls.records = enabled, timing, 0, []

    @classmethod
    def summary(cls) -> Dict[str, Dict[str, float]]:
        out: Dict[str, Dict[str, float]] = {}
        for name, a, b, units in cls.records:
            d = out.setdefault(name, {"calls": 0, "ms": 0.0, "units": 0})
            d["calls"] += 1
            d["ms"] += a.elapsed_time(b)
            d["units"] += int(units.reshape(-1)[0].item()) if isinstance(units, Tensor) else int(units)
        return out


def _call(name: str, units: int, *args, fn_name: Optional[str] = None) -> None:
    """Invoke a C-ABI entry point (kernels go to torch's current stream) and raise on error.  fn_name: the symbol to
    call when it differs from the name the launch is accounted under."""
    fn = getattr(_lib.load(), fn_name or name)
    if Profiler.enabled:
        Profiler.launches += _LAUNCHES[name]
        if Profiler.timing:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            Profiler.records.append((name, a, b, units))
            check(rc, name)
            return
    check(fn(*args), name)


# Data-parallel trainer, opt-in (UMHS_DP_SPLIT_LEVEL=8): scatter one half of the levels, call BETWEEN_SCATTER_HALVES()
# (the trainer starts that half's gradient exchange on a side stream), then the other half.  Measured on 8 x B200:
# 1.250 ms per step against 1.246 ms unsplit (2 GPUs: 1.227 / 1.229) -- what a step pays for data parallelism is the
# two barriers (the skew between ranks) and 58.8 MB per GPU over NVLink, which the second scatter half is too short
# and too busy (it shares the SMs with the exchange kernel) to hide.  0 = off.
SCATTER_SPLIT_LEVEL = 0
SCATTER_FINE_FIRST = os.environ.get("UMHS_DP_FINE_FIRST", "0") == "1"
BETWEEN_SCATTER_HALVES = None

# When True (set by UMHSTrainer), backward kernels accumulate straight into ``param.grad`` of leaf parameters
# that already own a contiguous fp32 grad buffer (views of the trainer's flat gradient), and autograd gets None.
DIRECT_GRAD = False


def _sink(p: Optional[Tensor]) -> Optional[Tensor]:
    if not DIRECT_GRAD or p is None or not isinstance(p, Tensor):
        return None
    g = p.grad if (p.is_leaf and p.requires_grad) else None
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or not g.is_cuda:
        return None
    return g


# --------------------------------------------------------------------------------------- points
@dataclass
class PointSet:
    """Sample points for the field kernels: explicit positions, or packed samples of a ray batch
    (position = o[ray] + d[ray] * (t0 + t1) / 2, nerfstudio Frustums.get_positions)."""
    n: int
    positions: Optional[Tensor] = None
    directions: Optional[Tensor] = None
    rays_o: Optional[Tensor] = None
    rays_d: Optional[Tensor] = None
    ray_indices: Optional[Tensor] = None
    t_starts: Optional[Tensor] = None
    t_ends: Optional[Tensor] = None
    contraction: bool = True
    aabb: Sequence[float] = (-1.0, -1.0, -1.0, 1.0, 1.0, 1.0)
    # optional device-side count (int64[>=1]): kernels process min(n, n_dev[0]) points; n is then a capacity
    n_dev: Optional[Tensor] = None
    # optional reuse of encodings computed for a superset of these points (the sampler's candidate pass):
    # enc_src [M, 32] and enc_index int32 [n] with enc(point i) = enc_src[enc_index[i]]
    enc_src: Optional[Tensor] = None
    enc_index: Optional[Tensor] = None

    def struct(self) -> _lib.Points:
        s = _lib.Points()
        s.positions = ptr(self.positions)
        s.directions = ptr(self.directions)
        s.rays_o = ptr(self.rays_o)
        s.rays_d = ptr(self.rays_d)
        s.ray_indices = ptr(self.ray_indices)
        s.t_starts = ptr(self.t_starts)
        s.t_ends = ptr(self.t_ends)
        s.contraction = 1 if self.contraction else 0
        s.n_dev = ptr(self.n_dev)
        s.enc_index = ptr(self.enc_index) if self.enc_src is not None else None
        for i, v in enumerate(self.aabb):
            s.aabb[i] = float(v)
        return s

    @property
    def device(self):
        return (self.positions if self.positions is not None else self.t_starts).device

    @staticmethod
    def from_positions(positions: Tensor, directions: Optional[Tensor] = None, contraction=True, aabb=None,
                       n_dev: Optional[Tensor] = None):
        p = f32(positions).reshape(-1, 3)
        d = f32(directions).reshape(-1, 3) if directions is not None else None
        return PointSet(n=p.shape[0], positions=p, directions=d, contraction=contraction,
                        aabb=tuple(aabb) if aabb is not None else PointSet.aabb, n_dev=n_dev)

    @staticmethod
    def from_packed(rays_o, rays_d, ray_indices, t_starts, t_ends, contraction=True, aabb=None, n_dev=None):
        return PointSet(n=int(t_starts.shape[0]), rays_o=f32(rays_o), rays_d=f32(rays_d),
                        ray_indices=ray_indices.to(torch.int64).contiguous(), t_starts=f32(t_starts).reshape(-1),
                        t_ends=f32(t_ends).reshape(-1), contraction=contraction,
                        aabb=tuple(aabb) if aabb is not None else PointSet.aabb, n_dev=n_dev)

    def units(self):
        """What the profiler should count as processed points (resolved after the timed region)."""
        return self.n_dev if self.n_dev is not None else self.n


def _grid_struct(table: Tensor, scalings: Tensor, log2_T: int) -> _lib.HashGrid:
    g = _lib.HashGrid()
    g.table = ptr(table)
    g.scalings = ptr(scalings)
    g.num_levels = int(scalings.shape[0])
    g.log2_table_size = int(log2_T)
    return g


# --------------------------------------------------------------------------------------- hash grid
def hash_encode_fwd(pts: PointSet, table: Tensor, scalings: Tensor, log2_T: int, tiles: bool = False) -> Tensor:
    """-> enc [n, 2L] fp32, or (tiles=True, L = 16) the bf16 tile buffer [ceil(n/128)*128, 32] (see ENC_TILES)."""
    lib = _lib.load()
    L = int(scalings.shape[0])
    ps, gs = pts.struct(), _grid_struct(table, scalings, log2_T)
    if tiles:
        enc = torch.empty(((pts.n + 127) // 128 * 128, 2 * L), dtype=torch.bfloat16, device=table.device)
        _call("umhs_hash_encode_fwd_tiles", pts.units(), ctypes.byref(ps), pts.n, ctypes.byref(gs), ptr(enc),
              stream_ptr(table.device))
        return enc
    enc = torch.empty((pts.n, 2 * L), dtype=torch.float32, device=table.device)
    _call("umhs_hash_encode_fwd", pts.units(), ctypes.byref(ps), pts.n, ctypes.byref(gs), ptr(enc), stream_ptr(table.device))
    return enc


def hash_encode_bwd(pts: PointSet, table: Tensor, scalings: Tensor, log2_T: int, d_enc: Tensor,
                    d_table: Optional[Tensor] = None, levels: Optional[Tuple[int, int]] = None) -> Tensor:
    """Scatter d_enc into d_table (accumulating).  Allocates a zero table gradient when none is given.
    levels = (begin, end): those levels only (tile buffers)."""
    lib = _lib.load()
    if d_table is None:
        d_table = torch.zeros_like(table, dtype=torch.float32)
    ps, gs = pts.struct(), _grid_struct(table, scalings, log2_T)
    if levels is not None:
        if not is_tile_buffer(d_enc):
            raise RuntimeError("a level range needs the bf16 tile buffer")
        _call("umhs_hash_encode_bwd_tiles_levels", pts.units(), ctypes.byref(ps), pts.n, ctypes.byref(gs), ptr(d_enc),
              ptr(d_table), int(levels[0]), int(levels[1]), stream_ptr(table.device))
        return d_table
    name = "umhs_hash_encode_bwd_tiles" if is_tile_buffer(d_enc) else "umhs_hash_encode_bwd"
    _call(name, pts.units(), ctypes.byref(ps), pts.n, ctypes.byref(gs), ptr(d_enc), ptr(d_table),
          stream_ptr(table.device))
    return d_table


def hash_encode_x01_fwd(x01: Tensor, table: Tensor, scalings: Tensor, log2_T: int) -> Tensor:
    lib = _lib.load()
    x = f32(x01).reshape(-1, 3)
    enc = torch.empty((x.shape[0], 2 * int(scalings.shape[0])), dtype=torch.float32, device=table.device)
    gs = _grid_struct(table, scalings, log2_T)
    _call("umhs_hash_encode_x01_fwd", x.shape[0], ptr(x), x.shape[0], ctypes.byref(gs), ptr(enc), stream_ptr(table.device))
    return enc


def hash_encode_x01_bwd(x01: Tensor, table: Tensor, scalings: Tensor, log2_T: int, d_enc: Tensor) -> Tensor:
    lib = _lib.load()
    x = f32(x01).reshape(-1, 3)
    d_table = torch.zeros_like(table, dtype=torch.float32)
    gs = _grid_struct(table, scalings, log2_T)
    _call("umhs_hash_encode_x01_bwd", x.shape[0], ptr(x), x.shape[0], ctypes.byref(gs), ptr(f32(d_enc)), ptr(d_table),
                                       stream_ptr(table.device))
    return d_table


class HashEncodeFunction(torch.autograd.Function):
    """HashEncoding.forward at its module boundary: x01 [N,3] in [0,1] -> [N, 2L]."""

    @staticmethod
    def forward(ctx, x01, table, scalings, log2_T):
        ctx.save_for_backward(x01, table, scalings)
        ctx.log2_T = log2_T
        return hash_encode_x01_fwd(x01, table, scalings, log2_T)

    @staticmethod
    def backward(ctx, d_enc):
        x01, table, scalings = ctx.saved_tensors
        return None, hash_encode_x01_bwd(x01, table, scalings, ctx.log2_T, d_enc.contiguous()), None, None


# --------------------------------------------------------------------------------------- field MLPs
MLP_NAMES = _lib._MLP_NAMES


def mlp_struct(weights: Dict[str, Optional[Tensor]], num_classes: int, specular: bool,
               temperature: float, enc_tiles: bool = False) -> _lib.MlpParams:
    s = _lib.MlpParams()
    for name in MLP_NAMES:
        setattr(s, name, ptr(weights.get(name)))
    s.num_classes = int(num_classes)
    s.specular = 1 if specular else 0
    s.temperature = float(temperature)
    s.enc_tiles = 1 if enc_tiles else 0
    if enc_tiles and MLP_PRECISION != "bf16":
        raise _lib.UmhsError("bf16 tile buffers are an interface of the tensor-core kernels only")
    return s


def density_fwd(pts: PointSet, enc: Tensor, weights: Dict[str, Tensor], zero_tail: bool = True) -> Tensor:
    lib = _lib.load()
    # with a device-side count the entries beyond it are not written: they read as zero density (zero_tail; a caller
    # that only reads the counted entries -- the sampler's visibility pass -- skips the fill)
    fill = pts.n_dev is not None and zero_tail
    sigma = (torch.zeros if fill else torch.empty)((pts.n,), dtype=torch.float32, device=enc.device)
    ps, ms = pts.struct(), mlp_struct(weights, 1, False, 1.0, is_tile_buffer(enc))
    _call("umhs_density_tc_fwd" if MLP_PRECISION == "bf16" else "umhs_density_fwd", pts.units(), ctypes.byref(ps), pts.n, ptr(enc), ctypes.byref(ms), ptr(sigma),
                               stream_ptr(enc.device))
    return sigma


ACT_TILE_BYTES = 90112  # include/umhs.h: UMHS_ACT_TILE_BYTES


def field_mlp_fwd(pts: PointSet, enc: Tensor, weights: Dict[str, Tensor], K: int, specular: bool,
                  temperature: float, want_embedding: bool = True, save_activations: bool = False) -> Dict[str, Tensor]:
    """save_activations (tensor-core path): also returns ``act_tiles``, the hidden activations as bf16 operand tiles
    (704 B per sample), for ``field_mlp_bwd(..., saved=...)``."""
    lib = _lib.load()
    dev, n = enc.device, pts.n
    out = {"sigma": torch.empty((n,), dtype=torch.float32, device=dev),
           "embedding": torch.empty((n, 15), dtype=torch.float32, device=dev) if want_embedding else None,
           "scalar": torch.empty((n, K), dtype=torch.float32, device=dev),
           "abund": torch.empty((n, K), dtype=torch.float32, device=dev),
           "s1": torch.empty((n,), dtype=torch.float32, device=dev) if specular else None,
           "spec_hidden": torch.empty((n, 16), dtype=torch.float32, device=dev) if specular else None}
    fo = _lib.FieldOut()
    for k, v in out.items():
        setattr(fo, k, ptr(v))
    if save_activations:
        if MLP_PRECISION != "bf16":
            raise RuntimeError("activations are saved by the tensor-core forward only")
        out["act_tiles"] = torch.empty(((n + 127) // 128) * ACT_TILE_BYTES, dtype=torch.uint8, device=dev)
        fo.act_tiles = ptr(out["act_tiles"])
    ps, ms = pts.struct(), mlp_struct(weights, K, specular, temperature, is_tile_buffer(enc))
    _call("umhs_field_mlp_tc_fwd" if MLP_PRECISION == "bf16" else "umhs_field_mlp_fwd", pts.units(), ctypes.byref(ps), n, ptr(enc), ctypes.byref(ms), ctypes.byref(fo), stream_ptr(dev))
    return out


def field_mlp_bwd(pts: PointSet, enc: Tensor, weights: Dict[str, Tensor], K: int, specular: bool, temperature: float,
                  grads: Dict[str, Optional[Tensor]], d_weights: Optional[Dict[str, Tensor]] = None,
                  want_d_enc: bool = True, scatter: Optional[tuple] = None,
                  saved: Optional[Dict[str, Tensor]] = None) -> Tuple[Optional[Tensor], Dict[str, Tensor]]:
    """-> (d_enc [n,32] or None, parameter grads).  Parameter grads are accumulated into d_weights if given.
    saved = the dictionary field_mlp_fwd(..., save_activations=True) returned: the backward then starts from the kept
    activation tiles and outputs instead of recomputing the forward (umhs_field_mlp_tc_bwd_saved, tile buffers only).
    scatter = (table, scalings, log2_T, d_table): the hash-grid scatter of d_enc into d_table (accumulating) runs inside
    the same kernel (tensor-core path with tile buffers only: umhs_field_mlp_tc_bwd_scatter)."""
    lib = _lib.load()
    dev, n = enc.device, pts.n
    if d_weights is None:
        d_weights = {k: torch.zeros_like(v, dtype=torch.float32) for k, v in weights.items() if v is not None}
    tiles = is_tile_buffer(enc)  # the gradient goes back in the layout the encodings came in
    d_enc = None
    if want_d_enc:
        d_enc = (torch.empty(((n + 127) // 128 * 128, ENC_DIM), dtype=torch.bfloat16, device=dev) if tiles else
                 torch.empty((n, ENC_DIM), dtype=torch.float32, device=dev))
    g = _lib.FieldOutGrads()
    for k in ("d_sigma", "d_embedding", "d_scalar", "d_abund", "d_s1", "d_spec_hidden"):
        setattr(g, k, ptr(grads.get(k)))
    dp = _lib.MlpGrads()
    for name in MLP_NAMES:
        setattr(dp, name, ptr(d_weights.get(name)))
    ps, ms = pts.struct(), mlp_struct(weights, K, specular, temperature, tiles)
    if saved is not None:
        if not (tiles and MLP_PRECISION == "bf16"):
            raise RuntimeError("the saved-activation backward needs the tensor-core path with tile buffers")
        sv = _lib.FieldSaved()
        for k in ("act_tiles", "sigma", "scalar", "abund", "s1", "spec_hidden"):
            setattr(sv, k, ptr(saved.get(k)))
        _call("umhs_field_mlp_tc_bwd_saved", pts.units(), ctypes.byref(ps), n, ptr(enc), ctypes.byref(ms), ctypes.byref(g),
              ctypes.byref(sv), ptr(d_enc), ctypes.byref(dp), stream_ptr(dev))
        return d_enc, d_weights
    if scatter is not None:
        if not (tiles and want_d_enc and MLP_PRECISION == "bf16"):
            raise RuntimeError("the fused scatter needs the tensor-core path with tile buffers")
        table, scalings, log2_T, d_table = scatter
        gs = _grid_struct(table, scalings, log2_T)
        _call("umhs_field_mlp_tc_bwd_scatter", pts.units(), ctypes.byref(ps), n, ptr(enc), ctypes.byref(ms),
              ctypes.byref(g), ptr(d_enc), ctypes.byref(dp), ctypes.byref(gs), ptr(d_table), stream_ptr(dev))
        return d_enc, d_weights
    _call("umhs_field_mlp_tc_bwd" if MLP_PRECISION == "bf16" else "umhs_field_mlp_bwd", pts.units(), ctypes.byref(ps), n, ptr(enc), ctypes.byref(ms), ctypes.byref(g), ptr(d_enc),
                                 ctypes.byref(dp), stream_ptr(dev))
    return d_enc, d_weights


class FieldFunction(torch.autograd.Function):
    """hash-grid encode + fused field MLPs.  inputs: table, then the MLP tensors in MLP_NAMES order
    (dir_* may be None).  outputs: sigma [N], embedding [N,15], scalar [N,K], abund [N,K], s1 [N], spec_hidden [N,16]
    (the last two are empty tensors when not specular)."""

    @staticmethod
    def forward(ctx, pts: PointSet, cfg: dict, table: Tensor, scalings: Tensor, *mlp: Optional[Tensor]):
        ctx.set_materialize_grads(False)  # undefined output gradients arrive as None, not as zero-filled tensors
        weights = {n: (w.detach() if w is not None else None) for n, w in zip(MLP_NAMES, mlp)}
        # kept samples are a subset of the sampler's candidates: reuse their encodings when the sampler kept them
        enc = pts.enc_src if pts.enc_src is not None else hash_encode_fwd(pts, table.detach(), scalings, cfg["log2_T"],
                                                                           tiles=_tiles_on())
        # training with the tensor-core kernels on tile buffers: keep the activation tiles for the backward
        save = (SAVE_ACTIVATIONS and cfg.get("save_activations", False) and is_tile_buffer(enc)
                and MLP_PRECISION == "bf16" and not FUSED_SCATTER and any(ctx.needs_input_grad))
        out = field_mlp_fwd(pts, enc, weights, cfg["K"], cfg["specular"], cfg["temperature"],
                            want_embedding=cfg.get("want_embedding", True), save_activations=save)
        ctx.saved_fwd = out if save else None
        ctx.pts, ctx.cfg = pts, cfg
        ctx.n_mlp = len(mlp)
        ctx.param_refs = (table, mlp)  # the Parameter objects themselves (for direct gradient accumulation)
        ctx.save_for_backward(enc, table, scalings, *[w for w in mlp if w is not None])
        ctx.mlp_present = [w is not None for w in mlp]
        empty = torch.empty((0,), device=enc.device)
        s1 = out["s1"] if out["s1"] is not None else empty
        sh = out["spec_hidden"] if out["spec_hidden"] is not None else empty
        emb = out["embedding"] if out["embedding"] is not None else empty
        return out["sigma"], emb, out["scalar"], out["abund"], s1, sh

    @staticmethod
    def backward(ctx, d_sigma, d_emb, d_scalar, d_abund, d_s1, d_sh):
        enc, table, scalings, *present = ctx.saved_tensors
        it = iter(present)
        weights = {n: (next(it) if p else None) for n, p in zip(MLP_NAMES, ctx.mlp_present)}
        cfg, pts = ctx.cfg, ctx.pts
        n, K, dev = pts.n, cfg["K"], enc.device

        def c(g, shape):
            return torch.zeros(shape, dtype=torch.float32, device=dev) if g is None else f32(g)

        grads = {"d_sigma": c(d_sigma, (n,)), "d_embedding": None if (d_emb is None or d_emb.numel() == 0) else f32(d_emb),
                 "d_scalar": c(d_scalar, (n, K)), "d_abund": c(d_abund, (n, K))}
        if cfg["specular"]:
            grads["d_s1"] = c(d_s1, (n,))
            grads["d_spec_hidden"] = c(d_sh, (n, 16))
        p_table, p_mlp = ctx.param_refs
        sinks = {n: _sink(w) for n, w in zip(MLP_NAMES, p_mlp)}
        d_w_in = {n: (sinks[n] if sinks[n] is not None else torch.zeros_like(w, dtype=torch.float32))
                  for n, w in weights.items() if w is not None}
        if BEFORE_MLP_BACKWARD is not None:
            BEFORE_MLP_BACKWARD()  # trainer hook: fork work that should run beside the MLP backward as well
        t_sink = _sink(p_table)
        if FUSED_SCATTER and is_tile_buffer(enc) and MLP_PRECISION == "bf16":
            d_table = t_sink if t_sink is not None else torch.zeros_like(table, dtype=torch.float32)
            d_enc, d_w = field_mlp_bwd(pts, enc, weights, K, cfg["specular"], cfg["temperature"], grads, d_w_in,
                                       scatter=(table, scalings, cfg["log2_T"], d_table))
            if AFTER_MLP_BACKWARD is not None:
                AFTER_MLP_BACKWARD()
        else:
            d_enc, d_w = field_mlp_bwd(pts, enc, weights, K, cfg["specular"], cfg["temperature"], grads, d_w_in,
                                       saved=ctx.saved_fwd)
            ctx.saved_fwd = None
            if AFTER_MLP_BACKWARD is not None:
                AFTER_MLP_BACKWARD()  # trainer hook: fork work that should run beside the hash scatter and the optimiser
            L = int(scalings.shape[0])
            if BETWEEN_SCATTER_HALVES is not None and is_tile_buffer(d_enc) and 0 < SCATTER_SPLIT_LEVEL < L:
                # one half of the levels first; the hook (data-parallel trainer) starts its gradient exchange on a side
                # stream while the other half is being scattered
                first, second = (0, SCATTER_SPLIT_LEVEL), (SCATTER_SPLIT_LEVEL, L)
                if SCATTER_FINE_FIRST:
                    first, second = second, first
                d_table = hash_encode_bwd(pts, table, scalings, cfg["log2_T"], d_enc, t_sink, levels=first)
                BETWEEN_SCATTER_HALVES()
                hash_encode_bwd(pts, table, scalings, cfg["log2_T"], d_enc, d_table, levels=second)
            else:
                d_table = hash_encode_bwd(pts, table, scalings, cfg["log2_T"], d_enc, t_sink)
        return (None, None, None if t_sink is not None else d_table, None,
                *[(d_w.get(n) if (p and sinks[n] is None) else None) for n, p in zip(MLP_NAMES, ctx.mlp_present)])


# --------------------------------------------------------------------------------------- compositing
def _composite_in(R, N, K, B, specular, grad_scaling, packed_info, t0, t1, sigma, scalar, abund, s1, spec_hidden,
                  endmembers, dir_w1, dir_b1, rgb_matrix, palette) -> _lib.CompositeIn:
    ci = _lib.CompositeIn()
    ci.n_rays, ci.n_samples, ci.num_classes, ci.num_bands = R, N, K, B
    ci.specular, ci.grad_scaling = (1 if specular else 0), (1 if grad_scaling else 0)
    ci.packed_info, ci.t_starts, ci.t_ends, ci.sigma = ptr(packed_info), ptr(t0), ptr(t1), ptr(sigma)
    ci.scalar, ci.abund = ptr(scalar), ptr(abund)
    ci.s1 = ptr(s1) if specular else None
    ci.spec_hidden = ptr(spec_hidden) if specular else None
    ci.endmembers, ci.rgb_matrix = ptr(endmembers), ptr(rgb_matrix)
    ci.dir_w1 = ptr(dir_w1) if specular else None
    ci.dir_b1 = ptr(dir_b1) if specular else None
    ci.palette = ptr(palette)
    return ci


_TMM_INIT: Dict[torch.device, Tensor] = {}


def _tminmax_init(dev) -> Tensor:
    t = _TMM_INIT.get(dev)
    if t is None:
        t = _TMM_INIT[dev] = torch.tensor([float("inf"), float("-inf")], dtype=torch.float32, device=dev)
    return t


def specular_tc_fwd(spec_hidden: Tensor, s1: Tensor, weights: Tensor, ray_indices: Tensor, n_rays: int, dir_w1: Tensor,
                    dir_b1: Tensor, n_dev: Optional[Tensor] = None) -> Tensor:
    """specular[r, b] = sum_{n in r} w_n s1_n sigmoid(W1[b,:] . h_n + b1[b]) on the tensor cores (include/umhs.h)."""
    N, B = int(spec_hidden.shape[0]), int(dir_w1.shape[0])
    out = torch.empty((n_rays, B), dtype=torch.float32, device=spec_hidden.device)
    _call("umhs_specular_tc_fwd", N, N, ptr(n_dev), n_rays, B, ptr(spec_hidden), ptr(s1), ptr(weights), ptr(ray_indices),
          ptr(dir_w1), ptr(dir_b1), ptr(out), stream_ptr(out.device))
    return out


def specular_tc_bwd(spec_hidden: Tensor, s1: Tensor, weights: Tensor, t0: Tensor, t1: Tensor, ray_indices: Tensor,
                    n_rays: int, dir_w1: Tensor, dir_b1: Tensor, d_spec: Tensor, grad_scaling: bool,
                    d_w1: Tensor, d_b1: Tensor, n_dev: Optional[Tensor] = None):
    """-> (d_spec_hidden [N,16], d_s1 [N], d_weights [N]); d_w1 / d_b1 are accumulated into."""
    N, B = int(spec_hidden.shape[0]), int(dir_w1.shape[0])
    dev = spec_hidden.device
    d_sh = torch.empty((N, 16), dtype=torch.float32, device=dev)
    d_s1 = torch.empty((N,), dtype=torch.float32, device=dev)
    d_w = torch.empty((N,), dtype=torch.float32, device=dev)
    _call("umhs_specular_tc_bwd", N, N, ptr(n_dev), n_rays, B, 1 if grad_scaling else 0, ptr(spec_hidden), ptr(s1),
          ptr(weights), ptr(t0), ptr(t1), ptr(ray_indices), ptr(dir_w1), ptr(dir_b1), ptr(d_spec), ptr(d_sh), ptr(d_s1),
          ptr(d_w), ptr(d_w1), ptr(d_b1), stream_ptr(dev))
    return d_sh, d_s1, d_w


COMPOSITE_OUTPUTS = ("spectral", "spectral2", "specular", "abundances", "accumulation", "depth", "rgb", "seg_probs",
                     "seg_raw", "seg_pred", "weights")


class CompositeFunction(torch.autograd.Function):
    """Per-ray compositing x unmixing x spec->sRGB x cluster probe over packed samples (fused kernel).
    Differentiable outputs: spectral, spectral2, rgb, accumulation.  Everything else is marked
    non-differentiable (the reference renders abundances / specular under no_grad and uses depth and the
    segmentation outputs in no loss)."""

    @staticmethod
    def forward(ctx, cfg: dict, packed_info, t0, t1, sigma, scalar, abund, s1, spec_hidden, endmembers, dir_w1,
                dir_b1, rgb_matrix, palette):
        lib = _lib.load()
        ctx.set_materialize_grads(False)  # undefined output gradients arrive as None, not as zero-filled tensors
        dev = sigma.device
        R, N, K, B, specular = int(packed_info.shape[0]), int(sigma.shape[0]), cfg["K"], cfg["B"], cfg["specular"]
        ctx.param_refs = (endmembers, dir_w1, dir_b1)
        tens = [f32(x) if x is not None and x.numel() else None
                for x in (t0, t1, sigma, scalar, abund, s1, spec_hidden, endmembers, dir_w1, dir_b1, rgb_matrix,
                          palette)]
        t0, t1, sigma, scalar, abund, s1, spec_hidden, endmembers, dir_w1, dir_b1, rgb_matrix, palette = tens
        z = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        # tensor-core specular path: weights first (the transmittance scan), then the [N x 16] . [16 x B] product +
        # sigmoid + per-ray sums on tcgen05, then the lane-group compositing kernel adds the result to the unmixed
        # spectrum before its colour / probe epilogue
        spec_tc = bool(specular and SPECULAR_TC and MLP_PRECISION == "bf16" and cfg.get("ray_indices") is not None)
        ctx.spec_tc = spec_tc
        w_pre = specular_rb = None
        if spec_tc:
            ri, n_dev = cfg["ray_indices"], cfg.get("n_dev")
            w_pre, t_pre = z(N), z(N)
            _call("umhs_render_weights_fwd", 0, ptr(packed_info), R, ptr(t0), ptr(t1), ptr(sigma), ptr(w_pre), ptr(t_pre),
                  stream_ptr(dev))
            specular_rb = z(R, B)
            _call("umhs_specular_tc_fwd", N, N, ptr(n_dev), R, B, ptr(spec_hidden), ptr(s1), ptr(w_pre), ptr(ri),
                  ptr(dir_w1), ptr(dir_b1), ptr(specular_rb), stream_ptr(dev))
        ci = _composite_in(R, N, K, B, specular and not spec_tc, cfg["grad_scaling"], packed_info, t0, t1, sigma, scalar,
                           abund, s1, spec_hidden, endmembers, dir_w1, dir_b1, rgb_matrix, palette)
        ci.specular_in = ptr(specular_rb)
        out = {"spectral": z(R, B), "spectral2": z(R, B) if specular else None,
               "specular": (specular_rb if spec_tc else z(R, B)) if specular else None, "abundances": z(R, K),
               "accumulation": z(R),
               "depth": z(R), "rgb": z(R, 3), "seg_probs": z(R, K), "seg_raw": z(R), "seg_pred": z(R, 3),
               "weights": z(N) if cfg.get("want_weights", False) else None}
        # per-sample transmittance, kept for the backward (4 B/sample instead of a recomputation pass there)
        trans = z(N) if any(ctx.needs_input_grad) else None
        want_depth = cfg.get("want_depth", True)  # a trainer that never reads the depth skips it (3 launches)
        if not want_depth:
            out["depth"] = None
        co = _lib.CompositeOut()
        for k, v in out.items():
            setattr(co, k, ptr(v))
        co.trans = ptr(trans)
        # depth clip bounds: the whole call (training batches, one reference forward) or, in evaluation, the reference's
        # 512-ray forward calls: cfg["depth_chunks"] = (chunk_rays, ray_ids | None), see umhs_depth_clip_chunks
        chunks = cfg.get("depth_chunks") if want_depth else None
        if want_depth and chunks is None:
            tmm = _tminmax_init(dev).clone()  # (+inf, -inf) without a host->device copy per call
            co.t_minmax = ptr(tmm)
        _call("umhs_composite_fwd", R, ctypes.byref(ci), ctypes.byref(co), stream_ptr(dev))
        if want_depth and chunks is None:
            _call("umhs_depth_clip", 0, ptr(out["depth"]), R, ptr(tmm), stream_ptr(dev))
        elif want_depth:
            chunk_rays, ray_ids = int(chunks[0]), chunks[1]
            n_chunks = (int(chunks[2]) if len(chunks) > 2 else R) // chunk_rays + 1
            bounds = _tminmax_init(dev).repeat(n_chunks)
            _call("umhs_depth_clip_chunks", 0, ptr(out["depth"]), R, ptr(packed_info), ptr(t0), ptr(t1),
                  ptr(ray_ids), chunk_rays, ptr(bounds), stream_ptr(dev))
        ctx.cfg = cfg
        ctx.save_for_backward(packed_info, t0, t1, sigma, scalar, abund, s1, spec_hidden, endmembers, dir_w1, dir_b1,
                              rgb_matrix, out["spectral"], trans, w_pre)
        empty = torch.empty((0,), device=dev)
        res = tuple(out[k] if out[k] is not None else empty for k in COMPOSITE_OUTPUTS)
        ctx.mark_non_differentiable(*[res[i] for i, k in enumerate(COMPOSITE_OUTPUTS)
                                      if k not in ("spectral", "spectral2", "rgb", "accumulation")])
        return res

    @staticmethod
    def backward(ctx, d_spectral, d_spectral2, d_specular, d_abundances, d_acc, d_depth, d_rgb, *_):
        lib = _lib.load()
        (packed_info, t0, t1, sigma, scalar, abund, s1, spec_hidden, endmembers, dir_w1, dir_b1, rgb_matrix,
         spectral, trans, w_pre) = ctx.saved_tensors
        cfg = ctx.cfg
        dev = sigma.device
        R, N, K, B, specular = int(packed_info.shape[0]), int(sigma.shape[0]), cfg["K"], cfg["B"], cfg["specular"]
        spec_tc = ctx.spec_tc
        ci = _composite_in(R, N, K, B, specular and not spec_tc, cfg["grad_scaling"], packed_info, t0, t1, sigma, scalar,
                           abund, s1, spec_hidden, endmembers, dir_w1, dir_b1, rgb_matrix, None)
        gi = _lib.CompositeGradsIn()
        keep = [f32(g) if g is not None else None for g in (d_spectral, d_spectral2 if specular else None, d_rgb,
                                                            d_acc)]
        z = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        d_sigma, d_scalar, d_abund = z(N), z(N, K), z(N, K)
        d_s1 = z(N) if specular else None
        d_sh = z(N, 16) if specular else None
        sE, sW, sB = (_sink(p) for p in ctx.param_refs)
        d_E = sE if sE is not None else torch.zeros((K, B), dtype=torch.float32, device=dev)
        d_w1 = (sW if sW is not None else torch.zeros((B, 16), dtype=torch.float32, device=dev)) if specular else None
        d_b1 = (sB if sB is not None else torch.zeros((B,), dtype=torch.float32, device=dev)) if specular else None
        if spec_tc:
            # dL / d (composited spectrum) first (loss gradient + the spec -> sRGB chain), then the tensor-core
            # backward of the specular term, whose dL/dw joins the compositing sweep below
            d_spec = z(R, B)
            have = keep[0] is not None or keep[2] is not None
            if have:
                _call("umhs_spectral_grad", R, R, B, ptr(keep[0]), ptr(keep[2]), ptr(spectral), ptr(rgb_matrix),
                      ptr(d_spec), stream_ptr(dev))
            else:
                d_spec.zero_()
            dw_x = z(N)
            _call("umhs_specular_tc_bwd", N, N, ptr(cfg.get("n_dev")), R, B, 1 if cfg["grad_scaling"] else 0,
                  ptr(spec_hidden), ptr(s1), ptr(w_pre), ptr(t0), ptr(t1), ptr(cfg["ray_indices"]), ptr(dir_w1),
                  ptr(dir_b1), ptr(d_spec), ptr(d_sh), ptr(d_s1), ptr(dw_x), ptr(d_w1), ptr(d_b1), stream_ptr(dev))
            d_unmixed = d_spec if keep[1] is None else d_spec + keep[1]  # spectral2 = the unmixed part alone
            keep = [d_unmixed, None, None, keep[3], dw_x]
            gi.d_weights = ptr(dw_x)
        gi.d_spectral, gi.d_spectral2, gi.d_rgb, gi.d_accumulation = (ptr(k) for k in keep[:4])
        gi.spectral = ptr(spectral)
        gi.trans = ptr(trans)
        go = _lib.CompositeGradsOut()
        go.d_sigma, go.d_scalar, go.d_abund = ptr(d_sigma), ptr(d_scalar), ptr(d_abund)
        go.d_s1, go.d_spec_hidden = (None, None) if spec_tc else (ptr(d_s1), ptr(d_sh))
        go.d_endmembers = ptr(d_E)
        go.d_dir_w1, go.d_dir_b1 = (None, None) if spec_tc else (ptr(d_w1), ptr(d_b1))
        _call("umhs_composite_bwd", R, ctypes.byref(ci), ctypes.byref(gi), ctypes.byref(go), stream_ptr(dev))
        return (None, None, None, None, d_sigma, d_scalar, d_abund, d_s1, d_sh, None if sE is not None else d_E,
                None if sW is not None else d_w1, None if sB is not None else d_b1, None, None)


# --------------------------------------------------------------------------------------- unfused module-boundary ops
class UnmixFunction(torch.autograd.Function):
    """Per-sample linear mixing model of UMHSField.get_outputs (ref umhs_field.py:229-251):
    spec[n,b] = sum_k s a E[k,b]; specular term s1 * sigmoid(W h + b).  -> (spec, specular_term)"""

    @staticmethod
    def forward(ctx, scalar, abund, s1, spec_hidden, endmembers, dir_w1, dir_b1):
        lib = _lib.load()
        specular = s1 is not None
        scalar, abund, endmembers = f32(scalar), f32(abund), f32(endmembers)
        n, K = scalar.shape
        B = endmembers.shape[1]
        dev = scalar.device
        spec = torch.empty((n, B), dtype=torch.float32, device=dev)
        sp = torch.empty((n, B), dtype=torch.float32, device=dev) if specular else None
        if specular:
            s1, spec_hidden, dir_w1, dir_b1 = f32(s1).reshape(-1), f32(spec_hidden), f32(dir_w1), f32(dir_b1)
        _call("umhs_unmix_fwd", 0, n, K, B, int(specular), ptr(scalar), ptr(abund), ptr(s1), ptr(spec_hidden),
                                 ptr(endmembers), ptr(dir_w1), ptr(dir_b1), ptr(spec), ptr(sp), stream_ptr(dev))
        ctx.specular = specular
        ctx.save_for_backward(scalar, abund, s1, spec_hidden, endmembers, dir_w1, dir_b1)
        return spec, (sp if specular else torch.empty((0,), device=dev))

    @staticmethod
    def backward(ctx, d_spec, d_sp):
        lib = _lib.load()
        scalar, abund, s1, spec_hidden, endmembers, dir_w1, dir_b1 = ctx.saved_tensors
        specular = ctx.specular
        n, K = scalar.shape
        B = endmembers.shape[1]
        dev = scalar.device
        d_spec = f32(d_spec) if d_spec is not None else torch.zeros((n, B), device=dev)
        d_sp = f32(d_sp) if (specular and d_sp is not None) else None
        d_scalar, d_abund = torch.empty_like(scalar), torch.empty_like(abund)
        d_E = torch.zeros_like(endmembers)
        d_s1 = torch.zeros_like(s1) if specular else None
        d_sh = torch.zeros_like(spec_hidden) if specular else None
        d_w1 = torch.zeros_like(dir_w1) if specular else None
        d_b1 = torch.zeros_like(dir_b1) if specular else None
        _call("umhs_unmix_bwd", 0, n, K, B, int(specular), ptr(scalar), ptr(abund), ptr(s1), ptr(spec_hidden),
                                 ptr(endmembers), ptr(dir_w1), ptr(dir_b1), ptr(d_spec), ptr(d_sp), ptr(d_scalar),
                                 ptr(d_abund), ptr(d_s1), ptr(d_sh), ptr(d_E), ptr(d_w1), ptr(d_b1), stream_ptr(dev))
        return d_scalar, d_abund, d_s1, d_sh, d_E, d_w1, d_b1


def pack_info(ray_indices: Tensor, n_rays: int) -> Tensor:
    """nerfacc.pack_info: [n_rays, 2] (start, count), int64."""
    lib = _lib.load()
    ri = ray_indices.to(torch.int64).contiguous()
    dev = ri.device
    counts = torch.empty((n_rays,), dtype=torch.int64, device=dev)
    packed = torch.empty((n_rays, 2), dtype=torch.int64, device=dev)
    _call("umhs_pack_info", 0, ptr(ri), ri.shape[0], n_rays, ptr(counts), ptr(packed), stream_ptr(dev))
    return packed


class RenderWeightsFunction(torch.autograd.Function):
    """nerfacc.render_weight_from_density on packed samples -> (weights, transmittance)."""

    @staticmethod
    def forward(ctx, t0, t1, sigma, packed_info):
        lib = _lib.load()
        t0, t1, sigma = f32(t0).reshape(-1), f32(t1).reshape(-1), f32(sigma).reshape(-1)
        w, T = torch.empty_like(sigma), torch.empty_like(sigma)
        _call("umhs_render_weights_fwd", 0, ptr(packed_info), packed_info.shape[0], ptr(t0), ptr(t1), ptr(sigma),
                                          ptr(w), ptr(T), stream_ptr(sigma.device))
        ctx.save_for_backward(t0, t1, sigma, packed_info, T)
        ctx.mark_non_differentiable(T)
        return w, T

    @staticmethod
    def backward(ctx, d_w, _):
        lib = _lib.load()
        t0, t1, sigma, packed_info, T = ctx.saved_tensors
        d_sigma = torch.empty_like(sigma)
        _call("umhs_render_weights_bwd", 0, ptr(packed_info), packed_info.shape[0], ptr(t0), ptr(t1), ptr(sigma),
                                          ptr(T), ptr(f32(d_w).reshape(-1)), ptr(d_sigma), stream_ptr(sigma.device))
        return None, None, d_sigma, None


class AccumulateFunction(torch.autograd.Function):
    """nerfacc.accumulate_along_rays on packed samples: out[ray[n], :] += w[n] * v[n, :]."""

    @staticmethod
    def forward(ctx, weights, values, ray_indices, n_rays):
        lib = _lib.load()
        w = f32(weights).reshape(-1)
        v = f32(values) if values is not None else None
        C = v.shape[-1] if v is not None else 1
        ri = ray_indices.to(torch.int64).contiguous()
        out = torch.zeros((n_rays, C), dtype=torch.float32, device=w.device)
        _call("umhs_accumulate_fwd", 0, ptr(w), ptr(v), ptr(ri), w.shape[0], C, ptr(out), stream_ptr(w.device))
        ctx.save_for_backward(w, v, ri)
        ctx.has_values = v is not None
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        w, v, ri = ctx.saved_tensors
        C = v.shape[-1] if ctx.has_values else 1
        d_w = torch.empty_like(w)
        d_v = torch.empty_like(v) if ctx.has_values else None
        _call("umhs_accumulate_bwd", 0, ptr(w), ptr(v), ptr(ri), w.shape[0], C, ptr(f32(d_out)), ptr(d_w), ptr(d_v),
                                      stream_ptr(w.device))
        return d_w, d_v, None, None


# --------------------------------------------------------------------------------------- packing
AFTER_MLP_BACKWARD = None  # optional callable, see FieldFunction.backward / trainer.UMHSTrainer._capture
BEFORE_MLP_BACKWARD = None
MARCH_STAGE = 192  # staged interval starts per ray (0 disables staging: the write pass marches again)
MARCH_BITS = True   # march on a bit-packed copy of the occupancy grid (identical results)


def occ_pack_bits(binaries: Tensor, bits: Optional[Tensor] = None, level_any: Optional[Tensor] = None):
    """Bit-packed copy of ``binaries`` [levels, res, res, res] (bit b of word w = cell 32 w + b) and the per-level
    "any occupied" flags, in one pass.  Writes into ``bits`` / ``level_any`` when given (persistent buffers)."""
    dev = binaries.device
    bn = binaries.contiguous().view(torch.uint8) if binaries.dtype == torch.bool else binaries.contiguous()
    a = _lib.MarchArgs()
    a.binaries, a.levels, a.resolution = ptr(bn), int(bn.shape[0]), int(bn.shape[1])
    words = (a.resolution ** 3 + 31) // 32
    if bits is None:
        bits = torch.empty((a.levels, words), dtype=torch.int32, device=dev)
    if level_any is None:
        level_any = torch.empty((8,), dtype=torch.int32, device=dev)
    _call("umhs_occ_pack_bits", 0, ctypes.byref(a), ptr(bits), ptr(level_any), stream_ptr(dev))
    return bits, level_any


def march(rays_o: Tensor, rays_d: Tensor, near_planes: Tensor, far_planes: Tensor, binaries: Tensor, aabbs: Tensor,
          step_size: float, cone_angle: float, capacity: Optional[int] = None,
          packed_grid: Optional[Tuple[Tensor, Tensor]] = None):
    """Occupancy-grid ray marching (nerfacc traverse_grids).  -> (ray_indices, t_starts, t_ends, packed_info, total).

    capacity=None: exact-sized outputs, one host sync on the sample count (as the reference does).
    capacity=C:    no host sync; outputs hold C entries of which the first total[0] (device int64) are valid, the
                   packing is truncated at C samples and total[1] keeps the unclamped count for a later check."""
    dev = rays_o.device
    R = int(rays_o.shape[0])
    a = _lib.MarchArgs()
    ro, rd, nr, fr = f32(rays_o), f32(rays_d), f32(near_planes).reshape(-1), f32(far_planes).reshape(-1)
    bn = binaries.contiguous().view(torch.uint8) if binaries.dtype == torch.bool else binaries.contiguous()
    ab = f32(aabbs)
    a.rays_o, a.rays_d, a.near_planes, a.far_planes = ptr(ro), ptr(rd), ptr(nr), ptr(fr)
    a.binaries, a.aabbs = ptr(bn), ptr(ab)
    a.levels, a.resolution = int(bn.shape[0]), int(bn.shape[1])
    a.step_size, a.cone_angle, a.n_rays = float(step_size), float(cone_angle), R
    counts = torch.empty((R,), dtype=torch.int64, device=dev)
    packed = torch.empty((R, 2), dtype=torch.int64, device=dev)
    total = torch.empty((2,), dtype=torch.int64, device=dev)  # both entries are written by umhs_pack_scan
    st = stream_ptr(dev)
    a.level_any = a.occ_bits = None
    if packed_grid is not None and MARCH_BITS:
        # the grid owner keeps the bit-packed copy up to date (OccGridEstimator.packed_grid): nothing to pack here
        bits, level_any = packed_grid
        a.occ_bits = ptr(bits)
    elif MARCH_BITS:
        level_any = torch.empty((8,), dtype=torch.int32, device=dev)
        # one pass over binaries: per-level "any occupied" flags + a bit-packed copy whose lookups hit L1
        words = (a.resolution ** 3 + 31) // 32
        bits = torch.empty((a.levels, words), dtype=torch.int32, device=dev)
        _call("umhs_occ_pack_bits", 0, ctypes.byref(a), ptr(bits), ptr(level_any), st)
        a.occ_bits = ptr(bits)
    else:
        level_any = torch.empty((8,), dtype=torch.int32, device=dev)
        _call("umhs_march_level_any", 0, ctypes.byref(a), ptr(level_any), st)
    a.level_any = ptr(level_any)
    # the count pass stages each ray's first MARCH_STAGE interval starts; the write pass copies them
    stage = torch.empty((R, MARCH_STAGE), dtype=torch.float32, device=dev) if MARCH_STAGE > 0 else None
    a.stage, a.stage_counts, a.stage_stride = ptr(stage), ptr(counts) if stage is not None else None, MARCH_STAGE
    _call("umhs_march_count", R, ctypes.byref(a), ptr(counts), st)
    _call("umhs_pack_scan", R, ptr(counts), R, ptr(packed), ptr(total), int(capacity or 0), st)
    N = int(total[0].item()) if capacity is None else int(capacity)
    ri = torch.empty((N,), dtype=torch.int64, device=dev)
    t0 = torch.empty((N,), dtype=torch.float32, device=dev)
    t1 = torch.empty((N,), dtype=torch.float32, device=dev)
    if N > 0:
        _call("umhs_march_write", R, ctypes.byref(a), ptr(packed), ptr(ri), ptr(t0), ptr(t1), st)
    return ri, t0, t1, packed, total


def visibility_compact(packed_info: Tensor, t0: Tensor, t1: Tensor, sigma: Tensor, early_stop_eps: float,
                       alpha_thre: float, occ_mean: Optional[Tensor] = None, capacity: Optional[int] = None):
    """render_visibility_from_density + boolean-mask compaction.  -> (ray_indices, t0, t1, packed_info, total).
    capacity: see ``march`` (no host sync; outputs sized `capacity`, total[0] valid entries)."""
    dev = t0.device
    R, N = int(packed_info.shape[0]), int(t0.shape[0])
    keep = torch.empty((N,), dtype=torch.uint8, device=dev)
    kc = torch.empty((R,), dtype=torch.int64, device=dev)
    packed2 = torch.empty((R, 2), dtype=torch.int64, device=dev)
    total = torch.empty((2,), dtype=torch.int64, device=dev)  # both entries are written by umhs_pack_scan
    st = stream_ptr(dev)
    _call("umhs_visibility_mask", R, ptr(packed_info), R, ptr(t0), ptr(t1), ptr(f32(sigma).reshape(-1)),
          float(early_stop_eps), float(alpha_thre), ptr(occ_mean), ptr(keep), ptr(kc), st)
    _call("umhs_pack_scan", R, ptr(kc), R, ptr(packed2), ptr(total), 0, st)
    M = int(total[0].item()) if capacity is None else int(capacity)
    ri = torch.empty((M,), dtype=torch.int64, device=dev)
    o0 = torch.empty((M,), dtype=torch.float32, device=dev)
    o1 = torch.empty((M,), dtype=torch.float32, device=dev)
    src = torch.empty((M,), dtype=torch.int32, device=dev)
    if M > 0:
        _call("umhs_compact_samples", R, ptr(packed_info), ptr(packed2), R, ptr(keep), ptr(t0), ptr(t1), ptr(ri),
              ptr(o0), ptr(o1), ptr(src), st)
    return ri, o0, o1, packed2, total, src


# --------------------------------------------------------------------------------------- optimiser
def adam_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, lr: float, beta1: float, beta2: float,
              eps: float, step: int, inv_scale: Optional[Tensor] = None, found_inf: Optional[Tensor] = None,
              clamp_range: Tuple[int, int] = (0, 0), zero_grad: bool = True) -> None:
    lib = _lib.load()
    _call("umhs_adam_step", param.numel(), ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), lr, beta1, beta2,
                             eps, step, ptr(inv_scale), ptr(found_inf), clamp_range[0], clamp_range[1],
                             1 if zero_grad else 0, stream_ptr(param.device))


def adam_step_dev(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, beta1: float, beta2: float,
                  eps: float, hyper: Tensor, inv_scale: Optional[Tensor] = None,
                  clamp_range: Tuple[int, int] = (0, 0), zero_grad: bool = True) -> None:
    """Adam step whose lr / bias corrections come from the device tensor hyper = [lr, 1-b1^t, sqrt(1-b2^t)]."""
    _call("umhs_adam_step_dev", param.numel(), ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(),
          beta1, beta2, eps, ptr(hyper), ptr(inv_scale), None, clamp_range[0], clamp_range[1], 1 if zero_grad else 0,
          stream_ptr(param.device))


def train_loss(spectral: Tensor, hs_image: Tensor, rgb: Optional[Tensor], accumulation: Optional[Tensor],
               image: Optional[Tensor], background: Optional[Tensor], w_spectral: float, w_rgb: float,
               blend_pred: bool = True):
    """Both training losses and d(sum)/d(spectral, rgb, accumulation) in one launch.  ``blend_pred``: the
    prediction is blended over ``background`` too (nerfstudio does that for background_color="random" only; a named
    colour blends just an RGBA ground truth).
    -> (losses [2], d_spectral [R,B], d_rgb [R,3] | None, d_accumulation [R] | None)"""
    dev = spectral.device
    R, B = int(spectral.shape[0]), int(spectral.shape[1])
    sp, hs = f32(spectral), f32(hs_image)
    losses = torch.empty((2,), dtype=torch.float32, device=dev)
    d_spec = torch.empty((R, B), dtype=torch.float32, device=dev)
    use_rgb = w_rgb != 0.0 and rgb is not None
    d_rgb = torch.empty((R, 3), dtype=torch.float32, device=dev) if use_rgb else None
    d_acc = torch.empty((R,), dtype=torch.float32, device=dev) if use_rgb else None
    keep = [f32(t) if (t is not None and use_rgb) else None for t in (rgb, accumulation, image, background)]
    ch = int(image.shape[-1]) if image is not None else 3
    _call("umhs_train_loss", R, R, B, ch, ptr(sp), ptr(hs), ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]),
          1 if blend_pred else 0, float(w_spectral), float(w_rgb if use_rgb else 0.0), ptr(losses), ptr(d_spec), ptr(d_rgb), ptr(d_acc),
          stream_ptr(dev))
    return losses, d_spec, d_rgb, d_acc


def dp_barrier(pads: Tensor, rank: int, world: int, channel: int, epochs: Tensor) -> None:
    """Cross-GPU barrier on peer-mapped signal pads (pads: device int64 [world] of pad addresses; epochs: device
    int32 [8] counters of this rank, one per channel).  Enqueued on the current stream; capturable."""
    _call("umhs_dp_barrier", 0, ptr(pads), rank, world, channel, epochs.data_ptr() + 4 * channel, stream_ptr(pads.device))


def dp_adam_step(peer_grads: List[int], peer_params: List[int], exp_avg: Tensor, exp_avg_sq: Tensor, n: int, rank: int,
                 world: int, beta1: float, beta2: float, eps: float, hyper: Tensor,
                 clamp_range: Tuple[int, int] = (0, 0), mc_grad: int = 0, mc_param: int = 0,
                 span: Optional[Tuple[int, int]] = None) -> None:
    """Fused reduce-scatter + Adam + all-gather over NVLink peer memory (include/umhs.h: umhs_dp_adam_step).
    mc_grad / mc_param: NVLS multicast addresses of the buffers (0: plain peer loads / stores).
    span = (begin, end): that slice of the flat buffers only (umhs_dp_adam_step_range)."""
    pg = (ctypes.c_void_p * world)(*peer_grads)
    pp = (ctypes.c_void_p * world)(*peer_params)
    lo, hi = (0, n) if span is None else span
    _call("umhs_dp_adam_step", (hi - lo) // world, pg, pp, ptr(exp_avg), ptr(exp_avg_sq), n, lo, hi, rank, world, beta1,
          beta2, eps, ptr(hyper), clamp_range[0], clamp_range[1], mc_grad or None, mc_param or None,
          stream_ptr(exp_avg.device), fn_name="umhs_dp_adam_step_range")


def tc05_selftest(mode: int, A: Tensor, B: Tensor, N: int, K: int) -> Tensor:
    """One-CTA tcgen05 GEMM self-test (see include/umhs.h)."""
    D = torch.zeros((64 if mode >= 2 else 128, N), dtype=torch.float32, device=A.device)
    _call("umhs_tc05_selftest", 0, mode, ptr(f32(A)), ptr(f32(B)), ptr(D), N, K, stream_ptr(A.device))
    return D
