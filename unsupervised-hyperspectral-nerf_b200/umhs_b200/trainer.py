"""Data-parallel training step for UMHSModel: one process per GPU, rays sharded across ranks.

What nerfstudio's Trainer does around ``get_train_loss_dict`` for this method (ref
umhsnerf/umhs_config.py:34-69: Adam lr 2e-2 eps 1e-15, exponential decay to 1e-5 over 30 k steps; ref
umhsnerf/umhs_model.py:542-591: occupancy update before, endmember clamp after each iteration), restated for the
fused path:

  * all field parameters live in ONE flat fp32 buffer (the nn.Parameters are views into it), with flat grad /
    exp_avg / exp_avg_sq twins, so the optimiser is a single fused kernel (csrc/optim.cu: Adam + endmember clamp +
    grad zeroing) and the gradient exchange is a single collective;
  * every rank renders its own shard of the ray batch; one NCCL all-reduce (sum) of the flat gradient per step over
    NVLink / NVSwitch, averaged (DDP semantics: losses are means over the local shard), then identical Adam on every
    rank.  Evaluation shards image tiles with no collective.

The reference itself never syncs gradients (its pipeline forces world_size = 1: ref umhsnerf/umhs_pipeline.py:86,
108-113); this is the data-parallel path the north star asks for.
"""
from __future__ import annotations

import os
import sys
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor, nn

from . import ops
from .rays import RayBundle
from .umhs_model import UMHSModel


def tile_pixel_indices(H: int, W: int, tile: int, rank: int, world_size: int, device=None) -> Tensor:
    """Flat pixel ids (row-major) of the tile x tile image tiles dealt round-robin to ``rank``.  The union over ranks
    covers every pixel exactly once: evaluation needs no collective."""
    ty, tx = (H + tile - 1) // tile, (W + tile - 1) // tile
    ids = []
    for t in range(rank, ty * tx, world_size):
        y0, x0 = (t // tx) * tile, (t % tx) * tile
        ys = torch.arange(y0, min(y0 + tile, H), device=device)
        xs = torch.arange(x0, min(x0 + tile, W), device=device)
        ids.append((ys[:, None] * W + xs[None, :]).reshape(-1))
    return torch.cat(ids) if ids else torch.zeros((0,), dtype=torch.int64, device=device)


def average_gradients(flat_grad: Tensor, world_size: int) -> None:
    """The one exchange step of data-parallel training: all-reduce(sum) of the flat gradient.  The 1/world_size
    factor (DDP semantics: each rank's loss is a mean over its own ray shard) is applied inside the fused Adam."""
    if world_size > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)


class PeerMemory:
    """Peer-mapped (symmetric) device buffers of one process group: every rank allocates the same sizes and gets the
    device addresses of all ranks' copies, mapped over NVLink (torch.distributed symmetric memory does the handle
    exchange -- plumbing; the kernels that use the pointers are csrc/optim.cu)."""

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self._sm = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self._keep = []

    def empty(self, n: int, dtype, device):
        """-> (local tensor, [device address of rank w's tensor for w in range(world)])"""
        t = self._sm.empty(n, dtype=dtype, device=device)
        hdl = self._sm.rendezvous(t, self.group)
        self._keep.append(hdl)
        # NVLS multicast address of the same buffers (0 when the fabric / driver has no multicast support)
        mc = 0
        try:
            # Per-direction link traffic of the exchange, in buffer sizes: peer loads / stores 2 (N-1)/N, multimem
            # 1 + 1/N (a rank's own shard also travels to the switch and back) -- multimem pays from 4 GPUs up
            # (measured on B200: N=2 0.115 vs 0.188 ms, N=4 0.162 vs 0.166 ms, N=8 0.208 vs 0.160 ms).
            want = os.environ.get("UMHS_MULTICAST")  # "1" / "0" override the rule
            use = (self.world >= 4) if want is None else want == "1"
            if use and getattr(hdl, "has_multicast_support", False):
                mc = int(hdl.multicast_ptr or 0)
        except Exception:  # noqa: BLE001
            mc = 0
        self.last_multicast_ptr = mc
        return t, [int(p) for p in hdl.buffer_ptrs]


class FlatParams:
    """Re-home a list of parameters into one contiguous fp32 buffer (16-byte aligned segments).  With ``peers`` the
    parameter and gradient buffers are peer-mapped, for the fused data-parallel optimiser step."""

    def __init__(self, params: List[nn.Parameter], peers: Optional[PeerMemory] = None):
        dev = params[0].device
        self.params = params
        self.offsets: List[int] = []
        n = 0
        for p in params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.peer_data = self.peer_grad = None
        self.mc_data = self.mc_grad = 0
        if peers is not None:
            self.data, self.peer_data = peers.empty(n, torch.float32, dev)
            self.mc_data = peers.last_multicast_ptr
            self.grad, self.peer_grad = peers.empty(n, torch.float32, dev)
            self.mc_grad = peers.last_multicast_ptr
            if not (self.mc_data and self.mc_grad):
                self.mc_data = self.mc_grad = 0
            self.data.zero_()
            self.grad.zero_()
        else:
            self.data = torch.zeros(n, dtype=torch.float32, device=dev)
            self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(params, self.offsets):
            view = self.data[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def range_of(self, p: nn.Parameter):
        for q, o in zip(self.params, self.offsets):
            if q is p:
                return o, o + q.numel()
        raise KeyError("parameter not in the flat buffer")


class UMHSTrainer:
    def __init__(self, model: UMHSModel, lr: float = 2e-2, eps: float = 1e-15, betas=(0.9, 0.999),
                 lr_final: float = 1e-5, max_steps: int = 30000, world_size: int = 1, rank: int = 0,
                 occ_update: bool = True, freeze_binaries: bool = False, start_step: int = 0,
                 sync_free: bool = True, capacity_check_every: int = 16, cuda_graph: bool = False,
                 fused_dp: bool = True, fused_loss: bool = True):
        self.model = model
        self.fused_loss = fused_loss
        self.world_size, self.rank = world_size, rank
        self.lr_init, self.lr_final, self.max_steps = lr, lr_final, max_steps
        self.eps, self.betas = eps, betas
        params = [p for p in model.field.parameters() if p.requires_grad]
        # fused_dp: gradient exchange + Adam as one kernel over NVLink peer memory (csrc/optim.cu); otherwise one
        # NCCL all-reduce followed by the single-GPU fused Adam on every rank
        self.fused_dp = bool(fused_dp and world_size > 1 and params[0].is_cuda)
        self._peers = PeerMemory() if self.fused_dp else None
        self.flat = FlatParams(params, self._peers)
        self._dp_cut, self._dp_side, self._dp_forked = 0, None, False
        if self.fused_dp:
            # split exchange: levels >= dp_split_level of the hash table (and the parameters behind it) first
            split = int(os.environ.get("UMHS_DP_SPLIT_LEVEL", ops.SCATTER_SPLIT_LEVEL))
            table = model.field.mlp_base.encoder.hash_table
            L = int(model.field.mlp_base.encoder.scalings.shape[0])
            if 0 < split < L and ops.ENC_TILES and L == 16:
                t_lo, t_hi = self.flat.range_of(table)
                self._dp_cut = t_lo + (t_hi - t_lo) // L * split
                assert self._dp_cut % 4 == 0
                self.dp_split_level = split
            dev = params[0].device
            pad, pad_ptrs = self._peers.empty(256, torch.int32, dev)   # 8 channels x 8 ranks (+ slack) of uint32
            pad.zero_()
            self._dp_pads = torch.tensor(pad_ptrs, dtype=torch.int64, device=dev)
            self._dp_epochs = torch.zeros(8, dtype=torch.int32, device=dev)
            self._dp_pad_keep = pad
            torch.cuda.synchronize(dev)
            dist.barrier()  # every pad is zero before anyone signals
        self.clamp_range = self.flat.range_of(model.field.endmembers)
        self.step = start_step
        self.opt_steps = 0
        self.occ_update, self.freeze_binaries = occ_update, freeze_binaries
        # sync-free steps: sample buffers are sized to a capacity learnt from a first (synchronising) step and
        # re-checked every `capacity_check_every` steps; between checks a step issues no host sync at all
        self.sync_free, self.capacity_check_every = sync_free, capacity_check_every
        self.truncated_steps = 0
        # whole-step CUDA graph: forward + loss + backward (+ Adam when single-GPU) captured once and replayed; the
        # per-step scalars (lr, Adam bias corrections) travel through a 3-float device buffer
        self.cuda_graph = cuda_graph and sync_free
        self._graph = None
        self._graph_piped = None
        self._occ_graph = None  # CUDAGraph of the steady-state occupancy update; False: capture failed, stay eager
        self._hyper = torch.zeros(3, dtype=torch.float32, device=self.flat.data.device)
        # ring of pinned staging rows: the CPU may run several steps ahead of the asynchronous copies
        self._hyper_host = torch.zeros((256, 3), dtype=torch.float32).pin_memory() if self.flat.data.is_cuda else None
        ops.DIRECT_GRAD = True  # backward kernels accumulate into the flat gradient views
        self._inv_world = torch.full((1,), 1.0 / world_size, dtype=torch.float32, device=self.flat.data.device)
        if world_size > 1:  # identical replicas
            dist.broadcast(self.flat.data, src=0)

    def dp_shard(self, rank: Optional[int] = None, span: Optional[tuple] = None):
        """[begin, end) of the flat buffers (or of their slice ``span``) whose Adam state rank ``rank`` owns in the
        fused data-parallel step (csrc/optim.cu::umhs_dp_adam_step_range: ceil(len/4 / world) float4 groups per rank)."""
        r = self.rank if rank is None else rank
        lo, hi = (0, self.flat.numel) if span is None else span
        n4 = (hi - lo) // 4
        per = -(-n4 // self.world_size)
        return lo + 4 * min(per * r, n4), lo + 4 * min(per * (r + 1), n4)

    def dp_spans(self):
        """The slices of the flat buffers that are exchanged separately: everything at once, or -- when the hash-grid
        scatter is split (``dp_split_level``) -- [0, cut) and [cut, n) with cut = first element of that level's rows:
        the fine levels (and the MLP weights behind the table) travel while the coarse levels are still being scattered."""
        n = self.flat.numel
        cut = getattr(self, "_dp_cut", 0)
        return [(0, n)] if not cut else [(0, cut), (cut, n)]

    def dp_shards(self, rank: Optional[int] = None):
        """Every [begin, end) whose Adam state rank ``rank`` owns (one per span)."""
        return [self.dp_shard(rank, sp) for sp in self.dp_spans()]

    def full_optimizer_state(self):
        """(exp_avg, exp_avg_sq) over ALL parameters.  In the fused data-parallel step every rank keeps the moments
        of its own shard only (the other entries of its buffers stay zero): the shards are gathered here (one
        collective; checkpoint saving).  Single GPU / all-reduce mode: the local buffers are already complete."""
        f = self.flat
        if not self.fused_dp:
            return f.exp_avg, f.exp_avg_sq
        out = []
        for buf in (f.exp_avg, f.exp_avg_sq):
            full = torch.zeros_like(buf)
            for lo, hi in self.dp_shards():
                full[lo:hi] = buf[lo:hi]
            dist.all_reduce(full, op=dist.ReduceOp.SUM)  # shards are disjoint: the sum is the concatenation
            out.append(full)
        return out[0], out[1]

    def lr_at(self, step: int) -> float:
        """nerfstudio ExponentialDecayScheduler without warm-up: lr_init * (lr_final/lr_init)^(step/max_steps)."""
        t = min(max(step / self.max_steps, 0.0), 1.0)
        return self.lr_init * (self.lr_final / self.lr_init) ** t

    # ------------------------------------------------------------------ input pipeline (host -> device)
    def stage(self, ray_bundle: RayBundle, batch: Dict[str, Tensor]):
        """Start the host->device copy of a batch (pinned host tensors) on a copy stream and return a token for
        ``train_step``.  Three staging slots: batch i is being trained, batch i+1 is already on the device (its rays
        are marched while step i runs, see ``train_step(next_rays=...)``), batch i+2 is being copied."""
        dev = self.flat.data.device
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_bufs = [None, None, None]
            self._slot_free = [None, None, None]
            self._stage_slot = 0
            self._stage_seq = 0
        slot = self._stage_slot
        self._stage_slot = (self._stage_slot + 1) % 3
        self._stage_seq += 1
        src = {"origins": ray_bundle.origins, "directions": ray_bundle.directions, "image": batch["image"],
               "hs_image": batch["hs_image"]}
        bufs = self._stage_bufs[slot]
        if bufs is None or any(bufs[k].shape != v.shape for k, v in src.items()):
            bufs = self._stage_bufs[slot] = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in src.items()}
            # The caching allocator may hand out blocks that kernels still queued on the compute stream use (freed
            # in stream order there); the copy stream must not write them before that work has drained.
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
        if self._slot_free[slot] is not None:
            self._copy_stream.wait_event(self._slot_free[slot])  # the step that last read this slot is done with it
        with torch.cuda.stream(self._copy_stream):
            for k, v in src.items():
                bufs[k].copy_(v, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        return ("staged", slot, ready, self._stage_seq)

    def _unstage(self, token):
        _, slot, ready, _seq = token
        torch.cuda.current_stream(self.flat.data.device).wait_event(ready)
        b = self._stage_bufs[slot]
        return slot, RayBundle(b["origins"], b["directions"]), {"image": b["image"], "hs_image": b["hs_image"]}

    # ------------------------------------------------------------------ result read-back (device -> host)
    def read_async(self, value: Tensor):
        """Start the device->host copy of a step result (e.g. ``train_step(...)["loss"]``) and return a handle whose
        ``.result()`` gives the float.  ``float(value)`` / ``.item()`` would queue the copy on the compute stream
        BEHIND whatever step the host has already enqueued and block until that later step is done; the copy issued
        here runs on a side stream as soon as the step that produced ``value`` has finished, so the host can stay a
        step ahead of the GPU."""
        dev = self.flat.data.device
        if getattr(self, "_rb_stream", None) is None:
            self._rb_stream = torch.cuda.Stream(device=dev)
            self._rb_host = torch.zeros((64,), dtype=torch.float32).pin_memory()
            self._rb_slot = 0
        snap = value.detach().reshape(-1)[:1].to(torch.float32).clone()  # graph outputs are overwritten by the next replay
        slot = self._rb_slot
        self._rb_slot = (slot + 1) % self._rb_host.shape[0]
        self._rb_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._rb_stream):
            self._rb_host[slot:slot + 1].copy_(snap, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._rb_stream)
        snap.record_stream(self._rb_stream)
        host = self._rb_host

        class _Handle:
            def result(self_inner) -> float:
                done.synchronize()
                return float(host[slot])

        return _Handle()

    def _set_hyper(self) -> None:
        """lr and Adam bias corrections of this step -> device (pinned staging buffer, async copy on the stream)."""
        t = self.opt_steps
        row = self._hyper_host[t % self._hyper_host.shape[0]]
        row[0] = self.lr_at(self.step)
        row[1] = 1.0 - self.betas[0] ** t
        row[2] = (1.0 - self.betas[1] ** t) ** 0.5
        self._hyper.copy_(row, non_blocking=True)

    def _forward_backward(self, ray_bundle: RayBundle, batch: Dict[str, Tensor], premarched=None):
        if not (self.fused_dp and self._dp_cut):
            return self._forward_backward_impl(ray_bundle, batch, premarched)
        ops.BETWEEN_SCATTER_HALVES, keep = self._exchange_fine_levels, ops.SCATTER_SPLIT_LEVEL
        ops.SCATTER_SPLIT_LEVEL = self.dp_split_level  # the level the spans of this trainer are cut at
        try:
            return self._forward_backward_impl(ray_bundle, batch, premarched)
        finally:
            ops.BETWEEN_SCATTER_HALVES, ops.SCATTER_SPLIT_LEVEL = None, keep

    def _forward_backward_impl(self, ray_bundle: RayBundle, batch: Dict[str, Tensor], premarched=None):
        m = self.model
        outputs = m.get_outputs(m.collider(ray_bundle), per_band_views=False, premarched=premarched, lean=True)
        if self.fused_loss and outputs["spectral"].is_cuda:
            # UMHSModel.get_loss_dict and its backward as one kernel (csrc/loss.cu): the gradients of the scalar loss
            # with respect to the rendered outputs go straight into the compositing backward
            cfg = m.config
            spectral, rgb, acc = outputs["spectral"], outputs["rgb"], outputs["accumulation"]
            w_spec = 1.0 if cfg.method == "spectral" else 5.0  # hard-coded 5 of ref umhs_model.py:369 (quirk 4)
            w_rgb = 0.0 if cfg.method == "spectral" else float(cfg.rgb_loss_weight)
            # nerfstudio RGBRenderer.blend_background_for_loss_computation: only the RANDOM background is
            # added to the prediction; a named colour just blends an RGBA ground truth (UMHSModel._blend_background)
            blend_pred = cfg.background_color == "random"
            if blend_pred:
                bg = m._bg_override if m._bg_override is not None else torch.rand_like(rgb)
            else:
                rgba = batch["image"].shape[-1] == 4
                bg = torch.ones_like(rgb) if (cfg.background_color == "white" and rgba) else None
            dev = spectral.device
            losses, d_spec, d_rgb, d_acc = ops.train_loss(
                spectral.detach(), batch["hs_image"].to(dev), rgb.detach(), acc.detach().reshape(-1),
                batch["image"].to(dev), bg, w_spec, w_rgb, blend_pred=blend_pred)
            tensors, grads = [spectral], [d_spec]
            if d_rgb is not None:
                tensors += [rgb, acc]
                grads += [d_rgb, d_acc.unsqueeze(-1)]
            torch.autograd.backward(tensors, grads)  # kernels accumulate straight into the flat gradient views
            loss_dict = {"spectral_loss": losses[0]}
            if w_rgb != 0.0:
                loss_dict["rgb_loss"] = losses[1]
            loss_dict["loss"] = losses[0] + losses[1]
            # kept samples of the step: the compaction's device-side count (no reduction launch)
            loss_dict["num_samples"] = m.occupancy_grid.last_totals[1][0]
            return loss_dict
        loss_dict = m.get_loss_dict(outputs, batch)
        loss = sum(loss_dict.values())
        loss.backward()  # kernels accumulate straight into the flat gradient buffer's views
        loss_dict["loss"] = loss.detach()
        loss_dict["num_samples"] = outputs["num_samples_per_ray"].sum()
        return loss_dict

    def _dp_exchange(self, channel: int, span) -> None:
        """barrier (the span's gradients are complete on every rank) + reduce-scatter / Adam / all-gather of the span"""
        f = self.flat
        ops.dp_barrier(self._dp_pads, self.rank, self.world_size, channel, self._dp_epochs)
        ops.dp_adam_step(f.peer_grad, f.peer_data, f.exp_avg, f.exp_avg_sq, f.numel, self.rank, self.world_size,
                         self.betas[0], self.betas[1], self.eps, self._hyper, clamp_range=self.clamp_range,
                         mc_grad=f.mc_grad, mc_param=f.mc_data, span=span)

    def _exchange_fine_levels(self) -> None:
        """ops.BETWEEN_SCATTER_HALVES: the table gradient of the first half of the levels is complete -- exchange that
        span on a side stream while the main stream scatters the other half."""
        dev = self.flat.data.device
        if self._dp_side is None:
            self._dp_side = torch.cuda.Stream(device=dev)
        self._dp_side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._dp_side):
            self._dp_exchange(2, self.dp_spans()[1 if ops.SCATTER_FINE_FIRST else 0])
        self._dp_forked = True

    def _optimizer(self) -> None:
        if self.fused_dp:
            spans = self.dp_spans()
            forked = len(spans) == 2 and self._dp_forked
            early = 1 if ops.SCATTER_FINE_FIRST else 0
            if len(spans) == 2 and not forked:  # no split backward ran (plain autograd path): in line
                self._dp_exchange(2, spans[early])
            self._dp_exchange(0, spans[-1] if len(spans) == 1 else spans[1 - early])
            if forked:
                torch.cuda.current_stream(self.flat.data.device).wait_stream(self._dp_side)
                self._dp_forked = False
            ops.dp_barrier(self._dp_pads, self.rank, self.world_size, 1, self._dp_epochs)  # parameters landed
            self.flat.grad.zero_()
            return
        ops.adam_step_dev(self.flat.data, self.flat.grad, self.flat.exp_avg, self.flat.exp_avg_sq, self.betas[0],
                          self.betas[1], self.eps, self._hyper,
                          inv_scale=self._inv_world if self.world_size > 1 else None, clamp_range=self.clamp_range,
                          zero_grad=True)

    @staticmethod
    def _batch_key(rays):
        """Identity of a batch of rays (a staged token or a RayBundle of device tensors), to recognise at step i+1 the
        batch announced as ``next_rays`` at step i."""
        if rays is None:
            return None
        if isinstance(rays, tuple) and rays[0] == "staged":
            return ("staged", rays[1], rays[3])
        return ("rays", rays.origins.data_ptr(), rays.origins._version, rays.directions.data_ptr(),
                rays.directions._version, tuple(rays.origins.shape))

    def train_step(self, ray_bundle, batch: Optional[Dict[str, Tensor]] = None, next_rays=None) -> Dict[str, Tensor]:
        """One training step.  ``ray_bundle`` is a RayBundle (device or pinned-host tensors) with its ``batch``, or a
        token returned by ``stage``.  ``next_rays`` (optional, experimental): the RayBundle / token the NEXT call will
        train on; its occupancy march -- which depends on the rays and the occupancy grid only -- then runs beside this
        step's hash scatter and optimiser (a parallel branch of the step's CUDA graph) instead of in front of the next
        step.  Same results; see ``_capture`` for the measured (small) gain."""
        m = self.model
        m.train()
        slot = None
        key_cur = self._batch_key(ray_bundle)
        if isinstance(ray_bundle, tuple) and ray_bundle[0] == "staged":
            slot, ray_bundle, batch = self._unstage(ray_bundle)
        _dbg = os.environ.get("UMHS_TRAINER_DEBUG")
        _t0 = time.perf_counter() if _dbg else 0.0
        if self.occ_update:
            self._occupancy_update()
        if _dbg and self.step % 16 == 0:
            torch.cuda.synchronize()
            print(f"[trainer] step {self.step}: occupancy update {1e3 * (time.perf_counter() - _t0):.2f} ms",
                  file=sys.stderr)
        n_rays = int(ray_bundle.origins.shape[0])
        check = self.sync_free and (m.sample_capacity is None or self.opt_steps % self.capacity_check_every == 0)
        self.opt_steps += 1
        self._set_hyper()
        pipelined = (next_rays is not None and self.cuda_graph and m.sample_capacity is not None
                     and (self.world_size == 1 or self.fused_dp))
        if pipelined:
            loss_dict = self._graph_step(ray_bundle, batch, key_cur, next_rays,
                                         remarch=self.occ_update and self.step % 16 == 0)
        elif self.cuda_graph and m.sample_capacity is not None:
            loss_dict = self._graph_step(ray_bundle, batch)
        else:
            dev = self.flat.data.device
            if ray_bundle.origins.device != dev:  # host batches (e2e): eager path copies them in first
                ray_bundle = RayBundle(ray_bundle.origins.to(dev, non_blocking=True),
                                       ray_bundle.directions.to(dev, non_blocking=True))
                batch = {k: v.to(dev, non_blocking=True) for k, v in batch.items()}
            loss_dict = self._forward_backward(ray_bundle, batch)
            if not self.fused_dp:
                average_gradients(self.flat.grad, self.world_size)
            self._optimizer()
        if slot is not None:  # the staging slot may be overwritten once this step has consumed it
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.data.device))
            self._slot_free[slot] = ev
        if check:
            self._update_capacity(n_rays)
        self.step += 1
        return loss_dict

    # ------------------------------------------------------------------ occupancy-grid update (ref umhs_model.py:549-554)
    def _occupancy_update(self) -> None:
        """Every 16 steps.  Past the grid's warm-up the update has a fixed shape and no host sync (occ_grid.py), so it
        is captured once as its own CUDA graph and replayed: one launch instead of ~150 from the host."""
        m = self.model
        if self.step % 16 != 0:
            return
        use_graph = self.cuda_graph and self.flat.data.is_cuda and self.step >= 256 and self._occ_graph is not False
        if not use_graph:
            m.update_occupancy_grid(self.step, freeze_binaries=self.freeze_binaries)
            return
        if self._occ_graph is None:
            g = torch.cuda.CUDAGraph()
            try:
                m.field._aabb6()  # host copy of the field's aabb: read back now, not inside the capture
                torch.cuda.synchronize(self.flat.data.device)
                with torch.cuda.graph(g):
                    m.update_occupancy_grid(self.step, freeze_binaries=self.freeze_binaries)
                self._occ_graph = g
            except Exception as e:  # noqa: BLE001 - an op that cannot be captured: stay eager
                import traceback

                print(f"[umhs_b200] occupancy update not captured ({type(e).__name__}: {e}); running it eagerly\n"
                      + "".join(traceback.format_tb(e.__traceback__)[-4:]), file=sys.stderr)
                self._occ_graph = False
                m.update_occupancy_grid(self.step, freeze_binaries=self.freeze_binaries)
                return
        m.step = self.step
        self._occ_graph.replay()

    # ------------------------------------------------------------------ whole-step CUDA graph
    def _graph_step(self, ray_bundle: RayBundle, batch: Dict[str, Tensor], key_cur=None, next_rays=None,
                    remarch: bool = False) -> Dict[str, Tensor]:
        piped = next_rays is not None
        name = "_graph_piped" if piped else "_graph"
        g = getattr(self, name, None)
        if g is None or g["capacity"] != self.model.sample_capacity or g["n_rays"] != int(ray_bundle.origins.shape[0]):
            self._capture(ray_bundle, batch, piped)
            g = getattr(self, name)
        g["origins"].copy_(ray_bundle.origins, non_blocking=True)
        g["directions"].copy_(ray_bundle.directions, non_blocking=True)
        g["image"].copy_(batch["image"], non_blocking=True)
        g["hs_image"].copy_(batch["hs_image"], non_blocking=True)
        if piped:
            dev = self.flat.data.device
            nxt = next_rays
            if isinstance(nxt, tuple) and nxt[0] == "staged":  # rays of a staged batch: its copy must have landed
                torch.cuda.current_stream(dev).wait_event(nxt[2])
                b = self._stage_bufs[nxt[1]]
                nxt = RayBundle(b["origins"], b["directions"])
            g["next_origins"].copy_(nxt.origins, non_blocking=True)
            g["next_directions"].copy_(nxt.directions, non_blocking=True)
            if remarch or key_cur is None or g["key_next"] != key_cur:
                # the candidates in flight are not this batch's (first step, caller changed its mind) or predate an
                # occupancy update: march this batch now, in line
                self.model.train()
                pm = self.model.premarch(self.model.collider(RayBundle(g["origins"], g["directions"])))
                for dst, src in zip(g["pm_next"], pm):
                    dst.copy_(src)
            g["key_next"] = self._batch_key(next_rays)
        elif getattr(self, "_graph_piped", None) is not None:
            self._graph_piped["key_next"] = None  # a plain step in between: the announced batch was not consumed
        g["fwd_bwd"].replay()
        if "opt" in g:
            average_gradients(self.flat.grad, self.world_size)
            g["opt"].replay()
        return g["out"]

    def _capture(self, ray_bundle: RayBundle, batch: Dict[str, Tensor], piped: bool = False) -> None:
        dev = self.flat.data.device
        if os.environ.get("UMHS_TRAINER_DEBUG"):
            print(f"[trainer] step {self.step}: capturing the {'pipelined ' if piped else ''}step graph (capacity "
                  f"{self.model.sample_capacity})", file=sys.stderr)
        st = {"capacity": self.model.sample_capacity, "n_rays": int(ray_bundle.origins.shape[0]),
              "origins": ray_bundle.origins.to(dev).clone(), "directions": ray_bundle.directions.to(dev).clone(),
              "image": batch["image"].to(dev).clone(), "hs_image": batch["hs_image"].to(dev).clone()}
        rb = RayBundle(st["origins"], st["directions"])
        bt = {"image": st["image"], "hs_image": st["hs_image"]}
        # a capture must not change the training state: snapshot, warm up on a side stream, capture, restore
        snap = [t.clone() for t in (self.flat.data, self.flat.grad, self.flat.exp_avg, self.flat.exp_avg_sq)]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._forward_backward(rb, bt)
                self.flat.grad.zero_()
            if piped:
                pm0 = self.model.premarch(self.model.collider(rb))
                st["next_origins"], st["next_directions"] = st["origins"].clone(), st["directions"].clone()
                st["pm_next"] = [t.clone() for t in pm0]   # candidates of the batch the next replay trains on
                st["pm_cur"] = [t.clone() for t in pm0]    # candidates of the batch this replay trains on
                st["key_next"] = None
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            if piped:
                for dst, src in zip(st["pm_cur"], st["pm_next"]):
                    dst.copy_(src)
                # Parallel branch: march the NEXT batch while this one computes.  It is forked right after the MLP
                # backward, beside the hash scatter and the optimiser: those kernels use no shared memory, so the march
                # CTAs can share their SMs, whereas an SM holding a march CTA cannot take a CTA of the persistent MLP
                # kernels (214-222 KB of shared memory; measured: forked at the start of the step the branch costs its
                # full serial time, 0.25 ms).  Measured gain of this placement on 1 x B200: 0.02-0.04 ms of a 1.8 ms
                # step -- the march is latency-bound and slows down under contention about as much as it overlaps -- so
                # bench.py does not use it by default (UMHS_BENCH_PIPELINE=1).
                main = torch.cuda.current_stream(dev)
                forked = []

                def fork_march():
                    cur = torch.cuda.current_stream(dev)
                    branch = torch.cuda.Stream(device=dev, priority=-1)  # its few CTAs go first when slots free up
                    branch.wait_stream(cur)
                    with torch.cuda.stream(branch):
                        pm = self.model.premarch(self.model.collider(RayBundle(st["next_origins"],
                                                                               st["next_directions"])))
                        for dst, src in zip(st["pm_next"], pm):
                            dst.copy_(src)
                    forked.append(branch)

                early = os.environ.get("UMHS_PIPE_FORK", "after") == "before"
                if early:
                    ops.BEFORE_MLP_BACKWARD = fork_march
                else:
                    ops.AFTER_MLP_BACKWARD = fork_march
                try:
                    out = self._forward_backward(rb, bt, premarched=tuple(st["pm_cur"]))
                finally:
                    ops.AFTER_MLP_BACKWARD = ops.BEFORE_MLP_BACKWARD = None
                if not forked:  # no field backward ran (cannot happen in training): march in line
                    fork_march()
            else:
                out = self._forward_backward(rb, bt)
            if self.world_size == 1 or self.fused_dp:
                self._optimizer()  # fused_dp: barrier + exchange/Adam kernel + barrier, all inside the graph
            if piped:
                for branch in forked:
                    main.wait_stream(branch)
        st["fwd_bwd"], st["out"] = g1, out
        if self.world_size > 1 and not self.fused_dp:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                self._optimizer()
            st["opt"] = g2
        for dst, src in zip((self.flat.data, self.flat.grad, self.flat.exp_avg, self.flat.exp_avg_sq), snap):
            dst.copy_(src)
        if piped:
            self._graph_piped = st
        else:
            self._graph = st

    def _update_capacity(self, n_rays: int) -> None:
        """One host read of the last step's unclamped candidate count; grows the capacity with 30 % headroom."""
        march_total, _ = self.model.occupancy_grid.last_totals
        raw = int(march_total[1].item())
        cap = self.model.sample_capacity
        if cap is not None and raw > cap:
            self.truncated_steps += 1
        if cap is None or raw > cap or raw * 2 < cap:
            self.model.sample_capacity = int(raw * 1.3) + n_rays

    @torch.no_grad()
    def render_tiles(self, camera_ray_bundle: RayBundle, tile: int = 32) -> Dict[str, Tensor]:
        """Evaluation rendering of an [H, W] ray bundle: tiles of tile x tile pixels dealt round-robin to the ranks,
        no collective.  Returns this rank's pixels: {"index": flat pixel ids, <output name>: [n, C]}."""
        m = self.model
        m.eval()
        H, W = camera_ray_bundle.origins.shape[:2]
        dev = m.device
        key = (H, W, tile, self.rank, self.world_size, str(dev))
        cache = self.__dict__.setdefault("_tile_index_cache", {})
        if key not in cache:  # a few hundred tiny launches per image otherwise
            cache[key] = tile_pixel_indices(H, W, tile, self.rank, self.world_size, dev)
        index = cache[key]
        o = camera_ray_bundle.origins.reshape(-1, 3).to(dev)[index]
        d = camera_ray_bundle.directions.reshape(-1, 3).to(dev)[index]
        out: Dict[str, List[Tensor]] = {}
        step = int(m.config.eval_tile_rays)
        chunk = int(m.config.eval_num_rays_per_chunk)  # depth is clipped per reference forward call (512 row-major rays)
        for s in range(0, index.numel(), step):
            rb = RayBundle(origins=o[s:s + step].contiguous(), directions=d[s:s + step].contiguous())
            res = m.get_outputs(m.collider(rb), per_band_views=False,
                                depth_chunks=(chunk, index[s:s + step].contiguous(), H * W))
            for k, v in res.items():
                out.setdefault(k, []).append(v.reshape(v.shape[0], -1))
        result = {k: torch.cat(v) for k, v in out.items()}
        result["index"] = index
        return result
