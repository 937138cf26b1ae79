"""2-GPU probe: torch symmetric memory (peer-mapped buffers + signal pads) availability and P2P copy bandwidth.
torchrun --nproc-per-node 2 tools/p2p_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

n = 16 * 1024 * 1024
try:
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok", type(hdl).__name__, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "pads",
          [hex(p) for p in hdl.signal_pad_ptrs], "pad size", hdl.signal_pad_size, flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    print(rank, "peer value", float(peer[123]), flush=True)
    dst = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(3):
        dst.copy_(peer)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        dst.copy_(peer)
    b.record()
    torch.cuda.synchronize()
    print(rank, "P2P read GB/s", 10 * n * 4 / (a.elapsed_time(b) * 1e-3) / 1e9, flush=True)
    # barrier under graph capture
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        hdl.barrier()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    try:
        with torch.cuda.graph(g):
            hdl.barrier()
            dst.copy_(peer)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        print(rank, "barrier captured in a CUDA graph: ok", flush=True)
    except Exception as e:  # noqa: BLE001
        print(rank, "graph capture of barrier failed:", repr(e)[:300], flush=True)
    # all-reduce baseline 64 MiB
    x = torch.zeros(n, dtype=torch.float32, device=dev)
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        dist.all_reduce(x)
    b.record()
    torch.cuda.synchronize()
    print(rank, "NCCL all_reduce 64 MiB ms", a.elapsed_time(b) / 10, flush=True)
except Exception as e:  # noqa: BLE001
    import traceback
    traceback.print_exc()
    print(rank, "symmetric memory failed:", repr(e)[:500], flush=True)
dist.barrier()
dist.destroy_process_group()
