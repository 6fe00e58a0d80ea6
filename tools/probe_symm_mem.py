"""Probe (multi-GPU box): does torch's symmetric memory rendezvous work here, is there an NVSwitch multicast address,
and what do NCCL / torch's own symm_mem all-reduces cost for the head's gradient bucket?  torchrun --nproc-per-node N tools/probe_symm_mem.py"""
import os
import sys

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3


for mb in (18, 26):
    n = mb * 1024 * 1024 // 4
    x = torch.randn(n, device=dev)
    us = timeit(lambda: dist.all_reduce(x))
    if rank == 0:
        print(f"nccl all_reduce {mb} MB fp32, {world} ranks: {us:.1f} us (eager launches)", flush=True)
    g = torch.cuda.CUDAGraph()
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            dist.all_reduce(x)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            dist.all_reduce(x)
        us = timeit(g.replay)
        if rank == 0:
            print(f"nccl all_reduce {mb} MB inside a CUDA graph: {us:.1f} us", flush=True)
    except Exception as exc:
        if rank == 0:
            print("nccl graph capture failed:", repr(exc), flush=True)

try:
    import torch.distributed._symmetric_memory as symm_mem
    n = 18 * 1024 * 1024 // 4
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    if rank == 0:
        print("symm_mem rendezvous ok: multicast_ptr", hex(hdl.multicast_ptr), "buffer_ptrs", len(hdl.buffer_ptrs), "signal pads", len(hdl.signal_pad_ptrs),
              "attrs", [a for a in dir(hdl) if not a.startswith("_")], flush=True)
    t.normal_()
    gname = dist.group.WORLD.group_name
    for name, fn in (("multimem_all_reduce_", lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname)),
                     ("two_shot_all_reduce_", lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", gname)),
                     ("one_shot_all_reduce", lambda: torch.ops.symm_mem.one_shot_all_reduce(t, "sum", gname))):
        try:
            us = timeit(fn)
            if rank == 0:
                print(f"symm_mem.{name} 18 MB: {us:.1f} us", flush=True)
        except Exception as exc:
            if rank == 0:
                print(f"symm_mem.{name} failed: {exc!r}"[:300], flush=True)
    us = timeit(lambda: hdl.barrier(channel=0))
    if rank == 0:
        print(f"symm_mem barrier: {us:.1f} us", flush=True)
except Exception as exc:
    if rank == 0:
        print("symm_mem unavailable:", repr(exc)[:500], flush=True)
dist.destroy_process_group()
