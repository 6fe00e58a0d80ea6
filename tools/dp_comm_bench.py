"""The gradient all-reduce of the data-parallel head, timed ALONE on N GPUs (torchrun --nproc-per-node N tools/dp_comm_bench.py):
NCCL all_reduce over the live span vs the hand-written one-kernel all-reduce (csrc/dp_comm.cuh) over the live ranges of a
symmetric-memory bucket - NVSwitch multimem and peer loads, with torch's two barrier kernels and with the barriers fused
into the kernel.  CUDA events on each rank, max over ranks, eager launches and one CUDA graph of 20 calls."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200")]
import torch
import torch.distributed as dist
import fusion_b200 as fb
from fusion_b200 import _lib

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
desc = fb.make_desc("crossattention", 4096, 2048, 85, 512, 512, 8, 6)
total, _ = _lib.grad_layout(desc)
ranges = _lib.grad_live_ranges(desc)
live = sum(e - b for b, e in ranges)
span = (min(b for b, _ in ranges), max(e for _, e in ranges))


def timeit(fn, n=40, warm=8):
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


def graphed(fn, reps=20):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    return lambda: g.replay(), reps


out = []
flat = torch.randn(total, device=dev)
out.append(("nccl all_reduce, live span (%.1f MB incl. W_q/W_k zeros)" % ((span[1] - span[0]) * 4 / 1e6), timeit(lambda: dist.all_reduce(flat[span[0]:span[1]]))))
try:
    fn, reps = graphed(lambda: dist.all_reduce(flat[span[0]:span[1]]))
    out.append(("  ... in a CUDA graph", timeit(fn, n=6, warm=2) / reps))
except Exception as exc:
    out.append((f"  ... graph capture failed: {exc!r}"[:120], float("nan")))
for mode in ("multimem", "peer"):
    try:
        bucket = fb.dp.SymmetricGradBucket(total, dev, mode=mode)
    except Exception as exc:
        out.append((f"{mode}: unavailable ({exc!r})"[:160], float("nan")))
        continue
    bucket.tensor.normal_()
    for fused in (False, True):
        bucket.fused_barriers = fused
        tag = f"fb200_dp_allreduce {mode}, live ranges ({live * 4 / 1e6:.1f} MB), " + ("barriers inside the kernel" if fused else "two torch barrier kernels")
        out.append((tag, timeit(lambda: bucket.all_reduce(ranges))))
        try:
            fn, reps = graphed(lambda: bucket.all_reduce(ranges))
            out.append(("  ... in a CUDA graph", timeit(fn, n=6, warm=2) / reps))
        except Exception as exc:
            out.append((f"  ... graph capture failed: {exc!r}"[:120], float("nan")))
    # correctness of the last call chain: all ranks must hold identical sums
    bucket.tensor.copy_(torch.arange(total, device=dev, dtype=torch.float32) % 7 + rank)
    torch.cuda.synchronize(); dist.barrier()
    bucket.all_reduce(ranges)
    torch.cuda.synchronize()
    b, e = ranges[0]
    expect = (torch.arange(b, e, device=dev, dtype=torch.float32) % 7) * world + sum(range(world))
    ok = torch.equal(bucket.tensor[b:e], expect)
    out.append((f"  {mode}: result check {'OK' if ok else 'WRONG'}", 0.0))
if rank == 0:
    print(f"== gradient all-reduce alone, {world} x B200, cfg2 head ({total * 4 / 1e6:.1f} MB flat buffer, {live * 4 / 1e6:.1f} MB live in {len(ranges)} ranges)")
    for k, v in out:
        print(f"{v:9.1f} us  {k}")
    sys.stdout.flush()
torch.cuda.synchronize(); dist.barrier()
os._exit(0)
