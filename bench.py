#!/usr/bin/env python
"""bench.py - fusion-head train samples/sec (forward + weighted CE + backward) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): ResNet-50 feature width F=2048, one-hot
PAD-UFES-20 metadata V=85, fusion "crossattention" (8 heads, COMMON_DIM=512), 6 classes,
fp32, random-init weights under torch.manual_seed(1234), synthetic N(0,1) inputs under
torch.manual_seed(4321+rank).  One step = one pass of the hot path over one batch of
--batch samples per GPU (weak scaling: per-GPU batch fixed; gradients all-reduced by NCCL).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: mechanism, F, V, C, T, text_mode, dtype
    "cfg1": ("concatenation", 512, 85, 6, 512, 0, "fp32"),
    "cfg2": ("crossattention", 2048, 85, 6, 512, 0, "fp32"),
    "cfg3a": ("metablock", 1664, 13, 8, 512, 0, "fp32"),
    "cfg3b": ("weighted", 1664, 13, 8, 512, 0, "fp32"),
    "cfg4a": ("gfcam", 768, 0, 2, 85, 1, "bf16"),
    "cfg4b": ("att-intramodal+residual+cross-attention-metadados", 768, 0, 2, 85, 1, "bf16"),
    "cfg5": ("att-intramodal+residual+cross-attention-metadados", 1024, 85, 6, 512, 0, "bf16"),
}
WORKLOAD_DESC = {
    "cfg2": "ResNet-50 (F=2048) + one-hot metadata (V=85) + crossattention (8 heads, COMMON_DIM=512), PAD-UFES-20 synthetic (6 classes)",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=4096, help="samples per GPU per step")
    ap.add_argument("--dtype", default=None, choices=[None, "fp32", "bf16"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --batch rows per GPU; strong: --batch rows in total, split over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying one CUDA graph per step")
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc"], help="GEMM engine policy (auto: tcgen05 from B > 32 in fp32, always in bf16)")
    ap.add_argument("--sweep", default="32,256,1024", help="comma-separated extra per-GPU batch sizes reported under 'sweep' ('' disables)")
    ap.add_argument("--no-extras", action="store_true", help="skip the token-attention and backbone end-to-end extras (profiling runs)")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the GPU-eager reference (torch.nn on cuda) and the eager drop-in route")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                smax = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons)}


def physical_cores():
    """Physical cores of the box (SURVEY 8d asks for physical, not logical, cores in `cpu_baseline.cores`)."""
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
        if n:
            return int(n)
    except Exception:
        pass
    return os.cpu_count() or 1


# ----------------------------------------------------------------------------- reference arm
def reference_arm(args, wl):
    """The reference's own CPU implementation of the path (torch.nn on the host cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle.torch_port import time_cpu_train_step
    mech, F, V, Cn, T, tm, dtype = wl
    threads = physical_cores()
    B = args.batch
    r = time_cpu_train_step(mech, F, V, Cn, B, T=T, one_hot=(tm == 0), steps=max(args.steps, 2), warmup=max(min(args.warmup, 3), 1),
                            threads=threads, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "fusion-head train samples/sec (fwd+bwd)", "value": r["samples_per_s"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "per_gpu_batch": B, "mechanism": mech,
                   "note": "reference torch.nn CPU path (oracle/torch_port.py: same modules, same dead work as the reference forward) on the host cores"},
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"], "kind": "port",
                         "sample": f"{r['steps']} train steps of batch {B} (median)"},
        "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- b200 arm
_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was re-routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # C-level writers to fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    wl = list(WORKLOADS[args.workload])
    if args.dtype:
        wl[6] = args.dtype
    if args.impl == "reference":
        return reference_arm(args, wl)

    import torch
    import torch.distributed as dist
    import fusion_b200 as fb
    from fusion_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = None
    if world > 1:
        numa_cpus = fb.dp.bind_to_gpu_numa_node(local)           # before any pinned allocation: host staging buffers become node-local
        dist.init_process_group("nccl", device_id=dev)
    mech, F, V, Cn, T, tm, dtype = wl
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    if args.scaling == "strong":
        if B % world:
            raise SystemExit(f"--scaling strong: --batch {B} must be a multiple of the {world} GPUs")
        B //= world

    def build(Bsz):
        torch.manual_seed(1234)
        model = fb.MultimodalModel(Cn, 8, dev, f"identity:{F}", "one-hot-encoder" if tm == 0 else "tab-transformer",
                                   vocab_size=V if V else 91, text_encoder_dim_output=T, attention_mecanism=mech, compute_dtype=dtype,
                                   engine_flags={"auto": 0, "simt": 4, "tc": 8}[args.engine]).to(dev)
        model.train()
        return model

    def make_pool(Bsz, nbatch, pinned=False):
        g = torch.Generator().manual_seed(4321 + rank)
        xs, ts, ys = [], [], []
        for _ in range(nbatch):
            x = torch.randn(Bsz, F, generator=g); t = torch.randn(Bsz, V if tm == 0 else T, generator=g)
            y = torch.randint(0, Cn, (Bsz,), generator=g)
            if pinned:
                xs.append(x.pin_memory()); ts.append(t.pin_memory()); ys.append(y.pin_memory())
            else:
                xs.append(x.to(dev)); ts.append(t.to(dev)); ys.append(y.to(dev))
        return xs, ts, ys

    def class_weights(ys):
        y = torch.cat([v.cpu() for v in ys])
        counts = torch.bincount(y, minlength=Cn).clamp_min(1).float()
        return (y.numel() / (Cn * counts)).to(dev)

    # ---- data-parallel gradient all-reduce: the hand-written one-kernel NVSwitch (multimem) / peer all-reduce over a
    #      symmetric-memory gradient bucket (csrc/dp_comm.cuh), captured into the step's CUDA graph; NCCL when symmetric
    #      memory is unavailable or FB200_DP_COMM=nccl.  One bucket serves every batch size (sized for the largest layout).
    comm = {"kind": "none"}
    if world > 1:
        want = os.environ.get("FB200_DP_COMM", "auto")
        comm = {"kind": "nccl"}
        if want != "nccl":
            try:
                total_max, _ = _lib.grad_layout(fb.make_desc(mech, B, F, V, T, 512, 8, Cn, text_mode=tm, dtype=dtype))
                bucket = fb.dp.SymmetricGradBucket(total_max, dev, mode="auto" if want == "auto" else want)
                comm = {"kind": bucket.mode, "bucket": bucket}
            except Exception as exc:
                if want != "auto":
                    raise
                comm = {"kind": "nccl", "symm_error": repr(exc)[:200]}

    def run_config(Bsz, steps, warm, sample_clocks=False):
        model = build(Bsz)
        in_bytes = Bsz * (F + (V if tm == 0 else T)) * 4
        nb = max(2, min(16, int(160e6 // in_bytes) + 1))        # pool > 126 MB L2 where affordable
        xs, ts, ys = make_pool(Bsz, nb)
        cw = class_weights(ys)

        use_graph = not args.no_graph
        live_ranges = _lib.grad_live_ranges(fb.make_desc(mech, Bsz, F, V, T, 512, 8, Cn, text_mode=tm, dtype=dtype))
        denoms = [None] * nb
        if world > 1:                                            # global weighted-CE denominator, static buffers
            denoms = [torch.zeros(1, device=dev) for _ in range(nb)]
        # data parallel: the gradient all-reduce (optionally in two buckets, the first one under the second half of the
        # weight-gradient launch: fb200_head_train_step_dp records `mids[j]` when the first bucket is final)
        mids = [None] * nb
        bar = None
        if world > 1:
            # Overlapping bucket 1 with the second half of the weight gradients (fb200_head_train_step_dp) measured SLOWER
            # here (2 GPUs: 0.905 vs 0.871 ms per step): NCCL's CTAs push the 136-tile launch into a second wave.
            # Default: one in-place all-reduce of the gradient span behind the step; FB200_DP_OVERLAP=1 for the overlapped path.
            if os.environ.get("FB200_DP_OVERLAP", "0") == "1":
                mids = [torch.cuda.Event() for _ in range(nb)]
                for ev in mids:
                    ev.record()
            bar = fb.dp.BucketedAllReduce(dev, bucket=comm.get("bucket"))
        split = _lib.dp_bucket_split(fb.make_desc(mech, Bsz, F, V, T, 512, 8, Cn, text_mode=tm, dtype=dtype, train=True,
                                                  flags={"auto": 0, "simt": 4, "tc": 8}[args.engine])) if world > 1 else 0
        graphs = []
        bucket = comm.get("bucket")
        flat_out = bucket.tensor if bucket is not None else None
        span = (min(b for b, _ in live_ranges), max(e for _, e in live_ranges))
        in_graph = world > 1 and use_graph and os.environ.get("FB200_DP_OVERLAP", "0") != "1" and os.environ.get("FB200_DP_IN_GRAPH", "1") == "1"

        def reduce_grads(flat):
            if bucket is not None:
                bucket.all_reduce(live_ranges)                      # barrier + ONE kernel over the live slices + barrier
            else:
                dist.all_reduce(flat[span[0]:span[1]])              # NCCL, one in-place collective over the contiguous live span
        if use_graph:
            for j in range(nb):
                graphs.append(fb.GraphedTrainStep(model, xs[j], ts[j], ys[j], cw, denom=denoms[j], mid_event=mids[j], flat_out=flat_out,
                                                  after_step=reduce_grads if in_graph else None))

        # The global weighted-CE denominator of a batch only needs its labels, which the loader has one step ahead:
        # it is all-reduced on a side stream while the previous step computes (dp.DenominatorPrefetcher).
        pref = fb.dp.DenominatorPrefetcher(dev, nb) if world > 1 else None

        def step(i):
            j = i % nb
            if world > 1:
                if i == 0:
                    pref.issue(j, ys[j], cw, denoms[j])
                pref.wait(j)
                jn = (i + 1) % nb
                pref.issue(jn, ys[jn], cw, denoms[jn])              # next batch's denominator: its tiny all-reduce rides under this step
            if use_graph:
                loss = graphs[j].run()
                flat = graphs[j].flat_grad
            else:
                loss, _ = model.forward_loss(xs[j], ts[j], ys[j], cw, denom=denoms[j], mid_event=mids[j], flat_out=flat_out)
                flat = model.flat_grad
            if world > 1:
                if in_graph:
                    pref.mark_consumed(j)                           # the collective was captured with the step: one graph launch did both
                elif os.environ.get("FB200_DP_OVERLAP", "0") != "1":
                    pref.mark_consumed(j)
                    reduce_grads(flat)
                else:
                    bar.start(flat, live_ranges, split, mids[j])   # bucket 1 on the communication stream, under the tail of the step
                    pref.mark_consumed(j)
                    bar.finish(flat)                                # bucket 2 (SUM: the global denominator already averages; only live slices travel)
            return loss

        for i in range(warm):
            step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start(); time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.time()
        e0.record()
        for i in range(steps):
            loss = step(warm + i)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        clocks = sampler.stop(t0, t1) if sampler else None
        return model, ms.item() / steps, clocks, float(loss), (nb, in_bytes)

    # ---- headline: inputs resident in HBM
    model, ms_step, clocks, last_loss, (nb, in_bytes) = run_config(B, K, W, sample_clocks=True)
    value = B * world / (ms_step * 1e-3)

    # ---- end to end through the public API with HOST buffers (H2D inside the timed region, loss read back)
    def run_e2e(Bsz, steps, warm):
        m2 = build(Bsz)
        nbh = 4
        hx, ht, hy = make_pool(Bsz, nbh, pinned=False)
        hx, ht, hy = [v.cpu() for v in hx], [v.cpu() for v in ht], [v.cpu() for v in hy]
        cw = class_weights(hy)
        copy_stream = torch.cuda.Stream(device=dev)
        Wt = ht[0].shape[1]
        nx, nt, ny = Bsz * F * 4, Bsz * Wt * 4, Bsz * 8
        # ONE packed pinned staging buffer per batch (features | metadata | labels) and ONE cudaMemcpyAsync per step into a
        # device buffer the step reads through typed views
        staged = []
        for j in range(nbh):
            hb = torch.empty(nx + nt + ny, dtype=torch.uint8).pin_memory()
            hb[:nx].view(torch.float32).copy_(hx[j].reshape(-1)); hb[nx:nx + nt].view(torch.float32).copy_(ht[j].reshape(-1))
            hb[nx + nt:].view(torch.int64).copy_(hy[j])
            staged.append(hb)
        slots = []
        for _ in range(2):
            db = torch.zeros(nx + nt + ny, dtype=torch.uint8, device=dev)
            slots.append(dict(buf=db, x=db[:nx].view(torch.float32).view(Bsz, F), t=db[nx:nx + nt].view(torch.float32).view(Bsz, Wt),
                              y=db[nx + nt:].view(torch.int64), ready=torch.cuda.Event(), free=torch.cuda.Event()))
        host_loss = torch.empty(warm + 3 * steps, dtype=torch.float32).pin_memory()

        def prefetch(i):
            s = slots[i % 2]; j = i % nbh
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(s["free"])
                s["buf"].copy_(staged[j], non_blocking=True)
                s["ready"].record(copy_stream)

        use_graph = not args.no_graph
        live_ranges2 = _lib.grad_live_ranges(fb.make_desc(mech, Bsz, F, V, T, 512, 8, Cn, text_mode=tm, dtype=dtype))
        bar2 = fb.dp.BucketedAllReduce(dev, bucket=comm.get("bucket")) if world > 1 else None
        split2 = _lib.dp_bucket_split(fb.make_desc(mech, Bsz, F, V, T, 512, 8, Cn, text_mode=tm, dtype=dtype, train=True,
                                                   flags={"auto": 0, "simt": 4, "tc": 8}[args.engine])) if world > 1 else 0
        bucket = comm.get("bucket")
        flat_out2 = bucket.tensor if bucket is not None else None
        span2 = (min(b for b, _ in live_ranges2), max(e for _, e in live_ranges2))

        def reduce2(flat):
            if bucket is not None:
                bucket.all_reduce(live_ranges2)
            else:
                dist.all_reduce(flat[span2[0]:span2[1]])
        in_graph2 = world > 1 and use_graph and os.environ.get("FB200_DP_OVERLAP", "0") != "1" and os.environ.get("FB200_DP_IN_GRAPH", "1") == "1"
        for s in slots:
            s["denom"] = torch.zeros(1, device=dev) if world > 1 else None
            s["mid"] = None
            if world > 1 and os.environ.get("FB200_DP_OVERLAP", "0") == "1":
                s["mid"] = torch.cuda.Event(); s["mid"].record()
            s["graph"] = fb.GraphedTrainStep(m2, s["x"], s["t"], s["y"], cw, denom=s["denom"], mid_event=s["mid"], flat_out=flat_out2,
                                             after_step=reduce2 if in_graph2 else None) if use_graph else None

        def run(n, base):
            cur = torch.cuda.current_stream()
            prefetch(base)
            for i in range(base, base + n):
                if i + 1 < base + n:
                    prefetch(i + 1)
                s = slots[i % 2]
                cur.wait_event(s["ready"])
                if world > 1:                                   # labels arrive with the H2D copy: the denominator cannot be earlier than that
                    s["denom"].copy_(cw[s["y"]].sum().reshape(1)); dist.all_reduce(s["denom"])
                if use_graph:
                    loss = s["graph"].run(); flat = s["graph"].flat_grad
                else:
                    loss, _ = m2.forward_loss(s["x"], s["t"], s["y"], cw, denom=s["denom"], mid_event=s["mid"], flat_out=flat_out2); flat = m2.flat_grad
                if world > 1 and not in_graph2:
                    if os.environ.get("FB200_DP_OVERLAP", "0") != "1":
                        reduce2(flat)
                    else:
                        bar2.start(flat, live_ranges2, split2, s["mid"])
                        bar2.finish(flat)
                host_loss[i:i + 1].copy_(loss.reshape(1), non_blocking=True)       # D2H read of the step's result
                s["free"].record(cur)
        for s in slots:
            s["free"].record(torch.cuda.current_stream())
        run(warm, 0)
        # K steps are a few tens of milliseconds of wall clock with PCIe in the loop: three timed repetitions of exactly
        # `steps` steps each, the median is reported (max over ranks per repetition)
        reps = []
        for rep in range(3):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            run(steps, warm + rep * steps)
            torch.cuda.synchronize()
            dtt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dtt, op=dist.ReduceOp.MAX)
            reps.append(dtt.item())
        h2d = Bsz * (F + ht[0].shape[1]) * 4 + Bsz * 8
        return Bsz * world * steps / statistics.median(reps), h2d, 4

    e2e_value, h2d, d2h = run_e2e(B, K, W)

    # ---- roofline of the dominant kernel (the GEMM), replayed live through the C ABI with CUDA events
    desc = fb.make_desc(mech, B, F, V, T, 512, 8, Cn, text_mode=tm, dtype=dtype, train=True, flags={"auto": 0, "simt": 4, "tc": 8}[args.engine])
    flops, nbytes, plive = _lib.algorithmic_work(desc)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roof = dominant_kernel_roofline(torch, _lib, desc, dev, peaks, dtype, model)
    t_roof_ms = max(flops / (roof["step_peak_tflops"] * 1e12), nbytes / (roof["hbm_gbs"] * 1e9)) * 1e3
    roof["step"] = {"algorithmic_flops": flops, "algorithmic_bytes": nbytes, "t_roof_ms": t_roof_ms, "frac": t_roof_ms / ms_step}

    fwd_l, bwd_l = _lib.launch_count(desc)
    launches = K * (fwd_l + bwd_l + 2)

    sweep = {}
    for bs in [int(v) for v in args.sweep.split(",") if v]:
        if bs == B:
            continue
        _, ms_b, _, _, _ = run_config(bs, max(20, K), W)
        d_b = fb.make_desc(mech, bs, F, V, T, 512, 8, Cn, text_mode=tm, dtype=dtype, train=True, flags={"auto": 0, "simt": 4, "tc": 8}[args.engine])
        fl_b, by_b, _ = _lib.algorithmic_work(d_b)
        t_roof_b = max(fl_b / (roof["step_peak_tflops"] * 1e12), by_b / (roof["hbm_gbs"] * 1e9)) * 1e3
        sweep[str(bs)] = {"ms_per_step": ms_b, "samples_per_s": bs * world / (ms_b * 1e-3), "t_roof_ms": t_roof_b, "roofline_frac": t_roof_b / ms_b}

    # ---- the incumbent on the same box: the reference torch.nn module on device='cuda' (stock ATen / cuBLAS
    #      kernels, fp32, TF32 off like the reference), eager and as one CUDA graph; and this repo's eager DROP-IN route
    #      model(x, meta) -> criterion -> loss.backward() (no graph, no fused step) - the call shape of train_pad_20.py:110-112
    incumbent = None
    if rank == 0 and not args.no_incumbent:
        incumbent = {}
        for bs in sorted({32, B}):
            try:
                incumbent[str(bs)] = time_incumbents(torch, fb, dev, wl, bs, build)
            except Exception as exc:
                incumbent[str(bs)] = {"error": repr(exc)}

    dp_check = None
    if world > 1:
        dp_check = run_dp_check(torch, dist, fb, _lib, dev, wl, build, rank, world, comm.get("bucket"))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.torch_port import time_cpu_train_step
        r = time_cpu_train_step(mech, F, V, Cn, B, T=T, one_hot=(tm == 0), steps=6, warmup=2, threads=physical_cores(), budget_s=25.0)
        cpu = {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"], "kind": "port",
               "sample": f"{r['steps']} train steps of batch {B} on the reference torch.nn CPU path (median), {r['ms_per_step']:.1f} ms/step"}

    # ---- the token-sequence attention kernel, timed alone: image tokens attending to metadata tokens
    extras = None
    if rank == 0 and not args.no_extras:
        try:
            extras = {"mha_tokens": time_token_attention(torch, fb, dev)}
        except Exception as exc:                                    # never lose the headline line to an extra
            extras = {"mha_tokens": {"error": repr(exc)}}
        try:
            extras["tab_transformer"] = time_tab_transformer(torch, fb, dev)
        except Exception as exc:
            extras["tab_transformer"] = {"error": repr(exc)}
        try:
            extras["e2e_backbone"] = time_backbone_e2e(torch, fb, dev, wl)
        except Exception as exc:
            extras["e2e_backbone"] = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "fusion-head train samples/sec (fwd+bwd)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32" if dtype == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "mechanism": mech, "per_gpu_batch": B, "global_batch": B * world,
                       "F": F, "V": V, "C": Cn, "D": 512, "heads": 8, "parallelism": f"dp{world}", "grad_allreduce": ({"multimem": "hand-written one-kernel NVSwitch multimem.ld_reduce / multimem.st all-reduce over the live gradient ranges of a symmetric-memory bucket (csrc/dp_comm.cuh), captured in the step graph",
                                                          "peer": "hand-written one-kernel peer-load / peer-store all-reduce over the live gradient ranges of a symmetric-memory bucket (csrc/dp_comm.cuh), captured in the step graph",
                                                          "nccl": "NCCL all_reduce over the live span, captured in the step graph", "none": None}[comm["kind"]]), "launch": ("eager" if args.no_graph else "one CUDA graph per train step") + "; every kernel launched with programmatic dependent launch; metadata chain on an internal side stream",
                       "l2": f"inputs rotate over a pool of {nb} batches = {nb * in_bytes / 1e6:.0f} MB (> 126 MB L2 when >= 127); weights ({plive * 4 / 1e6:.1f} MB) stay L2-resident by design"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "how": "one packed pinned staging buffer (features | metadata | labels) and one cudaMemcpyAsync per step, double-buffered on a copy stream; loss read back per step"
                           + (f"; rank pinned to the {len(numa_cpus)} cores of its GPU's NUMA node" if numa_cpus else "")},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "loss": last_loss, "sweep": sweep or None,
            "incumbent": incumbent, "dp_check": dp_check, "extras": extras,
        }
        emit(line)
    if world > 1:
        if comm.get("bucket") is not None:
            # tearing down a process group that holds symmetric-memory rendezvous state hung for minutes here (probe run,
            # gpurun_out/r02_probe_symm_2gpu.txt): everything is flushed and the line is printed - leave without the teardown
            torch.cuda.synchronize(); dist.barrier(); sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


def dominant_kernel_roofline(torch, _lib, desc, dev, peaks, dtype, model):
    """Times every GEMM launch shape of one train step (forward NT, dX NN, dW TN of each Linear)
    through fb200_gemm - the same kernels the step launches - with CUDA events on the launch
    stream, and reports algorithmic FLOPs / measured time for the GEMM kernel family."""
    L = _lib.lib()
    n = C.c_int(0)
    cap = 256
    arr = (C.c_int32 * (cap * 5))()
    L.fb200_list_gemms.restype = C.c_int
    L.fb200_list_gemms.argtypes = [C.POINTER(_lib.Desc), C.POINTER(C.c_int32), C.c_int]
    cnt = L.fb200_list_gemms(C.byref(desc), arr, cap)
    shapes = {}
    for i in range(cnt):
        key = tuple(arr[i * 5 + j] for j in range(5))           # layout, engine, M, N, K
        shapes[key] = shapes.get(key, 0) + 1
    stream = torch.cuda.current_stream()
    tot_flops, tot_ms, per = 0.0, 0.0, []
    engine_names = {0: "simt_fp32_ffma", 1: "tcgen05_3xtf32", 2: "tcgen05_bf16"}
    for (layout, engine, M, N, Kd), count in shapes.items():
        a_shape = (Kd, M) if layout == 2 else (M, Kd)
        b_shape = (N, Kd) if layout == 0 else (Kd, N)
        A = torch.randn(*a_shape, device=dev); Bm = torch.randn(*b_shape, device=dev); Cm = torch.empty(M, N, device=dev)
        wsz = C.c_size_t(0)
        L.fb200_gemm_workspace_bytes(layout, engine, M, N, Kd, C.byref(wsz))
        ws = torch.empty(max(wsz.value, 256), dtype=torch.uint8, device=dev)

        def call():
            _lib.check(L.fb200_gemm(layout, engine, M, N, Kd, A.data_ptr(), a_shape[1], Bm.data_ptr(), b_shape[1], Cm.data_ptr(), N,
                                    None, 0, 0, ws.data_ptr(), ws.numel(), stream.cuda_stream), "fb200_gemm")
        for _ in range(3):
            call()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            call()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * M * N * Kd
        tot_flops += fl * count; tot_ms += ms * count
        per.append({"layout": "NT NN TN".split()[layout], "engine": engine_names[engine], "M": M, "N": N, "K": Kd, "launches_per_step": count,
                    "us": ms * 1e3, "tflops": fl / (ms * 1e-3) / 1e12})
    per.sort(key=lambda r: -r["us"] * r["launches_per_step"])
    # The step's own GEMM launches (forward, dX chain, ONE grouped weight-gradient launch), replayed alone on the
    # launch stream between CUDA events: algorithmic FLOPs of those launches / their measured time.
    from fusion_b200.head import ParamTable
    table = ParamTable([p.detach() if p is not None else None for p in model._params_in_slot_order()])
    Bsz = desc.B
    x = torch.randn(Bsz, desc.F, device=dev); t = torch.randn(Bsz, desc.V if desc.text_mode == 0 else desc.T, device=dev)
    ws = torch.zeros(_lib.workspace_bytes(desc), dtype=torch.uint8, device=dev)
    total, _ = _lib.grad_layout(desc)
    flat = torch.zeros(total, device=dev); logits = torch.zeros(Bsz, desc.C, device=dev)

    def replay():
        _lib.check(L.fb200_debug_gemm_replay(C.byref(desc), table.arr, x.data_ptr(), t.data_ptr(), logits.data_ptr(), flat.data_ptr(),
                                             ws.data_ptr(), stream.cuda_stream), "fb200_debug_gemm_replay")
    for _ in range(3):
        replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(stream)
    for _ in range(reps):
        replay()
    e1.record(stream)
    torch.cuda.synchronize()
    tot_ms = e0.elapsed_time(e1) / reps
    achieved = tot_flops / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm = peaks.get("hbm_gbs", 6650.0)
    src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    traffic, traffic_src = None, None
    try:   # per-launch DRAM bytes of the dominant kernel: a profiler counter, so it comes from the committed ncu --set full
           # capture of this round (tools/final_r02e.sh), not from this run - the file name says which capture
        traffic_src = "profiles/r02e_dominant_kernel_traffic.json"
        traffic = json.load(open(os.path.join(ROOT, traffic_src))).get("dram_bytes_per_launch")
    except Exception:
        traffic_src = None
    if dtype == "fp32":
        # fp32-strict GEMMs run as 3 TF32 MMAs per product: peak = cuBLAS TF32 8192^3, measured here, / 3
        tf32 = measure_tf32_peak(torch, dev)
        peak = tf32 / 3.0
        peak_note = (f"cuBLAS TF32 8192^3 measured live = {tf32:.0f} TFLOP/s, / 3 MMAs per fp32-strict product (SURVEY 8d rule); "
                     f"for scale: fp32 FFMA pipe peak 74.4, dense bf16 sustained {bf16_peak:.0f} ({src})")
    else:
        peak, peak_note = bf16_peak, f"dense bf16 sustained, {src}"
    return {"bound": "tensor", "kernel": "tc_gemm_kernel (tcgen05 GEMM family: forward NT, dX NN, dW TN of one train step)", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_note": peak_note, "bf16_peak_tflops": bf16_peak,
            "hbm_gbs": hbm, "step_peak_tflops": peak, "gemm_ms_per_step": tot_ms, "gemm_launches_per_step": sum(shapes.values()),
            "how": "all GEMM launches of one train step replayed alone (fb200_debug_gemm_replay), CUDA events on the launch stream, eager launches",
            "top_shapes_single_launch": per[:6]}


def time_incumbents(torch, fb, dev, wl, Bsz, build, steps=30, warm=5):
    """ms per train step (zero_grad + forward + weighted CE + backward) at batch `Bsz`, inputs resident, CUDA events:
    gpu_eager / gpu_eager_graph = the reference torch.nn module (oracle/torch_port.py, bit-identical to the reference) on cuda,
    eager_dropin = this repo's model through the reference call shape, one library call per pass, no CUDA graph."""
    from oracle.torch_port import ReferencePort
    mech, F, V, Cn, T, tm, dtype = wl
    g = torch.Generator().manual_seed(99)
    x = torch.randn(Bsz, F, generator=g).to(dev); t = torch.randn(Bsz, V if tm == 0 else T, generator=g).to(dev)
    y = torch.randint(0, Cn, (Bsz,), generator=g).to(dev)
    cw = torch.ones(Cn, device=dev)
    out = {}

    def timed(fn):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False          # the reference runs fp32 with PyTorch's defaults
    try:
        torch.manual_seed(1234)
        ref = ReferencePort(mech, F, Cn, V=V if V else 85, T=T, one_hot=(tm == 0)).to(dev).train()
        crit = torch.nn.CrossEntropyLoss(weight=cw)

        def ref_step():
            ref.zero_grad(set_to_none=True)
            crit(ref(x, t), y).backward()
        out["gpu_eager_ms"] = timed(ref_step)
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    ref_step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            ref.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                crit(ref(x, t), y).backward()
            out["gpu_eager_graph_ms"] = timed(graph.replay)
            del graph
        except Exception as exc:
            out["gpu_eager_graph_ms"] = None
            out["gpu_eager_graph_error"] = repr(exc)
        del ref
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev

    model = build(Bsz)
    crit2 = fb.FusedCrossEntropyLoss(weight=cw)

    def dropin_step():
        model.zero_grad(set_to_none=True)
        crit2(model(x, t), y).backward()
    out["eager_dropin_ms"] = timed(dropin_step)

    def fused_step():
        model.forward_loss(x, t, y, cw)
    out["eager_fused_step_ms"] = timed(fused_step)
    out["note"] = ("gpu_eager*: reference torch.nn module on cuda (fp32, TF32 off), eager launches / one CUDA graph; eager_dropin: "
                   "fusion_b200 model(x, meta) -> FusedCrossEntropyLoss -> backward(), no graph; eager_fused_step: forward_loss, no graph")
    return out


def run_dp_check(torch, dist, fb, _lib, dev, wl, build, rank, world, bucket=None, rows_per_rank=192):
    """N ranks + NCCL + global denominator against ONE process at the global batch, on the hardware: every rank draws the same
    global batch, runs its shard through forward_loss with the global weighted-CE denominator and SUM-all-reduces the flat
    gradient; the result must equal forward_loss on the whole batch (eval mode: dropout draws depend on the local row index)."""
    mech, F, V, Cn, T, tm, dtype = wl
    Bg = rows_per_rank * world
    g = torch.Generator().manual_seed(777)
    x = torch.randn(Bg, F, generator=g).to(dev); t = torch.randn(Bg, V if tm == 0 else T, generator=g).to(dev)
    y = torch.randint(0, Cn, (Bg,), generator=g).to(dev)
    counts = torch.bincount(y.cpu(), minlength=Cn).clamp_min(1).float()
    cw = (Bg / (Cn * counts)).to(dev)
    model = build(Bg)
    model.eval()
    lo, hi = fb.dp.shard_rows(Bg, rank, world)
    denom = fb.dp.global_denominator(y[lo:hi], cw)
    if bucket is not None:           # the hand-written all-reduce kernel on the symmetric-memory bucket, live ranges only
        loss_s, _ = model.forward_loss(x[lo:hi], t[lo:hi], y[lo:hi], cw, denom=denom, flat_out=bucket.tensor)
        bucket.all_reduce(_lib.grad_live_ranges(model.last_desc))
        flat_s = model.flat_grad[: _lib.grad_layout(model.last_desc)[0]].clone()
    else:
        loss_s, _ = model.forward_loss(x[lo:hi], t[lo:hi], y[lo:hi], cw, denom=denom)
        flat_s = model.flat_grad.clone()
        fb.dp.allreduce_gradients(flat_s)
    loss_s = loss_s.clone(); dist.all_reduce(loss_s)
    loss_g, _ = model.forward_loss(x, t, y, cw)
    flat_g = model.flat_grad[: flat_s.numel()]
    torch.cuda.synchronize()
    err = ((flat_s - flat_g).abs().max() / flat_g.abs().max()).item()
    l2 = ((flat_s - flat_g).norm() / flat_g.norm()).item()
    worst = torch.tensor([err, l2, abs(loss_s.item() - loss_g.item()) / abs(loss_g.item())], device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    return {"max_rel_err": worst[0].item(), "rel_l2": worst[1].item(), "loss_rel_err": worst[2].item(), "global_batch": Bg,
            "collective": bucket.mode if bucket is not None else "nccl",
            "how": f"{world} ranks x {rows_per_rank} rows, global denominator, NCCL SUM all-reduce of the flat gradient vs one process on {Bg} rows (eval mode)"}


def time_backbone_e2e(torch, fb, dev, wl, Bsz=32, steps=20, warm=5):
    """End-to-end train step WITH the stock image backbone (north star: "end-to-end step time is reported"): 224x224 fp32
    images from pinned host memory -> torchvision ResNet-50 (random init: no network for pretrained weights; frozen, stock
    PyTorch kernels, outside the optimisation target) -> fused head step -> loss read back.  Batch 32 = conf/.env.test:2."""
    import warnings
    mech, F, V, Cn, T, tm, dtype = wl
    name = {2048: "resnet-50", 512: "resnet-18", 1664: "densenet169"}.get(F)
    if name is None or tm != 0:
        return {"skipped": f"no stock torchvision backbone of width {F} / one-hot metadata for this workload"}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(1234)
        model = fb.MultimodalModel(Cn, 8, dev, f"random:{name}", "one-hot-encoder", vocab_size=V, attention_mecanism=mech,
                                   unfreeze_weights="frozen_weights", compute_dtype=dtype).to(dev).train()
    model.image_encoder.eval()                                   # frozen backbone: no batch-norm statistics updates
    g = torch.Generator().manual_seed(5)
    img = torch.randn(Bsz, 3, 224, 224, generator=g).pin_memory(); meta = torch.randn(Bsz, V, generator=g).pin_memory()
    y = torch.randint(0, Cn, (Bsz,), generator=g).pin_memory()
    cw = torch.ones(Cn, device=dev)
    host_loss = torch.empty(1).pin_memory()
    d_img = torch.empty(Bsz, 3, 224, 224, device=dev); d_meta = torch.empty(Bsz, V, device=dev); d_y = torch.empty(Bsz, dtype=torch.int64, device=dev)

    def timed(fn):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def full():
        d_img.copy_(img, non_blocking=True); d_meta.copy_(meta, non_blocking=True); d_y.copy_(y, non_blocking=True)
        loss, _ = model.forward_loss(d_img, d_meta, d_y, cw)
        host_loss.copy_(loss.reshape(1), non_blocking=True)

    def backbone_only():
        with torch.no_grad():
            model.image_encoder(d_img)

    feat = torch.randn(Bsz, F, device=dev)
    head_model = fb.MultimodalModel(Cn, 8, dev, f"identity:{F}", "one-hot-encoder", vocab_size=V, attention_mecanism=mech, compute_dtype=dtype).to(dev).train()

    def head_only():
        head_model.forward_loss(feat, d_meta, d_y, cw)
    t_full, t_bb, t_head = timed(full), timed(backbone_only), timed(head_only)
    return {"backbone": f"torchvision {name} (random init, frozen, fp32, stock PyTorch eager)", "batch": Bsz, "image": "3x224x224 fp32",
            "ms_per_step": t_full, "backbone_forward_ms": t_bb, "head_step_ms": t_head, "samples_per_s": Bsz / (t_full * 1e-3),
            "h2d_bytes_per_step": Bsz * (3 * 224 * 224 + V) * 4 + Bsz * 8,
            "note": "head_step_ms is eager forward_loss (no CUDA graph); the frozen backbone needs no backward"}


def time_token_attention(torch, fb, dev, Sq=197, Skv=85, B=32, D=512, H=8, reps=20):
    """fb200_mha_forward + fb200_mha_backward through the MultiheadAttention drop-in on ViT-sized image tokens (197)
    attending to 85 metadata tokens, batch 32, COMMON_DIM 512, 8 heads; CUDA events, median of `reps`."""
    m = fb.MultiheadAttention(D, H).to(dev)
    q = torch.randn(Sq, B, D, device=dev, requires_grad=True)
    kv = torch.randn(Skv, B, D, device=dev, requires_grad=True)
    dy = torch.randn(Sq, B, D, device=dev)
    tf, tb = [], []
    for i in range(reps + 3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        q.grad = kv.grad = None; m.zero_grad(set_to_none=True)
        e0.record(); out, _ = m(q, kv, kv); e1.record(); out.backward(dy); e2.record()
        torch.cuda.synchronize()
        if i >= 3:
            tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
    hd = D // H
    core = 4.0 * B * H * Sq * Skv * hd                      # QK^T and PV, forward
    proj = 2.0 * D * D * B * (2 * Sq + 2 * Skv)             # q, k, v, out projections
    fwd_ms, bwd_ms = statistics.median(tf), statistics.median(tb)
    return {"shape": {"Sq": Sq, "Skv": Skv, "B": B, "D": D, "H": H}, "fwd_us": fwd_ms * 1e3, "bwd_us": bwd_ms * 1e3,
            "fwd_tflops": (core + proj) / (fwd_ms * 1e-3) / 1e12, "note": "eager launches through autograd (host overhead included); fp32, probabilities never stored"}


def time_tab_transformer(torch, fb, dev, B=1024, reps=10):
    """The fused TabTransformer (csrc/tabt.cu: embedding gather + both encoder layers in one kernel per pass, then the fc MLP on the
    GEMM engines) against the same module composed from stock torch.nn on the same GPU - what models/tab_transformer.py:6-60 runs
    with device='cuda' - at the reference's dimensions (82 columns x 10 categories, d 32, 4 heads, ff 128, 2 layers, 85 outputs),
    train mode, forward + backward, fp32 (TF32 off), CUDA events, median of `reps`."""
    import torch.nn as nn
    cards = [10] * 82

    class Stock(nn.Module):
        def __init__(self):
            super().__init__()
            self.embeddings = nn.ModuleList([nn.Embedding(c, 32) for c in cards])
            layer = nn.TransformerEncoderLayer(d_model=32, nhead=4, dim_feedforward=128, activation="relu", dropout=0.3, batch_first=True)
            self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=2)
            self.numeric_projection = nn.Linear(4, 32)
            self.fc = nn.Sequential(nn.Linear(82 * 32 + 32, 128), nn.ReLU(), nn.Dropout(0.3), nn.Linear(128, 85))

        def forward(self, xc, xn):
            tok = torch.stack([e(xc[:, i]) for i, e in enumerate(self.embeddings)], dim=1)
            return self.fc(torch.cat([self.transformer_encoder(tok).flatten(start_dim=1), self.numeric_projection(xn)], dim=1))

    fused = fb.TabTransformer(cards, 4, output_dim=85).to(dev).train()
    stock = Stock().to(dev).train()
    xc = torch.randint(0, 10, (B, 82), device=dev); xn = torch.randn(B, 4, device=dev); g = torch.randn(B, 85, device=dev)

    def timed(m):
        ts = []
        for i in range(reps + 3):
            m.zero_grad(set_to_none=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m(xc, xn).backward(g); e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        f_ms, s_ms = timed(fused), timed(stock)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    flops = 2 * (2 * 82 * 32 * 96 + 4 * 82 * 82 * 32 + 2 * 82 * 32 * 32 + 4 * 82 * 32 * 128) * B      # encoder stack, forward
    return {"batch": B, "fused_fwd_bwd_ms": f_ms, "torch_nn_cuda_fwd_bwd_ms": s_ms, "speedup": s_ms / f_ms,
            "samples_per_s": B / (f_ms * 1e-3), "encoder_fwd_gflop": flops / 1e9,
            "note": "eager launches through autograd; fused = 2 encoder kernels + GEMMs, stock = torch.nn.TransformerEncoder on cuda"}


def measure_tf32_peak(torch, dev, n=8192):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


if __name__ == "__main__":
    main()
