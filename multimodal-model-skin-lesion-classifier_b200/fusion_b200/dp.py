"""Data-parallel plumbing for the fused head: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch on the box, gloo in the CPU tests).  The head shards over the batch with no
data-path collective; two reductions keep N ranks equal to one process at the global batch
(SURVEY.md 8e):

  1. the weighted-CE denominator sum_i w[y_i] must be the GLOBAL one (nn.CrossEntropyLoss(weight)
     normalises by the batch's weight sum, train_pad_20.py:52) - one scalar all-reduce before the
     step, handed to the kernels as ``denom``;
  2. the flat gradient buffer (one contiguous fp32 bucket laid out by fb200_grad_offset, W_q/W_k rows
     included as the zeros autograd materialises) is all-reduced with SUM - the global denominator
     already makes the per-rank gradients partial sums of the global mean.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def global_denominator(labels, class_weights, group=None, out=None):
    """sum over ALL ranks of class_weights[labels]; returns a 1-element tensor on labels' device."""
    w = class_weights[labels].sum().reshape(1) if class_weights is not None else labels.new_tensor([labels.numel()], dtype=torch.float32)
    # (kept in the dtype of class_weights: the library boundary - head.denom_arg, used by forward_loss and the fused loss -
    # casts to the ONE fp32 device scalar the kernels read, so float64 class weights from numpy / sklearn are safe)
    if out is not None:
        out.copy_(w)
        w = out
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(w, op=dist.ReduceOp.SUM, group=group)
    return w


def allreduce_gradients(flat_grad, group=None, async_op=False, ranges=None):
    """SUM all-reduce of the head's flat gradient bucket (in place).

    ``ranges`` ([(begin, end)] from ``_lib.grad_live_ranges``): move only the slices that can be non-zero - the
    W_q / W_k rows of every S=1 attention are structural zeros (a third less traffic for `crossattention`).
    The live slices are packed into one staging tensor, reduced with ONE collective and scattered back."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return None
    if not ranges or len(ranges) == 1 and ranges[0] == (0, flat_grad.numel()):
        return dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if async_op:
        raise ValueError("allreduce_gradients: async_op is only supported for the whole-buffer all-reduce (ranges=None)")
    views = [flat_grad[b:e] for b, e in ranges]
    staging = torch.cat(views)
    dist.all_reduce(staging, op=dist.ReduceOp.SUM, group=group)
    torch._foreach_copy_(views, list(staging.split([v.numel() for v in views])))
    return None


class SymmetricGradBucket:
    """The flat gradient buffer in SYMMETRIC memory + the hand-written one-kernel all-reduce over it (csrc/dp_comm.cuh).

        bucket = SymmetricGradBucket(numel, device)                  # collective: every rank, once
        model.forward_loss(..., flat_out=bucket.tensor)              # the step writes its gradients straight into it
        bucket.all_reduce(live_ranges)                               # barrier, ONE kernel (NVSwitch multimem or peer loads), barrier

    torch.distributed._symmetric_memory provides the plumbing only (allocation, the rendezvous that maps every peer's
    buffer and the multicast address, the cross-rank barrier); the reduction itself is fb200_dp_allreduce.  `mode`:
    "multimem" (in-switch reduction), "peer" (plain peer loads / stores), "auto" = multimem when the rendezvous returns a
    multicast address."""

    def __init__(self, numel, device, group=None, mode="auto"):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        group = group or dist.group.WORLD
        self.tensor = symm_mem.empty(int(numel), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.tensor, group.group_name)
        self.rank, self.world = self.handle.rank, self.handle.world_size
        mc = int(self.handle.multicast_ptr)
        if mode == "auto":
            mode = "multimem" if mc else "peer"
        if mode == "multimem" and not mc:
            raise RuntimeError("no NVSwitch multicast address on this system (symmetric memory rendezvous returned 0)")
        self.mode = mode
        self._mc = C.c_void_p(mc if mode == "multimem" else 0)
        self._peers = (C.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        # in-kernel barriers (FB200_DP_FUSED_BARRIER=1; measured on 8 x B200, tools/dp_comm_bench.py: 77 us against 65-73 us with
        # torch's two barrier kernels around the reduction, so the default keeps those): the LAST 4 * world uint32 slots of every rank's signal pad (torch's own barrier channels grow from
        # the front), two all-reduce channels x (open, close) x world slots
        self._pads = (C.c_void_p * self.world)(*[int(p) for p in self.handle.signal_pad_ptrs])
        pad_words = int(self.handle.signal_pad_size) // 4
        self._slot0 = pad_words - 4 * self.world
        self._state = torch.zeros(2, 4, dtype=torch.int32, device=device)
        self.fused_barriers = os.environ.get("FB200_DP_FUSED_BARRIER", "0") == "1" and self._slot0 >= 8 * self.world
        self._lib, self._C = _lib, C
        self._channel = 0
        self._ranges = {}

    @staticmethod
    def normalize_ranges(ranges, numel):
        """128-bit vectors: widen every [begin, end) to multiples of 4 elements (the flat buffer pads every slice to 4 and the
        padding is zero-filled with the rest of the buffer), sort, and merge what then touches or overlaps."""
        norm = []
        for b, e in sorted((int(b) & ~3, (int(e) + 3) & ~3) for b, e in ranges):
            if e <= b:
                continue
            if norm and b <= norm[-1][1]:
                norm[-1][1] = max(norm[-1][1], e)
            else:
                norm.append([b, e])
        if not norm:
            raise ValueError("no gradient range to reduce")
        if norm[0][0] < 0 or norm[-1][1] > numel:
            raise ValueError("gradient range outside the symmetric bucket")
        return [tuple(be) for be in norm]

    def all_reduce(self, ranges, max_ctas=0, channel=0):
        """SUM over all ranks, in place, of the [begin, end) element ranges of `self.tensor`, on the current stream.
        `max_ctas` > 0 caps the grid (overlap with a GEMM); `channel` selects the pair of barrier channels (two all-reduces
        in flight on two streams need two pairs)."""
        C = self._C
        key = tuple(ranges)
        hit = self._ranges.get(key)
        if hit is None:
            norm = self.normalize_ranges(ranges, self.tensor.numel())
            flat = [v for be in norm for v in be]
            hit = self._ranges[key] = ((C.c_int64 * len(flat))(*flat), len(norm))
        arr, nr = hit
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L = self._lib.lib()
        if self.fused_barriers:          # ONE kernel: opening barrier, reduction, closing barrier
            self._lib.check(L.fb200_dp_allreduce(self._mc, self._peers, arr, nr, self.rank, self.world, int(max_ctas), self._pads,
                                                 C.c_void_p(self._state[channel].data_ptr()), self._slot0 + 2 * self.world * channel, stream), "fb200_dp_allreduce")
            return
        self.handle.barrier(channel=2 * channel)                   # every rank's gradients are complete and visible
        self._lib.check(L.fb200_dp_allreduce(self._mc, self._peers, arr, nr, self.rank, self.world, int(max_ctas), None, None, 0, stream), "fb200_dp_allreduce")
        self.handle.barrier(channel=2 * channel + 1)               # every rank has stored its slice into every copy


class BucketedAllReduce:
    """Gradient all-reduce in two buckets, the first overlapped with the tail of the backward pass.

    The head's weight gradients are all produced by the grouped launch that ends the backward pass; with a ``mid_event``
    the library (fb200_head_train_step_dp) launches the half below ``split`` first and records the event once every
    gradient below ``split`` is final.  ``start`` (called right after the step was enqueued) makes a communication stream
    wait for that event and all-reduces bucket 1 there - under the second half of the weight gradients; ``finish``
    all-reduces bucket 2 behind the step and joins the communication stream back."""

    def __init__(self, device, group=None, bucket=None, overlap_ctas=32):
        """`bucket`: a SymmetricGradBucket - both buckets then go through the hand-written all-reduce kernel on the live
        ranges (the first with at most `overlap_ctas` CTAs, beside the weight-gradient GEMM it overlaps) instead of NCCL."""
        device = torch.device(device)
        self.stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None     # None: host tensors (gloo tests)
        self.group = group
        self.bucket, self.overlap_ctas = bucket, overlap_ctas

    @staticmethod
    def split_ranges(ranges, split):
        lo = [(b, min(e, split)) for b, e in ranges if b < split]
        hi = [(max(b, split), e) for b, e in ranges if e > split]
        return lo, hi

    def start(self, flat_grad, ranges, split, mid_event):
        """Each bucket travels as ONE in-place all-reduce over its contiguous span of the flat buffer (the structural
        zeros of W_q / W_k inside a span ride along): measured at 2 and 8 GPUs, the pack / unpack kernels and the
        stream hand-offs around them cost more exposed latency than the extra bytes."""
        self._hi = None
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return
        if self.bucket is not None:
            lo, hi = self.split_ranges(ranges, split) if split and mid_event is not None else ([], list(ranges))
            self._hi = hi
            if lo:
                self.stream.wait_event(mid_event)
                with torch.cuda.stream(self.stream):
                    self.bucket.all_reduce(lo, max_ctas=self.overlap_ctas, channel=1)
            return
        b0, e1 = min(b for b, _ in ranges), max(e for _, e in ranges)
        self._hi = (b0, e1)
        if not split or not (b0 < split < e1) or (mid_event is None and self.stream is not None):
            return
        self._hi = (split, e1)
        if self.stream is None:                      # host tensors: same two buckets, nothing to overlap
            dist.all_reduce(flat_grad[b0:split], op=dist.ReduceOp.SUM, group=self.group)
            return
        self.stream.wait_event(mid_event)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(flat_grad[b0:split], op=dist.ReduceOp.SUM, group=self.group)
        flat_grad.record_stream(self.stream)

    def finish(self, flat_grad):
        if self.bucket is not None:
            if self._hi:
                self.bucket.all_reduce(self._hi, channel=0)
            torch.cuda.current_stream().wait_stream(self.stream)
            return
        if self._hi is not None:
            dist.all_reduce(flat_grad[self._hi[0]:self._hi[1]], op=dist.ReduceOp.SUM, group=self.group)
            if self.stream is not None:
                torch.cuda.current_stream().wait_stream(self.stream)


class DenominatorPrefetcher:
    """All-reduces the weighted-CE denominator of the NEXT batch on a side stream while the current step runs.
    `slots` static device buffers rotate; a slot is rewritten only after the step that read it has finished."""

    def __init__(self, device, slots, group=None):
        self.stream = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.consumed = [None] * slots
        self.group = group

    def issue(self, slot, labels, class_weights, out):
        with torch.cuda.stream(self.stream):
            if self.consumed[slot] is not None:
                self.stream.wait_event(self.consumed[slot])
            global_denominator(labels, class_weights, self.group, out=out)
            self.ready[slot].record(self.stream)

    def wait(self, slot):
        torch.cuda.current_stream().wait_event(self.ready[slot])

    def mark_consumed(self, slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.consumed[slot] = ev


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off (sysfs: the PCI device's local_cpulist), so that
    pinned host buffers allocated afterwards are node-local: with 8 ranks feeding 8 GPUs from one host, H2D copies that
    cross the socket interconnect were the end-to-end limiter.  Returns the cpu list used, or None when sysfs has no answer."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def shard_rows(n_rows, rank, world):
    """Contiguous row range [lo, hi) of this rank (rank r gets rows r*B/N .. (r+1)*B/N)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DataParallelHead:
    """Wraps a MultimodalModel for data-parallel training of the head with the fused step:

        dp = DataParallelHead(model)
        loss = dp.train_step(img_feat_shard, meta_shard, label_shard, class_weights)   # grads are global afterwards
    """

    def __init__(self, model, group=None):
        self.model, self.group = model, group

    def train_step(self, image, text_metadata, label, class_weights=None):
        """Returns (loss, logits): `loss` is the GLOBAL weighted-mean loss (the per-rank numerators over the global
        denominator, summed over the ranks), i.e. what one process would report on the global batch; `logits` are this
        rank's rows."""
        denom = global_denominator(label.to(self.model.device), None if class_weights is None else class_weights.to(self.model.device), self.group)
        loss, logits = self.model.forward_loss(image, text_metadata, label, class_weights, denom=denom)
        allreduce_gradients(self.model.flat_grad, self.group)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            loss = loss.clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss, logits
