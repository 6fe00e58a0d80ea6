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

import torch
import torch.distributed as dist


def global_denominator(labels, class_weights, group=None, out=None):
    """sum over ALL ranks of class_weights[labels]; returns a 1-element tensor on labels' device."""
    w = class_weights[labels].sum().reshape(1) if class_weights is not None else labels.new_tensor([labels.numel()], dtype=torch.float32)
    w = w.to(torch.float32)                       # the kernels read ONE fp32 scalar (float64 class weights are cast here)
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


class BucketedAllReduce:
    """Gradient all-reduce in two buckets, the first overlapped with the tail of the backward pass.

    The head's weight gradients are all produced by the grouped launch that ends the backward pass; with a ``mid_event``
    the library (fb200_head_train_step_dp) launches the half below ``split`` first and records the event once every
    gradient below ``split`` is final.  ``start`` (called right after the step was enqueued) makes a communication stream
    wait for that event and all-reduces bucket 1 there - under the second half of the weight gradients; ``finish``
    all-reduces bucket 2 behind the step and joins the communication stream back."""

    def __init__(self, device, group=None):
        device = torch.device(device)
        self.stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None     # None: host tensors (gloo tests)
        self.group = group

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
