"""CPU, world_size = 2 over gloo: the data-parallel host logic (row sharding, global weighted-CE
denominator, SUM all-reduce of the flat gradient bucket) makes two ranks on half batches equal to
one process on the global batch.  The per-rank arithmetic here is the numpy oracle standing in for
the CUDA kernels (same ``denom`` contract as fb200_cross_entropy / fb200_head_train_step)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, mech, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
    from fusion_b200 import _lib, dp, make_desc
    from oracle import head_oracle as ho
    from tests.golden import cases as C
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = dict(C.SMALL_DIMS, mechanism=mech)
    cfg = C.make_cfg(kw)
    B = 11                                             # odd: uneven shards
    params = C.gen_params(cfg, 77, np.float64)
    x, tin, labels, cw, masks = C.gen_inputs(cfg, B, 77, True, np.float64)
    lo, hi = dp.shard_rows(B, rank, world)
    den = dp.global_denominator(torch.from_numpy(labels[lo:hi]), torch.from_numpy(cw))
    sm = {k: v[lo:hi] for k, v in masks.items()}
    o = ho.head_forward_backward(cfg, params, x[lo:hi], tin[lo:hi], labels[lo:hi], cw, sm, denom=float(den))
    # flat bucket in the library's layout
    d = make_desc(mech, hi - lo, cfg.F, cfg.V, cfg.T, cfg.D, cfg.H, cfg.C)
    total, offs = _lib.grad_layout(d)
    names = _lib.param_names()
    flat = torch.zeros(total, dtype=torch.float64)
    for s, off in offs.items():
        g = o["grads"][names[s]]
        flat[off: off + g.size] = torch.from_numpy(g.ravel())
    ranges = _lib.grad_live_ranges(d)
    covered = torch.zeros(total, dtype=torch.bool)
    for b, e in ranges:
        covered[b:e] = True
    assert float(flat[~covered].abs().max()) == 0.0 if (~covered).any() else True     # what is not shipped is structurally zero
    local = flat.clone()
    dp.allreduce_gradients(flat, ranges=ranges)
    # what bench.py runs at N > 1: one in-place all-reduce of the live span, or two buckets around a split
    b0, e1 = ranges[0][0], ranges[-1][1]
    for split in (0, (b0 + e1) // 2 // 4 * 4):
        f2 = local.clone()
        bar = dp.BucketedAllReduce("cpu")
        bar.start(f2, ranges, split, None)
        bar.finish(f2)
        assert torch.equal(f2, flat), split
    num = torch.tensor([o["num"]]); dist.all_reduce(num)
    if rank == 0:
        full = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks)
        worst = 0.0
        for s, off in offs.items():
            g = full["grads"][names[s]]
            got = flat[off: off + g.size].numpy().reshape(g.shape)
            worst = max(worst, float(np.abs(got - g).max() / max(np.abs(g).max(), 1e-30)))
        q.put((worst, abs(float(num) / float(den) - full["loss"]), abs(float(den) - full["den"])))
    dist.destroy_process_group()


@pytest.mark.parametrize("mech", ["crossattention", "att-intramodal+residual+cross-attention-metadados", "metablock"])
def test_two_ranks_equal_one_process(mech):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mech, q)) for r in range(2)]
    for p in procs:
        p.start()
    worst, loss_err, den_err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst < 1e-12 and loss_err < 1e-12 and den_err < 1e-12


def test_shard_rows_cover_the_batch():
    from fusion_b200 import dp
    for B in (1, 7, 32, 1024):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_rows(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_bucket_split_of_live_ranges():
    from fusion_b200.dp import BucketedAllReduce
    ranges = [(0, 100), (150, 300), (300, 420)]
    lo, hi = BucketedAllReduce.split_ranges(ranges, 200)
    assert lo == [(0, 100), (150, 200)] and hi == [(200, 300), (300, 420)]
    lo, hi = BucketedAllReduce.split_ranges(ranges, 150)
    assert lo == [(0, 100)] and hi == [(150, 300), (300, 420)]
    assert sum(e - b for b, e in lo) + sum(e - b for b, e in hi) == sum(e - b for b, e in ranges)


def test_symmetric_bucket_range_normalisation_and_numa_binding():
    """Host logic of the hand-written all-reduce path: live ranges become 4-element aligned, sorted, merged vectors inside the
    bucket; the NUMA binding helper degrades to None without a GPU."""
    from fusion_b200.dp import SymmetricGradBucket, bind_to_gpu_numa_node
    nr = SymmetricGradBucket.normalize_ranges
    assert nr([(0, 6), (8, 12)], 16) == [(0, 12)]                      # 6 -> 8 touches the next range
    assert nr([(100, 110), (0, 4), (20, 22)], 128) == [(0, 4), (20, 24), (100, 112)]
    assert nr([(5, 7)], 8) == [(4, 8)]
    with pytest.raises(ValueError):
        nr([(0, 10)], 8)
    with pytest.raises(ValueError):
        nr([], 8)
    # the live ranges of every fusion string normalise inside their own flat buffer
    from fusion_b200 import _lib
    from fusion_b200.head import make_desc
    from oracle import head_oracle as ho
    for mech in ho.MECHANISMS:
        d = make_desc(mech, 32, 2048, 85, 512, 512, 8, 6)
        total, _ = _lib.grad_layout(d)
        if total:
            out = nr(_lib.grad_live_ranges(d), total)
            assert all(b % 4 == 0 and e % 4 == 0 and 0 <= b < e <= total for b, e in out) and out == sorted(out)
    import torch
    if not torch.cuda.is_available():
        assert bind_to_gpu_numa_node(0) is None
