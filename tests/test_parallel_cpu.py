"""CPU: the N > 1 host logic under gloo with world size 2 -- image sharding, max-over-ranks timing
reduction and the DDP gradient all-reduce (the only collective of the path)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import gkgnet_b200 as G
    from gkgnet_b200 import parallel as P

    w, r, _ = P.init_distributed("gloo")
    assert (w, r) == (world, rank)
    # timing reduction: the slowest rank wins; sums add up
    assert P.max_over_ranks(float(rank + 1)) == float(world)
    assert P.sum_over_ranks(1.0) == float(world)
    P.barrier()

    # data-parallel training step on the label head (CPU-capable part of the model): the averaged
    # gradients of the two half batches must equal the gradient of the full batch
    torch.manual_seed(0)
    head = G.LabelQueryHead(num_classes=6, in_channels=16)
    full_lab, full_gap = torch.randn(8, 6, 16), torch.randn(8, 16)
    tgt = (torch.rand(8, 6) < 0.3).float()
    lo, hi = P.shard_range(8, rank, world)

    class _Loss(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, lab, gap, t):
            out = self.m.forward_train((lab, gap), t)
            return out["bce_loss"] + out["asy_loss"]

    ddp = P.data_parallel(_Loss(head))
    loss = ddp(full_lab[lo:hi], full_gap[lo:hi], tgt[lo:hi])
    loss.backward()
    grads = {k: p.grad.clone() for k, p in head.named_parameters()}

    # the same exchange without DDP: gradients as views of one flat buffer, averaged with bucketed all-reduces
    # (parallel.FlatGradients -- what the captured training step of bench.py uses)
    torch.manual_seed(0)
    head2 = G.LabelQueryHead(num_classes=6, in_channels=16)
    fg = P.FlatGradients(list(head2.parameters()), bucket_bytes=256)        # several buckets, ragged last one
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in fg.params)
    for _ in range(2):                                                      # zero() makes the step repeatable
        fg.zero()
        out = _Loss(head2)(full_lab[lo:hi], full_gap[lo:hi], tgt[lo:hi])
        out.backward()
        fg.all_reduce_mean()
    flat = {k: p.grad.clone() for k, p in head2.named_parameters()}
    torch.save({"grads": grads, "flat": flat, "range": (lo, hi)}, os.path.join(out_dir, f"rank{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_shard_range_partitions_every_image_once():
    from gkgnet_b200.parallel import shard_range
    for total in (0, 1, 7, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_world2_gloo_gradient_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import gkgnet_b200 as G
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert r0["range"] == (0, 4) and r1["range"] == (4, 8)
    # both ranks hold the same (averaged) gradients after the all-reduce
    for k in r0["grads"]:
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k
    # and they equal the single-process gradient of the mean of the two half-batch losses
    torch.manual_seed(0)
    head = G.LabelQueryHead(num_classes=6, in_channels=16)
    lab, gap = torch.randn(8, 6, 16), torch.randn(8, 16)
    tgt = (torch.rand(8, 6) < 0.3).float()
    total = 0.0
    for lo, hi in ((0, 4), (4, 8)):
        out = head.forward_train((lab[lo:hi], gap[lo:hi]), tgt[lo:hi])
        total = total + 0.5 * (out["bce_loss"] + out["asy_loss"])
    total.backward()
    for k, p in head.named_parameters():
        assert torch.allclose(p.grad, r0["grads"][k], atol=1e-6, rtol=1e-5), k
        # FlatGradients: the same mean over ranks, on both ranks
        assert torch.allclose(p.grad, r0["flat"][k], atol=1e-6, rtol=1e-5), k
        assert torch.equal(r0["flat"][k], r1["flat"][k]), k


def test_pin_host_cores_partitions_the_affinity_mask():
    """Opt-in per-rank core pinning (GKG_PIN_CORES=1): disjoint contiguous slices that cover floor(n / world) cores each;
    checked in a child process so that the test runner's own affinity stays untouched."""
    import subprocess
    import sys
    code = (
        "import os, json, sys; sys.path.insert(0, %r)\n"
        "from gkgnet_b200 import parallel as P\n"
        "before = sorted(os.sched_getaffinity(0))\n"
        "mine = P.pin_host_cores(int(sys.argv[1]), int(sys.argv[2]))\n"
        "print(json.dumps([before, mine, sorted(os.sched_getaffinity(0))]))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import json
    world = 2
    seen = []
    for local in range(world):
        out = subprocess.run([sys.executable, "-c", code, str(local), str(world)], capture_output=True, text=True, check=True)
        before, mine, after = json.loads(out.stdout.strip().splitlines()[-1])
        if len(before) < world:
            assert mine == before
            return
        assert mine == after and len(mine) == len(before) // world
        assert set(mine) <= set(before) and not (set(mine) & set(seen))
        seen += mine
