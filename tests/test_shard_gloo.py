"""Host logic of the multi-GPU path on CPU: world_size-2 gloo, batch shard + the single all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusiondepth_b200 import shard


def test_shard_range_partitions_exactly():
    for gb in (0, 1, 7, 8, 32, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(gb, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == gb
            pos = 0
            for first, count in spans:
                assert first == pos
                pos += count
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, gb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = dict(rgb=torch.randn(gb, 3, 6, 8, generator=g), K=torch.zeros(gb, 4), tag=3)

        def fake_model(sample):  # per-sample function of the inputs, like the real path
            return {"pred": sample["rgb"].mean(1, keepdim=True) * 2 + 1}

        out = shard.run_sharded(fake_model, full, gb)
        want = full["rgb"].mean(1, keepdim=True) * 2 + 1
        ok = bool(torch.equal(out, want))
        # the off-stream, double-buffered variant used for steady-state serving (synchronous on CPU tensors)
        first, count = shard.shard_range(gb, rank, world)
        gat = shard.DepthGatherer(gb)
        tickets = [gat.submit(fake_model(shard.slice_sample(full, first, count))["pred"] + i) for i in range(3)]
        ok &= tickets == [0, 1, 0] and bool(torch.equal(gat.result(tickets[1]), want + 1))
        ok &= bool(torch.equal(gat.result(tickets[2]), want + 2))
        gat.drain()
        q.put((rank, ok, tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gb", [4, 5])
def test_two_rank_gloo_gather(gb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(shape == (gb, 1, 6, 8) for _, _, shape in res)
