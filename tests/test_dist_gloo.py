"""Host-side logic of the multi-GPU path on CPU: world size 2, gloo backend (no GPU needed)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pygradflow_b200.dist import gather_result, gather_rows, shard_range


def test_shard_range_partitions_the_batch():
    for B in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(B, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(B, rank, world)
        # every rank "solves" its block: deterministic stand-ins keyed by the global instance index
        idx = torch.arange(lo, hi, dtype=torch.float64)
        local = dict(
            x=torch.stack([idx, idx * 2, idx * 3], dim=1),
            y=idx[:, None] + 0.5,
            status=torch.ones(hi - lo, dtype=torch.int32),
            iterations=torch.arange(lo, hi, dtype=torch.int32),
            accepted_steps=torch.arange(lo, hi, dtype=torch.int32) // 2,
        )
        res = gather_result(local, B, (lo, hi))
        full = torch.arange(B, dtype=torch.float64)
        ok = (
            torch.equal(res.x, torch.stack([full, full * 2, full * 3], dim=1))
            and torch.equal(res.y, full[:, None] + 0.5)
            and torch.equal(res.iterations, torch.arange(B, dtype=torch.int32))
            and res.status.shape == (B,)
            and res.local_range == (lo, hi)
            and torch.equal(gather_rows(idx, B), full)
        )
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 9, 1])
def test_final_gather_world2_gloo(B):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r, False) for r in range(world))
