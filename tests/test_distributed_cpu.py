"""N > 1 host logic on CPU: two gloo ranks shard the games, broadcast the weights (C1), gather records (C2) and reduce
counters exactly the way bench.py / a multi-GPU run does over NCCL."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from alphagomoku_b200 import sharding, netblob
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = sharding.game_range(10, rank, world)
    blob = netblob.pack(netblob.random_tensors(15, 15, 1, 64, False), 15, 15, 1, 64, False) if rank == 0 else None
    got = sharding.broadcast_weights(blob)
    records = bytes([rank + 1] * (3 + 4 * rank))
    gathered = sharding.gather_records(records)
    total, peak = sharding.reduce_counters([float(last - first), float(rank)])
    # a pop interval in which only one rank finished games (the bench pops every 10 steps: most intervals look like this at small sizes)
    sparse = sharding.gather_records(b"" if rank == 0 else b"\x07\x08")
    out.put((rank, first, last, float(got.sum()), got.size, [len(g) for g in gathered], [g[:1] for g in gathered], total.tolist(), peak.tolist(),
             [bytes(g) for g in sparse]))
    dist.destroy_process_group()


def test_two_rank_sharding_broadcast_gather():
    from alphagomoku_b200 import netblob
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expected = netblob.pack(netblob.random_tensors(15, 15, 1, 64, False), 15, 15, 1, 64, False)
    ranges = [(r[1], r[2]) for r in results]
    assert ranges == [(0, 5), (5, 10)]  # disjoint, covering, rank order
    for r in results:
        assert r[4] == expected.size and abs(r[3] - float(expected.sum())) < 1e-3  # identical weights on every rank
        assert r[5] == [3, 7] and r[6] == [b"\x01", b"\x02"]  # ragged records gathered in rank order
        assert r[7] == [10.0, 1.0] and r[8] == [5.0, 1.0]
        assert r[9] == [b"", b"\x07\x08"]


def test_game_range_is_a_partition():
    from alphagomoku_b200 import sharding
    for total in (0, 1, 7, 4096, 32768):
        for world in (1, 2, 3, 8):
            spans = [sharding.game_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
