"""Multi-GPU plumbing of the lockstep engine: games are independent, so ranks own disjoint game-id ranges and exchange
only (C1) the network weights at (re)load and (C2) finished-game records / counters. torch.distributed is the transport
(NCCL on GPUs, gloo in the CPU tests); there is no collective on the search path.

Reference: one GeneratorThread per device sharing a mutex-protected GameDataBuffer (src/selfplay/GeneratorManager.cpp:29-53,
160-181); NetworkLoader::get loads the file once per thread (src/selfplay/NetworkLoader.cpp:41-53)."""
import numpy as np
import torch
import torch.distributed as dist


def game_range(total_games, rank, world):
    """Global game ids [first, last) owned by `rank`; per-game RNG streams are keyed by these ids, not by rank-local order."""
    per = total_games // world
    extra = total_games % world
    first = rank * per + min(rank, extra)
    return first, first + per + (1 if rank < extra else 0)


def broadcast_weights(blob, src=0, device=None):
    """C1: rank `src` holds the fp32 weight blob (numpy or None elsewhere); every rank returns an identical numpy copy."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.ascontiguousarray(blob, np.float32)
    device = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    n = torch.tensor([0 if blob is None else int(np.asarray(blob).size)], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src)
    t = torch.empty(int(n.item()), dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(blob, np.float32).reshape(-1)))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def gather_records(records, device=None):
    """C2: all-gather variable-length byte records (one bytes object per rank) -> list of bytes, rank order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [bytes(records)]
    device = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    world = dist.get_world_size()
    size = torch.tensor([len(records)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)  # size word first, then padded payloads
    cap = max(int(s.item()) for s in sizes)
    payload = torch.zeros(max(cap, 1), dtype=torch.uint8, device=device)
    if len(records):
        payload[:len(records)] = torch.frombuffer(bytearray(records), dtype=torch.uint8).to(device)
    out = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(out, payload)
    return [bytes(o[:int(s.item())].cpu().numpy().tobytes()) for o, s in zip(out, sizes)]


def reduce_counters(values, device=None):
    """Sum and max over ranks of a small vector of float64 counters (hasEnoughGames, GeneratorManager.cpp:177-181; bench totals)."""
    v = torch.tensor(values, dtype=torch.float64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return v.numpy().copy(), v.numpy().copy()
    device = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    s, m = v.to(device).clone(), v.to(device).clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return s.cpu().numpy(), m.cpu().numpy()
