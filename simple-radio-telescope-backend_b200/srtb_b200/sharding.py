"""Block sharding across the GPUs of one box: blocks are independent, so block k
goes to rank k mod G with every polarisation stream of a block on the same GPU; there is no
data-path collective. torch.distributed is used only for the barrier / max-over-ranks timing and
for gathering the (tiny) detection summaries on rank 0."""
from __future__ import annotations

from typing import Iterable, List, Sequence


def blocks_for_rank(n_blocks: int, world: int, rank: int) -> List[int]:
    """round-robin: block k -> rank k mod world"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    return list(range(rank, n_blocks, world))


def owner_of_block(block: int, world: int) -> int:
    return block % world


def merge_results(per_rank: Sequence[Iterable[dict]]) -> List[dict]:
    """merge per-rank detection summaries back into stream order: sorted by (block counter, stream id),
    the keys the reference's write_signal_pipe orders by (udp_packet_counter / data_stream_id)."""
    merged = [r for part in per_rank for r in part]
    merged.sort(key=lambda r: (r["block"], r.get("stream", 0)))
    return merged


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """device-timed numbers are reported as the max over ranks"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local: List[dict], dist=None) -> List[dict] | None:
    """rank 0 receives every rank's summaries merged in block order; other ranks get None"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return merge_results([local])
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [None] * world if rank == 0 else None
    dist.gather_object(local, out, dst=0)
    return merge_results(out) if rank == 0 else None
