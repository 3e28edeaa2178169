"""Multi-GPU sharding of a batch of independent NLP evaluations.

The problems are independent: the batch is split contiguously across ranks, every rank
evaluates its slice on its own GPU with no collective on the hot path, and NCCL is used only for the
optional final gather of per-problem scalars (8-16 B per problem against ~285 KB produced).
One process per GPU; ``torch.distributed`` provides the plumbing (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(B: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``range(B)``: the first ``B % world_size`` ranks get one extra."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    q, r = divmod(B, world_size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def evaluate_sharded(evaluate: Callable, Z, *, group=None, gather: Optional[str] = "f"):
    """Evaluate this rank's slice of the GLOBAL batch ``Z`` (same on every rank) and optionally
    all-gather one per-problem scalar output (default: the objective) so every rank sees all of them.

    ``evaluate(Z_local) -> dict`` is the local evaluator (``HybridNLP.eval_batch`` on a GPU rank).
    Returns ``(local_outputs, (lo, hi), gathered_or_None)``.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = Z.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    out = evaluate(Z[lo:hi])
    gathered = None
    if gather is not None:
        v = out[gather]
        if world == 1:
            gathered = v.clone()
        else:
            # ragged all-gather: pad every slice to the largest shard
            width = -(-B // world)
            pad = torch.zeros(width, dtype=v.dtype, device=v.device)
            pad[: hi - lo] = v
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad, group=group)
            gathered = torch.cat([parts[r][: shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0]]
                                  for r in range(world)])
    return out, (lo, hi), gathered
