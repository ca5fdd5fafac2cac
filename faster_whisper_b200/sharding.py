"""Chunk-parallel sharding across GPUs: independent 30 s chunks, one process per GPU, a full weight
replica each, **no collective on the data path**.  The only communication is an optional final gather of the per-chunk
results (a few hundred int32 tokens + two floats per chunk) so that rank 0 can emit segments in chunk order, and the
barrier / max-over-ranks timing in ``bench.py``.

The reference's multi-GPU story has the same shape — CTranslate2 replicas fed by the caller's threads
(``faster_whisper/transcribe.py:646-657``) — minus its host bounce of encoder outputs (``:1392-1394``).
"""

from __future__ import annotations

from typing import Any, Dict, List, Sequence


def shard_indices(n_chunks: int, rank: int, world_size: int, block: int = 1) -> List[int]:
    """Chunk ids owned by `rank`: blocks of `block` consecutive chunks dealt round-robin (block = batch_size keeps
    every generate() call full)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return [i for i in range(n_chunks) if (i // max(1, block)) % world_size == rank]


def gather_results(local: Dict[int, Any], group=None) -> Dict[int, Any]:
    """All ranks call this with {chunk_id: result}; every rank gets the merged dict (backend-agnostic:
    ``all_gather_object`` over NCCL on the GPU box, gloo in the CPU tests).  Without an initialised process group it is
    the identity."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    parts: List[Dict[int, Any]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, dict(local), group=group)
    merged: Dict[int, Any] = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise RuntimeError(f"chunk {k} was processed by two ranks")
            merged[k] = v
    return merged


def transcribe_sharded(transcribe_chunks, chunks: Sequence, rank: int, world_size: int, block: int = 1, group=None):
    """Runs `transcribe_chunks(list_of_chunks) -> list_of_results` on this rank's shard and returns the results of
    *all* chunks in chunk order (after the final gather)."""
    mine = shard_indices(len(chunks), rank, world_size, block)
    out = transcribe_chunks([chunks[i] for i in mine]) if mine else []
    merged = gather_results(dict(zip(mine, out)), group)
    return [merged[i] for i in range(len(chunks))]
