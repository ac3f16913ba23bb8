"""Test helper: host-side sharding of the overlap stage across ranks -- used by the gloo tests and
the profiling scripts; the product splits reads in the CLI (siga_main.cpp) and bench.py carries its own read_range.

Overlap search shards by independent units (reads): both FM-indices are replicated on every GPU, rank g searches the
contiguous read range [g*N/G, (g+1)*N/G).  VT records keep input order by concatenating the ranges in rank order; ED
order is free (the reference emits them in thread order, src/overlap_builder.cpp:447-480).  There is no data-path
collective; only the final gather of edge records to rank 0.
"""
import numpy as np


def read_range(n_reads, rank, world):
    """Contiguous, balanced index range of `rank` (first, count)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    first = (n_reads * rank) // world
    last = (n_reads * (rank + 1)) // world
    return first, last - first


def gather_edges(local_edges, dist=None):
    """Collect every rank's list of ED payload strings on rank 0 (returns None on the other ranks)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local_edges)
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object(list(local_edges), bucket, dst=0)
    if dist.get_rank() != 0:
        return None
    out = []
    for part in bucket:
        out.extend(part)
    return out


def gather_substring_flags(local_flags, n_reads, dist=None):
    """Per-read SS flags in input order on rank 0: ranges are contiguous, so a rank-ordered concatenation."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local_flags, dtype=np.uint8)
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object(np.asarray(local_flags, dtype=np.uint8), bucket, dst=0)
    if dist.get_rank() != 0:
        return None
    flags = np.concatenate(bucket)
    assert len(flags) == n_reads
    return flags
