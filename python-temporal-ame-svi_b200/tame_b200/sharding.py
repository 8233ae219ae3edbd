"""Row ownership of the multi-GPU fit (host-side logic; mirrors tame_lrow / tame_grow / tame_owned in
csrc/tame_kernels.cuh).

Nodes are dealt to ranks in panels of `panel` consecutive nodes: panel b belongs to rank b % world.  A rank stores the
rows of Y (and works on the X_cov / H / hab rows) of its panels back to back in increasing node order; X_mean is
replicated.  Y is never communicated.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

DEFAULT_PANEL = 64


def owner_of(i: int, panel: int, world: int) -> int:
    return (i // panel) % world


def owned_rows(n: int, panel: int, world: int, rank: int) -> List[Tuple[int, int]]:
    """[lo, hi) node ranges owned by `rank`, in storage order."""
    return [(b * panel, min(n, (b + 1) * panel)) for b in range((n + panel - 1) // panel) if b % world == rank]


def local_count(n: int, panel: int, world: int, rank: int) -> int:
    return sum(hi - lo for lo, hi in owned_rows(n, panel, world, rank))


def local_row(i: int, panel: int, world: int) -> int:
    """Storage row of node i on its owner."""
    b = i // panel
    return (b // world) * panel + (i - b * panel)


def global_row(l: int, panel: int, world: int, rank: int) -> int:
    """Node stored in row l of `rank`."""
    lb = l // panel
    return (lb * world + rank) * panel + (l - lb * panel)


def shard_rows(Y, panel: int, world: int, rank: int):
    """Rows of a full (n, n, T, 2) array that `rank` stores, concatenated in storage order (torch or numpy)."""
    parts = [Y[lo:hi] for lo, hi in owned_rows(Y.shape[0], panel, world, rank)]
    if hasattr(Y, "numpy") or type(Y).__module__.startswith("torch"):
        import torch
        return torch.cat(parts, 0)
    import numpy as np
    return np.concatenate(parts, 0)


def deal_fits(costs: Sequence[float], n_groups: int) -> List[List[int]]:
    """Independent fits over devices (BASELINE config 5 shards trivially): largest cost first, each to
    the group with the smallest load so far (ties: the lowest group).  Returns the fit indices per group, each in the
    caller's order; deterministic."""
    if n_groups < 1:
        raise ValueError("n_groups must be positive")
    load = [0.0] * n_groups
    groups: List[List[int]] = [[] for _ in range(n_groups)]
    for f in sorted(range(len(costs)), key=lambda k: (-float(costs[k]), k)):
        g = min(range(n_groups), key=lambda k: (load[k], k))
        groups[g].append(f)
        load[g] += float(costs[f])
    return [sorted(g) for g in groups]
