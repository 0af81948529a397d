"""Replica sharding across GPUs: one process per GPU, no collective on the force path.

Replicas (umbrella windows, walkers) are independent evaluations sharing read-only weights, so the
path shards by contiguous blocks of replicas: rank r of G owns replicas
[r*R/G, (r+1)*R/G).  The only cross-GPU exchange the application ever needs is the per-window
energies at replica-exchange intervals; `gather_energies` is that one NCCL all-gather (R*8 bytes,
latency-bound, kept off the per-step path).  The reference has no multi-GPU support at all
(CudaANN_KernelFactory.cpp:68 uses contexts[0]).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_replicas: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of replicas owned by `rank`; block sizes differ by at most one."""
    base, extra = divmod(int(num_replicas), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class ReplicaSharder:
    def __init__(self, num_replicas: int, rank: int = None, world_size: int = None):
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.num_replicas, self.rank, self.world_size = int(num_replicas), int(rank), int(world_size)
        self.begin, self.end = shard_bounds(num_replicas, world_size, rank)
        self.counts = [shard_bounds(num_replicas, world_size, r)[1] - shard_bounds(num_replicas, world_size, r)[0]
                       for r in range(world_size)]

    @property
    def local_count(self) -> int:
        return self.end - self.begin

    def local_slice(self, array):
        """This rank's block of a [R, ...] array (numpy or torch)."""
        return array[self.begin:self.end]

    def gather_energies(self, local_energies: torch.Tensor) -> torch.Tensor:
        """All ranks' per-replica energies, in global replica order, on every rank (replica exchange)."""
        assert local_energies.numel() == self.local_count
        if self.world_size == 1:
            return local_energies.clone()
        width = max(self.counts)
        padded = torch.zeros(width, dtype=local_energies.dtype, device=local_energies.device)
        padded[: self.local_count] = local_energies
        out = torch.empty(self.world_size * width, dtype=local_energies.dtype, device=local_energies.device)
        dist.all_gather_into_tensor(out, padded)
        out = out.view(self.world_size, width)
        return torch.cat([out[r, : self.counts[r]] for r in range(self.world_size)])
