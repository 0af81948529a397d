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

    def _buffers(self, like: torch.Tensor):
        """Send / receive buffers of the energy all-gather, allocated once per (dtype, device)."""
        key = (like.dtype, like.device)
        if getattr(self, "_buf_key", None) != key:
            width = max(self.counts)
            self._width = width
            self._send = torch.zeros(width, dtype=like.dtype, device=like.device)
            self._recv = torch.empty(self.world_size * width, dtype=like.dtype, device=like.device)
            self._even = all(c == width for c in self.counts)
            if not self._even:
                # global replica r lives at rank(r) * width + (r - begin(rank(r))) of the padded receive buffer
                idx = [rk * width + i for rk in range(self.world_size) for i in range(self.counts[rk])]
                self._index = torch.tensor(idx, dtype=torch.int64, device=like.device)
            self._buf_key = key
        return self._send, self._recv

    def gather_energies(self, local_energies: torch.Tensor) -> torch.Tensor:
        """All ranks' per-replica energies, in global replica order, on every rank (replica exchange).

        One `all_gather_into_tensor` over pre-allocated buffers.  With equal shards (the usual case) the local energies are
        gathered straight from the caller's tensor and the result is a VIEW of the receive buffer, valid until the next call;
        with unequal shards the local block is padded to the widest shard and the result is one index_select."""
        assert local_energies.numel() == self.local_count
        if self.world_size == 1:
            return local_energies
        send, recv = self._buffers(local_energies)
        if self._even:
            dist.all_gather_into_tensor(recv, local_energies.contiguous().view(-1))
            return recv
        send[: self.local_count].copy_(local_energies.view(-1))
        dist.all_gather_into_tensor(recv, send)
        return recv.index_select(0, self._index)
