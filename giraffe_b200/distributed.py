"""Interface-row exchange for mesh-partitioned assembly (one process per GPU).

Each rank evaluates a contiguous range of elements; rows of nodes that sit on a
partition interface receive partial sums on every rank that touches them.
``exchange_interface`` moves only those rows: partials travel from the
non-owning ranks to the owner (lowest rank touching the node) as
``torch.distributed`` point-to-point transfers -- NCCL over NVLink on the GPU
box, gloo in the CPU tests -- and the owner adds them in ascending peer order,
which keeps the summation order fixed.

The functions work on any tensor device so that the host-side logic can be
covered with world_size-2 gloo tests; pack / unpack are the library's
``gfa_interface_pack`` / ``gfa_interface_unpack`` on the GPU and plain index
operations in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def p2p_exchange(send_buf: torch.Tensor, send_counts, recv_buf: torch.Tensor, recv_counts):
    """send_buf / recv_buf are laid out by peer rank ascending (counts per peer)."""
    world = dist.get_world_size()
    ops, so, ro = [], 0, 0
    for r in range(world):
        ns, nr = int(send_counts[r]), int(recv_counts[r])
        if ns:
            ops.append(dist.P2POp(dist.isend, send_buf[so:so + ns], r))
        if nr:
            ops.append(dist.P2POp(dist.irecv, recv_buf[ro:ro + nr], r))
        so += ns
        ro += nr
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv_buf


class InterfaceExchange:
    """Binds an assembler-like object (interface_counts / interface_pack /
    interface_unpack taking raw device pointers) to its exchange buffers."""

    def __init__(self, asm, world: int, device: str = "cuda"):
        self.asm, self.world = asm, world
        self.send_counts, self.recv_counts = asm.interface_counts(world)
        self.send_buf = torch.empty(int(self.send_counts.sum()), dtype=torch.float64, device=device)
        self.recv_buf = torch.empty(int(self.recv_counts.sum()), dtype=torch.float64, device=device)

    @property
    def bytes_per_step(self) -> int:
        return 8 * int(self.send_counts.sum() + self.recv_counts.sum())

    def __call__(self):
        """pack -> NCCL -> unpack, all ordered on the library's interface stream (gfa_interface_stream): no host
        synchronisation.  The library scatters the interface rows first, so after gfa_assemble_enqueue the
        exchange overlaps the scatter of the interior rows; unpack makes the library's main stream wait."""
        if self.world == 1:
            return
        if self.recv_buf.is_cuda:
            with torch.cuda.stream(torch.cuda.ExternalStream(self.asm.interface_stream())):
                self.asm.interface_pack(self.send_buf.data_ptr())
                p2p_exchange(self.send_buf, self.send_counts, self.recv_buf, self.recv_counts)
                self.asm.interface_unpack(self.recv_buf.data_ptr())
        else:
            self.asm.interface_pack(self.send_buf.data_ptr())
            p2p_exchange(self.send_buf, self.send_counts, self.recv_buf, self.recv_counts)
            self.asm.interface_unpack(self.recv_buf.data_ptr())


def partition_ranges(type_counts, world: int):
    """Contiguous element range of every type per rank: [count*r/world, count*(r+1)/world)
    (the rule gfa_create applies; reference elements are independent, Solution.cpp:231-236)."""
    return [[(int(c) * r // world, int(c) * (r + 1) // world) for c in type_counts] for r in range(world)]
