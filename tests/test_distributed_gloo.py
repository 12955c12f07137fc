"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path --
partitioning by contiguous element ranges, owner selection (lowest rank that
touches a node), and the point-to-point exchange that adds the non-owners'
partial interface rows on the owner -- run with the CPU oracle standing in for
the per-rank assembly (the GPU kernels are exercised by tests/multi_gpu_check.py
under torchrun on the B200 box)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from giraffe_b200 import meshes as M                                   # noqa: E402
from giraffe_b200.distributed import p2p_exchange, partition_ranges   # noqa: E402


def _sub_model(m, keep):
    """Model with the same nodes but only the elements in `keep` (a rank's partition)."""
    sub = M.Model(xyz=m.xyz, hooke=m.hooke, sections=m.sections)
    sub.section_defs, sub.shell_thickness, sub.cs_defs = m.section_defs, m.shell_thickness, m.cs_defs
    sub.elem_type, sub.elem_mat = m.elem_type[keep], m.elem_mat[keep]
    sub.elem_sec, sub.elem_cs = m.elem_sec[keep], m.elem_cs[keep]
    ptr = m.elem_ptr
    sub.elem_nodes = np.concatenate([m.elem_nodes[ptr[e]:ptr[e + 1]] for e in keep]).astype(np.int32)
    sub.constraints, sub.gravity = m.constraints, m.gravity
    sub.pretension = m.pretension[keep] if m.pretension is not None else None
    return M._finish(sub)


def _owned_elements(m, rank, world):
    counts = [int((m.elem_type == t).sum()) for t in (M.SHELL_1, M.BEAM_1, M.SOLID_1)]
    ranges = partition_ranges(counts, world)[rank]
    keep = []
    seen = {M.SHELL_1: 0, M.BEAM_1: 0, M.SOLID_1: 0}
    slot = {M.SHELL_1: 0, M.BEAM_1: 1, M.SOLID_1: 2}
    for e, t in enumerate(m.elem_type):
        k = seen[int(t)]
        seen[int(t)] += 1
        lo, hi = ranges[slot[int(t)]]
        if lo <= k < hi:
            keep.append(e)
    return np.array(keep, np.int64)


def _worker(rank, world, port_file, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_file)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.portdrv import PortOracle
    m = M.concat_models([M.beam_line(21), M.shell_plate(6, 5, warp=0.01, gravity=(0.0, 0.0, -9.81))])
    d = M.mask_displacements(m, np.random.default_rng(3).uniform(-1e-4, 1e-4, (m.n_nodes, 6)))
    gls, nf, nx = M.number_dofs(m)

    # single-process truth
    full = PortOracle(threads=1).load(m)
    full.assemble(d)
    fo, fi, fv, _ = full.csr("AA")
    fpa = full.vectors()[0]

    # this rank's partial assembly: same DOF numbering, only its elements
    keep = _owned_elements(m, rank, world)
    part = PortOracle(threads=1)
    sub = _sub_model(m, keep)
    part.load(sub)
    # the sub-model activates fewer DOFs; map its rows back to the global numbering
    sgls = part.gls()
    part.assemble(d)
    po, pi, pv, _ = part.csr("AA")
    ppa = part.vectors()[0]
    to_global = np.zeros(part.n_free, np.int64)
    mask = sgls > 0
    to_global[sgls[mask] - 1] = gls[mask] - 1
    dense = np.zeros((nf, nf))
    rows = np.repeat(np.arange(part.n_free), np.diff(po))
    dense[to_global[rows], to_global[pi]] = pv
    pa = np.zeros(nf)
    pa[to_global] = ppa

    # ownership: lowest rank whose elements touch the node
    touch = np.zeros((world, m.n_nodes), bool)
    for r in range(world):
        for e in _owned_elements(m, r, world):
            touch[r, m.elem_nodes[m.elem_ptr[e]:m.elem_ptr[e + 1]] - 1] = True
    owner = np.argmax(touch, axis=0)
    shared = touch.sum(axis=0) > 1
    # rows exchanged: free DOFs of shared nodes this rank touches
    def rows_of(nodes):
        g = gls[nodes].reshape(-1)
        return np.sort(g[g > 0] - 1)
    send_counts, recv_counts = np.zeros(world, np.int64), np.zeros(world, np.int64)
    send_rows, recv_rows = {}, {}
    for r in range(world):
        if r == rank:
            continue
        mine_to_r = np.nonzero(shared & touch[rank] & (owner == r))[0]
        r_to_mine = np.nonzero(shared & touch[r] & (owner == rank))[0]
        send_rows[r], recv_rows[r] = rows_of(mine_to_r), rows_of(r_to_mine)
        send_counts[r] = len(send_rows[r]) * (nf + 1)
        recv_counts[r] = len(recv_rows[r]) * (nf + 1)
    send = torch.from_numpy(np.concatenate([np.concatenate([dense[send_rows[r]].reshape(-1), pa[send_rows[r]]]) for r in sorted(send_rows)] or [np.zeros(0)]))
    recv = torch.zeros(int(recv_counts.sum()), dtype=torch.float64)
    p2p_exchange(send, send_counts, recv, recv_counts)
    off = 0
    for r in sorted(recv_rows):                      # ascending peer order = fixed summation order
        k = len(recv_rows[r])
        dense[recv_rows[r]] += recv[off:off + k * nf].numpy().reshape(k, nf)
        pa[recv_rows[r]] += recv[off + k * nf:off + k * (nf + 1)].numpy()
        off += k * (nf + 1)

    # every row is owned exactly once; owned rows equal the single-process result
    own_nodes = np.nonzero(touch[rank] & (owner == rank))[0]
    own_rows = rows_of(own_nodes)
    n_owned = torch.tensor([len(own_rows)])
    dist.all_reduce(n_owned)
    ref = np.zeros((nf, nf))
    ref[np.repeat(np.arange(nf), np.diff(fo)), fi] = fv
    err_k = np.abs(dense[own_rows] - ref[own_rows]).max() / np.abs(ref).max()
    err_p = np.abs(pa[own_rows] - fpa[own_rows]).max() / np.abs(fpa).max()
    result[rank] = (int(n_owned.item()), nf, float(err_k), float(err_p), int(send_counts.sum()))
    dist.destroy_process_group()


def test_partition_ranges_cover_every_element_once():
    counts = [1000003, 17, 0]
    for world in (1, 2, 3, 8):
        r = partition_ranges(counts, world)
        for t in range(3):
            assert r[0][t][0] == 0 and r[-1][t][1] == counts[t]
            for a, b in zip(r[:-1], r[1:]):
                assert a[t][1] == b[t][0]


def test_interface_exchange_world2_gloo():
    world = 2
    mgr = mp.Manager()
    result = mgr.dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, result), nprocs=world, join=True)
    assert len(result) == world
    for rank in range(world):
        n_owned, nf, err_k, err_p, sent = result[rank]
        assert n_owned == nf, "every free row must be owned by exactly one rank"
        assert err_k < 1e-14 and err_p < 1e-13, (err_k, err_p)
    assert result[1][4] > 0 and result[0][4] == 0, "only the non-owner (higher rank) sends"
