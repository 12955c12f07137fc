"""Multi-GPU parity check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank assembles its element partition, the interface rows are exchanged
over NCCL, and the rows each rank owns are compared with the CPU oracle's
single-process result (pattern byte-equal per row, values to 1e-12).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from giraffe_b200 import capi, meshes as M           # noqa: E402
from giraffe_b200.distributed import InterfaceExchange  # noqa: E402
from oracle.portdrv import PortOracle                # noqa: E402
import util                                          # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [
        ("shell", M.shell_plate(24, 16, warp=0.01, gravity=(0.0, 0.0, -9.81))),
        ("beam", M.beam_line(101)),
        ("shell-large", M.shell_plate(160, 96, warp=0.01)),      # 30k shells: thousands of scatter blocks around the exchange
        ("mixed", M.concat_models([M.beam_line(40), M.shell_plate(9, 8), M.solid_block(4, 4, 3)])),
    ]
    cases.append(("dynamic", M.concat_models([M.beam_line(30), M.shell_plate(12, 9, warp=0.01, gravity=(0.0, 0.0, -9.81))])))
    # the ring pipeline on a partitioned model: elements that touch a partition interface are pinned and evaluated
    # first, so that their rows are scattered, packed and sent before the ring kernels start
    cases.append(("shell-ring", M.shell_plate(60, 40, warp=0.01, gravity=(0.0, 0.0, -9.81))))
    ok = True
    for name, m in cases:
        d = M.mask_displacements(m, np.random.default_rng(7).uniform(-1e-4, 1e-4, (m.n_nodes, 6)))
        port = PortOracle(threads=2).load(m)
        port.set_time(0.0, 1.0)
        dynamic = name == "dynamic"
        if dynamic:      # Newmark path: UpdateDyn + MountMass / MountDamping(true) / MountDyn, partitioned like the rest
            rng = np.random.default_rng(11)
            cv, ca = rng.uniform(-1, 1, (m.n_nodes, 6)), rng.uniform(-10, 10, (m.n_nodes, 6))
            zeros = np.zeros((m.n_nodes, 6))
            a = util.newmark_coefficients(0.005)
            port.set_dynamic(a, 0.4, 2.0e-4)
            port.set_kinematics(zeros, zeros, cv, ca)
            port.update_dyn(d)
            port.assemble_dynamic(d, True)
        else:
            port.assemble(d)
        ro, ri, rv, _ = port.csr("AA")
        rpa, ria, rpb = port.vectors()
        if name == "shell-ring":
            os.environ.update(GFA_RING="1", GFA_RING_CHUNK_KB="512")
        else:
            os.environ.pop("GFA_RING", None)
        asm = capi.Assembler(m, device=local, rank=rank, world=world).set_dofs()
        if name == "shell-ring":
            assert asm.pipeline_info()[0], asm.pipeline_info()[1]
        asm.set_time(0.0, 1.0)
        ex = InterfaceExchange(asm, world)
        if dynamic:
            asm.set_dynamic(a, 0.4, 2.0e-4)
            asm.set_kinematics(zeros, zeros, cv, ca)
            asm.update_dyn(d)
            asm.assemble_dynamic(d, True)
            ex()
        else:
            asm.assemble(d)
            ex()
            # a second Newton iteration on the same pattern (slots are rewritten, not accumulated), this time with the
            # partition-local upload: only the nodes this rank's elements reference travel
            nodes = asm.touched_nodes()
            packed = np.ascontiguousarray(d.reshape(-1, 6)[nodes]).reshape(-1)
            asm.assemble(np.full_like(d, 123.0))        # poison the device copy first
            asm.set_displacements_packed(packed.ctypes.data)
            asm.assemble(None)
            ex()
            for _ in range(3):   # and queued ones: the exchange overlaps the scatter of the interior rows
                asm.assemble_enqueue(None)
                ex()
        lo, li, lv, _ = asm.csr("AA")
        rows = asm.local_rows()
        owned = asm.owned_rows()
        pos = {int(r): i for i, r in enumerate(rows)}
        dA = util.csr_diag((ro, ri, rv, (len(ro) - 1, len(ro) - 1)))
        worst = 0.0
        for r in owned:
            i = pos[int(r)]
            a, b = lo[i], lo[i + 1]
            ra, rb = ro[r], ro[r + 1]
            assert (b - a) == (rb - ra) and (li[a:b] == ri[ra:rb]).all(), f"{name}: pattern of row {r} differs on rank {rank}"
            scale = np.sqrt(dA[r] * dA[ri[ra:rb]])
            worst = max(worst, util.parity_error(rv[ra:rb], lv[a:b], scale))
        pa, ia, pb = asm.vectors()
        e_pa = util.parity_error(rpa[owned], pa[owned], float(np.abs(rpa).max()))
        po = asm.vector_owned(capi.P_A, np.zeros(len(owned)))
        assert po.tobytes() == np.ascontiguousarray(pa[owned]).tobytes(), f"{name}: gfa_vector_owned differs from the owned entries of gfa_vector"
        counts = torch.tensor([len(owned)], device="cuda")
        dist.all_reduce(counts)
        good = worst <= util.TOL and e_pa <= util.TOL and int(counts.item()) == asm.n_free
        ok &= good
        print(f"[rank {rank}] {name}: local rows {len(rows)}, owned {len(owned)} (sum over ranks {int(counts.item())} of {asm.n_free}), "
              f"interface doubles sent {int(ex.send_counts.sum())} / received {int(ex.recv_counts.sum())}, "
              f"worst K parity {worst:.2e}, P_A parity {e_pa:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if int(flag.item()):
        raise SystemExit(1)
    if rank == 0:
        print("multi-GPU parity OK")


if __name__ == "__main__":
    main()
