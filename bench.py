#!/usr/bin/env python
"""Benchmark of the per-Newton-iteration element assembly (BASELINE.json metric:
elements assembled/sec, Kt+Fint -> CSR, FP64).

    python bench.py --gpus N --steps K --warmup W            # CUDA path, C-ABI
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU path

A "step" is one pass of the hot path (Clear + MountLocal + MountElementLoads +
MountGlobal + MountSparse) over one synthetic batch.  The headline line is the
1M-element Shell_1 plate of BASELINE.json configs[2] per GPU (weak scaling: N GPUs
assemble an N-times larger plate, partitioned by contiguous element ranges, the
interface rows exchanged with NCCL send/recv).  The same JSON line carries
`side_configs`: the other BASELINE configs (100k Beam_1 line, 4M Solid_1 block, the
mixed Beam_1 + Shell_1 + Solid_1 model of configs[4] at 1M elements per GPU) and,
for N > 1, the STRONG-scaling number of the fixed 1M-shell plate -- each with the
parity probe of the headline (`parity_ok`): rows of sampled nodes, partition
interfaces included, against a single-GPU assembly of the elements around them.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from giraffe_b200 import meshes as M  # noqa: E402

METRIC = "elements assembled/sec (Kt+Fint->CSR, FP64)"
UNIT = "elements/s"
# compulsory HBM traffic and structure-exploiting flop count
SHELL_ALG_FLOPS = 5.0e4
READ_BYTES = {"shell": 928.0, "beam": 524.0, "solid": 8 * (8 * 9 + 8) + 36.0}      # per element (Solid_1: builder-defined)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def profile_file(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# workloads = BASELINE.json configs
# ---------------------------------------------------------------------------
def workload(kind: str, n_gpus: int, scaling: str = "weak", cells=(1000, 500)):
    """(model, displacements, config dict).  Weak scaling multiplies the per-GPU size by N along the direction
    the element numbering runs last, so that contiguous element ranges are strips with one interface each."""
    f = n_gpus if scaling == "weak" else 1
    if kind == "shell":
        nx, ny = cells
        m = M.shell_plate(nx, ny * f)
        d = M.shell_plate_displacements(m)
        cfg = {"workload": f"shell_plate_{nx}x{ny * f}cells_Shell_1", "config": "BASELINE.json configs[2] (1M-element Shell_1 plate)" + (" per GPU" if scaling == "weak" else ", fixed size")}
    elif kind == "beam":
        m = M.beam_line(100_000 * f)
        d = M.beam_line_displacements(m)
        cfg = {"workload": f"beam_line_{100_000 * f}_Beam_1", "config": "BASELINE.json configs[1] (100k-element Beam_1 line)" + (" per GPU" if scaling == "weak" else "")}
    elif kind == "solid":
        m = M.solid_block(160, 160, 156 * f)
        d = M.solid_block_displacements(m)
        cfg = {"workload": f"solid_block_160x160x{156 * f}_Solid_1", "config": "BASELINE.json configs[3] (4M-element Solid_1 block; builder-defined hexahedron, reference bodies are empty)" + (" per GPU" if scaling == "weak" else "")}
    elif kind == "mixed":
        # configs[4]: 8M elements on 8 GPUs = 1M per GPU: 125k Beam_1 + 375k Shell_1 + 500k Solid_1 per GPU
        m = M.concat_models([M.beam_line(125_000 * f), M.shell_plate(750, 250 * f), M.solid_block(100, 100, 50 * f)])
        d = M.mask_displacements(m, np.random.default_rng(20240005).uniform(-1e-4, 1e-4, (m.n_nodes, 6)))
        cfg = {"workload": f"mixed_{125_000 * f}beams+{375_000 * f}shells+{500_000 * f}solids",
               "config": "BASELINE.json configs[4] (mixed Beam_1 + Shell_1 + Solid_1, 1M elements per GPU: 8M on 8 GPUs)", "mix": "12.5 % Beam_1, 37.5 % Shell_1, 50 % Solid_1"}
    else:
        raise SystemExit(f"unknown workload {kind}")
    cfg.update({"elements": int(m.n_elements), "per_gpu_elements": int(m.n_elements // n_gpus), "scaling": scaling,
                "partition": f"{n_gpus} contiguous element ranges per type", "l2": "inputs and outputs far larger than L2 (no flush needed between steps)"})
    return m, d, cfg


def algorithmic_bytes(m, nnz_local, n_free_local, world):
    """compulsory reads per element + 8 B per CSR non-zero + 16 B per free DOF (P_A, I_A)."""
    t = m.elem_type
    n = {"beam": int(np.count_nonzero((t == M.BEAM_1) | (t == M.PIPE_1))), "shell": int(np.count_nonzero(t == M.SHELL_1)), "solid": int(np.count_nonzero(t == M.SOLID_1))}
    return sum(READ_BYTES[k] * n[k] for k in n) / world + 8.0 * nnz_local + 16.0 * n_free_local


# ---------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path
# ---------------------------------------------------------------------------
def cpu_oracle(threads: int):
    from oracle import refdrv
    if refdrv.available():
        return refdrv.RefOracle(threads=threads), "reference"
    from oracle.portdrv import PortOracle
    return PortOracle(threads=threads), "port"


def time_cpu(steps: int, warmup: int, cells=(100, 50)):
    threads = os.cpu_count() or 1
    orc, kind = cpu_oracle(threads)
    m = M.shell_plate(*cells)
    d = M.shell_plate_displacements(m)
    orc.load(m)
    for _ in range(warmup):
        orc.assemble(d)
    t, local = [], []
    for _ in range(steps):
        s = orc.assemble(d)
        t.append(float(s[:4].sum()))       # MountLocal + MountElementLoads + MountGlobal + MountSparse
        local.append(float(s[0]))
    med = float(np.median(t))
    return {"value": m.n_elements / med, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"Shell_1 plate {cells[0]}x{cells[1]} cells = {m.n_elements} elements (the headline mesh family at a size the CPU finishes in seconds), median of {steps} steps",
            "sample_elements": int(m.n_elements), "ms_per_step": med * 1e3,
            "mount_local_only_elements_per_s": m.n_elements / float(np.median(local)),
            "note": "OpenMP MountLocal/MountElementLoads, serial MountGlobal + setFromTriplets as in the reference; "
                    "GEMM and setFromTriplets are restatements (no MKL/Eigen on the box)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cells = tuple(int(c) for c in args.cpu_cells.split("x"))
    cb = time_cpu(max(args.steps, 1), max(args.warmup, 1), cells)
    _, _, cfg = workload("shell", args.gpus)
    cfg = dict(cfg)
    # what is timed is the SAMPLE, not the 1M plate of the CUDA arm: say so in config itself
    cfg.update({"workload": f"shell_plate_{cells[0]}x{cells[1]}cells_Shell_1 (bounded CPU sample of the {cfg['workload']} family)",
                "elements": cb["sample_elements"], "per_gpu_elements": cb["sample_elements"],
                "same_config_as_cuda_arm": False, "extrapolates_to": "elements/s of the full plate (see full_size_step for one measured step at 1M elements)"})
    full = None
    if not args.no_full_size_step:
        # one real step of the reference at the CUDA arm's single-GPU size (1M shells fit in host memory)
        try:
            t0 = time.time()
            threads = os.cpu_count() or 1
            orc, kind = cpu_oracle(threads)
            m = M.shell_plate(1000, 500)
            d = M.shell_plate_displacements(m)
            orc.load(m)
            s = orc.assemble(d)
            sec = float(s[:4].sum())
            full = {"elements": int(m.n_elements), "seconds": sec, "value": m.n_elements / sec, "unit": UNIT, "steps": 1, "kind": kind,
                    "phases_s": {"MountLocal": float(s[0]), "MountElementLoads": float(s[1]), "MountGlobal": float(s[2]), "MountSparse": float(s[3])},
                    "wall_s_with_setup": time.time() - t0}
        except Exception as e:      # noqa: BLE001  (host memory, 32-bit triplet counters of the reference ...)
            full = {"error": repr(e)[:300]}
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "full_size_step": full,
            "note": cb["note"] + "; value = the bounded sample; full_size_step = one measured step of the same code at the 1M-element plate"}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------
class Runner:
    """One workload on this rank's GPU: set-up, exchange buffers, timed loops."""

    def __init__(self, m, d_host, rank, world, local_rank, dist):
        import torch
        from giraffe_b200 import capi
        self.torch, self.capi, self.dist = torch, capi, dist
        self.m, self.rank, self.world = m, rank, world
        t0 = time.time()
        self.asm = asm = capi.Assembler(m, device=local_rank, rank=rank, world=world)
        self.gls, self.nf, self.nx = M.number_dofs(m)
        asm.set_dofs(self.gls, self.nf, self.nx)
        self.setup_s = time.time() - t0
        self.d_host = d_host
        self.d_dev = torch.from_numpy(np.ascontiguousarray(d_host).reshape(-1)).cuda()
        self.lib_stream = torch.cuda.ExternalStream(asm.stream())
        self.if_stream = torch.cuda.ExternalStream(asm.interface_stream())
        self.send_cnt, self.recv_cnt = asm.interface_counts(world)
        self.send_buf = torch.empty(int(self.send_cnt.sum()), dtype=torch.float64, device="cuda") if world > 1 else None
        self.recv_buf = torch.empty(int(self.recv_cnt.sum()), dtype=torch.float64, device="cuda") if world > 1 else None
        self.nnz = [asm.csr_dims(w)[2] for w in ("AA", "AB", "BA", "BB")]
        self.eval_ms, self.scat_ms = [], []

    def close(self):
        self.asm.close()
        self.d_dev = self.send_buf = self.recv_buf = None
        self.torch.cuda.empty_cache()

    def exchange(self):
        """pack -> NCCL send/recv -> unpack, stream-ordered on the library's interface stream (no host syncs);
        the library scatters the interface rows first, so the exchange overlaps the interior rows' scatter."""
        if self.world == 1:
            return
        torch, dist = self.torch, self.dist
        with torch.cuda.stream(self.if_stream):
            self.asm.interface_pack(self.send_buf.data_ptr())
            ops, so, ro = [], 0, 0
            for r in range(self.world):
                if self.send_cnt[r]:
                    ops.append(dist.P2POp(dist.isend, self.send_buf[so:so + int(self.send_cnt[r])], r))
                if self.recv_cnt[r]:
                    ops.append(dist.P2POp(dist.irecv, self.recv_buf[ro:ro + int(self.recv_cnt[r])], r))
                so += int(self.send_cnt[r]); ro += int(self.recv_cnt[r])
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            self.asm.interface_unpack(self.recv_buf.data_ptr())

    def step_resident(self):
        if self.world > 1:
            # enqueue only: the host queues pack / NCCL / unpack behind the assembly instead of leaving the
            # GPU idle while it catches up; reads of the results wait for the stream
            self.asm.assemble_enqueue(self.d_dev.data_ptr())
        else:
            self.asm.assemble(None, device_ptr=self.d_dev.data_ptr())
        self.exchange()

    def step_resident_logged(self):
        self.step_resident()
        if self.world == 1:
            t = self.asm.timing()
            self.eval_ms.append(t["eval_ms"]); self.scat_ms.append(t["scatter_ms"])

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, host_side=False):
        """Device time of `steps` calls: CUDA events on the library's stream, bracketed by barrier + synchronize;
        max over ranks.  Paths with host-side pieces (exchange waits, D2H) take the larger of device and wall time."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.lib_stream)
        w0 = time.perf_counter()
        for _ in range(steps):
            fn()
        e1.record(self.lib_stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        self.barrier()
        ms = e0.elapsed_time(e1)
        ms = max(ms, wall * 1e3) if (self.world > 1 or host_side) else ms
        if self.dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def resident(self, steps, warmup):
        for _ in range(max(warmup, 3)):
            self.step_resident()
        ms = self.timed(self.step_resident_logged, steps)
        if self.world > 1:      # the timed loop of N > 1 does not read per-kernel times (that would wait for every step)
            for _ in range(3):
                self.step_resident()
                t = self.asm.timing()
                self.eval_ms.append(t["eval_ms"]); self.scat_ms.append(t["scatter_ms"])
        return ms / steps

    # ---- end to end through the C-ABI with HOST buffers ------------------------------------------------
    def e2e(self, steps):
        """Every step: H2D of the displacements of the nodes this rank's elements reference (pinned, packed),
        the assembly, the interface exchange, D2H of this rank's CSR values (AA, AB, BA, BB) and of its owned
        rows of P_A / I_A (+ P_B) into pinned host buffers -- what a host-side solver consumes."""
        torch, asm, capi = self.torch, self.asm, self.capi
        nodes = asm.touched_nodes()
        packed = torch.empty(len(nodes) * 6, dtype=torch.float64).pin_memory()
        packed.numpy()[:] = self.d_host.reshape(-1, 6)[nodes].reshape(-1)
        owned = len(asm.owned_rows()) if self.world > 1 else self.nf
        out_vals = [torch.empty(max(n, 1), dtype=torch.float64).pin_memory() for n in self.nnz]
        out_vecs = [torch.empty(max(n, 1), dtype=torch.float64).pin_memory() for n in (owned, owned, self.nx)]

        def step():
            asm.set_displacements_packed(packed.data_ptr())
            if self.world > 1:
                asm.assemble_enqueue(None)
            else:
                asm.assemble(None)
            self.exchange()
            for w, buf, n in zip(("AA", "AB", "BA", "BB"), out_vals, self.nnz):
                asm.values(w, out=buf.numpy()[:n])
            for w, buf, n in zip((capi.P_A, capi.I_A, capi.P_B), out_vecs, (owned, owned, self.nx)):
                asm.vector_owned(w, buf.numpy()[:n])

        step()
        ms = self.timed(step, steps, host_side=True) / steps
        h2d = len(nodes) * 48
        d2h = 8 * (sum(self.nnz) + 2 * owned + self.nx)
        if self.dist is not None:      # whole-job bytes per step
            t = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t)
            h2d, d2h = int(t[0].item()), int(t[1].item())
        return ms, h2d, d2h

    # ---- parity probe ---------------------------------------------------------------------------------
    def parity(self, n_nodes=24):
        """Rows of sampled nodes -- those of this rank's first and last elements (the partition interfaces when
        N > 1) and a few interior ones -- against a single-GPU assembly of all elements around them."""
        from giraffe_b200 import capi
        torch, asm, m = self.torch, self.asm, self.m
        ptr, en = m.elem_ptr, m.elem_nodes
        touched = asm.touched_nodes() + 1                       # 1-based ids of this rank's nodes
        owned = set(int(r) for r in (asm.owned_rows() if self.world > 1 else []))
        rng = np.random.default_rng(1000 + self.rank)
        # candidates: the nodes of the first and last elements of this rank's range of every element type (the library
        # partitions each type by contiguous ranges, gfa_create) -- the partition interfaces when N > 1 -- then random ones
        cand = []
        slot = np.where((m.elem_type == M.BEAM_1) | (m.elem_type == M.PIPE_1), 1, np.where(m.elem_type == M.SHELL_1, 0, 2))
        for sl in range(3):
            idx = np.nonzero(slot == sl)[0]
            if len(idx) == 0:
                continue
            lo, hi = len(idx) * self.rank // self.world, len(idx) * (self.rank + 1) // self.world
            for e in list(idx[hi - 3:hi][::-1]) + list(idx[lo:lo + 3]):
                cand.extend(int(x) for x in en[ptr[e]:ptr[e + 1]])
        first_last = np.concatenate([np.array(cand, np.int64), rng.choice(touched, size=min(n_nodes, len(touched)), replace=False)])
        sample = []
        for nd in first_last:
            g = self.gls[nd - 1]
            free = g[g > 0]
            if len(free) == 0:
                continue
            if self.world > 1 and not all(int(r - 1) in owned for r in free):
                continue                                         # completed on another rank
            if int(nd) not in sample:
                sample.append(int(nd))
            if len(sample) >= n_nodes:
                break
        ok, worst = True, 0.0
        if sample:
            elem_of_entry = np.repeat(np.arange(m.n_elements), np.diff(ptr))
            elems = np.unique(elem_of_entry[np.isin(en, np.array(sample, np.int32))])
            sub, nodes = M.submodel(m, elems)             # no constraints: every DOF of the sample is a free row there
            dev = torch.cuda.current_device()
            ref = capi.Assembler(sub, device=dev).set_dofs()
            ref.gravity_factor = asm.gravity_factor
            ref.assemble(self.d_host.reshape(-1, 6)[nodes - 1])
            so, si, sv, _ = ref.csr("AA")
            sgl = ref.gls
            inv = {}                                             # sub free id -> (parent node, dof)
            for i, nd in enumerate(nodes):
                for k in range(6):
                    if sgl[i, k] > 0:
                        inv[int(sgl[i, k]) - 1] = (int(nd), k)
            lo, li = asm.csr_pattern("AA")
            lv = asm.values("AA")
            rows = asm.local_rows() if self.world > 1 else None
            pos = {int(r): i for i, r in enumerate(rows)} if rows is not None else None
            sub_index = {int(nd): i for i, nd in enumerate(nodes)}
            for nd in sample:
                for k in range(6):
                    g = int(self.gls[nd - 1, k])
                    if g <= 0:
                        continue
                    i = pos[g - 1] if pos is not None else g - 1
                    cols, vals = li[lo[i]:lo[i + 1]], lv[lo[i]:lo[i + 1]]
                    got = dict(zip(cols.tolist(), vals.tolist()))
                    srow = int(sgl[sub_index[nd], k]) - 1
                    diag_r = abs(sv[so[srow]:so[srow + 1]][si[so[srow]:so[srow + 1]] == srow][0])
                    for c, v in zip(si[so[srow]:so[srow + 1]].tolist(), sv[so[srow]:so[srow + 1]].tolist()):
                        pn, pk = inv[c]
                        gc = int(self.gls[pn - 1, pk])
                        if gc <= 0:
                            continue                             # fixed in the full model: lives in AB
                        w = got.get(gc - 1)
                        if w is None:
                            ok = False
                            continue
                        dd = sv[so[c]:so[c + 1]][si[so[c]:so[c + 1]] == c]
                        scale = 0.1 * float(np.sqrt(diag_r * abs(dd[0]))) if len(dd) else 0.0
                        err = abs(v - w) / max(abs(v), abs(w), scale, 1e-300)
                        worst = max(worst, err)
            ref.close()
            ok = ok and worst <= 1e-12
        if self.dist is not None:
            t = torch.tensor([0.0 if ok else 1.0, worst, float(len(sample))], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ok, worst = t[0].item() == 0.0, float(t[1].item())
        return {"parity_ok": bool(ok), "worst_relative_error": worst, "sampled_nodes_per_rank": len(sample),
                "what": "CSR rows of sampled nodes (first / last nodes of every rank's partition = its interfaces, plus random ones) vs a single-GPU assembly of the elements around them, tolerance 1e-12"}


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank (and the pinned host buffers it is about to allocate) to the NUMA node its GPU hangs off:
    eight ranks copying 4 GB each over PCIe otherwise cross the socket interconnect at random."""
    try:
        q = subprocess.run(["nvidia-smi", f"--id={local_rank}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=20)
        bdf = q.stdout.strip().lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:      # noqa: BLE001  (containers without sysfs access: keep the default placement)
        return None


def run_cuda(args):
    import torch
    from giraffe_b200 import capi  # noqa: F401

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the assembly path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cells = tuple(int(c) for c in args.cells.split("x"))
    m, d_host, cfg = workload(args.workload, world, args.scaling, cells)
    run = Runner(m, d_host, rank, world, local_rank, dist)
    n_el_total = m.n_elements
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_per_step = run.resident(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    launches = run.asm.launch_count() * args.steps + (2 * args.steps if world > 1 else 0)
    value = n_el_total / (ms_per_step * 1e-3)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_ms, h2d, d2h = run.e2e(e2e_steps)
    par = run.parity()
    ring, pipeline = run.asm.pipeline_info()
    nnz, nf, nx = run.nnz, run.nf, run.nx
    ev, sc = float(np.mean(run.eval_ms)), float(np.mean(run.scat_ms))
    owned_free = len(run.asm.owned_rows()) if world > 1 else nf
    alg_bytes = algorithmic_bytes(m, nnz[0], owned_free, world)
    setup_s = run.setup_s
    run.close()

    # ---- the other BASELINE configs, same protocol, fewer steps ------------------------------------------
    side = []
    if not args.no_side_configs and args.workload == "shell" and args.scaling == "weak":
        plan = [("beam", "weak"), ("solid", "weak"), ("mixed", "weak")] if world == 1 else [("shell", "strong"), ("mixed", "weak")]
        for kind, scaling in plan:
            try:
                sm, sd, scfg = workload(kind, world, scaling, cells)
                r2 = Runner(sm, sd, rank, world, local_rank, dist)
                steps2 = max(3, min(args.steps, 20))
                ms2 = r2.resident(steps2, 3)
                p2 = r2.parity(12)
                e2, s2 = float(np.mean(r2.eval_ms)), float(np.mean(r2.scat_ms))
                ofree = len(r2.asm.owned_rows()) if world > 1 else r2.nf
                ab = algorithmic_bytes(sm, r2.nnz[0], ofree, world)
                entry = {"config": scfg, "value": sm.n_elements / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2, "steps": steps2,
                         "kernels_ms_rank0": {"evaluation": e2, "scatter": s2}, "nnz_AA_rank0": r2.nnz[0], "setup_seconds": r2.setup_s,
                         "gpu_launches_per_step": r2.asm.launch_count(),
                         "hbm_roofline": {"achieved_gbs": ab / ((e2 + s2) * 1e-3) / 1e9, "algorithmic_bytes_per_step_rank0": ab},
                         "parity_ok": p2["parity_ok"], "parity_worst": p2["worst_relative_error"]}
                r2.close()
                side.append(entry)
            except Exception as e:      # noqa: BLE001  (a side config must not take the headline down)
                side.append({"config": {"workload": kind, "scaling": scaling}, "error": repr(e)[:300]})

    if rank == 0:
        pk, pk_src = peaks()
        # One step is two launches (element evaluation, then the CSR scatter); the algorithmic bytes of SURVEY.md 8(d)
        # belong to the pair, so the roofline is taken over both kernels' device time (CUDA events on the library's
        # stream, gfa_last_timing).
        dom = "shell::eval_kernel + scatter_kernel (one step = these two launches)" if not ring else "ring pipeline (evaluation + scatter kernels, co-resident)"
        dom_ms = ev + sc
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        tj = profile_file("traffic_r02.json") or profile_file("traffic_r01.json")
        traffic, traffic_src = None, None
        if tj and tj.get("elements") == n_el_total // world and tj.get("pipeline", "classic") == ("ring" if ring else "classic"):
            traffic = tj["dram_bytes_per_step"]
            traffic_src = tj.get("source", "profiles/traffic_r01.json") + " (ncu capture of this command, not re-measured in this run)"
        fpk = profile_file("fp64_peak_r01.json") or {}
        fp64_peak = float(fpk.get("fp64_fma_tflops", fpk.get("tflops", 34.16)))
        evp = profile_file("eval_pipe_r02.json") or {}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cb = time_cpu(3, 1, tuple(int(c) for c in args.cpu_cells.split("x")))
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu["mount_local_only_elements_per_s"] = cb["mount_local_only_elements_per_s"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "e2e": {"value": n_el_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms,
                    "what": "per rank: pinned H2D of the displacements of the nodes its elements reference, gfa_assemble, interface exchange, "
                            "D2H of its CSR values (AA,AB,BA,BB) and its owned rows of P_A, I_A (+P_B); bytes are whole-job sums"},
            "gpu_launches": launches,
            "clocks": clocks,
            "parity_ok": par["parity_ok"], "parity": par,
            "pipeline": pipeline,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": dom, "peak_source": pk_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms,
                         "note": "algorithmic bytes of the path on rank 0 (SURVEY.md 8d: compulsory reads per element + 8 B per CSR non-zero + 16 B per owned free DOF) "
                                 "over the device time of the step's kernels"},
            "fp64_roofline": {"bound": "fp64", "achieved": (n_el_total // world) * SHELL_ALG_FLOPS / (ev * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": (n_el_total // world) * SHELL_ALG_FLOPS / (ev * 1e-3) / 1e12 / fp64_peak, "kernel": "shell::eval_kernel",
                              "peak_source": "tools/fp64_peak.cu on this pool (profiles/fp64_peak_r01.json); MEASURED_PEAKS.json has no FP64 entry",
                              "ncu_fp64_pipe_busy_pct": evp.get("fp64_pipe_busy_pct"), "ncu_executed_fp64_thread_instructions_per_element": evp.get("fp64_thread_inst_per_element"),
                              "note": "numerator = 5.0e4 ALGORITHMIC flops per Shell_1 element (SURVEY.md 8d), more than the kernel executes: the structure-exploiting "
                                      "kernel issues fewer, so this fraction overstates pipe use -- ncu's pipe-busy figure beside it is the utilisation"} if args.workload == "shell" else None,
            "kernels_ms": {"evaluation": ev, "scatter": sc},
            "cpu_baseline": cpu,
            "side_configs": side,
            "setup_seconds": setup_s, "numa_node_rank0": numa_node,
            "nnz_AA": nnz[0], "n_free": nf,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="shell", choices=["shell", "beam", "solid", "mixed"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cells", default="1000x500", help="per-GPU plate size in cells (2 Shell_1 per cell)")
    ap.add_argument("--cpu-cells", default="100x50", help="plate of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true")
    ap.add_argument("--no-full-size-step", action="store_true", help="reference arm: skip the one 1M-element step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
