#!/usr/bin/env python
"""Benchmark of the per-Newton-iteration element assembly (BASELINE.json metric:
elements assembled/sec, Kt+Fint -> CSR, FP64).

    python bench.py --gpus N --steps K --warmup W            # CUDA path, C-ABI
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU path

A "step" is one pass of the hot path (Clear + MountLocal + MountElementLoads +
MountGlobal + MountSparse) over one synthetic batch: the 1M-element Shell_1
plate of BASELINE.json configs[2] per GPU (weak scaling: N GPUs assemble an
N-times larger plate, partitioned by contiguous element ranges; interface rows
are exchanged with NCCL send/recv).  Prints ONE JSON line on rank 0.
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
SHELL_READ_BYTES = 928.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


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


def workload(n_gpus: int, per_gpu_cells=(1000, 500)):
    """N x (1000 x 500 cells x 2 triangles): the plate grows along y so that
    contiguous element ranges are strips with one interface line each."""
    nx, ny = per_gpu_cells
    return M.shell_plate(nx, ny * n_gpus), {"workload": f"shell_plate_{nx}x{ny * n_gpus}cells_Shell_1",
                                           "elements": 2 * nx * ny * n_gpus, "per_gpu_elements": 2 * nx * ny,
                                           "config": "BASELINE.json configs[2] (1M-element Shell_1 plate) per GPU",
                                           "partition": f"{n_gpus} strips by contiguous element range",
                                           "l2": "inputs larger than L2 (element blocks 3.6 GB, CSR values 4.2 GB per GPU)"}


# ---------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path
# ---------------------------------------------------------------------------
def cpu_sample(cells=(100, 50)):
    m = M.shell_plate(*cells)
    return m, M.shell_plate_displacements(m), f"Shell_1 plate {cells[0]}x{cells[1]} cells = {m.n_elements} elements of the same mesh family"


def cpu_oracle(threads: int):
    from oracle import refdrv
    if refdrv.available():
        return refdrv.RefOracle(threads=threads), "reference"
    from oracle.portdrv import PortOracle
    return PortOracle(threads=threads), "port"


def time_cpu(steps: int, warmup: int, cells=(100, 50)):
    threads = os.cpu_count() or 1
    orc, kind = cpu_oracle(threads)
    m, d, sample = cpu_sample(cells)
    orc.load(m)
    for _ in range(warmup):
        orc.assemble(d)
    t = []
    local = []
    for _ in range(steps):
        s = orc.assemble(d)
        t.append(float(s[:4].sum()))       # MountLocal + MountElementLoads + MountGlobal + MountSparse
        local.append(float(s[0]))
    med = float(np.median(t))
    return {"value": m.n_elements / med, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
            "ms_per_step": med * 1e3, "mount_local_only_elements_per_s": m.n_elements / float(np.median(local)),
            "note": "OpenMP MountLocal/MountElementLoads, serial MountGlobal + setFromTriplets as in the reference; "
                    "GEMM and setFromTriplets are restatements (no MKL/Eigen on the box)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = time_cpu(max(args.steps, 1), max(args.warmup, 1))
    _, cfg = workload(args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": cb["note"] + "; each step is a bounded sample of the workload (throughput is size-independent: the path is O(elements))"}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------
def run_cuda(args):
    import torch
    from giraffe_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the assembly path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cells = tuple(int(c) for c in args.cells.split("x"))
    m, cfg = workload(world, cells)
    d_host = M.shell_plate_displacements(m)
    t0 = time.time()
    asm = capi.Assembler(m, device=local_rank, rank=rank, world=world)
    gls, nf, nx = M.number_dofs(m)
    asm.set_dofs(gls, nf, nx)
    setup_s = time.time() - t0
    n_el_total = m.n_elements

    pinned = torch.empty(d_host.size, dtype=torch.float64).pin_memory()
    pinned.numpy()[:] = d_host.reshape(-1)
    d_dev = pinned.cuda(non_blocking=False)
    lib_stream = torch.cuda.ExternalStream(asm.stream())
    if_stream = torch.cuda.ExternalStream(asm.interface_stream())

    # interface exchange buffers (N > 1)
    send_cnt, recv_cnt = asm.interface_counts(world)
    send_buf = torch.empty(int(send_cnt.sum()), dtype=torch.float64, device="cuda") if world > 1 else None
    recv_buf = torch.empty(int(recv_cnt.sum()), dtype=torch.float64, device="cuda") if world > 1 else None

    def exchange():
        """pack -> NCCL send/recv -> unpack, stream-ordered on the library's interface stream (no host syncs);
        the library scatters the interface rows first, so the exchange overlaps the interior rows' scatter."""
        if world == 1:
            return
        with torch.cuda.stream(if_stream):
            asm.interface_pack(send_buf.data_ptr())
            ops, so, ro = [], 0, 0
            for r in range(world):
                if send_cnt[r]:
                    ops.append(dist.P2POp(dist.isend, send_buf[so:so + int(send_cnt[r])], r))
                if recv_cnt[r]:
                    ops.append(dist.P2POp(dist.irecv, recv_buf[ro:ro + int(recv_cnt[r])], r))
                so += int(send_cnt[r]); ro += int(recv_cnt[r])
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            asm.interface_unpack(recv_buf.data_ptr())

    def step_resident():
        if world > 1:
            # enqueue only: the host queues pack / NCCL / unpack behind the assembly instead of leaving the
            # GPU idle while it catches up; reads of the results wait for the stream
            asm.assemble_enqueue(d_dev.data_ptr())
        else:
            asm.assemble(None, device_ptr=d_dev.data_ptr())
        exchange()

    # results land here in the end-to-end leg (what the host-side solver consumes)
    nnz = [asm.csr_dims(w)[2] for w in ("AA", "AB", "BA", "BB")]
    out_vals = [torch.empty(max(n, 1), dtype=torch.float64).pin_memory() for n in nnz]
    out_vecs = [torch.empty(max(n, 1), dtype=torch.float64).pin_memory() for n in (nf, nf, nx)]

    def step_e2e():
        asm.assemble_raw(pinned.data_ptr())          # H2D of the displacements inside the call
        exchange()
        for w, buf in zip(("AA", "AB", "BA", "BB"), out_vals):
            asm.values(w, out=buf.numpy()[:asm.csr_dims(w)[2]])
        for w, buf, n in zip((capi.P_A, capi.I_A, capi.P_B), out_vecs, (nf, nf, nx)):
            asm.vector(w, buf.numpy()[:n])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Device time of `steps` calls: CUDA events on the library's stream,
        bracketed by barrier + synchronize; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        w0 = time.perf_counter()
        for _ in range(steps):
            fn()
        e1.record(lib_stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        barrier()
        ms = e0.elapsed_time(e1)
        # host-side pieces (exchange waits, D2H) are not on the library stream: take the larger
        ms = max(ms, wall * 1e3) if (world > 1 or fn is step_e2e) else ms
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eval_ms, scat_ms = [], []

    def step_resident_logged():
        step_resident()
        if world == 1:
            t = asm.timing()
            eval_ms.append(t["eval_ms"]); scat_ms.append(t["scatter_ms"])

    def sample_kernel_times():
        """N > 1: the timed loop does not read per-kernel times (that would wait for every step)."""
        for _ in range(3):
            step_resident()
            t = asm.timing()
            eval_ms.append(t["eval_ms"]); scat_ms.append(t["scatter_ms"])

    ms = timed(step_resident_logged, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        sample_kernel_times()
    launches = asm.launch_count() * args.steps + (2 * args.steps if world > 1 else 0)
    ms_per_step = ms / args.steps
    value = n_el_total / (ms_per_step * 1e-3)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    step_e2e()
    e2e_ms = timed(step_e2e, e2e_steps) / e2e_steps
    h2d = d_host.size * 8
    d2h = 8 * (sum(nnz) + 2 * nf + nx)

    if rank == 0:
        pk, pk_src = peaks()
        n_local = n_el_total // world
        nnz_local = nnz[0]                      # csr_dims are this rank's stored rows
        alg_bytes = n_local * (SHELL_READ_BYTES) + 8.0 * nnz_local + 16.0 * nf / world
        ev, sc = float(np.mean(eval_ms)), float(np.mean(scat_ms))
        # The path is two launches per step (element evaluation, then the CSR gather); the
        # algorithmic bytes of SURVEY.md 8(d) belong to the pair, so the roofline is taken over
        # both kernels' device time (CUDA events on the library's stream, gfa_last_timing).
        dom = "shell::eval_kernel + scatter_kernel (one step = these two launches)"
        dom_ms = ev + sc
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("elements") == n_local:
                traffic = tj["dram_bytes_per_step"]
        fp64_peak = 34.16      # measured on this pool with tools/fp64_peak.cu (profiles/fp64_peak_r01.json)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cb = time_cpu(3, 1)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu["mount_local_only_elements_per_s"] = cb["mount_local_only_elements_per_s"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "e2e": {"value": n_el_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "what": "gfa_assemble from pinned host displacements + D2H of all CSR values (AA,AB,BA,BB) and P_A,I_A,P_B"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": traffic, "kernel": dom, "peak_source": pk_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms,
                         "note": "algorithmic bytes of the path (SURVEY.md 8d: 928 B read/element + 8 B per CSR non-zero + 16 B per free DOF); traffic = ncu dram bytes of both launches"},
            "fp64_roofline": {"bound": "fp64", "achieved": n_local * SHELL_ALG_FLOPS / (ev * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": n_local * SHELL_ALG_FLOPS / (ev * 1e-3) / 1e12 / fp64_peak, "kernel": "shell::eval_kernel",
                              "note": "5.0e4 algorithmic flops per Shell_1 element (SURVEY.md 8d) over the evaluation kernel; peak = measured FP64 FMA rate"},
            "kernels_ms": {"shell_eval": ev, "scatter": sc},
            "cpu_baseline": cpu,
            "setup_seconds": setup_s,
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
    ap.add_argument("--cells", default="1000x500", help="per-GPU plate size in cells (2 Shell_1 per cell)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
