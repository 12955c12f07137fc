#!/usr/bin/env python
"""Timing probe of the assembly pipelines on one GPU (experiments; not the bench).
usage: python tools/ring_probe.py [cells=1000x500] [steps=20] KEY=VAL ...  (KEY=VAL are GFA_* environment settings)
Prints one line per run: pipeline description, ms per step (CUDA events over `steps` back-to-back steps), phases."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from giraffe_b200 import capi, meshes as M

def main():
    cells, steps, kind = (1000, 500), 20, "shell"
    for a in sys.argv[1:]:
        if a.startswith("cells="): cells = tuple(int(c) for c in a[6:].split("x"))
        elif a.startswith("steps="): steps = int(a[6:])
        elif a.startswith("kind="): kind = a[5:]
        elif "=" in a: k, v = a.split("=", 1); os.environ[k] = v
    if kind == "shell":
        m = M.shell_plate(*cells); d = M.shell_plate_displacements(m)
    elif kind == "beam":
        m = M.beam_line(cells[0]); d = M.beam_line_displacements(m)
    elif kind == "solid":
        m = M.solid_block(*cells); d = M.solid_block_displacements(m)
    else:
        m = M.mixed_model(*eval(kind)); d = M.mask_displacements(m, np.random.default_rng(1).uniform(-1e-4, 1e-4, (m.n_nodes, 6)))
    t0 = time.time()
    asm = capi.Assembler(m)
    gls, nf, nx = M.number_dofs(m)
    asm.set_dofs(gls, nf, nx)
    setup = time.time() - t0
    ring, note = asm.pipeline_info()
    dev = torch.from_numpy(np.ascontiguousarray(d).reshape(-1)).cuda()
    st = torch.cuda.ExternalStream(asm.stream())
    for _ in range(3):
        asm.assemble(None, device_ptr=dev.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    ph = np.zeros(4)
    for _ in range(steps):
        asm.assemble(None, device_ptr=dev.data_ptr())
        t = asm.timing(); ph += [t["h2d_ms"], t["eval_ms"], t["scatter_ms"], t["total_ms"]]
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ph /= steps
    chk = float(np.sum(asm.vectors()[0])) 
    print(f"RESULT {kind} {m.n_elements} el: {ms:.3f} ms/step = {m.n_elements / ms / 1e3:.1f} Mel/s | phases h2d {ph[0]:.3f} [1] {ph[1]:.3f} [2] {ph[2]:.3f} total {ph[3]:.3f} | launches {asm.launch_count()} | setup {setup:.1f}s | sumPA {chk:.6e} | {note} | env " +
          " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("GFA_")), flush=True)

if __name__ == "__main__":
    main()
