#!/usr/bin/env python
"""Side measurements for the other BASELINE.json configs (not the driver's bench
contract): device-resident assembly time of the 100k Beam_1 line, a Solid_1
block and a mixed model, through the C-ABI.  Prints one JSON line per config."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from giraffe_b200 import capi, meshes as M  # noqa: E402


def run(name, m, d, steps=30):
    import torch
    asm = capi.Assembler(m).set_dofs()
    dd = torch.from_numpy(np.ascontiguousarray(d).reshape(-1)).cuda()
    for _ in range(3):
        asm.assemble(None, device_ptr=dd.data_ptr())
    ev, sc, tot = [], [], []
    for _ in range(steps):
        asm.assemble(None, device_ptr=dd.data_ptr())
        t = asm.timing()
        ev.append(t["eval_ms"]); sc.append(t["scatter_ms"]); tot.append(t["total_ms"])
    ms = float(np.median(tot))
    print(json.dumps({"config": name, "elements": m.n_elements, "n_free": asm.n_free, "nnz_AA": asm.csr_dims("AA")[2],
                      "ms_per_step": ms, "elements_per_s": m.n_elements / (ms * 1e-3),
                      "eval_ms": float(np.median(ev)), "scatter_ms": float(np.median(sc)), "launches": asm.launch_count()}))
    asm.close()


if __name__ == "__main__":
    b = M.beam_line(100_000)
    run("configs[1]: 100k Beam_1 line", b, M.beam_line_displacements(b))
    v = M.solid_block(160, 160, 156)     # 3,993,600 hexahedra
    run("configs[3]: 4M Solid_1 block (builder-defined hexahedron)", v, M.solid_block_displacements(v), steps=10)
    del v
    mix = M.mixed_model(100_000, 400, 375, (100, 100, 60))
    d = M.mask_displacements(mix, np.random.default_rng(4).uniform(-1e-5, 1e-5, (mix.n_nodes, 6)))
    run("configs[4] (single-GPU slice): 100k Beam_1 + 300k Shell_1 + 600k Solid_1", mix, d, steps=10)
