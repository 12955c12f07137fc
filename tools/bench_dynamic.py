#!/usr/bin/env python
"""Side measurement (not the driver's bench contract): device time of one Newmark-dynamics assembly
(gfa_assemble_dynamic: MountLocal .. MountMass, MountDamping, MountDyn .. MountSparse) next to the static
one, on the Shell_1 plate of BASELINE.json configs[2] and the Beam_1 line of configs[1].  One JSON line each."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from giraffe_b200 import capi, meshes as M  # noqa: E402


def newmark(dt, beta=0.3, gamma=0.5):
    return [1.0 / (dt * dt * beta), 1.0 / (dt * beta), 1.0 / (2.0 * beta) - 1.0, gamma / (dt * beta), 1.0 - gamma / beta,
            dt * (1.0 - gamma / (2.0 * beta))]


def run(name, m, d, steps=20):
    asm = capi.Assembler(m).set_dofs()
    rng = np.random.default_rng(1)
    asm.set_dynamic(newmark(1e-3), 0.5, 1e-4)
    asm.set_kinematics(None, None, rng.uniform(-1, 1, (m.n_nodes, 6)), rng.uniform(-10, 10, (m.n_nodes, 6)))
    asm.update_dyn(d)
    asm.assemble(None)
    out = {"config": name, "elements": m.n_elements}
    for label, fn in (("static", lambda: asm.assemble(None)), ("dynamic_update_rayleigh", lambda: asm.assemble_dynamic(None, True)),
                      ("dynamic", lambda: asm.assemble_dynamic(None, False))):
        for _ in range(3):
            fn()
        tot, ev = [], []
        for _ in range(steps):
            fn()
            t = asm.timing()
            tot.append(t["total_ms"]); ev.append(t["eval_ms"])
        out[label + "_ms"] = float(np.median(tot))
        out[label + "_eval_ms"] = float(np.median(ev))
    out["dynamic_elements_per_s"] = m.n_elements / (out["dynamic_ms"] * 1e-3)
    print(json.dumps(out))
    asm.close()


if __name__ == "__main__":
    b = M.beam_line(100_000)
    run("configs[1]: 100k Beam_1 line", b, M.beam_line_displacements(b))
    s = M.shell_plate(1000, 500)
    run("configs[2]: 1M Shell_1 plate", s, M.shell_plate_displacements(s))
