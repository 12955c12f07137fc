"""Reader for the subset of GIRAFFE's ``.inp`` grammar that the assembly path
needs (reference ``IO.cpp:286-677``).

Tokens are whitespace separated; ``//`` and ``/* */`` comments may appear
between records; a block is ``Keyword N`` followed by N records.  Blocks the
assembly path does not consume (SolverOptions, Monitor, PostFiles,
ConvergenceCriteria, ...) are skipped up to the next known keyword.
Supported: Nodes, Elements (Beam_1 / Shell_1 / Solid_1), Materials (Hooke),
Sections (Rectangle / Tube), ShellSections (Homogeneous), CoordinateSystems,
NodeSets (List / Sequence), Constraints (NodalConstraint), Loads (NodalLoad
with a numeric table), Environment (GravityData), SolutionSteps (Static, for
the time-stepping data only).
"""
from __future__ import annotations

import re

import numpy as np

from .meshes import BEAM_1, SHELL_1, SOLID_1, Model, _finish

_TOP = {"Nodes", "Elements", "Materials", "Sections", "ShellSections", "CoordinateSystems", "NodeSets",
        "Constraints", "Loads", "Environment", "SolutionSteps", "SolverOptions", "Monitor", "PostFiles",
        "ConvergenceCriteria", "ElementSets", "ExecutionData", "EOF"}


def _tokens(text: str):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return text.split()


def read_inp(path: str):
    """Return (Model, info) where info holds node sets and the first Static
    step's time data ({'end_time','time_step'})."""
    with open(path, "r", errors="replace") as f:
        tk = _tokens(f.read())
    i = 0
    nodes, mats, secs, shsecs, csd = {}, {}, {}, {}, {}
    elems, nodesets, cons, loads = [], {}, [], []
    gravity = None
    info = {}

    def num(j):
        return float(tk[j])

    while i < len(tk):
        kw = tk[i]
        if kw == "EOF":
            break
        if kw == "Nodes":
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                assert tk[i] == "Node"
                nodes[int(tk[i + 1])] = (num(i + 2), num(i + 3), num(i + 4)); i += 5
        elif kw == "Materials":
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                assert tk[i] == "Hooke", f"unsupported material {tk[i]}"
                mats[int(tk[i + 1])] = (num(i + 3), num(i + 5), num(i + 7)); i += 8
        elif kw == "Sections":
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                kind = {"Rectangle": 0, "Tube": 1}.get(tk[i])
                assert kind is not None, f"unsupported section {tk[i]}"
                secs[int(tk[i + 1])] = (kind, num(i + 3), num(i + 5)); i += 6
                if i < len(tk) and tk[i] == "AD":
                    i += 7
        elif kw == "ShellSections":
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                assert tk[i] == "Homogeneous", f"unsupported shell section {tk[i]}"
                shsecs[int(tk[i + 1])] = num(i + 3); i += 4
        elif kw == "CoordinateSystems":
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                assert tk[i] == "CS"
                csd[int(tk[i + 1])] = ((num(i + 3), num(i + 4), num(i + 5)), (num(i + 7), num(i + 8), num(i + 9))); i += 10
        elif kw == "NodeSets":
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                assert tk[i] == "NodeSet"
                sid, cnt = int(tk[i + 1]), int(tk[i + 3])
                if tk[i + 4] == "List":
                    nodesets[sid] = [int(t) for t in tk[i + 5:i + 5 + cnt]]; i += 5 + cnt
                else:   # Sequence Initial a Increment k
                    a, k = int(tk[i + 6]), int(tk[i + 8])
                    nodesets[sid] = [a + k * q for q in range(cnt)]; i += 9
        elif kw == "Elements":
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                ty = tk[i]
                if ty == "Beam_1":
                    e = dict(type=BEAM_1, mat=int(tk[i + 3]), sec=int(tk[i + 5]), cs=int(tk[i + 7]),
                             nodes=[int(t) for t in tk[i + 9:i + 12]], T0=0.0)
                    i += 12
                    if i < len(tk) and tk[i] == "PreTension":
                        e["T0"] = num(i + 1); i += 2
                elif ty == "Shell_1":
                    e = dict(type=SHELL_1, mat=int(tk[i + 3]), sec=int(tk[i + 5]), cs=0, T0=0.0)
                    i += 6
                    if tk[i] == "CS":
                        e["cs"] = int(tk[i + 1]); i += 2
                    e["nodes"] = [int(t) for t in tk[i + 1:i + 7]]; i += 7
                elif ty == "Solid_1":
                    e = dict(type=SOLID_1, mat=int(tk[i + 3]), sec=0, cs=int(tk[i + 5]),
                             nodes=[int(t) for t in tk[i + 7:i + 15]], T0=0.0)
                    i += 15
                else:
                    raise ValueError(f"element type {ty} is outside the accelerated path")
                elems.append(e)
        elif kw == "Constraints":
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                assert tk[i] == "NodalConstraint"
                sid = int(tk[i + 3]); i += 4
                mask = 0
                names = ["UX", "UY", "UZ", "ROTX", "ROTY", "ROTZ"]
                while i < len(tk) and tk[i] in names:
                    k = names.index(tk[i]); i += 2          # keyword + 'BoolTable'
                    first = None
                    while i < len(tk) and tk[i][0].isdigit():
                        if first is None:
                            first = int(tk[i])
                        i += 1
                    if first == 1:
                        mask |= 1 << k
                cons.append((sid, mask))
        elif kw == "Loads":
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                assert tk[i] == "NodalLoad", f"load {tk[i]} stays on the host"
                sid, cs, nt = int(tk[i + 3]), int(tk[i + 5]), int(tk[i + 7]); i += 8
                table = np.array([float(t) for t in tk[i:i + 7 * nt]]).reshape(nt, 7); i += 7 * nt
                loads.append((sid, cs, table))
        elif kw == "Environment":
            i += 1
            if tk[i] == "GravityData":
                gravity = (num(i + 2), num(i + 3), num(i + 4)); i += 5
                if tk[i] == "BoolTable":
                    i += 1
                    while i < len(tk) and tk[i][0].isdigit():
                        i += 1
        elif kw == "SolutionSteps":
            i += 2
            if tk[i] == "Static":
                vals = {tk[j]: tk[j + 1] for j in range(i + 2, i + 20, 2)}
                info["end_time"] = float(vals["EndTime"]); info["time_step"] = float(vals["TimeStep"])
                i += 20
        else:
            i += 1
            while i < len(tk) and tk[i] not in _TOP:
                i += 1

    nn = max(nodes)
    xyz = np.array([nodes[k] for k in range(1, nn + 1)], float)
    m = Model(xyz=xyz, hooke=np.array([mats[k] for k in sorted(mats)], float), sections=np.zeros((0, 6)))
    m.section_defs = [secs[k] for k in sorted(secs)]
    m.shell_thickness = np.array([shsecs[k] for k in sorted(shsecs)], float)
    m.cs_defs = [csd[k] for k in sorted(csd)]
    m.elem_type = np.array([e["type"] for e in elems], np.int32)
    m.elem_mat = np.array([e["mat"] for e in elems], np.int32)
    m.elem_sec = np.array([e["sec"] for e in elems], np.int32)
    m.elem_cs = np.array([e["cs"] for e in elems], np.int32)
    m.elem_nodes = np.array([n for e in elems for n in e["nodes"]], np.int32)
    pre = np.array([e["T0"] for e in elems], float)
    m.pretension = pre if np.any(pre != 0.0) else None
    m.constraints = [(np.array(nodesets[s], np.int32), mask) for s, mask in cons]
    m.nodal_loads = [(np.array(nodesets[s], np.int32), cs, t) for s, cs, t in loads]
    m.gravity = gravity
    info["node_sets"] = nodesets
    return _finish(m), info
