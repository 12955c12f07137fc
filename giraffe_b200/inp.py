"""Reader for the subset of GIRAFFE's ``.inp`` grammar that the assembly path
needs (reference ``IO.cpp:286-677``).

Tokens are whitespace separated; ``//`` and ``/* */`` comments may appear
between records; a block is ``Keyword N`` followed by N records.  Blocks the
assembly path does not consume (SolverOptions, Monitor, PostFiles,
ConvergenceCriteria, ...) are skipped up to the next known keyword.
Supported: Nodes, Elements (Beam_1 / Pipe_1 / Shell_1 / Solid_1), Materials (Hooke),
Sections (Rectangle / Tube), PipeSections (PS), ShellSections (Homogeneous), CoordinateSystems,
NodeSets / ElementSets (List / Sequence), Constraints (NodalConstraint), Loads (NodalLoad,
ShellLoad with a numeric table), Environment (GravityData), SolutionSteps (Static / Dynamic, for the
time-stepping data, Rayleigh and Newmark coefficients only).
"""
from __future__ import annotations

import re

import numpy as np

from .meshes import BEAM_1, PIPE_1, SHELL_1, SOLID_1, Model, _finish

_TOP = {"Nodes", "Elements", "Materials", "Sections", "PipeSections", "ShellSections", "CoordinateSystems", "NodeSets",
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
    nodes, mats, secs, shsecs, csd, pipes = {}, {}, {}, {}, {}, {}
    elems, nodesets, cons, loads = [], {}, [], []
    elemsets, shloads, ploads, floads = {}, [], [], []
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
        elif kw == "PipeSections":          # PS id EA v EI v GJ v GA v Rho v CDt v CDn v CAt v CAn v De v Di v (PipeSection.cpp Read)
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                assert tk[i] == "PS", f"unsupported pipe section record {tk[i]}"
                names = [tk[i + 2 + 2 * k] for k in range(11)]
                assert names == ["EA", "EI", "GJ", "GA", "Rho", "CDt", "CDn", "CAt", "CAn", "De", "Di"], names
                pipes[int(tk[i + 1])] = tuple(num(i + 3 + 2 * k) for k in range(11)); i += 24
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
                elif ty == "Pipe_1":        # Pipe_1 id PipeSec s CS c Nodes a b c (Pipe_1.cpp:535-572)
                    assert tk[i + 2] == "PipeSec" and tk[i + 4] == "CS" and tk[i + 6] == "Nodes"
                    e = dict(type=PIPE_1, mat=0, sec=int(tk[i + 3]), cs=int(tk[i + 5]),
                             nodes=[int(t) for t in tk[i + 7:i + 10]], T0=0.0)
                    i += 10
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
                if tk[i] == "ShellLoad":      # ShellLoad id ElementSet s AreaUpdate b NTimes n (ShellLoad.cpp:28-87)
                    es, au, nt = int(tk[i + 3]), int(tk[i + 5]), int(tk[i + 7]); i += 8
                    table = np.array([float(t) for t in tk[i:i + 2 * nt]]).reshape(nt, 2); i += 2 * nt
                    shloads.append((es, bool(au), table))
                    continue
                if tk[i] == "PipeLoad":       # PipeLoad id ElementSet s NTimes n, rows time P0I P0E RhoI RhoE (PipeLoad.cpp:44-88)
                    es, nt = int(tk[i + 3]), int(tk[i + 5]); i += 6
                    table = np.array([float(t) for t in tk[i:i + 5 * nt]]).reshape(nt, 5); i += 5 * nt
                    ploads.append((es, table))
                    continue
                assert tk[i] in ("NodalLoad", "NodalFollowerLoad"), f"load {tk[i]} is outside the subset"
                follower = tk[i] == "NodalFollowerLoad"      # same format (NodalFollowerLoad.cpp:56-112)
                sid, cs, nt = int(tk[i + 3]), int(tk[i + 5]), int(tk[i + 7]); i += 8
                table = np.array([float(t) for t in tk[i:i + 7 * nt]]).reshape(nt, 7); i += 7 * nt
                (floads if follower else loads).append((sid, cs, table))
        elif kw == "ElementSets":        # ElementSet id Elements n List ... | Sequence Initial a Increment k (ElementSet.cpp:27-95)
            n = int(tk[i + 1]); i += 2
            for _ in range(n):
                assert tk[i] == "ElementSet"
                sid, cnt = int(tk[i + 1]), int(tk[i + 3])
                if tk[i + 4] == "List":
                    elemsets[sid] = [int(t) for t in tk[i + 5:i + 5 + cnt]]; i += 5 + cnt
                else:
                    a, inc = int(tk[i + 6]), int(tk[i + 8])
                    elemsets[sid] = [a + k * inc for k in range(cnt)]; i += 9
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
            elif tk[i] == "Dynamic":        # Dynamic::Read (Dynamic.cpp:65-222)
                vals = {tk[j]: tk[j + 1] for j in range(i + 2, i + 20, 2)}
                info["end_time"] = float(vals["EndTime"]); info["time_step"] = float(vals["TimeStep"])
                j = i + 20
                assert tk[j] == "RayleighDamping" and tk[j + 7] == "NewmarkCoefficients", "Dynamic step: RayleighDamping / NewmarkCoefficients expected"
                info["dynamic"] = {"alpha": float(tk[j + 2]), "beta": float(tk[j + 4]), "update": int(tk[j + 6]),
                                   "beta_new": float(tk[j + 9]), "gamma_new": float(tk[j + 11])}
                i = j + 12
        else:
            i += 1
            while i < len(tk) and tk[i] not in _TOP:
                i += 1

    nn = max(nodes)
    xyz = np.array([nodes[k] for k in range(1, nn + 1)], float)
    m = Model(xyz=xyz, hooke=np.array([mats[k] for k in sorted(mats)], float).reshape(-1, 3), sections=np.zeros((0, 6)))
    m.pipe_sections = np.array([pipes[k] for k in sorted(pipes)], float).reshape(-1, 11)
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
    m.shell_loads = [(np.array(elemsets[es], np.int32), au, t) for es, au, t in shloads]
    m.pipe_loads = [(np.array(elemsets[es], np.int32), t) for es, t in ploads]
    m.follower_loads = [(np.array(nodesets[s], np.int32), cs, t) for s, cs, t in floads]
    info["node_sets"] = nodesets
    return _finish(m), info


def _r(v) -> str:
    return repr(float(v))


def write_inp(m: Model, path: str, end_time: float = 1.0, time_step: float = 1.0, dynamic: dict | None = None) -> None:
    """Write a Model in the reference's ``.inp`` syntax, e.g. to
    hand a synthetic mesh to a GIRAFFE build or to the C++ host mirror.  `dynamic` =
    {alpha, beta, update, beta_new, gamma_new} writes a Dynamic solution step (Dynamic.cpp:65-222)."""
    names = ["UX", "UY", "UZ", "ROTX", "ROTY", "ROTZ"]
    with open(path, "w") as f:
        f.write(f"Nodes\t{m.n_nodes}\n")
        for i, x in enumerate(m.xyz):
            f.write(f"Node\t{i + 1}\t{_r(x[0])}\t{_r(x[1])}\t{_r(x[2])}\n")
        follower_loads = getattr(m, "follower_loads", [])
        sets = [np.asarray(n) for n, _ in m.constraints] + [np.asarray(n) for n, _, _ in m.nodal_loads] + [np.asarray(n) for n, _, _ in follower_loads]
        if sets:
            f.write(f"\nNodeSets\t{len(sets)}\n")
            for k, s in enumerate(sets):
                f.write(f"NodeSet\t{k + 1}\tNodes\t{len(s)}\tList\t" + "\t".join(str(int(v)) for v in s) + "\n")
        f.write(f"\nElements\t{m.n_elements}\n")
        for e in range(m.n_elements):
            nd = "\t".join(str(int(v)) for v in m.elem_nodes[m.elem_ptr[e]:m.elem_ptr[e + 1]])
            t = int(m.elem_type[e])
            if t == BEAM_1:
                f.write(f"Beam_1\t{e + 1}\tMat\t{m.elem_mat[e]}\tSec\t{m.elem_sec[e]}\tCS\t{m.elem_cs[e]}\tNodes\t{nd}")
                if m.pretension is not None and m.pretension[e] != 0.0:
                    f.write(f"\tPreTension\t{float(m.pretension[e])!r}")
                f.write("\n")
            elif t == PIPE_1:
                f.write(f"Pipe_1\t{e + 1}\tPipeSec\t{m.elem_sec[e]}\tCS\t{m.elem_cs[e]}\tNodes\t{nd}\n")
            elif t == SHELL_1:
                f.write(f"Shell_1\t{e + 1}\tMat\t{m.elem_mat[e]}\tSec\t{m.elem_sec[e]}\tNodes\t{nd}\n")
            else:
                f.write(f"Solid_1\t{e + 1}\tMat\t{m.elem_mat[e]}\tCS\t{m.elem_cs[e]}\tNodes\t{nd}\n")
        if len(m.hooke):
            f.write(f"\nMaterials\t{len(m.hooke)}\n")
            for k, (E, nu, rho) in enumerate(m.hooke):
                f.write(f"Hooke\t{k + 1}\tE\t{float(E)!r}\tNu\t{float(nu)!r}\tRho\t{float(rho)!r}\n")
        if len(m.pipe_sections):
            f.write(f"\nPipeSections\t{len(m.pipe_sections)}\n")
            for k, row in enumerate(np.asarray(m.pipe_sections, float).reshape(-1, 11)):
                f.write(f"PS\t{k + 1}\t" + "\t".join(f"{nm}\t{_r(v)}" for nm, v in zip(("EA", "EI", "GJ", "GA", "Rho", "CDt", "CDn", "CAt", "CAn", "De", "Di"), row)) + "\n")
        if m.section_defs:
            f.write(f"\nSections\t{len(m.section_defs)}\n")
            for k, (kind, a, b) in enumerate(m.section_defs):
                f.write((f"Rectangle\t{k + 1}\tB\t{_r(a)}\tH\t{_r(b)}\n") if kind == 0 else (f"Tube\t{k + 1}\tDe\t{_r(a)}\tDi\t{_r(b)}\n"))
        if len(m.shell_thickness):
            f.write(f"\nShellSections\t{len(m.shell_thickness)}\n")
            for k, t in enumerate(m.shell_thickness):
                f.write(f"Homogeneous\t{k + 1}\tThickness\t{float(t)!r}\n")
        if m.cs_defs:
            f.write(f"\nCoordinateSystems\t{len(m.cs_defs)}\n")
            for k, (e1, e3) in enumerate(m.cs_defs):
                f.write(f"CS\t{k + 1}\tE1\t{_r(e1[0])}\t{_r(e1[1])}\t{_r(e1[2])}\tE3\t{_r(e3[0])}\t{_r(e3[1])}\t{_r(e3[2])}\n")
        f.write(f"\nSolutionSteps\t1\n{'Dynamic' if dynamic else 'Static'}\t1\nEndTime\t{_r(end_time)}\nTimeStep\t{_r(time_step)}\nMaxTimeStep\t{_r(time_step)}\n"
                "MinTimeStep\t0.001\nMaxIt\t20\nMinIt\t3\nConvIncrease\t4\nIncFactor\t1.0\nSample\t1\n")
        if dynamic:
            f.write(f"RayleighDamping\tAlpha\t{_r(dynamic['alpha'])}\tBeta\t{_r(dynamic['beta'])}\tUpdate\t{int(dynamic['update'])}\n"
                    f"NewmarkCoefficients\tBeta\t{_r(dynamic['beta_new'])}\tGamma\t{_r(dynamic['gamma_new'])}\n")
        pipe_loads = getattr(m, "pipe_loads", [])
        if m.shell_loads or pipe_loads:
            f.write(f"\nElementSets\t{len(m.shell_loads) + len(pipe_loads)}\n")
            for k, elements in enumerate([l[0] for l in m.shell_loads] + [l[0] for l in pipe_loads]):
                f.write(f"ElementSet\t{k + 1}\tElements\t{len(elements)}\tList\t" + "\t".join(str(int(e)) for e in elements) + "\n")
        if m.nodal_loads or m.shell_loads or pipe_loads or follower_loads:
            f.write(f"\nLoads\t{len(m.nodal_loads) + len(m.shell_loads) + len(pipe_loads) + len(follower_loads)}\n")
            for k, (nodes, cs, table) in enumerate(m.nodal_loads):
                table = np.asarray(table, float)
                f.write(f"NodalLoad\t{k + 1}\tNodeSet\t{len(m.constraints) + k + 1}\tCS\t{cs}\tNTimes\t{len(table)}\n")
                for row in table:
                    f.write("\t".join(repr(float(v)) for v in row) + "\n")
            for k, (elements, area_update, table) in enumerate(m.shell_loads):
                table = np.asarray(table, float)
                f.write(f"ShellLoad\t{len(m.nodal_loads) + k + 1}\tElementSet\t{k + 1}\tAreaUpdate\t{1 if area_update else 0}\tNTimes\t{len(table)}\n")
                for row in table:
                    f.write("\t".join(repr(float(v)) for v in row) + "\n")
            for k, (elements, table) in enumerate(pipe_loads):
                table = np.asarray(table, float)
                f.write(f"PipeLoad\t{len(m.nodal_loads) + len(m.shell_loads) + k + 1}\tElementSet\t{len(m.shell_loads) + k + 1}\tNTimes\t{len(table)}\n")
                for row in table:
                    f.write("\t".join(repr(float(v)) for v in row) + "\n")
            for k, (nodes, cs, table) in enumerate(follower_loads):
                table = np.asarray(table, float)
                f.write(f"NodalFollowerLoad\t{len(m.nodal_loads) + len(m.shell_loads) + len(pipe_loads) + k + 1}\tNodeSet\t{len(m.constraints) + len(m.nodal_loads) + k + 1}"
                        f"\tCS\t{cs}\tNTimes\t{len(table)}\n")
                for row in table:
                    f.write("\t".join(repr(float(v)) for v in row) + "\n")
        if m.constraints:
            f.write(f"\nConstraints\t{len(m.constraints)}\n")
            for k, (_, mask) in enumerate(m.constraints):
                f.write(f"NodalConstraint\t{k + 1}\tNodeSet\t{k + 1}\n")
                for b, nm in enumerate(names):
                    f.write(f"\t{nm}\tBoolTable\t{(mask >> b) & 1}\n")
        if m.gravity is not None:
            f.write(f"\nEnvironment\nGravityData\tG\t{_r(m.gravity[0])}\t{_r(m.gravity[1])}\t{_r(m.gravity[2])}\tBoolTable\t1\n")
        f.write("\nSolverOptions\nProcessors\t1\tLinSys\tDirect\n")
