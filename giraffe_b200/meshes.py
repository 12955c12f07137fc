"""Model tables and synthetic meshes for the assembly path.

A ``Model`` is the plain-array form of what GIRAFFE's ``Database`` holds for the
in-scope entities (reference ``Database.h:210-331``): nodes, Hooke materials,
beam sections (constants after ``Section::PreCalc``), homogeneous shell
sections, coordinate systems, elements (type / material / section / CS /
connectivity), nodal constraints, gravity and nodal loads.  It is what the
C-ABI ``gfa_create`` receives, and what the oracle drivers receive.

The generators follow SURVEY.md section 8(d) (BASELINE.json ``configs``):
  * ``beam_line``   -- config 2: Beam_1 riser/cable line on the Z axis
  * ``shell_plate`` -- config 3: Shell_1 plate of 6-node triangles
  * ``solid_block`` -- config 4: Solid_1 block of 8-node hexahedra
  * ``mixed_model`` -- config 5: Beam_1 + Shell_1 + Solid_1 in one model
Element type ids are the reference's (``Element.h:8-15``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

BEAM_1 = 1
PIPE_1 = 2       # evaluated by the Beam_1 kernel (Pipe_1::Mount is Beam_1::Mount, Pipe_1.cpp:836-974)
SHELL_1 = 3
SOLID_1 = 7
NODES_PER_TYPE = {BEAM_1: 3, PIPE_1: 3, SHELL_1: 6, SOLID_1: 8}
DOFS_PER_TYPE = {BEAM_1: 18, PIPE_1: 18, SHELL_1: 27, SOLID_1: 24}


@dataclass
class Model:
    xyz: np.ndarray                                   # [n_nodes, 3] reference coordinates
    hooke: np.ndarray                                 # [n_mat, 3]  E, nu, rho
    sections: np.ndarray                              # [n_sec, 6]  A I11 I22 I12 I33 It
    section_defs: list = field(default_factory=list)  # [(kind, a, b)] kind 0 Rectangle(B,H) / 1 Tube(De,Di)
    shell_thickness: np.ndarray = field(default_factory=lambda: np.zeros(0))
    cs_defs: list = field(default_factory=list)       # [(E1, E3)] as given in the input
    cs: np.ndarray = field(default_factory=lambda: np.zeros((0, 9)))  # E1,E2,E3 normalised
    elem_type: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    elem_mat: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    elem_sec: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    elem_cs: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    elem_ptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    elem_nodes: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))  # 1-based
    pretension: np.ndarray | None = None
    constraints: list = field(default_factory=list)   # [(node ids 1-based, mask)]
    gravity: tuple | None = None
    nodal_loads: list = field(default_factory=list)   # [(node ids, cs id, table[n,7])]
    # [n, 11] EA EI GJ GA Rho CDt CDn CAt CAn De Di (PipeSection.h:13-23); Pipe_1's elem_sec points here
    pipe_sections: np.ndarray = field(default_factory=lambda: np.zeros((0, 11)))
    # ShellLoad (ShellLoad.h): [(element ids 1-based, area_update, table[n,2] = time, pressure)] -- a host-side
    # contributor (Load), not part of the device path
    shell_loads: list = field(default_factory=list)
    # PipeLoad (PipeLoad.h): [(element ids 1-based, table[n,5] = time, P0I, P0E, RhoI, RhoE)] -- internal pressure on
    # Pipe_1 elements (Pipe_1::MountPipeSpecialLoads uses P0I only), evaluated by gfa_apply_pipe_loads
    pipe_loads: list = field(default_factory=list)
    # NodalFollowerLoad (NodalFollowerLoad.h): [(node ids 1-based, CS id, table[n,7])] like nodal_loads; forces and moments
    # follow the node's rotation.  A host-side Load: it enters through gfa_add_host_triplets / gfa_add_host_vector
    follower_loads: list = field(default_factory=list)

    @property
    def n_nodes(self) -> int:
        return int(self.xyz.shape[0])

    @property
    def n_elements(self) -> int:
        return int(self.elem_type.shape[0])

    def constraint_mask(self) -> np.ndarray:
        """Per-node 6-bit mask of constrained DOFs (NodalConstraint.cpp:152-173)."""
        m = np.zeros(self.n_nodes, np.int32)
        for nodes, mask in self.constraints:
            m[np.asarray(nodes, np.int64) - 1] |= mask
        return m


# --------------------------------------------------------------------------
# section constants  (SecRectangle.cpp:82-96, SecTube.cpp:86-94)
# --------------------------------------------------------------------------
_PI = 3.1415926535897932384626433832795


def rectangle_constants(b: float, h: float) -> np.ndarray:
    a = b * h
    i11 = b * h * h * h / 12.0
    i22 = h * b * b * b / 12.0
    temp = 0.0
    for n in range(1, 22, 2):
        temp += (1.0 / (math.pow(float(n), 5))) * math.tanh(n * _PI * h / (2 * b))
    it = (1.0 / 3.0) * b * b * b * h * (1.0 - 192.0 * b * temp / (math.pow(_PI, 5) * h))
    return np.array([a, i11, i22, 0.0, i11 + i22, it])


def tube_constants(de: float, di: float) -> np.ndarray:
    a = (_PI / 4.0) * (de * de - di * di)
    i11 = (_PI / 64.0) * (de * de * de * de - di * di * di * di)
    i33 = (_PI / 32.0) * (de * de * de * de - di * di * di * di)
    return np.array([a, i11, i11, 0.0, i33, i33])


def section_constants(kind: int, a: float, b: float) -> np.ndarray:
    return rectangle_constants(a, b) if kind == 0 else tube_constants(a, b)


def normalise_cs(e1, e3) -> np.ndarray:
    """E2 = E3 x E1, then each normalised if its norm differs from 1
    (CoordinateSystem.cpp:64-77)."""
    e1 = np.asarray(e1, float).copy()
    e3 = np.asarray(e3, float).copy()
    e2 = np.array([e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]])

    def nrm(v):
        n = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
        return v if n == 1.0 else (1.0 / n) * v

    return np.concatenate([nrm(e1), nrm(e2), nrm(e3)])


def _finish(m: Model) -> Model:
    m.sections = np.array([section_constants(*d) for d in m.section_defs]).reshape(-1, 6)
    m.cs = np.array([normalise_cs(*d) for d in m.cs_defs]).reshape(-1, 9)
    counts = np.array([NODES_PER_TYPE[int(t)] for t in np.unique(m.elem_type)]) if m.n_elements else None
    npt = np.zeros(8, np.int32)
    for t, n in NODES_PER_TYPE.items():
        npt[t] = n
    m.elem_ptr = np.zeros(m.n_elements + 1, np.int32)
    np.cumsum(npt[m.elem_type], out=m.elem_ptr[1:])
    assert m.elem_ptr[-1] == m.elem_nodes.shape[0]
    del counts
    return m


# --------------------------------------------------------------------------
# config 2: Beam_1 line  (SURVEY.md 8d "Config 2 (B)")
# --------------------------------------------------------------------------
def beam_line(n_elements: int = 100_000, spacing: float = 0.5, tube=(0.65, 0.62),
              hooke=(2.07e11, 0.3, 7850.0), pretension: float = 0.0) -> Model:
    nn = 2 * n_elements + 1
    xyz = np.zeros((nn, 3))
    xyz[:, 2] = spacing * np.arange(nn)
    e = np.arange(n_elements, dtype=np.int64)
    conn = np.stack([2 * e + 1, 2 * e + 2, 2 * e + 3], axis=1).astype(np.int32)
    m = Model(xyz=xyz, hooke=np.array([hooke], float), sections=np.zeros((0, 6)))
    m.section_defs = [(1, tube[0], tube[1])]
    m.cs_defs = [((1.0, 0.0, 0.0), (0.0, 0.0, 1.0))]
    m.elem_type = np.full(n_elements, BEAM_1, np.int32)
    m.elem_mat = np.ones(n_elements, np.int32)
    m.elem_sec = np.ones(n_elements, np.int32)
    m.elem_cs = np.ones(n_elements, np.int32)
    m.elem_nodes = conn.reshape(-1)
    m.pretension = np.full(n_elements, pretension) if pretension else None
    m.constraints = [([1], 0x3F)]
    return _finish(m)


def beam_line_displacements(m: Model, seed: int = 20240001) -> np.ndarray:
    rng = np.random.Generator(np.random.MT19937(seed))
    d = np.zeros((m.n_nodes, 6))
    z = m.xyz[:, 2]
    wave = 1e-3 * np.sin(2 * np.pi * z / 50.0)
    d[:, 0] = wave + rng.uniform(-1e-5, 1e-5, m.n_nodes)
    d[:, 1] = wave + rng.uniform(-1e-5, 1e-5, m.n_nodes)
    d[:, 2] = rng.uniform(-1e-5, 1e-5, m.n_nodes)
    d[:, 3:] = rng.uniform(-1e-2, 1e-2, (m.n_nodes, 3))
    d[0, :] = 0.0  # clamped node
    return d


def pipe_line(n_elements: int = 1000, spacing: float = 0.5,
              pipe_section=(1.2e9, 6.0e7, 4.6e7, 4.6e8, 180.0, 0.1, 1.2, 0.0, 1.0, 0.65, 0.55), gravity=None) -> Model:
    """A straight Pipe_1 riser segment on the Z axis (same topology as beam_line): PipeSection constants
    EA EI GJ GA Rho CDt CDn CAt CAn De Di, no material (Pipe_1.cpp:535-572, PipeSection.cpp)."""
    m = beam_line(n_elements, spacing)
    m.hooke = np.zeros((0, 3))
    m.section_defs = []
    m.pipe_sections = np.array([pipe_section], float)
    m.elem_type = np.full(n_elements, PIPE_1, np.int32)
    m.elem_mat = np.zeros(n_elements, np.int32)
    m.pretension = None
    m.gravity = gravity
    return _finish(m)


# --------------------------------------------------------------------------
# config 3: Shell_1 plate  (SURVEY.md 8d "Config 3 (S)")
# --------------------------------------------------------------------------
def shell_plate(nx: int = 1000, ny: int = 500, cell: float = 0.0195, thickness: float = 0.002,
                hooke=(200e9, 0.3, 8000.0), gravity=None, warp: float = 0.0) -> Model:
    """nx x ny cells in the XY plane, two 6-node triangles per cell.

    Corner nodes first (row-major, (nx+1) x (ny+1)), then mid-side nodes of the
    x-edges, y-edges and cell diagonals.  Element node order c1 c2 c3 m12 m23
    m31 (reference Shell_1.cpp:2062-2070: node 4 on edge 1-2, 5 on 2-3, 6 on
    3-1).  ``warp`` adds a smooth out-of-plane shape so that element frames
    differ from the global axes (used by parity tests).
    """
    ncx, ncy = nx + 1, ny + 1
    n_corner = ncx * ncy
    n_ex = nx * ncy          # mids of edges along x
    n_ey = ncx * ny          # mids of edges along y
    n_d = nx * ny            # mids of cell diagonals
    nn = n_corner + n_ex + n_ey + n_d

    def corner(i, j):
        return j * ncx + i

    def ex(i, j):            # between corner(i,j) and corner(i+1,j)
        return n_corner + j * nx + i

    def ey(i, j):            # between corner(i,j) and corner(i,j+1)
        return n_corner + n_ex + j * ncx + i

    def dg(i, j):            # between corner(i,j) and corner(i+1,j+1)
        return n_corner + n_ex + n_ey + j * nx + i

    xyz = np.zeros((nn, 3))
    I, J = np.meshgrid(np.arange(ncx), np.arange(ncy), indexing="xy")
    xyz[:n_corner, 0] = (I * cell).reshape(-1)
    xyz[:n_corner, 1] = (J * cell).reshape(-1)
    if warp:
        lx, ly = nx * cell, ny * cell
        xyz[:n_corner, 2] = warp * np.sin(np.pi * xyz[:n_corner, 0] / lx) * np.cos(0.5 * np.pi * xyz[:n_corner, 1] / ly)

    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ii = ii.reshape(-1)
    jj = jj.reshape(-1)
    c00, c10, c11, c01 = corner(ii, jj), corner(ii + 1, jj), corner(ii + 1, jj + 1), corner(ii, jj + 1)
    # lower triangle: c00 c10 c11 ; upper triangle: c00 c11 c01
    t1 = np.stack([c00, c10, c11, ex(ii, jj), ey(ii + 1, jj), dg(ii, jj)], axis=1)
    t2 = np.stack([c00, c11, c01, dg(ii, jj), ex(ii, jj + 1), ey(ii, jj)], axis=1)
    conn = np.empty((2 * nx * ny, 6), np.int64)
    conn[0::2] = t1
    conn[1::2] = t2
    # straight-sided elements: mid nodes at edge midpoints
    for a, b, mid in ((0, 1, 3), (1, 2, 4), (2, 0, 5)):
        xyz[conn[:, mid]] = 0.5 * (xyz[conn[:, a]] + xyz[conn[:, b]])
    ne = conn.shape[0]

    m = Model(xyz=xyz, hooke=np.array([hooke], float), sections=np.zeros((0, 6)))
    m.shell_thickness = np.array([thickness])
    m.elem_type = np.full(ne, SHELL_1, np.int32)
    m.elem_mat = np.ones(ne, np.int32)
    m.elem_sec = np.ones(ne, np.int32)
    m.elem_cs = np.zeros(ne, np.int32)
    m.elem_nodes = (conn + 1).astype(np.int32).reshape(-1)
    clamped = np.nonzero(np.abs(xyz[:, 0]) < 1e-12 * max(1.0, nx * cell))[0] + 1
    m.constraints = [(clamped.astype(np.int32), 0x3F)]
    m.gravity = gravity
    return _finish(m)


def shell_plate_displacements(m: Model, seed: int = 20240002, amp_w: float = 1e-3) -> np.ndarray:
    rng = np.random.Generator(np.random.MT19937(seed))
    d = np.zeros((m.n_nodes, 6))
    lx = float(m.xyz[:, 0].max()) or 1.0
    d[:, :3] = rng.uniform(-1e-5, 1e-5, (m.n_nodes, 3))
    d[:, 2] += amp_w * np.sin(np.pi * m.xyz[:, 0] / lx)
    d[:, 3:] = rng.uniform(-1e-3, 1e-3, (m.n_nodes, 3))
    return mask_displacements(m, d)


def mask_displacements(m: Model, d: np.ndarray) -> np.ndarray:
    """Zero the entries of inactive DOFs (e.g. rotations of shell corner nodes)
    and of constrained DOFs (no prescribed motion in the synthetic configs)."""
    gls, _, _ = number_dofs(m)
    d = d.copy()
    d[gls <= 0] = 0.0
    return d


# --------------------------------------------------------------------------
# config 4: Solid_1 block of 8-node hexahedra (builder-defined formulation)
# --------------------------------------------------------------------------
def solid_block(nx: int = 200, ny: int = 200, nz: int = 100, cell: float = 0.01,
                hooke=(70e9, 0.33, 2700.0), gravity=None) -> Model:
    ncx, ncy, ncz = nx + 1, ny + 1, nz + 1
    K, J, I = np.meshgrid(np.arange(ncz), np.arange(ncy), np.arange(ncx), indexing="ij")
    xyz = np.stack([I.reshape(-1) * cell, J.reshape(-1) * cell, K.reshape(-1) * cell], axis=1).astype(float)

    def nid(i, j, k):
        return (k * ncy + j) * ncx + i

    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.reshape(-1), j.reshape(-1), k.reshape(-1)
    conn = np.stack([nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k),
                     nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)], axis=1)
    ne = conn.shape[0]
    m = Model(xyz=xyz, hooke=np.array([hooke], float), sections=np.zeros((0, 6)))
    m.cs_defs = [((1.0, 0.0, 0.0), (0.0, 0.0, 1.0))]
    m.elem_type = np.full(ne, SOLID_1, np.int32)
    m.elem_mat = np.ones(ne, np.int32)
    m.elem_sec = np.zeros(ne, np.int32)
    m.elem_cs = np.ones(ne, np.int32)
    m.elem_nodes = (conn + 1).astype(np.int32).reshape(-1)
    clamped = np.nonzero(xyz[:, 2] == 0.0)[0] + 1
    m.constraints = [(clamped.astype(np.int32), 0x07)]
    m.gravity = gravity
    return _finish(m)


def solid_block_displacements(m: Model, seed: int = 20240003, amp: float = 1e-4) -> np.ndarray:
    rng = np.random.Generator(np.random.MT19937(seed))
    d = np.zeros((m.n_nodes, 6))
    d[:, :3] = rng.uniform(-amp, amp, (m.n_nodes, 3))
    mask = m.constraint_mask()
    for k in range(3):
        d[(mask >> k) & 1 == 1, k] = 0.0
    return d


# --------------------------------------------------------------------------
# config 5: mixed model -- a solid block carrying a shell plate on its top
# face's edge line and a beam line hanging from one corner; the three parts
# share nodes only through their own connectivity (independent sub-meshes
# concatenated), which is what mesh-partitioning by element range needs.
# --------------------------------------------------------------------------
def concat_models(parts: list[Model]) -> Model:
    xyz, et, em, es, ec, en = [], [], [], [], [], []
    hooke, secdefs, thick, csdefs, cons = [], [], [], [], []
    pipes = []
    pret = []
    node_off = 0
    for p in parts:
        mo, so, to, co, po = len(hooke), len(secdefs), len(thick), len(csdefs), len(pipes)
        xyz.append(p.xyz)
        hooke.extend(p.hooke.tolist())
        secdefs.extend(p.section_defs)
        thick.extend(p.shell_thickness.tolist())
        csdefs.extend(p.cs_defs)
        pipes.extend(np.asarray(p.pipe_sections, float).reshape(-1, 11).tolist())
        et.append(p.elem_type)
        em.append(np.where(p.elem_mat > 0, p.elem_mat + mo, 0))
        is_shell = p.elem_type == SHELL_1
        is_pipe = p.elem_type == PIPE_1
        es.append(np.where(is_shell, p.elem_sec + to, np.where(is_pipe, p.elem_sec + po, np.where(p.elem_sec > 0, p.elem_sec + so, 0))).astype(np.int32))
        ec.append(np.where(p.elem_cs > 0, p.elem_cs + co, 0).astype(np.int32))
        en.append(p.elem_nodes + node_off)
        pret.append(p.pretension if p.pretension is not None else np.zeros(p.n_elements))
        for nodes, mask in p.constraints:
            cons.append((np.asarray(nodes, np.int32) + node_off, mask))
        node_off += p.n_nodes
    m = Model(xyz=np.concatenate(xyz), hooke=np.array(hooke, float), sections=np.zeros((0, 6)))
    m.section_defs = secdefs
    m.shell_thickness = np.array(thick, float)
    m.cs_defs = csdefs
    m.pipe_sections = np.array(pipes, float).reshape(-1, 11)
    m.elem_type = np.concatenate(et).astype(np.int32)
    m.elem_mat = np.concatenate(em).astype(np.int32)
    m.elem_sec = np.concatenate(es).astype(np.int32)
    m.elem_cs = np.concatenate(ec).astype(np.int32)
    m.elem_nodes = np.concatenate(en).astype(np.int32)
    m.pretension = np.concatenate(pret)
    m.constraints = cons
    m.gravity = next((p.gravity for p in parts if p.gravity is not None), None)
    return _finish(m)


def submodel(m: Model, elems) -> tuple[Model, np.ndarray]:
    """Sub-model made of the listed elements only (0-based indices; same node coordinates, no constraints);
    returns it with the 1-based ids of its nodes in the parent model."""
    ptr = m.elem_ptr
    elems = np.asarray(elems, np.int64)
    nodes = np.unique(np.concatenate([m.elem_nodes[ptr[e]:ptr[e + 1]] for e in elems]))
    remap = np.zeros(m.n_nodes + 1, np.int32)
    remap[nodes] = np.arange(1, len(nodes) + 1, dtype=np.int32)
    sub = Model(xyz=m.xyz[nodes - 1], hooke=m.hooke, sections=m.sections)
    sub.section_defs, sub.shell_thickness, sub.cs_defs = m.section_defs, m.shell_thickness, m.cs_defs
    sub.pipe_sections = getattr(m, "pipe_sections", np.zeros((0, 11)))
    sub.elem_type, sub.elem_mat = m.elem_type[elems], m.elem_mat[elems]
    sub.elem_sec, sub.elem_cs = m.elem_sec[elems], m.elem_cs[elems]
    sub.elem_nodes = np.concatenate([remap[m.elem_nodes[ptr[e]:ptr[e + 1]]] for e in elems]).astype(np.int32)
    sub.pretension = None if m.pretension is None else np.asarray(m.pretension)[elems]
    sub.gravity = m.gravity
    return _finish(sub), nodes


def mixed_model(n_beam: int, shell_nx: int, shell_ny: int, solid_n: tuple) -> Model:
    return concat_models([beam_line(n_beam), shell_plate(shell_nx, shell_ny), solid_block(*solid_n)])


# --------------------------------------------------------------------------
# DOF numbering (Solution.cpp:121-224 DOFsActive, :40-118 SetGlobalDOFs)
# --------------------------------------------------------------------------
def number_dofs(m: Model):
    """Return (GLs[n_nodes,6] int32, n_free, n_fixed) exactly as the reference
    numbers them: node-major, DOF-minor; free ids 1.., fixed ids -1, -2, ..."""
    active = np.zeros((m.n_nodes, 6), bool)
    conn_type = np.repeat(m.elem_type, np.diff(m.elem_ptr))
    local = np.arange(m.elem_nodes.size) - np.repeat(m.elem_ptr[:-1], np.diff(m.elem_ptr))
    nodes0 = m.elem_nodes.astype(np.int64) - 1
    active[nodes0, 0:3] = True
    rot = (conn_type == BEAM_1) | (conn_type == PIPE_1) | ((conn_type == SHELL_1) & (local >= 3))
    active[nodes0[rot], 3:6] = True
    mask = m.constraint_mask()
    fixed = ((mask[:, None] >> np.arange(6)[None, :]) & 1).astype(bool)
    free = active & ~fixed
    fix = active & fixed
    gls = np.zeros((m.n_nodes, 6), np.int32)
    gls.reshape(-1)[free.reshape(-1)] = np.arange(1, int(free.sum()) + 1, dtype=np.int32)
    gls.reshape(-1)[fix.reshape(-1)] = -np.arange(1, int(fix.sum()) + 1, dtype=np.int32)
    return gls, int(free.sum()), int(fix.sum())
