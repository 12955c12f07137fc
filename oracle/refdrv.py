"""TEST INFRASTRUCTURE (oracle) -- ctypes front end of ``oracle/_ref/libgiraffe_ref.so``.

That library is the reference's own UNMODIFIED element / solution sources
(``/root/reference/src/{Beam_1,Shell_1,Solid_1,Node,Solution,...}.cpp``) built
by ``oracle/Makefile`` against the shims in ``oracle/ref_shims/``.  This module
feeds it a :class:`giraffe_b200.meshes.Model` and drives the same call sequence
as ``Static::Solve`` (reference ``Static.cpp:161-163,203-212``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.  The product path never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgiraffe_ref.so")

_I = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_D = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(LIB_PATH)


class RefOracle:
    """One process-wide instance: the reference keeps its model in a global ``db``."""

    MATS = {"AA": 0, "AB": 1, "BA": 2, "BB": 3}

    def __init__(self, threads: int | None = None):
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(LIB_PATH)
        L = self.lib
        L.ref_set_nodes.argtypes = [C.c_int, _D]
        L.ref_add_hooke.argtypes = [C.c_double] * 3
        L.ref_add_section.argtypes = [C.c_int, C.c_double, C.c_double]
        L.ref_get_section.argtypes = [C.c_int, _D]
        L.ref_add_shell_section.argtypes = [C.c_double]
        L.ref_add_cs.argtypes = [_D, _D]
        L.ref_get_cs.argtypes = [C.c_int, _D]
        L.ref_set_elements.argtypes = [C.c_int, _I, _I, _I, _I, _I, C.c_void_p]
        L.ref_set_gravity.argtypes = [C.c_double] * 3
        L.ref_add_nodal_constraint.argtypes = [C.c_int, _I, C.c_int]
        L.ref_add_nodal_load.argtypes = [C.c_int, _I, C.c_int, C.c_int, _D]
        L.ref_add_shell_load.argtypes = [C.c_int, _I, C.c_int, C.c_int, _D]
        L.ref_add_pipe_load.argtypes = [C.c_int, _I, C.c_int, _D]
        L.ref_add_nodal_follower_load.argtypes = [C.c_int, _I, C.c_int, C.c_int, _D]
        L.ref_get_gls.argtypes = [_I]
        L.ref_set_time.argtypes = [C.c_double, C.c_double]
        L.ref_set_displacements.argtypes = [_D]
        L.ref_get_copy_coordinates.argtypes = [_D]
        L.ref_assemble.argtypes = [C.c_int, _D]
        L.ref_mount_local.argtypes = [_D]
        for f in (L.ref_triplet_count, L.ref_csr_nnz):
            f.argtypes = [C.c_int]
            f.restype = C.c_long
        L.ref_csr_rows.argtypes = [C.c_int]
        L.ref_csr_cols.argtypes = [C.c_int]
        L.ref_csr_get.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_get_vectors.argtypes = [_D, _D, _D]
        L.ref_get_element.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_get_state.argtypes = [C.c_int, _D]
        L.ref_add_pipe_section.argtypes = [_D]
        L.ref_residual.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(np.int32, flags='C')]
        L.ref_update_displacements.argtypes = [_D, np.ctypeslib.ndpointer(np.int32, flags='C'), _D]
        L.ref_get_results.argtypes = [C.c_int, _D]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_dynamic_begin.argtypes = [C.c_double] * 4 + [C.c_int]
        L.ref_newmark.argtypes = [C.c_double, _D]
        L.ref_set_kinematics.argtypes = [C.c_void_p] * 4
        L.ref_get_kinematics.argtypes = [C.c_void_p] * 4
        L.ref_assemble_dynamic.argtypes = [C.c_int, C.c_int]
        L.ref_get_alpha_i.argtypes = [C.c_int, _D]
        if threads:
            L.ref_set_threads(int(threads))
        self.model = None

    # ---- model ---------------------------------------------------------
    def load(self, m, check: bool = True):
        L = self.lib
        L.ref_reset()
        L.ref_set_nodes(m.n_nodes, np.ascontiguousarray(m.xyz, np.float64).reshape(-1))
        for E, nu, rho in m.hooke:
            L.ref_add_hooke(float(E), float(nu), float(rho))
        for kind, a, b in m.section_defs:
            L.ref_add_section(int(kind), float(a), float(b))
        for t in m.shell_thickness:
            L.ref_add_shell_section(float(t))
        for row in np.asarray(getattr(m, 'pipe_sections', np.zeros((0, 11))), float).reshape(-1, 11):
            L.ref_add_pipe_section(np.ascontiguousarray(row, np.float64))
        for e1, e3 in m.cs_defs:
            r = L.ref_add_cs(np.asarray(e1, np.float64), np.asarray(e3, np.float64))
            if r < 0:
                raise ValueError("reference rejected the coordinate system")
        pret = None
        if m.pretension is not None:
            self._pret = np.ascontiguousarray(m.pretension, np.float64)
            pret = self._pret.ctypes.data_as(C.c_void_p)
        L.ref_set_elements(m.n_elements, m.elem_type.astype(np.int32), m.elem_mat.astype(np.int32),
                           m.elem_sec.astype(np.int32), m.elem_cs.astype(np.int32),
                           np.ascontiguousarray(m.elem_nodes, np.int32), pret)
        if m.gravity is not None:
            L.ref_set_gravity(*[float(g) for g in m.gravity])
        for nodes, mask in m.constraints:
            nodes = np.ascontiguousarray(nodes, np.int32)
            L.ref_add_nodal_constraint(len(nodes), nodes, int(mask))
        for nodes, cs, table in m.nodal_loads:
            nodes = np.ascontiguousarray(nodes, np.int32)
            table = np.ascontiguousarray(table, np.float64)
            if L.ref_add_nodal_load(len(nodes), nodes, int(cs), table.shape[0], table.reshape(-1)) < 0:
                raise ValueError("reference rejected the nodal load")
        for nodes, cs, table in getattr(m, "follower_loads", []):
            nodes = np.ascontiguousarray(nodes, np.int32)
            table = np.ascontiguousarray(table, np.float64)
            if L.ref_add_nodal_follower_load(len(nodes), nodes, int(cs), table.shape[0], table.reshape(-1)) < 0:
                raise ValueError("reference rejected the nodal follower load")
        for elements, area_update, table in getattr(m, "shell_loads", []):
            elements = np.ascontiguousarray(elements, np.int32)
            table = np.ascontiguousarray(table, np.float64)
            if L.ref_add_shell_load(len(elements), elements, 1 if area_update else 0, table.shape[0], table.reshape(-1)) < 0:
                raise ValueError("reference rejected the shell load")
        for elements, table in getattr(m, "pipe_loads", []):
            elements = np.ascontiguousarray(elements, np.int32)
            table = np.ascontiguousarray(table, np.float64)
            if L.ref_add_pipe_load(len(elements), elements, table.shape[0], table.reshape(-1)) < 0:
                raise ValueError("reference rejected the pipe load")
        if check:
            bad = L.ref_check()
            if bad:
                raise ValueError(f"reference Element::Check failed on element {bad}")
        L.ref_precalc()
        L.ref_setup_dofs()
        self.model = m
        return self

    @property
    def n_free(self) -> int:
        return self.lib.ref_n_free()

    @property
    def n_fixed(self) -> int:
        return self.lib.ref_n_fixed()

    def gls(self) -> np.ndarray:
        g = np.zeros(self.model.n_nodes * 6, np.int32)
        self.lib.ref_get_gls(g)
        return g.reshape(-1, 6)

    def section(self, sid: int) -> np.ndarray:
        out = np.zeros(6)
        self.lib.ref_get_section(sid, out)
        return out

    def cs(self, cid: int) -> np.ndarray:
        out = np.zeros(9)
        self.lib.ref_get_cs(cid, out)
        return out

    # ---- per iteration -------------------------------------------------
    def set_time(self, last_converged: float, step: float):
        self.lib.ref_set_time(float(last_converged), float(step))

    def assemble(self, disp: np.ndarray, with_loads: bool = False) -> np.ndarray:
        """One Newton-iteration assembly; returns the 5 phase times in seconds
        (MountLocal, MountElementLoads, MountGlobal, MountSparse, Clear+MountLoads)."""
        self.lib.ref_set_displacements(np.ascontiguousarray(disp, np.float64).reshape(-1))
        sec = np.zeros(5)
        self.lib.ref_assemble(1 if with_loads else 0, sec)
        return sec

    def mount_local(self, disp: np.ndarray) -> np.ndarray:
        self.lib.ref_set_displacements(np.ascontiguousarray(disp, np.float64).reshape(-1))
        sec = np.zeros(2)
        self.lib.ref_mount_local(sec)
        return sec

    def commit(self):
        self.lib.ref_commit()

    def copy_coordinates(self) -> np.ndarray:
        c = np.zeros(self.model.n_nodes * 6)
        self.lib.ref_get_copy_coordinates(c)
        return c.reshape(-1, 6)

    # ---- results -------------------------------------------------------
    def csr(self, which: str = "AA"):
        w = self.MATS[which]
        nr, nz = self.lib.ref_csr_rows(w), self.lib.ref_csr_nnz(w)
        outer = np.zeros(nr + 1, np.int32)
        inner = np.zeros(nz, np.int32)
        val = np.zeros(nz, np.float64)
        self.lib.ref_csr_get(w, outer.ctypes.data, inner.ctypes.data, val.ctypes.data)
        return outer, inner, val, (nr, self.lib.ref_csr_cols(w))

    def triplets(self, which: str = "AA") -> int:
        return self.lib.ref_triplet_count(self.MATS[which])

    def vectors(self):
        pa, ia, pb = np.zeros(self.n_free), np.zeros(self.n_free), np.zeros(self.n_fixed)
        self.lib.ref_get_vectors(pa, ia, pb)
        return pa, ia, pb

    def element(self, e: int):
        n = self.lib.ref_get_element(e, None, None, None)
        K = np.zeros((n, n))
        P = np.zeros(n)
        en = C.c_double(0.0)
        self.lib.ref_get_element(e, K.ctypes.data, P.ctypes.data, C.addressof(en))
        return K, P, en.value

    def residual(self, X_B=None):
        """Static.cpp:210-217 + EstablishResidualCriteria / CheckResidualConvergence through the reference's
        own code; returns (node_force, node_moment, diverged); vectors() then holds the right-hand side."""
        out = np.zeros(4, np.int32)
        xb = np.ascontiguousarray(X_B, np.float64) if X_B is not None else None
        self.lib.ref_residual(xb.ctypes.data if xb is not None else None, out)
        return int(out[0]), int(out[1]), int(out[2])

    def update_displacements(self, x_A):
        """Solution::UpdateDisps + CheckGLConvergence; returns (displacements[n,6], node_disp, node_rot, diverged)."""
        out = np.zeros(4, np.int32)
        d = np.zeros(self.model.n_nodes * 6)
        self.lib.ref_update_displacements(np.ascontiguousarray(x_A, np.float64), out, d)
        return d.reshape(-1, 6), int(out[0]), int(out[1]), int(out[2])

    def results(self, e: int) -> np.ndarray:
        """Gauss-point results of element e after the last assemble, in the layout of
        gfa_gauss_point_results: [strain_energy, per point strains / resultants]."""
        buf = np.zeros(80)
        n = self.lib.ref_get_results(e, buf)
        return buf[:n].copy()

    def state(self, e: int) -> np.ndarray:
        buf = np.zeros(64)
        n = self.lib.ref_get_state(e, buf)
        return buf[:n].copy()

    # ---- Dynamic (Newmark) path: Dynamic.cpp:303-340 ----------------------
    def dynamic_begin(self, beta_new=0.3, gamma_new=0.5, rayleigh_alpha=0.0, rayleigh_beta=0.0, update=0):
        """Replace the solution object by the reference's own Dynamic (call after load())."""
        self.lib.ref_dynamic_begin(float(beta_new), float(gamma_new), float(rayleigh_alpha), float(rayleigh_beta), int(update))

    def newmark(self, time_step: float) -> np.ndarray:
        a = np.zeros(6)
        if self.lib.ref_newmark(float(time_step), a) < 0:
            raise RuntimeError("dynamic_begin() first")
        return a

    def set_kinematics(self, vel=None, accel=None, copy_vel=None, copy_accel=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64).reshape(-1) for a in (vel, accel, copy_vel, copy_accel)]
        self.lib.ref_set_kinematics(*[None if a is None else a.ctypes.data for a in arrs])

    def kinematics(self):
        """(vel, accel, copy_vel, copy_accel), each [n_nodes, 6]"""
        out = [np.zeros(self.model.n_nodes * 6) for _ in range(4)]
        self.lib.ref_get_kinematics(*[a.ctypes.data for a in out])
        return tuple(a.reshape(-1, 6) for a in out)

    def update_dyn(self, disp):
        """Dynamic::UpdateDyn for the given Node::displacements"""
        self.lib.ref_set_displacements(np.ascontiguousarray(disp, np.float64).reshape(-1))
        if self.lib.ref_update_dyn() < 0:
            raise RuntimeError("dynamic_begin() first")

    def assemble_dynamic(self, disp, update_rayleigh: bool, with_loads: bool = False):
        self.lib.ref_set_displacements(np.ascontiguousarray(disp, np.float64).reshape(-1))
        if self.lib.ref_assemble_dynamic(1 if with_loads else 0, 1 if update_rayleigh else 0) < 0:
            raise RuntimeError("dynamic_begin() first")

    def alpha_i(self, e: int) -> np.ndarray:
        buf = np.zeros(16)
        n = self.lib.ref_get_alpha_i(e, buf)
        return buf[:n].copy()
