"""TEST INFRASTRUCTURE (oracle) -- ctypes front end of ``oracle/libgfa_oracle.so``,
the builder's plain-C++ restatement of the assembly path (``oracle/port/``).

Same surface as :class:`oracle.refdrv.RefOracle`, so a test can run either
against the same :class:`giraffe_b200.meshes.Model`.  ``ensure_built()``
compiles the library with g++ when it is missing (source-only checkouts).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.  The product path never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgfa_oracle.so")
_I = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_D = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def ensure_built(force: bool = False) -> str:
    src = os.path.join(_HERE, "port", "gfa_oracle.cpp")
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class PortOracle:
    MATS = {"AA": 0, "AB": 1, "BA": 2, "BB": 3}

    def __init__(self, threads: int | None = None):
        self.lib = C.CDLL(ensure_built())
        L = self.lib
        L.gfo_set_nodes.argtypes = [C.c_int, _D]
        L.gfo_set_materials.argtypes = [C.c_int, _D]
        L.gfo_set_sections.argtypes = [C.c_int, C.c_void_p]
        L.gfo_set_shell_sections.argtypes = [C.c_int, C.c_void_p]
        L.gfo_set_cs.argtypes = [C.c_int, C.c_void_p]
        L.gfo_set_pipe_sections.argtypes = [C.c_int, C.c_void_p]
        L.gfo_set_elements.argtypes = [C.c_int, _I, _I, _I, _I, _I, _I, C.c_void_p]
        L.gfo_set_gravity.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        L.gfo_set_constraint_mask.argtypes = [_I]
        L.gfo_set_pipe_loads.argtypes = [C.c_int, _I, _I, _D]
        L.gfo_get_gls.argtypes = [_I]
        L.gfo_set_extra_triplets.argtypes = [C.c_int, C.c_long, _I, _I, _D]
        L.gfo_assemble.argtypes = [_D, C.c_double, _D]
        for f in (L.gfo_triplet_count, L.gfo_csr_nnz):
            f.argtypes = [C.c_int]
            f.restype = C.c_long
        L.gfo_csr_get.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gfo_get_vectors.argtypes = [_D, _D, _D]
        L.gfo_get_element.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gfo_get_state.argtypes = [C.c_int, _D]
        L.gfo_get_results.argtypes = [C.c_int, _D]
        L.gfo_get_copy_coordinates.argtypes = [_D]
        L.gfo_set_dynamic.argtypes = [_D, C.c_double, C.c_double]
        L.gfo_set_kinematics.argtypes = [C.c_void_p] * 4
        L.gfo_get_kinematics.argtypes = [C.c_void_p] * 4
        L.gfo_update_dyn.argtypes = [_D]
        L.gfo_assemble_dynamic.argtypes = [_D, C.c_double, C.c_int]
        L.gfo_get_alpha_i.argtypes = [C.c_int, _D]
        if threads:
            L.gfo_set_threads(int(threads))
        self.model = None
        self.gravity_factor = 1.0

    def load(self, m):
        L = self.lib
        L.gfo_reset()
        L.gfo_set_nodes(m.n_nodes, np.ascontiguousarray(m.xyz, np.float64).reshape(-1))
        L.gfo_set_materials(len(m.hooke), np.ascontiguousarray(m.hooke, np.float64).reshape(-1))
        sec = np.ascontiguousarray(m.sections, np.float64).reshape(-1)
        L.gfo_set_sections(len(m.sections), sec.ctypes.data)
        th = np.ascontiguousarray(m.shell_thickness, np.float64)
        L.gfo_set_shell_sections(len(th), th.ctypes.data)
        cs = np.ascontiguousarray(m.cs, np.float64).reshape(-1)
        L.gfo_set_cs(len(m.cs), cs.ctypes.data)
        pipes = np.ascontiguousarray(np.asarray(getattr(m, 'pipe_sections', np.zeros((0, 11))), float).reshape(-1, 11))
        L.gfo_set_pipe_sections(len(pipes), pipes.ctypes.data)
        pret = None
        if m.pretension is not None:
            self._pret = np.ascontiguousarray(m.pretension, np.float64)
            pret = self._pret.ctypes.data
        L.gfo_set_elements(m.n_elements, m.elem_type.astype(np.int32), m.elem_mat.astype(np.int32),
                           m.elem_sec.astype(np.int32), m.elem_cs.astype(np.int32),
                           m.elem_ptr.astype(np.int32), np.ascontiguousarray(m.elem_nodes, np.int32), pret)
        if m.gravity is not None:
            L.gfo_set_gravity(1, *[float(g) for g in m.gravity])
        L.gfo_set_constraint_mask(m.constraint_mask().astype(np.int32))
        if L.gfo_precalc() != 0:
            raise ValueError("unsupported element type")
        self.model = m
        self._set_pipe_loads(0.0)
        return self

    @property
    def n_free(self) -> int:
        return self.lib.gfo_n_free()

    @property
    def n_fixed(self) -> int:
        return self.lib.gfo_n_fixed()

    def gls(self) -> np.ndarray:
        g = np.zeros(self.model.n_nodes * 6, np.int32)
        self.lib.gfo_get_gls(g)
        return g.reshape(-1, 6)

    def set_time(self, last_converged: float, step: float, start: float = 0.0, end: float = 1.0):
        """Gravity ramp of a first solution step (BoolTable.cpp:84-106); PipeLoad pressures at the evaluation time
        (Load::GetValueAt(last_converged_time + current_time_step, 0): linear table, Table.cpp)."""
        self.gravity_factor = (last_converged + step - start) / (end - start)
        self._set_pipe_loads(last_converged + step)

    def _set_pipe_loads(self, time: float):
        loads = getattr(self.model, "pipe_loads", []) if self.model is not None else []
        ptr = np.zeros(len(loads) + 1, np.int32)
        for k, (elements, _) in enumerate(loads):
            ptr[k + 1] = ptr[k] + len(elements)
        el = np.concatenate([np.asarray(e, np.int32) - 1 for e, _ in loads]).astype(np.int32) if loads else np.zeros(0, np.int32)
        p = np.array([float(np.interp(time, np.asarray(t, float)[:, 0], np.asarray(t, float)[:, 1])) for _, t in loads], np.float64)
        if self.lib.gfo_set_pipe_loads(len(loads), ptr, np.ascontiguousarray(el), np.ascontiguousarray(p)) != 0:
            raise ValueError("a PipeLoad names an element that is not a Pipe_1")

    def set_extra_triplets(self, which: str, rows, cols, vals):
        self.lib.gfo_set_extra_triplets(self.MATS[which], len(vals), np.ascontiguousarray(rows, np.int32),
                                        np.ascontiguousarray(cols, np.int32), np.ascontiguousarray(vals, np.float64))

    def assemble(self, disp: np.ndarray, with_loads: bool = False) -> np.ndarray:
        sec = np.zeros(5)
        self.lib.gfo_assemble(np.ascontiguousarray(disp, np.float64).reshape(-1), float(self.gravity_factor), sec)
        return sec

    def commit(self):
        self.lib.gfo_commit()

    def copy_coordinates(self) -> np.ndarray:
        c = np.zeros(self.model.n_nodes * 6)
        self.lib.gfo_get_copy_coordinates(c)
        return c.reshape(-1, 6)

    def csr(self, which: str = "AA"):
        w = self.MATS[which]
        nr, nz = self.lib.gfo_csr_rows(w), self.lib.gfo_csr_nnz(w)
        outer = np.zeros(nr + 1, np.int32)
        inner = np.zeros(nz, np.int32)
        val = np.zeros(nz, np.float64)
        self.lib.gfo_csr_get(w, outer.ctypes.data, inner.ctypes.data, val.ctypes.data)
        return outer, inner, val, (nr, self.lib.gfo_csr_cols(w))

    def triplets(self, which: str = "AA") -> int:
        return self.lib.gfo_triplet_count(self.MATS[which])

    def vectors(self):
        pa, ia, pb = np.zeros(self.n_free), np.zeros(self.n_free), np.zeros(self.n_fixed)
        self.lib.gfo_get_vectors(pa, ia, pb)
        return pa, ia, pb

    def element(self, e: int):
        n = self.lib.gfo_get_element(e, None, None, None)
        K = np.zeros((n, n))
        P = np.zeros(n)
        en = C.c_double(0.0)
        self.lib.gfo_get_element(e, K.ctypes.data, P.ctypes.data, C.addressof(en))
        return K, P, en.value

    def results(self, e: int) -> np.ndarray:
        """Gauss-point results of element e after the last assemble, in the layout of
        gfa_gauss_point_results: [strain_energy, per point strains / resultants]."""
        buf = np.zeros(80)
        n = self.lib.gfo_get_results(e, buf)
        return buf[:n].copy()

    def state(self, e: int) -> np.ndarray:
        buf = np.zeros(64)
        n = self.lib.gfo_get_state(e, buf)
        return buf[:n].copy()

    # ---- Newmark dynamics (Dynamic.cpp:303-340) -------------------------------
    @staticmethod
    def newmark_coefficients(time_step: float, beta_new: float = 0.3, gamma_new: float = 0.5) -> np.ndarray:
        """Dynamic::CalculateNewmarkCoeff (Dynamic.cpp:582-590): a1..a6"""
        dt, b, g = float(time_step), float(beta_new), float(gamma_new)
        return np.array([1.0 / (dt * dt * b), 1.0 / (dt * b), 1.0 / (2.0 * b) - 1.0, g / (dt * b), 1.0 - g / b,
                         dt * (1.0 - g / (2.0 * b))])

    def set_dynamic(self, newmark6, rayleigh_alpha: float = 0.0, rayleigh_beta: float = 0.0):
        self.lib.gfo_set_dynamic(np.ascontiguousarray(newmark6, np.float64), float(rayleigh_alpha), float(rayleigh_beta))

    def set_kinematics(self, vel=None, accel=None, copy_vel=None, copy_accel=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64).reshape(-1) for a in (vel, accel, copy_vel, copy_accel)]
        self.lib.gfo_set_kinematics(*[None if a is None else a.ctypes.data for a in arrs])

    def kinematics(self):
        out = [np.zeros(self.model.n_nodes * 6) for _ in range(4)]
        self.lib.gfo_get_kinematics(*[a.ctypes.data for a in out])
        return tuple(a.reshape(-1, 6) for a in out)

    def update_dyn(self, disp):
        self.lib.gfo_update_dyn(np.ascontiguousarray(disp, np.float64).reshape(-1))

    def assemble_dynamic(self, disp, update_rayleigh: bool, with_loads: bool = False):
        r = self.lib.gfo_assemble_dynamic(np.ascontiguousarray(disp, np.float64).reshape(-1), float(self.gravity_factor),
                                          1 if update_rayleigh else 0)
        if r != 0:
            raise ValueError("dynamic assembly is restated for Beam_1, Pipe_1 and Shell_1 only")

    def alpha_i(self, e: int) -> np.ndarray:
        buf = np.zeros(16)
        n = self.lib.gfo_get_alpha_i(e, buf)
        return buf[:n].copy()
