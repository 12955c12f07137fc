"""ctypes binding of the C-ABI in ``include/gfa.h`` (``giraffe_b200/libgfa.so``).

This is the Python-side stub a maintainer would keep next to the C header; the
tests and ``bench.py`` call the CUDA path exclusively through it.  There is no
CPU implementation behind these calls: when the shared library is missing, or
when no CUDA device is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .meshes import Model

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GFA_LIB") or os.path.join(_HERE, "libgfa.so")      # GFA_LIB: another build of the same library (experiments)

AA, AB, BA, BB = 0, 1, 2, 3
P_A, I_A, P_B = 0, 1, 2
MATS = {"AA": AA, "AB": AB, "BA": BA, "BB": BB}

# every symbol declared in include/gfa.h
EXPORTS = [
    "gfa_last_error", "gfa_device_count", "gfa_create", "gfa_destroy", "gfa_number_dofs", "gfa_set_dofs",
    "gfa_csr_dims", "gfa_csr_pattern", "gfa_assemble", "gfa_add_host_triplets", "gfa_add_host_vector",
    "gfa_csr_values", "gfa_csr_values_device", "gfa_vector", "gfa_vector_device", "gfa_element_block",
    "gfa_commit_state", "gfa_element_state", "gfa_results_stride", "gfa_gauss_point_results",
    "gfa_residual", "gfa_update_displacements", "gfa_displacements", "gfa_copy_coordinates", "gfa_last_timing", "gfa_last_launch_count",
    "gfa_interface_counts", "gfa_interface_pack", "gfa_interface_unpack", "gfa_local_rows", "gfa_owned_rows", "gfa_stream",
    "gfa_set_kinematics", "gfa_kinematics", "gfa_update_dyn", "gfa_assemble_dynamic", "gfa_element_alpha_i", "gfa_assemble_enqueue", "gfa_interface_stream",
    "gfa_pipeline_info", "gfa_touched_nodes", "gfa_set_displacements_packed", "gfa_vector_owned",
    "gfa_set_shell_loads", "gfa_apply_shell_loads", "gfa_set_pipe_loads", "gfa_apply_pipe_loads",
]


class GfaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gfa error {code}: {msg}")
        self.code = code


class _ModelStruct(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32), ("ref_coordinates", C.c_void_p), ("copy_coordinates", C.c_void_p),
        ("n_materials", C.c_int32), ("hooke", C.c_void_p),
        ("n_sections", C.c_int32), ("sections", C.c_void_p),
        ("n_shell_sections", C.c_int32), ("shell_thickness", C.c_void_p),
        ("n_cs", C.c_int32), ("cs", C.c_void_p),
        ("n_elements", C.c_int32), ("elem_type", C.c_void_p), ("elem_material", C.c_void_p),
        ("elem_section", C.c_void_p), ("elem_cs", C.c_void_p), ("elem_node_ptr", C.c_void_p),
        ("elem_nodes", C.c_void_p), ("beam_pretension", C.c_void_p),
        ("gravity_on", C.c_int32), ("gravity", C.c_double * 3),
        ("part_rank", C.c_int32), ("part_world", C.c_int32),
        ("n_pipe_sections", C.c_int32), ("pipe_sections", C.c_void_p),
    ]


class _NormsStruct(C.Structure):
    _fields_ = [("max_force", C.c_double), ("max_moment", C.c_double), ("node_force", C.c_int32), ("node_moment", C.c_int32),
                ("max_disp_value", C.c_double), ("max_rot_value", C.c_double), ("nan_detected", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class _StepStruct(C.Structure):
    _fields_ = [("displacements", C.c_void_p), ("displacements_on_device", C.c_int32), ("gravity_factor", C.c_double)]


class _DynamicStruct(C.Structure):
    """gfa_dynamic_t: Dynamic::a1..a6, alpha, beta and the MountDamping(update_rayleigh) flag"""
    _fields_ = [("a1", C.c_double), ("a2", C.c_double), ("a3", C.c_double), ("a4", C.c_double), ("a5", C.c_double),
                ("a6", C.c_double), ("rayleigh_alpha", C.c_double), ("rayleigh_beta", C.c_double), ("update_rayleigh", C.c_int32)]


_lib = None


def load_library() -> C.CDLL:
    """Load libgfa.so; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GfaError(-2, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        lib.gfa_last_error.restype = C.c_char_p
        lib.gfa_create.argtypes = [C.POINTER(_ModelStruct), C.c_int, C.POINTER(C.c_void_p)]
        lib.gfa_destroy.argtypes = [C.c_void_p]
        lib.gfa_number_dofs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.gfa_set_dofs.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gfa_csr_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        lib.gfa_csr_pattern.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.gfa_assemble.argtypes = [C.c_void_p, C.POINTER(_StepStruct)]
        lib.gfa_assemble_enqueue.argtypes = [C.c_void_p, C.POINTER(_StepStruct)]
        lib.gfa_add_host_triplets.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gfa_add_host_vector.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        lib.gfa_csr_values.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.gfa_csr_values_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.gfa_vector.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.gfa_vector_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.gfa_element_block.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.gfa_commit_state.argtypes = [C.c_void_p]
        lib.gfa_element_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        lib.gfa_residual.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gfa_update_displacements.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gfa_displacements.argtypes = [C.c_void_p, C.c_void_p]
        lib.gfa_results_stride.argtypes = [C.c_int]
        lib.gfa_gauss_point_results.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        lib.gfa_gauss_point_results.restype = C.c_int64
        lib.gfa_copy_coordinates.argtypes = [C.c_void_p, C.c_void_p]
        lib.gfa_last_timing.argtypes = [C.c_void_p, C.c_void_p]
        lib.gfa_last_launch_count.argtypes = [C.c_void_p]
        lib.gfa_interface_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gfa_interface_pack.argtypes = [C.c_void_p, C.c_void_p]
        lib.gfa_interface_unpack.argtypes = [C.c_void_p, C.c_void_p]
        lib.gfa_owned_rows.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
        lib.gfa_local_rows.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
        lib.gfa_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        lib.gfa_interface_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        lib.gfa_set_kinematics.argtypes = [C.c_void_p] * 5
        lib.gfa_kinematics.argtypes = [C.c_void_p] * 5
        lib.gfa_update_dyn.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_DynamicStruct)]
        lib.gfa_assemble_dynamic.argtypes = [C.c_void_p, C.POINTER(_StepStruct), C.POINTER(_DynamicStruct)]
        lib.gfa_element_alpha_i.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        lib.gfa_pipeline_info.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
        lib.gfa_touched_nodes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
        lib.gfa_set_displacements_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        lib.gfa_vector_owned.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.gfa_set_shell_loads.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gfa_apply_shell_loads.argtypes = [C.c_void_p, C.c_void_p]
        lib.gfa_set_pipe_loads.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.gfa_apply_pipe_loads.argtypes = [C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Assembler:
    """Owns one ``gfa_t`` handle (one GPU).  Method names follow the reference's
    ``Solution`` steps they replace."""

    def __init__(self, model: Model, device: int = 0, rank: int = 0, world: int = 1):
        self.lib = load_library()
        self.model = model
        self._h = C.c_void_p()
        self._keep = []
        ms = _ModelStruct()

        def arr(x, dt):
            a = np.ascontiguousarray(x, dt)
            self._keep.append(a)
            return a

        xyz = arr(model.xyz, np.float64)
        ms.n_nodes = model.n_nodes
        ms.ref_coordinates = _ptr(xyz)
        ms.copy_coordinates = None
        hooke = arr(model.hooke, np.float64)
        ms.n_materials, ms.hooke = len(model.hooke), _ptr(hooke)
        sec = arr(model.sections, np.float64)
        ms.n_sections, ms.sections = len(model.sections), _ptr(sec) if len(model.sections) else None
        th = arr(model.shell_thickness, np.float64)
        ms.n_shell_sections, ms.shell_thickness = len(th), _ptr(th) if len(th) else None
        cs = arr(model.cs, np.float64)
        ms.n_cs, ms.cs = len(model.cs), _ptr(cs) if len(model.cs) else None
        ms.n_elements = model.n_elements
        ms.elem_type = _ptr(arr(model.elem_type, np.int32))
        ms.elem_material = _ptr(arr(model.elem_mat, np.int32))
        ms.elem_section = _ptr(arr(model.elem_sec, np.int32))
        ms.elem_cs = _ptr(arr(model.elem_cs, np.int32))
        ms.elem_node_ptr = _ptr(arr(model.elem_ptr, np.int32))
        ms.elem_nodes = _ptr(arr(model.elem_nodes, np.int32))
        ms.beam_pretension = _ptr(arr(model.pretension, np.float64)) if model.pretension is not None else None
        ms.gravity_on = 1 if model.gravity is not None else 0
        g = model.gravity if model.gravity is not None else (0.0, 0.0, 0.0)
        ms.gravity = (C.c_double * 3)(*[float(v) for v in g])
        ms.part_rank, ms.part_world = rank, world
        pipes = arr(np.asarray(getattr(model, 'pipe_sections', np.zeros((0, 11))), float).reshape(-1, 11), np.float64)
        ms.n_pipe_sections, ms.pipe_sections = len(pipes), _ptr(pipes) if len(pipes) else None
        self._check(self.lib.gfa_create(C.byref(ms), device, C.byref(self._h)))
        self.n_free = self.n_fixed = 0
        self.gravity_factor = 1.0

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise GfaError(rc, self.lib.gfa_last_error().decode())
        return rc

    def close(self):
        if self._h:
            self.lib.gfa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- SetGlobalDOFs / SetGlobalSize -----------------------------------
    def number_dofs(self, constraint_mask=None):
        m = self.model
        cm = np.ascontiguousarray(m.constraint_mask() if constraint_mask is None else constraint_mask, np.int32)
        gls = np.zeros(m.n_nodes * 6, np.int32)
        nf, nx = C.c_int32(), C.c_int32()
        self._check(self.lib.gfa_number_dofs(self._h, _ptr(cm), _ptr(gls), C.byref(nf), C.byref(nx)))
        return gls.reshape(-1, 6), nf.value, nx.value

    def set_dofs(self, gls=None, n_free=None, n_fixed=None, extra=None):
        if gls is None:
            gls, n_free, n_fixed = self.number_dofs()
        gls = np.ascontiguousarray(gls, np.int32).reshape(-1)
        n_extra, em, er, ec = 0, None, None, None
        if extra is not None and len(extra[0]):
            em, er, ec = (np.ascontiguousarray(x, np.int32) for x in extra)
            n_extra = len(em)
        self._check(self.lib.gfa_set_dofs(self._h, _ptr(gls), int(n_free), int(n_fixed), n_extra, _ptr(em), _ptr(er), _ptr(ec)))
        self.gls = gls.reshape(-1, 6)
        self.n_free, self.n_fixed = int(n_free), int(n_fixed)
        return self

    def csr_dims(self, which="AA"):
        r, c, nz = C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self.lib.gfa_csr_dims(self._h, MATS[which], C.byref(r), C.byref(c), C.byref(nz)))
        return r.value, c.value, nz.value

    def csr_pattern(self, which="AA"):
        r, _, nz = self.csr_dims(which)
        outer = np.zeros(r + 1, np.int32)
        inner = np.zeros(nz, np.int32)
        self._check(self.lib.gfa_csr_pattern(self._h, MATS[which], _ptr(outer), _ptr(inner)))
        return outer, inner

    # ---- per iteration -----------------------------------------------------
    def set_time(self, last_converged: float, step: float, start: float = 0.0, end: float = 1.0):
        """Gravity ramp of a first solution step (BoolTable.cpp:84-106)."""
        self.gravity_factor = (last_converged + step - start) / (end - start)

    def assemble(self, disp, device_ptr: int | None = None):
        st = _StepStruct()
        if device_ptr is not None:
            st.displacements, st.displacements_on_device = device_ptr, 1
        elif disp is None:           # keep the device copy (after update_displacements)
            st.displacements, st.displacements_on_device = None, 0
        else:
            d = np.ascontiguousarray(disp, np.float64).reshape(-1)
            self._keep_disp = d
            st.displacements, st.displacements_on_device = d.ctypes.data, 0
        st.gravity_factor = float(self.gravity_factor)
        self._check(self.lib.gfa_assemble(self._h, C.byref(st)))
        return self

    # ---- Newmark dynamics (Dynamic.cpp:303-340) ------------------------------
    def set_dynamic(self, newmark6, rayleigh_alpha: float = 0.0, rayleigh_beta: float = 0.0):
        """Dynamic::a1..a6 of the time step (Dynamic.cpp:582-590) and the Rayleigh coefficients."""
        d = _DynamicStruct()
        d.a1, d.a2, d.a3, d.a4, d.a5, d.a6 = [float(v) for v in newmark6]
        d.rayleigh_alpha, d.rayleigh_beta, d.update_rayleigh = float(rayleigh_alpha), float(rayleigh_beta), 0
        self._dyn = d
        return self

    def set_kinematics(self, vel=None, accel=None, copy_vel=None, copy_accel=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64).reshape(-1) for a in (vel, accel, copy_vel, copy_accel)]
        self._check(self.lib.gfa_set_kinematics(self._h, *[None if a is None else a.ctypes.data for a in arrs]))

    def kinematics(self):
        """(vel, accel, copy_vel, copy_accel), each [n_nodes, 6]"""
        out = [np.zeros(self.model.n_nodes * 6) for _ in range(4)]
        self._check(self.lib.gfa_kinematics(self._h, *[a.ctypes.data for a in out]))
        return tuple(a.reshape(-1, 6) for a in out)

    def update_dyn(self, disp=None):
        """Dynamic::UpdateDyn; disp None = the device copy of Node::displacements"""
        d = None if disp is None else np.ascontiguousarray(disp, np.float64).reshape(-1)
        self._check(self.lib.gfa_update_dyn(self._h, None if d is None else d.ctypes.data, C.byref(self._dyn)))

    def assemble_dynamic(self, disp, update_rayleigh: bool):
        st = _StepStruct()
        if disp is None:
            st.displacements, st.displacements_on_device = None, 0
        else:
            d = np.ascontiguousarray(disp, np.float64).reshape(-1)
            self._keep_disp = d
            st.displacements, st.displacements_on_device = d.ctypes.data, 0
        st.gravity_factor = float(self.gravity_factor)
        self._dyn.update_rayleigh = 1 if update_rayleigh else 0
        self._check(self.lib.gfa_assemble_dynamic(self._h, C.byref(st), C.byref(self._dyn)))
        return self

    def alpha_i(self, e: int):
        buf = np.zeros(16)
        n = self._check(self.lib.gfa_element_alpha_i(self._h, e, _ptr(buf)))
        return buf[:n].copy()

    def assemble_enqueue(self, device_ptr: int | None = None):
        """gfa_assemble_enqueue: queue the assembly on the library's stream and return (device displacements or
        the device copy); reads wait for the stream."""
        st = _StepStruct()
        st.displacements, st.displacements_on_device = device_ptr, 1 if device_ptr is not None else 0
        st.gravity_factor = float(self.gravity_factor)
        self._check(self.lib.gfa_assemble_enqueue(self._h, C.byref(st)))
        return self

    def assemble_raw(self, host_ptr: int):
        """Host pointer (e.g. pinned memory) without numpy marshalling."""
        st = _StepStruct()
        st.displacements, st.displacements_on_device = host_ptr, 0
        st.gravity_factor = float(self.gravity_factor)
        self._check(self.lib.gfa_assemble(self._h, C.byref(st)))

    def add_host_triplets(self, which, rows, cols, vals):
        rows, cols = np.ascontiguousarray(rows, np.int32), np.ascontiguousarray(cols, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        self._check(self.lib.gfa_add_host_triplets(self._h, MATS[which], len(vals), _ptr(rows), _ptr(cols), _ptr(vals)))

    def add_host_vector(self, which_vector, index, vals):
        index, vals = np.ascontiguousarray(index, np.int32), np.ascontiguousarray(vals, np.float64)
        self._check(self.lib.gfa_add_host_vector(self._h, which_vector, len(vals), _ptr(index), _ptr(vals)))

    # ---- ShellLoad follower pressure on the device ----------------------------
    def set_shell_loads(self, loads):
        """loads: [(element ids 1-based, area_update, table)] as in Model.shell_loads; call after set_dofs"""
        ptr = np.zeros(len(loads) + 1, np.int32)
        for k, (els, _, _) in enumerate(loads):
            ptr[k + 1] = ptr[k] + len(els)
        elems = np.concatenate([np.asarray(els, np.int32) - 1 for els, _, _ in loads]).astype(np.int32) if loads else np.zeros(0, np.int32)
        area = np.array([1 if au else 0 for _, au, _ in loads], np.int32)
        self._shell_load_tables = [np.asarray(t, float) for _, _, t in loads]
        self._check(self.lib.gfa_set_shell_loads(self._h, len(loads), _ptr(ptr), _ptr(elems), _ptr(area)))

    # ---- PipeLoad internal pressure on the device --------------------------------
    def set_pipe_loads(self, loads):
        """loads: [(element ids 1-based, table[n,5] = time P0I P0E RhoI RhoE)] as in Model.pipe_loads; call after set_dofs"""
        ptr = np.zeros(len(loads) + 1, np.int32)
        for k, (elements, _) in enumerate(loads):
            ptr[k + 1] = ptr[k] + len(elements)
        elems = np.concatenate([np.asarray(e, np.int32) - 1 for e, _ in loads]).astype(np.int32) if loads else np.zeros(1, np.int32)
        self._pipe_load_tables = [np.asarray(t, float) for _, t in loads]
        self._check(self.lib.gfa_set_pipe_loads(self._h, len(loads), _ptr(ptr), _ptr(elems)))

    def apply_pipe_loads(self, time: float):
        """Load::GetValueAt(time, 0) of every registered PipeLoad (linear table), then gfa_apply_pipe_loads"""
        p = np.array([float(np.interp(time, t[:, 0], t[:, 1])) for t in self._pipe_load_tables] or [0.0])
        self._check(self.lib.gfa_apply_pipe_loads(self._h, _ptr(p)))

    def apply_shell_loads(self, time: float):
        """ShellLoad::GetValueAt(time) of every registered load (linear table), then gfa_apply_shell_loads"""
        p = np.array([float(np.interp(time, t[:, 0], t[:, 1])) for t in self._shell_load_tables])
        self._check(self.lib.gfa_apply_shell_loads(self._h, _ptr(p)))

    def commit(self):
        self._check(self.lib.gfa_commit_state(self._h))

    # ---- results -----------------------------------------------------------
    def csr(self, which="AA"):
        outer, inner = self.csr_pattern(which)
        val = np.zeros(len(inner), np.float64)
        self._check(self.lib.gfa_csr_values(self._h, MATS[which], _ptr(val)))
        r, c, _ = self.csr_dims(which)
        return outer, inner, val, (r, c)

    def values(self, which="AA", out=None):
        _, _, nz = self.csr_dims(which)
        val = np.zeros(nz, np.float64) if out is None else out
        self._check(self.lib.gfa_csr_values(self._h, MATS[which], _ptr(val)))
        return val

    def values_device(self, which="AA") -> int:
        p = C.c_void_p()
        self._check(self.lib.gfa_csr_values_device(self._h, MATS[which], C.byref(p)))
        return p.value or 0

    def vector_device(self, which_vector) -> int:
        p = C.c_void_p()
        self._check(self.lib.gfa_vector_device(self._h, which_vector, C.byref(p)))
        return p.value or 0

    def vectors(self):
        out = []
        for w, n in ((P_A, self.n_free), (I_A, self.n_free), (P_B, self.n_fixed)):
            v = np.zeros(n, np.float64)
            self._check(self.lib.gfa_vector(self._h, w, _ptr(v)))
            out.append(v)
        return tuple(out)

    def vector(self, which_vector, out):
        self._check(self.lib.gfa_vector(self._h, which_vector, _ptr(out)))
        return out

    def element(self, e: int):
        from .meshes import DOFS_PER_TYPE
        n = DOFS_PER_TYPE[int(self.model.elem_type[e])]
        K, P = np.zeros((n, n)), np.zeros(n)
        self._check(self.lib.gfa_element_block(self._h, e, _ptr(K), _ptr(P)))
        return K, P

    def state(self, e: int):
        buf = np.zeros(64)
        n = self._check(self.lib.gfa_element_state(self._h, e, _ptr(buf)))
        return buf[:n].copy()

    def gauss_point_results(self, element_type: int):
        """Result read-back (gfa_gauss_point_results): array [n_elements_of_type, stride]; column 0 is
        the strain energy, the rest the per-point strains / resultants (see include/gfa.h)."""
        stride = self.lib.gfa_results_stride(int(element_type))
        if stride == 0:
            raise GfaError(-7, f"element type {element_type} keeps no Gauss-point results")
        n = int(np.count_nonzero(self.model.elem_type == element_type))      # upper bound: this rank holds a part
        buf = np.zeros((max(n, 1), stride))
        got = self.lib.gfa_gauss_point_results(self._h, int(element_type), _ptr(buf), buf.size)
        if got < 0:
            raise GfaError(int(got), self.lib.gfa_last_error().decode())
        return buf[:got]

    # ---- Newton-loop vector steps on the device copies (include/gfa.h) ----
    def residual(self, X_B=None) -> dict:
        """P_A = -P_A (- K_AB X_B); returns the max-norms CheckResidualConvergence reads."""
        n = _NormsStruct()
        xb = np.ascontiguousarray(X_B, np.float64) if X_B is not None else None
        self._check(self.lib.gfa_residual(self._h, _ptr(xb) if xb is not None else None, C.byref(n)))
        return n.as_dict()

    def update_displacements(self, x_A) -> dict:
        """Solution::UpdateDisps on the device copy; returns the norms CheckGLConvergence reads."""
        n = _NormsStruct()
        x = np.ascontiguousarray(x_A, np.float64)
        self._check(self.lib.gfa_update_displacements(self._h, _ptr(x), C.byref(n)))
        return n.as_dict()

    def displacements(self) -> np.ndarray:
        d = np.zeros(self.model.n_nodes * 6)
        self._check(self.lib.gfa_displacements(self._h, _ptr(d)))
        return d.reshape(-1, 6)

    def copy_coordinates(self):
        c = np.zeros(self.model.n_nodes * 6)
        self._check(self.lib.gfa_copy_coordinates(self._h, _ptr(c)))
        return c.reshape(-1, 6)

    def timing(self):
        t = np.zeros(4)
        self._check(self.lib.gfa_last_timing(self._h, _ptr(t)))
        return {"h2d_ms": t[0], "eval_ms": t[1], "scatter_ms": t[2], "total_ms": t[3]}

    def launch_count(self) -> int:
        return self.lib.gfa_last_launch_count(self._h)

    def pipeline_info(self):
        """(ring, description): which pipeline gfa_set_dofs chose (fused ring kernel or classic two kernels)"""
        buf = C.create_string_buffer(512)
        rc = self._check(self.lib.gfa_pipeline_info(self._h, buf, 512))
        return bool(rc), buf.value.decode()

    # ---- multi-GPU interface rows -----------------------------------------
    def interface_counts(self, world: int):
        s, r = np.zeros(world, np.int64), np.zeros(world, np.int64)
        self._check(self.lib.gfa_interface_counts(self._h, _ptr(s), _ptr(r)))
        return s, r

    def interface_pack(self, device_ptr: int):
        self._check(self.lib.gfa_interface_pack(self._h, device_ptr))

    def interface_unpack(self, device_ptr: int):
        self._check(self.lib.gfa_interface_unpack(self._h, device_ptr))

    def local_rows(self):
        n = C.c_int64()
        self._check(self.lib.gfa_local_rows(self._h, C.byref(n), None))
        rows = np.zeros(n.value, np.int32)
        self._check(self.lib.gfa_local_rows(self._h, C.byref(n), _ptr(rows)))
        return rows

    def owned_rows(self):
        n = C.c_int64()
        self._check(self.lib.gfa_owned_rows(self._h, C.byref(n), None))
        rows = np.zeros(n.value, np.int32)
        self._check(self.lib.gfa_owned_rows(self._h, C.byref(n), _ptr(rows)))
        return rows

    # ---- partition-local transfers -----------------------------------------
    def touched_nodes(self):
        """0-based indices of the nodes this rank's elements reference (ascending)"""
        n = C.c_int64()
        self._check(self.lib.gfa_touched_nodes(self._h, C.byref(n), None))
        nodes = np.zeros(n.value, np.int32)
        self._check(self.lib.gfa_touched_nodes(self._h, C.byref(n), _ptr(nodes)))
        return nodes

    def set_displacements_packed(self, ptr: int, on_device: bool = False):
        """Node::displacements of the touched nodes only, [n_touched*6], raw host (e.g. pinned) or device pointer"""
        self._check(self.lib.gfa_set_displacements_packed(self._h, ptr, 1 if on_device else 0))

    def vector_owned(self, which_vector, out):
        """P_A / I_A entries of the rows this rank owns (order of owned_rows()); P_B whole"""
        self._check(self.lib.gfa_vector_owned(self._h, which_vector, _ptr(out)))
        return out

    def stream(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.gfa_stream(self._h, C.byref(p)))
        return p.value or 0

    def interface_stream(self) -> int:
        """cudaStream_t of the interface exchange: pack / unpack are enqueued there, the caller's transport goes
        between them on the same stream"""
        p = C.c_void_p()
        self._check(self.lib.gfa_interface_stream(self._h, C.byref(p)))
        return p.value or 0
