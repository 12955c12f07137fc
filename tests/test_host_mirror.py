"""The C++ host mirror (giraffe_b200/host: GfaHost + gfa_run) and the .inp subset
reader/writer.  CPU: parsing and DOF numbering.  GPU: the reference's call
sequence (ReadFile, PreCalc, DOFsActive, SetGlobalDOFs, SetGlobalSize, Clear,
MountLocal, MountElementLoads, MountLoads, MountGlobal, MountSparse) driven from
C++ agrees with the Python binding of the same C-ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

import util
from giraffe_b200 import meshes as M
from giraffe_b200.inp import read_inp, write_inp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "giraffe_b200", "gfa_run")


def _model():
    m = M.concat_models([M.beam_line(6, pretension=1.0e4), M.shell_plate(3, 2, warp=0.01), M.pipe_line(4)])
    m.gravity = (0.0, 0.0, -9.81)
    m.nodal_loads = [(np.array([13], np.int32), 1, np.array([[0, 0, 0, 0, 0, 0, 0], [1, 100.0, 0, 50.0, 0, 3.0, 0]]))]
    return m


def test_inp_round_trip(tmp_path):
    m = _model()
    p = str(tmp_path / "model.inp")
    write_inp(m, p, end_time=1.0, time_step=0.25)
    m2, info = read_inp(p)
    assert info["time_step"] == 0.25
    for k in ("xyz", "hooke", "sections", "cs", "elem_type", "elem_mat", "elem_sec", "elem_cs", "elem_nodes", "shell_thickness", "pipe_sections"):
        assert np.array_equal(getattr(m, k), getattr(m2, k)), k
    assert np.array_equal(m.pretension, m2.pretension)
    assert (M.number_dofs(m)[0] == M.number_dofs(m2)[0]).all()
    assert m2.gravity == m.gravity and len(m2.nodal_loads) == 1


def test_cpp_reader_and_dof_numbering(tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("gfa_run not built")
    m = _model()
    p = str(tmp_path / "model.inp")
    write_inp(m, p, time_step=0.25)
    out = json.loads(subprocess.check_output([EXE, "--parse-only", p]))
    _, nf, nx = M.number_dofs(m)
    assert (out["nodes"], out["elements"], out["n_GL_free"], out["n_GL_fixed"]) == (m.n_nodes, m.n_elements, nf, nx)
    assert out["loads"] == 1 and out["time_step"] == 0.25


@pytest.mark.gpu
def test_cpp_host_sequence_matches_python_binding(tmp_path):
    from giraffe_b200 import capi
    m = _model()
    p = str(tmp_path / "model.inp")
    write_inp(m, p, time_step=0.25)
    out = json.loads(subprocess.check_output([EXE, p]))
    asm = capi.Assembler(m)
    gls, nf, nx = asm.number_dofs()
    asm.set_dofs(gls, nf, nx)
    asm.set_time(0.0, 0.25)
    d = np.zeros((m.n_nodes, 6))
    asm.assemble(d)
    trip, pa_add, pb_add = util.nodal_load_contribution(m, gls, d, 0.25)
    for w in ("AA", "AB", "BA", "BB"):
        if trip[w][0]:
            asm.add_host_triplets(w, *trip[w])
    asm.add_host_vector(capi.P_A, *pa_add)
    val = asm.values("AA")
    pa = asm.vectors()[0]
    assert out["n_GL_free"] == nf and out["nnz_AA"] == len(val)
    assert abs(out["sum_AA"] - val.sum()) <= 1e-9 * np.abs(val).sum()
    assert abs(out["max_AA"] - np.abs(val).max()) <= 1e-12 * np.abs(val).max()
    assert abs(out["max_P_A"] - np.abs(pa).max()) <= 1e-12 * np.abs(pa).max()


@pytest.mark.gpu
def test_cpp_host_pipe_load_matches_python_binding(tmp_path):
    """PipeLoad through the .inp writer, the C++ reader, GfaHost::SetGlobalSize (gfa_set_pipe_loads) and MountLoads
    (gfa_apply_pipe_loads) against the Python binding of the same entry points."""
    from giraffe_b200 import capi
    z = np.load(os.path.join(util.GOLDEN_DIR, "pipe_load.npz"))
    m = util.model_from_dict(z)
    p = str(tmp_path / "pipes.inp")
    write_inp(m, p, time_step=0.25)
    m2, _ = read_inp(p)
    assert len(m2.pipe_loads) == 2
    for (e1, t1), (e2, t2) in zip(m.pipe_loads, m2.pipe_loads):
        assert np.array_equal(e1, e2) and np.array_equal(np.asarray(t1, float), t2)
    out = json.loads(subprocess.check_output([EXE, p]))
    assert out["pipe_loads"] == 2
    asm = capi.Assembler(m).set_dofs()
    asm.set_time(0.0, 0.25)
    asm.set_pipe_loads(m.pipe_loads)
    asm.assemble(np.zeros((m.n_nodes, 6)))
    no_load = asm.values("AA").copy()
    asm.apply_pipe_loads(0.25)
    val = asm.values("AA")
    assert np.abs(val - no_load).max() > 0.0
    assert out["nnz_AA"] == len(val)
    assert abs(out["sum_AA"] - val.sum()) <= 1e-9 * np.abs(val).sum()
    assert abs(out["max_AA"] - np.abs(val).max()) <= 1e-12 * np.abs(val).max()


@pytest.mark.gpu
def test_cpp_static_solve_with_a_follower_load(tmp_path, ref):
    """NodalFollowerLoad in the C++ host mirror (GfaHost::MountLoads: Q_i from the committed rotations fetched with
    gfa_copy_coordinates): a cantilever bent by a follower force and moment at its tip, four increments of
    `gfa_run --solve`, against the same Newton loop through the reference's own sources."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    m = M.beam_line(8)
    tip = m.n_nodes
    m.follower_loads = [(np.array([tip], np.int32), 1, np.array([[0.0, 0, 0, 0, 0, 0, 0], [1.0, 1.2e6, -6.0e5, 0.0, 0.0, 1.8e6, 3.0e5]]))]
    p = str(tmp_path / "follower.inp")
    write_inp(m, p, end_time=1.0, time_step=0.25)
    m2, _ = read_inp(p)
    assert len(m2.follower_loads) == 1 and np.array_equal(m2.follower_loads[0][2], m.follower_loads[0][2])
    out = json.loads(subprocess.check_output([EXE, "--solve", p]))
    assert out["increments"] == 4
    got = np.array(out["copy_coordinates"]).reshape(-1, 6)
    ref.load(m)
    t = 0.0
    for inc in range(4):
        ref.set_time(t, 0.25)
        d = np.zeros((m.n_nodes, 6))
        for it in range(8):
            ref.assemble(d, with_loads=True)
            ref.residual(None)
            o, i, v, shape = ref.csr("AA")
            x = spla.spsolve(sp.csr_matrix((v, i, o), shape=shape).tocsc(), ref.vectors()[0])
            d = ref.update_displacements(x)[0]
        ref.commit()
        t += 0.25
    assert np.abs(ref.copy_coordinates()[:, 3:]).max() > 0.05, "the load must rotate the tip for the test to mean anything"
    util.assert_parity(ref.copy_coordinates(), got, "cantilever under a follower load, gfa_run --solve", tol=1e-9)


def _dynamic_model():
    m = M.concat_models([M.beam_line(6, pretension=1.0e4), M.shell_plate(3, 2, warp=0.01)])
    m.gravity = (0.0, 0.0, -9.81)
    return m


DYN = {"alpha": 0.4, "beta": 2.0e-4, "update": 0, "beta_new": 0.3, "gamma_new": 0.5}


def test_dynamic_step_round_trip(tmp_path):
    """Dynamic solution step (Dynamic::Read, Dynamic.cpp:65-222) through the .inp writer, the Python reader and
    the C++ reader."""
    m = _dynamic_model()
    p = str(tmp_path / "dyn.inp")
    write_inp(m, p, end_time=2.0, time_step=0.01, dynamic=DYN)
    _, info = read_inp(p)
    assert info["time_step"] == 0.01 and info["dynamic"] == DYN
    if not os.path.exists(EXE):
        pytest.skip("gfa_run not built")
    out = json.loads(subprocess.check_output([EXE, "--parse-only", p]))
    assert out["dynamic"] == 1 and out["time_step"] == 0.01
    assert (out["alpha"], out["beta"], out["update"], out["beta_new"], out["gamma_new"]) == (0.4, 2.0e-4, 0, 0.3, 0.5)


@pytest.mark.gpu
def test_cpp_host_dynamic_sequence_matches_python_binding(tmp_path):
    """CalculateNewmarkCoeff, UpdateDyn, MountLocal .. MountMass, MountDamping(true), MountDyn .. MountSparse
    driven from C++ (gfa_run) against the Python binding of the same C-ABI."""
    from giraffe_b200 import capi
    m = _dynamic_model()
    p = str(tmp_path / "dyn.inp")
    write_inp(m, p, end_time=2.0, time_step=0.01, dynamic=DYN)
    out = json.loads(subprocess.check_output([EXE, p]))
    asm = capi.Assembler(m).set_dofs()
    asm.set_time(0.0, 0.01, 0.0, 2.0)
    asm.set_dynamic(util.newmark_coefficients(0.01), DYN["alpha"], DYN["beta"])
    d = np.zeros((m.n_nodes, 6))
    asm.update_dyn(d)
    asm.assemble_dynamic(d, True)
    val = asm.values("AA")
    pa = asm.vectors()[0]
    assert out["nnz_AA"] == len(val)
    assert abs(out["sum_AA"] - val.sum()) <= 1e-9 * np.abs(val).sum()
    assert abs(out["max_AA"] - np.abs(val).max()) <= 1e-12 * np.abs(val).max()
    assert abs(out["max_P_A"] - np.abs(pa).max()) <= 1e-12 * np.abs(pa).max()


def _pressure_model():
    m = M.concat_models([M.beam_line(4), M.shell_plate(4, 3, warp=0.02)])
    m.gravity = (0.0, 0.0, -9.81)
    m.shell_loads = [(np.array([6, 9, 10, 20], np.int32), False, np.array([[0.0, 0.0], [1.0, 4.0e7]])),
                     (np.array([13, 14], np.int32), True, np.array([[0.0, 0.0], [1.0, -2.0e7]]))]
    return m


def test_shell_load_inp_round_trip(tmp_path):
    m = _pressure_model()
    p = str(tmp_path / "pressure.inp")
    write_inp(m, p, time_step=0.5)
    m2, _ = read_inp(p)
    assert len(m2.shell_loads) == 2
    for (e1, a1, t1), (e2, a2, t2) in zip(m.shell_loads, m2.shell_loads):
        assert np.array_equal(e1, e2) and a1 == a2 and np.array_equal(t1, t2)
    if os.path.exists(EXE):
        out = json.loads(subprocess.check_output([EXE, "--parse-only", p]))
        assert out["shell_loads"] == 2 and out["element_sets"] == 2


@pytest.mark.gpu
def test_cpp_host_shell_load_matches_python_host_step(tmp_path):
    """GfaHost::MountLoads with ShellLoad (C++) against the Python restatement of the same host step
    (tests/util.py, pinned to the reference by tests/golden/shell_load.npz), both over the C-ABI."""
    from giraffe_b200 import capi
    m = _pressure_model()
    p = str(tmp_path / "pressure.inp")
    write_inp(m, p, time_step=0.5)
    out = json.loads(subprocess.check_output([EXE, p]))
    asm = capi.Assembler(m).set_dofs()
    asm.set_time(0.0, 0.5)
    d = np.zeros((m.n_nodes, 6))
    asm.assemble(d)
    trip, pa_add, pb_add = util.shell_load_contribution(m, asm.gls, d, asm.copy_coordinates(), 0.5)
    for w in ("AA", "AB", "BA", "BB"):
        if trip[w][0]:
            asm.add_host_triplets(w, *trip[w])
    asm.add_host_vector(capi.P_A, *pa_add)
    val = asm.values("AA")
    pa = asm.vectors()[0]
    assert out["nnz_AA"] == len(val)
    assert abs(out["sum_AA"] - val.sum()) <= 1e-9 * np.abs(val).sum()
    assert abs(out["max_AA"] - np.abs(val).max()) <= 1e-12 * np.abs(val).max()
    assert abs(out["max_P_A"] - np.abs(pa).max()) <= 1e-12 * np.abs(pa).max()


@pytest.mark.gpu
def test_cpp_static_solve_matches_the_reference_newton_loop(tmp_path, ref):
    """`gfa_run --solve`: Static::Solve's loop in C++ over the host mirror -- device assembly, MountLoads, gfa_residual,
    a host solve, UpdateDisps, SaveConfiguration, ten increments of the shipped tutorial01 -- against the same loop
    through the reference's own sources (RefOracle)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    z = np.load(os.path.join(util.GOLDEN_DIR, "tutorial01.npz"))
    m = util.model_from_dict(z)
    dt = float(z["time"][1])
    p = str(tmp_path / "tutorial01.inp")
    write_inp(m, p, end_time=1.0, time_step=dt)
    out = json.loads(subprocess.check_output([EXE, "--solve", p]))
    assert out["increments"] == 10
    got = np.array(out["copy_coordinates"]).reshape(-1, 6)
    ref.load(m)
    t = 0.0
    for inc in range(10):
        ref.set_time(t, dt)
        d = np.zeros((m.n_nodes, 6))
        for it in range(8):
            ref.assemble(d, with_loads=True)
            ref.residual(None)
            o, i, v, shape = ref.csr("AA")
            x = spla.spsolve(sp.csr_matrix((v, i, o), shape=shape).tocsc(), ref.vectors()[0])
            d = ref.update_displacements(x)[0]
        ref.commit()
        t += dt
    util.assert_parity(ref.copy_coordinates(), got, "tutorial01 solved by gfa_run --solve", tol=1e-9)
    assert out["last_max_dx"] < 1e-9
