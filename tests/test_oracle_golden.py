"""CPU: the port restatement (oracle/port) against the committed golden
fixtures that the reference's own sources produced (tests/golden/make_golden.py).
This is what pins the oracle on boxes where /root/reference does not exist."""
import os

import numpy as np
import pytest

import util


def _load(name):
    return np.load(os.path.join(util.GOLDEN_DIR, name + ".npz"))


def test_tutorial01_structure_and_values(port):
    """BASELINE.json configs[0]: inputs/tutorial01 as shipped, iterations 1 and 2
    of increment 1.  Structural facts from SURVEY.md 8c: 60 free / 6 fixed DOF,
    1449 AA triplets (9 NodalLoad + 144 + 4*324), nnz_AA = 1296, AB 72, BB 36."""
    z = _load("tutorial01")
    m = util.model_from_dict(z)
    port.load(m)
    assert (port.n_free, port.n_fixed) == (60, 6)
    assert (port.gls() == z["gls"]).all()
    assert list(z["it1_triplets"]) == [1449, 72, 72, 36]
    t0, dt = z["time"]
    for tag in ("it1", "it2"):
        disp = z[f"{tag}_disp"]
        trip, pa_add, pb_add = util.nodal_load_contribution(m, port.gls(), disp, t0 + dt)
        for w in ("AA", "AB", "BA", "BB"):
            port.set_extra_triplets(w, *trip[w])
        port.assemble(disp)
        if tag == "it1":
            assert port.triplets("AA") == 1449 and port.triplets("AB") == 72 and port.triplets("BB") == 36
        util.assert_system_parity(lambda w: util.captured_csr(z, tag, w), port.csr, f"tutorial01 {tag}")
        pa, ia, pb = port.vectors()
        np.add.at(pa, pa_add[0], pa_add[1])
        np.add.at(pb, pb_add[0], pb_add[1])
        util.assert_parity(z[f"{tag}_PA"], pa, f"tutorial01 {tag} P_A")
        util.assert_parity(z[f"{tag}_IA"], ia, f"tutorial01 {tag} I_A")
        util.assert_parity(z[f"{tag}_PB"], pb, f"tutorial01 {tag} P_B")
    assert len(z["it1_AA_val"]) == 1296


@pytest.mark.parametrize("name", ["beam_line", "pipe_line", "shell_plate"])
def test_sequence_with_commit(port, name):
    """Two iterations, SaveLagrange/SaveConfiguration, one more iteration."""
    z = _load(name)
    m = util.model_from_dict(z)
    port.load(m)
    port.set_time(*z["time"])
    assert (port.gls() == z["gls"]).all()
    for tag, commit in (("it1", False), ("it2", True), ("it3", False)):
        port.assemble(z[f"{tag}_disp"])
        util.assert_system_parity(lambda w: util.captured_csr(z, tag, w), port.csr, f"{name} {tag}")
        for v, key in zip(port.vectors(), ("PA", "IA", "PB")):
            util.assert_parity(z[f"{tag}_{key}"], v, f"{name} {tag} {key}")
        K, P, en = port.element(1)
        util.assert_parity(z[f"{tag}_elem1_K"], K, f"{name} {tag} element K", util.block_scale(z[f"{tag}_elem1_K"]))
        util.assert_parity(z[f"{tag}_elem1_P"], P, f"{name} {tag} element P")
        assert abs(en - z[f"{tag}_elem1_energy"][0]) <= 1e-12 * abs(en) + 1e-300
        # Gauss-point results kept for WriteResults / WriteMonitor (Shell_1.cpp:624-707, Beam_1.cpp:444-497)
        util.assert_results_parity(z[f"{tag}_results"], np.array([port.results(e) for e in range(m.n_elements)]), f"{name} {tag} results")
        if commit:
            port.commit()
            util.assert_parity(z[f"{tag}_state1"], port.state(1), f"{name} committed state")
            util.assert_parity(z[f"{tag}_copy"], port.copy_coordinates(), f"{name} copy_coordinates")


def test_shell_gravity_is_applied_twice(port):
    """Shell_1::MountFieldLoads executes the self-weight block twice
    (reference Shell_1.cpp:1340-1375); the oracle must reproduce it: at zero
    displacement P = -2 * consistent weight, and the total vertical load
    equals 2 * rho * t * g * area."""
    from giraffe_b200 import meshes as M
    m = M.shell_plate(3, 2, gravity=(0.0, 0.0, -9.81))
    m.constraints = []
    port.load(m)
    port.set_time(0.0, 1.0)
    port.assemble(np.zeros((m.n_nodes, 6)))
    pa, _, _ = port.vectors()
    gls = port.gls()
    total_z = pa[gls[:, 2][gls[:, 2] > 0] - 1].sum()
    area = 3 * 2 * 0.0195 ** 2
    expect = 2.0 * 8000.0 * 0.002 * 9.81 * area      # P = Fint - Fext, Fext = 2 * weight (downwards)
    assert abs(total_z - expect) <= 1e-10 * expect


def test_newton_steps_restatement_against_reference_fixture(port):
    """oracle/newton_steps.py (sign flip, K_AB X_B, max-norms with first-node semantics, UpdateDisps) against
    what the reference's own Static / ConvergenceCriteria / Solution code produced (make_golden.py)."""
    from oracle import newton_steps as ns
    z = _load("newton_steps")
    m = util.model_from_dict(z)
    port.load(m)
    port.set_time(0.0, 1.0)
    port.assemble(z["disp"])
    assert (port.gls() == z["gls"]).all()
    pa = port.vectors()[0]
    rhs = ns.residual(pa, port.csr("AB"), z["X_B"])
    util.assert_parity(z["rhs"], rhs, "right-hand side")
    # the steps themselves are exact: replay them on the reference's own numbers
    n = ns.residual_norms(z["gls"], z["rhs"])
    assert (n["node_force"], n["node_moment"], n["nan_detected"]) == tuple(int(v) for v in z["residual_nodes"])
    d2 = ns.update_displacements(z["gls"], z["disp"], z["x_A"])
    assert np.array_equal(d2, z["disp_after"])
    inc = ns.increment_norms(z["gls"], z["x_A"], d2)
    assert (inc["node_force"], inc["node_moment"], inc["nan_detected"]) == tuple(int(v) for v in z["increment_nodes"])


@pytest.mark.parametrize("name", ["dynamic_beam", "dynamic_shell", "dynamic_pipe"])
def test_newmark_dynamics(port, name):
    """Dynamic::Solve's element contributions (MountMass, MountDamping with a Rayleigh update, MountDyn) and
    UpdateDyn, incl. nodes with partly-free rotations, against what the reference's own Dynamic produced."""
    z = _load(name)
    m = util.model_from_dict(z)
    port.load(m)
    port.set_time(*z["time"])
    assert (port.gls() == z["gls"]).all()
    util.run_dynamic(port, m, z, util.check_dynamic(z, (1, m.n_elements - 1), name))


def test_shell_load_host_contributor(port):
    """ShellLoad follower pressure stays on the host (it is a Load): the restatement in tests/util.py added to the
    oracle's element assembly reproduces what the reference's MountLoads + MountGlobal produced."""
    import scipy.sparse as sp
    z = _load("shell_load")
    m = util.model_from_dict(z)
    assert len(m.shell_loads) == 2
    port.load(m)
    port.set_time(*z["time"])
    assert (port.gls() == z["gls"]).all()
    t = float(z["time"][0] + z["time"][1])
    for tag, commit in (("it1", False), ("it2", True), ("it3", False)):
        disp = z[f"{tag}_disp"]
        util.assert_parity(z[f"{tag}_copy_before"], port.copy_coordinates(), f"shell_load {tag} copy coordinates")
        port.assemble(disp)
        trip, pa_add, pb_add = util.shell_load_contribution(m, port.gls(), disp, port.copy_coordinates(), t)

        def with_load(w):
            o, i, v, s = port.csr(w)
            if len(trip[w][0]) == 0:
                return o, i, v, s
            # the load only adds into positions of the element pattern: look each one up in it
            base = sp.csr_matrix((np.arange(1, len(v) + 1, dtype=float), i, o), shape=s)
            out = v.copy()
            add = sp.coo_matrix((trip[w][2], (trip[w][0], trip[w][1])), shape=s).tocsr()
            add.sum_duplicates()
            rows = np.repeat(np.arange(s[0]), np.diff(add.indptr))
            pos = np.asarray(base[rows, add.indices]).reshape(-1).astype(np.int64) - 1
            assert (pos >= 0).all(), "ShellLoad position outside the element pattern"
            out[pos] += add.data
            return o, i, out, s
        util.assert_system_parity(lambda w: util.captured_csr(z, tag, w), with_load, f"shell_load {tag}")
        pa, ia, pb = port.vectors()
        np.add.at(pa, pa_add[0], pa_add[1]); np.add.at(ia, pa_add[0], pa_add[1]); np.add.at(pb, pb_add[0], pb_add[1])
        util.assert_parity(z[f"{tag}_PA"], pa, f"shell_load {tag} P_A")
        util.assert_parity(z[f"{tag}_IA"], ia, f"shell_load {tag} I_A")
        util.assert_parity(z[f"{tag}_PB"], pb, f"shell_load {tag} P_B")
        if commit:
            port.commit()


def _run_pipe_fixture(z, m, b, what, commits, apply_loads=None, add_triplets=None, add_vector=None):
    """Shared by the port (CPU) and the library (GPU): iterations of a fixture whose loads (PipeLoad, and NodalLoad when
    the model has one) were mounted by the reference."""
    t = float(z["time"][0] + z["time"][1])
    for tag, commit in commits:
        disp = z[f"{tag}_disp"]
        pa_add = pb_add = None
        host_loads = bool(m.nodal_loads or getattr(m, "follower_loads", []))
        if host_loads:
            copy = z[f"{tag}_copy_before"] if f"{tag}_copy_before" in z.files else np.zeros((m.n_nodes, 6))
            trip, pa_add, pb_add = util.host_load_contribution(m, z["gls"], disp, copy, t)
            if add_triplets is None:                      # the port takes host triplets before its assembly
                for w in ("AA", "AB", "BA", "BB"):
                    b.set_extra_triplets(w, *trip[w])
        b.assemble(disp)
        if apply_loads is not None:
            apply_loads(t)
        if host_loads and add_triplets is not None:
            for w in ("AA", "AB", "BA", "BB"):
                if trip[w][0]:
                    add_triplets(w, *trip[w])
            add_vector("PA", *pa_add)
            if pb_add[0]:
                add_vector("PB", *pb_add)
        util.assert_system_parity(lambda w: util.captured_csr(z, tag, w), b.csr, f"{what} {tag}")
        pa, ia, pb = [v.copy() for v in b.vectors()]
        if host_loads and add_triplets is None:
            np.add.at(pa, pa_add[0], pa_add[1]); np.add.at(pb, pb_add[0], pb_add[1])
        util.assert_parity(z[f"{tag}_PA"], pa, f"{what} {tag} P_A")
        util.assert_parity(z[f"{tag}_IA"], ia, f"{what} {tag} I_A")
        util.assert_parity(z[f"{tag}_PB"], pb, f"{what} {tag} P_B")
        if commit:
            b.commit()


def test_pipe_load_internal_pressure(port):
    """PipeLoad -> Pipe_1::MountPipeSpecialLoads (Pipe_1.cpp:1443-1494) as restated in the oracle port, against what the
    reference's own MountLoads + MountGlobal produced on a bent pipe line (two loads, overlapping element sets,
    before and after a commit).  The pressure matters: the fixture differs from the same assembly without it."""
    z = _load("pipe_load")
    m = util.model_from_dict(z)
    assert len(m.pipe_loads) == 2
    port.load(m)
    port.set_time(*z["time"])
    assert (port.gls() == z["gls"]).all()
    _run_pipe_fixture(z, m, port, "pipe_load", (("it1", False), ("it2", True), ("it3", False)))
    # sensitivity: without the loads the same assembly is far outside the tolerance
    saved, m.pipe_loads = m.pipe_loads, []
    port.load(m)
    port.set_time(*z["time"])
    port.assemble(z["it1_disp"])
    m.pipe_loads = saved
    assert np.abs(port.vectors()[0] - z["it1_PA"]).max() > 1e-4 * np.abs(z["it1_PA"]).max()
    assert np.abs(port.csr("AA")[2] - z["it1_AA_val"]).max() > 1e-6 * np.abs(z["it1_AA_val"]).max()


def test_shipped_tutorial04_with_its_pipe_load(port):
    """inputs/tutorial04 as shipped (50 Pipe_1, NodalLoad perturbation + PipeLoad internal pressure), inside the
    pressure ramp of its second solution step."""
    z = _load("tutorial04")
    m = util.model_from_dict(z)
    assert len(m.pipe_loads) == 1 and m.n_elements == 50
    port.load(m)
    port.set_time(*z["time"])
    assert (port.gls() == z["gls"]).all()
    _run_pipe_fixture(z, m, port, "tutorial04", (("it1", True), ("it2", False)))


def test_shipped_tutorial03_with_its_follower_load(port):
    """inputs/tutorial03 as shipped (50 Pipe_1, NodalLoad + NodalFollowerLoad): the follower load is a host Load, restated
    in tests/util.py and pushed as host triplets; the second iteration sees committed rotations."""
    z = _load("tutorial03")
    m = util.model_from_dict(z)
    assert len(m.follower_loads) == 1 and m.n_elements == 50
    port.load(m)
    port.set_time(*z["time"])
    assert (port.gls() == z["gls"]).all()
    _run_pipe_fixture(z, m, port, "tutorial03", (("it1", True), ("it2", False)))


@pytest.mark.parametrize("name", ["tutorial05", "tutorial02"])
def test_shipped_shell_meshes(port, name):
    """inputs/tutorial05 (400 Shell_1) and inputs/tutorial02 (3036 Shell_1) as the reference ships them, assembled by
    the reference's own sources (fixture) and by the oracle port: pattern, values, vectors, a commit in between."""
    z = np.load(os.path.join(util.GOLDEN_DIR, name + "_shells.npz"))
    m = util.model_from_dict(z)
    port.load(m)
    assert (port.gls() == z["gls"]).all()
    port.set_time(*z["time"])
    util.check_shipped_shell_mesh(z, port, f"{name} (port)")
