"""CPU, build container only: the port restatement against the reference's own
sources (oracle/_ref) on seeded models that are larger / different from the
committed fixtures.  Skipped where /root/reference (hence oracle/_ref) is absent."""
import numpy as np
import pytest

import util
from giraffe_b200 import meshes as M


def _cases():
    b = M.beam_line(40, pretension=5.0e4)
    b.gravity = (0.1, 0.2, -9.81)
    s = M.shell_plate(5, 7, warp=0.02, gravity=(0.0, 0.0, -9.81))
    flat = M.shell_plate(4, 3)
    mixed = M.concat_models([M.beam_line(10), M.shell_plate(3, 3, warp=0.005)])
    return [("beam", b, M.beam_line_displacements(b)), ("shell", s, M.shell_plate_displacements(s)),
            ("flat", flat, M.shell_plate_displacements(flat, seed=7)),
            ("mixed", mixed, M.mask_displacements(mixed, np.random.default_rng(3).uniform(-1e-3, 1e-3, (mixed.n_nodes, 6))))]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_port_matches_reference_sources(ref, port, case):
    name, m, d = case
    ref.load(m)
    port.load(m)
    ref.set_time(0.0, 0.5)
    port.set_time(0.0, 0.5)
    assert (ref.gls() == port.gls()).all()
    assert (ref.n_free, ref.n_fixed) == (port.n_free, port.n_fixed)
    for it in range(3):
        ref.assemble(d)
        port.assemble(d)
        for w in ("AA", "AB", "BA", "BB"):
            assert ref.triplets(w) == port.triplets(w)
        util.assert_system_parity(ref.csr, port.csr, f"{name} it{it}")
        for a, b, key in zip(ref.vectors(), port.vectors(), ("PA", "IA", "PB")):
            util.assert_parity(a, b, f"{name} it{it} {key}")
        for e in (0, m.n_elements // 2, m.n_elements - 1):
            Kr, Pr, er = ref.element(e)
            Kp, Pp, ep = port.element(e)
            util.assert_parity(Kr, Kp, f"{name} element {e} K", util.block_scale(Kr))
            assert abs(er - ep) <= 1e-12 * abs(er) + 1e-300
        ref.commit()
        port.commit()
        for e in (0, m.n_elements - 1):
            util.assert_parity(ref.state(e), port.state(e), f"{name} state of element {e}")
        util.assert_parity(ref.copy_coordinates(), port.copy_coordinates(), f"{name} copy coordinates")
        d = 0.5 * d


def test_section_and_cs_constants(ref):
    """meshes.py restates SecRectangle/SecTube::PreCalc and CoordinateSystem::Read."""
    m = M.beam_line(2)
    m.section_defs = [(1, 0.65, 0.62), (0, 0.1, 0.25)]
    m.cs_defs = [((1.0, 0.0, 0.0), (0.0, 0.0, 1.0)), ((2.0, 0.0, 0.0), (0.0, 0.0, 0.5))]
    M._finish(m)
    ref.load(m)
    for k in (1, 2):
        assert np.array_equal(ref.section(k), m.sections[k - 1])
        assert np.array_equal(ref.cs(k), m.cs[k - 1])


@pytest.mark.parametrize("rayleigh", [(0.0, 0.0), (0.35, 2.0e-4)], ids=["undamped", "rayleigh"])
def test_port_newmark_dynamics_match_reference_sources(ref, port, rayleigh):
    """Seeded models other than the fixtures: the port's formulation-level restatement of the AceGen inertia
    code (forward-mode tangent) against the reference's own MountMass / MountDamping / MountDyn / UpdateDyn."""
    b = M.beam_line(30, pretension=5.0e4)
    b.gravity = (0.1, 0.2, -9.81)
    b.constraints = b.constraints + [([21], 0x20), ([22], 0x18)]
    s = M.shell_plate(4, 5, warp=0.02, gravity=(0.0, 0.0, -9.81))
    mixed = M.concat_models([M.beam_line(8), M.shell_plate(3, 3, warp=0.005)])
    cases = [("beam", b, M.beam_line_displacements(b)), ("shell", s, M.shell_plate_displacements(s)),
             ("mixed", mixed, np.random.default_rng(3).uniform(-1e-3, 1e-3, (mixed.n_nodes, 6)))]
    for k, (name, m, d) in enumerate(cases):
        scen = util.dynamic_scenario(m, M.mask_displacements(m, d), 77 + k, time_step=0.004, rayleigh=rayleigh)
        ref.load(m)
        ref.set_time(0.0, 0.5)
        z = dict(scen)
        els = (0, m.n_elements - 1)
        util.run_dynamic(ref, m, scen, util.capture_dynamic(z, els))
        port.load(m)
        port.set_time(0.0, 0.5)
        util.run_dynamic(port, m, scen, util.check_dynamic(z, els, name))


def _large_rotation_dynamic_case():
    """Rotation increments of ~0.5 rad over three committed time steps: the committed Rodrigues vector alpha_i,
    the gyroscopic terms and the forward-mode tangent far from the linear regime."""
    m = M.concat_models([M.beam_line(6), M.shell_plate(3, 2, warp=0.02)])
    rng = np.random.default_rng(4242)
    d = rng.uniform(-1.0, 1.0, (m.n_nodes, 6)) * np.array([2e-3, 2e-3, 2e-3, 0.5, 0.5, 0.5])
    return m, M.mask_displacements(m, d)


def test_port_newmark_dynamics_large_rotations(ref, port):
    m, d = _large_rotation_dynamic_case()
    scen = util.dynamic_scenario(m, d, 99, time_step=0.02, rayleigh=(0.2, 5.0e-5))
    scen["dyn_copy_vel"] = scen["dyn_copy_vel"] * np.array([1, 1, 1, 6.0, 6.0, 6.0])      # angular velocities of ~3 rad/s
    els = (0, 5, 6, m.n_elements - 1)
    z = dict(scen)
    ref.load(m); ref.set_time(0.0, 0.5)
    port.load(m); port.set_time(0.0, 0.5)
    # three time steps, each with a commit: run the 3-assembly scenario and commit after every assembly
    steps = (("s1", 0, True, True), ("s2", 1, False, True), ("s3", 2, False, True))
    old = util.DYN_STEPS
    util.DYN_STEPS = steps
    try:
        util.run_dynamic(ref, m, scen, util.capture_dynamic(z, els))
        util.run_dynamic(port, m, scen, util.check_dynamic(z, els, "large rotations"))
    finally:
        util.DYN_STEPS = old


def test_port_matches_reference_on_random_models(ref, port):
    """Seeded random variations the fixed cases do not reach: random constraint masks on random nodes (partly
    constrained translations and rotations, fully fixed and free-floating nodes), random plate / line sizes and
    warps, gravity on or off, displacement amplitudes from 1e-6 to 1e-2, with a commit between iterations."""
    rng = np.random.default_rng(20240031)
    for trial in range(12):
        nb = int(rng.integers(2, 9)); nx = int(rng.integers(1, 5)); ny = int(rng.integers(1, 4))
        parts = []
        if trial % 3 != 1:
            parts.append(M.beam_line(nb, pretension=float(rng.choice([0.0, 3.0e4]))))
        if trial % 3 != 2:
            parts.append(M.shell_plate(nx, ny, warp=float(rng.choice([0.0, 0.01, 0.05]))))
        if trial % 4 == 0:
            parts.append(M.pipe_line(int(rng.integers(2, 6))))
        m = M.concat_models(parts)
        m.gravity = None if trial % 2 else (float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), -9.81)
        extra = []
        for _ in range(int(rng.integers(1, 5))):
            node = int(rng.integers(1, m.n_nodes + 1))
            extra.append(([node], int(rng.integers(1, 64))))
        m.constraints = m.constraints + extra
        amp = 10.0 ** rng.uniform(-6, -2)
        d = M.mask_displacements(m, rng.uniform(-amp, amp, (m.n_nodes, 6)))
        ref.load(m); port.load(m)
        ref.set_time(0.0, 0.3); port.set_time(0.0, 0.3)
        assert (ref.gls() == port.gls()).all(), f"trial {trial}: DOF numbering"
        for it in range(2):
            ref.assemble(d); port.assemble(d)
            util.assert_system_parity(ref.csr, port.csr, f"random trial {trial} it{it}")
            for a, b, key in zip(ref.vectors(), port.vectors(), ("PA", "IA", "PB")):
                util.assert_parity(a, b, f"random trial {trial} it{it} {key}")
            ref.commit(); port.commit()
            d = -0.7 * d


def test_port_pipe_pressure_matches_reference_sources(ref, port):
    """PipeLoad internal pressure (Pipe_1::MountPipeSpecialLoads, Pipe_1.cpp:1443-1494) on seeded pipe lines bent by
    large displacements, random element sets and pressures of both signs, with a commit between iterations."""
    rng = np.random.default_rng(20240051)
    for trial in range(4):
        n = int(rng.integers(3, 15))
        m = M.pipe_line(n)
        m.gravity = (0.0, 0.3, -9.81) if trial % 2 else None
        loads = []
        for _ in range(int(rng.integers(1, 4))):
            els = np.sort(rng.choice(np.arange(1, n + 1), size=int(rng.integers(1, n + 1)), replace=False)).astype(np.int32)
            p = float(rng.uniform(-4.0e8, 4.0e8))
            loads.append((els, np.array([[0.0, 0.1 * p, 0, 0, 0], [1.0, p, 5.0e4, 800.0, 1025.0]])))
        m.pipe_loads = loads
        ref.load(m)
        port.load(m)
        ref.set_time(0.2, 0.5)
        port.set_time(0.2, 0.5)
        d = M.mask_displacements(m, rng.uniform(-5e-2, 5e-2, (m.n_nodes, 6)))
        for it in range(3):
            ref.assemble(d, with_loads=True)
            port.assemble(d)
            util.assert_system_parity(ref.csr, port.csr, f"pipe pressure trial {trial} it{it}")
            for a, b, key in zip(ref.vectors(), port.vectors(), ("PA", "IA", "PB")):
                util.assert_parity(a, b, f"pipe pressure trial {trial} it{it} {key}")
            if it == 0:
                ref.commit()
                port.commit()
            d = -0.6 * d
