"""GPU: the CUDA path, called through the C-ABI, against the CPU oracle
(oracle/port, pinned to the reference by tests/golden and test_oracle_vs_ref)
on the same seeded inputs, against the committed golden fixtures, and -- at
BASELINE.json's full sizes -- through size-independent properties."""
import os

import numpy as np
import pytest

import util
from giraffe_b200 import capi, meshes as M

pytestmark = pytest.mark.gpu


def _golden(name):
    return np.load(os.path.join(util.GOLDEN_DIR, name + ".npz"))


def _compare_system(port, asm, what):
    worst = util.assert_system_parity(port.csr, asm.csr, what)
    for a, b, key in zip(port.vectors(), asm.vectors(), ("P_A", "I_A", "P_B")):
        util.assert_parity(a, b, f"{what} {key}")
    return worst


# ---- committed reference fixtures ------------------------------------------
def test_tutorial01_against_reference_fixture():
    """BASELINE.json configs[0]: inputs/tutorial01 as shipped, Newton iterations
    1 and 2 of increment 1, including the host NodalLoad contribution."""
    z = _golden("tutorial01")
    m = util.model_from_dict(z)
    asm = capi.Assembler(m)
    gls, nf, nx = asm.number_dofs()
    assert (gls == z["gls"]).all() and (nf, nx) == (60, 6)
    asm.set_dofs(gls, nf, nx)
    assert asm.csr_dims("AA")[2] == 1296
    t0, dt = z["time"]
    for tag in ("it1", "it2"):
        disp = z[f"{tag}_disp"]
        asm.assemble(disp)
        trip, pa_add, pb_add = util.nodal_load_contribution(m, gls, disp, t0 + dt)
        for w in ("AA", "AB", "BA", "BB"):
            if trip[w][0]:
                asm.add_host_triplets(w, *trip[w])
        asm.add_host_vector(capi.P_A, *pa_add)
        if pb_add[0]:
            asm.add_host_vector(capi.P_B, *pb_add)
        util.assert_system_parity(lambda w: util.captured_csr(z, tag, w), asm.csr, f"tutorial01 {tag}")
        pa, ia, pb = asm.vectors()
        util.assert_parity(z[f"{tag}_PA"], pa, f"tutorial01 {tag} P_A")
        util.assert_parity(z[f"{tag}_IA"], ia, f"tutorial01 {tag} I_A")
        util.assert_parity(z[f"{tag}_PB"], pb, f"tutorial01 {tag} P_B")
    K, P = asm.element(0)
    # fixture element block is from iteration 2
    util.assert_parity(z["elem0_K"], K, "tutorial01 element 1 K", util.block_scale(z["elem0_K"]))


def test_tutorial01_static_solution_in_lockstep_with_the_reference(ref):
    """inputs/tutorial01 solved to the end -- ten increments of Static::Solve's Newton loop (Static.cpp:161-236) --
    twice: through the reference's own sources (assembly, MountLoads, sign flip, UpdateDisps, SaveConfiguration)
    and through the library (gfa_assemble, the host NodalLoad pushed with gfa_add_host_*, gfa_residual,
    gfa_update_displacements on the device copy, gfa_commit_state), the same sparse solve in between.  The
    configurations must stay together increment by increment."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    z = _golden("tutorial01")
    m = util.model_from_dict(z)
    dt = float(z["time"][1])
    ref.load(m)
    asm = capi.Assembler(m)
    gls, nf, nx = asm.number_dofs()
    asm.set_dofs(gls, nf, nx)
    t = 0.0
    for inc in range(10):
        ref.set_time(t, dt)
        asm.set_time(t, dt)
        d_ref = np.zeros((m.n_nodes, 6))
        asm.assemble(np.zeros((m.n_nodes, 6)))              # start of the increment: zero increments on the device
        first = True
        for it in range(8):
            # the reference's own loop
            ref.assemble(d_ref, with_loads=True)
            ref.residual(None)
            o, i, v, shape = ref.csr("AA")
            x = spla.spsolve(sp.csr_matrix((v, i, o), shape=shape).tocsc(), ref.vectors()[0])
            d_ref = ref.update_displacements(x)[0]
            # the library's
            if not first:
                asm.assemble(None)                              # the device copy left by gfa_update_displacements
            first = False
            d_dev = asm.displacements()
            trip, pa_add, pb_add = util.nodal_load_contribution(m, gls, d_dev, t + dt)
            for w in ("AA", "AB", "BA", "BB"):
                if trip[w][0]:
                    asm.add_host_triplets(w, *trip[w])
            asm.add_host_vector(capi.P_A, *pa_add)
            if pb_add[0]:
                asm.add_host_vector(capi.P_B, *pb_add)
            norms = asm.residual(None)
            o2, i2, v2, shape2 = asm.csr("AA")
            x2 = spla.spsolve(sp.csr_matrix((v2, i2, o2), shape=shape2).tocsc(), asm.vectors()[0])
            inc_norms = asm.update_displacements(x2)
            assert not norms["nan_detected"] and not inc_norms["nan_detected"]
        assert np.abs(x2).max() <= 1e-9 * max(np.abs(asm.displacements()).max(), 1e-30), "Newton did not converge in 8 iterations"
        util.assert_parity(d_ref, asm.displacements(), f"tutorial01 increment {inc + 1}: converged increments", tol=1e-9)
        ref.commit()
        asm.commit()
        util.assert_parity(ref.copy_coordinates(), asm.copy_coordinates(), f"tutorial01 increment {inc + 1}: configuration", tol=1e-10)
        t += dt
    tip = asm.copy_coordinates()[-1, :3] - m.xyz[-1]
    assert np.abs(tip).max() > 1e-6, "the load did not move the structure"


@pytest.mark.parametrize("name", ["beam_line", "pipe_line", "shell_plate"])
def test_sequence_against_reference_fixture(name):
    z = _golden(name)
    m = util.model_from_dict(z)
    asm = capi.Assembler(m).set_dofs()
    asm.set_time(*z["time"])
    assert (asm.gls == z["gls"]).all()
    for tag, commit in (("it1", False), ("it2", True), ("it3", False)):
        asm.assemble(z[f"{tag}_disp"])
        util.assert_system_parity(lambda w: util.captured_csr(z, tag, w), asm.csr, f"{name} {tag}")
        for v, key in zip(asm.vectors(), ("PA", "IA", "PB")):
            util.assert_parity(z[f"{tag}_{key}"], v, f"{name} {tag} {key}")
        K, P = asm.element(1)
        util.assert_parity(z[f"{tag}_elem1_K"], K, f"{name} {tag} element K", util.block_scale(z[f"{tag}_elem1_K"]))
        util.assert_parity(z[f"{tag}_elem1_P"], P, f"{name} {tag} element P")
        # result read-back: strains, stress resultants and strain energy of every element
        etype = int(m.elem_type[0])
        util.assert_results_parity(z[f"{tag}_results"], asm.gauss_point_results(etype), f"{name} {tag} results")
        if commit:
            asm.commit()
            util.assert_parity(z[f"{tag}_state1"], asm.state(1), f"{name} committed state")
            util.assert_parity(z[f"{tag}_copy"], asm.copy_coordinates(), f"{name} copy_coordinates")


@pytest.mark.parametrize("name", ["tutorial05", "tutorial02"])
def test_shipped_shell_meshes_against_reference_fixture(name):
    """The shell meshes the reference ships (inputs/tutorial05: 400 Shell_1; inputs/tutorial02: 3036 Shell_1),
    assembled by the reference's own sources (fixture) and by the CUDA path: two iterations with a commit."""
    z = _golden(name + "_shells")
    m = util.model_from_dict(z)
    asm = capi.Assembler(m).set_dofs()
    assert (asm.gls == z["gls"]).all()
    asm.set_time(*z["time"])
    util.check_shipped_shell_mesh(z, asm, name)
    K, P = asm.element(m.n_elements // 2)
    util.assert_parity(z["it2_elem_K"], K, f"{name} element K", util.block_scale(z["it2_elem_K"]))
    util.assert_parity(z["it2_elem_P"], P, f"{name} element P")


# ---- seeded models against the oracle ----------------------------------------
def _seeded_cases():
    b = M.beam_line(300, pretension=5.0e4)
    b.gravity = (0.1, 0.2, -9.81)
    s = M.shell_plate(17, 9, warp=0.02, gravity=(0.0, 0.0, -9.81))
    flat = M.shell_plate(8, 8)
    v = M.solid_block(5, 4, 3, gravity=(0.0, 0.0, -9.81))
    mixed = M.concat_models([M.beam_line(37), M.pipe_line(21), M.shell_plate(7, 5, warp=0.005), M.solid_block(3, 3, 2)])
    pipe = M.pipe_line(300, gravity=(0.1, 0.2, -9.81))
    rng = np.random.default_rng(11)
    return [
        ("beam", b, M.beam_line_displacements(b)),
        ("pipe", pipe, M.beam_line_displacements(pipe)),
        ("shell_warped_gravity", s, M.shell_plate_displacements(s)),
        ("shell_flat", flat, M.shell_plate_displacements(flat, seed=5)),
        ("solid", v, M.solid_block_displacements(v)),
        ("mixed", mixed, M.mask_displacements(mixed, rng.uniform(-1e-4, 1e-4, (mixed.n_nodes, 6)))),
    ]


@pytest.mark.parametrize("case", _seeded_cases(), ids=lambda c: c[0])
def test_against_oracle_with_commits(port, case):
    """Three iterations with a state commit after each (updated Lagrangian)."""
    name, m, d = case
    port.load(m)
    asm = capi.Assembler(m).set_dofs()
    assert (asm.gls == port.gls()).all() and (asm.n_free, asm.n_fixed) == (port.n_free, port.n_fixed)
    port.set_time(0.0, 0.5)
    asm.set_time(0.0, 0.5)
    for it in range(3):
        port.assemble(d)
        asm.assemble(d)
        _compare_system(port, asm, f"{name} it{it}")
        for e in sorted({0, m.n_elements // 3, m.n_elements // 2, m.n_elements - 1}):
            Kp, Pp, _ = port.element(e)
            Kg, Pg = asm.element(e)
            util.assert_parity(Kp, Kg, f"{name} element {e} K", util.block_scale(Kp))
            util.assert_parity(Pp, Pg, f"{name} element {e} P")
        port.commit()
        asm.commit()
        for e in (0, m.n_elements - 1):
            util.assert_parity(port.state(e), asm.state(e), f"{name} state of element {e}")
        util.assert_parity(port.copy_coordinates(), asm.copy_coordinates(), f"{name} copy coordinates")
        d = -0.6 * d


def test_large_rotation_increment(port):
    """Rotation increments of order 1 rad exercise every geometric term."""
    m = M.shell_plate(4, 4, warp=0.03)
    rng = np.random.default_rng(5)
    d = M.mask_displacements(m, np.concatenate([rng.uniform(-2e-3, 2e-3, (m.n_nodes, 3)), rng.uniform(-0.8, 0.8, (m.n_nodes, 3))], axis=1))
    port.load(m)
    asm = capi.Assembler(m).set_dofs()
    port.assemble(d)
    asm.assemble(d)
    _compare_system(port, asm, "large rotation shell")
    b = M.beam_line(20)
    db = M.mask_displacements(b, np.concatenate([rng.uniform(-1e-2, 1e-2, (b.n_nodes, 3)), rng.uniform(-0.9, 0.9, (b.n_nodes, 3))], axis=1))
    port.load(b)
    asmb = capi.Assembler(b).set_dofs()
    port.assemble(db)
    asmb.assemble(db)
    _compare_system(port, asmb, "large rotation beam")


def test_ragged_batches_and_tiny_models(port):
    """Element counts that do not fill a warp batch (10 shells / 16 beams / 4 solids)."""
    for m in (M.shell_plate(1, 1), M.beam_line(1), M.solid_block(1, 1, 1), M.beam_line(17), M.shell_plate(3, 2), M.solid_block(3, 1, 1)):
        d = M.mask_displacements(m, np.random.default_rng(m.n_elements).uniform(-1e-4, 1e-4, (m.n_nodes, 6)))
        port.load(m)
        asm = capi.Assembler(m).set_dofs()
        port.assemble(d)
        asm.assemble(d)
        _compare_system(port, asm, f"tiny model with {m.n_elements} elements")


def test_unconstrained_and_fully_fixed_nodes(port):
    """No fixed DOF at all (empty AB/BA/BB) and a model whose every DOF of some
    elements is fixed (rows that exist only in BB)."""
    m = M.shell_plate(4, 3)
    m.constraints = []
    d = M.mask_displacements(m, np.random.default_rng(1).uniform(-1e-4, 1e-4, (m.n_nodes, 6)))
    port.load(m)
    asm = capi.Assembler(m).set_dofs()
    assert asm.n_fixed == 0 and asm.csr_dims("BB")[2] == 0
    port.assemble(d)
    asm.assemble(d)
    _compare_system(port, asm, "unconstrained plate")
    m2 = M.beam_line(6)
    m2.constraints = [(np.arange(1, 6, dtype=np.int32), 0x3F), (np.array([9], np.int32), 0x15)]
    d2 = M.mask_displacements(m2, np.random.default_rng(2).uniform(-1e-3, 1e-3, (m2.n_nodes, 6)))
    port.load(m2)
    asm2 = capi.Assembler(m2).set_dofs()
    port.assemble(d2)
    asm2.assemble(d2)
    _compare_system(port, asm2, "partially and fully fixed beam nodes")


def test_assembly_is_bitwise_reproducible_and_reentrant():
    """Atomic-free scatter: repeated assemblies are bit-identical, and assembling
    again after a 'diverged' trial (no commit) reproduces the earlier result
    (failure recovery: RestoreConfiguration + bisection)."""
    m = M.shell_plate(20, 10, warp=0.01)
    d = M.shell_plate_displacements(m)
    asm = capi.Assembler(m).set_dofs()
    asm.assemble(d)
    v1 = asm.values("AA").copy()
    p1 = asm.vectors()[0].copy()
    asm.assemble(7.0 * d)          # a trial that is thrown away
    asm.assemble(d)
    assert asm.values("AA").tobytes() == v1.tobytes()
    assert asm.vectors()[0].tobytes() == p1.tobytes()
    # gfa_assemble_enqueue: same work, not waited for; the reads wait for the stream
    asm.assemble(7.0 * d)
    asm.assemble(d)                # leaves d as the device copy of the displacements
    asm.assemble_enqueue(None)
    assert asm.values("AA").tobytes() == v1.tobytes()
    assert asm.vectors()[0].tobytes() == p1.tobytes()
    assert asm.timing()["total_ms"] > 0.0
    import ctypes
    st = capi._StepStruct()
    st.displacements, st.displacements_on_device = d.ctypes.data, 0
    assert asm.lib.gfa_assemble_enqueue(asm._h, ctypes.byref(st)) == -1      # host pointers are refused


def test_host_triplets_outside_pattern_are_rejected():
    m = M.beam_line(8)
    asm = capi.Assembler(m).set_dofs()
    asm.assemble(np.zeros((m.n_nodes, 6)))
    asm.add_host_triplets("AA", [0, 0], [0, 0], [1.0, 2.0])          # duplicates are summed
    with pytest.raises(capi.GfaError) as ei:
        asm.add_host_triplets("AA", [0], [asm.n_free - 1], [1.0])
    assert ei.value.code == -5


def test_host_positions_outside_the_element_pattern(port):
    """contributors that stay on the host (joints, contacts, loads) push triplets into the same lists
    as the elements (Solution.cpp:268-281, 322-349) -- also at positions no element touches.  gfa_set_dofs takes
    them into the pattern (byte-equal to what setFromTriplets yields for elements + extras), gfa_add_host_triplets
    sums into them, and every assembly starts those slots from zero again."""
    m = M.concat_models([M.beam_line(12), M.shell_plate(5, 4, warp=0.01)])
    gls, nf, nx = M.number_dofs(m)
    rng = np.random.default_rng(31)
    d = M.mask_displacements(m, rng.uniform(-1e-4, 1e-4, (m.n_nodes, 6)))
    # a "joint" between a beam node and two shell nodes that share no element with it: every free-free pair, both ways
    a = 9                                                    # beam node (1-based)
    shell_nodes = np.unique(m.elem_nodes[m.elem_ptr[12 + 7]:m.elem_ptr[12 + 8]])
    b, c = int(shell_nodes[0]), int(shell_nodes[-1])
    rows, cols = [], []
    for n1, n2 in ((a, b), (b, a), (a, c), (c, a), (b, c)):     # (b, c) lies inside the shell pattern already
        for g1 in gls[n1 - 1][gls[n1 - 1] > 0]:
            for g2 in gls[n2 - 1][gls[n2 - 1] > 0]:
                rows.append(int(g1) - 1); cols.append(int(g2) - 1)
    rows.append(3); cols.append(nf - 1)                       # and one lone position far off the diagonal
    rows, cols = np.array(rows, np.int32), np.array(cols, np.int32)
    vals = rng.uniform(-1e6, 1e6, len(rows))
    port.load(m)
    port.set_time(0.0, 1.0)
    port.set_extra_triplets("AA", rows, cols, vals)
    asm = capi.Assembler(m).set_dofs(gls, nf, nx, extra=(np.zeros(len(rows), np.int32), rows, cols))
    asm.set_time(0.0, 1.0)
    for it in range(2):                                      # the second pass: slots are rewritten, not accumulated
        port.assemble(d)
        asm.assemble(d)
        asm.add_host_triplets("AA", rows, cols, vals)
        _compare_system(port, asm, f"joint positions outside the element pattern, pass {it}")
    port.set_extra_triplets("AA", np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
    # DOFs beyond the node table (Lagrange multipliers of a joint, Solution.cpp:73-107): rows that exist through host
    # positions only, and vector entries no element writes -- zeroed by every assembly like Solution::Clear does
    n_lag = 3
    lr, lc, lv = [], [], []
    for k in range(n_lag):
        g = gls[a - 1][gls[a - 1] > 0][k] - 1
        lr += [nf + k, int(g), nf + k]; lc += [int(g), nf + k, nf + k]; lv += [1.0 + k, 1.0 + k, 0.5]
    lr, lc, lv = np.array(lr, np.int32), np.array(lc, np.int32), np.array(lv)
    asm2 = capi.Assembler(m).set_dofs(gls, nf + n_lag, nx, extra=(np.zeros(len(lr), np.int32), lr, lc))
    base = capi.Assembler(m).set_dofs(gls, nf, nx)
    base.assemble(d)
    bo, bi, bv, _ = base.csr("AA")
    for it in range(2):
        asm2.assemble(d)
        asm2.add_host_triplets("AA", lr, lc, lv)
        asm2.add_host_vector(capi.P_A, np.array([nf, nf + 2], np.int32), np.array([7.0, -3.0]))
        o, i, v, shape = asm2.csr("AA")
        assert shape == (nf + n_lag, nf + n_lag)
        for k in range(n_lag):
            g = int(gls[a - 1][gls[a - 1] > 0][k] - 1)
            r0, r1 = o[nf + k], o[nf + k + 1]
            assert list(i[r0:r1]) == [g, nf + k] and list(v[r0:r1]) == [1.0 + k, 0.5]
            row = dict(zip(i[o[g]:o[g + 1]].tolist(), v[o[g]:o[g + 1]].tolist()))
            assert row[nf + k] == 1.0 + k
            ref_row = dict(zip(bi[bo[g]:bo[g + 1]].tolist(), bv[bo[g]:bo[g + 1]].tolist()))
            assert {c: x for c, x in row.items() if c < nf} == ref_row        # the element part of the row is untouched
        pa = asm2.vectors()[0]
        assert pa[nf] == 7.0 and pa[nf + 1] == 0.0 and pa[nf + 2] == -3.0
        np.testing.assert_array_equal(pa[:nf], base.vectors()[0])


def test_set_dofs_refuses_numberings_it_cannot_assemble():
    """The pattern builder relies on the reference's node-major ascending numbering (Solution.cpp:53-72); a permuted
    or interleaved map is refused instead of being mis-assembled."""
    m = M.beam_line(6)
    gls, nf, nx = M.number_dofs(m)
    asm = capi.Assembler(m)
    g2 = gls.copy()
    i, j = np.argwhere(g2 == 7)[0], np.argwhere(g2 == 20)[0]
    g2[tuple(i)], g2[tuple(j)] = 20, 7
    with pytest.raises(capi.GfaError) as ei:
        asm.set_dofs(g2, nf, nx)
    assert ei.value.code == -7
    g3 = gls.copy()
    g3[g3 > 8] += 1                                          # a foreign DOF numbered inside a node's translation group
    with pytest.raises(capi.GfaError) as ei:
        asm.set_dofs(g3, nf + 1, nx)
    assert ei.value.code == -7
    asm.set_dofs(gls, nf, nx)                                # the reference's own numbering still goes through


def test_call_order_errors():
    m = M.beam_line(3)
    asm = capi.Assembler(m)
    with pytest.raises(capi.GfaError) as ei:
        asm.assemble(np.zeros((m.n_nodes, 6)))
    assert ei.value.code == -4


# ---- full BASELINE sizes: size-independent properties -------------------------
def _sample_submodel(m, elems):
    """Sub-model made of the sampled elements only (same node coordinates)."""
    ptr = m.elem_ptr
    nodes = np.unique(np.concatenate([m.elem_nodes[ptr[e]:ptr[e + 1]] for e in elems]))
    remap = {int(n): i + 1 for i, n in enumerate(nodes)}
    sub = M.Model(xyz=m.xyz[nodes - 1], hooke=m.hooke, sections=m.sections)
    sub.section_defs, sub.shell_thickness, sub.cs_defs = m.section_defs, m.shell_thickness, m.cs_defs
    sub.elem_type, sub.elem_mat = m.elem_type[elems], m.elem_mat[elems]
    sub.elem_sec, sub.elem_cs = m.elem_sec[elems], m.elem_cs[elems]
    sub.elem_nodes = np.array([remap[int(n)] for e in elems for n in m.elem_nodes[ptr[e]:ptr[e + 1]]], np.int32)
    sub.gravity = m.gravity
    return M._finish(sub), nodes


def _sampled_elements_against_port(port, m, d, asm, what, n=64):
    """sampled elements of a full-size model against the oracle evaluated on a sub-model made of them"""
    rng = np.random.default_rng(123)
    elems = np.sort(rng.choice(m.n_elements, size=n, replace=False))
    sub, nodes = _sample_submodel(m, elems)
    port.load(sub)
    port.set_time(0.0, 1.0)
    port.assemble(d[nodes - 1])
    for k, e in enumerate(elems):
        Kp, Pp, _ = port.element(k)
        Kg, Pg = asm.element(int(e))
        util.assert_parity(Kp, Kg, f"{what}: element {e} K", util.block_scale(Kp))
        util.assert_parity(Pp, Pg, f"{what}: element {e} P")


def _full_size_checks(port, m, d, n_dof_el, what):
    asm = capi.Assembler(m).set_dofs()
    asm.assemble(d)
    # (1) sampled elements against the oracle evaluated on a sub-model
    _sampled_elements_against_port(port, m, d, asm, what)
    # (2) checksum of checksums: sum of all CSR values == sum of all element blocks
    #     (every element entry lands in exactly one slot of AA/AB/BA/BB)
    total = sum(float(np.sum(asm.values(w))) for w in ("AA", "AB", "BA", "BB"))
    # (3) pattern sanity: sorted, duplicate-free columns in every row
    outer, inner = asm.csr_pattern("AA")
    nnz = len(inner)
    dcol = np.diff(inner.astype(np.int64))
    row_starts = outer[1:-1]
    row_starts = row_starts[(row_starts > 0) & (row_starts < nnz)]
    interior = np.ones(nnz - 1, bool)
    interior[row_starts - 1] = False
    assert (dcol[interior] > 0).all(), f"{what}: AA columns not strictly ascending inside rows"
    # (4) translation invariance, the global check of the scatter: Fint(u + c) = Fint(u) for a rigid
    #     translation c, so K c = 0 row by row -- K_AA c_A + K_AB c_B = 0 and K_BA c_A + K_BB c_B = 0 --
    #     at ANY state (the shape-function derivatives sum to zero).  One misplaced, missing or doubled
    #     element entry anywhere in the 5e8 slots breaks it.
    _translation_invariance(asm, m, what)
    # (5) nothing is assembled at rest
    asm.assemble(np.zeros_like(d))
    pa0, _, pb0 = asm.vectors()
    assert np.abs(pa0).max() <= 1e-9 * max(1.0, np.abs(asm.values("AA")).max() * 1e-6), f"{what}: residual at rest"
    return asm, total


def _translation_invariance(asm, m, what, tol=1e-10):
    import scipy.sparse as sp
    gls, nf, nx = M.number_dofs(m)
    mats = {}
    for w in ("AA", "AB", "BA", "BB"):
        outer, inner = asm.csr_pattern(w)
        rows, cols, _ = asm.csr_dims(w)
        mats[w] = sp.csr_matrix((asm.values(w), inner, outer), shape=(rows, cols))
    for k in range(3):
        g = gls[:, k]
        cA = np.zeros(nf); cB = np.zeros(max(nx, 1))
        cA[g[g > 0] - 1] = 1.0
        cB[-g[g < 0] - 1] = 1.0
        cB = cB[:nx]
        for ra, rb in (("AA", "AB"), ("BA", "BB")):
            if mats[ra].shape[0] == 0:
                continue
            r = mats[ra] @ cA + (mats[rb] @ cB if nx else 0.0)
            scale = abs(mats[ra]) @ np.abs(cA) + (abs(mats[rb]) @ np.abs(cB) if nx else 0.0)
            bad = np.abs(r) > tol * np.maximum(scale, scale.max() * 1e-6)
            assert not bad.any(), f"{what}: K c != 0 for translation {k} in {int(bad.sum())} rows of {ra}|{rb} (worst {np.abs(r).max():.3e} vs {scale.max():.3e})"


def test_full_size_beam_line(port):
    """BASELINE.json configs[1]: 100k Beam_1 line."""
    m = M.beam_line(100_000)
    d = M.beam_line_displacements(m)
    asm, _ = _full_size_checks(port, m, d, 18, "100k beams")
    assert asm.n_free == 1_200_000
    # a 3-node beam couples its 18 DOFs: 324 entries per element, the 36 of the shared end node counted once;
    # the clamped first node takes its 6 rows and columns out (18 x 18 - 12 x 12 = 180 entries of element 1)
    assert asm.csr_dims("AA")[2] == 100_000 * 288 + 36 - 180


def test_full_size_shell_plate(port):
    """BASELINE.json configs[2]: 1M Shell_1 plate (1000 x 500 cells)."""
    m = M.shell_plate(1000, 500)
    d = M.shell_plate_displacements(m)
    asm, _ = _full_size_checks(port, m, d, 27, "1M shells")
    nnz = asm.csr_dims("AA")[2]
    assert abs(nnz / m.n_elements - 517.5) < 2.0        # ~517 non-zeros per element


def test_full_size_solid_block(port):
    """BASELINE.json configs[3]: 4M Solid_1 block (builder-defined hexahedron, parity unpinned by the
    reference: 64 sampled elements are compared with the oracle port, the builder's own CPU restatement).  The element arena passes 2^30 doubles here (1.3e9 with the upper-triangle storage; the
    unsigned 32-bit block offsets of the slot map were exercised beyond 2^31 with the full 64-block layout
    earlier in the round); the check is the global translation invariance of the assembled tangent --
    every lower block is read as the transpose of its stored twin."""
    m = M.solid_block(160, 160, 156)
    assert m.n_elements * 324 > 2 ** 30
    d = M.solid_block_displacements(m)
    asm = capi.Assembler(m).set_dofs()
    asm.assemble(d)
    _sampled_elements_against_port(port, m, d, asm, "4M solids")
    _translation_invariance(asm, m, "4M solids")
    asm.close()


def test_gauss_point_results_against_oracle(port):
    """gfa_gauss_point_results on a mixed model, before and after a state commit: the records come
    back per element type in element order; Solid_1 keeps none (as in the reference)."""
    m = M.concat_models([M.beam_line(37, pretension=1.0e4), M.pipe_line(11), M.shell_plate(7, 5, warp=0.005), M.solid_block(3, 3, 2), M.pipe_line(5)])
    rng = np.random.default_rng(3)
    asm = capi.Assembler(m).set_dofs()
    port.load(m)
    for it in range(2):
        d = M.mask_displacements(m, rng.uniform(-2e-4, 2e-4, (m.n_nodes, 6)))
        port.assemble(d)
        asm.assemble(d)
        for etype in (M.BEAM_1, M.PIPE_1, M.SHELL_1):
            ids = np.nonzero(m.elem_type == etype)[0]
            ref = np.array([port.results(int(e)) for e in ids])
            util.assert_results_parity(ref, asm.gauss_point_results(etype), f"iteration {it} type {etype}")
        port.commit()
        asm.commit()
    with pytest.raises(capi.GfaError):
        asm.gauss_point_results(M.SOLID_1)


def test_newton_vector_steps_on_device():
    """gfa_residual / gfa_update_displacements / assemble-from-the-device-copy against the reference-generated
    fixture: bit-exact right-hand side given the same P_A, first-node semantics of the max-norms, exact update."""
    from oracle import newton_steps as ns
    z = _golden("newton_steps")
    m = util.model_from_dict(z)
    asm = capi.Assembler(m).set_dofs()
    asm.set_time(0.0, 1.0)
    assert (asm.gls == z["gls"]).all()
    asm.assemble(z["disp"])
    pa = asm.vectors()[0]
    expect = ns.residual(pa, asm.csr("AB"), z["X_B"])          # same arithmetic on the device's own P_A
    got = asm.residual(z["X_B"])
    rhs = asm.vectors()[0]
    assert np.array_equal(rhs, expect)
    util.assert_parity(z["rhs"], rhs, "right-hand side vs reference")
    n = ns.residual_norms(z["gls"], rhs)
    assert (got["max_force"], got["max_moment"]) == (n["max_force"], n["max_moment"])
    assert (got["node_force"], got["node_moment"], got["nan_detected"]) == (n["node_force"], n["node_moment"], 0)
    assert (got["node_force"], got["node_moment"]) == tuple(int(v) for v in z["residual_nodes"][:2])
    inc = asm.update_displacements(z["x_A"])
    d2 = asm.displacements()
    assert np.array_equal(d2, z["disp_after"])
    ref_inc = ns.increment_norms(z["gls"], z["x_A"], z["disp_after"])
    for k in ("max_force", "max_moment", "node_force", "node_moment", "max_disp_value", "max_rot_value", "nan_detected"):
        assert inc[k] == ref_inc[k], k
    assert (inc["node_force"], inc["node_moment"]) == tuple(int(v) for v in z["increment_nodes"][:2])
    # the next iteration assembles the device copy: same system as handing the updated array over
    asm.assemble(None)
    v1 = asm.values("AA").copy(); p1 = asm.vectors()[0].copy()
    asm.assemble(z["disp_after"])
    assert np.array_equal(v1, asm.values("AA")) and np.array_equal(p1, asm.vectors()[0])
    # NaN in the increment is flagged, not propagated into the maxima
    x = z["x_A"].copy(); x[3] = np.nan
    assert asm.update_displacements(x)["nan_detected"] == 1


# ---- Newmark dynamics -----------------------------------------------------
@pytest.mark.parametrize("name", ["dynamic_beam", "dynamic_shell", "dynamic_pipe"])
def test_newmark_dynamics_against_reference_fixture(name):
    """gfa_update_dyn + gfa_assemble_dynamic + gfa_commit_state against what the reference's own Dynamic /
    MountMass / MountDamping / MountDyn / UpdateDyn produced (tests/golden/make_golden.py): CSR values,
    vectors, vel/accel (incl. nodes with partly-free rotations), alpha_i after the commit."""
    z = _golden(name)
    m = util.model_from_dict(z)
    asm = capi.Assembler(m).set_dofs()
    asm.set_time(*z["time"])
    assert (asm.gls == z["gls"]).all()
    util.run_dynamic(asm, m, z, util.check_dynamic(z, (1, m.n_elements - 1), name))


def test_newmark_dynamics_against_oracle(port):
    """Mixed Beam_1 + Shell_1 model, larger than the fixtures, with and without Rayleigh damping; the stored
    rayleigh_damping must survive iterations that do not update it."""
    m = M.concat_models([M.beam_line(70, pretension=3.0e4), M.shell_plate(9, 7, warp=0.015, gravity=(0.0, 0.0, -9.81))])
    d = M.mask_displacements(m, np.random.default_rng(11).uniform(-1e-3, 1e-3, (m.n_nodes, 6)))
    for k, rayleigh in enumerate([(0.0, 0.0), (0.5, 3.0e-4)]):
        scen = util.dynamic_scenario(m, d, 500 + k, time_step=0.002, rayleigh=rayleigh)
        port.load(m)
        port.set_time(0.0, 0.4)
        z = dict(scen)
        els = (0, 69, 70, m.n_elements - 1)
        util.run_dynamic(port, m, scen, util.capture_dynamic(z, els))
        asm = capi.Assembler(m).set_dofs()
        asm.set_time(0.0, 0.4)
        util.run_dynamic(asm, m, scen, util.check_dynamic(z, els, f"dynamic mixed rayleigh={rayleigh}"))
        asm.close()


def test_static_steps_keep_alpha_i_and_dynamic_rejects_unsupported_types(port):
    """SaveLagrange updates alpha_i in static steps too (Shell_1.cpp:1659, Beam_1.cpp:1502); Solid_1 has no
    dynamic path (the reference has no arithmetic for it)."""
    m = M.concat_models([M.beam_line(6), M.shell_plate(3, 2, warp=0.01)])
    d = M.mask_displacements(m, np.random.default_rng(2).uniform(-2e-3, 2e-3, (m.n_nodes, 6)))
    port.load(m)
    asm = capi.Assembler(m).set_dofs()
    for _ in range(2):
        port.assemble(d); asm.assemble(d)
        port.commit(); asm.commit()
    for e in (0, 5, 6, m.n_elements - 1):
        util.assert_parity(port.alpha_i(e), asm.alpha_i(e), f"alpha_i of element {e} after two static commits")
    asm.close()
    ms = M.solid_block(2, 2, 2)
    a2 = capi.Assembler(ms).set_dofs()
    a2.set_dynamic(util.newmark_coefficients(0.01))
    with pytest.raises(capi.GfaError):
        a2.assemble_dynamic(np.zeros((ms.n_nodes, 6)), True)
    a2.close()


def test_full_size_shell_plate_dynamics_conserve_mass():
    """BASELINE.json configs[2] through gfa_assemble_dynamic: size-independent properties of the Newmark path.
    For a rigid translation c_k the stiffness gives K c = 0, the consistent mass a1 * integral(rho t N_a) and the
    mass-proportional Rayleigh term a4 * alpha * (the same with the 6-point rule), so the sum over ALL rows of
    direction k of (K + M + a4 C) c_k is (a1 + a4 alpha) * rho * t * Area; a uniform previous acceleration g0
    gives inertial forces that sum to -a3 * rho * t * Area * g0.  One misplaced or doubled mass entry breaks it."""
    import scipy.sparse as sp
    m = M.shell_plate(1000, 500)
    asm = capi.Assembler(m).set_dofs()
    a = util.newmark_coefficients(1.0e-3)
    ray_alpha, g0 = 0.3, 2.5
    asm.set_dynamic(a, ray_alpha, 0.0)
    zeros = np.zeros((m.n_nodes, 6))
    acc = zeros.copy(); acc[:, 2] = g0
    asm.set_kinematics(zeros, zeros, zeros, acc)
    asm.assemble_dynamic(zeros, True)
    E, nu, rho = m.hooke[0]
    t = float(m.shell_thickness[0])
    area = 1000 * 500 * 0.0195 ** 2
    mass = rho * t * area
    gls, nf, nx = M.number_dofs(m)
    mats = {}
    for w in ("AA", "AB", "BA", "BB"):
        outer, inner = asm.csr_pattern(w)
        rows, cols, _ = asm.csr_dims(w)
        mats[w] = sp.csr_matrix((asm.values(w), inner, outer), shape=(rows, cols))
    for k in range(3):
        g = gls[:, k]
        cA = np.zeros(nf); cB = np.zeros(nx)
        cA[g[g > 0] - 1] = 1.0
        cB[-g[g < 0] - 1] = 1.0
        rA = mats["AA"] @ cA + mats["AB"] @ cB
        rB = mats["BA"] @ cA + mats["BB"] @ cB
        total = rA[g[g > 0] - 1].sum() + rB[-g[g < 0] - 1].sum()
        expect = (a[0] + a[3] * ray_alpha) * mass
        assert abs(total - expect) <= 1e-9 * expect, f"direction {k}: sum of (K+M+a4 C) c = {total!r}, expected {expect!r}"
        # no coupling into the other directions or the rotations
        others = np.abs(rA).sum() + np.abs(rB).sum() - np.abs(rA[g[g > 0] - 1]).sum() - np.abs(rB[-g[g < 0] - 1]).sum()
        assert others <= 1e-8 * expect
    pa, _, pb = asm.vectors()
    gz = gls[:, 2]
    fz = pa[gz[gz > 0] - 1].sum() + pb[-gz[gz < 0] - 1].sum()
    assert abs(fz - (-a[2] * mass * g0)) <= 1e-9 * abs(a[2] * mass * g0), f"inertial force sum {fz!r}"
    asm.close()


def test_newmark_dynamics_large_rotations(port):
    """Rotation increments of ~0.5 rad and angular velocities of ~3 rad/s over three committed time steps
    (the oracle is pinned to the reference in this regime by tests/test_oracle_vs_ref.py)."""
    m = M.concat_models([M.beam_line(6), M.shell_plate(3, 2, warp=0.02)])
    rng = np.random.default_rng(4242)
    d = M.mask_displacements(m, rng.uniform(-1.0, 1.0, (m.n_nodes, 6)) * np.array([2e-3, 2e-3, 2e-3, 0.5, 0.5, 0.5]))
    scen = util.dynamic_scenario(m, d, 99, time_step=0.02, rayleigh=(0.2, 5.0e-5))
    scen["dyn_copy_vel"] = scen["dyn_copy_vel"] * np.array([1, 1, 1, 6.0, 6.0, 6.0])
    els = (0, 5, 6, m.n_elements - 1)
    z = dict(scen)
    old = util.DYN_STEPS
    util.DYN_STEPS = (("s1", 0, True, True), ("s2", 1, False, True), ("s3", 2, False, True))
    try:
        port.load(m); port.set_time(0.0, 0.5)
        util.run_dynamic(port, m, scen, util.capture_dynamic(z, els))
        asm = capi.Assembler(m).set_dofs()
        asm.set_time(0.0, 0.5)
        util.run_dynamic(asm, m, scen, util.check_dynamic(z, els, "large rotations"))
        asm.close()
    finally:
        util.DYN_STEPS = old


def test_shell_load_host_contributor_through_the_c_abi():
    """ShellLoad follower pressure (a Load: host side) pushed with gfa_add_host_triplets / gfa_add_host_vector on top
    of the device assembly, against what the reference's MountLoads + MountGlobal produced (AreaUpdate 0 and 1,
    before and after a commit).  Its non-symmetric u-u blocks land in existing slots of the element pattern."""
    z = _golden("shell_load")
    m = util.model_from_dict(z)
    asm = capi.Assembler(m).set_dofs()
    asm.set_time(*z["time"])
    assert (asm.gls == z["gls"]).all()
    t = float(z["time"][0] + z["time"][1])
    for tag, commit in (("it1", False), ("it2", True), ("it3", False)):
        disp = z[f"{tag}_disp"]
        asm.assemble(disp)
        trip, pa_add, pb_add = util.shell_load_contribution(m, asm.gls, disp, asm.copy_coordinates(), t)
        for w in ("AA", "AB", "BA", "BB"):
            if trip[w][0]:
                asm.add_host_triplets(w, *trip[w])
        asm.add_host_vector(capi.P_A, *pa_add)
        asm.add_host_vector(capi.I_A, *pa_add)
        if pb_add[0]:
            asm.add_host_vector(capi.P_B, *pb_add)
        util.assert_system_parity(lambda w: util.captured_csr(z, tag, w), asm.csr, f"shell_load {tag}")
        for v, key in zip(asm.vectors(), ("PA", "IA", "PB")):
            util.assert_parity(z[f"{tag}_{key}"], v, f"shell_load {tag} {key}")
        if commit:
            asm.commit()
    asm.close()


def test_shell_load_on_the_device():
    """The same fixture with the follower pressure evaluated by the library's own kernel (gfa_set_shell_loads /
    gfa_apply_shell_loads): AreaUpdate 0 and 1 on different element sets, before and after a commit."""
    z = _golden("shell_load")
    m = util.model_from_dict(z)
    asm = capi.Assembler(m).set_dofs()
    asm.set_time(*z["time"])
    asm.set_shell_loads(m.shell_loads)
    t = float(z["time"][0] + z["time"][1])
    for tag, commit in (("it1", False), ("it2", True), ("it3", False)):
        asm.assemble(z[f"{tag}_disp"])
        asm.apply_shell_loads(t)
        util.assert_system_parity(lambda w: util.captured_csr(z, tag, w), asm.csr, f"device shell_load {tag}")
        for v, key in zip(asm.vectors(), ("PA", "IA", "PB")):
            util.assert_parity(z[f"{tag}_{key}"], v, f"device shell_load {tag} {key}")
        if commit:
            asm.commit()
    asm.close()


@pytest.mark.parametrize("name,commits", [("pipe_load", (("it1", False), ("it2", True), ("it3", False))),
                                          ("tutorial04", (("it1", True), ("it2", False))),
                                          ("tutorial03", (("it1", True), ("it2", False)))])
def test_pipe_load_on_the_device(name, commits):
    """PipeLoad internal pressure evaluated by the library's own kernel (gfa_set_pipe_loads / gfa_apply_pipe_loads)
    against the reference-generated fixtures: a bent pipe line with two loads, and inputs/tutorial04 as shipped
    (its NodalLoad stays a host contributor and enters through gfa_add_host_*), and inputs/tutorial03 (Pipe_1 with a
    NodalLoad and a NodalFollowerLoad, both host contributors)."""
    from test_oracle_golden import _run_pipe_fixture
    z = _golden(name)
    m = util.model_from_dict(z)
    asm = capi.Assembler(m)
    gls, nf, nx = asm.number_dofs()
    assert (gls.reshape(-1, 6) == z["gls"]).all()
    asm.set_dofs(gls, nf, nx)
    asm.set_time(*z["time"])
    asm.set_pipe_loads(m.pipe_loads)
    vec = {"PA": capi.P_A, "PB": capi.P_B}
    _run_pipe_fixture(z, m, asm, f"device {name}", commits, apply_loads=asm.apply_pipe_loads,
                      add_triplets=asm.add_host_triplets, add_vector=lambda k, i, v: asm.add_host_vector(vec[k], i, v))
    asm.close()


def test_pipe_load_random_against_oracle(port):
    """Device pipe pressure against the oracle port (pinned to the reference on the same generator by
    tests/test_oracle_vs_ref.py) on seeded pipe lines mixed with beams and shells: random element sets, pressures of
    both signs, large displacements, a commit between iterations."""
    rng = np.random.default_rng(20240052)
    for trial in range(4):
        n = int(rng.integers(3, 40))
        pipes = M.pipe_line(n)
        m = M.concat_models([M.beam_line(int(rng.integers(2, 6))), pipes, M.shell_plate(2, 2, warp=0.01)]) if trial % 2 else pipes
        first = int(np.nonzero(m.elem_type == M.PIPE_1)[0][0]) + 1              # 1-based id of the first pipe element
        loads = []
        for _ in range(int(rng.integers(1, 4))):
            els = (first + np.sort(rng.choice(np.arange(n), size=int(rng.integers(1, n + 1)), replace=False))).astype(np.int32)
            p = float(rng.uniform(-4.0e8, 4.0e8))
            loads.append((els, np.array([[0.0, 0.1 * p, 0, 0, 0], [1.0, p, 0, 0, 0]])))
        m.pipe_loads = loads
        port.load(m)
        port.set_time(0.2, 0.5)
        asm = capi.Assembler(m).set_dofs()
        asm.set_time(0.2, 0.5)
        asm.set_pipe_loads(m.pipe_loads)
        amp = 2e-3 if trial % 2 else 3e-2          # the shell cells of the mixed models are 2 cm wide
        d = M.mask_displacements(m, rng.uniform(-amp, amp, (m.n_nodes, 6)))
        for it in range(3):
            port.assemble(d)
            asm.assemble(d)
            asm.apply_pipe_loads(0.7)
            util.assert_system_parity(port.csr, asm.csr, f"device pipe pressure trial {trial} it{it}")
            for a, b, key in zip(port.vectors(), asm.vectors(), ("PA", "IA", "PB")):
                util.assert_parity(a, b, f"device pipe pressure trial {trial} it{it} {key}")
            if it == 0:
                port.commit(); asm.commit()
            d = -0.6 * d
        asm.close()


def test_random_models_against_oracle(port):
    """Seeded random variations (the oracle is pinned to the reference on the same generator by
    tests/test_oracle_vs_ref.py): random constraint masks on random nodes, sizes, warps, gravity, displacement
    amplitudes from 1e-6 to 1e-2, with a commit between iterations."""
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
        port.load(m)
        port.set_time(0.0, 0.3)
        asm = capi.Assembler(m).set_dofs()
        asm.set_time(0.0, 0.3)
        assert (asm.gls == port.gls()).all(), f"trial {trial}: DOF numbering"
        for it in range(2):
            port.assemble(d); asm.assemble(d)
            _compare_system(port, asm, f"random trial {trial} it{it}")
            port.commit(); asm.commit()
            d = -0.7 * d
        asm.close()


def test_multi_gpu_parity_under_torchrun():
    """Two ranks (one per GPU) assemble their element partitions, exchange the interface rows over NCCL and every
    rank compares the rows it owns with the single-process oracle: tests/multi_gpu_check.py under torchrun.
    Skipped on a box with one GPU (the driver's 1 -> 8 scaling run exercises the same path through bench.py)."""
    import subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(here, "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "multi-GPU parity OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
