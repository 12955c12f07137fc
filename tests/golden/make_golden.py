"""Generate the committed golden fixtures from the REFERENCE'S OWN sources.

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Every fixture is produced by ``oracle/_ref/libgiraffe_ref.so`` -- the unmodified
reference Beam_1/Shell_1/Node/Solution/NodalLoad translation units driven the
way ``Static::Solve`` drives them (see oracle/ref_shims/ref_driver.cpp).  The
fixtures hold the model tables, the nodal displacements of each captured
Newton iteration and the reference's CSR matrices (AA/AB/BA/BB) and vectors
(P_A, I_A, P_B) for that iteration.

  tutorial01.npz   inputs/tutorial01 as shipped: 5 Beam_1, clamped, FX ramp;
                   iterations 1 and 2 of increment 1 (BASELINE.json configs[0]).
                   Iteration 2 uses the displacements after the first Newton
                   update (K_AA solved here with scipy; the solve is not part
                   of the path).
  beam_line.npz    24 Beam_1, Tube section, pre-tension, gravity; two
                   iterations, a commit (SaveLagrange) and one more iteration.
  shell_plate.npz  6x4-cell warped Shell_1 plate with gravity (doubled
                   self-weight quirk), same sequence.
  shell_load.npz   ShellLoad follower pressure (AreaUpdate 0 and 1) folded into the shell blocks
  pipe_load.npz    PipeLoad internal pressure on a bent Pipe_1 line; tutorial04.npz / tutorial03.npz: the shipped inputs with their PipeLoad / NodalFollowerLoad
                   by MountLoads: the host-contributor seam of gfa_add_host_triplets.
  dynamic_beam.npz, dynamic_shell.npz, dynamic_pipe.npz
                   Newmark path (Dynamic.cpp:303-340): UpdateDyn, MountMass,
                   MountDamping (Rayleigh update on the first iteration), MountDyn
                   over two iterations, a commit and a third iteration.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from giraffe_b200 import meshes as M            # noqa: E402
from giraffe_b200.inp import read_inp            # noqa: E402
from oracle.refdrv import RefOracle              # noqa: E402
import util                                      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def tutorial01(R):
    m, info = read_inp("/root/reference/inputs/tutorial01/tutorial01.inp")
    R.load(m)
    dt = info["time_step"]
    R.set_time(0.0, dt)
    z = util.model_to_dict(m)
    z["time"] = np.array([0.0, dt])
    d = np.zeros((m.n_nodes, 6))
    R.assemble(d, with_loads=True)
    z["it1_disp"] = d.copy()
    z.update(util.capture(R, "it1"))
    z["it1_triplets"] = np.array([R.triplets(w) for w in ("AA", "AB", "BA", "BB")])
    # Newton update (Static.cpp:210-236): P_A = -P_A ; x = K_AA^-1 P_A ; UpdateDisps
    o, i, v, s = R.csr("AA")
    K = sp.csr_matrix((v, i, o), shape=s)
    pa, _, _ = R.vectors()
    x = spla.spsolve(K.tocsc(), -pa)
    gls = R.gls()
    d2 = d.copy()
    free = gls > 0
    d2[free] += x[gls[free] - 1]
    R.assemble(d2, with_loads=True)
    z["it2_disp"] = d2.copy()
    z.update(util.capture(R, "it2"))
    z["gls"] = gls
    z["elem0_K"], z["elem0_P"], _ = R.element(0)
    np.savez_compressed(os.path.join(OUT, "tutorial01.npz"), **z)
    print("tutorial01: n_free", R.n_free, "nnz_AA", len(v), "triplets", z["it1_triplets"])


def sequence(R, m, disp_fn, name, scale2=0.6):
    R.load(m)
    R.set_time(0.0, 0.7)
    z = util.model_to_dict(m)
    z["time"] = np.array([0.0, 0.7])
    z["gls"] = R.gls()
    d1 = disp_fn(m)
    steps = [("it1", d1, False), ("it2", scale2 * d1, True), ("it3", -0.35 * d1, False)]
    for tag, d, commit_after in steps:
        R.assemble(d)
        z[f"{tag}_disp"] = d.copy()
        z.update(util.capture(R, tag))
        K, P, en = R.element(1)
        z[f"{tag}_elem1_K"], z[f"{tag}_elem1_P"], z[f"{tag}_elem1_energy"] = K, P, np.array([en])
        # Gauss-point results the elements keep for WriteResults / WriteMonitor (layout of gfa_gauss_point_results)
        z[f"{tag}_results"] = np.array([R.results(e) for e in range(m.n_elements)])
        if commit_after:
            R.commit()
            z[f"{tag}_state1"] = R.state(1)
            z[f"{tag}_copy"] = R.copy_coordinates()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **z)
    print(name, "n_free", R.n_free, "nnz_AA", len(z["it1_AA_val"]))


def shipped_shell_mesh(R, name):
    """A shell mesh the reference ships (inputs/tutorial05: 400 Shell_1, SURVEY.md 7's minimum slice;
    inputs/tutorial02: 3036 Shell_1), read with giraffe_b200/inp.py and assembled by the reference's own sources at
    seeded nodal increments: an iteration, a commit, a second iteration.  tutorial05 keeps every CSR value;
    tutorial02 (1.6 M non-zeros per capture) keeps the pattern digest, the vectors, the row sums of AA (one
    misplaced, missing or doubled value changes one of them) and 20 000 sampled values."""
    import hashlib
    m, info = read_inp(f"/root/reference/inputs/{name}/{name}.inp")
    R.load(m)
    R.set_time(0.0, 0.5)
    z = util.model_to_dict(m)
    z["time"] = np.array([0.0, 0.5])
    z["gls"] = R.gls()
    rng = np.random.default_rng(20240031)
    d1 = M.mask_displacements(m, np.concatenate([rng.uniform(-2e-4, 2e-4, (m.n_nodes, 3)), rng.uniform(-5e-3, 5e-3, (m.n_nodes, 3))], axis=1))
    full = m.n_elements <= 500
    for tag, d, commit_after in (("it1", d1, True), ("it2", -0.4 * d1, False)):
        R.assemble(d)
        z[f"{tag}_disp"] = d.copy()
        if full:
            z.update(util.capture(R, tag))
        else:
            o, i, v, s = R.csr("AA")
            z[f"{tag}_AA_shape"] = np.array(s)
            z[f"{tag}_AA_pattern_sha256"] = np.frombuffer(hashlib.sha256(o.tobytes() + i.tobytes()).digest(), np.uint8)
            rows = np.repeat(np.arange(s[0]), np.diff(o))
            z[f"{tag}_AA_rowsum"] = np.bincount(rows, weights=v, minlength=s[0])
            z[f"{tag}_AA_rowabs"] = np.bincount(rows, weights=np.abs(v), minlength=s[0])
            z[f"{tag}_AA_diag"] = util.csr_diag((o, i, v, s))
            pick = np.sort(rng.choice(len(v), size=20000, replace=False))
            z[f"{tag}_AA_sample_idx"], z[f"{tag}_AA_sample_val"] = pick, v[pick]
            z[f"{tag}_nnz"] = np.array([len(R.csr(w)[2]) for w in ("AA", "AB", "BA", "BB")])
            pa, ia, pb = R.vectors()
            z[f"{tag}_PA"], z[f"{tag}_IA"], z[f"{tag}_PB"] = pa, ia, pb
        K, P, en = R.element(m.n_elements // 2)
        z[f"{tag}_elem_K"], z[f"{tag}_elem_P"] = K, P
        if commit_after:
            R.commit()
    np.savez_compressed(os.path.join(OUT, name + "_shells.npz"), **z)
    print(name, "elements", m.n_elements, "n_free", R.n_free, "nnz_AA", len(R.csr("AA")[2]))


def newton_steps(R):
    """The vector steps either side of the assembly, through the reference's own code (Static.cpp:210-217,
    ConvergenceCriteria.cpp, Solution::UpdateDisps): a beam + shell model with a prescribed-displacement set."""
    m = M.concat_models([M.beam_line(12, pretension=1.0e4), M.shell_plate(5, 3, warp=0.01)])
    m.constraints = m.constraints + [([7, 30], 0x07), ([40], 0x3F)]
    rng = np.random.default_rng(20240005)
    d = M.mask_displacements(m, rng.uniform(-1e-3, 1e-3, (m.n_nodes, 6)))
    R.load(m)
    R.set_time(0.0, 1.0)
    R.assemble(d)
    z = util.model_to_dict(m)
    z["gls"] = R.gls()
    z["disp"] = d
    z["X_B"] = rng.uniform(-1e-3, 1e-3, R.n_fixed)
    nf, nm, div = R.residual(z["X_B"])
    z["rhs"] = R.vectors()[0]
    z["residual_nodes"] = np.array([nf, nm, div])
    z["x_A"] = rng.uniform(-1e-4, 1e-4, R.n_free)
    d2, nd, nr, div = R.update_displacements(z["x_A"])
    z["disp_after"] = d2
    z["increment_nodes"] = np.array([nd, nr, div])
    np.savez_compressed(os.path.join(OUT, "newton_steps.npz"), **z)
    print("newton_steps: n_free", R.n_free, "n_fixed", R.n_fixed, "nodes", z["residual_nodes"], z["increment_nodes"])


def dynamic_models():
    """Beam_1 line and warped Shell_1 plate for the Newmark path; each has a node whose rotational DOFs are
    only partly free (UpdateDyn's vel_aux / ace_aux carry-over, Dynamic.cpp:493-556)."""
    mb = M.beam_line(14, pretension=2.0e5)
    mb.gravity = (0.4, -0.3, -9.81)
    mb.constraints = mb.constraints + [([9], 0x08), ([10], 0x38), ([16], 0x30)]
    ms = M.shell_plate(5, 3, warp=0.01, gravity=(0.0, 0.0, -9.81))
    mids = sorted(set(int(n) for n in ms.elem_nodes.reshape(-1, 6)[:, 3:].reshape(-1)))
    ms.constraints = ms.constraints + [([mids[7]], 0x10), ([mids[20]], 0x28)]
    db = M.mask_displacements(mb, M.beam_line_displacements(mb))
    db[8, 3] = 2.0e-3; db[9, 3:6] = (1.0e-3, -2.0e-3, 1.5e-3); db[15, 4:6] = (-1.0e-3, 0.5e-3)      # prescribed rotations
    ds = M.mask_displacements(ms, M.shell_plate_displacements(ms))
    ds[mids[7] - 1, 4] = 1.0e-3; ds[mids[20] - 1, 3] = -1.0e-3; ds[mids[20] - 1, 5] = 2.0e-3
    mp = M.pipe_line(10, gravity=(0.3, -0.2, -9.81))          # Pipe_1 with its structural mass (no ocean data)
    mp.constraints = mp.constraints + [([8], 0x10)]
    dp = M.mask_displacements(mp, M.beam_line_displacements(mp))
    dp[7, 4] = -1.5e-3
    return (("dynamic_beam", mb, db, 20240011), ("dynamic_shell", ms, ds, 20240012), ("dynamic_pipe", mp, dp, 20240013))


def dynamic(R):
    """Newmark path through the reference's own Dynamic / MountMass / MountDamping / MountDyn / UpdateDyn."""
    for name, m, d, seed in dynamic_models():
        R.load(m)
        R.set_time(0.0, 0.7)
        z = util.model_to_dict(m)
        z["time"] = np.array([0.0, 0.7])
        z["gls"] = R.gls()
        z.update(util.dynamic_scenario(m, d, seed))
        util.run_dynamic(R, m, z, util.capture_dynamic(z, (1, m.n_elements - 1)))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **z)
        print(name, "n_free", R.n_free, "nnz_AA", len(z["s1_AA_val"]))


def shell_load(R):
    """ShellLoad follower pressure (ShellLoad.cpp:133-148 -> Shell_1::MountShellSpecialLoads, Shell_1.cpp:1392-1467),
    one load with AreaUpdate 0 and one with AreaUpdate 1 on different element sets of a warped plate: two
    iterations, a commit (so that copy_coordinates differ from the reference ones) and a third iteration."""
    m = M.shell_plate(5, 4, warp=0.02, gravity=(0.0, 0.0, -9.81))
    m.shell_loads = [(np.array([2, 5, 6, 11, 17, 30], np.int32), False, np.array([[0.0, 0.0], [1.0, 6.0e7]])),
                     (np.array([9, 10, 23, 38], np.int32), True, np.array([[0.0, 1.0e7], [1.0, -4.0e7]]))]
    R.load(m)
    R.set_time(0.0, 0.7)
    z = util.model_to_dict(m)
    z["time"] = np.array([0.0, 0.7])
    z["gls"] = R.gls()
    rng = np.random.default_rng(20240021)
    d1 = M.mask_displacements(m, rng.uniform(-2e-3, 2e-3, (m.n_nodes, 6)))
    for tag, d, commit_after in (("it1", d1, False), ("it2", 0.6 * d1, True), ("it3", -0.35 * d1, False)):
        z[f"{tag}_copy_before"] = R.copy_coordinates()
        R.assemble(d, with_loads=True)
        z[f"{tag}_disp"] = d.copy()
        z.update(util.capture(R, tag))
        if commit_after:
            R.commit()
    np.savez_compressed(os.path.join(OUT, "shell_load.npz"), **z)
    print("shell_load: n_free", R.n_free, "nnz_AA", len(z["it1_AA_val"]))


def pipe_load(R):
    """PipeLoad internal pressure (PipeLoad.cpp:117-133 -> Pipe_1::MountPipeSpecialLoads, Pipe_1.cpp:1443-1494): two loads
    on overlapping element sets of a pipe line bent by its displacements, self-weight on: two iterations, a commit
    (so that Q_i, z'_i, kappa_i differ from the reference state) and a third iteration."""
    m = M.pipe_line(12)
    m.gravity = (0.0, 0.4, -9.81)
    m.pipe_loads = [(np.array([1, 2, 3, 5, 9], np.int32), np.array([[0.0, 0.0, 0, 0, 0], [1.0, 3.0e8, 1.0e5, 900.0, 1025.0]])),
                    (np.array([5, 6, 12], np.int32), np.array([[0.0, 2.0e7, 0, 0, 0], [1.0, -2.5e8, 0, 0, 0]]))]
    R.load(m)
    R.set_time(0.0, 0.7)
    z = util.model_to_dict(m)
    z["time"] = np.array([0.0, 0.7])
    z["gls"] = R.gls()
    rng = np.random.default_rng(20240041)
    d1 = M.mask_displacements(m, rng.uniform(-3e-2, 3e-2, (m.n_nodes, 6)))
    for tag, d, commit_after in (("it1", d1, False), ("it2", 0.6 * d1, True), ("it3", -0.35 * d1, False)):
        R.assemble(d, with_loads=True)
        z[f"{tag}_disp"] = d.copy()
        z.update(util.capture(R, tag))
        if commit_after:
            R.commit()
    np.savez_compressed(os.path.join(OUT, "pipe_load.npz"), **z)
    print("pipe_load: n_free", R.n_free, "nnz_AA", len(z["it1_AA_val"]))


def tutorial04(R):
    """inputs/tutorial04 as the reference ships it (50 Pipe_1, a NodalLoad perturbation and a PipeLoad internal pressure
    that ramps up in the second solution step): two Newton iterations at t = 1.5 + 0.005 -- inside the pressure ramp --
    with the loads mounted, a commit in between."""
    from giraffe_b200.inp import read_inp
    m, info = read_inp("/root/reference/inputs/tutorial04/tutorial04.inp")
    assert len(m.pipe_loads) == 1 and len(m.nodal_loads) == 1
    R.load(m)
    z = util.model_to_dict(m)
    z["time"] = np.array([1.5, 0.005])
    R.set_time(1.5, 0.005)
    z["gls"] = R.gls()
    rng = np.random.default_rng(20240042)
    d1 = M.mask_displacements(m, rng.uniform(-5e-3, 5e-3, (m.n_nodes, 6)))
    for tag, d, commit_after in (("it1", d1, True), ("it2", -0.4 * d1, False)):
        R.assemble(d, with_loads=True)
        z[f"{tag}_disp"] = d.copy()
        z.update(util.capture(R, tag))
        if commit_after:
            R.commit()
    np.savez_compressed(os.path.join(OUT, "tutorial04.npz"), **z)
    print("tutorial04: n_free", R.n_free, "nnz_AA", len(z["it1_AA_val"]))


def tutorial03(R):
    """inputs/tutorial03 as the reference ships it (50 Pipe_1, a NodalLoad imperfection and a NodalFollowerLoad that
    compresses the pipe in the second solution step): two Newton iterations at t = 1.5 + 0.005 with the loads
    mounted, a commit in between (the follower load reads the committed rotations)."""
    from giraffe_b200.inp import read_inp
    m, info = read_inp("/root/reference/inputs/tutorial03/tutorial03.inp")
    assert len(m.follower_loads) == 1 and len(m.nodal_loads) == 1
    R.load(m)
    z = util.model_to_dict(m)
    z["time"] = np.array([1.5, 0.005])
    R.set_time(1.5, 0.005)
    z["gls"] = R.gls()
    rng = np.random.default_rng(20240043)
    d1 = M.mask_displacements(m, rng.uniform(-5e-3, 5e-3, (m.n_nodes, 6)))
    for tag, d, commit_after in (("it1", d1, True), ("it2", -0.4 * d1, False)):
        z[f"{tag}_copy_before"] = R.copy_coordinates()
        R.assemble(d, with_loads=True)
        z[f"{tag}_disp"] = d.copy()
        z.update(util.capture(R, tag))
        if commit_after:
            R.commit()
    np.savez_compressed(os.path.join(OUT, "tutorial03.npz"), **z)
    print("tutorial03: n_free", R.n_free, "nnz_AA", len(z["it1_AA_val"]))


if __name__ == "__main__":
    R = RefOracle(threads=1)
    if sys.argv[1:] == ["pipe_load"]:
        pipe_load(R)
        tutorial04(R)
        tutorial03(R)
        sys.exit(0)
    if sys.argv[1:] == ["dynamic"]:          # only the fixtures of the Newmark path
        dynamic(R)
        sys.exit(0)
    if sys.argv[1:] == ["shell_load"]:
        shell_load(R)
        sys.exit(0)
    if sys.argv[1:] == ["shipped"]:
        shipped_shell_mesh(R, "tutorial05")
        shipped_shell_mesh(R, "tutorial02")
        sys.exit(0)
    shell_load(R)
    pipe_load(R)
    tutorial04(R)
    tutorial03(R)
    dynamic(R)
    newton_steps(R)
    tutorial01(R)
    shipped_shell_mesh(R, "tutorial05")
    shipped_shell_mesh(R, "tutorial02")
    mb = M.beam_line(24, pretension=2.0e5)
    mb.gravity = (0.4, -0.3, -9.81)
    sequence(R, mb, M.beam_line_displacements, "beam_line")
    mp = M.pipe_line(16, gravity=(0.3, -0.2, -9.81))      # Pipe_1 riser segment
    sequence(R, mp, M.beam_line_displacements, "pipe_line")
    ms = M.shell_plate(6, 4, warp=0.01, gravity=(0.0, 0.0, -9.81))
    sequence(R, ms, M.shell_plate_displacements, "shell_plate")
