"""Shared helpers of the parity tests: the parity metric, CSR comparison and
(de)serialisation of models / captured reference results as .npz fixtures."""
from __future__ import annotations

import os

import numpy as np

from giraffe_b200 import meshes as M

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: "Kt/Fint values must agree with the reference's CPU
# assembly to a relative 1e-12".  Pure relative error is undefined on the
# explicit zeros the reference pushes and is not attainable on cancellation
# residue by ANY re-ordering of a FP64 sum: an entry K_ij is a sum of products
# whose magnitudes are bounded by sqrt(K_ii K_jj), so its rounding error is
# ~1e-16 * sqrt(K_ii K_jj) however small K_ij itself is (e.g. the shell's
# membrane/bending couplings are the residue of +/-zeta terms cancelling in the
# thickness quadrature, Shell_1.cpp:1127-1143; the reference itself moves by
# that much between MKL versions).  The criterion is therefore
#       |a - b| <= TOL * max(|a|, |b|, FLOOR * s_ij),   s_ij = sqrt(|K_ii| |K_jj|)
# with the diagonals taken from the REFERENCE matrices (AA for free DOFs, BB
# for fixed ones): component-wise 1e-12 for every entry within one decade of
# its diagonal scale, and 1e-13 * s_ij (diagonally-scaled norm-wise) below.
# Vectors and single blocks without a diagonal use their max magnitude as s.
TOL = 1e-12
FLOOR = 1e-1


def parity_error(ref: np.ndarray, got: np.ndarray, scale=None) -> float:
    """max over entries of |a-b| / max(|a|, |b|, FLOOR*scale); <= TOL means parity.
    `scale` is a scalar or an array broadcastable to the entries."""
    ref = np.asarray(ref, float)
    got = np.asarray(got, float)
    assert ref.shape == got.shape, (ref.shape, got.shape)
    if ref.size == 0:
        return 0.0
    if scale is None:
        scale = float(max(np.abs(ref).max(), np.abs(got).max()))
    denom = np.maximum(np.maximum(np.abs(ref), np.abs(got)), FLOOR * np.asarray(scale, float))
    diff = np.abs(ref - got)
    err = np.where(denom > 0, diff / np.where(denom > 0, denom, 1.0), np.where(diff > 0, np.inf, 0.0))
    return float(err.max())


def assert_parity(ref, got, what: str, scale=None, tol: float = TOL):
    e = parity_error(ref, got, scale)
    assert e <= tol, f"{what}: parity error {e:.3e} > {tol:.1e}"


def block_scale(K: np.ndarray) -> np.ndarray:
    """s_ij = sqrt(|K_ii| |K_jj|) of a square element block."""
    d = np.abs(np.diag(K))
    return np.sqrt(np.outer(d, d))


def csr_diag(csr) -> np.ndarray:
    outer, inner, val, shape = csr
    d = np.zeros(shape[0])
    rows = np.repeat(np.arange(shape[0]), np.diff(outer))
    on = rows == inner
    d[rows[on]] = np.abs(val[on])
    return d


def assert_csr_parity(ref_csr, got_csr, what: str, drow=None, dcol=None, tol: float = TOL):
    """Pattern byte-equal (Eigen outerIndexPtr / innerIndexPtr, int32) and values
    within the diagonally scaled criterion above."""
    ro, ri, rv, rs = ref_csr
    go, gi, gv, gs = got_csr
    assert tuple(rs) == tuple(gs), f"{what}: shape {gs} != {rs}"
    assert ro.dtype == go.dtype == np.int32 and ri.dtype == gi.dtype == np.int32
    assert ro.tobytes() == go.tobytes(), f"{what}: outerIndexPtr differs"
    assert ri.tobytes() == gi.tobytes(), f"{what}: innerIndexPtr differs"
    if len(rv) == 0:
        return 0.0
    if drow is None:
        drow = dcol = csr_diag(ref_csr)
    rows = np.repeat(np.arange(rs[0]), np.diff(ro))
    scale = np.sqrt(drow[rows] * dcol[ri])
    e = parity_error(rv, gv, scale)
    assert e <= tol, f"{what}: value parity error {e:.3e} > {tol:.1e}"
    return e


def assert_system_parity(ref_get, got_get, what: str, tol: float = TOL):
    """All four matrices; ref_get / got_get map 'AA'|'AB'|'BA'|'BB' to a CSR tuple."""
    rAA, rBB = ref_get("AA"), ref_get("BB")
    dA, dB = csr_diag(rAA), csr_diag(rBB)
    worst = 0.0
    for w, dr, dc in (("AA", dA, dA), ("AB", dA, dB), ("BA", dB, dA), ("BB", dB, dB)):
        e = assert_csr_parity(ref_get(w), got_get(w), f"{what} {w}", dr, dc, tol)
        worst = max(worst, e or 0.0)
    return worst


# ---- model <-> npz -------------------------------------------------------
def model_to_dict(m: M.Model, prefix: str = "m_") -> dict:
    d = {
        "xyz": m.xyz, "hooke": m.hooke, "sections": m.sections,
        "section_defs": np.array(m.section_defs, float).reshape(-1, 3),
        "shell_thickness": m.shell_thickness,
        "cs_defs": np.array([list(a) + list(b) for a, b in m.cs_defs], float).reshape(-1, 6),
        "cs": m.cs, "elem_type": m.elem_type, "elem_mat": m.elem_mat, "elem_sec": m.elem_sec,
        "elem_cs": m.elem_cs, "elem_ptr": m.elem_ptr, "elem_nodes": m.elem_nodes,
        "pretension": m.pretension if m.pretension is not None else np.zeros(0),
        "gravity": np.array(m.gravity if m.gravity is not None else [], float),
        "n_constraints": np.array([len(m.constraints)]),
        "n_loads": np.array([len(m.nodal_loads)]),
        "pipe_sections": np.asarray(m.pipe_sections, float).reshape(-1, 11),
        "n_shell_loads": np.array([len(m.shell_loads)]),
        "n_pipe_loads": np.array([len(getattr(m, "pipe_loads", []))]),
        "n_follower_loads": np.array([len(getattr(m, "follower_loads", []))]),
    }
    for i, (nodes, cs, table) in enumerate(getattr(m, "follower_loads", [])):
        d[f"f{i}_nodes"] = np.asarray(nodes, np.int32)
        d[f"f{i}_cs"] = np.array([cs])
        d[f"f{i}_table"] = np.asarray(table, float)
    for i, (elements, table) in enumerate(getattr(m, "pipe_loads", [])):
        d[f"p{i}_elements"] = np.asarray(elements, np.int32)
        d[f"p{i}_table"] = np.asarray(table, float)
    for i, (elements, area_update, table) in enumerate(m.shell_loads):
        d[f"s{i}_elements"] = np.asarray(elements, np.int32)
        d[f"s{i}_area_update"] = np.array([1 if area_update else 0])
        d[f"s{i}_table"] = np.asarray(table, float)
    for i, (nodes, mask) in enumerate(m.constraints):
        d[f"c{i}_nodes"] = np.asarray(nodes, np.int32)
        d[f"c{i}_mask"] = np.array([mask])
    for i, (nodes, cs, table) in enumerate(m.nodal_loads):
        d[f"l{i}_nodes"] = np.asarray(nodes, np.int32)
        d[f"l{i}_cs"] = np.array([cs])
        d[f"l{i}_table"] = np.asarray(table, float)
    return {prefix + k: np.asarray(v) for k, v in d.items()}


def model_from_dict(z, prefix: str = "m_") -> M.Model:
    g = lambda k: z[prefix + k]
    m = M.Model(xyz=g("xyz"), hooke=g("hooke"), sections=g("sections"))
    m.section_defs = [(int(r[0]), float(r[1]), float(r[2])) for r in g("section_defs")]
    m.shell_thickness = g("shell_thickness")
    m.cs_defs = [(tuple(r[:3]), tuple(r[3:])) for r in g("cs_defs")]
    m.cs = g("cs")
    for k in ("elem_type", "elem_mat", "elem_sec", "elem_cs", "elem_ptr", "elem_nodes"):
        setattr(m, k, g(k).astype(np.int32))
    p = g("pretension")
    m.pretension = p if p.size else None
    gr = g("gravity")
    m.gravity = tuple(gr) if gr.size else None
    m.constraints = [(g(f"c{i}_nodes"), int(g(f"c{i}_mask")[0])) for i in range(int(g("n_constraints")[0]))]
    m.nodal_loads = [(g(f"l{i}_nodes"), int(g(f"l{i}_cs")[0]), g(f"l{i}_table")) for i in range(int(g("n_loads")[0]))]
    if prefix + "pipe_sections" in getattr(z, "files", z):
        m.pipe_sections = np.asarray(g("pipe_sections"), float).reshape(-1, 11)
    if prefix + "n_shell_loads" in getattr(z, "files", z):
        m.shell_loads = [(g(f"s{i}_elements"), bool(g(f"s{i}_area_update")[0]), g(f"s{i}_table")) for i in range(int(g("n_shell_loads")[0]))]
    if prefix + "n_follower_loads" in getattr(z, "files", z):
        m.follower_loads = [(g(f"f{i}_nodes"), int(g(f"f{i}_cs")[0]), g(f"f{i}_table")) for i in range(int(g("n_follower_loads")[0]))]
    if prefix + "n_pipe_loads" in getattr(z, "files", z):
        m.pipe_loads = [(g(f"p{i}_elements"), g(f"p{i}_table")) for i in range(int(g("n_pipe_loads")[0]))]
    return m


def capture(oracle, tag: str) -> dict:
    """All four CSR matrices and the three vectors of the oracle's last assembly."""
    out = {}
    for w in ("AA", "AB", "BA", "BB"):
        o, i, v, s = oracle.csr(w)
        out[f"{tag}_{w}_outer"], out[f"{tag}_{w}_inner"], out[f"{tag}_{w}_val"] = o, i, v
        out[f"{tag}_{w}_shape"] = np.array(s)
    pa, ia, pb = oracle.vectors()
    out[f"{tag}_PA"], out[f"{tag}_IA"], out[f"{tag}_PB"] = pa, ia, pb
    return out


def captured_csr(z, tag: str, w: str):
    return (z[f"{tag}_{w}_outer"], z[f"{tag}_{w}_inner"], z[f"{tag}_{w}_val"], tuple(z[f"{tag}_{w}_shape"]))


def nodal_load_contribution(m: M.Model, gls: np.ndarray, disp: np.ndarray, time: float):
    """Host restatement of NodalLoad::Mount (reference NodalLoad.cpp:322-401) for
    numeric tables: returns (triplets per matrix, additions to P_A / P_B).
    This is the 'coexisting host contributor' of SURVEY.md 8b: it stays on the
    host and enters the device matrix through gfa_add_host_triplets."""
    trip = {w: ([], [], []) for w in ("AA", "AB", "BA", "BB")}
    pa, pb = ([], []), ([], [])
    active = gls != 0
    for nodes, cs_id, table in m.nodal_loads:
        table = np.asarray(table, float)
        vals = np.array([np.interp(time, table[:, 0], table[:, k]) for k in range(1, 7)])
        nodes = np.asarray(nodes, int)
        nf = np.array([active[nodes - 1, k].sum() for k in range(6)], float)
        mult = 1.0 / nf
        Q = m.cs[cs_id - 1].reshape(3, 3)      # rows E1,E2,E3 = CoordinateSystem::Q
        for nd in nodes:
            f = Q.T @ (mult[:3] * vals[:3])
            mo = Q.T @ (mult[3:] * vals[3:])
            a = disp[nd - 1, 3:6]
            al = np.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])
            g = 4.0 / (4.0 + al * al)
            A = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            Xi = g * (np.eye(3) + 0.5 * A)
            mo = Xi.T @ mo
            h = g
            h2, h4, h8 = 0.5 * h, -0.25 * h * h, -0.5 * h * h
            skew_t = np.array([[0, -mo[2], mo[1]], [mo[2], 0, -mo[0]], [-mo[1], mo[0], 0]])
            V = np.outer(h8 * mo - h4 * (A @ mo), a) + h2 * skew_t
            for lin in range(3):
                gl = gls[nd - 1, lin]
                if active[nd - 1, lin]:
                    (pa if gl > 0 else pb)[0].append(abs(gl) - 1)
                    (pa if gl > 0 else pb)[1].append(-1.0 * f[lin])
            for lin in range(3):
                gl = gls[nd - 1, lin + 3]
                if active[nd - 1, lin + 3]:
                    (pa if gl > 0 else pb)[0].append(abs(gl) - 1)
                    (pa if gl > 0 else pb)[1].append(-1.0 * mo[lin])
                for col in range(3):
                    gc = gls[nd - 1, col + 3]
                    if not active[nd - 1, col + 3]:
                        continue
                    w = ("AA" if gc > 0 else "AB") if gl > 0 else ("BA" if gc > 0 else "BB")
                    if gl == 0:
                        continue
                    trip[w][0].append(abs(gl) - 1)
                    trip[w][1].append(abs(gc) - 1)
                    trip[w][2].append(-1.0 * V[lin, col])
    return trip, pa, pb


def nodal_follower_load_contribution(m: M.Model, gls: np.ndarray, disp: np.ndarray, copy: np.ndarray, time: float):
    """Host restatement of NodalFollowerLoad::Mount (reference NodalFollowerLoad.cpp:243-325): forces and moments given
    in a coordinate system that rotates with the node -- Q_i from the committed rotation (copy_coordinates[3:6]) times
    the increment's Q / Xi.  Returns (triplets per matrix, additions to P_A, additions to P_B) like
    nodal_load_contribution.  It pushes ALL 36 positions of the node's block (zeros in the translation columns)."""
    trip = {w: ([], [], []) for w in ("AA", "AB", "BA", "BB")}
    pa, pb = ([], []), ([], [])
    active = gls != 0
    I = np.eye(3)
    for nodes, cs_id, table in getattr(m, "follower_loads", []):
        table = np.asarray(table, float)
        vals = np.array([np.interp(time, table[:, 0], table[:, k]) for k in range(1, 7)])
        nodes = np.asarray(nodes, int)
        nf = np.array([active[nodes - 1, k].sum() for k in range(6)], float)
        with np.errstate(divide="ignore"):
            mult = 1.0 / nf
        Qcs = m.cs[cs_id - 1].reshape(3, 3)      # rows E1,E2,E3 = CoordinateSystem::Q
        for nd in nodes:
            a = copy[nd - 1, 3:6]
            A = _skew(a)
            g = 4.0 / (4.0 + a @ a)
            Qi = (I + g * (A + 0.5 * (A @ A))) @ Qcs.T
            a = disp[nd - 1, 3:6]
            A = _skew(a)
            g = 4.0 / (4.0 + a @ a)
            Q = I + g * (A + 0.5 * (A @ A))
            Xi = g * (I + 0.5 * A)
            f = mult[:3] * vals[:3]
            mo = mult[3:] * vals[3:]
            fip = Q @ Qi @ f
            mip = Xi @ Qi @ mo
            K12 = -1.0 * _skew(fip) @ Xi
            K22 = -0.5 * g * (_skew(Qi @ mo) + Xi @ np.outer(Qi @ mo, a))
            q = np.concatenate([fip, mip])
            dq = np.zeros((6, 6))
            dq[:3, 3:] = K12
            dq[3:, 3:] = K22
            for lin in range(6):
                gl = gls[nd - 1, lin]
                if active[nd - 1, lin]:
                    (pa if gl > 0 else pb)[0].append(abs(gl) - 1)
                    (pa if gl > 0 else pb)[1].append(-1.0 * q[lin])
                for col in range(6):
                    gc = gls[nd - 1, col]
                    if not active[nd - 1, col] or gl == 0:
                        continue
                    w = ("AA" if gc > 0 else "AB") if gl > 0 else ("BA" if gc > 0 else "BB")
                    trip[w][0].append(abs(gl) - 1)
                    trip[w][1].append(abs(gc) - 1)
                    trip[w][2].append(-1.0 * dq[lin, col])
    return trip, pa, pb


def host_load_contribution(m: M.Model, gls: np.ndarray, disp: np.ndarray, copy: np.ndarray, time: float):
    """NodalLoad + NodalFollowerLoad of a model, merged (what a host pushes after gfa_assemble)."""
    t1, a1, b1 = nodal_load_contribution(m, gls, disp, time)
    t2, a2, b2 = nodal_follower_load_contribution(m, gls, disp, copy, time)
    trip = {w: tuple(list(t1[w][k]) + list(t2[w][k]) for k in range(3)) for w in t1}
    return trip, (list(a1[0]) + list(a2[0]), list(a1[1]) + list(a2[1])), (list(b1[0]) + list(b2[0]), list(b1[1]) + list(b2[1]))


def assert_results_parity(ref, got, what):
    """Gauss-point results (layout of gfa_gauss_point_results): column 0 = strain energy, then
    3-vectors (Shell_1: 8 per point) or 6-vectors (Beam_1: epsilon_r, sigma_r); every group is
    compared on the scale of its largest entry over the model."""
    ref, got = np.asarray(ref), np.asarray(got)
    assert ref.shape == got.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    # Shell_1 evaluates its specific strain energy through a cancelling expression,
    # 0.5 lambda (0.5 (J^2 - 1) - log J) + 0.5 mu (I1 - 3 - 2 log J) with J, I1/3 = 1 + O(strain)
    # (Shell_1.cpp:1150-1161): its condition number is 1/strain^2 and one ulp of log() moves it by
    # 1e-11 relative at strains of 1e-5, so it is compared to 1e-9; Beam_1's 0.5 sigma.epsilon to 1e-12.
    energy_tol = 1e-9 if ref.shape[1] == 73 else TOL
    assert_parity(ref[:, 0], got[:, 0], f"{what} strain energy", float(np.abs(ref[:, 0]).max()), tol=energy_tol)
    width = 3 if ref.shape[1] == 73 else 6
    per_point = 8 if ref.shape[1] == 73 else 2
    body_r = ref[:, 1:].reshape(ref.shape[0], -1, per_point, width)
    body_g = got[:, 1:].reshape(ref.shape[0], -1, per_point, width)
    for k in range(per_point):
        scale = float(np.abs(body_r[:, :, k]).max())
        assert_parity(body_r[:, :, k].ravel(), body_g[:, :, k].ravel(), f"{what} group {k}", max(scale, 1e-300))


# ---- Newmark dynamics: one scenario driven identically on every backend -------------------------
# (RefOracle = the reference's own Dynamic object, PortOracle = the CPU restatement, Assembler = the
# C-ABI library).  Sequence of Dynamic::Solve (Dynamic.cpp:303-340, Solution.cpp:426-454):
#   time step 1: UpdateDyn, assembly with MountDamping(true); UpdateDyn, assembly with MountDamping(false);
#                SaveConfiguration (alpha_i, copy_vel/copy_accel)
#   time step 2: UpdateDyn, assembly with MountDamping(false)  -> non-zero alpha_i, stored rayleigh_damping
def newmark_coefficients(time_step: float, beta_new: float = 0.3, gamma_new: float = 0.5) -> np.ndarray:
    """Dynamic::CalculateNewmarkCoeff (Dynamic.cpp:582-590): a1..a6"""
    dt, b, g = float(time_step), float(beta_new), float(gamma_new)
    return np.array([1.0 / (dt * dt * b), 1.0 / (dt * b), 1.0 / (2.0 * b) - 1.0, g / (dt * b), 1.0 - g / b,
                     dt * (1.0 - g / (2.0 * b))])


def dynamic_scenario(m: M.Model, disp: np.ndarray, seed: int, time_step=0.01, rayleigh=(0.7, 1.0e-4)) -> dict:
    rng = np.random.default_rng(seed)
    n = m.n_nodes
    return {
        "dyn_dt": np.array([time_step]), "dyn_rayleigh": np.array(rayleigh, float),
        "dyn_copy_vel": rng.uniform(-1.0, 1.0, (n, 6)) * np.array([1, 1, 1, 0.5, 0.5, 0.5]),
        "dyn_copy_accel": rng.uniform(-10.0, 10.0, (n, 6)),
        "dyn_disp": np.stack([disp, 0.6 * disp, -0.4 * disp]),
    }


DYN_STEPS = (("s1", 0, True, False), ("s2", 1, False, True), ("s3", 2, False, False))   # tag, disp index, update_rayleigh, commit after


def run_dynamic(backend, m: M.Model, scen: dict, on_step):
    """Drives `backend` through the scenario; on_step(tag, backend) is called after every assembly and
    on_step(tag + '_commit', backend) after the commit."""
    dt = float(scen["dyn_dt"][0])
    ra, rb = [float(v) for v in scen["dyn_rayleigh"]]
    if hasattr(backend, "dynamic_begin"):         # the reference computes its own coefficients
        backend.dynamic_begin(0.3, 0.5, ra, rb, 0)
        a = backend.newmark(dt)
        assert np.array_equal(a, newmark_coefficients(dt))
    else:
        backend.set_dynamic(newmark_coefficients(dt), ra, rb)
    zeros = np.zeros((m.n_nodes, 6))
    backend.set_kinematics(zeros, zeros, scen["dyn_copy_vel"], scen["dyn_copy_accel"])
    for tag, k, update, commit in DYN_STEPS:
        d = scen["dyn_disp"][k]
        backend.update_dyn(d)
        backend.assemble_dynamic(d, update)
        on_step(tag, backend)
        if commit:
            backend.commit()
            on_step(tag + "_commit", backend)


def capture_dynamic(z: dict, elements):
    """on_step callback that records everything into z (fixture generation)."""
    def cb(tag, b):
        if tag.endswith("_commit"):
            for e in elements:
                z[f"{tag}_alpha_i{e}"] = b.alpha_i(e)
            z[f"{tag}_kin"] = np.stack(b.kinematics())
            return
        z.update(capture(b, tag))
        z[f"{tag}_kin"] = np.stack(b.kinematics())
        for e in elements:
            K, P = b.element(e)[:2]
            z[f"{tag}_elem{e}_K"], z[f"{tag}_elem{e}_P"] = K, P
    return cb


def check_dynamic(z, elements, what: str):
    """on_step callback that compares a backend with a recorded fixture / another capture."""
    def cb(tag, b):
        if tag.endswith("_commit"):
            for e in elements:
                assert_parity(z[f"{tag}_alpha_i{e}"], b.alpha_i(e), f"{what} {tag} alpha_i of element {e}")
            for r, g, key in zip(z[f"{tag}_kin"], b.kinematics(), ("vel", "accel", "copy_vel", "copy_accel")):
                assert_parity(r, g, f"{what} {tag} {key}")
            return
        assert_system_parity(lambda w: captured_csr(z, tag, w), b.csr, f"{what} {tag}")
        for v, key in zip(b.vectors(), ("PA", "IA", "PB")):
            assert_parity(z[f"{tag}_{key}"], v, f"{what} {tag} {key}")
        for r, g, key in zip(z[f"{tag}_kin"], b.kinematics(), ("vel", "accel", "copy_vel", "copy_accel")):
            assert_parity(r, g, f"{what} {tag} {key}")
        for e in elements:
            K, P = b.element(e)[:2]
            assert_parity(z[f"{tag}_elem{e}_K"], K, f"{what} {tag} element {e} K", block_scale(z[f"{tag}_elem{e}_K"]))
            assert_parity(z[f"{tag}_elem{e}_P"], P, f"{what} {tag} element {e} P")
    return cb


# ---- ShellLoad: follower pressure on Shell_1 elements, a coexisting HOST contributor ------------------
_COWPER = np.array([
    [0.816847572980459, 0.091576213509771, 0.091576213509771, 0.109951743655322],
    [0.091576213509771, 0.816847572980459, 0.091576213509771, 0.109951743655322],
    [0.091576213509771, 0.091576213509771, 0.816847572980459, 0.109951743655322],
    [0.108103018168070, 0.445948490915965, 0.445948490915965, 0.223381589678011],
    [0.445948490915965, 0.108103018168070, 0.445948490915965, 0.223381589678011],
    [0.445948490915965, 0.445948490915965, 0.108103018168070, 0.223381589678011]])


def _skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def shell_load_contribution(m: M.Model, gls: np.ndarray, disp: np.ndarray, copy: np.ndarray, time: float):
    """Host restatement of ShellLoad::Mount -> Shell_1::MountShellSpecialLoads (reference ShellLoad.cpp:133-148,
    Shell_1.cpp:1392-1467): follower pressure integrated with the 6-point rule on the current configuration
    (copy_coordinates + displacements), its non-symmetric load stiffness on the u-u blocks and its force on
    P_loading.  Returns (triplets per matrix, additions to P_A and I_A (the same), additions to P_B): what the
    host pushes through gfa_add_host_triplets / gfa_add_host_vector after gfa_assemble.  `copy` is
    Node::copy_coordinates [n,6] (gfa_copy_coordinates)."""
    trip = {w: ([], [], []) for w in ("AA", "AB", "BA", "BB")}
    pa, pb = ([], []), ([], [])
    conn = m.elem_nodes
    for elements, area_update, table in m.shell_loads:
        table = np.asarray(table, float)
        pressure = float(np.interp(time, table[:, 0], table[:, 1]))
        for e1 in np.asarray(elements, int):
            e = e1 - 1
            nd = conn[m.elem_ptr[e]:m.elem_ptr[e + 1]].astype(int) - 1
            x = m.xyz[nd]
            nvec = np.cross(x[1] - x[0], x[2] - x[0])
            A = 0.5 * np.linalg.norm(nvec)
            e3 = nvec / np.linalg.norm(nvec)
            eg = np.array([1.0, 0.0, 0.0])
            if abs(eg @ e3) >= 1.0 - 1e-4:
                eg = np.array([0.0, 1.0, 0.0])
            e1r = eg - (eg @ e3) * e3
            e1r = e1r / np.linalg.norm(e1r)
            e2r = np.cross(e3, e1r)
            Lx = 0.5 * np.array([(x[1] - x[2]) @ e2r, (x[2] - x[0]) @ e2r, (x[0] - x[1]) @ e2r]) / A
            Ly = 0.5 * np.array([(x[2] - x[1]) @ e1r, (x[0] - x[2]) @ e1r, (x[1] - x[0]) @ e1r]) / A
            u = copy[nd, :3] - x + disp[nd, :3]                       # pu_ip, global axes
            K = np.zeros((18, 18))
            P = np.zeros(18)
            for L1, L2, L3, wq in _COWPER:
                L = np.array([L1, L2, L3])
                w4 = A * wq
                N = np.array([(2 * L1 - 1) * L1, (2 * L2 - 1) * L2, (2 * L3 - 1) * L3, 4 * L1 * L2, 4 * L2 * L3, 4 * L3 * L1])

                def grad(D):
                    return np.array([4 * D[0] * L[0] - D[0], 4 * D[1] * L[1] - D[1], 4 * D[2] * L[2] - D[2],
                                     4 * D[0] * L[1] + 4 * L[0] * D[1], 4 * D[1] * L[2] + 4 * L[1] * D[2],
                                     4 * D[2] * L[0] + 4 * L[2] * D[0]])
                N1, N2 = grad(Lx), grad(Ly)
                t1 = e1r + N1 @ u
                t2 = e2r + N2 @ u
                c = np.cross(t1, t2)
                jac = np.linalg.norm(c)
                n = c / jac
                q = -1.0 * pressure * n
                if not area_update:
                    proj = (1.0 / jac) * (np.eye(3) - np.outer(n, n))
                    scale = w4
                else:
                    proj = np.eye(3)
                    scale = w4 * jac
                for a in range(6):
                    P[3 * a:3 * a + 3] -= scale * N[a] * q
                    for b in range(6):
                        Kp = proj @ (_skew(t1) * N2[b] - _skew(t2) * N1[b])
                        K[3 * a:3 * a + 3, 3 * b:3 * b + 3] += w4 * 1.0 * pressure * N[a] * Kp
            g = gls[nd, :3].reshape(-1)
            for i in range(18):
                gi = g[i]
                if gi > 0:
                    pa[0].append(gi - 1); pa[1].append(P[i])
                elif gi < 0:
                    pb[0].append(-gi - 1); pb[1].append(P[i])
                for j in range(18):
                    gj = g[j]
                    if gi > 0 and gj > 0:
                        w, r, cc = "AA", gi - 1, gj - 1
                    elif gi < 0 and gj < 0:
                        w, r, cc = "BB", -gi - 1, -gj - 1
                    elif gi > 0 and gj < 0:
                        w, r, cc = "AB", gi - 1, -gj - 1
                    elif gi < 0 and gj > 0:
                        w, r, cc = "BA", -gi - 1, gj - 1
                    else:
                        continue
                    trip[w][0].append(r); trip[w][1].append(cc); trip[w][2].append(K[i, j])
    return trip, pa, pb


# ---- shell meshes the reference ships (tests/golden/tutorial05_shells.npz, tutorial02_shells.npz) ----------
def check_shipped_shell_mesh(z, backend, what: str):
    """Replay the fixture's two iterations (with the commit in between) on `backend` (the oracle port or the CUDA
    Assembler) and compare with what the reference's own sources produced: every CSR value when the fixture
    holds them (tutorial05), otherwise the pattern digest, the vectors, AA's row sums, diagonal and 20 000
    sampled values (tutorial02)."""
    import hashlib
    full = "it1_AA_val" in z
    for tag, commit_after in (("it1", True), ("it2", False)):
        backend.assemble(z[f"{tag}_disp"])
        if full:
            assert_system_parity(lambda w: captured_csr(z, tag, w), backend.csr, f"{what} {tag}")
        else:
            o, i, v, s = backend.csr("AA")
            assert tuple(s) == tuple(z[f"{tag}_AA_shape"])
            digest = np.frombuffer(hashlib.sha256(o.tobytes() + i.tobytes()).digest(), np.uint8)
            assert (digest == z[f"{tag}_AA_pattern_sha256"]).all(), f"{what} {tag}: AA pattern differs from the reference's"
            assert [len(backend.csr(w)[2]) for w in ("AA", "AB", "BA", "BB")] == list(z[f"{tag}_nnz"])
            diag = z[f"{tag}_AA_diag"]
            rows = np.repeat(np.arange(s[0]), np.diff(o))
            pick = z[f"{tag}_AA_sample_idx"]
            scale = np.sqrt(diag[rows[pick]] * diag[i[pick]])
            assert_parity(z[f"{tag}_AA_sample_val"], v[pick], f"{what} {tag} sampled AA values", scale)
            # row sums: every value of a row enters; errors of 1e-12 relative to the row's absolute sum
            rs = np.bincount(rows, weights=v, minlength=s[0])
            assert_parity(z[f"{tag}_AA_rowsum"], rs, f"{what} {tag} AA row sums", 10.0 * z[f"{tag}_AA_rowabs"])
            assert_parity(diag, csr_diag((o, i, v, s)), f"{what} {tag} AA diagonal")
        for v, key in zip(backend.vectors(), ("PA", "IA", "PB")):
            assert_parity(z[f"{tag}_{key}"], v, f"{what} {tag} {key}")
        if commit_after:
            backend.commit()
