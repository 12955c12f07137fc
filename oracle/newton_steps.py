"""TEST INFRASTRUCTURE ONLY (see oracle/README or DESIGN.md section 5): numpy restatement of the
Newton-loop vector steps that sit either side of the assembly path in the reference.

  residual()              Static.cpp:210-217  P_A = -1.0*P_A ; P_A = P_A - 1.0*(K_AB * X_B) with the reference's own
                          row loop (SparseMatrix.cpp:186-190: y_i += a*x in column order, separate multiply and add)
  residual_norms()        ConvergenceCriteria.cpp:200-217, 474-505: max |P_A(GL-1)| over free translational /
                          rotational node DOFs, the FIRST node (node order) that reaches it, NaN flag
  update_displacements()  Solution.cpp:390-402: displacements[j] += x(GL-1) for GL > 0
  increment_norms()       ConvergenceCriteria.cpp:305-340: the same maxima over the increment x and over
                          |displacements| after the update
Pinned against the reference's own code paths through oracle/_ref (tests/test_oracle_vs_ref.py).
"""
from __future__ import annotations

import numpy as np


def residual(P_A: np.ndarray, AB=None, X_B: np.ndarray | None = None) -> np.ndarray:
    out = -1.0 * np.asarray(P_A, float)
    if X_B is not None and AB is not None:
        outer, inner, val = AB[0], AB[1], AB[2]
        for r in range(len(outer) - 1):
            if outer[r + 1] > outer[r]:
                y = 0.0
                for p in range(outer[r], outer[r + 1]):
                    y = y + val[p] * X_B[inner[p]]
                out[r] = out[r] - 1.0 * y
    return out


def _max_first(gls: np.ndarray, v: np.ndarray, cols):
    g = gls[:, cols]
    a = np.where(g > 0, np.abs(v[np.maximum(g, 1) - 1]), -1.0)
    a = np.where(np.isnan(a), -1.0, a)
    per_node = a.max(axis=1)
    m = per_node.max() if per_node.size else -1.0
    if m <= 0.0:
        return 0.0, 0
    return float(m), int(np.argmax(per_node == m)) + 1


def residual_norms(gls: np.ndarray, v: np.ndarray) -> dict:
    gls = np.asarray(gls).reshape(-1, 6)
    f, nf = _max_first(gls, v, [0, 1, 2])
    m, nm = _max_first(gls, v, [3, 4, 5])
    free = gls > 0
    nan = bool(np.isnan(v[gls[free] - 1]).any())
    return dict(max_force=f, max_moment=m, node_force=nf, node_moment=nm, nan_detected=int(nan))


def update_displacements(gls: np.ndarray, disp: np.ndarray, x: np.ndarray) -> np.ndarray:
    gls = np.asarray(gls).reshape(-1, 6)
    out = np.array(disp, float).reshape(-1, 6).copy()
    free = gls > 0
    out[free] = out[free] + x[gls[free] - 1]
    return out


def increment_norms(gls: np.ndarray, x: np.ndarray, disp_after: np.ndarray) -> dict:
    gls = np.asarray(gls).reshape(-1, 6)
    out = residual_norms(gls, x)
    d = np.abs(np.asarray(disp_after).reshape(-1, 6))
    free = gls > 0
    out["max_disp_value"] = float(np.where(free[:, :3], d[:, :3], 0.0).max()) if d.size else 0.0
    out["max_rot_value"] = float(np.where(free[:, 3:], d[:, 3:], 0.0).max()) if d.size else 0.0
    return out
