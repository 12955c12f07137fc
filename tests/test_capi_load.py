"""CPU: the C-ABI shared library loads, exports every symbol include/gfa.h
declares, and fails loudly (no CPU fallback) when no CUDA device is visible."""
import ctypes
import os
import re

import numpy as np
import pytest

from giraffe_b200 import capi, meshes as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "gfa.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gfa_[a-z_]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert _header_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    for name in _header_symbols():
        assert hasattr(lib, name), f"libgfa.so does not export {name}"


def test_no_cpu_fallback_without_device():
    """Without a GPU gfa_create must fail with GFA_ENODEVICE, never compute."""
    lib = capi.load_library()
    if lib.gfa_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(capi.GfaError) as ei:
        capi.Assembler(M.beam_line(4))
    assert ei.value.code == -2
    assert "no CUDA device" in str(ei.value)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under giraffe_b200/ may import,
    link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "giraffe_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in text.replace("oracle/_ref", "").lower() or f == "meshes.py", f"{f} mentions the oracle"
    out = os.popen(f"ldd {capi.LIB_PATH}").read()
    assert "oracle" not in out and "giraffe_ref" not in out


def test_dof_numbering_restatement_matches_reference_rule():
    """meshes.number_dofs follows Solution::DOFsActive/SetGlobalDOFs: node-major,
    DOF-minor, free ids from 1 up, fixed ids from -1 down, inactive 0."""
    m = M.shell_plate(2, 1)
    gls, nf, nx = M.number_dofs(m)
    assert gls.shape == (m.n_nodes, 6)
    flat = gls.reshape(-1)
    assert list(flat[flat > 0]) == list(range(1, nf + 1))
    assert list(flat[flat < 0]) == list(range(-1, -nx - 1, -1))
    corners = (2 + 1) * (1 + 1)
    assert (gls[:corners, 3:] == 0).all()          # corner nodes carry no rotations (Shell_1.cpp:57-67)
    assert (gls[corners:, :] != 0).all()
