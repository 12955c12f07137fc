"""CPU: the arena layouts of gfa_device.h (compiled with g++; the header is shared by host and device code).
Every stored block of an element must have its own 9 doubles, inside the element's region (compact layout) or the
batch's region (batch layout of the classic shell arena), and block (a, b) / its transposed twin must resolve to the
same storage wherever the tangent is symmetric -- the slot map, the gather lists and gfa_element_block all go
through these functions."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include "gfa_device.h"
using namespace gfa;
int main() {
    printf("{\"shell_arena\": %d, \"beam_arena\": %d, \"solid_arena\": %d, \"batch\": %d,\n", SHELL_ARENA, BEAM_ARENA, SOLID_ARENA, SHELL_BATCH);
    printf(" \"shell\": [");
    for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) { bool t; int o = shell_block_offset(a, b, t); printf("%s[%d,%d,%d,%d]", a + b ? "," : "", a, b, o, t ? 1 : 0); }
    printf("],\n \"shell_batch\": [");
    bool first = true;
    for (int l = 0; l < 19; l++) for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) { bool t; long long o = shell_batch_offset(l, a, b, t); printf("%s[%d,%d,%d,%lld,%d]", first ? "" : ",", l, a, b, o, t ? 1 : 0); first = false; }
    printf("],\n \"beam\": [");
    for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) { bool t; int o = beam_block_offset(a, b, t); printf("%s[%d,%d,%d,%d]", a + b ? "," : "", a, b, o, t ? 1 : 0); }
    printf("],\n \"solid\": [");
    for (int a = 0; a < 8; a++) for (int b = 0; b < 8; b++) { bool t; int o = solid_block_offset(a, b, t); printf("%s[%d,%d,%d,%d]", a + b ? "," : "", a, b, o, t ? 1 : 0); }
    printf("]}\n");
    return 0;
}
"""


def _tables(tmp_path):
    src = tmp_path / "layout.cpp"
    src.write_text(SRC)
    exe = tmp_path / "layout"
    subprocess.check_call(["g++", "-std=c++17", "-D__host__=", "-D__device__=", "-D__forceinline__=inline",
                           "-I", os.path.join(ROOT, "giraffe_b200", "csrc"), "-o", str(exe), str(src)])
    return json.loads(subprocess.check_output([str(exe)]))


def _check_compact(rows, n_groups, arena, symmetric_pairs):
    """rows: [a, b, offset, transposed]; stored blocks are the non-transposed ones"""
    stored = {}
    for a, b, off, tr in rows:
        assert 0 <= off and off + 9 <= arena
        if not tr:
            assert off not in stored, f"blocks {stored.get(off)} and {(a, b)} share offset {off}"
            stored[off] = (a, b)
    offs = sorted(stored)
    for x, y in zip(offs, offs[1:]):
        assert y - x >= 9, "stored blocks overlap"
    for a, b, off, tr in rows:
        if tr:       # a transposed block reads the storage of its twin
            assert stored[off] == (b, a)
    assert len(rows) == n_groups * n_groups
    return stored


def test_compact_layouts(tmp_path):
    t = _tables(tmp_path)
    shell = _check_compact(t["shell"], 9, t["shell_arena"], None)
    assert len(shell) == 48                       # upper triangle (45) + the three lower rotation-rotation blocks
    beam = _check_compact(t["beam"], 6, t["beam_arena"], None)
    assert len(beam) == 24
    solid = _check_compact(t["solid"], 8, t["solid_arena"], None)
    assert len(solid) == 36


def test_shell_batch_layout(tmp_path):
    t = _tables(tmp_path)
    batch, arena = t["batch"], t["shell_arena"]
    compact = {(a, b): (off, tr) for a, b, off, tr in t["shell"]}
    owner = {}
    for l, a, b, off, tr in t["shell_batch"]:
        assert tr == compact[(a, b)][1], "the batch layout stores the same blocks as the compact one"
        region = l // batch
        assert region * batch * arena <= off and off + 9 <= (region + 1) * batch * arena, "a block leaves its batch's region"
        key = (l, b, a) if tr else (l, a, b)
        if off in owner:
            assert owner[off] == key, f"{owner[off]} and {key} share offset {off}"
        owner[off] = key
    offs = sorted(owner)
    for x, y in zip(offs, offs[1:]):
        assert y - x >= 9, "stored blocks of the batch layout overlap"
    assert len(owner) == 19 * 48
    # the same stored block of the elements of a batch is contiguous: 72 bytes apart
    by_block = {}
    for off, (l, a, b) in owner.items():
        by_block.setdefault((l // batch, a, b), []).append((l % batch, off))
    for (_, a, b), lst in by_block.items():
        lst.sort()
        for (r0, o0), (r1, o1) in zip(lst, lst[1:]):
            assert o1 - o0 == 9 * (r1 - r0)
