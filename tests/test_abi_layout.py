"""CPU check of the C-ABI boundary: every struct that graph_marl_b200/_lib.py mirrors with ctypes must have the size and the
field offsets the C compiler gives the declaration in include/graphmarl_b200.h (a drifted mirror would pass garbage
pointers to the kernels -- found on the GPU box only)."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PAIRS = [("gm_routing_desc", "RoutingDesc"), ("gm_routing_io", "RoutingIO"), ("gm_cell_params", "CellParams"),
         ("gm_netmon_params", "NetmonParams"), ("gm_dqn_params", "DqnParams"), ("gm_replay_field", "ReplayField"),
         ("gm_mlp_desc", "MlpDesc"), ("gm_mlp_grads", "MlpGrads"), ("gm_cell_grads", "CellGrads"),
         ("gm_netmon_grads", "NetmonGrads")]


def test_ctypes_mirrors_match_the_header(tmp_path):
    from graph_marl_b200 import _lib

    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "graphmarl_b200.h"', "int main(void) {"]
    for cname, pyname in PAIRS:
        cls = getattr(_lib, pyname)
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    r = subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.fail("the header does not declare a field the ctypes mirror names:\n" + r.stderr[-2000:])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {}
    for ln in out.splitlines():
        s, f, v = ln.split()
        got[(s, f)] = int(v)
    for cname, pyname in PAIRS:
        cls = getattr(_lib, pyname)
        assert C.sizeof(cls) == got[(cname, "size")], (cname, C.sizeof(cls), got[(cname, "size")])
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == got[(cname, fname)], (cname, fname)
