"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/zsg_b200.h declares, and the ctypes binding knows each of them.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zsg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zsg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from zsg_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/zsg_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.load().zsg_abi_version() == 8


def test_row_table_layout_matches_header():
    from zsg_b200 import _lib, geometry
    assert ctypes.sizeof(_lib.RowT) == 16
    t = geometry.conv_rows(2, 6, 5, 8, 3, 3, 16, 2, 1, in_off=1000, out_off=64)
    assert t.shape == (18, 16)
    rows = (_lib.RowT * 18).from_buffer_copy(t.numpy().tobytes())
    r = rows[3 * 3 + 4]            # b=1, p=1, q=1
    assert (r.base, r.y0, r.x0, r.hin, r.win, r.out) == (1000 + 6 * 5 * 8, 1, 1, 6, 5, 64 + (9 + 4) * 16)
    d = geometry.dgrad_rows(1, 6, 5, 8, 3, 3, 16, 3, 2, 1)
    rows = (_lib.RowT * 30).from_buffer_copy(d.numpy().tobytes())
    assert (rows[0].y0, rows[0].x0, rows[0].hin, rows[0].win) == (-1, -1, 3, 3)


def test_missing_library_fails_loudly(monkeypatch):
    from zsg_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libzsg_b200.so")
    with pytest.raises(_lib.ZsgError):
        _lib.load()


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof / offsetof of the parameter blocks as a C compiler sees include/zsg_b200.h (gcc, plain C: the header must
    stay C-clean) against the ctypes mirrors in _lib.py and the numpy table layouts."""
    import subprocess
    from zsg_b200 import _lib, geometry, ops
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "zsg_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(zsg_row_t), sizeof(zsg_conv_params), '
                   'offsetof(zsg_conv_params, x_lo), offsetof(zsg_conv_params, dil), offsetof(zsg_conv_params, stats), '
                   'sizeof(zsg_wgrad_params), offsetof(zsg_wgrad_params, dy_pitch), offsetof(zsg_wgrad_params, dil), '
                   'sizeof(zsg_wtf_desc)); return 0; }\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    C, W = _lib.ConvParams, _lib.WgradParams
    want = [ctypes.sizeof(_lib.RowT), ctypes.sizeof(C), C.x_lo.offset, C.dil.offset, C.stats.offset, ctypes.sizeof(W),
            W.dy_pitch.offset, W.dil.offset, ops.WTF_DESC.itemsize]
    assert got == want, (got, want)
    assert geometry.ROW_DTYPE.itemsize == got[0]
