"""Status-file half of SURVEY.md 8 f4, without a GPU: the oracle restatement of get_total_circ / get_total_impulse against the
reference's own Points<float> (tests/golden/status.npz, minted by tests/golden/make_golden.py through oracle/_ref), and the
C-ABI status writer (host code in libo3d_cuda.so: csrc/status_writer.h) against the bytes the reference's StatusFile wrote."""
import os

import numpy as np

from conftest import golden
from omega3d_b200 import status as S


def test_restatement_totals_bit_identical_to_reference(restate):
    g = golden("status.npz")
    for name in ("ring", "leap", "cloud"):
        c, i = restate.totals(g[name + "_x"], g[name + "_s"])
        assert np.array_equal(c.view(np.uint32), g[name + "_circ"].view(np.uint32)), name
        assert np.array_equal(i.view(np.uint32), g[name + "_imp"].view(np.uint32)), name


def test_reference_totals_reproducible(reference_lib):
    g = golden("status.npz")
    c, i = reference_lib.totals(g["ring_x"], g["ring_s"])
    assert np.array_equal(c, g["ring_circ"]) and np.array_equal(i, g["ring_imp"])


def write_lines(path, fmt, vals, nv, reset):
    sf = S.StatusFile()
    assert not sf.is_active()
    sf.append_value("ignored", 1.0)          # an unarmed StatusFile collects nothing (the reference's writes nothing)
    sf.set_filename(path, fmt)
    assert sf.is_active() and sf.get_filename() == path
    for k in range(len(nv)):
        if reset[k]:
            sf.reset_sim()
        v = vals[k]
        sf.append_value("time", float(v[0]))
        sf.append_value("Nv", int(nv[k]))
        for name, x in zip(("gx", "gy", "gz", "fx", "fy", "fz"), v[1:]):
            sf.append_value(name, float(x))
        sf.write_line()
    sf.close()
    with open(path, "rb") as f:
        return f.read()


def test_status_writer_byte_identical_to_reference(tmp_path):
    """Two data sets in one file (reset_sim between them), both formats, %g corner cases (-0, 1e-05, 1.23457e+06, large ints),
    including the reference's habit of never clearing the name list (the second header repeats every name once per line
    written so far)."""
    g = golden("status.npz")
    for fmt, tag in ((S.dat, "dat"), (S.csv, "csv")):
        mine = write_lines(str(tmp_path / ("status." + tag)), fmt, g["status_vals"], g["status_nv"], g["status_reset"])
        assert mine == bytes(g["status_" + tag]), tag


def test_status_writer_appends_to_an_existing_file(tmp_path):
    """StatusFile opens in append mode: a second writer object on the same path continues the file with its own header."""
    g = golden("status.npz")
    path = str(tmp_path / "s.dat")
    a = write_lines(path, S.dat, g["status_vals"][:2], g["status_nv"][:2], [0, 0])
    b = write_lines(path, S.dat, g["status_vals"][:1], g["status_nv"][:1], [0])
    assert b.startswith(a) and b[len(a):].startswith(b"# time Nv gx gy gz fx fy fz\n")
    assert os.path.getsize(path) == len(b)


def test_anonymous_columns(tmp_path):
    sf = S.StatusFile()
    sf.set_filename(str(tmp_path / "a.csv"), S.csv)
    sf.append_value(1.5)
    sf.append_value(3)
    sf.write_line()
    sf.close()
    assert open(tmp_path / "a.csv").read() == "float,int\n1.5,3\n"
