"""CPU: the `.vtu` writer (csrc/vtu_writer.h through o3d_cuda_write_points_vtu - host I/O, no GPU needed) against the
bytes of the file the reference's Points<float>::write_vtk wrote (tests/golden/vtk.npz)."""
import numpy as np
import pytest

from conftest import golden

from omega3d_b200 import influence as I
from omega3d_b200 import vtk as V


@pytest.mark.parametrize("n", [1, 37])
def test_write_vtk_byte_identical(tmp_path, n):
    g = golden("vtk.npz")
    p = I.Points(g[f"x{n}"], g[f"s{n}"], g[f"r{n}"], I.active, I.lagrangian)
    p.u[:] = g[f"u{n}"]
    path = V.write_vtk(p, 3, 42, float(g[f"time{n}"]), str(tmp_path))
    assert path.endswith("part_03_00042.vtu")
    assert open(path, "rb").read() == g[f"file{n}"].tobytes()


def test_inert_points_file_and_errors(tmp_path):
    g = golden("vtk.npz")
    f = I.Points(g["x37"], e=I.inert, m=I.fixed)
    path = V.write_vtk(f, 0, 7, 1.5, str(tmp_path))
    text = open(path, "rb").read()
    assert path.endswith("fldpt_00_00007.vtu") and b"circulation" not in text and b"radius" not in text and b"velocity" in text
    with pytest.raises(I.O3DError):
        V.write_vtk(f, 0, 7, 1.5, str(tmp_path / "no_such_dir"))
