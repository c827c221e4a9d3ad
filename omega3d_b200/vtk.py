"""Field output for the device path: the reference's ``Points<S>::write_vtk`` file, byte for byte.

``write_vtk(points, index, frameno, time)`` mirrors src/Points.h:851-1039 (paths relative to /root/reference): file name
``part_<index>_<frameno>.vtu`` (``fldpt_`` for inert points), one vertex cell per particle, circulation / radius /
velocity as base64 DataArrays. The bytes come from ``omega3d_b200/csrc/vtu_writer.h`` through the C ABI; pure host I/O.
"""
from __future__ import annotations

import os

from . import _lib
from .influence import O3DError, Points, _ptr, inert


def vtk_name(points_are_inert: bool, index: int, frameno: int) -> str:
    """src/Points.h:857-868"""
    return f"{'fldpt_' if points_are_inert else 'part_'}{index:02d}_{frameno:05d}.vtu"


def write_vtk(p: Points, index: int, frameno: int, time: float, directory: str = ".") -> str:
    assert p.n > 0, "Inside write_vtk with no points"          # src/Points.h:852
    path = os.path.join(directory, vtk_name(p.E == inert, index, frameno))
    s = [None] * 3 if p.E == inert else [_ptr(p.s[k]) for k in range(3)]
    r = None if p.E == inert else _ptr(p.r)
    rc = _lib.load().o3d_cuda_write_points_vtu(path.encode(), p.n, _ptr(p.x[0]), _ptr(p.x[1]), _ptr(p.x[2]), *s, r, _ptr(p.u[0]),
                                               _ptr(p.u[1]), _ptr(p.u[2]), float(time))
    if rc != 0:
        raise O3DError(f"cannot write {path}")
    return path


def write_resident_vtk(dev, index: int, frameno: int, time: float, directory: str = ".") -> str:
    """Same file straight from a ``convection.DeviceParticles`` collection (one download, then the writer)."""
    path = os.path.join(directory, vtk_name(False, index, frameno))
    dev.ctx.check(dev.lib.o3d_cuda_particles_write_vtu(dev.ctx.h, dev.h, path.encode(), float(time)))
    return path
