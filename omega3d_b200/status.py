"""Status-file output for the device path: the reference's ``StatusFile`` (src/StatusFile.h:31-76, src/StatusFile.cpp:38-137;
paths relative to /root/reference) and the per-step line of ``Simulation::dump_stats_to_status`` (src/Simulation.cpp:851-897).

The writer itself is host code in the C-ABI library (``omega3d_b200/csrc/status_writer.h``); the quantities on a line - total
circulation and the impulse-derivative force estimate - are reduced on the device from resident particles
(``o3d_cuda_particles_totals``). Files are byte-identical to the reference's for the same values.
"""
from __future__ import annotations

from ctypes import byref, c_double, c_void_p

from . import _lib
from .influence import O3DError

dat, csv = 0, 1          # StatusFormat (src/StatusFile.h:29)


class StatusFile:
    """Same method names as the reference class; ``set_filename`` arms it, as there."""

    def __init__(self):
        self.lib = _lib.load()
        self.h = None
        self.format = dat
        self.fn = ""

    def is_active(self) -> bool:
        return self.h is not None

    def set_filename(self, fn: str, fmt: int = None):
        assert fn, "Filename is blank"                      # src/StatusFile.cpp:39
        if fmt is not None:
            self.format = fmt
        self.close()
        h = c_void_p()
        if self.lib.o3d_cuda_status_open(fn.encode(), int(self.format), byref(h)) != 0:
            raise O3DError(f"cannot open status file {fn}")
        self.h, self.fn = h, fn

    def get_filename(self) -> str:
        return self.fn

    def reset_sim(self):
        if self.h:
            self.lib.o3d_cuda_status_reset_sim(self.h)

    def append_value(self, *args):
        """append_value(name, value) or append_value(value); ints and floats keep their type, as the reference's overloads."""
        if not self.h:
            return
        name, val = (args[0].encode(), args[1]) if len(args) == 2 else (None, args[0])
        if isinstance(val, (int,)) and not isinstance(val, bool):
            self.lib.o3d_cuda_status_append_int(self.h, name, int(val))
        else:
            self.lib.o3d_cuda_status_append_float(self.h, name, float(val))

    def write_line(self):
        if self.h and self.lib.o3d_cuda_status_write_line(self.h) != 0:
            raise O3DError(f"cannot write {self.fn}")

    def close(self):
        if self.h:
            self.lib.o3d_cuda_status_close(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def totals(dev):
    """(total circulation[3], total impulse[3]) of a ``convection.DeviceParticles`` collection, as Python floats (FP64 sums)."""
    c, i = (c_double * 3)(), (c_double * 3)()
    dev.ctx.check(dev.lib.o3d_cuda_particles_totals(dev.ctx.h, dev.h, c, i))
    return list(c), list(i)


def dump_stats_to_status(dev, sf: StatusFile, time: float, dt: float):
    """One status line for a system that is one resident particle collection (src/Simulation.cpp:851-897)."""
    if sf.is_active():
        dev.ctx.check(dev.lib.o3d_cuda_particles_write_status(dev.ctx.h, dev.h, sf.h, float(time), float(dt)))
