"""Host-side mirror of the reference's convection driver for the part of it that sits on the Biot-Savart path.

``Convection<S,A,I>`` (src/Convection.h, paths relative to /root/reference) advances the vortex particles with a
1-, 2- (Ralston) or 3-stage Runge-Kutta scheme; every stage is ``find_vels`` = ``zero_vels`` -> influence sums ->
``finalize_vels`` (:130-184) followed by ``Points::move`` = advection + vortex stretching (src/Points.h:288-520).
For a particle-only system (no boundaries, no field points - BASELINE configs C1-C3, C5) that whole step runs here
on the device: the particles stay resident in HBM (``DeviceParticles``), the N^2 sums are the same kernels the
influence routines use, and the O(N) steps around them are ``omega3d_b200/csrc/convect.cuh``, which round exactly as
the reference's scalar build does. Steps of unchanged size replay one captured CUDA graph.

Anything this module cannot run on the GPU raises; there is no host arithmetic in here.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_float, c_void_p

import numpy as np

from .influence import (CudaContext, ExecEnv, O3DError, Points, ResultsType, _ptr, _require_cuda, active, default_context,
                        f32, lagrangian, results_t)


def _fs(fs):
    a = (c_double * 3)(*[float(v) for v in fs])
    return a


class DeviceParticles:
    """One vortex-particle collection (the reference's ``Points<float>``, active + lagrangian) resident on the GPU(s)
    of a ``CudaContext`` (include/o3d_cuda.h: ``o3d_particles``)."""

    def __init__(self, ctx: CudaContext = None):
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        h = c_void_p()
        self.ctx.check(self.lib.o3d_cuda_particles_create(self.ctx.h, byref(h)))
        self.h = h
        self.flops = 0.0

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.o3d_cuda_particles_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n(self) -> int:
        return int(self.lib.o3d_cuda_particles_count(self.h))

    def upload(self, x, s, r, elong=None):
        """x, s (3,n); r (n,); elong (n,) or None (= 1, a fresh collection)."""
        x = np.ascontiguousarray(x, f32)
        s = np.ascontiguousarray(s, f32)
        n = x.shape[1]
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, f32), (n,)), f32)
        e = None if elong is None else np.ascontiguousarray(elong, f32)
        self.ctx.check(self.lib.o3d_cuda_particles_upload(self.ctx.h, self.h, n, _ptr(x[0]), _ptr(x[1]), _ptr(x[2]), _ptr(s[0]),
                                                          _ptr(s[1]), _ptr(s[2]), _ptr(r), _ptr(e)))
        return self

    def download(self, want=("x", "s", "r", "elong", "u", "ug")):
        """Returns a dict of host arrays: x, s, u (3,n); ug (9,n); r, elong (n,)."""
        n = self.n
        out = {}
        if "x" in want: out["x"] = np.empty((3, n), f32)
        if "s" in want: out["s"] = np.empty((3, n), f32)
        if "r" in want: out["r"] = np.empty(n, f32)
        if "elong" in want: out["elong"] = np.empty(n, f32)
        if "u" in want: out["u"] = np.empty((3, n), f32)
        if "ug" in want: out["ug"] = np.empty((9, n), f32)

        def rows(key, k):
            a = out.get(key)
            return [None] * k if a is None else [_ptr(a[i]) for i in range(k)]

        gp = None
        if "ug" in out:
            keep = (c_void_p * 9)(*[out["ug"][k].ctypes.data for k in range(9)])
            gp = ctypes.cast(keep, c_void_p)
        self.ctx.check(self.lib.o3d_cuda_particles_download(self.ctx.h, self.h, *rows("x", 3), *rows("s", 3), _ptr(out.get("r")),
                                                            _ptr(out.get("elong")), *rows("u", 3), gp))
        return out

    def find_vels(self, fs=(0.0, 0.0, 0.0), results: results_t = results_t.velandgrad):
        """Convection::find_vels(fs, vort, {}, vort) for this collection (src/Convection.h:130-184)."""
        fl = c_double()
        self.ctx.check(self.lib.o3d_cuda_particles_find_vels(self.ctx.h, self.h, _fs(fs), int(ResultsType(results).compute_grad()),
                                                             byref(fl)))
        self.flops = fl.value
        return fl.value

    def advect(self, order: int, time: float, dt: float, fs=(0.0, 0.0, 0.0), nsteps: int = 1):
        """`nsteps` x Convection::advect (order 1: :232-262, 2: Ralston :349-425, 3: :431-556), all on the device."""
        fl = c_double()
        self.ctx.check(self.lib.o3d_cuda_particles_advect(self.ctx.h, self.h, int(order), float(time), float(dt), _fs(fs),
                                                          int(nsteps), byref(fl)))
        self.flops = fl.value
        return fl.value

    def graph_active(self) -> bool:
        return bool(self.lib.o3d_cuda_particles_graph_active(self.h))

    def stats(self):
        """(ElementBase::get_max_str, Points::get_max_elong)"""
        a, b = c_float(), c_float()
        self.ctx.check(self.lib.o3d_cuda_particles_stats(self.ctx.h, self.h, byref(a), byref(b)))
        return a.value, b.value


class Convection:
    """src/Convection.h:42-123 for systems made of vortex-particle collections only.

    ``advect(time, dt, fs, ips, vort, bdry, fldpt)`` has the reference's signature (:208-228); ``vort`` is a list with
    ONE active lagrangian ``Points`` (what every particle-only example holds), ``bdry`` and ``fldpt`` must be empty -
    boundaries bring the BEM solve, which stays the reference's host code and calls the influence routines of
    ``omega3d_b200.influence`` instead."""

    def __init__(self, order: int = 2, env: ExecEnv = None, ctx: CudaContext = None):
        assert 0 < order < 4, "Convection integrator orders over 3 unsupported"   # src/Convection.h:218
        self.convection_order = order
        self.conv_env = env or ExecEnv()
        self.ctx = ctx
        self._dev = None

    def _resident(self, p: Points) -> DeviceParticles:
        _require_cuda(self.conv_env)
        if p.E != active or p.M != lagrangian:
            raise O3DError("Convection: only active lagrangian particle collections move on the device")
        if self._dev is None:
            self._dev = DeviceParticles(self.ctx)
        if getattr(p, "elong", None) is None:
            p.elong = np.ones(p.n, f32)
        self._dev.upload(p.x, p.s, p.r, p.elong)
        return self._dev

    def _store(self, p: Points, want):
        out = self._dev.download(want)
        for k, v in out.items():
            if k == "ug":
                p.ug[:] = v
            else:
                getattr(p, k)[...] = v

    def find_vels(self, fs, vort, bdry, targets, results: results_t = results_t.velandgrad, force: bool = False):
        """src/Convection.h:130-184 for targets is vort (a particle system on itself)."""
        if bdry or len(vort) != 1 or len(targets) != 1 or targets[0] is not vort[0]:
            raise O3DError("Convection.find_vels on the device handles one particle collection acting on itself")
        d = self._resident(vort[0])
        d.find_vels(fs, results)
        self._store(vort[0], ("u", "ug") if ResultsType(results).compute_grad() else ("u",))

    def advect(self, time, dt, fs, ips, vort, bdry=(), fldpt=(), bem=None, nsteps: int = 1):
        if bdry or fldpt or len(vort) != 1:
            raise O3DError("Convection.advect on the device handles particle-only systems (one collection, no boundaries, "
                           "no field points)")
        d = self._resident(vort[0])
        d.advect(self.convection_order, time, dt, fs, nsteps)
        self._store(vort[0], ("x", "s", "elong", "u", "ug"))
        return d.flops
