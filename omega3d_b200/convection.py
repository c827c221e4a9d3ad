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
import math
from ctypes import POINTER, byref, c_double, c_float, c_int, c_int64, c_void_p

import numpy as np

from .influence import (CudaContext, ExecEnv, O3DError, Points, ResultsType, _ptr, _require_cuda, active, default_context,
                        f32, lagrangian, results_t)


def _fs(fs):
    a = (c_double * 3)(*[float(v) for v in fs])
    return a


# include/o3d_cuda.h: o3d_bem_solve_fn
SOLVE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int64, POINTER(c_float), POINTER(c_float), POINTER(c_float), POINTER(c_float),
                            POINTER(c_float), POINTER(c_int))
CLEAR_INNER_CUTOFF = 0.5 / math.sqrt(2.0 * math.pi)     # every clear_inner_layer call of src/Convection.h


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
        rc = self.lib.o3d_cuda_particles_advect(self.ctx.h, self.h, int(order), float(time), float(dt), _fs(fs), int(nsteps), byref(fl))
        if rc != 0 and getattr(self, "_cb_error", None) is not None:
            err, self._cb_error = self._cb_error, None
            raise err
        self.ctx.check(rc)
        self.flops = fl.value
        return fl.value

    # ---- a static body attached to the collection (include/o3d_cuda.h: o3d_cuda_particles_set_body) ----
    def set_body(self, surf, ips: float, solve=None, cutoff_mult: float = CLEAR_INNER_CUTOFF):
        """surf: influence.Surfaces. solve(pu (3,np) float32 - the raw panel-centre sums of the state: zeroed, then
        points_affect_panels) is the rest of solve_bem (finalize_vels, right-hand side, solve, set_str - host code as in the
        reference) and must return (ts (3,np), sss (np,) or None): the panels' total vortex and source strengths. None:
        strengths stay what set_body_strengths last set."""
        self._cb_error = None
        cb = None
        if solve is not None:
            def _cb(user, np_, pu, tsx, tsy, tsz, sss, have_source):
                try:
                    v = np.ctypeslib.as_array(pu, shape=(3 * np_,)).reshape(3, np_).copy()
                    ts, src = solve(v)
                    for dst, row in zip((tsx, tsy, tsz), ts):
                        np.ctypeslib.as_array(dst, shape=(np_,))[:] = row
                    if src is not None:
                        np.ctypeslib.as_array(sss, shape=(np_,))[:] = src
                        have_source[0] = 1
                    return 0
                except Exception as e:          # noqa: BLE001 - must not unwind through the C frames
                    self._cb_error = e
                    return 1
            cb = SOLVE_FN(_cb)
        self._cb = cb                            # keep the trampoline alive as long as the body is attached
        self._body = surf
        self.ctx.check(self.lib.o3d_cuda_particles_set_body(
            self.ctx.h, self.h, surf.x.shape[1], _ptr(surf.x[0]), _ptr(surf.x[1]), _ptr(surf.x[2]), surf.np_, _ptr(surf.idx),
            _ptr(surf.area), _ptr(surf.nrm), c_float(float(f32(cutoff_mult))), c_float(float(f32(ips))),
            ctypes.cast(cb, c_void_p) if cb is not None else None, None))
        return self

    def clear_body(self):
        self.ctx.check(self.lib.o3d_cuda_particles_clear_body(self.ctx.h, self.h))
        self._cb = self._body = None

    def set_body_strengths(self, ts, sss=None):
        ts = np.ascontiguousarray(ts, f32)
        sss = None if sss is None else np.ascontiguousarray(sss, f32)
        self.ctx.check(self.lib.o3d_cuda_particles_set_body_strengths(self.ctx.h, self.h, _ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]), _ptr(sss)))

    def body_vels(self):
        """Raw panel-centre sums induced by the resident particles: what solve_bem holds after zero_vels + points_affect_panels
        (src/BEMHelper.h:83-94), before finalize_vels. Returns (3,np) float32."""
        pu = np.zeros((3, self._body.np_), f32)
        self.ctx.check(self.lib.o3d_cuda_particles_body_vels(self.ctx.h, self.h, _ptr(pu[0]), _ptr(pu[1]), _ptr(pu[2])))
        return pu

    def clear_inner(self) -> int:
        n = c_int64()
        self.ctx.check(self.lib.o3d_cuda_particles_clear_inner(self.ctx.h, self.h, byref(n)))
        return n.value

    def body_counters(self):
        """(particles pushed out by the clear-inner passes, BEM solves requested) of the last advect call."""
        m, k = c_int64(), c_int()
        self.lib.o3d_cuda_particles_body_counters(self.h, byref(m), byref(k))
        return m.value, k.value

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
        """src/Convection.h:208-228. `bdry` may hold ONE static reactive Surfaces together with `bem`, an object with
        set_rhs / solve / getStrengths over that surface's self-influence system (bem.DenseBEM, bem.BEM): every derivative
        evaluation then solves the BEM for the state first (find_derivs) - right-hand side formed on the device, the solve
        here on the host as in the reference - and every move is followed by clear_inner_layer. The particles stay resident
        throughout."""
        from .bem import solve_bem_for
        if fldpt or len(vort) != 1 or len(bdry) > 1:
            raise O3DError("Convection.advect on the device handles one particle collection, at most one static body, no field points")
        d = self._resident(vort[0])
        if bdry:
            surf = bdry[0]
            if bem is None:
                raise O3DError("Convection.advect: a boundary needs its BEM system")

            def solve(pu):
                surf.pu[:] = pu
                surf.finalize_vels(fs)
                solve_bem_for(surf, bem)
                return surf.ts, surf.ps[2]
            d.set_body(surf, ips, solve)
        try:
            d.advect(self.convection_order, time, dt, fs, nsteps)
        finally:
            if bdry:
                d.clear_body()
        self._store(vort[0], ("x", "s", "elong", "u", "ug"))
        return d.flops
