"""ctypes front end to the parity oracles. TEST INFRASTRUCTURE ONLY.

Two libraries live under ``oracle/_ref/`` (git-ignored build outputs, see ``oracle/Makefile``):

* ``libo3d_oracle.so`` - the plain-C restatement (``biot_oracle.c``); builds anywhere.
* ``libo3d_ref.so`` / ``libo3d_ref_fast.so`` - the REFERENCE's own templates
  (``/root/reference/src/Influence.h``, ``Coefficients.h``) behind ``ref_driver.cpp``; built only in
  the container that has ``/root/reference`` and shipped prebuilt to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module. Nothing under ``omega3d_b200/`` does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_double, c_float, c_int, c_int64, c_long, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"


def build(want_ref: bool = True) -> None:
    """Compile the restatement (always) and the reference build (when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", HERE, "restate"], check=True)
    if want_ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
        if os.path.exists(os.path.join(HERE, "..", "omega3d_b200", "lib", "libo3d_cuda.so")):
            subprocess.run(["make", "-s", "-C", HERE, "dropin"], check=True)


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Restatement:
    """biot_oracle.c - every entry point takes SoA float32 arrays and accumulates in place."""

    def __init__(self):
        path = os.path.join(OUT, "libo3d_oracle.so")
        if not os.path.exists(path):
            build(want_ref=False)
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.o3d_oracle_pts_on_pts.argtypes = [c_int64] + [c_void_p] * 7 + [c_int64] + [c_void_p] * 6
        L.o3d_oracle_pts_on_pts_core.argtypes = [c_int, c_int64] + [c_void_p] * 7 + [c_int64] + [c_void_p] * 6
        L.o3d_oracle_pan_on_pts.argtypes = [c_int64] + [c_void_p] * 7 + [c_int64] + [c_void_p] * 5
        L.o3d_oracle_pts_on_pan.argtypes = [c_int64] + [c_void_p] * 6 + [c_int64] + [c_void_p] * 6
        L.o3d_oracle_pan_on_pan_coeff.argtypes = (
            [c_int64] + [c_void_p] * 7 + [c_int64] + [c_void_p] * 8 + [c_int, c_void_p])
        L.o3d_oracle_rkernel_2vs_0p.argtypes = [c_void_p] * 3 + [c_float, c_void_p]
        L.o3d_oracle_rkernel_2vs_0pg.argtypes = [c_void_p] * 3 + [c_float, c_void_p]
        L.o3d_oracle_set_threads.argtypes = [c_int]
        L.o3d_oracle_finalize_vels.argtypes = [c_int64, c_void_p, c_void_p, c_void_p]
        L.o3d_oracle_move.argtypes = [c_int, c_int64, c_double, c_void_p, c_void_p, c_void_p] + [c_void_p] * 4
        L.o3d_oracle_advect.argtypes = [c_int, c_int, c_double, c_void_p, c_int64] + [c_void_p] * 6
        L.o3d_oracle_set_advect_core.argtypes = [c_int]
        L.o3d_oracle_stats.argtypes = [c_int64] + [c_void_p] * 4
        L.o3d_oracle_closest_pass.argtypes = [c_int, c_int64] + [c_void_p] * 5 + [c_int64, c_void_p, c_float, c_float]
        L.o3d_oracle_closest_pass.restype = c_int64
        L.o3d_oracle_totals.argtypes = [c_int64] + [c_void_p] * 4

    def set_threads(self, n):
        self.lib.o3d_oracle_set_threads(int(n))

    # ---- convection (all arrays float32, rows contiguous: x, s, u (3,n); ug (9,n); r, elong (n,)) ----
    def finalize_vels(self, u, ug, fs):
        fs = np.asarray(fs, np.float64)
        self.lib.o3d_oracle_finalize_vels(u.shape[1], _p(u), _p(ug), _p(fs))

    def move(self, order, dt, wt, us, ugs, x, s, elong, uout=None):
        """Points::move with `order` stages; us/ugs: lists of (3,n) / (9,n)|None arrays; x, s, elong updated in place."""
        n = x.shape[1]
        wt = np.asarray(wt, np.float64)
        pu = (c_void_p * 3)(*[u.ctypes.data if u is not None else None for u in (list(us) + [None] * 3)[:3]])
        pg = (c_void_p * 3)(*[g.ctypes.data if g is not None else None for g in (list(ugs) + [None] * 3)[:3]])
        self.lib.o3d_oracle_move(order, n, float(dt), _p(wt), pu, pg, _p(x), _p(s), _p(elong), _p(uout))

    def advect(self, order, nsteps, dt, fs, x, s, r, elong, core=0):
        """nsteps x Convection::advect on one particle collection; returns (u, ug) as left in the collection."""
        n = x.shape[1]
        fs = np.asarray(fs, np.float64)
        u, ug = np.zeros((3, n), np.float32), np.zeros((9, n), np.float32)
        self.lib.o3d_oracle_set_advect_core(int(core))
        try:
            self.lib.o3d_oracle_advect(order, nsteps, float(dt), _p(fs), n, _p(x), _p(s), _p(r), _p(elong), _p(u), _p(ug))
        finally:
            self.lib.o3d_oracle_set_advect_core(0)
        return u, ug

    def stats(self, s, elong):
        a, b = c_float(), c_float()
        self.lib.o3d_oracle_stats(s.shape[1], _p(s), _p(elong), ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def totals(self, x, s):
        """(get_total_circ, get_total_impulse) of a particle collection: two float32[3]."""
        c, i = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self.lib.o3d_oracle_totals(x.shape[1], _p(_f32(x)), _p(_f32(s)), _p(c), _p(i))
        return c, i

    def reflect(self, nodes, idx, nrm, x):
        """reflect_panp2: nodes (3,nn) SoA, idx (np,3), nrm (3,np); x (3,nt) updated in place. Returns particles moved."""
        return int(self.lib.o3d_oracle_closest_pass(0, idx.shape[0], _p(nodes[0]), _p(nodes[1]), _p(nodes[2]), _p(idx), _p(nrm),
                                                    x.shape[1], _p(x), 0.0, 0.0))

    def clear_inner(self, nodes, idx, nrm, x, cutoff_mult, ips):
        """clear_inner_panp2 with _method 1."""
        return int(self.lib.o3d_oracle_closest_pass(1, idx.shape[0], _p(nodes[0]), _p(nodes[1]), _p(nodes[2]), _p(idx), _p(nrm),
                                                    x.shape[1], _p(x), cutoff_mult, ips))

    def max_threads(self):
        return int(self.lib.o3d_oracle_max_threads())

    def pts_on_pts(self, sx, sr, ss, tx, tr, tu, tug, core=0):
        """sx (3,ns), sr (ns,), ss (3,ns); tx (3,nt), tr (nt,)|None; tu (3,nt) and tug (9,nt)|None in/out.
        core: 0 Winckelmans-Leonard (the shipped build), 1 Rosenhead-Moore, 2 exponential, 3 Vatistas n=2."""
        ns, nt = sx.shape[1], tx.shape[1]
        if core:
            self.lib.o3d_oracle_pts_on_pts_core(
                int(core), ns, _p(sx[0]), _p(sx[1]), _p(sx[2]), _p(sr), _p(ss[0]), _p(ss[1]), _p(ss[2]),
                nt, _p(tx[0]), _p(tx[1]), _p(tx[2]), _p(tr), _p(tu), _p(tug))
            return
        self.lib.o3d_oracle_pts_on_pts(ns, _p(sx[0]), _p(sx[1]), _p(sx[2]), _p(sr), _p(ss[0]), _p(ss[1]), _p(ss[2]),
                                       nt, _p(tx[0]), _p(tx[1]), _p(tx[2]), _p(tr), _p(tu), _p(tug))

    def pan_on_pts(self, nodes, idx, ts, area, sss, tx, tu, tug):
        """nodes (3,nn) SoA, idx (np,3) uint32, ts (3,np), area (np,), sss (np,)|None."""
        np_, nt = idx.shape[0], tx.shape[1]
        self.lib.o3d_oracle_pan_on_pts(np_, _p(nodes[0]), _p(nodes[1]), _p(nodes[2]), _p(idx), _p(ts), _p(area),
                                       _p(sss), nt, _p(tx[0]), _p(tx[1]), _p(tx[2]), _p(tu), _p(tug))

    def pts_on_pan(self, sx, ss, nodes, idx, area, pu):
        ns, np_ = sx.shape[1], idx.shape[0]
        self.lib.o3d_oracle_pts_on_pan(ns, _p(sx[0]), _p(sx[1]), _p(sx[2]), _p(ss[0]), _p(ss[1]), _p(ss[2]),
                                       np_, _p(nodes[0]), _p(nodes[1]), _p(nodes[2]), _p(idx), _p(area), _p(pu))

    def pan_on_pan_coeff(self, snodes, sidx, sb1, sb2, sarea, tnodes, tidx, tb1, tb2, tnrm, tarea, self_block):
        nsp, ntp = sidx.shape[0], tidx.shape[0]
        out = np.zeros(9 * nsp * ntp, np.float32)
        self.lib.o3d_oracle_pan_on_pan_coeff(
            nsp, _p(snodes[0]), _p(snodes[1]), _p(snodes[2]), _p(sidx), _p(sb1), _p(sb2), _p(sarea),
            ntp, _p(tnodes[0]), _p(tnodes[1]), _p(tnodes[2]), _p(tidx), _p(tb1), _p(tb2), _p(tnrm), _p(tarea),
            int(bool(self_block)), _p(out))
        return out

    def kernel(self, name, s7, t):
        n = 12 if name.endswith("g") else 3
        out = np.zeros(n, np.float64)
        getattr(self.lib, "o3d_oracle_kernel_" + name)(_p(_f32(s7)), _p(_f32(t)), _p(out))
        return out

    def rkernel(self, grads, tri9, str4, t3, sa):
        out = np.zeros(12 if grads else 3, np.float64)
        fn = self.lib.o3d_oracle_rkernel_2vs_0pg if grads else self.lib.o3d_oracle_rkernel_2vs_0p
        flops = fn(_p(_f32(tri9)), _p(_f32(str4)), _p(_f32(t3)), c_float(sa), _p(out))
        return out, flops


class Reference:
    """ref_driver.cpp - the reference's real code. ``fast=True`` loads the stock-flags build (timing)."""

    TARG_FIELD, TARG_TRACER, TARG_BLOB = 0, 1, 2

    CORE_BUILDS = {0: "libo3d_ref.so", 1: "libo3d_ref_rm.so", 2: "libo3d_ref_exp.so", 3: "libo3d_ref_v2.so"}

    def __init__(self, fast: bool = False, dropin: bool = False, core: int = 0):
        """dropin=True: the patched, -DUSE_CUDA build of the same driver (oracle/_ref/libo3d_dropin.so); call
        set_accel(4) on it to route the reference's own routines into the CUDA arm.
        core != 0: the build whose src/CoreFunc.h has another core function #defined (oracle/Makefile core_build)."""
        # fast: True / "v3" = -march=x86-64-v3, "v4" = -march=x86-64-v4 (AVX-512 hosts only: check host_has_avx512())
        name = ("libo3d_ref_fast_v4.so" if fast == "v4" else "libo3d_ref_fast.so") if fast else self.CORE_BUILDS[core]
        if dropin:   # dropin="exp": the drop-in build of a reference with the exponential core #defined
            name = "libo3d_dropin_exp.so" if dropin == "exp" else "libo3d_dropin.so"
        path = os.path.join(OUT, name)
        if not os.path.exists(path):
            build(want_ref=True)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (needs /root/reference to build; ships prebuilt to the GPU box)")
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.o3d_ref_pts_on_pts.argtypes = [c_int] + [c_void_p] * 5 + [c_int, c_int] + [c_void_p] * 4 + [c_int] + [c_void_p] * 2
        L.o3d_ref_surface_props.argtypes = [c_int, c_void_p, c_int] + [c_void_p] * 7
        L.o3d_ref_pan_on_pts.argtypes = [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int] + [c_void_p] * 6
        L.o3d_ref_pts_on_pan.argtypes = [c_int] + [c_void_p] * 5 + [c_int, c_void_p, c_int] + [c_void_p] * 3
        L.o3d_ref_pan_on_pan.argtypes = [c_int, c_void_p, c_int, c_void_p, c_void_p] * 2 + [c_void_p]
        L.o3d_ref_pan_on_pan_coeff.argtypes = [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                               c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
        L.o3d_ref_pan_on_pan_coeff.restype = c_long
        L.o3d_ref_rkernel_2vs_0p.argtypes = [c_void_p] * 3 + [c_float, c_void_p]
        L.o3d_ref_rkernel_2vs_0pg.argtypes = [c_void_p] * 3 + [c_float, c_void_p]
        if hasattr(L, "o3d_ref_advect"):
            L.o3d_ref_finalize_vels.argtypes = [c_int, c_void_p, c_void_p, c_void_p]
            L.o3d_ref_move.argtypes = [c_int, c_int, c_double, c_void_p] + [c_void_p] * 10
            L.o3d_ref_advect.argtypes = [c_int, c_int, c_double, c_void_p, c_int] + [c_void_p] * 6
            L.o3d_ref_stats.argtypes = [c_int] + [c_void_p] * 4
        if hasattr(L, "o3d_ref_reflect"):
            L.o3d_ref_reflect.argtypes = [c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]
            L.o3d_ref_reflect.restype = c_long
            L.o3d_ref_clear_inner.argtypes = [c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float, c_float]
            L.o3d_ref_clear_inner.restype = c_long
        if hasattr(L, "o3d_ref_advect_body"):
            L.o3d_ref_advect_body.argtypes = ([c_int, c_double, c_void_p, c_float, c_int] + [c_void_p] * 4 + [c_int, c_void_p, c_int, c_void_p,
                                               c_void_p, c_void_p])
            L.o3d_ref_advect_body.restype = c_long
        if hasattr(L, "o3d_ref_totals"):
            L.o3d_ref_totals.argtypes = [c_int] + [c_void_p] * 4
            L.o3d_ref_status_lines.argtypes = [ctypes.c_char_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
        if hasattr(L, "o3d_ref_write_vtk"):
            L.o3d_ref_write_vtk.argtypes = [c_int] + [c_void_p] * 4 + [c_int, c_int, c_double, ctypes.c_char_p]
        if hasattr(L, "o3d_ref_has_features") and L.o3d_ref_has_features():
            L.o3d_ref_singular_ring.argtypes = [c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p, c_long]
            L.o3d_ref_singular_ring.restype = c_long
            L.o3d_ref_thick_ring.argtypes = [c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_long]
            L.o3d_ref_thick_ring.restype = c_long

    def set_threads(self, n):
        self.lib.o3d_ref_set_threads(int(n))

    # ---- convection through the reference's own Points<float> methods ----
    def finalize_vels(self, u, ug, fs):
        fs = np.asarray(fs, np.float64)
        self.lib.o3d_ref_finalize_vels(u.shape[1], _p(u), _p(ug), _p(fs))

    def move(self, order, dt, wt, us, ugs, x, s, elong, uout=None):
        n = x.shape[1]
        wt = np.asarray(wt, np.float64)
        us = (list(us) + [None] * 3)[:3]
        ugs = (list(ugs) + [None] * 3)[:3]
        self.lib.o3d_ref_move(order, n, float(dt), _p(wt), _p(us[0]), _p(ugs[0]), _p(us[1]), _p(ugs[1]), _p(us[2]), _p(ugs[2]),
                              _p(x), _p(s), _p(elong), _p(uout))

    def advect(self, order, nsteps, dt, fs, x, s, r, elong):
        n = x.shape[1]
        fs = np.asarray(fs, np.float64)
        u, ug = np.zeros((3, n), np.float32), np.zeros((9, n), np.float32)
        self.lib.o3d_ref_advect(order, nsteps, float(dt), _p(fs), n, _p(x), _p(s), _p(r), _p(elong), _p(u), _p(ug))
        return u, ug

    def stats(self, s, elong):
        a, b = c_float(), c_float()
        self.lib.o3d_ref_stats(s.shape[1], _p(s), _p(elong), ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    DENSE_SOLVE_FN = ctypes.CFUNCTYPE(None, ctypes.POINTER(c_float), c_int, ctypes.POINTER(c_float))

    def advect_body(self, order, dt, fs, ips, x, s, r, elong, nodes_i, idx, solve):
        """One Convection::advect step (order 1 or 2) of a particle collection around one static body through the reference's own
        routines; solve(rhs float32[3 np]) -> strengths float32[3 np] stands in for BEM::solve (Eigen). x, s, elong updated in
        place. Returns (BEM solves made, ts (3,np))."""
        fs = np.asarray(fs, np.float64)
        np_ = idx.shape[0]

        def _cb(rhs, n, out):
            sol = np.ascontiguousarray(solve(np.ctypeslib.as_array(rhs, shape=(n,)).copy()), np.float32)
            np.ctypeslib.as_array(out, shape=(n,))[:] = sol
        cb = self.DENSE_SOLVE_FN(_cb)
        ts = np.zeros((3, np_), np.float32)
        k = self.lib.o3d_ref_advect_body(order, float(dt), _p(fs), float(ips), x.shape[1], _p(x), _p(s), _p(r), _p(elong),
                                         nodes_i.shape[0], _p(nodes_i), np_, _p(idx), ctypes.cast(cb, c_void_p), _p(ts))
        return int(k), ts

    def totals(self, x, s):
        """(ElementBase::get_total_circ, Points::get_total_impulse) through the reference's Points<float>: two float32[3]."""
        c, i = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self.lib.o3d_ref_totals(x.shape[1], _p(_f32(x)), _p(_f32(s)), _p(c), _p(i))
        return c, i

    def status_lines(self, path, csv, vals, nv, reset_before):
        """The reference's StatusFile appending len(nv) lines of (time, Nv, g[3], f[3]); vals (nlines,7) float32."""
        vals = np.ascontiguousarray(vals, np.float32)
        nv = np.ascontiguousarray(nv, np.int32)
        rb = np.ascontiguousarray(reset_before, np.int32)
        self.lib.o3d_ref_status_lines(path.encode(), int(csv), len(nv), _p(vals), _p(nv), _p(rb))
        with open(path, "rb") as f:
            return f.read()

    # ---- particle x panel closest-point loops (src/Reflect.h) ----
    def reflect(self, nodes_i, idx, x):
        """reflect_panp2 on nodes_i (nn,3) interleaved, idx (np,3); x (3,nt) in place. Returns particles moved."""
        return int(self.lib.o3d_ref_reflect(nodes_i.shape[0], _p(nodes_i), idx.shape[0], _p(idx), x.shape[1], _p(x)))

    def clear_inner(self, method, nodes_i, idx, x, rad, cutoff_mult, ips):
        return int(self.lib.o3d_ref_clear_inner(method, nodes_i.shape[0], _p(nodes_i), idx.shape[0], _p(idx), x.shape[1], _p(x),
                                                _p(rad), cutoff_mult, ips))

    def write_vtk(self, x, s, r, u, index, frameno, time, directory):
        """Points<float>::write_vtk -> <directory>/part_<index>_<frameno>.vtu; returns the file's bytes."""
        rc = self.lib.o3d_ref_write_vtk(x.shape[1], _p(x), _p(s), _p(r), _p(u), index, frameno, float(time), directory.encode())
        assert rc == 0
        with open(os.path.join(directory, f"part_{index:02d}_{frameno:05d}.vtu"), "rb") as f:
            return f.read()

    # ---- initial conditions from the reference's feature generators (src/FlowFeature.cpp) ----
    def has_features(self) -> bool:
        return hasattr(self.lib, "o3d_ref_has_features") and bool(self.lib.o3d_ref_has_features())

    def singular_ring(self, center, normal, majrad, circ, ips):
        """SingularRing::init_elements(ips) -> x, s (3,n)"""
        c, nn = _f32(center), _f32(normal)
        n = self.lib.o3d_ref_singular_ring(_p(c), _p(nn), majrad, circ, ips, None, None, 0)
        x, s = np.zeros((3, n), np.float32), np.zeros((3, n), np.float32)
        self.lib.o3d_ref_singular_ring(_p(c), _p(nn), majrad, circ, ips, _p(x), _p(s), n)
        return x, s

    def thick_ring(self, center, normal, majrad, minrad, circ, ips):
        """ThickRing::init_elements(ips) -> x, s (3,n)"""
        c, nn = _f32(center), _f32(normal)
        n = self.lib.o3d_ref_thick_ring(_p(c), _p(nn), majrad, minrad, circ, ips, None, None, 0)
        x, s = np.zeros((3, n), np.float32), np.zeros((3, n), np.float32)
        self.lib.o3d_ref_thick_ring(_p(c), _p(nn), majrad, minrad, circ, ips, _p(x), _p(s), n)
        return x, s

    def max_threads(self):
        return int(self.lib.o3d_ref_max_threads())

    def set_accel(self, accel: int):
        """accel_t of the ExecEnv handed to the reference routines: 1 cpu_x86, 4 gpu_cuda."""
        self.lib.o3d_ref_set_accel(int(accel))

    def built_with_cuda(self) -> bool:
        return bool(self.lib.o3d_ref_built_with_cuda())

    def pts_on_pts(self, sx, sr, ss, tx, tr, tu, tug, targ_kind=None):
        ns, nt = sx.shape[1], tx.shape[1]
        if targ_kind is None:
            targ_kind = self.TARG_BLOB if tr is not None else (self.TARG_FIELD if tug is not None else self.TARG_TRACER)
        rc = self.lib.o3d_ref_pts_on_pts(ns, _p(sx[0]), _p(sx[1]), _p(sx[2]), _p(sr), _p(ss), targ_kind, nt,
                                         _p(tx[0]), _p(tx[1]), _p(tx[2]), _p(tr), int(tug is not None), _p(tu), _p(tug))
        if rc != 0:
            raise RuntimeError("reference asserts on this target-kind/result-type combination")

    def surface_props(self, nodes_i, idx, val):
        """nodes_i (nn,3) interleaved, idx (np,3), val (np,3) -> area, ts(3,np), b1, b2, nrm (3,np)."""
        nn, np_ = nodes_i.shape[0], idx.shape[0]
        area = np.zeros(np_, np.float32)
        ts, b1, b2, nrm = (np.zeros((3, np_), np.float32) for _ in range(4))
        self.lib.o3d_ref_surface_props(nn, _p(nodes_i), np_, _p(idx), _p(val), _p(area), _p(ts), _p(b1), _p(b2), _p(nrm))
        return area, ts, b1, b2, nrm

    def pan_on_pts(self, nodes_i, idx, val, tx, tr, tu, tug, targ_kind=None):
        nn, np_, nt = nodes_i.shape[0], idx.shape[0], tx.shape[1]
        if targ_kind is None:
            targ_kind = self.TARG_BLOB if tr is not None else (self.TARG_FIELD if tug is not None else self.TARG_TRACER)
        self.lib.o3d_ref_pan_on_pts(nn, _p(nodes_i), np_, _p(idx), _p(val), targ_kind, nt,
                                    _p(tx[0]), _p(tx[1]), _p(tx[2]), _p(tr), _p(tu), _p(tug))

    def pts_on_pan(self, sx, sr, ss, nodes_i, idx, val, pu):
        ns, nn, np_ = sx.shape[1], nodes_i.shape[0], idx.shape[0]
        self.lib.o3d_ref_pts_on_pan(ns, _p(sx[0]), _p(sx[1]), _p(sx[2]), _p(sr), _p(ss), nn, _p(nodes_i), np_,
                                    _p(idx), _p(val), _p(pu))

    def pan_on_pan(self, nodes_i, idx, val, tnodes_i, tidx, tval):
        pu = np.zeros((3, tidx.shape[0]), np.float32)
        self.lib.o3d_ref_pan_on_pan(nodes_i.shape[0], _p(nodes_i), idx.shape[0], _p(idx), _p(val),
                                    tnodes_i.shape[0], _p(tnodes_i), tidx.shape[0], _p(tidx), _p(tval), _p(pu))
        return pu

    def pan_on_pan_coeff(self, nodes_i, idx, bc, target=None):
        if target is None:
            n = 9 * idx.shape[0] ** 2
            out = np.zeros(n, np.float32)
            self.lib.o3d_ref_pan_on_pan_coeff(nodes_i.shape[0], _p(nodes_i), idx.shape[0], _p(idx), _p(bc), 1,
                                              0, None, 0, None, None, _p(out))
            return out
        tn, ti, tb = target
        out = np.zeros(9 * idx.shape[0] * ti.shape[0], np.float32)
        self.lib.o3d_ref_pan_on_pan_coeff(nodes_i.shape[0], _p(nodes_i), idx.shape[0], _p(idx), _p(bc), 0,
                                          tn.shape[0], _p(tn), ti.shape[0], _p(ti), _p(tb), _p(out))
        return out

    def kernel_0v_0bg(self, s7, t4):
        out = np.zeros(12, np.float64)
        self.lib.o3d_ref_kernel_0v_0bg(_p(_f32(s7)), _p(_f32(t4)), _p(out))
        return out

    def rkernel(self, grads, tri9, str4, t3, sa):
        out = np.zeros(12 if grads else 3, np.float64)
        fn = self.lib.o3d_ref_rkernel_2vs_0pg if grads else self.lib.o3d_ref_rkernel_2vs_0p
        flops = fn(_p(_f32(tri9)), _p(_f32(str4)), _p(_f32(t3)), c_float(sa), _p(out))
        return out, flops


def host_threads() -> int:
    """Host cores this process may run on. NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its ranks."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def host_has_avx512() -> bool:
    """The x86-64-v4 feature set (what -march=x86-64-v4 code needs), from /proc/cpuinfo."""
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((l for l in f if l.startswith("flags")), "").split()
    except OSError:
        return False
    return all(x in flags for x in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"))


def have_reference() -> bool:
    return os.path.exists(os.path.join(OUT, "libo3d_ref.so")) or os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))
