"""Host-side mirror of the reference's influence interface for the Biot-Savart path, over the C ABI.

The names, argument meaning and error behaviour follow the reference (paths relative to
/root/reference): ``ExecEnv``/``accel_t``/``summation_t`` (src/ExecEnv.h:26-123), ``ResultsType``
(src/ResultsType.h:27-95), the element containers ``Points``/``Surfaces`` reduced to the SoA arrays the
path reads and writes (src/Points.h, src/Surfaces.h, src/ElementBase.h), the four routines
``points_affect_points`` / ``panels_affect_points`` / ``points_affect_panels`` / ``panels_affect_panels``
(src/Influence.h:67,557,1107,1224), ``panels_on_panels_coeff`` (src/Coefficients.h:169) and the
``InfluenceVisitor`` double dispatch (src/Influence.h:1251-1261).

Every routine runs on the GPU through ``include/o3d_cuda.h`` (ctypes). There is no CPU arm here: an
``ExecEnv`` that does not select ``gpu_cuda`` raises, as does a missing library or device.
"""
from __future__ import annotations

import ctypes
import enum
import math
from ctypes import POINTER, byref, c_double, c_int, c_void_p

import numpy as np

from . import _lib

f32 = np.float32


# ---- src/ExecEnv.h ------------------------------------------------------------------------------------
class summation_t(enum.IntEnum):
    direct = 1
    barneshut = 2
    vic = 3
    fmm = 4


class accel_t(enum.IntEnum):
    cpu_x86 = 1
    cpu_vc = 2
    gpu_opengl = 3
    gpu_cuda = 4  # "unsupported internally" in the reference (src/ExecEnv.h:38); this package is that arm


class ExecEnv:
    """src/ExecEnv.h:43-123. The default here is what a ``-DUSE_CUDA`` build's default ctor selects."""

    def __init__(self, internal: bool = True, useomp: bool = True, sumtype: summation_t = summation_t.direct,
                 acceltype: accel_t = accel_t.gpu_cuda):
        self.m_internal, self.m_useomp, self.m_summ, self.m_accel = internal, useomp, sumtype, acceltype

    def is_internal(self):
        return self.m_internal

    def get_instrs(self):
        return self.m_accel

    def set_instrs(self, a: accel_t):
        self.m_accel = a

    def to_string(self):
        if not self.m_internal:
            return " external solver"
        s = {accel_t.cpu_x86: " native", accel_t.cpu_vc: " Vc-accelerated", accel_t.gpu_opengl: " OpenGL-accelerated",
             accel_t.gpu_cuda: " CUDA-accelerated"}[self.m_accel]
        s += " direct sums" if self.m_summ == summation_t.direct else " treecode"
        return s


# ---- src/ResultsType.h ---------------------------------------------------------------------------------
class core_t(enum.IntEnum):
    """The core functions of src/CoreFunc.h (USE_WL_KERNEL / USE_RM_KERNEL / USE_EXPONENTIAL_KERNEL / USE_V2_KERNEL,
    :35-38); values are include/o3d_cuda.h's O3D_CORE_*."""
    wl = 0
    rm = 1
    exp = 2
    v2 = 3


class results_t(enum.IntEnum):
    velonly = 1
    velandgrad = 2
    psionly = 3
    velandvort = 4


class ResultsType:
    def __init__(self, r: results_t = results_t.velonly):
        self.m_rtype = results_t(r)

    def compute_vel(self):
        return self.m_rtype in (results_t.velonly, results_t.velandgrad, results_t.velandvort)

    def compute_grad(self):
        return self.m_rtype == results_t.velandgrad

    def compute_psi(self):
        return self.m_rtype == results_t.psionly

    def compute_vort(self):
        return self.m_rtype == results_t.velandvort


velonly, velandgrad = results_t.velonly, results_t.velandgrad


# ---- src/Omega3D.h elem_t / move_t ----------------------------------------------------------------------
class elem_t(enum.IntEnum):
    active = 1
    reactive = 2
    inert = 3


class move_t(enum.IntEnum):
    lagrangian = 1
    bodybound = 2
    fixed = 3


active, reactive, inert = elem_t.active, elem_t.reactive, elem_t.inert
lagrangian, bodybound, fixed = move_t.lagrangian, move_t.bodybound, move_t.fixed


def _rows(a, nrow, n):
    a = np.ascontiguousarray(a, dtype=f32)
    if a.shape != (nrow, n):
        raise ValueError(f"expected shape ({nrow},{n}), got {a.shape}")
    return a


class Points:
    """The SoA arrays of ``Points<float>`` the path touches (src/Points.h:54-140, src/ElementBase.h).

    x (3,n) positions; s (3,n) strengths (not for inert); r (n,) radii (not for inert); u (3,n) velocity;
    ug (9,n) velocity gradient, slot 3*j+i = d u_i / d x_j - present for everything except inert
    lagrangian tracers, exactly as the reference allocates it (src/Points.h:97-106)."""

    def __init__(self, x, s=None, r=None, e: elem_t = active, m: move_t = lagrangian):
        x = np.ascontiguousarray(x, dtype=f32)
        self.n = x.shape[1]
        self.x = _rows(x, 3, self.n)
        self.E, self.M = elem_t(e), move_t(m)
        if self.E == inert:
            self.s, self.r = None, None
        else:
            self.s = _rows(s, 3, self.n)
            self.r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, f32), (self.n,)), dtype=f32).copy()
        self.u = np.zeros((3, self.n), f32)
        self.ug = None if (self.E == inert and self.M == lagrangian) else np.zeros((9, self.n), f32)
        self.elong = None if self.E == inert else np.ones(self.n, f32)   # src/Points.h:105-109

    def get_n(self): return self.n
    def get_elong(self): return self.elong
    def is_inert(self): return self.E == inert
    def get_pos(self): return self.x
    def get_str(self): return self.s
    def get_rad(self): return self.r
    def get_vel(self): return self.u
    def get_velgrad(self): return self.ug

    def zero_vels(self):
        """src/Points.h:252-262"""
        self.u[:] = 0
        if self.ug is not None:
            self.ug[:] = 0

    def finalize_vels(self, fs=(0.0, 0.0, 0.0)):
        """src/ElementBase.h:187-192 (u = fs + u/4pi in double) and src/Points.h:269-276 (grads * float(1/4pi))."""
        factor = 0.25 / math.pi
        for d in range(3):
            self.u[d] = (float(fs[d]) + self.u[d].astype(np.float64) * factor).astype(f32)
        if self.ug is not None:
            self.ug *= f32(factor)


class Surfaces:
    """The SoA arrays of ``Surfaces<float>`` the path touches (src/Surfaces.h:62-225).

    nodes x (3,nn); idx (np,3) uint32; val (np,3): vortex-sheet strength along x1, along x2, source-sheet
    strength for active surfaces, or the boundary condition for reactive ones. Derived per panel, as the
    reference's ctor does: basis b1,b2,nrm and area (compute_bases, :766-815), total vortex strength ts
    (vortex_sheet_to_panel_strength, :309-335), panel-centre velocity pu (3,np)."""

    def __init__(self, x, idx, val=None, e: elem_t = reactive, m: move_t = fixed):
        self.x = np.ascontiguousarray(x, dtype=f32)
        assert self.x.ndim == 2 and self.x.shape[0] == 3
        self.idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        self.np_ = self.idx.shape[0]
        if self.np_ and int(self.idx.max()) >= self.x.shape[1]:
            raise ValueError("panel node index out of range")
        self.E, self.M = elem_t(e), move_t(m)
        val = np.zeros((self.np_, 3), f32) if val is None else np.ascontiguousarray(val, dtype=f32).reshape(self.np_, 3)
        # ps = (vortex sheet x1, x2, source sheet); reactive surfaces start at zero strength, val is their BC
        self.ps = np.zeros((3, self.np_), f32)
        self.bc = np.zeros((3, self.np_), f32)
        if self.E == active:
            self.ps[:] = val.T
        elif self.E == reactive:
            self.bc[:] = val.T
        self.compute_bases()
        self.vortex_sheet_to_panel_strength()
        self.pu = np.zeros((3, self.np_), f32)

    def get_n(self): return self.x.shape[1]
    def get_npanels(self): return self.np_
    def get_pos(self): return self.x
    def get_idx(self): return self.idx
    def get_str(self): return self.ts
    def get_area(self): return self.area
    def get_x1(self): return self.b1
    def get_x2(self): return self.b2
    def get_norm(self): return self.nrm
    def get_vel(self): return self.pu
    def have_src_str(self): return True
    def get_src_str(self): return self.ps[2]
    def zero_vels(self): self.pu[:] = 0

    def finalize_vels(self, fs=(0.0, 0.0, 0.0)):
        """src/Surfaces.h:877-887: panel-centre velocities = fs + pu / 4pi, formed in double."""
        factor = 0.25 / math.pi
        for d in range(3):
            self.pu[d] = (float(fs[d]) + self.pu[d].astype(np.float64) * factor).astype(f32)

    def compute_bases(self):
        """src/Surfaces.h:766-815: x1 along node0->node1, x2 toward node2, normal x1 x x2, area = base*height/2.
        (float vectors; the two normalisations multiply by a double reciprocal, as the reference does)."""
        p0, p1, p2 = (self.x[:, self.idx[:, k]] for k in range(3))
        x1 = (p1 - p0).astype(f32)
        base = np.sqrt((x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2]).astype(f32)).astype(f32)
        x1 = (x1.astype(np.float64) * (1.0 / base.astype(np.float64))).astype(f32)
        x2 = (p2 - p0).astype(f32)
        dp = (x2[0] * x1[0] + x2[1] * x1[1] + x2[2] * x1[2]).astype(f32)
        x2 = (x2 - dp * x1).astype(f32)
        height = np.sqrt((x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2]).astype(f32)).astype(f32)
        x2 = (x2.astype(np.float64) * (1.0 / height.astype(np.float64))).astype(f32)
        self.area = (0.5 * base.astype(np.float64) * height.astype(np.float64)).astype(f32)
        nrm = np.stack([x1[1] * x2[2] - x1[2] * x2[1], x1[2] * x2[0] - x1[0] * x2[2], x1[0] * x2[1] - x1[1] * x2[0]]).astype(f32)
        self.b1, self.b2, self.nrm = np.ascontiguousarray(x1), np.ascontiguousarray(x2), np.ascontiguousarray(nrm)

    def vortex_sheet_to_panel_strength(self):
        """src/Surfaces.h:309-335: ts = (ps0 * x1 + ps1 * x2) * area."""
        self.ts = np.ascontiguousarray(((self.ps[0] * self.b1 + self.ps[1] * self.b2) * self.area).astype(f32))

    def num_unknowns_per_panel(self):
        """src/Surfaces.h: two vortex-sheet components plus the source sheet (flow_over_sphere.json's body)."""
        return 3

    def set_str(self, new_s):
        """src/Surfaces.h:267-306: the BEM-solved (x1, x2, source) sheet strengths per panel, interleaved, into ps; then
        the total panel strengths ts."""
        new_s = np.ascontiguousarray(new_s, f32).reshape(self.np_, 3)
        self.ps[:] = new_s.T
        self.vortex_sheet_to_panel_strength()

    def represent_as_particles(self, offset: float, ips: float = -1.0) -> "Points":
        """src/Surfaces.h:959-1010: one inert-position / total-strength particle per panel at
        centroid + offset * normal (used by panels_affect_panels with offset 1e-4)."""
        self.vortex_sheet_to_panel_strength()
        p0, p1, p2 = (self.x[:, self.idx[:, k]] for k in range(3))
        px = ((1.0 / 3.0) * ((p0 + p1).astype(f32) + p2).astype(f32).astype(np.float64)).astype(f32)  # (1./3.) is a double there
        px = (px + (f32(offset) * self.nrm).astype(f32)).astype(f32)
        val = self.ts.copy()
        if self.E == reactive:
            val = (val + (self.bc[0] * self.b1 + self.bc[1] * self.b2) * self.area).astype(f32)
        return Points(px, val, 0.0, active, lagrangian)


# ---- the C-ABI context -----------------------------------------------------------------------------------
class O3DError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


class CudaContext:
    """Owns one ``o3d_ctx`` (include/o3d_cuda.h). ``devices``: GPU ordinals driven by this process."""

    def __init__(self, devices=(0,)):
        self.lib = _lib.load()
        devs = (c_int * len(devices))(*devices)
        h = c_void_p()
        rc = self.lib.o3d_cuda_create(byref(h), len(devices), devs)
        if rc != 0:
            raise O3DError(f"o3d_cuda_create failed with code {rc} (no usable sm_100 device?) - there is no CPU fallback")
        self.h = h
        self.flops = 0.0

    def close(self):
        if getattr(self, "h", None):
            self.lib.o3d_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise O3DError(f"o3d_cuda error {rc}: {self.lib.o3d_cuda_last_error(self.h).decode()}")

    def device_props(self, k=0):
        sm, khz, peak = c_int(), c_int(), c_double()
        self.check(self.lib.o3d_cuda_device_props(self.h, k, byref(sm), byref(khz), byref(peak)))
        return {"sm_count": sm.value, "clock_khz": khz.value, "fp32_peak": peak.value}

    def set_tuned_kernels(self, on: bool):
        """Particles-on-points kernel from the SASS-post-processed cubin (default) or as compiled; same results bit for bit."""
        self.check(self.lib.o3d_cuda_set_tuned_kernels(self.h, int(on)))

    def tuned_kernels(self) -> bool:
        return bool(self.lib.o3d_cuda_tuned_kernels(self.h))

    def set_panel_queue(self, on: bool):
        """panels -> points with the warp-level work queue for subdividing pairs (default) or the per-lane round-1 kernel."""
        self.check(self.lib.o3d_cuda_set_panel_queue(self.h, int(on)))

    def set_host_staging(self, on: bool):
        """Pageable host arrays through the context's pinned staging ring (default) or straight to cudaMemcpyAsync."""
        self.check(self.lib.o3d_cuda_set_host_staging(self.h, int(on)))

    def set_core_func(self, core):
        """Core function of the particle kernels: a core_t / its number / its name ("wl", "rm", "exp", "v2"). The
        reference chooses it per build with the one active #define of src/CoreFunc.h:35-38 (USE_WL_KERNEL as shipped)."""
        if isinstance(core, str):
            core = core_t[core.lower()]
        self.check(self.lib.o3d_cuda_set_core_func(self.h, int(core)))

    def core_func(self) -> "core_t":
        return core_t(self.lib.o3d_cuda_core_func(self.h))

    def last_timing(self):
        k, a, b, n = c_double(), c_double(), c_double(), c_int()
        self.check(self.lib.o3d_cuda_last_timing(self.h, byref(k), byref(a), byref(b), byref(n)))
        return {"kernel_ms": k.value, "h2d_ms": a.value, "d2h_ms": b.value, "launches": n.value}

    # -- raw SoA entry points (arrays are float32, C-contiguous rows) --
    @staticmethod
    def _grad_ptrs(tug):
        if tug is None:
            return None, None
        arr = (c_void_p * 9)(*[tug[k].ctypes.data for k in range(9)])
        return arr, ctypes.cast(arr, c_void_p)

    def pts_on_pts(self, sx, sr, ss, tx, tr, tu, tug):
        ns, nt = sx.shape[1], tx.shape[1]
        keep, gp = self._grad_ptrs(tug)
        fl = c_double()
        self.check(self.lib.o3d_cuda_pts_on_pts(self.h, ns, _ptr(sx[0]), _ptr(sx[1]), _ptr(sx[2]), _ptr(sr), _ptr(ss[0]),
                                                _ptr(ss[1]), _ptr(ss[2]), nt, _ptr(tx[0]), _ptr(tx[1]), _ptr(tx[2]), _ptr(tr),
                                                _ptr(tu[0]), _ptr(tu[1]), _ptr(tu[2]), gp, byref(fl)))
        self.flops = fl.value

    def pan_on_pts(self, nodes, idx, ts, area, sss, tx, tu, tug):
        nn, np_, nt = nodes.shape[1], idx.shape[0], tx.shape[1]
        keep, gp = self._grad_ptrs(tug)
        fl = c_double()
        self.check(self.lib.o3d_cuda_pan_on_pts(self.h, nn, _ptr(nodes[0]), _ptr(nodes[1]), _ptr(nodes[2]), np_, _ptr(idx),
                                                _ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]), _ptr(area), _ptr(sss), nt, _ptr(tx[0]),
                                                _ptr(tx[1]), _ptr(tx[2]), _ptr(tu[0]), _ptr(tu[1]), _ptr(tu[2]), gp, byref(fl)))
        self.flops = fl.value

    def pts_on_pan(self, sx, ss, nodes, idx, area, pu):
        ns, nn, np_ = sx.shape[1], nodes.shape[1], idx.shape[0]
        fl = c_double()
        self.check(self.lib.o3d_cuda_pts_on_pan(self.h, ns, _ptr(sx[0]), _ptr(sx[1]), _ptr(sx[2]), _ptr(ss[0]), _ptr(ss[1]),
                                                _ptr(ss[2]), nn, _ptr(nodes[0]), _ptr(nodes[1]), _ptr(nodes[2]), np_, _ptr(idx),
                                                _ptr(area), _ptr(pu[0]), _ptr(pu[1]), _ptr(pu[2]), byref(fl)))
        self.flops = fl.value

    def pan_on_pan_coeff(self, src: "Surfaces", targ: "Surfaces", self_block: bool):
        out = np.empty(9 * src.np_ * targ.np_, f32)
        fl = c_double()
        self.check(self.lib.o3d_cuda_pan_on_pan_coeff(
            self.h, src.x.shape[1], _ptr(src.x[0]), _ptr(src.x[1]), _ptr(src.x[2]), src.np_, _ptr(src.idx), _ptr(src.b1),
            _ptr(src.b2), _ptr(src.area), targ.x.shape[1], _ptr(targ.x[0]), _ptr(targ.x[1]), _ptr(targ.x[2]), targ.np_,
            _ptr(targ.idx), _ptr(targ.b1), _ptr(targ.b2), _ptr(targ.nrm), _ptr(targ.area), int(bool(self_block)), _ptr(out),
            byref(fl)))
        self.flops = fl.value
        return out


_default_ctx = None


def default_context() -> CudaContext:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = CudaContext((0,))
    return _default_ctx


def _require_cuda(env: ExecEnv):
    if not env.is_internal() or env.get_instrs() != accel_t.gpu_cuda:
        raise O3DError("omega3d_b200 implements only the gpu_cuda arm of the influence routines; got" + env.to_string())


# ---- the four influence routines + the coefficient builder ---------------------------------------------
def points_affect_points(src: Points, targ: Points, restype: ResultsType, env: ExecEnv = None, ctx: CudaContext = None):
    """src/Influence.h:67-551. Kernel choice follows :213-533: inert targets -> kernel_0v_0p[g], blob targets ->
    kernel_0v_0b[g]; gradients iff restype says so and the target stores them; an inert target WITHOUT
    gradient storage asked for velandgrad hits the reference's assert(false) (:368-370) - here an error."""
    env = env or ExecEnv()
    _require_cuda(env)
    assert src.get_str() is not None, "sources must carry strengths"
    ctx = ctx or default_context()
    if targ.is_inert():
        if targ.ug is not None and restype.compute_grad():
            tug = targ.ug
        elif not restype.compute_grad():
            tug = None
        else:
            raise O3DError("points_affect_points: inert target without gradient storage asked for velandgrad "
                           "(the reference asserts here, src/Influence.h:368-370)")
        tr = None
    else:
        tr = targ.r
        tug = targ.ug if restype.compute_grad() else None
    ctx.pts_on_pts(src.x, src.r, src.s, targ.x, tr, targ.u, tug)
    return ctx.flops


def panels_affect_points(src: Surfaces, targ: Points, restype: ResultsType = None, env: ExecEnv = None, ctx: CudaContext = None):
    """src/Influence.h:557-1099. Gradients are produced iff the target stores them (:652,875), whatever restype says."""
    env = env or ExecEnv()
    _require_cuda(env)
    if restype is not None:
        assert not restype.compute_psi() and not restype.compute_vort()
    ctx = ctx or default_context()
    sss = src.get_src_str() if src.have_src_str() else None
    ctx.pan_on_pts(src.x, src.idx, src.ts, src.area, sss, targ.x, targ.u, targ.ug)
    return ctx.flops


def points_affect_panels(src: Points, targ: Surfaces, restype: ResultsType = None, env: ExecEnv = None, ctx: CudaContext = None):
    """src/Influence.h:1107-1221: panel-centre velocities, SUBTRACTED (:1210-1212)."""
    env = env or ExecEnv()
    _require_cuda(env)
    ctx = ctx or default_context()
    ctx.pts_on_pan(src.x, src.s, targ.x, targ.idx, targ.area, targ.pu)
    return ctx.flops


def panels_affect_panels(src: Surfaces, targ: Surfaces, restype: ResultsType = None, env: ExecEnv = None, ctx: CudaContext = None):
    """src/Influence.h:1224-1245: colocation points 1e-4 off the target panels, panels_affect_points, copy u -> pu."""
    env = env or ExecEnv()
    _require_cuda(env)
    volsrc = targ.represent_as_particles(0.0001, -1.0)
    volsrc.zero_vels()
    fl = panels_affect_points(src, volsrc, restype, env, ctx)
    targ.pu[:] = volsrc.u
    return fl


def panels_on_panels_coeff(src: Surfaces, targ: Surfaces, ctx: CudaContext = None):
    """src/Coefficients.h:169-483: column-major (3 ntarg) x (3 nsrc) float block; the diagonal override applies
    when src and targ are the same object (:414)."""
    ctx = ctx or default_context()
    return ctx.pan_on_pan_coeff(src, targ, src is targ)


class InfluenceVisitor:
    """src/Influence.h:1251-1261 - double dispatch over (source, target) element kinds."""

    def __init__(self, restype: ResultsType = None, env: ExecEnv = None, ctx: CudaContext = None):
        self.results, self.env, self.ctx = restype or ResultsType(), env or ExecEnv(), ctx

    def __call__(self, src, targ):
        if isinstance(src, Points) and isinstance(targ, Points):
            return points_affect_points(src, targ, self.results, self.env, self.ctx)
        if isinstance(src, Surfaces) and isinstance(targ, Points):
            return panels_affect_points(src, targ, self.results, self.env, self.ctx)
        if isinstance(src, Points) and isinstance(targ, Surfaces):
            return points_affect_panels(src, targ, self.results, self.env, self.ctx)
        if isinstance(src, Surfaces) and isinstance(targ, Surfaces):
            return panels_affect_panels(src, targ, self.results, self.env, self.ctx)
        raise TypeError("unknown element kinds")
