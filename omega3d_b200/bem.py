"""Matrix-free boundary-element operator: the product ``A x`` that ``BEM<S,I>::solve`` needs, without storing ``A``.

The reference assembles the dense influence matrix with ``panels_on_panels_coeff`` (src/Coefficients.h:169-483, paths
relative to /root/reference), copies it block by block into an Eigen matrix (``BEM::set_block``, src/BEM.h:117-136) and
hands it to Eigen's GMRES (:182-202), which touches ``A`` only through ``A * x``. ``(3 np)^2`` floats cap the panel
count (src/Simulation.cpp:675). ``PanelOperator`` is that product on the GPU, recomputing every 3 x 3 block from the
panel geometry with the same device function that assembles the matrix (``csrc/biot_panel.cuh: coef_block``).

``BEM`` mirrors the reference class around it: ``set_rhs`` / ``solve`` / ``getStrengths``. The Krylov recurrences of the
solver are the reference's host-side O(n m) work (Eigen there, numpy here - restarted GMRES(30) with the diagonal
preconditioner and the float-epsilon tolerance Eigen defaults to); every ``A * x`` is a GPU call.
"""
from __future__ import annotations

import math
from ctypes import byref, c_double, c_void_p

import numpy as np

from .influence import CudaContext, Surfaces, _ptr, default_context, f32


class PanelOperator:
    """y = A x, A the (3 ntarg) x (3 nsrc) block of ``panels_on_panels_coeff(src, targ)`` (include/o3d_cuda.h: o3d_bem_op)."""

    def __init__(self, src: Surfaces, targ: Surfaces, ctx: CudaContext = None):
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        self.nsp, self.ntp = src.np_, targ.np_
        self.self_block = src is targ
        h = c_void_p()
        self.ctx.check(self.lib.o3d_cuda_bem_op_create(
            self.ctx.h, src.x.shape[1], _ptr(src.x[0]), _ptr(src.x[1]), _ptr(src.x[2]), src.np_, _ptr(src.idx), _ptr(src.b1),
            _ptr(src.b2), _ptr(src.area), targ.x.shape[1], _ptr(targ.x[0]), _ptr(targ.x[1]), _ptr(targ.x[2]), targ.np_,
            _ptr(targ.idx), _ptr(targ.b1), _ptr(targ.b2), _ptr(targ.nrm), _ptr(targ.area), int(self.self_block), byref(h)))
        self.h = h
        self.flops = 0.0
        self.applies = 0

    @property
    def shape(self):
        return 3 * self.ntp, 3 * self.nsp

    def matvec(self, x):
        x = np.ascontiguousarray(x, f32)
        if x.shape != (3 * self.nsp,):
            raise ValueError(f"expected a vector of {3 * self.nsp} unknowns")
        y = np.empty(3 * self.ntp, f32)
        fl = c_double()
        self.ctx.check(self.lib.o3d_cuda_bem_op_apply(self.ctx.h, self.h, _ptr(x), _ptr(y), byref(fl)))
        self.flops = fl.value
        self.applies += 1
        return y

    __matmul__ = matvec

    def diagonal(self):
        """diag(A). The self block's diagonal 3 x 3 blocks are fixed by the reference (src/Coefficients.h:414-436, then
        * 1/4pi): (0, 0, 2pi/4pi); for a cross block the diagonal has no special meaning and is not needed."""
        if not self.self_block:
            raise ValueError("diagonal() is defined for the self-influence block only")
        return np.tile(np.array([0.0, 0.0, 0.5], f32), self.ntp)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.o3d_cuda_bem_op_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def vels_to_rhs_panels(targ: Surfaces):
    """src/RHS.h:55-121 with three unknowns per panel: the negated panel-centre velocity along x1, x2 and the normal, in
    float, interleaved per panel."""
    u = targ.pu
    rhs = np.empty((targ.np_, 3), f32)
    for k, b in enumerate((targ.b1, targ.b2, targ.nrm)):
        rhs[:, k] = -((u[0] * b[0] + u[1] * b[1]).astype(f32) + u[2] * b[2]).astype(f32)
    return rhs.reshape(-1)


class DenseBEM:
    """The reference's own arrangement (src/BEM.h:117-203): the dense matrix from panels_on_panels_coeff, assembled once on
    the GPU, and a host solve. `A` is column-major (3 np) x (3 np) as the C ABI returns it; the factorisation is numpy's
    (LAPACK, double) where the reference runs Eigen's GMRES to float epsilon - the same linear system."""

    def __init__(self, A, n):
        self.A = np.asarray(A, np.float64).reshape(n, n).T      # column-major -> rows are target unknowns
        self.b = None
        self.strengths = None

    def set_rhs(self, b):
        self.b = np.ascontiguousarray(b, f32)

    def getStrengths(self):
        return self.strengths

    def solve(self):
        self.strengths = np.linalg.solve(self.A, self.b.astype(np.float64)).astype(f32)
        return self.strengths


def solve_bem_for(targ: Surfaces, bem):
    """The host half of solve_bem (src/BEMHelper.h:103-262) for one static reactive surface whose panel-centre velocities
    targ.pu are already final: right-hand side, solve, strengths back into the surface."""
    bem.set_rhs(vels_to_rhs_panels(targ))
    bem.solve()
    targ.set_str(bem.getStrengths())


class BEM:
    """src/BEM.h:44-73 around a matrix-free operator: set_rhs, solve, getStrengths."""

    def __init__(self, op: PanelOperator, restart: int = 30, tol: float = float(np.finfo(f32).eps), max_iters: int = None):
        self.A, self.restart, self.tol = op, restart, tol
        self.max_iters = max_iters or 2 * op.shape[1]       # Eigen: 2 * cols
        self.b = None
        self.strengths = None
        self.iterations, self.error = 0, 0.0

    def set_rhs(self, b):
        self.b = np.ascontiguousarray(b, f32)

    def getRhs(self):
        return self.b

    def getStrengths(self):
        return self.strengths

    def solve(self):
        """Restarted GMRES with Eigen's defaults (DiagonalPreconditioner: 1/a_ii where a_ii != 0, else 1; restart 30;
        tolerance on the preconditioned residual relative to the preconditioned right-hand side)."""
        n = self.b.size
        d = self.A.diagonal().astype(np.float64)
        minv = np.where(d != 0.0, 1.0 / np.where(d != 0.0, d, 1.0), 1.0)
        b = self.b.astype(np.float64)
        x = np.zeros(n)
        bnorm = np.linalg.norm(minv * b)
        self.iterations = 0
        if bnorm == 0.0:
            self.strengths, self.error = x.astype(f32), 0.0
            return self.strengths
        while True:
            r = minv * (b - self.A.matvec(x.astype(f32)).astype(np.float64))
            beta = np.linalg.norm(r)
            self.error = beta / bnorm
            if self.error <= self.tol or self.iterations >= self.max_iters:
                break
            m = self.restart
            V = np.zeros((m + 1, n)); H = np.zeros((m + 1, m)); cs = np.zeros(m); sn = np.zeros(m); g = np.zeros(m + 1)
            V[0] = r / beta
            g[0] = beta
            k_used = 0
            for k in range(m):
                w = minv * self.A.matvec(V[k].astype(f32)).astype(np.float64)
                self.iterations += 1
                for i in range(k + 1):                      # modified Gram-Schmidt
                    H[i, k] = w @ V[i]
                    w -= H[i, k] * V[i]
                H[k + 1, k] = np.linalg.norm(w)
                if H[k + 1, k] > 0:
                    V[k + 1] = w / H[k + 1, k]
                for i in range(k):                          # previous Givens rotations
                    t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                    H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                    H[i, k] = t
                rho = math.hypot(H[k, k], H[k + 1, k])
                cs[k], sn[k] = H[k, k] / rho, H[k + 1, k] / rho
                H[k, k], H[k + 1, k] = rho, 0.0
                g[k + 1] = -sn[k] * g[k]
                g[k] = cs[k] * g[k]
                k_used = k + 1
                if abs(g[k + 1]) / bnorm <= self.tol or self.iterations >= self.max_iters:
                    break
            yk = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used])
            x += V[:k_used].T @ yk
        self.strengths = x.astype(f32)
        return self.strengths
