"""Host-side mirror of the reference's particle-vs-body clean-up loops, over the C ABI.

``reflect_panp2`` / ``reflect_interior`` (src/Reflect.h:194-335, paths relative to /root/reference) mirror particles
that ended up under a body surface back outside; ``clear_inner_panp2`` / ``clear_inner_layer`` (:446-655) push particles
out of the innermost layer. Both are O(particles x panels) closest-point searches the reference runs on the host every
step when bodies exist. Here each is one GPU call (``csrc/reflect.cuh``) returning the reference's bits.
"""
from __future__ import annotations

from ctypes import byref, c_int64

import numpy as np

from .influence import CudaContext, O3DError, Points, Surfaces, _ptr, default_context, f32, lagrangian


def reflect_panp2(src: Surfaces, targ: Points, ctx: CudaContext = None) -> int:
    """src/Reflect.h:194-311. Updates ``targ.x`` in place; returns the number of particles reflected."""
    ctx = ctx or default_context()
    n = c_int64()
    ctx.check(ctx.lib.o3d_cuda_reflect_pts(ctx.h, src.x.shape[1], _ptr(src.x[0]), _ptr(src.x[1]), _ptr(src.x[2]), src.np_,
                                           _ptr(src.idx), _ptr(src.nrm), targ.n, _ptr(targ.x[0]), _ptr(targ.x[1]), _ptr(targ.x[2]),
                                           byref(n)))
    return n.value


def clear_inner_panp2(method: int, src: Surfaces, targ: Points, cutoff_mult: float, ips: float, ctx: CudaContext = None) -> int:
    """src/Reflect.h:446-620. Only ``method`` 1 (push out, keep strength) exists in any call of the reference."""
    ctx = ctx or default_context()
    n = c_int64()
    ctx.check(ctx.lib.o3d_cuda_clear_inner_pts(ctx.h, int(method), src.x.shape[1], _ptr(src.x[0]), _ptr(src.x[1]), _ptr(src.x[2]),
                                               src.np_, _ptr(src.idx), _ptr(src.nrm), targ.n, _ptr(targ.x[0]), _ptr(targ.x[1]),
                                               _ptr(targ.x[2]), float(f32(cutoff_mult)), float(f32(ips)), byref(n)))
    return n.value


def reflect_interior(bdry, vort, ctx: CudaContext = None) -> int:
    """src/Reflect.h:313-335: every Points collection against every Surfaces collection."""
    moved = 0
    for targ in vort:
        if isinstance(targ, Points):
            for src in bdry:
                if isinstance(src, Surfaces):
                    moved += reflect_panp2(src, targ, ctx)
    return moved


def clear_inner_layer(method: int, bdry, vort, cutoff_factor: float, ips: float, ctx: CudaContext = None) -> int:
    """src/Reflect.h:625-655: only collections that move themselves are pushed."""
    moved = 0
    for targ in vort:
        if isinstance(targ, Points) and targ.M == lagrangian:
            for src in bdry:
                if isinstance(src, Surfaces):
                    moved += clear_inner_panp2(method, src, targ, cutoff_factor, ips, ctx)
    return moved
