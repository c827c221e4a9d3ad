"""GPU (-m gpu): reflect_panp2 / clear_inner_panp2 on the device (csrc/reflect.cuh through o3d_cuda_reflect_pts /
o3d_cuda_clear_inner_pts) against the golden vectors from the reference and against the oracle. Integer-like work:
which panel feature is closest, whether a particle is under the surface - the bar is bit-exact positions and counts."""
import numpy as np
import pytest

from conftest import golden

from omega3d_b200 import influence as I
from omega3d_b200 import reflect as R
from omega3d_b200 import workloads as W

pytestmark = pytest.mark.gpu
f32 = np.float32


def soa(nodes_i):
    return np.ascontiguousarray(nodes_i.T)


def sphere_from_golden(g):
    s = I.Surfaces(soa(g["nodes_i"]), g["idx"], None, I.reactive, I.fixed)
    s.nrm = np.ascontiguousarray(g["nrm"])      # the reference's own normals (the numpy mirror agrees to 3e-7, not to the bit)
    return s


def pts(x):
    return I.Points(x.copy(), np.zeros_like(x), 0.03, I.active, I.lagrangian)


def test_reflect_golden_bit_exact(cuda_ctx):
    g = golden("reflect.npz")
    surf, p = sphere_from_golden(g), pts(g["x0"])
    assert R.reflect_panp2(surf, p, cuda_ctx) == int(g["reflect_moved"])
    assert np.array_equal(p.x, g["reflect_x"])
    assert R.reflect_interior([surf], [p], cuda_ctx) == 0          # everything is outside now


def test_clear_inner_golden_bit_exact(cuda_ctx):
    g = golden("reflect.npz")
    surf, p = sphere_from_golden(g), pts(g["x0"])
    assert R.clear_inner_panp2(1, surf, p, float(g["clear_cm"]), float(g["clear_ips"]), cuda_ctx) == int(g["clear_moved"])
    assert np.array_equal(p.x, g["clear_x"])
    p = pts(g["reflect_x"])
    assert R.clear_inner_layer(1, [surf], [p], 0.2, 0.05, cuda_ctx) == int(g["both_moved"])
    assert np.array_equal(p.x, g["both_x"])
    with pytest.raises(I.O3DError):                                # method 0 has no caller in the reference
        R.clear_inner_panp2(0, surf, p, 0.2, 0.05, cuda_ctx)
    fixed = I.Points(g["x0"].copy(), np.zeros_like(g["x0"]), 0.03, I.active, I.fixed)
    assert R.clear_inner_layer(1, [surf], [fixed], 0.2, 0.05, cuda_ctx) == 0 and np.array_equal(fixed.x, g["x0"])


@pytest.mark.parametrize("levels,nt", [(1, 777), (3, 50000)])
def test_against_oracle_other_sizes(cuda_ctx, restate, levels, nt):
    """80 panels (one partial tile) and 1280 panels (five tiles) against the restatement, ragged particle counts."""
    nodes_i, idx = W.icosphere(levels, 0.5)
    surf = I.Surfaces(soa(nodes_i), idx, None, I.reactive, I.fixed)
    rng = np.random.default_rng(levels)
    x = rng.uniform(-0.7, 0.7, (3, nt)).astype(f32)
    p = pts(x)
    ref = x.copy()
    assert R.reflect_panp2(surf, p, cuda_ctx) == restate.reflect(surf.x, surf.idx, surf.nrm, ref)
    assert np.array_equal(p.x, ref)
    assert R.clear_inner_panp2(1, surf, p, 0.3, 0.04, cuda_ctx) == restate.clear_inner(surf.x, surf.idx, surf.nrm, ref, 0.3, 0.04)
    assert np.array_equal(p.x, ref)
    # size-independent property: after both passes every particle is outside the body and above the cutoff layer
    rad = np.sqrt((p.x.astype(np.float64) ** 2).sum(0))
    tri = surf.x[:, surf.idx].astype(np.float64)                       # (3, np, 3 nodes)
    inradius = np.abs((tri.mean(axis=2) * surf.nrm).sum(0)).min()      # the faceted sphere's inscribed radius
    assert rad.min() > inradius


def test_empty_inputs(cuda_ctx):
    nodes_i, idx = W.icosphere(0, 0.5)
    surf = I.Surfaces(soa(nodes_i), idx, None, I.reactive, I.fixed)
    empty = I.Points(np.zeros((3, 0), f32), np.zeros((3, 0), f32), 0.1, I.active, I.lagrangian)
    assert R.reflect_panp2(surf, empty, cuda_ctx) == 0


def test_two_devices_equal_one(cuda_ctx):
    if cuda_ctx.lib.o3d_cuda_device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx2 = I.CudaContext((0, 1))
    g = golden("reflect.npz")
    surf, p = sphere_from_golden(g), pts(g["x0"])
    assert R.reflect_panp2(surf, p, ctx2) == int(g["reflect_moved"]) and np.array_equal(p.x, g["reflect_x"])
    ctx2.close()
