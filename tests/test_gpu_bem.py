"""GPU (-m gpu): the matrix-free BEM operator (o3d_cuda_bem_op_*, csrc/biot_panel.cuh: pan_matvec_kernel) against the
golden influence blocks minted from the reference's panels_on_panels_coeff, and a GMRES solve through it against a dense
solve of the same system."""
import numpy as np
import pytest

from conftest import golden, rel_err

from omega3d_b200 import bem as B
from omega3d_b200 import influence as I
from omega3d_b200 import workloads as W

pytestmark = pytest.mark.gpu
f32 = np.float32


def soa(nodes_i):
    return np.ascontiguousarray(nodes_i.T)


def test_matvec_against_golden_blocks(cuda_ctx):
    g = golden("coeff_20.npz")
    s0 = I.Surfaces(soa(g["n0"]), g["i0"], None, I.reactive, I.fixed)
    s1 = I.Surfaces(soa(g["n1"]), g["i1"], None, I.reactive, I.fixed)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(60).astype(f32)
    for src, targ, key in ((s0, s0, "a_self"), (s0, s1, "a_cross")):
        op = B.PanelOperator(src, targ, cuda_ctx)
        A = g[key].reshape(60, 60).T.astype(np.float64)      # column-major (3 ntp) x (3 nsp)
        y = op.matvec(x)
        assert rel_err(y, A @ x.astype(np.float64)) <= 2e-5
        assert op.flops > 0
        with pytest.raises(ValueError):
            op.matvec(x[:-1])
        op.close()


def test_matvec_equals_assembled_matrix_times_vector(cuda_ctx):
    """Same device function builds the stored matrix and the matrix-free product: they agree to double rounding."""
    nodes, idx = W.icosphere(2, 0.5)                        # 320 panels: the flow_over_sphere body (BASELINE configs[3])
    surf = I.Surfaces(soa(nodes), idx, None, I.reactive, I.fixed)
    A = I.panels_on_panels_coeff(surf, surf, cuda_ctx).reshape(960, 960).T.astype(np.float64)
    op = B.PanelOperator(surf, surf, cuda_ctx)
    rng = np.random.default_rng(5)
    for _ in range(3):
        x = rng.standard_normal(960).astype(f32)
        assert rel_err(op.matvec(x), A @ x.astype(np.float64)) <= 2e-7
    # linearity, exactly representable scalings
    x = rng.standard_normal(960).astype(f32)
    assert np.array_equal(op.matvec(2 * x), 2 * op.matvec(x))
    assert np.array_equal(op.diagonal(), np.diag(A).astype(f32))


def test_gmres_through_the_operator_solves_the_bem_system(cuda_ctx):
    nodes, idx = W.icosphere(2, 0.5)
    surf = I.Surfaces(soa(nodes), idx, None, I.reactive, I.fixed)
    A = I.panels_on_panels_coeff(surf, surf, cuda_ctx).reshape(960, 960).T.astype(np.float64)
    # right-hand side: a uniform freestream resolved on the panel bases, as solve_bem builds it (src/BEMHelper.h:60-110)
    fs = np.array([1.0, 0.0, 0.0])
    b = np.stack([-(fs @ surf.b1), -(fs @ surf.b2), -(fs @ surf.nrm)], axis=1).reshape(-1).astype(f32)
    solver = B.BEM(B.PanelOperator(surf, surf, cuda_ctx), tol=1e-6)
    solver.set_rhs(b)
    x = solver.solve()
    ref = np.linalg.solve(A, b.astype(np.float64))
    assert solver.iterations < 200
    assert np.linalg.norm(A @ x.astype(np.float64) - b) / np.linalg.norm(b) <= 1e-5
    assert rel_err(x, ref) <= 1e-3                          # the system is mildly ill-conditioned; residual is the contract


def test_two_device_operator_equals_one_device(cuda_ctx):
    if cuda_ctx.lib.o3d_cuda_device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx2 = I.CudaContext((0, 1))
    nodes, idx = W.icosphere(2, 0.5)
    surf = I.Surfaces(soa(nodes), idx, None, I.reactive, I.fixed)
    x = np.random.default_rng(1).standard_normal(960).astype(f32)
    a = B.PanelOperator(surf, surf, cuda_ctx).matvec(x)
    b = B.PanelOperator(surf, surf, ctx2).matvec(x)
    assert rel_err(b, a) <= 1e-6
    ctx2.close()
