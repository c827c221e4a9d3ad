"""CPU: the restatement of the reference's closest-point loops (oracle/biot_oracle.c: o3d_oracle_closest_pass) against
golden vectors from the reference's own reflect_panp2 / clear_inner_panp2 (tests/golden/make_golden.py reflect)."""
import numpy as np

from conftest import golden

f32 = np.float32


def soa(nodes_i):
    return np.ascontiguousarray(nodes_i.T)


def test_reflect_bit_identical(restate):
    g = golden("reflect.npz")
    x = g["x0"].copy()
    moved = restate.reflect(soa(g["nodes_i"]), g["idx"], g["nrm"], x)
    assert moved == int(g["reflect_moved"]) == 635 and np.array_equal(x, g["reflect_x"])
    # reflected particles are outside now: a second pass moves nothing
    assert restate.reflect(soa(g["nodes_i"]), g["idx"], g["nrm"], x) == 0


def test_clear_inner_bit_identical(restate):
    g = golden("reflect.npz")
    x = g["x0"].copy()
    moved = restate.clear_inner(soa(g["nodes_i"]), g["idx"], g["nrm"], x, float(g["clear_cm"]), float(g["clear_ips"]))
    assert moved == int(g["clear_moved"]) and np.array_equal(x, g["clear_x"])
    x = g["reflect_x"].copy()
    moved = restate.clear_inner(soa(g["nodes_i"]), g["idx"], g["nrm"], x, 0.2, 0.05)
    assert moved == int(g["both_moved"]) and np.array_equal(x, g["both_x"])


def test_reference_reproduces_reflect_golden(reference_lib):
    if not hasattr(reference_lib.lib, "o3d_ref_reflect"):
        import pytest
        pytest.skip("stale reference build")
    g = golden("reflect.npz")
    x = g["x0"].copy()
    assert reference_lib.reflect(g["nodes_i"], g["idx"], x) == int(g["reflect_moved"])
    assert np.array_equal(x, g["reflect_x"])
