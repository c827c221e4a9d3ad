"""CPU: the C-ABI library builds, loads and exports every symbol include/o3d_cuda.h declares; the host-side
mirror of the reference interface behaves like the reference (no compute calls here - no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden

from omega3d_b200 import _lib, influence as I
from omega3d_b200 import workloads as W

f32 = np.float32


def header_symbols():
    src = open(os.path.join(ROOT, "include", "o3d_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(o3d_cuda_\w+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/o3d_cuda.h but not exported"
    # and the binding table covers exactly the header
    assert sorted(_lib.SYMBOLS) == names
    assert _lib.load().o3d_cuda_abi_version() == 3


def test_library_targets_sm100a_with_bulk_copy_and_packed_fma():
    """Evidence the product kernels are the sm_100a ones: UBLKCP (cp.async.bulk) and FFMA2 in the SASS."""
    import subprocess
    _lib.build()
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass and "FFMA2" in sass and "MUFU.RSQ" in sass


def test_no_device_means_error_not_fallback():
    lib = _lib.load()
    if lib.o3d_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(I.O3DError):
        I.CudaContext((0,))
    h = ctypes.c_void_p()
    assert lib.o3d_cuda_create(ctypes.byref(h), 1, None) == 4  # O3D_ERR_NODEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "omega3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in text and "biot_oracle" not in text and "libo3d_ref" not in text, f


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """oracle/ is the checker: besides tests/ only bench.py (CPU-baseline legs) and __graft_entry__.py (smoke, and build(),
    which compiles it) may name it - not the package, not integration/, not scripts/ or tools/."""
    for sub in ("omega3d_b200", "integration", "include", "scripts", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            if os.sep + "lib" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh", ".inc")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "oracle_py" not in text and "from oracle" not in text and "libo3d_oracle" not in text, os.path.join(dirpath, f)


def test_execenv_and_resultstype_mirror_reference():
    e = I.ExecEnv()
    assert e.is_internal() and e.get_instrs() == I.accel_t.gpu_cuda and int(I.accel_t.gpu_cuda) == 4
    assert e.to_string() == " CUDA-accelerated direct sums"
    e.set_instrs(I.accel_t.cpu_x86)
    with pytest.raises(I.O3DError):
        I._require_cuda(e)
    r = I.ResultsType(I.velandgrad)
    assert r.compute_vel() and r.compute_grad() and not r.compute_psi()


def test_points_storage_rules_match_reference():
    x = np.zeros((3, 5), f32)
    assert I.Points(x, x, 0.1, I.active, I.lagrangian).ug is not None
    assert I.Points(x, e=I.inert, m=I.fixed).ug is not None          # field points keep gradients
    assert I.Points(x, e=I.inert, m=I.lagrangian).ug is None         # tracers do not (src/Points.h:97-106)
    p = I.Points(x, x, 0.1)
    p.u[:] = 4 * np.pi
    p.finalize_vels((1.0, 0.0, 0.0))
    np.testing.assert_allclose(p.u[0], 2.0, rtol=1e-6)


def test_surfaces_bases_match_reference_ctor():
    g = golden("panels_80.npz")
    s = I.Surfaces(np.ascontiguousarray(g["nodes_i"].T), g["idx"], g["val"], I.active, I.fixed)
    np.testing.assert_allclose(s.area, g["area"], rtol=3e-7)
    for mine, ref in ((s.b1, g["b1"]), (s.b2, g["b2"]), (s.nrm, g["nrm"])):
        np.testing.assert_allclose(mine, ref, atol=3e-7)
    np.testing.assert_allclose(s.ts, g["ts"], rtol=2e-6, atol=1e-9)
    with pytest.raises(ValueError):
        I.Surfaces(np.zeros((3, 2), f32), np.array([[0, 1, 2]], np.uint32))


def test_workloads_are_seeded():
    a = W.random_cloud(1000)
    b = W.random_cloud(1000)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert abs(float(a[2][0]) - 1.5 * 1000 ** (-1 / 3)) < 1e-6
    n, i = W.icosphere(2)
    assert i.shape == (320, 3)  # the flow_over_sphere body of SURVEY.md 8 (C4)


def test_squared_threshold_identity():
    """csrc/biot_panel.cuh: sq_threshold. For a correctly rounded float sqrt, T = max{x : sqrt(x) <= thr} satisfies
    sqrt(x) > thr <=> x > T for every float x >= 0, and a subdivision level down (thr/2, T/4) it still does - so the panel
    kernels take the reference's stop decision (src/Kernels.h:979-1002) from the squared distance, bit for bit, without
    the square root. Brute force over the float neighbourhood of T, same stepping algorithm as the device function."""
    def t_of(thr):
        xi = np.array([f32(thr) * f32(thr)], f32).view(np.uint32)
        for _ in range(8):
            if xi.view(f32)[0] > 0 and np.sqrt(xi.view(f32)[0]) > thr:
                xi -= 1
            else:
                break
        for _ in range(8):
            if not (np.sqrt((xi + 1).view(f32)[0]) <= thr):
                break
            xi += 1
        return xi.view(f32)[0]

    rng = np.random.default_rng(0)
    thrs = np.concatenate([rng.random(600).astype(f32) * f32(2), rng.random(600).astype(f32) * f32(1e-3),
                           np.array([0, 1, 0.5, 4 * np.sqrt(f32(0.0098))], f32)])
    for thr in thrs:
        t = t_of(thr)
        for lev in range(4):
            thr_l, t_l = f32(thr) * f32(0.5) ** lev, t * f32(0.25) ** lev
            ti = int(np.array([t_l], f32).view(np.uint32)[0])
            xs = np.arange(max(ti - 40, 0), ti + 40, dtype=np.uint32).view(f32)
            assert np.array_equal(np.sqrt(xs) > thr_l, xs > t_l), (thr, lev)


def test_launch_plan_fills_the_gpu_at_the_benchmark_sizes():
    """Host logic of the launch shape (capi.cu: pp_shape through o3d_cuda_plan_pts_on_pts, no device needed): at every
    size of the BASELINE sweep, whole or sharded over 2/4/8 GPUs, the persistent CTAs of a 148-SM B200 all stream the
    same number of tiles +-1 (balance >= 99.95 % from 2 M up, >= 99.3 % at 256 K over 8 GPUs: stream-K: no last-wave quantisation at any size), at most grid-1 target blocks
    are shared between CTAs, and the workspace is the same few MB whatever the size."""
    lib = _lib.load()
    from ctypes import byref, c_double, c_int, c_int64
    for n in (1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24):
        for gpus in (1, 2, 4, 8):
            nt = n // gpus
            grid, sp, bal, ws = c_int64(), c_int(), c_double(), c_int64()
            assert lib.o3d_cuda_plan_pts_on_pts(148, n, nt, 1, byref(grid), byref(sp), byref(bal), byref(ws)) == 0
            assert grid.value == 148 and 0 <= sp.value <= 147   # one persistent 384-thread CTA per SM
            assert bal.value >= (0.9995 if n >= 1 << 21 else 0.993), (n, gpus, bal.value)   # one tile of granularity
            assert ws.value == 148 * 2 * 12 * 768 * 8
    # tiny target counts: the source tiles of the one target block are dealt out over the CTAs
    grid, sp, bal, ws = c_int64(), c_int(), c_double(), c_int64()
    assert lib.o3d_cuda_plan_pts_on_pts(148, 100000, 320, 0, byref(grid), byref(sp), byref(bal), byref(ws)) == 0
    assert grid.value == 196 and sp.value == 1            # a system below one 1536-target block runs as 128-thread CTAs: 196 tiles, one block of 512 targets
    assert lib.o3d_cuda_plan_pts_on_pts(148, 700, 5, 1, byref(grid), byref(sp), byref(bal), byref(ws)) == 0
    assert grid.value == 2 and sp.value == 1              # 700 sources are two 512-record tiles
    assert lib.o3d_cuda_plan_pts_on_pts(148, 100, 100, 1, byref(grid), byref(sp), byref(bal), byref(ws)) == 0
    assert grid.value == 1 and sp.value == 0 and bal.value == 1.0
    assert lib.o3d_cuda_plan_pts_on_pts(0, 10, 10, 1, None, None, None, None) == 1   # O3D_ERR_INVALID


def test_stream_k_bookkeeping_replayed_on_the_host():
    """o3d_cuda_plan_check walks every CTA's tiles with the kernels' own bookkeeping (pp_ring_start / pp2_walk) and
    every boundary with pp_fixup_kernel's: each (target block, source tile) unit consumed once, each partial segment in
    its own workspace slot, each shared block finished once from exactly the slots written. Ragged shapes, every regime:
    fewer units than CTAs, blocks spanning many CTAs, CTAs spanning many blocks, boundaries on block edges."""
    lib = _lib.load()
    rng = np.random.Generator(np.random.MT19937(5))
    shapes = [(1, 1), (512, 256), (513, 257), (700, 5), (100000, 320), (1 << 20, 1 << 20), (1 << 20, 1 << 17), (1 << 22, 1 << 22),
              (1 << 24, 1 << 21), (444 * 512, 256), (443 * 512, 512), (445 * 512, 256 * 3), (512 * 37, 256 * 444), (512 * 37, 256 * 12),
              (148 * 512, 768), (147 * 512, 1536), (149 * 512, 768 * 3), (512 * 37, 768 * 148), (512 * 37, 768 * 5)]
    shapes += [(int(rng.integers(1, 300000)), int(rng.integers(1, 300000))) for _ in range(60)]
    for ns, nt in shapes:
        for grad in (0, 1):
            for sms in (148, 132, 1, 7):
                assert lib.o3d_cuda_plan_check(sms, ns, nt, grad) == 0, (ns, nt, grad, sms)


def test_patch_keeps_the_reference_else_chain():
    """integration/omega3d_use_cuda.patch, Influence.h hunk: the reference's GL arm ends in a dangling `} else // if not
    gpu_opengl` whose statement is the CPU block. The CUDA arm inserted between the two must itself end in `else`, or - in a
    build with USE_OGL_COMPUTE and USE_CUDA both defined - the CPU block would run unconditionally after a GL dispatch that
    found its compute state busy."""
    import re
    text = open(os.path.join(ROOT, "integration", "omega3d_use_cuda.patch")).read()
    hunk = text[text.index("   } else // if not gpu_opengl"):text.index("   { // perform summations using internal CPU solver")]
    added = [l[1:] for l in hunk.splitlines() if l.startswith("+")]
    assert added[0] == "#ifdef USE_CUDA" and added[-2] == "#endif"
    body = "\n".join(added[1:-2])
    assert re.match(r"\s*if \(env\.is_internal\(\) and env\.get_instrs\(\) == gpu_cuda\) \{", body)
    assert re.search(r"return;\s*\} else\b[^\n]*$", body), "the CUDA arm must end in `} else`"
