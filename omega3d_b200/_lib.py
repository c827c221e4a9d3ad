"""Loader for the C-ABI shared library (``include/o3d_cuda.h``).

There is exactly one implementation behind this package - the sm_100a CUDA library
``omega3d_b200/lib/libo3d_cuda.so`` built by ``omega3d_b200/csrc/Makefile``. If it is missing or cannot
be loaded the import of any compute entry point raises: there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# O3D_CUDA_LIB: an alternative build of the SAME library (make OUT=... EXTRA=-D..., A/B measurements); never a fallback
LIB_PATH = os.environ.get("O3D_CUDA_LIB") or os.path.join(HERE, "lib", "libo3d_cuda.so")
CSRC = os.path.join(HERE, "csrc")

# every symbol include/o3d_cuda.h declares: (restype, argtypes)
_P = c_void_p
SYMBOLS = {
    "o3d_cuda_abi_version": (c_int, []),
    "o3d_cuda_device_count": (c_int, []),
    "o3d_cuda_create": (c_int, [POINTER(c_void_p), c_int, POINTER(c_int)]),
    "o3d_cuda_destroy": (None, [c_void_p]),
    "o3d_cuda_last_error": (c_char_p, [c_void_p]),
    "o3d_cuda_num_devices": (c_int, [c_void_p]),
    "o3d_cuda_device_props": (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_double)]),
    "o3d_cuda_last_timing": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_double), POINTER(c_int)]),
    "o3d_cuda_pts_on_pts": (c_int, [c_void_p, c_int64] + [_P] * 7 + [c_int64] + [_P] * 4 + [_P] * 3 + [_P, POINTER(c_double)]),
    "o3d_cuda_pan_on_pts": (c_int, [c_void_p, c_int64, _P, _P, _P, c_int64] + [_P] * 6 + [c_int64] + [_P] * 3 + [_P] * 3 + [_P, POINTER(c_double)]),
    "o3d_cuda_pts_on_pan": (c_int, [c_void_p, c_int64] + [_P] * 6 + [c_int64, _P, _P, _P, c_int64, _P, _P] + [_P] * 3 + [POINTER(c_double)]),
    "o3d_cuda_pan_on_pan_coeff": (c_int, [c_void_p, c_int64, _P, _P, _P, c_int64, _P, _P, _P, _P,
                                          c_int64, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, c_int, _P, POINTER(c_double)]),
    "o3d_cuda_packed_records": (c_int64, [c_int64]),
    "o3d_cuda_pack_sources_dev": (c_int, [c_void_p, c_void_p, c_int64] + [_P] * 7 + [c_int64, _P]),
    "o3d_cuda_pts_on_pts_dev": (c_int, [c_void_p, c_void_p, c_int64, _P, c_int64] + [_P] * 4 + [_P] * 3 + [_P, c_int64]),
    "o3d_cuda_pts_finalize_dev": (c_int, [c_void_p, c_void_p, c_int64, _P, _P, _P, _P, c_int64, POINTER(c_double)]),
    "o3d_cuda_pts_move_dev": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_double, POINTER(c_double), _P, _P, _P, _P, _P, _P, _P, _P]),
    "o3d_cuda_particles_create": (c_int, [c_void_p, POINTER(c_void_p)]),
    "o3d_cuda_particles_destroy": (None, [c_void_p, c_void_p]),
    "o3d_cuda_particles_count": (c_int64, [c_void_p]),
    "o3d_cuda_particles_upload": (c_int, [c_void_p, c_void_p, c_int64] + [_P] * 8),
    "o3d_cuda_particles_download": (c_int, [c_void_p, c_void_p] + [_P] * 12),
    "o3d_cuda_particles_find_vels": (c_int, [c_void_p, c_void_p, POINTER(c_double), c_int, POINTER(c_double)]),
    "o3d_cuda_particles_advect": (c_int, [c_void_p, c_void_p, c_int, c_double, c_double, POINTER(c_double), c_int, POINTER(c_double)]),
    "o3d_cuda_particles_stats": (c_int, [c_void_p, c_void_p, POINTER(ctypes.c_float), POINTER(ctypes.c_float)]),
    "o3d_cuda_particles_set_body": (c_int, [c_void_p, c_void_p, c_int64, _P, _P, _P, c_int64, _P, _P, _P, ctypes.c_float, ctypes.c_float, _P, _P]),
    "o3d_cuda_particles_clear_body": (c_int, [c_void_p, c_void_p]),
    "o3d_cuda_particles_set_body_strengths": (c_int, [c_void_p, c_void_p, _P, _P, _P, _P]),
    "o3d_cuda_particles_body_vels": (c_int, [c_void_p, c_void_p, _P, _P, _P]),
    "o3d_cuda_particles_clear_inner": (c_int, [c_void_p, c_void_p, POINTER(c_int64)]),
    "o3d_cuda_particles_body_counters": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int)]),
    "o3d_cuda_particles_totals": (c_int, [c_void_p, c_void_p, POINTER(c_double), POINTER(c_double)]),
    "o3d_cuda_status_open": (c_int, [c_char_p, c_int, POINTER(c_void_p)]),
    "o3d_cuda_status_close": (None, [c_void_p]),
    "o3d_cuda_status_reset_sim": (c_int, [c_void_p]),
    "o3d_cuda_status_append_float": (c_int, [c_void_p, c_char_p, ctypes.c_float]),
    "o3d_cuda_status_append_int": (c_int, [c_void_p, c_char_p, c_int]),
    "o3d_cuda_status_write_line": (c_int, [c_void_p]),
    "o3d_cuda_particles_write_status": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double]),
    "o3d_cuda_bem_op_create": (c_int, [c_void_p, c_int64, _P, _P, _P, c_int64, _P, _P, _P, _P,
                                       c_int64, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, c_int, POINTER(c_void_p)]),
    "o3d_cuda_bem_op_apply": (c_int, [c_void_p, c_void_p, _P, _P, POINTER(c_double)]),
    "o3d_cuda_bem_op_destroy": (None, [c_void_p, c_void_p]),
    "o3d_cuda_reflect_pts": (c_int, [c_void_p, c_int64, _P, _P, _P, c_int64, _P, _P, c_int64, _P, _P, _P, POINTER(c_int64)]),
    "o3d_cuda_clear_inner_pts": (c_int, [c_void_p, c_int, c_int64, _P, _P, _P, c_int64, _P, _P, c_int64, _P, _P, _P,
                                         ctypes.c_float, ctypes.c_float, POINTER(c_int64)]),
    "o3d_cuda_write_points_vtu": (c_int, [c_char_p, c_int64] + [_P] * 10 + [c_double]),
    "o3d_cuda_particles_write_vtu": (c_int, [c_void_p, c_void_p, c_char_p, c_double]),
    "o3d_cuda_set_graphs": (c_int, [c_void_p, c_int]),
    "o3d_cuda_particles_graph_active": (c_int, [c_void_p]),
    "o3d_cuda_set_tuned_kernels": (c_int, [c_void_p, c_int]),
    "o3d_cuda_tuned_kernels": (c_int, [c_void_p]),
    "o3d_cuda_plan_pts_on_pts": (c_int, [c_int, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "o3d_cuda_plan_check": (c_int, [c_int, c_int64, c_int64, c_int]),
    "o3d_cuda_set_host_staging": (c_int, [c_void_p, c_int]),
    "o3d_cuda_set_panel_queue": (c_int, [c_void_p, c_int]),
    "o3d_cuda_set_core_func": (c_int, [c_void_p, c_int]),
    "o3d_cuda_core_func": (c_int, [c_void_p]),
    "o3d_cuda_set_profiling": (c_int, [c_void_p, c_int]),
    "o3d_cuda_dev_kernel_ms": (c_int, [c_void_p, POINTER(c_double)]),
    "o3d_cuda_probe_fp32_peak": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", CSRC] + ([] if verbose else ["-s"]), check=True)
    return LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()). "
            "omega3d_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
