"""ctypes binding of the C ABI in include/rdc.h (librdcgpu.so).

This is the same binding a C++ caller gets by including rdc.h; Python is used here only because the test
and bench harnesses are Python.  There is no fallback: if the CUDA library cannot be loaded the import of
the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None


class RdcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rdc error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("ms_assemble", C.c_double), ("ms_solve", C.c_double), ("ms_clamp", C.c_double),
                ("ms_spmv_total", C.c_double), ("n_spmv", C.c_int), ("iterations", C.c_int),
                ("resnorm", C.c_double), ("resnorm0", C.c_double),
                ("n_nodes_local", C.c_int64), ("n_nodes_ghost", C.c_int64), ("n_elems_local", C.c_int64),
                ("nnzb_local", C.c_int64), ("bytes_assemble", C.c_int64), ("bytes_spmv", C.c_int64),
                ("bytes_index", C.c_int64), ("kernel_launches", C.c_int64), ("ripf_rt_total_max", C.c_int),
                ("sum_ms_assemble", C.c_double), ("sum_ms_solve", C.c_double), ("sum_ms_clamp", C.c_double),
                ("sum_ms_spmv", C.c_double), ("sum_iterations", C.c_int64), ("sum_n_spmv", C.c_int64),
                ("n_solves", C.c_int64), ("p2p_on", C.c_int), ("p2p_fused", C.c_int), ("bicg_persistent", C.c_int)]


EXPORTS = ["rdc_model_nvars", "rdc_model_nparams", "rdc_create", "rdc_create_distributed", "rdc_comm_unique_id",
           "rdc_destroy", "rdc_last_error", "rdc_set_params", "rdc_set_elem_field", "rdc_set_nodal_field",
           "rdc_update_coords", "rdc_set_solution", "rdc_get_solution", "rdc_get_solution_owned", "rdc_get_old_solution", "rdc_get_rhs", "rdc_n_dofs",
           "rdc_set_time", "rdc_set_dt", "rdc_rotate", "rdc_assemble", "rdc_solve", "rdc_clamp", "rdc_step",
           "rdc_spmv", "rdc_bench_spmv", "rdc_bench_stream", "rdc_bench_barrier", "rdc_bench_dfma", "rdc_download_csr", "rdc_free", "rdc_get_stats", "rdc_set_stream",
           "rdc_version", "rdc_probe_partition", "rdc_set_option", "rdc_set_subdomains", "rdc_region_volumes",
           "rdc_region_last_mean", "rdc_probe_spmv_tiles", "rdc_probe_region_chunks",
           "rdc_solid_set_reference", "rdc_solid_set_materials", "rdc_solid_set_fibres", "rdc_solid_set_symmetry", "rdc_solid_set_bcs", "rdc_solid_assemble",
           "rdc_solid_newton", "rdc_solid_post_process", "rdc_solid_probe_row", "rdc_solid_probe_bc_row", "rdc_solid_probe_post", "rdc_solid_probe_bc_rows"]


def load():
    """Load librdcgpu.so, (re)building it in-tree when nvcc and the sources are present."""
    global _LIB
    if _LIB is not None:
        return _LIB
    try:
        so = _build.build()
    except Exception as exc:  # noqa: BLE001
        if os.path.exists(_build.SO):
            so = _build.SO
        else:
            raise RuntimeError(f"librdcgpu.so is missing and could not be built: {exc}") from exc
    L = C.CDLL(so)
    for name in EXPORTS:
        getattr(L, name)  # AttributeError if the library does not export what rdc.h declares
    L.rdc_last_error.restype = C.c_char_p
    L.rdc_last_error.argtypes = [C.c_void_p]
    L.rdc_version.restype = C.c_char_p
    L.rdc_n_dofs.restype = C.c_int64
    L.rdc_n_dofs.argtypes = [C.c_void_p]
    L.rdc_destroy.restype = None
    L.rdc_destroy.argtypes = [C.c_void_p]
    L.rdc_free.restype = None
    L.rdc_free.argtypes = [C.c_void_p]
    vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "rdc_model_nvars": [i32], "rdc_model_nparams": [i32],
        "rdc_create": [C.POINTER(vp), i32, i32, i64, i64, vp, vp, vp, i32],
        "rdc_create_distributed": [C.POINTER(vp), i32, i32, i64, i64, vp, vp, vp, i32, i32, i32, i32, vp],
        "rdc_comm_unique_id": [vp],
        "rdc_set_params": [vp, vp, i32], "rdc_set_elem_field": [vp, i32, vp, i32],
        "rdc_set_nodal_field": [vp, i32, vp, i32], "rdc_update_coords": [vp, vp],
        "rdc_set_solution": [vp, vp], "rdc_get_solution": [vp, vp], "rdc_get_solution_owned": [vp, vp], "rdc_get_old_solution": [vp, vp], "rdc_get_rhs": [vp, vp],
        "rdc_set_time": [vp, f64], "rdc_set_dt": [vp, f64], "rdc_rotate": [vp], "rdc_assemble": [vp, f64, f64],
        "rdc_solve": [vp, i32, i32, f64, i32, i32, C.POINTER(i32), C.POINTER(f64)],
        "rdc_clamp": [vp],
        "rdc_step": [vp, f64, f64, i32, i32, f64, i32, i32, C.POINTER(i32), C.POINTER(f64)],
        "rdc_spmv": [vp, vp, vp], "rdc_bench_spmv": [vp, i32, C.POINTER(f64)],
        "rdc_bench_stream": [vp, i32, i32, C.POINTER(f64), C.POINTER(i64)],
        "rdc_bench_barrier": [vp, i32, i32, i32, C.POINTER(f64)], "rdc_bench_dfma": [vp, C.POINTER(f64)],
        "rdc_download_csr": [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                             C.POINTER(vp), C.POINTER(vp)],
        "rdc_get_stats": [vp, C.POINTER(Stats)], "rdc_set_stream": [vp, vp], "rdc_set_option": [vp, C.c_char_p, i32], "rdc_set_subdomains": [vp, vp, i32],
        "rdc_region_volumes": [vp, i32, vp, vp], "rdc_region_last_mean": [vp, i32, vp],
        "rdc_solid_set_reference": [vp, vp], "rdc_solid_set_materials": [vp, i32, vp, vp], "rdc_solid_set_fibres": [vp, vp], "rdc_solid_set_symmetry": [vp, i32],
        "rdc_solid_set_bcs": [vp, i32, vp, i64, vp, vp, vp, f64], "rdc_solid_assemble": [vp, f64],
        "rdc_solid_newton": [vp, f64, vp, i32, vp], "rdc_solid_post_process": [vp, f64, vp, vp, vp],
        "rdc_solid_probe_row": [i32, vp, vp, vp, f64, vp, i32, i32, vp, vp],
        "rdc_solid_probe_bc_row": [i32, vp, vp, vp, f64, f64, i32, vp, vp],
        "rdc_solid_probe_post": [i32, vp, vp, vp, f64, vp, vp],
        "rdc_solid_probe_bc_rows": [i32, i64, i64, vp, vp, i32, i32, i32, i64, vp, vp, C.POINTER(i32), C.POINTER(vp), C.POINTER(vp),
                                    C.POINTER(vp), C.POINTER(vp)],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _LIB = L
    return L
