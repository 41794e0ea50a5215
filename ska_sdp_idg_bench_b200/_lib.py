"""Loads libidgb200.so (built in-tree by csrc/Makefile) and declares the C ABI of
include/idg_b200.h for ctypes.  Missing library => ImportError (no fallback)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libidgb200.so")


class Params(C.Structure):  # idgb200_params
    _fields_ = [
        ("nr_subgrids", C.c_int32), ("grid_size", C.c_int32), ("subgrid_size", C.c_int32),
        ("image_size", C.c_float), ("w_step_in_lambda", C.c_float), ("nr_channels", C.c_int32),
        ("nr_stations", C.c_int32), ("sincos_mode", C.c_int32), ("variant", C.c_int32),
        ("flags", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class Perf(C.Structure):  # idgb200_perf
    _fields_ = [("seconds", C.c_double), ("gflops", C.c_double), ("gbytes", C.c_double),
                ("mvis", C.c_double), ("nr_subgrids", C.c_int32), ("iterations", C.c_int32),
                ("joules", C.c_double)]


# every symbol include/idg_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "idgb200_version": (C.c_int, []),
    "idgb200_error_string": (C.c_char_p, [C.c_int]),
    "idgb200_print_device_info": (C.c_int, []),
    "idgb200_device_name": (C.c_int, [C.c_char_p, C.c_size_t]),
    "idgb200_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "idgb200_flops_gridder": (C.c_uint64, [C.c_uint64] * 5),
    "idgb200_bytes_gridder": (C.c_uint64, [C.c_uint64] * 5),
    "idgb200_gridder": (C.c_int, [C.POINTER(Params)] + [_P] * 8),
    "idgb200_degridder": (C.c_int, [C.POINTER(Params)] + [_P] * 8),
    "idgb200_launch_count": (C.c_uint64, []),
    "idgb200_resolve_variant": (C.c_int, [C.POINTER(Params), C.c_int]),
    "idgb200_adder": (C.c_int, [C.POINTER(Params), _P, _P, C.POINTER(C.c_void_p), C.c_int, C.c_int, _P]),
    "idgb200_splitter": (C.c_int, [C.POINTER(Params), _P, _P, C.POINTER(C.c_void_p), C.c_int, C.c_int, _P]),
    "idgb200_reduce_parts": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_int64, _P, _P]),
    "idgb200_subgrid_fft": (C.c_int, [C.c_int64, C.c_int, C.c_int, _P, _P]),
    "idgb200_adder_rs_mode": (C.c_int, [C.c_int64, C.c_int, C.c_int]),
    "idgb200_c_run_gridder": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                        C.c_int, C.c_int64, C.c_int] + [_P] * 7),
    "idgb200_c_run_degridder": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                          C.c_int, C.c_int64, C.c_int] + [_P] * 7),
    "idgb200_c_run_gridder_ex": (C.c_int, [C.POINTER(Params), C.c_int64, C.c_int] + [_P] * 7),
    "idgb200_c_run_degridder_ex": (C.c_int, [C.POINTER(Params), C.c_int64, C.c_int] + [_P] * 7),
    "idgb200_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "idgb200_host_free": (C.c_int, [_P]),
    "idgb200_report": (None, [C.c_char_p] + [C.c_double] * 5),
    "idgb200_report_csv": (None, [C.c_char_p, C.c_char_p, C.c_char_p] + [C.c_double] * 5),
    "idgb200_p_run_gridder": (C.c_int, [C.POINTER(Perf)]),
    "idgb200_p_run_degridder": (C.c_int, [C.POINTER(Perf)]),
    "idgb200_init_uvw": (C.c_int, [C.c_uint32, C.c_int64, C.c_int, C.c_uint32, _P, _P]),
    "idgb200_init_wavenumbers": (C.c_int, [C.c_int, _P, _P]),
    "idgb200_init_visibilities": (C.c_int, [C.c_uint32, C.c_float, C.c_int64, C.c_int, _P, _P, _P]),
    "idgb200_init_spheroidal": (C.c_int, [C.c_int, _P, _P]),
    "idgb200_init_aterms": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint32, _P, _P]),
    "idgb200_init_metadata": (C.c_int, [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                        _P, _P]),
    "idgb200_init_subgrids": (C.c_int, [C.c_int64, C.c_int, _P, _P]),
}


def load(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `make -C ska_sdp_idg_bench_b200/csrc` "
            "(or __graft_entry__.build()).  There is no fallback implementation.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        f = getattr(lib, name)  # AttributeError if the library does not export it
        f.restype = res
        f.argtypes = args
    return lib


lib = load()
