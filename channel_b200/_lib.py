"""ctypes binding of libchannel_b200.so (the C ABI in include/channel_b200.h).

The product path has no CPU fallback: if the CUDA library is missing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libchannel_b200.so")

c_double_p = C.POINTER(C.c_double)


class HostTables(C.Structure):
    _fields_ = [("y", c_double_p), ("d0", c_double_p), ("d1", c_double_p), ("d2", c_double_p),
                ("d4", c_double_p), ("D0mat", c_double_p)] + [
        (n, C.c_double * 5) for n in ("d140", "d14m1", "d240", "d24m1", "d14n", "d14np1", "d24n", "d24np1",
                                      "v0bc", "v0m1bc", "vnbc", "vnp1bc", "eta0bc", "eta0m1bc", "etanbc", "etanp1bc")]


# every symbol include/channel_b200.h and include/channel_b200_host.h declare
SYMBOLS = {
    "chb_last_error": (C.c_char_p, []),
    "chb_version": (C.c_int, []),
    "chb_get_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "chb_create": (C.c_int, [C.POINTER(C.c_void_p)] + [C.c_int] * 5 + [C.c_double] * 6 + [C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "chb_destroy": (C.c_int, [C.c_void_p]),
    "chb_get_decomposition": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 4),
    "chb_set_tables": (C.c_int, [C.c_void_p] + [c_double_p] * 22),
    "chb_upload_V": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chb_download_V": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chb_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "chb_host_unregister": (C.c_int, [C.c_void_p]),
    "chb_upload_V_planes": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chb_download_V_planes": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chb_set_wall_velocity": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "chb_set_forcing": (C.c_int, [C.c_void_p] + [C.c_double] * 4 + [C.c_int, C.c_int, C.c_double]),
    "chb_cfl_prepass": (C.c_int, [C.c_void_p]),
    "chb_set_body_force_linear": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, C.c_int]),
    "chb_set_body_force_linear_yz": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_int]),
    "chb_upload_F": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chb_download_F": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chb_set_body_force": (C.c_int, [C.c_void_p]),
    "chb_buildrhs": (C.c_int, [C.c_void_p, c_double_p, C.c_double, C.c_int]),
    "chb_linsolve": (C.c_int, [C.c_void_p, C.c_double]),
    "chb_vetaTOuvw": (C.c_int, [C.c_void_p]),
    "chb_computeflowrate": (C.c_int, [C.c_void_p, C.c_double]),
    "chb_rk3_step": (C.c_int, [C.c_void_p, C.c_double]),
    "chb_get_step_scalars": (C.c_int, [C.c_void_p] + [c_double_p] * 10),
    "chb_download_rhs": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chb_debug_capture_products": (C.c_int, [C.c_void_p, C.c_int]),
    "chb_download_products": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chb_download_F_planes": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chb_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "chb_timing_report": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, c_double_p, C.POINTER(C.c_longlong), C.c_int]),
    "chb_launch_count": (C.c_longlong, [C.c_void_p]),
    "chb_sync": (C.c_int, [C.c_void_p]),
    "chb_stopwatch_begin": (C.c_int, [C.c_void_p]),
    "chb_stopwatch_end": (C.c_int, [C.c_void_p, c_double_p]),
    "chb_get_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "chb_device_bytes": (C.c_longlong, [C.c_void_p]),
    "chb_measure_device_peaks": (C.c_int, [c_double_p]),
    "chb_test_fft_lines": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "chb_save_restart_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double, C.c_int, C.c_int]),
    "chb_restart_wait": (C.c_int, [C.c_void_p]),
    "chb_restart_stats": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "chb_read_restart_file": (C.c_int, [C.c_void_p, C.c_char_p, c_double_p]),
    "chb_set_convvel": (C.c_int, [C.c_void_p, C.c_int]),
    "chb_get_convvel": (C.c_int, [C.c_void_p, c_double_p, C.POINTER(C.c_longlong)]),
    "chb_save_convvel_file": (C.c_int, [C.c_void_p, C.c_char_p]),
    "chb_host_restart_header": (C.c_int, [C.c_int] * 3 + [C.c_double] * 7 + [C.c_void_p]),
    "chb_host_restart_offset": (C.c_longlong, [C.c_int] * 5),
    "chb_host_restart_file_bytes": (C.c_longlong, [C.c_int] * 3),
    "chb_host_fft_fit": (C.c_int, [C.c_int]),
    "chb_host_padded_sizes": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "chb_host_setup_tables": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(HostTables)]),
    "chb_host_decomposition": (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_int)] * 4),
    "chb_host_transpose_index": (C.c_longlong, [C.c_int] * 9),
    "chb_host_apply_tables": (C.c_int, [C.c_void_p, C.POINTER(HostTables)]),
}

_lib = None


class ChannelB200Error(RuntimeError):
    pass


def load():
    """dlopen the library and declare every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ChannelB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().chb_last_error()
        raise ChannelB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
