"""Loader of libstrugepic_b200.so -- fails loudly, there is no fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SPIC_B200_LIBRARY selects another build of the same library, e.g. one made with
# `python -m strugepic_b200.build --user-w my_w.cu` (include/strugepic_user_w.h)
LIB_PATH = os.environ.get("SPIC_B200_LIBRARY") or os.path.join(HERE, "lib", "libstrugepic_b200.so")

_dp = C.POINTER(C.c_double)


class SpicConfig(C.Structure):
    _fields_ = [("n_cell", C.c_int32 * 3), ("periodic", C.c_int32 * 3), ("ng", C.c_int32), ("interp", C.c_int32),
                ("map4_mode", C.c_int32), ("engine", C.c_int32), ("device", C.c_int32), ("nranks", C.c_int32),
                ("rank", C.c_int32), ("reserved", C.c_int32 * 7)]


# every symbol include/strugepic_b200.h declares: name -> (restype, argtypes)
vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
SIGNATURES = {
    "spic_create": (i32, [C.POINTER(SpicConfig), C.POINTER(vp)]),
    "spic_destroy": (i32, [vp]),
    "spic_last_error": (C.c_char_p, [vp]),
    "spic_sync": (i32, [vp]),
    "spic_local_box": (i32, [vp, C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3)]),
    "spic_comm_unique_id": (i32, [vp]),
    "spic_comm_init": (i32, [vp, vp]),
    "spic_set_uniform_field": (i32, [vp, i32, _dp]),
    "spic_set_field": (i32, [vp, i32, _dp]),
    "spic_get_field": (i32, [vp, i32, _dp]),
    "spic_add_species": (i32, [vp, dbl, dbl, i64] + [_dp] * 6),
    "spic_load_uniform_plasma": (i32, [vp, dbl, dbl, C.c_int32, dbl, u64]),
    "spic_load_density_plasma": (i32, [vp, dbl, dbl, C.c_int32, C.c_int32, dbl, u64, C.POINTER(C.c_int32)]),
    "spic_num_species": (i32, [vp]),
    "spic_num_particles": (i32, [vp, i32, C.POINTER(i64)]),
    "spic_num_particles_global": (i32, [vp, i32, C.POINTER(i64)]),
    "spic_get_particles": (i32, [vp, i32] + [_dp] * 6),
    "spic_set_particles": (i32, [vp, i32, i64] + [_dp] * 6),
    "spic_theta_axis": (i32, [vp, i32, dbl]),
    "spic_theta_E": (i32, [vp, dbl]),
    "spic_theta_B": (i32, [vp, dbl]),
    "spic_source": (i32, [vp, i32, i32, dbl, dbl, dbl, dbl]),
    "spic_map": (i32, [vp, i32, dbl]),
    "spic_field_only_step": (i32, [vp, i32, i32, dbl, dbl, dbl, i32]),
    "spic_energy": (i32, [vp, _dp]),
    "spic_gauss_residual": (i32, [vp, _dp]),
    "spic_number_density": (i32, [vp, _dp]),
    "spic_plot_write": (i32, [vp, C.c_char_p]),
    "spic_checkpoint_write": (i32, [vp, C.c_char_p]),
    "spic_checkpoint_read": (i32, [vp, C.c_char_p]),
    "spic_W1": (dbl, [i32, dbl]),
    "spic_Wp": (dbl, [i32, dbl]),
    "spic_I_W1": (dbl, [i32, dbl, dbl]),
    "spic_I_Wp": (dbl, [i32, dbl, dbl]),
    "spic_interpolation_range": (i32, [i32]),
    "spic_tap_W1": (dbl, [i32, i32, dbl]),
    "spic_tap_Wp": (dbl, [i32, i32, dbl]),
    "spic_tap_IWp": (dbl, [i32, i32, dbl, dbl, i32]),
    "spic_launch_count": (i64, [vp]),
    "spic_kernel_time_ms": (i32, [vp, i32, _dp, C.POINTER(i64)]),
    "spic_kernel_times": (i32, [vp, i32, C.POINTER(C.c_double * 5), C.POINTER(C.c_int64 * 5)]),
    "spic_set_option": (i32, [vp, C.c_char_p, dbl]),
    "spic_stream": (vp, [vp]),
    "spic_probe_fp64_tflops": (i32, [i32, dbl, _dp]),
    "spic_probe_fp64_three_operand_tflops": (i32, [i32, dbl, _dp]),
    "spic_probe_fp64_immediate_tflops": (i32, [i32, dbl, _dp]),
}

_lib = None


def load():
    """dlopen the in-tree library and bind every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "strugepic_b200: %s is missing. Build it with `python -m strugepic_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
