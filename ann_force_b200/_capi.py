"""ctypes binding of include/annforce_b200.h.  Loading the library needs no GPU; every compute entry
fails loudly (RuntimeError) without one — there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ANNF_B200_LIB selects a diagnostic build of the same library (e.g. one made with EXTRA=-DANNF_PHASE_CLOCKS); there is still
# no fallback: whatever this names must exist and export every symbol
LIB_PATH = os.environ.get("ANNF_B200_LIB") or os.path.join(HERE, "libannforce_b200.so")

LINEAR, TANH, CIRCULAR, SOFTMAX = 0, 1, 2, 3
LAYER_CODES = {"Linear": LINEAR, "Tanh": TANH, "Circular": CIRCULAR, "Softmax": SOFTMAX}
PREC_AUTO, PREC_FP64, PREC_FP32, PREC_TF32X3, PREC_F16X3 = 0, 1, 2, 3, 4
PRECISIONS = {"auto": PREC_AUTO, "fp64": PREC_FP64, "fp32": PREC_FP32, "tf32x3": PREC_TF32X3, "f16x3": PREC_F16X3}
PRECISION_NAMES = {v: k for k, v in PRECISIONS.items()}
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = 0, 1, 2, 3, 4
FLAG_POSQ_DOUBLE, FLAG_ENERGY_DOUBLE, FLAG_SKIP_FORCES, FLAG_SKIP_ENERGY = 1, 2, 4, 8

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)


class Desc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("num_layers", C.c_int32),
                ("num_of_nodes", c_int_p), ("layer_types", c_int_p),
                ("coeff", C.POINTER(c_double_p)), ("bias", C.POINTER(c_double_p)),
                ("potential_center", c_double_p), ("num_center", C.c_int32),
                ("force_constant", C.c_double), ("scaling_factor", C.c_double),
                ("data_type_in_input_layer", C.c_int32),
                ("index_of_backbone_atoms", c_int_p), ("num_backbone_atoms", C.c_int32),
                ("dihedrals", c_int_p), ("num_dihedrals", C.c_int32),
                ("pairs", c_int_p), ("num_pairs", C.c_int32),
                ("num_system_atoms", C.c_int32), ("precision", C.c_int32)]


class Info(C.Structure):
    _fields_ = [("precision_single", C.c_int32), ("precision_batched", C.c_int32), ("batched_path", C.c_int32),
                ("cluster_size", C.c_int32), ("macs_per_eval", C.c_int64), ("kernel_launches", C.c_int64),
                ("sm_count", C.c_int32), ("single_path", C.c_int32)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("total_ms", C.c_double)]


# every symbol include/annforce_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "annf_version": (C.c_char_p, []),
    "annf_last_error": (C.c_char_p, []),
    "annf_device_count": (C.c_int, []),
    "annf_create": (C.c_int, [C.POINTER(Desc), C.POINTER(C.c_void_p)]),
    "annf_destroy": (C.c_int, [C.c_void_p]),
    "annf_get_info": (C.c_int, [C.c_void_p, C.POINTER(Info)]),
    "annf_update_parameters": (C.c_int, [C.c_void_p, c_double_p, C.c_int32, C.c_double, C.c_void_p]),
    "annf_reserve": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "annf_set_atom_order": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "annf_eval_openmm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32, C.c_void_p]),
    "annf_eval_batched": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annf_eval_batched_host": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annf_debug_single_phases": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "annf_debug_stream_phases": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "annf_debug_tensor_phases": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "annf_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "annf_get_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(KernelTime), C.c_int32, C.POINTER(C.c_int32), C.c_int32]),
}

_lib = None


class ANNForceError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def lib():
    """The loaded C-ABI library.  Raises if it has not been built — there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ANNForceError(ERR_NO_DEVICE, "%s is missing: build it with `make -C ann_force_b200/csrc` "
                                "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status):
    if status != OK:
        raise ANNForceError(status, lib().annf_last_error().decode())
