"""TEST INFRASTRUCTURE: ctypes front end of the plain-C restatement ``oracle/ann_oracle.c``
(built by ``make -C oracle port`` or ``__graft_entry__.build()`` into ``oracle/_build``)."""
import ctypes as C
import os
import subprocess
import time

import numpy as np

from ._common import HERE, LAYER_CODES, NetArrays, as_double_p, as_int_p, c_double_p, c_int_p, prepare_io

LIB_PATH = os.path.join(HERE, "_build", "libann_oracle.so")


class _Net(C.Structure):
    _fields_ = [("num_layers", C.c_int), ("num_of_nodes", c_int_p), ("layer_types", c_int_p),
                ("coeff", C.POINTER(c_double_p)), ("bias", C.POINTER(c_double_p)),
                ("potential_center", c_double_p), ("num_center", C.c_int),
                ("force_constant", C.c_double), ("scaling_factor", C.c_double),
                ("data_type_in_input_layer", C.c_int),
                ("index_of_backbone_atoms", c_int_p), ("num_backbone", C.c_int),
                ("dihedrals", c_int_p), ("num_dihedrals", C.c_int),
                ("pairs", c_int_p), ("num_pairs", C.c_int)]


_lib_handle = None


def build(force=False):
    src = os.path.join(HERE, "ann_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH


def _lib():
    global _lib_handle
    if _lib_handle is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.ann_oracle_eval_batch.restype = C.c_int
        lib.ann_oracle_eval_batch.argtypes = [C.POINTER(_Net), C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]
        lib.ann_oracle_forward_backward.restype = C.c_int
        lib.ann_oracle_forward_backward.argtypes = [C.POINTER(_Net), c_double_p, c_double_p, c_double_p]
        lib.ann_oracle_cos_sin_four_atoms.restype = None
        lib.ann_oracle_cos_sin_four_atoms.argtypes = [c_double_p] * 6
        _lib_handle = lib
    return _lib_handle


def _pack(force):
    arr = NetArrays(force)
    arr.type_codes = np.ascontiguousarray([LAYER_CODES[t] for t in arr.types], dtype=np.int32)
    net = _Net(arr.L, as_int_p(arr.nodes), as_int_p(arr.type_codes), arr.coeff_ptrs, arr.bias_ptrs,
               as_double_p(arr.center), int(arr.center.size), arr.k, arr.scale, arr.input_type,
               as_int_p(arr.atoms), int(arr.atoms.size),
               as_int_p(arr.dihedrals0), int(arr.dihedrals0.shape[0]),
               as_int_p(arr.pairs0), int(arr.pairs0.shape[0]))
    return arr, net


def evaluate(force, positions, centers=None, return_seconds=False):
    """Same contract as oracle.reference.evaluate (single-threaded)."""
    arr, net = _pack(force)
    pos, cen, single, R, N = prepare_io(positions, centers, arr.center.size)
    forces = np.zeros_like(pos)
    energies = np.zeros(R, dtype=np.float64)
    t0 = time.perf_counter()
    rc = _lib().ann_oracle_eval_batch(C.byref(net), N, R, as_double_p(pos), as_double_p(cen) if cen is not None else None,
                                      as_double_p(forces), as_double_p(energies))
    secs = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError("ann_oracle_eval_batch failed (%d)" % rc)
    out = (float(energies[0]), forces[0]) if single else (energies, forces)
    return out + (secs,) if return_seconds else out


def forward_backward(force, x):
    arr, net = _pack(force)
    x = np.ascontiguousarray(x, dtype=np.float64)
    total = int(arr.nodes.sum())
    outs = np.zeros(total)
    ders = np.zeros(total)
    _lib().ann_oracle_forward_backward(C.byref(net), as_double_p(x), as_double_p(outs), as_double_p(ders))
    cuts = np.cumsum(arr.nodes)[:-1]
    return np.split(outs, cuts), np.split(ders, cuts)


def cos_sin_four_atoms(xyz4):
    xyz = np.ascontiguousarray(xyz4, dtype=np.float64).reshape(4, 3)
    c, s = C.c_double(), C.c_double()
    _lib().ann_oracle_cos_sin_four_atoms(as_double_p(xyz[0]), as_double_p(xyz[1]), as_double_p(xyz[2]), as_double_p(xyz[3]),
                                         C.byref(c), C.byref(s))
    return c.value, s.value
