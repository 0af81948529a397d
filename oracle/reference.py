"""TEST INFRASTRUCTURE: ctypes front end of ``oracle/_ref/libannref_L<depth>.so`` — the reference's
own CPU kernel (platforms/reference/ANN_ReferenceKernels.cpp) compiled unmodified by oracle/Makefile."""
import ctypes as C
import os

import numpy as np

from ._common import HERE, NetArrays, as_double_p, as_int_p, c_double_p, c_int_p, prepare_io

REF_DIR = os.path.join(HERE, "_ref")


class _Net(C.Structure):
    _fields_ = [("num_layers", C.c_int), ("num_of_nodes", c_int_p),
                ("layer_types", C.POINTER(C.c_char_p)),
                ("coeff", C.POINTER(c_double_p)), ("bias", C.POINTER(c_double_p)),
                ("potential_center", c_double_p), ("num_center", C.c_int),
                ("force_constant", C.c_double), ("scaling_factor", C.c_double),
                ("data_type_in_input_layer", C.c_int),
                ("index_of_backbone_atoms", c_int_p), ("num_backbone", C.c_int),
                ("dihedrals", c_int_p), ("num_dihedrals", C.c_int),
                ("pairs", c_int_p), ("num_pairs", C.c_int)]


_libs = {}


def lib_path(depth):
    return os.path.join(REF_DIR, "libannref_L%d.so" % depth)


def available(depth=3):
    return os.path.exists(lib_path(depth))


def _lib(depth):
    if depth not in _libs:
        if not available(depth):
            raise FileNotFoundError("%s missing: run `make -C oracle ref` where /root/reference exists" % lib_path(depth))
        lib = C.CDLL(lib_path(depth))
        lib.annref_num_layers.restype = C.c_int
        assert lib.annref_num_layers() == depth
        lib.annref_eval_batch.restype = C.c_int
        lib.annref_eval_batch.argtypes = [C.POINTER(_Net), C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                          c_double_p, C.c_int, c_double_p]
        lib.annref_forward_backward.restype = C.c_int
        lib.annref_forward_backward.argtypes = [C.POINTER(_Net), c_double_p, c_double_p, c_double_p]
        lib.annref_cos_sin_four_atoms.restype = None
        lib.annref_cos_sin_four_atoms.argtypes = [c_double_p, c_double_p, c_double_p]
        _libs[depth] = lib
    return _libs[depth]


def _pack(force):
    arr = NetArrays(force)
    # the reference's setters want 1-based dihedral / pair indices
    arr.dihedrals1 = np.ascontiguousarray(arr.dihedrals0 + 1, dtype=np.int32)
    arr.pairs1 = np.ascontiguousarray(arr.pairs0 + 1, dtype=np.int32)
    arr.type_strs = (C.c_char_p * (arr.L - 1))(*[t.encode() for t in arr.types])
    net = _Net(arr.L, as_int_p(arr.nodes), arr.type_strs, arr.coeff_ptrs, arr.bias_ptrs,
               as_double_p(arr.center), int(arr.center.size), arr.k, arr.scale, arr.input_type,
               as_int_p(arr.atoms), int(arr.atoms.size),
               as_int_p(arr.dihedrals1), int(arr.dihedrals1.shape[0]),
               as_int_p(arr.pairs1), int(arr.pairs1.shape[0]))
    return arr, net


def evaluate(force, positions, centers=None, threads=1, return_seconds=False):
    """positions [N,3] or [R,N,3] (nm).  Returns (energy, forces) with matching leading shape; forces are
    THIS force's contribution only.  With return_seconds also the wall time of the evaluation loop."""
    arr, net = _pack(force)
    pos, cen, single, R, N = prepare_io(positions, centers, arr.center.size)
    forces = np.zeros_like(pos)
    energies = np.zeros(R, dtype=np.float64)
    secs = C.c_double(0.0)
    rc = _lib(arr.L).annref_eval_batch(C.byref(net), N, R, as_double_p(pos), as_double_p(cen) if cen is not None else None,
                                      as_double_p(forces), as_double_p(energies), int(threads), C.byref(secs))
    if rc != 0:
        raise RuntimeError("annref_eval_batch failed (%d)" % rc)
    out = (float(energies[0]), forces[0]) if single else (energies, forces)
    return out + (secs.value,) if return_seconds else out


def forward_backward(force, x):
    """Layer outputs and back-propagated derivatives for an explicit input vector
    (what test_ANN_package.cpp:98-173 exercises).  Returns two lists of per-layer arrays."""
    arr, net = _pack(force)
    x = np.ascontiguousarray(x, dtype=np.float64)
    total = int(arr.nodes.sum())
    outs = np.zeros(total)
    ders = np.zeros(total)
    rc = _lib(arr.L).annref_forward_backward(C.byref(net), as_double_p(x), as_double_p(outs), as_double_p(ders))
    if rc != 0:
        raise RuntimeError("annref_forward_backward failed (%d)" % rc)
    cuts = np.cumsum(arr.nodes)[:-1]
    return np.split(outs, cuts), np.split(ders, cuts)


def cos_sin_four_atoms(xyz4):
    xyz = np.ascontiguousarray(xyz4, dtype=np.float64).reshape(12)
    c, s = C.c_double(), C.c_double()
    _lib(3).annref_cos_sin_four_atoms(as_double_p(xyz), C.byref(c), C.byref(s))
    return c.value, s.value
