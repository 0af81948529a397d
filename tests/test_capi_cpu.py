"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header declares,
and refuses to compute without a GPU (no fallback).  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from ann_force_b200 import ANN_Force, _capi, synthetic
from ann_force_b200.kernel import CalcANNForce

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "annforce_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"ANNF_API[^;(]*?\b(annf_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 13
    assert sorted(_capi.SYMBOLS) == names


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", _capi.LIB_PATH]).decode()
    exported = set(re.findall(r" T (annf_\w+)", out))
    assert exported == set(declared_symbols())
    assert b"sm_100a" in lib.annf_version()


def test_library_contains_only_sm100a_code():
    out = subprocess.run(["cuobjdump", "--list-elf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_descriptor_layout_matches_header():
    # compile a probe against the header and compare sizeof / offsets with the ctypes mirror
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "annforce_b200.h"
    int main() { printf("%zu %zu %zu %zu %zu %zu\n", sizeof(annf_desc), offsetof(annf_desc, potential_center),
                 offsetof(annf_desc, force_constant), offsetof(annf_desc, precision), sizeof(annf_info), sizeof(annf_kernel_time)); return 0; }
    '''
    exe = os.path.join(ROOT, "oracle", "_build", "abi_probe")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    got = [int(v) for v in subprocess.check_output([exe]).split()]
    want = [C.sizeof(_capi.Desc), _capi.Desc.potential_center.offset, _capi.Desc.force_constant.offset,
            _capi.Desc.precision.offset, C.sizeof(_capi.Info), C.sizeof(_capi.KernelTime)]
    assert got == want


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _capi.lib().annf_device_count() == 0
    kernel = CalcANNForce()
    with pytest.raises(_capi.ANNForceError):
        kernel.initialize(7, synthetic.make_config("c2"))
    # and straight through the C ABI
    force = synthetic.make_config("c2")
    nodes = np.asarray(force.get_num_of_nodes(), dtype=np.int32)
    types = np.asarray([1, 2], dtype=np.int32)
    coeff = [np.asarray(w, dtype=np.float64) for w in force.get_coeffients_of_connections()]
    bias = [np.asarray(b, dtype=np.float64) for b in force.get_values_of_biased_nodes()]
    center = np.asarray(force.get_potential_center(), dtype=np.float64)
    atoms = np.asarray(force.get_index_of_backbone_atoms(), dtype=np.int32)
    d = lambda a: a.ctypes.data_as(_capi.c_double_p)
    i = lambda a: a.ctypes.data_as(_capi.c_int_p)
    desc = _capi.Desc(C.sizeof(_capi.Desc), -1, 3, i(nodes), i(types), (_capi.c_double_p * 2)(*map(d, coeff)),
                      (_capi.c_double_p * 2)(*map(d, bias)), d(center), 2, 500.0, 0.5, 1, i(atoms), 7, None, 0, None, 0, 7, 0)
    handle = C.c_void_p()
    assert _capi.lib().annf_create(C.byref(desc), C.byref(handle)) == _capi.ERR_NO_DEVICE
    assert b"no CPU fallback" in _capi.lib().annf_last_error()
    # argument errors are reported before any device is touched
    desc.num_center = 3
    assert _capi.lib().annf_create(C.byref(desc), C.byref(handle)) == _capi.ERR_INVALID
    desc.struct_size = 4
    assert _capi.lib().annf_create(C.byref(desc), C.byref(handle)) == _capi.ERR_INVALID
    # a selected atom listed twice: the reference would sum its two contributions (ANN_ReferenceKernels.cpp:401), the batched
    # kernels write one row per selected atom -> rejected up front, in the C ABI and in both validate() mirrors
    desc.struct_size = C.sizeof(_capi.Desc)
    desc.num_center = 2
    atoms[3] = atoms[5]
    assert _capi.lib().annf_create(C.byref(desc), C.byref(handle)) == _capi.ERR_INVALID
    assert b"more than once" in _capi.lib().annf_last_error()
    force.set_index_of_backbone_atoms([1, 2, 3, 4, 5, 6, 6])
    with pytest.raises(ValueError):
        force.validate()


def test_host_entry_rejects_buffers_it_would_misread():
    """execute_batched_host hands raw pointers to the C ABI: dtype, shape and contiguity are checked first (no GPU needed:
    the checks run before the library is called)."""
    kernel = CalcANNForce()
    kernel._handle = C.c_void_p(1)           # pretend-initialised: the argument checks come before any use of the handle
    kernel._num_center = 2
    try:
        good = np.zeros((4, 7, 3), dtype=np.float32)
        for bad in (good.astype(np.float64), good[:, ::2, :], np.zeros((4, 7, 2), dtype=np.float32)):
            with pytest.raises(ValueError):
                kernel.execute_batched_host(bad)
        with pytest.raises(ValueError):
            kernel.execute_batched_host(good, centers=np.zeros((4, 3), dtype=np.float32))
        with pytest.raises(ValueError):
            kernel.execute_batched_host(good, energies=np.zeros(4, dtype=np.float32))
    finally:
        kernel._handle = C.c_void_p()


def test_product_never_imports_the_oracle():
    for top in ("ann_force_b200", "include", "tools", "openmm_stub"):      # product, boundary, diagnostics, stand-in
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for name in files:
                if name.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    text = open(os.path.join(dirpath, name)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), name
                    assert "ann_oracle" not in text and "libannref" not in text, name


def test_force_mirror_validates_like_the_reference_asserts():
    f = synthetic.make_config("c2")
    f.validate()
    f.set_potential_center([0.0, 0.0, 0.0, 0.0])          # Circular output needs n/2 centres (ANN_ReferenceKernels.cpp:61-66)
    with pytest.raises(ValueError):
        f.validate()
    g = ANN_Force()
    g.set_num_of_nodes([4, 4, 4])
    g.set_layer_types(["Tanh", "Sigmoid"])
    g.set_coeffients_of_connections([[0.0] * 16, [0.0] * 16])
    g.set_values_of_biased_nodes([[0.0] * 4, [0.0] * 4])
    g.set_potential_center([0.0] * 4)
    with pytest.raises(ValueError):
        g.validate()
    # dihedral helper reproduces ANN_Force.cpp:118-154 for two residues
    h = ANN_Force()
    h.set_list_of_index_of_atoms_forming_dihedrals_from_index_of_backbone_atoms([1, 2, 3, 4, 5, 6])
    assert h.get_list_of_index_of_atoms_forming_dihedrals() == [[0, 1, 2, 3], [2, 3, 4, 5]]
    h.set_list_of_pair_index_for_distances([[1, 2], [3, 4]])
    assert h.get_list_of_pair_index_for_distances() == [[0, 1], [2, 3]]


def test_device_tanh_matches_long_double():
    """The short-dependency-chain fp64 tanh of the single-replica kernel (csrc/annf_math.cuh), host build, swept against
    tanhl: at most 4 ulp anywhere, exact at 0 and in saturation, sign-symmetric."""
    import json
    import subprocess
    cpp = os.path.join(ROOT, "tests", "cpp")
    subprocess.run(["make", "-C", cpp, os.path.join(cpp, "_build", "tanh_check")], check=True, capture_output=True)
    out = subprocess.run([os.path.join(cpp, "_build", "tanh_check")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert json.loads(out.stdout)["worst_ulp"] <= 4.0
