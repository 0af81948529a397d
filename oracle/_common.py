"""Shared ctypes plumbing for the two oracles (test infrastructure)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LAYER_CODES = {"Linear": 0, "Tanh": 1, "Circular": 2, "Softmax": 3}

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


def as_double_p(a):
    return a.ctypes.data_as(c_double_p)


def as_int_p(a):
    return a.ctypes.data_as(c_int_p)


class NetArrays:
    """Flattens an ANN_Force-like object (anything with the reference's getter names,
    openmmapi/include/openmm/ANN_Force.h:57-101) into contiguous arrays that outlive the call."""

    def __init__(self, force):
        self.nodes = np.ascontiguousarray(force.get_num_of_nodes(), dtype=np.int32)
        self.L = int(self.nodes.size)
        self.types = list(force.get_layer_types())
        coeff = force.get_coeffients_of_connections()
        bias = force.get_values_of_biased_nodes()
        assert len(self.types) == self.L - 1 and len(coeff) == self.L - 1 and len(bias) == self.L - 1
        self.coeff = [np.ascontiguousarray(c, dtype=np.float64).ravel() for c in coeff]
        self.bias = [np.ascontiguousarray(b, dtype=np.float64).ravel() for b in bias]
        for l in range(self.L - 1):
            assert self.coeff[l].size == int(self.nodes[l]) * int(self.nodes[l + 1]), "coefficient size mismatch"
            assert self.bias[l].size == int(self.nodes[l + 1]), "bias size mismatch"
        self.center = np.ascontiguousarray(force.get_potential_center(), dtype=np.float64)
        self.k = float(force.get_force_constant())
        self.scale = float(force.get_scaling_factor())
        self.input_type = int(force.get_data_type_in_input_layer())
        self.atoms = np.ascontiguousarray(force.get_index_of_backbone_atoms(), dtype=np.int32)     # 1-based
        # the getters return what the setters stored, i.e. 0-based (ANN_Force.cpp:111,164)
        self.dihedrals0 = np.ascontiguousarray(force.get_list_of_index_of_atoms_forming_dihedrals(),
                                               dtype=np.int32).reshape(-1, 4)
        self.pairs0 = np.ascontiguousarray(force.get_list_of_pair_index_for_distances(), dtype=np.int32).reshape(-1, 2)
        self.coeff_ptrs = (c_double_p * (self.L - 1))(*[as_double_p(c) for c in self.coeff])
        self.bias_ptrs = (c_double_p * (self.L - 1))(*[as_double_p(b) for b in self.bias])


def prepare_io(positions, centers, num_center):
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    single = pos.ndim == 2
    if single:
        pos = pos[None]
    R, N, three = pos.shape
    assert three == 3
    cen = None
    if centers is not None:
        cen = np.ascontiguousarray(centers, dtype=np.float64).reshape(R, num_center)
    return pos, cen, single, R, N
