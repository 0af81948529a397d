"""Networks and coordinates of the reference's own tests (openmmapi/tests/test_ANN_package.cpp),
rebuilt through the ANN_Force setters, plus the reference's glibc-rand() size sweep."""
import ctypes

import numpy as np

from ann_force_b200 import ANN_Force


def kat_2_3_1():
    """test_forward_and_backward_prop, test_ANN_package.cpp:98-133.  Input {1,2}."""
    f = ANN_Force()
    f.set_num_of_nodes([2, 3, 1])
    f.set_coeffients_of_connections([[0.01, 0.02, 0.03, 0.04, 0.05, 0.06], [-1, -2, -3]])
    f.set_values_of_biased_nodes([[0.01, 0.02, 0.03], [2]])
    f.set_layer_types(["Linear", "Tanh"])
    f.set_potential_center([0])
    f.set_force_constant(10)
    return f


def kat_2_3_2_circular():
    """test_forward_and_backward_prop_2, test_ANN_package.cpp:135-173.  Input {1,2}."""
    f = ANN_Force()
    f.set_num_of_nodes([2, 3, 2])
    f.set_coeffients_of_connections([[0.01, 0.02, 0.03, 0.04, 0.05, 0.06], [-1, -2, -3, 1, 2, 3]])
    f.set_values_of_biased_nodes([[0.01, 0.02, 0.03], [1.92, 1.08]])
    f.set_layer_types(["Linear", "Circular"])
    f.set_potential_center([0])
    f.set_force_constant(10)
    return f


_W44 = [[1, 0, 0, 0, 0, 1, 0, 0.7, 0, 0, 1, 0, 0, 0, 0, 1],
        [1, 0, 0.4, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]]
_B44 = [[0.1, 0.2, 0.3, 0.4], [0.5, 0.6, 0.4, 0.3]]


def cartesian_12_4_4(last_activation="Tanh"):
    """..._for_input_as_Cartesian_coordinates, test_ANN_package.cpp:422-502."""
    f = ANN_Force()
    f.set_num_of_nodes([12, 4, 4])
    f.set_layer_types(["Tanh", last_activation])
    w0 = np.zeros((4, 12))
    for j in range(4):
        w0[j, 3 * j:3 * j + 3] = 1.0
    f.set_coeffients_of_connections([w0.ravel().tolist(), _W44[1]])
    f.set_force_constant(10)
    f.set_scaling_factor(1)
    f.set_index_of_backbone_atoms([1, 2, 3, 4])
    f.set_potential_center([0, 0] if last_activation == "Circular" else [0, 0, 0, 0])
    f.set_values_of_biased_nodes(_B44)
    f.set_data_type_in_input_layer(1)
    pos = np.array([[-1, -2, -3], [0, 0, 0], [1, 0, 0], [0, 0, 2]], dtype=np.float64)
    return f, pos


def dihedral_4_4_4(last_activation="Tanh", center=None):
    """..._by_comparing_with_numerical_derivatives[_for_circular_layer], test_ANN_package.cpp:182-341."""
    f = ANN_Force()
    f.set_num_of_nodes([4, 4, 4])
    f.set_layer_types(["Tanh", last_activation])
    f.set_coeffients_of_connections(_W44)
    f.set_force_constant(10)
    if center is None:
        center = [0, 0] if last_activation == "Circular" else [0, 0, 0, 0]
    f.set_potential_center(center)
    f.set_values_of_biased_nodes(_B44)
    f.set_list_of_index_of_atoms_forming_dihedrals_from_index_of_backbone_atoms([1, 2, 3, 4, 5, 6])
    pos = np.array([[-1, -2, -3], [0, 0, 0], [1, 0, 0], [0, 0, 1], [0.5, 0, 0], [0, 0.3, 0.6]], dtype=np.float64)
    return f, pos


def pair_distance_6_3_3():
    """..._for_input_as_pairwise_distances, test_ANN_package.cpp:596-674."""
    f = ANN_Force()
    f.set_index_of_backbone_atoms([1, 2, 3, 4])
    f.set_num_of_nodes([6, 3, 3])
    f.set_layer_types(["Tanh", "Tanh"])
    f.set_list_of_pair_index_for_distances([[1, 2], [1, 3], [1, 4], [2, 3], [2, 4], [3, 4]])
    f.set_coeffients_of_connections([[1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0.7, 0, 0, 0, 0, 1, 1],
                                     [1, 0, 0.4, 0, 1, 0, 0, 0, 1]])
    f.set_force_constant(10)
    f.set_scaling_factor(10)
    f.set_potential_center([0, 0, 0])
    f.set_values_of_biased_nodes([[0.1, 0.2, 0.3], [0.5, 0.6, 0.4]])
    f.set_data_type_in_input_layer(2)
    pos = np.array([[-1, -2, -3], [0, 0, 0], [1, 0, 0], [0, 0, 2]], dtype=np.float64)
    return f, pos


def alanine_dipeptide_8_15_2(reference_test_source="/root/reference/openmmapi/tests/test_ANN_package.cpp"):
    """..._for_alanine_dipeptide, test_ANN_package.cpp:343-420: 7 backbone atoms, four dihedrals -> 8-15-2 Tanh/Tanh with
    TRAINED weights.  The 167 trained numbers are read from the reference's test source (build container only; the golden
    fixture generated from this case carries them to the GPU box), everything else is restated here."""
    import re
    text = open(reference_test_source).read()
    body = text[text.index("for_alanine_dipeptide() {"):]
    body = body[:body.index("system.addForce")]
    braces = re.findall(r"\{([-+0-9.eE,\s]+)\}", body)
    lists = [[float(v) for v in b.replace("\n", " ").split(",") if v.strip()] for b in braces]
    by_len = {len(v): v for v in lists}
    f = ANN_Force()
    f.set_num_of_nodes([8, 15, 2])
    f.set_layer_types(["Tanh", "Tanh"])
    f.set_coeffients_of_connections([by_len[120], by_len[30]])
    f.set_force_constant(10)
    f.set_potential_center([0.6, 0.5])
    f.set_values_of_biased_nodes([by_len[15], [v for v in lists if len(v) == 2 and v != [0.6, 0.5]][0]])
    f.set_list_of_index_of_atoms_forming_dihedrals([[1, 2, 3, 4], [2, 3, 4, 5], [3, 4, 5, 6], [4, 5, 6, 7]])
    pos = np.array([[-1, -2, -3], [0, 0, 0], [1.5, 0, 0], [0, 0.57, 1], [0.5, 0, 0.1], [0, 0.3, 0.6], [0, 0.4, 0.5]], dtype=np.float64)
    return f, pos


def _glibc():
    return ctypes.CDLL("libc.so.6")


def size_sweep(num_of_atoms, seed=1):
    """..._Cartesian_coordinates_larger_system, test_ANN_package.cpp:504-594: 3N-150-4 Tanh/Tanh with
    weights (rand() % 10)/100 re-seeded per array and positions (rand() % 100)/20 continuing the last stream."""
    libc = _glibc()

    def random_array(size):
        libc.srand(ctypes.c_uint(seed))
        return [(libc.rand() % 10) / 100.0 for _ in range(size)]

    n0 = 3 * num_of_atoms
    f = ANN_Force()
    f.set_num_of_nodes([n0, 150, 4])
    f.set_layer_types(["Tanh", "Tanh"])
    coeff = [random_array(n0 * 150), random_array(150 * 4)]
    f.set_coeffients_of_connections(coeff)
    f.set_force_constant(10)
    f.set_scaling_factor(1)
    f.set_index_of_backbone_atoms(list(range(1, num_of_atoms + 1)))
    f.set_potential_center([0, 0, 0, 0])
    f.set_values_of_biased_nodes([random_array(150), random_array(4)])
    f.set_data_type_in_input_layer(1)
    pos = np.array([[(libc.rand() % 100) / 20.0 for _ in range(3)] for _ in range(num_of_atoms)], dtype=np.float64)
    return f, pos


def finite_difference_forces(evaluate, force, pos, h=1e-6):
    """Central differences of the energy: -dE/dx for every coordinate."""
    out = np.zeros_like(pos)
    for a in range(pos.shape[0]):
        for c in range(3):
            p = pos.copy(); p[a, c] += h
            m = pos.copy(); m[a, c] -= h
            out[a, c] = -(evaluate(force, p)[0] - evaluate(force, m)[0]) / (2 * h)
    return out
