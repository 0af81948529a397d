"""Host-side mirror of the reference's user-facing parameter object.

``ANN_Force`` here exposes exactly the setters/getters the reference's C++ class and its SWIG
wrapper expose (openmmapi/include/openmm/ANN_Force.h:57-101, python/ANN.i:47-91) — same names
(including the historical spelling ``coeffients``), same argument meaning, same 1-based atom
conventions — so scripts written against the SWIG module ``ANN`` only change their import.

Differences, all deliberate (SURVEY.md §7 "API generalisation"):

* depth comes from the arguments, not from the compile-time ``NUM_OF_LAYERS`` macro
  (ANN_Force.h:36); the reference's setters overflow or truncate when given another depth
  (ANN_Force.cpp:13-18,46-52,58-64);
* ``validate()`` raises ``ValueError`` where the reference ``assert``s at kernel initialisation
  (ANN_ReferenceKernels.cpp:61-66,74) or silently prints (``:235-238``).

The object is pure host data; ``ann_force_b200.kernel.CalcANNForce`` turns it into a device
handle through the C-ABI (include/annforce_b200.h).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

LAYER_TYPES = ("Linear", "Tanh", "Circular", "Softmax")

INPUT_COS_SIN_OF_DIHEDRALS = 0   # the reference's default (ANN_Force.h:137)
INPUT_CARTESIAN = 1
INPUT_PAIR_DISTANCES = 2


def _floats(values):
    """Lists stay lists (like the SWIG vectors); numpy arrays are kept as flat float64 arrays so that
    multi-million-weight layers do not become Python float lists."""
    if isinstance(values, np.ndarray):
        return np.ascontiguousarray(values, dtype=np.float64).ravel()
    return [float(v) for v in values]


class ANN_Force:
    def __init__(self):
        self._num_of_nodes: List[int] = []
        self._index_of_backbone_atoms: List[int] = []
        self._dihedrals: List[List[int]] = []        # stored 0-based, like the reference
        self._pairs: List[List[int]] = []            # stored 0-based, like the reference
        self._coeff: List[List[float]] = []
        self._layer_types: List[str] = []
        self._bias: List[List[float]] = []
        self._potential_center: List[float] = []
        self._force_constant: float = 0.0
        self._scaling_factor: float = 1.0
        self._data_type_in_input_layer: int = INPUT_COS_SIN_OF_DIHEDRALS
        self._force_group: int = 0

    # ---- OpenMM::Force surface used by ANN_ForceImpl.cpp:26-30 -------------------------------
    def getForceGroup(self) -> int:
        return self._force_group

    def setForceGroup(self, group: int) -> None:
        self._force_group = int(group)

    def usesPeriodicBoundaryConditions(self) -> bool:      # ANN_Force.h:119-121
        return False

    # ---- network shape ----------------------------------------------------------------------
    def get_num_of_nodes(self) -> List[int]:
        return self._num_of_nodes

    def set_num_of_nodes(self, num: Sequence[int]) -> None:
        self._num_of_nodes = [int(n) for n in num]

    def get_layer_types(self) -> List[str]:
        return self._layer_types

    def set_layer_types(self, temp_layer_types: Sequence[str]) -> None:
        self._layer_types = [str(t) for t in temp_layer_types]

    def get_coeffients_of_connections(self) -> List[List[float]]:
        return self._coeff

    def set_coeffients_of_connections(self, coefficients: Sequence[Sequence[float]]) -> None:
        """One flat list per connection, row-major [n_out][n_in] (ANN_ReferenceKernels.cpp:81-83)."""
        self._coeff = [_floats(layer) for layer in coefficients]

    def get_values_of_biased_nodes(self) -> List[List[float]]:
        return self._bias

    def set_values_of_biased_nodes(self, bias: Sequence[Sequence[float]]) -> None:
        self._bias = [_floats(layer) for layer in bias]

    # ---- umbrella potential -----------------------------------------------------------------
    def get_potential_center(self) -> List[float]:
        return self._potential_center

    def set_potential_center(self, temp_potential_center: Sequence[float]) -> None:
        self._potential_center = _floats(temp_potential_center)

    def get_force_constant(self) -> float:
        return self._force_constant

    def set_force_constant(self, temp_force_constant: float) -> None:
        self._force_constant = float(temp_force_constant)

    def get_scaling_factor(self) -> float:
        return self._scaling_factor

    def set_scaling_factor(self, temp_scaling_factor: float) -> None:
        self._scaling_factor = float(temp_scaling_factor)

    # ---- input features ---------------------------------------------------------------------
    def get_data_type_in_input_layer(self) -> int:
        return self._data_type_in_input_layer

    def set_data_type_in_input_layer(self, temp_data_type_in_input_layer: int) -> None:
        self._data_type_in_input_layer = int(temp_data_type_in_input_layer)

    def get_index_of_backbone_atoms(self) -> List[int]:
        return self._index_of_backbone_atoms

    def set_index_of_backbone_atoms(self, indices: Sequence[int]) -> None:
        """1-based (PDB-style) atom numbers; kept 1-based, as the reference keeps them (ANN_ReferenceKernels.cpp:132)."""
        self._index_of_backbone_atoms = [int(i) for i in indices]

    def get_list_of_index_of_atoms_forming_dihedrals(self) -> List[List[int]]:
        return self._dihedrals

    def set_list_of_index_of_atoms_forming_dihedrals(self, temp_list_of_index: Sequence[Sequence[int]]) -> None:
        """Appends (the reference appends too, ANN_Force.cpp:105-115) 1-based quadruples, stored 0-based."""
        for quad in temp_list_of_index:
            self._dihedrals.append([int(quad[j]) - 1 for j in range(4)])

    def set_list_of_index_of_atoms_forming_dihedrals_from_index_of_backbone_atoms(self, index_of_backbone_atoms: Sequence[int]) -> None:
        """phi/psi quadruples of a backbone given as N, CA, C triples per residue (ANN_Force.cpp:118-154)."""
        idx = [int(i) for i in index_of_backbone_atoms]
        num_of_residues = len(idx) // 3
        for res in range(num_of_residues):
            if res != 0:                       # phi: C(i-1) N CA C
                self._dihedrals.append([idx[3 * res - 1] - 1, idx[3 * res] - 1, idx[3 * res + 1] - 1, idx[3 * res + 2] - 1])
            if res != num_of_residues - 1:     # psi: N CA C N(i+1)
                self._dihedrals.append([idx[3 * res] - 1, idx[3 * res + 1] - 1, idx[3 * res + 2] - 1, idx[3 * res + 3] - 1])

    def get_list_of_pair_index_for_distances(self) -> List[List[int]]:
        return self._pairs

    def set_list_of_pair_index_for_distances(self, temp_list_of_index: Sequence[Sequence[int]]) -> None:
        for pair in temp_list_of_index:
            self._pairs.append([int(pair[0]) - 1, int(pair[1]) - 1])

    # ---- checks the reference leaves to assert() ----------------------------------------------
    def num_layers(self) -> int:
        return len(self._num_of_nodes)

    def validate(self) -> None:
        L = self.num_layers()
        if L < 2:
            raise ValueError("set_num_of_nodes: need at least an input and an output layer")
        if any(n <= 0 for n in self._num_of_nodes):
            raise ValueError("set_num_of_nodes: layer sizes must be positive")
        if len(self._layer_types) != L - 1 or len(self._coeff) != L - 1 or len(self._bias) != L - 1:
            raise ValueError("layer_types / coefficients / biases must each have num_layers-1 entries")
        for l in range(L - 1):
            n_in, n_out = self._num_of_nodes[l], self._num_of_nodes[l + 1]
            if self._layer_types[l] not in LAYER_TYPES:
                raise ValueError("unknown layer type %r (expected one of %s)" % (self._layer_types[l], ", ".join(LAYER_TYPES)))
            if len(self._coeff[l]) != n_in * n_out:          # ANN_ReferenceKernels.cpp:74
                raise ValueError("connection %d: expected %d coefficients, got %d" % (l, n_in * n_out, len(self._coeff[l])))
            if len(self._bias[l]) != n_out:
                raise ValueError("connection %d: expected %d biases, got %d" % (l, n_out, len(self._bias[l])))
            if self._layer_types[l] == "Circular" and n_out % 2 != 0:   # ANN_ReferenceKernels.cpp:211
                raise ValueError("Circular layer %d needs an even number of nodes" % (l + 1))
        n_last = self._num_of_nodes[-1]
        want = n_last // 2 if self._layer_types[-1] == "Circular" else n_last   # ANN_ReferenceKernels.cpp:61-66
        if len(self._potential_center) != want:
            raise ValueError("potential_center: expected %d values, got %d" % (want, len(self._potential_center)))
        if self._scaling_factor == 0.0:
            raise ValueError("scaling_factor must be non-zero")
        t = self._data_type_in_input_layer
        if t == INPUT_CARTESIAN:
            if 3 * len(self._index_of_backbone_atoms) != self._num_of_nodes[0]:
                raise ValueError("Cartesian input: 3*len(index_of_backbone_atoms) must equal num_of_nodes[0]")
            if any(i < 1 for i in self._index_of_backbone_atoms):
                raise ValueError("index_of_backbone_atoms is 1-based")
            # the reference accumulates per list entry (ANN_ReferenceKernels.cpp:401); the batched kernels write one force
            # row per selected atom, so a repeated atom is rejected instead of being evaluated two different ways
            if len(set(self._index_of_backbone_atoms)) != len(self._index_of_backbone_atoms):
                raise ValueError("index_of_backbone_atoms lists an atom more than once")
        elif t == INPUT_COS_SIN_OF_DIHEDRALS:
            if 2 * len(self._dihedrals) != self._num_of_nodes[0]:
                raise ValueError("dihedral input: 2*number of dihedrals must equal num_of_nodes[0]")
        elif t == INPUT_PAIR_DISTANCES:
            if len(self._pairs) != self._num_of_nodes[0]:
                raise ValueError("pair-distance input: number of pairs must equal num_of_nodes[0]")
        else:
            raise ValueError("data_type_in_input_layer must be 0, 1 or 2")
