"""Loads tests/golden/*.npz (made by tests/golden/make_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np

from ann_force_b200 import ANN_Force

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    """Returns (force, positions[R,N,3] f64, centers[R,C] f64 or None, energies[R], forces[R,N,3], meta)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    f = ANN_Force()
    f.set_num_of_nodes(meta["num_of_nodes"])
    f.set_layer_types(meta["layer_types"])
    n_conn = len(meta["num_of_nodes"]) - 1
    f.set_coeffients_of_connections([z["coeff%d" % l] for l in range(n_conn)])
    f.set_values_of_biased_nodes([z["bias%d" % l] for l in range(n_conn)])
    f.set_potential_center(meta["potential_center"])
    f.set_force_constant(meta["force_constant"])
    f.set_scaling_factor(meta["scaling_factor"])
    f.set_data_type_in_input_layer(meta["data_type_in_input_layer"])
    f.set_index_of_backbone_atoms(meta["index_of_backbone_atoms"])
    # the setters take 1-based indices; the fixtures store what the getters return (0-based)
    f.set_list_of_index_of_atoms_forming_dihedrals([[i + 1 for i in q] for q in meta["dihedrals0"]])
    f.set_list_of_pair_index_for_distances([[i + 1 for i in p] for p in meta["pairs0"]])
    centers = z["centers"] if "centers" in z.files else None
    return f, z["positions"], centers, z["energies"], z["forces"], meta
