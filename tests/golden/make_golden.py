"""Regenerates tests/golden/*.npz by running the reference's UNMODIFIED CPU kernel (oracle/_ref,
built by `make -C oracle ref` from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture stores the full network + inputs + the reference's fp64 outputs, so the tests that read
them need neither /root/reference nor oracle/_ref (neither exists by default on the GPU box).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from ann_force_b200 import synthetic  # noqa: E402
from oracle import reference as ref  # noqa: E402


def dump(name, force, positions, centers=None, extra=None):
    pos = np.asarray(positions, dtype=np.float64)
    if pos.ndim == 2:
        pos = pos[None]
    energies, forces = ref.evaluate(force, pos, centers)
    meta = dict(
        num_of_nodes=list(force.get_num_of_nodes()), layer_types=list(force.get_layer_types()),
        force_constant=force.get_force_constant(), scaling_factor=force.get_scaling_factor(),
        data_type_in_input_layer=force.get_data_type_in_input_layer(),
        index_of_backbone_atoms=list(force.get_index_of_backbone_atoms()),
        dihedrals0=[list(q) for q in force.get_list_of_index_of_atoms_forming_dihedrals()],
        pairs0=[list(p) for p in force.get_list_of_pair_index_for_distances()],
        potential_center=[float(v) for v in force.get_potential_center()],
        source="oracle/_ref (reference platforms/reference/ANN_ReferenceKernels.cpp, unmodified, g++ -O2)",
    )
    if extra:
        meta.update(extra)
    arrays = dict(positions=pos, energies=energies, forces=forces, meta=np.array(json.dumps(meta)))
    if centers is not None:
        arrays["centers"] = np.asarray(centers, dtype=np.float64)
    for l, (w, b) in enumerate(zip(force.get_coeffients_of_connections(), force.get_values_of_biased_nodes())):
        arrays["coeff%d" % l] = np.asarray(w, dtype=np.float64)
        arrays["bias%d" % l] = np.asarray(b, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("%-28s R=%-4d E[0]=%.15g" % (name, pos.shape[0], energies[0]))


def main():
    # --- the reference's own test networks (test_ANN_package.cpp) ---
    for act in ["Tanh", "Softmax", "Linear", "Circular"]:
        f, pos = cases.cartesian_12_4_4(act)
        dump("ref_cartesian_12_4_4_%s" % act.lower(), f, pos)
    f, pos = cases.pair_distance_6_3_3()
    dump("ref_pairdist_6_3_3", f, pos)
    for tag, act, cen in [("tanh", "Tanh", None), ("circ_00", "Circular", [0, 0]), ("circ_2p4_2p3", "Circular", [2.4, 2.3])]:
        f, pos = cases.dihedral_4_4_4(act, cen)
        dump("ref_dihedral_4_4_4_%s" % tag, f, pos)
    # the full size sweep main() runs, 20 .. 180 atoms step 20 (test_ANN_package.cpp:689-694)
    for n in range(20, 200, 20):
        f, pos = cases.size_sweep(n, seed=1)
        dump("ref_sweep_%datoms" % n, f, pos)
    # alanine dipeptide, 8-15-2 with trained weights (test_ANN_package.cpp:343-420)
    f, pos = cases.alanine_dipeptide_8_15_2()
    dump("ref_alanine_dipeptide_8_15_2", f, pos)
    # --- BASELINE.json shapes, synthetic, fp32-rounded positions ---
    for name, R in [("c1", 8), ("c2", 8), ("c4", 4)]:
        cfg = synthetic.CONFIGS[name]
        f = synthetic.make_config(name)
        pos = synthetic.make_positions(cfg["atoms"], R).astype(np.float64)
        cen = synthetic.make_centers(f, R).astype(np.float64)
        dump("syn_%s" % name, f, pos, cen, extra=dict(config=name))
    # shuffled, non-contiguous atom list inside a larger system
    rng = np.random.default_rng(99)
    atom_index = (rng.permutation(40)[:7] + 1).tolist()
    f = synthetic.make_force([21, 40, 4], ["Tanh", "Circular"], 7, seed=5, atom_index=atom_index)
    dump("syn_shuffled_atoms", f, synthetic.make_positions(40, 4, seed=11).astype(np.float64))
    # 5-layer net with a Circular hidden layer and a Softmax hidden layer (generic back_prop branches)
    f = synthetic.make_force([12, 8, 6, 6, 3], ["Tanh", "Circular", "Softmax", "Linear"], 4, seed=21)
    dump("syn_mixed_depth5", f, synthetic.make_positions(4, 4, seed=12).astype(np.float64))
    # single-layer network (depth 2)
    f = synthetic.make_force([9, 2], ["Tanh"], 3, seed=22)
    dump("syn_depth2", f, synthetic.make_positions(3, 4, seed=13).astype(np.float64))


if __name__ == "__main__":
    main()
