"""Seeded synthetic workloads of the shapes BASELINE.json names.

Used by bench.py, __graft_entry__.smoke() and the tests so that every implementation — the CUDA
path, the CPU oracle, the reference arm — sees identical networks and identical fp32-rounded
positions.  Pure numpy; nothing here touches the GPU.

Recipe: weights Xavier-uniform U(+-sqrt(6/(n_in+n_out))), biases 0.1*U(-1,1),
seed 1234; positions N(0, 0.5 nm) per coordinate rounded to fp32; k = 500, scaling factor 0.5;
harmonic centres 0.3, angular centres uniform in (-pi, pi); per-replica centres jitter around those.
"""
from __future__ import annotations

import numpy as np

from .force import ANN_Force, INPUT_CARTESIAN

# name -> (layer sizes, layer types, selected atoms, replicas per GPU at the headline setting)
CONFIGS = {
    "c1": dict(nodes=[21, 40, 2], types=["Tanh", "Tanh"], atoms=7, replicas=1,
               label="alanine dipeptide vacuum, 7 backbone heavy atoms, 21-40-2 Tanh encoder, 1 replica"),
    "c2": dict(nodes=[21, 40, 4], types=["Tanh", "Circular"], atoms=7, replicas=1,
               label="alanine dipeptide, 21-40-4 Circular-output encoder, 1 replica, OpenMM CUDA buffers"),
    "c3": dict(nodes=[21, 40, 4], types=["Tanh", "Circular"], atoms=7, replicas=256,
               label="alanine dipeptide umbrella sampling, 256 windows per launch, 21-40-4 Circular"),
    "c4": dict(nodes=[180, 100, 100, 3], types=["Tanh", "Tanh", "Tanh"], atoms=60, replicas=1024,
               label="Trp-cage backbone (60 atoms), 180-100-100-3 Tanh encoder, 1024 walkers"),
    "c5": dict(nodes=[4500, 512, 512, 8], types=["Tanh", "Tanh", "Tanh"], atoms=1500, replicas=1024,
               label="synthetic wide selection, 1500 atoms, 4500-512-512-8 encoder, 1024 replicas per GPU (8192 on 8)"),
}


def flops_per_eval(nodes):
    """4*M: one multiply-add = 2 flops, forward + input-gradient backward."""
    return 4 * sum(int(a) * int(b) for a, b in zip(nodes[:-1], nodes[1:]))


def bytes_per_eval(num_atoms_selected, num_cv):
    """24*A + 4*n_cv + 8: 3 fp32 coords in, 3 fp32 forces out per atom, centres in, one fp64 energy out."""
    return 24 * int(num_atoms_selected) + 4 * int(num_cv) + 8


def make_force(nodes, types, atoms=None, seed=1234, force_constant=500.0, scaling_factor=0.5,
               atom_index=None, total_atoms=None) -> ANN_Force:
    """Cartesian-input ANN_Force with Xavier-uniform weights.  `atom_index` (1-based) defaults to 1..A."""
    rng = np.random.default_rng(seed)
    nodes = [int(n) for n in nodes]
    if atoms is None:
        atoms = nodes[0] // 3
    assert 3 * atoms == nodes[0]
    coeff, bias = [], []
    for n_in, n_out in zip(nodes[:-1], nodes[1:]):
        bound = np.sqrt(6.0 / (n_in + n_out))
        # round weights to fp32-representable values the way a network trained in fp32 would be
        coeff.append(rng.uniform(-bound, bound, size=n_in * n_out).astype(np.float32).astype(np.float64))
        bias.append((0.1 * rng.uniform(-1.0, 1.0, size=n_out)).astype(np.float32).astype(np.float64))
    n_last = nodes[-1]
    if types[-1] == "Circular":
        center = rng.uniform(-np.pi, np.pi, size=n_last // 2)
    else:
        center = np.full(n_last, 0.3)
    f = ANN_Force()
    f.set_num_of_nodes(nodes)
    f.set_layer_types(types)
    f.set_coeffients_of_connections(coeff)
    f.set_values_of_biased_nodes(bias)
    f.set_potential_center(center)
    f.set_force_constant(force_constant)
    f.set_scaling_factor(scaling_factor)
    f.set_data_type_in_input_layer(INPUT_CARTESIAN)
    if atom_index is None:
        atom_index = list(range(1, atoms + 1))
    f.set_index_of_backbone_atoms(atom_index)
    f.validate()
    return f


def make_config(name: str, seed=1234) -> ANN_Force:
    cfg = CONFIGS[name]
    return make_force(cfg["nodes"], cfg["types"], cfg["atoms"], seed=seed)


def make_positions(num_atoms: int, replicas: int, seed=4321) -> np.ndarray:
    """[R, num_atoms, 3] float32, N(0, 0.5 nm) per coordinate."""
    rng = np.random.default_rng(seed)
    return (0.5 * rng.standard_normal(size=(replicas, num_atoms, 3))).astype(np.float32)


def make_centers(force: ANN_Force, replicas: int, seed=777) -> np.ndarray:
    """Per-replica umbrella centres [R, n_center] float32: the force's centre plus a window-dependent shift."""
    rng = np.random.default_rng(seed)
    base = np.asarray(force.get_potential_center(), dtype=np.float64)
    if force.get_layer_types()[-1] == "Circular":
        cen = rng.uniform(-np.pi, np.pi, size=(replicas, base.size))
    else:
        cen = base[None, :] + rng.uniform(-0.5, 0.5, size=(replicas, base.size))
    return cen.astype(np.float32)
