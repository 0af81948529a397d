"""Pins the CPU oracles (test infrastructure) before anything is compared with them.

  1. the reference's own known-answer vectors (openmmapi/tests/test_ANN_package.cpp),
  2. golden outputs of the reference's unmodified kernel (tests/golden/*.npz),
  3. oracle/_ref itself on random networks, when that library is present,
  4. finite differences — the check the reference's own integration tests make (:175-180).
"""
import math

import numpy as np
import pytest

import cases
import golden_io
from ann_force_b200 import synthetic
from oracle import port
from oracle import reference as ref

ORACLES = [pytest.param(port, id="port")] + (
    [pytest.param(ref, id="reference")] if ref.available(3) else [])

TOL = 1e-5          # test_ANN_package.cpp:20
FORCE_TOL = 5e-2    # test_ANN_package.cpp:21


@pytest.mark.parametrize("oracle", ORACLES)
def test_forward_kat_linear_tanh(oracle):
    """test_ANN_package.cpp:98-133: layer outputs {1,2} {0.06,0.13,0.20} {tanh(1.08)}."""
    outs, ders = oracle.forward_backward(cases.kat_2_3_1(), [1, 2])
    expected = [[1, 2], [0.06, 0.13, 0.20], [math.tanh(1.08)]]
    for got, want in zip(outs, expected):
        np.testing.assert_allclose(got, want, atol=TOL, rtol=0)
    # back_prop of the same network, measured on the unmodified reference
    want_d = [[-0.647121507, -0.823609191], [-2.9414614, -5.8829228, -8.82438419], [7.93199097]]
    for got, want in zip(ders, want_d):
        np.testing.assert_allclose(got, want, rtol=2e-8)


@pytest.mark.parametrize("oracle", ORACLES)
def test_forward_kat_circular(oracle):
    """test_ANN_package.cpp:135-173: Circular output {0.447214, 0.894427}."""
    outs, ders = oracle.forward_backward(cases.kat_2_3_2_circular(), [1, 2])
    np.testing.assert_allclose(outs[1], [0.06, 0.13, 0.20], atol=TOL, rtol=0)
    np.testing.assert_allclose(outs[2], [0.447214, 0.894427], atol=TOL, rtol=0)
    want_d = [[1.46143631, 1.86000985], [6.64289231, 13.2857846, 19.9286769], [-12.378299, 0]]
    for got, want in zip(ders, want_d):
        np.testing.assert_allclose(got, want, rtol=2e-8, atol=1e-12)


@pytest.mark.parametrize("oracle", ORACLES)
def test_dihedral_cos_sin_kat(oracle):
    """test_sincos_of_dihedrals_four_atom (test_ANN_package.cpp:80-95) expects cos=0, sin=-1, but it is
    never called from main (:676-705) and the reference's own kernel returns sin=+1 for these four
    points (n1 x n2 = (1,0,0), middle bond = (1,0,0): ANN_ReferenceKernels.cpp:612).  The oracle is
    pinned to what the reference computes, and the discrepancy is recorded here."""
    c, s = oracle.cos_sin_four_atoms([[0, 1, 0], [0, 0, 0], [1, 0, 0], [0, 0, 1]])
    assert abs(c) < TOL and abs(s - 1.0) < TOL


GOLDEN_ENERGIES = {   # SURVEY.md §8c, produced from the unmodified reference
    "ref_cartesian_12_4_4_tanh": 11.421839171580899,
    "ref_cartesian_12_4_4_softmax": 1.4608776635136895,
    "ref_cartesian_12_4_4_linear": 26.790165804903218,
    "ref_cartesian_12_4_4_circular": 16.572144252757681,
    "ref_pairdist_6_3_3": 10.832122286301122,
    "ref_dihedral_4_4_4_tanh": 0.5707411857967678,
    "ref_dihedral_4_4_4_circ_00": 34.049960896473365,
    "ref_dihedral_4_4_4_circ_2p4_2p3": 50.906053293232567,
}


def test_golden_files_hold_the_survey_energies():
    for name, want in GOLDEN_ENERGIES.items():
        _, _, _, energies, forces, _ = golden_io.load(name)
        assert energies[0] == pytest.approx(want, rel=1e-14)
    _, _, _, _, forces, _ = golden_io.load("ref_cartesian_12_4_4_tanh")
    np.testing.assert_allclose(forces[0, :, 0], [0.266341673478, -0.681947964612, 0.165733598395, 0.249872692739], rtol=1e-11)
    np.testing.assert_allclose(forces[0, :, 0], forces[0, :, 1], rtol=1e-14)


@pytest.mark.parametrize("name", golden_io.names())
def test_port_matches_golden(name):
    force, pos, centers, energies, forces, _ = golden_io.load(name)
    e, f = port.evaluate(force, pos, centers)
    np.testing.assert_allclose(e, energies, rtol=1e-12)
    scale = np.abs(forces).max()
    assert np.abs(f - forces).max() <= 1e-11 * scale


@pytest.mark.skipif(not ref.available(3), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", golden_io.names())
def test_reference_library_reproduces_golden(name):
    force, pos, centers, energies, forces, _ = golden_io.load(name)
    e, f = ref.evaluate(force, pos, centers)
    np.testing.assert_array_equal(e, energies)
    np.testing.assert_array_equal(f, forces)


@pytest.mark.skipif(not (ref.available(3) and ref.available(4)), reason="oracle/_ref not built")
@pytest.mark.parametrize("nodes,types", [
    ([21, 40, 2], ["Tanh", "Tanh"]),
    ([21, 40, 4], ["Tanh", "Circular"]),
    ([30, 17, 6], ["Linear", "Softmax"]),
    ([180, 100, 100, 3], ["Tanh", "Tanh", "Tanh"]),
    ([60, 32, 8, 4], ["Tanh", "Circular", "Circular"]),
])
def test_port_matches_reference_on_random_networks(nodes, types):
    force = synthetic.make_force(nodes, types, seed=7)
    pos = synthetic.make_positions(nodes[0] // 3 + 5, 6, seed=8).astype(np.float64)
    cen = synthetic.make_centers(force, 6, seed=9).astype(np.float64)
    e_ref, f_ref = ref.evaluate(force, pos, cen)
    e, f = port.evaluate(force, pos, cen)
    np.testing.assert_allclose(e, e_ref, rtol=1e-12)
    assert np.abs(f - f_ref).max() <= 1e-11 * np.abs(f_ref).max()
    # atoms outside the selection feel nothing
    assert np.all(f[:, nodes[0] // 3:, :] == 0.0)


@pytest.mark.skipif(not ref.available(3), reason="oracle/_ref not built")
def test_reference_threads_agree():
    force = synthetic.make_config("c2")
    pos = synthetic.make_positions(7, 33).astype(np.float64)
    e1, f1 = ref.evaluate(force, pos, threads=1)
    e4, f4 = ref.evaluate(force, pos, threads=4)
    np.testing.assert_array_equal(e1, e4)
    np.testing.assert_array_equal(f1, f4)


@pytest.mark.parametrize("oracle", ORACLES)
@pytest.mark.parametrize("builder", [
    lambda: cases.cartesian_12_4_4("Tanh"), lambda: cases.cartesian_12_4_4("Softmax"),
    lambda: cases.cartesian_12_4_4("Circular"), cases.pair_distance_6_3_3,
    lambda: cases.dihedral_4_4_4("Tanh"), lambda: cases.dihedral_4_4_4("Circular", [0, 0]),
    lambda: cases.dihedral_4_4_4("Circular", [2.4, 2.3]), lambda: cases.size_sweep(20),
], ids=["cart_tanh", "cart_softmax", "cart_circular", "pairdist", "dih_tanh", "dih_circ00", "dih_circ2p4", "sweep20"])
def test_forces_are_minus_energy_gradient(oracle, builder):
    """The reference's integration tests: analytic force vs finite differences (test_ANN_package.cpp:175-180),
    here with central differences (h = 1e-6) at a tolerance four orders tighter than its FORCE_TOL = 5e-2.
    (The reference uses forward differences with delta = 0.005, whose truncation error alone is 6.5e-2 on the
    Circular {2.4, 2.3} case — the central form is the stricter and the meaningful check.)"""
    force, pos = builder()
    _, analytic = oracle.evaluate(force, pos)
    numeric = cases.finite_difference_forces(oracle.evaluate, force, pos, h=1e-6)
    assert np.abs(analytic - numeric).max() <= 1e-6 * max(1.0, np.abs(analytic).max())
