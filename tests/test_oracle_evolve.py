"""CPU checks of oracle/evolve.py (the checker of the state-evolution / QSE / natural-gradient rows): rotations against
scipy.linalg.expm, the rotation form of a Trotter slice against the gate-by-gate restatement of the reference's circuit
(/root/reference/qat/fermion/trotterisation.py:85-150), QSE against its definition."""

import numpy as np
import scipy.linalg

from oracle import evolve, spin


def _state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return v / np.linalg.norm(v)


def test_rotation_is_the_matrix_exponential():
    n = 5
    rng = np.random.default_rng(0)
    psi = _state(n, 1)
    for _ in range(12):
        x, z = int(rng.integers(0, 1 << n)), int(rng.integers(0, 1 << n))
        theta = float(rng.uniform(-3, 3))
        want = scipy.linalg.expm(-1j * theta * evolve.pauli_matrix(n, x, z)) @ psi
        assert np.max(np.abs(evolve.rotate(n, x, z, theta, psi) - want)) < 1e-13


def test_trotter_slice_matches_the_reference_circuit_gate_by_gate():
    n = 4
    terms = [(0.7, "XYZ", [0, 2, 3]), (-0.3, "ZZ", [1, 2]), (1.1, "Y", [1]), (0.5, "XIXY", [0, 1, 2, 3]), (0.25, "Z", [3])]
    xs, zs, cs = [], [], []
    for c, op, qb in terms:
        x, z, ph = spin.string_to_packed(n, op, qb)
        xs.append(x)
        zs.append(z)
        cs.append(c)
        assert abs(ph * (-1j) ** spin.popcount(x & z) - 1) < 1e-15  # Y-form coefficient of a plain Pauli string is 1
    psi = _state(n, 2)
    got = evolve.trotter_slice(n, xs, zs, cs, 0.37, psi)
    want = evolve.trotter_slice_circuit(n, terms, 0.37, psi)
    assert np.max(np.abs(got - want)) < 1e-13


def test_qse_reduces_to_the_energy_and_overlaps():
    n = 3
    rng = np.random.default_rng(3)
    hx = rng.integers(0, 1 << n, size=6)
    hz = rng.integers(0, 1 << n, size=6)
    hc = rng.standard_normal(6)
    h = (hx, hz, hc, 0.3)
    ident = ([], [], [], 1.0)
    zop = ([0], [4], [1.0], 0.0)
    psi = _state(n, 4)
    mh, ms = evolve.qse_matrices(n, h, [ident, zop], psi)
    hm = spin.csr_from_packed(n, *h).toarray()
    assert abs(mh[0, 0] - np.vdot(psi, hm @ psi).real) < 1e-13
    assert abs(ms[0, 0] - 1) < 1e-13 and abs(ms[1, 1] - 1) < 1e-13
    assert abs(ms[0, 1] - np.vdot(psi, spin.apply_packed(n, [0], [4], [1.0], 0.0, psi)).real) < 1e-13
