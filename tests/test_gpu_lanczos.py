"""Kernel (3) parity: Lanczos ground state on the matrix-free apply vs the reference's
eigvalsh / eigsh(get_matrix(sparse=True), which="SA") path
(tests/integration/test_integration_models.py:102-103, test_integration_adaptvqe.py:97).
Tolerance 1e-10 relative (fp64), 1e-5 (fp32).  Runs on the B200 (-m gpu)."""

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import models, spin, transforms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import myqlm_fermion_b200 as qfe

    return qfe.get_engine(0)


def _hubbard_chain(ns, U=4.0, mu=2.0):
    t = np.zeros((ns, ns))
    for i in range(ns - 1):
        t[i, i + 1] = t[i + 1, i] = -1.0
    return models.hubbard_integrals(t, U, mu)


def test_config1_hubbard_2x2_ground_energy(eng):
    import myqlm_fermion_b200 as qfe

    t = np.zeros((4, 4))
    for a, b in [(0, 1), (1, 3), (3, 2), (2, 0)]:
        t[a, b] = t[b, a] = -1.0
    for enc in ("jordan-wigner", "bravyi-kitaev", "parity"):
        hs = qfe.make_hubbard_model(t, 4.0, 2.0).to_spin(enc)
        assert len(hs.terms) == 20
        e0, psi0, info = eng.ground_state(hs.packed(eng), tol=1e-12)
        assert abs(e0 - (-10.102748483462088)) <= 1e-10 * 10.1, (enc, e0, info)
        # Ritz vector: residual and Rayleigh quotient
        e_rq = eng.energy(hs.packed(eng), psi0)
        assert abs(e_rq - e0) <= 1e-9
        # the facade's get_matrix agrees with the oracle's matrix
        x, z, c, k = hs.packed_arrays()
        ref = spin.csr_from_packed(8, x, z, c, k).toarray()
        assert np.max(np.abs(hs.get_matrix() - ref)) <= 1e-12
        assert abs(np.linalg.eigvalsh(ref).min() - e0) <= 1e-10 * 10.1


@pytest.mark.parametrize("ns", [5, 7])
def test_hubbard_chain_vs_eigsh(eng, ns):
    """Config 3 family (12 sites = 24 qubits is the bench size; the oracle's eigsh needs the CSR
    matrix, so parity runs at 10 and 14 qubits)."""
    n = 2 * ns
    hpq, hpqrs = _hubbard_chain(ns)
    terms = eng.terms_from_integrals(hpq, hpqrs)
    ref = transforms.transform(n, models.electronic_terms(hpq, hpqrs), 0.0, "jordan-wigner")
    rx, rz, rc, rk = ref.canonical()
    mat = spin.csr_from_packed(n, rx, rz, rc, rk)
    want = spla.eigsh(mat, k=1, which="SA", tol=1e-13)[0][0]
    e0, psi0, info = eng.ground_state(terms, max_iter=400, tol=1e-11)
    assert abs(e0 - want) <= 1e-10 * abs(want), (e0, want, info)
    # eigen-residual of the returned vector
    phi = eng.apply(terms, psi0)
    eng.axpy(-e0, psi0, phi)
    assert abs(eng.dot(phi, phi)) ** 0.5 <= 1e-5
    import torch

    e32, _, info32 = eng.ground_state(terms, max_iter=400, tol=1e-6, dtype=torch.complex64)
    assert abs(e32 - want) <= 1e-5 * abs(want), (e32, want, info32)


def test_molecular_ground_state(eng):
    n = 10
    one, two = models.synthetic_integrals(5, seed=1234)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    for enc in ("jordan-wigner", "bravyi-kitaev"):
        terms = eng.terms_from_integrals(hpq, hpqrs, constant=1.25, encoding=enc)
        x, z, c, _, k = terms.arrays()
        want = np.linalg.eigvalsh(spin.csr_from_packed(n, x, z, c, complex(k[0])).toarray()).min()
        e0, _, info = eng.ground_state(terms, max_iter=600, tol=1e-12, want_vector=False)
        assert abs(e0 - want) <= 1e-10 * abs(want), (enc, e0, want, info)


def test_h2_fixture_ground_energy(eng, goldens, golden_dir):
    """UCC H2 golden of the reference (tests/integration/test_integration_ucc.py:112): -1.13727017466 @ 8 decimals."""
    d = np.load(golden_dir / "h2_integrals.npz")
    hpq, hpqrs = models.convert_to_h_integrals(d["one_body_integrals"], d["two_body_integrals"])
    terms = eng.terms_from_integrals(hpq, hpqrs, constant=float(d["nuclear_repulsion"]))
    e0, _, _ = eng.ground_state(terms, tol=1e-13)
    np.testing.assert_almost_equal(e0, goldens["scalars"]["h2_ucc_vqe_energy"]["value"], decimal=8)


def test_zne_golden_energy_on_gpu(eng, goldens):
    """The reference's 12-decimal golden <psi|H|psi> (tests/integration/test_integration_zne_plugin.py:16-49)."""
    from oracle import circuits

    hpq = np.array([[0.5, 1, 0, 0], [1, 1, 0, 0], [0, 0, 0.5, 1], [0, 0, 1, 1]], dtype=float)
    hpqrs = np.zeros((4, 4, 4, 4))
    hpqrs[0, 2, 0, 2] = hpqrs[2, 0, 2, 0] = -1
    terms = eng.terms_from_integrals(hpq, hpqrs)
    np.random.seed(0)
    psi = circuits.shallow_circuit_state(np.random.random(8))
    import torch

    e = eng.expectation(terms, torch.from_numpy(psi.astype(np.complex128)).to("cuda:0"))
    g = goldens["scalars"]["zne_energy_shallow_seed0"]
    np.testing.assert_almost_equal(e.real, g["value"], decimal=g["decimals"])
    assert abs(e.imag) < 1e-14
