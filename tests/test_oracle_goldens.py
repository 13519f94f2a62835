"""Pins the CPU oracle against every golden the reference's tests / executed notebooks hold for
the Hamiltonian hot path.  CPU only."""

import numpy as np
import pytest

from oracle import circuits, fock, models, plugins, spin, transforms
from oracle.spin import PauliSum


def _triples(lst):
    return [(complex(c[0], c[1]) if isinstance(c, list) else c, op, qb) for c, op, qb in lst]


def _assert_same_sum(a: PauliSum, b: PauliSum, tol=1e-12):
    keys = set(a.pruned(tol).d) | set(b.pruned(tol).d)
    for k in keys:
        assert abs(a.d.get(k, 0) - b.d.get(k, 0)) <= tol, (k, a.d.get(k), b.d.get(k))


ZNE_HPQ = np.array([[0.5, 1, 0, 0], [1, 1, 0, 0], [0, 0, 0.5, 1], [0, 0, 1, 1]], dtype=float)


def _zne_hamiltonian():
    hpqrs = np.zeros((4, 4, 4, 4))
    hpqrs[0, 2, 0, 2] = -1
    hpqrs[2, 0, 2, 0] = -1
    return transforms.transform(4, models.electronic_terms(ZNE_HPQ, hpqrs), 0.0, "jordan-wigner")


def test_spin_matrix_golden(goldens):
    g = goldens["spin_matrix_4x4"]
    expect = np.array(g["matrix_re"]) + 1j * np.array(g["matrix_im"])
    terms = [(c, op, qb) for c, op, qb in g["terms"]]
    np.testing.assert_equal(spin.spin_get_matrix(2, terms), expect)
    np.testing.assert_equal(spin.spin_get_matrix(2, terms, sparse=True).toarray(), expect)
    ps = PauliSum.from_terms(2, terms)
    xs, zs, cs, const = ps.canonical()
    np.testing.assert_allclose(spin.csr_from_packed(2, xs, zs, cs, const).toarray(), expect, atol=0)
    assert abs(np.linalg.eigvalsh(expect).min() - goldens["scalars"]["spin_min_eigenvalue"]["value"]) < 1e-13


def test_fermion_matrix_golden(goldens):
    g = goldens["fermion_matrix_4x4"]
    terms = [(c, op, qb) for c, op, qb in g["terms"]]
    np.testing.assert_equal(fock.fermion_get_matrix(2, terms), np.array(g["matrix_re"], dtype=complex))
    np.testing.assert_equal(fock.fermion_get_matrix(2, terms, sparse=True).toarray(), np.array(g["matrix_re"], dtype=complex))


@pytest.mark.parametrize("key", ["parity_n3", "parity_n4_electronic"])
def test_parity_notebook_term_lists(goldens, key):
    g = goldens[key]
    n = g["nqbits"]
    if "fermion_terms" in g:
        fterms = [(c, op, qb) for c, op, qb in g["fermion_terms"]]
    else:
        hpqrs = np.zeros((n,) * 4)
        for idx, v in g["hpqrs_nonzero"]:
            hpqrs[tuple(idx)] = v
        fterms = models.electronic_terms(np.array(g["hpq"]), hpqrs)
    got = transforms.transform(n, fterms, 0.0, g["encoding"])
    want = PauliSum.from_terms(n, _triples(g["terms"]), complex(*g["constant"]))
    _assert_same_sum(got, want, 1e-15)
    assert len(got.pruned(1e-12).d) == len(g["terms"]) + 1


def test_one_site_hubbard_jw(goldens):
    g = goldens["hubbard_one_site_jw"]
    hpq, hpqrs = models.hubbard_integrals(np.zeros((1, 1)), g["U"], g["mu"])
    got = transforms.transform(2, models.electronic_terms(hpq, hpqrs), 0.0, "jordan-wigner")
    want = PauliSum.from_terms(2, _triples(g["terms"]), complex(*g["constant"]))
    _assert_same_sum(got, want, 1e-15)
    xs, zs, cs, const = got.canonical()
    assert abs(np.linalg.eigvalsh(spin.csr_from_packed(2, xs, zs, cs, const).toarray()).min() + 1.0) < 1e-14


def test_zne_golden_energy(goldens):
    np.random.seed(0)
    theta = np.random.random(8)
    psi = circuits.shallow_circuit_state(theta)
    h = _zne_hamiltonian()
    xs, zs, cs, const = h.canonical()
    e = spin.expectation_packed(4, xs, zs, cs, const, psi)
    g = goldens["scalars"]["zne_energy_shallow_seed0"]
    assert abs(e.imag) < 1e-14
    np.testing.assert_almost_equal(e.real, g["value"], decimal=g["decimals"])
    # and through the reference's own matrix builder
    m = spin.spin_get_matrix(4, h.term_triples(), h.constant)
    np.testing.assert_almost_equal(np.vdot(psi, m @ psi).real, g["value"], decimal=g["decimals"])


def test_seqoptim_golden_energy(goldens):
    np.random.seed(1)
    x0 = np.random.random(8)
    h = _zne_hamiltonian()
    m = spin.spin_get_matrix(4, h.term_triples(), h.constant)

    def energy(params):
        psi = circuits.shallow_circuit_state(params)
        return np.vdot(psi, m @ psi)

    e, _ = plugins.rotosolve(energy, x0, ncycles=10)
    g = goldens["scalars"]["seqoptim_energy_seed1"]
    np.testing.assert_almost_equal(e, g["value"], decimal=g["decimals"])


def test_hubbard_dimer_goldens(goldens):
    t_mat = 0.25 * np.array([[0.0, 1.0], [1.0, 0.0]])
    hpq, hpqrs = models.hubbard_integrals(t_mat, 2.0, 1.0)
    mat = fock.fermion_get_matrix(4, models.electronic_terms(hpq, hpqrs))
    assert abs(np.linalg.eigvalsh(mat).min() - goldens["scalars"]["hubbard_dimer_ground"]["value"]) < 1e-14
    mat0 = fock.fermion_get_matrix(4, models.electronic_terms(hpq, None))
    assert abs(np.linalg.eigvalsh(mat0).min() - goldens["scalars"]["hubbard_dimer_ground_no_interaction"]["value"]) < 1e-14
    for enc in transforms.ENCODINGS:
        ps = transforms.transform(4, models.electronic_terms(hpq, hpqrs), 0.0, enc)
        m = spin.spin_get_matrix(4, ps.term_triples(), ps.constant)
        assert abs(np.linalg.eigvalsh(m).min() - goldens["scalars"]["hubbard_dimer_ground"]["value"]) < 1e-13


def test_h2_fixture_energy(goldens, golden_dir):
    d = np.load(golden_dir / "h2_integrals.npz")
    hpq, hpqrs = models.convert_to_h_integrals(d["one_body_integrals"], d["two_body_integrals"])
    terms = models.electronic_terms(hpq, hpqrs)
    for enc in transforms.ENCODINGS:
        ps = transforms.transform(4, terms, float(d["nuclear_repulsion"]), enc)
        assert len(ps.pruned(1e-12).d) == 15
        m = spin.spin_get_matrix(4, ps.term_triples(), ps.constant)
        e0 = np.linalg.eigvalsh(m).min()
        # the fixture's FCI field and the reference's UCC-VQE golden agree with it to 8 decimals
        np.testing.assert_almost_equal(e0, goldens["scalars"]["h2_fci"]["value"], decimal=8)
        np.testing.assert_almost_equal(e0, goldens["scalars"]["h2_ucc_vqe_energy"]["value"], decimal=8)


# --- the reference's own transform criterion: tests/unitary/test_unitary_transforms.py:23-52 -----
N5 = 5
H_LIST = [
    [(1.0, "C", [0])],
    [(1.0, "C", [1])],
    [(1.0, "CCcc", [0, 1, 1, 0]), (1.0, "CCcc", [1, 0, 0, 1])],
    [(1.0, "CCcc", [0, 1, 1, 0]), (1.0, "CCcc", [1, 0, 0, 1]), (0.1, "Cc", [1, 0])],
]


@pytest.mark.parametrize("enc", transforms.ENCODINGS)
@pytest.mark.parametrize("fterms", H_LIST)
def test_transform_vs_change_encoding(enc, fterms):
    ps = transforms.transform(N5, fterms, 0.0, enc)
    a = spin.spin_get_matrix(N5, ps.term_triples(), ps.constant)
    b = transforms.change_encoding(fock.fermion_get_matrix(N5, fterms), transforms.code_for(N5, enc))
    np.testing.assert_almost_equal(np.linalg.norm(a - b), 0, decimal=10)


def test_fenwick_sets_known_answers():
    # SURVEY.md appendix A probe vectors (P, C, U per qubit)
    want5 = [(set(), set(), {1, 2, 4}), ({0}, set(), {2, 4}), ({1}, set(), {4}), ({2}, {2}, {4}), ({2, 3}, set(), set())]
    assert transforms.pcu_sets(5) == want5
    want8 = [
        (set(), set(), {1, 3, 7}), ({0}, set(), {3, 7}), ({1}, {1}, {3, 7}), ({1, 2}, set(), {7}),
        ({3}, {3}, {5, 7}), ({3, 4}, {3}, {7}), ({3, 5}, {3, 5}, {7}), ({3, 5, 6}, set(), set()),
    ]
    assert transforms.pcu_sets(8) == want8


def test_config1_hubbard_2x2():
    t = np.zeros((4, 4))
    for a, b in [(0, 1), (1, 3), (3, 2), (2, 0)]:
        t[a, b] = t[b, a] = -1.0
    hpq, hpqrs = models.hubbard_integrals(t, 4.0, 2.0)
    fterms = models.electronic_terms(hpq, hpqrs)
    assert len(fterms) == 24 + 8
    e_ref = np.linalg.eigvalsh(fock.fermion_get_matrix(8, fterms)).min()
    assert abs(e_ref - (-10.102748483462088)) < 1e-12
    for enc in transforms.ENCODINGS:
        ps = transforms.transform(8, fterms, 0.0, enc).pruned(1e-12)
        assert len(ps.d) == 21
        xs, zs, cs, const = ps.canonical()
        m = spin.csr_from_packed(8, xs, zs, cs, const).toarray()
        assert abs(np.linalg.eigvalsh(m).min() - e_ref) < 1e-12


def test_matrix_free_equals_kron_chain():
    one, two = models.synthetic_integrals(3, seed=7)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    n = 6
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    for enc in transforms.ENCODINGS:
        ps = transforms.transform(n, models.electronic_terms(hpq, hpqrs), 0.3, enc)
        xs, zs, cs, const = ps.canonical()
        m = spin.spin_get_matrix(n, ps.term_triples(), ps.constant, sparse=True)
        np.testing.assert_allclose(spin.apply_packed(n, xs, zs, cs, const, psi), m @ psi, atol=1e-12)
        np.testing.assert_allclose(spin.csr_from_packed(n, xs, zs, cs, const).toarray(), m.toarray(), atol=1e-13)
        f = fock.fermion_get_matrix(n, models.electronic_terms(hpq, hpqrs), 0.3)
        np.testing.assert_allclose(m.toarray(), transforms.change_encoding(f, transforms.code_for(n, enc)), atol=1e-12)


def test_uccsd_pool_sizes_and_single_x_mask():
    # SURVEY.md appendix B
    assert len(models.excitation_tuples(2, 4)) == 3
    assert len(models.excitation_tuples(4, 12)) == 92
    assert len(models.excitation_tuples(10, 14)) == 140
    n = 8
    for op in models.cluster_ops(4, n):
        ps = transforms.transform(n, op, 0.0, "jordan-wigner").pruned(1e-12)
        assert len({k[0] for k in ps.d}) == 1
        assert len(ps.d) in (2, 8)
        # Hermitian generator
        _assert_same_sum(ps, ps.dag())


def test_commutator_gradient_identity():
    """-i<[H,tau]> == 2 Im <H psi|tau|psi> for Hermitian H, tau."""
    n = 6
    one, two = models.synthetic_integrals(3, seed=3)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    h = transforms.transform(n, models.electronic_terms(hpq, hpqrs), 0.0, "jordan-wigner")
    pool = [transforms.transform(n, op, 0.0, "jordan-wigner") for op in models.cluster_ops(2, n)]
    rng = np.random.default_rng(5)
    psi = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    psi /= np.linalg.norm(psi)
    g = plugins.commutator_gradients(h, pool, psi)
    hx, hz, hc, hconst = h.canonical()
    phi = spin.apply_packed(n, hx, hz, hc, hconst, psi)
    for k, tau in enumerate(pool):
        tx, tz, tc, tconst = tau.canonical()
        val = 2 * np.imag(np.vdot(phi, spin.apply_packed(n, tx, tz, tc, tconst, psi)))
        assert abs(g[k].imag) < 1e-12
        assert abs(g[k].real - val) < 1e-12
