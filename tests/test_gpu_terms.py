"""Kernel (1) parity: GPU fermion -> Pauli term expansion vs the oracle's dictionary algebra.
Masks must match bit-exactly, coefficients to 1e-12.  Runs on the B200 (-m gpu)."""

import numpy as np
import pytest

from oracle import models, spin, transforms
from oracle.spin import PauliSum

pytestmark = pytest.mark.gpu

ENCODINGS = ("jordan-wigner", "bravyi-kitaev", "parity")


@pytest.fixture(scope="module")
def eng():
    import myqlm_fermion_b200 as qfe

    return qfe.get_engine(0)


def _assert_terms_equal(packed, ref: PauliSum, tol=1e-12):
    gx, gz, gc, _, gk = packed.arrays()
    rx, rz, rc, rk = ref.canonical(eps=tol)
    # a term may be missing on one side only if its coefficient is below the drop threshold
    gd = {(int(a), int(b)): c for a, b, c in zip(gx, gz, gc)}
    rd = {(int(a), int(b)): c for a, b, c in zip(rx, rz, rc)}
    full = {k: v * (-1j) ** (spin.popcount(k[0] & k[1]) % 4) for k, v in ref.d.items()}
    for k in set(gd) | set(rd):
        a = gd.get(k, 0.0)
        b = rd.get(k, full.get(k, 0.0))
        assert abs(a - b) <= 2 * tol, (k, a, b)
        if k not in gd or k not in rd:
            assert abs(full.get(k, 0.0)) <= 4 * tol, f"term {k} present on one side only"
    assert abs(complex(gk[0]) - rk) <= tol * max(1.0, abs(rk))
    # canonical export order: ascending (x, z)
    keys = list(zip(gx.tolist(), gz.tolist()))
    assert keys == sorted(keys)
    assert len(set(keys)) == len(keys)


def test_config1_hubbard_terms(eng):
    t = np.zeros((4, 4))
    for a, b in [(0, 1), (1, 3), (3, 2), (2, 0)]:
        t[a, b] = t[b, a] = -1.0
    hpq, hpqrs = models.hubbard_integrals(t, 4.0, 2.0)
    for enc in ENCODINGS:
        got = eng.terms_from_integrals(hpq, hpqrs, encoding=enc)
        assert got.n_terms == 20  # 21 with the identity
        ref = transforms.transform(8, models.electronic_terms(hpq, hpqrs), 0.0, enc)
        _assert_terms_equal(got, ref)
        gx, gz, gc, _, gk = got.arrays()
        rx, rz, rc, rk = ref.canonical()
        assert np.array_equal(gx, rx) and np.array_equal(gz, rz)  # bit-exact masks
        assert np.array_equal(gc, rc)  # these coefficients are dyadic rationals: exact


@pytest.mark.parametrize("m", [2, 3, 5, 7])
@pytest.mark.parametrize("enc", ENCODINGS)
def test_synthetic_molecular_terms(eng, m, enc):
    """Config 2 family: random 8-fold symmetric integrals, n = 2m qubits (m=7: 3 382 terms)."""
    one, two = models.synthetic_integrals(m, seed=1234)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    n = 2 * m
    got = eng.terms_from_integrals(hpq, hpqrs, constant=0.25, encoding=enc)
    ref = transforms.transform(n, models.electronic_terms(hpq, hpqrs), 0.25, enc)
    _assert_terms_equal(got, ref)
    if m == 7:
        assert got.n_terms == 3381
        assert got.n_xmasks == 554  # 553 + the diagonal group


@pytest.mark.parametrize("key", ["parity_n3", "parity_n4_electronic"])
def test_notebook_parity_term_lists(eng, goldens, key):
    """Term lists printed in the reference's executed notebook
    (doc/notebooks/qat_fermion_spin_fermion_transforms.ipynb cells 5/7/12/14)."""
    g = goldens[key]
    n = g["nqbits"]
    tr = lambda lst: [(complex(c[0], c[1]) if isinstance(c, list) else c, op, qb) for c, op, qb in lst]
    if "fermion_terms" in g:
        got = eng.terms_from_fermion(n, tr(g["fermion_terms"]), encoding="parity")
    else:
        hpqrs = np.zeros((n,) * 4)
        for idx, v in g["hpqrs_nonzero"]:
            hpqrs[tuple(idx)] = v
        got = eng.terms_from_integrals(np.array(g["hpq"], dtype=float), hpqrs, encoding="parity")
    want = PauliSum.from_terms(n, tr(g["terms"]), complex(*g["constant"]))
    _assert_terms_equal(got, want, tol=1e-15)
    assert got.n_terms == len(g["terms"])


def test_h2_and_lih_fixtures(eng, golden_dir):
    for name, n_terms in (("h2_integrals.npz", 14),):
        d = np.load(golden_dir / name)
        hpq, hpqrs = models.convert_to_h_integrals(d["one_body_integrals"], d["two_body_integrals"])
        for enc in ENCODINGS:
            got = eng.terms_from_integrals(hpq, hpqrs, constant=float(d["nuclear_repulsion"]), encoding=enc)
            assert got.n_terms == n_terms  # 15 with the identity
            ref = transforms.transform(4, models.electronic_terms(hpq, hpqrs), float(d["nuclear_repulsion"]), enc)
            _assert_terms_equal(got, ref)


def test_general_fermion_terms_and_pool(eng):
    """transform of arbitrary ladder-operator strings (the reference's H_LIST of
    tests/unitary/test_unitary_transforms.py:23-33) and of the UCCSD pool operators."""
    n = 5
    h_list = [
        [(1.0, "C", [0])],
        [(1.0, "C", [1])],
        [(1.0, "CCcc", [0, 1, 1, 0]), (1.0, "CCcc", [1, 0, 0, 1])],
        [(1.0, "CCcc", [0, 1, 1, 0]), (1.0, "CCcc", [1, 0, 0, 1]), (0.1, "Cc", [1, 0])],
        [(0.5 - 0.25j, "CcCcC", [4, 2, 2, 0, 3]), (1.5, "cC", [3, 3])],
    ]
    for fterms in h_list:
        for enc in ENCODINGS:
            got = eng.terms_from_fermion(n, fterms, constant=-0.5, encoding=enc)
            _assert_terms_equal(got, transforms.transform(n, fterms, -0.5, enc))
    for op in models.cluster_ops(4, 12)[::7]:
        got = eng.terms_from_fermion(12, op, encoding="jordan-wigner")
        _assert_terms_equal(got, transforms.transform(12, op, 0.0, "jordan-wigner"))
        assert got.n_xmasks == 1 and got.n_terms in (2, 8)


def test_edge_cases(eng):
    # empty Hamiltonian, all integrals below TOL, constant only
    z2 = np.zeros((4, 4))
    got = eng.terms_from_integrals(z2, None, constant=1.5)
    assert got.n_terms == 0 and got.constant == 1.5
    tiny = np.full((4, 4), 1e-13)
    assert eng.terms_from_integrals(tiny, np.full((4,) * 4, 1e-13)).n_terms == 0
    assert eng.terms_from_fermion(3, [], constant=2.0).n_terms == 0
    # n = 1 and the 64-qubit limit of the mask words
    got = eng.terms_from_fermion(1, [(1.0, "Cc", [0, 0])])
    _assert_terms_equal(got, transforms.transform(1, [(1.0, "Cc", [0, 0])], 0.0, "jordan-wigner"))
    f64 = [(1.0, "Cc", [0, 63]), (1.0, "Cc", [63, 0]), (0.5, "CCcc", [1, 62, 62, 1])]
    for enc in ENCODINGS:
        _assert_terms_equal(eng.terms_from_fermion(64, f64, encoding=enc), transforms.transform(64, f64, 0.0, enc))
    # duplicate / cancelling Pauli strings are merged and dropped
    x = np.array([1, 1, 2], dtype=np.uint64)
    z = np.array([0, 0, 2], dtype=np.uint64)
    c = np.array([1.0, -1.0, 0.5j])
    got = eng.terms_from_pauli(2, x, z, c, constant=0.0)
    gx, gz, gc, _, _ = got.arrays()
    assert gx.tolist() == [2] and gz.tolist() == [2] and gc[0] == 0.5j
    # errors map to the reference's exception types
    import myqlm_fermion_b200 as qfe

    with pytest.raises(KeyError):
        eng.terms_from_integrals(z2, None, encoding="nope")
    with pytest.raises(qfe.QfeError):
        eng.terms_from_fermion(3, [(1.0, "CX", [0, 1])])
    with pytest.raises(qfe.QfeError):
        eng.terms_from_fermion(3, [(1.0, "Cc", [0, 3])])


def test_facade_transforms_match_reference_criterion(eng):
    """tests/unitary/test_unitary_transforms.py:44-52 through the facade objects:
    ||T(H).get_matrix() - change_encoding(H.get_matrix(), code)|| < 1e-10."""
    import myqlm_fermion_b200 as qfe
    from oracle import fock

    n = 5
    h_list = [
        [(1.0, "C", [0])],
        [(1.0, "CCcc", [0, 1, 1, 0]), (1.0, "CCcc", [1, 0, 0, 1]), (0.1, "Cc", [1, 0])],
    ]
    for fterms in h_list:
        h = qfe.FermionHamiltonian(n, [qfe.Term(c, op, qb) for c, op, qb in fterms], normal_order=False)
        a_f = fock.fermion_get_matrix(n, fterms)
        for fn, enc in ((qfe.transform_to_jw_basis, "jordan-wigner"), (qfe.transform_to_parity_basis, "parity"), (qfe.transform_to_bk_basis, "bravyi-kitaev")):
            a = fn(h).get_matrix()
            b = transforms.change_encoding(a_f, transforms.code_for(n, enc))
            np.testing.assert_almost_equal(np.linalg.norm(a - b), 0, decimal=10)
            # the facade's own code matrices and change_encoding (transforms.py:260-407) give the same relabelling
            code = {"jordan-wigner": qfe.get_jw_code, "parity": qfe.get_parity_code, "bravyi-kitaev": qfe.get_bk_code}[enc](n)
            assert np.array_equal(qfe.change_encoding(a_f, code), b)
