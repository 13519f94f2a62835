"""The Hamiltonian objects of the facade against the behaviour the reference's own unit tests pin
(/root/reference/tests/unitary/test_unitary_hamiltonians.py, test_unitary_transforms.py, test_unitary_trotterization.py):
same constructors, arithmetic, conversions, exception types, and -- on the GPU -- the same matrices, which the facade obtains
from the matrix-free engine instead of kron chains."""

import numpy as np
import pytest

import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200 import ElectronicStructureHamiltonian, FermionHamiltonian, SpinHamiltonian, Term


# ---- host-side object algebra (no GPU) -----------------------------------------------------------------------------------------
def _h1():
    return FermionHamiltonian(2, terms=[Term(1.0 + 2j, "CCcc", [1, 0, 1, 0])])


def test_adding_and_subtracting_constants():
    """test_unitary_hamiltonians.py:104-146: a number lands in constant_coeff, on either side."""
    assert _h1().constant_coeff == 0
    assert (_h1() + 2).constant_coeff == 2.0
    assert (2 + _h1()).constant_coeff == 2.0
    assert (_h1() - 2).constant_coeff == -2.0
    h = 2 - _h1()
    assert h.constant_coeff == 2.0
    t = h.terms[0]
    assert (t.coeff, t.op, t.qbits) == ((-1 - 2j), "CCcc", [0, 1, 0, 1])


def test_adding_subtracting_and_scaling_fermionic_hamiltonians():
    """:120-127, :149-187: terms are normal ordered (CCcc on ascending creation / annihilation indices, sign from the swaps) and merged."""
    h2 = FermionHamiltonian(2, terms=[Term(1.0, "CCcc", [0, 1, 1, 0])])
    t = (_h1() + h2).terms[0]
    assert (t.coeff, t.op, t.qbits) == (2j, "CCcc", [0, 1, 0, 1])
    t = (_h1() - h2).terms[0]
    assert (t.coeff, t.op, t.qbits) == ((2 + 2j), "CCcc", [0, 1, 0, 1])
    for h in (2 * _h1(), _h1() * 2):
        t = h.terms[0]
        assert (t.coeff, t.op, t.qbits) == ((2 + 4j), "CCcc", [0, 1, 0, 1])
    h = _h1() * h2  # (n_0 n_1)^2 = n_0 n_1
    t = h.terms[0]
    assert (t.coeff, t.op, t.qbits, h.constant_coeff) == ((1 + 2j), "CCcc", [0, 1, 0, 1], 0.0)


def test_to_electronic_collects_normal_ordered_terms():
    """:61-73: c+_0 c_0 + (c+_0 c_1 c+_1 c_0 = n_0 - n_0 n_1) -> hpq[0,0] = 2, hpqrs term on [0, 1, 0, 1] with coefficient 1."""
    h = FermionHamiltonian(2, terms=[Term(1, "Cc", [0, 0]), Term(1, "CcCc", [0, 1, 1, 0])])
    h_elec = h.to_electronic()
    assert isinstance(h_elec, ElectronicStructureHamiltonian) and len(h.terms) == 2
    seen = set()
    for t in h_elec.terms:
        if t.op == "Cc":
            assert t.coeff == 2 and t.qbits == [0, 0]
        if t.op == "CCcc":
            assert t.coeff == 1 and t.qbits == [0, 1, 0, 1]
        seen.add(t.op)
    assert seen == {"Cc", "CCcc"}


def test_operator_letters_are_checked():
    """:76-85: a fermionic letter in a spin Hamiltonian and a Pauli letter in a fermionic one raise TypeError."""
    with pytest.raises(TypeError):
        SpinHamiltonian(2, terms=[Term(1.0, "Cc", [0, 1])])
    with pytest.raises(TypeError):
        FermionHamiltonian(2, terms=[Term(1.0, "XY", [0, 1])])


def test_unknown_transformation_is_a_key_error():
    """hamiltonians.py:454: to_spin looks the method up in a dict."""
    with pytest.raises(KeyError):
        FermionHamiltonian(2, [Term(0.3, "Cc", [0, 1])]).to_spin("not-a-transform")


# ---- matrices (GPU: the facade computes them with the matrix-free engine) -------------------------------------------------------------
def _spin_example():
    return SpinHamiltonian(2, [Term(0.5, "X", [0]), Term(0.25, "Y", [1]), Term(1.0, "ZZ", [0, 1])], 0.0)


_SPIN_MATRIX = np.array(
    [[1.0, -0.25j, 0.5, 0.0], [0.25j, -1.0, 0.0, 0.5], [0.5, 0.0, -1.0, -0.25j], [0.0, 0.5, 0.25j, 1.0]], dtype=np.complex128
)
_FERMI_MATRIX = np.zeros((4, 4), dtype=np.complex128)
_FERMI_MATRIX[2, 1] = 1.0
_FERMI_MATRIX[3, 3] = -0.5


@pytest.mark.gpu
def test_spin_matrix_golden_and_lowest_eigenvalue():
    """:14-34 and :252-276: the 4x4 golden matrix, dense and sparse, and min eigenvalue -1.25 to 13 decimals."""
    h = _spin_example()
    np.testing.assert_equal(h.get_matrix(), _SPIN_MATRIX)
    np.testing.assert_equal(h.get_matrix(sparse=True).toarray(), _SPIN_MATRIX)
    np.testing.assert_almost_equal(min(np.linalg.eigh(h.get_matrix())[0]), -1.25, decimal=13)


@pytest.mark.gpu
def test_fermionic_and_electronic_matrix_goldens():
    """:279-311: Fock-space matrix of c+_0 c_1 + 0.5 c+_0 c+_1 c_0 c_1, unchanged by to_electronic; dag() conjugates."""
    h = FermionHamiltonian(2, [Term(1.0, "Cc", [0, 1]), Term(0.5, "CCcc", [0, 1, 0, 1])])
    for obj in (h, h.to_electronic()):
        np.testing.assert_equal(obj.get_matrix(), _FERMI_MATRIX)
        np.testing.assert_equal(obj.get_matrix(sparse=True).toarray(), _FERMI_MATRIX)
    hj = FermionHamiltonian(2, [Term(1.0j, "Cc", [0, 1]), Term(0.5j, "CCcc", [0, 1, 0, 1])]).to_electronic()
    np.testing.assert_equal(np.conj(hj.get_matrix()), hj.dag().get_matrix())


@pytest.mark.gpu
def test_to_spin_and_to_electronic_keep_the_matrix_and_spectrum():
    """:37-58, :88-101."""
    h = FermionHamiltonian(2, [Term(0.3, "Cc", [0, 1]), Term(1.4, "CcCc", [0, 1, 1, 0])])
    spin_h = h.to_spin()
    assert all("C" not in t.op and "c" not in t.op for t in spin_h.terms)
    np.testing.assert_equal(h.get_matrix(), spin_h.get_matrix())
    compatible = FermionHamiltonian(2, [Term(0.3, "Cc", [0, 1]), Term(1.4, "CCcc", [0, 1, 1, 0])])
    np.testing.assert_equal(compatible.get_matrix(), compatible.to_electronic().get_matrix())
    h3 = FermionHamiltonian(3, [Term(1.0, "Cc", [0, 1]), Term(0.5, "CCcc", [0, 1, 0, 1])])
    spectra = [sorted(np.linalg.eigvals(o.get_matrix())) for o in (h3, h3.to_spin(), h3.to_electronic())]
    np.testing.assert_equal(spectra[0], spectra[1])
    np.testing.assert_equal(spectra[0], spectra[2])


@pytest.mark.gpu
def test_matrices_follow_changed_terms_integrals_and_constants():
    """:190-248: set_terms, the hpqrs setter and constant_coeff are picked up by the next get_matrix."""
    s = _spin_example()
    m1 = s.get_matrix()
    s.set_terms([Term(0.5, "Z", [0]), Term(0.5, "X", [1]), Term(1.0, "XX", [0, 1])])
    assert not np.all(np.equal(m1, s.get_matrix()))
    f = FermionHamiltonian(2, [Term(1.0, "Cc", [0, 1]), Term(0.5, "CCcc", [0, 1, 0, 1])])
    m1 = f.get_matrix()
    f.set_terms([Term(0.2, "Cc", [0, 1]), Term(2, "CCcc", [0, 1, 1, 0])])
    assert not np.all(np.equal(m1, f.get_matrix()))
    e = FermionHamiltonian(2, [Term(1.0, "Cc", [0, 1]), Term(0.5, "CCcc", [0, 1, 0, 1])]).to_electronic()
    m1 = e.get_matrix()
    e.hpqrs = np.zeros(e.hpqrs.shape)
    m2 = e.get_matrix()
    assert not np.all(np.equal(m1, m2))
    e.constant_coeff += 3
    assert np.all(np.equal(e.get_matrix(), m2 + 3 * np.eye(4, 4)))


@pytest.mark.gpu
@pytest.mark.parametrize("transform, code", [(qfe.transform_to_jw_basis, qfe.get_jw_code), (qfe.transform_to_parity_basis, qfe.get_parity_code),
                                             (qfe.transform_to_bk_basis, qfe.get_bk_code)])
def test_transforms_match_the_recoded_fock_matrix(transform, code):
    """test_unitary_transforms.py:23-52: ||T(H).get_matrix() - change_encoding(H.get_matrix(), code)|| = 0 to 10 decimals."""
    n = 5
    for terms in ([Term(1.0, "C", [0])], [Term(1.0, "CCcc", [0, 1, 1, 0]), Term(1.0, "CCcc", [1, 0, 0, 1]), Term(0.1, "Cc", [1, 0])]):
        h = FermionHamiltonian(n, terms, normal_order=False)
        np.testing.assert_almost_equal(np.linalg.norm(transform(h).get_matrix() - qfe.change_encoding(h.get_matrix(), code(n))), 0, decimal=10)


# ---- normal ordering (test_unitary_fermion_algebra.py) -------------------------------------------------------------------------
def test_make_pcu_sets_keeps_the_reference_signature():
    """make_PCU_sets(nqbits) -> [(P, C, U)] for every qubit (/root/reference/qat/fermion/transforms.py:65-79); the known answers are
    SURVEY.md appendix A's probe vectors, the same ones the oracle is pinned to."""
    from myqlm_fermion_b200 import make_PCU_sets

    assert make_PCU_sets(5) == [(set(), set(), {1, 2, 4}), ({0}, set(), {2, 4}), ({1}, set(), {4}), ({2}, {2}, {4}), ({2, 3}, set(), set())]
    got8 = make_PCU_sets(8)
    assert len(got8) == 8 and got8[6] == ({3, 5}, {3, 5}, {7}) and got8[7] == ({3, 5, 6}, set(), set())
    assert make_PCU_sets(1) == [(set(), set(), set())]


def test_normal_ordering_of_single_terms():
    from myqlm_fermion_b200 import FermionicTerm, normal_order_fermionic_term

    new_terms = normal_order_fermionic_term(FermionicTerm(1, "CcCc", [0, 1, 1, 0]))
    assert len(new_terms) == 2
    for t in new_terms:
        if t.op == "Cc":
            assert t.qbits == [0, 0] and t.coeff == 1
        if t.op == "CCcc":
            assert t.qbits == [0, 1, 0, 1] and t.coeff == 1
    for letter in "cC":
        (t,) = normal_order_fermionic_term(FermionicTerm(1.0, letter, [0]))
        assert (t.op, t.coeff, t.qbits) == (letter, 1.0, [0])
    for op in ("CC", "cc"):  # one swap: a sign, indices ascending
        (t,) = normal_order_fermionic_term(FermionicTerm(2.0, op, [1, 0]))
        assert (t.coeff, t.op, t.qbits) == (-2.0, op, [0, 1])


# ---- Trotter slices (test_unitary_trotterization.py) ------------------------------------------------------------------------------
def _slice_unitary(hs, coeff=1.0):
    """Columns = the slice applied to every basis state, on the device."""
    import torch

    eng = qfe.get_engine(0)
    n = hs.nbqbits
    cols = []
    for b in range(1 << n):
        psi = eng.basis_state(n, b)
        eng.trotter_slice(hs, psi, coeff=coeff)
        cols.append(psi)
    return torch.stack(cols, dim=1).cpu().numpy()


def _check_slice_is_the_exponential(hs, decimal=10):
    import scipy.linalg

    hm = hs.get_matrix() - hs.constant_coeff * np.eye(1 << hs.nbqbits)  # a spin slice has no gate for the identity term
    np.testing.assert_almost_equal(np.linalg.norm(scipy.linalg.expm(-1j * hm) - _slice_unitary(hs)), 0, decimal=decimal)


def _electronic(n, one=(), two=()):
    hpq = np.zeros((n, n))
    hpqrs = np.zeros((n, n, n, n))
    for idx, v in one:
        hpq[idx] = v
    for idx, v in two:
        hpqrs[idx] = v
    return ElectronicStructureHamiltonian(hpq, hpqrs)


@pytest.mark.gpu
def test_trotter_slices_of_single_fermionic_operators_are_exact():
    """The operator families of test_unitary_trotterization.py:79-198 (number, excitation, Coulomb / exchange, number-excitation,
    double excitation): the Jordan-Wigner strings of each commute, so one slice exp(-i c_t P_t) ... is exp(-iH) itself, up to the
    global phase of the identity term."""
    cases = [
        _electronic(2, one=[((1, 1), 2)]),
        _electronic(2, one=[((0, 1), 2), ((1, 0), 2)]),
        _electronic(5, one=[((4, 0), 2), ((0, 4), 2)]),
        _electronic(3, two=[((0, 2, 2, 0), -1.5)]),
        _electronic(4, two=[((3, 0, 0, 3), 6.5489)]),
        _electronic(3, two=[((2, 1, 1, 0), -1.5), ((0, 1, 1, 2), -1.5)]),
        _electronic(5, two=[((4, 0, 0, 1), -1.5), ((1, 0, 0, 4), -1.5)]),
        _electronic(5, two=[((2, 4, 4, 0), -1.5), ((0, 4, 4, 2), -1.5)]),
        _electronic(4, two=[((3, 2, 1, 0), -2), ((0, 1, 2, 3), -2)]),
        _electronic(6, two=[((5, 3, 2, 0), 2.25), ((0, 2, 3, 5), 2.25)]),
        _electronic(6, two=[((5, 3, 1, 0), 4.587), ((0, 1, 3, 5), 4.587)]),
    ]
    for h in cases:
        _check_slice_is_the_exponential(h.to_spin("jordan-wigner"))


@pytest.mark.gpu
def test_spin_trotter_slices():
    """:201-260: single ZZ / XX terms exactly; a random sum of ten two-qubit strings with small angles to 3 decimals like the reference."""
    _check_slice_is_the_exponential(SpinHamiltonian(2, terms=[Term(0.543, "ZZ", [0, 1])]))
    _check_slice_is_the_exponential(SpinHamiltonian(2, terms=[Term(0.3, "XX", [0, 1])]))
    rng = np.random.default_rng(0)
    terms = [Term(0.01 * rng.standard_normal(), "".join(rng.choice(list("XYZ"), 2)), [0, 1]) for _ in range(10)]
    _check_slice_is_the_exponential(SpinHamiltonian(2, terms=terms), decimal=3)


# ---- model generators (tests/integration/test_integration_models.py) ---------------------------------------------------------------
def test_embedded_model_index_maps_are_inverse():
    for n_orb in (1, 2, 5):
        assert [qfe.ind_spins_ord(qfe.ind_clusters_ord(i, n_orb), n_orb) for i in range(2 * n_orb)] == list(range(2 * n_orb))
        assert sorted(qfe.ind_clusters_ord(i, n_orb) for i in range(2 * n_orb)) == list(range(2 * n_orb))


@pytest.mark.gpu
def test_anderson_model_is_an_embedded_model():
    """:13-21: one bath level; the embedded model carries tr(lambda_c) in constant_coeff, set to zero as in the reference test."""
    U = 1
    mu = U / 2
    anderson = qfe.make_anderson_model(U, mu, [0.4], [0.04])
    embedded = qfe.make_embedded_model(U, mu, np.diag([0.4, 0.4]), -0.04 * np.eye(2), grouping="clusters")
    embedded.constant_coeff = 0
    diff = np.real(anderson.get_matrix(sparse=True)) - np.real(embedded.get_matrix(sparse=True))
    assert np.max(np.abs(diff)) < 1e-8


@pytest.mark.gpu
def test_risb_embedding_against_fock_space_operators():
    """:24-106: two impurity orbitals + bath, random D / lambda_c / t_loc, an explicit interaction kernel; the lowest eigenvalue equals
    the one of the Hamiltonian assembled from Fock-space creation operators (oracle/fock.py restates util.init_creation_ops)."""
    import itertools

    from scipy.sparse.linalg import eigsh

    from oracle import fock

    Nc = 2
    M = 2 * Nc
    U = 0.1
    mu = U / 2
    rng = np.random.RandomState(0)
    D = np.diag(rng.random_sample(4))
    lambda_c = np.diag(rng.random_sample(4))
    t_pm = np.diag(rng.random_sample(4))
    d_dag = fock.creation_operators(2 * M, sparse=True)
    d = {i: m.conj().T for i, m in d_dag.items()}
    ii, jj, kk, ll = np.indices((M, M, M, M))
    int_kernel = 1.0 * (((ii == jj) & (ii % 2 == 0) & (kk == ll) & (kk == ii + 1)) | ((ii == jj) & (ii % 2 == 1) & (kk == ll) & (kk == ii - 1)))
    kernel_cccc = 1.0 * (((ii == kk) & (ii % 2 == 0) & (jj == ll) & (jj == ii + 1)) | ((ii == kk) & (ii % 2 == 1) & (jj == ll) & (jj == ii - 1)))
    extended = np.zeros((2 * M, 2 * M, 2 * M, 2 * M))
    extended[:M, :M, :M, :M] = kernel_cccc
    int_term = U / 2 * sum(int_kernel[i, j, k, l] * d_dag[i].dot(d[j]).dot(d_dag[k]).dot(d[l])
                           for i, j, k, l in itertools.product(range(M), repeat=4) if int_kernel[i, j, k, l])
    non_int = (
        sum(lambda_c[i, j] * d[M + j].dot(d_dag[M + i]) for i, j in itertools.product(range(M), repeat=2) if lambda_c[i, j])
        + sum(D[i, j] * d_dag[i].dot(d[M + j]) + np.conj(D[i, j]) * d_dag[M + j].dot(d[i]) for i, j in itertools.product(range(M), repeat=2) if D[i, j])
        + sum(t_pm[i, j] * d_dag[i].dot(d[j]) for i, j in itertools.product(range(M), repeat=2) if t_pm[i, j])
        - mu * sum(d_dag[i].dot(d[i]) for i in range(M))
    )
    h_emb = qfe.make_embedded_model(U, mu, D, lambda_c, t_loc=t_pm, int_kernel=extended, grouping="clusters").get_matrix(sparse=True)
    eig = eigsh(h_emb, k=6, which="SA")[0]
    want = eigsh(non_int + int_term, k=6, which="SA")[0]
    np.testing.assert_almost_equal(min(eig), min(want), decimal=7)


# ---- H2 through the facade (tests/integration/test_integration_fermionic_algebra.py:13-61) ---------------------------------------------
def test_cluster_ops_and_hf_ket_match_the_oracle_generators():
    from oracle import models

    for n_e, n in ((2, 4), (4, 8), (2, 6)):
        ops = qfe.get_cluster_ops(n_e, nqbits=n)
        ref = models.cluster_ops(n_e, n)
        assert len(ops) == len(ref)
        raw = qfe.workloads.uccsd_pool_fermion_terms(n_e, n)  # what the facade wraps, before FermionHamiltonian normal-orders it
        assert [[(complex(c), o, list(q)) for c, o, q in t] for t in raw] == [[(complex(c), o, list(q)) for c, o, q in t] for t in ref]
        for op, terms in zip(ops, ref):
            assert len(op.terms) == 2 and {t.op for t in op.terms} == {terms[0][1]}
            assert sorted(abs(t.coeff) for t in op.terms) == [1.0, 1.0] and sum(t.coeff for t in op.terms) == 0  # i(T - T^dagger)
            assert sorted(sorted(t.qbits) for t in op.terms) == sorted(sorted(q) for _, _, q in terms)
        assert qfe.get_hf_ket(n_e, n) == models.hf_ket(n_e, n)
    assert len(qfe.get_cluster_ops(4, noons=[1.9, 1.8, 0.2, 0.1])) == len(qfe.get_cluster_ops(4, nqbits=8))
    with pytest.raises(TypeError):
        qfe.get_cluster_ops(2)
    with pytest.raises(ValueError):
        qfe.get_cluster_ops(2, nqbits=5)
    with pytest.raises(ValueError):
        qfe.get_cluster_ops(3, nqbits=4)


@pytest.mark.gpu
def test_h2_spectra_and_commutators_across_encodings(golden_dir):
    """H = molecule file -> ElectronicStructureHamiltonian; its spectrum survives to_fermion(); for every UCCSD cluster operator the
    commutator [H, tau] has the same spectrum in Fock space and in the JW / BK / parity spin encodings."""
    h, n_e = qfe.molecule_hamiltonian(golden_dir / "h2_integrals.npz")
    np.testing.assert_almost_equal(sorted(np.linalg.eigvals(h.get_matrix())), sorted(np.linalg.eigvals(h.to_fermion().get_matrix())))
    h_sp = h.to_spin()
    for c_op in qfe.get_cluster_ops(n_e, nqbits=h.nbqbits):
        fermion = np.linalg.eigvals((h | c_op).get_matrix())
        spins = [(h_sp | c_op.to_spin()).get_matrix()] + [(h.to_spin(enc) | c_op.to_spin(enc)).get_matrix() for enc in ("jordan-wigner", "bravyi-kitaev", "parity")]
        for m in spins:
            ev = np.linalg.eigvals(m)
            np.testing.assert_almost_equal(sorted(np.real(fermion)), sorted(np.real(ev)))
            np.testing.assert_almost_equal(sorted(np.imag(fermion)), sorted(np.imag(ev)))
