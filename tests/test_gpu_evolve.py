"""SURVEY 8(f) rows on the B200 (-m gpu): basis states and fused Pauli rotations (Trotter slices, ansatz layers), overlap / QFIM /
energy-gradient evaluation, QSE matrices and the natural-gradient optimizer, each against oracle/evolve.py.  Tolerances as for the
hot path: 1e-10 relative in complex128, 1e-5 in complex64."""

import numpy as np
import pytest

from oracle import evolve, models, spin, transforms

pytestmark = pytest.mark.gpu

RTOL64, RTOL32 = 1e-10, 1e-5


@pytest.fixture(scope="module")
def eng():
    import myqlm_fermion_b200 as qfe

    return qfe.get_engine(0)


def _torch():
    import torch

    return torch


def _state(n, seed=0):
    rng = np.random.default_rng(seed)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return psi / np.linalg.norm(psi)


def _dev(a):
    return _torch().from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def test_basis_state(eng):
    for n, idx in ((1, 1), (5, 19), (12, 0), (12, 4095)):
        psi = eng.basis_state(n, idx).cpu().numpy()
        want = np.zeros(1 << n, dtype=np.complex128)
        want[idx] = 1
        assert np.array_equal(psi, want)
    # a shard that does not hold the occupied amplitude is all zero; the one that does holds it at its local index
    assert not eng.basis_state(6, 40, n_local=4, shard=1).cpu().numpy().any()
    assert eng.basis_state(6, 40, n_local=4, shard=2).cpu().numpy()[8] == 1
    assert eng.basis_state(4, 3, dtype=_torch().complex64).dtype == _torch().complex64


@pytest.mark.parametrize("n", [3, 9, 13, 15])
def test_random_rotations_match_the_oracle(eng, n):
    """Arbitrary X/Z masks (diagonal ones included), many more rotations than one fused batch can hold when n > 13."""
    rng = np.random.default_rng(n)
    K = 60
    x = rng.integers(0, 1 << n, size=K, dtype=np.uint64)
    if n > 13:  # a rotation may flip at most 10 qubits besides the tile's low bits
        keep = rng.integers(0, 1 << n, size=K, dtype=np.uint64) & rng.integers(0, 1 << n, size=K, dtype=np.uint64)
        x &= keep
    x[::7] = 0
    z = rng.integers(0, 1 << n, size=K, dtype=np.uint64)
    theta = rng.uniform(-np.pi, np.pi, size=K)
    psi = _state(n, 1)
    want = evolve.rotations(n, x, z, theta, psi)
    d = _dev(psi)
    eng.pauli_rotations(d, x, z, theta)
    got = d.cpu().numpy()
    assert np.max(np.abs(got - want)) <= RTOL64 * np.max(np.abs(want))
    d32 = _dev(psi.astype(np.complex64))
    eng.pauli_rotations(d32, x, z, theta)
    assert np.max(np.abs(d32.cpu().numpy() - want)) <= RTOL32 * np.max(np.abs(want)) * 10


@pytest.mark.parametrize("n", [15, 17])
def test_wide_rotations_take_the_direct_pass(eng, n):
    """Rotations that flip more qubits than a tile holds (a parity-encoded a+_p a_q carries an X string of length |p - q|;
    the reference's _grow_ansatz / make_spin_hamiltonian_trotter_slice have no width limit): unrestricted random X masks, the full
    register, and a parity-encoded long-range excitation, interleaved with narrow rotations so that the order is exercised."""
    rng = np.random.default_rng(100 + n)
    K = 24
    x = rng.integers(0, 1 << n, size=K, dtype=np.uint64)  # ~n/2 flipped qubits each: most do not fit a tile
    x[3] = np.uint64((1 << n) - 1)
    x[5] = np.uint64(1)
    x[9] = np.uint64(0)
    z = rng.integers(0, 1 << n, size=K, dtype=np.uint64)
    theta = rng.uniform(-np.pi, np.pi, size=K)
    psi = _state(n, 3)
    want = evolve.rotations(n, x, z, theta, psi)
    d = _dev(psi)
    eng.pauli_rotations(d, x, z, theta)
    assert np.max(np.abs(d.cpu().numpy() - want)) <= RTOL64 * np.max(np.abs(want))
    d32 = _dev(psi.astype(np.complex64))
    eng.pauli_rotations(d32, x, z, theta)
    assert np.max(np.abs(d32.cpu().numpy() - want)) <= RTOL32 * np.max(np.abs(want)) * 10
    # i (a+_1 a_{n-2} - h.c.) in the parity encoding: every term flips qubits 1 .. n-2
    op = transforms.transform(n, [(1j, "Cc", [1, n - 2]), (-1j, "Cc", [n - 2, 1])], 0.0, "parity")
    xs, zs, cs, _ = op.canonical()
    assert min(bin(int(v)).count("1") for v in xs) >= n - 3 > 10
    th = 0.37 * np.real(cs)
    want = evolve.rotations(n, xs, zs, th, psi)
    d = _dev(psi)
    eng.pauli_rotations(d, xs, zs, th)
    assert np.max(np.abs(d.cpu().numpy() - want)) <= RTOL64 * np.max(np.abs(want))


def test_rotation_argument_errors(eng):
    import myqlm_fermion_b200 as qfe

    d = _dev(_state(15, 0))
    with pytest.raises(qfe.QfeError):  # acts outside the register
        eng.pauli_rotations(d, [1 << 20], [0], [0.1])
    with pytest.raises(ValueError):
        eng.pauli_rotations(d, [1, 2], [0], [0.1])


def test_trotter_slice_of_a_hubbard_chain(eng):
    """make_spin_hamiltonian_trotter_slice semantics: term order of the Hamiltonian, theta_t = coeff * Re c_t; many steps of a small
    step approach exp(-i t H) (first-order Trotter error, checked loosely) while one slice matches the oracle to rounding."""
    import myqlm_fermion_b200 as qfe

    t_mat = np.zeros((3, 3))
    for i in range(2):
        t_mat[i, i + 1] = t_mat[i + 1, i] = -1.0
    hs = qfe.make_hubbard_model(t_mat, 4.0, 2.0).to_spin("jordan-wigner")
    n = hs.nbqbits
    x, z, c = qfe.pack_terms(n, hs.terms)
    psi = _state(n, 3)
    d = _dev(psi)
    eng.trotter_slice(hs, d, coeff=0.21)
    want = evolve.trotter_slice(n, x, z, c, 0.21, psi)
    assert np.max(np.abs(d.cpu().numpy() - want)) <= RTOL64
    # gate-level restatement of the reference circuit on the same term list
    want_gates = evolve.trotter_slice_circuit(n, [(t.coeff, t.op, t.qbits) for t in hs.terms], 0.21, psi)
    assert np.max(np.abs(d.cpu().numpy() - want_gates)) <= 1e-10
    import scipy.linalg

    hm = spin.csr_from_packed(n, x, z, c, hs.constant_coeff).toarray()
    d = _dev(psi)
    steps, total = 400, 0.5
    eng.trotter_slice(hs, d, coeff=total / steps, n_steps=steps)
    exact = scipy.linalg.expm(-1j * total * (hm - hs.constant_coeff * np.eye(1 << n))) @ psi
    assert np.max(np.abs(d.cpu().numpy() - exact)) < 2e-2
    assert abs(np.linalg.norm(d.cpu().numpy()) - 1) < 1e-12


def test_uccsd_layer_from_the_hartree_fock_state(eng):
    """prod_k exp(-i theta_k tau_k)|HF> with tau_k the (Hermitian) UCCSD cluster operators in JW form, as
    AdaptVQEPlugin._grow_ansatz builds it (adapt_vqe.py:64-141).  The strings of one cluster operator commute, so each factor is
    the exact exponential: compared with dense expm."""
    import scipy.linalg

    import myqlm_fermion_b200 as qfe

    n, n_e = 8, 4
    pool_terms = models.cluster_ops(n_e, n)[:6]
    pool = [transforms.transform(n, op, 0.0, "jordan-wigner").canonical() for op in pool_terms]
    ops = []
    for px, pz, pc, pk in pool:
        assert np.max(np.abs(np.imag(pc))) < 1e-14  # Hermitian: real Pauli coefficients
        terms = []
        for xx, zz, cc in zip(px, pz, pc):
            op, qb = qfe.unpack_term(n, int(xx), int(zz))
            terms.append(qfe.Term(complex(cc), op, qb))
        ops.append(qfe.SpinHamiltonian(n, terms))
    thetas = np.linspace(0.1, 0.6, len(ops))
    hf = models.hf_ket(n_e, n)
    psi = eng.basis_state(n, hf)
    eng.evolve_operators(psi, ops, thetas)
    want = np.zeros(1 << n, dtype=np.complex128)
    want[hf] = 1
    for (px, pz, pc, pk), th in zip(pool, thetas):
        want = scipy.linalg.expm(-1j * th * spin.csr_from_packed(n, px, pz, pc, pk).toarray()) @ want
    assert np.max(np.abs(psi.cpu().numpy() - want)) <= 1e-10
    assert abs(np.linalg.norm(psi.cpu().numpy()) - 1) < 1e-12


def test_gram_qfim_and_energy_gradient(eng):
    n, P = 10, 5
    states = [_state(n, 10 + k) for k in range(P)]
    psi = _state(n, 99)
    d_states = [_dev(s) for s in states]
    d_psi = _dev(psi)
    g = eng.gram(d_states)
    want_g = np.array(states).conj() @ np.array(states).T
    assert np.max(np.abs(g - want_g)) <= 1e-12
    g2 = eng.gram(d_states[:2], d_states[2:])
    assert np.max(np.abs(g2 - want_g[:2, 2:])) <= 1e-12
    assert np.max(np.abs(eng.qfim(d_states, d_psi) - evolve.qfim(states, psi))) <= 1e-11
    one, two = models.synthetic_integrals(5, seed=7)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    terms = eng.terms_from_integrals(hpq, hpqrs)
    x, z, c, _, k = terms.arrays()
    phi = spin.apply_packed(n, x, z, c, complex(k[0]), psi)
    want_grad = np.array([2 * np.vdot(s, phi).real for s in states])
    got_grad = eng.energy_gradient(terms, d_states, d_psi)
    assert np.max(np.abs(got_grad - want_grad)) <= RTOL64 * np.max(np.abs(want_grad))


@pytest.mark.parametrize("n,nb,nk", [(5, 3, 2), (10, 5, 3), (12, 40, 0), (11, 33, 35), (4, 2, 2)])
def test_batched_gram_kernel(eng, n, nb, nk):
    """qfe_vec_gram: G[i, j] = <b_i|k_j> in one pass per 32 x 32 panel, against numpy; nk = 0 stands for the Hermitian Gram matrix of
    the bras.  Panel edges (33, 35, 40 vectors), registers below the 32-amplitude chunk (n = 4) and complex64 included."""
    rng = np.random.default_rng(n * 100 + nb)
    B = rng.standard_normal((nb, 1 << n)) + 1j * rng.standard_normal((nb, 1 << n))
    Kk = B if nk == 0 else rng.standard_normal((nk, 1 << n)) + 1j * rng.standard_normal((nk, 1 << n))
    want = np.conj(B) @ Kk.T
    db = [_dev(v) for v in B]
    dk = None if nk == 0 else [_dev(v) for v in Kk]
    got = eng.gram(db, dk)
    assert got.shape == want.shape and np.max(np.abs(got - want)) <= RTOL64 * np.max(np.abs(want))
    if nk == 0:
        assert np.array_equal(got, np.conj(got.T))  # mirrored, not recomputed
    B32, K32 = B.astype(np.complex64), Kk.astype(np.complex64)
    got32 = eng.gram([_dev(v) for v in B32], None if nk == 0 else [_dev(v) for v in K32])
    want32 = np.conj(B32.astype(np.complex128)) @ K32.astype(np.complex128).T
    assert np.max(np.abs(got32 - want32)) <= RTOL64 * np.max(np.abs(want32))  # complex64 inputs, double accumulation


def test_qse_matrices(eng):
    """qse.py:191-214 on the 2x2 Hubbard plaquette ground state region: H and S from device overlaps vs the dense definition."""
    import myqlm_fermion_b200 as qfe

    n = 6
    t_mat = np.zeros((3, 3))
    for i in range(2):
        t_mat[i, i + 1] = t_mat[i + 1, i] = -1.0
    hs = qfe.make_hubbard_model(t_mat, 2.0, 1.0).to_spin("jordan-wigner")
    ops = [
        qfe.SpinHamiltonian(n, [], constant_coeff=1.0),
        qfe.SpinHamiltonian(n, [qfe.Term(1.0, "XX", [0, 2]), qfe.Term(0.5, "YY", [0, 2])]),
        qfe.SpinHamiltonian(n, [qfe.Term(1.0, "Z", [1]), qfe.Term(-0.25j, "XY", [3, 5])]),
    ]
    psi = _state(n, 5)
    mh, ms = qfe.plugins.qse_matrices(hs, _dev(psi), ops, engine=eng)
    hx, hz, hc, hk = hs.packed_arrays()
    packed_ops = [o.packed_arrays() for o in ops]
    want_h, want_s = evolve.qse_matrices(n, (hx, hz, hc, hk), packed_ops, psi)
    assert np.max(np.abs(mh - want_h)) <= 1e-11 and np.max(np.abs(ms - want_s)) <= 1e-12


@pytest.mark.parametrize("natural", [True, False])
def test_gradient_descent_optimizer_follows_the_reference_update(eng, natural):
    """Two-parameter ansatz exp(-i t1 Y0) exp(-i t0 Y1 X0)|00> on H = XX + YY + ZZ (the Hamiltonian of the reference's natural
    gradient test, tests/integration/test_integration_natural_gradient_plugin.py:30): same trajectory as the dense restatement of
    GradientDescentOptimizer.optimize, and the energy decreases."""
    import myqlm_fermion_b200 as qfe

    n = 2
    hs = qfe.SpinHamiltonian(n, [qfe.Term(1, op, [0, 1]) for op in ["XX", "YY", "ZZ"]])
    gx, gz, _ = qfe.pack_terms(n, [qfe.Term(1, "XY", [0, 1]), qfe.Term(1, "Y", [0])])

    def dense(params):
        psi0 = np.zeros(4, dtype=np.complex128)
        psi0[0] = 1
        a = evolve.rotate(n, int(gx[0]), int(gz[0]), params[0], psi0)
        psi = evolve.rotate(n, int(gx[1]), int(gz[1]), params[1], a)
        # d/dt exp(-i t P) = -i P exp(-i t P)
        d0 = evolve.rotate(n, int(gx[1]), int(gz[1]), params[1], -1j * spin.apply_packed(n, [gx[0]], [gz[0]], [1.0], 0.0, a))
        d1 = -1j * spin.apply_packed(n, [gx[1]], [gz[1]], [1.0], 0.0, psi)
        return psi, [d0, d1]

    def device(params):
        psi, dps = dense(params)  # derivative states are supplied by the state-preparation side
        check = eng.basis_state(n, 0)
        eng.pauli_rotations(check, gx, gz, params)
        assert np.max(np.abs(check.cpu().numpy() - psi)) < 1e-12
        return check, [_dev(d) for d in dps]

    x0 = [0.3, -0.4]
    hx, hz, hc, hk = hs.packed_arrays()
    hm = spin.csr_from_packed(n, hx, hz, hc, hk).toarray()
    want_e, want_p, want_trace = evolve.gradient_descent(hm, dense, x0, maxiter=25, natural_gradient=natural, lambda_step=0.2, tol=1e-10)
    opt = qfe.plugins.GradientDescentOptimizer(maxiter=25, natural_gradient=natural, lambda_step=0.2, tol=1e-10, x0=x0, engine=eng)
    e, p, info = opt.optimize(["t0", "t1"], hs, device)
    assert info["iterations"] == len(want_trace) - 1
    assert np.max(np.abs(np.array(opt.trace) - np.array(want_trace))) <= 1e-9
    assert np.max(np.abs(np.array(p) - want_p)) <= 1e-8
    assert e < opt.trace[0] - 0.1


def test_adapt_vqe_loop_on_h2_reaches_the_exact_energy(eng, golden_dir):
    """tests/integration/test_integration_adaptvqe.py:78-121 with every step on the device: H2 fixture, UCCSD pool, Hartree-Fock start,
    gradients -> grow ansatz -> COBYLA, until the gradient vanishes.  The reference asserts 2 decimals against the exact
    diagonalisation; UCCSD is exact for two electrons, so the loop is held to 1e-6 here."""
    import warnings

    import myqlm_fermion_b200 as qfe

    d = qfe.load_molecule_npz(golden_dir / "h2_integrals.npz")
    h_el, n_e = qfe.molecule_hamiltonian(d)
    hs = h_el.to_spin("jordan-wigner")
    n = hs.nbqbits
    pool = []
    for op in models.cluster_ops(n_e, n):
        px, pz, pc, pk = transforms.transform(n, op, 0.0, "jordan-wigner").canonical()
        pool.append(qfe.SpinHamiltonian(n, [qfe.Term(complex(c), *qfe.unpack_term(n, int(x), int(z))) for x, z, c in zip(px, pz, pc)]))
    hx, hz, hc, hk = hs.packed_arrays()
    exact = np.linalg.eigvalsh(spin.csr_from_packed(n, hx, hz, hc, hk).toarray()).min()
    assert abs(exact - float(d["fci"])) < 1e-8
    adapt = qfe.plugins.AdaptVQE(pool, n_iterations=30, engine=eng)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # "Norm of the energy gradient is below the set threshold": the normal way out
        res = adapt.run(hs, models.hf_ket(n_e, n))
    np.testing.assert_almost_equal(exact, res.value, decimal=2)  # the reference's criterion
    assert abs(res.value - exact) < 1e-6
    order = eval(res.meta_data["operator_order"])
    trace = eval(res.meta_data["energy_trace"])
    assert 1 <= len(order) <= len(pool) and len(trace) == len(order) and all(isinstance(v, str) for v in res.meta_data.values())
    assert all(b <= a + 1e-9 for a, b in zip(trace, trace[1:]))  # the energy never rises when the ansatz grows


def _shallow_circuit_on_device(eng, theta):
    """make_shallow_circ (/root/reference/qat/fermion/circuits.py:207-232) on |0000> with Pauli rotations only:
    RY(t) = exp(-i t/2 Y);  CNOT(c, t) = e^{-i pi/4} exp(+i pi/4 Z_c) exp(+i pi/4 X_t) exp(-i pi/4 Z_c X_t) (global phase dropped)."""
    import myqlm_fermion_b200 as qfe

    n = 4
    gates = []

    def ry(q, t):
        gates.append((qfe.Term(1.0, "Y", [q]), t / 2))

    def cnot(c, t):
        gates.extend([(qfe.Term(1.0, "Z", [c]), -np.pi / 4), (qfe.Term(1.0, "X", [t]), -np.pi / 4), (qfe.Term(1.0, "ZX", [c, t]), np.pi / 4)])

    for i in range(4):
        ry(i, theta[i])
    cnot(2, 3)
    ry(2, theta[4])
    ry(3, theta[5])
    cnot(0, 2)
    ry(0, theta[6])
    ry(2, theta[7])
    cnot(0, 1)
    x, z, c = qfe.pack_terms(n, [g for g, _ in gates])
    assert np.all(c == 1)
    psi = eng.basis_state(n, 0)
    eng.pauli_rotations(psi, x, z, [a for _, a in gates])
    return psi


def _embedded_hamiltonian(eng):
    hpq = np.array([[0.5, 1, 0, 0], [1, 1, 0, 0], [0, 0, 0.5, 1], [0, 0, 1, 1]], dtype=float)
    hpqrs = np.zeros((4, 4, 4, 4))
    hpqrs[0, 2, 0, 2] = hpqrs[2, 0, 2, 0] = -1
    return eng.terms_from_integrals(hpq, hpqrs)


def test_golden_energies_with_the_circuit_built_on_the_device(eng, goldens):
    """The reference's two 12-decimal goldens with every step on the GPU: the ZNE test's energy of the shallow circuit at
    np.random.seed(0) angles (tests/integration/test_integration_zne_plugin.py:16-49) and the SeqOptim (rotosolve) end energy from
    seed-1 angles after 10 cycles (tests/integration/test_integration_seqoptim.py:16-33)."""
    import myqlm_fermion_b200 as qfe
    from oracle import circuits

    terms = _embedded_hamiltonian(eng)
    plan = eng.plan(terms)
    np.random.seed(0)
    theta = np.random.random(8)
    psi = _shallow_circuit_on_device(eng, theta)
    want_psi = circuits.shallow_circuit_state(theta)
    got_psi = psi.cpu().numpy()
    phase = np.vdot(want_psi, got_psi)
    assert abs(abs(phase) - 1) < 1e-12 and np.max(np.abs(got_psi - phase * want_psi)) < 1e-12  # equal up to the dropped global phase
    g = goldens["scalars"]["zne_energy_shallow_seed0"]
    np.testing.assert_almost_equal(eng.energy(plan, psi), g["value"], decimal=g["decimals"])

    np.random.seed(1)
    x0 = np.random.random(8)
    names = [f"theta_{i}" for i in range(8)]
    opt = qfe.plugins.SeqOptim(ncycles=10, x0=x0)
    e, params, _ = opt.optimize(names, lambda p: eng.energy(plan, _shallow_circuit_on_device(eng, [p[k] for k in names])))
    g = goldens["scalars"]["seqoptim_energy_seed1"]
    np.testing.assert_almost_equal(e, g["value"], decimal=g["decimals"])
    assert opt.n_evaluations == 1 + 10 * 8 * 3


@pytest.mark.parametrize("natural", [False, True])
def test_gradient_descent_reaches_the_exact_energy_like_the_reference_test(eng, natural):
    """tests/integration/test_integration_natural_gradient_plugin.py:16-83: one-site Hubbard model (2 qubits), ansatz
    H(0) RY(t0, 0) CNOT(0, 1) RX(t1, 1), x0 from np.random.seed(1234), maxiter 100, step 0.2, tol 1e-7; the optimised energy equals the
    exact one to 4 decimals.  State and derivative states are built on the device from Pauli rotations and single-string applies."""
    import myqlm_fermion_b200 as qfe

    hs = qfe.make_hubbard_model(np.zeros((1, 1)), 2.0, mu=1.0).to_spin()
    exact = min(np.linalg.eigh(hs.get_matrix())[0])
    T = qfe.Term
    pack = lambda terms: qfe.pack_terms(2, terms)[:2]
    # H = RY(pi/2) Z and CNOT = exp(i pi/4 Z_c) exp(i pi/4 X_t) exp(-i pi/4 Z_c X_t), both up to a global phase
    pre = [(T(1.0, "Z", [0]), np.pi / 2), (T(1.0, "Y", [0]), np.pi / 4)]
    cnot = [(T(1.0, "Z", [0]), -np.pi / 4), (T(1.0, "X", [1]), -np.pi / 4), (T(1.0, "ZX", [0, 1]), np.pi / 4)]
    y0 = eng.terms_from_pauli(2, *pack([T(1.0, "Y", [0])]), [1.0])
    x1 = eng.terms_from_pauli(2, *pack([T(1.0, "X", [1])]), [1.0])

    def rot(psi, gates):
        x, z = pack([g for g, _ in gates])
        eng.pauli_rotations(psi, x, z, [a for _, a in gates])

    def state_and_derivatives(theta):
        a = eng.basis_state(2, 0)
        rot(a, pre + [(T(1.0, "Y", [0]), theta[0] / 2)])
        psi = a.clone()
        rot(psi, cnot + [(T(1.0, "X", [1]), theta[1] / 2)])
        d0 = eng.apply(y0, a)  # d/dt0 RY(t0) = -i/2 Y RY(t0)
        eng.scale(d0, -0.5j)
        rot(d0, cnot + [(T(1.0, "X", [1]), theta[1] / 2)])
        d1 = eng.apply(x1, psi)
        eng.scale(d1, -0.5j)
        return psi, [d0, d1]

    np.random.seed(1234)
    x0 = np.random.uniform(0, 2 * np.pi, 2)
    # the derivative states are consistent with finite differences of the state
    psi, dps = state_and_derivatives(x0)
    for k in range(2):
        step = np.zeros(2)
        step[k] = 1e-6
        fd = (state_and_derivatives(x0 + step)[0] - state_and_derivatives(x0 - step)[0]).cpu().numpy() / 2e-6
        assert np.max(np.abs(fd - dps[k].cpu().numpy())) < 1e-8
    opt = qfe.plugins.GradientDescentOptimizer(maxiter=100, lambda_step=0.2, natural_gradient=natural, tol=1e-7, x0=x0, engine=eng)
    value, params, info = opt.optimize(["\\theta_0", "\\theta_1"], hs, state_and_derivatives)
    np.testing.assert_almost_equal(value, exact, decimal=4)
    assert info["iterations"] <= 100 and len(info["energies"]) == info["iterations"] + 1


def test_qse_on_the_vqe_state_of_the_reference_example(eng):
    """tests/integration/test_qse.py: H = XX + YY + ZZ, ansatz RY(t0, 0) RY(t1, 1) RZ(t2, 1) CNOT(0, 1) optimised by SeqOptim from
    [0, 0.5, 0] (10 cycles), expansion {1, ZZ}: the QSE energy is -3; and the linear Pauli expansion's matrices."""
    import myqlm_fermion_b200 as qfe

    T = qfe.Term
    hs = qfe.SpinHamiltonian(2, [T(1, op, [0, 1]) for op in ["XX", "YY", "ZZ"]])
    plan = eng.plan(hs.packed(eng))
    cnot = [(T(1.0, "Z", [0]), -np.pi / 4), (T(1.0, "X", [1]), -np.pi / 4), (T(1.0, "ZX", [0, 1]), np.pi / 4)]

    def state(p):
        gates = [(T(1.0, "Y", [0]), p[0] / 2), (T(1.0, "Y", [1]), p[1] / 2), (T(1.0, "Z", [1]), p[2] / 2)] + cnot
        x, z, _ = qfe.pack_terms(2, [g for g, _ in gates])
        psi = eng.basis_state(2, 0)
        eng.pauli_rotations(psi, x, z, [a for _, a in gates])
        return psi

    names = ["t0", "t1", "t2"]
    e_vqe, params, _ = qfe.plugins.SeqOptim(ncycles=10, x0=[0, 0.5, 0]).optimize(names, lambda p: eng.energy(plan, state([p[k] for k in names])))
    assert e_vqe <= -2.9
    expansion = [qfe.SpinHamiltonian(2, [], 1.0), qfe.SpinHamiltonian(2, [T(1.0, "ZZ", [0, 1])])]
    e_qse = qfe.plugins.apply_quantum_subspace_expansion(hs, state(params), expansion, engine=eng)
    np.testing.assert_almost_equal(e_qse, -3.0)
    e2, mh, ms = qfe.plugins.apply_quantum_subspace_expansion(hs, state(params), expansion, engine=eng, return_matrices=True)
    assert e2 == e_qse and mh.shape == ms.shape == (2, 2) and abs(ms[0, 0] - 1) < 1e-12
    ops = qfe.plugins.build_linear_pauli_expansion(["X", "Y", "Z", "Z", "I"], 1)
    want = [np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.array([[1, 0], [0, -1]]), np.array([[1, 0], [0, -1]]), np.eye(2)]
    np.testing.assert_equal([o.get_matrix() for o in ops], [w.astype(np.complex128) for w in want])


def test_ucc_on_h2_reproduces_the_reference_energy_to_8_decimals(eng, golden_dir, goldens):
    """tests/integration/test_integration_ucc.py:115-175 (test_more_basic): full H2 Hamiltonian in JW form, UCCSD cluster operators,
    construct_ucc_ansatz on the Hartree-Fock state, COBYLA(tol=1e-15, maxiter=1000) -> -1.1372701746609022 to 8 decimals.  (The
    reference starts from an MP2 guess; chemistry pre-processing is out of scope here and the search starts from zero.)"""
    import scipy.optimize

    import myqlm_fermion_b200 as qfe

    d = qfe.load_molecule_npz(golden_dir / "h2_integrals.npz")
    hpq, hpqrs = qfe.convert_to_h_integrals(d["one_body_integrals"], d["two_body_integrals"])
    h = qfe.ElectronicStructureHamiltonian(hpq, hpqrs, constant_coeff=float(d["nuclear_repulsion"]))
    h_sp = qfe.transform_to_jw_basis(h)
    cluster_ops_sp = [qfe.transform_to_jw_basis(t) for t in qfe.get_cluster_ops(int(d["n_electrons"]), nqbits=h_sp.nbqbits)]
    hf_init_sp = qfe.recode_integer(qfe.get_hf_ket(int(d["n_electrons"]), h_sp.nbqbits), qfe.get_jw_code(h_sp.nbqbits))
    ansatz = qfe.construct_ucc_ansatz(cluster_ops_sp, hf_init_sp, engine=eng)
    plan = eng.plan(h_sp.packed(eng))
    res = scipy.optimize.minimize(lambda theta: eng.energy(plan, ansatz(theta)), x0=np.zeros(len(cluster_ops_sp)), tol=1e-15, method="COBYLA",
                                  options={"maxiter": 1000})
    np.testing.assert_almost_equal(res.fun, -1.1372701746609022, decimal=8)
    g = goldens["scalars"].get("h2_ucc_vqe_energy")
    if g is not None:
        np.testing.assert_almost_equal(res.fun, g["value"], decimal=g["decimals"])


def test_vqe_function_on_the_reference_examples(eng):
    """tests/integration/test_integration_vqe.py:33-73: one- and two-qubit Hamiltonians from integrals through transform_to_jw_basis,
    RY / RX ansatz circuits, a derivative-free optimiser with the reference's (theta, energy, nfev, trace) return convention (the
    reference's SPSA lives in the absent qat.vsolve; Nelder-Mead plays its role), energies to 3 decimals / 5e-3."""
    import random

    import scipy.optimize

    import myqlm_fermion_b200 as qfe

    T = qfe.Term

    def optimizer(fun, theta0):
        res = scipy.optimize.minimize(fun, theta0, method="Nelder-Mead", options={"xatol": 1e-7, "fatol": 1e-9, "maxiter": 2000})
        return res.x, float(res.fun), res.nfev, []

    def rotations(n, gates):
        x, z, _ = qfe.pack_terms(n, [g for g, _ in gates])
        psi = eng.basis_state(n, 0)
        eng.pauli_rotations(psi, x, z, [a for _, a in gates])
        return psi

    random.seed(7)
    hamilt_sp = qfe.transform_to_jw_basis(qfe.ElectronicStructureHamiltonian(hpq=np.ones((1, 1)), hpqrs=np.zeros((1, 1, 1, 1))))
    energy, theta, _, _ = qfe.plugins.VQE(hamilt_sp, optimizer, lambda x: rotations(1, [(T(1.0, "Y", [0]), x[0] / 2)]),
                                          np.array([random.random() * 2 * np.pi]), engine=eng)
    assert abs(energy - 0.0) < 1e-3
    hpq = np.array([[-3.0, 2.0], [2.0, -3.0]])
    hamiltonian = qfe.ElectronicStructureHamiltonian(hpq, np.zeros((2, 2, 2, 2)))
    want = min(np.linalg.eigvalsh(hamiltonian.get_matrix()))
    energy, theta, _, _ = qfe.plugins.VQE(qfe.transform_to_jw_basis(hamiltonian), optimizer,
                                          lambda x: rotations(2, [(T(1.0, "Y", [0]), x[0] / 2), (T(1.0, "X", [1]), x[1] / 2)]),
                                          np.array([random.random() * 2 * np.pi, random.random() * 2 * np.pi]), engine=eng)
    assert abs(want + 6.0) < 1e-12 and abs(energy - want) < 5e-3  # |11> (both modes filled) is a product state the ansatz reaches
