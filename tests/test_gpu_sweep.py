"""Kernel (2) parity: matrix-free <psi|H|psi>, H|psi>, batched transition elements and the ADAPT
commutator-gradient sweep vs the oracle's get_matrix()-style path.  Tolerances from
BASELINE.json: 1e-10 relative in complex128, 1e-5 in complex64.  Runs on the B200 (-m gpu)."""

import numpy as np
import pytest

from oracle import models, plugins, spin, transforms

pytestmark = pytest.mark.gpu

RTOL64, RTOL32 = 1e-10, 1e-5


@pytest.fixture(scope="module")
def eng():
    import myqlm_fermion_b200 as qfe

    return qfe.get_engine(0)


def _torch():
    import torch

    return torch


def _state(n, seed=0, dtype=np.complex128):
    rng = np.random.default_rng(seed)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    return psi.astype(dtype)


def _dev(a):
    return _torch().from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def _molecule(eng, m, enc="jordan-wigner", seed=1234, constant=0.0):
    one, two = models.synthetic_integrals(m, seed=seed)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    return eng.terms_from_integrals(hpq, hpqrs, constant=constant, encoding=enc)


def _random_pauli(n, T, seed, hermitian=True):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 1 << n, size=T, dtype=np.uint64)
    z = rng.integers(0, 1 << n, size=T, dtype=np.uint64)
    c = rng.standard_normal(T) + (0 if hermitian else 1j * rng.standard_normal(T))
    return x, z, c.astype(np.complex128)


@pytest.mark.parametrize("enc", ["jordan-wigner", "bravyi-kitaev", "parity"])
@pytest.mark.parametrize("m", [4, 5, 6, 7])
def test_molecular_expectation_and_apply(eng, m, enc):
    """Config 2 family (m=7 is the 14-qubit H2O-sized case): energy and H|psi> in both dtypes,
    tile kernel and direct kernel."""
    n = 2 * m
    terms = _molecule(eng, m, enc, constant=0.7)
    x, z, c, _, k = terms.arrays()
    psi = _state(n)
    want_phi = spin.apply_packed(n, x, z, c, complex(k[0]), psi)
    want_e = np.vdot(psi, want_phi)
    scale = np.max(np.abs(want_phi))
    for tile_bits in (0, 9, -1):  # default tiles, small tiles (many sweeps), direct kernel
        plan = eng.plan(terms, tile_bits=tile_bits)
        d = _dev(psi)
        got_e = eng.expectation(plan, d)
        assert abs(got_e - want_e) <= RTOL64 * abs(want_e), (tile_bits, got_e, want_e)
        got_phi = eng.apply(plan, d).cpu().numpy()
        assert np.max(np.abs(got_phi - want_phi)) <= RTOL64 * scale, tile_bits
        d32 = _dev(psi.astype(np.complex64))
        got_e32 = eng.expectation(plan, d32)
        psi32 = psi.astype(np.complex64).astype(np.complex128)
        want_e32 = np.vdot(psi32, spin.apply_packed(n, x, z, c, complex(k[0]), psi32))
        assert abs(got_e32 - want_e32) <= RTOL32 * abs(want_e32), (tile_bits, got_e32, want_e32)
        got_phi32 = eng.apply(plan, d32).cpu().numpy()
        assert np.max(np.abs(got_phi32 - want_phi)) <= RTOL32 * scale * 10


def test_energy_matches_reference_matrix_path(eng):
    """The reference's own path: get_matrix(sparse=True) as a kron chain, then vdot(psi, H psi)."""
    n = 8
    terms = _molecule(eng, 4)
    x, z, c, _, k = terms.arrays()
    ps = spin.PauliSum(n)
    for xi, zi, ci in zip(x.tolist(), z.tolist(), c.tolist()):
        ps.d[(xi, zi)] = ci * (1j) ** (spin.popcount(xi & zi) % 4)
    mat = spin.spin_get_matrix(n, ps.term_triples(eps=0.0), complex(k[0]), sparse=True)
    psi = _state(n, 3)
    want = np.vdot(psi, mat @ psi)
    got = eng.expectation(terms, _dev(psi))
    assert abs(got - want) <= RTOL64 * abs(want)


@pytest.mark.parametrize("n", [1, 2, 5, 7, 8, 9, 13, 16])
def test_random_pauli_sums_all_sizes(eng, n):
    """Arbitrary (non-molecular, complex-coefficient) Pauli sums, including shards smaller than a
    tile and X masks wider than a tile."""
    for herm in (True, False):
        T = min(60, 4**n - 1)
        x, z, c = _random_pauli(n, T, seed=n, hermitian=herm)
        terms = eng.terms_from_pauli(n, x, z, c, constant=0.3 - 0.1j)
        gx, gz, gc, _, gk = terms.arrays()
        psi, bra = _state(n, 1), _state(n, 2)
        want_phi = spin.apply_packed(n, gx, gz, gc, complex(gk[0]), psi)
        for tile_bits in (0, 8, -1):
            plan = eng.plan(terms, tile_bits=tile_bits)
            got = eng.expectation(plan, _dev(psi), bra=_dev(bra))
            want = np.vdot(bra, want_phi)
            assert abs(got - want) <= RTOL64 * max(abs(want), 1e-3), (n, tile_bits, got, want)
            got = eng.expectation(plan, _dev(psi))
            want = np.vdot(psi, want_phi)
            assert abs(got - want) <= RTOL64 * max(abs(want), 1e-3), (n, tile_bits, got, want)
            got_phi = eng.apply(plan, _dev(psi)).cpu().numpy()
            assert np.max(np.abs(got_phi - want_phi)) <= RTOL64 * np.max(np.abs(want_phi))


def test_basis_state_and_product_state_energies(eng):
    """Size-independent checks usable at 32 qubits, validated here at 14."""
    n = 14
    terms = _molecule(eng, 7)
    x, z, c, _, k = terms.arrays()
    # Hartree-Fock-like basis state: only diagonal terms contribute
    b = models.hf_ket(6, n)
    e_basis = complex(k[0]) + sum(ci * (-1) ** spin.popcount(int(zi) & b) for xi, zi, ci in zip(x, z, c) if xi == 0)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[b] = 1.0
    got = eng.expectation(terms, _dev(psi))
    assert abs(got - e_basis) <= RTOL64 * abs(e_basis)
    # random product state: <P> = prod_q <sigma_q>
    rng = np.random.default_rng(11)
    ab = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    ab /= np.linalg.norm(ab, axis=1, keepdims=True)
    d = eng.product_state(n, ab)
    sig = {
        (0, 0): np.ones(n),
        (1, 0): 2 * np.real(np.conj(ab[:, 0]) * ab[:, 1]),
        (1, 1): 2 * np.imag(np.conj(ab[:, 0]) * ab[:, 1]),
        (0, 1): np.abs(ab[:, 0]) ** 2 - np.abs(ab[:, 1]) ** 2,
    }
    want = complex(k[0])
    for xi, zi, ci in zip(x.tolist(), z.tolist(), c.tolist()):
        v = ci
        for q in range(n):
            bit = 1 << (n - 1 - q)
            v *= sig[(1 if xi & bit else 0, 1 if zi & bit else 0)][q]
        want += v
    got = eng.expectation(terms, d)
    assert abs(got - want) <= RTOL64 * abs(want)
    # <psi|H|psi> is real for a Hermitian H and equals <psi|phi>
    psi_r = eng.random_state(n, seed=5)
    e = eng.expectation(terms, psi_r)
    phi = eng.apply(terms, psi_r)
    assert abs(e.imag) <= 1e-12 * abs(e)
    assert abs(eng.dot(psi_r, phi) - e) <= RTOL64 * abs(e)


@pytest.mark.parametrize("n,ne", [(8, 4), (12, 4), (14, 10)])
def test_commutator_gradients_vs_expanded_commutators(eng, n, ne):
    """ADAPT-VQE gradient sweep (adapt_vqe.py:50-62,181-185) over the FULL UCCSD pools (K = 92 at n = 12, K = 140
    at n = 14): the oracle expands every [H, tau_k] as the reference does and evaluates -i<psi|[H,tau_k]|psi>.  At n = 8 the oracle runs
    here (matrix route); the two larger sweeps take minutes on a CPU and come from the committed fixture
    tests/golden/adapt_gradients_full_pools.json (made by tests/golden/make_golden_adapt.py from the same generators and seed)."""
    one, two = models.synthetic_integrals(n // 2, seed=1234)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    pool_f = models.cluster_ops(ne, n)
    assert len(pool_f) == {(8, 4): 26, (12, 4): 92, (14, 10): 140}[(n, ne)]
    psi = _state(n, 9)
    if n == 8:
        h_ref = transforms.transform(n, models.electronic_terms(hpq, hpqrs), 0.0, "jordan-wigner")
        pool_ref = [transforms.transform(n, op, 0.0, "jordan-wigner") for op in pool_f]
        want = plugins.commutator_gradients(h_ref, pool_ref, psi)
        assert np.max(np.abs(want.imag)) < 1e-10
        want = want.real
    else:
        import json
        from pathlib import Path

        gold = json.loads((Path(__file__).parent / "golden" / "adapt_gradients_full_pools.json").read_text())
        case = [c for c in gold["cases"] if c["n"] == n and c["n_electrons"] == ne][0]
        assert case["K"] == len(pool_f)
        want = np.array(case["gradients"])
    h = eng.terms_from_integrals(hpq, hpqrs)
    pool = eng.concat([eng.terms_from_fermion(n, op) for op in pool_f])
    assert pool.n_slots == len(pool_f)
    for tile_bits in (0, -1):
        got = eng.commutator_gradients(eng.plan(h, tile_bits=tile_bits), eng.plan(pool, tile_bits=tile_bits), _dev(psi))
        assert np.max(np.abs(got - want)) <= RTOL64 * np.max(np.abs(want)), tile_bits
    sel, nrm = plugins.select_operator(want)
    assert sel == int(np.argmax(np.abs(got))) and abs(nrm - np.linalg.norm(got)) <= RTOL64 * nrm
    got32 = eng.commutator_gradients(h, pool, _dev(psi.astype(np.complex64)))
    assert np.max(np.abs(got32 - want)) <= 10 * RTOL32 * np.max(np.abs(want))


def test_batched_expectations(eng):
    """K observables on one state in one call (the natural-gradient / QSE evaluation shape)."""
    n = 10
    ops = []
    want = []
    psi = _state(n, 4)
    for s in range(7):
        x, z, c = _random_pauli(n, 5 + s, seed=100 + s, hermitian=False)
        t = eng.terms_from_pauli(n, x, z, c, constant=0.1 * s)
        ops.append(t)
        gx, gz, gc, _, gk = t.arrays()
        want.append(spin.expectation_packed(n, gx, gz, gc, complex(gk[0]), psi))
    batch = eng.concat(ops)
    for tile_bits in (0, -1):
        got = eng.expectation(eng.plan(batch, tile_bits=tile_bits), _dev(psi))
        assert np.max(np.abs(got - np.array(want))) <= RTOL64 * np.max(np.abs(want))


def test_sharded_classes_reassemble(eng):
    """Shard geometry on ONE GPU: split a 12-qubit vector into 4 shards by its 2 high qubits and
    recombine the per-class partial results; must equal the unsharded result."""
    n, g = 12, 2
    n_local = n - g
    terms = _molecule(eng, 6, "bravyi-kitaev", constant=0.2)
    x, z, c, _, k = terms.arrays()
    psi = _state(n, 6)
    want_phi = spin.apply_packed(n, x, z, c, complex(k[0]), psi)
    want_e = np.vdot(psi, want_phi)
    shards = [_dev(psi[r << n_local : (r + 1) << n_local]) for r in range(1 << g)]
    total = 0.0
    phi = np.zeros_like(psi)
    for r in range(1 << g):
        plan = eng.plan(terms, n_local=n_local, shard=r)
        y = _torch().zeros_like(shards[r])
        first = True
        for x_hi in plan.classes:
            total += eng.expectation_class(plan, x_hi, shards[r], shards[r ^ x_hi])
            eng.apply_class(plan, x_hi, shards[r ^ x_hi], y, accumulate=not first)
            first = False
        phi[r << n_local : (r + 1) << n_local] = y.cpu().numpy()
    assert abs(total - want_e) <= RTOL64 * abs(want_e)
    assert np.max(np.abs(phi - want_phi)) <= RTOL64 * np.max(np.abs(want_phi))


def test_host_path_and_errors(eng):
    import myqlm_fermion_b200 as qfe

    n = 10
    terms = _molecule(eng, 5)
    psi = _state(n, 2)
    x, z, c, _, k = terms.arrays()
    want = spin.expectation_packed(n, x, z, c, complex(k[0]), psi)
    got = eng.expectation_host(terms, psi)
    assert abs(got - want) <= RTOL64 * abs(want)
    with pytest.raises(ValueError):
        eng.expectation(terms, _dev(psi[:-1].copy()))
    with pytest.raises(TypeError):
        eng.expectation(terms, _dev(psi.real.copy()))
    with pytest.raises(qfe.QfeError):
        d = _dev(psi)
        eng.apply(terms, d, out=d)
