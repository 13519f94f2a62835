"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle's matrix path is infeasible there):

* config 4, 32 qubits / 94 848 terms: energy of a random product state = sum_t c_t prod_q <sigma_q> (O(T n) on the CPU) and energy
  of a Hartree-Fock basis state = constant + sum over the diagonal terms;
* config 3, 24-qubit Hubbard chain: Lanczos eigenpair residual |H psi - E psi|, <psi|H|psi> real and equal to <psi|phi>, complex64
  run within 1e-5;
* config 5 shape on one GPU (26 qubits, the same UCCSD pool generator): the batched gradient sweep against the same numbers from
  single-operator plans (different kernel mode), and g_k = 2 Im <H psi|tau_k|psi>.

Runs on the B200 (-m gpu); about two minutes."""

import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL64, RTOL32 = 1e-10, 1e-5


@pytest.fixture(scope="module")
def eng():
    import myqlm_fermion_b200 as qfe

    return qfe.get_engine(0)


def _free_gb():
    import gc

    import torch

    gc.collect()
    torch.cuda.empty_cache()  # earlier tests' vectors sit in torch's caching allocator
    return torch.cuda.mem_get_info(0)[0] / 2**30


def test_config4_product_and_basis_state_energies_at_32_qubits(eng):
    import torch

    from myqlm_fermion_b200 import workloads
    from oracle import models

    n = 32
    if _free_gb() < 80:
        pytest.skip("needs 69 GB for a 32-qubit complex128 state")
    hpq, hpqrs = workloads.synthetic_h_integrals(n // 2, seed=1234)
    terms = eng.terms_from_integrals(hpq, hpqrs, constant=0.25)
    del hpqrs
    assert terms.n_terms == 94848 and terms.n_xmasks == 18281  # T = 94 849 with the identity, 18 281 X masks
    x, z, c, _, k = terms.arrays()
    plan = eng.plan(terms)
    # random product state: <P> = prod_q <sigma_q>, vectorised over the terms
    rng = np.random.default_rng(11)
    ab = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    ab /= np.linalg.norm(ab, axis=1, keepdims=True)
    sx = 2 * np.real(np.conj(ab[:, 0]) * ab[:, 1])
    sy = 2 * np.imag(np.conj(ab[:, 0]) * ab[:, 1])
    sz = np.abs(ab[:, 0]) ** 2 - np.abs(ab[:, 1]) ** 2
    val = c.astype(np.complex128).copy()
    for q in range(n):
        bit = np.uint64(1 << (n - 1 - q))
        hx, hz = (x & bit) != 0, (z & bit) != 0
        val *= np.where(hx & hz, sy[q], np.where(hx, sx[q], np.where(hz, sz[q], 1.0)))
    want = complex(k[0]) + val.sum()
    psi = eng.product_state(n, ab)
    got = eng.expectation(plan, psi)
    assert abs(got - want) <= RTOL64 * abs(want), (got, want)
    del psi
    torch.cuda.empty_cache()
    # Hartree-Fock determinant (16 electrons): only the diagonal terms contribute
    b = models.hf_ket(16, n)
    diag = x == 0
    par = np.array([bin(int(v) & b).count("1") & 1 for v in z[diag]])
    want_b = complex(k[0]) + np.sum(c[diag] * np.where(par == 1, -1.0, 1.0))
    psi = eng.basis_state(n, b)
    got_b = eng.expectation(plan, psi)
    assert abs(got_b - want_b) <= RTOL64 * abs(want_b), (got_b, want_b)


def test_config4_random_state_energy_is_real_and_equals_psi_dot_hpsi_at_32_qubits(eng):
    """SURVEY.md 8d (iii) at full size: <psi|H|psi> of a random 32-qubit state is real and equals <psi|(H psi)> from the apply path
    (a different kernel mode: full visits, write-back through the tile).  Two 68.7 GB vectors."""
    import torch

    from myqlm_fermion_b200 import workloads

    n = 32
    if _free_gb() < 150:
        pytest.skip("needs two 32-qubit complex128 vectors (137 GB)")
    hpq, hpqrs = workloads.synthetic_h_integrals(n // 2, seed=1234)
    terms = eng.terms_from_integrals(hpq, hpqrs, constant=0.25)
    del hpqrs
    plan = eng.plan(terms)
    psi = eng.random_state(n, seed=0)
    e = eng.expectation(plan, psi)
    assert abs(e.imag) <= 1e-12 * abs(e), e
    phi = eng.apply(plan, psi)
    e2 = eng.dot(psi, phi)
    assert abs(e2 - e) <= RTOL64 * abs(e), (e, e2)
    del psi, phi
    torch.cuda.empty_cache()


def test_config3_hubbard_chain_24_qubits_lanczos(eng):
    import torch

    from myqlm_fermion_b200 import make_hubbard_model, workloads

    ns = 12
    hs = make_hubbard_model(workloads.hubbard_chain_t(ns), 4.0, 2.0).to_spin("jordan-wigner")
    terms = hs.packed(eng)
    assert terms.n_terms == 56  # T = 57 with the identity
    # the reference's own route, precomputed on the CPU by tests/golden/make_golden_lanczos.py: eigsh(get_matrix(sparse=True), which="SA")
    # (tests/integration/test_integration_models.py:102-103) on the oracle's 24-qubit matrix
    golden = json.loads((Path(__file__).parent / "golden" / "hubbard_chain_24q_eigsh.json").read_text())
    assert golden["n_qubits"] == 24 and golden["pauli_terms"] == 56
    e0, psi0, info = eng.ground_state(terms, max_iter=600, tol=1e-11)
    assert abs(e0 - golden["e0"]) <= RTOL64 * abs(golden["e0"]), (e0, golden["e0"], info)
    phi = eng.apply(terms, psi0)
    e = eng.expectation(terms, psi0)
    assert abs(e.imag) <= 1e-12 * abs(e) and abs(e.real - e0) <= RTOL64 * abs(e0)
    assert abs(eng.dot(psi0, phi) - e) <= RTOL64 * abs(e)
    eng.axpy(-e0, psi0, phi)
    assert abs(eng.dot(phi, phi)) ** 0.5 <= 2e-5, info  # eigen-residual
    assert abs(eng.dot(psi0, psi0) - 1.0) < 1e-10
    e32, _, _ = eng.ground_state(terms, max_iter=600, tol=1e-6, dtype=torch.complex64)
    assert abs(e32 - golden["e0"]) <= RTOL32 * abs(golden["e0"])


def test_config5_pool_gradients_at_26_qubits(eng):
    from myqlm_fermion_b200 import workloads

    n, n_e = 26, 12
    hpq, hpqrs = workloads.synthetic_h_integrals(n // 2, seed=1234)
    h = eng.terms_from_integrals(hpq, hpqrs)
    del hpqrs
    pool_f = workloads.uccsd_pool_fermion_terms(n_e, n)
    assert len(pool_f) == len(workloads.uccsd_excitations(n_e, n)) > 1000  # singles + doubles of 12 electrons in 26 spin-orbitals
    ops = [eng.terms_from_fermion(n, op) for op in pool_f]
    pool = eng.concat(ops)
    psi = eng.random_state(n, seed=2)
    g = eng.commutator_gradients(eng.plan(h), eng.plan(pool), psi)
    assert g.shape == (len(pool_f),) and np.all(np.isfinite(g))
    phi = eng.apply(h, psi)
    scale = np.max(np.abs(g))
    for kk in (0, 1, len(pool_f) // 3, len(pool_f) // 2, len(pool_f) - 1):
        single = eng.expectation(eng.plan(ops[kk]), psi, bra=phi)  # <H psi|tau_k|psi> with a single-operator plan
        assert abs(g[kk] - 2.0 * single.imag) <= RTOL64 * scale, kk
    assert scale > 1e-3


def test_host_upload_pipeline_matches_the_device_path(eng):
    """qfe_expectation_host at a size where the upload is pipelined (state >= 1 GiB: 16 pieces on a copy stream, the first sweeps run
    piece by piece behind it): same energy as the device-resident call, and the piecewise launches really happened."""
    from myqlm_fermion_b200 import workloads

    n = 26
    hpq, hpqrs = workloads.synthetic_h_integrals(n // 2, seed=1234)
    plan = eng.plan(eng.terms_from_integrals(hpq, hpqrs, constant=-0.5))
    del hpqrs
    psi = eng.random_state(n, seed=7)
    l0 = eng.launch_count()
    e_dev = eng.expectation(plan, psi)
    l1 = eng.launch_count()
    host = psi.cpu().numpy()
    e_host = eng.expectation_host(plan, host)
    l2 = eng.launch_count()
    assert abs(e_host - e_dev) <= 1e-12 * abs(e_dev), (e_host, e_dev)
    assert (l2 - l1) >= (l1 - l0) + 15, "no sweep ran piece by piece"
    host[: 1 << 20] = 0.0  # a second call sees the new host data (no stale landing buffer, no race with the previous call)
    psi[: 1 << 20] = 0.0
    assert abs(eng.expectation_host(plan, host) - eng.expectation(plan, psi)) <= 1e-12 * abs(e_dev)
