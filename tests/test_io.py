"""On-disk formats: molecule files with the keys of the reference's tests/resources/*_data.npz, packed term sets."""

import numpy as np
import pytest

from oracle import models, spin, transforms


def test_molecule_file_round_trip_and_hamiltonian(golden_dir, tmp_path):
    import myqlm_fermion_b200 as qfe

    d = qfe.load_molecule_npz(golden_dir / "h2_integrals.npz")
    h, n_e = qfe.molecule_hamiltonian(d)
    assert h.nbqbits == 4 and n_e == 2
    hpq, hpqrs = models.convert_to_h_integrals(d["one_body_integrals"], d["two_body_integrals"])
    assert np.array_equal(h.hpq, hpq) and np.array_equal(h.hpqrs, hpqrs)
    assert h.constant_coeff == float(d["nuclear_repulsion"])
    qfe.save_molecule_npz(tmp_path / "m.npz", d["one_body_integrals"], d["two_body_integrals"], d["nuclear_repulsion"], d["n_electrons"], rdm1=d["rdm1"])
    e = qfe.load_molecule_npz(tmp_path / "m.npz")
    for k in ("one_body_integrals", "two_body_integrals", "nuclear_repulsion", "n_electrons", "rdm1"):
        assert np.array_equal(e[k], d[k]), k
    np.savez(tmp_path / "bad.npz", one_body_integrals=np.zeros((2, 2)))
    with pytest.raises(KeyError):
        qfe.load_molecule_npz(tmp_path / "bad.npz")


@pytest.mark.gpu
def test_molecule_file_to_ground_energy_and_term_files(golden_dir, tmp_path):
    """H2 fixture -> JW terms on the GPU -> Lanczos ground energy = the FCI energy stored with the fixture; term sets (single operator
    and a batch) survive a save / load cycle bit for bit."""
    import myqlm_fermion_b200 as qfe

    eng = qfe.get_engine(0)
    d = qfe.load_molecule_npz(golden_dir / "h2_integrals.npz")
    h, _ = qfe.molecule_hamiltonian(d)
    hs = h.to_spin("jordan-wigner")
    e0, _, _ = eng.ground_state(hs.packed(eng), max_iter=100, tol=1e-12)
    assert abs(e0 - float(d["fci"])) < 1e-8
    qfe.save_terms_npz(tmp_path / "h2_terms.npz", hs)
    back = qfe.load_terms_npz(tmp_path / "h2_terms.npz", eng)
    x, z, c, k = hs.packed_arrays()
    bx, bz, bc, _, bk = back.arrays()
    assert np.array_equal(bx, x) and np.array_equal(bz, z) and np.array_equal(bc, c) and bk[0] == k
    # a batch: the UCCSD pool of 4 electrons in 8 spin-orbitals
    pool = [transforms.transform(8, op, 0.0, "jordan-wigner").canonical() for op in models.cluster_ops(4, 8)]
    batch = eng.concat([eng.terms_from_pauli(8, p[0], p[1], p[2], constant=p[3]) for p in pool])
    qfe.save_terms_npz(tmp_path / "pool.npz", batch)
    again = qfe.load_terms_npz(tmp_path / "pool.npz", eng)
    for a, b in zip(batch.arrays(), again.arrays()):
        assert np.array_equal(a, b)



def test_change_encoding_matches_the_reference_loop():
    """transforms.py:380-407 restated as the double loop, against the facade's vectorised permutation."""
    import itertools

    import myqlm_fermion_b200 as qfe

    n = 4
    rng = np.random.default_rng(0)
    mat = rng.standard_normal((1 << n, 1 << n)) + 1j * rng.standard_normal((1 << n, 1 << n))
    for code in (qfe.get_jw_code(n), qfe.get_parity_code(n), qfe.get_bk_code(n)):
        table = [qfe.recode_integer(i, code) for i in range(1 << n)]
        want = np.zeros_like(mat)
        for i, j in itertools.product(range(1 << n), repeat=2):
            want[table[i], table[j]] = mat[i, j]
        assert np.array_equal(qfe.change_encoding(mat, code), want)
    for enc, fn in (("jordan-wigner", qfe.get_jw_code), ("parity", qfe.get_parity_code), ("bravyi-kitaev", qfe.get_bk_code)):
        assert np.array_equal(fn(5), transforms.code_for(5, enc)), enc  # the oracle's code matrices (transforms.py:260-346)
