"""Synthetic inputs of the BASELINE.json configs (vectorised; no Python n^4 loops).

``convert_to_h_integrals`` follows /root/reference/qat/fermion/chemistry/ucc.py:128-229 (spatial ->
spin-orbital, index 2*orbital + spin; mixed-spin entries always, same-spin entries only when
p != q and r != s); ``synthetic_integrals`` is the generator SURVEY.md 8(d) defines for configs 2/4/5.
"""

from __future__ import annotations

import numpy as np


def synthetic_integrals(m: int, seed: int = 1234):
    """Random 8-fold-symmetric spatial integrals: I1 symmetric, g symmetrised as a chemist (ij|kl)
    tensor, I2 = g.transpose(0, 2, 3, 1) (the reference's order, chemistry/pyscf_tools.py:32)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((m, m))
    one = 0.5 * (a + a.T)
    g = rng.standard_normal((m, m, m, m))
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    return one, np.ascontiguousarray(g.transpose(0, 2, 3, 1))


def convert_to_h_integrals(one_body: np.ndarray, two_body: np.ndarray):
    """(I_uv, I_uvwx) -> (h_pq, h_pqrs) on 2m spin-orbitals (ucc.py:128-229)."""
    m = one_body.shape[0]
    n = 2 * m
    hpq = np.zeros((n, n), dtype=np.complex128)
    hpqrs = np.zeros((n, n, n, n), dtype=np.complex128)
    for s in (0, 1):
        hpq[s::2, s::2] = one_body
        # mixed spin: h[2p+s, 2q+1-s, 2r+1-s, 2s'+s] = I[p, q, r, s']
        hpqrs[s::2, 1 - s :: 2, 1 - s :: 2, s::2] = two_body
    p, q, r, t = np.ogrid[:m, :m, :m, :m]
    same = np.where((p != q) & (r != t), two_body, 0.0)
    keep = np.broadcast_to((p != q) & (r != t), two_body.shape)
    for s in (0, 1):
        view = hpqrs[s::2, s::2, s::2, s::2]
        view[keep] = same[keep]
    return hpq, hpqrs


def synthetic_h_integrals(m: int, seed: int = 1234):
    return convert_to_h_integrals(*synthetic_integrals(m, seed))


def synthetic_h_integrals_n(n: int, seed: int = 1234):
    """The generator on n spin-orbitals; an odd n keeps the first n of the 2*ceil(n/2) spin-orbitals."""
    hpq, hpqrs = synthetic_h_integrals((n + 1) // 2, seed)
    if n % 2:
        hpq, hpqrs = np.ascontiguousarray(hpq[:n, :n]), np.ascontiguousarray(hpqrs[:n, :n, :n, :n])
    return hpq, hpqrs


def hubbard_chain_t(ns: int, t: float = -1.0) -> np.ndarray:
    t_mat = np.zeros((ns, ns))
    for i in range(ns - 1):
        t_mat[i, i + 1] = t_mat[i + 1, i] = t
    return t_mat


def uccsd_excitations(n_electrons: int, nqbits: int):
    """(a, i) and (b, a, j, i) spin-orbital index tuples of the UCCSD pool in the reference's order
    (/root/reference/qat/fermion/chemistry/ucc.py:554-580, 691-713): occupied = the first n_electrons
    spin-orbitals, spin-preserving singles, spin-matching doubles, each list sorted descending, singles first."""
    if nqbits % 2 or n_electrons % 2:
        raise ValueError("Only even values of nqbits / n_electrons are allowed.")
    occ = list(range(n_electrons))
    unocc = list(range(n_electrons, nqbits))
    singles = []
    for a in unocc[::2]:
        for i in occ[::2]:
            singles.append((a, i))
            singles.append((a + 1, i + 1))
    doubles = []
    for ia, a in enumerate(unocc):
        for b in unocc[ia + 1 :]:
            for ii, i in enumerate(occ):
                for j in occ[ii + 1 :]:
                    if (a % 2 == i % 2 and b % 2 == j % 2) or (a % 2 == j % 2 and b % 2 == i % 2):
                        doubles.append((b, a, j, i))
    return sorted(singles)[::-1] + sorted(doubles)[::-1]


def uccsd_pool_fermion_terms(n_electrons: int, nqbits: int):
    """Hermitian generators i(T - T^dagger) as fermionic term lists (ucc.py:232-288): one list per excitation."""
    pool = []
    for ex in uccsd_excitations(n_electrons, nqbits):
        if len(ex) == 2:
            op, idx, conj = "Cc", list(ex), [ex[1], ex[0]]
        else:
            op, idx, conj = "CCcc", list(ex), [ex[2], ex[3], ex[0], ex[1]]
        pool.append([(1j, op, idx), (-1j, op, conj)])
    return pool
