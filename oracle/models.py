"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  Input generators of the oracle.

Restates, in plain numpy / Python loops:
  * /root/reference/qat/fermion/hamiltonians.py:625-648  integrals -> fermionic terms
    (|h| > 1e-12 kept; two-body terms carry 1/2 h_pqrs and the operator a+_p a+_q a_r a_s),
  * /root/reference/qat/fermion/hamiltonians.py:874-916  make_hubbard_model,
  * /root/reference/qat/fermion/chemistry/ucc.py:128-229 spatial -> spin-orbital integrals,
  * /root/reference/qat/fermion/chemistry/ucc.py:232-288, 554-580, 583-713, 779-852  UCCSD pool,
plus the synthetic 8-fold-symmetric integral generator BASELINE.md section 4 / SURVEY.md 8(d) define.
"""

from __future__ import annotations

import itertools

import numpy as np

TOL = 1e-12  # hamiltonians.py:541


def electronic_terms(hpq: np.ndarray, hpqrs: np.ndarray | None):
    """[(coeff, op, qbits)] of H = sum h_pq a+_p a_q + 1/2 sum h_pqrs a+_p a+_q a_r a_s."""
    n = hpq.shape[0]
    terms = []
    for p, q in itertools.product(range(n), repeat=2):
        if abs(hpq[p, q]) > TOL:
            terms.append((hpq[p, q], "Cc", [p, q]))
    if hpqrs is not None:
        nz = np.argwhere(np.abs(hpqrs) > TOL)  # C order == itertools.product order
        for p, q, r, s in nz:
            terms.append((0.5 * hpqrs[p, q, r, s], "CCcc", [int(p), int(q), int(r), int(s)]))
    return terms


def hubbard_integrals(t_mat: np.ndarray, U: float, mu: float):
    """(hpq, hpqrs) of the Hubbard model; spin-orbital index = 2*site + sigma."""
    ns = t_mat.shape[0]
    n = 2 * ns
    hpq = np.zeros((n, n))
    for i, j in itertools.product(range(ns), repeat=2):
        for sig in (0, 1):
            hpq[2 * i + sig, 2 * j + sig] = t_mat[i, j]
    for i in range(ns):
        for sig in (0, 1):
            hpq[2 * i + sig, 2 * i + sig] += -mu
    hpqrs = np.zeros((n, n, n, n))
    for i in range(ns):
        for sig in (0, 1):
            hpqrs[2 * i + sig, 2 * i + 1 - sig, 2 * i + sig, 2 * i + 1 - sig] = -U
    return hpq, hpqrs


def convert_to_h_integrals(one_body: np.ndarray, two_body: np.ndarray):
    """Spatial integrals I_uv, I_uvwx -> spin-orbital h_pq, h_pqrs (index 2*orbital + spin).
    Mixed-spin entries always; same-spin entries only when p != q and r != s (ucc.py:176-185)."""
    m = one_body.shape[0]
    n = 2 * m
    hpq = np.zeros((n, n), dtype=np.complex128)
    hpqrs = np.zeros((n, n, n, n), dtype=np.complex128)
    for p, q in itertools.product(range(m), repeat=2):
        for s in (0, 1):
            hpq[2 * p + s, 2 * q + s] = one_body[p, q]
    for p, q, r, s in itertools.product(range(m), repeat=4):
        v = two_body[p, q, r, s]
        for sp in (0, 1):
            hpqrs[2 * p + sp, 2 * q + 1 - sp, 2 * r + 1 - sp, 2 * s + sp] = v
        if p != q and r != s:
            for sp in (0, 1):
                hpqrs[2 * p + sp, 2 * q + sp, 2 * r + sp, 2 * s + sp] = v
    return hpq, hpqrs


def synthetic_integrals(m: int, seed: int = 1234):
    """Random 8-fold-symmetric spatial integrals:
    I1 symmetric, g symmetrised as a chemist (ij|kl) tensor, I2 = g.transpose(0,2,3,1)
    (the reference's integral order, /root/reference/qat/fermion/chemistry/pyscf_tools.py:32)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((m, m))
    one = 0.5 * (a + a.T)
    g = rng.standard_normal((m, m, m, m))
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    two = np.ascontiguousarray(g.transpose(0, 2, 3, 1))
    return one, two


# ---------------------------------------------------------------------------------------------
# UCCSD operator pool
# ---------------------------------------------------------------------------------------------
def hf_ket(n_electrons: int, n: int) -> int:
    """Hartree-Fock basis state: the first n_electrons spin-orbitals occupied (ucc.py:779-798)."""
    v = 0
    for i in range(n_electrons):
        v |= 1 << (n - 1 - i)
    return v


def excitation_tuples(n_electrons: int, n: int):
    """(a, i) and (b, a, j, i) index tuples in the reference's order (ucc.py:554-580, 691-713):
    occupied = first n_electrons spin-orbitals, spin-preserving singles, spin-matching doubles,
    each list ``sorted(...)[::-1]``, singles first."""
    if n % 2 or n_electrons % 2:
        raise ValueError("Only even values of nqbits / n_electrons are allowed.")
    orbitals = list(range(n))
    occ = orbitals[:n_electrons]
    unocc = orbitals[n_electrons:]
    singles = []
    for a, i in itertools.product(unocc[::2], occ[::2]):
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


def cluster_ops(n_electrons: int, n: int):
    """Pool of Hermitian generators i(T - T^dagger) as fermionic term lists (ucc.py:232-288):
    one entry per excitation, [(1j, op, idx), (-1j, op, idx_conj)]."""
    pool = []
    for ex in excitation_tuples(n_electrons, n):
        if len(ex) == 2:
            op, idx, conj = "Cc", list(ex), [ex[1], ex[0]]
        else:
            op, idx, conj = "CCcc", list(ex), [ex[2], ex[3], ex[0], ex[1]]
        pool.append([(1j, op, idx), (-1j, op, conj)])
    return pool
