"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  CPU baseline of the hot path for bench.py's
``cpu_baseline`` leg and ``--impl reference`` arm: the reference's algorithm for <psi|H|psi> --
``SpinHamiltonian.get_matrix(sparse=True)`` as a chain of kron(I, sigma, I) products per term
(/root/reference/qat/fermion/hamiltonians.py:163-250, restated in oracle/spin.py) followed by
``vdot(psi, H psi)`` -- on a bounded sample of the bench workload (same integral generator, fewer
orbitals).  The reference is single-threaded; ``workers`` > 1 splits the term list across processes
and adds the partial matrices, which is the most parallelism the algorithm admits.
"""

from __future__ import annotations

import os
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

from . import models, spin, vectorised


def _partial_matrix(args):
    n, triples = args
    return spin.spin_get_matrix(n, triples, 0.0, sparse=True)


def sample_terms(n: int, seed: int = 1234, encoding: str = "jordan-wigner"):
    """Term triples [(coeff, op, qbits)] + constant of the bench generator on n = 2m spin-orbitals."""
    m = (n + 1) // 2
    hpq, hpqrs = models.convert_to_h_integrals(*models.synthetic_integrals(m, seed))
    hpq, hpqrs = hpq[:n, :n], hpqrs[:n, :n, :n, :n]
    x, z, c, k = vectorised.expand_electronic(hpq, hpqrs, 0.0, encoding)
    triples = []
    for xi, zi, ci in zip(x.tolist(), z.tolist(), c.tolist()):
        op, qb, _ = spin.packed_to_string(n, xi, zi)
        triples.append((ci, op, qb))
    return triples, k


def reference_matrix(n: int, triples, constant, workers: int = 1, pool: ProcessPoolExecutor | None = None):
    """get_matrix(sparse=True) of the term list; with a pool the term list is split across processes and the partial matrices added."""
    if workers <= 1 or pool is None:
        return spin.spin_get_matrix(n, triples, constant, sparse=True)
    chunks = [triples[i::workers] for i in range(workers)]
    parts = list(pool.map(_partial_matrix, [(n, ch) for ch in chunks if ch]))
    mat = parts[0]
    for p in parts[1:]:
        mat = mat + p
    import scipy.sparse as sp

    return (mat + constant * sp.identity(2**n, dtype=complex, format="csr")).tocsr()


def reference_energy(n: int, triples, constant, psi: np.ndarray, workers: int = 1, pool: ProcessPoolExecutor | None = None):
    """(energy, t_build, t_eval): get_matrix(sparse=True) then vdot(psi, H psi)."""
    t0 = time.perf_counter()
    mat = reference_matrix(n, triples, constant, workers, pool)
    t1 = time.perf_counter()
    e = np.vdot(psi, mat.dot(psi))
    t2 = time.perf_counter()
    return complex(e), t1 - t0, t2 - t1


def usable_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
