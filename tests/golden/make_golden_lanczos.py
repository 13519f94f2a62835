"""
Generates tests/golden/hubbard_chain_24q_eigsh.json: the lowest eigenvalue of BASELINE.json config 3 (12-site open Hubbard chain,
t = -1, U = 4, mu = 2, Jordan-Wigner, 24 qubits) computed the way the reference's own test computes it
(/root/reference/tests/integration/test_integration_models.py:102-103): ``scipy.sparse.linalg.eigsh(H.get_matrix(sparse=True), which="SA")``.

The matrix comes from the oracle (make_hubbard_model -> transform_to_jw_basis -> sparse matrix of the Pauli sum, the same assembly that
tests/test_oracle_goldens.py checks against the reference's kron-chain get_matrix at small sizes).  Run ONCE in the build container
(``python tests/golden/make_golden_lanczos.py``, ~10 GB of RAM, a few minutes); the GPU test only reads the JSON.
"""

from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np
import scipy.sparse.linalg as spla

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import models, spin, transforms  # noqa: E402

OUT = Path(__file__).resolve().parent / "hubbard_chain_24q_eigsh.json"


def chain_t(ns: int) -> np.ndarray:
    t = np.zeros((ns, ns))
    for i in range(ns - 1):
        t[i, i + 1] = t[i + 1, i] = -1.0
    return t


def main(ns: int = 12) -> None:
    n = 2 * ns
    t0 = time.time()
    hpq, hpqrs = models.hubbard_integrals(chain_t(ns), 4.0, 2.0)
    ps = transforms.transform(n, models.electronic_terms(hpq, hpqrs), 0.0, "jordan-wigner")
    xs, zs, cs, const = ps.canonical()
    print(f"n={n}: {len(xs)} Pauli terms, constant {const}", flush=True)
    mat = spin.csr_from_packed(n, xs, zs, cs, const)
    assert abs(mat.imag).max() == 0.0  # JW Hubbard chain: real symmetric
    mat = mat.real.tocsr()
    mat.eliminate_zeros()
    print(f"matrix: nnz={mat.nnz} ({time.time() - t0:.0f} s)", flush=True)
    v0 = np.random.default_rng(0).standard_normal(1 << n)
    vals, vecs = spla.eigsh(mat, k=1, which="SA", tol=1e-13, v0=v0, ncv=40, maxiter=5000)
    e0 = float(vals[0])
    resid = float(np.linalg.norm(mat @ vecs[:, 0] - e0 * vecs[:, 0]))
    print(f"E0 = {e0!r}  residual {resid:.2e}  ({time.time() - t0:.0f} s)", flush=True)
    OUT.write_text(json.dumps({
        "config": "BASELINE.json configs[2]: 1D Hubbard 12 sites, t=-1 (open chain), U=4, mu=2, Jordan-Wigner, 24 qubits",
        "reference_call": "scipy.sparse.linalg.eigsh(H.get_matrix(sparse=True), k=1, which='SA', tol=1e-13)  "
                          "(/root/reference/tests/integration/test_integration_models.py:102-103)",
        "n_qubits": n, "pauli_terms": int(len(xs)), "constant": [const.real, const.imag], "nnz": int(mat.nnz),
        "e0": e0, "eigsh_residual_norm": resid,
    }, indent=1) + "\n")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 12)
