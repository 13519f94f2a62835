"""
Generates tests/golden/adapt_gradients_full_pools.json: the ADAPT-VQE gradient sweep of BASELINE.json config 5 at the two sizes SURVEY.md
8d names for full-oracle parity -- n = 12 (4 electrons, K = 92 pool operators) and n = 14 (10 electrons, K = 140) -- computed the way
the reference computes it (/root/reference/qat/plugins/adapt_vqe.py:50-62, 181-185): every commutator observable [H, tau_k] is expanded
term by term with the Pauli algebra (H | tau = H*tau - tau*H, qat/fermion/hamiltonians.py:134-135), then g_k = -i <psi|[H, tau_k]|psi>.

Inputs are the same generators the GPU test uses (oracle.models: synthetic integrals seed 1234, UCCSD pool, state default_rng(9)).
The expanded observables hold ~10^4 terms each; the full sweep takes minutes on the CPU, which is why the numbers are a fixture.
Run ONCE in the build container: python tests/golden/make_golden_adapt.py
"""

from __future__ import annotations

import json
import sys
import time
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import models, spin, transforms  # noqa: E402

OUT = Path(__file__).resolve().parent / "adapt_gradients_full_pools.json"
_CTX = {}


def _setup(n, ne):
    one, two = models.synthetic_integrals(n // 2, seed=1234)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    h = transforms.transform(n, models.electronic_terms(hpq, hpqrs), 0.0, "jordan-wigner")
    pool = [transforms.transform(n, op, 0.0, "jordan-wigner") for op in models.cluster_ops(ne, n)]
    rng = np.random.default_rng(9)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    return h, pool, psi


def _one(args):
    n, ne, k = args
    if (n, ne) not in _CTX:
        _CTX[(n, ne)] = _setup(n, ne)
    h, pool, psi = _CTX[(n, ne)]
    comm = h.commutator(pool[k])  # the expanded observable [H, tau_k]
    xs, zs, cs, const = comm.canonical(eps=0.0)
    val = -1j * spin.expectation_packed(n, xs, zs, cs, const, psi)
    return k, val.real, val.imag, len(xs)


def main():
    out = {"how": "g_k = -i <psi|[H, tau_k]|psi> with every commutator expanded as a Pauli sum (oracle.spin.PauliSum.commutator), "
                  "JW synthetic molecular H (seed 1234), UCCSD pool models.cluster_ops(ne, n), psi = normalised N(0,1)+iN(0,1) from default_rng(9)",
           "reference": "qat/plugins/adapt_vqe.py:50-62,181-185", "cases": []}
    for n, ne in ((12, 4), (14, 10)):
        t0 = time.time()
        K = len(models.cluster_ops(ne, n))
        with ProcessPoolExecutor(8) as ex:
            rows = sorted(ex.map(_one, [(n, ne, k) for k in range(K)], chunksize=4))
        g = [r[1] for r in rows]
        assert max(abs(r[2]) for r in rows) < 1e-10  # the gradients are real
        out["cases"].append({"n": n, "n_electrons": ne, "K": K, "gradients": g, "max_terms_in_a_commutator": max(r[3] for r in rows),
                             "norm": float(np.linalg.norm(g)), "argmax": int(np.argmax(np.abs(g)))})
        print(f"n={n} ne={ne} K={K}: {time.time() - t0:.0f} s, |g|={np.linalg.norm(g):.6f}", flush=True)
    OUT.write_text(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
