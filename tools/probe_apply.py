"""Quick timing probe (one GPU, development): H|psi> and <psi|H|psi> of the synthetic molecular JW Hamiltonian at n = 2M qubits and H|psi> of the
n-qubit Hubbard chain (the Lanczos matvec), CUDA events on the launching stream.  usage: python tools/probe_apply.py M [reps] [hubbard]"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200 import make_hubbard_model
from myqlm_fermion_b200.workloads import synthetic_h_integrals

m = int(sys.argv[1]) if len(sys.argv) > 1 else 14
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 2 * m
eng = qfe.get_engine(0)
eng.use_torch_stream()


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


hubbard_only = "hubbard" in sys.argv[3:]
psi = eng.random_state(n, dtype=torch.complex128, seed=0)
phi = torch.empty_like(psi)
if not hubbard_only:
    hpq, hpqrs = synthetic_h_integrals(m, seed=1234)
    terms = eng.terms_from_integrals(hpq, hpqrs)
    plan = eng.plan(terms)
    ms_e, e = timed(lambda: eng.expectation(plan, psi))
    ms_a, _ = timed(lambda: eng.apply(plan, psi, out=phi))
    d = eng.dot(psi, phi)
import numpy as np

t = np.zeros((m, m))
for i in range(m - 1):
    t[i, i + 1] = t[i + 1, i] = -1.0
hub = make_hubbard_model(t, 4.0, 2.0).to_spin("jordan-wigner")
hterms = hub.packed(eng)
hplan = eng.plan(hterms)
ms_h, _ = timed(lambda: eng.apply(hplan, psi, out=phi))
dh = eng.dot(psi, phi)
out = dict(n=n, env={k: v for k, v in os.environ.items() if k.startswith("QFE_")}, hubbard_apply_ms=ms_h, hubbard_sweeps=hplan.n_sweeps, hubbard_psi_H_psi=dh.real)
if not hubbard_only:
    out.update(expect_ms=ms_e, apply_ms=ms_a, energy=e.real, psi_H_psi_from_apply=d.real, apply_terms_amps_per_s=terms.n_terms * float(1 << n) / (ms_a / 1e3))
print(json.dumps(out))
