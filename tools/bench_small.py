"""Latency of the engine at the register sizes the reference's own tests and notebooks use (8-22 qubits): microseconds per call of
<psi|H|psi>, H|psi> and one Trotter slice, wall clock around synchronous calls (what a VQE loop sees).
usage: python tools/bench_small.py > gpurun_out/small.jsonl"""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200 import workloads

eng = qfe.get_engine(0)


def per_call(fn, reps):
    fn()
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6, (eng.launch_count() - l0) // reps


for n in (8, 10, 12, 14, 16, 18, 20, 22):
    hpq, hpqrs = workloads.synthetic_h_integrals_n(n, seed=1234)
    h = eng.terms_from_integrals(hpq, hpqrs)
    plan = eng.plan(h)
    psi = eng.random_state(n, seed=0)
    phi = torch.empty_like(psi)
    reps = 200 if n <= 16 else 50
    us_e, l_e = per_call(lambda: eng.expectation(plan, psi), reps)
    us_a, l_a = per_call(lambda: eng.apply(plan, psi, out=phi), reps)
    x, z, c, _, _ = h.arrays()
    sel = slice(0, min(len(x), 400))
    th = 0.01 * c[sel].real
    us_r, l_r = per_call(lambda: eng.pauli_rotations(psi, x[sel], z[sel], th), max(reps // 10, 5))
    print(json.dumps({"n_qubits": n, "pauli_terms": h.n_terms, "sweeps": plan.n_sweeps, "expectation_us": us_e, "expectation_launches": l_e,
                      "apply_us": us_a, "apply_launches": l_a, "rotations": int(len(th)), "rotations_us": us_r, "rotation_passes": l_r}), flush=True)
