"""Quick timing probe of kernels 1-2 on one GPU (not the bench; used while developing).
usage: python tools/probe_perf.py M [tile_bits] [dtype] -> n = 2M qubits synthetic molecular JW Hamiltonian."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200.workloads import synthetic_h_integrals

m = int(sys.argv[1]) if len(sys.argv) > 1 else 12
tile_bits = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dtype = torch.complex64 if len(sys.argv) > 3 and sys.argv[3] == "c64" else torch.complex128
n = 2 * m
eng = qfe.get_engine(0)
t0 = time.time()
hpq, hpqrs = synthetic_h_integrals(m, seed=1234)
t1 = time.time()
terms = eng.terms_from_integrals(hpq, hpqrs)
eng.sync()
t2 = time.time()
print(f"n={n}: integrals {t1-t0:.2f}s, term expansion {t2-t1:.3f}s -> T={terms.n_terms} xmasks={terms.n_xmasks}", flush=True)
plan = eng.plan(terms, tile_bits=tile_bits)
t3 = time.time()
print(f"plan: {t3-t2:.2f}s sweeps={plan.n_sweeps} groups={plan.n_groups} tile_bits={plan.tile_bits}", flush=True)
psi = eng.random_state(n, dtype=dtype, seed=0)
torch.cuda.synchronize()
for rep in range(2):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t4 = time.time()
    e = eng.expectation(plan, psi)
    t5 = time.time()
    amps = float(1 << n)
    print(f"expectation: {t5-t4:.3f}s  E={e.real:.10f} im={e.imag:.2e}  terms*amps/s={terms.n_terms*amps/(t5-t4):.3e}  "
          f"sweep-GB/s={plan.n_sweeps*amps*psi.element_size()/(t5-t4)/1e9:.1f}", flush=True)
if n <= 30:
    t6 = time.time()
    phi = eng.apply(plan, psi)
    t7 = time.time()
    print(f"apply: {t7-t6:.3f}s  <psi|phi>={eng.dot(psi, phi).real:.10f}", flush=True)
