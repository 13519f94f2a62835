"""Multi-GPU parity at sizes where no single GPU holds the register (BASELINE.json configs 4/5 at 33-34 qubits): size-independent checks
of the sharded <psi|H|psi> and H|psi> against closed forms computed on the CPU:
  * random product state: <P> = prod_q <sigma_q>, O(T n) on the host;
  * Hartree-Fock basis state: constant + the diagonal terms;
  * random state: <psi|H|psi> is real and equals <psi|(H psi)> from the sharded apply.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29557 tools/check_sharded_large.py N_QUBITS"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200 import workloads
from myqlm_fermion_b200.sharded import ShardedEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = qfe.get_engine(local)
se = ShardedEngine(eng, n)
hpq, hpqrs = workloads.synthetic_h_integrals_n(n, seed=1234)
terms = eng.terms_from_integrals(hpq, hpqrs, constant=0.25)
del hpqrs
x, z, c, _, k = terms.arrays()
plan = se.plan(terms)
out = {"n_qubits": n, "n_gpus": world, "pauli_terms": terms.n_terms, "classes": plan.classes, "sweeps_per_rank": plan.n_sweeps, "comm": se.comm_info()}

# (i) random product state
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
psi = se.product_state(ab)
t0 = time.perf_counter()
got = se.expectation(plan, psi)
out["product_state"] = {"got": [got.real, got.imag], "want": [want.real, want.imag], "rel_dev": abs(got - want) / abs(want), "seconds": time.perf_counter() - t0}
assert abs(got - want) <= 1e-10 * abs(want), (got, want)
x_hi, ms = np.zeros(64, np.int32), np.zeros(64)
import ctypes as C

cnt = C.c_int32()
eng.lib.qfe_comm_class_times(eng.ctx, x_hi.ctypes.data_as(C.POINTER(C.c_int32)), ms.ctypes.data_as(C.POINTER(C.c_double)), 64, C.byref(cnt))
out["class_ms_rank0"] = {int(a): float(b) for a, b in zip(x_hi[: cnt.value], ms[: cnt.value])}

# (ii) Hartree-Fock determinant: the first n/2 qubits set
from oracle import models  # test infrastructure: the closed form only

b = models.hf_ket(n // 2, n)
diag = x == 0
par = np.array([bin(int(v) & b).count("1") & 1 for v in z[diag]])
want_b = complex(k[0]) + np.sum(c[diag] * np.where(par == 1, -1.0, 1.0))
del psi
psi = eng.basis_state(n, b, n_local=se.n_local, shard=rank)
got_b = se.expectation(plan, psi)
out["basis_state"] = {"got": [got_b.real, got_b.imag], "want": [want_b.real, want_b.imag], "rel_dev": abs(got_b - want_b) / abs(want_b)}
assert abs(got_b - want_b) <= 1e-10 * abs(want_b), (got_b, want_b)

# (iii) random state: E real, E = <psi|H psi>  ("noapply" skips it: an apply costs about two expectation steps)
if "noapply" not in sys.argv[2:]:
    del psi
    psi = se.random_state(seed=0)
    e = se.expectation(plan, psi)
    phi = se.apply(plan, psi)
    e2 = se.dot(psi, phi)
    out["random_state"] = {"E": [e.real, e.imag], "psi_dot_Hpsi": [e2.real, e2.imag], "rel_dev": abs(e - e2) / abs(e)}
    assert abs(e.imag) <= 1e-10 * abs(e) and abs(e - e2) <= 1e-10 * abs(e), (e, e2)
out["exchanges"] = se.exchanges
if rank == 0:
    print(json.dumps(out), flush=True)
    print("sharded large-register parity ok", flush=True)
dist.barrier()
dist.destroy_process_group()
