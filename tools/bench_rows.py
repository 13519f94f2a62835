"""Measurement of the SURVEY.md 8 rows next to the headline one (bench.py measures <psi|H|psi>): every kernel family of libqfe timed on
one GPU with CUDA events on the launching stream, after a warm-up call, against the bound that applies to it.  One JSON line per
row; the HBM rows carry achieved GB/s (algorithmic bytes / time) and the fraction of MEASURED_PEAKS.json's hbm_gbs.
usage: python tools/bench_rows.py [N_QUBITS=30] > gpurun_out/rows.jsonl"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import myqlm_fermion_b200 as qfe  # noqa: E402
from myqlm_fermion_b200 import make_hubbard_model, workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
try:
    HBM = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    HBM_SRC = "MEASURED_PEAKS.json hbm_gbs"
except Exception:
    HBM, HBM_SRC = 7700.0, "nominal (MEASURED_PEAKS.json absent)"

eng = qfe.get_engine(0)
eng.use_torch_stream()
N = 1 << n
ESZ = 16  # complex128


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3, (eng.launch_count() - l0) // reps


def emit(row, seconds, launches, **kw):
    d = {"row": row, "n_qubits": n, "seconds": seconds, "launches": launches}
    if "bytes" in kw:
        gbs = kw["bytes"] / seconds / 1e9
        d["hbm"] = {"achieved_gbs": gbs, "peak_gbs": HBM, "frac": gbs / HBM, "peak_source": HBM_SRC}
    d.update(kw)
    print(json.dumps(d), flush=True)


# ---- kernel (1): integrals -> packed Pauli terms (32 spin-orbitals, the headline Hamiltonian) -------------------------------------
hpq, hpqrs = workloads.synthetic_h_integrals(16, seed=1234)
t0 = time.perf_counter()
terms32 = eng.terms_from_integrals(hpq, hpqrs)
eng.sync()
cold = time.perf_counter() - t0
s, l = timed(lambda: eng.terms_from_integrals(hpq, hpqrs), reps=3, warm=0)
emit("terms: h_pq/h_pqrs (32 spin-orbitals) -> Jordan-Wigner Pauli sum, host arrays in, packed terms resident on the device", s, l,
     pauli_terms=terms32.n_terms, fermionic_integrals=int(np.count_nonzero(hpqrs) + np.count_nonzero(hpq)), first_call_seconds=cold,
     h2d_bytes=int(hpq.nbytes + hpqrs.nbytes), note="includes the host->device copy of the dense integral tensors; launch-latency bound at this size")
del hpqrs, terms32

# ---- kernel (2): H|psi> and <psi|H|psi> on the n-qubit member of the same generator ------------------------------------------------
hpq, hpqrs = workloads.synthetic_h_integrals_n(n, seed=1234)
h = eng.terms_from_integrals(hpq, hpqrs)
del hpqrs
t0 = time.perf_counter()
plan = eng.plan(h)
t_plan = time.perf_counter() - t0
psi = eng.random_state(n, seed=0)
s, l = timed(lambda: eng.expectation(plan, psi), reps=1)
emit("expectation: <psi|H|psi>, molecular Jordan-Wigner H", s, l, pauli_terms=h.n_terms, x_masks=h.n_xmasks, sweeps=plan.n_sweeps,
     terms_amps_per_s=h.n_terms * N / s, plan_seconds=t_plan, bound="shared memory (see bench.py roofline.smem)")
phi = torch.empty_like(psi)
s, l = timed(lambda: eng.apply(plan, psi, out=phi), reps=1)
emit("apply: phi = H|psi>, same H", s, l, pauli_terms=h.n_terms, sweeps=plan.n_sweeps, terms_amps_per_s=h.n_terms * N / s,
     fp64_fma_per_s=2.0 * h.n_terms * N / s, bound="FP64 pipe + shared memory: every (term, amplitude) is one real x complex multiply-add")

# ---- vector kernels of the Lanczos loop (HBM bound) ---------------------------------------------------------------------------------
s, l = timed(lambda: eng.dot(psi, phi), reps=5)
emit("vec_dot: <a|b>", s, l, bytes=2 * N * ESZ)
s, l = timed(lambda: eng.axpy(0.5 - 0.25j, psi, phi), reps=5)
emit("vec_axpy: y += alpha x", s, l, bytes=3 * N * ESZ)
s, l = timed(lambda: eng.scale(phi, 1.0 + 0.0j), reps=5)
emit("vec_scale: x *= alpha", s, l, bytes=2 * N * ESZ)
del phi
torch.cuda.empty_cache()

# ---- Lanczos on the Hubbard chain (config 3 family): a fixed number of iterations ---------------------------------------------------
hub = make_hubbard_model(workloads.hubbard_chain_t(n // 2), 4.0, 2.0).to_spin("jordan-wigner").packed(eng)
hub_plan = eng.plan(hub)
iters = 20
info = {}


def lanczos():
    e0, _, inf = eng.ground_state(hub_plan, v0=psi, max_iter=iters, tol=0.0, want_vector=False)
    info.update(inf, e0=e0)


s, l = timed(lanczos, reps=1)
s_apply, _ = timed(lambda: eng.expectation(hub_plan, psi), reps=1)
emit(f"lanczos: {iters} iterations on the {n // 2}-site Hubbard chain ({hub.n_terms} Pauli terms)", s, l, iterations=info.get("iterations"),
     seconds_per_iteration=s / max(info.get("iterations", iters), 1), sweeps_per_apply=hub_plan.n_sweeps, expectation_seconds=s_apply,
     bytes=int(info.get("iterations", iters)) * (hub_plan.n_sweeps * 3 - 1 + 8) * N * ESZ,
     note="bytes = iterations x (apply: every sweep reads the vector and accumulates into w, the first only writes; + 8 vector passes: dot, fused update, scale)")

# ---- state preparation ---------------------------------------------------------------------------------------------------------------
s, l = timed(lambda: eng.basis_state(n, 12345), reps=3)
emit("basis_state fill", s, l, bytes=N * ESZ)
ab = np.random.default_rng(3).standard_normal((n, 2)) + 1j * np.random.default_rng(4).standard_normal((n, 2))
ab /= np.linalg.norm(ab, axis=1, keepdims=True)
s, l = timed(lambda: eng.product_state(n, ab), reps=3)
emit("product_state fill", s, l, bytes=N * ESZ)
s, l = timed(lambda: eng.random_state(n, seed=5, normalize=False), reps=3)
emit("random_state fill (counter-based generator)", s, l, bytes=N * ESZ)

# ---- state evolution: fused Pauli rotations ------------------------------------------------------------------------------------------
x, z, c, _, _ = hub.arrays()
s, l = timed(lambda: eng.trotter_slice(hub, psi, coeff=0.05), reps=2)
emit(f"trotter_slice: one first-order slice of the Hubbard chain = {len(x)} Pauli rotations", s, l, rotations=len(x), passes=l,
     rotations_per_pass=len(x) / max(l, 1), bytes=l * 2 * N * ESZ, note="bytes = passes x (read + write) of the state")
n_e = (n // 2) & ~1
pool = workloads.uccsd_pool_fermion_terms(n_e, n)
sel = pool[:: max(len(pool) // 96, 1)][:96]  # a spread of singles and doubles
xs, zs, th = [], [], []
for k, op in enumerate(sel):
    t = eng.terms_from_fermion(n, op)
    xx, zz, cc, _, _ = t.arrays()
    # tau = T - T^dagger is anti-Hermitian: exp(theta tau) = prod exp(-i theta (i c_t) P_t), i c_t real
    xs.append(xx), zs.append(zz), th.append(0.01 * (k + 1) * np.real(1j * cc))
xs, zs, th = np.concatenate(xs), np.concatenate(zs), np.concatenate(th)
s, l = timed(lambda: eng.pauli_rotations(psi, xs, zs, th), reps=1)
emit(f"uccsd layer: exp(theta_k tau_k) for {len(sel)} cluster operators = {len(xs)} Pauli rotations", s, l, rotations=len(xs), passes=l,
     rotations_per_pass=len(xs) / max(l, 1), bytes=l * 2 * N * ESZ, norm=abs(eng.dot(psi, psi)) ** 0.5)
