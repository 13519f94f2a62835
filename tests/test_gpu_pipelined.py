"""Parity of the pipelined EXPECT and APPLY kernels (csrc/sweep.cu:tile_expect_kernel, tile_apply_kernel: next tile requested before the
barrier, table-driven tile offsets) against the oracle at sizes the oracle finishes in seconds.  The kernel is the default for real-coefficient sweeps over full tiles when every SM
has its own tile (>= 20 qubits); QFE_SPLIT=1 makes small registers take the same path, so the check runs in a child process with that
knob and with QFE_PROF=1, whose counters prove which kernel ran.  Tolerances: BASELINE.json (1e-10 complex128, 1e-5 complex64)."""

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

CHILD = r"""
import ctypes as C, json, sys
import numpy as np, torch
sys.path.insert(0, %r)
import myqlm_fermion_b200 as qfe
from oracle import models, spin
eng = qfe.get_engine(0)
out = []
def state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return v / np.linalg.norm(v)
def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
def cycles():
    c = np.zeros(128, dtype=np.uint64)
    eng.lib.qfe_dev_phase_cycles(eng.ctx, c.ctypes.data_as(C.POINTER(C.c_uint64)), 1)
    return c.reshape(16, 8)
for m, enc in [(7, "jordan-wigner"), (7, "bravyi-kitaev"), (7, "parity"), (8, "jordan-wigner")]:
    n = 2 * m
    one, two = models.synthetic_integrals(m, seed=1234)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    terms = eng.terms_from_integrals(hpq, hpqrs, constant=0.7, encoding=enc)
    x, z, c, _, k = terms.arrays()
    psi, bra = state(n, 0), state(n, 5)
    phi = spin.apply_packed(n, x, z, c, complex(k[0]), psi)
    plan = eng.plan(terms)
    cycles()
    e = eng.expectation(plan, dev(psi))
    tiles = int(cycles()[:, 7].sum())
    t = eng.expectation(plan, dev(psi), bra=dev(bra))
    psi32 = psi.astype(np.complex64)
    e32 = eng.expectation(plan, dev(psi32))
    l0 = eng.launch_count()
    phi_dev = eng.apply(plan, dev(psi)).cpu().numpy()        # pipelined APPLY kernel (tile_apply_kernel): unsplit full tiles
    apply_launches = eng.launch_count() - l0
    phi_acc = dev(psi.copy())                                 # y += H psi on a non-zero y: two applies of the same plan through apply_class
    eng.apply_class(plan, 0, dev(psi), phi_acc, accumulate=True)
    phi32 = eng.apply(plan, dev(psi32)).cpu().numpy()
    want32 = np.vdot(psi32.astype(np.complex128), spin.apply_packed(n, x, z, c, complex(k[0]), psi32.astype(np.complex128)))
    out.append(dict(m=m, enc=enc, tile_bits=plan.tile_bits, async_tiles=tiles, e=[e.real, e.imag], want_e=[np.vdot(psi, phi).real, np.vdot(psi, phi).imag],
                    t=[t.real, t.imag], want_t=[np.vdot(bra, phi).real, np.vdot(bra, phi).imag], e32=[e32.real, e32.imag], want32=[want32.real, want32.imag],
                    apply_dev=float(np.max(np.abs(phi_dev - phi)) / np.max(np.abs(phi))), apply_launches=apply_launches, sweeps=plan.n_sweeps,
                    apply_acc_dev=float(np.max(np.abs(phi_acc.cpu().numpy() - (psi + phi))) / np.max(np.abs(phi))),
                    apply32_dev=float(np.max(np.abs(phi32 - phi)) / np.max(np.abs(phi)))))
print("RESULT " + json.dumps(out))
"""


def test_pipelined_expect_kernel_matches_the_oracle():
    env = dict(os.environ, QFE_SPLIT="1", QFE_PROF="1", QFE_PIPELINED="1")
    r = subprocess.run([sys.executable, "-c", CHILD % str(ROOT)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    for row in json.loads(line[7:]):
        assert row["tile_bits"] == 13
        assert row["async_tiles"] > 0, "the pipelined kernel did not run"
        e, w = complex(*row["e"]), complex(*row["want_e"])
        assert abs(e - w) <= 1e-10 * abs(w), row
        t, w = complex(*row["t"]), complex(*row["want_t"])
        assert abs(t - w) <= 1e-10 * max(abs(w), 1e-3), row
        e, w = complex(*row["e32"]), complex(*row["want32"])
        assert abs(e - w) <= 1e-5 * abs(w), row
        assert row["apply_launches"] == row["sweeps"], row  # one launch per sweep: no slab combine, i.e. the unsplit path ran
        assert row["apply_dev"] <= 1e-10 and row["apply_acc_dev"] <= 1e-10 and row["apply32_dev"] <= 1e-4, row
