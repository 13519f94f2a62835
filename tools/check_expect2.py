"""Development aid (one GPU): a newer generation of the EXPECT kernel against the one before it (csrc/sweep.cu: tile_expect2_kernel against
tile_expect_kernel with the knob QFE_EXPECT2; any other on/off knob of the library can be named) on the same inputs -- <psi|H|psi>, the transition <phi|H|psi>, the real-part-only instance the sharded driver uses (QFE_FORCE_REAL_ONLY=1),
complex128 and complex64, three encodings -- and their timings.  usage: python tools/check_expect2.py [M=11] [reps=3] [KNOB=QFE_EXPECT2]"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

CHILD = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, %r)
import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200.workloads import synthetic_h_integrals
m, reps = %d, %d
n = 2 * m
eng = qfe.get_engine(0)
eng.use_torch_stream()
out = []
for enc in ("jordan-wigner", "bravyi-kitaev", "parity"):
    hpq, hpqrs = synthetic_h_integrals(m, seed=1234)
    terms = eng.terms_from_integrals(hpq, hpqrs, encoding=enc)
    plan = eng.plan(terms)
    row = dict(enc=enc, n=n, terms=terms.n_terms, sweeps=plan.n_sweeps)
    for name, dt in (("c128", torch.complex128), ("c64", torch.complex64)):
        psi = eng.random_state(n, dtype=dt, seed=0)
        bra = eng.random_state(n, dtype=dt, seed=1)
        e = eng.expectation(plan, psi)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            e = eng.expectation(plan, psi)
        ev1.record()
        torch.cuda.synchronize()
        t = eng.expectation(plan, psi, bra=bra)
        r = eng.expectation_class(plan, 0, bra, psi)
        row[name] = dict(e=[e.real, e.imag], t=[t.real, t.imag], r=[r.real, r.imag], ms=ev0.elapsed_time(ev1) / reps)
    out.append(row)
print("RESULT " + json.dumps(out))
"""


def run(env_extra, m, reps):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", CHILD % (str(ROOT), m, reps)], env=env, capture_output=True, text=True, timeout=150)
    if r.returncode != 0:
        print(r.stderr[-3000:])
        raise SystemExit(1)
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    knob = sys.argv[3] if len(sys.argv) > 3 else "QFE_EXPECT2"  # the kernel generation under test: knob=0 against knob=1
    old = run({knob: "0"}, m, reps)
    new = run({knob: "1"}, m, reps)
    old_r = run({knob: "0", "QFE_FORCE_REAL_ONLY": "1"}, m, 1)
    new_r = run({knob: "1", "QFE_FORCE_REAL_ONLY": "1"}, m, 1)
    ok = True
    for a, b, ar, br in zip(old, new, old_r, new_r):
        for name, tol in (("c128", 1e-11), ("c64", 2e-5)):
            ea, eb = complex(*a[name]["e"]), complex(*b[name]["e"])
            ta, tb = complex(*a[name]["t"]), complex(*b[name]["t"])
            ra, rb = ar[name]["r"][0], br[name]["r"][0]
            scale = max(abs(ea), 1e-3)
            good = abs(ea - eb) <= tol * scale and abs(ta - tb) <= tol * scale and abs(ra - rb) <= tol * scale and abs(rb - tb.real) <= tol * scale
            ok = ok and good
            print(json.dumps(dict(enc=a["enc"], n=a["n"], dtype=name, terms=a["terms"], sweeps=a["sweeps"], ok=good, e_old=a[name]["e"], e_new=b[name]["e"],
                                  t_old=a[name]["t"], t_new=b[name]["t"], re_only_old=ra, re_only_new=rb, ms_old=a[name]["ms"], ms_new=b[name]["ms"],
                                  speedup=a[name]["ms"] / b[name]["ms"])), flush=True)
    print("expect2 parity ok" if ok else "expect2 parity FAILED", flush=True)
    raise SystemExit(0 if ok else 1)


if __name__ == "__main__":
    main()
