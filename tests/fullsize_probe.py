"""Full-size parity probe (not a pytest test: minutes of GPU time): BASELINE.json configs[3] (C4,
staircase m=n=20 000) and configs[4] (C5, degenerate transportation 5 050 x 250 000) solved on the
GPU from the all-slack basis and audited: KKT on the true data (oracle.kkt_violations) and the
optimum against HiGHS dual simplex (tests/golden/fullsize_highs.json, produced in the build
container by tests/golden/make_fullsize_highs.py).  python tests/fullsize_probe.py [c5] [c4]"""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np
import clp_b200
from clp_b200 import generators as G
from oracle import oracle as O

ref = {}
p = os.path.join("tests", "golden", "fullsize_highs.json")
if os.path.exists(p):
    ref = json.load(open(p))
which = sys.argv[1:] or ["c5", "c4"]
cases = {"c5": ("C5 transport-50x5000", lambda: G.transportation_lp(50, 5000, 20260926), [{}, {"perturbation": 50}]),
         "c4": ("C4 staircase-20000", lambda: G.staircase_lp(40, 500, 20260925), [{}]),
         "c4s": ("C4 staircase-20000", lambda: G.staircase_lp(40, 500, 20260925), [{"scaling": 3, "perturbation": 50}])}
out = []
for w in which:
    name, gen, variants = cases[w]
    lp = gen()
    for params in variants:
        s = clp_b200.ClpSimplex(); s.loadLP(lp)
        for k, v in params.items():
            s.setParameter(k, v)
        s.setParameter("maximumSeconds", 600)
        t = time.time(); st = s.dual(); el = time.time() - t
        kkt = O.kkt_violations(lp, s.primalColumnSolution(), s.primalRowSolution(), s.dualColumnSolution()) if st == 0 else None
        r = {"case": name, "params": params, "m": lp.m, "n": lp.n, "nnz": lp.nnz, "status": st, "objective": s.objectiveValue(),
             "iterations": s.numberIterations(), "refactorizations": s.numberRefactorizations(), "seconds": round(el, 2),
             "iterations_per_sec": round(s.numberIterations() / max(1e-9, s.secondsInLoop()), 1), "nucleus": s.nucleusSize(),
             "kkt_violations": kkt, "highs_objective": ref.get(name, {}).get("objective")}
        if r["highs_objective"] is not None and st == 0:
            r["rel_diff_vs_highs"] = abs(r["objective"] - r["highs_objective"]) / (1 + abs(r["highs_objective"]))
        print(json.dumps(r), flush=True)
        out.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.path.join("gpurun_out", "fullsize_probe.json"), "w"), indent=1)
