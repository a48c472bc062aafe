"""Scratch probe (not a test): solve a planted random LP on the GPU with phase timing."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import clp_b200
from clp_b200 import generators as G

m, n = int(sys.argv[1]), int(sys.argv[2])
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
timing = int(sys.argv[4]) if len(sys.argv) > 4 else 1
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 16
freq = int(sys.argv[6]) if len(sys.argv) > 6 else 0
t = time.time(); lp = G.random_sparse_lp(m, n, 0.01, 20260923); print("gen", round(time.time() - t, 2), "nnz", lp.nnz, flush=True)
s = clp_b200.ClpSimplex(); s.loadLP(lp)
s.setParameter("timing", timing); s.setParameter("batch", batch); s.setMaximumIterations(maxit)
if freq: s.setFactorizationFrequency(freq)
s.setLogLevel(int(sys.argv[7]) if len(sys.argv) > 7 else 0)
t = time.time(); st = s.dual(); el = time.time() - t
ph = s.phaseTimes()
print(json.dumps({"status": st, "obj": s.objectiveValue(), "known": lp.known_objective, "iters": s.numberIterations(),
                  "refactors": s.numberRefactorizations(), "wall_s": el, "loop_s": s.secondsInLoop(),
                  "it_per_s": s.numberIterations() / max(1e-9, s.secondsInLoop()), "k": s.nucleusSize(), "phase_ms": ph}))
if ph["samples"]:
    print({k: round(v / ph["samples"] * 1000, 1) for k, v in ph.items() if k not in ("samples", "refactor")}, "us/iter; refactor total ms", ph["refactor"])
