"""A/B probe (not a test): one full factorization cycle of the C2 window in timing mode for each
value of an engine parameter, same process, same GPU.  python tests/ab_probe.py <param> <v1> <v2> ..."""
import sys
sys.path.insert(0, ".")
import clp_b200
from bench import build_workload, default_cycle

key, values = sys.argv[1], [float(v) for v in sys.argv[2:]]
lp, status, start = build_workload("c2")
cycle = default_cycle(lp.m)
for rep in range(2):
    for v in values:
        s = clp_b200.ClpSimplex(); s.loadLP(lp); s.copyinStatus(status)
        s.setParameter("timing", 1); s.setParameter("batch", 16)
        s.setMaximumIterations(cycle); s.setFactorizationFrequency(cycle)
        s.setParameter(key, v)
        s.dual()
        ph = s.phaseTimes(); ns = max(1.0, ph["samples"])
        print(key, v, {k: round(1000.0 * ph[k] / ns, 1) for k in ("btran", "price", "chuzc", "ftran", "update")},
              "sum", round(1000.0 * sum(ph[k] for k in ("chuzr", "btran", "price", "chuzc", "dualUpdate", "ftran", "update")) / ns, 1),
              "gemv_us", {q: round(1000.0 * ph[q] / ns, 2) for q in ("ftranGemv", "btranGemv", "priceKernel")},
              "refactor_ms", round(ph["refactor"], 1), flush=True)
