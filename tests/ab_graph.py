"""A/B probe in GRAPH mode (not a test): iterations/s of the C2 window (device-timed, refactorizations
included) for each value of an engine parameter.  python tests/ab_graph.py <param> <v1> <v2> ..."""
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
        s.setParameter("batch", 32); s.setParameter("warmupIterations", cycle)
        s.setMaximumIterations(5 * cycle); s.setFactorizationFrequency(cycle)
        s.setParameter(key, v)
        s.dual()
        ms, its = s.timedWindow()
        print(key, v, "it/s", round(its / (ms / 1000.0), 1), "us/iter", round(1000.0 * ms / its, 1), "nucleus", s.nucleusSize(), flush=True)
