"""Scratch tool: run N iterations on GPU and on the oracle (GPU-mimicking ratio test) and
report where the pivot sequences diverge."""
import sys, subprocess, os, re, io, contextlib, tempfile
sys.path.insert(0, ".")
import numpy as np
import clp_b200
from clp_b200 import generators as G
from oracle.oracle import OracleSimplex

m, n, dens, N = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
lp = G.random_sparse_lp(m, n, dens, 20260923)

def run(kind):
    # traces go to the C stderr: capture by redirecting fd 2 to a file
    tf = tempfile.TemporaryFile(mode="w+b")
    old = os.dup(2); os.dup2(tf.fileno(), 2)
    try:
        if kind == "gpu":
            s = clp_b200.ClpSimplex(); s.loadLP(lp); s.setLogLevel(3); s.setMaximumIterations(N); s.setParameter("batch", 16)
            st = s.dual(); obj = s.objectiveValue(); it = s.numberIterations()
        else:
            o = OracleSimplex(lp); o.set_option("logLevel", 3); o.set_option("bucketedRatioTest", 1); o.set_option("maximumIterations", N)
            st = o.dual(); obj = o.objective_value; it = o.iterations
    finally:
        os.dup2(old, 2); os.close(old)
    tf.seek(0); txt = tf.read().decode(errors="ignore")
    tr = [l for l in txt.splitlines() if l.startswith("TRACE")]
    return st, obj, it, tr

g = run("gpu"); c = run("cpu")
print("gpu", g[:3], "cpu", c[:3])
pat = re.compile(r"TRACE (\d+) out=(\d+) in=(\d+) sigma=(-?\d+) thetaD=(\S+) thetaP=(\S+) alpha=(\S+) infeas=(\S+)")
for i, (a, b) in enumerate(zip(g[3], c[3])):
    ma, mb = pat.match(a), pat.match(b)
    if ma.group(2, 3, 4) != mb.group(2, 3, 4):
        print("first divergence at iteration", i); 
        for j in range(max(0, i - 3), min(len(g[3]), i + 3)):
            print(" G", g[3][j]); print(" C", c[3][j])
        break
    else:
        rel = abs(float(ma.group(5)) - float(mb.group(5))) / (1e-12 + abs(float(mb.group(5))))
        if rel > 1e-6:
            print("theta mismatch at", i, a, b)
else:
    print("no divergence in", min(len(g[3]), len(c[3])), "iterations")
