"""Netlib AFIRO (BASELINE.json configs[0], C1): tests/golden/afiro.mps is the public Netlib model (27 rows + the
objective row, 32 columns, 83 matrix entries) typed in by hand -- it is not under /root/reference, which only
names it (src/unitTest.cpp:480-486: mpsName "afiro", nRows 28, nCols 32, objValue -4.6475314286e+02,
objValueTol 1e-8).  That published optimum is what pins the transcription: a single wrong coefficient moves
it.  This script parses the file with THIS repo's MPS reader, checks the oracle and HiGHS against the
reference's value and freezes the fixture afiro.npz + its manifest entry.   python tests/golden/make_afiro.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import clp_b200  # noqa: E402
from make_golden import highs_objective  # noqa: E402
from oracle.oracle import OracleSimplex  # noqa: E402

REFERENCE_VALUE = -4.6475314286e+02  # src/unitTest.cpp:486

s = clp_b200.ClpSimplex()
assert s.readMps(os.path.join(HERE, "afiro.mps")) == 0
lp = s.getProblem()
lp.name = "afiro"
assert (lp.m + 1, lp.n, lp.nnz) == (28, 32, 83)          # nRows counts the objective row
o = OracleSimplex(lp)
assert o.dual() == 0
assert abs(o.objective_value - REFERENCE_VALUE) <= 1e-8 * (1 + abs(REFERENCE_VALUE)), o.objective_value
st, hobj = highs_objective(lp)
assert st == 0 and abs(hobj - REFERENCE_VALUE) <= 1e-8 * (1 + abs(REFERENCE_VALUE))
lp.known_objective = REFERENCE_VALUE
lp.save(os.path.join(HERE, "afiro.npz"))
mp = os.path.join(HERE, "manifest.json")
manifest = json.load(open(mp))
manifest["afiro"] = {"m": lp.m, "n": lp.n, "nnz": lp.nnz, "expect_status": 0, "known_objective": REFERENCE_VALUE,
                     "objective_source": "reference", "oracle_objective": o.objective_value,
                     "oracle_iterations": o.iterations}
json.dump(manifest, open(mp, "w"), indent=1, sort_keys=True)
print("afiro", manifest["afiro"], "highs", hobj)
