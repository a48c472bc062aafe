"""Full-size BASELINE.json configurations solved on the GPU from the all-slack basis at default
settings (no scaling, no perturbation) and audited (-m gpu; each takes up to a few minutes):
  C2  random 10 000 x 100 000, 1 % -- self-certifying planted optimum c^T x* (generators.py)
  C4  staircase 20 000 x 20 000    -- optimum pinned by HiGHS dual simplex (tests/golden/fullsize_highs.json)
  C5  transportation 5 050 x 250 000 (fully degenerate) -- optimum pinned by HiGHS
Checks: status 0, objective within the reference's CoinRelFltEq(1e-8) (src/unitTest.cpp:1930) of the
pinned value, the reference's KKT audit (test/test_racing_lp.cpp:36-116 tolerances, oracle.kkt_violations)
on the TRUE data, and a square basis (exactly m basic variables)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT

import clp_b200
from clp_b200 import generators as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu

HIGHS = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_highs.json")))


def _solve_and_audit(lp, expected, rel_tol=1e-8, **params):
    s = clp_b200.ClpSimplex()
    s.loadLP(lp)
    for k, v in params.items():
        s.setParameter(k, v)
    s.setParameter("maximumSeconds", 900)
    st = s.dual()
    assert st == 0, (st, s.objectiveValue(), s.numberIterations())
    obj = s.objectiveValue()
    assert abs(obj - expected) <= rel_tol * (1.0 + abs(expected)), (obj, expected)
    assert O.kkt_violations(lp, s.primalColumnSolution(), s.primalRowSolution(), s.dualColumnSolution()) == 0
    assert int((s.statusArray() == 1).sum()) == lp.m
    return s


def test_c2_full_size_reaches_planted_optimum():
    lp = G.random_sparse_lp(10000, 100000, 0.01, 20260923)
    _solve_and_audit(lp, lp.known_objective)


def test_c5_full_size_transportation_default_settings():
    lp = G.transportation_lp(50, 5000, 20260926)
    _solve_and_audit(lp, HIGHS["C5 transport-50x5000"]["objective"])


@pytest.mark.skipif("C4 staircase-20000" not in HIGHS, reason="no independent optimum recorded for C4")
def test_c4_full_size_staircase_default_settings():
    lp = G.staircase_lp(40, 500, 20260925)
    _solve_and_audit(lp, HIGHS["C4 staircase-20000"]["objective"])
