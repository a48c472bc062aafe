"""Full-size BASELINE.json configurations solved on the GPU from the all-slack basis at default
settings (no scaling, no perturbation) and audited (-m gpu; each takes up to a few minutes):
  C2  random 10 000 x 100 000, 1 % -- self-certifying planted optimum c^T x* (generators.py)
  C4  staircase 20 000 x 20 000    -- optimum pinned by HiGHS dual simplex (tests/golden/fullsize_highs.json)
  C5  transportation 5 050 x 250 000 (fully degenerate) -- optimum pinned by HiGHS
Checks: status 0, objective within the reference's CoinRelFltEq(1e-8) (src/unitTest.cpp:1930) of the
pinned value, the reference's KKT audit (test/test_racing_lp.cpp:36-116 tolerances, oracle.kkt_violations)
on the TRUE data, a square basis (exactly m basic variables), and a duality certificate computed here in
numpy from the returned (x, pi, d) alone: primal objective == dual objective to 1e-8 with both sides
feasible proves optimality without a second solver (HiGHS serial dual simplex does not finish C4 in two
hours, its IPM hangs in the crossover start basis; C2 itself is the subject of
tests/fullsize_probe.py -- millions of iterations -- and is represented here by the same generator at
3 000 x 30 000, which the CPU oracle also solves)."""
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


def duality_gap(lp, x, pi, dj):
    """(primal objective, dual objective, worst dual sign violation) on the true data, min form with
    d = c - A^T pi: pi_i > 0 prices the row's lower bound, pi_i < 0 its upper bound; d_j likewise for the
    column bounds.  A price of the wrong sign against an infinite bound is a dual infeasibility."""
    A = lp.to_scipy().tocsr()
    primal = float(lp.objective @ x)
    d = lp.objective - A.T @ pi
    assert np.allclose(d, dj, rtol=0, atol=1e-7 * (1 + np.abs(d).max()))
    inf = 1e29

    def side(price, lo, up):
        lo_ok, up_ok = lo > -inf, up < inf
        val = np.where(price > 0, np.where(lo_ok, lo, 0.0), np.where(up_ok, up, 0.0)) * price
        bad = np.where(price > 0, np.where(lo_ok, 0.0, price), np.where(up_ok, 0.0, -price))
        return float(val.sum()), float(bad.max(initial=0.0))
    vr, br = side(pi, lp.row_lower, lp.row_upper)
    vc, bc = side(d, lp.col_lower, lp.col_upper)
    return primal, vr + vc, max(br, bc)


def _solve_and_audit(lp, expected, rel_tol=1e-8, **params):
    s = clp_b200.ClpSimplex()
    s.loadLP(lp)
    for k, v in params.items():
        s.setParameter(k, v)
    s.setParameter("maximumSeconds", 900)
    st = s.dual()
    assert st == 0, (st, s.objectiveValue(), s.numberIterations())
    obj = s.objectiveValue()
    primal, dual, bad = duality_gap(lp, s.primalColumnSolution(), s.dualRowSolution(), s.dualColumnSolution())
    assert bad <= 1e-6, bad
    assert abs(primal - dual) <= 1e-7 * (1.0 + abs(primal)), (primal, dual)
    assert abs(obj - primal) <= 1e-9 * (1.0 + abs(primal))
    if expected is not None:
        assert abs(obj - expected) <= rel_tol * (1.0 + abs(expected)), (obj, expected)
    assert O.kkt_violations(lp, s.primalColumnSolution(), s.primalRowSolution(), s.dualColumnSolution()) == 0
    assert int((s.statusArray() == 1).sum()) == lp.m
    return s


def test_c2_family_reaches_planted_optimum():
    lp = G.random_sparse_lp(3000, 30000, 0.01, 20260923)
    _solve_and_audit(lp, lp.known_objective)


def test_c5_full_size_transportation_default_settings():
    lp = G.transportation_lp(50, 5000, 20260926)
    _solve_and_audit(lp, HIGHS["C5 transport-50x5000"]["objective"])


def test_c4_full_size_staircase_default_settings():
    lp = G.staircase_lp(40, 500, 20260925)
    _solve_and_audit(lp, HIGHS.get("C4 staircase-20000", {}).get("objective"))
