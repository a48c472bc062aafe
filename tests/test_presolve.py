"""Presolve / postsolve (clp_b200/csrc/presolve.cpp) checked on the CPU: the reduced model is solved
by the CPU oracle, the solution is handed to Clpb_postsolve, and the result is audited on the
ORIGINAL problem (KKT incl. row duals, same optimum as the oracle on the original problem)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden

import clp_b200
from clp_b200 import generators as G
from oracle import oracle as O

INF = 1e29


def decorate(lp, seed):
    """Add what the elementary actions remove: fixed columns, singleton rows (some binding, some
    not, one equality), an empty column, an empty row."""
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    A = lp.to_scipy().tocsc()
    m, n = lp.m, lp.n
    o = O.OracleSimplex(lp)
    assert o.dual() == 0
    x0 = o.column_solution()
    col_lower, col_upper, obj = lp.col_lower.copy(), lp.col_upper.copy(), lp.objective.copy()
    # fix a few columns at their optimal value
    for j in rng.choice(n, size=min(5, n), replace=False):
        col_lower[j] = col_upper[j] = x0[j]
    # singleton rows through the optimum x0 of the undecorated LP (so the LP stays feasible); the
    # cost of the column is pushed against the row, so that the row is active with a nonzero dual
    rows, rl, ru = [], [], []
    fixed = col_lower == col_upper
    picks = [j for j in rng.permutation(n) if not fixed[j]][:8]
    for t, j in enumerate(picks):
        a = float(rng.choice([-2.0, 0.5, 3.0]))
        rows.append(sp.csr_matrix(([a], ([0], [j])), shape=(1, n)))
        v = a * x0[j]
        kind = t % 4
        if kind == 0:      # x_j <= x0_j through the row, and the cost wants x_j larger
            obj[j] -= 1.0
            (rl.append(-1e30), ru.append(v)) if a > 0 else (rl.append(v), ru.append(1e30))
        elif kind == 1:    # x_j >= x0_j through the row, and the cost wants x_j smaller
            obj[j] += 1.0
            (rl.append(v), ru.append(1e30)) if a > 0 else (rl.append(-1e30), ru.append(v))
        elif kind == 2:    # slack row
            rl.append(v - 10.0); ru.append(v + 10.0)
        else:              # equality row: fixes the column
            rl.append(v); ru.append(v)
    rows.append(sp.csr_matrix((1, n)))   # empty row
    rl.append(-1.0); ru.append(2.0)
    A2 = sp.vstack([A] + rows).tocsc()
    # an empty column with positive cost and one with zero cost
    A2 = sp.hstack([A2, sp.csc_matrix((A2.shape[0], 2))]).tocsc()
    col_lower = np.concatenate([col_lower, [1.5, -3.0]])
    col_upper = np.concatenate([col_upper, [4.0, 1e30]])
    obj = np.concatenate([obj, [2.0, 0.0]])
    # keep the singleton rows compatible with the column bounds (the LP must stay feasible)
    lp2 = G.LP(lp.name + "+presolvable", A2.shape[0], A2.shape[1], A2.indptr.astype(np.int32),
               A2.indices.astype(np.int32), A2.data.astype(np.float64), col_lower, col_upper, obj,
               np.concatenate([lp.row_lower, rl]), np.concatenate([lp.row_upper, ru]))
    return lp2


def full_audit(lp, s, tol=1e-6):
    x, act, dj, pi, st = (s.primalColumnSolution(), s.primalRowSolution(), s.dualColumnSolution(),
                          s.dualRowSolution(), s.statusArray())
    assert O.kkt_violations(lp, x, act, dj) == 0
    A = lp.to_scipy()
    np.testing.assert_allclose(dj, lp.objective - A.T @ pi, atol=tol * (1 + np.abs(lp.objective).max()))
    # row duals: >= 0 needs the row at its lower bound, <= 0 at its upper bound
    at_lo = np.abs(act - lp.row_lower) <= 1e-5 * (1 + np.abs(act))
    at_up = np.abs(act - lp.row_upper) <= 1e-5 * (1 + np.abs(act))
    assert not np.any((pi > tol) & ~at_lo)
    assert not np.any((pi < -tol) & ~at_up)
    assert int((st == 1).sum()) == lp.m                     # square basis
    assert np.all(np.abs(dj[st[: lp.n] == 1]) <= tol)       # basic columns have zero reduced cost
    assert np.all(np.abs(pi[st[lp.n:] == 1]) <= tol)        # basic rows have zero price


@pytest.mark.parametrize("name,seed", [("UFL-10x30", 1), ("TSP-MTZ-20", 2), ("modified_afiro", 3), ("SetCover-30x100", 4),
                                       ("staircase-480", 5), ("transport-10x200", 6)])
def test_presolve_postsolve_against_oracle(name, seed):
    lp = decorate(load_golden(name), seed)
    ref = O.OracleSimplex(lp)
    st_ref = ref.dual()
    s = clp_b200.ClpSimplex(); s.loadLP(lp)
    st, red = s.presolvedModel()
    assert st_ref == 0 and st == 0
    rlp = red.getProblem()
    assert rlp.n <= lp.n - 5 and rlp.m <= lp.m - 3           # something was removed
    o = O.OracleSimplex(rlp)
    assert o.dual() == 0
    red.setSolution(o.column_solution(), o.row_price(), o.status(), 0)
    assert s.postsolve(red) == 0
    assert abs(s.objectiveValue() - ref.objective_value) <= 1e-8 * (1 + abs(ref.objective_value))
    full_audit(lp, s)


def test_presolve_detects_trivial_infeasibility_and_unboundedness():
    import scipy.sparse as sp

    # singleton rows 2x >= 3 and 2x <= 1
    A = sp.csc_matrix(np.array([[2.0, 0.0], [2.0, 0.0], [0.0, 1.0]]))
    lp = G.LP("inf", 3, 2, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, np.zeros(2), np.full(2, 1e30),
              np.ones(2), np.array([3.0, -1e30, 0.0]), np.array([1e30, 1.0, 5.0]))
    s = clp_b200.ClpSimplex(); s.loadLP(lp)
    assert s.presolvedModel() == (1, None)
    # empty column with negative cost and no upper bound
    A = sp.csc_matrix(np.array([[1.0, 0.0]]))
    lp = G.LP("unb", 1, 2, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, np.zeros(2), np.full(2, 1e30),
              np.array([1.0, -1.0]), np.array([0.0]), np.array([4.0]))
    s = clp_b200.ClpSimplex(); s.loadLP(lp)
    assert s.presolvedModel() == (2, None)


def test_presolve_leaves_irreducible_models_alone():
    lp = load_golden("NQueens-8")
    s = clp_b200.ClpSimplex(); s.loadLP(lp)
    st, red = s.presolvedModel()
    assert st == 0
    r = red.getProblem()
    assert (r.m, r.n, r.nnz) == (lp.m, lp.n, lp.nnz)


@pytest.mark.gpu
@pytest.mark.parametrize("name,seed", [("UFL-10x30", 1), ("SetCover-30x100", 4), ("staircase-480", 5)])
def test_initial_solve_with_presolve_on_the_device(name, seed):
    """ClpSimplex::initialSolve: presolve -> dual simplex on the GPU -> postsolve."""
    lp = decorate(load_golden(name), seed)
    ref = O.OracleSimplex(lp)
    assert ref.dual() == 0
    s = clp_b200.ClpSimplex(); s.loadLP(lp)
    assert s.initialSolve(presolve=True) == 0
    assert abs(s.objectiveValue() - ref.objective_value) <= 1e-8 * (1 + abs(ref.objective_value))
    full_audit(lp, s)


def test_presolve_to_an_empty_model_is_solved_on_the_host():
    """every row a singleton: presolve folds all rows into column bounds and the reduced model has no
    rows -- Engine::dual must not launch zero-block kernels for it (host-only trivial solve), and
    initialSolve reports the optimum of the ORIGINAL model"""
    n = 12
    rng = np.random.default_rng(3)
    start = np.arange(n + 1, dtype=np.int32)
    rows = np.arange(n, dtype=np.int32)
    el = rng.choice([-2.0, 0.5, 3.0], size=n)
    cost = rng.standard_normal(n)
    lp = G.LP("all-singleton", n, n, start, rows, el, np.full(n, -5.0), np.full(n, 7.0), cost,
              np.full(n, -4.0), np.full(n, 6.0))
    s = clp_b200.ClpSimplex()
    s.loadLP(lp)
    assert s.initialSolve() == 0          # no CUDA device is needed for this model
    assert s.status() == 0
    o = O.OracleSimplex(lp)
    assert o.dual() == 0
    assert abs(s.objectiveValue() - o.objective_value) <= 1e-9 * (1 + abs(o.objective_value))
    assert O.kkt_violations(lp, s.primalColumnSolution(), s.primalRowSolution(), s.dualColumnSolution()) == 0


def test_presolve_infeasibility_is_reported_on_the_original_model():
    """a singleton row that contradicts the column bounds: presolve proves infeasibility itself and
    the original model must say so (status 1), not keep a stale -1"""
    start = np.array([0, 1, 2], dtype=np.int32)
    lp = G.LP("contradiction", 2, 2, start, np.array([0, 1], dtype=np.int32), np.array([1.0, 1.0]),
              np.array([0.0, 0.0]), np.array([1.0, 1.0]), np.array([1.0, 1.0]),
              np.array([3.0, -1e30]), np.array([1e30, 5.0]))
    s = clp_b200.ClpSimplex()
    s.loadLP(lp)
    assert s.initialSolve() == 1
    assert s.status() == 1 and s.isProvenPrimalInfeasible()


def test_dual_fixing_removes_dominated_columns():
    """remove_dual_action: columns whose reduced-cost sign follows from the row types alone are fixed at
    the bound that sign prefers; the postsolved solution must be optimal for the ORIGINAL LP (oracle) with
    a clean KKT audit, and the reduced model must really be smaller"""
    import scipy.sparse as sp

    rng = np.random.default_rng(12)
    base = G.random_sparse_lp(60, 300, 0.08, 9)
    A = base.to_scipy().tocsc()
    m, n = base.m, base.n
    # make every row one-sided (>=) so that all row duals are sign restricted, keep the LP feasible
    row_lower = base.row_lower.copy()
    row_upper = np.full(m, 1e30)
    # dominated columns: only positive entries in >= rows and a positive cost -> d_j = c_j - a^T pi can be
    # anything, so NOT dominated; only NEGATIVE entries and a positive cost -> d_j >= c_j > 0: fixed at lower;
    # only positive entries and a negative cost -> d_j <= c_j < 0: fixed at upper
    extra_cols, extra_cost, extra_lo, extra_up = [], [], [], []
    for t in range(20):
        rows = rng.choice(m, size=4, replace=False)
        vals = rng.uniform(0.2, 1.0, size=4)
        if t % 2 == 0:
            extra_cols.append(sp.csc_matrix((-vals, (rows, np.zeros(4, dtype=int))), shape=(m, 1)))
            extra_cost.append(float(rng.uniform(0.1, 1.0)))
        else:
            extra_cols.append(sp.csc_matrix((vals, (rows, np.zeros(4, dtype=int))), shape=(m, 1)))
            extra_cost.append(-float(rng.uniform(0.1, 1.0)))
        extra_lo.append(0.0)
        extra_up.append(2.0)
    A2 = sp.hstack([A] + extra_cols).tocsc()
    lp = G.LP("dominated", m, n + 20, A2.indptr.astype(np.int32), A2.indices.astype(np.int32), A2.data.astype(float),
              np.concatenate([base.col_lower, extra_lo]), np.concatenate([base.col_upper, extra_up]),
              np.concatenate([base.objective, extra_cost]), row_lower, row_upper)
    o = O.OracleSimplex(lp)
    assert o.dual() == 0
    s = clp_b200.ClpSimplex()
    s.loadLP(lp)
    st, red = s.presolvedModel()
    assert st == 0 and red is not None
    assert red.numberColumns() <= lp.n - 20           # at least the planted dominated columns are gone
    rlp = red.getProblem()
    ro = O.OracleSimplex(rlp)
    assert ro.dual() == 0
    red.setSolution(ro.column_solution(), ro.row_price(), ro.status())
    s.postsolve(red)
    assert abs(s.objectiveValue() - o.objective_value) <= 1e-8 * (1 + abs(o.objective_value))
    x = s.primalColumnSolution()
    assert np.all(x[n:][0::2] == 0.0) and np.all(x[n:][1::2] == 2.0)   # where the sign argument puts them
    assert O.kkt_violations(lp, x, s.primalRowSolution(), s.dualColumnSolution()) == 0
    assert int((s.statusArray() == 1).sum()) == lp.m


def test_forcing_rows_fix_their_columns():
    """forcing_constraint_action: rows whose bound equals the extreme activity the column bounds allow pin
    all their columns; postsolve must give the row the dual that keeps every pinned column dual feasible
    (and make one of them basic when that dual is nonzero)"""
    import scipy.sparse as sp

    rng = np.random.default_rng(4)
    base = G.random_sparse_lp(50, 260, 0.08, 3)
    A = base.to_scipy().tocsr()
    m, n = base.m, base.n
    extra_rows, rl, ru = [], [], []
    used = set()
    for t in range(6):
        cols = [j for j in rng.permutation(n) if j not in used][:3]
        used.update(cols)
        coef = rng.choice([-1.5, 0.7, 2.0], size=3)
        extra_rows.append(sp.csr_matrix((coef, (np.zeros(3, dtype=int), cols)), shape=(1, n)))
        lo, up = base.col_lower[cols], base.col_upper[cols]
        min_act = float(np.where(coef > 0, coef * lo, coef * up).sum())
        max_act = float(np.where(coef > 0, coef * up, coef * lo).sum())
        if t % 2 == 0:      # activity <= its minimum: pinned at the lower corner
            rl.append(-1e30); ru.append(min_act)
        else:               # activity >= its maximum
            rl.append(max_act); ru.append(1e30)
    A2 = sp.vstack([A] + extra_rows).tocsc()
    lp = G.LP("forcing", m + 6, n, A2.indptr.astype(np.int32), A2.indices.astype(np.int32), A2.data.astype(float),
              base.col_lower, base.col_upper, base.objective,
              np.concatenate([base.row_lower, rl]), np.concatenate([base.row_upper, ru]))
    o = O.OracleSimplex(lp)
    ost = o.dual()
    s = clp_b200.ClpSimplex()
    s.loadLP(lp)
    st, red = s.presolvedModel()
    if ost != 0:
        pytest.skip("pinning made this instance infeasible")
    assert st == 0 and red is not None
    assert red.numberRows() <= lp.m - 6 and red.numberColumns() <= lp.n - 18
    rlp = red.getProblem()
    ro = O.OracleSimplex(rlp)
    assert ro.dual() == 0
    red.setSolution(ro.column_solution(), ro.row_price(), ro.status())
    s.postsolve(red)
    assert abs(s.objectiveValue() - o.objective_value) <= 1e-8 * (1 + abs(o.objective_value))
    assert O.kkt_violations(lp, s.primalColumnSolution(), s.primalRowSolution(), s.dualColumnSolution()) == 0
    assert int((s.statusArray() == 1).sum()) == lp.m
