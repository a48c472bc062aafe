"""Synthetic LP generators.

Two families:

1. The reference's own programmatic known-answer LPs, restated from
   /root/reference/test/test_racing_lp.cpp (N-Queens :122-192, TSP-MTZ :199-272, infeasible
   :277-317, UFL :326-372, set cover/pack/partition :377-467, unbounded :471-491).  They use
   glibc ``srand/rand`` exactly like the C++ file, so the expected bounds of
   test/test_racing_reference.txt apply.  ``tests/golden/make_golden.py`` freezes them to
   .npz fixtures so that nothing depends on libc at test time.

2. The BASELINE.json benchmark shapes (C2..C5), seeded with numpy's PCG64.

Every generator returns an ``LP`` (CSC matrix + rim), the input format of both the CUDA
engine (``clp_b200.ClpSimplex.loadProblem``) and the CPU oracle.
"""
from __future__ import annotations

import ctypes
import dataclasses

import numpy as np

COIN_DBL_MAX = 1.0e30  # we clamp infinities to +-1e30 like the engine does


@dataclasses.dataclass
class LP:
    name: str
    m: int
    n: int
    col_start: np.ndarray  # int32 [n+1]
    row_index: np.ndarray  # int32 [nnz]
    element: np.ndarray  # float64 [nnz]
    col_lower: np.ndarray
    col_upper: np.ndarray
    objective: np.ndarray
    row_lower: np.ndarray
    row_upper: np.ndarray
    known_objective: float | None = None
    expect_status: int = 0  # 0 optimal, 1 primal infeasible, 2 dual infeasible

    @property
    def nnz(self) -> int:
        return int(self.col_start[-1])

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csc_matrix((self.element, self.row_index, self.col_start), shape=(self.m, self.n))

    def save(self, path):
        np.savez_compressed(
            path, name=self.name, m=self.m, n=self.n, col_start=self.col_start,
            row_index=self.row_index, element=self.element, col_lower=self.col_lower,
            col_upper=self.col_upper, objective=self.objective, row_lower=self.row_lower,
            row_upper=self.row_upper,
            known_objective=np.nan if self.known_objective is None else self.known_objective,
            expect_status=self.expect_status)

    @staticmethod
    def load(path) -> "LP":
        z = np.load(path, allow_pickle=False)
        ko = float(z["known_objective"])
        return LP(str(z["name"]), int(z["m"]), int(z["n"]), z["col_start"], z["row_index"],
                  z["element"], z["col_lower"], z["col_upper"], z["objective"], z["row_lower"],
                  z["row_upper"], None if np.isnan(ko) else ko, int(z["expect_status"]))


def from_rows(name, n_cols, rows, row_lb, row_ub, col_lb, col_ub, obj, **kw) -> LP:
    """rows: list of (indices, values) -- the row-major appendRow style of the reference."""
    m = len(rows)
    counts = np.zeros(n_cols + 1, dtype=np.int64)
    for idx, _ in rows:
        for j in idx:
            counts[j + 1] += 1
    start = np.cumsum(counts)
    nnz = int(start[-1])
    ri = np.zeros(nnz, dtype=np.int32)
    el = np.zeros(nnz, dtype=np.float64)
    fill = start[:-1].copy()
    for i, (idx, val) in enumerate(rows):
        for j, v in zip(idx, val):
            ri[fill[j]] = i
            el[fill[j]] = v
            fill[j] += 1
    clamp = lambda a: np.clip(np.asarray(a, dtype=np.float64), -COIN_DBL_MAX, COIN_DBL_MAX)
    return LP(name, m, n_cols, start.astype(np.int32), ri, el, clamp(col_lb), clamp(col_ub),
              np.asarray(obj, dtype=np.float64), clamp(row_lb), clamp(row_ub), **kw)


# ----------------------------------------------------------------------------- glibc rand
class _GlibcRand:
    def __init__(self):
        self.libc = ctypes.CDLL("libc.so.6")
        self.libc.rand.restype = ctypes.c_int

    def srand(self, seed):
        self.libc.srand(ctypes.c_uint(seed))

    def rand(self):
        return self.libc.rand()


# ----------------------------------------------------------------------------- reference LPs
def nqueens(n) -> LP:
    """test_racing_lp.cpp:122-192 ; LP bound -n (:578-581)."""
    ncols = n * n
    rows, lb, ub = [], [], []
    for i in range(n):
        rows.append(([i * n + j for j in range(n)], [1.0] * n)); lb.append(1.0); ub.append(1.0)
    for j in range(n):
        rows.append(([i * n + j for i in range(n)], [1.0] * n)); lb.append(-COIN_DBL_MAX); ub.append(1.0)
    for k in range(-(n - 2), n - 1):
        idx = [i * n + (i - k) for i in range(n) if 0 <= i - k < n]
        if len(idx) > 1:
            rows.append((idx, [1.0] * len(idx))); lb.append(-COIN_DBL_MAX); ub.append(1.0)
    for k in range(1, 2 * n - 2):
        idx = [i * n + (k - i) for i in range(n) if 0 <= k - i < n]
        if len(idx) > 1:
            rows.append((idx, [1.0] * len(idx))); lb.append(-COIN_DBL_MAX); ub.append(1.0)
    return from_rows(f"NQueens-{n}", ncols, rows, lb, ub, [0.0] * ncols, [1.0] * ncols,
                     [-1.0] * ncols, known_objective=-float(n))


def tsp_mtz(n, seed, known=None) -> LP:
    """test_racing_lp.cpp:199-272."""
    r = _GlibcRand(); r.srand(seed)
    nX = n * (n - 1); nU = n - 1; ncols = nX + nU
    xidx = lambda i, j: i * (n - 1) + (j - 1 if j > i else j)
    obj = [0.0] * ncols
    for i in range(n):
        for j in range(n):
            if i != j:
                obj[xidx(i, j)] = 1.0 + (r.rand() % 100)
    clb = [0.0] * ncols; cub = [1.0] * ncols
    for i in range(nU):
        clb[nX + i] = 1.0; cub[nX + i] = float(n - 1)
    rows, lb, ub = [], [], []
    for i in range(n):
        idx = [xidx(i, j) for j in range(n) if j != i]
        rows.append((idx, [1.0] * len(idx))); lb.append(1.0); ub.append(1.0)
    for j in range(n):
        idx = [xidx(i, j) for i in range(n) if i != j]
        rows.append((idx, [1.0] * len(idx))); lb.append(1.0); ub.append(1.0)
    for i in range(1, n):
        for j in range(1, n):
            if i != j:
                rows.append(([nX + i - 1, nX + j - 1, xidx(i, j)], [1.0, -1.0, float(n)]))
                lb.append(-COIN_DBL_MAX); ub.append(float(n - 1))
    return from_rows(f"TSP-MTZ-{n}", ncols, rows, lb, ub, clb, cub, obj, known_objective=known)


def infeasible(n) -> LP:
    """test_racing_lp.cpp:277-317."""
    rows = [(list(range(n)), [1.0] * n), (list(range(n)), [1.0] * n)]
    lb = [1.0, 2.0]; ub = [1.0, 2.0]
    for i in range(n - 1):
        rows.append(([i, i + 1], [1.0, 1.0])); lb.append(-COIN_DBL_MAX); ub.append(5.0)
    return from_rows(f"Infeasible-{n}", n, rows, lb, ub, [0.0] * n, [10.0] * n, [1.0] * n,
                     expect_status=1)


def ufl(nf, nc, seed, known=None) -> LP:
    """test_racing_lp.cpp:326-372."""
    r = _GlibcRand(); r.srand(seed)
    ncols = nf + nf * nc
    obj = [0.0] * ncols
    for i in range(nf):
        obj[i] = 50.0 + (r.rand() % 101)
    for i in range(nf):
        for j in range(nc):
            obj[nf + i * nc + j] = 1.0 + (r.rand() % 50)
    rows, lb, ub = [], [], []
    for j in range(nc):
        rows.append(([nf + i * nc + j for i in range(nf)], [1.0] * nf)); lb.append(1.0); ub.append(1.0)
    for i in range(nf):
        for j in range(nc):
            rows.append(([nf + i * nc + j, i], [1.0, -1.0])); lb.append(-COIN_DBL_MAX); ub.append(0.0)
    return from_rows(f"UFL-{nf}x{nc}", ncols, rows, lb, ub, [0.0] * ncols, [1.0] * ncols, obj,
                     known_objective=known)


def _random01(kind, nrows, ncols, density, seed) -> LP:
    """test_racing_lp.cpp:377-467 (set covering / packing / partitioning)."""
    r = _GlibcRand(); r.srand(seed)
    if kind == "pack":
        obj = [-(1.0 + (r.rand() % 20)) for _ in range(ncols)]
    else:
        obj = [1.0 + (r.rand() % 20) for _ in range(ncols)]
    rows, lb, ub = [], [], []
    thr = int(density * 1000)
    for _ in range(nrows):
        idx = [j for j in range(ncols) if (r.rand() % 1000) < thr]
        if not idx:
            idx = [r.rand() % ncols]
        rows.append((idx, [1.0] * len(idx)))
        if kind == "cover":
            lb.append(1.0); ub.append(COIN_DBL_MAX)
        elif kind == "pack":
            lb.append(-COIN_DBL_MAX); ub.append(1.0)
        else:
            lb.append(1.0); ub.append(1.0)
    name = {"cover": "SetCover", "pack": "SetPack", "part": "SetPart"}[kind]
    return from_rows(f"{name}-{nrows}x{ncols}", ncols, rows, lb, ub, [0.0] * ncols,
                     [1.0] * ncols, obj)


def set_cover(nr, nc, d, seed): return _random01("cover", nr, nc, d, seed)
def set_pack(nr, nc, d, seed): return _random01("pack", nr, nc, d, seed)
def set_part(nr, nc, d, seed): return _random01("part", nr, nc, d, seed)


def unbounded(n) -> LP:
    """test_racing_lp.cpp:471-491."""
    return from_rows(f"Unbounded-{n}", n, [([0, 1], [1.0, -1.0])], [-COIN_DBL_MAX], [10.0],
                     [0.0] * n, [COIN_DBL_MAX] * n, [-1.0] * n, expect_status=2)


def unit_test_3x5() -> LP:
    """src/unitTest.cpp:1415-1431 (the 3x5 LP of ClpSimplexUnitTest)."""
    start = np.array([0, 2, 5, 6, 7, 8], dtype=np.int32)
    rows = np.array([0, 2, 0, 1, 2, 0, 1, 2], dtype=np.int32)
    el = np.array([7.0, 2.0, -2.0, 1.0, -2.0, 1.0, 1.0, 1.0])
    return LP("unitTest-3x5", 3, 5, start, rows, el, np.zeros(5), np.full(5, 100.0),
              np.array([-4.0, 1.0, 0.0, 0.0, 0.0]), np.array([14.0, 3.0, 3.0]),
              np.array([14.0, 3.0, 3.0]))


# ----------------------------------------------------------------------------- BASELINE shapes
def random_sparse_lp(m, n, density, seed, name=None, tight_frac=0.3) -> LP:
    """BASELINE.json configs[1]/[2] (C2/C3): random columns, ~density*m nonzeros each, values
    U(-1,1) with |a|>=0.05, boxed 0<=x<=1, rows ranged around a planted point, costs from a
    planted dual so the LP is feasible and bounded.  Generated column-block-wise in numpy."""
    rng = np.random.default_rng(seed)
    k = np.maximum(1, rng.binomial(m, density, size=n)).astype(np.int64)
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(k, out=start[1:])
    nnz = int(start[-1])
    # distinct rows per column: sample with replacement then de-duplicate by perturbing
    row = rng.integers(0, m, size=nnz, dtype=np.int64)
    col_of = np.repeat(np.arange(n, dtype=np.int64), k)
    key = col_of * m + row
    order = np.argsort(key, kind="stable")
    key = key[order]
    dup = np.concatenate(([False], key[1:] == key[:-1]))
    while dup.any():
        idx = np.nonzero(dup)[0]
        newrow = rng.integers(0, m, size=idx.size, dtype=np.int64)
        key[idx] = (key[idx] // m) * m + newrow
        order2 = np.argsort(key, kind="stable")
        key = key[order2]
        dup = np.concatenate(([False], key[1:] == key[:-1]))
    row = (key % m).astype(np.int32)
    mag = rng.uniform(0.05, 1.0, size=nnz)
    sgn = rng.choice([-1.0, 1.0], size=nnz)
    el = mag * sgn
    import scipy.sparse as sp

    A = sp.csc_matrix((el, row, start), shape=(m, n))
    return _planted_rim(rng, A, m, n, tight_frac, name or f"rand-{m}x{n}", start, row, el)


def _planted_rim(rng, A, m, n, tight_frac, name, start, row, el) -> LP:
    """Bounds, row ranges and costs around a planted primal vertex / dual vector (shared by the two
    random generators; the order of the rng calls is part of the fixtures' identity)."""
    # planted primal vertex: a few columns strictly inside their box (fewer than the number of
    # tight rows, so the planted point is a non-degenerate-ish vertex), the rest at a bound
    tight = rng.uniform(size=m) < tight_frac
    n_inside = min(n // 2, int(0.8 * tight.sum()))
    xs = np.where(rng.uniform(size=n) < 0.5, 0.0, 1.0)
    inside = rng.choice(n, size=n_inside, replace=False)
    xs[inside] = rng.uniform(0.1, 0.9, size=n_inside)
    act = A @ xs
    slack = rng.uniform(0.05, 1.0, size=m)
    row_lower = np.where(tight, act, act - slack)
    row_upper = np.where(rng.uniform(size=m) < 0.5, COIN_DBL_MAX, act + rng.uniform(0.5, 2.0, size=m))
    # planted dual: y>=0 on tight rows (row at lower bound)
    y = np.where(tight, rng.uniform(0.05, 1.0, size=m), 0.0)
    red = rng.uniform(0.05, 1.0, size=n)
    red = np.where(xs == 0.0, red, np.where(xs == 1.0, -red, 0.0))
    c = A.T @ y + red
    return LP(name, m, n, np.asarray(start, dtype=np.int32), row, el, np.zeros(n),
              np.ones(n), c, row_lower, row_upper, known_objective=float(c @ xs))


def random_sparse_lp_large(m, n, density, seed, name=None, tight_frac=0.3) -> LP:
    """BASELINE.json configs[2] (C3, m=50k n=500k, 2.5e8 nonzeros): the recipe of random_sparse_lp
    with a sort-free, chunked matrix generator (the argsort de-duplication of random_sparse_lp needs
    ~15 GB and 3 minutes at this size).  Column j gets k_j ~ Binomial(m, density) (>= 1) rows, one per
    stratum of width m/k_j (distinct and ascending by construction), values U(0.05,1) with a random
    sign; rim vectors from _planted_rim."""
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    k = np.maximum(1, rng.binomial(m, density, size=n)).astype(np.int64)
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(k, out=start[1:])
    nnz = int(start[-1])
    row = np.empty(nnz, dtype=np.int32)
    el = np.empty(nnz, dtype=np.float64)
    chunk = 20000
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        e0, e1 = int(start[c0]), int(start[c1])
        kk = k[c0:c1]
        col_of = np.repeat(np.arange(c1 - c0, dtype=np.int64), kk)
        within = np.arange(e1 - e0, dtype=np.int64) - (start[c0:c1] - e0)[col_of]
        kc = kk[col_of]
        lo = within * m // kc            # integer strata [lo, hi): disjoint, non-empty because k_j <= m
        hi = (within + 1) * m // kc
        r = lo + np.floor(rng.uniform(size=e1 - e0) * (hi - lo)).astype(np.int64)
        np.minimum(r, hi - 1, out=r)
        row[e0:e1] = r
        el[e0:e1] = rng.uniform(0.05, 1.0, size=e1 - e0) * rng.choice([-1.0, 1.0], size=e1 - e0)
    A = sp.csc_matrix((el, row, start), shape=(m, n))
    return _planted_rim(rng, A, m, n, tight_frac, name or f"rand-{m}x{n}", start, row, el)


def staircase_lp(stages=40, block=500, seed=0) -> LP:
    """BASELINE.json configs[3] (C4): staircase, ~10 nz in own stage + ~10 in next stage."""
    rng = np.random.default_rng(seed)
    m = n = stages * block
    rows_l, cols_l, vals_l = [], [], []
    for s in range(stages):
        for part, cnt in ((s, 10), (s + 1, 10)):
            if part >= stages:
                continue
            cc = np.repeat(np.arange(s * block, (s + 1) * block), cnt)
            rr = part * block + rng.integers(0, block, size=cc.size)
            rows_l.append(rr); cols_l.append(cc)
            vals_l.append(rng.uniform(0.1, 1.0, size=cc.size) * rng.choice([-1.0, 1.0], size=cc.size))
    import scipy.sparse as sp

    A = sp.coo_matrix((np.concatenate(vals_l), (np.concatenate(rows_l), np.concatenate(cols_l))),
                      shape=(m, n)).tocsc()
    A.sum_duplicates()
    xs = rng.uniform(0.0, 2.0, size=n) * (rng.uniform(size=n) < 0.5)
    b = A @ xs
    c = rng.uniform(0.0, 1.0, size=n)
    return LP(f"staircase-{m}", m, n, A.indptr.astype(np.int32), A.indices.astype(np.int32),
              A.data.astype(np.float64), np.zeros(n), np.full(n, COIN_DBL_MAX), c, b, b.copy())


def transportation_lp(S=50, D=5000, seed=0) -> LP:
    """BASELINE.json configs[4] (C5): degenerate transportation LP, costs 1+rand%100."""
    rng = np.random.default_rng(seed)
    m, n = S + D, S * D
    demand = rng.integers(1, 20, size=D).astype(np.float64)
    total = demand.sum()
    supply = np.floor(total / S) * np.ones(S)
    supply[: int(total - supply.sum())] += 1.0
    start = np.arange(0, 2 * n + 1, 2, dtype=np.int32)
    row = np.empty(2 * n, dtype=np.int32)
    src = np.repeat(np.arange(S), D); dst = np.tile(np.arange(D), S)
    row[0::2] = src; row[1::2] = S + dst
    el = np.ones(2 * n)
    c = 1.0 + rng.integers(0, 100, size=n).astype(np.float64)
    rl = np.concatenate((np.full(S, -COIN_DBL_MAX), demand))
    ru = np.concatenate((supply, demand))
    return LP(f"transport-{S}x{D}", m, n, start, row, el, np.zeros(n), np.full(n, COIN_DBL_MAX),
              c, rl, ru)


def racing_suite(small_only=False):
    """The cases of test_racing_lp.cpp main() (:572-737) with test_racing_reference.txt bounds."""
    cases = [nqueens(8), nqueens(20), tsp_mtz(20, 42, 172.283333), ufl(10, 30, 99, 560.0),
             infeasible(10), infeasible(50), unbounded(10), unbounded(50),
             set_cover(30, 100, 0.15, 11), set_pack(40, 120, 0.12, 44), set_part(20, 80, 0.20, 77)]
    if not small_only:
        cases += [nqueens(50), nqueens(100), tsp_mtz(40, 123, 189.0), tsp_mtz(60, 7, 197.116667),
                  ufl(20, 60, 77, 770.5), ufl(30, 100, 55, 1040.962), ufl(50, 200, 33, 1433.0516),
                  infeasible(200), set_cover(50, 200, 0.10, 22), set_cover(100, 500, 0.08, 33),
                  set_pack(80, 300, 0.08, 55), set_pack(150, 600, 0.06, 66),
                  set_part(40, 200, 0.12, 88), set_part(60, 400, 0.08, 99)]
    return cases
