"""CPU restatement (numpy) of ClpPackedMatrix::scale -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/ClpPackedMatrix.cpp:4120-4640 step by step (useful columns :4185-4240,
"don't bother" :4262, equilibrium row maxima :4340-4358, geometric passes :4365-4452, range fix
:4459-4469, overall ratio :4471-4491, auto choice :4493-4513, final column pass :4531-4581).
Only tests/ may import this module; the product computes its own factors in
clp_b200/csrc/engine.cu (Engine::computeScaling) and the two are compared in
tests/test_oracle_golden.py::test_scale_factors_match_restatement.
"""
import numpy as np


def _group_max_min(values, groups, ngroups, big0, small0):
    big = np.full(ngroups, big0)
    small = np.full(ngroups, small0)
    np.maximum.at(big, groups, values)
    np.minimum.at(small, groups, values)
    return big, small


def scale_factors(lp, mode, primal_tolerance=1e-7):
    """Returns (rc, rowScale, columnScale); rc == 1: not worth scaling (all ones)."""
    m, n = lp.m, lp.n
    ones = (1, np.ones(m), np.ones(n))
    if mode <= 0 or m == 0 or n == 0:
        return ones
    start = np.asarray(lp.col_start, dtype=np.int64)
    length = np.diff(start)
    col_of = np.repeat(np.arange(n), length)
    row_of = np.asarray(lp.row_index, dtype=np.int64)
    a = np.abs(np.asarray(lp.element, dtype=np.float64))
    free_col = lp.col_upper > lp.col_lower + 1.0e-12
    nz = a > 1.0e-20
    in_useful = free_col[col_of] & nz
    useful = np.zeros(n, dtype=bool)
    useful[col_of[in_useful]] = True
    if not in_useful.any():
        largest, smallest = 0.0, 1.0e50
    else:
        largest, smallest = a[in_useful].max(), a[in_useful].min()
    if smallest >= 0.5 and largest <= 2.0:
        return ones
    ue = useful[col_of]  # entries of useful columns (tiny ones included, as in the reference loops)
    method = {4: 3}.get(mode, 2 if mode >= 5 else mode)
    saved = 0.0
    tol = 5.0 * primal_tolerance
    rdiff = lp.row_upper - lp.row_lower
    while True:
        r = np.ones(m)
        c = np.ones(n)
        if method in (1, 3):
            big, _ = _group_max_min(a[ue], row_of[ue], m, 1.0e-10, 1.0e50)
            r = 1.0 / big
        else:
            passes = 3
            while passes:
                passes -= 1
                v = a[ue] * c[col_of[ue]]
                big, small = _group_max_min(v, row_of[ue], m, 1.0e-50, 1.0e50)
                r = 1.0 / np.sqrt(small * big)
                if passes == 1:
                    break
                v = a[ue] * r[row_of[ue]]
                big, small = _group_max_min(v, col_of[ue], n, 1.0e-50, 1.0e50)
                c = np.where(useful, 1.0 / np.sqrt(small * big), c)
        with np.errstate(invalid="ignore", over="ignore"):
            sd = rdiff * r
        fix = (sd > tol) & (sd < 1.0e-4)
        r = np.where(fix, np.clip(r * (1.0e-4 / np.where(fix, sd, 1.0)), 1.0e-10, 1.0e10), r)
        v = a[ue] * r[row_of[ue]]
        big, small = _group_max_min(v, col_of[ue], n, 1.0e-20, 1.0e50)
        # sequential rule "if (overallSmallest*largest > smallest) overallSmallest = smallest/largest"
        # is a running minimum of smallest/largest
        overall_smallest = min(1.0e50, (small[useful] / big[useful]).min()) if useful.any() else 1.0e50
        if method in (1, 2):
            break
        if saved == 0.0 and method != 4:
            saved = overall_smallest
            method = 4
        elif overall_smallest > 2.0 * saved:
            break
        else:
            method = 1
    overall_largest = 1.0
    if overall_smallest < 1.0e-1:
        overall_largest = 1.0 / np.sqrt(overall_smallest)
    overall_largest = min(100.0, overall_largest)
    scaled_col = free_col & (length > 0)
    se = scaled_col[col_of]
    big, _ = _group_max_min(a[se] * r[row_of[se]], col_of[se], n, 1.0e-20, 1.0e50)
    c = np.where(scaled_col, overall_largest / big, 1.0)
    cdiff = lp.col_upper - lp.col_lower
    c = np.where(scaled_col & (cdiff < 1.0e-5 * c), cdiff / 1.0e-5, c)
    used_row = np.zeros(m, dtype=bool)
    used_row[row_of[se]] = True
    r = np.where(used_row, r, 1.0)
    return 0, r, c
