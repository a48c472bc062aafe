// presolve.cpp -- a small presolve / postsolve pair around the dual simplex path (host only).
//
// ClpPresolve::presolvedModel / postsolve (/root/reference/src/ClpPresolve.hpp:40,61) drive a list
// of CoinPresolveAction objects that live in CoinUtils (not in the reference tree).  This file
// restates the four elementary ones the driver always applies (src/ClpPresolve.cpp: make_fixed :966,
// slack_doubleton_action :1141, drop_empty_cols_action :1448, drop_empty_rows_action :1449) and the dual
// fixing of remove_dual_action (:1158, :1296) and forcing_constraint_action (:1182), iterated to a fixed point:
//   F  fixed column (l == u)       : removed, row bounds shifted by -a_ij x_j, constant c_j x_j
//   S  singleton row a_ij x_j      : becomes bounds on x_j, row removed
//   C  empty column                : set to the bound its cost prefers (unbounded if that is infinite)
//   R  empty row                   : removed (infeasible if 0 is outside its bounds)
//   G  forcing row                 : the smallest (largest) activity the column bounds allow equals the
//                                    row's upper (lower) bound: every column of the row sits on the bound
//                                    that attains it; they are fixed there and the row is removed
//   D  dominated column            : the sign restrictions of the row duals (pi_i >= 0 on a row with only a
//                                    lower bound, <= 0 with only an upper bound, 0 on a free row) already
//                                    force d_j = c_j - sum a_ij pi_i >= 0 (<= 0): some optimal solution has
//                                    x_j at its lower (upper) bound, so it is fixed there like an F column
// postsolve undoes the stack in reverse order on (x, pi, status): a removed row comes back basic with
// pi = 0, except a singleton row whose implied bound is the one x_j sits on -- then the row takes
// over (pi_i = d_j / a_ij, d_j = 0, the column becomes basic, the row nonbasic), which keeps the
// basis square and the solution dual feasible.  Conventions: d_j = c_j - sum_i a_ij pi_i, the dual
// of a row at its lower bound is >= 0 (engine.cu / update.cu reduced_cost_kernel).
#include "engine.hpp"

#include <algorithm>
#include <cmath>

namespace clpb {

namespace {
constexpr double kFeasTol = 1.0e-9;
}

int Presolve::presolve(const Engine &src, Engine &dst)
{
  m = src.m;
  n = src.n;
  const std::vector<int> &cs = src.hColStart, &ri = src.hRow;
  const std::vector<double> &va = src.hVal;
  lower = src.hLower;
  upper = src.hUpper;
  cost.assign(src.hCost.begin(), src.hCost.begin() + n);
  colStart = cs;
  rowIdx = ri;
  val = va;
  actions.clear();
  colAlive.assign(n, 1);
  rowAlive.assign(m, 1);
  offset = src.objectiveOffset;
  // row-wise copy for the singleton test
  std::vector<int> rowStart(m + 1, 0), colIdx(ri.size());
  std::vector<double> rval(ri.size());
  for (size_t e = 0; e < ri.size(); e++)
    rowStart[ri[e] + 1]++;
  for (int i = 0; i < m; i++)
    rowStart[i + 1] += rowStart[i];
  {
    std::vector<int> fill(rowStart.begin(), rowStart.end() - 1);
    for (int j = 0; j < n; j++)
      for (int e = cs[j]; e < cs[j + 1]; e++) {
        const int at = fill[ri[e]]++;
        colIdx[at] = j;
        rval[at] = va[e];
      }
  }
  std::vector<int> rowCount(m, 0), colCount(n, 0);
  for (int j = 0; j < n; j++)
    for (int e = cs[j]; e < cs[j + 1]; e++)
      if (va[e] != 0.0) {
        rowCount[ri[e]]++;
        colCount[j]++;
      }
  bool changed = true;
  while (changed) {
    changed = false;
    // ---- F: fixed columns
    for (int j = 0; j < n; j++) {
      if (!colAlive[j] || !(lower[j] > -kInf) || upper[j] - lower[j] > 0.0)
        continue;
      if (upper[j] < lower[j] - kFeasTol)
        return 1;
      const double x = lower[j];
      for (int e = cs[j]; e < cs[j + 1]; e++) {
        const int i = ri[e];
        if (!rowAlive[i] || va[e] == 0.0)
          continue;
        if (lower[n + i] > -kInf)
          lower[n + i] -= va[e] * x;
        if (upper[n + i] < kInf)
          upper[n + i] -= va[e] * x;
        rowCount[i]--;
      }
      offset += cost[j] * x;
      colAlive[j] = 0;
      actions.push_back({'F', j, -1, x, 0.0, 0.0, 0.0, 0.0});
      changed = true;
    }
    // ---- G: forcing rows
    for (int i = 0; i < m; i++) {
      if (!rowAlive[i] || rowCount[i] < 2)
        continue;
      double minAct = 0.0, maxAct = 0.0;
      bool minFinite = true, maxFinite = true;
      for (int e = rowStart[i]; e < rowStart[i + 1]; e++) {
        const int j = colIdx[e];
        const double a = rval[e];
        if (!colAlive[j] || a == 0.0)
          continue;
        const double lo = lower[j], up = upper[j];
        if (a > 0.0) {
          lo > -kInf ? (void)(minAct += a * lo) : (void)(minFinite = false);
          up < kInf ? (void)(maxAct += a * up) : (void)(maxFinite = false);
        } else {
          up < kInf ? (void)(minAct += a * up) : (void)(minFinite = false);
          lo > -kInf ? (void)(maxAct += a * lo) : (void)(maxFinite = false);
        }
      }
      const double rlo = lower[n + i], rup = upper[n + i];
      const double tolUp = kFeasTol * (1.0 + std::fabs(rup)), tolLo = kFeasTol * (1.0 + std::fabs(rlo));
      if ((minFinite && rup < kInf && minAct > rup + 1.0e3 * tolUp) || (maxFinite && rlo > -kInf && maxAct < rlo - 1.0e3 * tolLo))
        return 1; // the column bounds cannot satisfy the row
      int side = 0; // +1: activity pinned at the row's upper bound by the minimum, -1: at the lower by the maximum
      if (minFinite && rup < kInf && std::fabs(minAct - rup) <= tolUp)
        side = +1;
      else if (maxFinite && rlo > -kInf && std::fabs(maxAct - rlo) <= tolLo)
        side = -1;
      if (side == 0)
        continue;
      for (int e = rowStart[i]; e < rowStart[i + 1]; e++) {
        const int j = colIdx[e];
        const double a = rval[e];
        if (!colAlive[j] || a == 0.0)
          continue;
        // side +1: a > 0 -> lower bound, a < 0 -> upper bound; side -1: the other way round
        const bool toLower = (a > 0.0) == (side > 0);
        const double x = toLower ? lower[j] : upper[j];
        for (int q = cs[j]; q < cs[j + 1]; q++) {
          const int r2 = ri[q];
          if (!rowAlive[r2] || va[q] == 0.0 || r2 == i)
            continue;
          if (lower[n + r2] > -kInf)
            lower[n + r2] -= va[q] * x;
          if (upper[n + r2] < kInf)
            upper[n + r2] -= va[q] * x;
          rowCount[r2]--;
        }
        offset += cost[j] * x;
        colAlive[j] = 0;
        // value = x, oldLo: 1 if the column went to its lower bound, row = the forcing row
        actions.push_back({'g', j, i, x, toLower ? 1.0 : 0.0, a, 0.0, 0.0});
      }
      rowAlive[i] = 0;
      rowCount[i] = 0;
      actions.push_back({'G', -1, i, (double)side, 0.0, 0.0, 0.0, 0.0});
      changed = true;
    }
    // ---- D: columns whose reduced cost sign is decided by the row types alone
    for (int j = 0; j < n; j++) {
      if (!colAlive[j] || colCount[j] == 0 || !(upper[j] - lower[j] > 0.0))
        continue;
      bool canBeNegative = cost[j] < 0.0, canBePositive = cost[j] > 0.0;
      for (int e = cs[j]; e < cs[j + 1] && !(canBeNegative && canBePositive); e++) {
        const int i = ri[e];
        if (!rowAlive[i] || va[e] == 0.0)
          continue;
        const bool hasLo = lower[n + i] > -kInf, hasUp = upper[n + i] < kInf;
        if (hasLo && hasUp) { // equality or range: the dual is free
          canBeNegative = canBePositive = true;
        } else if (hasLo) { // pi_i >= 0: the term -a pi is <= 0 for a > 0, >= 0 for a < 0
          (va[e] > 0.0 ? canBeNegative : canBePositive) = true;
        } else if (hasUp) { // pi_i <= 0
          (va[e] > 0.0 ? canBePositive : canBeNegative) = true;
        } // free row: pi_i = 0
      }
      double x;
      if (!canBeNegative && lower[j] > -kInf)
        x = lower[j]; // d_j >= 0 whatever the duals are
      else if (!canBePositive && upper[j] < kInf)
        x = upper[j]; // d_j <= 0
      else
        continue;
      for (int e = cs[j]; e < cs[j + 1]; e++) {
        const int i = ri[e];
        if (!rowAlive[i] || va[e] == 0.0)
          continue;
        if (lower[n + i] > -kInf)
          lower[n + i] -= va[e] * x;
        if (upper[n + i] < kInf)
          upper[n + i] -= va[e] * x;
        rowCount[i]--;
      }
      offset += cost[j] * x;
      colAlive[j] = 0;
      actions.push_back({'D', j, -1, x, 0.0, 0.0, 0.0, 0.0});
      changed = true;
    }
    // ---- S: singleton rows
    for (int i = 0; i < m; i++) {
      if (!rowAlive[i] || rowCount[i] != 1)
        continue;
      int j = -1;
      double a = 0.0;
      for (int e = rowStart[i]; e < rowStart[i + 1]; e++)
        if (colAlive[colIdx[e]] && rval[e] != 0.0) {
          j = colIdx[e];
          a = rval[e];
        }
      if (j < 0)
        continue;
      double lo = lower[n + i], up = upper[n + i];
      double nlo = a > 0 ? (lo > -kInf ? lo / a : -kInf) : (up < kInf ? up / a : -kInf);
      double nup = a > 0 ? (up < kInf ? up / a : kInf) : (lo > -kInf ? lo / a : kInf);
      const double oldLo = lower[j], oldUp = upper[j];
      const double newLo = std::max(oldLo, nlo), newUp = std::min(oldUp, nup);
      if (newLo > newUp + kFeasTol * (1.0 + std::fabs(newLo)))
        return 1;
      lower[j] = newLo;
      upper[j] = std::max(newLo, newUp);
      rowAlive[i] = 0;
      colCount[j]--;
      actions.push_back({'S', j, i, a, oldLo, oldUp, nlo, nup});
      changed = true;
    }
    // ---- C: empty columns
    for (int j = 0; j < n; j++) {
      if (!colAlive[j] || colCount[j] != 0)
        continue;
      double x;
      if (cost[j] > 0.0) {
        if (!(lower[j] > -kInf))
          return 2;
        x = lower[j];
      } else if (cost[j] < 0.0) {
        if (!(upper[j] < kInf))
          return 2;
        x = upper[j];
      } else {
        x = lower[j] > -kInf ? lower[j] : (upper[j] < kInf ? upper[j] : 0.0);
      }
      offset += cost[j] * x;
      colAlive[j] = 0;
      actions.push_back({'C', j, -1, x, 0.0, 0.0, 0.0, 0.0});
      changed = true;
    }
    // ---- R: empty rows
    for (int i = 0; i < m; i++) {
      if (!rowAlive[i] || rowCount[i] != 0)
        continue;
      if (lower[n + i] > kFeasTol || upper[n + i] < -kFeasTol)
        return 1;
      rowAlive[i] = 0;
      actions.push_back({'R', -1, i, 0.0, 0.0, 0.0, 0.0, 0.0});
      changed = true;
    }
  }
  // ---- reduced model
  colMap.assign(n, -1);
  rowMap.assign(m, -1);
  int nr = 0, mr = 0;
  for (int j = 0; j < n; j++)
    if (colAlive[j])
      colMap[j] = nr++;
  for (int i = 0; i < m; i++)
    if (rowAlive[i])
      rowMap[i] = mr++;
  std::vector<int> start(nr + 1, 0), idx;
  std::vector<double> el, cl(nr), cu(nr), ob(nr), rl(mr), ru(mr);
  for (int j = 0; j < n; j++) {
    if (!colAlive[j])
      continue;
    const int jj = colMap[j];
    for (int e = cs[j]; e < cs[j + 1]; e++)
      if (rowAlive[ri[e]] && va[e] != 0.0) {
        idx.push_back(rowMap[ri[e]]);
        el.push_back(va[e]);
      }
    start[jj + 1] = (int)idx.size();
    cl[jj] = lower[j];
    cu[jj] = upper[j];
    ob[jj] = cost[j];
  }
  for (int i = 0; i < m; i++)
    if (rowAlive[i]) {
      rl[rowMap[i]] = lower[n + i];
      ru[rowMap[i]] = upper[n + i];
    }
  dst.loadProblem(nr, mr, start.data(), idx.data(), el.data(), cl.data(), cu.data(), ob.data(), rl.data(),
                  ru.data());
  dst.objectiveOffset = offset;
  dst.problemName = src.problemName;
  return 0;
}

// x[n], rowActivity[m], pi[m], status[n+m] of the ORIGINAL problem from the reduced solution
// (xr[nr], pir[mr], statusR[nr+mr]); reducedCost[n+m] follows from pi.
void Presolve::postsolve(const std::vector<double> &xr, const std::vector<double> &pir,
                         const std::vector<unsigned char> &statusR, const Engine &orig,
                         std::vector<double> &solution, std::vector<double> &reducedCost,
                         std::vector<double> &rowPrice, std::vector<unsigned char> &status) const
{
  const int nm = n + m;
  int nr = 0;
  for (int j = 0; j < n; j++)
    if (colAlive[j])
      nr++;
  solution.assign(nm, 0.0);
  rowPrice.assign(m, 0.0);
  status.assign(nm, atLowerBound);
  for (int j = 0; j < n; j++)
    if (colAlive[j]) {
      solution[j] = xr[colMap[j]];
      status[j] = statusR[colMap[j]];
    }
  for (int i = 0; i < m; i++)
    if (rowAlive[i]) {
      rowPrice[i] = pir[rowMap[i]];
      status[n + i] = statusR[nr + rowMap[i]];
    } else
      status[n + i] = basic; // until the action below says otherwise
  auto dj = [&](int j) {
    double d = cost[j];
    for (int e = colStart[j]; e < colStart[j + 1]; e++)
      d -= val[e] * rowPrice[rowIdx[e]];
    return d;
  };
  for (int a = (int)actions.size() - 1; a >= 0; a--) {
    const Action &ac = actions[a];
    switch (ac.kind) {
    case 'F':
      solution[ac.col] = ac.value;
      status[ac.col] = isFixed;
      break;
    case 'G': {
      // the forced columns of this row are the 'g' actions right below on the stack; their values and
      // bound sides are known, the row dual is the tightest one that keeps all their reduced costs on
      // the feasible side (row at upper: pi <= 0, at lower: pi >= 0); if it is nonzero the column that
      // attains it becomes basic and the row nonbasic at that bound, otherwise the row stays basic
      const int i = ac.row;
      const int side = (int)ac.value;
      for (int b = a - 1; b >= 0 && actions[b].kind == 'g' && actions[b].row == i; b--) {
        solution[actions[b].col] = actions[b].value;
        status[actions[b].col] = actions[b].oldLo != 0.0 ? atLowerBound : atUpperBound;
      }
      rowPrice[i] = 0.0;
      status[n + i] = basic;
      double best = 0.0;
      int bestCol = -1;
      for (int b = a - 1; b >= 0 && actions[b].kind == 'g' && actions[b].row == i; b--) {
        const int j = actions[b].col;
        const double aij = actions[b].oldUp; // coefficient of the column in the forcing row
        const double ratio = dj(j) / aij;    // pi_i that makes d_j exactly zero
        if ((side > 0 && ratio < best) || (side < 0 && ratio > best)) {
          best = ratio;
          bestCol = j;
        }
      }
      if (bestCol >= 0) {
        rowPrice[i] = best;
        status[bestCol] = basic;
        status[n + i] = side > 0 ? atUpperBound : atLowerBound;
      }
      break;
    }
    case 'g':
      break; // handled by the 'G' entry of its row
    case 'D':
    case 'C': {
      solution[ac.col] = ac.value;
      const double lo = orig.hLower[ac.col], up = orig.hUpper[ac.col];
      status[ac.col] = (lo > -kInf && ac.value == lo) ? atLowerBound
                       : (up < kInf && ac.value == up) ? atUpperBound
                                                       : isFree;
      break;
    }
    case 'R':
      status[n + ac.row] = basic;
      rowPrice[ac.row] = 0.0;
      break;
    case 'S': {
      const int j = ac.col, i = ac.row;
      const double aij = ac.value, x = solution[j];
      status[n + i] = basic;
      rowPrice[i] = 0.0;
      if (status[j] != basic) {
        // does x_j sit on a bound that only this row implies, and is that the bound holding it?
        // (d > 0: held from below, d < 0: held from above; a column at its OWN bound stays nonbasic)
        const double tol = 1.0e-9 * (1.0 + std::fabs(x));
        const bool onRowLower = ac.impliedLo > ac.oldLo && std::fabs(x - ac.impliedLo) <= tol;
        const bool onRowUpper = ac.impliedUp < ac.oldUp && std::fabs(x - ac.impliedUp) <= tol;
        const bool atOwnBound = (ac.oldLo > -kInf && std::fabs(x - ac.oldLo) <= tol) ||
                                (ac.oldUp < kInf && std::fabs(x - ac.oldUp) <= tol);
        const double d = dj(j);
        bool take = false, rowAtLower = true;
        if (d > 0.0 && onRowLower) {
          take = true;
          rowAtLower = aij > 0.0; // x_j's lower bound is the row's lower (a > 0) / upper (a < 0) bound
        } else if (d < 0.0 && onRowUpper) {
          take = true;
          rowAtLower = aij < 0.0;
        } else if (!atOwnBound && (onRowLower || onRowUpper)) {
          take = true; // strictly inside its own bounds: must be basic (d is zero up to rounding)
          rowAtLower = onRowLower ? aij > 0.0 : aij < 0.0;
        }
        if (take) {
          rowPrice[i] = d / aij;
          status[j] = basic;
          status[n + i] = rowAtLower ? atLowerBound : atUpperBound;
        }
      }
      break;
    }
    default:
      break;
    }
  }
  // row activities and reduced costs of the full problem
  for (int i = 0; i < m; i++)
    solution[n + i] = 0.0;
  for (int j = 0; j < n; j++)
    for (int e = colStart[j]; e < colStart[j + 1]; e++)
      solution[n + rowIdx[e]] += val[e] * solution[j];
  reducedCost.assign(nm, 0.0);
  for (int j = 0; j < n; j++)
    reducedCost[j] = status[j] == basic ? 0.0 : dj(j);
  for (int i = 0; i < m; i++)
    reducedCost[n + i] = status[n + i] == basic ? 0.0 : rowPrice[i];
}

} // namespace clpb
