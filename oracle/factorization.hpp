// factorization.hpp -- CPU oracle: sparse LU of the simplex basis with Forrest-Tomlin update.
//
// TEST INFRASTRUCTURE ONLY (see clp_oracle.h).  Restates, on the CPU, the behaviour of
//   ClpFactorization::factorize            /root/reference/src/ClpFactorization.cpp:1649
//   CoinAbcTypeFactorization::factor       src/CoinAbcBaseFactorization1.cpp:683
//     factorSparse (Markowitz, singletons) src/CoinAbcBaseFactorization2.cpp:18
//     factorDense  (dense tail)            src/CoinAbcBaseFactorization2.cpp:976
//   updateColumn / updateColumnFT (FTRAN)  src/CoinAbcBaseFactorization3.cpp:68-2348
//   updateColumnTranspose (BTRAN)          src/CoinAbcBaseFactorization4.cpp:3216
//   checkReplacePart1/2, replaceColumnPart3 (Forrest-Tomlin) ...4.cpp:1254,1844,1870
// It is an independent restatement (own data structures), not a copy: U is held by columns in
// an append-only arena, row deletions of the FT update are represented by time stamps
// (entry (i,j) is live iff column j was born after row i was last eliminated), and the FT
// row-eta is taken from the U-part of the BTRAN of the pivot row, as the reference's
// checkReplacePart1 computes it (row r of U -> BTRAN-U -> multipliers).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>

namespace orc {

class Factorization {
public:
  // --- parameters (ClpFactorization.hpp:149-508 scalar accessors) ---
  double pivotTolerance = 0.1;   // threshold partial pivoting
  double zeroTolerance = 1.0e-13;
  double singularTolerance = 1.0e-11;
  int maximumPivots = 200;
  double denseThreshold = 0.25; // switch nucleus to dense LU above this density
  int threads = 1;

  // --- state ---
  int m = 0;
  int numberPivots = 0; // FT updates since last factorize
  long nnzL = 0, nnzU = 0, nnzR = 0;
  int numberDense = 0; // size of dense tail at last factorize
  std::vector<int> rowOfPos, posOfRow;

  // L etas by pivot position (entries at later positions)
  std::vector<long> Lstart;
  std::vector<int> Lidx;
  std::vector<double> Lval;
  // U columns in an append-only arena (off-diagonal entries), diag separate
  std::vector<long> Ustart;
  std::vector<int> Ulen;
  std::vector<int> Uidx;
  std::vector<double> Uval;
  std::vector<double> diag;
  std::vector<int> Ubirth;  // stamp at which the column was installed
  std::vector<int> rowElim; // stamp at which the row was last eliminated (-1 never)
  std::vector<int> order;   // positions in triangular order, -1 = tombstone
  std::vector<int> slotOfPos;
  int stamp = 0;
  // R row-etas of the FT updates
  std::vector<int> Rpos;
  std::vector<long> Rstart; // size nR+1
  std::vector<int> Ridx;
  std::vector<double> Rval;
  // saved spike (after L and R, before U) and saved BTRAN-U vector for the FT update
  std::vector<double> spike;
  bool spikeValid = false;
  std::vector<double> btranU;
  int btranUPos = -1;
  // scratch
  std::vector<double> work;
  // singular info from last factorize: columns of the input that could not be pivoted
  std::vector<int> rejectedColumns, unpivotedRows;

  /* Factorize an m x m basis given by columns (CSC).  rowOfColumn[c] receives the pivot
     row assigned to input column c (or -1 if rejected).  Returns 0 or -1 (singular). */
  int factorize(int mIn, const std::vector<long> &bStart, const std::vector<int> &bRow,
                const std::vector<double> &bVal, std::vector<int> &rowOfColumn);

  /* FTRAN: v (row indexed, length m) -> B^-1 v indexed by pivot row label.
     saveSpike keeps the partially transformed column for replaceColumn. */
  void ftran(double *v, bool saveSpike = false);
  /* BTRAN: v (indexed by pivot row label) -> B^-T v (row indexed).
     unitRow>=0 says v==e_unitRow (lets us keep the U-part for the FT update). */
  void btran(double *v, int unitRow = -1);
  /* Forrest-Tomlin: replace the column pivoting on 'pivotRow' by the column whose spike
     was saved by the last ftran(...,true).  alphaCheck = tableau pivot element.
     0 ok, 1 ok but inaccurate, 2 singular (nothing changed), 3 no room, 5 max pivots. */
  int replaceColumn(int pivotRow, double alphaCheck);

  long numberElements() const { return nnzL + nnzU + nnzR; }

private:
  void denseTail(const std::vector<int> &actRows, const std::vector<int> &actCols,
                 std::vector<std::vector<int>> &rowCols, std::vector<std::vector<double>> &rowVals,
                 std::vector<int> &pivRow, std::vector<int> &pivCol,
                 std::vector<std::vector<std::pair<int, double>>> &Lsteps,
                 std::vector<std::vector<std::pair<int, double>>> &Usteps,
                 std::vector<char> &colDone, std::vector<char> &rowDone);
};

// ---------------------------------------------------------------------------------------
inline void Factorization::denseTail(
    const std::vector<int> &actRows, const std::vector<int> &actCols,
    std::vector<std::vector<int>> &rowCols, std::vector<std::vector<double>> &rowVals,
    std::vector<int> &pivRow, std::vector<int> &pivCol,
    std::vector<std::vector<std::pair<int, double>>> &Lsteps,
    std::vector<std::vector<std::pair<int, double>>> &Usteps, std::vector<char> &colDone,
    std::vector<char> &rowDone)
{
  // Dense partial-pivoting LU of the remaining active block (factorDense analogue).
  const int d = static_cast<int>(actRows.size());
  numberDense = d;
  std::vector<int> colLocal(m, -1);
  for (int c = 0; c < d; c++)
    colLocal[actCols[c]] = c;
  std::vector<double> D(static_cast<size_t>(d) * d, 0.0); // column major
  for (int r = 0; r < d; r++) {
    int i = actRows[r];
    for (size_t e = 0; e < rowCols[i].size(); e++) {
      int c = colLocal[rowCols[i][e]];
      if (c >= 0)
        D[static_cast<size_t>(c) * d + r] = rowVals[i][e];
    }
  }
  std::vector<int> rowLabel(actRows);
  const int NB = 48;
  std::vector<char> stepOk(d, 1);
  for (int kb = 0; kb < d; kb += NB) {
    int nb = std::min(NB, d - kb);
    // panel
    for (int j = kb; j < kb + nb; j++) {
      double *colj = &D[static_cast<size_t>(j) * d];
      int piv = j;
      double best = std::fabs(colj[j]);
      for (int i = j + 1; i < d; i++) {
        double a = std::fabs(colj[i]);
        if (a > best) {
          best = a;
          piv = i;
        }
      }
      if (best < singularTolerance) {
        stepOk[j] = 0;
        continue; // leave column; handled as singular below
      }
      if (piv != j) {
        for (int c = 0; c < d; c++)
          std::swap(D[static_cast<size_t>(c) * d + j], D[static_cast<size_t>(c) * d + piv]);
        std::swap(rowLabel[j], rowLabel[piv]);
      }
      double inv = 1.0 / colj[j];
      for (int i = j + 1; i < d; i++)
        colj[i] *= inv;
      for (int c = j + 1; c < kb + nb; c++) {
        double *colc = &D[static_cast<size_t>(c) * d];
        double u = colc[j];
        if (u != 0.0)
          for (int i = j + 1; i < d; i++)
            colc[i] -= u * colj[i];
      }
    }
    int rest = d - kb - nb;
    if (rest <= 0)
      continue;
      // U12 = L11^-1 A12 and A22 -= L21 U12, column by column of the trailing matrix
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1 && rest > 64)
    for (int c = kb + nb; c < d; c++) {
      double *colc = &D[static_cast<size_t>(c) * d];
      for (int l = kb; l < kb + nb; l++) {
        if (!stepOk[l])
          continue;
        double u = colc[l];
        if (u == 0.0)
          continue;
        const double *coll = &D[static_cast<size_t>(l) * d];
        for (int i = l + 1; i < d; i++)
          colc[i] -= u * coll[i];
      }
    }
  }
  // emit
  for (int k = 0; k < d; k++) {
    if (!stepOk[k])
      continue; // rejected: row rowLabel[k]/col actCols[k] stay unpivoted
    int prow = rowLabel[k], pcol = actCols[k];
    pivRow.push_back(prow);
    pivCol.push_back(pcol);
    rowDone[prow] = 1;
    colDone[pcol] = 1;
    std::vector<std::pair<int, double>> lcol, urow;
    const double *colk = &D[static_cast<size_t>(k) * d];
    for (int i = k + 1; i < d; i++)
      if (std::fabs(colk[i]) > zeroTolerance)
        lcol.emplace_back(rowLabel[i], colk[i]);
    urow.emplace_back(pcol, colk[k]);
    for (int c = k + 1; c < d; c++) {
      double v = D[static_cast<size_t>(c) * d + k];
      if (std::fabs(v) > zeroTolerance && stepOk[c])
        urow.emplace_back(actCols[c], v);
    }
    Lsteps.push_back(std::move(lcol));
    Usteps.push_back(std::move(urow));
  }
}

inline int Factorization::factorize(int mIn, const std::vector<long> &bStart,
                                    const std::vector<int> &bRow, const std::vector<double> &bVal,
                                    std::vector<int> &rowOfColumn)
{
  m = mIn;
  numberPivots = 0;
  numberDense = 0;
  stamp = 0;
  spikeValid = false;
  btranUPos = -1;
  rejectedColumns.clear();
  unpivotedRows.clear();
  // active submatrix: rows hold (col,val); columns hold row patterns
  std::vector<std::vector<int>> rowCols(m), colRows(m);
  std::vector<std::vector<double>> rowVals(m);
  for (int c = 0; c < m; c++) {
    for (long e = bStart[c]; e < bStart[c + 1]; e++) {
      if (bVal[e] == 0.0)
        continue;
      int i = bRow[e];
      rowCols[i].push_back(c);
      rowVals[i].push_back(bVal[e]);
      colRows[c].push_back(i);
    }
  }
  std::vector<char> rowDone(m, 0), colDone(m, 0);
  std::vector<int> pivRow, pivCol;
  pivRow.reserve(m);
  pivCol.reserve(m);
  std::vector<std::vector<std::pair<int, double>>> Lsteps, Usteps; // labels, not positions
  Lsteps.reserve(m);
  Usteps.reserve(m);
  long nnzActive = 0;
  for (int i = 0; i < m; i++)
    nnzActive += static_cast<long>(rowCols[i].size());

  auto removeFromCol = [&](int c, int i) {
    std::vector<int> &v = colRows[c];
    for (size_t e = 0; e < v.size(); e++)
      if (v[e] == i) {
        v[e] = v.back();
        v.pop_back();
        return;
      }
  };
  auto removeFromRow = [&](int i, int c, double &valOut) {
    std::vector<int> &rc = rowCols[i];
    std::vector<double> &rv = rowVals[i];
    for (size_t e = 0; e < rc.size(); e++)
      if (rc[e] == c) {
        valOut = rv[e];
        rc[e] = rc.back();
        rv[e] = rv.back();
        rc.pop_back();
        rv.pop_back();
        return true;
      }
    valOut = 0.0;
    return false;
  };

  std::vector<int> colStack, rowStack;
  for (int c = 0; c < m; c++)
    if (colRows[c].size() == 1)
      colStack.push_back(c);
  for (int i = 0; i < m; i++)
    if (rowCols[i].size() == 1)
      rowStack.push_back(i);
  int nActive = m;

  // do one pivot on (i,c) where either column c is a singleton or row i is a singleton
  auto pivotColumnSingleton = [&](int i, int c) {
    std::vector<std::pair<int, double>> urow;
    double pv = 0.0;
    for (size_t e = 0; e < rowCols[i].size(); e++) {
      int c2 = rowCols[i][e];
      if (c2 == c)
        pv = rowVals[i][e];
    }
    urow.emplace_back(c, pv);
    for (size_t e = 0; e < rowCols[i].size(); e++) {
      int c2 = rowCols[i][e];
      if (c2 == c)
        continue;
      urow.emplace_back(c2, rowVals[i][e]);
      removeFromCol(c2, i);
      if (colRows[c2].size() == 1)
        colStack.push_back(c2);
    }
    nnzActive -= static_cast<long>(rowCols[i].size());
    rowCols[i].clear();
    rowVals[i].clear();
    colRows[c].clear();
    rowDone[i] = 1;
    colDone[c] = 1;
    pivRow.push_back(i);
    pivCol.push_back(c);
    Lsteps.emplace_back();
    Usteps.push_back(std::move(urow));
    nActive--;
  };
  auto pivotRowSingleton = [&](int i, int c) {
    double pv = rowVals[i][0];
    std::vector<std::pair<int, double>> lcol, urow;
    urow.emplace_back(c, pv);
    for (size_t e = 0; e < colRows[c].size(); e++) {
      int i2 = colRows[c][e];
      if (i2 == i)
        continue;
      double v;
      removeFromRow(i2, c, v);
      nnzActive--;
      if (std::fabs(v) > zeroTolerance)
        lcol.emplace_back(i2, v / pv);
      if (rowCols[i2].size() == 1)
        rowStack.push_back(i2);
    }
    nnzActive--;
    rowCols[i].clear();
    rowVals[i].clear();
    colRows[c].clear();
    rowDone[i] = 1;
    colDone[c] = 1;
    pivRow.push_back(i);
    pivCol.push_back(c);
    Lsteps.push_back(std::move(lcol));
    Usteps.push_back(std::move(urow));
    nActive--;
  };
  auto drainSingletons = [&]() {
    bool any = true;
    while (any) {
      any = false;
      while (!colStack.empty()) {
        int c = colStack.back();
        colStack.pop_back();
        if (colDone[c] || colRows[c].size() != 1)
          continue;
        int i = colRows[c][0];
        double pv = 0.0;
        for (size_t e = 0; e < rowCols[i].size(); e++)
          if (rowCols[i][e] == c)
            pv = rowVals[i][e];
        if (std::fabs(pv) < singularTolerance)
          continue;
        pivotColumnSingleton(i, c);
        any = true;
      }
      while (!rowStack.empty()) {
        int i = rowStack.back();
        rowStack.pop_back();
        if (rowDone[i] || rowCols[i].size() != 1)
          continue;
        int c = rowCols[i][0];
        if (std::fabs(rowVals[i][0]) < singularTolerance)
          continue;
        // threshold test against the rest of the column
        double cmax = 0.0;
        for (size_t e = 0; e < colRows[c].size(); e++) {
          int i2 = colRows[c][e];
          for (size_t f = 0; f < rowCols[i2].size(); f++)
            if (rowCols[i2][f] == c)
              cmax = std::max(cmax, std::fabs(rowVals[i2][f]));
        }
        if (std::fabs(rowVals[i][0]) < pivotTolerance * cmax)
          continue; // leave for Markowitz
        pivotRowSingleton(i, c);
        any = true;
        if (!colStack.empty())
          break;
      }
      if (!colStack.empty())
        any = true;
    }
  };
  drainSingletons();

  // Markowitz on the nucleus, dense switch when it fills
  std::vector<double> wval(m, 0.0);
  std::vector<int> wmark(m, -1);
  int markStamp = 0;
  std::vector<int> activeCols;
  for (int c = 0; c < m; c++)
    if (!colDone[c])
      activeCols.push_back(c);
  while (nActive > 0) {
    // compact active column list
    {
      size_t w = 0;
      for (size_t e = 0; e < activeCols.size(); e++)
        if (!colDone[activeCols[e]])
          activeCols[w++] = activeCols[e];
      activeCols.resize(w);
    }
    if (activeCols.empty())
      break;
    double density = static_cast<double>(nnzActive) / (static_cast<double>(nActive) * nActive);
    if ((density > denseThreshold && nActive >= 16) || (nActive < 16 && density > 0.5)) {
      std::vector<int> actRows;
      for (int i = 0; i < m; i++)
        if (!rowDone[i])
          actRows.push_back(i);
      std::vector<int> actCols(activeCols);
      if (actRows.size() == actCols.size())
        denseTail(actRows, actCols, rowCols, rowVals, pivRow, pivCol, Lsteps, Usteps, colDone,
                  rowDone);
      break;
    }
    // candidate columns: the few with smallest count
    int bestC = -1, bestI = -1;
    double bestCost = 1e300, bestAbs = 0.0;
    int cand[4];
    int ncand = 0;
    {
      size_t cnt[4] = {0, 0, 0, 0};
      for (size_t e = 0; e < activeCols.size(); e++) {
        int c = activeCols[e];
        size_t len = colRows[c].size();
        if (len == 0)
          continue;
        int posn = ncand;
        while (posn > 0 && cnt[posn - 1] > len)
          posn--;
        if (posn < 4) {
          for (int q = std::min(ncand, 3); q > posn; q--) {
            cand[q] = cand[q - 1];
            cnt[q] = cnt[q - 1];
          }
          cand[posn] = c;
          cnt[posn] = len;
          if (ncand < 4)
            ncand++;
        }
      }
    }
    for (int t = 0; t < ncand; t++) {
      int c = cand[t];
      double cmax = 0.0;
      std::vector<std::pair<int, double>> ents;
      for (size_t e = 0; e < colRows[c].size(); e++) {
        int i = colRows[c][e];
        double v = 0.0;
        for (size_t f = 0; f < rowCols[i].size(); f++)
          if (rowCols[i][f] == c) {
            v = rowVals[i][f];
            break;
          }
        ents.emplace_back(i, v);
        cmax = std::max(cmax, std::fabs(v));
      }
      if (cmax < singularTolerance)
        continue;
      for (auto &pr : ents) {
        if (std::fabs(pr.second) < pivotTolerance * cmax)
          continue;
        double cost = static_cast<double>(rowCols[pr.first].size() - 1) *
                      static_cast<double>(colRows[c].size() - 1);
        if (cost < bestCost || (cost == bestCost && std::fabs(pr.second) > bestAbs)) {
          bestCost = cost;
          bestAbs = std::fabs(pr.second);
          bestC = c;
          bestI = pr.first;
        }
      }
    }
    if (bestC < 0) {
      // all candidate columns numerically empty: reject the emptiest and go on
      int c = -1;
      for (size_t e = 0; e < activeCols.size(); e++) {
        int cc = activeCols[e];
        if (c < 0 || colRows[cc].size() < colRows[c].size())
          c = cc;
      }
      for (size_t e = 0; e < colRows[c].size(); e++) {
        double v;
        removeFromRow(colRows[c][e], c, v);
        nnzActive--;
      }
      colRows[c].clear();
      colDone[c] = 2; // rejected
      nActive--;       // keeps counts square-ish; matching row is found at the end
      continue;
    }
    // eliminate with pivot (bestI,bestC)
    int pi = bestI, pc = bestC;
    double pv = 0.0;
    std::vector<std::pair<int, double>> urow, lcol;
    for (size_t e = 0; e < rowCols[pi].size(); e++)
      if (rowCols[pi][e] == pc)
        pv = rowVals[pi][e];
    urow.emplace_back(pc, pv);
    for (size_t e = 0; e < rowCols[pi].size(); e++) {
      int c2 = rowCols[pi][e];
      if (c2 == pc)
        continue;
      urow.emplace_back(c2, rowVals[pi][e]);
      removeFromCol(c2, pi);
    }
    nnzActive -= static_cast<long>(rowCols[pi].size());
    std::vector<int> others(colRows[pc]);
    for (int i2 : others) {
      if (i2 == pi)
        continue;
      double v;
      removeFromRow(i2, pc, v);
      nnzActive--;
      double mult = v / pv;
      if (std::fabs(mult) <= zeroTolerance)
        continue;
      lcol.emplace_back(i2, mult);
      // row_i2 -= mult * (pivot row without pc)
      markStamp++;
      std::vector<int> &rc = rowCols[i2];
      std::vector<double> &rv = rowVals[i2];
      for (size_t f = 0; f < rc.size(); f++) {
        wmark[rc[f]] = markStamp;
        wval[rc[f]] = static_cast<double>(f);
      }
      for (size_t e = 1; e < urow.size(); e++) {
        int c2 = urow[e].first;
        double delta = -mult * urow[e].second;
        if (wmark[c2] == markStamp) {
          rv[static_cast<size_t>(wval[c2])] += delta;
        } else {
          rc.push_back(c2);
          rv.push_back(delta);
          colRows[c2].push_back(i2);
          nnzActive++;
        }
      }
      // drop cancellations
      for (size_t f = 0; f < rc.size();) {
        if (std::fabs(rv[f]) <= zeroTolerance) {
          removeFromCol(rc[f], i2);
          if (colRows[rc[f]].size() == 1)
            colStack.push_back(rc[f]);
          rc[f] = rc.back();
          rv[f] = rv.back();
          rc.pop_back();
          rv.pop_back();
          nnzActive--;
        } else
          f++;
      }
      if (rc.size() == 1)
        rowStack.push_back(i2);
    }
    for (size_t e = 1; e < urow.size(); e++)
      if (colRows[urow[e].first].size() == 1)
        colStack.push_back(urow[e].first);
    rowCols[pi].clear();
    rowVals[pi].clear();
    colRows[pc].clear();
    rowDone[pi] = 1;
    colDone[pc] = 1;
    pivRow.push_back(pi);
    pivCol.push_back(pc);
    Lsteps.push_back(std::move(lcol));
    Usteps.push_back(std::move(urow));
    nActive--;
    drainSingletons();
  }

  // ---- singular? ----
  const int nPiv = static_cast<int>(pivRow.size());
  rowOfColumn.assign(m, -1);
  for (int k = 0; k < nPiv; k++)
    rowOfColumn[pivCol[k]] = pivRow[k];
  if (nPiv < m) {
    for (int c = 0; c < m; c++)
      if (rowOfColumn[c] < 0)
        rejectedColumns.push_back(c);
    for (int i = 0; i < m; i++)
      if (!rowDone[i])
        unpivotedRows.push_back(i);
    return -1;
  }

  // ---- build position-space structures ----
  rowOfPos = pivRow;
  posOfRow.assign(m, -1);
  std::vector<int> posOfCol(m, -1);
  for (int k = 0; k < m; k++) {
    posOfRow[pivRow[k]] = k;
    posOfCol[pivCol[k]] = k;
  }
  Lstart.assign(m + 1, 0);
  Lidx.clear();
  Lval.clear();
  for (int k = 0; k < m; k++) {
    for (auto &pr : Lsteps[k]) {
      Lidx.push_back(posOfRow[pr.first]);
      Lval.push_back(pr.second);
    }
    Lstart[k + 1] = static_cast<long>(Lidx.size());
  }
  nnzL = static_cast<long>(Lidx.size());
  diag.assign(m, 0.0);
  std::vector<int> ucount(m, 0);
  for (int k = 0; k < m; k++)
    for (size_t e = 1; e < Usteps[k].size(); e++)
      ucount[posOfCol[Usteps[k][e].first]]++;
  Ustart.assign(m, 0);
  Ulen.assign(m, 0);
  long tot = 0;
  for (int k = 0; k < m; k++) {
    Ustart[k] = tot;
    tot += ucount[k];
  }
  Uidx.assign(tot, 0);
  Uval.assign(tot, 0.0);
  for (int k = 0; k < m; k++) {
    diag[k] = Usteps[k][0].second;
    for (size_t e = 1; e < Usteps[k].size(); e++) {
      int kc = posOfCol[Usteps[k][e].first];
      long at = Ustart[kc] + Ulen[kc]++;
      Uidx[at] = k;
      Uval[at] = Usteps[k][e].second;
    }
  }
  nnzU = tot + m;
  nnzR = 0;
  Ubirth.assign(m, 0);
  rowElim.assign(m, -1);
  order.resize(m);
  slotOfPos.resize(m);
  for (int k = 0; k < m; k++) {
    order[k] = k;
    slotOfPos[k] = k;
  }
  Rpos.clear();
  Rstart.assign(1, 0);
  Ridx.clear();
  Rval.clear();
  spike.assign(m, 0.0);
  btranU.assign(m, 0.0);
  work.assign(m, 0.0);
  return 0;
}

inline void Factorization::ftran(double *v, bool saveSpike)
{
  double *w = work.data();
  for (int k = 0; k < m; k++)
    w[k] = v[rowOfPos[k]];
  // L (updateColumnL, CoinAbcBaseFactorization3.cpp:279)
  for (int k = 0; k < m; k++) {
    double t = w[k];
    if (t != 0.0) {
      for (long e = Lstart[k]; e < Lstart[k + 1]; e++)
        w[Lidx[e]] -= Lval[e] * t;
    }
  }
  // R row-etas (updateColumnR, ...5.cpp:220)
  const int nR = static_cast<int>(Rpos.size());
  for (int t = 0; t < nR; t++) {
    double s = 0.0;
    for (long e = Rstart[t]; e < Rstart[t + 1]; e++)
      s += Rval[e] * w[Ridx[e]];
    w[Rpos[t]] -= s;
  }
  if (saveSpike) {
    std::memcpy(spike.data(), w, sizeof(double) * m);
    spikeValid = true;
  }
  // U (updateColumnU, ...3.cpp:1392) in pivot-list order, back to front
  for (int s = static_cast<int>(order.size()) - 1; s >= 0; s--) {
    int k = order[s];
    if (k < 0)
      continue;
    double x = w[k];
    if (x == 0.0)
      continue;
    x /= diag[k];
    if (std::fabs(x) < zeroTolerance) {
      w[k] = 0.0;
      continue;
    }
    w[k] = x;
    const int birth = Ubirth[k];
    const long st = Ustart[k];
    const int len = Ulen[k];
    for (int e = 0; e < len; e++) {
      int i = Uidx[st + e];
      if (birth > rowElim[i])
        w[i] -= Uval[st + e] * x;
    }
  }
  for (int k = 0; k < m; k++)
    v[rowOfPos[k]] = w[k];
}

inline void Factorization::btran(double *v, int unitRow)
{
  double *w = work.data();
  for (int k = 0; k < m; k++)
    w[k] = v[rowOfPos[k]];
  // U^T (updateColumnTransposeU, ...4.cpp:3362) front to back
  int s0 = 0;
  if (unitRow >= 0)
    s0 = slotOfPos[posOfRow[unitRow]];
  const int nslots = static_cast<int>(order.size());
  for (int s = s0; s < nslots; s++) {
    int k = order[s];
    if (k < 0)
      continue;
    double sum = w[k];
    const int birth = Ubirth[k];
    const long st = Ustart[k];
    const int len = Ulen[k];
    for (int e = 0; e < len; e++) {
      int i = Uidx[st + e];
      if (birth > rowElim[i])
        sum -= Uval[st + e] * w[i];
    }
    w[k] = sum / diag[k];
  }
  if (unitRow >= 0) {
    std::memcpy(btranU.data(), w, sizeof(double) * m);
    btranUPos = posOfRow[unitRow];
  }
  // R^T (updateColumnTransposeR, ...4.cpp:4217) newest first
  for (int t = static_cast<int>(Rpos.size()) - 1; t >= 0; t--) {
    double tp = w[Rpos[t]];
    if (tp != 0.0)
      for (long e = Rstart[t]; e < Rstart[t + 1]; e++)
        w[Ridx[e]] -= Rval[e] * tp;
  }
  // L^T (updateColumnTransposeL, ...4.cpp:3920) back to front
  for (int k = m - 1; k >= 0; k--) {
    double sum = w[k];
    for (long e = Lstart[k]; e < Lstart[k + 1]; e++)
      sum -= Lval[e] * w[Lidx[e]];
    w[k] = sum;
  }
  for (int k = 0; k < m; k++) {
    double x = w[k];
    v[rowOfPos[k]] = (std::fabs(x) < zeroTolerance) ? 0.0 : x;
  }
}

inline int Factorization::replaceColumn(int pivotRow, double alphaCheck)
{
  if (!spikeValid)
    return 2;
  if (numberPivots >= maximumPivots)
    return 5;
  const int p = posOfRow[pivotRow];
  if (btranUPos != p) {
    // recompute U^-T e_p (checkReplacePart1: row of U -> BTRAN-U)
    std::vector<double> e(m, 0.0);
    e[pivotRow] = 1.0;
    btran(e.data(), pivotRow);
  }
  const double *z = btranU.data();
  const double zp = z[p];
  if (zp == 0.0)
    return 2;
  // multipliers r_j = -z_j / z_p ; new diagonal = spike_p - sum r_j spike_j
  double newDiag = spike[p];
  std::vector<int> ri;
  std::vector<double> rv;
  for (int s = slotOfPos[p] + 1; s < static_cast<int>(order.size()); s++) {
    int j = order[s];
    if (j < 0)
      continue;
    double zj = z[j];
    if (zj != 0.0) {
      double r = -zj / zp;
      if (std::fabs(r) > zeroTolerance) {
        ri.push_back(j);
        rv.push_back(r);
        newDiag -= r * spike[j];
      }
    }
  }
  // checkPivot (CoinAbcBaseFactorization4.cpp:94): compare with alpha * old diagonal
  const double expected = alphaCheck * diag[p];
  int status = 0;
  if (std::fabs(newDiag) < singularTolerance)
    return 2;
  if (alphaCheck != 0.0) {
    double rel = std::fabs(newDiag - expected) / std::max(std::fabs(newDiag), std::fabs(expected));
    if (rel > 1.0e-6)
      status = 1;
    if (rel > 1.0e-2)
      return 2;
  }
  // install (replaceColumnPart3)
  stamp++;
  Rpos.push_back(p);
  for (size_t e = 0; e < ri.size(); e++) {
    Ridx.push_back(ri[e]);
    Rval.push_back(rv[e]);
  }
  Rstart.push_back(static_cast<long>(Ridx.size()));
  nnzR += static_cast<long>(ri.size());
  rowElim[p] = stamp;
  nnzU -= Ulen[p];
  Ustart[p] = static_cast<long>(Uidx.size());
  int len = 0;
  for (int i = 0; i < m; i++) {
    if (i == p)
      continue;
    double sv = spike[i];
    if (std::fabs(sv) > zeroTolerance) {
      Uidx.push_back(i);
      Uval.push_back(sv);
      len++;
    }
  }
  Ulen[p] = len;
  nnzU += len;
  diag[p] = newDiag;
  Ubirth[p] = stamp;
  order[slotOfPos[p]] = -1;
  slotOfPos[p] = static_cast<int>(order.size());
  order.push_back(p);
  numberPivots++;
  spikeValid = false;
  btranUPos = -1;
  return status;
}

} // namespace orc
