// clp_oracle.cpp -- CPU oracle: restatement of Clp's revised dual simplex (dual steepest edge,
// bound-flipping ratio test, Forrest-Tomlin LU update).
//
// TEST INFRASTRUCTURE ONLY (see clp_oracle.h): only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library.
//
// Each function names the reference function it follows (paths relative to /root/reference):
//   DualSimplex::dual              ClpSimplexDual::dual                src/ClpSimplexDual.cpp:637
//   DualSimplex::whileIterating    ClpSimplexDual::whileIterating      src/ClpSimplexDual.cpp:973
//   DualSimplex::statusOfProblem   ClpSimplexDual::statusOfProblemInDual   :4996
//   DualSimplex::pivotRow          ClpDualRowSteepest::pivotRow        src/ClpDualRowSteepest.cpp:179
//   DualSimplex::transposeTimes    ClpPackedMatrix::transposeTimes     src/ClpPackedMatrix.cpp:706
//                                  (by column :961/:1640, by row :1307)
//   dualColumn (free function)     ClpSimplexDual::dualColumn0/dualColumn  :3665/:4192
//   dseUpdate (free function)      ClpDualRowSteepest::updateWeights   src/ClpDualRowSteepest.cpp:375
//   DualSimplex::updateDualsInDual ClpSimplexDual::updateDualsInDual   :2430
//   DualSimplex::updatePrimal      ClpDualRowSteepest::updatePrimalSolution :630
//   DualSimplex::computePrimals/Duals  ClpSimplex::computePrimals/computeDuals src/ClpSimplex.cpp:914/1164
//   DualSimplex::changeBounds      ClpSimplexDual::changeBounds        :3148
//   defaultFactorizationFrequency  ClpSimplex::defaultFactorizationFrequency src/ClpSimplex.cpp:11401
//
// Conventions (ClpSimplex): variables 0..n-1 are columns, n..n+m-1 are row activities with
// column -e_i ("slacks as -1 singletons", ClpFactorization.cpp:2233-2240); [A | -I][x;y] = 0.
// Scaling, perturbation and presolve are OFF (stated in DESIGN.md); costs are shifted only when
// the ratio test needs it (ClpSimplexDual.cpp:4705-4772) and shifts are removed before the
// final optimality check.
#include "clp_oracle.h"
#include "factorization.hpp"

#include <chrono>
#include <cstdlib>
#include <limits>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

static const double kInf = 1.0e30;
static const double kDevexTryNorm = 1.0e-4; // DEVEX_TRY_NORM, ClpSimplex.hpp:2056

// ------------------------------------------------------------------ ratio test (BFRT)
// Candidates: alpha already multiplied by the leaving direction so that dj' = dj - theta*alpha.
// atLower candidates have alpha>0, atUpper alpha<0, free any sign.
int dualColumn(int count, const double *alpha, const double *dj, const double *range,
               const unsigned char *stat, double infeasibility, double dualTolerance,
               double acceptablePivot, double *thetaOut, unsigned char *passed)
{
  std::vector<int> remaining;
  remaining.reserve(count);
  for (int k = 0; k < count; k++) {
    if (passed)
      passed[k] = 0;
    double a = alpha[k];
    unsigned char s = stat[k];
    bool ok = false;
    if (s == ORC_atLowerBound)
      ok = a > 0.0;
    else if (s == ORC_atUpperBound)
      ok = a < 0.0;
    else if (s == ORC_isFree || s == ORC_superBasic)
      ok = a != 0.0;
    if (ok)
      remaining.push_back(k);
  }
  *thetaOut = 0.0;
  if (remaining.empty())
    return -1;
  double slope = infeasibility;
  int chosen = -1;
  int lastBest = -1; // best acceptable pivot among batches already passed (fallback)
  std::vector<int> batch, next;
  for (int pass = 0; pass < 1000 && !remaining.empty(); pass++) {
    // Harris bound over elements allowed to be pivots
    double thetaMax = 1.0e50;
    for (int k : remaining) {
      double a = std::fabs(alpha[k]);
      if (a < acceptablePivot)
        continue;
      unsigned char s = stat[k];
      double d = (s == ORC_atLowerBound) ? std::max(dj[k], 0.0)
                 : (s == ORC_atUpperBound) ? std::max(-dj[k], 0.0)
                                           : 0.0;
      double t = (d + dualTolerance) / a;
      if (t < thetaMax)
        thetaMax = t;
    }
    batch.clear();
    next.clear();
    double batchSlope = 0.0;
    int best = -1;
    double bestAbs = 0.0;
    for (int k : remaining) {
      double a = std::fabs(alpha[k]);
      unsigned char s = stat[k];
      double d = (s == ORC_atLowerBound) ? std::max(dj[k], 0.0)
                 : (s == ORC_atUpperBound) ? std::max(-dj[k], 0.0)
                                           : 0.0;
      if (d <= thetaMax * a) {
        batch.push_back(k);
        bool boxed = range[k] < 1.0e29 && (s == ORC_atLowerBound || s == ORC_atUpperBound);
        if (a >= acceptablePivot) {
          if (boxed)
            batchSlope += a * range[k];
          else
            batchSlope = 1.0e100;
          if (a > bestAbs) {
            bestAbs = a;
            best = k;
          }
        } else if (boxed) {
          batchSlope += a * range[k];
        }
      } else {
        next.push_back(k);
      }
    }
    if (batch.empty())
      break; // only tiny pivots left
    if (best >= 0 && (slope - batchSlope < 0.0 || next.empty())) {
      chosen = best;
      break;
    }
    if (best < 0 && next.empty())
      break;
    // pass the whole batch (they will flip)
    slope -= batchSlope;
    if (best >= 0)
      lastBest = best;
    if (passed)
      for (int k : batch)
        passed[k] = 1;
    remaining.swap(next);
    if (slope < 0.0) {
      chosen = lastBest;
      break;
    }
  }
  if (chosen < 0)
    chosen = lastBest;
  if (chosen < 0)
    return -1;
  if (passed)
    passed[chosen] = 0;
  double d = dj[chosen];
  double t = d / alpha[chosen];
  *thetaOut = t > 0.0 ? t : 0.0;
  return chosen;
}

// ------------------------------------------------------------------ bucketed BFRT
// Restatement of the GPU's sort-free variant of the same ratio test (clp_b200/csrc/price.cu):
// slope contributions are accumulated per ratio bucket (top 15 bits of the double), the step
// stops at the smallest ratio of the bucket in which the slope is exhausted, then Harris +
// largest |alpha|.  Lets tests follow the GPU's pivot sequence on the CPU.
int dualColumnBucketed(int count, const double *alpha, const double *dj, const double *range,
                       const unsigned char *stat, double infeasibility, double dualTolerance,
                       double acceptablePivot, double *thetaOut, unsigned char *passed)
{
  const int NB = 32768;
  static thread_local std::vector<unsigned long long> hw, hm;
  hw.assign(NB, 0ull);
  hm.assign(NB, ~0ull);
  std::vector<double> A(count), D(count), R(count);
  std::vector<char> ok(count, 0);
  bool any = false;
  for (int k = 0; k < count; k++) {
    if (passed)
      passed[k] = 0;
    double ab = alpha[k], a = std::fabs(ab);
    unsigned char s = stat[k];
    if (a <= 1.0e-12)
      continue;
    double dt;
    bool boxed = false;
    if (s == ORC_atLowerBound) {
      if (ab <= 0.0)
        continue;
      dt = dj[k] > 0.0 ? dj[k] : 0.0;
      boxed = range[k] < 1.0e29;
    } else if (s == ORC_atUpperBound) {
      if (ab >= 0.0)
        continue;
      dt = dj[k] < 0.0 ? -dj[k] : 0.0;
      boxed = range[k] < 1.0e29;
    } else if (s == ORC_isFree || s == ORC_superBasic) {
      dt = 0.0;
    } else
      continue;
    ok[k] = 1;
    any = true;
    A[k] = a;
    D[k] = dt;
    double ratio = dt / a;
    R[k] = ratio;
    unsigned long long bits;
    std::memcpy(&bits, &ratio, 8);
    int b = (int)(bits >> 48) & (NB - 1);
    unsigned long long w = 1ull << 41;
    if (boxed) {
      double v = a * range[k] / infeasibility * 1099511627776.0;
      w = v >= 2199023255552.0 ? (1ull << 41) : (unsigned long long)v;
      if (w == 0ull)
        w = 1ull;
    }
    hw[b] += w;
    hm[b] = std::min(hm[b], bits);
  }
  *thetaOut = 0.0;
  if (!any)
    return -1;
  unsigned long long c = 0, before = 0;
  int cross = -1, last = -1;
  for (int b = 0; b < NB; b++) {
    if (hm[b] != ~0ull)
      last = b;
    if (cross < 0 && c + hw[b] >= (1ull << 40)) {
      cross = b;
      before = c;
    }
    c += hw[b];
  }
  double thetaStar;
  bool neverExhausted = false;
  if (cross < 0) { // slope never exhausted: last break point group (last sub-bucket of the last bucket)
    cross = last;
    before = 0;
    neverExhausted = true;
  }
  {
    // second level: the next 12 bits of the ratio inside the crossing bucket
    const int NB2 = 4096;
    std::vector<unsigned long long> hw2(NB2, 0ull), hm2(NB2, ~0ull);
    for (int k = 0; k < count; k++) {
      if (!ok[k])
        continue;
      unsigned long long bits;
      std::memcpy(&bits, &R[k], 8);
      if (((int)(bits >> 48) & (NB - 1)) != cross)
        continue;
      int sb = (int)(bits >> 36) & (NB2 - 1);
      bool boxed = range[k] < 1.0e29 && (stat[k] == ORC_atLowerBound || stat[k] == ORC_atUpperBound);
      unsigned long long w = 1ull << 41;
      if (boxed) {
        double v = A[k] * range[k] / infeasibility * 1099511627776.0;
        w = v >= 2199023255552.0 ? (1ull << 41) : (unsigned long long)v;
        if (w == 0ull)
          w = 1ull;
      }
      hw2[sb] += w;
      hm2[sb] = std::min(hm2[sb], bits);
    }
    const unsigned long long resid = neverExhausted ? ~0ull : (1ull << 40) - before;
    unsigned long long c2 = 0;
    int cross2 = -1, last2 = -1;
    for (int b = 0; b < NB2; b++) {
      if (hm2[b] != ~0ull)
        last2 = b;
      c2 += hw2[b];
      if (cross2 < 0 && c2 >= resid)
        cross2 = b;
    }
    if (cross2 < 0)
      cross2 = last2;
    std::memcpy(&thetaStar, &hm2[cross2], 8);
  }
  double harris = std::numeric_limits<double>::infinity();
  for (int k = 0; k < count; k++)
    if (ok[k] && A[k] >= acceptablePivot && R[k] >= thetaStar)
      harris = std::min(harris, (D[k] + dualTolerance) / A[k]);
  int best = -1;
  unsigned long long bestKey = 0;
  for (int k = 0; k < count; k++)
    if (ok[k] && A[k] >= acceptablePivot && R[k] >= thetaStar && R[k] <= harris) {
      unsigned long long bits;
      std::memcpy(&bits, &A[k], 8);
      unsigned long long key = (bits & ~0xFFFFFull) | (unsigned long long)(0xFFFFF - k);
      if (key > bestKey) {
        bestKey = key;
        best = k;
      }
    }
  if (best < 0)
    return -1;
  if (passed)
    for (int k = 0; k < count; k++)
      if (ok[k] && R[k] < thetaStar)
        passed[k] = 1;
  double t = dj[best] / alpha[best];
  *thetaOut = t > 0.0 ? t : 0.0;
  return best;
}

// ------------------------------------------------------------------ DSE recurrence
void dseUpdate(int m, double *weights, const double *alphaColumn, const double *tau, int pivotRow,
               double rhoNorm2)
{
  const double alphaR = alphaColumn[pivotRow];
  const double wr = rhoNorm2;
  for (int i = 0; i < m; i++) {
    double a = alphaColumn[i];
    if (a == 0.0 || i == pivotRow)
      continue;
    double ratio = a / alphaR;
    double w = weights[i] + ratio * (ratio * wr - 2.0 * tau[i]);
    weights[i] = w < kDevexTryNorm ? kDevexTryNorm : w;
  }
  double w = wr / (alphaR * alphaR);
  weights[pivotRow] = w < kDevexTryNorm ? kDevexTryNorm : w;
}

// ------------------------------------------------------------------ the model
struct DualSimplex {
  int m = 0, n = 0;
  // column copy (CSC) and row copy (CSR)
  std::vector<long> colStart;
  std::vector<int> rowIdx;
  std::vector<double> elem;
  std::vector<long> rowStart;
  std::vector<int> colIdx;
  std::vector<double> relem;
  // rim (n+m)
  std::vector<double> cost, costTrue, lower, upper, lowerTrue, upperTrue, sol, dj;
  std::vector<unsigned char> status;
  std::vector<unsigned char> fake; // 1 lower is fake, 2 upper is fake
  std::vector<int> pivotVariable;
  std::vector<double> weights;
  Factorization fac;
  // options
  double primalTolerance = 1e-7, dualTolerance = 1e-7, dualBound = 1e10;
  double acceptablePivot = 1e-7;
  int maximumIterations = 2147483647;
  int factorizationFrequency = 0;
  int logLevel = 0;
  int threads = 1;
  double maximumSeconds = 1e30;
  int bucketedRatioTest = 0; // 1 = mimic the GPU's sort-free ratio test
  int warmupIterations = 0;  // timed window starts after this many iterations
  double timedSeconds = 0.0;
  int timedIterations = 0;
  // results
  int problemStatus = -1;
  int numberIterations = 0, numberRefactorizations = 0;
  double objectiveValue = 0.0;
  double secondsInLoop = 0.0;
  bool haveUserStatus = false;
  bool costsShifted = false;
  int numberFake = 0;
  // work
  std::vector<double> rho, alphaRow, alphaCol, tau, flipRhs, piWork;

  int defaultFactorizationFrequency() const
  {
    // ClpSimplex.cpp:11401-11431
    const int cutoff1 = 10000, cutoff2 = 100000, base = 75, freq0 = 50, freq1 = 150,
              maximum = 10000;
    int frequency;
    if (m < cutoff1)
      frequency = base + m / freq0;
    else
      frequency = base + cutoff1 / freq0 + (m - cutoff1) / freq1;
    (void)cutoff2;
    return std::min(maximum, frequency);
  }

  void buildRowCopy()
  {
    rowStart.assign(m + 1, 0);
    for (long e = 0; e < colStart[n]; e++)
      rowStart[rowIdx[e] + 1]++;
    for (int i = 0; i < m; i++)
      rowStart[i + 1] += rowStart[i];
    colIdx.resize(colStart[n]);
    relem.resize(colStart[n]);
    std::vector<long> fill(rowStart.begin(), rowStart.end() - 1);
    for (int j = 0; j < n; j++)
      for (long e = colStart[j]; e < colStart[j + 1]; e++) {
        long at = fill[rowIdx[e]]++;
        colIdx[at] = j;
        relem[at] = elem[e];
      }
  }

  // z[n] = scalar * A^T pi   (by column; gutsOfTransposeTimesUnscaled ClpPackedMatrix.cpp:1484)
  void transposeTimes(double scalar, const double *pi, double *z) const
  {
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int j = 0; j < n; j++) {
      double s = 0.0;
      for (long e = colStart[j]; e < colStart[j + 1]; e++)
        s += pi[rowIdx[e]] * elem[e];
      z[j] = scalar * s;
    }
  }
  // nonbasic only, choosing by-row when pi is sparse (transposeTimesByRow :1307)
  void priceRow(const double *pi, double *z) const
  {
    int nz = 0;
    for (int i = 0; i < m; i++)
      if (pi[i] != 0.0)
        nz++;
    if (nz * 3 < m && threads == 1) {
      std::fill(z, z + n, 0.0);
      for (int i = 0; i < m; i++) {
        double p = pi[i];
        if (p == 0.0)
          continue;
        for (long e = rowStart[i]; e < rowStart[i + 1]; e++)
          z[colIdx[e]] += p * relem[e];
      }
      for (int j = 0; j < n; j++)
        if (status[j] == ORC_basic)
          z[j] = 0.0;
    } else {
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
      for (int j = 0; j < n; j++) {
        if (status[j] == ORC_basic) {
          z[j] = 0.0;
          continue;
        }
        double s = 0.0;
        for (long e = colStart[j]; e < colStart[j + 1]; e++)
          s += pi[rowIdx[e]] * elem[e];
        z[j] = s;
      }
    }
  }
  // y += scalar * A x (ClpPackedMatrix::times :296)
  void times(double scalar, const double *x, double *y) const
  {
    for (int j = 0; j < n; j++) {
      double v = scalar * x[j];
      if (v != 0.0)
        for (long e = colStart[j]; e < colStart[j + 1]; e++)
          y[rowIdx[e]] += v * elem[e];
    }
  }
  // unpack column of [A|-I] into dense m-vector (ClpPackedMatrix::unpack :4803)
  void unpack(int seq, double *v) const
  {
    std::fill(v, v + m, 0.0);
    if (seq < n) {
      for (long e = colStart[seq]; e < colStart[seq + 1]; e++)
        v[rowIdx[e]] = elem[e];
    } else
      v[seq - n] = -1.0;
  }

  // ClpFactorization::factorize: gather basis, factor, permute pivotVariable_, repair if singular
  int factorize()
  {
    for (int attempt = 0; attempt < 10; attempt++) {
      std::vector<long> bs(m + 1, 0);
      std::vector<int> br;
      std::vector<double> bv;
      for (int c = 0; c < m; c++) {
        int seq = pivotVariable[c];
        if (seq < n) {
          for (long e = colStart[seq]; e < colStart[seq + 1]; e++) {
            br.push_back(rowIdx[e]);
            bv.push_back(elem[e]);
          }
        } else {
          br.push_back(seq - n);
          bv.push_back(-1.0);
        }
        bs[c + 1] = static_cast<long>(br.size());
      }
      std::vector<int> rowOfColumn;
      fac.threads = threads;
      int rc = fac.factorize(m, bs, br, bv, rowOfColumn);
      numberRefactorizations++;
      if (rc == 0) {
        std::vector<int> pv(m);
        std::vector<double> w(m);
        for (int c = 0; c < m; c++) {
          pv[rowOfColumn[c]] = pivotVariable[c];
          w[rowOfColumn[c]] = weights[c];
        }
        pivotVariable.swap(pv);
        weights.swap(w);
        return 0;
      }
      // singular: throw rejected variables out, bring slacks of unpivoted rows in
      // (ClpFactorization.cpp:2382-2532)
      if (logLevel > 0)
        fprintf(stderr, "oracle: singular basis, %d rejected\n",
                static_cast<int>(fac.rejectedColumns.size()));
      size_t k = 0;
      for (int c : fac.rejectedColumns) {
        int seq = pivotVariable[c];
        int row = fac.unpivotedRows[k++];
        // variable leaves at nearest bound
        setNonbasic(seq);
        pivotVariable[c] = n + row;
        status[n + row] = ORC_basic;
        weights[c] = 1.0;
      }
    }
    return -1;
  }
  void setNonbasic(int seq)
  {
    double lo = lower[seq], up = upper[seq], v = sol[seq];
    if (lo > -kInf && up < kInf) {
      if (lo == up) {
        status[seq] = ORC_isFixed;
        sol[seq] = lo;
      } else if (std::fabs(v - lo) <= std::fabs(v - up)) {
        status[seq] = ORC_atLowerBound;
        sol[seq] = lo;
      } else {
        status[seq] = ORC_atUpperBound;
        sol[seq] = up;
      }
    } else if (lo > -kInf) {
      status[seq] = ORC_atLowerBound;
      sol[seq] = lo;
    } else if (up < kInf) {
      status[seq] = ORC_atUpperBound;
      sol[seq] = up;
    } else {
      status[seq] = ORC_isFree;
      sol[seq] = 0.0;
    }
  }

  // ClpSimplex::computePrimals : x_B = B^-1 ( -N x_N ), one step of refinement
  void computePrimals()
  {
    std::vector<double> rhs(m, 0.0), xn(n);
    for (int j = 0; j < n; j++)
      xn[j] = status[j] == ORC_basic ? 0.0 : sol[j];
    times(-1.0, xn.data(), rhs.data());
    for (int i = 0; i < m; i++)
      if (status[n + i] != ORC_basic)
        rhs[i] += sol[n + i];
    std::vector<double> x(rhs);
    fac.ftran(x.data());
    // refinement: r = rhs - B x
    std::vector<double> r(rhs);
    for (int p = 0; p < m; p++) {
      double v = x[p];
      if (v == 0.0)
        continue;
      int seq = pivotVariable[p];
      if (seq < n) {
        for (long e = colStart[seq]; e < colStart[seq + 1]; e++)
          r[rowIdx[e]] -= v * elem[e];
      } else
        r[seq - n] += v;
    }
    fac.ftran(r.data());
    for (int p = 0; p < m; p++)
      sol[pivotVariable[p]] = x[p] + r[p];
  }
  // ClpSimplex::computeDuals : pi = B^-T c_B ; dj = c - A^T pi ; row dj = pi
  void computeDuals()
  {
    std::vector<double> pi(m);
    for (int p = 0; p < m; p++)
      pi[p] = cost[pivotVariable[p]];
    fac.btran(pi.data());
    piWork = pi;
    std::vector<double> z(n);
    transposeTimes(1.0, pi.data(), z.data());
    for (int j = 0; j < n; j++)
      dj[j] = status[j] == ORC_basic ? 0.0 : cost[j] - z[j];
    for (int i = 0; i < m; i++)
      dj[n + i] = status[n + i] == ORC_basic ? 0.0 : cost[n + i] + pi[i];
  }

  // Make every nonbasic variable dual feasible by choosing its bound; put a fake bound of
  // width dualBound where the needed bound is infinite (ClpSimplexDual::changeBounds :3148).
  // Returns number of variables whose value moved.
  int makeDualFeasible()
  {
    int moved = 0;
    numberFake = 0;
    for (int j = 0; j < n + m; j++) {
      if (status[j] == ORC_basic) {
        // basic variables always carry their true bounds (originalBound :1828)
        lower[j] = lowerTrue[j];
        upper[j] = upperTrue[j];
        fake[j] = 0;
        continue;
      }
      double lo = lowerTrue[j], up = upperTrue[j];
      double d = dj[j];
      double old = sol[j];
      unsigned char f = 0;
      if (lo == up) {
        status[j] = ORC_isFixed;
        sol[j] = lo;
      } else if (d > dualTolerance) {
        // wants lower bound
        if (lo <= -kInf) {
          lo = (up < kInf ? up : 0.0) - dualBound;
          f = 1;
        }
        status[j] = ORC_atLowerBound;
        sol[j] = lo;
      } else if (d < -dualTolerance) {
        if (up >= kInf) {
          up = (lo > -kInf ? lo : 0.0) + dualBound;
          f = 2;
        }
        status[j] = ORC_atUpperBound;
        sol[j] = up;
      } else {
        // dual degenerate: keep current side if it is a real bound
        if (status[j] == ORC_atUpperBound && up < kInf && !(fake[j] & 2)) {
          sol[j] = up;
        } else if (lo > -kInf) {
          status[j] = ORC_atLowerBound;
          sol[j] = lo;
        } else if (up < kInf) {
          status[j] = ORC_atUpperBound;
          sol[j] = up;
        } else {
          status[j] = ORC_isFree;
          sol[j] = 0.0;
        }
      }
      lower[j] = lo;
      upper[j] = up;
      fake[j] = f;
      if (f)
        numberFake++;
      if (sol[j] != old)
        moved++;
    }
    return moved;
  }

  // ClpDualRowSteepest::pivotRow : argmax infeas^2 / weight
  int pivotRow(double &infeasOut, int &directionOut) const
  {
    int best = -1;
    double bestScore = 0.0;
    const double tol = primalTolerance;
    for (int p = 0; p < m; p++) {
      int seq = pivotVariable[p];
      double v = sol[seq];
      double inf = 0.0;
      if (v < lower[seq] - tol)
        inf = lower[seq] - v;
      else if (v > upper[seq] + tol)
        inf = v - upper[seq];
      else
        continue;
      double score = inf * inf / weights[p];
      if (score > bestScore) {
        bestScore = score;
        best = p;
      }
    }
    if (best >= 0) {
      int seq = pivotVariable[best];
      double v = sol[seq];
      if (v < lower[seq]) {
        infeasOut = lower[seq] - v;
        directionOut = -1; // leaves to lower bound: sigma = -1
      } else {
        infeasOut = v - upper[seq];
        directionOut = +1; // leaves to upper bound: sigma = +1
      }
    }
    return best;
  }

  double computeObjective() const
  {
    double s = 0.0;
    for (int j = 0; j < n + m; j++)
      s += costTrue[j] * sol[j];
    return s;
  }

  // refactorize and recompute everything (statusOfProblemInDual :4996 + gutsOfSolution)
  int refresh()
  {
    if (factorize() != 0)
      return -1;
    computeDuals();
    int moved = makeDualFeasible();
    (void)moved;
    computePrimals();
    return 0;
  }

  int dual();
};

int DualSimplex::dual()
{
  auto t0 = std::chrono::steady_clock::now();
  const int nm = n + m;
  numberIterations = 0;
  numberRefactorizations = 0;
  problemStatus = -1;
  cost = costTrue;
  lower = lowerTrue;
  upper = upperTrue;
  fake.assign(nm, 0);
  dj.assign(nm, 0.0);
  costsShifted = false;
  rho.assign(m, 0.0);
  alphaRow.assign(nm, 0.0);
  alphaCol.assign(m, 0.0);
  tau.assign(m, 0.0);
  flipRhs.assign(m, 0.0);
  weights.assign(m, 1.0);
  pivotVariable.clear();
  if (!haveUserStatus) {
    status.assign(nm, ORC_atLowerBound);
    for (int i = 0; i < m; i++)
      status[n + i] = ORC_basic;
    sol.assign(nm, 0.0);
  }
  {
    std::vector<int> basics;
    for (int j = 0; j < nm; j++)
      if (status[j] == ORC_basic)
        basics.push_back(j);
    if (static_cast<int>(basics.size()) != m) {
      // bad user basis: fall back to all slack
      status.assign(nm, ORC_atLowerBound);
      basics.clear();
      for (int i = 0; i < m; i++) {
        status[n + i] = ORC_basic;
        basics.push_back(n + i);
      }
    }
    pivotVariable = basics;
    for (int j = 0; j < nm; j++)
      if (status[j] != ORC_basic) {
        // honour a user-supplied side (Clp_copyinStatus) when that bound exists
        if (haveUserStatus && status[j] == ORC_atUpperBound && upper[j] < kInf)
          sol[j] = upper[j];
        else if (haveUserStatus && status[j] == ORC_atLowerBound && lower[j] > -kInf)
          sol[j] = lower[j];
        setNonbasic(j);
      }
  }
  fac.maximumPivots = factorizationFrequency > 0 ? factorizationFrequency
                                                 : defaultFactorizationFrequency();
  if (refresh() != 0) {
    problemStatus = 4;
    return problemStatus;
  }
  int dualBoundIncreases = 0;
  bool needRefresh = false;
  int consecutiveBad = 0;
  std::vector<int> candIdx;
  std::vector<double> candAlpha, candDj, candRange;
  std::vector<unsigned char> candStat, candPassed;
  std::vector<int> flipList;

  bool windowOpen = false;
  auto tWindow = t0;
  int windowStart = 0;
  while (problemStatus < 0) {
    if (!windowOpen && numberIterations >= warmupIterations) {
      windowOpen = true;
      tWindow = std::chrono::steady_clock::now();
      windowStart = numberIterations;
    }
    if (numberIterations >= maximumIterations) {
      problemStatus = 3;
      break;
    }
    if ((numberIterations & 63) == 0) {
      double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (el > maximumSeconds) {
        problemStatus = 3;
        break;
      }
    }
    if (needRefresh || fac.numberPivots >= fac.maximumPivots) {
      if (refresh() != 0) {
        problemStatus = 4;
        break;
      }
      needRefresh = false;
    }
    // ---- CHUZR
    double infeas = 0.0;
    int sigma = 0;
    int r = pivotRow(infeas, sigma);
    if (r < 0) {
      if (fac.numberPivots > 0) {
        needRefresh = true;
        continue;
      }
      // primal feasible on a fresh factorization
      if (costsShifted) {
        cost = costTrue;
        costsShifted = false;
        computeDuals();
        makeDualFeasible();
        computePrimals();
        continue;
      }
      // any nonbasic sitting on a fake bound?
      int atFake = 0;
      for (int j = 0; j < nm; j++)
        if (status[j] != ORC_basic && fake[j])
          atFake++;
      if (atFake) {
        if (dualBoundIncreases < 2) {
          dualBoundIncreases++;
          dualBound *= 1000.0;
          makeDualFeasible();
          computePrimals();
          continue;
        }
        problemStatus = 2; // dual infeasible / unbounded
        break;
      }
      problemStatus = 0;
      break;
    }
    const int seqOut = pivotVariable[r];
    // ---- BTRAN  rho = B^-T e_r   (ClpSimplexDual.cpp:1286-1288)
    std::fill(rho.begin(), rho.end(), 0.0);
    rho[r] = 1.0;
    fac.btran(rho.data(), r);
    double rhoNorm2 = 0.0;
    for (int i = 0; i < m; i++)
      rhoNorm2 += rho[i] * rho[i];
    // ---- PRICE  alpha_j = rho^T a_j (ClpSimplexDual.cpp:1300)
    priceRow(rho.data(), alphaRow.data());
    for (int i = 0; i < m; i++)
      alphaRow[n + i] = status[n + i] == ORC_basic ? 0.0 : -rho[i];
    // ---- CHUZC (dualColumn :4192)
    candIdx.clear();
    candAlpha.clear();
    candDj.clear();
    candRange.clear();
    candStat.clear();
    for (int j = 0; j < nm; j++) {
      double a = alphaRow[j];
      if (std::fabs(a) <= 1.0e-12)
        continue;
      unsigned char s = status[j];
      if (s == ORC_basic || s == ORC_isFixed)
        continue;
      candIdx.push_back(j);
      candAlpha.push_back(sigma * a);
      candDj.push_back(dj[j]);
      candRange.push_back(upper[j] - lower[j]);
      candStat.push_back(s);
    }
    candPassed.assign(candIdx.size(), 0);
    double thetaDual = 0.0;
    int kq = (bucketedRatioTest ? dualColumnBucketed : dualColumn)(
        static_cast<int>(candIdx.size()), candAlpha.data(), candDj.data(), candRange.data(),
        candStat.data(), infeas, dualTolerance, acceptablePivot, &thetaDual, candPassed.data());
    if (logLevel > 3 && kq >= 0) {
      double t2 = 0.0;
      std::vector<unsigned char> p2(candIdx.size());
      int k2 = (bucketedRatioTest ? dualColumn : dualColumnBucketed)(
          static_cast<int>(candIdx.size()), candAlpha.data(), candDj.data(), candRange.data(),
          candStat.data(), infeas, dualTolerance, acceptablePivot, &t2, p2.data());
      int np1 = 0, np2 = 0;
      for (size_t k = 0; k < candIdx.size(); k++) {
        np1 += candPassed[k];
        np2 += p2[k];
      }
      fprintf(stderr, "CMP it=%d infeas=%.6g used: k=%d theta=%.8g passed=%d alpha=%.4g | other: k=%d theta=%.8g passed=%d alpha=%.4g\n",
              numberIterations, infeas, kq, thetaDual, np1, candAlpha[kq], k2, t2, np2, k2 >= 0 ? candAlpha[k2] : 0.0);
    }
    if (kq < 0) {
      if (fac.numberPivots > 0) {
        needRefresh = true;
        continue;
      }
      // no entering variable on a fresh factorization: dual unbounded => primal infeasible,
      // unless fake bounds took part in this row
      bool fakeInvolved = false;
      for (size_t k = 0; k < candIdx.size(); k++)
        if (fake[candIdx[k]])
          fakeInvolved = true;
      for (int j = 0; j < nm && !fakeInvolved; j++)
        if (status[j] != ORC_basic && fake[j] && std::fabs(alphaRow[j]) > 1e-9)
          fakeInvolved = true;
      if (fakeInvolved && dualBoundIncreases < 2) {
        dualBoundIncreases++;
        dualBound *= 1000.0;
        makeDualFeasible();
        computePrimals();
        continue;
      }
      problemStatus = 1;
      break;
    }
    const int seqIn = candIdx[kq];
    const double alphaBtran = alphaRow[seqIn];
    // ---- FTRAN (with FT spike) of entering column and of rho (updateWeights :375)
    unpack(seqIn, alphaCol.data());
    fac.ftran(alphaCol.data(), true);
    const double alphaFtran = alphaCol[r];
    // accuracy gate (ClpSimplexDual.cpp:1447-1501)
    {
      double err = std::fabs(alphaBtran - alphaFtran) / (1.0 + std::fabs(alphaFtran));
      if (err > 1.0e-6 || std::fabs(alphaFtran) < 1.0e-9) {
        consecutiveBad++;
        if (fac.numberPivots > 0) {
          needRefresh = true;
          continue;
        }
        if (consecutiveBad > 3 && logLevel > 0)
          fprintf(stderr, "oracle: inaccurate pivot accepted it=%d err=%g\n", numberIterations, err);
        if (std::fabs(alphaFtran) < 1.0e-11) {
          problemStatus = 4;
          break;
        }
      } else {
        consecutiveBad = 0;
        if (err > 1.0e-9 && fac.numberPivots > 20)
          needRefresh = true; // refactor after this pivot
      }
    }
    tau = rho;
    fac.ftran(tau.data(), false);
    // ---- dual update + flips (updateDualsInDual :2430)
    flipList.clear();
    std::fill(flipRhs.begin(), flipRhs.end(), 0.0);
    bool anyFlip = false;
    for (size_t k = 0; k < candIdx.size(); k++) {
      int j = candIdx[k];
      if (j == seqIn)
        continue;
      double dnew = dj[j] - thetaDual * candAlpha[k];
      unsigned char s = status[j];
      bool flip = false;
      if (s == ORC_atLowerBound && dnew < -dualTolerance)
        flip = true;
      else if (s == ORC_atUpperBound && dnew > dualTolerance)
        flip = true;
      if (flip) {
        if (upper[j] - lower[j] < 1.0e29) {
          double delta = (s == ORC_atLowerBound) ? (upper[j] - lower[j]) : (lower[j] - upper[j]);
          status[j] = (s == ORC_atLowerBound) ? ORC_atUpperBound : ORC_atLowerBound;
          sol[j] += delta;
          // rhs of x_B changes by -a_j * delta
          if (j < n) {
            for (long e = colStart[j]; e < colStart[j + 1]; e++)
              flipRhs[rowIdx[e]] -= delta * elem[e];
          } else
            flipRhs[j - n] += delta;
          anyFlip = true;
        } else {
          // cannot flip: shift cost so that dj becomes exactly zero (:4705-4772)
          cost[j] -= dnew;
          dnew = 0.0;
          costsShifted = true;
        }
      } else if ((s == ORC_isFree || s == ORC_superBasic) && std::fabs(dnew) > dualTolerance) {
        cost[j] -= dnew;
        dnew = 0.0;
        costsShifted = true;
      }
      dj[j] = dnew;
    }
    // also non-candidate entries of the row (wrong-sign alpha) just move
    // (they were included in candIdx already: candIdx holds every nonzero of the row)
    if (anyFlip) {
      fac.ftran(flipRhs.data(), false); // third FTRAN (ClpSimplexDual.cpp:1533-1537)
      for (int p = 0; p < m; p++)
        if (flipRhs[p] != 0.0)
          sol[pivotVariable[p]] += flipRhs[p];
    }
    // ---- DSE weights (updateWeights :501-538)
    dseUpdate(m, weights.data(), alphaCol.data(), tau.data(), r, rhoNorm2);
    // ---- primal step
    double valueOut = sol[seqOut];
    double boundOut = sigma < 0 ? lower[seqOut] : upper[seqOut];
    double thetaPrimal = (valueOut - boundOut) / alphaFtran;
    for (int p = 0; p < m; p++) {
      double a = alphaCol[p];
      if (a != 0.0)
        sol[pivotVariable[p]] -= thetaPrimal * a;
    }
    sol[seqIn] += thetaPrimal;
    sol[seqOut] = boundOut;
    // ---- FT update (replaceColumn :1599)
    int rc = fac.replaceColumn(r, alphaFtran);
    if (rc == 2 || rc == 3) {
      // singular / no room: undo is messy; refactorize from the new basis instead
      needRefresh = true;
    } else if (rc == 1 || rc == 5) {
      needRefresh = true;
    }
    // ---- housekeeping (ClpSimplex::housekeeping :2065)
    dj[seqIn] = 0.0;
    dj[seqOut] = -sigma * thetaDual;
    status[seqIn] = ORC_basic;
    lower[seqIn] = lowerTrue[seqIn];
    upper[seqIn] = upperTrue[seqIn];
    fake[seqIn] = 0;
    if (lower[seqOut] == upper[seqOut])
      status[seqOut] = ORC_isFixed;
    else
      status[seqOut] = sigma < 0 ? ORC_atLowerBound : ORC_atUpperBound;
    pivotVariable[r] = seqIn;
    if (logLevel > 2) {
      int nfl = 0;
      for (size_t k = 0; k < candIdx.size(); k++)
        nfl += 0;
      fprintf(stderr, "TRACE %d out=%d in=%d sigma=%d thetaD=%.12g thetaP=%.12g alpha=%.12g infeas=%.12g\n",
              numberIterations, seqOut, seqIn, sigma, thetaDual, thetaPrimal, alphaFtran, infeas);
    }
    numberIterations++;
    if (logLevel > 1 && (numberIterations % 100) == 0)
      fprintf(stderr, "oracle: it %d obj %.10g infeas %g theta %g\n", numberIterations,
              computeObjective(), infeas, thetaDual);
  }
  if (windowOpen) {
    timedSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tWindow).count();
    timedIterations = numberIterations - windowStart;
  }
  objectiveValue = computeObjective();
  secondsInLoop = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return problemStatus;
}

} // namespace orc

// ======================================================================= C API
struct orc_model {
  orc::DualSimplex s;
};

extern "C" {

orc_model *orc_create(int m, int n, const int *columnStart, const int *row, const double *element,
                      const double *columnLower, const double *columnUpper, const double *objective,
                      const double *rowLower, const double *rowUpper)
{
  orc_model *h = new orc_model();
  orc::DualSimplex &s = h->s;
  s.m = m;
  s.n = n;
  s.colStart.assign(columnStart, columnStart + n + 1);
  long nnz = columnStart[n];
  s.rowIdx.assign(row, row + nnz);
  s.elem.assign(element, element + nnz);
  s.buildRowCopy();
  const int nm = n + m;
  s.costTrue.assign(nm, 0.0);
  s.lowerTrue.assign(nm, 0.0);
  s.upperTrue.assign(nm, 0.0);
  auto clampLo = [](double v) { return v < -1.0e29 ? -orc::kInf : v; };
  auto clampUp = [](double v) { return v > 1.0e29 ? orc::kInf : v; };
  for (int j = 0; j < n; j++) {
    s.costTrue[j] = objective ? objective[j] : 0.0;
    s.lowerTrue[j] = columnLower ? clampLo(columnLower[j]) : 0.0;
    s.upperTrue[j] = columnUpper ? clampUp(columnUpper[j]) : orc::kInf;
  }
  for (int i = 0; i < m; i++) {
    s.lowerTrue[n + i] = rowLower ? clampLo(rowLower[i]) : -orc::kInf;
    s.upperTrue[n + i] = rowUpper ? clampUp(rowUpper[i]) : orc::kInf;
  }
  s.sol.assign(nm, 0.0);
  s.dj.assign(nm, 0.0);
  s.status.assign(nm, ORC_atLowerBound);
  s.weights.assign(m, 1.0);
  s.cost = s.costTrue;
  s.lower = s.lowerTrue;
  s.upper = s.upperTrue;
  s.fake.assign(nm, 0);
  return h;
}
void orc_destroy(orc_model *h) { delete h; }

void orc_set_option(orc_model *h, const char *key, double value)
{
  std::string k(key);
  orc::DualSimplex &s = h->s;
  if (k == "primalTolerance")
    s.primalTolerance = value;
  else if (k == "dualTolerance")
    s.dualTolerance = value;
  else if (k == "dualBound")
    s.dualBound = value;
  else if (k == "maximumIterations")
    s.maximumIterations = static_cast<int>(value);
  else if (k == "factorizationFrequency")
    s.factorizationFrequency = static_cast<int>(value);
  else if (k == "logLevel")
    s.logLevel = static_cast<int>(value);
  else if (k == "threads")
    s.threads = std::max(1, static_cast<int>(value));
  else if (k == "maximumSeconds")
    s.maximumSeconds = value;
  else if (k == "warmupIterations")
    s.warmupIterations = static_cast<int>(value);
  else if (k == "bucketedRatioTest")
    s.bucketedRatioTest = static_cast<int>(value);
}
void orc_set_status(orc_model *h, const unsigned char *status)
{
  h->s.status.assign(status, status + h->s.n + h->s.m);
  h->s.haveUserStatus = true;
}
int orc_dual(orc_model *h) { return h->s.dual(); }
double orc_objective_value(const orc_model *h) { return h->s.objectiveValue; }
int orc_number_iterations(const orc_model *h) { return h->s.numberIterations; }
int orc_number_refactorizations(const orc_model *h) { return h->s.numberRefactorizations; }
double orc_seconds_in_loop(const orc_model *h) { return h->s.secondsInLoop; }
double orc_timed_seconds(const orc_model *h) { return h->s.timedSeconds; }
int orc_timed_iterations(const orc_model *h) { return h->s.timedIterations; }
void orc_get_column_solution(const orc_model *h, double *x)
{
  std::copy(h->s.sol.begin(), h->s.sol.begin() + h->s.n, x);
}
void orc_get_row_activity(const orc_model *h, double *y)
{
  std::copy(h->s.sol.begin() + h->s.n, h->s.sol.end(), y);
}
void orc_get_reduced_cost(const orc_model *h, double *d)
{
  std::copy(h->s.dj.begin(), h->s.dj.begin() + h->s.n, d);
}
void orc_get_row_price(const orc_model *h, double *pi)
{
  // dual of row i = dj of the row variable (computeDuals: dj[n+i] = pi_i when nonbasic),
  // recomputed from the final basis for all rows
  orc::DualSimplex &s = const_cast<orc_model *>(h)->s;
  std::vector<double> p(s.m);
  for (int k = 0; k < s.m; k++)
    p[k] = s.costTrue[s.pivotVariable[k]];
  s.fac.btran(p.data());
  std::copy(p.begin(), p.end(), pi);
}
void orc_get_status(const orc_model *h, unsigned char *st)
{
  std::copy(h->s.status.begin(), h->s.status.end(), st);
}

int orc_factorize(orc_model *h, const int *basicSequence, int *pivotVariableOut)
{
  orc::DualSimplex &s = h->s;
  s.pivotVariable.assign(basicSequence, basicSequence + s.m);
  s.weights.assign(s.m, 1.0);
  for (int j = 0; j < s.n + s.m; j++)
    if (s.status[j] == ORC_basic)
      s.status[j] = ORC_atLowerBound;
  for (int p = 0; p < s.m; p++)
    s.status[s.pivotVariable[p]] = ORC_basic;
  if (s.fac.maximumPivots < 1000)
    s.fac.maximumPivots = 1000;
  int rc = s.factorize();
  std::copy(s.pivotVariable.begin(), s.pivotVariable.end(), pivotVariableOut);
  return rc;
}
void orc_ftran(orc_model *h, double *v) { h->s.fac.ftran(v, false); }
void orc_btran(orc_model *h, double *v) { h->s.fac.btran(v, -1); }
int orc_replace_column(orc_model *h, int sequenceIn, int pivotRow)
{
  orc::DualSimplex &s = h->s;
  std::vector<double> col(s.m);
  s.unpack(sequenceIn, col.data());
  s.fac.ftran(col.data(), true);
  int rc = s.fac.replaceColumn(pivotRow, col[pivotRow]);
  if (rc == 0 || rc == 1) {
    int out = s.pivotVariable[pivotRow];
    s.status[out] = ORC_atLowerBound;
    s.status[sequenceIn] = ORC_basic;
    s.pivotVariable[pivotRow] = sequenceIn;
  }
  return rc;
}
void orc_transpose_times(const orc_model *h, double scalar, const double *pi, double *z)
{
  h->s.transposeTimes(scalar, pi, z);
}
void orc_times(const orc_model *h, double scalar, const double *x, double *y)
{
  h->s.times(scalar, x, y);
}
int orc_dual_column(int count, const double *alpha, const double *dj, const double *range,
                    const unsigned char *stat, double infeasibility, double dualTolerance,
                    double acceptablePivot, double *theta_out, unsigned char *flips_out)
{
  return orc::dualColumn(count, alpha, dj, range, stat, infeasibility, dualTolerance,
                         acceptablePivot, theta_out, flips_out);
}
int orc_dual_column_bucketed(int count, const double *alpha, const double *dj, const double *range,
                             const unsigned char *stat, double infeasibility, double dualTolerance,
                             double acceptablePivot, double *theta_out, unsigned char *flips_out)
{
  return orc::dualColumnBucketed(count, alpha, dj, range, stat, infeasibility, dualTolerance,
                                 acceptablePivot, theta_out, flips_out);
}
void orc_dse_update(int m, double *weights, const double *alphaColumn, const double *tau,
                    int pivotRow, double rhoNorm2)
{
  orc::dseUpdate(m, weights, alphaColumn, tau, pivotRow, rhoNorm2);
}
}
