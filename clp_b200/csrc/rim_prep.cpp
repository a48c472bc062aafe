// rim_prep.cpp -- host-side preparation of the working problem: scale factors and the dual (cost)
// perturbation.  Both follow the reference's RULES (constants and decision thresholds are what make a
// scaled / perturbed Clp run reproducible: ClpPackedMatrix::scale, src/ClpPackedMatrix.cpp:4120-4640;
// ClpSimplexDual::perturb, src/ClpSimplexDual.cpp:6533-6964) but are organised the way the rest of this
// engine is: whole-vector passes over the CSC / CSR copies, small pure helpers, no shared mutable state.
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <functional>
#include <numeric>

namespace clpb {

namespace {

// ---------------------------------------------------------------------------------------------
// scaling
struct Extrema {
  double hi, lo;
};

// One sweep target: for every line (row or column) of a sparse copy, the largest and smallest
// |a_ij| * weight[other index] over the entries whose column is in play.
struct SparseLines {
  const std::vector<int> &start, &other;
  const std::vector<double> &value;
  template <class Keep>
  Extrema extrema(int line, const std::vector<double> &weight, Keep keep, double hi0, double lo0) const
  {
    Extrema x{hi0, lo0};
    for (int e = start[line]; e < start[line + 1]; e++)
      if (keep(other[e])) {
        const double v = std::fabs(value[e]) * weight[other[e]];
        x.hi = std::max(x.hi, v);
        x.lo = std::min(x.lo, v);
      }
    return x;
  }
};

struct ScaleCandidate {
  std::vector<double> row, col;
  double spread = 0.0; // smallest / largest entry of the worst column once rows are scaled (:4471)
};

} // namespace

// Returns 1 (problem left unscaled) when there is nothing to gain: scaling switched off, or every
// entry of the columns in play already lies in [0.5, 2] (:4262).  Entries <= 1e-20 are ignored.
int Engine::computeScaling()
{
  rowScale.clear();
  columnScale.clear();
  if (scalingFlag <= 0 || m == 0 || n == 0)
    return 1;
  // columns "in play": not fixed (or basic in a user basis) and holding at least one real entry
  std::vector<char> inPlay(n, 0);
  Extrema all{0.0, 1.0e50};
  for (int j = 0; j < n; j++) {
    if (!(hUpper[j] > hLower[j] + 1.0e-12 || (haveUserStatus && hStatus[j] == basic)))
      continue;
    for (int e = hColStart[j]; e < hColStart[j + 1]; e++) {
      const double v = std::fabs(hVal[e]);
      if (v > 1.0e-20) {
        inPlay[j] = 1;
        all.hi = std::max(all.hi, v);
        all.lo = std::min(all.lo, v);
      }
    }
  }
  if (all.lo >= 0.5 && all.hi <= 2.0)
    return 1;

  std::vector<int> rowStart, colIdx;
  std::vector<double> rval;
  buildRowCopy(hVal, rowStart, colIdx, rval);
  const SparseLines byRow{rowStart, colIdx, rval};
  const SparseLines byCol{hColStart, hRow, hVal};
  const auto playing = [&](int j) { return inPlay[j] != 0; };
  const auto anyRow = [](int) { return true; };
  const double rangeFloor = 5.0 * primalTolerance;

  // A candidate = row factors from one strategy, then the guard against ranged rows whose scaled
  // range would fall below 1e-4 (:4459), then the figure of merit the automatic mode compares.
  auto finish = [&](ScaleCandidate &c) {
    for (int i = 0; i < m; i++) {
      const double scaledRange = (hUpper[n + i] - hLower[n + i]) * c.row[i];
      if (scaledRange > rangeFloor && scaledRange < 1.0e-4)
        c.row[i] = std::clamp(c.row[i] * (1.0e-4 / scaledRange), 1.0e-10, 1.0e10);
    }
    c.spread = 1.0e50;
    for (int j = 0; j < n; j++)
      if (inPlay[j]) {
        const Extrema x = byCol.extrema(j, c.row, anyRow, 1.0e-20, 1.0e50);
        if (c.spread * x.hi > x.lo)
          c.spread = x.lo / x.hi;
      }
  };
  auto equilibrium = [&]() { // every row divided by its largest entry
    ScaleCandidate c{std::vector<double>(m, 1.0), std::vector<double>(n, 1.0)};
    for (int i = 0; i < m; i++)
      c.row[i] = 1.0 / byRow.extrema(i, c.col, playing, 1.0e-10, 1.0e50).hi;
    finish(c);
    return c;
  };
  auto geometric = [&]() { // sqrt(min*max): rows, columns, rows, columns, rows
    ScaleCandidate c{std::vector<double>(m, 1.0), std::vector<double>(n, 1.0)};
    for (int sweep = 0; sweep < 3; sweep++) {
      for (int i = 0; i < m; i++) {
        const Extrema x = byRow.extrema(i, c.col, playing, 1.0e-50, 1.0e50);
        c.row[i] = 1.0 / std::sqrt(x.lo * x.hi);
      }
      if (sweep == 1)
        break; // the reference skips the last column round and the row round after it
      for (int j = 0; j < n; j++)
        if (inPlay[j]) {
          const Extrema x = byCol.extrema(j, c.row, anyRow, 1.0e-50, 1.0e50);
          c.col[j] = 1.0 / std::sqrt(x.lo * x.hi);
        }
    }
    finish(c);
    return c;
  };

  const int mode = scalingFlag == 4 ? 3 : scalingFlag >= 5 ? 2 : scalingFlag;
  ScaleCandidate chosen;
  if (mode == 1) {
    chosen = equilibrium();
  } else if (mode == 2) {
    chosen = geometric();
  } else {
    // automatic: geometric only if its worst column spread is more than twice as good (:4493-4513)
    ScaleCandidate e = equilibrium();
    ScaleCandidate g = geometric();
    chosen = g.spread > 2.0 * e.spread ? std::move(g) : std::move(e);
  }
  rowScale = std::move(chosen.row);

  // final column factors (:4531-4581): the largest scaled entry of every non-fixed, non-empty column
  // becomes 'target' (1 .. 100, larger when the spread is poor); narrow boxes are widened to 1e-5
  const double target = std::min(100.0, chosen.spread < 1.0e-1 ? 1.0 / std::sqrt(chosen.spread) : 1.0);
  columnScale.assign(n, 1.0);
  std::vector<char> rowSeen(m, 0);
  for (int j = 0; j < n; j++) {
    if (!(hUpper[j] > hLower[j] + 1.0e-12) || hColStart[j + 1] == hColStart[j])
      continue;
    double hi = 1.0e-20;
    for (int e = hColStart[j]; e < hColStart[j + 1]; e++) {
      rowSeen[hRow[e]] = 1;
      hi = std::max(hi, std::fabs(hVal[e] * rowScale[hRow[e]]));
    }
    const double box = hUpper[j] - hLower[j];
    columnScale[j] = box < 1.0e-5 * (target / hi) ? box / 1.0e-5 : target / hi;
  }
  for (int i = 0; i < m; i++)
    if (!rowSeen[i])
      rowScale[i] = 1.0;
  return 0;
}

// ClpSimplex::createRim (src/ClpSimplex.cpp:7895ff) for the scaled problem: columns x' = x/c,
// bounds/c, cost*c; rows activity' = r*activity, bounds*r.  Infinite bounds stay infinite.
void Engine::prepareWorkingProblem()
{
  wVal = hVal;
  wLower = hLower;
  wUpper = hUpper;
  wCost = hCost;
  if (computeScaling() != 0)
    return;
  const auto finite = [](double v) { return v > -kInf && v < kInf; };
  for (int j = 0; j < n; j++) {
    const double c = columnScale[j];
    for (int e = hColStart[j]; e < hColStart[j + 1]; e++)
      wVal[e] = hVal[e] * rowScale[hRow[e]] * c;
    if (finite(wLower[j]))
      wLower[j] /= c;
    if (finite(wUpper[j]))
      wUpper[j] /= c;
    wCost[j] *= c;
  }
  for (int i = 0; i < m; i++) {
    if (finite(wLower[n + i]))
      wLower[n + i] *= rowScale[i];
    if (finite(wUpper[n + i]))
      wUpper[n + i] *= rowScale[i];
  }
}

// ---------------------------------------------------------------------------------------------
// dual perturbation
namespace {

struct SplitMix64 { // fixed stream: the reference draws from CoinThreadRandom (CoinUtils, not in its tree)
  unsigned long long state = 0x9E3779B97F4A7C15ull;
  double next()
  {
    unsigned long long z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
  }
};

// "all finite non-zero bounds have the same magnitude" detector (:6640-6700)
struct SameMagnitude {
  double first = 0.0;
  bool same = true;
  void see(double bound)
  {
    if (bound == 0.0 || std::fabs(bound) >= 1.0e10)
      return;
    const double b = std::fabs(bound);
    if (first == 0.0)
      first = b;
    else if (std::fabs(b - first) > 1.0e-7)
      same = false;
  }
};

// push |v| into (floor, ceil] by whole decades, keeping the sign (:6862-6874)
double intoDecadeRange(double v, double floor, double ceil)
{
  while (std::fabs(v) <= floor)
    v *= 10.0;
  while (std::fabs(v) > ceil)
    v *= 0.1;
  return v;
}

} // namespace

// Cost perturbation before the first iteration (perturbation_ = 50 always, 100 only when the costs
// have few distinct values).  Every nonbasic, non-fixed column gets a change of the sign that keeps its
// reduced cost on the feasible side; its size is a fraction of |cost| (at least 100*dualTolerance),
// randomised in [0.5, 1], weighted by the column length (:6790 table) and kept inside
// [min(0.01*dualTol, fraction), max(1000*dualTol, fraction*average|cost|)].  Row costs stay (:6754).
// The individual values are not the reference's (different random stream); what is pinned is that the
// perturbed solve ends at the true optimum: the changes are removed like cost shifts at the first
// "optimal" basis and the dual simplex continues on the true costs (Engine::dual).
// Returns 0 when cost[] was perturbed, 1 when the rule says not to.
int Engine::perturbCosts(std::vector<double> &cost) const
{
  const double dualTol = dualTolerance;
  const double hugeBound = 1.0e15; // ClpSimplex::largeValue_
  const auto boxed = [&](int j) { return wLower[j] < wUpper[j]; };
  const auto length = [&](int j) { return hColStart[j + 1] - hColStart[j]; };

  // ---- is it worth it?  statistics of the objective as loaded (before scaling)
  std::vector<double> magnitude(n);
  for (int j = 0; j < n; j++)
    magnitude[j] = std::fabs(hCost[j]);
  const int nonZero = (int)std::count_if(magnitude.begin(), magnitude.end(), [](double v) { return v != 0.0; });
  const double averageCost = nonZero ? std::accumulate(magnitude.begin(), magnitude.end(), 0.0) / nonZero : 1.0;
  std::sort(magnitude.begin(), magnitude.end());
  const int distinct = (int)(std::unique(magnitude.begin(), magnitude.end()) - magnitude.begin());
  if (!nonZero && perturbation < 55)
    return 1; // no objective: the reference says "safer to use primal"
  if (perturbation >= 100 && distinct * 4 > n)
    return 1; // plenty of distinct costs already

  // ---- scale of the perturbation
  int longest = 0, shortest = m;
  for (int j = 0; j < n; j++)
    if (boxed(j) && length(j) > 2) {
      longest = std::max(longest, length(j));
      shortest = std::min(shortest, length(j));
    }
  double fraction = 1.0e-5;
  double largestCost = 1.0e-8, smallestCost = 1.0e100;
  SameMagnitude rowBounds, colBounds;
  for (int i = 0; i < m; i++) {
    rowBounds.see(wLower[n + i]);
    rowBounds.see(wUpper[n + i]);
  }
  for (int j = 0; j < n; j++) {
    if (boxed(j)) {
      const double v = std::fabs(cost[j]);
      largestCost = std::max(largestCost, v);
      if (v != 0.0)
        smallestCost = std::min(smallestCost, v);
    }
    colBounds.see(wLower[j]);
    colBounds.see(wUpper[j]);
  }
  if (rowBounds.same && colBounds.same) {
    // all bounds alike; if the matrix has a single positive and a single negative value as well
    // (set covering and friends) the reference "really hits" the perturbation (:6703)
    double negLo = 0.0, negHi = 0.0, posLo = 0.0, posHi = 0.0;
    bool anyNeg = false, anyPos = false;
    for (double v : wVal) {
      if (v < 0.0) {
        negLo = anyNeg ? std::min(negLo, v) : v;
        negHi = anyNeg ? std::max(negHi, v) : v;
        anyNeg = true;
      } else if (v > 0.0) {
        posLo = anyPos ? std::min(posLo, v) : v;
        posHi = anyPos ? std::max(posHi, v) : v;
        anyPos = true;
      }
    }
    if (negLo == negHi && posLo == posHi)
      fraction = std::max(fraction, std::min(100.0 * fraction, 1.0e-3 * std::max(rowBounds.first, colBounds.first)));
  }
  const double size = std::min(largestCost, smallestCost / fraction);
  const double constant = 100.0 * dualTol;
  const double floorAllowed = std::min(1.0e-2 * dualTol, fraction);
  const double ceilAllowed = std::max(1.0e3 * dualTol, fraction * averageCost);
  const double lengthFactor = longest ? 3.0 / (double)shortest : 1.0;
  static const double lengthWeight[] = {1.0e-4, 1.0e-2, 5.0e-1, 1.0, 2.0, 5.0, 10.0, 20.0, 30.0, 40.0, 100.0};
  const auto weightOf = [&](int len) {
    if (len > 3)
      len = std::max(3, (int)((double)len * lengthFactor));
    return lengthWeight[std::min(len, 10)];
  };

  // ---- per column
  SplitMix64 rng;
  double largestOnZeroCost = 0.0, largestOnNonZero = 0.0;
  for (int j = 0; j < n; j++) {
    const double u1 = 0.5 + 0.5 * rng.next(), u2 = 0.5 + 0.5 * rng.next(); // one pair per column, drawn always
    if (!boxed(j) || hStatus[j] == basic)
      continue;
    // the side the reduced cost has to stay on: +1 cost goes up (column sits at / near its lower
    // bound), -1 cost goes down; 0: the nearer bound is the upper one of a two-sided box -- left alone
    int side = 0;
    if (wLower[j] > -hugeBound)
      side = std::fabs(wLower[j]) < std::fabs(wUpper[j]) ? +1 : 0;
    else if (wUpper[j] < hugeBound)
      side = -1;
    if (side == 0)
      continue;
    const double c = cost[j];
    double delta = std::min(size, constant + fraction * (std::fabs(c) + 1.0e-1 * size + 1.0e-8)) * u1 * side;
    const double cap = (constant + 1.0e-1 * smallestCost) * u2 * side;
    delta *= weightOf(length(j));
    delta = std::min(delta, cap); // signed comparison, as the reference does
    delta = intoDecadeRange(delta, floorAllowed, ceilAllowed);
    (c != 0.0 ? largestOnNonZero : largestOnZeroCost) =
        std::max(c != 0.0 ? largestOnNonZero : largestOnZeroCost, std::fabs(delta));
    cost[j] += hStatus[j] == atUpperBound ? -delta : delta;
  }
  if (largestOnZeroCost > largestOnNonZero && largestOnNonZero != 0.0) {
    // zero-cost columns must not be perturbed more than the others
    const double limit = std::max(1.0e-8, largestOnNonZero);
    for (int j = 0; j < n; j++)
      if (hCost[j] == 0.0)
        while (std::fabs(cost[j]) > limit)
          cost[j] *= 0.5;
  }
  return 0;
}

// Host-only preview used by the CPU test-suite: builds the working problem (scaling included),
// normalises the status array the way resetStateForRun does, and applies perturbCosts.
int Engine::previewPerturbation(double *costOut)
{
  prepareWorkingProblem();
  if (std::count(hStatus.begin(), hStatus.end(), (unsigned char)basic) != m) {
    hStatus.assign(nm, atLowerBound);
    std::fill(hStatus.begin() + n, hStatus.end(), (unsigned char)basic);
  }
  std::vector<double> pc(wCost.begin(), wCost.begin() + n);
  const int rc = perturbCosts(pc);
  std::copy(pc.begin(), pc.end(), costOut);
  return rc;
}

} // namespace clpb
