// mps_reader.cpp -- MPS (fixed or free format, whitespace separated names) -> CSC + rim.
//
// Host-side replacement for the part of ClpModel::readMps (/root/reference/src/ClpModel.cpp:2884,
// which delegates to CoinMpsIO from CoinUtils -- not in the reference tree) that the
// "readMps -> dual()" drop-in needs: ROWS / COLUMNS / RHS / RANGES / BOUNDS sections, the
// first N row as objective, RHS on the objective row as (negated) constant, OBJSENSE MAX.
#include "engine.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <unordered_map>

namespace clpb {

int readMpsFile(const char *fileName, int &m, int &n, std::vector<int> &colStart,
                std::vector<int> &row, std::vector<double> &val, std::vector<double> &colLower,
                std::vector<double> &colUpper, std::vector<double> &obj,
                std::vector<double> &rowLower, std::vector<double> &rowUpper, double &objOffset,
                std::string &name)
{
  std::ifstream in(fileName);
  if (!in) {
    fprintf(stderr, "clp_b200: cannot open %s\n", fileName);
    return -1;
  }
  enum Section { NONE, ROWS, COLUMNS, RHS, RANGES, BOUNDS, OBJSENSE, ENDATA };
  Section sec = NONE;
  std::unordered_map<std::string, int> rowIndex, colIndex;
  std::vector<char> rowType;
  std::string objName;
  bool haveObj = false;
  bool maximize = false;
  std::vector<std::vector<std::pair<int, double>>> cols;
  std::vector<double> rhs, range;
  std::vector<char> hasRange;
  std::vector<double> clo, cup;
  std::vector<char> boundSetLower;
  std::vector<double> objCoef;
  objOffset = 0.0;
  bool inInteger = false;
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '*')
      continue;
    std::istringstream ss(line);
    std::vector<std::string> tok;
    std::string t;
    while (ss >> t)
      tok.push_back(t);
    if (tok.empty())
      continue;
    if (line[0] != ' ' && line[0] != '\t') {
      const std::string &h = tok[0];
      if (h == "NAME") {
        name = tok.size() > 1 ? tok[1] : "";
        sec = NONE;
      } else if (h == "ROWS")
        sec = ROWS;
      else if (h == "COLUMNS")
        sec = COLUMNS;
      else if (h == "RHS")
        sec = RHS;
      else if (h == "RANGES")
        sec = RANGES;
      else if (h == "BOUNDS")
        sec = BOUNDS;
      else if (h == "OBJSENSE") {
        sec = OBJSENSE;
        if (tok.size() > 1 && (tok[1] == "MAX" || tok[1] == "MAXIMIZE"))
          maximize = true;
      } else if (h == "ENDATA") {
        sec = ENDATA;
        break;
      } else
        sec = NONE;
      continue;
    }
    switch (sec) {
    case OBJSENSE:
      if (tok[0] == "MAX" || tok[0] == "MAXIMIZE")
        maximize = true;
      break;
    case ROWS: {
      if (tok.size() < 2)
        break;
      char ty = tok[0][0];
      if (ty == 'N') {
        if (!haveObj) {
          haveObj = true;
          objName = tok[1];
        }
        // other free rows are dropped (CoinMpsIO keeps them as free rows; they never bind)
        else
          rowIndex[tok[1]] = -2;
      } else {
        rowIndex[tok[1]] = (int)rowType.size();
        rowType.push_back(ty);
      }
      break;
    }
    case COLUMNS: {
      if (tok.size() >= 3 && tok[1] == "'MARKER'") {
        inInteger = (tok[2] == "'INTORG'");
        (void)inInteger;
        break;
      }
      if (tok.size() < 3)
        break;
      int j;
      auto it = colIndex.find(tok[0]);
      if (it == colIndex.end()) {
        j = (int)cols.size();
        colIndex[tok[0]] = j;
        cols.emplace_back();
        objCoef.push_back(0.0);
      } else
        j = it->second;
      for (size_t k = 1; k + 1 < tok.size(); k += 2) {
        double v = atof(tok[k + 1].c_str());
        if (haveObj && tok[k] == objName) {
          objCoef[j] += v;
          continue;
        }
        auto r = rowIndex.find(tok[k]);
        if (r == rowIndex.end()) {
          fprintf(stderr, "clp_b200: unknown row %s in COLUMNS\n", tok[k].c_str());
          return -2;
        }
        if (r->second >= 0 && v != 0.0)
          cols[j].emplace_back(r->second, v);
      }
      break;
    }
    case RHS: {
      if (rhs.empty())
        rhs.assign(rowType.size(), 0.0);
      size_t k0 = (tok.size() % 2 == 1) ? 1 : 0; // optional set name
      for (size_t k = k0; k + 1 < tok.size(); k += 2) {
        double v = atof(tok[k + 1].c_str());
        if (haveObj && tok[k] == objName) {
          objOffset = -v;
          continue;
        }
        auto r = rowIndex.find(tok[k]);
        if (r != rowIndex.end() && r->second >= 0)
          rhs[r->second] = v;
      }
      break;
    }
    case RANGES: {
      if (range.empty()) {
        range.assign(rowType.size(), 0.0);
        hasRange.assign(rowType.size(), 0);
      }
      size_t k0 = (tok.size() % 2 == 1) ? 1 : 0;
      for (size_t k = k0; k + 1 < tok.size(); k += 2) {
        auto r = rowIndex.find(tok[k]);
        if (r != rowIndex.end() && r->second >= 0) {
          range[r->second] = atof(tok[k + 1].c_str());
          hasRange[r->second] = 1;
        }
      }
      break;
    }
    case BOUNDS: {
      if (clo.empty()) {
        clo.assign(cols.size(), 0.0);
        cup.assign(cols.size(), kInf);
        boundSetLower.assign(cols.size(), 0);
      }
      const std::string &ty = tok[0];
      // forms: TYPE SET COL VALUE | TYPE COL VALUE | TYPE SET COL | TYPE COL
      std::string cname;
      double v = 0.0;
      bool needVal = !(ty == "FR" || ty == "MI" || ty == "PL" || ty == "BV");
      if (needVal) {
        if (tok.size() >= 4) {
          cname = tok[2];
          v = atof(tok[3].c_str());
        } else if (tok.size() == 3) {
          cname = tok[1];
          v = atof(tok[2].c_str());
        } else
          break;
      } else {
        if (tok.size() >= 3)
          cname = tok[2];
        else if (tok.size() == 2)
          cname = tok[1];
        else
          break;
        if (colIndex.find(cname) == colIndex.end() && tok.size() >= 3)
          cname = tok[1];
      }
      auto c = colIndex.find(cname);
      if (c == colIndex.end()) {
        fprintf(stderr, "clp_b200: unknown column %s in BOUNDS\n", cname.c_str());
        return -3;
      }
      int j = c->second;
      if (ty == "UP" || ty == "UI") {
        cup[j] = v;
        if (v < 0.0 && !boundSetLower[j] && clo[j] == 0.0)
          clo[j] = -kInf; // classic MPS convention (CoinMpsIO does the same)
      } else if (ty == "LO" || ty == "LI") {
        clo[j] = v;
        boundSetLower[j] = 1;
      } else if (ty == "FX") {
        clo[j] = cup[j] = v;
        boundSetLower[j] = 1;
      } else if (ty == "FR") {
        clo[j] = -kInf;
        cup[j] = kInf;
      } else if (ty == "MI") {
        clo[j] = -kInf;
        boundSetLower[j] = 1;
      } else if (ty == "PL") {
        cup[j] = kInf;
      } else if (ty == "BV") {
        clo[j] = 0.0;
        cup[j] = 1.0;
        boundSetLower[j] = 1;
      }
      break;
    }
    default:
      break;
    }
  }
  m = (int)rowType.size();
  n = (int)cols.size();
  if (rhs.empty())
    rhs.assign(m, 0.0);
  if (range.empty()) {
    range.assign(m, 0.0);
    hasRange.assign(m, 0);
  }
  if (clo.empty()) {
    clo.assign(n, 0.0);
    cup.assign(n, kInf);
  }
  clo.resize(n, 0.0);
  cup.resize(n, kInf);
  rowLower.assign(m, 0.0);
  rowUpper.assign(m, 0.0);
  for (int i = 0; i < m; i++) {
    double lo, up;
    switch (rowType[i]) {
    case 'E':
      lo = up = rhs[i];
      if (hasRange[i]) {
        if (range[i] > 0)
          up = rhs[i] + std::fabs(range[i]);
        else if (range[i] < 0)
          lo = rhs[i] - std::fabs(range[i]);
      }
      break;
    case 'L':
      up = rhs[i];
      lo = hasRange[i] ? rhs[i] - std::fabs(range[i]) : -kInf;
      break;
    case 'G':
      lo = rhs[i];
      up = hasRange[i] ? rhs[i] + std::fabs(range[i]) : kInf;
      break;
    default:
      lo = -kInf;
      up = kInf;
    }
    rowLower[i] = lo;
    rowUpper[i] = up;
  }
  colLower = clo;
  colUpper = cup;
  obj = objCoef;
  if (maximize) {
    for (double &c : obj)
      c = -c;
    objOffset = -objOffset;
  }
  colStart.assign(n + 1, 0);
  row.clear();
  val.clear();
  for (int j = 0; j < n; j++) {
    // merge duplicates, keep row order of appearance
    std::map<int, double> merged;
    for (auto &pr : cols[j])
      merged[pr.first] += pr.second;
    for (auto &pr : merged)
      if (pr.second != 0.0) {
        row.push_back(pr.first);
        val.push_back(pr.second);
      }
    colStart[j + 1] = (int)row.size();
  }
  return 0;
}

// ClpModel::writeMps (/root/reference/src/ClpModel.cpp:3986 -> CoinMpsIO::writeMps, CoinUtils) for
// the model as loaded: default names R%7.7d / C%7.7d (what Clp uses when a model has none), one
// entry per line, 17 significant digits (formatType 1, "extra accuracy").  Row types: E (equal
// bounds), L / G (one finite bound), L + RANGES (both finite), N (free row: the reader, like
// CoinMpsIO, drops such rows).  Column bounds: default [0, inf) is not written; FR, MI(+UP), FX,
// LO, UP otherwise.  objOffset goes to the RHS of the objective row, negated.
int writeMpsFile(const char *fileName, int m, int n, const std::vector<int> &colStart,
                 const std::vector<int> &row, const std::vector<double> &val,
                 const std::vector<double> &lower, const std::vector<double> &upper,
                 const std::vector<double> &cost, double objOffset, const std::string &name)
{
  FILE *fp = fopen(fileName, "w");
  if (!fp)
    return -1;
  fprintf(fp, "NAME          %s\nROWS\n N  OBJROW\n", name.empty() ? "BLANK" : name.c_str());
  std::vector<char> type(m);
  for (int i = 0; i < m; i++) {
    const double lo = lower[n + i], up = upper[n + i];
    const bool flo = lo > -kInf, fup = up < kInf;
    type[i] = (flo && fup) ? (lo == up ? 'E' : 'L') : fup ? 'L' : flo ? 'G' : 'N';
    fprintf(fp, " %c  R%7.7d\n", type[i], i);
  }
  fprintf(fp, "COLUMNS\n");
  for (int j = 0; j < n; j++) {
    if (cost[j] != 0.0 || colStart[j + 1] == colStart[j])
      fprintf(fp, "    C%7.7d  OBJROW    %.17g\n", j, cost[j]);
    for (int e = colStart[j]; e < colStart[j + 1]; e++)
      fprintf(fp, "    C%7.7d  R%7.7d  %.17g\n", j, row[e], val[e]);
  }
  fprintf(fp, "RHS\n");
  if (objOffset != 0.0)
    fprintf(fp, "    RHS       OBJROW    %.17g\n", -objOffset);
  for (int i = 0; i < m; i++) {
    const double rhs = type[i] == 'G' ? lower[n + i] : type[i] == 'N' ? 0.0 : upper[n + i];
    if (rhs != 0.0)
      fprintf(fp, "    RHS       R%7.7d  %.17g\n", i, rhs);
  }
  bool anyRange = false;
  for (int i = 0; i < m; i++)
    if (type[i] == 'L' && lower[n + i] > -kInf) {
      if (!anyRange)
        fprintf(fp, "RANGES\n");
      anyRange = true;
      fprintf(fp, "    RANGE     R%7.7d  %.17g\n", i, upper[n + i] - lower[n + i]);
    }
  bool anyBound = false;
  auto head = [&]() {
    if (!anyBound)
      fprintf(fp, "BOUNDS\n");
    anyBound = true;
  };
  for (int j = 0; j < n; j++) {
    const double lo = lower[j], up = upper[j];
    const bool flo = lo > -kInf, fup = up < kInf;
    if (!flo && !fup) {
      head();
      fprintf(fp, " FR BOUND     C%7.7d\n", j);
    } else if (flo && fup && lo == up) {
      head();
      fprintf(fp, " FX BOUND     C%7.7d  %.17g\n", j, lo);
    } else {
      if (!flo) {
        head();
        fprintf(fp, " MI BOUND     C%7.7d\n", j);
      } else if (lo != 0.0) {
        head();
        fprintf(fp, " LO BOUND     C%7.7d  %.17g\n", j, lo);
      }
      if (fup) {
        head();
        fprintf(fp, " UP BOUND     C%7.7d  %.17g\n", j, up);
      }
    }
  }
  fprintf(fp, "ENDATA\n");
  fclose(fp);
  return 0;
}

} // namespace clpb
