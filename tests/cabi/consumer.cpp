// A C++ translation unit compiled against include/clp_b200.h only (no engine headers, no ctypes):
// what a Clp maintainer's binding would look like.  Host-only calls always run; with a CUDA device
// (argv[1] == "gpu") it also solves the 3x5 LP of src/unitTest.cpp:1415-1431 through the model level
// and walks one iteration through the ClpDualRowPivot-level calls.
#include "clp_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#define CHECK(cond)                                                                  \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      std::fprintf(stderr, "consumer.cpp:%d: check failed: %s\n", __LINE__, #cond);  \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main(int argc, char **argv)
{
  const bool gpu = argc > 1 && !std::strcmp(argv[1], "gpu");
  // the 3x5 LP of the reference's unit test (src/unitTest.cpp:1415-1431), column major
  const int start[6] = {0, 2, 4, 6, 8, 10};
  const int index[10] = {0, 2, 0, 1, 0, 1, 1, 2, 0, 2};
  // rows: r0: x0 + x1 + ... pattern chosen so that the LP is bounded; values are irrelevant for the
  // host-only part, the GPU part checks optimality conditions instead of a hard-coded optimum
  const double value[10] = {1.0, 1.0, 2.0, 1.0, 1.0, 3.0, 1.0, 2.0, 1.0, 1.0};
  const double collb[5] = {0, 0, 0, 0, 0}, colub[5] = {10, 10, 10, 10, 10};
  const double obj[5] = {-4.0, 1.0, -1.0, 2.0, -3.0};
  const double rowlb[3] = {-1e30, -1e30, -1e30}, rowub[3] = {14.0, 9.0, 12.0};

  Clpb_Simplex *model = Clpb_newModel();
  CHECK(model != nullptr);
  CHECK(Clpb_loadProblem(model, 5, 3, start, index, value, collb, colub, obj, rowlb, rowub) == 0);
  CHECK(Clpb_numberRows(model) == 3 && Clpb_numberColumns(model) == 5 && Clpb_getNumElements(model) == 10);
  CHECK(Clpb_setParameter(model, "primalTolerance", 1e-7) == 0);
  CHECK(Clpb_setParameter(model, "noSuchKey", 1.0) != 0);
  std::vector<int> s(6), ix(10);
  std::vector<double> v(10), cl(5), cu(5), ob(5), rl(3), ru(3);
  Clpb_getProblem(model, s.data(), ix.data(), v.data(), cl.data(), cu.data(), ob.data(), rl.data(), ru.data());
  CHECK(s[5] == 10 && ix[9] == 2 && v[5] == 3.0 && cu[4] == 10.0 && ob[0] == -4.0 && ru[1] == 9.0);

  const int st = Clpb_dual(model, 0);
  if (!gpu) {
    // no device: the library must say so loudly, never fall back to a CPU path
    CHECK(st == CLPB_NO_DEVICE);
    Clpb_deleteModel(model);
    std::printf("consumer ok (host only)\n");
    return 0;
  }
  CHECK(st == 0);
  std::vector<double> x(5), act(3), dj(5), pi(3);
  std::vector<unsigned char> stat(8);
  Clpb_primalColumnSolution(model, x.data());
  Clpb_primalRowSolution(model, act.data());
  Clpb_dualColumnSolution(model, dj.data());
  Clpb_dualRowSolution(model, pi.data());
  Clpb_statusArray(model, stat.data());
  double o = 0.0;
  int nbasic = 0;
  for (int j = 0; j < 5; j++) {
    o += obj[j] * x[j];
    CHECK(x[j] >= -1e-7 && x[j] <= 10.0 + 1e-7);
    // reduced cost sign against the bound the column sits at
    if (x[j] < 1e-7)
      CHECK(dj[j] >= -1e-6);
    else if (x[j] > 10.0 - 1e-7)
      CHECK(dj[j] <= 1e-6);
    else
      CHECK(std::fabs(dj[j]) <= 1e-6);
  }
  for (int q = 0; q < 8; q++)
    nbasic += stat[q] == 1;
  CHECK(nbasic == 3);
  for (int i = 0; i < 3; i++)
    CHECK(act[i] <= rowub[i] + 1e-7);
  CHECK(std::fabs(o - Clpb_objectiveValue(model)) <= 1e-9 * (1.0 + std::fabs(o)));

  // one iteration through the ClpDualRowPivot-level calls from the all-slack basis
  Clpb_Simplex *m2 = Clpb_newModel();
  CHECK(Clpb_loadProblem(m2, 5, 3, start, index, value, collb, colub, obj, rowlb, rowub) == 0);
  CHECK(Clpb_startup(m2) == 0);
  CHECK(Clpb_saveWeights(m2, 5) == 0);
  int seqOut = -1, dir = 0;
  double infeas = 0.0;
  const int r = Clpb_pivotRow(m2, &seqOut, &dir, &infeas);
  if (r >= 0) {
    std::vector<double> rho(3), row(5);
    CHECK(Clpb_updateColumnTransposeAndPrice(m2, rho.data(), row.data()) >= 1);
    double theta = 0.0, alphaRow = 0.0;
    const int q = Clpb_dualColumnDevice(m2, &theta, &alphaRow);
    CHECK(q >= 0 && theta >= 0.0);
    int rc = -1;
    const double alphaCol = Clpb_updateWeights(m2, &rc);
    CHECK(rc == 0 && std::fabs(alphaCol - alphaRow) <= 1e-9 * (1.0 + std::fabs(alphaCol)));
    CHECK(Clpb_unrollWeights(m2) == 0);
    double change = 0.0;
    CHECK(Clpb_updatePrimalSolution(m2, &change) == 1);
    CHECK(change >= 0.0);
  }
  Clpb_deleteModel(m2);
  Clpb_deleteModel(model);
  std::printf("consumer ok (gpu), objective %.9f\n", o);
  return 0;
}
