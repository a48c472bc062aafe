/* clp_b200.h -- C ABI of the B200-native revised dual simplex engine.
 *
 * Drop-in boundary for ONE path of coin-or/Clp: ClpModel::readMps/loadProblem ->
 * ClpSimplex::dual().  Every entry point names the reference interface it replaces
 * (paths relative to the coin-or/Clp source tree).  Conventions follow
 * src/Clp_C_Interface.h: one opaque model pointer, int status returns, caller-owned HOST
 * buffers (plain pointers and sizes), no exceptions across the boundary, single caller thread.
 * Sequence numbering is ClpSimplex's: 0..n-1 columns, n..n+m-1 rows; status bytes are
 * ClpSimplex::Status (src/ClpSimplex.hpp:119-126).
 *
 * The library has NO CPU fallback: every solve/plug-in call needs a CUDA device and returns
 * CLPB_NO_DEVICE (-100) otherwise.
 */
#ifndef CLP_B200_H
#define CLP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct Clpb_Simplex Clpb_Simplex;

#define CLPB_NO_DEVICE (-100)

/* ---------------- model level: replaces src/Clp_C_Interface.h ---------------- */
/* Clp_newModel :77 / Clp_deleteModel :79 */
Clpb_Simplex *Clpb_newModel(void);
void Clpb_deleteModel(Clpb_Simplex *model);
/* Clp_loadProblem :101 (column-major matrix without gaps; NULL rim arrays take Clp's defaults) */
int Clpb_loadProblem(Clpb_Simplex *model, int numcols, int numrows, const int *start,
                     const int *index, const double *value, const double *collb,
                     const double *colub, const double *obj, const double *rowlb,
                     const double *rowub);
/* Clp_readMps :116 (ClpModel::readMps src/ClpModel.cpp:2884) */
int Clpb_readMps(Clpb_Simplex *model, const char *filename, int keepNames, int ignoreErrors);
/* ClpPresolve::presolvedModel / postsolve (src/ClpPresolve.hpp:40,61) restricted to the elementary
   actions (fixed columns, singleton rows, empty columns, empty rows; src/ClpPresolve.cpp:966,1141,
   1448,1449).  Clpb_presolvedModel returns a NEW model (delete it with Clpb_deleteModel) holding the
   reduced problem, or NULL with *status = 1 (primal infeasible) / 2 (dual infeasible); the original
   model keeps the postsolve information.  After solving the reduced model, Clpb_postsolve writes
   the solution of the original problem (primal, dual, status, objective) into the original model.
   Clpb_setSolution hands a solution of a model in from outside (x[n], rowPrice[m], status[n+m]).
   Host only. */
Clpb_Simplex *Clpb_presolvedModel(Clpb_Simplex *model, int *status);
int Clpb_postsolve(Clpb_Simplex *model, Clpb_Simplex *presolved);
void Clpb_setSolution(Clpb_Simplex *model, const double *x, const double *rowPrice,
                      const unsigned char *status, int problemStatus);
/* Clp_writeMps :120 (ClpModel::writeMps src/ClpModel.cpp:3986): the model as loaded, default names
   R%7.7d / C%7.7d, 17 significant digits; formatType / numberAcross / objSense are accepted for
   signature compatibility.  0 ok, -1 cannot open.  Host only. */
int Clpb_writeMps(Clpb_Simplex *model, const char *filename, int formatType, int numberAcross,
                  double objSense);
/* Clp_numberRows :174, Clp_numberColumns :176, Clp_getNumElements :246 */
int Clpb_numberRows(Clpb_Simplex *model);
int Clpb_numberColumns(Clpb_Simplex *model);
long long Clpb_getNumElements(Clpb_Simplex *model);
/* problem data as loaded (for callers that used readMps): Clp_getColLower etc. :236-262 */
void Clpb_getProblem(Clpb_Simplex *model, int *start, int *index, double *value, double *collb,
                     double *colub, double *obj, double *rowlb, double *rowub);
/* Clp_setPrimalTolerance :179, Clp_setDualTolerance :182, Clp_setDualBound :382,
   Clp_setMaximumIterations :199, Clp_setMaximumSeconds :202, Clp_setLogLevel :314,
   ClpFactorization::maximumPivots (src/ClpFactorization.hpp:149).  Keys: "primalTolerance",
   "dualTolerance", "dualBound", "maximumIterations", "maximumSeconds", "logLevel",
   "factorizationFrequency", "scaling", "perturbation" (Clp_setPerturbation :302: 50 on,
   100 automatic, 102 off = default here; ClpSimplexDual::perturb src/ClpSimplexDual.cpp:6533), "batch" (iterations enqueued per host sync), "timing" (0/1: per-phase
   CUDA events, no graph replay), "useGraph" (0/1), "warmupIterations", "objectiveOffset". */
int Clpb_setParameter(Clpb_Simplex *model, const char *key, double value);
/* Clp_scaling :268 (ClpModel::scaling(int mode), src/ClpModel.hpp): 0 off (default here; the
   benchmark configuration is unscaled), 1 equilibrium, 2 geometric, 3 automatic (Clp's default),
   4 automatic-dynamic (treated as 3).  The same value can be set with the key "scaling".
   Clpb_scaleFactors runs ClpPackedMatrix::scale (src/ClpPackedMatrix.cpp:4120) on the host and
   copies rowScale[m] / columnScale[n] out (all 1 and return value 1 when the matrix is not
   worth scaling, :4262); dual() applies them to the matrix and the rim and returns the solution
   in the caller's units. */
void Clpb_scaling(Clpb_Simplex *model, int mode);
int Clpb_scaleFactors(Clpb_Simplex *model, double *rowScale, double *columnScale);
/* Host-only preview of ClpSimplexDual::perturb (src/ClpSimplexDual.cpp:6533) for the current
   "perturbation" setting and status: cost[n] = the (scaled) working objective after perturbation.
   Returns 0 if costs were perturbed, 1 if the rule decided not to (copy of the objective). */
int Clpb_perturbedCosts(Clpb_Simplex *model, double *cost);
/* Clp_copyinStatus :280 : status[n+m], columns first */
void Clpb_copyinStatus(Clpb_Simplex *model, const unsigned char *statusArray);
/* ClpSimplex::writeBasis / readBasis (src/ClpSimplex.cpp:6569 / :6577 -> ClpSimplexOther.cpp:1018 /
   :1136): MPS basis file (XU/XL/UL/LL/BS records) in the reference's no-names form C%7.7d / R%7.7d;
   writeValues / formatType are accepted for signature compatibility, values are not written.
   readBasis returns 0, -1 (cannot open) or the number of records it could not interpret; the basis
   becomes the starting basis of the next Clpb_dual (like Clpb_copyinStatus).  Host only. */
int Clpb_writeBasis(Clpb_Simplex *model, const char *filename, int writeValues, int formatType);
int Clpb_readBasis(Clpb_Simplex *model, const char *filename);
/* Clp_dual :346 (ClpSimplex::dual src/ClpSimplex.cpp:5631).  Returns Clp_status :212:
   0 optimal, 1 primal infeasible, 2 dual infeasible, 3 stopped on iterations/time,
   4 stopped due to errors. */
int Clpb_dual(Clpb_Simplex *model, int ifValuesPass);
int Clpb_status(Clpb_Simplex *model);
/* Clp_objectiveValue :256, Clp_numberIterations :195 */
double Clpb_objectiveValue(Clpb_Simplex *model);
int Clpb_numberIterations(Clpb_Simplex *model);
int Clpb_numberRefactorizations(Clpb_Simplex *model);
/* Clp_primalColumnSolution :230, Clp_primalRowSolution :228, Clp_dualColumnSolution :234,
   Clp_dualRowSolution :232, Clp_statusArray :278 -- copied into caller buffers */
void Clpb_primalColumnSolution(Clpb_Simplex *model, double *x /* n */);
void Clpb_primalRowSolution(Clpb_Simplex *model, double *rowActivity /* m */);
void Clpb_dualColumnSolution(Clpb_Simplex *model, double *reducedCost /* n */);
void Clpb_dualRowSolution(Clpb_Simplex *model, double *rowPrice /* m */);
void Clpb_statusArray(Clpb_Simplex *model, unsigned char *status /* n+m */);
/* measurement: seconds in the iteration loop, kernels launched, per-phase device ms
   (order: chuzr, btran, price, chuzc, dualUpdate, ftran, update, refactor, samples, then the
   single kernels priceKernel, ftranGemv, btranGemv, then the algorithmic bytes summed over the
   timed FTRAN / BTRAN GEMV launches; 14 doubles) */
double Clpb_secondsInLoop(Clpb_Simplex *model);
long long Clpb_kernelLaunches(Clpb_Simplex *model);
void Clpb_phaseTimes(Clpb_Simplex *model, double *out14);
int Clpb_nucleusSize(Clpb_Simplex *model);
/* CUDA-event time (on the engine's stream) and iteration count of the window that starts after
   "warmupIterations" iterations and ends when Clpb_dual returns (refactorizations included) */
void Clpb_timedWindow(Clpb_Simplex *model, double *milliseconds, int *iterations);

/* column-sharded pricing across GPUs (one process per GPU).  ncclUniqueId (128 bytes) is
   created by rank 0 with Clpb_ncclUniqueId and shipped to the other ranks by the caller
   (torch.distributed broadcast); then every rank calls Clpb_initSharding. */
int Clpb_ncclUniqueId(unsigned char *id128);
int Clpb_initSharding(Clpb_Simplex *model, int rank, int worldSize, const unsigned char *id128);

/* ---------------- plug-in level: the three interfaces whileIterating calls ------------- */
/* ClpFactorization::factorize (src/ClpFactorization.hpp:54, .cpp:1649).  basicSequence[m] in,
   pivotVariable[m] out: pivotVariable[i] pivots on row i.  0 ok, -1 singular. */
int Clpb_factorize(Clpb_Simplex *model, const int *basicSequence, int *pivotVariable);
/* ClpFactorization::updateColumn (hpp:117, FTRAN) / updateColumnTranspose (hpp:135, BTRAN);
   region[m] dense, in place */
int Clpb_updateColumn(Clpb_Simplex *model, double *region);
int Clpb_updateColumnTranspose(Clpb_Simplex *model, double *region);
/* ClpFactorization::replaceColumn (hpp:89): 0 ok, 2 singular (nothing changed), 5 max pivots */
int Clpb_replaceColumn(Clpb_Simplex *model, int sequenceIn, int pivotRow);
/* ClpMatrixBase::transposeTimes (src/ClpMatrixBase.hpp:287) z[n] = scalar*A^T pi ;
   ClpMatrixBase::times (hpp:275) y[m] = scalar*A x */
int Clpb_transposeTimes(Clpb_Simplex *model, double scalar, const double *pi, double *z);
int Clpb_times(Clpb_Simplex *model, double scalar, const double *x, double *y);
/* ClpSimplexDual::dualColumn (src/ClpSimplexDual.cpp:4192) on an explicit tableau row:
   alphaRow[n+m], dj[n+m], status[n+m]; direction +1 leaving to upper / -1 to lower.
   Returns sequenceIn or -1; *theta = dual step. */
int Clpb_dualColumn(Clpb_Simplex *model, const double *alphaRow, const double *dj,
                    const unsigned char *status, int direction, double infeasibility,
                    double *theta);
/* The dense kernel of the refactorization on its own (CoinAbcDgetrf + inverse,
   src/AbcSimplexParallel.cpp:2491): a[k*k] column-major in, x = a^-1 column-major out.
   Returns 0, or 1 + the index of the first column without an acceptable pivot. */
int Clpb_denseInvert(int k, const double *a, double *x);
/* run the startup of dual() (basis from status, factorize, computePrimals/Duals) and then
   'count' iterations; DSE weights by pivot row (ClpDualRowSteepest::weights_) */
int Clpb_startup(Clpb_Simplex *model);
int Clpb_iterate(Clpb_Simplex *model, int count);
void Clpb_getWeights(Clpb_Simplex *model, double *weights /* m */);
/* debug reads of device vectors: "sol","dj" (n+m), "rho" (m), "alphaRow" (n+m),
   "pivotVariable" (m, as doubles), "status" (n+m, as doubles) */
void Clpb_getDeviceVector(Clpb_Simplex *model, const char *name, double *out);

#ifdef __cplusplus
}
#endif
#endif
