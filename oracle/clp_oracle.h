/* clp_oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of coin-or/Clp's revised
 * dual-simplex path (ClpSimplexDual / ClpDualRowSteepest / ClpPackedMatrix /
 * ClpFactorization / CoinAbcBaseFactorization).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (clp_b200/) never includes, links or calls anything in oracle/.
 *
 * Parity pin: objectives / solutions are pinned against the reference's own
 * known-answer fixtures (unitTest.cpp:1415-1482 3x5 LP; test/test_racing_reference.txt
 * TSP-MTZ / UFL / NQueens bounds) -- see tests/test_oracle_golden.py.  The LU itself is
 * "parity unpinned" at entry level (the reference's default LU lives in CoinUtils, which is
 * not in /root/reference): any correct LU is acceptable, solutions through it are pinned.
 */
#ifndef CLP_ORACLE_H
#define CLP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model;

/* status codes follow ClpSimplex::Status (ClpSimplex.hpp:119-126) */
enum { ORC_isFree = 0, ORC_basic = 1, ORC_atUpperBound = 2, ORC_atLowerBound = 3,
       ORC_superBasic = 4, ORC_isFixed = 5 };

/* problem status follows ClpModel::status(): 0 optimal, 1 primal infeasible,
 * 2 dual infeasible (unbounded), 3 stopped on iterations/time, 4 stopped due to errors */

orc_model *orc_create(int numberRows, int numberColumns, const int *columnStart,
                      const int *row, const double *element, const double *columnLower,
                      const double *columnUpper, const double *objective,
                      const double *rowLower, const double *rowUpper);
void orc_destroy(orc_model *);

/* option keys: "primalTolerance","dualTolerance","dualBound","maximumIterations",
 * "factorizationFrequency","logLevel","threads","maximumSeconds","bucketedRatioTest" */
void orc_set_option(orc_model *, const char *key, double value);

/* optional starting basis: status[numberColumns+numberRows] (columns first, Clp order) */
void orc_set_status(orc_model *, const unsigned char *status);

/* ClpSimplex::dual() equivalent.  Returns problem status. */
int orc_dual(orc_model *);

double orc_objective_value(const orc_model *);
int orc_number_iterations(const orc_model *);
int orc_number_refactorizations(const orc_model *);
double orc_seconds_in_loop(const orc_model *);
/* window that starts after option "warmupIterations" iterations */
double orc_timed_seconds(const orc_model *);
int orc_timed_iterations(const orc_model *);
/* iteration index and seconds at which the steady-state window began (after 1st refactor) */
void orc_get_column_solution(const orc_model *, double *x);      /* n */
void orc_get_row_activity(const orc_model *, double *y);         /* m */
void orc_get_reduced_cost(const orc_model *, double *dj);        /* n */
void orc_get_row_price(const orc_model *, double *pi);           /* m */
void orc_get_status(const orc_model *, unsigned char *status);   /* n+m */

/* ---- kernel-level entry points (for GPU kernel parity tests) ----
 * All vectors are dense, row-indexed (length m) or column-indexed (length n). */

/* ClpFactorization::factorize : basic sequence list (m entries, 0..n-1 structural,
 * n+i slack of row i).  On return pivotVariable[m] holds the list permuted so that
 * pivotVariable[i] pivots on row i (ClpFactorization.cpp:2300-2321).  0 ok, -1 singular. */
int orc_factorize(orc_model *, const int *basicSequence, int *pivotVariable);
/* ClpFactorization::updateColumn (FTRAN), in place */
void orc_ftran(orc_model *, double *vec);
/* ClpFactorization::updateColumnTranspose (BTRAN), in place */
void orc_btran(orc_model *, double *vec);
/* FTRAN of structural/slack column 'sequenceIn' keeping the FT spike, then
 * ClpFactorization::replaceColumn on pivot row.  returns 0,1,2,3,5 */
int orc_replace_column(orc_model *, int sequenceIn, int pivotRow);
/* ClpPackedMatrix::transposeTimes : z[n] = scalar * A^T pi */
void orc_transpose_times(const orc_model *, double scalar, const double *pi, double *z);
/* ClpPackedMatrix::times : y[m] += scalar * A x */
void orc_times(const orc_model *, double scalar, const double *x, double *y);

/* ClpSimplexDual::dualColumn restated on explicit inputs: for k in [0,count):
 * seq[k] candidate sequence, alpha[k] signed tableau-row entry (already multiplied by the
 * leaving direction), dj[k], range[k] (upper-lower, >=1e30 if not boxed), stat[k] status.
 * Returns chosen k (or -1); theta_out = dual step. flips_out[k]=1 if k passed. */
int orc_dual_column(int count, const double *alpha, const double *dj, const double *range,
                    const unsigned char *stat, double infeasibility, double dualTolerance,
                    double acceptablePivot, double *theta_out, unsigned char *flips_out);

/* Same ratio test evaluated the way the GPU evaluates it (two-level ratio histogram instead of
 * sorted passes; clp_b200/csrc/price.cu) -- lets tests follow the GPU's pivot sequence. */
int orc_dual_column_bucketed(int count, const double *alpha, const double *dj,
                             const double *range, const unsigned char *stat,
                             double infeasibility, double dualTolerance, double acceptablePivot,
                             double *theta_out, unsigned char *flips_out);

/* ClpDualRowSteepest::updateWeights recurrence on explicit arrays (in place on weights) */
void orc_dse_update(int m, double *weights, const double *alphaColumn, const double *tau,
                    int pivotRow, double rhoNorm2);

#ifdef __cplusplus
}
#endif
#endif
