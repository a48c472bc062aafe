// kernels_common.cuh -- device helpers shared by solve.cu / price.cu / update.cu.
#pragma once
#include "engine.cuh"

namespace clpb {

__device__ __forceinline__ bool iter_active(const IterState *st) { return st->stop == 0; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// indices into DeviceModel::tailCounter
enum : int { TAIL_DUAL_UPDATE = 0, TAIL_ITER_UPDATE = 1, TAIL_PFI_APPLY = 2, TAIL_HIST2 = 3,
             TAIL_SELECT = 4 };

// "last block done": returns true in exactly one CTA of a 1-D grid, after every other CTA of the
// grid has passed this point (and hence finished the work before it).  What the tail then reads
// of other CTAs' results must bypass L1 (__ldcg / atomics): L1 is not coherent inside a kernel.
__device__ __forceinline__ bool last_block_done(unsigned int *counter)
{
  __shared__ int sIsLast;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(counter, 1u);
    const int last = (ticket == gridDim.x - 1);
    if (last)
      *counter = 0u; // ready for the next launch
    sIsLast = last;
  }
  __syncthreads();
  if (sIsLast)
    __threadfence();
  return sIsLast != 0;
}

// out[j] = scale * sum_{i>=j, i<t} Ginv[i][j] * vec[i]   for j in [32*jblock, 32*jblock+32), j < t
//   mode 0 : nu (BTRAN eta transposes), vec = W[pivot row][:]
//   mode 1 : new row t of Ginv = -out / alphaCol, diagonal 1/alphaCol, vec = W[pivot row][:]
//   mode 2 : nu for a general BTRAN, vec = d.mu (the t dot products W_i . v)
// 256 threads; 'part' is 8 x 33 doubles of shared memory.
__device__ __forceinline__ void eta_rowvec_body(const DeviceModel &d, int mode, int jblock,
                                                double (*part)[33])
{
  const int t = d.st->numEtas;
  const int j0 = jblock * 32;
  if (j0 >= t && !(mode == 1 && jblock == 0))
    return;
  const int r = d.st->pivotRow;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = j0 + lane;
  const double *wrow = mode == 2 ? d.mu : d.W + (size_t)r * d.tmax;
  double acc = 0.0;
  if (j < t)
    for (int i = j0 + warp; i < t; i += 8) // rows below j0 contribute nothing (lower triangular)
      if (i >= j)
        acc = fma(d.Ginv[(size_t)i * d.tmax + j], wrow[i], acc);
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++)
      s += part[w][lane];
    if (mode != 1) {
      if (j < t)
        d.nu[j] = s;
    } else {
      const double dinv = 1.0 / d.st->alphaCol;
      if (j < t)
        d.Ginv[(size_t)t * d.tmax + j] = -s * dinv;
      if (jblock == 0 && lane == 0)
        d.Ginv[(size_t)t * d.tmax + t] = dinv;
    }
  }
}

// packed (score, position) key of the dual steepest edge row choice; 0 = primal feasible
__device__ __forceinline__ unsigned long long chuzr_key(double x, double lo, double up, double weight,
                                                        double tol, int p)
{
  double inf = 0.0;
  if (x < lo - tol)
    inf = lo - x;
  else if (x > up + tol)
    inf = x - up;
  if (!(inf > 0.0))
    return 0ull;
  const double score = inf * inf / weight;
  return ((unsigned long long)__double_as_longlong(score) & ~0xFFFFFull) | (unsigned long long)(0xFFFFF - p);
}

// accuracy gate (ClpSimplexDual.cpp:1447-1501) and primal step length.  Runs as the tail of the
// eta-panel kernel of the FTRAN (solve.cu), i.e. single thread, results of other CTAs via L2.
__device__ __forceinline__ void pivot_scalars_body(const DeviceModel &d)
{
  IterState *st = d.st;
  const int r = st->pivotRow;
  const double ac = __ldcg(d.rhs3 + r);
  st->alphaCol = ac;
  const double ar = st->alphaRow;
  const double err = fabs(ar - ac) / (1.0 + fabs(ac));
  const bool bad = !(fabs(ac) >= 1.0e-9) || !(err <= 1.0e-6);
  if (bad) {
    if (st->numEtas > 0) {
      st->stop = STOP_INACCURATE; // the host refactorizes, recomputes and retries
      return;
    }
    if (!(fabs(ac) >= 1.0e-11) || !(err <= 1.0e-3)) {
      st->stop = STOP_TINY_PIVOT; // fresh factors and still no usable pivot
      return;
    }
  }
  const int seqOut = st->seqOut;
  double valueOut = d.sol[seqOut];
  if (st->numFlips > 0)
    valueOut += __ldcg(d.rhs3 + (size_t)2 * d.m + r);
  const double bound = st->sigma < 0 ? d.lower[seqOut] : d.upper[seqOut];
  st->thetaPrimal = (valueOut - bound) / ac;
}

// order independent histogram add with warp aggregation: lanes of a warp that hit the same bucket
// are summed first (one atomic per distinct bucket and warp).  Degenerate LPs put thousands of
// candidates into the ratio-0 bucket; without aggregation those atomics serialise in L2.
// Must be called by all 32 lanes of a converged warp; w < 2^42.
__device__ __forceinline__ void hist_add_aggregated(unsigned long long *hist, int bucket,
                                                    unsigned long long w, bool valid)
{
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (!valid)
    return;
  const unsigned peers = __match_any_sync(act, bucket);
  const unsigned lo = (unsigned)(w & 0xFFFFFull), hi = (unsigned)(w >> 20);
  const unsigned slo = __reduce_add_sync(peers, lo), shi = __reduce_add_sync(peers, hi);
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1)
    atomicAdd(hist + bucket, ((unsigned long long)shi << 20) + (unsigned long long)slo);
}
// same for atomicMin of 64-bit keys (called by the lanes with valid == true of the call above)
__device__ __forceinline__ void hist_min_aggregated(unsigned long long *hist, int bucket,
                                                    unsigned long long key, bool valid)
{
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (!valid)
    return;
  const unsigned peers = __match_any_sync(act, bucket);
  const unsigned hi = (unsigned)(key >> 32);
  const unsigned mh = __reduce_min_sync(peers, hi);
  const unsigned lo = hi == mh ? (unsigned)key : 0xFFFFFFFFu;
  const unsigned ml = __reduce_min_sync(peers, lo);
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1)
    atomicMin(hist + bucket, ((unsigned long long)mh << 32) | (unsigned long long)ml);
}

} // namespace clpb
