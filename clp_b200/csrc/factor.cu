// factor.cu -- dense refactorization of the basis nucleus on the GPU.
//
// Replaces the numerical part of ClpFactorization::factorize
// (/root/reference/src/ClpFactorization.cpp:1649 -> CoinAbcTypeFactorization::factor
// src/CoinAbcBaseFactorization1.cpp:683; the dense tail factorDense ...2.cpp:976 ->
// CoinAbcDgetrf src/AbcSimplexParallel.cpp:2491 with its CoinAbcDgemm trailing update).
//
// The k x k nucleus (structural basic columns restricted to rows whose slack is nonbasic) is
// LU-factorized with partial pivoting by a blocked right-looking algorithm (panel kernel +
// row swaps + unit-lower TRSM + rank-NB DGEMM update) and the factors are then turned into
// the explicit inverse X = U^-1 L^-1 P by blocked forward/backward substitution (TRSM + DGEMM).
// Everything is fp64 FMA; the matrix is column major with leading dimension ld.
#include "engine.cuh"

namespace clpb {

constexpr int NB = 32;

// ---- panel factorization: columns [j0, j0+nb) rows [j0, k), single CTA ------------------
__global__ void __launch_bounds__(1024)
    lu_panel_kernel(double *__restrict__ A, int k, int ld, int j0, int nb, int *__restrict__ ipiv,
                    int *__restrict__ info, double singularTol)
{
  __shared__ double sVal[32];
  __shared__ int sIdx[32];
  __shared__ double urow[NB];
  __shared__ int sPiv;
  __shared__ double sPivVal;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int jj = 0; jj < nb; jj++) {
    const int j = j0 + jj;
    double *colj = A + (size_t)j * ld;
    // pivot search
    double best = -1.0;
    int bi = j;
    for (int i = j + tid; i < k; i += 1024) {
      double a = fabs(colj[i]);
      if (a > best) {
        best = a;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if (lane == 0) {
      sVal[warp] = best;
      sIdx[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
      best = sVal[lane];
      bi = sIdx[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (lane == 0) {
        sPiv = bi;
        sPivVal = best;
        ipiv[j] = bi;
      }
    }
    __syncthreads();
    const int piv = sPiv;
    const bool singular = !(sPivVal >= singularTol);
    if (singular) {
      if (tid == 0 && *info == 0)
        *info = j + 1;
      __syncthreads();
      continue; // leave the column untouched; the caller repairs the basis
    }
    // swap rows j and piv inside the panel, cache row j of U
    if (tid < nb) {
      double *c = A + (size_t)(j0 + tid) * ld;
      double a = c[j], b = c[piv];
      if (piv != j) {
        c[j] = b;
        c[piv] = a;
        a = b;
      }
      urow[tid] = a; // row j of the panel after the swap
    }
    __syncthreads();
    const double inv = 1.0 / urow[jj];
    for (int i = j + 1 + tid; i < k; i += 1024) {
      double l = colj[i] * inv;
      colj[i] = l;
      for (int c = jj + 1; c < nb; c++) {
        double *cc = A + (size_t)(j0 + c) * ld;
        cc[i] = fma(-l, urow[c], cc[i]);
      }
    }
    __syncthreads();
  }
}

// apply the panel's row interchanges to columns [c0,c1) (thread per column)
__global__ void lu_swap_kernel(double *__restrict__ A, int ld, int j0, int nb,
                               const int *__restrict__ ipiv, int c0, int c1)
{
  int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c1)
    return;
  double *col = A + (size_t)c * ld;
  for (int jj = 0; jj < nb; jj++) {
    int j = j0 + jj, p = ipiv[j];
    if (p != j) {
      double a = col[j];
      col[j] = col[p];
      col[p] = a;
    }
  }
}

// B[j0..j0+nb, c] := T^-1 B[.., c] for columns c in [c0,c1); T = nb x nb triangle of A at (j0,j0)
//   lower=true : unit lower triangular (forward substitution)
//   lower=false: upper triangular with diagonal (backward substitution)
__global__ void __launch_bounds__(128)
    trsm_kernel(const double *__restrict__ A, int lda, double *__restrict__ B, int ldb, int j0,
                int nb, int c0, int c1, bool lower)
{
  __shared__ double T[NB][NB + 1];
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    int i = e % NB, l = e / NB;
    T[i][l] = (i < nb && l < nb) ? A[(size_t)(j0 + l) * lda + j0 + i] : (i == l ? 1.0 : 0.0);
  }
  __syncthreads();
  int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c1)
    return;
  double *col = B + (size_t)c * ldb + j0;
  double x[NB];
#pragma unroll
  for (int i = 0; i < NB; i++)
    x[i] = i < nb ? col[i] : 0.0;
  if (lower) {
#pragma unroll
    for (int i = 1; i < NB; i++) {
      double s = x[i];
#pragma unroll
      for (int l = 0; l < i; l++)
        s = fma(-T[i][l], x[l], s);
      x[i] = s;
    }
  } else {
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
      double s = x[i];
#pragma unroll
      for (int l = i + 1; l < NB; l++)
        s = fma(-T[i][l], x[l], s);
      x[i] = s / T[i][i];
    }
  }
#pragma unroll
  for (int i = 0; i < NB; i++)
    if (i < nb)
      col[i] = x[i];
}

// C[M x N] -= A[M x K] * B[K x N]  (column major).  64x64 tile per CTA, 4x4 per thread.
__global__ void __launch_bounds__(256)
    gemm_sub_kernel(double *__restrict__ C, int ldc, const double *__restrict__ A, int lda,
                    const double *__restrict__ B, int ldb, int M, int N, int K)
{
  __shared__ double As[16][64 + 1];
  __shared__ double Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++)
      acc[a][b] = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    // A tile: 64 rows x 16 cols ; B tile: 16 rows x 64 cols
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      int i = e & 63, l = e >> 6;
      int gi = m0 + i, gl = k0 + l;
      As[l][i] = (gi < M && gl < K) ? A[(size_t)gl * lda + gi] : 0.0;
    }
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      int l = e & 15, jn = e >> 4;
      int gl = k0 + l, gj = n0 + jn;
      Bs[l][jn] = (gl < K && gj < N) ? B[(size_t)gj * ldb + gl] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < 16; l++) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; a++)
        av[a] = As[l][tx + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; b++)
        bv[b] = Bs[l][ty + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
          acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int b = 0; b < 4; b++) {
    int gj = n0 + ty + 16 * b;
    if (gj >= N)
      continue;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      int gi = m0 + tx + 16 * a;
      if (gi < M)
        C[(size_t)gj * ldc + gi] -= acc[a][b];
    }
  }
}

static void gemm_sub(double *C, int ldc, const double *A, int lda, const double *B, int ldb, int M,
                     int N, int K, cudaStream_t s)
{
  if (M <= 0 || N <= 0 || K <= 0)
    return;
  dim3 grid((M + 63) / 64, (N + 63) / 64);
  gemm_sub_kernel<<<grid, 256, 0, s>>>(C, ldc, A, lda, B, ldb, M, N, K);
}

__global__ void set_permuted_identity_kernel(double *__restrict__ X, int k, int ld,
                                             const int *__restrict__ perm)
{
  // X = P : row i of P has its one in column perm[i]
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)ld * k;
  if (idx >= total)
    return;
  int i = (int)(idx % ld), c = (int)(idx / ld);
  X[idx] = (i < k && perm[i] == c) ? 1.0 : 0.0;
}

__global__ void transpose_kernel(const double *__restrict__ src, double *__restrict__ dst, int k,
                                 int ld)
{
  __shared__ double tile[32][33];
  int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int y = y0 + r;
    tile[r][threadIdx.x] = (x < k && y < k) ? src[(size_t)y * ld + x] : 0.0;
  }
  __syncthreads();
  int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int yo = yo0 + r;
    if (xo < ld && yo < k)
      dst[(size_t)yo * ld + xo] = (xo < k) ? tile[threadIdx.x][r] : 0.0;
  }
}
void launch_transpose(const double *src, double *dst, int k, int ld, cudaStream_t s)
{
  if (k <= 0)
    return;
  dim3 grid((ld + 31) / 32, (k + 31) / 32);
  transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(src, dst, k, ld);
}

// zero the padding columns/rows [k, ld) of every row so GEMVs can run over ld
__global__ void zero_padding_kernel(double *__restrict__ M, int k, int ld)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k)
    return;
  for (int j = k; j < ld; j++)
    M[(size_t)i * ld + j] = 0.0;
}

/* In: A = k x k matrix, column major, leading dimension ld (device).  Out: X = A^-1 (same
   layout), A is overwritten by its LU factors.  hostIpiv/hostPerm: host scratch of k ints,
   dIpiv/dPerm/dInfo: device scratch.  Returns 0 or (1 + index of the first column without an
   acceptable pivot).  Synchronizes the stream once (pivot vector -> permutation). */
int dense_invert(double *A, double *X, int k, int ld, int *dIpiv, int *dPerm, int *dInfo,
                 int *hostIpiv, int *hostPerm, double singularTol, cudaStream_t s)
{
  if (k <= 0)
    return 0;
  cudaMemsetAsync(dInfo, 0, sizeof(int), s);
  for (int j0 = 0; j0 < k; j0 += NB) {
    int nb = k - j0 < NB ? k - j0 : NB;
    lu_panel_kernel<<<1, 1024, 0, s>>>(A, k, ld, j0, nb, dIpiv, dInfo, singularTol);
    // interchanges on the columns left and right of the panel
    if (j0 > 0)
      lu_swap_kernel<<<(j0 + 127) / 128, 128, 0, s>>>(A, ld, j0, nb, dIpiv, 0, j0);
    int c0 = j0 + nb;
    if (c0 < k) {
      lu_swap_kernel<<<(k - c0 + 127) / 128, 128, 0, s>>>(A, ld, j0, nb, dIpiv, c0, k);
      trsm_kernel<<<(k - c0 + 127) / 128, 128, 0, s>>>(A, ld, A, ld, j0, nb, c0, k, true);
      gemm_sub(A + (size_t)c0 * ld + c0, ld, A + (size_t)j0 * ld + c0, ld,
               A + (size_t)c0 * ld + j0, ld, k - c0, k - c0, nb, s);
    }
  }
  int info = 0;
  cudaMemcpyAsync(hostIpiv, dIpiv, sizeof(int) * k, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(&info, dInfo, sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  if (info != 0)
    return info;
  for (int i = 0; i < k; i++)
    hostPerm[i] = i;
  for (int j = 0; j < k; j++) {
    int p = hostIpiv[j];
    if (p != j) {
      int t = hostPerm[j];
      hostPerm[j] = hostPerm[p];
      hostPerm[p] = t;
    }
  }
  cudaMemcpyAsync(dPerm, hostPerm, sizeof(int) * k, cudaMemcpyHostToDevice, s);
  {
    size_t total = (size_t)ld * k;
    set_permuted_identity_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(X, k, ld, dPerm);
  }
  // forward: X := L^-1 X
  for (int j0 = 0; j0 < k; j0 += NB) {
    int nb = k - j0 < NB ? k - j0 : NB;
    trsm_kernel<<<(k + 127) / 128, 128, 0, s>>>(A, ld, X, ld, j0, nb, 0, k, true);
    int r0 = j0 + nb;
    if (r0 < k)
      gemm_sub(X + r0, ld, A + (size_t)j0 * ld + r0, ld, X + j0, ld, k - r0, k, nb, s);
  }
  // backward: X := U^-1 X
  int lastBlock = ((k - 1) / NB) * NB;
  for (int j0 = lastBlock; j0 >= 0; j0 -= NB) {
    int nb = k - j0 < NB ? k - j0 : NB;
    trsm_kernel<<<(k + 127) / 128, 128, 0, s>>>(A, ld, X, ld, j0, nb, 0, k, false);
    if (j0 > 0)
      gemm_sub(X, ld, A + (size_t)j0 * ld, ld, X + j0, ld, j0, k, nb, s);
  }
  zero_padding_kernel<<<(k + 255) / 256, 256, 0, s>>>(X, k, ld);
  return 0;
}

} // namespace clpb
