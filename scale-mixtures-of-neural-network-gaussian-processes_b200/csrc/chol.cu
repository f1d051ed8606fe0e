// Blocked right-looking FP64 Cholesky on the row-major lower triangle (replaces lax.linalg.cholesky at
// spax/utils.py:179 and the cho_factor inside neural_tangents' predict, spax/kernels.py:29-32).
//   diagonal block : one-CTA fused potf2 + triangular inverse in shared memory
//   TRSM           : panel <- panel * inv(L_kk)^T on the DMMA GEMM core (in place, one column tile)
//   trailing update: C -= P * P^T on the DMMA GEMM core, lower tiles only
// Rows below the square part (the trapezoid) are carried through TRSM + update, so appended rows y^T and K_td
// come out as (L^-1 y)^T and (L^-1 K_dt)^T: the triangular solves of spax/utils.py:180 and of cho_solve are
// fused into the factorisation and run on the tensor pipe.
#include "gemm_core.cuh"
#include "kernels.cuh"

namespace smnngp {

namespace {

constexpr int EPI_STORE = 0, EPI_SUB = 1;

template <bool ALIGN16, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_kernel(const GemmParams p) {
  extern __shared__ __align__(16) double smem[];
  const int ntn = (p.N + BN - 1) / BN;
  int ti, tj;
  decode_tile(blockIdx.x, ntn, p.lower, ti, tj);
  const int r0 = ti * BM, c0 = tj * BN;
  double acc[MI][NI][2];
  gemm_mainloop<ALIGN16>(acc, p.A + (long long)r0 * p.lda, p.lda, min(BM, p.M - r0),
                         p.B + (long long)c0 * p.ldb, p.ldb, min(BN, p.N - c0), p.K, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rbase = r0 + (warp >> 2) * 64 + (lane >> 2);
  const int cbase = c0 + (warp & 3) * 32 + (lane & 3) * 2;
  const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
  for (int mi = 0; mi < MI; mi++) {
    const int r = rbase + mi * 8;
    if (r >= p.M) continue;
#pragma unroll
    for (int ni = 0; ni < NI; ni++) {
      const int c = cbase + ni * 8;
      const bool ok0 = c < p.N && (!p.lower || c <= r);
      const bool ok1 = (c + 1) < p.N && (!p.lower || (c + 1) <= r);
      double* dst = p.C + (long long)r * p.ldc + c;
      if (ok0 && ok1 && vec_ok) {
        double2* d2 = reinterpret_cast<double2*>(dst);
        if (EPI == EPI_STORE) {
          *d2 = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        } else {
          double2 v = *d2;
          v.x -= acc[mi][ni][0];
          v.y -= acc[mi][ni][1];
          *d2 = v;
        }
      } else {
        if (ok0) dst[0] = (EPI == EPI_STORE) ? acc[mi][ni][0] : dst[0] - acc[mi][ni][0];
        if (ok1) dst[1] = (EPI == EPI_STORE) ? acc[mi][ni][1] : dst[1] - acc[mi][ni][1];
      }
    }
  }
}

template <int EPI>
cudaError_t launch_gemm_t(cudaStream_t s, const GemmParams& p) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  long long tiles = count_tiles(p.M, p.N, p.lower);
  bool a16 = (p.lda % 2 == 0) && (p.ldb % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);
  auto kern = a16 ? gemm_kernel<true, EPI> : gemm_kernel<false, EPI>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)tiles, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Diagonal block: Cholesky + inverse of the factor in one in-place elimination.
// State G (lower, shared memory).  At step j the pivot vector v holds  v[c<j] = W[j][c] (partial inverse
// row), v[j] = d (pivot), v[c>j] = S[c][j] (current column).  Every row i > j then does
//   G[i][c] -= v[i] * v[c] / d      (c <= i, c != j),      G[i][j] = -v[i] / d
// which is simultaneously the Schur update of the trailing matrix (c > j) and the forward substitution for
// inv(L) (c < j).  L^T is parked in the unused upper triangle, inv(L) ends up in the lower triangle.
// One __syncthreads per column.
// ---------------------------------------------------------------------------------------------------------
constexpr int PF_THREADS = 512;
constexpr int GLD = PB + 1;
constexpr int PF_SMEM_BYTES = (PB * GLD + 3 * PB + 32) * 8;

__global__ void __launch_bounds__(PF_THREADS, 1)
potf2_trtri_kernel(double* __restrict__ A, long long lda, int w, double* __restrict__ Linv,
                   double* __restrict__ logdet, int* __restrict__ info, int gcol0) {
  extern __shared__ __align__(16) double sm[];
  double* G = sm;
  double* v0 = G + PB * GLD;
  double* v1 = v0 + PB;
  double* dl = v1 + PB;
  double* red = dl + PB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int idx = tid; idx < PB * PB; idx += PF_THREADS) {
    int i = idx >> 7, c = idx & (PB - 1);
    double g = 0.0;
    if (i < w && c <= i) g = A[(long long)i * lda + c];
    else if (i == c) g = 1.0;
    G[i * GLD + c] = g;
  }
  if (tid < PB) { dl[tid] = 1.0; v1[tid] = 0.0; }
  __syncthreads();
  if (tid < PB) v0[tid] = G[tid * GLD];
  __syncthreads();

  int bad = 0;
  for (int j = 0; j < w; j++) {
    const double* vb = (j & 1) ? v1 : v0;
    double* vn = (j & 1) ? v0 : v1;
    double d = vb[j];
    if (!(d > 0.0)) {               // non-PD (or NaN input): mirror lax.linalg.cholesky -> NaN, no abort
      if (bad == 0) bad = j + 1;
      d = __longlong_as_double(0x7ff8000000000000ll);
    }
    const double dinv = 1.0 / d;
    if (tid < PB) {
      const double r = sqrt(d);
      const double rinv = 1.0 / r;
      if (tid == j) {
        dl[j] = r;
        G[j * GLD + j] = rinv;
      } else {
        G[j * GLD + tid] = vb[tid] * rinv;   // tid > j: L[tid][j] (stored transposed); tid < j: inv(L)[j][tid]
      }
    }
    for (int i = j + 1 + warp; i < w; i += PF_THREADS / 32) {
      const double f = vb[i] * dinv;
      double* gi = G + i * GLD;
      for (int c = lane; c <= i; c += 32) {
        double g = (c == j) ? -f : gi[c] - f * vb[c];
        gi[c] = g;
        if (i == j + 1) vn[c] = g;
        else if (c == j + 1) vn[i] = g;
      }
    }
    __syncthreads();
  }

  for (int idx = tid; idx < PB * PB; idx += PF_THREADS) {
    int i = idx >> 7, c = idx & (PB - 1);
    if (i < w && c <= i) A[(long long)i * lda + c] = (c == i) ? dl[i] : G[c * GLD + i];
    Linv[idx] = (c <= i) ? G[i * GLD + c] : 0.0;
  }
  // sum of log L_ii in a fixed order
  double lg = (tid < w) ? log(dl[tid]) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
  if (lane == 0) red[warp] = lg;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int k = 0; k < PB / 32; k++) s += red[k];
    *logdet += s;
    if (bad != 0 && *info == 0) *info = gcol0 + bad;
  }
}

}  // namespace

cudaError_t launch_gemm_store(cudaStream_t s, const GemmParams& p) { return launch_gemm_t<EPI_STORE>(s, p); }
cudaError_t launch_gemm_sub(cudaStream_t s, const GemmParams& p) { return launch_gemm_t<EPI_SUB>(s, p); }

cudaError_t launch_potf2_trtri(cudaStream_t s, double* A, long long lda, int w, double* Linv, double* logdet,
                               int* info, int global_col0) {
  cudaError_t e = cudaFuncSetAttribute(potf2_trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       PF_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  potf2_trtri_kernel<<<1, PF_THREADS, PF_SMEM_BYTES, s>>>(A, lda, w, Linv, logdet, info, global_col0);
  return cudaGetLastError();
}

cudaError_t potrf_trapezoid(cudaStream_t s, double* A, long long lda, long long Mtot, long long N, int NB,
                            double* Linv_ws, double* logdet, int* info) {
  if (NB < PB) NB = PB;
  NB = (NB / PB) * PB;
  cudaError_t e;
  for (long long c0 = 0; c0 < N; c0 += NB) {
    const long long c1 = (c0 + NB < N) ? c0 + NB : N;
    for (long long j0 = c0; j0 < c1; j0 += PB) {
      const long long j1 = (j0 + PB < N) ? j0 + PB : N;
      const int w = (int)(j1 - j0);
      e = launch_potf2_trtri(s, A + j0 * lda + j0, lda, w, Linv_ws, logdet, info, (int)j0);
      if (e != cudaSuccess) return e;
      if (Mtot > j1) {
        GemmParams t{};   // panel rows <- panel rows * inv(L_jj)^T   (in place: one CTA owns whole rows)
        t.A = A + j1 * lda + j0; t.lda = lda;
        t.B = Linv_ws;           t.ldb = PB;
        t.C = A + j1 * lda + j0; t.ldc = lda;
        t.M = (int)(Mtot - j1); t.N = w; t.K = w; t.lower = 0;
        e = launch_gemm_store(s, t);
        if (e != cudaSuccess) return e;
        if (j1 < c1) {    // remaining columns of the outer panel
          GemmParams u{};
          u.A = A + j1 * lda + j0; u.lda = lda;
          u.B = A + j1 * lda + j0; u.ldb = lda;
          u.C = A + j1 * lda + j1; u.ldc = lda;
          u.M = (int)(Mtot - j1); u.N = (int)(c1 - j1); u.K = w; u.lower = 1;
          e = launch_gemm_sub(s, u);
          if (e != cudaSuccess) return e;
        }
      }
    }
    if (c1 < N) {         // trailing update with the whole outer panel, K = NB
      GemmParams u{};
      u.A = A + c1 * lda + c0; u.lda = lda;
      u.B = A + c1 * lda + c0; u.ldb = lda;
      u.C = A + c1 * lda + c1; u.ldc = lda;
      u.M = (int)(Mtot - c1); u.N = (int)(N - c1); u.K = (int)(c1 - c0); u.lower = 1;
      e = launch_gemm_sub(s, u);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

}  // namespace smnngp
