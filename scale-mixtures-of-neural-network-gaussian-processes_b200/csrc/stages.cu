// Stage-level C-ABI (declared in include/smnngp.h): the same kernels as the fused entry points, exposed one
// pipeline stage at a time so the multi-GPU driver (distributed.py, one process per GPU, torch.distributed /
// NCCL for the exchange steps) can interleave them with collectives.  Every function only enqueues on `stream`.
#include "../../include/smnngp.h"

#include <string>

#include "context.cuh"
#include "kernels.cuh"

using namespace smnngp;

namespace {
int fail_stage(cudaError_t e) { return e == cudaSuccess ? SMNNGP_OK : SMNNGP_ECUDA; }
}  // namespace

extern "C" {

int smnngp_stage_qtable_f64(void* stream, const double* X, int64_t N, int64_t D, int n_hidden, int act, int arch,
                            const double* hp_dev, double* tab, int64_t tab_ld, double* q, double* scal) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!X || !tab || !q || !hp_dev || N <= 0 || D <= 0) return SMNNGP_EINVAL;
  cudaError_t e = launch_qtable(s, X, D, (int)N, (int)D, n_hidden, act, arch, hp_dev, tab, tab_ld, q);
  if (e != cudaSuccess) return SMNNGP_ECUDA;
  if (scal) e = launch_scalars(s, q, (int)N, hp_dev, scal);
  return fail_stage(e);
}

int smnngp_stage_gram_f64(void* stream, const double* X1, int64_t n1, const double* X2, int64_t n2, int64_t D,
                          int n_hidden, int act, int arch, const double* hp_dev, const double* tab1,
                          int64_t tab_ld1, const double* tab2, int64_t tab_ld2, const double* scal, int shift,
                          int symmetric_lower, double* K, int64_t ldk) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!X1 || !X2 || !K || !hp_dev || !tab1 || !tab2 || n1 < 0 || n2 < 0 || D <= 0) return SMNNGP_EINVAL;
  GramParams g{};
  g.X1 = X1; g.X2 = X2; g.ld1 = D; g.ld2 = D; g.N = (int)n1; g.M = (int)n2; g.D = (int)D;
  g.tab1 = tab1; g.tab2 = tab2; g.tab_ld1 = tab_ld1; g.tab_ld2 = tab_ld2;
  g.n_hidden = n_hidden; g.act = act; g.arch = arch; g.hp = hp_dev; g.scal = scal;
  g.shift = symmetric_lower ? shift : SHIFT_NONE;
  g.symmetric = symmetric_lower ? 1 : 0; g.out_full = symmetric_lower ? 0 : 1; g.K = K; g.ldk = ldk;
  return fail_stage(launch_gram(s, g));
}

// factor the w x w diagonal block (w <= any multiple of 128) in place, keep inv(L) of every 128-block
int smnngp_stage_factor_diag_f64(void* stream, double* A, int64_t lda, int64_t w, double* linv_blocks,
                                 double* logdet_dev, int* info_dev, int64_t gcol0) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!A || !linv_blocks || !logdet_dev || !info_dev || w <= 0) return SMNNGP_EINVAL;
  (void)gcol0;
  return fail_stage(potrf_trapezoid(s, A, lda, w, w, (int)((w + PB - 1) / PB * PB), linv_blocks, logdet_dev, info_dev,
                                    (long long)PB * PB));   // single panel: serial path, one inverse block per 128 columns
}

// R [m, w] <- R * L^-T by 128-block substitution with the factor L [w, w] (ldl) and its block inverses
int smnngp_stage_trsm_f64(void* stream, double* R, int64_t ldr, int64_t m, int64_t w, const double* L, int64_t ldl,
                          const double* linv_blocks) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!R || !L || !linv_blocks || m < 0 || w <= 0) return SMNNGP_EINVAL;
  if (m == 0) return SMNNGP_OK;
  for (int64_t j0 = 0; j0 < w; j0 += PB) {
    const int64_t j1 = (j0 + PB < w) ? j0 + PB : w;
    GemmParams t{};
    t.A = R + j0; t.lda = ldr;
    t.B = linv_blocks + (j0 / PB) * PB * PB; t.ldb = PB;
    t.C = R + j0; t.ldc = ldr;
    t.M = (int)m; t.N = (int)(j1 - j0); t.K = (int)(j1 - j0); t.lower = 0;
    cudaError_t e = launch_gemm_store(s, t);
    if (e != cudaSuccess) return SMNNGP_ECUDA;
    if (j1 < w) {
      GemmParams u{};
      u.A = R + j0; u.lda = ldr;
      u.B = L + j1 * ldl + j0; u.ldb = ldl;
      u.C = R + j1; u.ldc = ldr;
      u.M = (int)m; u.N = (int)(w - j1); u.K = (int)(j1 - j0); u.lower = 0;
      e = launch_gemm_sub(s, u);
      if (e != cudaSuccess) return SMNNGP_ECUDA;
    }
  }
  return SMNNGP_OK;
}

// C [M, N] -= A [M, K] * B [N, K]^T with the block-row-cyclic lower mask (cyc_db = 0: plain lower / no mask);
// cyc_alt: extra column shift of the rows of odd local blocks (snake distribution, see GemmParams)
int smnngp_stage_update2_f64(void* stream, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                            int64_t ldc, int64_t M, int64_t N, int64_t K, int lower, int64_t cyc_db, int64_t cyc_p,
                            int64_t base_shift, int64_t cyc_alt, int sm_reserve) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Enter scope(s);
  if (!A || !B || !C || M < 0 || N < 0 || K <= 0) return SMNNGP_EINVAL;
  GemmParams u{};
  u.A = A; u.lda = lda; u.B = B; u.ldb = ldb; u.C = C; u.ldc = ldc;
  u.M = (int)M; u.N = (int)N; u.K = (int)K; u.lower = lower;
  u.cyc_db = (int)cyc_db; u.cyc_p = (int)cyc_p; u.base_shift = (int)base_shift; u.cyc_alt = (int)cyc_alt;
  u.sm_reserve = sm_reserve > 0 ? sm_reserve : 0;
  const bool timed = instr().time_updates && K >= 256;      // outer trailing updates only
  if (timed) {
    double pairs = 0.0;                                      // active (row, col) pairs of this rank's part
    for (int64_t r = 0; r < M; r++) {
      long long lim = diag_limit(u, (int)r) + 1;
      pairs += (double)(lim < N ? (lim > 0 ? lim : 0) : N);
    }
    instr_begin_update(s, 2.0 * pairs * (double)K);
  }
  cudaError_t e = launch_gemm_sub(s, u);
  if (timed) instr_end_update(s);
  return fail_stage(e);
}

int smnngp_stage_update_f64(void* stream, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                            int64_t ldc, int64_t M, int64_t N, int64_t K, int lower, int64_t cyc_db, int64_t cyc_p,
                            int64_t base_shift, int sm_reserve) {
  return smnngp_stage_update2_f64(stream, A, lda, B, ldb, C, ldc, M, N, K, lower, cyc_db, cyc_p, base_shift, 0, sm_reserve);
}

// predictive tail for carried rows: V [T, ldv] = K_td L^-T rows, Z [C, ldz] = (L^-1 Y)^T rows, ktt [T] prior variances
// -> mean [T, C] = V Z^T, var [T] = ktt - ||v||^2   (NaN when *info_dev != 0)
int smnngp_stage_predict_finalize_f64(void* stream, const double* V, int64_t ldv, const double* Z, int64_t ldz,
                                      const double* ktt, int64_t T, int64_t C, int64_t N, const int* info_dev,
                                      double* mean, double* var) {
  if (!V || !Z || !ktt || !info_dev || !mean || !var || T < 0 || C <= 0 || N <= 0 || T > INT32_MAX || C > INT32_MAX)
    return SMNNGP_EINVAL;
  if (T == 0) return SMNNGP_OK;
  Enter scope(static_cast<cudaStream_t>(stream));
  return fail_stage(launch_predict_finalize(static_cast<cudaStream_t>(stream), V, ldv, Z, ldz, ktt, (int)T, (int)C, N,
                                            info_dev, mean, var));
}

// SPR.test_nll tail (spax/models.py:114-119, spax/likelihoods.py:30-33 / :52-65) on gathered predictive moments:
// quad2_dev = ||L2^-1 y||^2 of the second factorisation K + 1e-6 (alpha/beta) I (Student-t only; may be NULL for gauss)
int smnngp_stage_test_nll_finalize_f64(void* stream, const double* mean, const double* var, const double* ytest,
                                       int64_t T, int64_t N, double y_mean, double y_std, const double* hp_dev, int kind,
                                       const double* quad2_dev, const int* info_dev, double* logp, double* nll_out_dev) {
  if (!mean || !var || !ytest || !hp_dev || !info_dev || !nll_out_dev || T <= 0 || T > INT32_MAX ||
      (kind != KIND_GAUSS && kind != KIND_STUDENT_T) || (kind == KIND_STUDENT_T && !quad2_dev))
    return SMNNGP_EINVAL;
  Enter scope(static_cast<cudaStream_t>(stream));
  return fail_stage(launch_test_nll_finalize(static_cast<cudaStream_t>(stream), mean, var, ytest, (int)T, N, y_mean,
                                             y_std, hp_dev, kind, quad2_dev ? quad2_dev : mean, info_dev, logp,
                                             nll_out_dev));
}

int smnngp_stage_sumsq_f64(void* stream, const double* z, int64_t n, double* out_dev) {
  if (!z || !out_dev || n < 0) return SMNNGP_EINVAL;
  Enter scope(static_cast<cudaStream_t>(stream));
  return fail_stage(launch_sumsq(static_cast<cudaStream_t>(stream), z, n, out_dev));
}

// sums[2] = { sum log L_ii, ||L^-1 y||^2 } -> out_dev[4] as in smnngp_lml_f64
int smnngp_stage_lml_finalize_f64(void* stream, const double* sums_dev, const double* hp_dev, int kind, int64_t N,
                                  const int* info_dev, double* out_dev) {
  if (!sums_dev || !hp_dev || !info_dev || !out_dev) return SMNNGP_EINVAL;
  Enter scope(static_cast<cudaStream_t>(stream));
  // lml_finalize reads scal[SC_LOGDET], scal[SC_QUAD]: present the two sums at those offsets
  return fail_stage(launch_lml_finalize(static_cast<cudaStream_t>(stream), sums_dev - SC_LOGDET, hp_dev, kind, N,
                                        info_dev, out_dev));
}

}  // extern "C"
