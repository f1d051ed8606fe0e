/* CPU restatement (plain C + OpenMP) of the NNGP layer recursion - TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Same arithmetic as oracle/nngp_oracle.py::_recursion (which restates experiments/nt_kernels.py:21-31 and
 * :83-103 with neural_tangents' Dense / Relu / Erf kernel maps).  It exists because the NumPy version is
 * single-threaded and allocation-bound, whereas the reference's JAX CPU backend would fuse the element-wise
 * layers and spread them over the host cores; this file gives the "port" CPU baseline the same advantage.
 * PARITY UNPINNED: see the header of nngp_oracle.py.  Never linked into the product library.
 *
 * In-place on k [n, m] (row-major, ld = m): k holds X1.X2^T / D on entry and the kernel on exit.
 * q1 [n], q2 [m]: input marginal variances |x|^2 / D.  act: 0 relu, 1 erf.  arch: 0 mlp, 1 resnet.
 */
#include <math.h>
#include <stdlib.h>
#include <stdint.h>

static const double PI = 3.14159265358979323846;

static inline double act_diag(double u, int act) {
  return act == 0 ? 0.5 * u : (2.0 / PI) * asin(2.0 * u / (1.0 + 2.0 * u));
}

static inline double act_off(double k, double u1, double u2, int act) {
  if (act == 0) {
    double s = sqrt(fmax(u1 * u2 - k * k, 0.0));
    double th = (s == 0.0 && k == 0.0) ? PI / 2 : atan2(s, k);
    return s / (2.0 * PI) + (0.5 - th / (2.0 * PI)) * k;
  }
  return (2.0 / PI) * asin(2.0 * k / sqrt((1.0 + 2.0 * u1) * (1.0 + 2.0 * u2)));
}

/* per-point pre-activation variances of every nonlinearity application: tab [n_act][n] */
static void layer_table(const double* q, int64_t n, int n_hidden, int act, int arch, double w2, double b2,
                        double* tab) {
  for (int64_t i = 0; i < n; i++) {
    double v = q[i];
    if (arch == 0) {
      for (int a = 0; a < n_hidden; a++) { double u = w2 * v + b2; tab[a * n + i] = u; v = act_diag(u, act); }
    } else {
      double u = w2 * v + b2;
      for (int a = 0; a < n_hidden; a++) { tab[a * n + i] = u; u = u + (w2 * act_diag(u, act) + b2); }
      tab[(int64_t)n_hidden * n + i] = u;
    }
  }
}

int nngp_recursion_inplace(double* k, int64_t n, int64_t m, const double* q1, const double* q2, int n_hidden,
                           int act, int arch, double w_std, double b_std, double last_w_std) {
  const double w2 = w_std * w_std, b2 = b_std * b_std, v2 = last_w_std * last_w_std;
  const int n_act = arch == 1 ? n_hidden + 1 : n_hidden;
  double* t1 = (double*)malloc(sizeof(double) * (size_t)(n_act > 0 ? n_act : 1) * n);
  double* t2 = (double*)malloc(sizeof(double) * (size_t)(n_act > 0 ? n_act : 1) * m);
  if (!t1 || !t2) { free(t1); free(t2); return 1; }
  layer_table(q1, n, n_hidden, act, arch, w2, b2, t1);
  layer_table(q2, m, n_hidden, act, arch, w2, b2, t2);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; i++) {
    double* row = k + i * m;
    for (int64_t j = 0; j < m; j++) {
      double z = row[j];
      if (arch == 0) {
        for (int a = 0; a < n_hidden; a++) z = act_off(w2 * z + b2, t1[a * n + i], t2[a * m + j], act);
      } else {
        z = w2 * z + b2;
        for (int a = 0; a < n_hidden; a++) z = z + (w2 * act_off(z, t1[a * n + i], t2[a * m + j], act) + b2);
        z = act_off(z, t1[(int64_t)n_hidden * n + i], t2[(int64_t)n_hidden * m + j], act);
      }
      row[j] = v2 * z;
    }
  }
  free(t1);
  free(t2);
  return 0;
}
