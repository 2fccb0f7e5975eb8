/*
 * C restatement of the LDS + GPC rollout (TEST INFRASTRUCTURE / CPU baseline; see oracle/__init__.py).
 *
 * Parity pinned to the reference's own files run here: reproduces tests/golden/ref_lds.npz (deluca/agents/_gpc.py +
 * deluca/envs/_lds.py executed literally on tests/refshim) to 1e-9 in double
 * (tests/test_ref_fixtures_oracle.py::test_c_oracle_equals_reference_file), and agrees with oracle/lds.py (the
 * line-by-line NumPy restatement) in tests/test_oracle_lds.py.  The reference ships no test of its own for LDS/GPC.
 *
 * Follows, per env and per step t:
 *   u = -K x + sum_i M[i] hist[HH+i]                       deluca/agents/_gpc.py:171-183
 *   noise = x - A x_prev - B u ; hist <- roll(-1)           _gpc.py:155-157
 *   dM = grad_M policy_loss(M, hist)                        _gpc.py:109-128 (hand adjoint, SURVEY A-3.5)
 *   M -= lr_scale/(t+1) * dM                                _gpc.py:161-163
 *   cost = x'x + u'u                                        _gpc.py:29-40
 *   x <- A x + B u + w_t                                    deluca/envs/_lds.py:102-111
 * Plain scalar loops over envs, OpenMP across envs (`REAL` = float or double).  Requires HH >= H-1.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL float
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#ifndef SUFFIX
#define SUFFIX f32
#endif

int CAT(lds_gpc_max_threads_, SUFFIX)(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* A[N,d,d] B[N,d,m] K[N,m,d] x[N,d] (in/out) W[T,N,d] costs[T,N] (out); fresh agent state (M = 0, hist = 0). */
int CAT(lds_gpc_rollout_, SUFFIX)(int64_t N, int T, int d, int m, int H, int HH, const REAL* A, const REAL* B,
                                  const REAL* K, REAL* x_io, const REAL* W, REAL* costs, double lr_scale,
                                  int decay, int nthreads) {
  if (HH < H - 1 || d > 64 || m > 16 || H > 32) return -1;
  const int P = H + HH;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    REAL* M = (REAL*)malloc(sizeof(REAL) * H * m * d);
    REAL* hist = (REAL*)malloc(sizeof(REAL) * P * d);
    REAL x[64], xprev[64], y[64], yn[64], lam[64], ln[64], u[16], v[16], uf[16], du[16], lu[16], noise[64];
#pragma omp for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
      const REAL* An = A + n * d * d;
      const REAL* Bn = B + n * d * m;
      const REAL* Kn = K + n * m * d;
      memset(M, 0, sizeof(REAL) * H * m * d);
      memset(hist, 0, sizeof(REAL) * P * d);
      for (int i = 0; i < d; ++i) { x[i] = x_io[n * d + i]; xprev[i] = 0; }
      for (int t = 0; t < T; ++t) {
        /* action */
        for (int c = 0; c < m; ++c) {
          REAL s = 0, mw = 0;
          for (int j = 0; j < d; ++j) s += Kn[c * d + j] * x[j];
          for (int i = 0; i < H; ++i)
            for (int j = 0; j < d; ++j) mw += M[(i * m + c) * d + j] * hist[(HH + i) * d + j];
          u[c] = -s + mw;
        }
        /* noise, ring */
        for (int i = 0; i < d; ++i) {
          REAL s = 0, sb = 0;
          for (int j = 0; j < d; ++j) s += An[i * d + j] * xprev[j];
          for (int c = 0; c < m; ++c) sb += Bn[i * m + c] * u[c];
          noise[i] = x[i] - s - sb;
        }
        memmove(hist, hist + d, sizeof(REAL) * (P - 1) * d);
        memcpy(hist + (P - 1) * d, noise, sizeof(REAL) * d);
        /* policy_loss forward */
        for (int i = 0; i < d; ++i) y[i] = 0;
        for (int h = 0; h < H - 1; ++h) {
          for (int c = 0; c < m; ++c) {
            REAL s = 0, ky = 0;
            for (int i = 0; i < H; ++i)
              for (int j = 0; j < d; ++j) s += M[(i * m + c) * d + j] * hist[(h + i) * d + j];
            for (int j = 0; j < d; ++j) ky += Kn[c * d + j] * y[j];
            v[c] = s - ky;
          }
          for (int i = 0; i < d; ++i) {
            REAL s = 0, sb = 0;
            for (int j = 0; j < d; ++j) s += An[i * d + j] * y[j];
            for (int c = 0; c < m; ++c) sb += Bn[i * m + c] * v[c];
            yn[i] = (s + sb) + hist[(h + H) * d + i];
          }
          for (int i = 0; i < d; ++i) y[i] = yn[i];
        }
        for (int c = 0; c < m; ++c) {
          REAL s = 0, ky = 0;
          for (int i = 0; i < H; ++i)
            for (int j = 0; j < d; ++j) s += M[(i * m + c) * d + j] * hist[(HH - 1 + i) * d + j];
          for (int j = 0; j < d; ++j) ky += Kn[c * d + j] * y[j];
          uf[c] = s - ky;
          du[c] = 2 * uf[c];
        }
        /* adjoint, M updated in place (the adjoint never reads M) */
        const REAL lr = (REAL)(decay ? lr_scale * (1.0 / (t + 1)) : lr_scale);
        for (int j = 0; j < d; ++j) {
          REAL s = 2 * y[j];
          for (int c = 0; c < m; ++c) s -= Kn[c * d + j] * du[c];
          lam[j] = s;
        }
        for (int i = 0; i < H; ++i)
          for (int c = 0; c < m; ++c)
            for (int j = 0; j < d; ++j) M[(i * m + c) * d + j] -= lr * (du[c] * hist[(HH - 1 + i) * d + j]);
        for (int h = H - 2; h >= 0; --h) {
          for (int c = 0; c < m; ++c) {
            REAL s = 0;
            for (int i = 0; i < d; ++i) s += Bn[i * m + c] * lam[i];
            lu[c] = s;
          }
          for (int i = 0; i < H; ++i)
            for (int c = 0; c < m; ++c)
              for (int j = 0; j < d; ++j) M[(i * m + c) * d + j] -= lr * (lu[c] * hist[(h + i) * d + j]);
          for (int j = 0; j < d; ++j) {
            REAL s = 0;
            for (int i = 0; i < d; ++i) s += An[i * d + j] * lam[i];
            for (int c = 0; c < m; ++c) s -= Kn[c * d + j] * lu[c];
            ln[j] = s;
          }
          for (int j = 0; j < d; ++j) lam[j] = ln[j];
        }
        /* cost, env step */
        REAL cost = 0;
        for (int i = 0; i < d; ++i) cost += x[i] * x[i];
        for (int c = 0; c < m; ++c) cost += u[c] * u[c];
        costs[(int64_t)t * N + n] = cost;
        for (int i = 0; i < d; ++i) xprev[i] = x[i];
        for (int i = 0; i < d; ++i) {
          REAL s = 0, sb = 0;
          for (int j = 0; j < d; ++j) s += An[i * d + j] * xprev[j];
          for (int c = 0; c < m; ++c) sb += Bn[i * m + c] * u[c];
          yn[i] = (s + sb) + W[((int64_t)t * N + n) * d + i];
        }
        for (int i = 0; i < d; ++i) x[i] = yn[i];
      }
      for (int i = 0; i < d; ++i) x_io[n * d + i] = x[i];
    }
    free(M);
    free(hist);
  }
  return 0;
}
