// Classic env models on the device (deluca/envs/classic): step functions and closed-form / forward-mode Jacobians,
// shared by the igpc kernels (igpc.cu) and the closed-loop env x agent rollout (env_agent.cu).
#pragma once
#include "common.cuh"

namespace dlc {

// ----------------------------------------------------------------------------------- env models
template <int KIND>
struct Model;

template <>
struct Model<DLC_ENV_PENDULUM> {  // deluca/envs/classic/_pendulum.py:54-73 (as written, SURVEY A-6)
  static constexpr int D = 3, M = 1;
  float k1, k2, max_torque, dt;
  __device__ explicit Model(const DlcEnvSpec& s) {
    const float m = s.p[0], l = s.p[1], g = s.p[2];
    max_torque = s.p[3];
    dt = s.p[4];
    k1 = -3.0f * g / (2.0f * l);
    k2 = 3.0f / (m * l * l);
  }
  __device__ void step(const float* x, const float* u, float* xn) const {
    const float th = atan2f(x[0], x[1]);
    const float a = max_torque * tanhf(u[0]);
    const float nthd = th + (k1 * sinf(th + 3.14159265358979323846f) + k2 * a);
    const float nth = th + nthd * dt;
    float s, c;
    sincosf(nth, &s, &c);
    xn[0] = s; xn[1] = c; xn[2] = nthd;
  }
  // Fx[i*D+j] = d xn_i / d x_j, Fu[i*M+c] = d xn_i / d u_c
  __device__ void jac(const float* x, const float* u, float* Fx, float* Fu) const {
    const float s = x[0], co = x[1];
    const float r2 = s * s + co * co;
    const float th = atan2f(s, co);
    const float dth0 = co / r2, dth1 = -s / r2;
    const float tu = tanhf(u[0]);
    const float a = max_torque * tu;
    float sp, cp;
    sincosf(th + 3.14159265358979323846f, &sp, &cp);
    const float nthd = th + (k1 * sp + k2 * a);
    const float nth = th + nthd * dt;
    const float dthd_dth = 1.0f + k1 * cp;
    const float dnth_dth = 1.0f + dt * dthd_dth;
    float sn, cn;
    sincosf(nth, &sn, &cn);
    Fx[0] = cn * dnth_dth * dth0;  Fx[1] = cn * dnth_dth * dth1;  Fx[2] = 0.f;
    Fx[3] = -sn * dnth_dth * dth0; Fx[4] = -sn * dnth_dth * dth1; Fx[5] = 0.f;
    Fx[6] = dthd_dth * dth0;       Fx[7] = dthd_dth * dth1;       Fx[8] = 0.f;
    const float dthd_du = k2 * max_torque * (1.0f - tu * tu);
    Fu[0] = cn * dt * dthd_du; Fu[1] = -sn * dt * dthd_du; Fu[2] = dthd_du;
  }
};

template <>
struct Model<DLC_ENV_CARTPOLE> {  // _cartpole.py:60-89, intent-fixed (SURVEY A-7)
  static constexpr int D = 4, M = 1;
  float a21, a31, b2, b3, dt, offset;
  __device__ explicit Model(const DlcEnvSpec& s) {
    const float m = s.p[0], Mc = s.p[1], l = s.p[2], g = s.p[3];
    dt = s.p[4];
    offset = s.p[5];
    a21 = dt * m * g / Mc;
    a31 = dt * (m + Mc) * g / (Mc * l);
    b2 = dt / Mc;
    b3 = dt / (Mc * l);
  }
  __device__ void step(const float* x, const float* u, float* xn) const {
    const float v = u[0] + offset;
    xn[0] = x[0] + dt * x[2];
    xn[1] = x[1] + dt * x[3];
    xn[2] = (a21 * x[1] + x[2]) + b2 * v;
    xn[3] = (a31 * x[1] + x[3]) + b3 * v;
  }
  __device__ void jac(const float*, const float*, float* Fx, float* Fu) const {
#pragma unroll
    for (int i = 0; i < 16; ++i) Fx[i] = 0.f;
    Fx[0] = 1.f; Fx[2] = dt; Fx[5] = 1.f; Fx[7] = dt; Fx[9] = a21; Fx[10] = 1.f; Fx[13] = a31; Fx[15] = 1.f;
    Fu[0] = 0.f; Fu[1] = 0.f; Fu[2] = b2; Fu[3] = b3;
  }
};

template <>
struct Model<DLC_ENV_QUADROTOR> {  // deluca/envs/classic/_planar_quadrotor.py:92-104 (explicit Euler, wind field)
  static constexpr int D = 6, M = 2;
  float im, g, dt, wind, kth;
  int wind_kind;
  __device__ explicit Model(const DlcEnvSpec& s) {
    const float m = s.p[0], l = s.p[1];
    im = 1.0f / m;
    g = s.p[2];
    dt = s.p[3];
    wind = s.p[4];
    wind_kind = (int)s.p[5];
    kth = l / (m * l * l);
  }
  __device__ void step(const float* x, const float* u, float* xn) const {
    float sn, cs;
    sincosf(x[2], &sn, &cs);
    const float wx = wind_kind == 0 ? wind * x[0] : wind * 0.70710678118654752f;
    const float wy = wind_kind == 0 ? wind * x[1] : wind * 0.70710678118654752f;
    const float tu = u[0] + u[1];
    const float xdd = -tu * sn * im + wx * im;
    const float ydd = tu * cs * im - g + wy * im;
    const float tdd = kth * (u[1] - u[0]);
    xn[0] = x[0] + x[3] * dt; xn[1] = x[1] + x[4] * dt; xn[2] = x[2] + x[5] * dt;
    xn[3] = x[3] + xdd * dt;  xn[4] = x[4] + ydd * dt;  xn[5] = x[5] + tdd * dt;
  }
  __device__ void jac(const float* x, const float* u, float* Fx, float* Fu) const {
    float sn, cs;
    sincosf(x[2], &sn, &cs);
    const float tu = u[0] + u[1];
#pragma unroll
    for (int i = 0; i < 36; ++i) Fx[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) Fx[i * 6 + i] = 1.f;
    Fx[0 * 6 + 3] = dt; Fx[1 * 6 + 4] = dt; Fx[2 * 6 + 5] = dt;
    const float dw = wind_kind == 0 ? wind * im * dt : 0.f;
    Fx[3 * 6 + 0] = dw;
    Fx[3 * 6 + 2] = -tu * cs * im * dt;
    Fx[4 * 6 + 1] = dw;
    Fx[4 * 6 + 2] = -tu * sn * im * dt;
#pragma unroll
    for (int i = 0; i < 12; ++i) Fu[i] = 0.f;
    Fu[3 * 2 + 0] = -sn * im * dt; Fu[3 * 2 + 1] = -sn * im * dt;
    Fu[4 * 2 + 0] = cs * im * dt;  Fu[4 * 2 + 1] = cs * im * dt;
    Fu[5 * 2 + 0] = -kth * dt;     Fu[5 * 2 + 1] = kth * dt;
  }
};


// Forward-mode scalar (value + NV partials): what jax.jacfwd does to Reacher's step (rollout.py:79-80)
template <int NV>
struct Dual {
  float v, d[NV];
};
template <int NV>
__device__ __forceinline__ Dual<NV> dconst(float c) {
  Dual<NV> r;
  r.v = c;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = 0.f;
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> dvar(float c, int idx) {
  Dual<NV> r = dconst<NV>(c);
  r.d[idx] = 1.f;
  return r;
}
__device__ __forceinline__ float dconst_like(float, float c) { return c; }
template <int NV>
__device__ __forceinline__ Dual<NV> dconst_like(const Dual<NV>&, float c) { return dconst<NV>(c); }
template <int NV>
__device__ __forceinline__ Dual<NV> operator+(const Dual<NV>& a, const Dual<NV>& b) {
  Dual<NV> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> operator-(const Dual<NV>& a, const Dual<NV>& b) {
  Dual<NV> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> operator*(const Dual<NV>& a, const Dual<NV>& b) {
  Dual<NV> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = fmaf(a.v, b.d[i], a.d[i] * b.v);
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> operator*(float a, const Dual<NV>& b) {
  Dual<NV> r;
  r.v = a * b.v;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = a * b.d[i];
  return r;
}
template <int NV>
__device__ __forceinline__ Dual<NV> operator/(const Dual<NV>& a, const Dual<NV>& b) {
  Dual<NV> r;
  const float ib = 1.0f / b.v;
  r.v = a.v * ib;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
__device__ __forceinline__ void dsincos(float a, float* s, float* c) { sincosf(a, s, c); }
template <int NV>
__device__ __forceinline__ void dsincos(const Dual<NV>& a, Dual<NV>* s, Dual<NV>* c) {
  float sv, cv;
  sincosf(a.v, &sv, &cv);
  s->v = sv;
  c->v = cv;
#pragma unroll
  for (int i = 0; i < NV; ++i) { s->d[i] = cv * a.d[i]; c->d[i] = -sv * a.d[i]; }
}

template <>
struct Model<DLC_ENV_REACHER> {  // deluca/envs/classic/_reacher.py:98-141 (two-link arm, state carries the tip offset)
  static constexpr int D = 6, M = 2;
  float m1, m2, l1, l2, g, dt, gx, gy;
  __device__ explicit Model(const DlcEnvSpec& s) {
    m1 = s.p[0]; m2 = s.p[1]; l1 = s.p[2]; l2 = s.p[3]; g = s.p[4]; dt = s.p[5]; gx = s.p[6]; gy = s.p[7];
  }
  template <class T>
  __device__ void step_t(const T* x, const T* u, T* xn) const {
    const T th1 = x[0], th2 = x[1], dth1 = x[2], dth2 = x[3];
    T s2, c2, s1, c1, s12, c12;
    dsincos(th2, &s2, &c2);
    dsincos(th1, &s1, &c1);
    dsincos(th1 + th2, &s12, &c12);
    const float mll = m2 * l1 * l2;
    const T a11 = dconst_like(th1, (m1 + m2) * l1 * l1 + m2 * l2 * l2) + (2.0f * mll) * c2;
    const T a12 = dconst_like(th1, m2 * l2 * l2) + mll * c2;
    const float a22 = m2 * l2 * l2;
    const T b1 = ((u[0] + mll * ((2.0f * dth1 + dth2) * dth2 * s2)) - (m2 * l2 * g) * s12) - ((m1 + m2) * l1 * g) * s1;
    const T b2 = (u[1] - mll * (dth1 * dth1 * s2)) - (m2 * l2 * g) * s12;
    const T det = a22 * a11 - a12 * a12;                          // inv([[a11,a12],[a12,a22]]) @ b   :128-129
    const T dd1 = (a22 * b1 - a12 * b2) / det;
    const T dd2 = (a11 * b2 - a12 * b1) / det;
    const T nth1 = th1 + dt * dth1, nth2 = th2 + dt * dth2;
    T sn1, cn1, sn12, cn12;
    dsincos(nth1, &sn1, &cn1);
    dsincos(nth1 + nth2, &sn12, &cn12);
    xn[0] = nth1;
    xn[1] = nth2;
    xn[2] = dth1 + dt * dd1;
    xn[3] = dth2 + dt * dd2;
    xn[4] = (l1 * cn1 + l2 * cn12) - dconst_like(th1, gx);
    xn[5] = (l1 * sn1 + l2 * sn12) - dconst_like(th1, gy);
  }
  __device__ void step(const float* x, const float* u, float* xn) const { step_t<float>(x, u, xn); }
  __device__ void jac(const float* x, const float* u, float* Fx, float* Fu) const {
    typedef Dual<8> T;
    T xd[D], ud[M], xo[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xd[i] = dvar<8>(x[i], i);
#pragma unroll
    for (int c = 0; c < M; ++c) ud[c] = dvar<8>(u[c], D + c);
    step_t<T>(xd, ud, xo);
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int j = 0; j < D; ++j) Fx[i * D + j] = xo[i].d[j];
#pragma unroll
      for (int c = 0; c < M; ++c) Fu[i * M + c] = xo[i].d[D + c];
    }
  }
};


}  // namespace dlc
