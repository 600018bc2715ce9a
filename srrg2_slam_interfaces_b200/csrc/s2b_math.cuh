// s2b_math.cuh -- small deterministic SE(d) / elementary math shared by the device solve step and
// the host glue.  fp64 with plain (unfused) operations except where fma() is spelled out: the TU is
// compiled with -fmad=false so nvcc fuses nothing on its own, and results are bitwise reproducible
// between the GPU, the host, and the CPU oracle used by the tests.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define S2B_HD __host__ __device__ __forceinline__

namespace s2b {

struct Mat4f {  // isometry embedded in 4x4 row-major; SE(2) lives in rows/cols {0,1,3}
  float m[16];
};

S2B_HD void sincos_det(double x, double& s, double& c) {
  const double k = rint(x * 6.36619772367581382433e-01);
  double r = fma(-k, 1.57079632673412561417e+00, x);
  r = fma(-k, 6.07710050650619224932e-11, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double sr = fma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
  const int quad = (int) (((long long) k) & 3);
  if (quad == 0) { s = sr; c = cr; }
  else if (quad == 1) { s = cr; c = -sr; }
  else if (quad == 2) { s = -sr; c = -cr; }
  else { s = -cr; c = sr; }
}

S2B_HD double atan_unit_det(double z) {
  z = z / (1.0 + sqrt(1.0 + z * z));
  z = z / (1.0 + sqrt(1.0 + z * z));
  z = z / (1.0 + sqrt(1.0 + z * z));
  const double z2 = z * z;
  double p = 1.0 / 19.0;
  p = fma(-p, z2, 1.0 / 17.0);
  p = fma(-p, z2, 1.0 / 15.0);
  p = fma(-p, z2, 1.0 / 13.0);
  p = fma(-p, z2, 1.0 / 11.0);
  p = fma(-p, z2, 1.0 / 9.0);
  p = fma(-p, z2, 1.0 / 7.0);
  p = fma(-p, z2, 1.0 / 5.0);
  p = fma(-p, z2, 1.0 / 3.0);
  p = fma(-p, z2, 1.0);
  return 8.0 * (z * p);
}

S2B_HD double atan2_det(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  if (ax == 0.0 && ay == 0.0) return 0.0;
  double a = (ay <= ax) ? atan_unit_det(ay / ax) : 1.57079632679489655800e+00 - atan_unit_det(ax / ay);
  if (x < 0.0) a = 3.14159265358979311600e+00 - a;
  return (y < 0.0) ? -a : a;
}

S2B_HD double log_det(double x) {
  // x = m * 2^e, m in [sqrt(1/2), sqrt(2)); ln x = e ln2 + 2 atanh((m-1)/(m+1))
  long long bits;
#ifdef __CUDA_ARCH__
  bits = __double_as_longlong(x);
#else
  memcpy(&bits, &x, 8);
#endif
  int e = (int) ((bits >> 52) & 0x7ff) - 1022;
  bits = (bits & 0x800fffffffffffffLL) | 0x3fe0000000000000LL;
  double m;
#ifdef __CUDA_ARCH__
  m = __longlong_as_double(bits);
#else
  memcpy(&m, &bits, 8);
#endif
  if (m < 7.07106781186547572737e-01) { m = m * 2.0; e -= 1; }
  const double z = (m - 1.0) / (m + 1.0);
  const double z2 = z * z;
  double p = 1.0 / 25.0;
  for (int d = 23; d >= 1; d -= 2) p = fma(p, z2, 1.0 / (double) d);
  return fma((double) e, 6.93147180559945286227e-01, 2.0 * (z * p));
}

// ---- quaternion <-> rotation (fp64) ------------------------------------------------------------
S2B_HD void quat_of(const double* R /*9*/, double* q /*x y z w*/) {
  const double tr = R[0] + R[4] + R[8];
  double x, y, z, w, s;
  if (tr > 0.0) {
    s = sqrt(tr + 1.0) * 2.0;
    w = 0.25 * s; x = (R[7] - R[5]) / s; y = (R[2] - R[6]) / s; z = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0;
    w = (R[7] - R[5]) / s; x = 0.25 * s; y = (R[1] + R[3]) / s; z = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0;
    w = (R[2] - R[6]) / s; x = (R[1] + R[3]) / s; y = 0.25 * s; z = (R[5] + R[7]) / s;
  } else {
    s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0;
    w = (R[3] - R[1]) / s; x = (R[2] + R[6]) / s; y = (R[5] + R[7]) / s; z = 0.25 * s;
  }
  const double n = sqrt(x * x + y * y + z * z + w * w);
  x = x / n; y = y / n; z = z / n; w = w / n;
  if (w < 0.0) { x = -x; y = -y; z = -z; w = -w; }
  q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

S2B_HD void rot_of(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double xx = x * x, yy = y * y, zz = z * z;
  const double xy = x * y, xz = x * z, yz = y * z;
  const double wx = w * x, wy = w * y, wz = w * z;
  R[0] = 1.0 - 2.0 * (yy + zz); R[1] = 2.0 * (xy - wz); R[2] = 2.0 * (xz + wy);
  R[3] = 2.0 * (xy + wz); R[4] = 1.0 - 2.0 * (xx + zz); R[5] = 2.0 * (yz - wx);
  R[6] = 2.0 * (xz - wy); R[7] = 2.0 * (yz + wx); R[8] = 1.0 - 2.0 * (xx + yy);
}

// ---- isometry algebra on the 4x4 embedding -----------------------------------------------------
S2B_HD void set_identity(Mat4f& A) {
  for (int i = 0; i < 16; ++i) A.m[i] = (i % 5 == 0) ? 1.f : 0.f;
}

S2B_HD void embed(int dim, const float* M, Mat4f& A) {
  if (dim == 3) {
    for (int i = 0; i < 16; ++i) A.m[i] = M[i];
    return;
  }
  for (int i = 0; i < 16; ++i) A.m[i] = 0.f;
  A.m[0] = M[0]; A.m[1] = M[1]; A.m[3] = M[2];
  A.m[4] = M[3]; A.m[5] = M[4]; A.m[7] = M[5];
  A.m[10] = 1.f; A.m[15] = 1.f;
}

S2B_HD void unembed(int dim, const Mat4f& A, float* M) {
  if (dim == 3) {
    for (int i = 0; i < 16; ++i) M[i] = A.m[i];
    return;
  }
  M[0] = A.m[0]; M[1] = A.m[1]; M[2] = A.m[3];
  M[3] = A.m[4]; M[4] = A.m[5]; M[5] = A.m[7];
  M[6] = 0.f; M[7] = 0.f; M[8] = 1.f;
}

// C = A * B, each entry accumulated in fp64 left to right, rounded once to fp32
S2B_HD void compose(const Mat4f& A, const Mat4f& B, Mat4f& C) {
  Mat4f o;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double c = (double) A.m[i * 4] * (double) B.m[j];
      c = c + (double) A.m[i * 4 + 1] * (double) B.m[4 + j];
      c = c + (double) A.m[i * 4 + 2] * (double) B.m[8 + j];
      if (j == 3) c = c + (double) A.m[i * 4 + 3];
      o.m[i * 4 + j] = (float) c;
    }
  }
  o.m[12] = 0.f; o.m[13] = 0.f; o.m[14] = 0.f; o.m[15] = 1.f;
  C = o;
}

S2B_HD void invert(const Mat4f& A, Mat4f& Ai) {
  Mat4f o;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o.m[i * 4 + j] = A.m[j * 4 + i];
  for (int i = 0; i < 3; ++i) {
    double c = (double) A.m[i] * (double) A.m[3];
    c = c + (double) A.m[4 + i] * (double) A.m[7];
    c = c + (double) A.m[8 + i] * (double) A.m[11];
    o.m[i * 4 + 3] = (float) (-c);
  }
  o.m[12] = 0.f; o.m[13] = 0.f; o.m[14] = 0.f; o.m[15] = 1.f;
  Ai = o;
}

// perturbation vector -> [Rd | td] (3x4, fp64).  VariableSE{2,3}*Right::applyPerturbation uses it
// as X <- X * v2t(dx)  (variable types: R/registration/aligners/multi_aligner.h:152-158)
S2B_HD void exp_right(int dim, int variable, const double* dx, double* D /*12*/) {
  for (int i = 0; i < 12; ++i) D[i] = 0.0;
  if (dim == 2) {
    double s, c;
    sincos_det(dx[2], s, c);
    D[0] = c; D[1] = -s; D[3] = dx[0];
    D[4] = s; D[5] = c; D[7] = dx[1];
    D[10] = 1.0;
    return;
  }
  double R[9];
  if (variable == 0) {  // quaternion vector part
    double q[4];
    const double x = dx[3], y = dx[4], z = dx[5];
    const double n2 = x * x + y * y + z * z;
    if (n2 < 1.0) {
      q[3] = sqrt(1.0 - n2); q[0] = x; q[1] = y; q[2] = z;
    } else {
      const double n = sqrt(n2);
      q[3] = 0.0; q[0] = x / n; q[1] = y / n; q[2] = z / n;
    }
    rot_of(q, R);
  } else {  // Euler, R = Rx Ry Rz
    double sx, cx, sy, cy, sz, cz;
    sincos_det(dx[3], sx, cx);
    sincos_det(dx[4], sy, cy);
    sincos_det(dx[5], sz, cz);
    R[0] = cy * cz;                R[1] = -cy * sz;               R[2] = sy;
    R[3] = cx * sz + sx * sy * cz; R[4] = cx * cz - sx * sy * sz; R[5] = -sx * cy;
    R[6] = sx * sz - cx * sy * cz; R[7] = sx * cz + cx * sy * sz; R[8] = cx * cy;
  }
  for (int i = 0; i < 3; ++i) {
    D[i * 4] = R[i * 3]; D[i * 4 + 1] = R[i * 3 + 1]; D[i * 4 + 2] = R[i * 3 + 2];
    D[i * 4 + 3] = dx[i];
  }
}

S2B_HD void box_plus(int dim, int variable, const double* dx_in, Mat4f& X) {
  const int P = (dim == 3) ? 6 : 3;
  double dx[6] = {0, 0, 0, 0, 0, 0}, D[12];
  for (int i = 0; i < P; ++i) dx[i] = (double) ((float) dx_in[i]);  // upstream perturbations are float
  exp_right(dim, variable, dx, D);
  Mat4f o;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double c = (double) X.m[i * 4] * D[j];
      c = c + (double) X.m[i * 4 + 1] * D[4 + j];
      c = c + (double) X.m[i * 4 + 2] * D[8 + j];
      if (j == 3) c = c + (double) X.m[i * 4 + 3];
      o.m[i * 4 + j] = (float) c;
    }
  }
  o.m[12] = 0.f; o.m[13] = 0.f; o.m[14] = 0.f; o.m[15] = 1.f;
  X = o;
}

// fixTransform (called at R/registration/aligners/multi_aligner_impl.cpp:92)
S2B_HD void fix_transform(int dim, Mat4f& X) {
  if (dim == 3) {
    double R[9], q[4];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i * 3 + j] = (double) X.m[i * 4 + j];
    quat_of(R, q);
    rot_of(q, R);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) X.m[i * 4 + j] = (float) R[i * 3 + j];
  } else {
    double c = 0.5 * ((double) X.m[0] + (double) X.m[5]);
    double s = 0.5 * ((double) X.m[4] - (double) X.m[1]);
    const double n = sqrt(c * c + s * s);
    c = c / n; s = s / n;
    X.m[0] = (float) c; X.m[1] = (float) (-s);
    X.m[4] = (float) s; X.m[5] = (float) c;
  }
}

// dense SPD solve H dx = -b via LL^T; false when H is not positive definite / not finite.
// Compile-time P and full unrolling keep L, y in registers on the device.
template <int P>
S2B_HD bool spd_solve_t(const double* H, const double* b, double* dx) {
  double L[P * P], inv[P];  // inv[j] = 1 / L_jj: one division per column, the rest are products
#pragma unroll
  for (int i = 0; i < P * P; ++i) L[i] = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double d = H[j * P + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = d - L[j * P + k] * L[j * P + k];
    if (!(d > 0.0) || !(d < 1e300)) return false;
    const double ljj = sqrt(d);
    L[j * P + j] = ljj;
    inv[j] = 1.0 / ljj;
#pragma unroll
    for (int i = j + 1; i < P; ++i) {
      double s = H[i * P + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s = s - L[i * P + k] * L[j * P + k];
      L[i * P + j] = s * inv[j];
    }
  }
  double y[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    double s = -b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = s - L[i * P + k] * y[k];
    y[i] = s * inv[i];
  }
#pragma unroll
  for (int i = P - 1; i >= 0; --i) {
    double s = y[i];
#pragma unroll
    for (int k = i + 1; k < P; ++k) s = s - L[k * P + i] * dx[k];
    dx[i] = s * inv[i];
  }
#pragma unroll
  for (int i = 0; i < P; ++i)
    if (!(dx[i] == dx[i]) || fabs(dx[i]) > 1e300) return false;
  return true;
}

S2B_HD bool spd_solve(int P, const double* H, const double* b, double* dx) {
  return P == 6 ? spd_solve_t<6>(H, b, dx) : spd_solve_t<3>(H, b, dx);
}

// SE(d) prior factor e = t2v(Z^-1 X), diagonal information; adds J^T W J / J^T W e, returns chi.
// (factor types SE2PriorErrorFactor / SE3PriorErrorFactorAD:
//  R/registration/aligners/aligner_slice_odometry_prior.h:9,33)
S2B_HD bool prior_accumulate(int dim, int variable, const Mat4f& Z, const Mat4f& X, const float* info,
                             double* H, double* b, double& chi) {
  const int P = (dim == 3) ? 6 : 3;
  Mat4f Zi;
  invert(Z, Zi);
  double E[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double c = (double) Zi.m[i * 4] * (double) X.m[j];
      c = c + (double) Zi.m[i * 4 + 1] * (double) X.m[4 + j];
      c = c + (double) Zi.m[i * 4 + 2] * (double) X.m[8 + j];
      if (j == 3) c = c + (double) Zi.m[i * 4 + 3];
      E[i * 4 + j] = c;
    }
  }
  double e[6], J[36];
  for (int i = 0; i < 36; ++i) J[i] = 0.0;
  if (dim == 3) {
    if (variable != 0) return false;
    double R[9], q[4];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i * 3 + j] = E[i * 4 + j];
    quat_of(R, q);
    e[0] = E[3]; e[1] = E[7]; e[2] = E[11]; e[3] = q[0]; e[4] = q[1]; e[5] = q[2];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) J[i * 6 + j] = R[i * 3 + j];
    J[21] = q[3];  J[22] = -q[2]; J[23] = q[1];
    J[27] = q[2];  J[28] = q[3];  J[29] = -q[0];
    J[33] = -q[1]; J[34] = q[0];  J[35] = q[3];
  } else {
    e[0] = E[3]; e[1] = E[7]; e[2] = atan2_det(E[4], E[0]);
    J[0] = E[0]; J[1] = E[1]; J[3] = E[4]; J[4] = E[5]; J[8] = 1.0;
  }
  double c = 0.0;
  for (int r = 0; r < P; ++r) c = c + ((double) info[r] * e[r]) * e[r];
  chi = c;
  for (int i = 0; i < P; ++i) {
    for (int j = 0; j < P; ++j) {
      double h = 0.0;
      for (int r = 0; r < P; ++r) h = h + (J[r * P + i] * (double) info[r]) * J[r * P + j];
      H[i * P + j] = H[i * P + j] + h;
    }
    double g = 0.0;
    for (int r = 0; r < P; ++r) g = g + (J[r * P + i] * (double) info[r]) * e[r];
    b[i] = b[i] + g;
  }
  return true;
}

// ---- fixed-point scale exponents for the exact integer accumulation ----------------------------
// Per-TERM magnitude bounds of everything that is accumulated, derived from the DATA (global quantities
// only), so that no term can leave its fixed-point range -- there is nothing to clamp:
//   nb = max(1, |n|_max)   normal norms of both clouds      rm = |m|_max   moving point norms
//   E  = max_distance (1 + 2^-10): a correspondence with |S m - f| > E is suppressed and counted as
//        saturated (cannot happen for finder output; guards externally supplied pairs)
//   Gt = nb (1 + 2^-10) >= |R^T n_f|,   Gr = c rm Gt >= |c m x a|   (c = 2 quaternion, 1 otherwise)
//   PLANE: H_tt <= ip Gt^2, H_tr <= ip Gt Gr, H_rr <= ip Gr^2 + in c^2 nb^2, e0 <= nb E,
//          b_t <= ip Gt nb E, b_r <= ip Gr nb E + in c nb Gt, chi <= ip (nb E)^2 + in (2 nb)^2
//   P2P:   H_tt <= ip, H_tr <= ip c rm, H_rr <= ip (c rm)^2, b_t <= ip E, b_r <= ip c rm E, chi <= ip E^2
// (robust weights are <= 1).  A term of class X is stored as rint(term * 2^k_X), k_X = 21 - ceil(log2(1.01 B_X)),
// so |term * 2^k| < 2^21.  chi gets a second, 2^20 times finer, residual word.  All fp32, this order
// (the same sequence of operations as orc_scales in the oracle; host code is compiled without contraction).
enum { kKHtt = 0, kKHtr = 1, kKHrr = 2, kKBt = 3, kKBr = 4, kKChi = 5, kKChiLo = 6, kKCount = 7 };
struct Scales {
  int k[kKCount];
  float err_bound;  // E
};

inline int ceil_log2_float(float x) {
  int e = 0;
  if (!(x > 0.f)) return 0;
  (void) frexpf(x, &e);
  return e;
}

inline Scales choose_scales(int dim, int variable, int factor, float radius_bound2, float normal_bound2,
                            float max_distance, float info_point, float info_normal) {
  const float c = (dim == 3 && variable == 0) ? 2.f : 1.f;
  const float slack = 1.0009765625f;
  float nb = sqrtf(normal_bound2);
  if (!(nb > 1.f)) nb = 1.f;
  const float rm = sqrtf(radius_bound2);
  const float E = max_distance * slack;
  const float ip = info_point, in_ = info_normal;
  float B[kKCount];
  if (factor == 1) {  // PLANE
    const float Gt = nb * slack, Gr = (c * rm) * Gt, e0 = nb * E;
    B[kKHtt] = (ip * Gt) * Gt;
    B[kKHtr] = (ip * Gt) * Gr;
    B[kKHrr] = (ip * Gr) * Gr + ((in_ * c) * c) * (nb * nb);
    B[kKBt] = (ip * Gt) * e0;
    B[kKBr] = (ip * Gr) * e0 + (in_ * c) * (nb * Gt);
    B[kKChi] = (ip * e0) * e0 + (in_ * 4.f) * (nb * nb);
  } else {
    const float Gr = c * rm;
    B[kKHtt] = ip;
    B[kKHtr] = ip * Gr;
    B[kKHrr] = (ip * Gr) * Gr;
    B[kKBt] = ip * E;
    B[kKBr] = (ip * Gr) * E;
    B[kKChi] = (ip * E) * E;
  }
  Scales s;
  for (int k = 0; k < kKChiLo; ++k) {
    float v = B[k] * 1.01f;
    if (!(v > 1e-30f)) v = 1e-30f;  // degenerate inputs (empty cloud, zero information)
    if (!(v < 1e30f)) v = 1e30f;
    s.k[k] = 21 - ceil_log2_float(v);
  }
  s.k[kKChiLo] = s.k[kKChi] + 20;
  s.err_bound = E;
  return s;
}

}  // namespace s2b
