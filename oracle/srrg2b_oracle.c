/*
 * srrg2b_oracle.c -- CPU ORACLE (test infrastructure, NOT product code; see srrg2b_oracle.h).
 *
 * Restates, in plain C:
 *   - the MultiAligner control flow        R/registration/aligners/multi_aligner_impl.cpp:46-263
 *   - the slice glue                       R/registration/aligners/aligner_slice_processor_impl.cpp:19-93
 *   - prior slices                         R/registration/aligners/aligner_slice_odometry_prior.cpp:6-37
 *   - the termination criterion            R/registration/aligners/aligner_termination_criteria_impl.cpp:10-65
 *   - the finder contract                  R/registration/correspondence_finder.h:41-124
 * and the arithmetic those call in srrg2_solver / srrg2_core / srrg2_laser_slam_2d (absent from
 * /root/reference, unpinned HEAD per srrg2_slam_interfaces/README.md:20-23): exact nearest
 * neighbour inside max_distance + normal gate, point-to-point and point+normal ("plane") error
 * factors with right-multiplicative SE(d) perturbation, Saturated/Cauchy/Clamp/Huber robustifiers,
 * one Gauss-Newton step H dx = -b by Cholesky, X <- X * v2t(dx).  PARITY UNPINNED for that part.
 *
 * Numerics contract (shared with the CUDA product so results can be compared bit for bit):
 *   * per-term arithmetic is fp32 with an explicitly written operation order; every fused
 *     multiply-add is spelled fmaf(); compile with -ffp-contract=off so nothing else fuses;
 *   * every per-correspondence contribution to H, b, chi is converted to 64-bit fixed point
 *     (rint(term * 2^k) by one fp32 fma, see to_fix; ranges k from orc_scales, derived from the data so that no
 *     term can leave them) and summed as integers: the sums are exact, hence
 *     independent of summation order, thread count and GPU count;
 *   * the 6x6 / 3x3 solve, pose update and prior factors are fp64 with plain (unfused) operations
 *     and in-house sin/cos/atan2/log so that host libm differences cannot leak in.
 */
#include "srrg2b_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* deterministic elementary functions (fp64, only + - * / sqrt fma)                            */
/* ------------------------------------------------------------------------------------------ */
void orc_sincos(double x, double* s, double* c) {
  const double two_over_pi = 6.36619772367581382433e-01;
  const double pio2_1 = 1.57079632673412561417e+00; /* first 33 bits of pi/2 */
  const double pio2_2 = 6.07710050650619224932e-11; /* pi/2 - pio2_1 */
  double k = nearbyint(x * two_over_pi);
  double r = fma(-k, pio2_1, x);
  r = fma(-k, pio2_2, r);
  double z = r * r;
  /* Taylor/minimax kernels on |r| <= pi/4 (fdlibm __kernel_sin/__kernel_cos coefficients) */
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  double sr = fma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
  long long q = (long long) k;
  switch ((int) (q & 3)) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}

static double orc_atan_unit(double z) { /* 0 <= z <= 1 */
  /* three half-angle reductions: atan z = 2 atan( z / (1 + sqrt(1+z^2)) ) */
  for (int i = 0; i < 3; ++i) {
    z = z / (1.0 + sqrt(1.0 + z * z));
  }
  double z2 = z * z;
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

double orc_atan2(double y, double x) {
  const double pi = 3.14159265358979311600e+00;
  const double pio2 = 1.57079632679489655800e+00;
  double ax = fabs(x), ay = fabs(y);
  double a;
  if (ax == 0.0 && ay == 0.0) {
    return 0.0;
  }
  if (ay <= ax) {
    a = orc_atan_unit(ay / ax);
  } else {
    a = pio2 - orc_atan_unit(ax / ay);
  }
  if (x < 0.0) {
    a = pi - a;
  }
  return (y < 0.0) ? -a : a;
}

double orc_log(double x) { /* x > 0, finite */
  const double ln2 = 6.93147180559945286227e-01;
  int e;
  double m = frexp(x, &e); /* m in [0.5,1) */
  if (m < 7.07106781186547572737e-01) {
    m = m * 2.0;
    e -= 1;
  }
  double z = (m - 1.0) / (m + 1.0);
  double z2 = z * z;
  double p = 1.0 / 25.0;
  for (int d = 23; d >= 1; d -= 2) {
    p = fma(p, z2, 1.0 / (double) d);
  }
  return fma((double) e, ln2, 2.0 * (z * p));
}

/* ------------------------------------------------------------------------------------------ */
/* small SE(d) helpers; matrices are row-major (dim+1)x(dim+1) floats                          */
/* ------------------------------------------------------------------------------------------ */
static void embed4(int dim, const float* M, float* M4) {
  if (dim == 3) {
    memcpy(M4, M, 16 * sizeof(float));
    return;
  }
  memset(M4, 0, 16 * sizeof(float));
  M4[0] = M[0]; M4[1] = M[1]; M4[3] = M[2];
  M4[4] = M[3]; M4[5] = M[4]; M4[7] = M[5];
  M4[10] = 1.f; M4[15] = 1.f;
}

static void unembed4(int dim, const float* M4, float* M) {
  if (dim == 3) {
    memcpy(M, M4, 16 * sizeof(float));
    return;
  }
  M[0] = M4[0]; M[1] = M4[1]; M[2] = M4[3];
  M[3] = M4[4]; M[4] = M4[5]; M[5] = M4[7];
  M[6] = 0.f; M[7] = 0.f; M[8] = 1.f;
}

/* C = A*B for isometries embedded in 4x4; fp64 plain ops, rounded to float */
static void mul4(const float* A, const float* B, float* C) {
  float out[16];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double c = (double) A[i * 4 + 0] * (double) B[0 * 4 + j];
      c = c + (double) A[i * 4 + 1] * (double) B[1 * 4 + j];
      c = c + (double) A[i * 4 + 2] * (double) B[2 * 4 + j];
      if (j == 3) {
        c = c + (double) A[i * 4 + 3];
      }
      out[i * 4 + j] = (float) c;
    }
  }
  out[12] = 0.f; out[13] = 0.f; out[14] = 0.f; out[15] = 1.f;
  memcpy(C, out, sizeof(out));
}

/* inverse of an isometry: [R t]^-1 = [R^T, -R^T t] */
static void inv4(const float* A, float* Ai) {
  float out[16];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      out[i * 4 + j] = A[j * 4 + i];
    }
  }
  for (int i = 0; i < 3; ++i) {
    double c = (double) A[0 * 4 + i] * (double) A[0 * 4 + 3];
    c = c + (double) A[1 * 4 + i] * (double) A[1 * 4 + 3];
    c = c + (double) A[2 * 4 + i] * (double) A[2 * 4 + 3];
    out[i * 4 + 3] = (float) (-c);
  }
  out[12] = 0.f; out[13] = 0.f; out[14] = 0.f; out[15] = 1.f;
  memcpy(Ai, out, sizeof(out));
}

void orc_mul(int dim, const float* A, const float* B, float* C) {
  float A4[16], B4[16], C4[16];
  embed4(dim, A, A4); embed4(dim, B, B4);
  mul4(A4, B4, C4);
  unembed4(dim, C4, C);
}

void orc_inverse(int dim, const float* A, float* Ai) {
  float A4[16], I4[16];
  embed4(dim, A, A4);
  inv4(A4, I4);
  unembed4(dim, I4, Ai);
}

static void quat_from_R(const double R[9], double q[4] /* x y z w */) {
  double tr = R[0] + R[4] + R[8];
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
  double n = sqrt(x * x + y * y + z * z + w * w);
  x = x / n; y = y / n; z = z / n; w = w / n;
  if (w < 0.0) {
    x = -x; y = -y; z = -z; w = -w;
  }
  q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

static void R_from_quat(const double q[4], double R[9]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double xx = x * x, yy = y * y, zz = z * z;
  double xy = x * y, xz = x * z, yz = y * z;
  double wx = w * x, wy = w * y, wz = w * z;
  R[0] = 1.0 - 2.0 * (yy + zz); R[1] = 2.0 * (xy - wz); R[2] = 2.0 * (xz + wy);
  R[3] = 2.0 * (xy + wz); R[4] = 1.0 - 2.0 * (xx + zz); R[5] = 2.0 * (yz - wx);
  R[6] = 2.0 * (xz - wy); R[7] = 2.0 * (yz + wx); R[8] = 1.0 - 2.0 * (xx + yy);
}

/* srrg2_core fixTransform (called at R/registration/aligners/multi_aligner_impl.cpp:92):
 * re-orthonormalise the rotation. Upstream body is not in /root/reference; restated as
 * R -> unit quaternion -> R (3D) and (c,s)/hypot (2D). */
void orc_fix_transform(int dim, float* T) {
  if (dim == 3) {
    double R[9], q[4];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        R[i * 3 + j] = (double) T[i * 4 + j];
      }
    }
    quat_from_R(R, q);
    R_from_quat(q, R);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        T[i * 4 + j] = (float) R[i * 3 + j];
      }
    }
    T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
  } else {
    double c = 0.5 * ((double) T[0] + (double) T[4]);
    double s = 0.5 * ((double) T[3] - (double) T[1]);
    double n = sqrt(c * c + s * s);
    c = c / n; s = s / n;
    T[0] = (float) c; T[1] = (float) (-s);
    T[3] = (float) s; T[4] = (float) c;
    T[6] = 0.f; T[7] = 0.f; T[8] = 1.f;
  }
}

/* geometry{2,3}d::t2v : 3D -> [t ; unit-quaternion vector part, w >= 0], 2D -> [t ; theta] */
void orc_t2v(int dim, const float* T, float* v) {
  if (dim == 3) {
    double R[9], q[4];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        R[i * 3 + j] = (double) T[i * 4 + j];
      }
    }
    quat_from_R(R, q);
    v[0] = T[3]; v[1] = T[7]; v[2] = T[11];
    v[3] = (float) q[0]; v[4] = (float) q[1]; v[5] = (float) q[2];
  } else {
    v[0] = T[2]; v[1] = T[5];
    v[2] = (float) orc_atan2((double) T[3], (double) T[0]);
  }
}

/* perturbation -> isometry (fp64), D = [Rd td] 3x4 row-major (4x4 embedding rows 0..2) */
static void v2t_d(int dim, int variable, const double* dx, double D[12]) {
  memset(D, 0, 12 * sizeof(double));
  if (dim == 2) {
    double s, c;
    orc_sincos(dx[2], &s, &c);
    D[0] = c; D[1] = -s; D[3] = dx[0];
    D[4] = s; D[5] = c; D[7] = dx[1];
    D[10] = 1.0;
    return;
  }
  double R[9];
  if (variable == ORC_VAR_SE3_QUAT_RIGHT) {
    double q[4];
    double x = dx[3], y = dx[4], z = dx[5];
    double n2 = x * x + y * y + z * z;
    if (n2 < 1.0) {
      q[3] = sqrt(1.0 - n2);
      q[0] = x; q[1] = y; q[2] = z;
    } else {
      double n = sqrt(n2);
      q[3] = 0.0;
      q[0] = x / n; q[1] = y / n; q[2] = z / n;
    }
    R_from_quat(q, R);
  } else { /* Euler: R = Rx * Ry * Rz */
    double sx, cx, sy, cy, sz, cz;
    orc_sincos(dx[3], &sx, &cx);
    orc_sincos(dx[4], &sy, &cy);
    orc_sincos(dx[5], &sz, &cz);
    R[0] = cy * cz;                R[1] = -cy * sz;               R[2] = sy;
    R[3] = cx * sz + sx * sy * cz; R[4] = cx * cz - sx * sy * sz; R[5] = -sx * cy;
    R[6] = sx * sz - cx * sy * cz; R[7] = sx * cz + cx * sy * sz; R[8] = cx * cy;
  }
  for (int i = 0; i < 3; ++i) {
    D[i * 4 + 0] = R[i * 3 + 0]; D[i * 4 + 1] = R[i * 3 + 1]; D[i * 4 + 2] = R[i * 3 + 2];
    D[i * 4 + 3] = dx[i];
  }
}

void orc_v2t(int dim, int variable, const float* v, float* T) {
  double dx[6] = {0, 0, 0, 0, 0, 0}, D[12];
  int P = (dim == 3) ? 6 : 3;
  for (int i = 0; i < P; ++i) {
    dx[i] = (double) v[i];
  }
  v2t_d(dim, variable, dx, D);
  float T4[16];
  for (int i = 0; i < 12; ++i) {
    T4[i] = (float) D[i];
  }
  T4[12] = 0.f; T4[13] = 0.f; T4[14] = 0.f; T4[15] = 1.f;
  unembed4(dim, T4, T);
}

/* X <- X * v2t(dx)  (Variable*Right::applyPerturbation), X embedded 4x4 float */
static void apply_perturbation(int dim, int variable, const double* dx, float* X4) {
  double D[12];
  float dxf[6];
  double dxr[6] = {0, 0, 0, 0, 0, 0};
  int P = (dim == 3) ? 6 : 3;
  for (int i = 0; i < P; ++i) { /* upstream perturbation vectors are float */
    dxf[i] = (float) dx[i];
    dxr[i] = (double) dxf[i];
  }
  v2t_d(dim, variable, dxr, D);
  float out[16];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double c = (double) X4[i * 4 + 0] * D[0 * 4 + j];
      c = c + (double) X4[i * 4 + 1] * D[1 * 4 + j];
      c = c + (double) X4[i * 4 + 2] * D[2 * 4 + j];
      if (j == 3) {
        c = c + (double) X4[i * 4 + 3];
      }
      out[i * 4 + j] = (float) c;
    }
  }
  out[12] = 0.f; out[13] = 0.f; out[14] = 0.f; out[15] = 1.f;
  memcpy(X4, out, sizeof(out));
}

/* Cholesky solve of H dx = -b (H full row-major PxP). returns 1 on success */
static int chol_solve(int P, const double* H, const double* b, double* dx) {
  double L[36], inv[6]; /* inv[j] = 1 / L_jj: one division per column, the rest are products */
  memset(L, 0, sizeof(L));
  for (int j = 0; j < P; ++j) {
    double d = H[j * P + j];
    for (int k = 0; k < j; ++k) {
      d = d - L[j * P + k] * L[j * P + k];
    }
    if (!(d > 0.0) || !(d < 1e300)) {
      return 0;
    }
    double ljj = sqrt(d);
    L[j * P + j] = ljj;
    inv[j] = 1.0 / ljj;
    for (int i = j + 1; i < P; ++i) {
      double s = H[i * P + j];
      for (int k = 0; k < j; ++k) {
        s = s - L[i * P + k] * L[j * P + k];
      }
      L[i * P + j] = s * inv[j];
    }
  }
  double y[6];
  for (int i = 0; i < P; ++i) {
    double s = -b[i];
    for (int k = 0; k < i; ++k) {
      s = s - L[i * P + k] * y[k];
    }
    y[i] = s * inv[i];
  }
  for (int i = P - 1; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < P; ++k) {
      s = s - L[k * P + i] * dx[k];
    }
    dx[i] = s * inv[i];
  }
  for (int i = 0; i < P; ++i) {
    if (!(dx[i] == dx[i]) || fabs(dx[i]) > 1e300) {
      return 0;
    }
  }
  return 1;
}

int orc_solve_update(int dim, int variable, const double* H, const double* b, float* T) {
  int P = (dim == 3) ? 6 : 3;
  double dx[6];
  if (!chol_solve(P, H, b, dx)) {
    return 0;
  }
  float T4[16];
  embed4(dim, T, T4);
  apply_perturbation(dim, variable, dx, T4);
  unembed4(dim, T4, T);
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* point access + fp32 transform with pinned operation order                                   */
/* ------------------------------------------------------------------------------------------ */
static inline void get3(const float* a, int dim, int64_t i, float* o) {
  if (dim == 3) {
    o[0] = a[3 * i]; o[1] = a[3 * i + 1]; o[2] = a[3 * i + 2];
  } else {
    o[0] = a[2 * i]; o[1] = a[2 * i + 1]; o[2] = 0.f;
  }
}

static inline void xf_point(const float* S4, const float* m, float* q) {
  for (int r = 0; r < 3; ++r) {
    float t = S4[r * 4 + 0] * m[0];
    t = fmaf(S4[r * 4 + 1], m[1], t);
    t = fmaf(S4[r * 4 + 2], m[2], t);
    q[r] = t + S4[r * 4 + 3];
  }
}

static inline void xf_dir(const float* S4, const float* n, float* o) {
  for (int r = 0; r < 3; ++r) {
    float t = S4[r * 4 + 0] * n[0];
    t = fmaf(S4[r * 4 + 1], n[1], t);
    t = fmaf(S4[r * 4 + 2], n[2], t);
    o[r] = t;
  }
}

/* N1: the clip of the tracker slice, TrackerSliceProcessor_::clip
 * (R/trackers/tracker_slice_processor_impl.cpp:194-205) -> SceneClipper::compute, whose contract is
 * "clipped scene in the robot frame + the indices of its points in the full scene" (R/mapping/scene_clipper.h:104-107;
 * the concrete clippers live in srrg2_laser_slam_2d / srrg2_proslam, not in the tree).  Range clipper: a scene point
 * is kept iff it is valid and |T p| <= max_range (T = scene in robot); the kept point is T p, its normal R n; order =
 * ascending scene index.  fp32, operation order as in the finder's query transform.  Returns the number kept. */
int64_t orc_scene_clip(int dim, const orc_cloud* scene, const float* T, float max_range, float* out_coords, float* out_normals,
                       int32_t* global_indices) {
  float T4[16];
  embed4(dim, T, T4);
  const float r2 = max_range * max_range;
  int64_t k = 0;
  for (int64_t i = 0; i < scene->n; ++i) {
    if (scene->valid && !scene->valid[i]) continue;
    float m[3], q[3];
    get3(scene->coords, dim, i, m);
    xf_point(T4, m, q);
    float d2 = fmaf(q[1], q[1], q[0] * q[0]);
    if (dim == 3) d2 = fmaf(q[2], q[2], d2);
    if (!(d2 <= r2)) continue;
    for (int c = 0; c < dim; ++c) out_coords[k * dim + c] = q[c];
    if (scene->normals && out_normals) {
      float n[3], o[3];
      get3(scene->normals, dim, i, n);
      xf_dir(T4, n, o);
      for (int c = 0; c < dim; ++c) out_normals[k * dim + c] = o[c];
    }
    if (global_indices) global_indices[k] = (int32_t) i;
    ++k;
  }
  return k;
}

/* N2: MergerCorrespondenceHomo_::compute(), R/mapping/merger_correspondence_homo_impl.cpp:11-126, on
 * flat arrays.  corr_scene / corr_meas / corr_resp: the correspondences as the tracker hands them over -- flipped and
 * mapped to the global scene (R/trackers/tracker_slice_processor_impl.cpp:159-191): (scene index, measurement index,
 * response).  n_corr < 0: "no correspondences set" (:31-42, the initial frame): every valid measurement point is
 * appended.  The scene arrays must hold n_scene + n_meas points.  Arithmetic (upstream: Eigen fp32, order unpinned):
 * q = T m in the finder's operation order, d2 = fma chain, merged coordinates (q + s) * 0.5; a merged scene point takes
 * the measurement point's other fields as they are -- its normal is NOT rotated (:71 copies the point, :74 overwrites
 * only the coordinates) --, appended points are transformed in place (coordinates and normal, :112-113).
 * Returns the new scene size. */
int64_t orc_scene_merge(int dim, float* scene_coords, float* scene_normals, uint8_t* scene_valid, int64_t n_scene,
                        const orc_cloud* meas, const float* T, int64_t n_corr, const int32_t* corr_scene, const int32_t* corr_meas,
                        const float* corr_resp, float maximum_response, float maximum_distance_geometry_squared,
                        int64_t target_number_of_merges, int64_t* n_merged_out, int64_t* n_added_out) {
  float T4[16];
  embed4(dim, T, T4);
  uint8_t* merged = (uint8_t*) calloc((size_t) (meas->n > 0 ? meas->n : 1), 1);
  int64_t n_merged = 0, n_added = 0, n = n_scene;
  int append = 1;
  if (n_corr >= 0) {
    for (int64_t c = 0; c < n_corr; ++c) {
      const int64_t g = corr_scene[c], j = corr_meas[c];
      if (!(corr_resp[c] < maximum_response)) continue;
      float m[3], q[3], sp[3];
      get3(meas->coords, dim, j, m);
      xf_point(T4, m, q);
      get3(scene_coords, dim, g, sp);
      const float dx = q[0] - sp[0], dy = q[1] - sp[1], dz = q[2] - sp[2];
      float d2 = fmaf(dy, dy, dx * dx);
      if (dim == 3) d2 = fmaf(dz, dz, d2);
      if (!(d2 < maximum_distance_geometry_squared)) continue;
      for (int k = 0; k < dim; ++k) {
        scene_coords[g * dim + k] = (q[k] + sp[k]) * 0.5f;
        if (scene_normals && meas->normals) scene_normals[g * dim + k] = meas->normals[j * dim + k];
      }
      if (!merged[j]) { merged[j] = 1; ++n_merged; }
    }
    append = n_merged < target_number_of_merges;
  }
  if (append) {
    for (int64_t j = 0; j < meas->n; ++j) {
      if (merged[j] || (meas->valid && !meas->valid[j])) continue;
      float m[3], q[3];
      get3(meas->coords, dim, j, m);
      xf_point(T4, m, q);
      for (int k = 0; k < dim; ++k) scene_coords[n * dim + k] = q[k];
      if (scene_normals && meas->normals) {
        float nn[3], o[3];
        get3(meas->normals, dim, j, nn);
        xf_dir(T4, nn, o);
        for (int k = 0; k < dim; ++k) scene_normals[n * dim + k] = o[k];
      }
      if (scene_valid) scene_valid[n] = 1;
      ++n; ++n_added;
    }
  }
  free(merged);
  if (n_merged_out) *n_merged_out = n_merged;
  if (n_added_out) *n_added_out = n_added;
  return n;
}



static inline float dist2(const float* q, const float* f) {
  float dx = q[0] - f[0], dy = q[1] - f[1], dz = q[2] - f[2];
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

static inline float dot3(const float* a, const float* b) {
  return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]));
}

/* ------------------------------------------------------------------------------------------ */
/* exact NN index: kd-tree with full backtracking, lexicographic (d2, index) minimum            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t lo, hi, left, right, axis;
  float split;
} kdnode;

struct orc_index {
  int dim, method;
  int64_t n;        /* fixed cloud size */
  float* pts;       /* n x 3 padded */
  int32_t* perm;    /* valid point indices, kd order */
  float* lpts;      /* points in perm order (x3) */
  int64_t nvalid;
  kdnode* nodes;
  int32_t nnodes, cap;
};

#define KD_LEAF 8

static void kd_select(const float* pts, int32_t* perm, int64_t lo, int64_t hi, int64_t k, int axis) {
  /* quickselect: afterwards perm[k] holds the k-th smallest along axis in [lo,hi) */
  while (hi - lo > 1) {
    float pv = pts[3 * (int64_t) perm[lo + (hi - lo) / 2] + axis];
    int64_t i = lo, j = hi - 1;
    while (i <= j) {
      while (pts[3 * (int64_t) perm[i] + axis] < pv) ++i;
      while (pts[3 * (int64_t) perm[j] + axis] > pv) --j;
      if (i <= j) {
        int32_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
        ++i; --j;
      }
    }
    if (k <= j) {
      hi = j + 1;
    } else if (k >= i) {
      lo = i;
    } else {
      return;
    }
  }
}

static int32_t kd_build(orc_index* ix, int64_t lo, int64_t hi) {
  if (ix->nnodes == ix->cap) {
    ix->cap = ix->cap ? ix->cap * 2 : 1024;
    ix->nodes = (kdnode*) realloc(ix->nodes, sizeof(kdnode) * (size_t) ix->cap);
  }
  int32_t id = ix->nnodes++;
  kdnode nd;
  nd.lo = (int32_t) lo; nd.hi = (int32_t) hi; nd.left = -1; nd.right = -1; nd.axis = 0; nd.split = 0.f;
  if (hi - lo > KD_LEAF) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = lo; i < hi; ++i) {
      const float* p = ix->pts + 3 * (int64_t) ix->perm[i];
      for (int a = 0; a < 3; ++a) {
        if (p[a] < mn[a]) mn[a] = p[a];
        if (p[a] > mx[a]) mx[a] = p[a];
      }
    }
    int axis = 0;
    float ext = mx[0] - mn[0];
    for (int a = 1; a < 3; ++a) {
      if (mx[a] - mn[a] > ext) {
        ext = mx[a] - mn[a];
        axis = a;
      }
    }
    if (ext > 0.f) {
      int64_t mid = lo + (hi - lo) / 2;
      kd_select(ix->pts, ix->perm, lo, hi, mid, axis);
      nd.axis = axis;
      nd.split = ix->pts[3 * (int64_t) ix->perm[mid] + axis];
      ix->nodes[id] = nd;
      int32_t l = kd_build(ix, lo, mid);
      int32_t r = kd_build(ix, mid, hi);
      nd.left = l; nd.right = r;
    }
  }
  ix->nodes[id] = nd;
  return id;
}

orc_index* orc_index_create(int dim, const orc_cloud* fixed, int method) {
  orc_index* ix = (orc_index*) calloc(1, sizeof(orc_index));
  ix->dim = dim; ix->method = method; ix->n = fixed->n;
  ix->pts = (float*) malloc(sizeof(float) * 3 * (size_t) (fixed->n + 1));
  ix->perm = (int32_t*) malloc(sizeof(int32_t) * (size_t) (fixed->n + 1));
  int64_t nv = 0;
  for (int64_t i = 0; i < fixed->n; ++i) {
    get3(fixed->coords, dim, i, ix->pts + 3 * i);
    if (!fixed->valid || fixed->valid[i]) {
      ix->perm[nv++] = (int32_t) i;
    }
  }
  ix->nvalid = nv;
  if (method == ORC_NN_KDTREE && nv > 0) {
    kd_build(ix, 0, nv);
  }
  ix->lpts = (float*) malloc(sizeof(float) * 3 * (size_t) (nv + 1));
  for (int64_t i = 0; i < nv; ++i) {
    memcpy(ix->lpts + 3 * i, ix->pts + 3 * (int64_t) ix->perm[i], 3 * sizeof(float));
  }
  return ix;
}

void orc_index_free(orc_index* ix) {
  if (!ix) return;
  free(ix->pts); free(ix->perm); free(ix->lpts); free(ix->nodes); free(ix);
}

static void kd_search(const orc_index* ix, int32_t node, const float* q, float* best_d2, int32_t* best_i) {
  const kdnode* nd = &ix->nodes[node];
  if (nd->left < 0) {
    for (int32_t i = nd->lo; i < nd->hi; ++i) {
      float d2 = dist2(q, ix->lpts + 3 * (int64_t) i);
      int32_t id = ix->perm[i];
      if (d2 < *best_d2 || (d2 == *best_d2 && id < *best_i)) {
        *best_d2 = d2; *best_i = id;
      }
    }
    return;
  }
  float diff = q[nd->axis] - nd->split;
  int32_t near = diff < 0.f ? nd->left : nd->right;
  int32_t far = diff < 0.f ? nd->right : nd->left;
  kd_search(ix, near, q, best_d2, best_i);
  if (diff * diff <= *best_d2) {
    kd_search(ix, far, q, best_d2, best_i);
  }
}

static int32_t nn_query(const orc_index* ix, const float* q, float md2, float* d2_out) {
  float best = md2;
  int32_t bi = INT32_MAX;
  if (ix->method == ORC_NN_KDTREE) {
    if (ix->nvalid > 0) {
      kd_search(ix, 0, q, &best, &bi);
    }
  } else {
    for (int64_t i = 0; i < ix->nvalid; ++i) {
      float d2 = dist2(q, ix->lpts + 3 * i);
      int32_t id = ix->perm[i];
      if (d2 < best || (d2 == best && id < bi)) {
        best = d2; bi = id;
      }
    }
  }
  if (bi == INT32_MAX) {
    return -1;
  }
  *d2_out = best;
  return bi;
}

/* projective association: index image of the fixed cloud (nearest depth wins, then lowest index) */
static int proj_pixel(const orc_finder_params* fp, const float* p, int* col, int* row) {
  float z = p[2];
  if (!(z > fp->min_depth) || !(z < fp->max_depth)) {
    return 0;
  }
  float u = fmaf(fp->fx, p[0] / z, fp->cx);
  float v = fmaf(fp->fy, p[1] / z, fp->cy);
  float uf = floorf(u + 0.5f), vf = floorf(v + 0.5f);
  if (!(uf >= 0.f) || !(vf >= 0.f) || !(uf < (float) fp->width) || !(vf < (float) fp->height)) {
    return 0;
  }
  *col = (int) uf; *row = (int) vf;
  return 1;
}

int orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void) n;
  return 1;
#endif
}

/* a3 -- CorrespondenceFinder_::compute() (interface R/registration/correspondence_finder.h:56,
 * called from R/registration/aligners/aligner_slice_processor_impl.cpp:43).  For every valid moving
 * point j: q = S m_j; nearest valid fixed point (lowest index on ties) with |q-f|^2 <= max_d^2;
 * rejected if n_f . (R_S n_m) < normal_cos; response = |q - f|. */
int orc_find(const orc_index* index, int dim, const orc_cloud* fixed, const orc_cloud* moving,
             const float* S, const orc_finder_params* fp, int32_t* fidx, float* resp) {
  float S4[16];
  embed4(dim, S, S4);
  const float md2 = fp->max_distance * fp->max_distance;
  const int use_normals = (fixed->normals && moving->normals && fp->normal_cos > -1.f);
  int64_t* image = NULL;
  if (fp->kind == ORC_FINDER_PROJECTIVE) {
    if (dim != 3) return 1;
    int64_t npx = (int64_t) fp->width * fp->height;
    image = (int64_t*) malloc(sizeof(int64_t) * (size_t) npx);
    for (int64_t i = 0; i < npx; ++i) image[i] = INT64_MAX;
    for (int64_t i = 0; i < fixed->n; ++i) {
      if (fixed->valid && !fixed->valid[i]) continue;
      float f[3];
      int c, r;
      get3(fixed->coords, dim, i, f);
      if (!proj_pixel(fp, f, &c, &r)) continue;
      uint32_t zb;
      memcpy(&zb, &f[2], 4);
      int64_t key = ((int64_t) zb << 32) | (int64_t) (uint32_t) i;
      int64_t* px = &image[(int64_t) r * fp->width + c];
      if (key < *px) *px = key;
    }
  }
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < moving->n; ++j) {
    fidx[j] = -1;
    resp[j] = 0.f;
    if (moving->valid && !moving->valid[j]) continue;
    float m[3], q[3];
    get3(moving->coords, dim, j, m);
    xf_point(S4, m, q);
    float d2 = 0.f;
    int32_t bi = -1;
    if (fp->kind == ORC_FINDER_PROJECTIVE) {
      int c, r;
      if (!proj_pixel(fp, q, &c, &r)) continue;
      int64_t key = image[(int64_t) r * fp->width + c];
      if (key == INT64_MAX) continue;
      bi = (int32_t) (key & 0xffffffffLL);
      float f[3];
      get3(fixed->coords, dim, bi, f);
      d2 = dist2(q, f);
      if (!(d2 <= md2)) continue;
    } else {
      bi = nn_query(index, q, md2, &d2);
      if (bi < 0) continue;
    }
    if (use_normals) {
      float nm[3], nq[3], nf[3];
      get3(moving->normals, dim, j, nm);
      xf_dir(S4, nm, nq);
      get3(fixed->normals, dim, bi, nf);
      if (dot3(nf, nq) < fp->normal_cos) continue;
    }
    fidx[j] = bi;
    resp[j] = sqrtf(d2);
  }
  free(image);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* fixed-point scales                                                                          */
/* ------------------------------------------------------------------------------------------ */
static int ceil_log2f_(float x) { /* smallest e with x <= 2^e (x > 0) */
  int e;
  if (!(x > 0.f)) return 0;
  (void) frexpf(x, &e);
  return e;
}

float orc_coord_bound(int dim, const orc_cloud* moving) {
  float b = 0.f;
  for (int64_t i = 0; i < moving->n * dim; ++i) {
    float a = fabsf(moving->coords[i]);
    if (a > b) b = a;
  }
  return b;
}

/* max |n|^2 over the valid points of a cloud (0 without normals): the fixed-point ranges below are
 * derived from the data, so that no accumulated term can leave its range */
float orc_normal_bound2(int dim, const orc_cloud* c) {
  float b = 0.f;
  if (!c->normals) return 0.f;
  for (int64_t i = 0; i < c->n; ++i) {
    if (c->valid && !c->valid[i]) continue;
    const float* n = c->normals + i * dim;
    float v = n[0] * n[0];
    v = fmaf(n[1], n[1], v);
    if (dim == 3) v = fmaf(n[2], n[2], v);
    if (v > b) b = v;
  }
  return b;
}

/* max |m|^2 over the valid points of the moving cloud (the rotation columns of J scale with it) */
float orc_radius_bound2(int dim, const orc_cloud* c) {
  float b = 0.f;
  for (int64_t i = 0; i < c->n; ++i) {
    if (c->valid && !c->valid[i]) continue;
    const float* p = c->coords + i * dim;
    float v = p[0] * p[0];
    v = fmaf(p[1], p[1], v);
    if (dim == 3) v = fmaf(p[2], p[2], v);
    if (v > b) b = v;
  }
  return b;
}

/* Per-TERM magnitude bounds of everything that is accumulated, derived from the DATA (global quantities
 * only), so that no term can leave its fixed-point range -- there is nothing to clamp:
 *   nb = max(1, |n|_max)   normal norms of both clouds      rm = |m|_max   moving point norms
 *   E  = max_distance (1 + 2^-10): a correspondence with |S m - f| > E is suppressed and counted as
 *        saturated (cannot happen for finder output; guards externally supplied pairs)
 *   Gt = nb (1 + 2^-10) >= |R^T n_f|,   Gr = c rm Gt >= |c m x a|   (c = 2 quaternion, 1 otherwise)
 *   PLANE: H_tt <= ip Gt^2, H_tr <= ip Gt Gr, H_rr <= ip Gr^2 + in c^2 nb^2, e0 <= nb E,
 *          b_t <= ip Gt nb E, b_r <= ip Gr nb E + in c nb Gt, chi <= ip (nb E)^2 + in (2 nb)^2
 *   P2P:   H_tt <= ip, H_tr <= ip c rm, H_rr <= ip (c rm)^2, b_t <= ip E, b_r <= ip c rm E, chi <= ip E^2
 * (robust weights are <= 1).  A term of class X is stored as rint(term * 2^k_X), k_X = 21 - ceil(log2(1.01 B_X)),
 * so |term * 2^k| < 2^21.  chi gets a second, 2^20 times finer, residual word.  All fp32, this order. */
int orc_scales(int dim, int variable, float radius_bound2, float normal_bound2, const orc_finder_params* fp,
               const orc_factor_params* fa, orc_scales_t* out) {
  const float c = (dim == 3 && variable == ORC_VAR_SE3_QUAT_RIGHT) ? 2.f : 1.f;
  const float slack = 1.0009765625f;
  float nb = sqrtf(normal_bound2);
  if (!(nb > 1.f)) nb = 1.f;
  const float rm = sqrtf(radius_bound2);
  const float E = fp->max_distance * slack;
  const float ip = fa->info_point, in_ = fa->info_normal;
  float B[ORC_K_COUNT];
  if (fa->factor == ORC_FACTOR_PLANE) {
    const float Gt = nb * slack, Gr = (c * rm) * Gt, e0 = nb * E;
    B[ORC_K_HTT] = (ip * Gt) * Gt;
    B[ORC_K_HTR] = (ip * Gt) * Gr;
    B[ORC_K_HRR] = (ip * Gr) * Gr + ((in_ * c) * c) * (nb * nb);
    B[ORC_K_BT] = (ip * Gt) * e0;
    B[ORC_K_BR] = (ip * Gr) * e0 + (in_ * c) * (nb * Gt);
    B[ORC_K_CHI] = (ip * e0) * e0 + (in_ * 4.f) * (nb * nb);
  } else {
    const float Gr = c * rm;
    B[ORC_K_HTT] = ip;
    B[ORC_K_HTR] = ip * Gr;
    B[ORC_K_HRR] = (ip * Gr) * Gr;
    B[ORC_K_BT] = ip * E;
    B[ORC_K_BR] = (ip * Gr) * E;
    B[ORC_K_CHI] = (ip * E) * E;
  }
  for (int k = 0; k < ORC_K_CHI_LO; ++k) {
    float v = B[k] * 1.01f;
    if (!(v > 1e-30f)) v = 1e-30f;   /* degenerate inputs (empty cloud, zero information) */
    if (!(v < 1e30f)) v = 1e30f;
    out->k[k] = 21 - ceil_log2f_(v);
  }
  out->k[ORC_K_CHI_LO] = out->k[ORC_K_CHI] + 20;
  out->err_bound = E;
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* a5 -- per-correspondence linearisation (FactorCorrespondenceDriven_<F,...> inside            */
/* Solver::compute(), reached from R/registration/aligners/multi_aligner_impl.cpp:112)          */
/* ------------------------------------------------------------------------------------------ */
/* Fixed-point value of one term with k fractional bits, |v| <= 2^(21-k): ONE fp32 operation,
 * u = fma(v, 2^(k-22), 3.5) lies in [3, 4] where the fp32 spacing is 2^-22, i.e. one unit of 2^-k of v, so
 * bits(u) - bits(3.5f) = rint(v * 2^k) (ties to even).  The ranges are guaranteed by orc_scales. */
static inline int32_t to_fix(float v, int k) {
  const float s = ldexpf(1.f, k - 22);
  const float u = fmaf(v, s, 3.5f);
  int32_t ui;
  memcpy(&ui, &u, 4);
  return ui - 0x40600000;
}

/* chi as a (coarse, residual) pair: the residual of the coarse rounding is exact in fp32 */
static inline void to_fix2(float v, int k_hi, int k_lo, int64_t* hi, int64_t* lo) {
  const float s = ldexpf(1.f, k_hi - 22), inv_s = ldexpf(1.f, 22 - k_hi);
  const float u = fmaf(v, s, 3.5f);
  int32_t ui;
  memcpy(&ui, &u, 4);
  *hi += ui - 0x40600000;
  const float rem = v - (u - 3.5f) * inv_s;
  *lo += to_fix(rem, k_lo);
}

/* robustifier on chi with threshold tau -> weight w, robustified rho; returns 1 if kernelized.
 * RobustifierBase::param_chi_threshold is the in-tree evidence (multi_aligner_impl.cpp:194-196);
 * bodies restated from srrg2_solver conventions (unpinned). */
static int robustify(int kind, float tau, float chi, float* w, float* rho) {
  *w = 1.f; *rho = chi;
  if (kind == ORC_ROB_NONE) return 0;
  if (!(chi > tau)) return 0;
  switch (kind) {
    case ORC_ROB_SATURATED:
    case ORC_ROB_CLAMP:
      *w = 0.f; *rho = tau;
      return 1;
    case ORC_ROB_CAUCHY: {
      float r = chi / tau;
      *w = 1.f / (1.f + r);
      *rho = (float) ((double) tau * orc_log(1.0 + (double) r));
      return 1;
    }
    case ORC_ROB_HUBER: {
      float delta = sqrtf(tau);
      float sc = sqrtf(chi);
      *w = delta / sc;
      *rho = fmaf(2.f * delta, sc, -tau);
      return 1;
    }
    default: return 0;
  }
}

/* One correspondence: chi (before the robustifier) and the 21 | 6 upper-triangle entries of J^T Om J and
 * the 6 | 3 entries of J^T Om e for unit robust weight scale w, in the REDUCED form that holds for a
 * rotation matrix R (R^T R = I), right perturbation X <- X v2t(dx), c = rs (2: quaternion, 1: Euler / SE2):
 *   P2P    e = q - f,  J = [R | -c R [m]x]
 *          H = s [ I, -c [m]x ; ., c^2 (|m|^2 I - m m^T) ],  b = s [ r ; c m x r ],  r = R^T e,  s = w ip
 *   PLANE  e = [ n_f^T (q - f) ; R n_m - n_f ],  J = [ a^T, c (m x a)^T ; 0, -c R [n_m]x ],  a = R^T n_f
 *          g = [ a ; c m x a ],  H = s0 g g^T + [ 0, 0 ; 0, s1 c^2 (|n_m|^2 I - n_m n_m^T) ],
 *          b = s0 e0 g + [ 0 ; s1 c (a x n_m) ],  s0 = w ip,  s1 = w in
 * (2D: the analogous forms with the single rotation column R (-m_y, m_x)^T.)  Every fused multiply-add is
 * an explicit fmaf; the CUDA kernels perform the same operations in the same order.
 * The weight-independent part is returned: e rows / chi first (terms_chi), then the H/b terms for weight w. */
typedef struct {
  float q[3], nq[3], d[3];
  float a[3], g[6], e0, en[3], r[3];
} lin_geo;

static float lin_chi(int dim, int factor, const float* S4, const float* m, const float* nm, const float* f,
                     const float* nf, float ip, float in_, float rs, lin_geo* G) {
  xf_point(S4, m, G->q);
  G->nq[0] = G->nq[1] = G->nq[2] = 0.f;
  xf_dir(S4, nm, G->nq);
  for (int k = 0; k < 3; ++k) G->d[k] = G->q[k] - f[k];
  float chi;
  if (factor == ORC_FACTOR_P2P) {
    chi = (ip * G->d[0]) * G->d[0];
    chi = fmaf(ip * G->d[1], G->d[1], chi);
    if (dim == 3) chi = fmaf(ip * G->d[2], G->d[2], chi);
    /* r = R^T d */
    for (int c = 0; c < dim; ++c) {
      float t = S4[0 * 4 + c] * G->d[0];
      t = fmaf(S4[1 * 4 + c], G->d[1], t);
      if (dim == 3) t = fmaf(S4[2 * 4 + c], G->d[2], t);
      G->r[c] = t;
    }
    return chi;
  }
  /* a = R^T n_f */
  for (int c = 0; c < dim; ++c) {
    float t = S4[0 * 4 + c] * nf[0];
    t = fmaf(S4[1 * 4 + c], nf[1], t);
    if (dim == 3) t = fmaf(S4[2 * 4 + c], nf[2], t);
    G->a[c] = t;
  }
  float t;
  if (dim == 3) {
    G->g[0] = G->a[0]; G->g[1] = G->a[1]; G->g[2] = G->a[2];
    t = m[2] * G->a[1]; G->g[3] = rs * fmaf(m[1], G->a[2], -t);
    t = m[0] * G->a[2]; G->g[4] = rs * fmaf(m[2], G->a[0], -t);
    t = m[1] * G->a[0]; G->g[5] = rs * fmaf(m[0], G->a[1], -t);
    G->e0 = fmaf(nf[2], G->d[2], fmaf(nf[1], G->d[1], nf[0] * G->d[0]));
  } else {
    G->g[0] = G->a[0]; G->g[1] = G->a[1];
    t = G->a[0] * m[1]; G->g[2] = fmaf(G->a[1], m[0], -t);
    G->e0 = fmaf(nf[1], G->d[1], nf[0] * G->d[0]);
  }
  for (int k = 0; k < dim; ++k) G->en[k] = G->nq[k] - nf[k];
  chi = (ip * G->e0) * G->e0;
  for (int k = 0; k < dim; ++k) chi = fmaf(in_ * G->en[k], G->en[k], chi);
  return chi;
}

/* Ht[21|6] (upper triangle, row-major), bt[6|3] for robust weight w */
static void lin_terms(int dim, int factor, const lin_geo* G, const float* m, const float* nm, float ip, float in_,
                      float rs, float w, float* Ht, float* bt) {
  float t;
  if (dim == 3 && factor == ORC_FACTOR_P2P) {
    const float s = w * ip, sr = s * rs, srr = sr * rs;
    const float* r = G->r;
    /* rows 0..2 (translation): [ s I | sr * (-[m]x) ] */
    Ht[0] = s;   Ht[1] = 0.f; Ht[2] = 0.f; Ht[3] = 0.f;          Ht[4] = sr * m[2];    Ht[5] = -(sr * m[1]);
    Ht[6] = s;   Ht[7] = 0.f; Ht[8] = -(sr * m[2]);              Ht[9] = 0.f;          Ht[10] = sr * m[0];
    Ht[11] = s;  Ht[12] = sr * m[1];   Ht[13] = -(sr * m[0]);    Ht[14] = 0.f;
    /* rotation block: srr (|m|^2 I - m m^T) */
    Ht[15] = srr * fmaf(m[1], m[1], m[2] * m[2]); Ht[16] = -(srr * (m[0] * m[1])); Ht[17] = -(srr * (m[0] * m[2]));
    Ht[18] = srr * fmaf(m[0], m[0], m[2] * m[2]); Ht[19] = -(srr * (m[1] * m[2]));
    Ht[20] = srr * fmaf(m[0], m[0], m[1] * m[1]);
    bt[0] = s * r[0]; bt[1] = s * r[1]; bt[2] = s * r[2];
    t = m[2] * r[1]; bt[3] = sr * fmaf(m[1], r[2], -t);
    t = m[0] * r[2]; bt[4] = sr * fmaf(m[2], r[0], -t);
    t = m[1] * r[0]; bt[5] = sr * fmaf(m[0], r[1], -t);
    return;
  }
  if (dim == 3) { /* PLANE */
    const float s0 = w * ip, s1 = w * in_, s1r = s1 * rs, s1rr = s1r * rs;
    const float* g = G->g;
    const float* a = G->a;
    float u[6];
    for (int i = 0; i < 6; ++i) u[i] = s0 * g[i];
    /* normal rows: s1 c^2 (|n_m|^2 I - n_m n_m^T), s1 c (a x n_m) */
    float N[6];
    N[0] = s1rr * fmaf(nm[1], nm[1], nm[2] * nm[2]); N[1] = -(s1rr * (nm[0] * nm[1])); N[2] = -(s1rr * (nm[0] * nm[2]));
    N[3] = s1rr * fmaf(nm[0], nm[0], nm[2] * nm[2]); N[4] = -(s1rr * (nm[1] * nm[2]));
    N[5] = s1rr * fmaf(nm[0], nm[0], nm[1] * nm[1]);
    float bn[3];
    t = a[2] * nm[1]; bn[0] = s1r * fmaf(a[1], nm[2], -t);
    t = a[0] * nm[2]; bn[1] = s1r * fmaf(a[2], nm[0], -t);
    t = a[1] * nm[0]; bn[2] = s1r * fmaf(a[0], nm[1], -t);
    int slot = 0, nslot = 0;
    for (int i = 0; i < 6; ++i) {
      for (int j = i; j < 6; ++j) {
        if (i >= 3) Ht[slot++] = fmaf(u[i], g[j], N[nslot++]);
        else Ht[slot++] = u[i] * g[j];
      }
    }
    for (int i = 0; i < 3; ++i) bt[i] = u[i] * G->e0;
    for (int i = 0; i < 3; ++i) bt[3 + i] = fmaf(u[3 + i], G->e0, bn[i]);
    return;
  }
  if (factor == ORC_FACTOR_P2P) { /* 2D */
    const float s = w * ip;
    const float* r = G->r;
    Ht[0] = s; Ht[1] = 0.f; Ht[2] = -(s * m[1]);
    Ht[3] = s; Ht[4] = s * m[0];
    Ht[5] = s * fmaf(m[0], m[0], m[1] * m[1]);
    bt[0] = s * r[0]; bt[1] = s * r[1];
    t = m[1] * r[0]; bt[2] = s * fmaf(m[0], r[1], -t);
    return;
  }
  { /* 2D PLANE */
    const float s0 = w * ip, s1 = w * in_;
    const float* g = G->g;
    const float* a = G->a;
    float u[3];
    for (int i = 0; i < 3; ++i) u[i] = s0 * g[i];
    const float N = s1 * fmaf(nm[0], nm[0], nm[1] * nm[1]);
    t = nm[0] * a[1];
    const float bn = s1 * fmaf(nm[1], a[0], -t);
    Ht[0] = u[0] * g[0]; Ht[1] = u[0] * g[1]; Ht[2] = u[0] * g[2];
    Ht[3] = u[1] * g[1]; Ht[4] = u[1] * g[2];
    Ht[5] = fmaf(u[2], g[2], N);
    bt[0] = u[0] * G->e0; bt[1] = u[1] * G->e0;
    bt[2] = fmaf(u[2], G->e0, bn);
  }
}

/* slots of the 40-entry accumulator */
#define ACC_B 21
#define ACC_CHI_IN 27      /* coarse word; +1 = residual word */
#define ACC_CHI_OUT 29     /* coarse word; +1 = residual word */
#define ACC_N_IN 31
#define ACC_N_OUT 32
#define ACC_N_SUP 33
#define ACC_N_SAT 34       /* suppressed because |S m - f| exceeds the fixed-point error range (also in N_SUP) */
#define ACC_SLOTS ORC_ACC_SLOTS

/* plain: when non-NULL, receives the un-quantised fp64 sums of the same fp32 terms (27 | 9 H/b entries,
 * chi_in, chi_out): the independent check of the fixed-point accumulation */
static void lin_one(int dim, int variable, const orc_factor_params* fa, const orc_scales_t* sc,
                    const float* S4, const float* m, const float* nm, const float* f, const float* nf,
                    int64_t* acc, double* plain, uint8_t* status, float* chi_out) {
  const int P = (dim == 3) ? 6 : 3, NH = P * (P + 1) / 2;
  const float rs = (dim == 3 && variable == ORC_VAR_SE3_QUAT_RIGHT) ? 2.f : 1.f;
  lin_geo G;
  memset(&G, 0, sizeof(G));
  const float chi = lin_chi(dim, fa->factor, S4, m, nm, f, nf, fa->info_point, fa->info_normal, rs, &G);
  float d2 = G.d[0] * G.d[0];
  d2 = fmaf(G.d[1], G.d[1], d2);
  if (dim == 3) d2 = fmaf(G.d[2], G.d[2], d2);
  const int sat = !(d2 <= sc->err_bound * sc->err_bound);  /* also NaN */
  if (sat || !(chi == chi) || isinf(chi)) {
    acc[ACC_N_SUP] += 1;
    if (sat) acc[ACC_N_SAT] += 1;
    if (status) *status = ORC_STAT_SUPPRESSED;
    if (chi_out) *chi_out = chi;
    return;
  }
  float w, rho;
  int kern = robustify(fa->robustifier, fa->chi_threshold, chi, &w, &rho);
  if (kern) {
    acc[ACC_N_OUT] += 1;
    to_fix2(rho, sc->k[ORC_K_CHI], sc->k[ORC_K_CHI_LO], &acc[ACC_CHI_OUT], &acc[ACC_CHI_OUT + 1]);
    if (plain) plain[28] += (double) rho;
  } else {
    acc[ACC_N_IN] += 1;
    to_fix2(chi, sc->k[ORC_K_CHI], sc->k[ORC_K_CHI_LO], &acc[ACC_CHI_IN], &acc[ACC_CHI_IN + 1]);
    if (plain) plain[27] += (double) chi;
  }
  if (status) *status = kern ? ORC_STAT_KERNELIZED : ORC_STAT_INLIER;
  if (chi_out) *chi_out = chi;
  float Ht[21], bt[6];
  lin_terms(dim, fa->factor, &G, m, nm, fa->info_point, fa->info_normal, rs, w, Ht, bt);
  const int T = dim; /* columns < T are the translation part of the perturbation */
  int slot = 0;
  for (int i = 0; i < P; ++i) {
    for (int j = i; j < P; ++j) {
      acc[slot] += to_fix(Ht[slot], sc->k[(j < T) ? ORC_K_HTT : ((i < T) ? ORC_K_HTR : ORC_K_HRR)]);
      if (plain) plain[slot] += (double) Ht[slot];
      ++slot;
    }
  }
  for (int i = 0; i < P; ++i) {
    acc[ACC_B + i] += to_fix(bt[i], sc->k[(i < T) ? ORC_K_BT : ORC_K_BR]);
    if (plain) plain[NH + i] += (double) bt[i];
  }
}

static void acc_to_Hb(int dim, const int64_t* acc, const orc_scales_t* sc, double* H, double* b) {
  const int P = (dim == 3) ? 6 : 3;
  int slot = 0;
  for (int i = 0; i < P; ++i) {
    for (int j = i; j < P; ++j) {
      double v = ldexp((double) acc[slot++], -sc->k[(j < dim) ? ORC_K_HTT : ((i < dim) ? ORC_K_HTR : ORC_K_HRR)]);
      H[i * P + j] = v;
      H[j * P + i] = v;
    }
  }
  for (int i = 0; i < P; ++i) {
    b[i] = ldexp((double) acc[ACC_B + i], -sc->k[(i < dim) ? ORC_K_BT : ORC_K_BR]);
  }
}

int orc_linearize(int dim, int variable, const orc_cloud* fixed, const orc_cloud* moving,
                  const int32_t* fidx, const float* S, const orc_finder_params* fp,
                  const orc_factor_params* fa, float radius_bound2, float normal_bound2,
                  int64_t* acc_out, double* H, double* b, orc_iter_stats* stats, uint8_t* status_dense,
                  float* chi_dense, double* plain_out) {
  float S4[16];
  embed4(dim, S, S4);
  orc_scales_t sc;
  /* radius_bound2 / normal_bound2 describe the WHOLE moving cloud when `moving` is one shard of it */
  if (!(normal_bound2 > 0.f)) {
    const float nf2 = orc_normal_bound2(dim, fixed), nm2 = orc_normal_bound2(dim, moving);
    normal_bound2 = nf2 > nm2 ? nf2 : nm2;
  }
  if (!(radius_bound2 > 0.f)) radius_bound2 = orc_radius_bound2(dim, moving);
  orc_scales(dim, variable, radius_bound2, normal_bound2, fp, fa, &sc);
  int64_t acc[ACC_SLOTS];
  memset(acc, 0, sizeof(acc));
  double plain[32];
  memset(plain, 0, sizeof(plain));
  const int have_n = (fixed->normals && moving->normals);
  if (fa->factor == ORC_FACTOR_PLANE && !have_n) return 1;
#pragma omp parallel
  {
    int64_t loc[ACC_SLOTS];
    double ploc[32];
    memset(loc, 0, sizeof(loc));
    memset(ploc, 0, sizeof(ploc));
#pragma omp for schedule(static)
    for (int64_t j = 0; j < moving->n; ++j) {
      if (status_dense) status_dense[j] = ORC_STAT_NONE;
      if (chi_dense) chi_dense[j] = 0.f;
      int32_t i = fidx[j];
      if (i < 0) continue;
      if (i >= fixed->n || (fixed->valid && !fixed->valid[i]) || (moving->valid && !moving->valid[j])) {
        loc[ACC_N_SUP] += 1;
        if (status_dense) status_dense[j] = ORC_STAT_SUPPRESSED;
        continue;
      }
      float m[3], nm[3] = {0, 0, 0}, f[3], nf[3] = {0, 0, 0};
      get3(moving->coords, dim, j, m);
      get3(fixed->coords, dim, i, f);
      if (have_n) {
        get3(moving->normals, dim, j, nm);
        get3(fixed->normals, dim, i, nf);
      }
      lin_one(dim, variable, fa, &sc, S4, m, nm, f, nf, loc, plain_out ? ploc : NULL,
              status_dense ? status_dense + j : NULL, chi_dense ? chi_dense + j : NULL);
    }
#pragma omp critical
    {
      for (int k = 0; k < ACC_SLOTS; ++k) acc[k] += loc[k];
      for (int k = 0; k < 32; ++k) plain[k] += ploc[k];
    }
  }
  if (acc_out) memcpy(acc_out, acc, sizeof(acc));
  if (plain_out) memcpy(plain_out, plain, sizeof(plain));
  if (H && b) acc_to_Hb(dim, acc, &sc, H, b);
  if (stats) {
    stats->num_inliers = acc[ACC_N_IN];
    stats->num_outliers = acc[ACC_N_OUT];
    stats->num_suppressed = acc[ACC_N_SUP];
    stats->num_saturated = acc[ACC_N_SAT];
    stats->num_correspondences = acc[ACC_N_IN] + acc[ACC_N_OUT] + acc[ACC_N_SUP];
    stats->chi_inliers = ldexp((double) acc[ACC_CHI_IN], -sc.k[ORC_K_CHI]) + ldexp((double) acc[ACC_CHI_IN + 1], -sc.k[ORC_K_CHI_LO]);
    stats->chi_outliers = ldexp((double) acc[ACC_CHI_OUT], -sc.k[ORC_K_CHI]) + ldexp((double) acc[ACC_CHI_OUT + 1], -sc.k[ORC_K_CHI_LO]);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* a9 -- SE(d) prior factor e = t2v(Z^-1 X) with diagonal information (fp64)                    */
/* (factor types: R/registration/aligners/aligner_slice_odometry_prior.h:9,33)                  */
/* ------------------------------------------------------------------------------------------ */
static int prior_terms(int dim, int variable, const float* Z4, const float* X4, const float* info,
                       double* H, double* b, double* chi) {
  const int P = (dim == 3) ? 6 : 3;
  float Zi[16];
  inv4(Z4, Zi);
  double E[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double c = (double) Zi[i * 4 + 0] * (double) X4[0 * 4 + j];
      c = c + (double) Zi[i * 4 + 1] * (double) X4[1 * 4 + j];
      c = c + (double) Zi[i * 4 + 2] * (double) X4[2 * 4 + j];
      if (j == 3) c = c + (double) Zi[i * 4 + 3];
      E[i * 4 + j] = c;
    }
  }
  double e[6], J[36];
  memset(J, 0, sizeof(J));
  if (dim == 3) {
    if (variable != ORC_VAR_SE3_QUAT_RIGHT) return 0;
    double R[9], q[4];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) R[i * 3 + j] = E[i * 4 + j];
    }
    quat_from_R(R, q);
    e[0] = E[3]; e[1] = E[7]; e[2] = E[11];
    e[3] = q[0]; e[4] = q[1]; e[5] = q[2];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) J[i * 6 + j] = R[i * 3 + j];
    }
    /* d vec(q_e * (1,dq)) / d dq = w I + [v]x */
    J[3 * 6 + 3] = q[3];  J[3 * 6 + 4] = -q[2]; J[3 * 6 + 5] = q[1];
    J[4 * 6 + 3] = q[2];  J[4 * 6 + 4] = q[3];  J[4 * 6 + 5] = -q[0];
    J[5 * 6 + 3] = -q[1]; J[5 * 6 + 4] = q[0];  J[5 * 6 + 5] = q[3];
  } else {
    e[0] = E[3]; e[1] = E[7];
    e[2] = orc_atan2(E[4], E[0]);
    J[0 * 3 + 0] = E[0]; J[0 * 3 + 1] = E[1];
    J[1 * 3 + 0] = E[4]; J[1 * 3 + 1] = E[5];
    J[2 * 3 + 2] = 1.0;
  }
  double c = 0.0;
  for (int r = 0; r < P; ++r) {
    c = c + ((double) info[r] * e[r]) * e[r];
  }
  *chi = c;
  for (int i = 0; i < P; ++i) {
    for (int j = 0; j < P; ++j) {
      double h = 0.0;
      for (int r = 0; r < P; ++r) {
        h = h + (J[r * P + i] * (double) info[r]) * J[r * P + j];
      }
      H[i * P + j] = H[i * P + j] + h;
    }
    double g = 0.0;
    for (int r = 0; r < P; ++r) {
      g = g + (J[r * P + i] * (double) info[r]) * e[r];
    }
    b[i] = b[i] + g;
  }
  return 1;
}


/* ------------------------------------------------------------------------------------------ */
/* a6 -- AlignerTerminationCriteriaStandard_ (aligner_termination_criteria_impl.cpp:10-65)      */
/* ------------------------------------------------------------------------------------------ */
#define TC_MAXW 64
typedef struct {
  int window, count, head;
  double v[TC_MAXW];
} tc_ring;

static void ring_reset(tc_ring* r, int w) {
  r->window = w > TC_MAXW ? TC_MAXW : (w < 1 ? 1 : w);
  r->count = 0; r->head = 0;
}
static void ring_add(tc_ring* r, double x) {
  r->v[r->head] = x;
  r->head = (r->head + 1) % r->window;
  if (r->count < r->window) r->count++;
}
static double ring_max(const tc_ring* r) {
  double m = r->v[0];
  for (int i = 1; i < r->count; ++i) if (r->v[i] > m) m = r->v[i];
  return m;
}
static double ring_min(const tc_ring* r) {
  double m = r->v[0];
  for (int i = 1; i < r->count; ++i) if (r->v[i] < m) m = r->v[i];
  return m;
}

typedef struct {
  tc_ring ncorr, ninl, nout, chi;
  int num_iteration;
} term_crit;

static void tc_init(term_crit* tc, const orc_aligner_params* ap) { /* :10-21 */
  tc->num_iteration = 0;
  ring_reset(&tc->ncorr, ap->window_size);
  ring_reset(&tc->ninl, ap->window_size);
  ring_reset(&tc->nout, ap->window_size);
  ring_reset(&tc->chi, ap->window_size);
}

static int tc_has_to_stop(term_crit* tc, const orc_aligner_params* ap, const orc_iter_stats* st,
                          int64_t ncorr) { /* :24-65, quirks kept */
  ++tc->num_iteration;
  int ninl = (int) st->num_inliers;
  int nout = (int) st->num_outliers;
  float chi = (float) st->chi_inliers / (float) ninl;
  if (!ninl) return 0;
  ring_add(&tc->ncorr, (double) ncorr);
  ring_add(&tc->ninl, (double) ninl);
  ring_add(&tc->nout, (double) nout);
  ring_add(&tc->chi, (double) chi);
  if (tc->ncorr.count < ap->window_size) return 0;
  if (ring_max(&tc->nout) - ring_min(&tc->nout) > (double) ap->num_correspondences_range) return 0; /* :46 */
  if (ring_max(&tc->ninl) - ring_min(&tc->ninl) > (double) ap->num_inliers_range) return 0;
  float crange = (float) ring_max(&tc->chi) - (float) ring_min(&tc->chi);
  float cmax = (float) ring_max(&tc->chi);
  if (crange > (float) ap->num_outliers_range) return 0; /* :53 */
  if (crange / cmax > ap->chi_epsilon) return 0;
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* a1/a2/a7/a8 -- MultiAlignerBase_::compute / _runSolver / _postCompute / _pruneCorrespondences */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int dim, n_slices, nn_method;
  const orc_slice* slices;
  const orc_aligner_params* ap;
  orc_index** index;
  int32_t** fidx;   /* dense per slice */
  float** resp;
  uint8_t** fstat;
  int64_t* ncorr;
  float X4[16];     /* variable 0 estimate (moving in fixed) */
  term_crit tc;
  orc_iter_stats* stats;
  int n_stats, cap_stats;
  int status;
} run_state;

/* _runSolver, multi_aligner_impl.cpp:97-128 */
static void run_solver(run_state* rs, int iterations, int use_tc, int clamp) {
  const int dim = rs->dim, P = (dim == 3) ? 6 : 3;
  float backup[16];
  memcpy(backup, rs->X4, sizeof(backup)); /* :102 */
  for (int it = 0; it < iterations; ++it) {
    int good = 0;
    /* _computeCorrespondencesPerSlices, multi_aligner.h:126-138 */
    for (int s = 0; s < rs->n_slices; ++s) {
      const orc_slice* sl = &rs->slices[s];
      if (sl->kind == ORC_SLICE_PRIOR) {
        rs->ncorr[s] = 1;  /* aligner_slice_processor_prior.h:75-77 */
        good |= 1;         /* :65-67 */
        continue;
      }
      float ris4[16], S4[16], S[16];
      embed4(dim, sl->robot_in_sensor, ris4);
      mul4(ris4, backup, S4); /* aligner_slice_processor_impl.cpp:35 */
      unembed4(dim, S4, S);
      orc_find(rs->index[s], dim, &sl->fixed, &sl->moving, S, &sl->finder, rs->fidx[s], rs->resp[s]);
      int64_t n = 0;
      for (int64_t j = 0; j < sl->moving.n; ++j) n += (rs->fidx[s][j] >= 0);
      rs->ncorr[s] = n;
      good |= ((int) n > sl->min_num_correspondences); /* :77-79, strict > */
    }
    if (!good) { /* :107-111 */
      rs->status = ORC_ALIGNER_NOT_ENOUGH_CORR;
      memcpy(rs->X4, backup, sizeof(backup));
      break;
    }
    /* solver->compute(): one GN iteration (multi_aligner.h:61-62) */
    double H[36], b[6];
    memset(H, 0, sizeof(H)); memset(b, 0, sizeof(b));
    orc_iter_stats st;
    memset(&st, 0, sizeof(st));
    st.iteration = rs->n_stats;
    for (int s = 0; s < rs->n_slices; ++s) {
      const orc_slice* sl = &rs->slices[s];
      if (sl->kind == ORC_SLICE_PRIOR) {
        float Z4[16];
        double chi = 0.0;
        embed4(dim, sl->prior_measurement, Z4);
        prior_terms(dim, rs->ap->variable, Z4, rs->X4, sl->prior_info_diag, H, b, &chi);
        st.num_inliers += 1; st.num_correspondences += 1; st.chi_inliers += chi;
        continue;
      }
      float ris4[16], S4[16], S[16];
      embed4(dim, sl->robot_in_sensor, ris4);
      mul4(ris4, rs->X4, S4);
      unembed4(dim, S4, S);
      orc_factor_params fa = sl->factor;
      if (clamp && fa.robustifier != ORC_ROB_NONE) fa.robustifier = ORC_ROB_CLAMP; /* :193-199 */
      double Hs[36], bs[6];
      orc_iter_stats ss;
      orc_linearize(dim, rs->ap->variable, &sl->fixed, &sl->moving, rs->fidx[s], S, &sl->finder, &fa, 0.f, 0.f,
                    NULL, Hs, bs, &ss, rs->fstat[s], NULL, NULL);
      for (int k = 0; k < P * P; ++k) H[k] = H[k] + Hs[k];
      for (int k = 0; k < P; ++k) b[k] = b[k] + bs[k];
      st.num_inliers += ss.num_inliers; st.num_outliers += ss.num_outliers;
      st.num_suppressed += ss.num_suppressed; st.num_correspondences += ss.num_correspondences;
      st.num_saturated += ss.num_saturated;
      st.chi_inliers += ss.chi_inliers; st.chi_outliers += ss.chi_outliers;
    }
    float Xn[16], Xu[16];
    memcpy(Xn, rs->X4, sizeof(Xn));
    unembed4(dim, Xn, Xu);
    st.solver_status = orc_solve_update(dim, rs->ap->variable, H, b, Xu);
    if (st.solver_status) embed4(dim, Xu, rs->X4);
    if (rs->n_stats < rs->cap_stats) rs->stats[rs->n_stats] = st; /* :113-115 */
    rs->n_stats++;
    if (st.solver_status) memcpy(backup, rs->X4, sizeof(backup)); /* :118-121 */
    int64_t total = 0;
    for (int s = 0; s < rs->n_slices; ++s) total += rs->ncorr[s];
    if (use_tc && tc_has_to_stop(&rs->tc, rs->ap, &st, total)) break; /* :124-126 */
  }
}

int orc_icp_run(int dim, int n_slices, const orc_slice* slices, const orc_aligner_params* ap,
                float* T, orc_iter_stats* stats_out, int32_t* n_stats, int32_t* status_out,
                orc_corr_out* corr_out, int nn_method) {
  run_state rs;
  memset(&rs, 0, sizeof(rs));
  rs.dim = dim; rs.n_slices = n_slices; rs.slices = slices; rs.ap = ap; rs.nn_method = nn_method;
  rs.index = (orc_index**) calloc((size_t) n_slices, sizeof(void*));
  rs.fidx = (int32_t**) calloc((size_t) n_slices, sizeof(void*));
  rs.resp = (float**) calloc((size_t) n_slices, sizeof(void*));
  rs.fstat = (uint8_t**) calloc((size_t) n_slices, sizeof(void*));
  rs.ncorr = (int64_t*) calloc((size_t) n_slices, sizeof(int64_t));
  rs.stats = stats_out; rs.cap_stats = *n_stats; rs.n_stats = 0;
  rs.status = ORC_ALIGNER_FAIL;
  embed4(dim, T, rs.X4);
  for (int s = 0; s < n_slices; ++s) {
    if (slices[s].kind != ORC_SLICE_POINTS) continue;
    if (slices[s].finder.kind == ORC_FINDER_NN) {
      rs.index[s] = orc_index_create(dim, &slices[s].fixed, nn_method);
    }
    rs.fidx[s] = (int32_t*) malloc(sizeof(int32_t) * (size_t) (slices[s].moving.n + 1));
    rs.resp[s] = (float*) malloc(sizeof(float) * (size_t) (slices[s].moving.n + 1));
    rs.fstat[s] = (uint8_t*) malloc((size_t) (slices[s].moving.n + 1));
    for (int64_t j = 0; j < slices[s].moving.n; ++j) rs.fidx[s][j] = -1;
  }
  if (ap->use_termination_criteria) tc_init(&rs.tc, ap); /* compute() :52-57 */
  /* _preCompute :130-141 -- prior slices overwrite the guess in slice order
   * (aligner_slice_odometry_prior.cpp:19,34; aligner_slice_motion_model.hpp:70) */
  for (int s = 0; s < n_slices; ++s) {
    if (slices[s].kind == ORC_SLICE_PRIOR) embed4(dim, slices[s].prior_measurement, rs.X4);
  }
  run_solver(&rs, ap->max_iterations, ap->use_termination_criteria, 0); /* :72 */
  int done = 0;
  if (rs.n_stats == 0) { /* :75-78 */
    rs.status = ORC_ALIGNER_FAIL;
    done = 1;
  }
  if (!done) {
    int last = rs.n_stats <= rs.cap_stats ? rs.n_stats - 1 : rs.cap_stats - 1;
    if (rs.stats[last].num_inliers < ap->min_num_inliers) { /* :81-85 */
      rs.status = ORC_ALIGNER_NOT_ENOUGH_INLIERS;
      done = 1;
    }
  }
  if (!done) {
    if (ap->enable_inlier_only_runs) { /* _postCompute :162-175 */
      run_solver(&rs, ap->max_iterations, ap->use_termination_criteria, 1);
    }
    float Tf[16];
    unembed4(dim, rs.X4, Tf);
    orc_fix_transform(dim, Tf); /* :90-93 */
    embed4(dim, Tf, rs.X4);
    rs.status = ORC_ALIGNER_SUCCESS;
  }
  unembed4(dim, rs.X4, T);
  *n_stats = rs.n_stats;
  *status_out = rs.status;
  if (corr_out) {
    for (int s = 0; s < n_slices; ++s) {
      corr_out[s].n = 0;
      if (slices[s].kind != ORC_SLICE_POINTS || !corr_out[s].fixed_idx) continue;
      const int prune = (!done && ap->keep_only_inlier_correspondences); /* :177-180, :213-263 */
      int64_t k = 0;
      for (int64_t j = 0; j < slices[s].moving.n; ++j) {
        if (rs.fidx[s][j] < 0) continue;
        if (prune && rs.fstat[s][j] != ORC_STAT_INLIER) continue;
        corr_out[s].fixed_idx[k] = rs.fidx[s][j];
        corr_out[s].moving_idx[k] = (int32_t) j;
        corr_out[s].response[k] = rs.resp[s][j];
        ++k;
      }
      corr_out[s].n = k;
    }
  }
  for (int s = 0; s < n_slices; ++s) {
    orc_index_free(rs.index[s]);
    free(rs.fidx[s]); free(rs.resp[s]); free(rs.fstat[s]);
  }
  free(rs.index); free(rs.fidx); free(rs.resp); free(rs.fstat); free(rs.ncorr);
  return 0;
}
