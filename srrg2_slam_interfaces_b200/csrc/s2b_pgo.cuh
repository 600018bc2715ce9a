// s2b_pgo.cuh -- device code of the pose-graph Gauss-Newton / Levenberg-Marquardt step (SURVEY.md section 8
// row a10): what MultiGraphSLAM_::optimize() reaches through Solver::compute()
// (R/system/multi_graph_slam_impl.cpp:299-317) for SE{2,3}PosePoseGeodesicErrorFactor
// (R/registration/loop_closure.h:110-111) between VariableSE2RightAD / VariableSE3QuaternionRightAD poses
// (R/mapping/local_map.h:64,75).  D = 6 (SE(3)) or 3 (SE(2)) degrees of freedom per pose.
//
//   k3a pgo_factor_kernel    one thread per factor of this rank: e = t2v(Z^-1 Xi^-1 Xj), J_i, J_j, the blocks
//                            J^T Omega J (ii, ij, jj), J^T Omega e (i, j) and chi into a PER-FACTOR scratch record
//   k3b pgo_gather_*         one thread per matrix / vector entry sums the records of the factors that touch it,
//                            in factor order: the block-CSR matrix and b are assembled WITHOUT atomics, bit for
//                            bit the same on every run (and for every thread-block shape)
//   k4  preconditioned conjugate gradients on the block-CSR matrix (block-Jacobi preconditioner), three fused
//       kernels per iteration; the dot products are two-stage sums in a fixed order (per-CTA partials, the last
//       CTA to arrive adds them up), scalars stay on the device
//   k4u pgo_update_kernel    X_v <- X_v * v2t(dx_v)
// Arithmetic is fp64 throughout (documented deviation from the fp32 upstream: the parity bars are
// tolerances here, and CG needs the head-room).  SE(3) poses are kept as 3x4 doubles (R row-major, then t),
// SE(2) poses as (x, y, theta).
#pragma once
#include "s2b_math.cuh"

namespace s2b {

constexpr int kPgoMaxParts = 8192;  // per-CTA partial sums of a dot product (n / 192 CTAs: up to 1.5M scalars)

struct PgoScalars {       // two parities of {rz, pAp, rr} + bookkeeping
  double rz[2], pAp[2], rr[2];
  double chi, b_norm2, dx_max;
  double dot_b_dx, dot_dx_Hdx, dot_dx_Ddx;  // gain-ratio terms of the Levenberg-Marquardt step
  unsigned int ticket[8];
};

template <int D>
struct PgoDim {
  static constexpr int BB = D * D;
  static constexpr int POSE = (D == 6) ? 12 : 3;
  static constexpr int REC = 3 * BB + 2 * D;  // scratch record of a factor: H_ii, H_ij, H_jj, b_i, b_j
};

__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3_mulT(const double* A, const double* B, double* C) {  // A^T B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
__device__ __forceinline__ void quat_mul(const double* a, const double* b, double* o) {  // x y z w
  o[0] = a[3] * b[0] + b[3] * a[0] + (a[1] * b[2] - a[2] * b[1]);
  o[1] = a[3] * b[1] + b[3] * a[1] + (a[2] * b[0] - a[0] * b[2]);
  o[2] = a[3] * b[2] + b[3] * a[2] + (a[0] * b[1] - a[1] * b[0]);
  o[3] = a[3] * b[3] - (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
}

// error and Jacobians of one pose-pose factor, right perturbation X <- X v2t(dx)
template <int D>
__device__ __forceinline__ void pgo_factor_terms(const double* Xi, const double* Xj, const double* Z, double* e, double* Ji, double* Jj);

template <>
__device__ __forceinline__ void pgo_factor_terms<6>(const double* Xi, const double* Xj, const double* Z, double* e, double* Ji,
                                                    double* Jj) {
  double Ra[9], ta[3], Rzi[9], tzi[3], Re[9], te[3];
  {  // A = Xi^-1 Xj
    mat3_mulT(Xi, Xj, Ra);
    const double d[3] = {Xj[9] - Xi[9], Xj[10] - Xi[10], Xj[11] - Xi[11]};
#pragma unroll
    for (int k = 0; k < 3; ++k) ta[k] = Xi[k] * d[0] + Xi[3 + k] * d[1] + Xi[6 + k] * d[2];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {  // Z^-1
#pragma unroll
    for (int j = 0; j < 3; ++j) Rzi[i * 3 + j] = Z[j * 3 + i];
    tzi[i] = -(Z[i] * Z[9] + Z[3 + i] * Z[10] + Z[6 + i] * Z[11]);
  }
  mat3_mul(Rzi, Ra, Re);
#pragma unroll
  for (int i = 0; i < 3; ++i) te[i] = Rzi[i * 3] * ta[0] + Rzi[i * 3 + 1] * ta[1] + Rzi[i * 3 + 2] * ta[2] + tzi[i];
  double qe[4], qz[4], qa[4], prod[4];
  quat_of(Re, qe);
  quat_of(Rzi, qz);
  quat_of(Ra, qa);
  quat_mul(qz, qa, prod);
  const double sgn = (prod[0] * qe[0] + prod[1] * qe[1] + prod[2] * qe[2] + prod[3] * qe[3]) < 0.0 ? -1.0 : 1.0;
  e[0] = te[0]; e[1] = te[1]; e[2] = te[2]; e[3] = qe[0]; e[4] = qe[1]; e[5] = qe[2];
#pragma unroll
  for (int k = 0; k < 36; ++k) { Ji[k] = 0.0; Jj[k] = 0.0; }
  // J_j = [R_e 0; 0 w I + [v]x]
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Jj[i * 6 + j] = Re[i * 3 + j];
  Jj[21] = qe[3];  Jj[22] = -qe[2]; Jj[23] = qe[1];
  Jj[27] = qe[2];  Jj[28] = qe[3];  Jj[29] = -qe[0];
  Jj[33] = -qe[1]; Jj[34] = qe[0];  Jj[35] = qe[3];
  // J_i: translation rows [-R_z^-1 | 2 R_z^-1 [t_a]x], rotation rows [0 | -sgn vec(q_z^-1 (x) e_k (x) q_a)]
  const double Sk[9] = {0.0, -ta[2], ta[1], ta[2], 0.0, -ta[0], -ta[1], ta[0], 0.0};
  double RS[9];
  mat3_mul(Rzi, Sk, RS);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Ji[i * 6 + j] = -Rzi[i * 3 + j];
      Ji[i * 6 + 3 + j] = 2.0 * RS[i * 3 + j];
    }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double ek[4] = {0.0, 0.0, 0.0, 0.0}, t1[4], t2[4];
    ek[k] = 1.0;
    quat_mul(qz, ek, t1);
    quat_mul(t1, qa, t2);
    Ji[18 + 3 + k] = -sgn * t2[0];
    Ji[24 + 3 + k] = -sgn * t2[1];
    Ji[30 + 3 + k] = -sgn * t2[2];
  }
}

// SE(2): X = (x, y, theta).  A = Xi^-1 Xj, E = Z^-1 A, e = (t_E, theta_E);
// J_j = [R_E 0; 0 1],  J_i = [-R_z^T, -R_z^T S t_A; 0, -1]  with S = [0 -1; 1 0]
template <>
__device__ __forceinline__ void pgo_factor_terms<3>(const double* Xi, const double* Xj, const double* Z, double* e, double* Ji,
                                                    double* Jj) {
  double si, ci, sz, cz;
  sincos_det(Xi[2], si, ci);
  sincos_det(Z[2], sz, cz);
  const double dx = Xj[0] - Xi[0], dy = Xj[1] - Xi[1];
  const double tax = ci * dx + si * dy, tay = -si * dx + ci * dy;  // R_i^T (t_j - t_i)
  const double tha = Xj[2] - Xi[2];
  const double ux = tax - Z[0], uy = tay - Z[1];
  e[0] = cz * ux + sz * uy;
  e[1] = -sz * ux + cz * uy;
  double th = tha - Z[2], s, c;
  sincos_det(th, s, c);
  e[2] = atan2_det(s, c);  // angle of R_E, in (-pi, pi]
  Jj[0] = c;  Jj[1] = -s; Jj[2] = 0.0;
  Jj[3] = s;  Jj[4] = c;  Jj[5] = 0.0;
  Jj[6] = 0.0; Jj[7] = 0.0; Jj[8] = 1.0;
  const double stx = -tay, sty = tax;  // S t_A
  Ji[0] = -cz; Ji[1] = -sz; Ji[2] = -(cz * stx + sz * sty);
  Ji[3] = sz;  Ji[4] = -cz; Ji[5] = -(-sz * stx + cz * sty);
  Ji[6] = 0.0; Ji[7] = 0.0; Ji[8] = -1.0;
}

// k3a: one thread per factor of this rank (f = f_begin + l f_stride): the factor's scratch record and chi.
// chi_only: candidate evaluation of the Levenberg-Marquardt step (no record).
template <int D>
__global__ void __launch_bounds__(128) pgo_factor_kernel(const double* __restrict__ poses, const int* __restrict__ ij,
                                                         const double* __restrict__ Zm, const double* __restrict__ Om, int F,
                                                         int f_begin, int f_stride, int chi_only, double* __restrict__ rec,
                                                         double* __restrict__ chi_f) {
  constexpr int BB = PgoDim<D>::BB, POSE = PgoDim<D>::POSE, REC = PgoDim<D>::REC;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = f_begin + l * f_stride;
  if (f >= F) return;
  const int vi = ij[2 * f], vj = ij[2 * f + 1];
  double e[D], Ji[BB], Jj[BB];
  pgo_factor_terms<D>(poses + (size_t) vi * POSE, poses + (size_t) vj * POSE, Zm + (size_t) f * POSE, e, Ji, Jj);
  const double* O = Om + (size_t) f * BB;
  double Oe[D], OJi[BB], OJj[BB], chi = 0.0;
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) s += O[r * D + k] * e[k];
    Oe[r] = s;
    chi += e[r] * s;
    if (!chi_only) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double si = 0.0, sj = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) { si += O[r * D + k] * Ji[k * D + c]; sj += O[r * D + k] * Jj[k * D + c]; }
        OJi[r * D + c] = si;
        OJj[r * D + c] = sj;
      }
    }
  }
  chi_f[l] = chi;
  if (chi_only) return;
  double* R = rec + (size_t) l * REC;
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double bi = 0.0, bj = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) { bi += Ji[k * D + r] * Oe[k]; bj += Jj[k * D + r] * Oe[k]; }
    R[3 * BB + r] = bi;
    R[3 * BB + D + r] = bj;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      double hii = 0.0, hij = 0.0, hjj = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        hii += Ji[k * D + r] * OJi[k * D + c];
        hij += Ji[k * D + r] * OJj[k * D + c];
        hjj += Jj[k * D + r] * OJj[k * D + c];
      }
      R[r * D + c] = hii;
      R[BB + r * D + c] = hij;
      R[2 * BB + r * D + c] = hjj;
    }
  }
}

// k3b: entry e of block-CSR slot s = sum over the records that touch the slot, in factor order.
// src code = 4 l + which (0: H_ii, 1: H_ij, 2: H_ij^T, 3: H_jj).  Rows / columns of fixed variables stay zero.
template <int D>
__global__ void pgo_gather_blocks_kernel(const int* __restrict__ blk_ptr, const int* __restrict__ blk_src,
                                         const int* __restrict__ row_of, const int* __restrict__ col_idx,
                                         const unsigned char* __restrict__ fixed, const double* __restrict__ rec, int nnzb,
                                         double* __restrict__ vals) {
  constexpr int BB = PgoDim<D>::BB, REC = PgoDim<D>::REC;
  const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long) nnzb * BB) return;
  const int s = (int) (g / BB), e = (int) (g - (long long) s * BB);
  const int r = e / D, c = e - r * D;
  double v = 0.0;
  if (!fixed[row_of[s]] && !fixed[col_idx[s]]) {
    for (int k = blk_ptr[s]; k < blk_ptr[s + 1]; ++k) {
      const int code = blk_src[k], l = code >> 2, which = code & 3;
      const double* R = rec + (size_t) l * REC;
      v += which == 0 ? R[e] : (which == 1 ? R[BB + e] : (which == 2 ? R[BB + c * D + r] : R[2 * BB + e]));
    }
  }
  vals[g] = v;
}

// entry k of b_v = sum over the records of the factors at v, in factor order (src code = 2 l + side)
template <int D>
__global__ void pgo_gather_b_kernel(const int* __restrict__ var_ptr, const int* __restrict__ var_src,
                                    const unsigned char* __restrict__ fixed, const double* __restrict__ rec, int V,
                                    double* __restrict__ b) {
  constexpr int BB = PgoDim<D>::BB, REC = PgoDim<D>::REC;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= V * D) return;
  const int v = g / D, k = g - v * D;
  double s = 0.0;
  if (!fixed[v]) {
    for (int q = var_ptr[v]; q < var_ptr[v + 1]; ++q) {
      const int code = var_src[q];
      s += rec[(size_t) (code >> 1) * REC + 3 * BB + (code & 1) * D + k];
    }
  }
  b[g] = s;
}

// deterministic sum of n doubles by ONE CTA (fixed strided partials, fixed tree)
__global__ void __launch_bounds__(1024) pgo_sum_kernel(const double* __restrict__ x, int n, double* __restrict__ out) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += x[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 512; w; w >>= 1) {
    if ((int) threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// Levenberg-Marquardt damping of the diagonal: keep the undamped diagonal entries, scale by (1 + lambda)
template <int D>
__global__ void pgo_damp_kernel(const int* __restrict__ diag_slot, int V, double lambda, int save, double* __restrict__ diag0,
                                double* __restrict__ vals) {
  constexpr int BB = PgoDim<D>::BB;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= V * D) return;
  const int v = g / D, k = g - v * D;
  double* p = vals + (size_t) diag_slot[v] * BB + k * (D + 1);
  if (save) diag0[g] = *p;
  *p = diag0[g] * (1.0 + lambda);
}

// gauge: the diagonal block of a fixed variable is the identity (its other blocks stay zero)
template <int D>
__global__ void pgo_fix_diag_kernel(const unsigned char* __restrict__ fixed, const int* __restrict__ diag_slot, int V,
                                    double* __restrict__ vals) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V || !fixed[v]) return;
  for (int k = 0; k < D; ++k) vals[(size_t) diag_slot[v] * PgoDim<D>::BB + k * (D + 1)] = 1.0;
}

// block-Jacobi preconditioner: inverse of every DxD diagonal block (via LL^T); identity on failure
template <int D>
__global__ void pgo_block_inverse_kernel(const double* __restrict__ vals, const int* __restrict__ diag_slot, int V,
                                         double* __restrict__ Minv) {
  constexpr int BB = PgoDim<D>::BB;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double H[BB], col[6], e[D];
  for (int k = 0; k < BB; ++k) H[k] = vals[(size_t) diag_slot[v] * BB + k];
  for (int c = 0; c < D; ++c) {
    for (int k = 0; k < D; ++k) e[k] = (k == c) ? -1.0 : 0.0;  // spd_solve solves H x = -b
    const bool ok = spd_solve_t<D>(H, e, col);
    for (int r = 0; r < D; ++r) Minv[(size_t) v * BB + r * D + c] = ok ? col[r] : (r == c ? 1.0 : 0.0);
  }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  }
  return v;  // valid in thread 0
}

// Second stage of a grid-wide sum in a FIXED order: every CTA parks its partial, the last CTA to arrive (ticket)
// adds the partials up in index order (one warp: lane-strided partial sums, then a shuffle tree) and publishes.
// `v` is the CTA's partial (thread 0); returns true in thread 0 of the last CTA with the total in `total`.
__device__ __forceinline__ bool grid_sum_ordered(double v, double* parts, unsigned int* ticket, double* sh, double& total) {
  __shared__ int s_last;
  __syncthreads();  // (the flag of a previous call in the same kernel has been read by everybody)
  if (threadIdx.x == 0) {
    parts[blockIdx.x] = v;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double s = 0.0;
  if (threadIdx.x < 32) {
    for (int k = threadIdx.x; k < (int) gridDim.x; k += 32) s += __ldcg(parts + k);
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  }
  if (threadIdx.x == 0) { total = s; *ticket = 0u; }
  return threadIdx.x == 0;
}

// CG start: x = 0, r = -b, z = M^-1 r, p = z, rz[0] = r.z, rr[0] = r.r
template <int D>
__global__ void pgo_cg_init_kernel(const double* __restrict__ b, const double* __restrict__ Minv, int n, double* x,
                                   double* r, double* z, double* p, PgoScalars* sc, double* parts) {
  __shared__ double sh[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double rz = 0.0, rr = 0.0;
  if (i < n) {
    const int v = i / D, c = i - D * v;
    double zi = 0.0;
    for (int k = 0; k < D; ++k) zi += Minv[(size_t) v * PgoDim<D>::BB + c * D + k] * (-b[v * D + k]);
    const double ri = -b[i];
    x[i] = 0.0; r[i] = ri; z[i] = zi; p[i] = zi;
    rz = ri * zi; rr = ri * ri;
  }
  const double a = block_sum(rz, sh);
  const double c2 = block_sum(rr, sh);
  double t;
  if (grid_sum_ordered(a, parts, &sc->ticket[0], sh, t)) sc->rz[0] = t;
  if (grid_sum_ordered(c2, parts + kPgoMaxParts, &sc->ticket[1], sh, t)) { sc->rr[0] = t; sc->b_norm2 = t; }
}

// k1: Ap = A p (one thread per scalar row of the block-CSR matrix), pAp[par] = p.Ap
template <int D>
__global__ void pgo_cg_spmv_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                   const double* __restrict__ vals, const double* __restrict__ p, int n, int par,
                                   double* __restrict__ Ap, PgoScalars* sc, double* parts) {
  __shared__ double sh[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0, pi = 0.0;
  if (i < n) {
    const int v = i / D, c = i - D * v;
    // (unrolled by 4 blocks: the loads of four blocks and their p entries are in flight together; the adds keep
    // their order, so the sum is the same to the bit)
    const int s_end = row_ptr[v + 1];
#pragma unroll 4
    for (int s = row_ptr[v]; s < s_end; ++s) {
      const double* blk = vals + (size_t) s * PgoDim<D>::BB + c * D;
      const double* pv = p + (size_t) __ldg(col_idx + s) * D;
#pragma unroll
      for (int k = 0; k < D; ++k) acc += __ldg(blk + k) * pv[k];
    }
    Ap[i] = acc;
    pi = p[i];
  }
  const double d = block_sum(pi * acc, sh);
  double t;
  if (grid_sum_ordered(d, parts, &sc->ticket[2], sh, t)) sc->pAp[par] = t;
}

// k2: alpha = rz/pAp; x += alpha p; r -= alpha Ap; z = M^-1 r; rz[par^1] = r.z; rr[par^1] = r.r
template <int D>
__global__ void pgo_cg_update_kernel(const double* __restrict__ Minv, const double* __restrict__ p,
                                     const double* __restrict__ Ap, int n, int par, double* x, double* r, double* z,
                                     PgoScalars* sc, double* parts) {
  __shared__ double sh[32];
  __shared__ double rs[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // blockDim.x is a multiple of D
  const double pAp = sc->pAp[par];
  const double alpha = pAp > 0.0 ? sc->rz[par] / pAp : 0.0;
  double ri = 0.0;
  if (i < n) {
    x[i] += alpha * p[i];
    ri = r[i] - alpha * Ap[i];
    r[i] = ri;
  }
  rs[threadIdx.x] = ri;
  __syncthreads();
  double rz = 0.0, rr = 0.0;
  if (i < n) {
    const int v = i / D, c = i - D * v;
    const int base = threadIdx.x - c;
    double zi = 0.0;
    for (int k = 0; k < D; ++k) zi += Minv[(size_t) v * PgoDim<D>::BB + c * D + k] * rs[base + k];
    z[i] = zi;
    rz = ri * zi; rr = ri * ri;
  }
  const double a = block_sum(rz, sh);
  const double c2 = block_sum(rr, sh);
  double t;
  if (grid_sum_ordered(a, parts, &sc->ticket[3], sh, t)) sc->rz[par ^ 1] = t;
  if (grid_sum_ordered(c2, parts + kPgoMaxParts, &sc->ticket[4], sh, t)) sc->rr[par ^ 1] = t;
}

// k3: beta = rz_new / rz_old; p = z + beta p
__global__ void pgo_cg_direction_kernel(const double* __restrict__ z, int n, int par, double* p, const PgoScalars* sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double rz_old = sc->rz[par];
  const double beta = rz_old > 0.0 ? sc->rz[par ^ 1] / rz_old : 0.0;
  if (i < n) p[i] = z[i] + beta * p[i];
}

// gain-ratio terms of the damped step: b.dx, dx.(H_damped dx), dx.(D dx) with D = the undamped diagonal
template <int D>
__global__ void pgo_model_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                 const double* __restrict__ vals, const double* __restrict__ dx, const double* __restrict__ b,
                                 const double* __restrict__ diag0, int n, PgoScalars* sc, double* parts) {
  __shared__ double sh[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0, di = 0.0, bi = 0.0, d0 = 0.0;
  if (i < n) {
    const int v = i / D, c = i - D * v;
    for (int s = row_ptr[v]; s < row_ptr[v + 1]; ++s) {
      const double* blk = vals + (size_t) s * PgoDim<D>::BB + c * D;
      const double* pv = dx + (size_t) col_idx[s] * D;
#pragma unroll
      for (int k = 0; k < D; ++k) acc += blk[k] * pv[k];
    }
    di = dx[i]; bi = b[i]; d0 = diag0[i];
  }
  const double s0 = block_sum(bi * di, sh);
  const double s1 = block_sum(di * acc, sh);
  const double s2 = block_sum(di * d0 * di, sh);
  double t;
  if (grid_sum_ordered(s0, parts, &sc->ticket[5], sh, t)) sc->dot_b_dx = t;
  if (grid_sum_ordered(s1, parts + kPgoMaxParts, &sc->ticket[6], sh, t)) sc->dot_dx_Hdx = t;
  if (grid_sum_ordered(s2, parts + 2 * kPgoMaxParts, &sc->ticket[7], sh, t)) sc->dot_dx_Ddx = t;
}

// X_v <- X_v * v2t(dx_v) (VariableSE{2,3}*Right::applyPerturbation), |dx|_inf into scalars (a max: any order)
template <int D>
__global__ void pgo_update_kernel(const double* __restrict__ dx, const unsigned char* __restrict__ fixed, int V,
                                  const double* poses_in, double* poses, PgoScalars* sc) {
  constexpr int POSE = PgoDim<D>::POSE;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (v < V) {
    double X[POSE];
    for (int k = 0; k < POSE; ++k) X[k] = poses_in[(size_t) v * POSE + k];
    if (!fixed[v]) {
      double d[6] = {0, 0, 0, 0, 0, 0};
      for (int k = 0; k < D; ++k) { d[k] = dx[(size_t) v * D + k]; m = fmax(m, fabs(d[k])); }
      if (D == 6) {
        double Dm[12], o[12];
        exp_right(3, 0, d, Dm);
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) o[i * 3 + j] = X[i * 3] * Dm[j] + X[i * 3 + 1] * Dm[4 + j] + X[i * 3 + 2] * Dm[8 + j];
          o[9 + i] = X[i * 3] * Dm[3] + X[i * 3 + 1] * Dm[7] + X[i * 3 + 2] * Dm[11] + X[9 + i];
        }
        for (int k = 0; k < 12; ++k) X[k] = o[k];
      } else {
        double s, c;
        sincos_det(X[2], s, c);
        const double x = X[0] + c * d[0] - s * d[1], y = X[1] + s * d[0] + c * d[1];
        double th = X[2] + d[2], s2, c2;
        sincos_det(th, s2, c2);
        X[0] = x; X[1] = y; X[2] = atan2_det(s2, c2);
      }
    }
    for (int k = 0; k < POSE; ++k) poses[(size_t) v * POSE + k] = X[k];
  }
  for (int off = 16; off; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.0) {
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(&sc->dx_max), (unsigned long long) __double_as_longlong(m));
  }
}

__global__ void pgo_pack_poses_kernel(const float* __restrict__ in16, int V, double* __restrict__ out12) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float* T = in16 + (size_t) v * 16;
  double* o = out12 + (size_t) v * 12;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) o[i * 3 + j] = (double) T[i * 4 + j];
    o[9 + i] = (double) T[i * 4 + 3];
  }
}

__global__ void pgo_unpack_poses_kernel(const double* __restrict__ in12, int V, float* __restrict__ out16) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const double* o = in12 + (size_t) v * 12;
  float* T = out16 + (size_t) v * 16;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[i * 4 + j] = (float) o[i * 3 + j];
    T[i * 4 + 3] = (float) o[9 + i];
  }
  T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}

// SE(2): 3x3 float matrices <-> (x, y, theta)
__global__ void pgo_pack_poses2_kernel(const float* __restrict__ in9, int V, double* __restrict__ out3) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float* T = in9 + (size_t) v * 9;
  out3[(size_t) v * 3] = (double) T[2];
  out3[(size_t) v * 3 + 1] = (double) T[5];
  out3[(size_t) v * 3 + 2] = atan2_det((double) T[3], (double) T[0]);
}
__global__ void pgo_unpack_poses2_kernel(const double* __restrict__ in3, int V, float* __restrict__ out9) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double s, c;
  sincos_det(in3[(size_t) v * 3 + 2], s, c);
  float* T = out9 + (size_t) v * 9;
  T[0] = (float) c; T[1] = (float) -s; T[2] = (float) in3[(size_t) v * 3];
  T[3] = (float) s; T[4] = (float) c;  T[5] = (float) in3[(size_t) v * 3 + 1];
  T[6] = 0.f; T[7] = 0.f; T[8] = 1.f;
}

__global__ void pgo_cast_kernel(const float* __restrict__ in, size_t n, double* __restrict__ out) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double) in[i];
}

}  // namespace s2b
