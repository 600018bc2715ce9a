// s2b_pgo.cuh -- device code of the pose-graph Gauss-Newton step:
// what MultiGraphSLAM_::optimize() reaches through Solver::compute()
// (R/system/multi_graph_slam_impl.cpp:299-317) for SE3PosePoseGeodesicErrorFactor
// (R/registration/loop_closure.h:110-111) between VariableSE3QuaternionRightAD poses.
//
//   k3  pgo_linearize_kernel   one thread per factor: e = t2v(Z^-1 Xi^-1 Xj), J_i, J_j, the four
//                              6x6 blocks J^T Omega J into a block-CSR matrix, J^T Omega e into b
//   k4  preconditioned conjugate gradients on the block-CSR matrix (block-Jacobi preconditioner),
//       three fused kernels per iteration, scalars stay on the device
//   k4u pgo_update_kernel      X_v <- X_v * v2t(dx_v)
// Arithmetic is fp64 throughout (documented deviation from the fp32 upstream: the parity bars are
// tolerances here, and CG needs the head-room); poses are kept as 3x4 doubles.
#pragma once
#include "s2b_math.cuh"

namespace s2b {

struct PgoScalars {       // two parities of {rz, pAp, rr} + bookkeeping
  double rz[2], pAp[2], rr[2];
  double chi, b_norm2, dx_max;
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

// poses: V x 12 (R row-major, then t); Zm: F x 12; Om: F x 36
__global__ void __launch_bounds__(128) pgo_linearize_kernel(const double* __restrict__ poses, const unsigned char* __restrict__ fixed,
                                                            const int* __restrict__ ij, const double* __restrict__ Zm,
                                                            const double* __restrict__ Om, const int* __restrict__ slots,
                                                            int F, int f_begin, int f_stride, double* __restrict__ vals,
                                                            double* __restrict__ b, PgoScalars* __restrict__ sc) {
  const int f = f_begin + (blockIdx.x * blockDim.x + threadIdx.x) * f_stride;
  double chi = 0.0;
  if (f < F) {
    const int vi = ij[2 * f], vj = ij[2 * f + 1];
    const double* Xi = poses + (size_t) vi * 12;
    const double* Xj = poses + (size_t) vj * 12;
    const double* Z = Zm + (size_t) f * 12;
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
    double e[6] = {te[0], te[1], te[2], qe[0], qe[1], qe[2]};
    double Ji[36], Jj[36];
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
    const double* O = Om + (size_t) f * 36;
    double Oe[6], OJi[36], OJj[36];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += O[r * 6 + k] * e[k];
      Oe[r] = s;
      chi += e[r] * s;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double si = 0.0, sj = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) { si += O[r * 6 + k] * Ji[k * 6 + c]; sj += O[r * 6 + k] * Jj[k * 6 + c]; }
        OJi[r * 6 + c] = si;
        OJj[r * 6 + c] = sj;
      }
    }
    const bool fi = fixed[vi] != 0, fj = fixed[vj] != 0;
    const int sii = slots[4 * f], sij = slots[4 * f + 1], sji = slots[4 * f + 2], sjj = slots[4 * f + 3];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double bi = 0.0, bj = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) { bi += Ji[k * 6 + r] * Oe[k]; bj += Jj[k * 6 + r] * Oe[k]; }
      if (!fi) atomicAdd(&b[(size_t) vi * 6 + r], bi);
      if (!fj) atomicAdd(&b[(size_t) vj * 6 + r], bj);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double hii = 0.0, hij = 0.0, hjj = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          hii += Ji[k * 6 + r] * OJi[k * 6 + c];
          hij += Ji[k * 6 + r] * OJj[k * 6 + c];
          hjj += Jj[k * 6 + r] * OJj[k * 6 + c];
        }
        if (!fi) atomicAdd(&vals[(size_t) sii * 36 + r * 6 + c], hii);
        if (!fj) atomicAdd(&vals[(size_t) sjj * 36 + r * 6 + c], hjj);
        if (!fi && !fj) {
          atomicAdd(&vals[(size_t) sij * 36 + r * 6 + c], hij);
          atomicAdd(&vals[(size_t) sji * 36 + c * 6 + r], hij);  // H_ji = H_ij^T
        }
      }
    }
  }
  chi = [](double v) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  }(chi);
  if ((threadIdx.x & 31) == 0 && chi != 0.0) atomicAdd(&sc->chi, chi);
}

// gauge: the diagonal block of a fixed variable is the identity (its other blocks stay zero)
__global__ void pgo_fix_diag_kernel(const unsigned char* __restrict__ fixed, const int* __restrict__ diag_slot, int V,
                                    double* __restrict__ vals) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V || !fixed[v]) return;
  for (int k = 0; k < 6; ++k) vals[(size_t) diag_slot[v] * 36 + k * 7] = 1.0;
}

// block-Jacobi preconditioner: inverse of every 6x6 diagonal block (via LL^T); identity on failure
__global__ void pgo_block_inverse_kernel(const double* __restrict__ vals, const int* __restrict__ diag_slot, int V,
                                         double* __restrict__ Minv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double H[36], col[6], e[6];
  for (int k = 0; k < 36; ++k) H[k] = vals[(size_t) diag_slot[v] * 36 + k];
  for (int c = 0; c < 6; ++c) {
    for (int k = 0; k < 6; ++k) e[k] = (k == c) ? -1.0 : 0.0;  // spd_solve solves H x = -b
    const bool ok = spd_solve_t<6>(H, e, col);
    for (int r = 0; r < 6; ++r) Minv[(size_t) v * 36 + r * 6 + c] = ok ? col[r] : (r == c ? 1.0 : 0.0);
  }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  }
  return v;  // valid in thread 0
}

// CG start: x = 0, r = -b, z = M^-1 r, p = z, rz[0] = r.z, rr[0] = r.r
__global__ void pgo_cg_init_kernel(const double* __restrict__ b, const double* __restrict__ Minv, int n, double* x,
                                   double* r, double* z, double* p, PgoScalars* sc) {
  __shared__ double sh[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double rz = 0.0, rr = 0.0;
  if (i < n) {
    const int v = i / 6, c = i - 6 * v;
    double zi = 0.0;
    for (int k = 0; k < 6; ++k) zi += Minv[(size_t) v * 36 + c * 6 + k] * (-b[v * 6 + k]);
    const double ri = -b[i];
    x[i] = 0.0; r[i] = ri; z[i] = zi; p[i] = zi;
    rz = ri * zi; rr = ri * ri;
  }
  const double a = block_sum(rz, sh);
  __syncthreads();
  const double c2 = block_sum(rr, sh);
  if (threadIdx.x == 0) { atomicAdd(&sc->rz[0], a); atomicAdd(&sc->rr[0], c2); atomicAdd(&sc->b_norm2, c2); }
}

// k1: Ap = A p (one thread per scalar row of the block-CSR matrix), pAp[par] += p.Ap
__global__ void pgo_cg_spmv_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                   const double* __restrict__ vals, const double* __restrict__ p, int n, int par,
                                   double* __restrict__ Ap, PgoScalars* sc) {
  __shared__ double sh[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { sc->rz[par ^ 1] = 0.0; sc->pAp[par ^ 1] = 0.0; sc->rr[par ^ 1] = 0.0; }
  double acc = 0.0, pi = 0.0;
  if (i < n) {
    const int v = i / 6, c = i - 6 * v;
    for (int s = row_ptr[v]; s < row_ptr[v + 1]; ++s) {
      const double* blk = vals + (size_t) s * 36 + c * 6;
      const double* pv = p + (size_t) col_idx[s] * 6;
      acc += blk[0] * pv[0] + blk[1] * pv[1] + blk[2] * pv[2] + blk[3] * pv[3] + blk[4] * pv[4] + blk[5] * pv[5];
    }
    Ap[i] = acc;
    pi = p[i];
  }
  const double d = block_sum(pi * acc, sh);
  if (threadIdx.x == 0 && d != 0.0) atomicAdd(&sc->pAp[par], d);
}

// k2: alpha = rz/pAp; x += alpha p; r -= alpha Ap; z = M^-1 r; rz[par^1] += r.z; rr[par^1] += r.r
__global__ void pgo_cg_update_kernel(const double* __restrict__ Minv, const double* __restrict__ p,
                                     const double* __restrict__ Ap, int n, int par, double* x, double* r, double* z,
                                     PgoScalars* sc) {
  __shared__ double sh[32];
  __shared__ double rs[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // blockDim.x is a multiple of 6
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
    const int v = i / 6, c = i - 6 * v;
    const int base = threadIdx.x - c;
    double zi = 0.0;
    for (int k = 0; k < 6; ++k) zi += Minv[(size_t) v * 36 + c * 6 + k] * rs[base + k];
    z[i] = zi;
    rz = ri * zi; rr = ri * ri;
  }
  const double a = block_sum(rz, sh);
  __syncthreads();
  const double c2 = block_sum(rr, sh);
  if (threadIdx.x == 0) { atomicAdd(&sc->rz[par ^ 1], a); atomicAdd(&sc->rr[par ^ 1], c2); }
}

// k3: beta = rz_new / rz_old; p = z + beta p
__global__ void pgo_cg_direction_kernel(const double* __restrict__ z, int n, int par, double* p, const PgoScalars* sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double rz_old = sc->rz[par];
  const double beta = rz_old > 0.0 ? sc->rz[par ^ 1] / rz_old : 0.0;
  if (i < n) p[i] = z[i] + beta * p[i];
}

// X_v <- X_v * v2t(dx_v) (VariableSE3QuaternionRight::applyPerturbation), |dx|_inf into scalars
__global__ void pgo_update_kernel(const double* __restrict__ dx, const unsigned char* __restrict__ fixed, int V,
                                  double* __restrict__ poses, PgoScalars* sc) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (v < V && !fixed[v]) {
    double d[6], D[12], X[12], o[12];
    for (int k = 0; k < 6; ++k) { d[k] = dx[(size_t) v * 6 + k]; m = fmax(m, fabs(d[k])); }
    exp_right(3, 0, d, D);
    for (int k = 0; k < 12; ++k) X[k] = poses[(size_t) v * 12 + k];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) o[i * 3 + j] = X[i * 3] * D[j] + X[i * 3 + 1] * D[4 + j] + X[i * 3 + 2] * D[8 + j];
      o[9 + i] = X[i * 3] * D[3] + X[i * 3 + 1] * D[7] + X[i * 3 + 2] * D[11] + X[9 + i];
    }
    for (int k = 0; k < 12; ++k) poses[(size_t) v * 12 + k] = o[k];
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

__global__ void pgo_cast_kernel(const float* __restrict__ in, size_t n, double* __restrict__ out) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double) in[i];
}

}  // namespace s2b
