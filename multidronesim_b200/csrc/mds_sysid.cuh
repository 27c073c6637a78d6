// mds_sysid.cuh -- per-drone recursive-least-squares model learning and the per-drone-gain LQR law of the
// reference's decentralised LQR:
//   control/dlqr/decentralized_lqr_omega.py:125-139 (theta_update), :110-123 (theta_update2), :185-204 (compute)
//   control/dlqr/decentralized_lqr_yank_omega.py:112-126
//   control/dlqr/decentralized_lqr.py:132-183 (theta_update2 / theta_update), :185-240 (est_x_dot, approx_theta_update,
//   project_theta), control/dlqr/decentralized_yolqr_crazyflie.py:228-290 (same for the 10-dim yank model)
// One thread per drone.  theta [(m+4) x m] and P [(m+4) x (m+4)] live in HBM as planes of D Reals (element k of drone d at
// [k * D + d]): every access of a warp is one coalesced 128-byte line.  Both matrices are swept twice (gain / prediction
// first, rank-1 update second); the second sweep re-reads lines the same SM touched microseconds earlier (L1 / L2 hits),
// so HBM sees each matrix once in and once out.
#pragma once
#include "mds_common.cuh"

namespace mds {

struct RlsP {
  int target;             // MDS_RLS_TARGET_PREDICT / MDS_RLS_TARGET_XDOT
  int predict_from_xtp1;  // PREDICT: integrate from x_{t+1} (omega / yank-omega quirk) instead of phi[:m]
  int normalize_gain;     // 1: L = P phi / (1 + phi' P phi) (RLS); 0: L = P phi (information form, P holds V^-1)
  int project;            // MDS_RLS_PROJECT_*
  int drones_per_env;
  double dt;
  unsigned long long zero_mask[3], one_mask[3];  // project_theta: bit k of entry k = i * m + j of theta [(m+4)][m]: force 0 / force 1
};

// project_theta (decentralized_lqr.py:230-240, decentralized_yolqr_crazyflie.py:245-257) on one drone's theta planes.
// A run-time loop over bit masks: unrolled, the ~190 predicated stores keep as many addresses live (255 registers).
template <typename Real, int M> MDS_DEV void rls_project(const RlsP& c, Real* __restrict__ theta, size_t D, size_t d) {
#pragma unroll 1
  for (int k = 0; k < (M + 4) * M; ++k) {
    const unsigned long long z = k < 64 ? c.zero_mask[0] : (k < 128 ? c.zero_mask[1] : c.zero_mask[2]);
    const unsigned long long o = k < 64 ? c.one_mask[0] : (k < 128 ? c.one_mask[1] : c.one_mask[2]);
    if ((z >> (k & 63)) & 1ull) theta[(size_t)k * D + d] = Real(0);
    else if ((o >> (k & 63)) & 1ull) theta[(size_t)k * D + d] = Real(1);
  }
}

// One RLS step for drone d.  phi = [e_t (m), u_t (4)], x1 = e_{t+1} (m).
// Loops over matrix ROWS are run-time loops (a fully unrolled sweep lets the scheduler hoist all (m+4)^2 loads at once:
// 255 registers and spills); the vectors a row loop indexes (phi, w = P phi, the Taylor term) therefore sit in shared
// memory as [index][thread] columns, and everything indexed by the unrolled column loop stays in registers.
#define MDS_RLS_THREADS 128
template <typename Real, int M>
__global__ void __launch_bounds__(MDS_RLS_THREADS) rls_update_kernel(RlsP c, const Real* __restrict__ phi_in, const Real* __restrict__ x1_in,
                                                                     Real* __restrict__ theta, Real* __restrict__ Pm, Real* __restrict__ resid_out, int D_) {
  constexpr int MN = M + 4;
  __shared__ Real s_phi[MN][MDS_RLS_THREADS], s_w[MN][MDS_RLS_THREADS], s_term[M][MDS_RLS_THREADS];
  const int tid = threadIdx.x;
  const size_t d = (size_t)blockIdx.x * blockDim.x + tid, D = (size_t)D_;
  if (d >= D) return;  // no block-level synchronisation below: every thread only touches its own shared column
  Real x1[M];
#pragma unroll
  for (int i = 0; i < MN; ++i) s_phi[i][tid] = phi_in[d * MN + i];
#pragma unroll
  for (int i = 0; i < M; ++i) x1[i] = x1_in[d * M + i];
  // robots after the first of an env see the projection of the earlier robots' loop iterations before their own update
  if (c.project == MDS_RLS_PROJECT_LOOP && (d % (size_t)c.drones_per_env) != 0) rls_project<Real, M>(c, theta, D, d);
  // ---- gain: w = P phi, v = phi' P, s = 1 + phi' P phi
  Real v[MN], phi[MN];
#pragma unroll
  for (int j = 0; j < MN; ++j) { v[j] = Real(0); phi[j] = s_phi[j][tid]; }
  Real s = Real(1);
#pragma unroll 1
  for (int i = 0; i < MN; ++i) {
    const Real phi_i = s_phi[i][tid];
    Real wi = Real(0);
#pragma unroll
    for (int j = 0; j < MN; ++j) {
      const Real p = Pm[(size_t)(i * MN + j) * D + d];
      wi += p * phi[j];
      v[j] += phi_i * p;
    }
    s_w[i][tid] = wi;
    s += phi_i * wi;
  }
  const Real inv_s = Real(1) / s;
  // ---- regression residual r (m)
  Real r[M];
  if (c.target == MDS_RLS_TARGET_XDOT) {
    // est_x_dot, then x_dot - theta' phi
    const Real inv_dt = Real(1.0 / c.dt);
    if (M == 12) {  // decentralized_lqr.py:185-198
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        r[k] = x1[3 + k];
        r[3 + k] = (x1[3 + k] - phi[3 + k]) * inv_dt;
        r[6 + k] = (x1[6 + k] - phi[6 + k]) * inv_dt;
        r[9 + k] = x1[6 + k];
      }
    } else {  // decentralized_yolqr_crazyflie.py:228-243 (m = 10; m = 9 is rejected on the host)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        r[k] = phi[M + 1 + k];
        r[4 + k] = (x1[4 + k] - phi[4 + k]) * inv_dt;
        r[7 + k] = x1[4 + k];
      }
      r[3] = phi[M];
    }
#pragma unroll 1
    for (int i = 0; i < MN; ++i) {
      const Real phi_i = s_phi[i][tid];
#pragma unroll
      for (int j = 0; j < M; ++j) r[j] -= theta[(size_t)(i * M + j) * D + d] * phi_i;
    }
  } else {
    // forward_predict: e' = Ahat e + Bhat u over dt from e0, Ahat = theta[:m]^T, Bhat = theta[m:]^T.  The reference runs
    // scipy RK45 (its error over one 1/240 s step is far below its 1e-3 tolerance); here the exact solution
    // e(dt) = e0 + sum_{k>=1} dt^k / k! A^(k-1) (A e0 + B u), summed until the terms vanish in Real.
    Real acc[M], nt[M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      acc[j] = c.predict_from_xtp1 ? x1[j] : phi[j];
      s_term[j][tid] = acc[j];
      nt[j] = Real(0);
    }
#pragma unroll 1
    for (int i = 0; i < MN; ++i) {  // A e0 + B u = theta' [e0; u]
      const Real zi = i < M ? s_term[i][tid] : s_phi[i][tid];
#pragma unroll
      for (int j = 0; j < M; ++j) nt[j] += theta[(size_t)(i * M + j) * D + d] * zi;
    }
    Real coef = Real(c.dt);
    const Real eps = sizeof(Real) == 4 ? Real(1e-9) : Real(1e-18);
#pragma unroll 1
    for (int k = 1; k <= 16; ++k) {
      Real big = Real(0), mag = Real(0);
#pragma unroll
      for (int j = 0; j < M; ++j) {
        const Real inc = coef * nt[j];
        acc[j] += inc;
        big = max_(big, abs_(inc)); mag = max_(mag, abs_(acc[j]));
        s_term[j][tid] = nt[j]; nt[j] = Real(0);
      }
      if (big <= eps * mag) break;  // the series has converged in Real (|A| dt ~ 0.05: 5-7 terms)
      coef *= Real(c.dt) / Real(k + 1);
#pragma unroll 1
      for (int i = 0; i < M; ++i) {
        const Real ti = s_term[i][tid];
#pragma unroll
        for (int j = 0; j < M; ++j) nt[j] += theta[(size_t)(i * M + j) * D + d] * ti;
      }
    }
#pragma unroll
    for (int j = 0; j < M; ++j) r[j] = x1[j] - acc[j];
  }
  if (resid_out) {
#pragma unroll
    for (int j = 0; j < M; ++j) resid_out[d * M + j] = r[j];
  }
  // ---- theta += L r',  P -= (w / s) v
#pragma unroll 1
  for (int i = 0; i < MN; ++i) {
    const Real wi = s_w[i][tid];
    const Real Li = c.normalize_gain ? wi * inv_s : wi;
#pragma unroll
    for (int j = 0; j < M; ++j) theta[(size_t)(i * M + j) * D + d] += Li * r[j];
  }
  if (c.project != MDS_RLS_PROJECT_NONE) rls_project<Real, M>(c, theta, D, d);
#pragma unroll 1
  for (int i = 0; i < MN; ++i) {
    const Real Li = s_w[i][tid] * inv_s;
#pragma unroll
    for (int j = 0; j < MN; ++j) Pm[(size_t)(i * MN + j) * D + d] -= Li * v[j];
  }
}

}  // namespace mds
