// mds_cbf.cuh -- exponential-CBF rows in closed form + per-env dual active-set QP.
// Replaces CBF.custom_hdots / custom_control_affine_terms / _build_ineq_const
// (cbf/cbf.py:135-283,308-464) and QPTracker._rectify (cbf/qptracker.py:86-114).
//
// The reference builds dense 2*xdim Jacobians/Hessians and a (2 xdim)^3 tensor per pair;
// with the hover-linearised A, B only e = p_i - p_j, the relative velocity error dv and
// the relative linearised acceleration da survive (derivation: DESIGN.md "CBF rows"):
//   d  = dh/de = (4 ex rho, 4 ey rho, 4 ez^3/c^4),  rho = ex^2 + ey^2,  H = d2h/de2
//   order 2:  Lf2h = d.da + dv'H dv            LgLfh  = (dz/m, 0, 0)
//   order 3:  Lf3h = 3 da'H dv + q(dv).dv      LgLf2h = (dz/m, -g dy, g dx)
// and each pair row is  -a on drone i's input block, +a on drone j's (cbf.py:299-300).
// The QP  min 1/2|u - u_nom|^2  s.t.  G u <= h  has P = I, so the dual active-set method
// (Goldfarb-Idnani) needs only sparse dot products between rows; the 4th input (wz) never
// appears in a barrier row and is a closed-form clamp.
#pragma once
#include "mds_common.cuh"
#include "mds_ctrl.cuh"

namespace mds {

#define MDS_QP_QMAX 24  // max simultaneously active constraints handled (else status ITER_CAP)

// what a barrier row needs from one agent: position, velocity error, linearised acceleration error
template <typename Real> struct CbfAgent {
  V3<Real> p, dv, da;
};

// obs + xdes (linear-model layout) -> CbfAgent.  order 2: xdes = [0,0,yaw,vel,pos];
// order 3: xdes = [0,0,yaw,F_des,vel,pos].  Also returns the current thrust F (order 3 force rows).
template <typename Real>
MDS_DEV CbfAgent<Real> cbf_agent(const DroneP<Real>& P, const CbfP<Real>& C, const Obs<Real>& o, const Real* xd, Real* F_out) {
  CbfAgent<Real> a;
  a.p = o.p;
  Real roll = o.rpy.x - xd[0], pitch = o.rpy.y - xd[1];
  if (C.order == 2) {
    a.dv = {o.v.x - xd[3], o.v.y - xd[4], o.v.z - xd[5]};
    a.da = {P.g * pitch, -P.g * roll, Real(0)};
    *F_out = Real(0);
  } else {
    Real F = z_thrust(P, o.rpm);
    a.dv = {o.v.x - xd[4], o.v.y - xd[5], o.v.z - xd[6]};
    a.da = {P.g * pitch, -P.g * roll, (F - xd[3]) / P.m};
    *F_out = F;
  }
  return a;
}

// One ECBF row between agent i and agent/obstacle j (obstacle: dv = da = 0).  a3 = LgLf^{r-1}h on
// i's block (columns u0, wx, wy); rhs = Kcbf.[h, hdot, (hddot)] + Lf^r h; h0 = barrier value.
template <typename Real>
MDS_DEV void cbf_row(const DroneP<Real>& P, const CbfP<Real>& C, const CbfAgent<Real>& ai, const CbfAgent<Real>& aj,
                     Real Ds, Real a3[3], Real* rhs, Real* h0_out) {
  V3<Real> e = ai.p - aj.p, dv = ai.dv - aj.dv, da = ai.da - aj.da;
  Real ex2 = e.x * e.x, ey2 = e.y * e.y, ez2 = e.z * e.z;
  Real rho = ex2 + ey2;
  V3<Real> d = {Real(4) * e.x * rho, Real(4) * e.y * rho, Real(4) * e.z * ez2 * C.c4inv};
  Real Hxx = Real(12) * ex2 + Real(4) * ey2, Hxy = Real(8) * e.x * e.y, Hyy = Real(4) * ex2 + Real(12) * ey2;
  Real Hzz = Real(12) * ez2 * C.c4inv;
  V3<Real> Hdv = {Hxx * dv.x + Hxy * dv.y, Hxy * dv.x + Hyy * dv.y, Hzz * dv.z};
  Real Ds2 = Ds * Ds;
  Real h0 = rho * rho + ez2 * ez2 * C.c4inv - Ds2 * Ds2;
  Real h1 = dot(d, dv);
  Real inv_m = Real(1) / P.m;
  *h0_out = h0;
  if (C.order == 2) {
    Real Lf = dot(d, da) + dot(dv, Hdv);
    a3[0] = d.z * inv_m; a3[1] = Real(0); a3[2] = Real(0);
    *rhs = C.k0 * h0 + C.k1 * h1 + Lf;
    return;
  }
  // hdots[2] with the reference's hard-coded indices 6,7,8 of the 10-dim state (quirk B12)
  Real h2 = d.y * da.x + d.z * da.y + Hxx * da.z * da.z + Real(2) * Hxy * da.z * dv.x + Hyy * dv.x * dv.x + Hzz * dv.y * dv.y;
  V3<Real> q = {Real(24) * e.x * dv.x * dv.x + Real(16) * e.y * dv.x * dv.y + Real(8) * e.x * dv.y * dv.y,
                Real(8) * e.y * dv.x * dv.x + Real(16) * e.x * dv.x * dv.y + Real(24) * e.y * dv.y * dv.y,
                Real(24) * e.z * C.c4inv * dv.z * dv.z};
  Real Lf = Real(3) * dot(da, Hdv) + dot(q, dv);
  a3[0] = d.z * inv_m; a3[1] = -P.g * d.y; a3[2] = P.g * d.x;
  *rhs = C.k0 * h0 + C.k1 * h1 + C.k2 * h2 + Lf;
}

// pair index r in [0, N(N-1)/2) -> (i, j), i < j, lexicographic (cbf.py:342-346)
MDS_DEV void pair_from_index(int r, int N, int* i, int* j) {
  int a = 0;
  while (r >= N - 1 - a) { r -= N - 1 - a; ++a; }
  *i = a; *j = a + 1 + r;
}

// ----------------------------------------------------------------------------------------
// Per-env QP over the coupled inputs x[4n + c], c in {0,1,2}, n < N (c == 3 is decoupled).
// Constraint index space:  [0, n_rows): barrier rows (pairs, then obstacles i*n_obs + o)
//                          [n_rows, n_rows + 6N): box  s * x[i,c] <= umax[c], k = i*6 + (s<0)*3 + c
// rows[r*4 + 0..2] = a3, rows[r*4 + 3] = rhs.  All arrays may live in shared memory.
template <typename Real> struct QpCon {
  int i, j;      // drone blocks (j < 0: single block)
  Real gi[3];    // coefficients on block i; block j carries -gi (pair rows)
  Real rhs;
};

template <typename Real>
MDS_DEV QpCon<Real> qp_get(const Real* rows, const CbfP<Real>& C, int idx, int N, int n_pairs, int n_rows, int n_obs) {
  QpCon<Real> c;
  if (idx < n_rows) {
    const Real* r = rows + 4 * idx;
    c.gi[0] = -r[0]; c.gi[1] = -r[1]; c.gi[2] = -r[2]; c.rhs = r[3];
    if (idx < n_pairs) pair_from_index(idx, N, &c.i, &c.j);
    else { c.i = (idx - n_pairs) / n_obs; c.j = -1; }
  } else {
    int k = idx - n_rows;
    c.i = k / 6; c.j = -1;
    int rem = k - 6 * c.i;
    int comp = rem % 3;
    Real s = rem < 3 ? Real(1) : Real(-1);
    c.gi[0] = comp == 0 ? s : Real(0); c.gi[1] = comp == 1 ? s : Real(0); c.gi[2] = comp == 2 ? s : Real(0);
    c.rhs = C.umax[comp];
  }
  return c;
}
template <typename Real> MDS_DEV Real qp_dot_x(const QpCon<Real>& c, const Real* x, Real* mag) {
  const Real* xi = x + 4 * c.i;
  Real t0 = c.gi[0] * xi[0], t1 = c.gi[1] * xi[1], t2 = c.gi[2] * xi[2];
  Real s = t0 + t1 + t2, m = abs_(t0) + abs_(t1) + abs_(t2);
  if (c.j >= 0) {
    const Real* xj = x + 4 * c.j;
    Real u0 = c.gi[0] * xj[0], u1 = c.gi[1] * xj[1], u2 = c.gi[2] * xj[2];
    s -= u0 + u1 + u2;
    m += abs_(u0) + abs_(u1) + abs_(u2);
  }
  *mag = m;
  return s;
}
template <typename Real> MDS_DEV Real qp_dot_g(const QpCon<Real>& a, const QpCon<Real>& b) {
  Real ab = a.gi[0] * b.gi[0] + a.gi[1] * b.gi[1] + a.gi[2] * b.gi[2];
  Real s = Real(0);
  if (a.i == b.i) s += ab;
  if (a.j >= 0 && a.j == b.j) s += ab;
  if (a.j >= 0 && a.j == b.i) s -= ab;
  if (b.j >= 0 && b.j == a.i) s -= ab;
  return s;
}

// Goldfarb-Idnani dual active set, P = I.  x holds u_nom on entry, the minimiser on exit.
// z: work vector of 4N Reals.  Returns MDS_QP_*; *iters_out = inner iterations.
template <typename Real>
MDS_DEV int qp_solve(const CbfP<Real>& C, const Real* rows, Real* x, Real* z, int N, int n_pairs, int n_rows, int n_obs, int* iters_out) {
  const Real tol = sizeof(Real) == 4 ? Real(2e-6) : Real(1e-11);
  const Real INF = Real(1e30);
  int act[MDS_QP_QMAX];
  Real lam[MDS_QP_QMAX], r[MDS_QP_QMAX];
  Real Lc[MDS_QP_QMAX * (MDS_QP_QMAX + 1) / 2];
  int q = 0, iters = 0;
  const int n_con = n_rows + 6 * N;
  for (;;) {
    // ---- most violated inactive constraint (normalised by |g|)
    int p = -1;
    Real worst = Real(0);
    for (int idx = 0; idx < n_con; ++idx) {
      bool is_act = false;
      for (int k = 0; k < q; ++k) is_act |= (act[k] == idx);
      if (is_act) continue;
      QpCon<Real> c = qp_get(rows, C, idx, N, n_pairs, n_rows, n_obs);
      Real mag, gx = qp_dot_x(c, x, &mag);
      Real s = c.rhs - gx;
      if (s < -tol * (abs_(c.rhs) + mag + Real(1e-12))) {
        Real g2 = qp_dot_g(c, c);
        if (g2 <= Real(0)) { *iters_out = iters; return MDS_QP_INFEASIBLE; }  // 0 <= rhs < 0
        Real v = s * rsqrt_(g2);
        if (v < worst) { worst = v; p = idx; }
      }
    }
    if (p < 0) { *iters_out = iters; return MDS_QP_OPTIMAL; }
    QpCon<Real> cp = qp_get(rows, C, p, N, n_pairs, n_rows, n_obs);
    Real g2p = qp_dot_g(cp, cp);
    Real lam_p = Real(0);
    for (;;) {
      if (++iters > C.max_iter) { *iters_out = iters; return MDS_QP_ITER_CAP; }
      // ---- r = (Na Na')^-1 Na g_p by Cholesky of the active Gram matrix
      for (int a = 0; a < q; ++a) {
        QpCon<Real> ca = qp_get(rows, C, act[a], N, n_pairs, n_rows, n_obs);
        for (int b = 0; b <= a; ++b) {
          QpCon<Real> cb = qp_get(rows, C, act[b], N, n_pairs, n_rows, n_obs);
          Real s = qp_dot_g(ca, cb);
          for (int k = 0; k < b; ++k) s -= Lc[a * (a + 1) / 2 + k] * Lc[b * (b + 1) / 2 + k];
          if (a == b) {
            if (s <= Real(0)) { *iters_out = iters; return MDS_QP_ITER_CAP; }
            Lc[a * (a + 1) / 2 + a] = sqrt_(s);
          } else {
            Lc[a * (a + 1) / 2 + b] = s / Lc[b * (b + 1) / 2 + b];
          }
        }
        r[a] = qp_dot_g(ca, cp);
      }
      for (int a = 0; a < q; ++a) {  // forward
        Real s = r[a];
        for (int k = 0; k < a; ++k) s -= Lc[a * (a + 1) / 2 + k] * r[k];
        r[a] = s / Lc[a * (a + 1) / 2 + a];
      }
      for (int a = q - 1; a >= 0; --a) {  // backward
        Real s = r[a];
        for (int k = a + 1; k < q; ++k) s -= Lc[k * (k + 1) / 2 + a] * r[k];
        r[a] = s / Lc[a * (a + 1) / 2 + a];
      }
      // ---- z = g_p - Na' r  (dense over the 3 coupled inputs of every drone)
      for (int k = 0; k < 4 * N; ++k) z[k] = Real(0);
      for (int c = 0; c < 3; ++c) {
        z[4 * cp.i + c] += cp.gi[c];
        if (cp.j >= 0) z[4 * cp.j + c] -= cp.gi[c];
      }
      for (int a = 0; a < q; ++a) {
        QpCon<Real> ca = qp_get(rows, C, act[a], N, n_pairs, n_rows, n_obs);
        for (int c = 0; c < 3; ++c) {
          z[4 * ca.i + c] -= r[a] * ca.gi[c];
          if (ca.j >= 0) z[4 * ca.j + c] += r[a] * ca.gi[c];
        }
      }
      Real mag, zn = qp_dot_x(cp, z, &mag);
      // ---- step lengths
      Real t1 = INF;
      int kdrop = -1;
      for (int a = 0; a < q; ++a)
        if (r[a] > tol) {
          Real cnd = lam[a] / r[a];
          if (cnd < t1) { t1 = cnd; kdrop = a; }
        }
      Real gx = qp_dot_x(cp, x, &mag);
      Real s_p = cp.rhs - gx;
      Real t2 = (zn > Real(sizeof(Real) == 4 ? 1e-5 : 1e-10) * g2p) ? -s_p / zn : INF;
      Real t = min_(t1, t2);
      if (t >= INF) { *iters_out = iters; return MDS_QP_INFEASIBLE; }
      if (t2 < INF)
        for (int k = 0; k < 4 * N; ++k) x[k] -= t * z[k];
      for (int a = 0; a < q; ++a) lam[a] -= t * r[a];
      lam_p += t;
      if (t2 <= t1) {
        if (q == MDS_QP_QMAX) { *iters_out = iters; return MDS_QP_ITER_CAP; }
        act[q] = p; lam[q] = lam_p; ++q;
        break;
      }
      for (int a = kdrop; a < q - 1; ++a) { act[a] = act[a + 1]; lam[a] = lam[a + 1]; }
      --q;
    }
  }
}

// decoupled 4th input (wz): box +-umax[3] merged with the order-3 force-bound rows, which the
// reference places on column 4i+3 (cbf.py:456-460).  Returns false when the interval is empty.
template <typename Real> MDS_DEV bool cbf_wz_bounds(const CbfP<Real>& C, Real F, Real* lo, Real* hi) {
  *lo = -C.umax[3]; *hi = C.umax[3];
  if (C.order == 3) {
    *hi = min_(*hi, C.k2 * (C.fmax - F));
    *lo = max_(*lo, -(C.k2 * (F - C.fmin)));
  }
  return *lo <= *hi;
}

MDS_DEV int cbf_num_barrier_rows(int N, int n_obs) { return N * (N - 1) / 2 + N * n_obs; }

}  // namespace mds
