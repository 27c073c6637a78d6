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

#define MDS_QP_QMAX 12  // active constraints the in-shared-memory solver holds; larger sets go to the scratch solver (qp_solve_group<.., true>)

// what a barrier row needs from one agent: position, velocity error, linearised acceleration error
template <typename Real> struct CbfAgent {
  V3<Real> p, dv, da;
};

// ORD: relative degree as a compile-time constant (2 or 3), or 0 = read C.order at run time
template <int ORD, typename Real> MDS_DEV int cbf_order(const CbfP<Real>& C) { return ORD ? ORD : C.order; }

// obs + xdes (linear-model layout) -> CbfAgent.  order 2: xdes = [0,0,yaw,vel,pos];
// order 3: xdes = [0,0,yaw,F_des,vel,pos].  Also returns the current thrust F (order 3 force rows).
template <int ORD, typename Real>
MDS_DEV CbfAgent<Real> cbf_agent(const DroneP<Real>& P, const CbfP<Real>& C, const Obs<Real>& o, const Real* xd, Real* F_out) {
  CbfAgent<Real> a;
  a.p = o.p;
  Real roll = o.rpy.x - xd[0], pitch = o.rpy.y - xd[1];
  if (cbf_order<ORD>(C) == 2) {
    a.dv = {o.v.x - xd[3], o.v.y - xd[4], o.v.z - xd[5]};
    a.da = {P.g * pitch, -P.g * roll, Real(0)};
    *F_out = Real(0);
  } else {
    Real F = z_thrust(P, o.rpm);
    a.dv = {o.v.x - xd[4], o.v.y - xd[5], o.v.z - xd[6]};
    a.da = {P.g * pitch, -P.g * roll, (F - xd[3]) * P.inv_m};
    *F_out = F;
  }
  return a;
}

// One ECBF row between agent i and agent/obstacle j (obstacle: dv = da = 0).  a3 = LgLf^{r-1}h on
// i's block (columns u0, wx, wy); rhs = Kcbf.[h, hdot, (hddot)] + Lf^r h; h0 = barrier value.  Ds4 = Ds^4;
// c4inv = 1 / zscale^4 for agents and sphere obstacles; 0 for a vertical-cylinder obstacle (the z terms vanish).
// The closed forms of the header are grouped around  rho = ex^2 + ey^2,  s = e_xy . dv_xy,  w = |dv_xy|^2:
//   H dv = (4 rho dvx + 8 ex s, 4 rho dvy + 8 ey s, Hzz dvz),  dv'H dv = 4 rho w + 8 s^2 + Hzz dvz^2,
//   q(dv).dv = 24 (s w + ez dvz^3 / c^4)
// so that a row is ~60 instructions, most of them FMAs (products-then-sums cost 95).
// The row as a function of the RELATIVE state (e, dv, da) = agent i - agent j.  T = Real evaluates one row; T = F2
// (mds_common.cuh) evaluates two rows of the same owner at once with Blackwell's packed fp32 instructions: the parameters
// (Ds4, c4inv, gains, 1/m, g) are scalars and broadcast inside the instruction.
template <int ORD, typename Real, typename T>
MDS_DEV void cbf_row_rel(const DroneP<Real>& P, const CbfP<Real>& C, T ex, T ey, T ez, T dvx, T dvy, T dvz, T dax, T day, T daz,
                         Real Ds4, Real c4inv, T a3[3], T* rhs, T* h0_out) {
  const T rho = fma_(ey, ey, ex * ex), ez2 = ez * ez, ez2c = ez2 * T(c4inv), rho4 = T(Real(4)) * rho;
  const T s = fma_(ey, dvy, ex * dvx);
  const T dvx2 = dvx * dvx, dvy2 = dvy * dvy;
  const T dz = T(Real(4)) * (ez * ez2c), dx = rho4 * ex, dy = rho4 * ey;  // d = dh/de
  const T h0 = fma_(rho, rho, fma_(ez2, ez2c, T(-Ds4)));
  const T h1 = fma_(rho4, s, dz * dvz);
  const T Hzz = T(Real(12)) * ez2c;
  *h0_out = h0;
  if (cbf_order<ORD>(C) == 2) {
    const T Lf = fma_(dx, dax, dy * day) + fma_(rho4, dvx2 + dvy2, fma_(T(Real(8)) * s, s, Hzz * dvz * dvz));
    a3[0] = dz * T(P.inv_m); a3[1] = T(Real(0)); a3[2] = T(Real(0));
    *rhs = fma_(T(C.k0), h0, fma_(T(C.k1), h1, Lf));
    return;
  }
  const T ex8 = T(Real(8)) * ex;
  const T Hxx = fma_(ex8, ex, rho4), Hyy = fma_(T(Real(8)) * ey, ey, rho4), Hxy = ex8 * ey;
  // hdots[2] with the reference's hard-coded indices 6,7,8 of the 10-dim state (quirk B12)
  const T h2 = fma_(dy, dax, fma_(dz, day, fma_(daz, fma_(Hxx, daz, T(Real(2)) * Hxy * dvx), fma_(Hyy, dvx2, Hzz * dvy2))));
  const T daHdv = fma_(rho4, fma_(dax, dvx, day * dvy), fma_(T(Real(8)) * s, fma_(ex, dax, ey * day), Hzz * dvz * daz));
  const T qdv = T(Real(24)) * fma_(s, dvx2 + dvy2, (T(c4inv) * ez) * (dvz * dvz * dvz));
  const T Lf = fma_(T(Real(3)), daHdv, qdv);
  a3[0] = dz * T(P.inv_m); a3[1] = T(-P.g) * dy; a3[2] = T(P.g) * dx;
  *rhs = fma_(T(C.k0), h0, fma_(T(C.k1), h1, fma_(T(C.k2), h2, Lf)));
}
template <int ORD, typename Real>
MDS_DEV void cbf_row(const DroneP<Real>& P, const CbfP<Real>& C, const CbfAgent<Real>& ai, const CbfAgent<Real>& aj,
                     Real Ds4, Real c4inv, Real a3[3], Real* rhs, Real* h0_out) {
  cbf_row_rel<ORD, Real, Real>(P, C, ai.p.x - aj.p.x, ai.p.y - aj.p.y, ai.p.z - aj.p.z, ai.dv.x - aj.dv.x, ai.dv.y - aj.dv.y,
                               ai.dv.z - aj.dv.z, ai.da.x - aj.da.x, ai.da.y - aj.da.y, ai.da.z - aj.da.z, Ds4, c4inv, a3, rhs, h0_out);
}
// Two rows of one owner at once: (ai - aj0) in the low halves, (ai - aj1) in the high halves.  The nine differences are scalar
// subtractions written straight into the halves (no packing moves); everything after them is packed.
template <int ORD>
MDS_DEV void cbf_row2(const DroneP<float>& P, const CbfP<float>& C, const CbfAgent<float>& ai, const CbfAgent<float>& aj0,
                      const CbfAgent<float>& aj1, float Ds4, float c4inv, F2 a3[3], F2* rhs, F2* h0_out) {
  cbf_row_rel<ORD, float, F2>(P, C, F2(ai.p.x - aj0.p.x, ai.p.x - aj1.p.x), F2(ai.p.y - aj0.p.y, ai.p.y - aj1.p.y),
                              F2(ai.p.z - aj0.p.z, ai.p.z - aj1.p.z), F2(ai.dv.x - aj0.dv.x, ai.dv.x - aj1.dv.x),
                              F2(ai.dv.y - aj0.dv.y, ai.dv.y - aj1.dv.y), F2(ai.dv.z - aj0.dv.z, ai.dv.z - aj1.dv.z),
                              F2(ai.da.x - aj0.da.x, ai.da.x - aj1.da.x), F2(ai.da.y - aj0.da.y, ai.da.y - aj1.da.y),
                              F2(ai.da.z - aj0.da.z, ai.da.z - aj1.da.z), Ds4, c4inv, a3, rhs, h0_out);
}

// ----------------------------------------------------------------------------------------
// Row ownership.  The env's lane group (NP = next power of two >= N consecutive lanes of one warp, lane n
// owns drone n) splits the barrier rows so that every row has its owner's drone as one end point:
//   slot s <  K1 = (N-1)/2          : pair (n, n+s+1 mod N)          -- every lane
//   slot s == K1, N even            : pair (n, n+N/2) for n < N/2    -- the "diameters"
//   slot s >= S0 = K1 + (N even)    : obstacle s - S0 against drone n
// RPL = S0 + n_obs slots per lane; row id r = n * RPL + s.  Each unordered pair appears exactly once, evaluated
// as (owner - partner): a row is odd in (e, dv, da) in its coefficients and even in its right-hand side, and
// negation is exact in floating point, so  G = -a on the owner's block, +a on the partner's  is bit-identical
// to the reference's i < j orientation (cbf.py:342-346, 299-300).  Unused slots hold a never-violated row.
struct RowMap {
  int K1, S0, RPL, half;  // half = N/2 if N even else 0
};
MDS_DEV RowMap row_map(int N, int n_obs) {
  RowMap m;
  m.K1 = (N - 1) >> 1;
  m.half = (N & 1) ? 0 : (N >> 1);
  m.S0 = m.K1 + (m.half ? 1 : 0);
  m.RPL = m.S0 + n_obs;
  return m;
}
// partner drone of (lane n, slot s): >= 0 pair partner, -1 obstacle row, -2 unused slot
// m in [-N, 2N) wrapped into [0, N): a mask when N is a power of two (N folds to a constant in the NT kernels)
MDS_DEV int wrap_n(int m, int N) {
  if ((N & (N - 1)) == 0) return m & (N - 1);
  return m >= N ? m - N : (m < 0 ? m + N : m);
}
MDS_DEV int row_partner(const RowMap& M, int N, int n, int s) {
  if (s < M.K1) return wrap_n(n + s + 1, N);
  if (s < M.S0) return (n < M.half) ? n + M.half : -2;
  return -1;
}

// ----------------------------------------------------------------------------------------
// Per-env QP over the coupled inputs x[4n + c], c in {0,1,2}, n < N (c == 3 is decoupled), solved
// COOPERATIVELY by the env's lane group.  Only group-level synchronisation is used (__syncwarp /
// shuffles with the group's lane mask), so the groups of a warp iterate independently; no block barrier.
//
// Constraint ids:  [0, NP*RPL)          barrier rows (above);  smem record = (a0, a1, a2, rhs), G = -a on i, +a on j
//                  MDS_QP_BOX0 + 6n + k  box  s * x[n,c] <= umax[c],  k = (s<0)*3 + c
// An active-set entry packs (id | i << 16 | j << 24), j = 0x7f for single-block constraints.
#define MDS_QP_BOX0 4096
#define MDS_QP_WS_WORDS (4 * MDS_QP_QMAX + MDS_QP_QMAX * (MDS_QP_QMAX + 1) / 2 + 4)  // act, lam, d, r, chol, header

// umax[comp] for a run-time comp in 0..2 by selects: a dynamically indexed member would make nvcc copy the whole
// (kernel-parameter) CbfP block to local memory and turn every C.* read of the kernel into an LDL.
template <typename Real> MDS_DEV Real cbf_umax(const CbfP<Real>& C, int comp) { return comp == 0 ? C.umax[0] : (comp == 1 ? C.umax[1] : C.umax[2]); }

template <typename Real> struct QpCon {
  int i, j;      // drone blocks (j < 0: single block)
  Real gi[3];    // coefficients on block i; block j carries -gi (pair rows)
  Real rhs, g2;
};
MDS_DEV int pack_con(int id, int i, int j) { return id | (i << 16) | ((j < 0 ? 0x7f : j) << 24); }

template <typename Real>
MDS_DEV QpCon<Real> qp_get(const typename Vec4T<Real>::type* rows, const CbfP<Real>& C, int packed) {
  QpCon<Real> c;
  const int id = packed & 0xffff;
  c.i = (packed >> 16) & 0xff;
  c.j = (packed >> 24) & 0x7f;
  if (c.j == 0x7f) c.j = -1;
  if (id < MDS_QP_BOX0) {
    auto r = rows[id];
    c.gi[0] = -r.x; c.gi[1] = -r.y; c.gi[2] = -r.z; c.rhs = r.w;
    Real a2 = r.x * r.x + r.y * r.y + r.z * r.z;
    c.g2 = c.j >= 0 ? Real(2) * a2 : a2;
  } else {
    int rem = id - MDS_QP_BOX0 - 6 * c.i;
    int comp = rem >= 3 ? rem - 3 : rem;
    Real s = rem < 3 ? Real(1) : Real(-1);
    c.gi[0] = comp == 0 ? s : Real(0); c.gi[1] = comp == 1 ? s : Real(0); c.gi[2] = comp == 2 ? s : Real(0);
    c.rhs = cbf_umax(C, comp); c.g2 = Real(1);
  }
  return c;
}
template <typename Real> MDS_DEV Real qp_dot_x(const QpCon<Real>& c, const typename Vec4T<Real>::type* x, Real* mag) {
  auto xi = x[c.i];
  Real t0 = c.gi[0] * xi.x, t1 = c.gi[1] * xi.y, t2 = c.gi[2] * xi.z;
  Real s = t0 + t1 + t2, m = abs_(t0) + abs_(t1) + abs_(t2);
  if (c.j >= 0) {
    auto xj = x[c.j];
    Real u0 = c.gi[0] * xj.x, u1 = c.gi[1] * xj.y, u2 = c.gi[2] * xj.z;
    s -= u0 + u1 + u2;
    m += abs_(u0) + abs_(u1) + abs_(u2);
  }
  *mag = m;
  return s;
}
// g_a . g_b accumulated in W (the scratch solver and the polish work in double whatever the row precision)
template <typename W, typename Real> MDS_DEV W qp_dot_g(const QpCon<Real>& a, const QpCon<Real>& b) {
  W ab = (W)a.gi[0] * (W)b.gi[0] + (W)a.gi[1] * (W)b.gi[1] + (W)a.gi[2] * (W)b.gi[2];
  W s = W(0);
  if (a.i == b.i) s += ab;
  if (a.j >= 0 && a.j == b.j) s += ab;
  if (a.j >= 0 && a.j == b.i) s -= ab;
  if (b.j >= 0 && b.j == a.i) s -= ab;
  return s;
}
// coefficient of constraint c on drone block n, component k
template <typename Real> MDS_DEV Real qp_coef(const QpCon<Real>& c, int n, int k) {
  return (c.i == n) ? c.gi[k] : ((c.j == n) ? -c.gi[k] : Real(0));
}

// slack test shared by the row builder and the scans: violated <=> rhs - G x < -tol (|rhs| + sum |terms|)
template <typename Real> MDS_DEV Real qp_tol() { return sizeof(Real) == 4 ? Real(2e-6) : Real(1e-11); }

// running "most violated constraint" of one lane: (normalised slack, packed constraint), ties -> lowest id
template <typename Real> struct QpWorst {
  Real v;
  int con;
};
// owner lane n tests its own barrier row (a0, a1, a2, rhs) with partner m (< 0: obstacle) at the iterate
// (own block xn, partner block from shared memory):  G x = -a.x_n + a.x_m
template <typename Real>
MDS_DEV void qp_test_row(QpWorst<Real>& w, const typename Vec4T<Real>::type& r, const Real xn[3], const typename Vec4T<Real>::type* x,
                         int n, int m, int id) {
  Real gx = -(r.x * xn[0] + r.y * xn[1] + r.z * xn[2]);
  typename Vec4T<Real>::type xm;
  xm.x = Real(0); xm.y = Real(0); xm.z = Real(0); xm.w = Real(0);
  if (m >= 0) {
    xm = x[m];
    gx += r.x * xm.x + r.y * xm.y + r.z * xm.z;
  }
  Real sl = r.w - gx;
  if (sl < Real(0)) {  // the magnitude of the terms (for the relative tolerance) only when the slack is negative
    Real mag = abs_(r.x * xn[0]) + abs_(r.y * xn[1]) + abs_(r.z * xn[2]) + abs_(r.x * xm.x) + abs_(r.y * xm.y) + abs_(r.z * xm.z);
    if (sl < -qp_tol<Real>() * (abs_(r.w) + mag + Real(1e-12))) {
      Real a2 = r.x * r.x + r.y * r.y + r.z * r.z;
      if (m >= 0) a2 *= Real(2);
      Real v = (a2 > Real(0)) ? sl * rsqrt_(a2) : Real(-1e30);  // zero row with rhs < 0: infeasible
      int con = pack_con(id, n, m);
      if (v < w.v || (v == w.v && con < w.con)) { w.v = v; w.con = con; }
    }
  }
}
// own box bounds not in boxmask: per component only the bound on the side x is on can be violated
template <typename Real> MDS_DEV void qp_test_box(QpWorst<Real>& w, const CbfP<Real>& C, const Real xn[3], int n, unsigned boxmask) {
#pragma unroll
  for (int comp = 0; comp < 3; ++comp) {
    const Real ax = abs_(xn[comp]);
    const Real sl = C.umax[comp] - ax;
    if (sl < -qp_tol<Real>() * (C.umax[comp] + ax + Real(1e-12))) {
      const int k = (xn[comp] < Real(0) ? 3 : 0) + comp;
      if (!(boxmask & (1u << k))) {
        int con = pack_con(MDS_QP_BOX0 + 6 * n + k, n, -1);
        if (sl < w.v || (sl == w.v && con < w.con)) { w.v = sl; w.con = con; }
      }
    }
  }
}
// group-wide argmin -> packed constraint, or -1 (none violated) / -2 (a zero row with negative rhs: infeasible)
template <typename Real> MDS_DEV int qp_worst_of_group(QpWorst<Real> w, int NP, unsigned gmask) {
  for (int off = NP >> 1; off > 0; off >>= 1) {
    Real ov = __shfl_xor_sync(gmask, w.v, off);
    int oc = __shfl_xor_sync(gmask, w.con, off);
    if (ov < w.v || (ov == w.v && oc < w.con)) { w.v = ov; w.con = oc; }
  }
  if (w.con == 0x7fffffff) return -1;
  if (w.v <= Real(-1e30)) return -2;
  return w.con;
}

// Most violated inactive constraint among every lane's own rows (slots not in rowmask) and own box bounds
// (not in boxmask) at the iterate x.
template <typename Real>
MDS_DEV int qp_scan(const CbfP<Real>& C, const typename Vec4T<Real>::type* rows, const typename Vec4T<Real>::type* x, const Real xn[3],
                    const RowMap& M, int N, int NP, int n, bool valid, unsigned rowmask, unsigned boxmask, unsigned gmask) {
  QpWorst<Real> w = {Real(0), 0x7fffffff};
  if (valid) {
#pragma unroll
    for (int s = 0; s < M.S0; ++s) {  // pair slots
      const int m = row_partner(M, N, n, s);
      if (m >= 0 && !(rowmask & (1u << s))) qp_test_row(w, rows[n * M.RPL + s], xn, x, n, m, n * M.RPL + s);
    }
    for (int s = M.S0; s < M.RPL; ++s)  // obstacle slots
      if (!(rowmask & (1u << s))) qp_test_row(w, rows[n * M.RPL + s], xn, x, n, -1, n * M.RPL + s);
    qp_test_box(w, C, xn, n, boxmask);
  }
  return qp_worst_of_group(w, NP, gmask);
}

// int stored in a workspace slot of either width (the active list shares the Real / double workspace)
template <typename W> MDS_DEV int& ws_int(W* slot) { return *reinterpret_cast<int*>(slot); }
template <typename W> MDS_DEV int ws_int(const W* slot) { return *reinterpret_cast<const int*>(slot); }

// fp32 polish, lane-0 part (solves that end with >= 3 active constraints): solve (A A') lam = A u_nom - b for the final
// active set in double and leave lam in the workspace.  act = packed active list (one int per W slot); Lm, y = double
// scratch of q (q + 1) / 2 and q entries.
template <typename Real, typename W>
MDS_DEV bool qp_polish_core(Real umax0, Real umax1, Real umax2, const typename Vec4T<Real>::type* rows, const typename Vec4T<Real>::type* xnom,
                            const W* act, double* Lm, double* y, W* lam, int q) {
  CbfP<Real> C;  // only the box bounds are read (qp_get); taking the caller's block by reference would force it into local memory
  C.umax[0] = umax0; C.umax[1] = umax1; C.umax[2] = umax2;
  bool ok = true;
  for (int a = 0; a < q && ok; ++a) {
    QpCon<Real> ca = qp_get(rows, C, ws_int(act + a));
    for (int b2 = 0; b2 <= a; ++b2) {
      QpCon<Real> cb = qp_get(rows, C, ws_int(act + b2));
      double sacc = qp_dot_g<double>(ca, cb);
      for (int k = 0; k < b2; ++k) sacc -= Lm[a * (a + 1) / 2 + k] * Lm[b2 * (b2 + 1) / 2 + k];
      if (a == b2) {
        if (sacc <= 0.0) { ok = false; break; }
        Lm[a * (a + 1) / 2 + a] = sqrt(sacc);
      } else {
        Lm[a * (a + 1) / 2 + b2] = sacc / Lm[b2 * (b2 + 1) / 2 + b2];
      }
    }
    auto ui = xnom[ca.i];
    double r = (double)ca.gi[0] * ui.x + (double)ca.gi[1] * ui.y + (double)ca.gi[2] * ui.z;
    if (ca.j >= 0) { auto uj = xnom[ca.j]; r -= (double)ca.gi[0] * uj.x + (double)ca.gi[1] * uj.y + (double)ca.gi[2] * uj.z; }
    y[a] = r - (double)ca.rhs;
  }
  if (ok) {
    for (int a = 0; a < q; ++a) {  // forward, then backward substitution
      double sacc = y[a];
      for (int k = 0; k < a; ++k) sacc -= Lm[a * (a + 1) / 2 + k] * y[k];
      y[a] = sacc / Lm[a * (a + 1) / 2 + a];
    }
    for (int a = q - 1; a >= 0; --a) {
      double sacc = y[a];
      for (int k = a + 1; k < q; ++k) sacc -= Lm[k * (k + 1) / 2 + a] * y[k];
      y[a] = sacc / Lm[a * (a + 1) / 2 + a];
    }
    for (int a = 0; a < q; ++a) lam[a] = (W)y[a];
  }
  return ok;
}
// in-shared-memory solver: the double arrays live in this function's own frame (out of line, off the step loop's registers)
template <typename Real>
__device__ __noinline__ bool qp_polish_small(Real umax0, Real umax1, Real umax2, const typename Vec4T<Real>::type* rows,
                                             const typename Vec4T<Real>::type* xnom, Real* ws, Real* lam, int q) {
  double Lm[MDS_QP_QMAX * (MDS_QP_QMAX + 1) / 2], y[MDS_QP_QMAX];
  return qp_polish_core<Real, Real>(umax0, umax1, umax2, rows, xnom, ws, Lm, y, lam, q);
}

// Goldfarb-Idnani dual active set, P = I, executed by the env's lane group.
//   rows : smem barrier rows;  x : smem iterate, one Vec4 per drone (u_nom on entry, minimiser on exit);
//   xnom : smem copy of u_nom that stays untouched (fp32 polish);
//   ws   : workspace -- BIG = false: MDS_QP_WS_WORDS Reals of shared memory, at most MDS_QP_QMAX active constraints;
//          BIG = true: a slot of the global scratch in double (QpScratch), at most qmax = 3 N active constraints (the number of
//          coupled variables, i.e. no cap at all);
//   p0   : the most violated constraint at u_nom (packed).
// Work split: scans and the primal update are spread over the lanes (lane n owns drone n's inputs, rows and
// box bounds).  BIG = false computes the first iteration (empty active set: z = g_p, t = -s_p / |g_p|^2) redundantly on
// every lane with no workspace traffic.  From then on the O(q^2) scalar part (triangular solves with the Cholesky factor
// of the active Gram matrix, step lengths, multiplier and factor updates) is done by the group's lane 0 and published
// through the workspace between __syncwarp(gmask) points.
// Returns MDS_QP_ITER_CAP when the workspace / iteration budget is exhausted or the factor breaks down; the caller then
// repeats the solve with BIG = true, whose own ITER_CAP is final.
enum { MDS_QP_ACT_FULL = 0, MDS_QP_ACT_DROP = 1, MDS_QP_ACT_STOP = 2 };

template <typename Real, typename W, bool BIG>
MDS_DEV int qp_solve_group(const CbfP<Real>& C, const typename Vec4T<Real>::type* rows, typename Vec4T<Real>::type* x,
                           const typename Vec4T<Real>::type* xnom, W* ws, int qmax_rt, const RowMap& M, int N, int NP, int n, bool valid,
                           unsigned gmask, int p0, int* iters_out) {
  using R4 = typename Vec4T<Real>::type;
  const int qmax = BIG ? qmax_rt : MDS_QP_QMAX;
  const int max_iter = BIG ? 16 * qmax + 64 : C.max_iter;
  const W tol = (W)qp_tol<Real>();
  const W zn_eps = sizeof(Real) == 4 ? W(1e-5) : W(1e-10);
  const W INF = W(1e30);
  // workspace: act | lam | d | r (qmax each) | Lc (packed lower-triangular Cholesky factor of the active Gram matrix) | hdr
  // (| Lm | y : double scratch of the polish, BIG only)
  W* lam = ws + qmax;
  W* dv = ws + 2 * qmax;
  W* rv = ws + 3 * qmax;
  W* Lc = ws + 4 * qmax;
  W* hdr = Lc + qmax * (qmax + 1) / 2;  // [0] t, [1] action, [2] dropped index, [3] status
  unsigned rowmask = 0, boxmask = 0;  // this lane's own active rows (slots) / box bounds (6 bits)
  auto set_active = [&](int con, bool on) {
    const int id = con & 0xffff;
    if (id >= MDS_QP_BOX0) {
      const int k = id - MDS_QP_BOX0 - 6 * n;
      if (k >= 0 && k < 6) boxmask = on ? (boxmask | (1u << k)) : (boxmask & ~(1u << k));
    } else {
      const int s = id - n * M.RPL;
      if (s >= 0 && s < M.RPL) rowmask = on ? (rowmask | (1u << s)) : (rowmask & ~(1u << s));
    }
  };
  Real xn[3] = {Real(0), Real(0), Real(0)};
  Real x3 = Real(0);
  if (valid) { R4 v = x[n]; xn[0] = v.x; xn[1] = v.y; xn[2] = v.z; x3 = v.w; }
  auto publish_x = [&]() {
    if (valid) { R4 v; v.x = xn[0]; v.y = xn[1]; v.z = xn[2]; v.w = x3; x[n] = v; }
  };
  int q = 0, iters = 0, status = MDS_QP_OPTIMAL;
  int p = p0;
  QpCon<Real> cp = qp_get(rows, C, p);
  W lam_p = W(0);  // meaningful on lane 0 only
  bool need_scan = false;
  if (!BIG) {  // ---- first iteration, empty active set
    Real mag, gx = qp_dot_x(cp, x, &mag);
    Real t = -(cp.rhs - gx) / cp.g2;  // g2 > 0: zero rows are reported by the scan as infeasible
    __syncwarp(gmask);                // every lane has read x before anyone overwrites it
#pragma unroll
    for (int k = 0; k < 3; ++k) xn[k] -= t * qp_coef(cp, n, k);
    publish_x();
    if (n == 0) { ws_int(ws) = p; lam[0] = (W)t; Lc[0] = (W)sqrt_(cp.g2); }
    set_active(p, true);
    q = 1; iters = 1;
    need_scan = true;
    __syncwarp(gmask);
  }
  // ONE flat loop (scan-if-needed -> scalar part -> primal step -> add or drop) instead of nested
  // outer/inner loops: the lane groups of a warp then stay converged on the same instructions even when
  // one group takes a full step and another a partial step (nested loops serialised the groups, ~4x).
  for (;;) {
    if (need_scan) {
      p = qp_scan(C, rows, x, xn, M, N, NP, n, valid, rowmask, boxmask, gmask);
      if (p == -1) break;  // optimal
      if (p == -2) { status = MDS_QP_INFEASIBLE; break; }
      cp = qp_get(rows, C, p);
      lam_p = W(0);
    }
    ++iters;
    // ---- scalar part (lane 0): d = L^-1 Na g_p, zn = |g_p|^2 - |d|^2, r = L^-T d, step lengths, multipliers
    if (n == 0) {
      int action = MDS_QP_ACT_FULL, st = MDS_QP_OPTIMAL, kdrop = -1;
      W t = W(0), zn = (W)cp.g2;
      if (iters > max_iter) {
        action = MDS_QP_ACT_STOP; st = MDS_QP_ITER_CAP;
      } else {
        W dd = W(0);
        for (int a = 0; a < q; ++a) {
          QpCon<Real> ca = qp_get(rows, C, ws_int(ws + a));
          W sacc = qp_dot_g<W>(ca, cp);
          for (int k = 0; k < a; ++k) sacc -= Lc[a * (a + 1) / 2 + k] * dv[k];
          sacc /= Lc[a * (a + 1) / 2 + a];
          dv[a] = sacc;
          dd += sacc * sacc;
        }
        zn = (W)cp.g2 - dd;
        for (int a = q - 1; a >= 0; --a) {
          W sacc = dv[a];
          for (int k = a + 1; k < q; ++k) sacc -= Lc[k * (k + 1) / 2 + a] * rv[k];
          rv[a] = sacc / Lc[a * (a + 1) / 2 + a];
        }
        W t1 = INF;
        for (int a = 0; a < q; ++a) {
          W ra = rv[a];
          if (ra > tol) {
            W cnd = lam[a] / ra;
            if (cnd < t1) { t1 = cnd; kdrop = a; }
          }
        }
        Real mag, gx = qp_dot_x(cp, x, &mag);
        W s_p = (W)(cp.rhs - gx);
        W t2 = (zn > zn_eps * (W)cp.g2) ? -s_p / zn : INF;
        t = t1 < t2 ? t1 : t2;
        if (t >= INF) {
          action = MDS_QP_ACT_STOP; st = MDS_QP_INFEASIBLE;
        } else {
          for (int a = 0; a < q; ++a) lam[a] -= t * rv[a];
          lam_p += t;
          if (t2 <= t1) {
            action = (q == qmax) ? MDS_QP_ACT_STOP : MDS_QP_ACT_FULL;
            if (q == qmax) st = MDS_QP_ITER_CAP;
          } else {
            action = MDS_QP_ACT_DROP;
          }
          if (t2 >= INF) t = -t;  // sign bit tells the lanes "dual step only, no primal move"
        }
      }
      hdr[0] = t;
      ws_int(hdr + 1) = action;
      ws_int(hdr + 2) = kdrop;
      ws_int(hdr + 3) = st;
    }
    __syncwarp(gmask);
    const int action = ws_int(hdr + 1), kdrop = ws_int(hdr + 2);
    if (action == MDS_QP_ACT_STOP) { status = ws_int(hdr + 3); break; }
    const Real t = (Real)hdr[0];
    // ---- own block of z = g_p - Na' r and the primal step (skipped for a pure dual step)
    if (!(t < Real(0)) && valid) {
      Real zk[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) zk[k] = qp_coef(cp, n, k);
      for (int a = 0; a < q; ++a) {
        QpCon<Real> ca = qp_get(rows, C, ws_int(ws + a));
        Real ra = (Real)rv[a];
#pragma unroll
        for (int k = 0; k < 3; ++k) zk[k] -= ra * qp_coef(ca, n, k);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) xn[k] -= t * zk[k];
      publish_x();
    }
    const int pd = (action == MDS_QP_ACT_DROP) ? ws_int(ws + kdrop) : -1;
    __syncwarp(gmask);  // x updated; every lane has consumed act / r of this iteration
    if (action == MDS_QP_ACT_FULL) {  // constraint p becomes active: append its row to the Cholesky factor
      if (n == 0) {
        W dd = W(0);
        for (int k = 0; k < q; ++k) { W v = dv[k]; Lc[q * (q + 1) / 2 + k] = v; dd += v * v; }
        Lc[q * (q + 1) / 2 + q] = sqrt_((W)cp.g2 - dd);
        ws_int(ws + q) = p;
        lam[q] = lam_p;
      }
      set_active(p, true);
      ++q;
      need_scan = true;
    } else {  // partial step: drop constraint kdrop, rebuild the (small) factor, keep working on p
      set_active(pd, false);
      --q;
      need_scan = false;
      if (n == 0) {
        for (int a = kdrop; a < q; ++a) { ws_int(ws + a) = ws_int(ws + a + 1); lam[a] = lam[a + 1]; }
        int st = MDS_QP_OPTIMAL;
        for (int a = 0; a < q; ++a) {
          QpCon<Real> ca = qp_get(rows, C, ws_int(ws + a));
          for (int b2 = 0; b2 <= a; ++b2) {
            QpCon<Real> cb = qp_get(rows, C, ws_int(ws + b2));
            W sacc = qp_dot_g<W>(ca, cb);
            for (int k = 0; k < b2; ++k) sacc -= Lc[a * (a + 1) / 2 + k] * Lc[b2 * (b2 + 1) / 2 + k];
            if (a == b2) {
              if (sacc <= W(0)) { st = MDS_QP_ITER_CAP; sacc = W(1); }
              Lc[a * (a + 1) / 2 + a] = sqrt_(sacc);
            } else {
              Lc[a * (a + 1) / 2 + b2] = sacc / Lc[b2 * (b2 + 1) / 2 + b2];
            }
          }
        }
        ws_int(hdr + 3) = st;
      }
    }
    __syncwarp(gmask);
    if (action == MDS_QP_ACT_DROP && ws_int(hdr + 3) != MDS_QP_OPTIMAL) { status = ws_int(hdr + 3); break; }
  }
  __syncwarp(gmask);
  if (sizeof(Real) == 4 && status == MDS_QP_OPTIMAL && q >= 3) {
    // fp32 polish: the iterate has been moved by `iters` incremental steps and the factor grown row by row, which
    // loses ~1e-4 on long solves.  With the final active set A the minimiser is u_nom - A' lam, (A A') lam = A u_nom - b;
    // lane 0 solves this small system once in double and every lane rebuilds its own block from u_nom.
    if (n == 0) {
      bool ok;
      if (BIG) {
        double* Lm = reinterpret_cast<double*>(hdr + 4);
        ok = qp_polish_core<Real, W>(C.umax[0], C.umax[1], C.umax[2], rows, xnom, ws, Lm, Lm + qmax * (qmax + 1) / 2, lam, q);
      } else {
        ok = qp_polish_small<Real>(C.umax[0], C.umax[1], C.umax[2], rows, xnom, reinterpret_cast<Real*>(ws), reinterpret_cast<Real*>(lam), q);
      }
      ws_int(hdr + 1) = ok ? 1 : 0;
    }
    __syncwarp(gmask);
    if (ws_int(hdr + 1) && valid) {
      auto un = xnom[n];
      double z0 = un.x, z1 = un.y, z2 = un.z;
      for (int a = 0; a < q; ++a) {
        QpCon<Real> ca = qp_get(rows, C, ws_int(ws + a));
        const double la = (double)lam[a];
        z0 -= la * (double)qp_coef(ca, n, 0); z1 -= la * (double)qp_coef(ca, n, 1); z2 -= la * (double)qp_coef(ca, n, 2);
      }
      xn[0] = (Real)z0; xn[1] = (Real)z1; xn[2] = (Real)z2;
      publish_x();
    }
    __syncwarp(gmask);
  }
  if (status == MDS_QP_OPTIMAL && q > 1) {
    // certify: rows held active must still be satisfied (guards breakdown on nearly dependent active sets)
    const Real ctol = sizeof(Real) == 4 ? Real(1e-3) : Real(1e-7);
    for (int a = 0; a < q; ++a) {
      QpCon<Real> ca = qp_get(rows, C, ws_int(ws + a));
      Real mag, gx = qp_dot_x(ca, x, &mag);
      if (ca.rhs - gx < -ctol * (abs_(ca.rhs) + mag + Real(1e-12))) status = MDS_QP_ITER_CAP;
    }
  }
  *iters_out = iters;
  return status;
}

// The scratch solver, out of line: rare (active sets beyond MDS_QP_QMAX, factor breakdown), so its code stays off the step loop.
// Lane 0 of the group claims a slot of the global scratch (flags: 0 free / 1 taken; a busy pool is waited for -- holders
// never wait for anything, so the wait ends), the group repeats the solve from u_nom with the scalar part in double, and
// lane 0 releases the slot.  Everything comes BY VALUE: a reference to the caller's parameter block (or to any of its
// locals) would pin that object in local memory for the whole kernel.  Returns status | iterations << 8.
template <typename Real>
__device__ __noinline__ int qp_solve_group_big(Real umax0, Real umax1, Real umax2, QpScratch S, const typename Vec4T<Real>::type* rows,
                                               typename Vec4T<Real>::type* x, const typename Vec4T<Real>::type* xnom, RowMap M, int N, int NP, int n,
                                               bool valid, unsigned gmask, int p0) {
  if (S.base == nullptr || S.slots <= 0) return MDS_QP_ITER_CAP;
  CbfP<Real> C;  // the solver reads the box bounds only
  C.umax[0] = umax0; C.umax[1] = umax1; C.umax[2] = umax2; C.max_iter = 0;
  const int lane0 = __ffs(gmask) - 1;
  int slot = -1;
  if (n == 0) {
    unsigned h = (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
    for (;;) {
      slot = (int)(h % (unsigned)S.slots);
      if (atomicCAS(S.flags + slot, 0, 1) == 0) break;
      h = h * 1664525u + 1013904223u;
    }
    __threadfence();
  }
  slot = __shfl_sync(gmask, slot, lane0);
  double* ws = S.base + (size_t)slot * (size_t)S.slot_doubles;
  int iters = 0;
  const int st = qp_solve_group<Real, double, true>(C, rows, x, xnom, ws, S.qmax, M, N, NP, n, valid, gmask, p0, &iters);
  __syncwarp(gmask);
  if (n == 0) {
    __threadfence();
    atomicExch(S.flags + slot, 0);
  }
  return st | (iters << 8);
}

// Obstacle record (cx, cy, cz, r): r > 0 is the reference's sphere (super-ellipsoid barrier with the agents' zscale,
// Ds = r_safe + r; cbf.py:380-383); r < 0 encodes a VERTICAL CYLINDER of radius |r| and unbounded height through
// (cx, cy) -- the zscale -> infinity limit of the same barrier, h = (ex^2 + ey^2)^2 - Ds^4 (builder extension:
// the reference has no cylinder primitive; parity is against oracle/cbf.py only).
template <typename Real> MDS_DEV void obstacle_shape(const CbfP<Real>& C, Real r, Real* Ds, Real* c4inv) {
  *Ds = C.rs + abs_(r);
  *c4inv = r < Real(0) ? Real(0) : C.c4inv;
}

// decoupled 4th input (wz): box +-umax[3] merged with the order-3 force-bound rows, which the
// reference places on column 4i+3 (cbf.py:456-460).  Returns false when the interval is empty.
template <int ORD, typename Real> MDS_DEV bool cbf_wz_bounds(const CbfP<Real>& C, Real F, Real* lo, Real* hi) {
  *lo = -C.umax[3]; *hi = C.umax[3];
  if (cbf_order<ORD>(C) == 3 && C.state_bounds) {  // custom_force_bound_const, emitted only with do_state_bounds (cbf.py:473-476)
    *hi = min_(*hi, C.k2 * (C.fmax - F));
    *lo = max_(*lo, -(C.k2 * (F - C.fmin)));
  }
  return *lo <= *hi;
}

MDS_DEV int cbf_num_barrier_rows(int N, int n_obs) { return N * (N - 1) / 2 + N * n_obs; }

}  // namespace mds
