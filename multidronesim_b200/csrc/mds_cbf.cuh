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

#define MDS_QP_QMAX 12  // max simultaneously active constraints handled (else status ITER_CAP)

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
// c4inv = 1 / zscale^4 for agents and sphere obstacles; 0 for a vertical-cylinder obstacle (the z terms vanish).
template <typename Real>
MDS_DEV void cbf_row(const DroneP<Real>& P, const CbfP<Real>& C, const CbfAgent<Real>& ai, const CbfAgent<Real>& aj,
                     Real Ds, Real c4inv, Real a3[3], Real* rhs, Real* h0_out) {
  V3<Real> e = ai.p - aj.p, dv = ai.dv - aj.dv, da = ai.da - aj.da;
  Real ex2 = e.x * e.x, ey2 = e.y * e.y, ez2 = e.z * e.z;
  Real rho = ex2 + ey2;
  V3<Real> d = {Real(4) * e.x * rho, Real(4) * e.y * rho, Real(4) * e.z * ez2 * c4inv};
  Real Hxx = Real(12) * ex2 + Real(4) * ey2, Hxy = Real(8) * e.x * e.y, Hyy = Real(4) * ex2 + Real(12) * ey2;
  Real Hzz = Real(12) * ez2 * c4inv;
  V3<Real> Hdv = {Hxx * dv.x + Hxy * dv.y, Hxy * dv.x + Hyy * dv.y, Hzz * dv.z};
  Real Ds2 = Ds * Ds;
  Real h0 = rho * rho + ez2 * ez2 * c4inv - Ds2 * Ds2;
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
                Real(24) * e.z * c4inv * dv.z * dv.z};
  Real Lf = Real(3) * dot(da, Hdv) + dot(q, dv);
  a3[0] = d.z * inv_m; a3[1] = -P.g * d.y; a3[2] = P.g * d.x;
  *rhs = C.k0 * h0 + C.k1 * h1 + C.k2 * h2 + Lf;
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
MDS_DEV int row_partner(const RowMap& M, int N, int n, int s) {
  if (s < M.K1) { int m = n + s + 1; return m >= N ? m - N : m; }
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
template <typename Real> MDS_DEV Real qp_dot_g(const QpCon<Real>& a, const QpCon<Real>& b) {
  Real ab = a.gi[0] * b.gi[0] + a.gi[1] * b.gi[1] + a.gi[2] * b.gi[2];
  Real s = Real(0);
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

// fp32 polish, lane-0 part (rare: solves that end with >= 3 active constraints): solve (A A') lam = A u_nom - b for the
// final active set in double and leave lam in the workspace.  Out of line: its fp64 arrays and code stay out of the
// step loop's registers and instruction stream.
template <typename Real>
__device__ __noinline__ bool qp_polish_lane0(Real umax0, Real umax1, Real umax2, const typename Vec4T<Real>::type* rows,
                                             const typename Vec4T<Real>::type* xnom, Real* ws, Real* lam, int q) {
  auto iref = [](Real* slot) -> int& { return *reinterpret_cast<int*>(slot); };
  CbfP<Real> C;  // only the box bounds are read (qp_get); taking the caller's block by reference would force it into local memory
  C.umax[0] = umax0; C.umax[1] = umax1; C.umax[2] = umax2;
  double Lm[MDS_QP_QMAX * (MDS_QP_QMAX + 1) / 2], y[MDS_QP_QMAX];
  bool ok = true;
  for (int a = 0; a < q && ok; ++a) {
    QpCon<Real> ca = qp_get(rows, C, iref(ws + a));
    for (int b2 = 0; b2 <= a; ++b2) {
      QpCon<Real> cb = qp_get(rows, C, iref(ws + b2));
      double ab = (double)ca.gi[0] * cb.gi[0] + (double)ca.gi[1] * cb.gi[1] + (double)ca.gi[2] * cb.gi[2], sacc = 0.0;
      if (ca.i == cb.i) sacc += ab;
      if (ca.j >= 0 && ca.j == cb.j) sacc += ab;
      if (ca.j >= 0 && ca.j == cb.i) sacc -= ab;
      if (cb.j >= 0 && cb.j == ca.i) sacc -= ab;
      for (int k = 0; k < b2; ++k) sacc -= Lm[a * (a + 1) / 2 + k] * Lm[b2 * (b2 + 1) / 2 + k];
      if (a == b2) {
        if (sacc <= 0.0) { ok = false; break; }
        Lm[a * (a + 1) / 2 + a] = sqrt(sacc);
      } else {
        Lm[a * (a + 1) / 2 + b2] = sacc / Lm[b2 * (b2 + 1) / 2 + b2];
      }
    }
    // right-hand side A u_nom - b from the nominal inputs kept in the 4th ... see below: unom is passed in xnom
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
    for (int a = 0; a < q; ++a) lam[a] = (Real)y[a];
  }
  return ok;
}

// Goldfarb-Idnani dual active set, P = I, executed by the env's lane group.
//   rows : smem barrier rows;  x : smem iterate, one Vec4 per drone (u_nom on entry, minimiser on exit);
//   xnom : smem copy of u_nom that stays untouched (fp32 polish);
//   ws   : smem workspace of MDS_QP_WS_WORDS Reals;  p0 : the most violated constraint at u_nom (packed),
//   found by the row builder's own scan.
// Work split: scans and the primal update are spread over the lanes (lane n owns drone n's inputs, rows and
// box bounds).  The first iteration (empty active set: z = g_p, t = -s_p / |g_p|^2) is computed redundantly by
// every lane with no workspace traffic -- in the C5 workload it is the only one for most environments.  From
// the second iteration on, the O(q^2) scalar part (triangular solves with the Cholesky factor of the active
// Gram matrix, step lengths, multiplier and factor updates) is done by the group's lane 0 and published
// through shared memory between __syncwarp(gmask) points.
enum { MDS_QP_ACT_FULL = 0, MDS_QP_ACT_DROP = 1, MDS_QP_ACT_STOP = 2 };

template <typename Real>
MDS_DEV int qp_solve_group(const CbfP<Real>& C, const typename Vec4T<Real>::type* rows, typename Vec4T<Real>::type* x,
                           const typename Vec4T<Real>::type* xnom, Real* ws, const RowMap& M, int N, int NP, int n, bool valid,
                           unsigned gmask, int p0, int* iters_out) {
  using R4 = typename Vec4T<Real>::type;
  const Real tol = qp_tol<Real>();
  const Real zn_eps = sizeof(Real) == 4 ? Real(1e-5) : Real(1e-10);
  const Real INF = Real(1e30);
  // workspace: one int per Real slot for act / header ints
  Real* lam = ws + MDS_QP_QMAX;
  Real* dv = ws + 2 * MDS_QP_QMAX;
  Real* rv = ws + 3 * MDS_QP_QMAX;
  Real* Lc = ws + 4 * MDS_QP_QMAX;  // packed lower-triangular Cholesky factor of the active Gram matrix
  Real* hdr = Lc + MDS_QP_QMAX * (MDS_QP_QMAX + 1) / 2;  // [0] t, [1] action, [2] dropped index, [3] status
  auto iref = [](Real* slot) -> int& { return *reinterpret_cast<int*>(slot); };
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
  int q = 0, iters = 1, status = MDS_QP_OPTIMAL;
  int p = p0;
  QpCon<Real> cp = qp_get(rows, C, p);
  Real lam_p = Real(0);  // meaningful on lane 0 only
  // ---- first iteration, empty active set
  {
    Real mag, gx = qp_dot_x(cp, x, &mag);
    Real t = -(cp.rhs - gx) / cp.g2;  // g2 > 0: zero rows are reported by the scan as infeasible
    __syncwarp(gmask);                // every lane has read x before anyone overwrites it
#pragma unroll
    for (int k = 0; k < 3; ++k) xn[k] -= t * qp_coef(cp, n, k);
    publish_x();
    if (n == 0) { iref(ws) = p; lam[0] = t; Lc[0] = sqrt_(cp.g2); }
    set_active(p, true);
    q = 1;
    __syncwarp(gmask);
  }
  bool need_scan = true;
  // ONE flat loop (scan-if-needed -> scalar part -> primal step -> add or drop) instead of nested
  // outer/inner loops: the lane groups of a warp then stay converged on the same instructions even when
  // one group takes a full step and another a partial step (nested loops serialised the groups, ~4x).
  for (;;) {
    if (need_scan) {
      p = qp_scan(C, rows, x, xn, M, N, NP, n, valid, rowmask, boxmask, gmask);
      if (p == -1) break;  // optimal
      if (p == -2) { status = MDS_QP_INFEASIBLE; break; }
      cp = qp_get(rows, C, p);
      lam_p = Real(0);
    }
    ++iters;
    // ---- scalar part (lane 0): d = L^-1 Na g_p, zn = |g_p|^2 - |d|^2, r = L^-T d, step lengths, multipliers
    if (n == 0) {
      int action = MDS_QP_ACT_FULL, st = MDS_QP_OPTIMAL, kdrop = -1;
      Real t = Real(0), zn = cp.g2;
      if (iters > C.max_iter) {
        action = MDS_QP_ACT_STOP; st = MDS_QP_ITER_CAP;
      } else {
        Real dd = Real(0);
        for (int a = 0; a < q; ++a) {
          QpCon<Real> ca = qp_get(rows, C, iref(ws + a));
          Real sacc = qp_dot_g(ca, cp);
          for (int k = 0; k < a; ++k) sacc -= Lc[a * (a + 1) / 2 + k] * dv[k];
          sacc /= Lc[a * (a + 1) / 2 + a];
          dv[a] = sacc;
          dd += sacc * sacc;
        }
        zn = cp.g2 - dd;
        for (int a = q - 1; a >= 0; --a) {
          Real sacc = dv[a];
          for (int k = a + 1; k < q; ++k) sacc -= Lc[k * (k + 1) / 2 + a] * rv[k];
          rv[a] = sacc / Lc[a * (a + 1) / 2 + a];
        }
        Real t1 = INF;
        for (int a = 0; a < q; ++a) {
          Real ra = rv[a];
          if (ra > tol) {
            Real cnd = lam[a] / ra;
            if (cnd < t1) { t1 = cnd; kdrop = a; }
          }
        }
        Real mag, gx = qp_dot_x(cp, x, &mag);
        Real s_p = cp.rhs - gx;
        Real t2 = (zn > zn_eps * cp.g2) ? -s_p / zn : INF;
        t = min_(t1, t2);
        if (t >= INF) {
          action = MDS_QP_ACT_STOP; st = MDS_QP_INFEASIBLE;
        } else {
          for (int a = 0; a < q; ++a) lam[a] -= t * rv[a];
          lam_p += t;
          if (t2 <= t1) {
            action = (q == MDS_QP_QMAX) ? MDS_QP_ACT_STOP : MDS_QP_ACT_FULL;
            if (q == MDS_QP_QMAX) st = MDS_QP_ITER_CAP;
          } else {
            action = MDS_QP_ACT_DROP;
          }
          if (t2 >= INF) t = -t;  // sign bit tells the lanes "dual step only, no primal move"
        }
      }
      hdr[0] = t;
      iref(hdr + 1) = action;
      iref(hdr + 2) = kdrop;
      iref(hdr + 3) = st;
    }
    __syncwarp(gmask);
    const int action = iref(hdr + 1), kdrop = iref(hdr + 2);
    if (action == MDS_QP_ACT_STOP) { status = iref(hdr + 3); break; }
    const Real t = hdr[0];
    // ---- own block of z = g_p - Na' r and the primal step (skipped for a pure dual step)
    if (!(t < Real(0)) && valid) {
      Real zk[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) zk[k] = qp_coef(cp, n, k);
      for (int a = 0; a < q; ++a) {
        QpCon<Real> ca = qp_get(rows, C, iref(ws + a));
        Real ra = rv[a];
#pragma unroll
        for (int k = 0; k < 3; ++k) zk[k] -= ra * qp_coef(ca, n, k);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) xn[k] -= t * zk[k];
      publish_x();
    }
    const int pd = (action == MDS_QP_ACT_DROP) ? iref(ws + kdrop) : -1;
    __syncwarp(gmask);  // x updated; every lane has consumed act / r of this iteration
    if (action == MDS_QP_ACT_FULL) {  // constraint p becomes active: append its row to the Cholesky factor
      if (n == 0) {
        Real dd = Real(0);
        for (int k = 0; k < q; ++k) { Real v = dv[k]; Lc[q * (q + 1) / 2 + k] = v; dd += v * v; }
        Lc[q * (q + 1) / 2 + q] = sqrt_(cp.g2 - dd);
        iref(ws + q) = p;
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
        for (int a = kdrop; a < q; ++a) { iref(ws + a) = iref(ws + a + 1); lam[a] = lam[a + 1]; }
        int st = MDS_QP_OPTIMAL;
        for (int a = 0; a < q; ++a) {
          QpCon<Real> ca = qp_get(rows, C, iref(ws + a));
          for (int b2 = 0; b2 <= a; ++b2) {
            QpCon<Real> cb = qp_get(rows, C, iref(ws + b2));
            Real sacc = qp_dot_g(ca, cb);
            for (int k = 0; k < b2; ++k) sacc -= Lc[a * (a + 1) / 2 + k] * Lc[b2 * (b2 + 1) / 2 + k];
            if (a == b2) {
              if (sacc <= Real(0)) { st = MDS_QP_ITER_CAP; sacc = Real(1); }
              Lc[a * (a + 1) / 2 + a] = sqrt_(sacc);
            } else {
              Lc[a * (a + 1) / 2 + b2] = sacc / Lc[b2 * (b2 + 1) / 2 + b2];
            }
          }
        }
        iref(hdr + 3) = st;
      }
    }
    __syncwarp(gmask);
    if (action == MDS_QP_ACT_DROP && iref(hdr + 3) != MDS_QP_OPTIMAL) { status = iref(hdr + 3); break; }
  }
  __syncwarp(gmask);
  if (sizeof(Real) == 4 && status == MDS_QP_OPTIMAL && q >= 3) {
    // fp32 polish: the iterate has been moved by `iters` incremental steps and the factor grown row by row, which
    // loses ~1e-4 on long solves.  With the final active set A the minimiser is u_nom - A' lam, (A A') lam = A u_nom - b;
    // lane 0 solves this small system once in double and every lane rebuilds its own block from u_nom.
    if (n == 0) {
      const bool ok = qp_polish_lane0<Real>(C.umax[0], C.umax[1], C.umax[2], rows, xnom, ws, lam, q);
      iref(hdr + 1) = ok ? 1 : 0;
    }
    __syncwarp(gmask);
    if (iref(hdr + 1) && valid) {
      auto un = xnom[n];
      double z0 = un.x, z1 = un.y, z2 = un.z;
      for (int a = 0; a < q; ++a) {
        QpCon<Real> ca = qp_get(rows, C, iref(ws + a));
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
      QpCon<Real> ca = qp_get(rows, C, iref(ws + a));
      Real mag, gx = qp_dot_x(ca, x, &mag);
      if (ca.rhs - gx < -ctol * (abs_(ca.rhs) + mag + Real(1e-12))) status = MDS_QP_ITER_CAP;
    }
  }
  *iters_out = iters;
  return status;
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
