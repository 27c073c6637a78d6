// mds_ctrl.cuh -- tracking controllers, mixer and inner-loop PID as device functions.
// Replaces control/geometric.py:59-115, control/lqr/*.py compute()/cap_u()/compute_low_level(),
// control/low_level/thrust_omega_ctrl.py:81-132, yank_omega_ctrl.py:39-55 and
// utils/model_conversions.py:69-103,137-143 of the reference.  Appendix-B quirks of
// SURVEY.md are reproduced on purpose (parity is against the reference, not the textbook).
#pragma once
#include "mds_common.cuh"
#include "mds_traj.cuh"

namespace mds {

#define MDS_MIN_RPM 9440.3
#define MDS_PWM2RPM_SCALE 0.2685
#define MDS_PWM2RPM_CONST 4070.3
#define MDS_MIN_PWM 20000.0
#define MDS_MAX_PWM 65535.0

// scipy Rotation.from_quat(q).as_matrix(): normalises first (model_conversions.py:110)
template <typename Real> MDS_DEV M3<Real> quat_to_rot_scipy(Real x, Real y, Real z, Real w) {
  Real inv = rsqrt_(x * x + y * y + z * z + w * w);
  x *= inv; y *= inv; z *= inv; w *= inv;
  M3<Real> R;
  R.m[0] = Real(1) - Real(2) * (y * y + z * z); R.m[1] = Real(2) * (x * y - z * w); R.m[2] = Real(2) * (x * z + y * w);
  R.m[3] = Real(2) * (x * y + z * w); R.m[4] = Real(1) - Real(2) * (x * x + z * z); R.m[5] = Real(2) * (y * z - x * w);
  R.m[6] = Real(2) * (x * z - y * w); R.m[7] = Real(2) * (y * z + x * w); R.m[8] = Real(1) - Real(2) * (x * x + y * y);
  return R;
}

// calc_z_thrust: model_conversions.py:137-143
template <typename Real> MDS_DEV Real z_thrust(const DroneP<Real>& P, const Real rpm[4]) {
  return P.kf * rpm[0] * rpm[0] + P.kf * rpm[1] * rpm[1] + P.kf * rpm[2] * rpm[2] + P.kf * rpm[3] * rpm[3];  // summation order of calc_z_thrust
}

// action_to_input: RPM -> [f, tx, ty, tz], PLUS-frame mixer (model_conversions.py:69-83)
template <int SPEC = 0, typename Real> MDS_DEV void action_to_input(const DroneP<Real>& P, const Real rpm_in[4], Real u[4]) {
  Real T[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Real r = clamp_(rpm_in[i], Real(0), P.max_rpm);
    T[i] = P.kf * r * r;
  }
  Real kr = P.km / P.kf;
  u[0] = T[0] + T[1] + T[2] + T[3];
  u[3] = kr * (-T[0] + T[1] - T[2] + T[3]);
  if (PhysSpec<SPEC>::x_frame_mixer(P) && PhysSpec<SPEC>::drone_model(P) == MDS_DRONE_CF2X) {  // the allocation physics_substep applies to a CF2X
    const Real l2 = P.arm_l * Real(0.70710678118654752);
    u[1] = Real(P.cf2x_torque_sign) * l2 * (T[0] + T[1] - T[2] - T[3]);
    u[2] = l2 * (-T[0] + T[1] + T[2] - T[3]);
    return;
  }
  u[1] = P.arm_l * (T[1] - T[3]);
  u[2] = P.arm_l * (T[2] - T[0]);
}

// input_to_action: [f, tx, ty, tz] -> RPM (model_conversions.py:85-103); clamps u[0] >= 0 in
// place; closed-form inverse of the mixer; per-motor clip to [MIN_RPM^2 kf, MAX_THRUST] (B6).
template <int SPEC = 0, typename Real> MDS_DEV void input_to_action(const DroneP<Real>& P, Real u[4], Real rpm[4]) {
  u[0] = max_(u[0], Real(0));
  Real kr = P.km / P.kf;
  Real q = Real(0.25) * u[0], a = u[1] / (Real(2) * P.arm_l), b = u[2] / (Real(2) * P.arm_l), c = u[3] / (Real(4) * kr);
  Real T[4] = {q - b - c, q + a + c, q + b - c, q - a + c};
  if (PhysSpec<SPEC>::x_frame_mixer(P) && PhysSpec<SPEC>::drone_model(P) == MDS_DRONE_CF2X) {  // inverse of the X-frame allocation (extension; the reference is PLUS-only)
    const Real l2 = P.arm_l * Real(0.70710678118654752);
    a = Real(P.cf2x_torque_sign) * u[1] / (Real(4) * l2);
    b = u[2] / (Real(4) * l2);
    T[0] = q + a - b - c; T[1] = q + a + b + c; T[2] = q - a + b - c; T[3] = q - a - b + c;
  }
  const Real tmin = Real(MDS_MIN_RPM * MDS_MIN_RPM) * P.kf;
#pragma unroll
  for (int i = 0; i < 4; ++i) rpm[i] = sqrt_(clamp_(T[i], tmin, P.max_thrust) / P.kf);
}

// GeometricControl.compute (control/geometric.py:59-114) -> u = [f, tau]
template <typename Real>
MDS_DEV void geometric_input(const DroneP<Real>& P, const GeoP<Real>& G, const Obs<Real>& o, const Ref<Real>& r, Real u[4]) {
  const Real m = P.m;
  M3<Real> R = quat_to_rot_scipy(o.qx, o.qy, o.qz, o.qw);
  V3<Real> w = o.av;  // world rates used as body rates (B3)
  V3<Real> RTvd = mulT(R, r.v);
  V3<Real> ev = mulT(R, o.v) - RTvd;
  V3<Real> ep = o.p - r.p;
  V3<Real> RTa = mulT(R, r.a);
  V3<Real> grav = mulT(R, v3(Real(0), Real(0), m * G.g_ctrl));
  V3<Real> kp_ep = mulT(R, G.kp * ep);
  V3<Real> f_b = grav - m * kp_ep - (m * G.kv) * ev + m * (RTa - cross(w, RTvd));
  V3<Real> f_w = mul(R, f_b);
  Real fn = norm(f_w);
  Real tilt = acos_(f_w.z / fn);
  if (tilt > G.max_tilt) {  // :79-84 (B5)
    Real xy = sqrt_(f_w.x * f_w.x + f_w.y * f_w.y);
    Real sc = f_w.z * G.tan_max_tilt / xy;
    f_w.x *= sc;
    f_w.y *= sc;
    fn = norm(f_w);
  }
  f_b = mulT(R, f_w);
  Real sy, cy;
  sincos_(r.yaw, &sy, &cy);
  V3<Real> b1c = {cy, sy, Real(0)};
  V3<Real> b3 = (Real(1) / fn) * f_w;
  V3<Real> c1 = cross(b3, b1c);
  V3<Real> b2 = (Real(1) / norm(c1)) * c1;
  V3<Real> c2 = cross(b2, b3);
  V3<Real> b1 = (Real(1) / norm(c2)) * c2;
  M3<Real> Rd = from_cols(b1, b2, b3);
  V3<Real> b1c_dot = {-sy * r.yaw_rate, cy * r.yaw_rate, Real(0)};
  V3<Real> f_dot = (m / fn) * mul(R, G.kp * ev);  // :96 (B4: Kp)
  V3<Real> b3_dot = cross(cross(b3, f_dot), b3);
  V3<Real> num = cross(b1c_dot, b3) + cross(b1c, b3_dot);
  V3<Real> b2_dot = cross(cross(b2, (Real(1) / norm(cross(b1c, b3))) * num), b2);
  V3<Real> b1_dot = cross(b3_dot, b2) + cross(b3, b2_dot);
  M3<Real> Rd_dot = from_cols(b1_dot, b2_dot, b3_dot);
  M3<Real> W = matmul(Rd, Rd_dot);  // :102 (B1: transpose(0,1) is a numpy no-op)
  V3<Real> w_d = {W.m[7], W.m[2], W.m[3]};
  M3<Real> A = matmulTN(Rd, R);  // Rd^T R ; R^T Rd is its transpose
  // vee(A - A^T) with the reference's vee = (-M12, M02, -M01)
  V3<Real> eR = {Real(0.5) * G.kr * -(A.m[5] - A.m[7]), Real(0.5) * G.kr * (A.m[2] - A.m[6]), Real(0.5) * G.kr * -(A.m[1] - A.m[3])};
  V3<Real> ew = w - mulT(R, mul(Rd, w_d));
  V3<Real> t0 = {-eR.x - G.kw * ew.x, -eR.y - G.kw * ew.y, -eR.z - G.kw * ew.z};
  V3<Real> Jw = {P.ixx * w.x, P.iyy * w.y, P.izz * w.z};
  V3<Real> gy = cross(w, Jw);
  u[0] = max_(Real(0), f_b.z);
  u[1] = P.ixx * t0.x - gy.x;
  u[2] = P.iyy * t0.y - gy.y;
  u[3] = P.izz * t0.z - gy.z;
}

// LQR error state + u = -K e (+ hover) for the three parametrisations (B8-B10).
// variant: MDS_CTRL_LQR_TORQUE (dim 12), _OMEGA (9), _YANK (10).  Returns the UN-capped u.
template <typename Real>
MDS_DEV int lqr_error_state(const DroneP<Real>& P, int variant, const Obs<Real>& o, const Ref<Real>& r, Real e[12]) {
  Real sy = Real(0), cy = Real(1);
  if (r.yaw != Real(0)) sincos_(r.yaw, &sy, &cy);  // a zero reference yaw (the reference's mains) needs no sincos
  // Error attitude (lqr_omega_controller.py:97-104): as_euler('xyz') of R_eq^T R with R = from_euler('xyz', rpy)
  // = Rz(yaw) Ry(pitch) Rx(roll) and R_eq = Rz(yaw_d).  Rz(yaw_d)^T Rz(yaw) = Rz(yaw - yaw_d), and the obs pitch is
  // an asin() in [-pi/2, pi/2], so the Euler angles of the product are (roll, pitch, wrap(yaw - yaw_d)) in closed
  // form -- no rotation matrix, sincos or atan2 round trip (it was ~10 % of the controller kernel's instructions).
  e[0] = o.rpy.x;
  e[1] = o.rpy.y;
  {
    const Real two_pi = Real(6.283185307179586476925286766559), inv_two_pi = Real(0.15915494309189533576888376337251);
    Real d = o.rpy.z - r.yaw;
    e[2] = d - two_pi * rint_(d * inv_two_pi);
  }
  V3<Real> dp = o.p - r.p, dv = o.v - r.v;
  V3<Real> ep = {cy * dp.x + sy * dp.y, -sy * dp.x + cy * dp.y, dp.z};
  V3<Real> ev = {cy * dv.x + sy * dv.y, -sy * dv.x + cy * dv.y, dv.z};
  int dim;
  if (variant == MDS_CTRL_LQR_TORQUE) {
    dim = 12;
    V3<Real> dw = {o.av.x, o.av.y, o.av.z - r.yaw_rate};
    e[3] = cy * dw.x + sy * dw.y; e[4] = -sy * dw.x + cy * dw.y; e[5] = dw.z;
    e[6] = ev.x; e[7] = ev.y; e[8] = ev.z; e[9] = ep.x; e[10] = ep.y; e[11] = ep.z;
  } else if (variant == MDS_CTRL_LQR_OMEGA) {
    dim = 9;
    e[3] = ev.x; e[4] = ev.y; e[5] = ev.z; e[6] = ep.x; e[7] = ep.y; e[8] = ep.z;
  } else {
    dim = 10;
    e[3] = z_thrust(P, o.rpm) - P.m * P.g;
    e[4] = ev.x; e[5] = ev.y; e[6] = ev.z; e[7] = ep.x; e[8] = ep.y; e[9] = ep.z;
  }
  return dim;
}
template <typename Real>
MDS_DEV void lqr_input(const DroneP<Real>& P, const LqrP<Real>& L, int variant, const Obs<Real>& o, const Ref<Real>& r, Real u[4]) {
  Real e[12];
  const int dim = lqr_error_state(P, variant, o, r, e);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Real acc = Real(0);
    for (int k = 0; k < dim; ++k) acc += L.K[i * dim + k] * e[k];
    u[i] = -acc;
  }
  if (variant != MDS_CTRL_LQR_YANK) u[0] += P.m * P.g;
}
// fp32: two inputs per instruction (packed FMA on the transposed gain; the error component broadcasts)
#ifndef MDS_NO_LQR_PACK
MDS_DEV void lqr_input(const DroneP<float>& P, const LqrP<float>& L, int variant, const Obs<float>& o, const Ref<float>& r, float u[4]) {
  float e[12];
  const int dim = lqr_error_state(P, variant, o, r, e);
  const float2* Kt = reinterpret_cast<const float2*>(L.Kt);
  F2 a01(0.f), a23(0.f);
  for (int k = 0; k < dim; ++k) {
    F2 k01, k23;
    k01.v = Kt[2 * k]; k23.v = Kt[2 * k + 1];
    a01 = fma_(k01, F2(e[k]), a01);
    a23 = fma_(k23, F2(e[k]), a23);
  }
  u[0] = -a01.v.x; u[1] = -a01.v.y; u[2] = -a23.v.x; u[3] = -a23.v.y;
  if (variant != MDS_CTRL_LQR_YANK) u[0] += P.m * P.g;
}
#endif
// The same law with a gain of the drone's own (decentralised LQR: every drone carries the K of its learned model;
// decentralized_lqr_omega.py:212-231, decentralized_lqr.py:326-342).  K planes: element (i, k) of drone d at [(i*dim+k)*D + d].
template <typename Real>
MDS_DEV void dlqr_input(const DroneP<Real>& P, const Real* __restrict__ K, size_t D, size_t d, int variant, const Obs<Real>& o, const Ref<Real>& r,
                        Real u[4]) {
  Real e[12];
  const int dim = lqr_error_state(P, variant, o, r, e);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Real acc = Real(0);
    for (int k = 0; k < dim; ++k) acc += K[(size_t)(i * dim + k) * D + d] * e[k];
    u[i] = -acc;
  }
  if (variant != MDS_CTRL_LQR_YANK) u[0] += P.m * P.g;
}

// LQROmegaController.cap_u (lqr_omega_controller.py:116-119)
template <typename Real> MDS_DEV Real cap_thrust(const DroneP<Real>& P, Real f) {
  return clamp_(f, Real(4.0 * MDS_MIN_RPM * MDS_MIN_RPM) * P.kf, P.max_thrust);
}

template <typename Real> struct Pid {
  V3<Real> last_w, integ;
};
template <typename Real> struct PidP {
  typename Vec4T<Real>::type* a;
  typename Vec2T<Real>::type* b;
};
template <typename Real> inline PidP<Real> to_dev(const MdsPidState& s) {
  return PidP<Real>{(typename Vec4T<Real>::type*)s.a, (typename Vec2T<Real>::type*)s.b};
}
template <typename Real> MDS_DEV Pid<Real> load_pid(const PidP<Real>& s, int d) {
  auto a = s.a[d]; auto b = s.b[d];
  return Pid<Real>{{a.x, a.y, a.z}, {a.w, b.x, b.y}};
}
template <typename Real> MDS_DEV void store_pid(const PidP<Real>& s, int d, const Pid<Real>& p) {
  typename Vec4T<Real>::type a; typename Vec2T<Real>::type b;
  a.x = p.last_w.x; a.y = p.last_w.y; a.z = p.last_w.z; a.w = p.integ.x; b.x = p.integ.y; b.y = p.integ.z;
  s.a[d] = a; s.b[d] = b;
}

// ThrustOmegaController.computeControlFromInput + omega_PID (thrust_omega_ctrl.py:81-132, B11).
// `thrust` already resolved (yank variant integrates before calling).  w_b = body rates.
template <int SPEC = 0, typename Real>
MDS_DEV void thrust_omega_pid(const DroneP<Real>& P, Pid<Real>& s, Real thrust, V3<Real> w_target, V3<Real> w_b, Real rpm[4]) {
  const Real dt = P.dt_ctrl;
  thrust = max_(thrust, Real(0));
  // divisions by constants as multiplications by their host-side reciprocals (an fp64 division is a ~28-instruction sequence)
  Real pwm_t = clamp_((sqrt_(thrust * P.inv_4kf) - Real(MDS_PWM2RPM_CONST)) * Real(1.0 / MDS_PWM2RPM_SCALE),
                      Real(MDS_MIN_PWM), Real(MDS_MAX_PWM));
  V3<Real> e = w_target - w_b;
  s.last_w = w_b;
  s.integ = s.integ - dt * e;
  s.integ.x = clamp_(s.integ.x, Real(-1), Real(1));  // the reference clips to +-1500 and then x, y to +-1: the second clip subsumes the first
  s.integ.y = clamp_(s.integ.y, Real(-1), Real(1));
  s.integ.z = clamp_(s.integ.z, Real(-1500), Real(1500));
  // the reference's derivative gain is 0 (control/low_level/thrust_omega_ctrl.py:42, 124: D_COEFF_OMEGA_TOR = 0): its term kd * (w_b - last_w) / (-dt) adds an
  // exact +0 for every finite rate and is left out (last_w is still carried as state)
  const Real kp = Real(17500), ki = Real(10);
  V3<Real> tq = {clamp_(kp * e.x + ki * s.integ.x, Real(-3200), Real(3200)),
                 clamp_(kp * e.y + ki * s.integ.y, Real(-3200), Real(3200)),
                 clamp_(kp * e.z + ki * s.integ.z, Real(-3200), Real(3200))};
  Real pw[4];
  if (PhysSpec<SPEC>::drone_model(P) == MDS_DRONE_CF2X) {
    pw[0] = pwm_t + (Real(-0.5) * tq.x - Real(0.5) * tq.y - tq.z);
    pw[1] = pwm_t + (Real(-0.5) * tq.x + Real(0.5) * tq.y + tq.z);
    pw[2] = pwm_t + (Real(0.5) * tq.x + Real(0.5) * tq.y - tq.z);
    pw[3] = pwm_t + (Real(0.5) * tq.x - Real(0.5) * tq.y + tq.z);
  } else {
    pw[0] = pwm_t + (-tq.y - tq.z);
    pw[1] = pwm_t + (tq.x + tq.z);
    pw[2] = pwm_t + (tq.y - tq.z);
    pw[3] = pwm_t + (-tq.x + tq.z);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) rpm[i] = Real(MDS_PWM2RPM_SCALE) * clamp_(pw[i], Real(MDS_MIN_PWM), Real(MDS_MAX_PWM)) + Real(MDS_PWM2RPM_CONST);
}

// ---------------------------------------------------------------------------------------------------
// Upstream gym-pybullet-drones DSLPIDControl.computeControlFromState (external to the reference; call site
// MultiDroneExample.py:111-114, gains halved at :85-92; SURVEY.md App. A.5 -- PARITY UNPINNED like the env step):
// Crazyflie position PID -> thrust + desired attitude -> attitude PID -> PWM mixer -> RPM.
template <typename Real> struct DslP { Real pf[3], if_[3], df[3], pt[3], it[3], dtq[3]; };
template <typename Real> inline DslP<Real> to_dev(const MdsDslPidGains& g) {
  DslP<Real> d;
  for (int i = 0; i < 3; ++i) {
    d.pf[i] = Real(g.p_for[i]); d.if_[i] = Real(g.i_for[i]); d.df[i] = Real(g.d_for[i]);
    d.pt[i] = Real(g.p_tor[i]); d.it[i] = Real(g.i_tor[i]); d.dtq[i] = Real(g.d_tor[i]);
  }
  return d;
}
template <typename Real> struct DslState {
  V3<Real> ipos, last_rpy, irpy;
};
// SoA planes: a Real4 = (ipos, last_roll), b Real4 = (last_pitch, last_yaw, irpy.x, irpy.y), c Real = irpy.z
template <typename Real> struct DslStateP {
  typename Vec4T<Real>::type *a, *b;
  Real* c;
};
template <typename Real> inline DslStateP<Real> to_dev(const MdsDslPidState& s) {
  using R4 = typename Vec4T<Real>::type;
  return DslStateP<Real>{(R4*)s.a, (R4*)s.b, (Real*)s.c};
}
template <typename Real> MDS_DEV DslState<Real> load_dsl(const DslStateP<Real>& p, int d) {
  auto a = p.a[d]; auto b = p.b[d];
  return DslState<Real>{{a.x, a.y, a.z}, {a.w, b.x, b.y}, {b.z, b.w, p.c[d]}};
}
template <typename Real> MDS_DEV void store_dsl(const DslStateP<Real>& p, int d, const DslState<Real>& s) {
  typename Vec4T<Real>::type a, b;
  a.x = s.ipos.x; a.y = s.ipos.y; a.z = s.ipos.z; a.w = s.last_rpy.x;
  b.x = s.last_rpy.y; b.y = s.last_rpy.z; b.z = s.irpy.x; b.w = s.irpy.y;
  p.a[d] = a; p.b[d] = b; p.c[d] = s.irpy.z;
}
template <typename Real>
MDS_DEV void dslpid_control(const DroneP<Real>& P, const DslP<Real>& G, DslState<Real>& s, const Obs<Real>& o, V3<Real> tpos, V3<Real> trpy,
                            V3<Real> tvel, V3<Real> trates, Real rpm[4], V3<Real>* pos_e_out) {
  const Real dt = P.dt_ctrl;
  M3<Real> R = quat_to_rot_scipy(o.qx, o.qy, o.qz, o.qw);
  // position loop: PID on position / velocity error + gravity feed-forward; integral clamps +-2 and +-0.15 on z
  V3<Real> pe = tpos - o.p, ve = tvel - o.v;
  s.ipos.x = clamp_(s.ipos.x + pe.x * dt, Real(-2), Real(2));
  s.ipos.y = clamp_(s.ipos.y + pe.y * dt, Real(-2), Real(2));
  s.ipos.z = clamp_(clamp_(s.ipos.z + pe.z * dt, Real(-2), Real(2)), Real(-0.15), Real(0.15));
  V3<Real> tt = {G.pf[0] * pe.x + G.if_[0] * s.ipos.x + G.df[0] * ve.x, G.pf[1] * pe.y + G.if_[1] * s.ipos.y + G.df[1] * ve.y,
                 G.pf[2] * pe.z + G.if_[2] * s.ipos.z + G.df[2] * ve.z + P.g * P.m};
  Real scalar_thrust = max_(Real(0), tt.x * R.m[2] + tt.y * R.m[5] + tt.z * R.m[8]);
  Real thrust = (sqrt_(scalar_thrust / (Real(4) * P.kf)) - Real(MDS_PWM2RPM_CONST)) / Real(MDS_PWM2RPM_SCALE);
  V3<Real> zax = (Real(1) / norm(tt)) * tt;
  Real sy, cy;
  sincos_(trpy.z, &sy, &cy);
  V3<Real> xc = {cy, sy, Real(0)};
  V3<Real> yc = cross(zax, xc);
  V3<Real> yax = (Real(1) / norm(yc)) * yc;
  V3<Real> xax = cross(yax, zax);
  M3<Real> Rd = from_cols(xax, yax, zax);
  // attitude loop: rot_e = vee(Rd^T R - R^T Rd), rate error from finite-differenced rpy
  M3<Real> A = matmulTN(Rd, R);  // Rd^T R; R^T Rd is its transpose
  V3<Real> rot_e = {A.m[7] - A.m[5], A.m[2] - A.m[6], A.m[3] - A.m[1]};
  Real inv_dt = Real(1) / dt;
  V3<Real> rate_e = {trates.x - (o.rpy.x - s.last_rpy.x) * inv_dt, trates.y - (o.rpy.y - s.last_rpy.y) * inv_dt,
                     trates.z - (o.rpy.z - s.last_rpy.z) * inv_dt};
  s.last_rpy = o.rpy;
  s.irpy.x = clamp_(clamp_(s.irpy.x - rot_e.x * dt, Real(-1500), Real(1500)), Real(-1), Real(1));
  s.irpy.y = clamp_(clamp_(s.irpy.y - rot_e.y * dt, Real(-1500), Real(1500)), Real(-1), Real(1));
  s.irpy.z = clamp_(s.irpy.z - rot_e.z * dt, Real(-1500), Real(1500));
  V3<Real> tq = {clamp_(-G.pt[0] * rot_e.x + G.dtq[0] * rate_e.x + G.it[0] * s.irpy.x, Real(-3200), Real(3200)),
                 clamp_(-G.pt[1] * rot_e.y + G.dtq[1] * rate_e.y + G.it[1] * s.irpy.y, Real(-3200), Real(3200)),
                 clamp_(-G.pt[2] * rot_e.z + G.dtq[2] * rate_e.z + G.it[2] * s.irpy.z, Real(-3200), Real(3200))};
  Real pw[4];
  if (P.drone_model == MDS_DRONE_CF2X) {
    pw[0] = thrust + (Real(-0.5) * tq.x - Real(0.5) * tq.y - tq.z);
    pw[1] = thrust + (Real(-0.5) * tq.x + Real(0.5) * tq.y + tq.z);
    pw[2] = thrust + (Real(0.5) * tq.x + Real(0.5) * tq.y - tq.z);
    pw[3] = thrust + (Real(0.5) * tq.x - Real(0.5) * tq.y + tq.z);
  } else {
    pw[0] = thrust + (-tq.y - tq.z);
    pw[1] = thrust + (tq.x + tq.z);
    pw[2] = thrust + (tq.y - tq.z);
    pw[3] = thrust + (-tq.x + tq.z);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) rpm[i] = Real(MDS_PWM2RPM_SCALE) * clamp_(pw[i], Real(MDS_MIN_PWM), Real(MDS_MAX_PWM)) + Real(MDS_PWM2RPM_CONST);
  *pos_e_out = pe;
}

// compute_low_level (lqr_omega_controller.py:78-88, lqr_YO_controller.py:87-98): world ->
// body rates with scipy's normalised R, then the inner loop.  R_out (optional) receives that matrix: it is the attitude the
// physics step starts from, so the fused rollout hands it on instead of rebuilding it from the quaternion.
template <int SPEC = 0, typename Real>
MDS_DEV void low_level(const DroneP<Real>& P, int variant, Pid<Real>& s, const Real u[4], const Obs<Real>& o, Real rpm[4], M3<Real>* R_out = nullptr) {
  M3<Real> R = quat_to_rot_scipy(o.qx, o.qy, o.qz, o.qw);
  V3<Real> w_b = mulT(R, o.av);
  Real thrust = u[0];
  if (variant == MDS_CTRL_LQR_YANK) thrust = z_thrust(P, o.rpm) + u[0] * P.dt_ctrl;
  thrust_omega_pid<SPEC>(P, s, thrust, v3(u[1], u[2], u[3]), w_b, rpm);
  if (R_out) *R_out = R;
}

}  // namespace mds
