// mds_physics.cuh -- explicit rigid-body update of one drone for one physics sub-step.
// Replaces upstream BaseAviary._dynamics/_integrateQ (Physics.DYN)
// and the composite DYN_GND_DRAG_DW of App. A.4 (_groundEffect/_drag/_downwash force
// models fed into the same integrator, plus a ground-plane contact clamp).
#pragma once
#include "mds_common.cuh"

namespace mds {

// Downwash on the LOWER drone of a pair: dz > 0 its distance below the other, dxy2 the squared distance in
// the plane; only pairs within 10 m in the plane interact:  alpha * exp(-(dxy / beta)^2 / 2).
template <typename Real> MDS_DEV Real downwash_pair(const DroneP<Real>& P, Real dz, Real dxy2) {
  Real q = div_(P.prop_radius, Real(4) * max_(dz, P.dw_dz_clip));  // clip: App. A.4 regularisation (0 = upstream); callers pass dz > 0
  Real alpha = P.dw1 * q * q;
  Real beta = P.dw2 * dz + P.dw3;
  return alpha * exp_(div_(Real(-0.5) * dxy2, beta * beta));
}
// two pairs at once (fp32, packed arithmetic; the reciprocals and the exponentials are per half: MUFU has no packed form)
MDS_DEV F2 downwash_pair2(const DroneP<float>& P, F2 dz, F2 dxy2) {
  const float pr4 = 0.25f * P.prop_radius;
  const F2 q(pr4 / max_(dz.v.x, P.dw_dz_clip), pr4 / max_(dz.v.y, P.dw_dz_clip));
  const F2 alpha = F2(P.dw1) * q * q;
  const F2 beta = fma_(F2(P.dw2), dz, F2(P.dw3));
  const F2 b2 = beta * beta;
  const F2 arg = dxy2 * F2(-0.72134752044448170f) * F2(1.f / b2.v.x, 1.f / b2.v.y);  // -0.5 log2(e)
  return alpha * F2(exp2f(arg.v.x), exp2f(arg.v.y));
}
// the same as seen from drone i: contribution of a drone at pj
template <typename Real> MDS_DEV Real downwash_term(const DroneP<Real>& P, V3<Real> pi, V3<Real> pj) {
  Real dz = pj.z - pi.z;
  Real dx = pj.x - pi.x, dy = pj.y - pi.y;
  Real dxy2 = dx * dx + dy * dy;
  if (dz > Real(0) && dxy2 < Real(100)) return downwash_pair(P, dz, dxy2);
  return Real(0);
}

// cos(x) and sin(x) / x * h for the quaternion update, x = |w| h, h = dt / 2.  fp32: the polynomials of sincos_ on
// x^2 (|x| < pi/4, i.e. |w| < 377 rad/s at 240 Hz -- beyond that the general path), no division by |w| and no
// small-|w| branch; fp64: upstream's literal form.
MDS_DEV bool quat_step_coeffs(float wn2, float h, float* cs, float* k) {
  const float x2 = wn2 * h * h;
  if (x2 < 0.6f) {
    *k = h * fmaf(fmaf(fmaf(-1.9515295891e-4f, x2, 8.3321608736e-3f), x2, -1.6666654611e-1f), x2, 1.0f);
    *cs = fmaf(fmaf(fmaf(2.443315711809948e-5f, x2, -1.388731625493765e-3f), x2, 4.166664568298827e-2f), x2 * x2, fmaf(-0.5f, x2, 1.0f));
    return true;
  }
  const float wn = sqrtf(wn2);
  float sn;
  sincos_(wn * h, &sn, cs);
  *k = sn / wn;
  return true;
}
MDS_DEV bool quat_step_coeffs(double wn2, double h, double* cs, double* k) {
  const double wn = sqrt(wn2);
  if (!(wn > 1e-8)) return false;  // upstream leaves the quaternion alone
  double sn;
  sincos(wn * h, &sn, cs);
  *k = sn / wn;
  return true;
}

// Per-rotor thrust kf rpm^2 and drag torque km rpm^2, and the ground effect on the thrusts (per-prop height above the plane,
// clipped): generic form, and fp32 with two rotors per packed instruction (mds_common.cuh F2).
template <typename Real> MDS_DEV void rotor_forces(const DroneP<Real>& P, const Real rpm[4], Real f[4], Real zt[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[i] = P.kf * rpm[i] * rpm[i]; zt[i] = P.km * rpm[i] * rpm[i]; }
}
template <typename Real> MDS_DEV void ground_effect(const DroneP<Real>& P, Real pz, Real r6, Real r7, Real f[4]) {
  const Real pr4 = Real(0.25) * P.prop_radius;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Real h = pz + r6 * P.prop_x[i] + r7 * P.prop_y[i];
    h = max_(h, P.gnd_eff_h_clip);
    Real q = pr4 / h;
    f[i] = fma_(f[i] * P.gnd_eff_coeff, q * q, f[i]);
  }
}
#ifndef MDS_NO_ROTOR_PACK
MDS_DEV void rotor_forces(const DroneP<float>& P, const float rpm[4], float f[4], float zt[4]) {
  const F2 r01(rpm[0], rpm[1]), r23(rpm[2], rpm[3]);
  const F2 s01 = r01 * r01, s23 = r23 * r23;
  const F2 f01 = F2(P.kf) * s01, f23 = F2(P.kf) * s23, z01 = F2(P.km) * s01, z23 = F2(P.km) * s23;
  f[0] = f01.v.x; f[1] = f01.v.y; f[2] = f23.v.x; f[3] = f23.v.y;
  zt[0] = z01.v.x; zt[1] = z01.v.y; zt[2] = z23.v.x; zt[3] = z23.v.y;
}
MDS_DEV void ground_effect(const DroneP<float>& P, float pz, float r6, float r7, float f[4]) {
  const float pr4 = 0.25f * P.prop_radius;
  const F2 h01 = fma_(F2(r7), F2(P.prop_y[0], P.prop_y[1]), fma_(F2(r6), F2(P.prop_x[0], P.prop_x[1]), F2(pz)));
  const F2 h23 = fma_(F2(r7), F2(P.prop_y[2], P.prop_y[3]), fma_(F2(r6), F2(P.prop_x[2], P.prop_x[3]), F2(pz)));
  const F2 q01(pr4 / max_(h01.v.x, P.gnd_eff_h_clip), pr4 / max_(h01.v.y, P.gnd_eff_h_clip));
  const F2 q23(pr4 / max_(h23.v.x, P.gnd_eff_h_clip), pr4 / max_(h23.v.y, P.gnd_eff_h_clip));
  const F2 f01(f[0], f[1]), f23(f[2], f[3]);
  const F2 g01 = fma_(f01 * F2(P.gnd_eff_coeff), q01 * q01, f01), g23 = fma_(f23 * F2(P.gnd_eff_coeff), q23 * q23, f23);
  f[0] = g01.v.x; f[1] = g01.v.y; f[2] = g23.v.x; f[3] = g23.v.y;
}
#endif

// One sub-step.  `rpm` is the clipped action; `dw` the summed downwash (0 for DYN); `fext` an optional world-frame
// force; R = the rotation matrix of the CURRENT attitude (quat_to_mat(s.q), or the controller's copy of it).
// Returns the world angular velocity that PyBullet would be handed: R(q_old) * w_new.
template <int SPEC, typename Real>
MDS_DEV V3<Real> physics_substep(const DroneP<Real>& P, Drone<Real>& s, const Real rpm[4], Real dw, V3<Real> fext, const M3<Real>& R) {
  using S = PhysSpec<SPEC>;
  const Real dt = P.dt_phys;
  Real f[4], zt[4];
  rotor_forces(P, rpm, f, zt);
  V3<Real> extra = fext;
  if (S::physics(P) == MDS_PHYSICS_DYN_GND_DRAG_DW) {
    // ground effect: per-prop height above the plane, gated on |roll|, |pitch| < pi/2 of Bullet's getEulerFromQuaternion
    // without the atan2/asin: pitch is asin(sarg) (+-pi/2 exactly inside the +-0.99999 gimbal branches), roll = atan2(A, B)
    // is inside (-pi/2, pi/2) iff B > 0 (or A = B = 0).
    const Real sarg = Real(-2) * (s.qx * s.qz - s.qw * s.qy);
    const Real rollA = Real(2) * (s.qy * s.qz + s.qw * s.qx);
    const Real rollB = s.qw * s.qw - s.qx * s.qx - s.qy * s.qy + s.qz * s.qz;
    if (abs_(sarg) < Real(0.99999) && (rollB > Real(0) || (rollB == Real(0) && rollA == Real(0)))) {
      ground_effect(P, s.p.z, R.m[6], R.m[7], f);
    }
    // rotor-speed-scaled linear drag from the PREVIOUS clipped RPM
    Real wsum = Real(0.10471975511965977) * (s.rpm[0] + s.rpm[1] + s.rpm[2] + s.rpm[3]);  // 2 pi / 60
    extra.x -= P.drag_xy * wsum * s.v.x;
    extra.y -= P.drag_xy * wsum * s.v.y;
    extra.z -= P.drag_z * wsum * s.v.z;
  }
  Real thrust = (f[0] + f[1] + f[2] + f[3]) - dw;
  V3<Real> F = {fma_(R.m[2], thrust, extra.x), fma_(R.m[5], thrust, extra.y), fma_(R.m[8], thrust, extra.z) - P.m * P.g};
  V3<Real> tau;
  tau.z = -zt[0] + zt[1] - zt[2] + zt[3];
  if (S::drone_model(P) == MDS_DRONE_CF2X) {
    const Real l2 = P.arm_l * Real(0.70710678118654752);
    tau.x = Real(P.cf2x_torque_sign) * (f[0] + f[1] - f[2] - f[3]) * l2;
    tau.y = (-f[0] + f[1] + f[2] - f[3]) * l2;
  } else {
    tau.x = (f[1] - f[3]) * P.arm_l;
    tau.y = (-f[0] + f[2]) * P.arm_l;
  }
  V3<Real> Jw = {P.ixx * s.w.x, P.iyy * s.w.y, P.izz * s.w.z};
  tau = tau - cross(s.w, Jw);
  const Real dtm = dt * P.inv_m;
  s.v = {fma_(dtm, F.x, s.v.x), fma_(dtm, F.y, s.v.y), fma_(dtm, F.z, s.v.z)};
  s.w = {fma_(dt * P.inv_ixx, tau.x, s.w.x), fma_(dt * P.inv_iyy, tau.y, s.w.y), fma_(dt * P.inv_izz, tau.z, s.w.z)};
  s.p = {fma_(dt, s.v.x, s.p.x), fma_(dt, s.v.y, s.p.y), fma_(dt, s.v.z, s.p.z)};  // semi-implicit Euler: uses the NEW velocity
  // quaternion exponential update with the NEW body rates (upstream _integrateQ)
  Real cs, k;
  if (quat_step_coeffs(dot(s.w, s.w), dt * Real(0.5), &cs, &k)) {
    Real p = s.w.x, q = s.w.y, r = s.w.z;
    Real x = s.qx, y = s.qy, z = s.qz, w = s.qw;
    s.qx = fma_(cs, x, k * (r * y - q * z + p * w));
    s.qy = fma_(cs, y, k * (-r * x + p * z + q * w));
    s.qz = fma_(cs, z, k * (q * x - p * y + r * w));
    s.qw = fma_(cs, w, k * (-p * x - q * y - r * z));
    if (S::renormalize(P)) {
      Real inv = rsqrt_(s.qx * s.qx + s.qy * s.qy + s.qz * s.qz + s.qw * s.qw);
      s.qx *= inv; s.qy *= inv; s.qz *= inv; s.qw *= inv;
    }
  }
  if (S::ground_clamp(P) && s.p.z < P.z_floor) {
    s.p.z = P.z_floor;
    s.v.z = max_(s.v.z, Real(0));
  }
  return mul(R, s.w);
}

}  // namespace mds
