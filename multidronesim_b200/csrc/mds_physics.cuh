// mds_physics.cuh -- explicit rigid-body update of one drone for one physics sub-step.
// Replaces upstream BaseAviary._dynamics/_integrateQ (Physics.DYN)
// and the composite DYN_GND_DRAG_DW of App. A.4 (_groundEffect/_drag/_downwash force
// models fed into the same integrator, plus a ground-plane contact clamp).
#pragma once
#include "mds_common.cuh"

namespace mds {

// Pairwise downwash on a drone at `pi` from one drone at `pj`: only drones
// above (dz > 0) and within 10 m in the plane contribute alpha * exp(-(dxy/beta)^2 / 2).
template <typename Real>
MDS_DEV Real downwash_term(const DroneP<Real>& P, V3<Real> pi, V3<Real> pj) {
  Real dz = pj.z - pi.z;
  Real dx = pj.x - pi.x, dy = pj.y - pi.y;
  Real dxy2 = dx * dx + dy * dy;
  if (dz > Real(0) && dxy2 < Real(100)) {
    Real q = P.prop_radius / (Real(4) * max_(dz, P.dw_dz_clip));  // clip: App. A.4 regularisation (0 = upstream)
    Real alpha = P.dw1 * q * q;
    Real beta = P.dw2 * dz + P.dw3;
    return alpha * exp_(Real(-0.5) * dxy2 / (beta * beta));
  }
  return Real(0);
}

// One sub-step.  `rpm` is the clipped action; `dw` the summed downwash (0 for DYN);
// `fext` an optional world-frame force.  Returns the world angular velocity that
// PyBullet would be handed: R(q_old) * w_new.
template <typename Real>
MDS_DEV V3<Real> physics_substep(const DroneP<Real>& P, Drone<Real>& s, const Real rpm[4], Real dw, V3<Real> fext) {
  const Real dt = P.dt_phys;
  M3<Real> R = quat_to_mat(s.qx, s.qy, s.qz, s.qw);
  Real f[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = P.kf * rpm[i] * rpm[i];
  V3<Real> extra = fext;
  if (P.physics == MDS_PHYSICS_DYN_GND_DRAG_DW) {
    // ground effect: per-prop height above the plane, gated on |roll|, |pitch| < pi/2
    // |roll| < pi/2 and |pitch| < pi/2 of Bullet's getEulerFromQuaternion without the atan2/asin: pitch is
    // asin(sarg) (+-pi/2 exactly inside the +-0.99999 gimbal branches), roll = atan2(A, B) is inside (-pi/2, pi/2)
    // iff B > 0 (or A = B = 0).
    const Real sarg = Real(-2) * (s.qx * s.qz - s.qw * s.qy);
    const Real rollA = Real(2) * (s.qy * s.qz + s.qw * s.qx);
    const Real rollB = s.qw * s.qw - s.qx * s.qx - s.qy * s.qy + s.qz * s.qz;
    if (abs_(sarg) < Real(0.99999) && (rollB > Real(0) || (rollB == Real(0) && rollA == Real(0)))) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        Real h = s.p.z + R.m[6] * P.prop_x[i] + R.m[7] * P.prop_y[i];
        h = max_(h, P.gnd_eff_h_clip);
        Real q = P.prop_radius / (Real(4) * h);
        f[i] = f[i] + f[i] * P.gnd_eff_coeff * q * q;
      }
    }
    // rotor-speed-scaled linear drag from the PREVIOUS clipped RPM
    Real wsum = Real(0.10471975511965977) * (s.rpm[0] + s.rpm[1] + s.rpm[2] + s.rpm[3]);  // 2 pi / 60
    extra.x -= P.drag_xy * wsum * s.v.x;
    extra.y -= P.drag_xy * wsum * s.v.y;
    extra.z -= P.drag_z * wsum * s.v.z;
  }
  Real thrust = (f[0] + f[1] + f[2] + f[3]) - dw;
  V3<Real> F = {R.m[2] * thrust + extra.x, R.m[5] * thrust + extra.y, R.m[8] * thrust + extra.z - P.m * P.g};
  Real zt0 = P.km * rpm[0] * rpm[0], zt1 = P.km * rpm[1] * rpm[1], zt2 = P.km * rpm[2] * rpm[2], zt3 = P.km * rpm[3] * rpm[3];
  V3<Real> tau;
  tau.z = -zt0 + zt1 - zt2 + zt3;
  if (P.drone_model == MDS_DRONE_CF2X) {
    const Real l2 = P.arm_l * Real(0.70710678118654752);
    tau.x = Real(P.cf2x_torque_sign) * (f[0] + f[1] - f[2] - f[3]) * l2;
    tau.y = (-f[0] + f[1] + f[2] - f[3]) * l2;
  } else {
    tau.x = (f[1] - f[3]) * P.arm_l;
    tau.y = (-f[0] + f[2]) * P.arm_l;
  }
  V3<Real> Jw = {P.ixx * s.w.x, P.iyy * s.w.y, P.izz * s.w.z};
  tau = tau - cross(s.w, Jw);
  V3<Real> wdot = {tau.x / P.ixx, tau.y / P.iyy, tau.z / P.izz};
  Real inv_m = Real(1) / P.m;
  s.v = s.v + dt * (inv_m * F);
  s.w = s.w + dt * wdot;
  s.p = s.p + dt * s.v;  // semi-implicit Euler: uses the NEW velocity
  // quaternion exponential update with the NEW body rates (upstream _integrateQ)
  Real wn = norm(s.w);
  if (wn > Real(1e-8)) {
    Real sn, cs;
    sincos_(wn * dt * Real(0.5), &sn, &cs);
    Real k = sn / wn;
    Real p = s.w.x, q = s.w.y, r = s.w.z;
    Real x = s.qx, y = s.qy, z = s.qz, w = s.qw;
    s.qx = cs * x + k * (r * y - q * z + p * w);
    s.qy = cs * y + k * (-r * x + p * z + q * w);
    s.qz = cs * z + k * (q * x - p * y + r * w);
    s.qw = cs * w + k * (-p * x - q * y - r * z);
    if (P.renormalize_quat) {
      Real inv = rsqrt_(s.qx * s.qx + s.qy * s.qy + s.qz * s.qz + s.qw * s.qw);
      s.qx *= inv; s.qy *= inv; s.qz *= inv; s.qw *= inv;
    }
  }
  if (P.ground_clamp && s.p.z < P.z_floor) {
    s.p.z = P.z_floor;
    s.v.z = max_(s.v.z, Real(0));
  }
  return mul(R, s.w);
}

}  // namespace mds
