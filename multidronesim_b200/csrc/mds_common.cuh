// mds_common.cuh -- Real-templated small math, parameter blocks and the SoA state
// accessors shared by every kernel of the hot path (sm_100a, no tensor cores: nothing
// here is a dense contraction; the rules that matter are coalesced 128-bit planes,
// shared-memory staging of env neighbours and grids sized from the SM count).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/mds_b200.h"

#define MDS_DEV __device__ __forceinline__

namespace mds {

// ---------------------------------------------------------------- vector types
template <typename Real> struct Vec4T;
template <> struct Vec4T<float> { using type = float4; };
template <> struct Vec4T<double> { using type = double4; };
template <typename Real> struct Vec2T;
template <> struct Vec2T<float> { using type = float2; };
template <> struct Vec2T<double> { using type = double2; };

template <typename Real> struct V3 { Real x, y, z; };
template <typename Real> struct M3 { Real m[9]; };  // row-major

template <typename Real> MDS_DEV V3<Real> v3(Real x, Real y, Real z) { return V3<Real>{x, y, z}; }
template <typename Real> MDS_DEV V3<Real> operator+(V3<Real> a, V3<Real> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename Real> MDS_DEV V3<Real> operator-(V3<Real> a, V3<Real> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename Real> MDS_DEV V3<Real> operator*(Real s, V3<Real> a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename Real> MDS_DEV Real dot(V3<Real> a, V3<Real> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename Real> MDS_DEV V3<Real> cross(V3<Real> a, V3<Real> b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename Real> MDS_DEV V3<Real> mul(const M3<Real>& A, V3<Real> v) {
  return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
          A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
template <typename Real> MDS_DEV V3<Real> mulT(const M3<Real>& A, V3<Real> v) {  // A^T v
  return {A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z,
          A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z};
}
template <typename Real> MDS_DEV M3<Real> matmul(const M3<Real>& A, const M3<Real>& B) {
  M3<Real> C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
template <typename Real> MDS_DEV M3<Real> matmulTN(const M3<Real>& A, const M3<Real>& B) {  // A^T B
  M3<Real> C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[i] * B.m[j] + A.m[3 + i] * B.m[3 + j] + A.m[6 + i] * B.m[6 + j];
  return C;
}
template <typename Real> MDS_DEV M3<Real> from_cols(V3<Real> a, V3<Real> b, V3<Real> c) {
  M3<Real> R;
  R.m[0] = a.x; R.m[1] = b.x; R.m[2] = c.x;
  R.m[3] = a.y; R.m[4] = b.y; R.m[5] = c.y;
  R.m[6] = a.z; R.m[7] = b.z; R.m[8] = c.z;
  return R;
}

// ---------------------------------------------------------------- scalar math
MDS_DEV float rsqrt_(float x) { return rsqrtf(x); }
MDS_DEV double rsqrt_(double x) { return 1.0 / sqrt(x); }
MDS_DEV float sqrt_(float x) { return sqrtf(x); }
MDS_DEV double sqrt_(double x) { return sqrt(x); }
// fp32 sincos for the bounded arguments of this path (|x| < ~1e4): Cody-Waite reduction by pi/2 and
// minimax polynomials on [-pi/4, pi/4] (max error ~1 ulp).  libdevice's sincosf drags a Payne-Hanek slow
// path (local array, ~200 instructions) into every call site, which the fused kernels pay in I-cache.
MDS_DEV void sincos_(float x, float* s, float* c) {
  float j = rintf(x * 0.636619772367581343f);
  int q = (int)j;
  float r = fmaf(j, -1.5703125f, x);
  r = fmaf(j, -4.837512969970703125e-4f, r);
  r = fmaf(j, -7.54978995489188216e-8f, r);
  float r2 = r * r;
  float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
  float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2 * r2, fmaf(-0.5f, r2, 1.0f));
  float ss = (q & 1) ? cp : sp, cc = (q & 1) ? sp : cp;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}
MDS_DEV void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
// fp32 atan2 from ONE division: the ratio of the smaller to the larger magnitude, reduced to |t| <= tan(pi/8) by
// atan(a) = pi/4 + atan((a - 1) / (a + 1)) with numerator and denominator selected before dividing, then cephes atanf's
// degree-9 odd polynomial (peak relative error 2e-7) and the octant fix-ups.  libdevice's atan2f is ~42 instructions and
// runs twice per drone-step (roll and yaw of the observation); this is ~24.
MDS_DEV float atan2_(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const bool mid = mn > 0.41421356237f * mx;
  const float num = mid ? mn - mx : mn, den = mid ? mn + mx : mx;
  const float t = den > 0.f ? num / den : 0.f;  // atan2(0, 0) = 0
  const float z = t * t;
  float r = fmaf(fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f), z * t, t);
  if (mid) r += 0.78539816339744831f;
  if (ay > ax) r = 1.57079632679489662f - r;
  if (x < 0.f) r = 3.14159265358979324f - r;
  return copysignf(r, y);
}
MDS_DEV float asin_(float x) { return asinf(x); }
MDS_DEV float acos_(float x) { return acosf(x); }
MDS_DEV double acos_(double x) { return acos(x); }
MDS_DEV float exp_(float x) { return expf(x); }
// fp64 exp for finite arguments in [-700, 700] (clamped: the one caller is the downwash Gaussian, argument <= 0, and
// exp(-700) = 1e-304 is zero to it): 2^k e^r with k = rint(x log2 e) by the 1.5 * 2^52 trick, r = x - k ln 2 in two parts,
// |r| <= 0.347, e^r as its degree-13 Taylor polynomial (truncation 4e-18), the exponent added as an integer.  24 instructions
// against libdevice's 65 (special cases, sub-normal results); measured 1 ulp against glibc on 2e7 arguments
// (tests/test_device_math.py re-evaluates these tables in numpy).
MDS_DEV double exp_(double x) {
  x = fmin(fmax(x, -700.0), 700.0);
  const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
  const int k = __double2loint(t);
  const double kd = t - 6755399441055744.0;
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;
  p = fma(p, r, 2.08767569878681e-09);
  p = fma(p, r, 2.505210838544172e-08);
  p = fma(p, r, 2.755731922398589e-07);
  p = fma(p, r, 2.7557319223985893e-06);
  p = fma(p, r, 2.48015873015873e-05);
  p = fma(p, r, 1.984126984126984e-04);
  p = fma(p, r, 1.388888888888889e-03);
  p = fma(p, r, 8.333333333333333e-03);
  p = fma(p, r, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __longlong_as_double(__double_as_longlong(p) + ((long long)k << 52));
}
// a / b and 1 / b.  fp32: the plain operators (approximate under -use_fast_math, IEEE otherwise).  fp64, for normal finite
// b != 0 and results in the normal range: MUFU.RCP64H + two Newton steps + one residual correction, 8 instructions against the
// 15 + out-of-line slow path nvcc emits for `/`; bit-equal to IEEE division on 2e7 random pairs in the CPU emulation.
MDS_DEV float rcp_(float b) { return 1.f / b; }
MDS_DEV float div_(float a, float b) { return a / b; }
MDS_DEV double rcp_(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  return fma(r, e, r);
}
MDS_DEV double div_(double a, double b) {
  const double r = rcp_(b), q = a * r;
  return fma(fma(-b, q, a), r, q);
}
// fp64 atan2 / asin of the observation's Euler angles: the fp32 scheme above (one division, reduction to |t| <= tan(pi/8),
// octant fix-ups) with an 11-term minimax polynomial in t^2, and asin as x + x z R(z) on |x| <= 1/2, pi/2 - 2 asin(sqrt((1 - |x|) / 2))
// beyond (13 terms; coefficients by Chebyshev fit at 60 digits).  ~40 and ~30 instructions against libdevice's 88 and 129;
// 1.3 / 2 ulp against glibc on 2e7 arguments.  Finite arguments only (the callers' are: quaternion products).
MDS_DEV double atan2_(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const double mx = fmax(ax, ay), mn = fmin(ax, ay);
  const bool mid = mn > 0.41421356237309503 * mx;
  const double num = mid ? mn - mx : mn, den = mid ? mn + mx : mx;
  const double t = den > 0.0 ? div_(num, den) : 0.0;  // atan2(0, 0) = 0
  const double z = t * t;
  double p = -0.019176887119062259;
  p = fma(p, z, 0.0392316582955871913);
  p = fma(p, z, -0.0508544973794025986);
  p = fma(p, z, 0.0585814891280220988);
  p = fma(p, z, -0.06664511447381948);
  p = fma(p, z, 0.0769218319082608656);
  p = fma(p, z, -0.0909090457812390189);
  p = fma(p, z, 0.111111110152563617);
  p = fma(p, z, -0.142857142846665429);
  p = fma(p, z, 0.199999999999955207);
  p = fma(p, z, -0.333333333333333302);
  double r = fma(t * z, p, t);
  if (mid) r += 0.78539816339744831;
  if (ay > ax) r = 1.5707963267948966 - r;
  if (x < 0.0) r = 3.1415926535897932 - r;
  return copysign(r, y);
}
MDS_DEV double asin_(double x) {
  const double a = fabs(x);
  const bool big = a > 0.5;
  const double z = big ? 0.5 * (1.0 - a) : a * a;
  const double s = big ? sqrt(z) : a;
  double p = 0.0287578513674215647;
  p = fma(p, z, -0.0148518870712472032);
  p = fma(p, z, 0.0174008794426940224);
  p = fma(p, z, 0.0054575067186403583);
  p = fma(p, z, 0.0103228143501857793);
  p = fma(p, z, 0.0114791774151849059);
  p = fma(p, z, 0.0139712129735529333);
  p = fma(p, z, 0.0173523927208699729);
  p = fma(p, z, 0.0223721729421498879);
  p = fma(p, z, 0.0303819441385312473);
  p = fma(p, z, 0.0446428571463554298);
  p = fma(p, z, 0.0749999999999843293);
  p = fma(p, z, 0.166666666666666678);
  double r = fma(s * z, p, s);
  if (big) r = 1.5707963267948966 - 2.0 * r;
  return copysign(r, x);
}
MDS_DEV float tan_(float x) { return tanf(x); }
MDS_DEV double tan_(double x) { return tan(x); }
MDS_DEV float fma_(float a, float b, float c) { return fmaf(a, b, c); }
MDS_DEV double fma_(double a, double b, double c) { return fma(a, b, c); }
MDS_DEV float rint_(float x) { return rintf(x); }
MDS_DEV double rint_(double x) { return rint(x); }
MDS_DEV float abs_(float x) { return fabsf(x); }
MDS_DEV double abs_(double x) { return fabs(x); }
MDS_DEV float min_(float a, float b) { return fminf(a, b); }
MDS_DEV double min_(double a, double b) { return fmin(a, b); }
MDS_DEV float max_(float a, float b) { return fmaxf(a, b); }
MDS_DEV double max_(double a, double b) { return fmax(a, b); }
template <typename Real> MDS_DEV Real clamp_(Real x, Real lo, Real hi) { return min_(max_(x, lo), hi); }

// Two independent fp32 values in one 64-bit register pair.  Blackwell's packed fp32 instructions (FFMA2 / FMUL2 / FADD2,
// PTX fma.rn.f32x2) do both lanes in ONE issue slot; the FMA pipe is busy for two cycles, so the flop rate is that of two
// scalar instructions (tools/ffma2_probe.cu: 74.1 vs 72.8 TFLOP/s) but the second issue slot goes to other work -- which
// is what an issue-bound kernel like the fused rollout is short of.  A scalar operand broadcasts for free (SASS `R7.F32`,
// `UR6.F32`, immediates), there is no negate modifier (a - b is written fma(b, -1, a): exact), and the halves stay
// addressable as ordinary registers.  The same expression templates run scalar (T = Real) or on two rows / rotors / angles
// at once (T = F2); each lane of a packed result is the IEEE result of the scalar instruction.
struct F2 {
  float2 v;
  F2() = default;
  MDS_DEV F2(float s) : v(make_float2(s, s)) {}
  MDS_DEV F2(float a, float b) : v(make_float2(a, b)) {}
};
MDS_DEV F2 operator+(F2 a, F2 b) { F2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
MDS_DEV F2 operator*(F2 a, F2 b) { F2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
MDS_DEV F2 fma_(F2 a, F2 b, F2 c) { F2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
MDS_DEV F2 operator-(F2 a, F2 b) { return fma_(b, F2(-1.f), a); }
// number of fp32 lanes of an arithmetic type, and lane access
template <typename T> struct Lanes { static constexpr int n = 1; };
template <> struct Lanes<F2> { static constexpr int n = 2; };
MDS_DEV float lane(float x, int) { return x; }
MDS_DEV double lane(double x, int) { return x; }
MDS_DEV float lane(F2 x, int i) { return i ? x.v.y : x.v.x; }
template <typename Real> MDS_DEV Real norm(V3<Real> a) { return sqrt_(dot(a, a)); }
template <typename Real> MDS_DEV Real sign_(Real x) { return (x > Real(0)) ? Real(1) : ((x < Real(0)) ? Real(-1) : Real(0)); }

// ---------------------------------------------------------------- device parameter blocks
template <typename Real> struct DroneP {
  Real m, g, kf, km, arm_l, ixx, iyy, izz, max_rpm, max_thrust;
  Real gnd_eff_coeff, prop_radius, gnd_eff_h_clip, drag_xy, drag_z, dw1, dw2, dw3, dw_dz_clip;
  Real prop_x[4], prop_y[4];
  Real z_floor, dt_phys, dt_ctrl;
  Real inv_m, inv_ixx, inv_iyy, inv_izz, inv_4kf;  // derived on the host (to_dev): no divisions by constants in the kernels
  int substeps, drone_model, physics, cf2x_torque_sign, renormalize_quat, ground_clamp, x_frame_mixer;
};
template <typename Real> inline DroneP<Real> to_dev(const MdsDroneParams& p) {
  DroneP<Real> d;
  d.m = Real(p.m); d.g = Real(p.g); d.kf = Real(p.kf); d.km = Real(p.km); d.arm_l = Real(p.arm_l);
  d.ixx = Real(p.ixx); d.iyy = Real(p.iyy); d.izz = Real(p.izz);
  d.max_rpm = Real(p.max_rpm); d.max_thrust = Real(p.max_thrust);
  d.gnd_eff_coeff = Real(p.gnd_eff_coeff); d.prop_radius = Real(p.prop_radius);
  d.gnd_eff_h_clip = Real(p.gnd_eff_h_clip); d.drag_xy = Real(p.drag_xy); d.drag_z = Real(p.drag_z);
  d.dw1 = Real(p.dw1); d.dw2 = Real(p.dw2); d.dw3 = Real(p.dw3); d.dw_dz_clip = Real(p.dw_dz_clip);
  for (int i = 0; i < 4; ++i) { d.prop_x[i] = Real(p.prop_x[i]); d.prop_y[i] = Real(p.prop_y[i]); }
  d.z_floor = Real(p.z_floor); d.dt_phys = Real(p.dt_phys); d.dt_ctrl = Real(p.dt_ctrl);
  d.inv_m = Real(1.0 / p.m); d.inv_ixx = Real(1.0 / p.ixx); d.inv_iyy = Real(1.0 / p.iyy); d.inv_izz = Real(1.0 / p.izz);
  d.inv_4kf = Real(1.0 / (4.0 * p.kf));
  d.substeps = p.substeps; d.drone_model = p.drone_model; d.physics = p.physics;
  d.cf2x_torque_sign = p.cf2x_torque_sign; d.renormalize_quat = p.renormalize_quat; d.ground_clamp = p.ground_clamp; d.x_frame_mixer = p.x_frame_mixer;
  return d;
}
// Compile-time view of the run-time switches of the parameter block.  SPEC = 0: read them from P (any configuration).
// SPEC = 1: the swarm configuration of the reference's CBF mains -- physics DYN_GND_DRAG_DW, CF2P, one physics sub-step
// per control period, no quaternion renormalisation, ground clamp on, PLUS-frame mixer; the host selects it when P matches
// exactly (mds_kernels.cu phys_spec_of), so the step loop carries none of the mode branches.
template <int SPEC> struct PhysSpec {
  template <typename Real> static MDS_DEV int physics(const DroneP<Real>& P) { return SPEC == 1 ? MDS_PHYSICS_DYN_GND_DRAG_DW : P.physics; }
  template <typename Real> static MDS_DEV int drone_model(const DroneP<Real>& P) { return SPEC == 1 ? MDS_DRONE_CF2P : P.drone_model; }
  template <typename Real> static MDS_DEV int substeps(const DroneP<Real>& P) { return SPEC == 1 ? 1 : P.substeps; }
  template <typename Real> static MDS_DEV bool renormalize(const DroneP<Real>& P) { return SPEC == 1 ? false : P.renormalize_quat != 0; }
  template <typename Real> static MDS_DEV bool ground_clamp(const DroneP<Real>& P) { return SPEC == 1 ? true : P.ground_clamp != 0; }
  template <typename Real> static MDS_DEV bool x_frame_mixer(const DroneP<Real>& P) { return SPEC == 1 ? false : P.x_frame_mixer != 0; }
};

template <typename Real> struct GeoP { Real kp, kv, kr, kw, g_ctrl, max_tilt, tan_max_tilt; };
template <typename Real> inline GeoP<Real> to_dev(const MdsGeoGains& g) {
  return GeoP<Real>{Real(g.kp), Real(g.kv), Real(g.kr), Real(g.kw), Real(g.g_ctrl), Real(g.max_tilt), Real(tan(g.max_tilt))};
}
// K: the reference's row-major 4 x dim gain; Kt: its transpose (Kt[4 k + i] = K[i dim + k]), so that the two gain entries a
// packed fp32 FMA multiplies one error component with are one 64-bit kernel-parameter operand (lqr_input, fp32).
template <typename Real> struct LqrP { Real K[48]; alignas(16) Real Kt[48]; int dim; };
template <typename Real> inline LqrP<Real> to_dev(const MdsLqrGains& g) {
  LqrP<Real> d;
  for (int i = 0; i < 48; ++i) { d.K[i] = Real(g.K[i]); d.Kt[i] = Real(0); }
  d.dim = g.dim;
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < g.dim && k < 12; ++k) d.Kt[4 * k + i] = Real(g.K[i * g.dim + k]);
  return d;
}
// Scratch of the large-active-set QP solver (mds_cbf.cuh qp_solve_group_big): `slots` slots of `slot_doubles` doubles in
// global memory + one claim flag each; owned by the library (mds_kernels.cu qp_scratch_for), sized for qmax = 3 N.
struct QpScratch {
  double* base;
  int* flags;
  int slots, qmax;
  long long slot_doubles;
};
template <typename Real> struct CbfP {
  int order, max_iter, state_bounds;
  Real c4inv, inv_c, rs, ds4_pair, k0, k1, k2, umax[4], fmin, fmax;
  Real umax_hi[3];  // |u_c| beyond this violates the box bound by more than the solver's relative tolerance
  QpScratch scr;
};
template <typename Real> inline CbfP<Real> to_dev(const MdsCbfParams& c) {
  CbfP<Real> d;
  d.order = c.order; d.max_iter = c.max_iter > 0 ? c.max_iter : 64;
  double c4 = c.zscale * c.zscale * c.zscale * c.zscale;
  d.c4inv = Real(1.0 / c4); d.inv_c = Real(1.0 / c.zscale); d.rs = Real(c.safety_radius);
  d.ds4_pair = Real(16.0 * c.safety_radius * c.safety_radius * c.safety_radius * c.safety_radius);  // (2 r_safe)^4
  d.state_bounds = c.no_state_bounds ? 0 : 1;
  d.scr = QpScratch{nullptr, nullptr, 0, 0, 0};
  d.k0 = Real(c.kcbf[0]); d.k1 = Real(c.kcbf[1]); d.k2 = Real(c.order == 3 ? c.kcbf[2] : 0.0);
  for (int i = 0; i < 4; ++i) d.umax[i] = Real(c.umax[i]);
  {
    const double tol = sizeof(Real) == 4 ? 2e-6 : 1e-11;  // qp_tol<Real>() of mds_cbf.cuh
    for (int i = 0; i < 3; ++i) d.umax_hi[i] = Real((c.umax[i] * (1.0 + tol) + tol * 1e-12) / (1.0 - tol));
  }
  d.fmin = Real(c.fmin); d.fmax = Real(c.fmax);
  return d;
}

// ---------------------------------------------------------------- per-drone state in registers
template <typename Real> struct Drone {
  V3<Real> p;       // world position
  Real qx, qy, qz, qw;
  V3<Real> v;       // world velocity
  V3<Real> w;       // BODY rates (upstream rpy_rates)
  Real rpm[4];      // last clipped action
};
template <typename Real> struct StateP {
  typename Vec4T<Real>::type *pos_wx, *quat, *vel_wy, *rpm;
  Real* wz;
};
template <typename Real> inline StateP<Real> to_dev(const MdsState& s) {
  using R4 = typename Vec4T<Real>::type;
  return StateP<Real>{(R4*)s.pos_wx, (R4*)s.quat, (R4*)s.vel_wy, (R4*)s.rpm, (Real*)s.wz};
}
template <typename Real> MDS_DEV Drone<Real> load_drone(const StateP<Real>& s, int d) {
  auto a = s.pos_wx[d]; auto q = s.quat[d]; auto c = s.vel_wy[d]; auto r = s.rpm[d];
  Drone<Real> o;
  o.p = {a.x, a.y, a.z}; o.qx = q.x; o.qy = q.y; o.qz = q.z; o.qw = q.w;
  o.v = {c.x, c.y, c.z}; o.w = {a.w, c.w, s.wz[d]};
  o.rpm[0] = r.x; o.rpm[1] = r.y; o.rpm[2] = r.z; o.rpm[3] = r.w;
  return o;
}
template <typename Real> MDS_DEV void store_drone(const StateP<Real>& s, int d, const Drone<Real>& o) {
  typename Vec4T<Real>::type a, q, c, r;
  a.x = o.p.x; a.y = o.p.y; a.z = o.p.z; a.w = o.w.x;
  q.x = o.qx; q.y = o.qy; q.z = o.qz; q.w = o.qw;
  c.x = o.v.x; c.y = o.v.y; c.z = o.v.z; c.w = o.w.y;
  r.x = o.rpm[0]; r.y = o.rpm[1]; r.z = o.rpm[2]; r.w = o.rpm[3];
  s.pos_wx[d] = a; s.quat[d] = q; s.vel_wy[d] = c; s.rpm[d] = r; s.wz[d] = o.w.z;
}

// Bullet getMatrixFromQuaternion (xyzw; self-normalising s = 2/|q|^2) -- SURVEY A.2
template <typename Real> MDS_DEV M3<Real> quat_to_mat(Real x, Real y, Real z, Real w) {
  Real d = x * x + y * y + z * z + w * w;
  Real s = Real(2) * rcp_(d);
  Real xs = x * s, ys = y * s, zs = z * s;
  Real wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
  M3<Real> R;
  R.m[0] = Real(1) - (yy + zz); R.m[1] = xy - wz; R.m[2] = xz + wy;
  R.m[3] = xy + wz; R.m[4] = Real(1) - (xx + zz); R.m[5] = yz - wx;
  R.m[6] = xz - wy; R.m[7] = yz + wx; R.m[8] = Real(1) - (xx + yy);
  return R;
}
// Bullet getEulerFromQuaternion (no normalisation; +-0.99999 gimbal branches)
template <typename Real> MDS_DEV V3<Real> quat_to_rpy(Real x, Real y, Real z, Real w) {
  Real sqx = x * x, sqy = y * y, sqz = z * z, squ = w * w;
  Real sarg = Real(-2) * (x * z - w * y);
  const Real half_pi = Real(1.5707963267948966);
  if (sarg <= Real(-0.99999)) return {Real(0), -half_pi, Real(2) * atan2_(x, -y)};
  if (sarg >= Real(0.99999)) return {Real(0), half_pi, Real(2) * atan2_(-x, y)};
  return {atan2_(Real(2) * (y * z + w * x), squ - sqx - sqy + sqz), asin_(sarg),
          atan2_(Real(2) * (x * y + w * z), squ + sqx - sqy - sqz)};
}

// The 20-float observation of one drone, in registers
template <typename Real> struct Obs {
  V3<Real> p; Real qx, qy, qz, qw; V3<Real> rpy, v, av; Real rpm[4];
};
template <typename Real> MDS_DEV Obs<Real> load_obs(const Real* __restrict__ obs, int d) {
  using R4 = typename Vec4T<Real>::type;
  const R4* o4 = reinterpret_cast<const R4*>(obs + (size_t)d * MDS_OBS_DIM);
  R4 a = o4[0], b = o4[1], c = o4[2], e = o4[3], f = o4[4];
  Obs<Real> o;
  o.p = {a.x, a.y, a.z}; o.qx = a.w; o.qy = b.x; o.qz = b.y; o.qw = b.z;
  o.rpy = {b.w, c.x, c.y}; o.v = {c.z, c.w, e.x}; o.av = {e.y, e.z, e.w};
  o.rpm[0] = f.x; o.rpm[1] = f.y; o.rpm[2] = f.z; o.rpm[3] = f.w;
  return o;
}
MDS_DEV void store_vec4(float4* p, float4 v, bool stream) { if (stream) __stcs(p, v); else *p = v; }
MDS_DEV void store_vec4(double4* p, double4 v, bool stream) {
  if (stream) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); __stcs(reinterpret_cast<double2*>(p) + 1, make_double2(v.z, v.w)); }
  else *p = v;
}
// stream = true: write-once data nobody on this SM reads back (the observation log): streaming stores, so that they
// do not displace the L1 lines the step loop lives on
template <typename Real> MDS_DEV void store_obs(Real* __restrict__ obs, int d, const Obs<Real>& o, bool stream = false) {
  using R4 = typename Vec4T<Real>::type;
  R4* o4 = reinterpret_cast<R4*>(obs + (size_t)d * MDS_OBS_DIM);
  R4 a, b, c, e, f;
  a.x = o.p.x; a.y = o.p.y; a.z = o.p.z; a.w = o.qx;
  b.x = o.qy; b.y = o.qz; b.z = o.qw; b.w = o.rpy.x;
  c.x = o.rpy.y; c.y = o.rpy.z; c.z = o.v.x; c.w = o.v.y;
  e.x = o.v.z; e.y = o.av.x; e.z = o.av.y; e.w = o.av.z;
  f.x = o.rpm[0]; f.y = o.rpm[1]; f.z = o.rpm[2]; f.w = o.rpm[3];
  store_vec4(o4, a, stream); store_vec4(o4 + 1, b, stream); store_vec4(o4 + 2, c, stream); store_vec4(o4 + 3, e, stream); store_vec4(o4 + 4, f, stream);
}
template <typename Real> MDS_DEV Obs<Real> make_obs(const Drone<Real>& s, V3<Real> ang_v_world) {
  Obs<Real> o;
  o.p = s.p; o.qx = s.qx; o.qy = s.qy; o.qz = s.qz; o.qw = s.qw;
  o.rpy = quat_to_rpy(s.qx, s.qy, s.qz, s.qw);
  o.v = s.v; o.av = ang_v_world;
#pragma unroll
  for (int i = 0; i < 4; ++i) o.rpm[i] = s.rpm[i];
  return o;
}

}  // namespace mds
