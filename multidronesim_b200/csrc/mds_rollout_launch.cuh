// mds_rollout_launch.cuh -- host-side launchers of the three heavy rollout kernels.  Declared here, defined in
// mds_rollout_tu.cu, which the Makefile compiles once per (precision, kernel kind): the ~120 instantiations build in
// parallel instead of inside one translation unit.
#pragma once
#include <cuda_runtime.h>
#include "mds_rollout.cuh"

template <typename Real> struct RolloutLaunch {
  DroneP<Real> Pd;
  RolloutP<Real> R;
  GeoP<Real> G;
  LqrP<Real> L;
  CbfP<Real> C;
  DslP<Real> Dg;
  DslStateP<Real> Ds;
  StateP<Real> Sd;
  PidP<Real> Pi;
  const typename TrajSpecT<Real>::spec* specs;
  const typename TrajSpecT<Real>::seg* segs;
  Real* action;
  const Real* fext;
  Real* obs;
  Real* obs_log;
  double* stats;
  int E, N, NP, blocks, threads;
  int spec;  // PhysSpec instantiation the parameter block qualifies for (0 = none: run-time switches)
  size_t smem;
  cudaStream_t cs;
};

// Each returns the error of the shared-memory opt-in (cudaSuccess when it was not needed); launch errors are
// picked up by the caller's cudaGetLastError().
template <typename Real> cudaError_t launch_loop_kernel(const RolloutLaunch<Real>& a, double t0, double dt_ctrl, int K);
template <typename Real> cudaError_t launch_ctrl_kernel(const RolloutLaunch<Real>& a, double t, const Real* obs_in, bool set_attr);
template <typename Real> cudaError_t launch_fused_kernel(const RolloutLaunch<Real>& a, double t, Real* obs_out, bool set_attr);
// the work-queue form of the K-step rollout (rollout_queue_kernel); q: device queue already zeroed on a.cs, tiles / chunks set by the caller.
// *grid_blocks (in: 0 = ask) receives / provides the persistent grid size
template <typename Real> cudaError_t launch_queue_kernel(const RolloutLaunch<Real>& a, double t0, double dt_ctrl, int K, RolloutQueue q, int sm_count, bool query_only,
                                                         int* blocks_per_sm);
