#!/usr/bin/env python
"""Headline benchmark: drone-steps/s of the full hot path on synthetic swarms (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's algorithms on the host cores

Workload (config.workload): SURVEY.md 8(d) C5 -- per GPU 125 000 independent environments x 8 drones
(1M drones), 240 Hz; every drone-step = Lemniscate reference -> LQR-yank-omega nominal -> order-3
CBF-QP (28 pair rows + 8 obstacle rows + box/force bounds per env) -> YankOmega inner loop ->
DYN_GND_DRAG_DW physics (ground effect, drag, pairwise downwash) -> 20-float observation.

One bench "step" = one ``mds_rollout`` call = ONE kernel launch = ``--fuse`` control steps of every drone with the
observation and body rates in registers from step to step (environments are independent; a lane group owns its env);
every step's observation is also written to HBM (obs log), as the reference's loop keeps every obs.
``value``   : drone-steps/s, state resident in HBM, CUDA events on the launching stream, max over ranks.
``e2e``     : the same metric through the per-call API with HOST buffers (multidronesim_b200.HostPipeline): every
              control step copies the step's references host -> device from pinned memory and the new observation
              device -> host, copies overlapped with the neighbouring steps' kernels on their own streams.
``roofline``: the dominant (only) kernel of the timed region, rollout_loop_kernel.  It is FP32-pipe-bound (the state
              never leaves the registers: ~20 flop per HBM byte with the per-step obs log, ridge 9.8), so achieved = algorithmic flop per launch / launch
              duration against the FMA-chain peak measured live (MEASURED_PEAKS.json has no non-tensor FP32 figure);
              its HBM view is reported inside.  ``roofline_ctrl`` / ``roofline_physics``: the per-call kernels (what
              ``e2e`` launches) against the measured HBM copy bandwidth, from a launch-by-launch replay.
              No tensor cores on this path: nothing is a dense contraction.
``cpu_baseline``: oracle/ (numpy restatement of the reference's algorithms) on all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

# one BLAS thread per worker process: the CPU arm parallelises over processes (one env per core)
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_DRONES = 8
SWARM_OBSTACLES = [[0.2, 0.0, 0.5, 0.1]]  # == multidronesim_b200.scenarios.SWARM_OBSTACLES (checked in run_gpu_arm)
CBF_ORDER = 3
# Algorithmic work per drone-step of the C5 path (DESIGN.md "Roofline"): FP32 operations of the closed-form math
# (add / mul / compare / min-max = 1, fma = 2, transcendental / div / sqrt = 1), hand-counted per stage at the lap-average
# 1.5 QP iterations per solve.  Cross-check (profiles/r1_ncu_kernels.txt): ncu's executed fadd + fmul + 2 ffma per
# drone-step in a 1.0-iteration phase is 1423 for the rollout kernel (1016 controller stack + 419 physics); the
# table adds the compares / min-max / MUFU ops those counters leave out and the extra half iteration.
ALGO_FLOP_PER_DRONE_STEP = {"traj": 70, "lqr_yank": 120, "cbf_rows_and_first_scan": 620, "qp_iterations_and_rescans": 230, "lowlevel": 150,
                            "physics_gnd_drag_dw_n8": 480, "obs": 60}
# Algorithmic HBM bytes per drone-step (fp32), per kernel of the rollout (DESIGN.md "Roofline"):
ALGO_BYTES_CTRL_F32 = {"read_obs": 80, "read_traj_spec": 48, "read_pid": 24, "write_pid": 24, "write_action": 16}
ALGO_BYTES_PHYS_F32 = {"read_state": 68, "read_action": 16, "write_state": 68, "write_obs": 80}  # 232 B
# the K-steps-in-one-launch kernel, per drone and per LAUNCH (not per step): initial obs + body rates + trajectory spec,
# PID state in and out, final state + obs + action
ALGO_BYTES_OBS_LOG_PER_STEP_F32 = 80  # the observation of every control step, written to its log slot
ALGO_BYTES_LOOP_PER_LAUNCH_F32 = {"read_obs": 80, "read_body_rates": 12, "read_traj_spec": 48, "read_pid": 24, "write_pid": 24,
                                  "write_state": 68, "write_obs": 80, "write_action": 16}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=125000, help="environments per GPU (weak scaling)")
    ap.add_argument("--fuse", type=int, default=24, help="control steps per fused launch (one bench step)")
    ap.add_argument("--settle", type=int, default=3024, help="untimed control steps that take the swarm from its start at rest into its steady "
                                                               "lap before warm-up, in both arms (default: one lemniscate period); the start-up "
                                                               "transient has heavy-tailed QP iteration counts")
    ap.add_argument("--e2e-steps", type=int, default=48, help="control steps timed on the host-buffer path")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="wall budget of the cpu_baseline sample")
    ap.add_argument("--ref-steps-per-step", type=int, default=48, help="control steps per reference-arm step and worker")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle's reference-style loop (numpy, per-drone Python loops) on the host cores
# ----------------------------------------------------------------------------------------------
def _oracle_env(env_index, seed=3):
    import numpy as np
    from oracle import trajectories as otj
    from oracle.aviary import OracleCtrlAviary
    from oracle.constants import DroneModel as ODM, Physics as OPH
    N = N_DRONES
    phase = (2 * np.pi / (N + 0.25)) * np.arange(N)
    specs = [dict(a=1.0, center=np.array([0, 0, 0.5]), omega=0.5, yaw_rate=0.0, phase_shift=float(p)) for p in phase]
    rng = np.random.default_rng(np.random.SeedSequence([seed, 10_000_000 + env_index]))
    init = np.array([otj.Lemniscate(**sp)(0.0)[0] for sp in specs]) + rng.normal(0, 0.02, (N, 3))
    init[:, 2] += 0.04 * np.arange(N)
    env = OracleCtrlAviary(ODM.CF2P, N, initial_xyzs=init, physics=OPH.DYN_GND_DRAG_DW)
    return env, [otj.Lemniscate(**sp) for sp in specs]


_WORKER = {}


def _cpu_worker_step(args):
    """advance this worker's private environment by `steps` control steps; returns drone-steps done"""
    from oracle import pipeline as opl
    env_index, steps = args
    key = env_index
    if key not in _WORKER:
        env, trajs = _oracle_env(env_index)
        _WORKER[key] = dict(env=env, trajs=trajs, ctrls=opl.make_controllers(env, "yank10"), t=0.0)
    w = _WORKER[key]
    opl.run_cbf(w["env"], w["trajs"], CBF_ORDER, steps, obstacles=SWARM_OBSTACLES, ctrls=w["ctrls"], t0=w["t"], log=False)
    w["t"] += steps * w["env"].CTRL_TIMESTEP
    return N_DRONES * steps


def cpu_baseline(seconds, steps_per_task=24, settle=3024):
    """bounded sample: every host core advances its own C5 environment in 24-step tasks for ~`seconds`"""
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker_step, [(i, max(2, settle)) for i in range(cores)], chunksize=1)  # imports, gains, settle into steady flight
        done, t0 = 0, time.perf_counter()
        rounds = 0
        while time.perf_counter() - t0 < seconds:
            done += sum(pool.map(_cpu_worker_step, [(i, steps_per_task) for i in range(cores)], chunksize=1))
            rounds += 1
        dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "drone-steps/s", "cores": cores, "kind": "port",
            "sample": f"{cores} C5 environments x {N_DRONES} drones (one per core), {rounds * steps_per_task} control steps each "
                      f"after {settle} settling steps, oracle/pipeline.run_cbf (reference algorithms in numpy; cvxopt -> oracle active-set QP, "
                      f"PyBullet env -> restated DYN_GND_DRAG_DW step), {dt:.1f} s wall"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    S = args.ref_steps_per_step
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker_step, [(i, max(1, args.settle)) for i in range(cores)], chunksize=1)  # settle into steady flight
        for _ in range(max(1, args.warmup)):
            pool.map(_cpu_worker_step, [(i, 4) for i in range(cores)], chunksize=1)
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            done += sum(pool.map(_cpu_worker_step, [(i, S) for i in range(cores)], chunksize=1))
        dt = time.perf_counter() - t0
    value = done / dt
    sample = (f"each step: {cores} C5 environments x {N_DRONES} drones (one per host core) advance {S} control steps; "
              f"oracle/pipeline.run_cbf = the reference's algorithms restated in numpy (reference is pure Python and cannot "
              f"travel to the GPU box; cvxopt and gym-pybullet-drones are not installable offline)")
    line = {"impl": "reference", "metric": "drone-steps/sec (DYN_GND_DRAG_DW + LQR + order-3 CBF-QP, 8 drones/env)", "value": value,
            "unit": "drone-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cores * 1, per_gpu=False),
            "cpu_baseline": {"value": value, "unit": "drone-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "drone-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(args, envs, per_gpu=True):
    return {"workload": "C5 swarm sweep: envs x 8 drones, Physics.DYN_GND_DRAG_DW 240 Hz, Lemniscate refs, LQR-yank-omega nominal, "
                        "order-3 CBF-QP (r_safe 0.125, zscale 2, poles -3/-3.6/-5.6) + sphere obstacle r=0.1 at (0.2, 0, 0.5), YankOmega inner loop",
            "envs_per_gpu" if per_gpu else "envs": envs, "drones_per_env": N_DRONES, "drone_model": "cf2p", "cbf_order": CBF_ORDER,
            "control_steps_per_step": args.fuse if per_gpu else args.ref_steps_per_step, "settle_steps": args.settle, "parallelism": f"env-sharded x{args.gpus}",
            "obs": "every control step's observation is written to HBM (log ring of control_steps_per_step slots)",
            "l2": "working set (state+obs+traj specs+PID+obs log > 2 GB per GPU) exceeds the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        """clocks over the samples that arrived inside [t0, t1] (the timed region); if the region is too short for two
        samples, over everything since the sampler started (settle + warm-up + timed: the same kernels, same load)"""
        rows, window = self.rows, "all samples"
        if t0 is not None:
            inside = [r for r in self.rows if t0 <= r[0] <= t1 + 0.12]
            if len(inside) >= 2:
                rows, window = inside, "timed region"
            else:
                window = "settle + warm-up + timed region (timed region shorter than two sampling periods)"
        sm, mx, reasons = [], [], set()
        for _, r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return None


def snapshot(env, ctrl, ro):
    return (env.state_dict(), ctrl.low_level._a.clone(), ctrl.low_level._b.clone(), ro.t, ro.stats.clone())


def restore(env, ctrl, ro, snap):
    env.load_state_dict(snap[0])
    ctrl.low_level._a.copy_(snap[1])
    ctrl.low_level._b.copy_(snap[2])
    ro.t = snap[3]
    ro.stats.copy_(snap[4])


def ncu_traffic(kernel, envs, dtype):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/ncu_traffic.json: written by tools/ncu_traffic.py from the .ncu-rep), or None if no capture
    of this kernel at this size exists."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except (OSError, ValueError):
        return None
    for row in tab.get("kernels", []):
        if row["kernel"].startswith(kernel) and row["envs"] == envs and row["dtype"] == dtype:
            return row["dram_bytes_per_launch"]
    return None


def run_gpu_arm(args):
    cpu = None
    rank_env = int(os.environ.get("RANK", "0"))
    if rank_env == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.cpu_seconds, settle=args.settle)  # before CUDA is initialised in this process (fork-safe)

    import torch
    import torch.distributed as dist

    import multidronesim_b200 as mds
    from multidronesim_b200 import scenarios

    assert scenarios.SWARM_OBSTACLES == SWARM_OBSTACLES
    rank, local_rank, world = mds.dist.init_from_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    esz = 4 if dtype == torch.float32 else 8
    E, N, F, K, W = args.envs, N_DRONES, args.fuse, args.steps, max(3, args.warmup)
    D = E * N

    sc = scenarios.cbf_swarm(E, N, order=CBF_ORDER, dtype=dtype, device=dev, env_offset=rank * E)
    env, ro, ctrl = sc["env"], sc["rollout"], sc["ctrl"]
    fma_peak = mds._lib.fma_peak_tflops(use_f64=(dtype == torch.float64))

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- device-resident headline -------------------------------------------------------------
    # every control step's observation is materialised in HBM (a ring of F log slots, reused by each launch), as the
    # reference's loop does (observations.append(obs)); a drone-step therefore includes its 80 B observation write
    obs_ring = torch.empty(F, E, N, 20, device=dev, dtype=dtype)
    step = lambda: ro.run(F, obs_log=obs_ring, log_every=1)
    with ClockSampler(local_rank) as clk:  # sampling starts with the settle phase: nvidia-smi needs ~0.2 s to deliver its first line
        for _ in range(args.settle // F):
            step()
        for _ in range(W):
            step()
        torch.cuda.synchronize()
        ro.reset_stats()
        snap = snapshot(env, ctrl, ro)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        t_wall0 = time.perf_counter()
        ev0.record()
        for _ in range(K):
            step()
        ev1.record()
        torch.cuda.synchronize()
        t_wall1 = time.perf_counter()
        barrier()
        time.sleep(0.12)  # let the sample that covers the end of the region arrive
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * D * F * K / (ms_max * 1e-3)
    stats_all = mds.dist.gather_stats(ro.stats)  # the path's only collective (NCCL all-gather of 8 doubles per rank)
    stats = dict(zip(mds._lib.STAT_NAMES, mds.dist.reduce_stats(stats_all).tolist()))
    clocks = clk.summary(t_wall0, t_wall1)

    # ---- roofline of the dominant kernel: the K-steps-in-one-launch rollout kernel (this rank) ----------------
    # Every bench step is ONE launch of rollout_loop_kernel (F control steps with the state in registers), so its
    # mean launch duration is the timed region / K, on the launching stream.  It moves ~15 B per drone-step through
    # HBM for ~1.73 kflop: FP32-pipe-bound (SURVEY.md 8d "fused K-step rollout"), peak = FMA-chain microbenchmark.
    if ro.plan() != 6:
        raise SystemExit("bench.py assumes the K-steps-in-one-launch plan")
    n_steps = K * F
    step_ms = ms / n_steps
    launch_ms = ms / K
    peaks = measured_peaks()
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
    peak_src = "MEASURED_PEAKS.json hbm_gbs, of measured (sustained copy)" if peaks else "fallback 6.65 TB/s, of fallback"
    sfx = "float" if dtype == torch.float32 else "double"
    flop_step = sum(ALGO_FLOP_PER_DRONE_STEP.values())
    bytes_launch = (sum(ALGO_BYTES_LOOP_PER_LAUNCH_F32.values()) + F * ALGO_BYTES_OBS_LOG_PER_STEP_F32) * esz // 4
    tf = flop_step * D * F / (launch_ms * 1e-3) / 1e12
    gbs = bytes_launch * D / (launch_ms * 1e-3) / 1e9
    roofline = {"bound": "fp32", "kernel": f"rollout_loop_kernel<{sfx}, MDS_CTRL_LQR_YANK, true, 8>", "achieved": tf, "peak": fma_peak,
                "unit": "TFLOP/s", "frac": tf / fma_peak,
                "peak_source": "measured live: mds_fma_peak dependent-FMA chains on all SMs (no non-tensor FP32 figure in MEASURED_PEAKS.json)",
                "algorithmic_flop_per_drone_step": flop_step, "control_steps_per_launch": F, "launch_ms": launch_ms, "share_of_step": 1.0,
                "traffic": ncu_traffic("rollout_loop_kernel", E, args.dtype),
                "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "algorithmic_bytes_per_drone_per_launch": bytes_launch,
                        "peak_source": peak_src},
                "note": f"arithmetic intensity {flop_step * F / bytes_launch:.0f} flop/B > ridge {fma_peak * 1e3 / hbm_peak:.1f} flop/B: the state lives in "
                        "registers for the whole launch and HBM sees the per-step observation log plus one state load/store per launch, so the "
                        "FP32 pipe (not HBM, not tensor cores: nothing is a dense contraction) bounds it"}

    # ---- the per-call kernels against the HBM roofline: replay with one launch per kernel ----------------------
    # (MdsRolloutCfg.stages 1 = controller kernel, 2 = physics kernel; a CUDA event pair around every launch)
    restore(env, ctrl, ro, snap)
    torch.cuda.synchronize()
    n_rep = min(n_steps, 240)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_rep)]
    for k in range(n_rep):
        evs[k][0].record()
        ro.run(1, stages=1)
        evs[k][1].record()
        ro.run(1, stages=2)
        evs[k][2].record()
    torch.cuda.synchronize()
    ctrl_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / n_rep
    phys_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / n_rep
    b_ctrl, b_phys = sum(ALGO_BYTES_CTRL_F32.values()) * esz // 4, sum(ALGO_BYTES_PHYS_F32.values()) * esz // 4

    def roof(kernel, bytes_unit, launch):
        g = bytes_unit * D / (launch * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kernel, "achieved": g, "peak": hbm_peak, "unit": "GB/s", "frac": g / hbm_peak, "peak_source": peak_src,
                "algorithmic_bytes_per_drone_step": bytes_unit, "launch_ms": launch, "traffic": ncu_traffic(kernel.split("<")[0], E, args.dtype),
                "note": "per-call path (two launches per control step); replay of the first %d control steps of the timed region" % n_rep}

    roofline_ctrl = roof(f"ctrl_step_kernel<{sfx}, MDS_CTRL_LQR_YANK, true, 8>", b_ctrl, ctrl_ms)
    roofline_physics = roof(f"physics_step_kernel<{sfx}, 8>", b_phys, phys_ms)

    # ---- end to end through the per-call API with host buffers --------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, mds, sc, dev, dtype, world, barrier)

    if rank == 0:
        line = {"metric": "drone-steps/sec (DYN_GND_DRAG_DW + LQR + order-3 CBF-QP, 8 drones/env)", "value": value, "unit": "drone-steps/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, E),
                "clocks": clocks, "e2e": e2e, "gpu_launches": K, "roofline": roofline, "roofline_ctrl": roofline_ctrl,
                "roofline_physics": roofline_physics, "cpu_baseline": cpu, "rollout_stats": stats, "sm_count": mds._lib.device_info()["sm_count"]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, mds, sc, dev, dtype, world, barrier):
    """Reference-facing per-call sequence with HOST buffers (multidronesim_b200.HostPipeline), every control step:
       H2D refs (what the reference passes to set_desired_trajectory) -> LQR (skip_low_level) -> caller glue
       (nominal -= mg, xdes) -> CBF-QP -> inner loop -> env.step -> D2H observations; the copies of neighbouring
       steps overlap the kernels on their own streams."""
    import torch
    import torch.distributed as dist
    env, ctrl, trk, trajs = sc["env"], sc["ctrl"], sc["tracker"], sc["trajs"]
    E, N = env.NUM_ENVS, env.NUM_DRONES
    D = E * N
    S = args.e2e_steps
    # continues from the swarm's current (steady-flight) state; the host owns the references (pre-evaluated for the
    # S steps, as a host-side planner would) and receives the observations
    t_start = sc["rollout"].t
    ref_host = torch.empty(S + 4, D, 11, dtype=dtype).pin_memory()
    for k in range(S + 4):
        ref_host[k].copy_(trajs.eval(t_start + k * env.CTRL_TIMESTEP))
    torch.cuda.synchronize()
    obs_host = [torch.empty(E, N, 20, dtype=dtype).pin_memory() for _ in range(2)]
    pipe = mds.HostPipeline(env, ctrl, trk, sc["obstacles"])
    for k in range(4):
        pipe.step(ref_host[k], obs_host[k & 1])
    pipe.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    cs = torch.cuda.current_stream(dev)
    ev0.record()
    for k in range(S):
        pipe.step(ref_host[4 + k], obs_host[k & 1])
    cs.wait_stream(pipe.s_out)  # the last observation has landed in host memory
    cs.wait_stream(pipe.s_in)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    esz = 4 if dtype == torch.float32 else 8
    return {"value": world * D * S / (ms * 1e-3), "unit": "drone-steps/s", "h2d_bytes_per_step": D * 11 * esz, "d2h_bytes_per_step": D * 20 * esz,
            "control_steps": S, "ms_per_control_step": ms / S, "kernels_per_control_step": pipe.launches_per_step,
            "path": "pinned host refs -> H2D (copy stream) -> mds_lqr_ctrl -> mds_cbf_prepare -> mds_cbf_qp -> mds_lowlevel -> mds_physics_step "
                    "-> D2H obs to pinned host (copy stream); two slots, copies overlap the next step's kernels"}


_JSON_FD = None


def emit(line):
    """the ONE JSON line goes to the process's original stdout; everything else (NCCL banners, warnings) to stderr"""
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    global _JSON_FD
    args = parse_args()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)  # libraries that print to stdout (e.g. "NCCL version ...") must not pollute the JSON line
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
