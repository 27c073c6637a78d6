"""Analytic trajectory generators -- TEST INFRASTRUCTURE ONLY.

numpy fp64 restatement of the reference's trajectories/*.py (each class cites
the lines it follows); pinned against the imported reference by
tests/test_oracle_vs_reference.py and tests/golden/trajectories.npz.
Every generator maps ``t -> (pos3, vel3, acc3, yaw, yaw_rate)``.
"""
from __future__ import annotations

import math

import numpy as np

TWO_PI = 2.0 * math.pi


class Circle:
    """trajectories/Circle.py:5-45.  Quirk B18: yaw wrapped into [pi, 3pi)."""

    def __init__(self, r=1.0, v=0.5, center=(0, 0, 0), yaw_rate=0.0, revolutions=None, duration=None):
        self.r, self.v, self.yaw_rate = float(r), float(v), float(yaw_rate)
        self.center = np.asarray(center, dtype=float)
        if revolutions is not None:
            self.total_time = TWO_PI * r * revolutions / self.v
        elif duration is not None:
            self.total_time = duration
        else:
            self.total_time = TWO_PI * self.r / self.v

    def get_total_time(self):
        return self.total_time

    def __call__(self, t):
        w = self.v / self.r
        c, s = math.cos(w * t), math.sin(w * t)
        cen = self.v ** 2 / self.r
        pos = self.center + np.array([self.r * c, self.r * s, 0.0])
        vel = np.array([-self.v * s, self.v * c, 0.0])
        acc = np.array([-cen * c, -cen * s, 0.0])
        yaw = (self.yaw_rate * t - math.pi) % TWO_PI + math.pi
        return pos, vel, acc, yaw, self.yaw_rate


class Lemniscate:
    """trajectories/Lemniscate.py:3-63 (Bernoulli lemniscate; quirk B19 yaw law)."""

    def __init__(self, a=1, omega=0.5, center=(0, 0, 0), yaw_rate=0, revolutions=None, duration=None, phase_shift=0):
        self.a, self.omega, self.yaw_rate, self.phase_shift = a, omega, yaw_rate, phase_shift
        self.center = np.asarray(center, dtype=float)
        if revolutions is not None:
            self.total_time = TWO_PI * revolutions / omega
        elif duration is not None:
            self.total_time = duration
        else:
            self.total_time = TWO_PI / omega

    def get_total_time(self):
        return self.total_time

    def __call__(self, t):
        a, om = self.a, self.omega
        th = t * om + self.phase_shift
        s, c = math.sin(th), math.cos(th)
        c2, c4 = math.cos(2 * th), math.cos(4 * th)
        den = 1 + s ** 2
        pos = self.center + np.array([a * s * c / den, a * c / den, 0.0])
        vel = np.array([-a * om * (s ** 4 + s ** 2 + (s ** 2 - 1) * c ** 2) / den ** 2,
                        -a * om * s * (s ** 2 + 2 * c ** 2 + 1) / den ** 2, 0.0])
        acc = np.array([4 * a * om ** 2 * math.sin(2 * th) * (3 * c2 + 7) / (c2 - 3) ** 3,
                        a * om ** 2 * c * (44 * c2 + c4 - 21) / (c2 - 3) ** 3, 0.0])
        yaw = math.pi * math.sin(self.yaw_rate * t)
        yaw_rate = math.pi * self.yaw_rate * math.cos(self.yaw_rate * t)
        return pos, vel, acc, yaw, yaw_rate


class Wait:
    """trajectories/LineTrajectory.py:4-14."""

    def __init__(self, position, duration, yaw=0):
        self.position, self.duration, self.yaw = np.asarray(position, dtype=float), duration, yaw

    def get_total_time(self):
        return self.duration

    def __call__(self, t):
        return self.position, np.zeros(3), np.zeros(3), self.yaw, 0


class Line:
    """trajectories/LineTrajectory.py:16-103 -- trapezoidal speed profile, a_max = 1.
    Quirk B20: ``speed`` is mandatory, ``dist_end`` uses |v0|, per-axis sign() acceleration."""

    def __init__(self, start, end, speed=None, duration=None, s0=0, sf=0):
        if not speed > 0:
            raise AssertionError("Speed must be positive")
        if duration is not None and not duration > 0:
            raise AssertionError("Duration must be positive")
        self.start, self.end = np.asarray(start, dtype=float), np.asarray(end, dtype=float)
        delta = self.end - self.start
        dist = float(np.linalg.norm(delta))
        self.a = 1.0
        self.speed = speed
        self.dir = delta / dist
        self.v0, self.vf = s0 * self.dir, sf * self.dir
        self._ramps()
        if self.dist_init + self.dist_end > dist:
            self.time_middle = 0
            self.speed = sf + math.sqrt(dist * self.a) + 0.5 * s0 ** 2 - 0.5 * sf ** 2
            self._ramps()
        else:
            self.time_middle = (dist - self.dist_init - self.dist_end) / self.speed
        self.total_time = self.time_init + self.time_middle + self.time_end

    def _ramps(self):
        self.dv_init = self.speed * self.dir - self.v0
        self.dv_end = self.vf - self.speed * self.dir
        self.time_init = float(np.linalg.norm(self.dv_init)) / self.a
        self.time_end = float(np.linalg.norm(self.dv_end)) / self.a
        n0 = float(np.linalg.norm(self.v0))
        self.dist_init = n0 * self.time_init + 0.5 * self.a * self.time_init ** 2
        self.dist_end = n0 * self.time_end + 0.5 * self.a * self.time_end ** 2

    def get_total_time(self):
        return self.total_time

    def __call__(self, t):
        if t > self.total_time:
            return self.end, self.vf, np.zeros(3), 0, 0
        sg_i, sg_e = np.sign(self.dv_init), np.sign(self.dv_end)
        cruise = self.speed * self.dir
        if t < self.time_init:
            return (self.start + self.v0 * t + 0.5 * sg_i * self.a * t ** 2,
                    self.v0 + sg_i * self.a * t, sg_i * self.a, 0, 0)
        d_init = self.v0 * self.time_init + 0.5 * sg_i * self.a * self.time_init ** 2
        if t < self.time_init + self.time_middle:
            tau = t - self.time_init
            return self.start + d_init + cruise * tau, cruise, np.zeros(3), 0, 0
        tau = t - self.time_middle - self.time_init
        d_mid = d_init + cruise * self.time_middle
        return (self.start + d_mid + cruise * tau + 0.5 * sg_e * self.a * tau ** 2,
                cruise + sg_e * self.a * tau, sg_e * self.a, 0, 0)


class Compound:
    """trajectories/CompoundTrajectory.py:5-40 -- piecewise dispatcher with the
    reference's forward-only cursor that resets at the end / on time reversal (B21)."""

    def __init__(self, trajectories):
        self.trajectories = list(trajectories)
        durs = [tr.get_total_time() for tr in self.trajectories]
        self.total_time = sum(durs)
        self.times = np.cumsum(durs)
        self.reset()

    def get_total_time(self):
        return self.total_time

    def reset(self):
        self.idx, self.t0 = 0, 0

    def __call__(self, t):
        if t >= self.total_time:
            self.reset()
            last = self.trajectories[-1]
            return last(last.get_total_time())
        while t > self.times[self.idx]:
            self.t0 = self.times[self.idx]
            self.idx += 1
        if self.idx >= 1 and t < self.times[self.idx - 1]:
            self.reset()
            return self(t)
        return self.trajectories[self.idx](t - self.t0)


class Rotate:
    """trajectories/RotateTrajectory.py:5-24 -- rotate pos about ``center``, vel/acc by R."""

    def __init__(self, trajectory, R, center):
        self.trajectory, self.R, self.center = trajectory, np.asarray(R, dtype=float), np.asarray(center, dtype=float)

    def get_total_time(self):
        return self.trajectory.get_total_time()

    def __call__(self, t):
        pos, vel, acc, yaw, om = self.trajectory(t)
        return self.R @ (pos - self.center) + self.center, self.R @ vel, self.R @ acc, yaw, om
