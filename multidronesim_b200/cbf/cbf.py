"""Exponential-CBF safety-filter parameters (reference cbf/cbf.py:27-125,540-580).

``DroneCBF`` keeps the reference's constructor and the quantities it precomputes (Kcbf by
pole placement on the integrator chain, umax, force bounds); the per-step row building and
the QP run on device (``mds_cbf_qp``; dense G, h via ``mds_cbf_rows`` for inspection)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def place_poles_chain(poles):
    """Gain of scipy.signal.place_poles(F, G, poles) for the single-input integrator chain built at
    cbf/cbf.py:119-124: det(sI - (F - G K)) = prod(s - p_k)  =>  K = reversed polynomial coefficients."""
    c = np.poly(np.asarray(poles, dtype=float))
    return np.real(c[1:][::-1]).reshape(1, -1).copy()


class CBF:
    def __init__(self, xy_only, zscale=2.0, order=3, umax=None, safety_radius=1.0, cbf_poles=np.array([-2.2, -2.4, -2.6]),
                 room_bounds=np.array([-4.25, 4.5, -3.5, 4.25, 1.0, 2.0]), Fmax=3, Fmin=-3,
                 RPY_max=np.array([np.pi / 6, np.pi / 6, 2 * np.pi]), vmax=np.array([2, 2, 2]), A=None, B=None,
                 do_state_bounds=True, num_agents=1):
        if order not in (2, 3):
            raise AssertionError("device CBF supports relative degree 2 (thrust-omega) and 3 (yank-omega)")
        if room_bounds is not None:
            assert len(room_bounds) == 6, "Require xmin, xmax, ymin, ymax, zmin, zmax (6 values), but len(room_bounds)={}".format(len(room_bounds))
        assert len(cbf_poles) == order, "Number of specified CBF poles ({})does not match order ({})".format(len(cbf_poles), order)
        if xy_only:
            raise _lib.MdsError("xy_only CBFs are not supported on device")
        self.xdim = A.shape[0] // num_agents
        if self.xdim != (9 if order == 2 else 10):
            raise AssertionError("order 2 needs the 9-dim omega model, order 3 the 10-dim yank-omega model")
        self.zscale, self.order, self.umax = zscale, order, umax
        self.safety_radius, self.cbf_poles, self.xy_only, self.dim = safety_radius, np.asarray(cbf_poles, dtype=float), xy_only, 3
        self.num_agents = num_agents
        self.Fmax, self.Fmin, self.Vmax, self.RPY_max = Fmax, Fmin, vmax, RPY_max
        self.do_state_bounds = do_state_bounds
        self.A, self.B = A, B
        self.Kcbf = place_poles_chain(self.cbf_poles)
        self.max_iter = 64

    def c_params(self):
        c = _lib.CbfParams()
        c.order, c.zscale, c.safety_radius = self.order, float(self.zscale), float(self.safety_radius)
        for i in range(self.order):
            c.kcbf[i] = float(self.Kcbf[0, i])
        um = np.broadcast_to(np.asarray(self.umax, dtype=float), (4,))
        for i in range(4):
            c.umax[i] = float(um[i])
        c.fmin, c.fmax, c.max_iter = float(self.Fmin), float(self.Fmax), int(self.max_iter)
        c.no_state_bounds = 0 if self.do_state_bounds else 1
        return c

    def update_cbf_gain(self, cbf_poles):
        cbf_poles = np.asarray(cbf_poles, dtype=float)
        if len(cbf_poles) != self.order or len(set(cbf_poles)) != len(cbf_poles) or not (cbf_poles < 0).all():
            return False
        self.cbf_poles, self.Kcbf = cbf_poles, place_poles_chain(cbf_poles)
        return True

    def update_zscale(self, zscale):
        if not zscale > 0:
            return False
        self.zscale = zscale
        return True

    def update_safety_radius(self, safety_radius):
        if not safety_radius > 0:
            return False
        self.safety_radius = safety_radius
        return True


class DroneCBF(CBF):
    def __init__(self, env, lin_models, zscale=2.0, safety_radius=1, cbf_poles=np.array([-2.2, -2.4]), room_bounds=None,
                 omega_max=np.array([10, 10, 10]), order=2, allow_extra_obstacles=False):
        self.num_agents = len(lin_models)
        # The reference fails with more obstacles than drones (cbf.py:388 indexes agent blocks by obstacle id, quirk B14);
        # True lifts that: up to MAX_OBSTACLES for any N (builder extension).
        self.allow_extra_obstacles = bool(allow_extra_obstacles)
        if self.num_agents != env.NUM_DRONES:
            raise ValueError("one linear model per drone of an environment is required")
        self.xdim = lin_models[0].A.shape[0]
        n, N = self.xdim, self.num_agents
        A, B = np.zeros((n * N, n * N)), np.zeros((n * N, 4 * N))
        for i, mdl in enumerate(lin_models):
            A[n * i:n * (i + 1), n * i:n * (i + 1)] = mdl.A
            B[n * i:n * (i + 1), 4 * i:4 * (i + 1)] = mdl.B
            if not (np.array_equal(mdl.A, lin_models[0].A) and np.array_equal(mdl.B, lin_models[0].B)):
                raise _lib.MdsError("device CBF rows assume identical hover-linearised models for all drones")
        Fmin, Fmax = -env.M * env.G, env.MAX_THRUST
        Ymax = (env.MAX_THRUST / env.CTRL_TIMESTEP) / 100
        umax = np.array([env.MAX_THRUST if order == 2 else Ymax, omega_max[0], omega_max[1], omega_max[2]], dtype=float)
        super().__init__(xy_only=False, zscale=zscale, order=order, umax=umax, safety_radius=safety_radius, cbf_poles=cbf_poles,
                         room_bounds=room_bounds, Fmin=Fmin, Fmax=Fmax, RPY_max=np.array([np.pi / 6, np.pi / 6, 2 * np.pi]),
                         A=A, B=B, num_agents=N)
        self.lin_models, self.env = lin_models, env

    def check_obstacle_count(self, n_obs):
        if n_obs > self.num_agents and not getattr(self, "allow_extra_obstacles", False):
            raise IndexError("more obstacles than drones: the reference's builder fails here (cbf/cbf.py:388); "
                             "pass DroneCBF(..., allow_extra_obstacles=True) for the extension")

    def num_rows(self, n_obs=0):
        m = _lib.load_library().mds_cbf_num_rows(self.order, self.num_agents, n_obs)
        if self.order == 3 and not self.do_state_bounds:   # no force-bound rows (cbf/cbf.py:473-476)
            m -= 2 * self.num_agents
        return m

    def build_ineq_const(self, obs, xdes, obstacles=None):
        """Dense (G [E,m,4N], h [E,m]) in the reference's row order (cbf/cbf.py:308-367) for inspection /
        parity tests; the solver itself never materialises G."""
        env = self.env
        E, N = env.NUM_ENVS, env.NUM_DRONES
        n_obs = 0 if obstacles is None else obstacles.shape[0]
        self.check_obstacle_count(n_obs)
        m = self.num_rows(n_obs)
        G = torch.empty(E, m, 4 * N, device=env.device, dtype=env.dtype)
        h = torch.empty(E, m, device=env.device, dtype=env.dtype)
        _lib.call("mds_cbf_rows", env.dtype, env._prm, self.c_params(), _lib.ptr(obs), _lib.ptr(xdes), _lib.ptr(obstacles), n_obs,
                  _lib.ptr(G), _lib.ptr(h), E, N, _lib.stream_ptr(env.device))
        return G, h
