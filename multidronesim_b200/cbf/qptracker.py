"""CBF-QP safety filter (reference cbf/qptracker.py:13-34,86-114) solved per environment on
device: min |u - u_nom|^2 s.t. G u <= h with a dual active-set method; envs whose QP is
infeasible or hits the iteration cap get the nominal input back, as the reference's
except-branch does, and are flagged in ``status``."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def pack_obstacles(x_obs, obs_r_list, device, dtype):
    """Reference obstacle lists (x_obs [N_obs, order, 3] with row 0 = centre, radii list;
    simulations/CBFTest.py:421-424) -> device [N_obs, 4] = cx, cy, cz, r.  A negative radius marks a vertical
    cylinder of radius |r| (obstacles.Cylinder.as_row(); builder extension)."""
    if x_obs is None or obs_r_list is None:
        return None
    assert len(x_obs) == len(obs_r_list), \
        "The lists for Obstacle positions and radii must have the same length. Right now {} & {}".format(len(x_obs), len(obs_r_list))
    if len(x_obs) == 0:
        return None
    rows = []
    for xo, r in zip(x_obs, obs_r_list):
        c = np.asarray(xo, dtype=float).reshape(-1, 3)[0]
        rows.append([c[0], c[1], c[2], float(r)])
    return torch.tensor(rows, device=device, dtype=dtype)


class DroneQPTracker(object):
    def __init__(self, cbf, order=2, num_robots=1, xdim=9, env=None):
        self.cbf, self.order, self.num_robots, self.xdim = cbf, order, num_robots, xdim
        self.env = env if env is not None else cbf.env
        if cbf.order != order or cbf.xdim != xdim:
            raise AssertionError("Require that qptracker and cbf have the same order / state dimension.")
        if num_robots != self.env.NUM_DRONES:
            raise ValueError("num_robots must equal the env's drones per environment")
        E, N = self.env.NUM_ENVS, self.env.NUM_DRONES
        kw = dict(device=self.env.device)
        self.u_safe = torch.zeros(E, N, 4, dtype=self.env.dtype, **kw)
        self.status = torch.zeros(E, dtype=torch.int32, **kw)
        self.iters = torch.zeros(E, dtype=torch.int32, **kw)
        self.qp_tracker = self
        self._obst_key, self._obst = None, None

    def compute_control(self, obs, xdes, u_nominal, ignore_zmin=False, x_obs=None, obs_r_list=None):
        """obs [E,N,20], xdes [E,N,xdim], u_nominal [E,N,4] (device) -> u_safe [E,N,4] (library-owned).
        ``self.status`` [E] int32: 0 optimal, 1 infeasible -> nominal used, 2 iteration cap -> nominal used."""
        env = self.env
        E, N = env.NUM_ENVS, env.NUM_DRONES
        if isinstance(x_obs, torch.Tensor):
            obst = x_obs  # already packed [N_obs, 4]
        else:
            key = (id(x_obs), id(obs_r_list))
            if key != self._obst_key:
                self._obst_key, self._obst = key, pack_obstacles(x_obs, obs_r_list, env.device, env.dtype)
            obst = self._obst
        n_obs = 0 if obst is None else int(obst.shape[0])
        self.cbf.check_obstacle_count(n_obs)
        _lib.call("mds_cbf_qp", env.dtype, env._prm, self.cbf.c_params(),
                  _lib.ptr(_lib.require_cuda(obs, "obs", env.dtype, (E, N, _lib.OBS_DIM))),
                  _lib.ptr(_lib.require_cuda(xdes, "xdes", env.dtype, (E, N, self.xdim))),
                  _lib.ptr(_lib.require_cuda(u_nominal, "u_nominal", env.dtype, (E, N, 4))),
                  _lib.ptr(obst), n_obs, _lib.ptr(self.u_safe), _lib.ptr(self.status), _lib.ptr(self.iters), E, N,
                  _lib.stream_ptr(env.device))
        return self.u_safe


QPTracker = DroneQPTracker
