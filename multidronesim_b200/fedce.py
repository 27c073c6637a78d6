"""Federated certainty-equivalence learning wrapper of the reference (FedCE/FederatedLearning.py:9-67), batched on device.

The reference feeds it lists of ``YOState`` objects (one per drone: roll, pitch, yaw, thrust, velocity, position) that come
from outside the simulator; here a state is a row ``[r, p, y, T, vx, vy, vz, px, py, pz]`` of a device tensor [E, N, 10]
(``YOState.get_state_vec`` order, control/dlqr/decentralized_yolqr_crazyflie.py:105-116).  ``update`` forms
phi = [e_t, u_t] and e_{t+1} against the PREVIOUS desired state (FederatedLearning.py:25-48) with ``mds_state_feedback`` and
runs one ``approx_theta_update`` (``mds_rls_update``); ``lqr_control`` is u = -K_d e_d (``mds_state_feedback``)."""
from __future__ import annotations

import torch

from . import _lib
from .control.dlqr import DecentralizedYOLQRCrazyflie


class FederatedLearning:
    def __init__(self, env, lin_models, Q, R, P=None, num_drones=1):
        if num_drones != env.NUM_DRONES:
            raise ValueError("num_drones must equal the env's drones per environment")
        self.num_drones, self.env, self.linear_models = num_drones, env, lin_models
        self.m, self.n = 10, 4
        self.dLQR = DecentralizedYOLQRCrazyflie(env, lin_models, Q, R, P=P)
        self.x_prev = self.x_des_prev = self.u_prev = None
        D = env.NUM_TOTAL
        kw = dict(device=env.device, dtype=env.dtype)
        self._e_t, self._e_tp1 = torch.zeros(D, self.m, **kw), torch.zeros(D, self.m, **kw)
        self._phi = torch.zeros(D, self.m + self.n, **kw)
        self._u = torch.zeros(D, 4, **kw)

    def make_desired_state(self, pos=None, vel=None, yaw=0.0):
        """[E, N, 10] desired states with thrust = m g, the equilibrium (FederatedLearning.py:21-23); pos / vel [E,N,3] or [3]."""
        env = self.env
        x = torch.zeros(env.NUM_ENVS, env.NUM_DRONES, self.m, device=env.device, dtype=env.dtype)
        x[..., 2] = yaw
        x[..., 3] = env.M * env.G
        if vel is not None:
            x[..., 4:7] = torch.as_tensor(vel, device=env.device, dtype=env.dtype)
        if pos is not None:
            x[..., 7:10] = torch.as_tensor(pos, device=env.device, dtype=env.dtype)
        return x

    def _states(self, t, name, width):
        return _lib.require_cuda(t, name, self.env.dtype).reshape(self.env.NUM_TOTAL, width)

    def error_state(self, x, x_des, out=None):
        env = self.env
        out = torch.zeros(env.NUM_TOTAL, self.m, device=env.device, dtype=env.dtype) if out is None else out
        _lib.call("mds_state_feedback", env.dtype, _lib.CTRL_LQR_YANK, None, _lib.ptr(self._states(x, "x", self.m)),
                  _lib.ptr(self._states(x_des, "x_des", self.m)), _lib.ptr(out), None, env.NUM_TOTAL, _lib.stream_ptr(env.device))
        return out

    def update(self, xtp1, x_des, u):
        """One learning step from the new states; the first call only stores them (returns (None, None))."""
        if self.x_prev is None:
            self.x_prev, self.x_des_prev, self.u_prev = xtp1.clone(), x_des.clone(), u.clone()
            return None, None
        self.error_state(xtp1, self.x_des_prev, self._e_tp1)       # against the PREVIOUS desired state (:37)
        self.error_state(self.x_prev, self.x_des_prev, self._e_t)
        self._phi[:, :self.m] = self._e_t
        self._phi[:, self.m:] = self._states(self.u_prev, "u", self.n)
        self.x_des_prev, self.x_prev, self.u_prev = x_des.clone(), xtp1.clone(), u.clone()
        self.dLQR.approx_theta_update(self._phi, self._e_tp1, project=True)
        return self._phi, self._e_tp1

    def lqr_control(self, x, x_des):
        """u [E, N, 4] = -K_d e_d (DecentralizedYOLQRCrazyflie.compute, :350-362: no hover offset, yank input)."""
        env = self.env
        _lib.call("mds_state_feedback", env.dtype, _lib.CTRL_LQR_YANK, _lib.ptr(self.dLQR.K_planes), _lib.ptr(self._states(x, "x", self.m)),
                  _lib.ptr(self._states(x_des, "x_des", self.m)), None, _lib.ptr(self._u), env.NUM_TOTAL, _lib.stream_ptr(env.device))
        return self._u.view(env.NUM_ENVS, env.NUM_DRONES, 4)

    def calc_controller(self):
        self.dLQR.compute_controller()

    def theta_str(self, env_idx=0):
        import numpy as np
        out = []
        for i in range(self.num_drones):
            th = self.dLQR.get_thetai(i, env_idx).cpu().numpy()
            out.append("Theta A (robot %d):\n%s\nTheta B:\n%s" % (i, np.array_str(th[:self.m].T, precision=3, suppress_small=True, max_line_width=100000),
                                                                   np.array_str(th[self.m:].T, precision=3, suppress_small=True, max_line_width=100000)))
        return "\n".join(out)

    def save_theta(self, filename="theta.npy"):
        import numpy as np
        np.save(filename, self.dLQR.theta.cpu().numpy())
