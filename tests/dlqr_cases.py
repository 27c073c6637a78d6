"""The reference's model-learning entry points exercised by tests/golden/dlqr.npz (oracle/make_golden.py:golden_dlqr):
tag -> (m, target, predict_from_xtp1, normalize_gain, project mode, product class name, product method, kwargs)."""
CASES = {
    "omega9_update": (9, 0, True, True, 0, "DecentralizedLQROmega", "theta_update", {}),
    "omega9_update2": (9, 0, True, False, 0, "DecentralizedLQROmega", "theta_update2", {}),   # golden P is V; state is V^-1
    "yank10_update": (10, 0, True, True, 0, "DecentralizedLQRYankOmega", "theta_update", {}),
    "torque12_update": (12, 0, False, True, 1, "DecentralizedLQR", "theta_update", {}),
    "torque12_approx": (12, 1, False, True, 2, "DecentralizedLQR", "approx_theta_update", {}),
    "cf10_approx": (10, 1, False, True, 2, "DecentralizedYOLQRCrazyflie", "approx_theta_update", {"project": True}),
    "cf10_approx_noproj": (10, 1, False, True, 0, "DecentralizedYOLQRCrazyflie", "approx_theta_update", {"project": False}),
}
