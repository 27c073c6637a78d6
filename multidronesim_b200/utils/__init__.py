from .model_conversions import (obs_to_lin_model, obs_to_geo_model, action_to_input, input_to_action, calc_z_thrust,
                                geo_x_dot_to_linear)

__all__ = ["obs_to_lin_model", "obs_to_geo_model", "action_to_input", "input_to_action", "calc_z_thrust", "geo_x_dot_to_linear"]
