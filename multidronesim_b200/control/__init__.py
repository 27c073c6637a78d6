from .base_controller import BaseController
from .dslpid import DSLPIDControl
from .geometric import GeometricControl
from .low_level import ThrustOmegaController, YankOmegaController
from .lqr import LQRController, LQROmegaController, LQRYankOmegaController
from .dlqr import DecentralizedLQR, DecentralizedLQROmega, DecentralizedLQRYankOmega, DecentralizedYOLQRCrazyflie
from . import dlqr

__all__ = ["BaseController", "GeometricControl", "DSLPIDControl", "ThrustOmegaController", "YankOmegaController",
           "LQRController", "LQROmegaController", "LQRYankOmegaController",
           "DecentralizedLQR", "DecentralizedLQROmega", "DecentralizedLQRYankOmega", "DecentralizedYOLQRCrazyflie"]
