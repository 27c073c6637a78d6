"""String-valued enums with the upstream gym-pybullet-drones values the reference
passes as ``type=DroneModel`` / ``type=Physics`` to argparse
(reference simulations/EnvGeometric.py:9,21-22,39-44)."""
from enum import Enum


class DroneModel(Enum):
    CF2X = "cf2x"
    CF2P = "cf2p"


class Physics(Enum):
    DYN = "dyn"                          # upstream explicit dynamics
    DYN_GND_DRAG_DW = "dyn_gnd_drag_dw"  # composite defined in SURVEY.md App. A.4
