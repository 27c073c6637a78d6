from .cbf import CBF, DroneCBF
from .qptracker import DroneQPTracker, QPTracker

__all__ = ["CBF", "DroneCBF", "DroneQPTracker", "QPTracker"]
