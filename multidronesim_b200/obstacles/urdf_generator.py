"""Obstacle primitives.  The reference writes a sphere URDF for PyBullet's renderer/collider
(obstacles/urdf_generator.py:4-30); the CBF only ever consumes (centre, radius)
(cbf/cbf.py:380-383), which is what ``Sphere`` carries to the device."""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass(frozen=True)
class Sphere:
    center: tuple
    radius: float

    def as_row(self):
        return [float(self.center[0]), float(self.center[1]), float(self.center[2]), float(self.radius)]


_TEMPLATE = """<?xml version="1.0"?>
<robot name="sphere_obstacle">
  <link name="base_link">
    <inertial><origin xyz="0 0 0"/><mass value="0"/><inertia ixx="0" ixy="0" ixz="0" iyy="0" iyz="0" izz="0"/></inertial>
    <visual><origin xyz="0 0 0"/><geometry><sphere radius="{r}"/></geometry><material name="red"><color rgba="1 0 0 1"/></material></visual>
    <collision><origin xyz="0 0 0"/><geometry><sphere radius="{r}"/></geometry></collision>
  </link>
</robot>
"""


def generate_sphere(radius, folder="/tmp"):
    """Write ``sphere_<radius>.urdf`` and return its path (same contract as the reference)."""
    path = os.path.join(folder, f"sphere_{radius}.urdf")
    with open(path, "w") as f:
        f.write(_TEMPLATE.format(r=radius))
    return path
