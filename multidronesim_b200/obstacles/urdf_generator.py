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


@dataclass(frozen=True)
class Cylinder:
    """Vertical cylinder through (center[0], center[1]) of unbounded height for the CBF (builder extension: the
    reference has spheres only); ``height`` is used by the URDF (render / collision shape) alone."""
    center: tuple
    radius: float
    height: float = 2.0

    def as_row(self):
        return [float(self.center[0]), float(self.center[1]), float(self.center[2]), -float(self.radius)]


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


def generate_cylinder(radius, height=2.0, folder="/tmp"):
    """Write ``cylinder_<radius>_<height>.urdf`` (z-axis cylinder) and return its path."""
    path = os.path.join(folder, f"cylinder_{radius}_{height}.urdf")
    geom = f'<cylinder radius="{radius}" length="{height}"/>'
    with open(path, "w") as f:
        f.write(_TEMPLATE.replace("sphere_obstacle", "cylinder_obstacle").replace('<sphere radius="{r}"/>', geom))
    return path
