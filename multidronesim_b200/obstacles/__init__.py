from .urdf_generator import Cylinder, Sphere, generate_cylinder, generate_sphere

__all__ = ["generate_sphere", "generate_cylinder", "Sphere", "Cylinder"]
