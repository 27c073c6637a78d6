from .urdf_generator import generate_sphere, Sphere

__all__ = ["generate_sphere", "Sphere"]
