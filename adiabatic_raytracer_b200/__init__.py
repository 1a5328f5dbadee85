"""adiabatic_raytracer_b200 — B200-native (CUDA sm_100a, FP64) hot path of SamWitte/Adiabatic_RayTracer.

Only what the path needs: the CUDA library + C ABI (csrc/, include/art.h), its ctypes binding
(lib.py) and the host-side mirror of the reference entry point (main_runner.py).
Importing the package does not load the library; the first use does, and raises if it is missing.
"""
from .params import ArtParams, ArtStats  # noqa: F401

__all__ = ["ArtParams", "ArtStats", "main_runner_tree", "main_runner_tree_multi", "Handle"]


def __getattr__(name):
    if name == "main_runner_tree":
        from .main_runner import main_runner_tree
        return main_runner_tree
    if name == "main_runner_tree_multi":
        from .main_runner import main_runner_tree_multi
        return main_runner_tree_multi
    if name == "Handle":
        from .lib import Handle
        return Handle
    raise AttributeError(name)
