"""orbithunter_b200: B200-native engine for orbithunter's spatiotemporal Kuramoto-Sivashinsky hot path.

Drop-in for that path only: the OrbitKS-family classes and ``hunt`` keep the reference's names and semantics
(orbithunter/ks/orbits.py, orbithunter/optimize.py); ``BatchedOrbitKS`` evaluates many independent orbits per launch.
All numerics run in hand-written sm_100a CUDA kernels behind the C ABI in include/ks_b200.h; there is no CPU path.
"""
from ._lib import KsB200Error  # noqa: F401
from .batch import BatchedOrbitKS  # noqa: F401
from .ks import (AntisymmetricOrbitKS, OrbitKS, RelativeOrbitKS, ShiftReflectionOrbitKS)  # noqa: F401

__all__ = ["OrbitKS", "RelativeOrbitKS", "ShiftReflectionOrbitKS", "AntisymmetricOrbitKS", "BatchedOrbitKS", "KsB200Error"]
