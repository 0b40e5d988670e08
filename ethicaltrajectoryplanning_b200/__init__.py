"""B200-native candidate-trajectory scoring path of TUMFTM/EthicalTrajectoryPlanning.

Host side: Python, same call surface as the reference's Frenet planner step; device side:
hand-written sm_100a CUDA kernels behind the C ABI of ``include/etp_b200.h``.
"""
from . import configs, spline, synthetic, vehicleparams  # noqa: F401

__all__ = ["configs", "spline", "synthetic", "vehicleparams"]
