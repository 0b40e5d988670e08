"""Mirror of the reference's ``risk_assessment`` package: same entry points, CUDA underneath."""
