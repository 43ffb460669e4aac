"""graphik_b200: B200-native batched distance-geometric inverse kinematics.

Drop-in for the Riemannian IK path of utiasSTARS/GraphIK
(`ProblemGraph` / `RiemannianSolver` / `solve_with_riemannian`); the hot path
runs as hand-written sm_100a CUDA kernels behind the C ABI declared in
include/graphik_b200.h.  See DESIGN.md.
"""
__version__ = "0.1.0"
