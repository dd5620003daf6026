"""juliafem.jl_b200 -- B200-native replacement for JuliaFEM's 3D-elasticity
assemble -> K.u -> CG hot path.  Device code lives in csrc/ behind the C ABI of
include/jfem_b200.h; this package is the Python host mirror of the reference's interface."""
from . import mesh  # noqa: F401
