"""B200-native constitutive + heat-source update of solidStateAMFoam's chtMultiRegionInterFoam.

Package contents: csrc/ (hand-written sm_100a kernels + the C-ABI of include/ssam_gpu.h),
capi.py (ctypes binding = what an FFI host binds), params.py (POD mirrors), mesh.py + meshgen/
(synthetic OpenFOAM-convention meshes), cases.py (BASELINE.json's configs), host/ (C++ mirror of the
reference's run-time-selectable interfaces).  The oracle lives outside the package, in oracle/.
"""
from . import params  # noqa: F401

__all__ = ["params"]
__version__ = "0.1.0"
