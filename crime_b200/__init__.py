"""crime_b200 -- B200-native GetHI hot path (damonge/CRIME), host-side bindings.

The compute lives in crime_b200/csrc/libgh_cuda.so (hand-written sm_100a CUDA behind the C-ABI of
include/gh_cuda.h).  This package only binds it; importing it does not load the library.
"""
from .abi import GhCudaParams, params_from_dict  # noqa: F401
from .gethi import GetHI, GetHIError, params_from_tables  # noqa: F401
