"""gogp_b200 -- B200-native GP hot path behind GoGP's gp.GP API.

``gp``      host-side mirror of gp.GP / gp.Model (reference gp/gp.go, gp/model.go)
``kernel``  stock kernels and their compositions, lowered to the device descriptor
``_lib``    ctypes binding of the C-ABI (include/gogp_b200.h)
"""
from . import _lib, gp, kernel  # noqa: F401
from .gp import GP, Model, GoGPError, GoGPPanic  # noqa: F401
