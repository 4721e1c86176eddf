"""gogp_b200 -- B200-native GP hot path behind GoGP's gp.GP API.

``gp``      host-side mirror of gp.GP / gp.Model (reference gp/gp.go, gp/model.go)
``grid``    gp.GP over a pr x pc grid of GPUs (2D block-cyclic K, NCCL inside the library)
``tutorial`` tutorial.Evaluate's expanding window and CSV formats (reference tutorial/tutorial.go)
``kernel``  stock kernels and their compositions, lowered to the device descriptor
``_lib``    ctypes binding of the C-ABI (include/gogp_b200.h)
"""
from . import _lib, gp, grid, kernel, tutorial  # noqa: F401
from .gp import GP, Model, GoGPError, GoGPPanic  # noqa: F401
from .grid import GridGP  # noqa: F401
