"""quanto_b200 -- B200-native (sm_100a) kernels behind optimum-quanto's quantized-linear operator surface.

Importing the package registers the `torch.ops.quanto.*` CUDA implementations (library.py) and exposes the host-side
mirror of the reference classes on the hot path.  The native library is loaded lazily at the first op call and
its absence is an error (no CPU / eager fallback).
"""
from . import library  # noqa: F401  (op registration side effect)
from .calibrate import *  # noqa: F401,F403
from .nn import *  # noqa: F401,F403
from .pipeline import *  # noqa: F401,F403
from .tensor import *  # noqa: F401,F403

__version__ = "0.1.0"
