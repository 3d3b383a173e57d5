"""Device engine of the discrete-GP hot path (ctypes over libtal_b200.so)."""
from . import _abi  # noqa: F401
from .engine import DeviceGP  # noqa: F401
