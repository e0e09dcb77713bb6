"""b200sd — sm_100a compute path of the local-GPU worker."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
