"""densehead -- B200-native dense-head path (target encode, losses, decode + NMS) behind the
reference's Python entry points.

    from densehead import fcos, retinanet, centernet

Each module mirrors the reference module it replaces (same function names, argument meaning and error
behaviour -- see the docstrings for file:line) and adds a `*_batch` variant that takes a padded batch and
keeps everything on the device.  All arithmetic runs in libdensehead.so (hand-written CUDA for sm_100a,
`include/densehead.h`); there is no CPU fallback.
"""
from . import _capi
from ._capi import DenseHeadError, launch_count, raise_for_status, set_option, status  # noqa: F401
from . import fcos, retinanet, centernet, prep, distributed  # noqa: F401

__all__ = ["fcos", "retinanet", "centernet", "prep", "distributed", "DenseHeadError", "launch_count", "set_option", "version"]


def version():
    return _capi.lib().dh_version().decode()
