"""graph_pooling_b200 -- B200-native (sm_100a) DiffPool forward/backward hot path.

Drop-in for the reference's ``encoders.GcnEncoderGraph`` / ``encoders.SoftPoolingGcnEncoder``
(/root/reference/encoders.py:976-1334).  All compute lives in libgp_b200.so (C ABI in
include/gp_b200.h); this package is the Python host side that mirrors the reference interface.
"""
from . import _lib  # noqa: F401
from .encoders import GcnEncoderGraph, GcnSet2SetEncoder, GraphConv, SoftPoolingGcnEncoder  # noqa: F401

__version__ = '0.1.0'
