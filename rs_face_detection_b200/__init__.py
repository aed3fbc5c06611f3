"""rs_face_detection_b200 — B200-native detection hot path for okieraised/rs-face-detection.

The product is libfd_b200.so (hand-written CUDA for sm_100a, C ABI in include/fd_b200.h).  This package is the
host-side mirror of the reference's operator interface over that ABI (same names and argument meaning as the Rust
modules `processing`, `rcnn` and `pipeline::module`), used by the tests and the benchmark.  No CPU fallback exists.
"""
from .ffi import Context, FdError, default_config, device_count, load  # noqa: F401

__all__ = ["Context", "FdError", "default_config", "device_count", "load", "default_context"]

_default_ctx = {}


def default_context(device=0):
    """Lazily created per-device context used by the free functions in `processing` / `rcnn`."""
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
