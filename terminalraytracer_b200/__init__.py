"""terminalraytracer_b200 — B200-native (sm_100a) implementation of TerminalRayTracer's per-pixel
render path behind the reference's own C entry points.  See DESIGN.md."""
from . import abi, sharding  # noqa: F401

__all__ = ["abi", "sharding"]
