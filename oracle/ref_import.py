"""Import the UNMODIFIED reference torch backend from /root/reference (build container only).

The reference's projection/rasterization modules execute `from max.torch import
CustomOpLibrary` and `CustomOpLibrary(<dir>)` at import time (mojosplat/projection.py:9-13,
rasterization.py:5-10).  MAX is not installed, so a no-op stub is injected into
sys.modules first; the torch code paths never touch it.  /root/reference does not exist
on the GPU box: nothing under tests -m gpu / smoke() / bench.py may import this module.
TEST INFRASTRUCTURE ONLY.
"""
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "mojosplat" / "binning.py").exists()


def load():
    """-> the reference `mojosplat` package (projection, binning, utils importable)."""
    if not available():
        raise RuntimeError("reference checkout not present (expected on the build container only)")
    if "max" not in sys.modules:
        m = types.ModuleType("max")
        mt = types.ModuleType("max.torch")

        class CustomOpLibrary:  # noqa: D401 - stub
            def __init__(self, *a, **k):
                pass

        mt.CustomOpLibrary = CustomOpLibrary
        m.torch = mt
        sys.modules["max"] = m
        sys.modules["max.torch"] = mt
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import mojosplat  # noqa: F401
    import mojosplat.binning
    import mojosplat.projection
    import mojosplat.utils
    return sys.modules["mojosplat"]
