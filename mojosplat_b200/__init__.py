"""mojosplat_b200 -- B200-native (sm_100a) backend of the MojoSplat forward path.

Public surface = the reference's (mojosplat/render.py, projection.py, binning.py,
rasterization.py, utils.py) plus the additive batched / multi-GPU entry points.
Importing the package needs neither a GPU nor the built library; the first call that
selects the CUDA backend loads ``csrc/libbsplat.so`` and fails loudly if it is missing.
"""
from .utils import Camera  # noqa: F401
