"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/bsplat.h declares.
No compute call is made (no GPU here)."""
import re
from pathlib import Path

import pytest

import mojosplat_b200 as ms
from mojosplat_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    if not _lib.LIB_PATH.exists():
        _lib.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    header = (ROOT / "include" / "bsplat.h").read_text()
    declared = set(re.findall(r"\b(bsplat_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bsplat_version() == 1
    assert b"workspace" in lib.bsplat_error_string(-2)


def test_struct_sizes_match_header():
    import ctypes
    assert ctypes.sizeof(_lib.BsplatCamera) == 16 * 4 + 4 * 4 + 2 * 4 + 2 * 4
    assert ctypes.sizeof(_lib.BsplatBinInfo) == 32
    assert ctypes.sizeof(_lib.BsplatKeyLayout) == 12


def test_workspace_queries_need_no_gpu(lib):
    assert lib.bsplat_bin_scan_workspace_bytes(1_000_000) >= (1_000_000 // 2048) * 8
    small = lib.bsplat_render_workspace_bytes(1000, 5000, 256, 256, 16)
    big = lib.bsplat_render_workspace_bytes(1000, 5_000_000, 256, 256, 16)
    assert 0 < small < big
    assert lib.bsplat_radix_sort_workspace_bytes(0, 0, 40) > 0


def test_key_layout_host_helper(lib):
    import ctypes
    info = _lib.BsplatBinInfo()
    info.n_isect = 10
    info.min_depth_key = 0x80000000 | 0x3DCCCCCD  # 0.1
    info.max_depth_key = 0x80000000 | 0x42C80000  # 100.0
    lay = lib.bsplat_make_key_layout(ctypes.byref(info), 1920, 1080, 16)
    assert lay.tile_bits == 13 and lay.depth_bits == 27 and lay.depth_bias == info.min_depth_key
    lay = lib.bsplat_make_key_layout(ctypes.byref(info), 256, 256, 16)
    assert lay.tile_bits == 8


def test_no_cpu_fallback_and_api_surface():
    import torch
    z = torch.zeros
    cam = ms.Camera(R=torch.eye(3), T=z(3), H=8, W=8, fx=1.0, fy=1.0, cx=4.0, cy=4.0)
    assert cam.view_matrix.shape == (4, 4) and cam.Ks.shape == (3, 3) and ms.TILE_SIZE == 16
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ms.project_gaussians(z(1, 3), z(1, 3), z(1, 4), z(1, 1), cam, backend="cuda")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ms.bin_gaussians_to_tiles(z(1, 2), z(1, 2), z(1), 8, 8, 16, backend="cuda")
    for fn, args in ((ms.project_gaussians, (z(1, 3), z(1, 3), z(1, 4), z(1, 1), cam)),
                     (ms.bin_gaussians_to_tiles, (z(1, 2), z(1, 2), z(1), 8, 8, 16))):
        with pytest.raises(ValueError, match="Invalid backend"):
            fn(*args, backend="bogus")
    with pytest.raises(ValueError, match="CUDA tensors"):
        ms.render_gaussians(z(1, 3), z(1, 3), z(1, 4), z(1), z(1, 3), cam)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (parity claims would be void)."""
    for p in (ROOT / "mojosplat_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
    for p in (ROOT / "mojosplat_b200" / "csrc").glob("*.cu*"):
        assert "oracle" not in p.read_text(), p


def test_camera_struct_reads_intrinsics_from_Ks():
    """Like the reference (projection.py:137-140, 156-159), the intrinsics come from camera.Ks."""
    import pytest
    import torch

    import mojosplat_b200 as ms
    from mojosplat_b200 import _lib
    cam = ms.Camera(R=torch.eye(3), T=torch.zeros(3), H=48, W=64, fx=100.0, fy=100.0, cx=32.0, cy=24.0)
    c = _lib.camera_struct(cam)
    assert (c.fx, c.fy, c.cx, c.cy, c.width, c.height) == (100.0, 100.0, 32.0, 24.0, 64, 48)
    cam2 = ms.Camera(R=torch.eye(3), T=torch.zeros(3), H=48, W=64, fx=1.0, fy=1.0, cx=0.0, cy=0.0,
                     Ks=torch.tensor([[120.0, 0.0, 30.0], [0.0, 110.0, 20.0], [0.0, 0.0, 1.0]]))
    c2 = _lib.camera_struct(cam2)
    assert (c2.fx, c2.fy, c2.cx, c2.cy) == (120.0, 110.0, 30.0, 20.0)
    cam3 = ms.Camera(R=torch.eye(3), T=torch.zeros(3), H=48, W=64, fx=1.0, fy=1.0, cx=0.0, cy=0.0,
                     Ks=torch.tensor([[120.0, 0.5, 30.0], [0.0, 110.0, 20.0], [0.0, 0.0, 1.0]]))
    with pytest.raises(ValueError, match="pinhole"):
        _lib.camera_struct(cam3)
