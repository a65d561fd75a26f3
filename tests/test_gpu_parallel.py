"""GPU: the sharded paths give the single-GPU result bit for bit (bands emulated on one device, as the
profiling guide asks when fewer GPUs than ranks are available)."""
import numpy as np
import pytest
import torch

import mojosplat_b200 as ms
from helpers import scene_on
from mojosplat_b200 import parallel, synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,N,world", [("config2_100k_1080p", 20_000, 4), ("config3_1m_1080p", 300_000, 8),
                                         ("config5_6m_4k", 200_000, 3)])
def test_row_bands_reassemble_bit_identical(cuda_device, cfg, N, world):
    sc = synthetic.make_scene(cfg, N=N)
    (m, s, q, o, c), cam = scene_on(sc, cuda_device)
    bg = sc.background.to(cuda_device)
    full = ms.render_fused(m, s, q, o, c, cam, bg)
    proj = ms.projection.project_gaussians_cuda(m, s, q, o, cam)
    cost = parallel.tile_row_cost(proj[0], proj[3], cam.H, cam.W, 16)
    bands = parallel.balanced_row_bands(cost.tolist(), world)
    assert bands[0][0] == 0 and bands[-1][1] == (cam.H + 15) // 16
    image = torch.full((cam.H, cam.W, 3), -1.0, device=cuda_device)
    for band in bands:  # "ranks" one after the other on one GPU, each writing only its rows
        parallel.render_band(m, s, q, o, c, cam, bg, band, out=image, projected=proj)
    assert torch.equal(image, full)
    # balance: no band carries more than ~2x the mean cost
    sums = [float(cost[b0:b1].sum()) for b0, b1 in bands]
    assert max(sums) <= 2.0 * (sum(sums) / world) + float(cost.max())
    # the single-process entry point
    assert torch.equal(parallel.render_frame_row_split(m, s, q, o, c, cam, bg), full)


def test_batched_views_equal_single_renders(cuda_device):
    sc = synthetic.make_scene("config3_1m_1080p", N=100_000)
    (m, s, q, o, c), _ = scene_on(sc, cuda_device)
    cams = synthetic.orbit_cameras(4, 480, 270, 250.0)
    bg = sc.background.to(cuda_device)
    imgs = parallel.render_gaussians_batched(m, s, q, o, c, cams, bg)
    assert imgs.shape == (4, 270, 480, 3)
    for k, cam in enumerate(cams):
        assert torch.equal(imgs[k], ms.render_gaussians(m, s, q, o, c, cam, background_color=bg))
    ids, mine = parallel.render_views(m, s, q, o, c, cams, bg)  # world size 1: everything is mine
    assert ids == [0, 1, 2, 3] and torch.equal(mine, imgs)
    assert not torch.equal(imgs[0], imgs[1])


def test_pipelined_views_equal_sequential(cuda_device):
    from mojosplat_b200.pipeline import FramePipeline
    sc = synthetic.make_scene("config3_1m_1080p", N=200_000)
    (m, s, q, o, c), _ = scene_on(sc, cuda_device)
    cams = synthetic.orbit_cameras(7, 960, 540, 500.0)
    bg = sc.background.to(cuda_device)
    # tiny initial capacity: exercises the grow-and-redo path as well
    pipe = FramePipeline(cuda_device, sc.N, 960, 540, m_capacity=1000)
    imgs = pipe.render(m, s, q, o, c, cams, bg)
    torch.cuda.synchronize()
    for k, cam in enumerate(cams):
        assert torch.equal(imgs[k], ms.render_fused(m, s, q, o, c, cam, bg)), k
    # second run with warm workspaces, ring output of 2 slots
    ring = torch.empty((2, 540, 960, 3), device=cuda_device)
    pipe.render(m, s, q, o, c, cams, bg, out=ring)
    torch.cuda.synchronize()
    assert torch.equal(ring[0], imgs[6]) and torch.equal(ring[1], imgs[5])


def test_overlapped_pipeline_equals_sequential(cuda_device):
    """Sync-free frames (device-side M, binning on a high-priority stream, rasterizer on a second one)
    give exactly the frame-by-frame images; a too-small capacity is detected and redone."""
    from mojosplat_b200.pipeline import OverlappedPipeline
    sc = synthetic.make_scene("config3_1m_1080p", N=200_000)
    (m, s, q, o, c), _ = scene_on(sc, cuda_device)
    cams = synthetic.orbit_cameras(9, 960, 540, 500.0)
    bg = sc.background.to(cuda_device)
    pipe = OverlappedPipeline(cuda_device, sc.N, 960, 540)
    imgs = pipe.render(m, s, q, o, c, cams, bg)
    assert pipe.check() == 0
    ref = [ms.render_gaussians(m, s, q, o, c, cam, background_color=bg) for cam in cams]
    for k in range(len(cams)):
        assert torch.equal(imgs[k], ref[k]), k
    assert pipe.last_M > 0
    # capacity far too small: every frame overflows, is flagged on the device and re-rendered
    small = OverlappedPipeline(cuda_device, sc.N, 960, 540, m_capacity=5000)
    imgs2 = small.render(m, s, q, o, c, cams[:3], bg)
    assert small.check() == 3
    for k in range(3):
        assert torch.equal(imgs2[k], ref[k]), k
    # second batch runs with the adapted capacity, nothing to redo
    imgs3 = small.render(m, s, q, o, c, cams[3:6], bg)
    assert small.check() == 0
    for k in range(3):
        assert torch.equal(imgs3[k], ref[3 + k]), k


def test_enqueue_empty_scene_gives_zero_image(cuda_device):
    """gsplat rules + everything behind the camera: M == 0 is only known on the device; the image must
    be all zeros (render.py:73-76), not the background."""
    from mojosplat_b200 import _lib
    from mojosplat_b200.pipeline import OverlappedPipeline
    sc = synthetic.make_scene("config1_1k_256")
    (m, s, q, o, c), cam = scene_on(sc, cuda_device)
    m = m.clone(); m[:, 2] += 1000.0  # far beyond the far plane for this pose
    pipe = OverlappedPipeline(cuda_device, sc.N, cam.W, cam.H, semantics=_lib.SEM_GSPLAT)
    img = pipe.render(m, s, q, o, c, [cam], sc.background.to(cuda_device))
    assert pipe.check() == 0
    if pipe.last_M == 0:
        assert float(img.abs().max()) == 0.0
    else:
        pytest.skip("scene still produced intersections")


def test_graph_renderer_replays_equal_direct_renders(cuda_device):
    """One captured frame (CUDA graph, indirect camera) replayed for several poses == direct renders."""
    from mojosplat_b200.pipeline import GraphRenderer
    sc = synthetic.make_scene("config3_1m_1080p", N=150_000)
    (m, s, q, o, c), _ = scene_on(sc, cuda_device)
    cams = synthetic.orbit_cameras(5, 640, 360, 330.0)
    bg = sc.background.to(cuda_device)
    gr = GraphRenderer(m, s, q, o, c, cams[0], bg)
    for cam in cams:
        img = gr.render(cam).clone()
        assert gr.check() > 0
        assert torch.equal(img, ms.render_gaussians(m, s, q, o, c, cam, background_color=bg))
    # new Gaussian values in the same buffers
    c2 = c.flip(0).contiguous()
    gr.update_gaussians(features=c2)
    assert torch.equal(gr.render(cams[1]), ms.render_gaussians(m, s, q, o, c2, cams[1], background_color=bg))
    with pytest.raises(ValueError):
        gr.render(synthetic.orbit_cameras(1, 320, 200, 100.0)[0])
    small = GraphRenderer(m, s, q, o, c, cams[0], bg, m_capacity=1000)
    small.render(cams[0])
    with pytest.raises(RuntimeError):
        small.check()


def test_host_frame_pipeline_equals_host_api(cuda_device):
    """Pipelined end-to-end path (H2D | render | D2H overlapped) == render_gaussians_host frame by frame."""
    from mojosplat_b200.pipeline import HostFramePipeline
    sc = synthetic.make_scene("config3_1m_1080p", N=120_000)
    host = [t.pin_memory() for t in sc.gaussians()]
    host2 = [host[0], host[1], host[2], host[3], host[4].flip(0).contiguous().pin_memory()]
    cams = synthetic.orbit_cameras(7, 640, 360, 330.0)
    pipe = HostFramePipeline(cuda_device, sc.N, 640, 360)
    out = torch.empty((7, 360, 640, 3), dtype=torch.float32).pin_memory()
    scenes = lambda k: host2 if k % 3 == 2 else host
    pipe.render(scenes, cams, sc.background, out)
    for k, cam in enumerate(cams):
        ref = ms.render_gaussians_host(*scenes(k), cam, background_color=sc.background, device=cuda_device)
        assert torch.equal(out[k], ref), k
    # static scene: uploaded once, every frame still gets its own camera and download
    out2 = torch.empty_like(out)
    pipe.render(lambda k: host, cams, sc.background, out2, upload="once")
    for k, cam in enumerate(cams):
        ref = ms.render_gaussians_host(*host, cam, background_color=sc.background, device=cuda_device)
        assert torch.equal(out2[k], ref), k


@pytest.mark.parametrize("cfg,N,world", [("config3_1m_1080p", 300_000, 8), ("config5_6m_4k", 200_000, 3),
                                         ("config2_100k_1080p", 20_000, 2)])
def test_row_band_renderer_bands_reassemble(cuda_device, cfg, N, world):
    """RowBandRenderer (sync-free band frames, band-compacted depth sort) with the ranks emulated one after the
    other on one GPU: the bands written into one buffer are the single-GPU image bit for bit."""
    sc = synthetic.make_scene(cfg, N=N)
    (m, s, q, o, c), cam = scene_on(sc, cuda_device)
    bg = sc.background.to(cuda_device)
    full = ms.render_fused(m, s, q, o, c, cam, bg)
    with torch.cuda.device(cuda_device):
        rb = parallel.RowBandRenderer(sc.N, cam, m_capacity=100 * sc.N)
    proj = ms.projection.project_gaussians_cuda(m, s, q, o, cam)
    cost = parallel.tile_row_cost(proj[0], proj[3], cam.H, cam.W, 16)
    bands = parallel.balanced_row_bands(cost.tolist(), world)
    rb.image.fill_(-1.0)
    total = 0
    for band in bands:
        rb.bands = [band]
        rb.render(m, s, q, o, c, cam, bg)
        total += rb.check()
    assert torch.equal(rb.image, full)
    _, aux = ms.render_fused(m, s, q, o, c, cam, bg, return_aux=True)
    assert total == aux["n_isect"]  # every (Gaussian, tile) pair lands in exactly one band


@pytest.mark.parametrize("pretest", [True, False])
def test_band_pretest_is_conservative(cuda_device, pretest):
    """Row-band frames only project the Gaussians a cheap pre-test cannot rule out of the band.  Adversarial inputs
    (huge / tiny / NaN / inf scales, zero and unnormalised quaternions, means behind and on the camera plane, far
    outside the frustum, non-finite means): every band must still hold exactly the pairs of the single-GPU frame."""
    g = torch.Generator().manual_seed(7)
    sc = synthetic.make_scene("config3_1m_1080p", N=120_000)
    (m, s, q, o, c), cam = scene_on(sc, cuda_device)
    m, s, q = m.clone(), s.clone(), q.clone()
    n = sc.N
    pick = lambda k: torch.randperm(n, generator=g)[:k].to(cuda_device)
    s[pick(2000)] += 3.0                       # e^3 larger: rectangles spanning many bands
    s[pick(2000), 1] = 6.0                     # one huge axis
    s[pick(500)] = -30.0                       # vanishing
    s[pick(50), 0] = float("nan")
    s[pick(50), 2] = float("inf")
    q[pick(300)] = 0.0                         # degenerate quaternion (normalisation clamp)
    q[pick(300)] *= 1e-14
    q[pick(300)] *= 1e6                        # unnormalised
    q[pick(20), 1] = float("nan")
    m[pick(3000), 2] *= -1.0                   # mirrored through the origin plane: many land behind the camera
    m[pick(200)] *= 50.0                       # far outside the frustum
    m[pick(20), 1] = float("inf")
    m[pick(20), 0] = float("nan")
    bg = sc.background.to(cuda_device)
    full, aux = ms.render_fused(m, s, q, o, c, cam, bg, return_aux=True)
    with torch.cuda.device(cuda_device):
        rb = parallel.RowBandRenderer(sc.N, cam, m_capacity=200 * sc.N, pretest=pretest)
    th = (cam.H + 15) // 16
    cuts = [0, 1, 7, th // 3, th // 2, th - 9, th - 1, th]
    rb.image.fill_(-1.0)
    total = 0
    for b0, b1 in zip(cuts[:-1], cuts[1:]):
        rb.bands = [(b0, b1)]
        rb.render(m, s, q, o, c, cam, bg)
        total += rb.check()
    assert total == aux["n_isect"]
    assert torch.equal(rb.image.nan_to_num(nan=-7.0), full.nan_to_num(nan=-7.0))


@pytest.mark.parametrize("cfg,N,W,H,f", [("config3_1m_1080p", 150_000, 800, 450, 420.0), ("config2_100k_1080p", 3_000, 640, 360, 170.0),
                                         ("config1_1k_256", 1_000, 256, 256, 66.7)])
def test_sync_free_paths_gsplat_rules(cuda_device, cfg, N, W, H, f):
    """The sync-free frame (pipeline, graph) under the gsplat rule set equals render_gaussians(backend="cuda_gsplat")."""
    from mojosplat_b200 import _lib
    from mojosplat_b200.pipeline import GraphRenderer, OverlappedPipeline
    sc = synthetic.make_scene(cfg, N=N)
    (m, s, q, o, c), _ = scene_on(sc, cuda_device)
    cams = synthetic.orbit_cameras(4, W, H, f)
    bg = sc.background.to(cuda_device)
    ref = [ms.render_gaussians(m, s, q, o, c, cam, background_color=bg, backend="cuda_gsplat") for cam in cams]
    pipe = OverlappedPipeline(cuda_device, sc.N, W, H, semantics=_lib.SEM_GSPLAT, m_capacity=200 * sc.N)
    imgs = pipe.render(m, s, q, o, c, cams, bg)
    assert pipe.check() == 0
    for k in range(len(cams)):
        assert torch.equal(imgs[k], ref[k]), k
    gr = GraphRenderer(m, s, q, o, c, cams[0], bg, semantics=_lib.SEM_GSPLAT, m_capacity=200 * sc.N)
    for k, cam in enumerate(cams):
        assert torch.equal(gr.render(cam), ref[k]), k
        gr.check()
    assert not torch.equal(ref[0], ms.render_gaussians(m, s, q, o, c, cams[0], background_color=bg, backend="cuda"))
