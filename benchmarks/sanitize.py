"""compute-sanitizer workload: every kernel of libbsplat.so on small scenes (both rule sets, packed layout, row bands,
long-list pre-pass, sync-free / captured frames, backward, SH).  Run as

    compute-sanitizer --tool memcheck  python benchmarks/sanitize.py
    compute-sanitizer --tool racecheck python benchmarks/sanitize.py
    compute-sanitizer --tool synccheck python benchmarks/sanitize.py

and keep the summaries under profiles/ (the raster staging double buffer, the onesweep / scan look-back chains and
the multi-CTA tile finish are the kernels that need it)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import mojosplat_b200 as ms
from mojosplat_b200 import binning, parallel, rasterization, synthetic
from mojosplat_b200.pipeline import GraphRenderer, OverlappedPipeline

dev = torch.device("cuda:0")
small = "--small" in sys.argv
for cfg, N, sem in [("config1_1k_256", None, "cuda"), ("config3_1m_1080p", 8000 if small else 30000, "cuda"),
                    ("config3_1m_1080p", 8000 if small else 30000, "cuda_gsplat"),
                    ("config2_100k_1080p", 1500 if small else 3000, "cuda")]:
    sc = synthetic.make_scene(cfg, N=N)
    g = [t.to(dev) for t in sc.gaussians()]
    bg = sc.background.to(dev)
    img = ms.render_gaussians(*g, sc.camera, background_color=bg, backend=sem)
    semv = ms.projection.CUDA_BACKENDS[sem]
    img2, aux = ms.render_fused(*g, sc.camera, bg, 16, semantics=semv, return_aux=True)
    img2, aux = ms.render_fused(*g, sc.camera, bg, 16, semantics=semv, return_aux=True)
    img3, _ = ms.render_fused(*g, sc.camera, bg, 16, semantics=semv, return_aux=True, packed=True)
    img4, _ = ms.render_fused(*g, sc.camera, bg, 16, semantics=semv, return_aux=True, bin_algo="single")
    assert torch.equal(img2, img3) and torch.equal(img2, img4)
    for mode in ["fast", "faithful", "fast_nocull"]:
        rasterization.rasterize_gaussians_cuda(aux["means2d"], aux["conics"], g[4], g[3], bg, aux["tile_ranges"],
                                               aux["sorted_ids"], sc.camera, 16, mode=mode)
    # small tiles: more than 65 536 tile ids (3-pass tile sort), stage-level binning with and without compaction
    for ts, packed in ((4, False), (4, True), (16, True)):
        binning.bin_gaussians_to_tiles_cuda(aux["means2d"], aux["radii"], aux["depths"], sc.camera.H, sc.camera.W, ts,
                                            semantics=semv, packed=packed)
    t = [aux["means2d"].clone().requires_grad_(True), aux["conics"].clone().requires_grad_(True),
         g[4].clone().requires_grad_(True), g[3].clone().requires_grad_(True)]
    for mode in ("fast", "faithful"):  # pair-layout kernels / generic kernels
        out = rasterization.rasterize_gaussians_diff(*t, bg, aux["tile_ranges"], aux["sorted_ids"], sc.camera, 16,
                                                     mode=mode)
        out.sum().backward()
    pipe = OverlappedPipeline(dev, sc.N, sc.camera.W, sc.camera.H, semantics=semv, m_capacity=200 * sc.N + 4096)
    cams = synthetic.orbit_cameras(4, sc.camera.W, sc.camera.H, sc.camera.fx)
    pipe.render(*g, cams, bg); pipe.check()
    tiny = OverlappedPipeline(dev, sc.N, sc.camera.W, sc.camera.H, m_capacity=500)
    tiny.render(*g, cams[:2], bg); tiny.check()
    gr = GraphRenderer(*g, sc.camera, bg, semantics=semv, m_capacity=200 * sc.N + 4096)
    gr.render(cams[1]); gr.check()
    full = ms.render_fused(*g, sc.camera, bg, 16, semantics=semv)
    assert torch.equal(parallel.render_frame_row_split(*g, sc.camera, bg, semantics=semv), full)
    rb_img = torch.zeros_like(full)
    th = (sc.camera.H + 15) // 16
    for bi, band in enumerate(((0, th // 3), (th // 3, th // 3), (th // 3, th // 2), (th // 2, th))):   # (one empty band)
        # (torch rules: bands 0..2 through the band pre-test + candidate-list projection, the last one without)
        rb = parallel.RowBandRenderer(sc.N, sc.camera, semantics=semv, pretest=bi < 3, m_capacity=200 * sc.N + 4096)
        rb.bands = [band]
        rb.render(*g, sc.camera, bg); rb.check()
        r0, r1 = band[0] * 16, min(band[1] * 16, sc.camera.H)
        rb_img[r0:r1] = rb.image[r0:r1]
    assert torch.equal(rb_img, full)
    coeffs = torch.randn(sc.N, 16, 3, device=dev)
    ms.eval_sh(3, coeffs, g[0], sc.camera)
    torch.cuda.synchronize()
    print(cfg, sem, "ok", float(img.mean()), flush=True)
