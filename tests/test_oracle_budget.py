"""CPU: the error-budget side of the oracle -- fp64 variants and the outlier audit.
No GPU: these pin the checker itself; the GPU tests then use it (tests/test_gpu_budget.py)."""
import numpy as np

from conftest import load_golden
from helpers import audit_inputs, camera_from_golden, image_gate
from mojosplat_b200 import synthetic
from oracle import oracle


def _scene(cfg="config2_100k_1080p", N=3000):
    sc = synthetic.make_scene(cfg, N=N)
    cam = sc.camera
    ref = oracle.render(sc.means3d.numpy(), sc.log_scales.numpy(), sc.quats.numpy(), sc.opacities.numpy(),
                        sc.colors.numpy(), cam.view_matrix.numpy(), cam.fx, cam.fy, cam.cx, cam.cy, cam.W, cam.H,
                        cam.near, cam.far, background=sc.background.numpy(), return_all=True)
    return sc, ref


def test_projection_f64_agrees_with_fp32_oracle_and_golden():
    """fp32 torch output (golden, made by the unmodified reference) and the fp32 oracle sit at the same distance
    from the fp64 evaluation: the fp64 routine restates the same mathematics."""
    g = load_golden("garden_6k_1080p")
    cam = camera_from_golden(g)
    vm = cam.view_matrix.numpy()
    m64, c64, d64, r64 = oracle.project_f64(g["means3d"], g["log_scales"], g["quats"], vm, cam.fx, cam.fy, cam.cx,
                                            cam.cy, cam.W, cam.H, cam.near, cam.far)
    vis = (g["radii"] > 0).all(-1)
    assert vis.sum() > 1000
    # 1 ulp at |x| ~ 2000 px is 1.2e-4 -- the reference's own fp32 error on means2d
    assert np.abs(g["means2d"][vis] - m64[vis]).max() < 5e-3
    assert np.abs(g["depths"][vis] - d64[vis]).max() < 1e-5
    rel = np.abs(g["conics"][vis] - c64[vis]) / (np.abs(c64[vis]) + 1e-6)
    assert rel.max() < 1e-3
    # radii = ceil(real radius) except at integer crossings
    assert (np.abs(np.ceil(r64[vis]) - g["radii"][vis]) <= 1).all()
    assert (np.ceil(r64[vis]) != g["radii"][vis]).mean() < 1e-3
    o = oracle.project(g["means3d"], g["log_scales"], g["quats"], g["opacities"], vm, cam.fx, cam.fy, cam.cx, cam.cy,
                       cam.W, cam.H, cam.near, cam.far)
    e_ref = np.abs(g["means2d"][vis] - m64[vis]).mean()
    e_orc = np.abs(o[0][vis] - m64[vis]).mean()
    assert e_orc <= 1.25 * e_ref + 1e-7


def test_raster_f64_vs_fp32_oracle():
    sc, ref = _scene()
    cam = sc.camera
    img64 = oracle.rasterize_f64(ref["means2d"], ref["conics"], sc.colors.numpy(), sc.opacities.numpy(),
                                 sc.background.numpy(), ref["tile_ranges"], ref["sorted_ids"], cam.W, cam.H, 16)
    err = np.abs(ref["image"].astype(np.float64) - img64)
    assert np.sqrt((err ** 2).mean()) < 2e-6           # rounding noise ...
    assert (err > 1e-4).mean() < 1e-4                  # ... plus a few threshold flips (H4)


def test_audit_explains_flips_and_rejects_corruption():
    """Another correct fp32 implementation = the oracle with opacities perturbed by a few ulp: every pixel that leaves
    the tolerance must be explained by flipped threshold decisions; a corrupted pixel must not be."""
    sc, ref = _scene("config2_100k_1080p", 6000)
    cam = sc.camera
    op2 = (sc.opacities.numpy().astype(np.float64) * (1.0 + 3e-7)).astype(np.float32)
    img2 = oracle.rasterize(ref["means2d"], ref["conics"], sc.colors.numpy(), op2, sc.background.numpy(),
                            ref["tile_ranges"], ref["sorted_ids"], cam.W, cam.H, 16)
    a = audit_inputs(ref, sc)
    r = image_gate(img2, ref["image"], audit=a)
    assert r["ok"], r
    assert r["n_unexplained"] == 0
    bad = img2.copy()
    bad[100, 200, 1] += 0.01          # not a threshold effect
    bad[500, 900, :] += 2e-3
    r2 = image_gate(bad, ref["image"], audit=a)
    assert not r2["ok"] and r2["n_unexplained"] == 2, r2
    # the max-error bound alone catches gross errors
    bad2 = ref["image"].copy()
    bad2[10, 10, 0] += 0.5
    assert not image_gate(bad2, ref["image"])["ok"]
