"""GPU parity tests of the photon path: emission + random walk, grid build, kNN radiance gather.

The reference's photon pass is unseeded (ThreadLocalRandom), so parity is defined against the oracle run with the SAME counter-based
sampler: stored photons must be the same set, the k-nearest estimate at any point must equal both the oracle's kd-tree answer and a
brute-force numpy selection, and rendered photon scenes must meet the image bars of BASELINE.json (<= 2/255 on >= 99.9 % of pixels,
PSNR >= 40 dB)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,photons", [("t05", 200000), ("t11", 20000), ("t10", 30000), ("t08", 100000), ("t04", 50000)])
def test_photon_emission_matches_oracle(drt, orc, gpu_ctx_factory, name, photons):
    ctx = gpu_ctx_factory()
    s = drt.Scene.from_cli(ctx, name + ".cli", photons=photons)
    st = s.emit_photons()
    g = s.photons()
    o = orc.OracleScene(name + ".cli", photons=photons)
    r = o.photons()
    assert len(g) == st.photons_stored
    # a hit decided by the last ulp of sin/cos/acos (CUDA libm vs glibc) may end a walk differently: allow 1e-4 of the set
    assert abs(len(g) - len(r)) <= max(2, 1e-4 * len(r)), (len(g), len(r))
    # same set of records (order differs: the oracle emits light-major and its kd build permutes the list): match by position
    from scipy.spatial import cKDTree
    dist, idx = cKDTree(r[:, :3]).query(g[:, :3], k=1)
    matched = (dist <= 1e-9) & np.isclose(g[:, 3:], r[idx, 3:], rtol=1e-9, atol=0).all(axis=1)
    assert matched.mean() >= 1 - 1e-3, matched.mean()
    ctx.close()


def brute_force(ph, pts, k, r2):
    out = np.zeros((len(pts), 4))
    for i, p in enumerate(pts):
        d2 = ((ph[:, 0] - p[0]) ** 2 + (ph[:, 1] - p[1]) ** 2) + (ph[:, 2] - p[2]) ** 2
        idx = np.nonzero(d2 < r2)[0]
        if len(idx) == 0:
            continue
        idx = idx[np.argsort(d2[idx], kind="stable")][:k]
        out[i, :3] = ph[idx, 3:].sum(axis=0)
        out[i, 3] = d2[idx].max()
    return out


@pytest.mark.parametrize("name,photons,k,r,tiers", [("t05", 400000, 80, 0.05, {1}), ("t11", 40000, 200, 0.1, {2}), ("t11", 400000, 200, 0.1, {3, 4})])
def test_knn_gather_matches_kdtree_and_brute_force(drt, orc, gpu_ctx_factory, name, photons, k, r, tiers):
    """Every tier of the gather (1: lane sums <= k candidates, 2: lane-serial histogram selection, 3: warp-cooperative over the coarse rows,
    4: warp-cooperative over a fine cube with a reduced radius) must return the exact k-nearest set."""
    ctx = gpu_ctx_factory()
    s = drt.Scene.from_cli(ctx, name + ".cli", photons=photons)
    ph = s.photons()
    assert len(ph) > k
    rng = np.random.default_rng(11)
    sel = rng.integers(0, len(ph), size=1500)
    pts = np.concatenate([ph[sel, :3] + rng.normal(0, 0.02, size=(1500, 3)),       # near the photon-carrying surfaces
                          rng.uniform(-3, 3, size=(200, 3)),                        # mostly empty space (outside the grid, too)
                          ph[:50, :3]])                                             # exactly on photons (d2 == 0 candidates)
    got = s.photon_probe(pts)
    assert tiers & set(int(t) for t in np.unique(got[:, 4])), np.unique(got[:, 4], return_counts=True)     # the tier this case is meant to exercise was taken
    r2 = float(np.float32(r)) ** 2                                                  # the reference keeps the radius as a float (myScene.java:927)
    want = brute_force(ph, pts, k, r2)
    assert np.array_equal(got[:, 3], want[:, 3])                                    # the k-th neighbour's d^2: bit-exact
    assert np.allclose(got[:, :3], want[:, :3], rtol=1e-12, atol=0)                 # sums differ only in addition order
    # and the oracle's kd-tree (its own photon set is the same set, see the emission test)
    if photons > 100000 and name == "t11":
        ctx.close()
        return                                                                     # (the oracle's single-threaded emission of 400 k diffuse photons takes too long for a test)
    o = orc.OracleScene(name + ".cli", photons=photons)
    ref = o.photon_probe(pts)
    same = np.isclose(ref[:, 3], got[:, 3], rtol=1e-9, atol=1e-15)
    assert same.mean() >= 0.995
    assert np.allclose(ref[same, :3], got[same, :3], rtol=1e-9, atol=1e-15)
    ctx.close()


@pytest.mark.parametrize("name,photons,spp", [("t05", 200000, 0), ("t10", 30000, 0), ("t11", 20000, 2), ("t08", 100000, 0), ("t04", 50000, 0),
                                              ("box_caustics", 200000, 0), ("box_gi", 30000, 0)])      # BASELINE configs[4] wrappers
def test_photon_scene_image_matches_oracle(drt, orc, gpu_ctx_factory, name, photons, spp):
    cols = rows = 150
    ctx = gpu_ctx_factory(cols, rows)
    s = drt.Scene.from_cli(ctx, name + ".cli", spp=spp, photons=photons)
    g = s.draw(aov=True)
    o = orc.OracleScene(name + ".cli", cols=cols, rows=rows, spp=spp if spp > 0 else -1, photons=photons)
    r = o.render(threads=os.cpu_count())
    ga, ra = orc.argb_to_rgb8(g["argb"]).astype(int), orc.argb_to_rgb8(r["argb"]).astype(int)
    d = np.abs(ga - ra).max(axis=-1)
    assert (d > 2).mean() <= 1e-3, (name, (d > 2).mean(), d.max())
    mse = ((ga - ra) ** 2).mean()
    assert mse == 0 or 10 * np.log10(255 ** 2 / mse) >= 40.0
    assert g["stats"].photons_stored > 0 and g["stats"].rays_photon > 0
    ctx.close()


def test_box_caustics_converges_to_oracle_reference(drt, orc, gpu_ctx_factory):
    """BASELINE north star for configs[4], second form: the GPU render at its own photon count against a far better converged oracle render
    (8x the photons, hence different photon paths).  The reference's estimate is k-nearest-neighbour density estimation (myLight.java:389-445,
    myObjShader.java:441-458): its variance is ~1/k whatever the photon count (k = 80 -> ~11 % per lit pixel), so two renders with different
    photon sets cannot agree to 40 dB however many photons are cast -- 40 dB is met where it is meaningful, against the oracle with the SAME
    seeded photons (test_photon_scene_image_matches_oracle[box_caustics]).  What this test pins is convergence: the error against the 12.8 M
    oracle render must shrink as the GPU casts more photons and reach the estimator's noise floor (measured 32.5 dB at 1.6 M)."""
    cols = rows = 160
    r = orc.OracleScene("box_caustics.cli", cols=cols, rows=rows, photons=12800000).render(threads=os.cpu_count(), want=("argb",))
    ra = orc.argb_to_rgb8(r["argb"]).astype(float)
    ps = []
    for n in (25000, 200000, 1600000):
        ctx = gpu_ctx_factory(cols, rows)
        g, _ = drt.Scene.from_cli(ctx, "box_caustics.cli", photons=n).draw()
        ctx.close()
        mse = ((orc.argb_to_rgb8(g).astype(float) - ra) ** 2).mean()
        ps.append(99.0 if mse == 0 else 10 * np.log10(255 ** 2 / mse))
    assert ps[0] < ps[1] < ps[2] and ps[2] >= 31.0, ps


def test_knn_gather_when_grid_cells_are_wider_than_the_radius(drt, gpu_ctx_factory, tmp_path):
    """A tiny photon radius on a large extent makes the grid double its cell size until it fits 2^22 coarse cells: the fine sub-cells are then
    wider than r/4, and a fine-cube search plan would look beyond the scene's radius (find_near requires d^2 < r^2, myLight.java:389-445).
    Scene: a glass ball focuses a dense caustic spot (thousands of photons inside one coarse cell), a mirror ball scatters photons over a
    1000-unit floor (the extent).  Every query must return the brute-force answer."""
    (tmp_path / "wide.cli").write_text(
        "fov 60\nbackground 0 0 0\npoint_light 0 5 -6 1 1 1\ncaustic_photons 1000000 40 0.05\ndiffuse .8 .8 .8 .1 .1 .1\n"
        "begin\nvertex -500 -1 -500\nvertex 500 -1 -500\nvertex 500 -1 500\nend\nbegin\nvertex 500 -1 500\nvertex -500 -1 500\nvertex -500 -1 -500\nend\n"
        "surface .1 .1 .1  0 0 0  1 1 1  20 1.0 1.9 1.8 .9 1.0 1.0\nsphere 1 0 0.05 -6\nreflective .1 .1 .1  0 0 0  0.95\nsphere 1 3 2 -6\nwrite x.png\n")
    ctx = gpu_ctx_factory(64, 64)
    s = drt.Scene.from_cli(ctx, "wide.cli", data_dir=str(tmp_path))
    ph = s.photons()
    assert len(ph) > 5000
    ext = ph[:, :3].max(axis=0) - ph[:, :3].min(axis=0)
    assert (np.floor(ext / 0.05) + 1).prod() > 2 ** 22                              # the grid cannot keep cell = r
    from scipy.spatial import cKDTree
    tree = cKDTree(ph[:, :3])
    dense = np.array([len(x) for x in tree.query_ball_point(ph[:4000, :3], 0.05)])
    assert (dense >= 40).sum() > 100                                               # neighbourhoods with >= k photons inside r exist
    rng = np.random.default_rng(3)
    sel = np.argsort(-dense)[:600]
    pts = np.concatenate([ph[sel, :3] + rng.normal(0, 0.008, size=(600, 3)), ph[rng.integers(0, len(ph), 400), :3]])
    got = s.photon_probe(pts)
    want = brute_force(ph, pts, 40, float(np.float32(0.05)) ** 2)
    assert np.array_equal(got[:, 3], want[:, 3])
    assert np.allclose(got[:, :3], want[:, :3], rtol=1e-12, atol=0)
    ctx.close()


def test_photon_split_emission_is_partition_independent(drt, gpu_ctx_factory):
    """Multi-GPU contract on one device: photon index ranges emitted separately and concatenated in rank order give the same record
    sequence -- and therefore the same grid and the same image -- as a single emission."""
    import torch
    from distraytracer_old_b200 import dist as D
    n = 30000
    ctx = gpu_ctx_factory(96, 96)
    s = drt.Scene.from_cli(ctx, "t11.cli", photons=n)
    whole = s.photons()
    a, _ = s.draw()
    parts = []
    for rank in range(3):
        i0, i1 = D.photon_range(n, 3, rank)
        s.emit_photons_range(i0, i1)
        cnt = s.photons_export_device(None, 0)
        buf = torch.empty((max(cnt, 1), 6), dtype=torch.float64, device="cuda")
        s.photons_export_device(buf.data_ptr(), cnt)
        parts.append(buf[:cnt].clone())
    allrec = torch.cat(parts).contiguous()
    assert np.array_equal(allrec.cpu().numpy(), whole)
    s.photons_build_device(allrec.data_ptr(), allrec.shape[0])
    b, _ = s.draw()
    assert np.array_equal(a, b)
    ctx.close()


def test_photon_map_edge_cases(drt, gpu_ctx_factory, tmp_path):
    # no photon ever stored (nothing specular in a caustic scene): render must not fail and must equal the no-photon image
    (tmp_path / "nocaustic.cli").write_text("fov 60\nbackground 0 0 0\npoint_light 0 2 0 1 1 1\ncaustic_photons 1000 10 0.1\ndiffuse .8 .8 .8 .1 .1 .1\nsphere 1 0 0 -4\nwrite x.png\n")
    (tmp_path / "plain.cli").write_text("fov 60\nbackground 0 0 0\npoint_light 0 2 0 1 1 1\ndiffuse .8 .8 .8 .1 .1 .1\nsphere 1 0 0 -4\nwrite x.png\n")
    ctx = gpu_ctx_factory(64, 64)
    a, st = drt.Scene.from_cli(ctx, "nocaustic.cli", data_dir=str(tmp_path)).draw()
    b, _ = drt.Scene.from_cli(ctx, "plain.cli", data_dir=str(tmp_path)).draw()
    assert st.photons_stored == 0 and np.array_equal(a, b)
    # zero photons requested
    s = drt.Scene.from_cli(ctx, "t05.cli", photons=0)
    c, st = s.draw()
    assert st.photons_stored == 0
    ctx.close()
