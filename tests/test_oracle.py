"""CPU tests of the oracle (the CPU restatement of the reference).  PARITY UNPINNED by reference tests: the
reference ships none; the oracle is pinned by (a) independent known answers (java.util.Random, Philox), (b) the
reference's one mapped output image t11c.png (loose PSNR), (c) frozen oracle outputs (regression), (d) ledger
behaviours Q1..Q21 of SURVEY.md checked as properties."""
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_java_util_random_known_answer(orc):
    # published JDK behaviour: new Random(0).nextDouble() == 0.730967787376657
    v = orc.java_random(0, 2)
    assert v[0] == 0.730967787376657
    assert v[1] == 0.24053641567148587
    assert orc.java_random(42, 1)[0] == 0.7275636800328681


def test_philox_reference_vectors(orc):
    # Random123 known-answer: philox4x32-10, counter 0, key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    # u01 packs (c0>>5)<<26 | (c1>>6) into 53 bits
    c0, c1 = 0x6627e8d5, 0xe169c58d
    expect = float(((c0 >> 5) << 26) | (c1 >> 6)) / 2.0 ** 53
    assert orc.u01(0, 0, 0, 0, 0, 0) == expect
    # counter ffffffff x4, key ffffffff x2 -> 408f276d 41c83b0e a20bc7c6 6d5451fd
    c0, c1 = 0x408f276d, 0x41c83b0e
    expect = float(((c0 >> 5) << 26) | (c1 >> 6)) / 2.0 ** 53
    assert orc.u01(0xFFFFFFFFFFFFFFFF, 0, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == expect


def test_golden_regression(orc, golden):
    for name, n in (("t01", 64), ("t03", 64), ("p3_t08", 64), ("p3_t02_sierp", 64), ("p3_t12", 64), ("p3_t06", 48), ("c5Fish", 48), ("planets3Ortho", 40), ("p2_t06", 32), ("t06", 48)):
        r = orc.OracleScene(name + ".cli", cols=n, rows=n).render()
        assert np.array_equal(r["argb"], golden[name + "_argb"]), name
        assert np.array_equal(r["hit_prim"], golden[name + "_hit_prim"]), name
        assert np.array_equal(r["hit_inst"], golden[name + "_hit_inst"]), name
    assert np.array_equal(orc.java_random(0, 4), golden["javarand0"])


def test_golden_textures_and_noise(orc, golden):
    pts = golden["probe_pts"]
    got = np.array([orc.perlin(*[np.float32(v) for v in p]) for p in pts], dtype=np.float32)
    assert np.array_equal(got, golden["perlin"])
    assert np.abs(got).max() <= 1.5 and np.abs(got).max() > 0.1
    for key in [k for k in golden.files if k.startswith("tex_")]:
        sc = key[4:]
        serial = 2 if sc.startswith("p4_st") or sc == "p3_t08" else 0
        out = orc.OracleScene(sc + ".cli").eval_texture(serial, pts)
        assert np.array_equal(out, golden[key]), key
        assert out.min() >= 0 and out.max() <= 1


def test_reference_image_t11c_loose_pin(orc):
    """The only reference-produced image with a scene in the checkout: t11c.png <- data/t11.cli (50 spp, 1M photons,
    unseeded RNG).  The oracle at 4 spp / 100k photons must be the same picture: PSNR after a 4x4 box filter."""
    from PIL import Image
    ref = np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "ref_t11c.png")).convert("RGB")).astype(float)
    s = orc.OracleScene("t11.cli", photons=100000, spp=4)
    img = orc.argb_to_rgb8(s.render(threads=os.cpu_count())["argb"]).astype(float)

    def box(a):
        return a.reshape(75, 4, 75, 4, 3).mean(axis=(1, 3))
    mse = ((box(img) - box(ref)) ** 2).mean()
    psnr = 10 * np.log10(255 ** 2 / mse)
    assert psnr > 24.0, psnr


def test_reference_image_t11_sierp_pin(orc):
    """t11_sierp.png is the reference's own render of data/project3/p3_t11_sierp.cli (Sierpinski depth 6 of bun69k instances, skydome, two
    lights, per-level pastel shaders) -- a DETERMINISTIC scene, so the only differences to the oracle are the stand-in mesh (bun69k.cli is
    missing from the checkout; tools/make_bun69k.py) and the unknown code revision.  It exercises the instance generator (Appendix B), the
    instance-level BVH with its non-conservative boxes (Q5), mixed t units (Q7), the CTM double application (Q6: the black bunnies), shading
    and the skydome lookup.  Measured: 23.7 dB per pixel, 34.6 dB after a 4x4 box filter."""
    from PIL import Image
    ref = np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "ref_t11_sierp.png")).convert("RGB")).astype(float)
    s = orc.OracleScene("p3_t11_sierp_d6.cli")
    img = orc.argb_to_rgb8(s.render(threads=os.cpu_count(), want=("argb",))["argb"]).astype(float)

    def box(a):
        return a.reshape(75, 4, 75, 4, 3).mean(axis=(1, 3))
    psnr_px = 10 * np.log10(255 ** 2 / ((img - ref) ** 2).mean())
    psnr_box = 10 * np.log10(255 ** 2 / ((box(img) - box(ref)) ** 2).mean())
    assert psnr_px > 21.0 and psnr_box > 31.0, (psnr_px, psnr_box)


def test_reference_image_t11_sierp_sky_and_silhouette_pin(orc):
    """Bit-level pin of the oracle against the reference's own render where the missing mesh cannot interfere (tests/golden/ref_masks.py):
    every pixel the reference shows as sky (62 % of the frame) must come out within 2/255 -- measured: ALL of them identical -- which pins the
    camera, the skydome lookup and the texel decode; the union of instance silhouettes must overlap the reference's with IoU >= 0.9
    (measured 0.992 with the stand-in mesh), which pins the Sierpinski transform chain and the instance-level BVH."""
    from tests.golden import ref_masks
    ref = ref_masks.load_ref()
    assert 0.60 < ref["sky"].mean() < 0.64
    r = orc.OracleScene("p3_t11_sierp_d6.cli").render(threads=os.cpu_count(), want=("argb", "hit_prim"))
    res = ref_masks.compare(ref, orc.argb_to_rgb8(r["argb"]), r["hit_prim"] >= 0)
    assert res["sky_bad_frac"] <= 1e-3 and res["sky_exact_frac"] >= 0.999, res
    assert res["sky_raw_bad_frac"] <= 5e-3, res          # outline band included: the stand-in's outline differs from the real bunny's by a few pixels
    assert res["iou"] >= 0.9, res


def test_bvh_median_split_ledger(orc, golden):
    s = orc.OracleScene("p3_t08.cli")
    d, box = s.dump_bvh(2)
    assert zlib.crc32(d.tobytes()) == int(golden["p3_t08_bvh_crc"][0])
    assert np.array_equal(box, golden["p3_t08_bvh_box"])
    leaf_ids = []
    i = 0
    while i < len(d):
        if d[i] == -1:
            i += 1
        else:
            n = d[i + 1]
            assert 1 <= n <= 5            # maxPrimsPerLeaf (DistRayTracer.java:35)
            leaf_ids += list(d[i + 2:i + 2 + n])
            i += 2 + n
    # Q2: exactly one of the 966 triangles is dropped by the root call (endIDX = size-1)
    assert len(leaf_ids) == 965 and len(set(leaf_ids)) == 965
    s2 = orc.OracleScene("p3_t02_sierp.cli")
    d2, _ = s2.dump_bvh(0)
    assert zlib.crc32(d2.tobytes()) == int(golden["sierp_bvh_crc"][0])
    inst = [x & 0x3FFFFFFF for x in d2 if x >= 0x40000000]
    assert len(inst) == 340               # (4^5-1)/3 = 341 instances, one dropped


def test_box_rule_origin_inside_misses(orc):
    # Q1: a ray starting inside an accel structure's root box never hits it (entry t must be > 0)
    s = orc.OracleScene("p3_t08.cli")
    ids, t = s.trace_rays(np.array([[0.0, 0.0, -3.0], [0.0, 0.0, 0.0]]), np.array([[0.0, 0.3, -1.0], [0.0, 0.0, -1.0]]))
    assert ids[0, 0] < 2                  # inside the bunny's box: only floor (0/1) or nothing
    assert ids[1, 0] >= 2                 # from the eye: a bunny triangle


def test_two_sided_polygons_and_clamp(orc):
    # Q9: triangles are hit from both sides; Q14: output channels are (int)(min(1,c)*255)
    s = orc.OracleScene("t03.cli")
    r = s.render()
    a = r["argb"].astype(np.uint32)
    assert ((a >> 24) == 255).all()
    ids_front, _ = s.trace_rays(np.array([[0.0, 0.0, 0.0]]), np.array([[0.0, -0.3, -1.0]]))
    ids_back, _ = s.trace_rays(np.array([[0.0, -5.0, -3.0]]), np.array([[0.0, 1.0, 0.0]]))
    assert ids_front[0, 0] >= 0 and ids_back[0, 0] >= 0


def test_literal_renormalisation_is_ulp_level(orc):
    # canonical mode (normalise once) vs literal mode (re-normalise on every getTransformedRay, myRay.java:93)
    a = orc.OracleScene("p3_t08.cli", cols=96, rows=96).render()
    b = orc.OracleScene("p3_t08.cli", cols=96, rows=96, literal_renorm=True).render()
    assert (a["hit_prim"] != b["hit_prim"]).mean() < 1e-3
    d = np.abs(orc.argb_to_rgb8(a["argb"]).astype(int) - orc.argb_to_rgb8(b["argb"]).astype(int))
    assert d.max() <= 1


def test_seeded_sampler_is_deterministic_and_thread_invariant(orc):
    a = orc.OracleScene("p2_t06.cli", cols=40, rows=40).render(threads=1)
    b = orc.OracleScene("p2_t06.cli", cols=40, rows=40).render(threads=4)
    assert np.array_equal(a["argb"], b["argb"])
    c = orc.OracleScene("p2_t06.cli", cols=40, rows=40, seed=1).render()
    assert not np.array_equal(a["argb"], c["argb"])


def test_photon_emission_statistics(orc):
    s = orc.OracleScene("t05.cli", photons=20000)
    ph = s.photons()
    assert 200 < len(ph) < 20000          # caustic photons only survive via the mirror
    assert np.isfinite(ph).all()
    # stored on the floor plane y = -1 or on the green sphere
    assert (np.abs(ph[:, 1] + 1) < 1e-6).mean() > 0.5
    s2 = orc.OracleScene("t11.cli", photons=5000)
    ph2 = s2.photons()
    assert len(ph2) > 5000                # diffuse photons: several stores per emitted photon, two lights
    assert (np.abs(ph2[:, :2]).max() <= 1.0 + 1e-6)


def test_empty_and_degenerate_scenes(orc, tmp_path):
    p = tmp_path / "empty.cli"
    p.write_text("fov 60\nbackground 0.2 0.4 1\nwrite x.png\n")
    r = orc.OracleScene("empty.cli", cols=8, rows=8, data_dir=str(tmp_path)).render()
    assert (orc.argb_to_rgb8(r["argb"]) == np.array([51, 102, 255])).all()
    assert (r["hit_prim"] == -1).all()
    # no fov line: rays_per_pixel stays 0 -> 0/0 -> NaN -> black (myRTFileReader.java:32, myScene.java:1509)
    p2 = tmp_path / "nofov.cli"
    p2.write_text("background 1 1 1\nsphere 1 0 0 -4\n")
    r2 = orc.OracleScene("nofov.cli", cols=8, rows=8, data_dir=str(tmp_path)).render()
    assert (orc.argb_to_rgb8(r2["argb"]) == 0).all()
    with pytest.raises(RuntimeError):
        orc.OracleScene("does_not_exist.cli")
