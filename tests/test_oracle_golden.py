"""Pin the CPU oracle (oracle/paresis_oracle.py) to the reference's own outputs.

The golden vectors were produced by the unmodified reference (oracle/make_golden.py); the
reference has no tests of its own.  All checks are fp64 vs fp64, so tolerances are at
round-off level: a restatement that deviates anywhere fails here, not on the GPU box.
"""
import numpy as np
import pytest

import paresis_oracle as po
from conftest import rel_l2
from ref_harness import synthetic_sphere_rows

TIGHT = 1e-12


def test_wavenumber_constants():
    # getk.py:12-20 at 25 keV (its own __main__ example)
    assert po.wavenumber(25000) == 2 * np.pi * 25000 * 1.6e-19 / (6.626e-34 * 2.998e8)
    assert abs(po.wavenumber_from_lambda(25.0) / po.wavenumber(25000) - 1) < 1e-15


def test_splat_known_answers(golden):
    g = golden("splat_kernel")
    for row in g["known_answers"]:
        dxv, dyv, want = row[0], row[1], row[2:].reshape(9, 9)
        I = np.zeros((9, 9)); I[4, 4] = 1.0
        Dx = np.zeros((9, 9)); Dx[4, 4] = dxv
        Dy = np.zeros((9, 9)); Dy[4, 4] = dyv
        for fn in (po.splat, po.splat_python):
            got = fn(I, Dx, Dy, 0)
            assert np.allclose(got, want, rtol=0, atol=1e-15), (dxv, dyv, fn.__name__)
    # integer shifts move the whole ray
    ka = {(r[0], r[1]): r[2:].reshape(9, 9) for r in g["known_answers"]}
    assert ka[(1.0, 0.0)][5, 4] == 1.0 and ka[(-1.0, 0.0)][3, 4] == 1.0
    assert ka[(-0.25, 0.0)][3, 4] == 0.25 and ka[(-0.25, 0.0)][4, 4] == 0.75
    assert ka[(-1.25, 0.0)][2, 4] == 0.25 and ka[(-1.25, 0.0)][3, 4] == 0.75
    assert abs(ka[(1e-13, 0.0)][4, 4] - 1.0) < 1e-12  # the raw kernel keeps tiny shifts; fastRefraction zeroes them (:59)


def test_splat_edge_quirk(golden):
    want = golden("splat_kernel")["edge_quirk"]
    I = np.zeros((5, 5)); I[4, 2] = 1.0
    Dx = np.zeros((5, 5)); Dy = np.zeros((5, 5)); Dy[4, 2] = 0.5
    assert want.sum() == 0.5  # the reference drops the neighbour deposit on the last row
    assert np.array_equal(po.splat(I, Dx, Dy, 0), want)
    assert np.array_equal(po.splat_python(I, Dx, Dy, 0), want)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_splat_random_frames(golden, tag):
    g = golden("splat_kernel")
    got = po.splat(g["I_" + tag], g["Dx_" + tag], g["Dy_" + tag], 0)
    assert rel_l2(got, g["out_" + tag]) < TIGHT
    assert rel_l2(po.splat_python(g["I_" + tag], g["Dx_" + tag], g["Dy_" + tag], 0), g["out_" + tag]) < TIGHT


@pytest.mark.parametrize("tag", ["sub", "mid", "far", "huge"])
def test_fast_refraction(golden, tag):
    g = golden("fast_refraction")
    pix, z, E, M = g["params"]
    out, dx, dy = po.fast_refraction(g["I"].copy(), g["phi_" + tag], z, E, M, pix)
    assert dx.shape == (97 + 30, 113 + 30)
    assert rel_l2(dx, g["Dx_" + tag]) < TIGHT and rel_l2(dy, g["Dy_" + tag]) < TIGHT
    assert rel_l2(out, g["out_" + tag]) < TIGHT
    # v1 (margin 10, fixed clamp) is the same map whenever |D| stays below min(1e3, N)
    out10, _, _ = po.fast_refraction(g["I"].copy(), g["phi_" + tag], z, E, M, pix, margin=10)
    if tag != "huge":
        assert rel_l2(out10, g["outv1_" + tag]) < TIGHT


def test_gaussian_kernels(golden):
    g = golden("detector")
    widths = []
    for k, s in enumerate(g["sigmas"]):
        got = po.gaussian_kernel(float(s))
        assert got.shape == g["g%d" % k].shape
        assert np.allclose(got, g["g%d" % k], rtol=1e-14, atol=0)
        widths.append(got.shape[0])
    assert widths == [1, 3, 5, 5, 9, 9, 13]  # round-half-even: 3*0.8333=2.5 -> 2, 3*1.5=4.5 -> 4


def test_bin_sum(golden):
    g = golden("detector")
    assert np.allclose(po.bin_sum(g["resize_in"], 30, 45), g["resize_2"], rtol=1e-14)
    assert np.allclose(po.bin_sum(g["resize_in"], 20, 30), g["resize_3"], rtol=1e-14)
    assert np.array_equal(po.bin_sum(g["resize_in"], 60, 90), g["resize_id"])


def test_detection(golden):
    g = golden("detector")
    for k in range(int(g["n_det"])):
        os_, d0, d1, fwhm, psf = g["det%d_cfg" % k]
        got = po.detection(g["det%d_in" % k], float(fwhm), int(os_), (int(d0), int(d1)), float(psf))
        assert got.shape == (int(d0), int(d1))
        assert rel_l2(got, g["det%d_out" % k]) < 1e-13, k


def test_transmission_and_propagation(golden):
    g = golden("waves")
    E = float(g["E"])
    i_rt, phi_rt = po.set_wave_rt(g["I"], g["phi0"], g["t"], g["delta"], g["beta"], E)
    assert rel_l2(i_rt, g["I_rt"]) < 1e-15 and rel_l2(phi_rt, g["phi_rt"]) < 1e-15
    wave = po.set_wave(g["wave0"], g["t"], g["delta"], g["beta"], E)
    assert rel_l2(np.abs(wave), np.abs(g["wave"])) < 1e-14
    assert np.abs(wave - g["wave"]).max() / np.abs(g["wave"]).max() < 1e-12
    for k in range(3):
        z, M, pix = g["prop%d_cfg" % k]
        got = po.wave_propagation(g["wave"], float(z), E, float(M), g["wave"].shape, float(pix))
        assert np.abs(got - g["prop%d" % k]).max() / np.abs(g["prop%d" % k]).max() < 1e-12


def test_membrane_raster(golden):
    g = golden("geometry")
    rows = synthetic_sphere_rows(0, 60000)
    for tag in ("mem0", "mem1"):
        mean_r, layers, dx, dy, pix, support, seed = g[tag + "_cfg"]
        np.random.seed(int(seed))
        got = po.membrane_segmented(rows, float(mean_r), int(layers), int(dx), int(dy), float(pix), float(support))
        assert got.shape == g[tag].shape
        assert np.array_equal(got[1], g[tag][1])
        assert rel_l2(got[0], g[tag][0]) < 1e-13, tag
        assert got[0].max() > 0


def test_sample_shapes(golden):
    g = golden("geometry")
    r, dx, dy, pix = g["sphere_cfg"]
    assert rel_l2(po.sample_sphere(float(r), int(dx), int(dy), float(pix)), g["sphere"]) < 1e-15
    for tag in ("cyl", "cyl2"):
        r, ang, dx, dy, pix = g[tag + "_cfg"]
        got = po.sample_cylinder(float(r), float(ang), int(dx), int(dy), float(pix))
        assert got.shape == g[tag].shape
        # OpenCV warpAffine restated (third-party step, cv2 4.13 pinned by this fixture)
        assert rel_l2(got, g[tag]) < 1e-12, tag


def _setup_from(g, sim):
    d1, d2, d3, det_pix, os_, shots, src, psf, esamp = g["cfg"]
    spectrum = [(float(e), float(w)) for e, w in g["spectrum"]]
    return po.Setup(d1, d2, d3, g["det_dims"], det_pix, int(os_), shots, spectrum, src, psf,
                    energy_sampling=esamp, bin_thresholds=list(g["thresholds"]))


def _db(table, n_mat):
    return {float(r[0]): (r[1:1 + n_mat], r[1 + n_mat:1 + 2 * n_mat]) for r in table}


@pytest.mark.parametrize("name", ["e2e_rt_cylinder", "e2e_rt_sphere", "e2e_rt_poly3", "e2e_fresnel_sphere"])
def test_end_to_end(golden, name):
    g = golden(name)
    s = _setup_from(g, name)
    if str(g["sample_name"]) == "PMMA_sphere":
        sample_t = po.sample_sphere(1000.0, s.study_dims[0], s.study_dims[1], s.study_pixel_um)
    else:
        sample_t = po.sample_cylinder(700.0, 30.0, s.study_dims[0], s.study_dims[1], s.study_pixel_um)
    probe = g["sample_t_probe"]
    assert abs(sample_t.sum() / probe[0] - 1) < 1e-12 and abs(sample_t.max() / probe[1] - 1) < 1e-12
    rows = synthetic_sphere_rows(0, 60000)
    mean_r, layers, mem_pix = g["membrane_cfg"]
    assert abs(mem_pix / s.membrane_pixel_um - 1) < 1e-15
    mdb, sdb = _db(g["membrane_db"], 2), _db(g["sample_db"], 1)
    fn = po.compute_fresnel if "fresnel" in name else po.compute_rt
    for point in (0, 1):
        np.random.seed(int(g["membrane_seed_p%d" % point]))
        mem_t = po.membrane_segmented(rows, float(mean_r), int(layers), s.study_dims[0], s.study_dims[1],
                                      s.membrane_pixel_um, float(g["support_um"]))
        mp = g["membrane_probe_p%d" % point]
        assert abs(mem_t[0].sum() / mp[0] - 1) < 1e-12 and mem_t[0][5, 9] == mp[2]
        sample, ref, propag, white, _ = fn(s, mem_t, mdb, sample_t, sdb, point)
        tol = 1e-10
        assert rel_l2(sample, g["sample_p%d" % point]) < tol
        assert rel_l2(ref, g["reference_p%d" % point]) < tol
        if point == 0:
            assert rel_l2(propag, g["propag_p0"]) < tol
            assert rel_l2(white, g["white_p0"]) < tol
        else:
            assert not propag.any()


def test_dark_field_branch(golden):
    """fastRefractionDF (refractionFileNumba2.py:88-196) and setWaveRT with the Lung model (Sample.py:322-343)."""
    g = golden("darkfield")
    pix, z, E, M = g["params"]
    for tag in ("narrow", "wide"):
        out, dx, dy = po.fast_refraction_df(g["I"].copy(), g["phi"], z, E, M, pix, g["df_" + tag])
        assert dx.shape == g["Dx_" + tag].shape
        assert rel_l2(dx, g["Dx_" + tag]) < 1e-12
        assert rel_l2(out, g["out_" + tag]) < 1e-12, tag
    i_out, phi_out, df = po.set_wave_rt_df(g["I"], g["phi"], g["t"], g["sw_db"][0], g["sw_db"][1], E, ["Lung", None])
    assert rel_l2(i_out, g["sw_I"]) < 1e-13 and rel_l2(phi_out, g["sw_phi"]) < 1e-13 and rel_l2(df, g["sw_df"]) < 1e-13


def test_two_sphere_phantoms(golden):
    """createSampGeom.py:110-260 (OpenCV's warpAffine restated for the tilted one)."""
    g = golden("phantoms")
    for tag, kind in (("cyl_a", 0), ("cyl_b", 0), ("par_a", 1), ("par_b", 1)):
        dx, dy, pix = g[tag + "_cfg"]
        got = po.sample_two_spheres(kind, int(dx), int(dy), float(pix))
        assert got.shape == g[tag].shape
        assert rel_l2(got, g[tag]) < 1e-12, tag
