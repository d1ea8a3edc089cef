"""End-to-end parity of the drop-in shim against the reference's own images.

Same flow as PARESIS's main.py:63-71 (new membrane per point, then compute), Poisson noise
off on both sides, same seeds for the membrane offsets.  Goldens come from the unmodified
reference (oracle/make_golden.py).  Bound: 1e-5 relative L2 (north star).
"""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from paresis_b200 import workspace
    ws = workspace.make_workspace(str(tmp_path_factory.mktemp("ws")))
    old = os.getcwd()
    workspace.enter(ws)
    yield importlib.import_module("Experiment")
    os.chdir(old)


CASES = [
    ("e2e_rt_cylinder", "Fil_Nylon_ID17", "RayT"),
    ("e2e_rt_sphere", "Small_sphere_mono", "RayT"),
    ("e2e_rt_poly3", "B200_small_poly3", "RayT"),
    ("e2e_fresnel_sphere", "Small_sphere_mono", "Fresnel"),
]


@pytest.mark.parametrize("gold,name,model", CASES)
def test_matches_reference_images(shim, golden, gold, name, model):
    g = golden(gold)
    d = dict(experimentName=name, filepath="unused/", overSampling=2, nbExpPoints=2, simulation_type=model,
             expID="t", poissonNoise=False)
    e = shim.Experiment(d)
    assert np.allclose(d["magnification"], 145.2 / 141.6)
    # the delta/beta the shim derived are the ones the reference used
    for m in range(2):
        assert np.allclose([v for _, v in e.myMembrane.delta[m]], g["membrane_db"][:, 1 + m], rtol=1e-12)
        assert np.allclose([v for _, v in e.myMembrane.beta[m]], g["membrane_db"][:, 3 + m], rtol=1e-12)
    probe = g["sample_t_probe"]
    st = np.asarray(e.mySampleofInterest.myGeometry)
    assert st.ndim == 3 and abs(st.sum() / probe[0] - 1) < 1e-6
    for point in (0, 1):
        np.random.seed(int(g["membrane_seed_p%d" % point]))
        e.myMembrane.myGeometry = []
        e.myMembrane.getMyGeometry(e.exp_dict['studyDimensions'], e.myMembrane.membranePixelSize, 2, point, 2)
        mt = e.myMembrane.myGeometry[0]
        mp = g["membrane_probe_p%d" % point]
        assert abs(mt.astype(np.float64).sum() / mp[0] - 1) < 1e-6 and abs(mt[5, 9] / mp[2] - 1) < 1e-6 if mp[2] else True
        if model == "RayT":
            res = e.computeSampleAndReferenceImages_RT(point)
            assert len(res) == 7 and res[6].shape == tuple(e.exp_dict['studyDimensions'])
        else:
            res = e.computeSampleAndReferenceImages_Fresnel(point)
            assert len(res) == 4
        sample, ref, propag, white = res[:4]
        assert sample.dtype == np.float64 and sample.shape == g["sample_p%d" % point].shape
        assert rel_l2(ref, g["reference_p%d" % point]) < TOL
        assert rel_l2(sample, g["sample_p%d" % point]) < TOL
        if point == 0:
            assert rel_l2(propag, g["propag_p0"]) < TOL
            assert rel_l2(white, g["white_p0"]) < TOL
            if model == "RayT":
                assert res[4].shape == (st.shape[1] + 30, st.shape[2] + 30)
                # auxiliary output (main.py discards it): differences of fp32 thickness maps carry ~6e-8 * t/dt
                assert rel_l2(res[4][::8, ::8], g["Dx_p0_s8"]) < 5e-5 and rel_l2(res[5][::8, ::8], g["Dy_p0_s8"]) < 5e-5
        else:
            assert not propag.any() and not white.any()
    assert abs(e.exp_dict["meanEnergy"] / float(g["mean_energy"]) - 1) < 1e-5
    # main.py:84-85 relies on the in-place bin closing
    assert e.myDetector.det_param["myBinsThersholds"][-1] == e.mySource.mySpectrum[-1][0]


def test_api_pieces_and_errors(shim, golden):
    """The module-level functions keep their numpy-in / numpy-out contracts."""
    r2 = importlib.import_module("refractionFileNumba2")
    r1 = importlib.import_module("refractionFileNumba")
    det = importlib.import_module("Detector")
    g = golden("fast_refraction")
    pix, z, E, M = g["params"]
    out, dx, dy = r2.fastRefraction(g["I"].copy(), g["phi_mid"], z, E, M, pix)
    assert out.dtype == np.float64 and dx.shape == (127, 143)
    assert rel_l2(out, g["out_mid"]) < TOL and rel_l2(dx, g["Dx_mid"]) < 1e-6
    out1, dx1, _ = r1.fastRefraction(g["I"].copy(), g["phi_mid"], z, E, M, pix)
    assert dx1.shape == (117, 133) and rel_l2(out1, g["outv1_mid"]) < TOL
    bad = g["I"].copy(); bad[4, 4] = np.nan
    with pytest.raises(Exception, match="nans or insane values"):
        r2.fastRefraction(bad, g["phi_mid"], z, E, M, pix)
    k = golden("splat_kernel")
    acc = np.zeros((37, 41))
    ret = r2.fastloopNumba(37, 41, k["I_b"], acc, k["Dy_b"], k["Dx_b"], None, None)
    assert ret is acc and rel_l2(acc, k["out_b"]) < 2e-6
    d = golden("detector")
    assert np.allclose(det.create_gaussian_shape(1.2), d["g4"], rtol=1e-14)
    assert np.array_equal(r2.gaussian_shape(0.5), det.create_gaussian_shape(0.5))
    assert rel_l2(det.resize(d["resize_in"], 30, 45), d["resize_2"]) < 1e-6
    assert det.resize(d["resize_in"], 60, 90) is d["resize_in"] or np.array_equal(det.resize(d["resize_in"], 60, 90), d["resize_in"])
    with pytest.raises(ValueError, match="experiment not found in xml file"):
        shim.Experiment(dict(experimentName="nope", filepath="", overSampling=2, nbExpPoints=1, simulation_type="RayT", expID="x"))


def test_public_wave_methods(shim, golden):
    g = golden("waves")
    Sample = importlib.import_module("Sample")
    s = Sample.AnalyticalSample()
    s.myType, s.myName, s.myMaterials = "membrane", "golden", ["A", "B"]
    E = float(g["E"])
    s.delta = [[(E, g["delta"][0]), (40.0, 1.0)], [(E, g["delta"][1])]]
    s.beta = [[(E, g["beta"][0])], [(40.0, 1.0), (E, g["beta"][1])]]
    s.myGeometry = g["t"]
    i_rt, phi_rt, df = s.setWaveRT(g["I"], E, g["phi0"])
    assert df == 0 and rel_l2(i_rt, g["I_rt"]) < 1e-6 and rel_l2(phi_rt, g["phi_rt"]) < 1e-6
    w = s.setWave(g["wave0"], E)
    assert rel_l2(np.abs(w) ** 2, np.abs(g["wave"]) ** 2) < 1e-6
    s.myGeometry = g["t"][0]
    with pytest.raises(Exception, match="wrong nb of dim"):
        s.setWave(g["wave0"], E)
    e = shim.Experiment.__new__(shim.Experiment)
    z, M, pix = g["prop1_cfg"]
    e.exp_dict = {"studyDimensions": np.array(g["wave"].shape), "studyPixelSize": pix}
    wave = g["wave"]
    out = e.wavePropagation(wave, z, E, M)
    assert rel_l2(np.abs(out) ** 2, np.abs(g["prop1"]) ** 2) < TOL
    assert e.wavePropagation(wave, 0, E, M) is wave          # Experiment.py:233-234


def test_main_script_writes_the_output_tree(shim, tmp_path):
    """main.py flow: membraneThickness/, ref/, sample/, propag/, White_, DF, report (main.py:76-113)."""
    import argparse
    from paresis_b200 import workspace
    from paresis_b200.hostio import imageio
    sys.path.insert(0, workspace.SHIM_DIR)
    main = importlib.import_module("main")
    args = argparse.Namespace(experiment="Small_sphere_mono", results=str(tmp_path / "Results"), oversampling=2, points=2,
                              model="RayT", format=".tif", seed=3)
    root = main.run(args)
    importlib.import_module("InputOutput.pagailleIO").wait_for_writes()
    files = sorted(os.path.relpath(os.path.join(dp, f), root) for dp, _, fs in os.walk(root) for f in fs)
    assert sum(f.startswith("sample/") for f in files) == 2 and sum(f.startswith("ref/") for f in files) == 2
    assert sum(f.startswith("propag/") for f in files) == 1 and sum(f.startswith("membraneThickness/") for f in files) == 2
    assert any(f.startswith("White_") for f in files) and "DF.tif" in files
    img = imageio.read_tiff(os.path.join(root, [f for f in files if f.startswith("sample/")][0]))
    assert img.shape == (96, 128) and img.dtype == np.float32 and img.mean() > 1000
    assert np.array_equal(img, np.round(img))            # Poisson counts
    reports = [f for f in os.listdir(os.path.dirname(os.path.dirname(root))) if f.endswith(".txt")]
    assert len(reports) == 1


@pytest.mark.parametrize("name", ["Fil_Nylon_ID17", "B200_small_poly3"])
def test_positions_in_one_call_match_one_by_one(shim, name):
    """paresis_rt_run_positions (raster + pipeline of several positions, two in flight) gives the images
    and the mean-energy sums of the position-by-position API; membrane offsets from the same seed."""
    from paresis_b200 import geometry
    d = dict(experimentName=name, filepath="unused/", overSampling=2, nbExpPoints=3, simulation_type="RayT",
             expID="t", poissonNoise=False)
    e = shim.Experiment(d)
    mem = e.myMembrane
    dims = e.exp_dict['studyDimensions']
    np.random.seed(77)
    single = []
    for point in range(3):
        mem.myGeometry = []
        mem.getMyGeometry(dims, mem.membranePixelSize, 2, point, 3)
        thick = np.array(mem.myGeometry[0])
        e.exp_dict['meanEnergy'] = 0
        res = e.computeSampleAndReferenceImages_RT(point)
        single.append((thick, [np.array(r) for r in res[:4]], e.exp_dict['meanEnergy']))
    thresholds = e.myDetector.det_param["myBinsThersholds"]
    scene = e._scene(thresholds, per_position_membrane=True)
    plan = geometry.MembranePlan(mem, dims[0], dims[1], mem.membranePixelSize)
    np.random.seed(77)
    offsets = [plan.draw_offsets() for _ in range(3)]
    eng = e._get_engine()
    # (streams, positions per launch): positions on 1-3 streams, and several positions per kernel launch (blockIdx.z)
    for slots, per in ((1, 0), (2, 0), (3, 0), (3, 3), (4, 2)):
        out = eng.compute_rt_positions(scene, plan, offsets, [0, 1, 2], n_slots=slots, per_launch=per)
        torch.cuda.synchronize()
        eng.check_flag()
        energies = np.array([en for en, _ in scene.spectrum])
        for p in range(3):
            thick, imgs, mean_e = single[p]
            assert rel_l2(out["thickness"][p].cpu().numpy(), thick) < 1e-6
            assert rel_l2(out["sample"][p].cpu().numpy(), imgs[0]) < 1e-6
            assert rel_l2(out["reference"][p].cpu().numpy(), imgs[1]) < 1e-6
            sums = out["sums"][p].cpu().numpy()
            assert abs(np.dot(sums, energies) / sums.sum() / mean_e - 1) < 1e-6
        assert out["firsts"] == [0]
        assert rel_l2(out["propag"][0].cpu().numpy(), single[0][1][2]) < 1e-6
        assert rel_l2(out["white"][0].cpu().numpy(), single[0][1][3]) < 1e-6


def test_dark_field_branch(shim, golden):
    """fastRefractionDF / refraction(..., darkField) / setWaveRT with the Lung model, and one end-to-end
    position with a scattering sample, against the unmodified reference."""
    r2 = importlib.import_module("refractionFileNumba2")
    g = golden("darkfield")
    pix, z, E, M = g["params"]
    for tag in ("narrow", "wide"):
        out, dx, dy = r2.fastRefractionDF(g["I"].copy(), g["phi"], z, E, M, pix, g["df_" + tag])
        assert out.dtype == np.float64 and dx.shape == g["Dx_" + tag].shape
        assert rel_l2(dx, g["Dx_" + tag]) < 1e-6
        assert rel_l2(out, g["out_" + tag]) < TOL, tag
    Sample = importlib.import_module("Sample")
    s = Sample.AnalyticalSample()
    s.myType, s.myName, s.myMaterials = "sample_of_interest", "probe", ["Lung", "PMMA"]
    s.delta = [[(E, g["sw_db"][0][0])], [(E, g["sw_db"][0][1])]]
    s.beta = [[(E, g["sw_db"][1][0])], [(E, g["sw_db"][1][1])]]
    s.myGeometry = g["t"]
    i_out, phi_out, df = s.setWaveRT(g["I"], E, g["phi"])
    assert rel_l2(i_out, g["sw_I"]) < 1e-6 and rel_l2(phi_out, g["sw_phi"]) < 1e-6 and rel_l2(df, g["sw_df"]) < 1e-6
    # end to end: the sphere sample made of Lung
    ge = golden("e2e_rt_lung")
    d = dict(experimentName="Small_sphere_mono", filepath="unused/", overSampling=2, nbExpPoints=1, simulation_type="RayT",
             expID="t", poissonNoise=False)
    e = shim.Experiment(d)
    smp = e.mySampleofInterest
    smp.myMaterials = ["Lung"]
    smp.delta, smp.beta = [], []
    smp.getDeltaBeta(e.mySource.mySpectrum)
    assert np.allclose([v for _, v in smp.delta[0]], ge["sample_db"][:, 1], rtol=1e-12)
    np.random.seed(int(ge["membrane_seed"]))
    e.myMembrane.myGeometry = []
    e.myMembrane.getMyGeometry(e.exp_dict['studyDimensions'], e.myMembrane.membranePixelSize, 2, 0, 1)
    res = e.computeSampleAndReferenceImages_RT(0)
    assert len(res) == 7
    for k, name in enumerate(("sample", "reference", "propag", "white")):
        assert rel_l2(res[k], ge[name]) < TOL, name
    assert rel_l2(res[6][::4, ::4], ge["df_s4"]) < 1e-6
    assert abs(e.exp_dict["meanEnergy"] / float(ge["mean_energy"]) - 1) < 1e-5
    # the sample image really differs from a non-scattering one
    assert rel_l2(res[0], ge["reference"]) > 1e-3


def test_float32_results_option(shim):
    """exp_dict['resultDtype'] = 'float32' (what the drop-in main.py asks for): same counts, half the bytes."""
    outs = {}
    for dt in ("float64", "float32"):
        d = dict(experimentName="Small_sphere_mono", filepath="unused/", overSampling=2, nbExpPoints=1, simulation_type="RayT",
                 expID="t", poissonNoise=False, resultDtype=dt)
        e = shim.Experiment(d)
        np.random.seed(4)
        e.myMembrane.myGeometry = []
        e.myMembrane.getMyGeometry(e.exp_dict['studyDimensions'], e.myMembrane.membranePixelSize, 2, 0, 1)
        outs[dt] = e.computeSampleAndReferenceImages_RT(0)[:4]
    for a, b in zip(outs["float64"], outs["float32"]):
        # (two runs: the fp32 accumulation order of the few rays that bypass the tiles is not reproducible)
        assert a.dtype == np.float64 and b.dtype == np.float32 and rel_l2(b, a) < 1e-6


def test_reference_main_script_runs_unmodified(golden, tmp_path, monkeypatch):
    """PARESIS's OWN main.py (tests/golden/reference_main_py.txt: a verbatim copy made by oracle/make_golden.py), run with
    runpy from a CodePython-shaped directory whose modules are the drop-in's: same output tree, same image files, same
    images (noise off on both sides) and the same saveAllParameters report (Experiment.py:530-607) as the reference
    produced from the same script and seed.  main.py:58-113."""
    import re
    import runpy
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from paresis_b200 import workspace
    from paresis_b200.hostio import imageio
    g = golden("main_script")
    ws = workspace.make_workspace(str(tmp_path / "CodePython"))
    results = tmp_path / "Results" / "Fil_Nylon_ID17"          # main.py writes to ../Results/Fil_Nylon_ID17/
    results.mkdir(parents=True)
    monkeypatch.setenv("PARESIS_B200_POISSON", "0")
    monkeypatch.chdir(ws)
    monkeypatch.syspath_prepend(workspace.SHIM_DIR)
    for name in ("Experiment", "Sample", "Detector", "Source"):      # fresh module state, as a new interpreter would have
        sys.modules.pop(name, None)
    script = os.path.join(os.path.dirname(__file__), "golden", "reference_main_py.txt")
    np.random.seed(77)
    runpy.run_path(script, run_name="__main__")
    importlib.import_module("InputOutput.pagailleIO").wait_for_writes()      # save_image is asynchronous
    report = [f for f in os.listdir(results) if f.endswith(".txt")]
    assert len(report) == 1
    exp_id = report[0][len("Fil_Nylon_ID17_"):-4]
    root = results / ("RayTracing_" + exp_id)
    tree = sorted(os.path.relpath(os.path.join(dp, d), results).replace(exp_id, "<ID>") + "/" for dp, ds, _ in os.walk(root) for d in ds)
    assert tree == list(g["tree"])
    files = sorted(os.path.relpath(os.path.join(dp, f), results).replace(exp_id, "<ID>") for dp, _, fs in os.walk(root) for f in fs)
    assert files == sorted(g["image_names"])
    for k, name in enumerate(g["image_names"]):
        img = imageio.read_tiff(str(results / name.replace("<ID>", exp_id)))
        want = g["img_%d" % k]
        assert img.dtype == np.float32 and img.shape == want.shape, name
        if want.any():
            assert rel_l2(img, want) < TOL, name
        else:
            assert not img.any(), name
    # the report: everything but the wall-clock line
    got = open(results / report[0]).read().replace(exp_id, "<ID>")
    strip = lambda t: re.sub(r"Entire computing time: [0-9.e+-]+s", "Entire computing time: Xs", t)
    got_lines, want_lines = strip(got).splitlines(), strip(str(g["report"])).splitlines()
    assert len(got_lines) == len(want_lines)
    for a, b in zip(got_lines, want_lines):
        if a.strip().startswith("meanEnergy:"):
            # (m * E) / m in floating point: 52.0 or 52.00000000000001 depending on the last bit of the image mean m
            assert b.strip().startswith("meanEnergy:") and a.split()[-1] == b.split()[-1] == "keV"
            assert abs(float(a.split()[1]) / float(b.split()[1]) - 1) < 1e-12
        else:
            assert a == b


@pytest.mark.parametrize("model", ["RayT", "Fresnel"])
def test_air_plate_and_scintillator_factors(shim, golden, model):
    """Air volume (Experiment.py:452-453), scintillator efficiency (:456-459; Fresnel :326-333) and detector protection
    plate (:478-480), end to end against the unmodified reference: experiment Small_sphere_air_plate_CsI."""
    g = golden("e2e_air_plate_csi_" + model.lower())
    d = dict(experimentName="Small_sphere_air_plate_CsI", filepath="unused/", overSampling=2, nbExpPoints=2, simulation_type=model,
             expID="t", poissonNoise=False)
    e = shim.Experiment(d)
    assert not e.exp_dict["inVacuum"] and e.myPlate is not None
    assert np.allclose(e.myAirVolume.myThickness, float(g["air_thickness_um"]))
    assert np.allclose([v for _, v in e.myAirVolume.delta[0]], g["air_db"][:, 1], rtol=1e-12)
    assert np.allclose([v for _, v in e.myAirVolume.beta[0]], g["air_db"][:, 2], rtol=1e-12)
    assert np.allclose([v for _, v in e.myPlate.beta[0]], g["plate_db"][:, 2], rtol=1e-12)
    assert np.allclose(np.array(e.myDetector.beta, dtype=float), g["scintillator_beta"], rtol=1e-12)
    assert np.allclose(np.array(e.myDetector.mySpectralEfficiency, dtype=float), g["efficiency"], rtol=1e-12)
    for point in (0, 1):
        np.random.seed(300 + point)
        e.myMembrane.myGeometry = []
        e.myMembrane.getMyGeometry(e.exp_dict['studyDimensions'], e.myMembrane.membranePixelSize, 2, point, 2)
        res = e.computeSampleAndReferenceImages_RT(point) if model == "RayT" else e.computeSampleAndReferenceImages_Fresnel(point)
        assert rel_l2(res[0], g["sample_p%d" % point]) < TOL
        assert rel_l2(res[1], g["reference_p%d" % point]) < TOL
        if point == 0:
            assert rel_l2(res[2], g["propag_p0"]) < TOL
            assert rel_l2(res[3], g["white_p0"]) < TOL
            # the factors really bite: the white field is far below the shot count
            assert res[3].mean() < 0.8 * e.exp_dict["meanShotCount"]
    assert abs(e.exp_dict["meanEnergy"] / float(g["mean_energy"]) - 1) < 1e-5
