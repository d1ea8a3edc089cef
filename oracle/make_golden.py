"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py

Every fixture stores the seeded inputs next to what the reference returned for them, so the
oracle (oracle/paresis_oracle.py) and the CUDA path can be checked against the reference on
a box where the reference itself is absent.  The reference has no tests or golden vectors
of its own: these files are the pins.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

REPO = os.path.dirname(HERE)
GOLD = os.path.join(REPO, "tests", "golden")


def smooth_field(rng, shape, cells=6):
    """Band-limited random field in [-1, 1] (sum of a few random cosines)."""
    x = np.linspace(0, 1, shape[0])[:, None]
    y = np.linspace(0, 1, shape[1])[None, :]
    f = np.zeros(shape)
    for _ in range(12):
        kx, ky = rng.uniform(-cells, cells, 2)
        f += rng.uniform(0.3, 1.0) * np.cos(2 * np.pi * (kx * x + ky * y) + rng.uniform(0, 2 * np.pi))
    return f / np.abs(f).max()


def save(name, **arrays):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("  wrote %-28s %7.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024))


def golden_splat_kernel(ref):
    """fastloopNumba called directly (refractionFileNumba2.py:198-263): known answers of
    SURVEY App. B-3 plus random frames, edge quirks included."""
    loop = ref["refractionFileNumba2"].fastloopNumba
    out = {}
    # single unit ray at (4,4) on a 9x9 frame
    shifts = [1.0, -1.0, -0.25, 0.25, -1.25, 2.0, -2.0, 1e-13, 0.5, -0.5]
    ka = []
    for dxv in shifts:
        for dyv in (0.0, 0.75, -1.5):
            I = np.zeros((9, 9)); I[4, 4] = 1.0
            Dx = np.zeros((9, 9)); Dx[4, 4] = dxv
            Dy = np.zeros((9, 9)); Dy[4, 4] = dyv
            res = loop(9, 9, I, np.zeros((9, 9)), Dy, Dx, Dx.astype(int), Dy.astype(int))
            ka.append(np.concatenate(([dxv, dyv], res.ravel())))
    out["known_answers"] = np.array(ka)
    # the edge quirk: ray on the last row of an unpadded 5x5 frame
    I = np.zeros((5, 5)); I[4, 2] = 1.0
    Dx = np.zeros((5, 5)); Dy = np.zeros((5, 5)); Dy[4, 2] = 0.5
    out["edge_quirk"] = loop(5, 5, I, np.zeros((5, 5)), Dy, Dx, Dx.astype(int), Dy.astype(int))
    rng = np.random.default_rng(11)
    for tag, amp in (("a", 0.9), ("b", 6.0), ("c", 45.0)):
        I = rng.uniform(0.5, 2.0, (37, 41))
        Dx = rng.normal(0, amp / 2, (37, 41)); Dy = rng.normal(0, amp / 2, (37, 41))
        Dx[rng.random((37, 41)) < 0.2] = 0.0
        Dy[rng.random((37, 41)) < 0.2] = 0.0
        Dx[3, 3] = 1.0; Dy[3, 3] = -1.0; Dx[5, 5] = -3.0; Dy[5, 6] = 2.0
        res = loop(37, 41, I.copy(), np.zeros((37, 41)), Dy, Dx, Dx.astype(int), Dy.astype(int))
        out["I_" + tag], out["Dx_" + tag], out["Dy_" + tag], out["out_" + tag] = I, Dx, Dy, res
    save("splat_kernel", **out)


def golden_fast_refraction(ref):
    """fastRefraction v2 and v1 (refractionFileNumba2.py:25-86, refractionFileNumba.py:11-68)."""
    v2 = ref["refractionFileNumba2"].fastRefraction
    v1 = ref["refractionFileNumba"].fastRefraction
    rng = np.random.default_rng(5)
    shape = (97, 113)
    pix, z, E, M = 2.9, 3.6, 52.0, 1.0254
    out = dict(params=np.array([pix, z, E, M]))
    I = rng.uniform(0.5, 1.5, shape) * 7500.0
    base = smooth_field(rng, shape, cells=5)
    for tag, amp in (("sub", 1.3), ("mid", 14.0), ("far", 150.0), ("huge", 3000.0)):
        phi = base * amp
        phi[10:14, 20:30] = phi[10, 20]          # flat patch -> exact zeros in D
        res, Dx, Dy = v2(I.copy(), phi.copy(), z, E, M, pix)
        res1, _, _ = v1(I.copy(), phi.copy(), z, E, M, pix)
        out["phi_" + tag], out["out_" + tag], out["Dx_" + tag], out["Dy_" + tag] = phi, res, Dx, Dy
        out["outv1_" + tag] = res1
        print("    fastRefraction %-5s max|D| = %8.2f px  kept %.4f" % (tag, max(np.abs(Dx).max(), np.abs(Dy).max()),
                                                                       res.sum() / I.sum()))
    out["I"] = I
    save("fast_refraction", **out)


def golden_detector(ref):
    """create_gaussian_shape, resize, detection (Detector.py:79-119, 185-220)."""
    D = ref["Detector"]
    sig = np.array([0.036, 0.18, 0.5, 2.5 / 3, 1.2, 1.5, 2.0])
    out = dict(sigmas=sig)
    for k, s in enumerate(sig):
        out["g%d" % k] = D.create_gaussian_shape(float(s))
        assert np.array_equal(out["g%d" % k], ref["refractionFileNumba2"].gaussian_shape(float(s)))
    rng = np.random.default_rng(3)
    img = rng.uniform(0, 10, (60, 90))
    out["resize_in"] = img
    out["resize_2"] = D.resize(img, 30, 45)
    out["resize_3"] = D.resize(img, 20, 30)
    out["resize_id"] = D.resize(img, 60, 90)
    cases = [(2, (64, 48), 0.424, 1.2), (1, (70, 50), 0.0, 0.0), (3, (40, 44), 2.5, 0.0), (2, (50, 50), 0.036 * 2.355, 0.5),
             (2, (40, 40), 30.0, 5.5)]
    with rh.identity_poisson():
        for k, (os_, dims, fwhm, psf) in enumerate(cases):
            det = D.Detector({})
            det.det_param["myDimensions"] = np.array(dims)
            det.det_param["myPSF"] = psf
            im = rng.uniform(0.0, 1.0, (dims[0] * os_, dims[1] * os_)) * 7500 + 100 * smooth_field(rng, (dims[0] * os_, dims[1] * os_))
            out["det%d_in" % k] = im
            out["det%d_cfg" % k] = np.array([os_, dims[0], dims[1], fwhm, psf])
            out["det%d_out" % k] = np.asarray(det.detection(im.copy(), fwhm, {"overSampling": os_}), dtype=np.float64)
    out["n_det"] = np.array(len(cases))
    save("detector", **out)


def golden_waves(ref):
    """setWave / setWaveRT (Sample.py:248-351) and wavePropagation (Experiment.py:219-252)."""
    S = ref["Sample"]
    rng = np.random.default_rng(9)
    shape = (64, 80)
    t = np.array([np.abs(smooth_field(rng, shape)) * 4e-4, np.ones(shape) * 6e-3])
    delta, beta, E = [5.97e-07, 9.85e-08], [5.37e-09, 3.16e-12], 52.0
    samp = S.AnalyticalSample()
    samp.myType = "membrane"; samp.myName = "golden"
    samp.myMaterials = ["A", "B"]
    samp.delta = [[(E, delta[0]), (40.0, 1.0)], [(E, delta[1])]]
    samp.beta = [[(E, beta[0])], [(40.0, 1.0), (E, beta[1])]]
    samp.myGeometry = t
    I = rng.uniform(0.5, 1.5, shape) * 7500
    phi0 = smooth_field(rng, shape) * 3
    i_rt, phi_rt, df = samp.setWaveRT(I, E, phi0)
    assert df == 0
    wave0 = np.sqrt(I) * np.exp(1j * phi0)
    wave = samp.setWave(wave0, E)
    out = dict(t=t, delta=np.array(delta), beta=np.array(beta), E=np.array(E), I=I, phi0=phi0,
               I_rt=i_rt, phi_rt=phi_rt, wave0=wave0, wave=wave)
    Exp = ref["Experiment"].Experiment
    e = Exp.__new__(Exp)
    pix = 2.9256
    e.exp_dict = {"studyDimensions": np.array(shape), "studyPixelSize": pix}
    for k, (z, M) in enumerate(((1.6, 1.0114), (3.6, 1.0254), (0, 1.0))):
        out["prop%d" % k] = np.asarray(e.wavePropagation(wave, z, E, M))
        out["prop%d_cfg" % k] = np.array([z, M, pix])
    save("waves", **out)


def golden_geometry(ref):
    """getMembraneSegmentedFromFile, CreateSampleSphere, CreateSampleCylindre."""
    mem = ref["membrane"].getMembraneSegmentedFromFile
    geom = ref["geom"]

    class M:
        pass

    out = {}
    m = M(); m.myMeanSphereRadius = 50.0; m.myNbOfLayers = 3
    np.random.seed(7)
    state_probe = np.random.get_state()[1][:4].copy()
    g, _ = mem(m, 120, 150, 2.88, 0, 6000.0)
    out["mem0"] = np.array(g); out["mem0_cfg"] = np.array([50.0, 3, 120, 150, 2.88, 6000.0, 7])
    m = M(); m.myMeanSphereRadius = 5.0; m.myNbOfLayers = 2
    np.random.seed(8)
    g, _ = mem(m, 400, 100, 10.0, 0, 4000.0)  # file narrower than the field of view -> stitched along x
    out["mem1"] = np.array(g); out["mem1_cfg"] = np.array([5.0, 2, 400, 100, 10.0, 4000.0, 8])
    out["seed_probe"] = state_probe
    s, _ = geom.CreateSampleSphere("PMMA_sphere", 200, 240, 12.0)
    out["sphere"] = s; out["sphere_cfg"] = np.array([1000.0, 200, 240, 12.0])
    c, _ = geom.CreateSampleCylindre("filNylon", 200, 240, 12.0)
    out["cyl"] = c; out["cyl_cfg"] = np.array([700.0, 30.0, 200, 240, 12.0])
    c, _ = geom.CreateSampleCylindre("filNylon", 160, 160, 9.5)
    out["cyl2"] = c; out["cyl2_cfg"] = np.array([700.0, 30.0, 160, 160, 9.5])
    save("geometry", **out)


def _experiment(ref, sim, sample_name=None, spectrum=None, thresholds=None, psf=None, source_um=None, det_dims=None):
    Exp = ref["Experiment"].Experiment
    Det = ref["Detector"].Detector
    orig_dims = Det.getMyDimensions
    if det_dims is not None:
        Det.getMyDimensions = lambda self, node: np.array(det_dims)
    d = dict(experimentName="Fil_Nylon_ID17", filepath="../Results/Fil_Nylon_ID17/", overSampling=2,
             nbExpPoints=1, simulation_type=sim, expID="golden")
    if sample_name is not None:
        # same experiment entry, other sample of interest: patch the XML lookup result
        orig = Exp.defineCorrectValues

        def patched(self, exp_dict):
            orig(self, exp_dict)
            self.mySampleofInterest.myName = sample_name

        Exp.defineCorrectValues = patched
    try:
        e = Exp(d)
    finally:
        Det.getMyDimensions = orig_dims
        if sample_name is not None:
            Exp.defineCorrectValues = orig
    if psf is not None:
        e.myDetector.det_param["myPSF"] = psf
    if source_um is not None:
        e.mySource.source_dict["mySize"] = source_um
    if spectrum is not None:
        e.mySource.mySpectrum = list(spectrum)
        e.mySource.source_dict["myEnergySampling"] = 10
        for obj in (e.myAirVolume, e.mySampleofInterest, e.myMembrane):
            obj.delta, obj.beta = [], []
            obj.getDeltaBeta(e.mySource.mySpectrum)
        e.myDetector.det_param["myBinsThersholds"] = list(thresholds)
    return e


def _db(obj):
    """{E: (deltas, betas)} as flat arrays: rows = energies, cols = [E, d0.., b0..]."""
    energies = [e for e, _ in obj.delta[0]]
    rows = []
    for k, e in enumerate(energies):
        rows.append([e] + [obj.delta[m][k][1] for m in range(len(obj.delta))] + [obj.beta[m][k][1] for m in range(len(obj.beta))])
    return np.array(rows)


def golden_end_to_end(ref):
    """main.py:63-71 flow at the bundled configuration (N=400), Poisson patched to identity."""
    small = (96, 128)
    runs = [
        ("e2e_rt_cylinder", "RayT", None, None, None, None, None, None),
        ("e2e_rt_sphere", "RayT", "PMMA_sphere", None, None, 1.2, 50.0, small),
        ("e2e_fresnel_sphere", "Fresnel", "PMMA_sphere", None, None, 1.2, 50.0, small),
        ("e2e_rt_poly3", "RayT", "PMMA_sphere", [(30.0, 0.2), (40.0, 0.5), (52.0, 0.3)], [35.0], 1.2, 50.0, small),
    ]
    for name, sim, sample_name, spectrum, thresholds, psf, source_um, dims in runs:
        e = _experiment(ref, sim, sample_name, spectrum, thresholds, psf, source_um, dims)
        st = np.array(e.mySampleofInterest.myGeometry)
        out = dict(sample_t_probe=np.array([st.sum(), st.max(), st[0, 7, 11], st[0, st.shape[1] // 2, st.shape[2] // 2]]),
                   sample_name=np.array(e.mySampleofInterest.myName), membrane_db=_db(e.myMembrane),
                   sample_db=_db(e.mySampleofInterest), spectrum=np.array(e.mySource.mySpectrum, dtype=float),
                   cfg=np.array([e.exp_dict["distSourceToMembrane"], e.exp_dict["distMembraneToObject"],
                                 e.exp_dict["distObjectToDetector"], e.myDetector.det_param["myPixelSize"],
                                 e.exp_dict["overSampling"], e.exp_dict["meanShotCount"], e.mySource.source_dict["mySize"],
                                 e.myDetector.det_param["myPSF"], e.mySource.source_dict["myEnergySampling"]]),
                   thresholds=np.array(thresholds if thresholds else [], dtype=float),
                   det_dims=np.array(e.myDetector.det_param["myDimensions"]))
        for point in (0, 1):
            np.random.seed(100 + point)
            e.myMembrane.myGeometry = []
            e.myMembrane.getMyGeometry(e.exp_dict["studyDimensions"], e.myMembrane.membranePixelSize, 2, point, 2)
            with rh.identity_poisson():
                if sim == "RayT":
                    res = e.computeSampleAndReferenceImages_RT(point)
                else:
                    res = e.computeSampleAndReferenceImages_Fresnel(point)
            mt = np.asarray(e.myMembrane.myGeometry[0], dtype=np.float64)
            out["membrane_probe_p%d" % point] = np.array([mt.sum(), mt.max(), mt[5, 9], mt[mt.shape[0] // 2, mt.shape[1] // 3]])
            out["membrane_seed_p%d" % point] = np.array(100 + point)
            out["support_um"] = np.array(e.myMembrane.myPMMAThickness)
            out["membrane_cfg"] = np.array([e.myMembrane.myMeanSphereRadius, e.myMembrane.myNbOfLayers,
                                            e.myMembrane.membranePixelSize])
            for tag, arr in zip(("sample", "reference", "propag", "white"), res[:4]):
                arr = np.asarray(arr, dtype=np.float64)
                if point == 0 or tag in ("sample", "reference"):
                    out["%s_p%d" % (tag, point)] = arr
                else:
                    assert not arr.any() or tag == "white"
                    out["%s_p%d_sum" % (tag, point)] = np.array(arr.sum())
            if sim == "RayT" and point == 0:
                out["Dx_p0_s8"], out["Dy_p0_s8"] = np.asarray(res[4])[::8, ::8], np.asarray(res[5])[::8, ::8]
        out["mean_energy"] = np.array(e.exp_dict["meanEnergy"])
        save(name, **out)


def golden_darkfield(ref):
    """The dark-field branch: fastRefractionDF (refractionFileNumba2.py:88-196), setWaveRT with the Lung model
    (Sample.py:322-343) and one end-to-end position with a scattering sample."""
    rng = np.random.default_rng(23)
    r2 = ref["refractionFileNumba2"]
    shape = (72, 90)
    pix, z, E, M = 2.9256, 3.6, 52.0, 1.0254
    x = np.arange(shape[0])[:, None]; y = np.arange(shape[1])[None, :]
    phi = 40.0 * smooth_field(rng, shape, cells=3)
    I = 7500.0 * (0.8 + 0.4 * rng.random(shape))
    blob = np.clip(1.0 - ((x - 36.0) ** 2 + (y - 50.0) ** 2) / 24.0 ** 2, 0.0, None)
    to_px = z / (pix * 1e-6 * M)
    out = dict(params=np.array([pix, z, E, M]), I=I, phi=phi)
    for tag, peak_px in (("narrow", 0.9), ("wide", 2.6)):
        df = np.sqrt(blob) * peak_px / to_px                     # radians; zero outside the blob
        if tag == "wide":
            df[5, 7] = 30.0 / to_px                              # above Nx/4 pixels: dropped at :130
        res, dx, dy = r2.fastRefractionDF(I.copy(), phi.copy(), z, E, M, pix, df.copy())
        out["df_" + tag], out["out_" + tag], out["Dx_" + tag] = df, res, dx
    # setWaveRT with a scattering material
    S = ref["Sample"].AnalyticalSample
    smp = object.__new__(S)
    smp.myMaterials, smp.myType, smp.myName = ["Lung", "PMMA"], "sample_of_interest", "probe"
    t = np.stack([1.2e-3 * blob, 4e-4 * np.ones(shape)])
    smp.myGeometry = t
    smp.delta, smp.beta = [[(E, 2.1e-7)], [(E, 9.9e-8)]], [[(E, 1.3e-10)], [(E, 4.4e-11)]]
    i_out, phi_out, new_df = smp.setWaveRT(I.copy(), E, phi.copy(), 0)
    out.update(t=t, sw_I=i_out, sw_phi=phi_out, sw_df=new_df, sw_db=np.array([[2.1e-7, 9.9e-8], [1.3e-10, 4.4e-11]]))
    save("darkfield", **out)

    # end to end: the bundled experiment with the sphere sample made of Lung
    e = _experiment(ref, "RayT", "PMMA_sphere", None, None, 1.2, 50.0, (96, 128))
    e.mySampleofInterest.myMaterials = ["Lung"]
    # (upstream resolves "Lung" through xraylib's compound parser, absent here: take the coefficients from the
    # reference's own TablesDeltaBeta.xls, as exported to delta_beta_tables.npz -- the values the shim uses too)
    sys.path.insert(0, os.path.dirname(HERE))
    from paresis_b200.hostio import tables
    energies = [en for en, _ in e.mySource.mySpectrum]
    db = tables.interpolate("Lung", energies)
    e.mySampleofInterest.delta = [[(en, float(db[k][0])) for k, en in enumerate(energies)]]
    e.mySampleofInterest.beta = [[(en, float(db[k][1])) for k, en in enumerate(energies)]]
    np.random.seed(100)
    e.myMembrane.myGeometry = []
    e.myMembrane.getMyGeometry(e.exp_dict["studyDimensions"], e.myMembrane.membranePixelSize, 2, 0, 1)
    with rh.identity_poisson():
        res = e.computeSampleAndReferenceImages_RT(0)
    st = np.array(e.mySampleofInterest.myGeometry)
    save("e2e_rt_lung", sample=np.asarray(res[0], float), reference=np.asarray(res[1], float), propag=np.asarray(res[2], float),
         white=np.asarray(res[3], float), df_s4=np.asarray(res[6], float)[::4, ::4], membrane_seed=np.array(100),
         sample_db=_db(e.mySampleofInterest), sample_t_probe=np.array([st.sum(), st.max()]),
         mean_energy=np.array(e.exp_dict["meanEnergy"]))


def golden_phantoms(ref):
    """CreateSampleSpheresInCylinder / CreateSampleSpheresInParallelepiped (createSampGeom.py:110-260)."""
    geom = ref["geom"]
    out = {}
    for tag, fn, dx, dy, pix in (("cyl_a", geom.CreateSampleSpheresInCylinder, 420, 300, 10.0),
                                 ("cyl_b", geom.CreateSampleSpheresInCylinder, 500, 260, 8.3),
                                 ("par_a", geom.CreateSampleSpheresInParallelepiped, 300, 300, 12.0),
                                 ("par_b", geom.CreateSampleSpheresInParallelepiped, 260, 340, 9.7)):
        g, params = fn("x", dx, dy, pix)
        out[tag] = np.asarray(g, dtype=np.float64)
        out[tag + "_cfg"] = np.array([dx, dy, pix])
    save("phantoms", **out)


def golden_air_plate_scintillator(ref):
    """Air volume, detector protection plate and scintillator efficiency (Experiment.py:451-459, :478-480, :320-333 for
    Fresnel; Detector.py:131-182): the bundled experiment with the sphere sample, taken out of vacuum, with the C_plate of
    the reference's Samples.xml in front of the detector and a 400 um CsI scintillator -- both models, two positions."""
    for sim in ("RayT", "Fresnel"):
        e = _experiment(ref, sim, "PMMA_sphere", None, None, 1.2, 50.0, (96, 128))
        e.exp_dict["inVacuum"] = False
        plate = ref["Sample"].AnalyticalSample()
        plate.myName = "C_plate"
        plate.defineCorrectValuesSample()
        plate.getDeltaBeta(e.mySource.mySpectrum)
        plate.getMyGeometry(e.exp_dict["studyDimensions"], e.exp_dict["studyPixelSize"], e.exp_dict["overSampling"])
        e.myPlate = plate
        e.myDetector.det_param["myScintillatorMaterial"] = "CsI"
        e.myDetector.det_param["myScintillatorThickness"] = 400.0
        e.myDetector.getBeta(e.mySource.mySpectrum)
        e.myDetector.getSpectralEfficiency()
        out = dict(air_db=_db(e.myAirVolume), plate_db=_db(e.myPlate), air_thickness_um=np.array(e.myAirVolume.myThickness),
                   scintillator_beta=np.array(e.myDetector.beta, dtype=float), efficiency=np.array(e.myDetector.mySpectralEfficiency, dtype=float))
        for point in (0, 1):
            np.random.seed(300 + point)
            e.myMembrane.myGeometry = []
            e.myMembrane.getMyGeometry(e.exp_dict["studyDimensions"], e.myMembrane.membranePixelSize, 2, point, 2)
            with rh.identity_poisson():
                res = e.computeSampleAndReferenceImages_RT(point) if sim == "RayT" else e.computeSampleAndReferenceImages_Fresnel(point)
            for tag, arr in zip(("sample", "reference", "propag", "white"), res[:4]):
                if point == 0 or tag in ("sample", "reference"):
                    out["%s_p%d" % (tag, point)] = np.asarray(arr, dtype=np.float64)
        out["mean_energy"] = np.array(e.exp_dict["meanEnergy"])
        save("e2e_air_plate_csi_" + sim.lower(), **out)


def golden_spectrum_xls(ref):
    """Source.setMySpectrum on the reference's own Sources/W_50kVp.xls (Source.py:131-240), source 'simap2' of its
    Sources.xml, at several energy samplings.  The sheet itself (28 KB of measured data, not code) is copied next to
    the shim's parameter files so that the drop-in reads the same bytes."""
    import shutil
    Source = ref["Source"].Source
    out = {}
    for sampling in (4.0, 2.0, 1.0):
        s = Source()
        s.myName = "simap2"
        s.defineCorrectValuesSource()
        s.source_dict["myEnergySampling"] = sampling
        s.setMySpectrum()
        out["spectrum_%g" % sampling] = np.array(s.mySpectrum, dtype=float)
    save("spectrum_xls", **out)
    dst = os.path.join(REPO, "paresis_b200", "CodePython", "Sources", "W_50kVp.xls")
    shutil.copyfile(os.path.join(rh.REFERENCE_ROOT, "Sources", "W_50kVp.xls"), dst)
    os.chmod(dst, 0o644)


def golden_main_script(ref):
    """The reference's main.py, unmodified, through runpy in the reference tree (Poisson patched to identity): the output
    tree it writes, the images it saves and the report of saveAllParameters (Experiment.py:530-607).  The script text is
    stored verbatim as a fixture (tests/golden/reference_main_py.txt) so that the GPU test can run THE SAME script against
    the drop-in modules -- that is the "main.py is a drop-in" claim, tested."""
    import runpy
    import shutil
    src = os.path.join(rh.REFERENCE_ROOT, "main.py")
    results = os.path.abspath(os.path.join(ref["scratch"], "..", "Results", "Fil_Nylon_ID17"))
    before = set(os.listdir(results))
    rh.SAVED_IMAGES.clear()
    np.random.seed(77)
    with rh.identity_poisson():
        runpy.run_path(src, run_name="__main__")
    new = sorted(set(os.listdir(results)) - before)
    report = [f for f in new if f.endswith(".txt")][0]
    exp_id = report[len("Fil_Nylon_ID17_"):-4]
    tree = []
    for root, dirs, files in os.walk(os.path.join(results, "RayTracing_" + exp_id)):
        for d in dirs:
            tree.append(os.path.relpath(os.path.join(root, d), results).replace(exp_id, "<ID>") + "/")
    out = dict(tree=np.array(sorted(tree)), report=np.array(open(os.path.join(results, report)).read().replace(exp_id, "<ID>")))
    names = []
    for fn, arr in rh.SAVED_IMAGES.items():
        key = os.path.relpath(fn, "../Results/Fil_Nylon_ID17").replace(exp_id, "<ID>")
        names.append(key)
        out["img_%d" % (len(names) - 1)] = np.asarray(arr, dtype=np.float32)
    out["image_names"] = np.array(names)
    save("main_script", **out)
    shutil.copyfile(src, os.path.join(GOLD, "reference_main_py.txt"))
    os.chmod(os.path.join(GOLD, "reference_main_py.txt"), 0o644)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    ref = rh.load_reference()
    only = sys.argv[1:]
    for fn in (golden_splat_kernel, golden_fast_refraction, golden_detector, golden_waves, golden_geometry,
               golden_end_to_end, golden_darkfield, golden_phantoms, golden_air_plate_scintillator, golden_spectrum_xls,
               golden_main_script):
        if not only or fn.__name__ in only:
            print(fn.__name__)
            fn(ref)
