"""Size-independent properties at BASELINE.json's full grids (4096^2, 8192^2): the oracle cannot run there in
seconds, so the CUDA path is checked through conservation, linearity, kernel-against-kernel agreement and
statistics.  Every call goes through the C ABI."""
import importlib
import os

import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from paresis_b200 import _cabi
    return _cabi


def _membrane_like(n, period=37.0):
    """Thickness of touching sphere caps (torn gradient field), built on the device."""
    i = torch.arange(n, device="cuda", dtype=torch.float32)
    u = torch.remainder(i, period) / period - 0.5
    d2 = 0.22 - u[:, None] ** 2 - u[None, :] ** 2
    return (4e-4 * torch.sqrt(torch.clamp(d2, min=0.0))).contiguous()


@pytest.mark.parametrize("n", [4096, 8192])
def test_hops_conserve_and_scale_at_full_size(abi, n):
    from paresis_b200 import hostmath as hm
    t = _membrane_like(n)
    g = -5.97e-7 * hm.refraction_gradient_scale(3.6, 1.0254, 2.9256)
    i0 = 7500.0
    margin = 64                                   # rays of the inner region cannot leave the frame
    inten = torch.zeros((n, n), device="cuda")
    inten[margin:-margin, margin:-margin] = i0 * (0.75 + 0.5 * torch.rand((n - 2 * margin, n - 2 * margin), device="cuda"))
    total = inten.double().sum().item()
    layers = [(t, g, g, 0.0)]
    tile_s = torch.zeros((n, n), device="cuda"); tile_r = torch.zeros((n, n), device="cuda")
    ssum = torch.zeros(1, device="cuda", dtype=torch.float64)
    abi.refract_layers(inten, 0.0, layers, tile_s, tile_r, sum_ref=ssum, intensity_scale=i0)
    # conservation: nothing leaves the frame, both beams keep the total; the running sum agrees
    assert abs(tile_s.double().sum().item() / total - 1) < 2e-6
    assert abs(tile_r.double().sum().item() / total - 1) < 2e-6
    assert abs(ssum.item() / total - 1) < 2e-6
    # the fixed-point tile kernel against the fp32 direct-to-L2 kernel
    direct = torch.zeros((n, n), device="cuda")
    abi.refract_layers(inten, 0.0, [(t, g, 0.0, 0.0)], direct)
    assert rel_l2(tile_s[::7, ::5].cpu().numpy(), direct[::7, ::5].cpu().numpy()) < 3e-6
    del direct
    # linearity in the intensity (the displacement does not depend on it)
    twice = torch.zeros((n, n), device="cuda")
    abi.refract_layers(2 * inten, 0.0, [(t, g, 0.0, 0.0)], twice, intensity_scale=2 * i0)
    assert rel_l2(twice[::7, ::5].cpu().numpy(), 2 * tile_s[::7, ::5].cpu().numpy()) < 3e-6
    # the field really moves intensity around
    assert rel_l2(tile_s[margin:-margin:7, margin:-margin:5].cpu().numpy(), inten[margin:-margin:7, margin:-margin:5].cpu().numpy()) > 0.05


@pytest.mark.parametrize("n", [4096, 8192])
def test_detector_at_full_size(abi, n):
    from paresis_b200 import hostmath as hm
    det = n // 2
    src, psf = torch.as_tensor(hm.gaussian_1d(0.18), device="cuda", dtype=torch.float32), \
        torch.as_tensor(hm.gaussian_1d(1.2), device="cuda", dtype=torch.float32)
    # a flat image stays flat: 4 source pixels per detector pixel, both kernels normalised, reflect padding
    flat = torch.full((n, n), 7500.0, device="cuda")
    out = torch.empty((det, det), device="cuda")
    abi.detect_counts(flat, 2, det, det, src, psf, None, out, False)
    assert abs(out.min().item() / 30000.0 - 1) < 1e-5 and abs(out.max().item() / 30000.0 - 1) < 1e-5
    # Poisson counts around it: mean and variance of 4-16 M draws, integers, two images in one launch
    a = torch.empty((det, det), device="cuda"); b = torch.empty((det, det), device="cuda")
    abi.detect_counts_multi([flat, flat], 2, det, det, src, psf, None, [a, b], True, 3, [10, 11])
    for c in (a, b):
        m, v = c.double().mean().item(), c.double().var().item()
        npx = det * det
        assert abs(m - 30000.0) < 6 * np.sqrt(30000.0 / npx) + 0.05 and abs(v / 30000.0 - 1) < 6 * np.sqrt(2.0 / npx) + 1e-3
        assert torch.equal(c, c.round())
    assert not torch.equal(a, b)


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from paresis_b200 import workspace
    ws = workspace.make_workspace(str(tmp_path_factory.mktemp("ws_full")))
    old = os.getcwd()
    workspace.enter(ws)
    yield importlib.import_module("Experiment")
    os.chdir(old)


@pytest.mark.parametrize("name,n,energies", [("B200_4096_poly64", 4096, 64), ("B200_8192_poly128", 8192, 128)])
def test_polychromatic_configs_run_at_full_size(shim, name, n, energies):
    """BASELINE.json configs 3 and 5: one membrane position with the whole spectrum, noise off.  Checks the
    bookkeeping the reference does around the hops: shot count recovered in the reference image, the sample
    only removes intensity, mean energy inside the spectrum, propagation / white images only at position 0."""
    d = dict(experimentName=name, filepath="unused/", overSampling=2, nbExpPoints=2, simulation_type="RayT",
             expID="t", poissonNoise=False, returnDisplacement=False)
    e = shim.Experiment(d)
    assert tuple(e.exp_dict['studyDimensions']) == (n, n) and len(e.mySource.mySpectrum) == energies
    mem = e.myMembrane
    np.random.seed(5)
    for point in (0, 1):          # position 0 closes the last detector bin in place (Experiment.py:429), as main.py relies on
        mem.myGeometry = []
        mem.getMyGeometry(e.exp_dict['studyDimensions'], mem.membranePixelSize, 2, point, 2)
        e.exp_dict['meanEnergy'] = 0
        sample, ref, propag, white = e.computeSampleAndReferenceImages_RT(point)[:4]
        if point == 0:
            assert propag.any() and abs(white[0, 64:-64, 64:-64].mean() / e.exp_dict['meanShotCount'] - 1) < 1e-4
    assert sample.shape == (1, n // 2, n // 2) and not propag.any() and not white.any()
    lo, hi = e.mySource.mySpectrum[0][0], e.mySource.mySpectrum[-1][0]
    assert lo < e.exp_dict['meanEnergy'] < hi
    # the membrane attenuates and redistributes, it does not create intensity: mean counts below the shot count
    inner = ref[0, 64:-64, 64:-64]
    assert 0.2 * e.exp_dict['meanShotCount'] < inner.mean() < e.exp_dict['meanShotCount']
    assert inner.std() / inner.mean() > 0.02                       # speckle
    assert sample[0, 64:-64, 64:-64].sum() < inner.sum()            # the fibre absorbs
    assert np.isfinite(sample).all() and (sample >= 0).all()


@pytest.mark.parametrize("n", [4096, 8192])
def test_fresnel_against_refraction_model_at_full_size(shim, n, capsys):
    """BASELINE.json config 4: the Fresnel propagator and the ray-tracing model on the same membrane and
    sample, noise off.  The two models legitimately differ in the fine speckle structure (PARESIS UserGuide,
    'Additional remarks'), so this is a physics cross-check, not a parity gate: same flux, same large-scale
    image, correlated speckle; the relative L2 distance is printed for the record."""
    out = {}
    for model in ("RayT", "Fresnel"):
        d = dict(experimentName="B200_%d_mono" % n, filepath="unused/", overSampling=2, nbExpPoints=1,
                 simulation_type=model, expID="t", poissonNoise=False, returnDisplacement=False)
        e = shim.Experiment(d)
        assert tuple(e.exp_dict['studyDimensions']) == (n, n)
        mem = e.myMembrane
        np.random.seed(11)
        mem.myGeometry = []
        mem.getMyGeometry(e.exp_dict['studyDimensions'], mem.membranePixelSize, 2, 0, 1)
        res = e.computeSampleAndReferenceImages_RT(0) if model == "RayT" else e.computeSampleAndReferenceImages_Fresnel(0)
        out[model] = [np.asarray(r[0], dtype=np.float64) for r in res[:4]]
        del e
    for k, name in enumerate(("sample", "reference", "propagation", "white")):
        a, b = out["RayT"][k][64:-64, 64:-64], out["Fresnel"][k][64:-64, 64:-64]
        assert np.isfinite(b).all()
        assert abs(b.mean() / a.mean() - 1) < 0.02, name                       # same flux
        if name != "white":
            # same image once the speckle grains (~17 detector pixels) are averaged over 64 x 64 pixels
            m = (a.shape[0] // 64) * 64
            ca = a[:m, :m].reshape(m // 64, 64, m // 64, 64).mean(axis=(1, 3))
            cb = b[:m, :m].reshape(m // 64, 64, m // 64, 64).mean(axis=(1, 3))
            assert rel_l2(cb, ca) < 0.05, name
        with capsys.disabled():
            print("\n[config 4, %d^2] %s: rel L2 Fresnel vs RayT = %.3f, correlation = %.3f"
                  % (n, name, rel_l2(b, a), np.corrcoef(a.ravel()[::17], b.ravel()[::17])[0, 1] if a.std() > 0 else 1.0))
