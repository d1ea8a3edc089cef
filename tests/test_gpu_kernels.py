"""GPU parity tests, kernel by kernel, through the C ABI (paresis_b200._cabi).

Oracle = oracle/paresis_oracle.py (fp64 CPU restatement, pinned to the reference by
tests/test_oracle_golden.py) and the committed reference goldens themselves.  Tolerance: the
north-star bound, 1e-5 relative L2 for noise-free fp32 results; bit-level properties (known
answers, conservation) are checked where the domain offers them.
"""
import numpy as np
import pytest

import paresis_oracle as po
from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def abi():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from paresis_b200 import _cabi
    return _cabi


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=dtype).contiguous()


def run_splat(abi, I, Dx, Dy, margin, variant):
    out = torch.zeros(I.shape, device="cuda", dtype=torch.float32)
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    abi.splat(dev(I), dev(Dx), dev(Dy), out, margin=margin, variant=variant, flag=flag)
    return out.cpu().numpy().astype(np.float64), int(flag.item())


# variant 3 (fixed-point shared-memory tiles) quantises deposits to 2^-19 ... 2^-18 of the mean ray of a tile;
# variants 4 / 5 (owner-computes rolling strips, out = / out +=) to 2^-22 of twice the mean ray of the image
SPLAT_TOL = {0: 1.0, 1: 1.0, 2: 1.0, 3: 10.0, 4: 2.0, 5: 2.0}
VARIANTS = [0, 1, 2, 3, 4, 5]


@pytest.mark.parametrize("variant", VARIANTS)
def test_splat_known_answers(abi, golden, variant):
    g = golden("splat_kernel")
    for row in g["known_answers"]:
        dxv, dyv, want = row[0], row[1], row[2:].reshape(9, 9)
        I = np.zeros((9, 9)); I[4, 4] = 1.0
        Dx = np.zeros((9, 9)); Dx[4, 4] = dxv
        Dy = np.zeros((9, 9)); Dy[4, 4] = dyv
        got, flag = run_splat(abi, I, Dx, Dy, 0, variant)
        assert flag == 0
        assert np.allclose(got, want, rtol=0, atol=1e-7), (dxv, dyv)
    I = np.zeros((5, 5)); I[4, 2] = 1.0
    Dx = np.zeros((5, 5)); Dy = np.zeros((5, 5)); Dy[4, 2] = 0.5
    got, _ = run_splat(abi, I, Dx, Dy, 0, variant)
    assert np.array_equal(got, g["edge_quirk"])  # the reference's edge quirk, bit for bit


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_splat_reference_frames(abi, golden, variant, tag):
    g = golden("splat_kernel")
    got, flag = run_splat(abi, g["I_" + tag], g["Dx_" + tag], g["Dy_" + tag], 0, variant)
    assert flag == 0
    assert rel_l2(got, g["out_" + tag]) < 2e-6 * max(1.0, SPLAT_TOL[variant] / 3)  # inputs are fp64 goldens rounded to fp32 (|D| up to ~100 px)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("shape,amp,margin", [((257, 300), 0.7, 15), ((64, 1030), 9.0, 15), ((513, 96), 30.0, 0),
                                               ((40, 33), 100.0, 10), ((3, 3), 1.0, 0), ((1, 70), 2.0, 0),
                                               ((700, 1100), 2.5, 15)])
def test_splat_vs_oracle(abi, variant, shape, amp, margin):
    rng = np.random.default_rng(hash((shape, margin)) % 1000)
    x = np.linspace(0, 6, shape[0])[:, None]
    y = np.linspace(0, 6, shape[1])[None, :]
    Dx = (amp * np.sin(1.3 * x + 0.4 * y) * np.cos(0.9 * y)).astype(np.float32).astype(np.float64)
    Dy = (amp * np.cos(0.7 * x) * np.sin(1.1 * y + 0.3 * x)).astype(np.float32).astype(np.float64)
    Dx[rng.random(shape) < 0.1] = 0.0
    Dy[rng.random(shape) < 0.1] = 0.0
    jump = rng.random(shape) < 0.05                      # rays that tear away from their neighbours
    Dx[jump] += np.float32(5.0) * rng.standard_normal(jump.sum()).astype(np.float32)
    I = rng.uniform(0.5, 2.0, shape).astype(np.float32).astype(np.float64)
    want = po.splat(I, Dx, Dy, margin)
    got, flag = run_splat(abi, I, Dx, Dy, margin, variant)
    assert flag == 0
    assert rel_l2(got, want) < 5e-7 * SPLAT_TOL[variant]
    # conservation: what stays inside the frame is what the oracle says stays
    assert abs(got.sum() / want.sum() - 1) < 1e-6


def test_splat_flags_nonfinite(abi):
    I = np.ones((16, 40)); I[3, 3] = np.nan
    for variant in (2, 3, 4, 5):
        _, flag = run_splat(abi, I, np.zeros((16, 40)), np.zeros((16, 40)), 15, variant)
        assert flag & abi.FLAG_NONFINITE


def test_splat_large_conservation(abi):
    """Full-size property test (2048^2): rays that stay in the frame conserve the total."""
    n = 2048
    x = torch.linspace(0, 40, n, device="cuda")
    Dx = (3.0 * torch.sin(x)[:, None] * torch.cos(0.5 * x)[None, :]).contiguous()
    Dy = (2.0 * torch.cos(0.7 * x)[:, None] * torch.sin(x)[None, :]).contiguous()
    I = torch.ones((n, n), device="cuda")
    I[:8] = 0; I[-8:] = 0; I[:, :8] = 0; I[:, -8:] = 0
    outs = []
    for variant in (0, 2, 3, 4):
        # variant 4 OVERWRITES (fastRefraction's contract: refractionFileNumba2.py:70,77): start it from garbage
        out = torch.zeros((n, n), device="cuda") if variant != 4 else torch.full((n, n), 123.0, device="cuda")
        abi.splat(I, Dx, Dy, out, margin=15, variant=variant)
        outs.append(out)
        assert abs(out.double().sum().item() / I.double().sum().item() - 1) < 1e-6
    assert rel_l2(outs[1].cpu().numpy(), outs[0].cpu().numpy()) < 1e-6
    assert rel_l2(outs[2].cpu().numpy(), outs[0].cpu().numpy()) < 5e-6      # unit = 2^-18 of these rays, weights floored
    assert rel_l2(outs[3].cpu().numpy(), outs[0].cpu().numpy()) < 1e-6      # unit = 2^-22 of twice the mean ray
    # negative and very unequal intensities, `out +=` on a non-zero image: the tile variant against plain REDs
    I2 = (I * (1.0 + 50.0 * (torch.rand((n, n), device="cuda") > 0.999))).contiguous()
    I2[100:110, 200:260] = -3.0
    base = torch.rand((n, n), device="cuda")
    a, b, c = base.clone(), base.clone(), base.clone()
    abi.splat(I2, Dx, Dy, a, margin=15, variant=0)
    abi.splat(I2, Dx, Dy, b, margin=15, variant=3)
    abi.splat(I2, Dx, Dy, c, margin=15, variant=5)
    assert rel_l2((b - base).cpu().numpy(), (a - base).cpu().numpy()) < 1e-5
    assert rel_l2((c - base).cpu().numpy(), (a - base).cpu().numpy()) < 1e-5


@pytest.mark.parametrize("tag", ["sub", "mid", "far", "huge"])
def test_refract_phi_vs_reference(abi, golden, tag):
    g = golden("fast_refraction")
    pix, z, E, M = g["params"]
    I32 = g["I"].astype(np.float32)
    out = torch.zeros(I32.shape, device="cuda")
    dxp = torch.zeros((I32.shape[0] + 30, I32.shape[1] + 30), device="cuda")
    dyp = torch.zeros_like(dxp)
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    abi.refract_phi(dev(I32), dev(g["phi_" + tag], torch.float64), out, z, E, M, pix, 15, dxp, dyp, flag)
    assert int(flag.item()) == 0
    assert rel_l2(dxp.cpu().numpy(), g["Dx_" + tag]) < 1e-6
    assert rel_l2(dyp.cpu().numpy(), g["Dy_" + tag]) < 1e-6
    assert rel_l2(out.cpu().numpy(), g["out_" + tag]) < TOL
    # v1 (margin 10) is the same map below its clamp
    if tag != "huge":
        out1 = torch.zeros(I32.shape, device="cuda")
        abi.refract_phi(dev(I32), dev(g["phi_" + tag], torch.float64), out1, z, E, M, pix, 10)
        assert rel_l2(out1.cpu().numpy(), g["outv1_" + tag]) < TOL


def _layer_coeffs(deltas, betas, E, z, M, pix):
    from paresis_b200 import hostmath as hm
    k = hm.wavenumber(E * 1000)
    s = hm.refraction_gradient_scale(z, M, pix)
    return [d * s for d in deltas], [2 * k * b for b in betas]


@pytest.mark.parametrize("kernel", ["direct", "tile", "strip"])
@pytest.mark.parametrize("shape", [(130, 290), (300, 257), (700, 1100), (96, 2500)])
def test_refract_layers_vs_oracle(abi, shape, kernel):
    """Experiment.py:463-474: membrane hop, then the fused sample + reference hop -- each hop kernel against the
    ORACLE, not against another CUDA kernel.  "tile": the round-1 fixed-point tile kernels (refract_lean.cuh, out +=);
    "strip": the production owner-computes rolling-strip kernels (refract_strip.cuh), which OVERWRITE their outputs
    (fastRefraction scatters into fresh zeros, refractionFileNumba2.py:70,77) -- they start from garbage here."""
    scale = {"direct": {}, "tile": dict(intensity_scale=7500.0), "strip": dict(intensity_scale=7500.0, mode=1)}[kernel]
    fresh = (lambda: torch.full(shape, 321.0, device="cuda")) if kernel == "strip" else (lambda: torch.zeros(shape, device="cuda"))
    rng = np.random.default_rng(17)
    x = np.linspace(0, 5, shape[0])[:, None]
    y = np.linspace(0, 5, shape[1])[None, :]
    t_mem = (2e-4 * (1 + np.sin(3 * x) * np.cos(2.3 * y)) * (rng.random(shape) > 0.02)).astype(np.float32)
    t_smp = (9e-4 * np.exp(-((x - 2.5) ** 2 + (y - 2.2) ** 2))).astype(np.float32)
    E, pix, M, d2, d3 = 52.0, 2.9256, 1.0254, 1.6, 3.6
    dm, bm, ds, bs = [5.97e-7], [5.37e-9], [9.52e-8], [4.4e-11]
    i0 = 7500.0
    # oracle, fp64, on the same fp32-rounded maps
    tm64, ts64 = t_mem.astype(np.float64), t_smp.astype(np.float64)
    i_m, phi_m = po.set_wave_rt(np.full(shape, i0), np.zeros(shape), [tm64], dm, bm, E)
    i_bs, _, _ = po.fast_refraction(i_m, phi_m, d2, E, M, pix)
    i_s, phi_ms = po.set_wave_rt(i_bs, phi_m, [ts64], ds, bs, E)
    want_s, _, _ = po.fast_refraction(i_s, phi_ms, d3, E, M, pix)
    want_r, _, _ = po.fast_refraction(i_bs.copy(), phi_m, d3, E, M, pix)
    # GPU
    tm, ts = dev(t_mem), dev(t_smp)
    g2, a2 = _layer_coeffs(dm, bm, E, d2, M, pix)
    ibs = fresh()
    abi.refract_layers(None, i0, [(tm, g2[0], 0.0, a2[0])], ibs, reach=8, **scale)
    assert rel_l2(ibs.cpu().numpy(), i_bs) < TOL
    g3m, _ = _layer_coeffs(dm, bm, E, d3, M, pix)
    g3s, a3s = _layer_coeffs(ds, bs, E, d3, M, pix)
    out_s, out_r = fresh(), fresh()
    ssum = torch.zeros(1, device="cuda", dtype=torch.float64)
    abi.refract_layers(ibs, 0.0, [(tm, g3m[0], g3m[0], 0.0), (ts, g3s[0], 0.0, a3s[0])], out_s, out_r, sum_ref=ssum, **scale)
    assert rel_l2(out_r.cpu().numpy(), want_r) < TOL
    assert rel_l2(out_s.cpu().numpy(), want_s) < TOL
    assert abs(ssum.item() / want_r.sum() - 1) < 1e-5          # Experiment.py:485-486, formed while depositing
    if kernel == "strip":
        # mode 2: a second energy of the same detector bin is ADDED to the accumulators (Experiment.py:482-483)
        abi.refract_layers(ibs, 0.0, [(tm, g3m[0], g3m[0], 0.0), (ts, g3s[0], 0.0, a3s[0])], out_s, out_r, sum_ref=ssum,
                           intensity_scale=7500.0, mode=2)
        assert rel_l2(out_r.cpu().numpy(), 2 * want_r) < TOL and rel_l2(out_s.cpu().numpy(), 2 * want_s) < TOL
        assert abs(ssum.item() / (2 * want_r.sum()) - 1) < 1e-5
        # the single-beam object hop (propagation image, Experiment.py:490-492) at reach 12
        out_p = fresh()
        abi.refract_layers(ibs, 0.0, [(tm, g3m[0], 0.0, 0.0), (ts, g3s[0], 0.0, a3s[0])], out_p, intensity_scale=7500.0, mode=1, reach=12)
        assert rel_l2(out_p.cpu().numpy(), want_s) < TOL


@pytest.mark.parametrize("mode", [0, 1], ids=["tile", "strip"])
def test_refract_layers_dark_regions(abi, mode):
    """A sample with bands of 1 ... 1e-6 transmission: the global relative L2 is dominated by the
    bright background, so each band is held to the oracle LOCALLY.  Rays too dim for the fixed-point tile
    (below 2^-10 of the intensity scale) must take the fp32 path instead of being quantised away."""
    from paresis_b200 import hostmath as hm
    shape = (384, 1536)
    rng = np.random.default_rng(5)
    x = np.linspace(0, 5, shape[0])[:, None]
    y = np.linspace(0, 9, shape[1])[None, :]
    t_mem = (1.5e-4 * (1 + np.sin(3 * x) * np.cos(2.3 * y)) * (rng.random(shape) > 0.02)).astype(np.float32)
    E, pix, M, d3 = 52.0, 2.9256, 1.0254, 3.6
    k = hm.wavenumber(E * 1000)
    beta = 4.4e-9
    trans = [1.0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-6]
    band = shape[1] // len(trans)
    t_smp = np.zeros(shape, dtype=np.float32)
    for b, tr in enumerate(trans):
        t_smp[:, b * band:(b + 1) * band] = -np.log(tr) / (2 * k * beta)
    t_smp *= (1 + 0.02 * np.sin(4 * x) * np.sin(3 * y)).astype(np.float32)     # a little refraction inside the bands
    dm, ds, bs = [5.97e-7], [9.52e-8], [beta]
    i0 = 7500.0
    i_in = (i0 * rng.uniform(0.8, 1.2, shape)).astype(np.float32)
    phi_m = -k * dm[0] * t_mem.astype(np.float64)
    i_s, phi_ms = po.set_wave_rt(i_in.astype(np.float64), phi_m, [t_smp.astype(np.float64)], ds, bs, E)
    want_s, _, _ = po.fast_refraction(i_s, phi_ms, d3, E, M, pix)
    want_r, _, _ = po.fast_refraction(i_in.astype(np.float64), phi_m, d3, E, M, pix)
    g3m, _ = _layer_coeffs(dm, [0.0], E, d3, M, pix)
    g3s, a3s = _layer_coeffs(ds, bs, E, d3, M, pix)
    out_s = torch.zeros(shape, device="cuda"); out_r = torch.zeros(shape, device="cuda")
    abi.refract_layers(dev(i_in), 0.0, [(dev(t_mem), g3m[0], g3m[0], 0.0), (dev(t_smp), g3s[0], 0.0, a3s[0])], out_s, out_r,
                       intensity_scale=i0, mode=mode)
    got_s, got_r = out_s.cpu().numpy(), out_r.cpu().numpy()
    assert rel_l2(got_r, want_r) < TOL and rel_l2(got_s, want_s) < TOL
    for b, tr in enumerate(trans):
        sl = (slice(16, -16), slice(b * band + 16, (b + 1) * band - 16))      # band interior: no light from the neighbours
        assert want_s[sl].mean() < 1.3 * i0 * tr
        # fixed-point bands: a unit is 2^-19 (tile) / 2^-21 (strip, reach 12) of the beam, so a band at transmission tr is
        # held to ~1 unit per pixel; bands whose rays fall below 2^9 units take the fp32 path
        unit = 2.0 ** -19 if mode == 0 else 2.0 ** -21
        assert rel_l2(got_s[sl], want_s[sl]) < max(2e-5, 1.2 * unit / tr), tr
        assert abs(got_s[sl].sum() / want_s[sl].sum() - 1) < 2e-5, tr               # unbiased: the band total is kept


def test_transmission(abi, golden):
    from paresis_b200 import hostmath as hm
    g = golden("waves")
    E = float(g["E"])
    k = hm.wavenumber(E * 1000)
    t = [dev(g["t"][0]), dev(g["t"][1])]
    shape = g["I"].shape
    i_out = torch.empty(shape, device="cuda")
    phi_out = torch.empty(shape, device="cuda", dtype=torch.float64)
    abi.transmit_rt(dev(g["I"]), dev(g["phi0"], torch.float64), t, [2 * k * b for b in g["beta"]],
                    [k * d for d in g["delta"]], i_out, phi_out)
    assert rel_l2(i_out.cpu().numpy(), g["I_rt"]) < 1e-6
    assert rel_l2(phi_out.cpu().numpy(), g["phi_rt"]) < 1e-6
    w_out = torch.empty(shape, device="cuda", dtype=torch.complex64)
    abi.transmit_wave(dev(g["wave0"], torch.complex64), 0.0, t, [k * b for b in g["beta"]], [k * d for d in g["delta"]], w_out)
    got = w_out.cpu().numpy()
    assert np.abs(got - g["wave"]).max() / np.abs(g["wave"]).max() < 2e-5  # fp32 thickness carries ~1e-5 rad
    assert rel_l2(np.abs(got) ** 2, np.abs(g["wave"]) ** 2) < 1e-6


def test_fresnel_propagation(abi, golden):
    from paresis_b200 import hostmath as hm
    g = golden("waves")
    E = float(g["E"])
    nx, ny = g["wave"].shape
    plan = abi.FresnelPlan(nx, ny, 15)
    w_in = dev(g["wave"], torch.complex64)
    for kcase in range(2):
        z, M, pix = g["prop%d_cfg" % kcase]
        hx, hy, phase = hm.fresnel_vectors(nx, ny, 15, (nx, ny), pix, z, E, M)
        w_out = torch.empty((nx, ny), device="cuda", dtype=torch.complex64)
        acc = torch.zeros((nx, ny), device="cuda")
        plan.propagate(w_in, dev(hx, torch.complex64), dev(hy, torch.complex64), phase, w_out, acc)
        want = g["prop%d" % kcase]
        assert rel_l2(np.abs(w_out.cpu().numpy()) ** 2, np.abs(want) ** 2) < TOL
        assert rel_l2(acc.cpu().numpy(), np.abs(want) ** 2) < TOL
        # complex field incl. the global phase exp(ikz/M) (k z ~ 1e12 rad: only fp64 on the host can carry it)
        assert np.abs(w_out.cpu().numpy() - want).max() / np.abs(want).max() < 1e-4
    plan.close()


@pytest.mark.parametrize("shape,margin", [((96, 128), 15), ((70, 111), 15), ((129, 64), 7), ((50, 50), 0), ((300, 260), 16),
                                          ((256, 512), 15), ((1000, 1300), 15), ((2048, 1024), 15)])
def test_fresnel_separable_path_is_the_padded_transform(abi, shape, margin):
    """The production propagator (per-axis circular convolution of period n + 2m through line transforms plus the reflect-margin
    terms, csrc/fresnel.cu; the larger shapes take the in-shared-memory transform of fresnel_lines.cuh at M = 512 ... 4096, the
    others batched cuFFT) against numpy's literal pad -> fft2 -> transfer -> ifft2 -> crop (Experiment.py:236-251) in fp64, and
    against the library's own literal chain (paresis_fresnel_spectrum / _from_spectrum)."""
    from paresis_b200 import hostmath as hm
    nx, ny = shape
    rng = np.random.default_rng(nx * 1000 + ny)
    wave = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)
    plan = abi.FresnelPlan(nx, ny, margin)
    w_in = dev(wave, torch.complex64)
    for z, E, M, pix in ((0.7, 22.0, 1.3, 0.9), (3.0, 17.0, 2.0, 2.5)):
        hx, hy, phase = hm.fresnel_vectors(nx, ny, margin, (nx, ny), pix, z, E, M)
        padded = np.pad(wave.astype(np.complex128), margin, mode="reflect") if margin else wave.astype(np.complex128)
        h2 = hx.astype(np.complex128)[:, None] * hy.astype(np.complex128)[None, :]
        full = np.fft.ifft2(np.fft.fft2(padded) * h2) * (padded.shape[0] * padded.shape[1])     # the vectors carry the 1/(Px Py)
        want = full[margin:margin + nx, margin:margin + ny]
        hx_d, hy_d = dev(hx, torch.complex64), dev(hy, torch.complex64)
        got = torch.empty(shape, device="cuda", dtype=torch.complex64)
        acc = torch.full(shape, 2.0, device="cuda")
        kern = plan.kernel(hx_d, hy_d)
        plan.convolve(w_in, kern, 1.0, got, acc)
        def dist(a, b):      # complex-aware (conftest.rel_l2 is for real images)
            return float(np.linalg.norm(a.astype(np.complex128) - b) / np.linalg.norm(b))
        assert dist(got.cpu().numpy(), want) < 2e-6
        assert rel_l2(acc.cpu().numpy() - 2.0, np.abs(want) ** 2) < 1e-5
        again = torch.empty_like(got)
        plan.propagate(w_in, hx_d, hy_d, 1.0, again, None)
        assert torch.equal(again, got)
        inplace = w_in.clone()
        plan.convolve(inplace, kern, 1.0, inplace, None)
        assert torch.equal(inplace, got)
        lit = torch.empty_like(got)
        plan.spectrum(w_in)
        plan.from_spectrum(hx_d, hy_d, 1.0, lit, None)
        assert dist(lit.cpu().numpy(), want) < 2e-6
        if (nx, ny) == (96, 128):      # a kernel only fits plans of its own size
            other = abi.FresnelPlan(64, 128, margin)
            with pytest.raises(RuntimeError, match="prepared for a 96 x 128 plan"):
                other.convolve(torch.zeros((64, 128), device="cuda", dtype=torch.complex64), kern, 1.0,
                               torch.empty((64, 128), device="cuda", dtype=torch.complex64), None)
            other.close()
        kern.close()
    plan.close()


@pytest.mark.parametrize("n", [4096, 8192])
def test_fresnel_benchmark_grids_match_the_literal_chain(abi, n):
    """M = 8192 and 16384 of the in-shared-memory line transform (the sizes bench.py times) against cuFFT's Bluestein transform
    of the reference's (n + 30)^2 grid, on a field with the benchmark's dynamic range."""
    from paresis_b200 import hostmath as hm
    g = torch.Generator(device="cuda").manual_seed(n)
    amp = 1.0 + 0.3 * torch.rand((n, n), device="cuda", generator=g)
    ph = 6.0 * torch.rand((n, n), device="cuda", generator=g)
    w_in = torch.polar(amp, ph).to(torch.complex64)
    del amp, ph
    plan = abi.FresnelPlan(n, n, 15)
    hx, hy, phase = hm.fresnel_vectors(n, n, 15, (n, n), 0.75 * 4096 / n, 1.0, 52.0, 1.5)
    hx_d, hy_d = dev(hx, torch.complex64), dev(hy, torch.complex64)
    got = torch.empty_like(w_in)
    plan.propagate(w_in, hx_d, hy_d, 1.0, got, None)
    lit = torch.empty_like(w_in)
    plan.spectrum(w_in)
    plan.from_spectrum(hx_d, hy_d, 1.0, lit, None)
    num = torch.linalg.vector_norm((got - lit).to(torch.complex128)).item()
    den = torch.linalg.vector_norm(lit.to(torch.complex128)).item()
    assert num / den < 3e-6, num / den
    # intensities, the quantity the images are made of
    assert rel_l2((got.abs() ** 2).cpu().numpy()[::7, ::5], (lit.abs() ** 2).cpu().numpy()[::7, ::5]) < 3e-6
    plan.close()
    if n == 4096:
        # ... and against the literal chain in float64 on the host (Experiment.py:236-251), the whole 4126^2 transform
        del lit
        wave = w_in.cpu().numpy().astype(np.complex128)
        spec = np.fft.fft2(np.pad(wave, 15, mode="reflect"))
        del wave
        spec *= hx.astype(np.complex128)[:, None]
        spec *= hy.astype(np.complex128)[None, :]
        want = (np.fft.ifft2(spec) * float(spec.shape[0] * spec.shape[1]))[15:15 + n, 15:15 + n]     # the vectors carry 1/(Px Py)
        del spec
        g = got.cpu().numpy()
        assert float(np.linalg.norm(g - want) / np.linalg.norm(want)) < 3e-6
        assert rel_l2(np.abs(g) ** 2, np.abs(want) ** 2) < 3e-6


def test_detection(abi, golden):
    from paresis_b200 import hostmath as hm
    g = golden("detector")
    for kcase in range(int(g["n_det"])):
        os_, d0, d1, fwhm, psf = g["det%d_cfg" % kcase]
        os_, d0, d1 = int(os_), int(d0), int(d1)
        img = dev(g["det%d_in" % kcase])
        src = dev(hm.gaussian_1d(fwhm / 2.355)) if fwhm != 0 else None
        pk = dev(hm.gaussian_1d(psf)) if psf != 0 else None
        work = torch.empty(abi.detect_work_floats(d0 * os_, d1 * os_, os_, d0, d1), device="cuda")
        out = torch.empty((d0, d1), device="cuda")
        abi.detect(img, os_, d0, d1, src, pk, work, out)
        assert rel_l2(out.cpu().numpy(), g["det%d_out" % kcase]) < 1e-6, kcase
        # the fused single-kernel path (production) gives the same expectation ...
        fused = torch.empty((d0, d1), device="cuda")
        abi.detect_counts(img, os_, d0, d1, src, pk, work, fused, False)
        assert rel_l2(fused.cpu().numpy(), g["det%d_out" % kcase]) < 1e-6, kcase
        # ... and Poisson counts around it
        counts = torch.empty((d0, d1), device="cuda")
        abi.detect_counts(img, os_, d0, d1, src, pk, work, counts, True, 99, kcase)
        ok = fused > 10.0          # (one synthetic input dips below zero: lambda <= 0 draws 0)
        z = ((counts - fused) / fused.clamp_min(1.0).sqrt())[ok].double()
        assert torch.equal(counts, counts.round()) and abs(z.mean().item()) < 0.08 and abs(z.std().item() - 1) < 0.08, kcase
    out = torch.empty((30, 45), device="cuda")
    abi.bin_sum(dev(g["resize_in"]), 30, 45, out)
    assert rel_l2(out.cpu().numpy(), g["resize_2"]) < 1e-6
    out = torch.empty((20, 30), device="cuda")
    abi.bin_sum(dev(g["resize_in"]), 20, 30, out)
    assert rel_l2(out.cpu().numpy(), g["resize_3"]) < 1e-6


def test_membrane_and_samples(abi, golden):
    from ref_harness import synthetic_sphere_rows
    g = golden("geometry")
    rows = synthetic_sphere_rows(0, 60000)
    for tag in ("mem0", "mem1"):
        mean_r, layers, dx, dy, pix, support, seed = g[tag + "_cfg"]
        dx, dy, layers = int(dx), int(dy), int(layers)
        tab, ex, ey = po.membrane_sphere_table(rows, float(mean_r), dx, dy, float(pix))
        np.random.seed(int(seed))
        offs = po.draw_membrane_offsets(layers, float(mean_r), ex, ey, dx, dy, float(pix))
        margin, _ = po.membrane_margin(float(mean_r), float(pix))
        out = torch.empty((dx, dy), device="cuda")
        abi.raster_spheres(dev(tab, torch.float64), float(pix), offs, dx, dy, margin, out)
        assert rel_l2(out.cpu().numpy(), g[tag][0]) < 1e-6, tag
    r, dx, dy, pix = g["sphere_cfg"]
    out = torch.empty((int(dx), int(dy)), device="cuda")
    abi.sphere_map(float(r), int(dx), int(dy), float(pix), out)
    assert rel_l2(out.cpu().numpy(), g["sphere"][0]) < 1e-6
    for tag in ("cyl", "cyl2"):
        r, ang, dx, dy, pix = g[tag + "_cfg"]
        out = torch.empty((int(dx), int(dy)), device="cuda")
        abi.cylinder_map(float(r), float(ang), int(dx), int(dy), float(pix), out)
        assert rel_l2(out.cpu().numpy(), g[tag][0]) < 1e-6, tag


def test_poisson_statistics(abi):
    """rs.poisson (Detector.py:113-115) cannot be compared draw for draw (wall-clock seed);
    check moments for small / medium / large lambda, a chi-square at low lambda, and the
    counter-based reproducibility."""
    from scipy import stats
    n = 1 << 20
    for lam in (0.3, 4.0, 9.99, 10.0, 37.5, 7500.0, 3.0e5):
        expect = torch.full((n,), lam, device="cuda")
        counts = torch.empty(n, device="cuda")
        abi.poisson(expect, counts, 1234, 7)
        c = counts.double()
        m, v = c.mean().item(), c.var().item()
        assert abs(m - lam) < 6 * np.sqrt(lam / n), (lam, m)
        assert abs(v / lam - 1) < 6 * np.sqrt(2.0 / n) + 2.0 / lam / np.sqrt(n) + 1e-3, (lam, v)
        assert torch.equal(c, torch.round(c)) and c.min().item() >= 0
        if lam < 40:
            kmax = int(stats.poisson.ppf(1 - 1e-7, lam)) + 1
            obs = torch.bincount(counts.long(), minlength=kmax + 1).cpu().numpy().astype(np.float64)
            pmf = stats.poisson.pmf(np.arange(len(obs)), lam) * n
            keep = pmf > 20
            chi2 = ((obs[keep] - pmf[keep]) ** 2 / pmf[keep]).sum()
            assert chi2 < stats.chi2.ppf(1 - 1e-6, keep.sum()), (lam, chi2)
    lam = torch.rand(4096, device="cuda") * 50
    a = torch.empty(4096, device="cuda"); b = torch.empty(4096, device="cuda"); c2 = torch.empty(4096, device="cuda")
    abi.poisson(lam, a, 5, 1); abi.poisson(lam, b, 5, 1); abi.poisson(lam, c2, 5, 2)
    assert torch.equal(a, b) and not torch.equal(a, c2)
    # sharding independence: the draw for pixel p does not depend on the launch extent
    d = torch.empty(1000, device="cuda")
    abi.poisson(lam[:1000].contiguous(), d, 5, 1)
    assert torch.equal(d, a[:1000])


@pytest.fixture
def tile_config(abi, request):
    """0 = the tile kernel (refract_lean.cuh)."""
    abi.set_tuning(2, request.param)
    yield request.param
    abi.set_tuning(2, 0)


@pytest.mark.parametrize("tile_config", [0], indirect=True)
@pytest.mark.parametrize("shape", [(130, 290), (300, 257), (40, 36), (517, 1030), (96, 2048)])
def test_tile_hop_kernel_matches_direct_kernel(abi, shape, tile_config):
    """The fixed-point shared-memory tile kernels (intensity_scale given) against the fp32 direct-to-L2
    kernel on the same inputs: membrane-like torn field, rays brighter than 2 x scale, negative rays,
    ragged / odd sizes (border warps and rows, the miss list, unaligned zero-fill); plus the pipeline
    extras (zero-fill, clear-input, reference-beam sum)."""
    rng = np.random.default_rng(5)
    x = np.linspace(0, 9, shape[0])[:, None]
    y = np.linspace(0, 9, shape[1])[None, :]
    caps = np.sqrt(np.maximum(0.0, 0.2 - (np.mod(x, 1.0) - 0.5) ** 2 - (np.mod(y, 1.0) - 0.5) ** 2))   # torn gradients
    t_mem = (6e-4 * caps).astype(np.float32)
    t_smp = (9e-4 * np.exp(-((x - 4.5) ** 2 + (y - 4.2) ** 2))).astype(np.float32)
    E, pix, M, d3 = 52.0, 2.9256, 1.0254, 3.6
    g3m, _ = _layer_coeffs([5.97e-7], [5.37e-9], E, d3, M, pix)
    g3s, a3s = _layer_coeffs([9.52e-8], [4.4e-11], E, d3, M, pix)
    i0 = 7500.0
    i_in = (i0 * (0.6 + 0.8 * rng.random(shape))).astype(np.float32)
    i_in[rng.random(shape) > 0.995] *= 4.0          # brighter than 2 x intensity_scale: fp32 path
    i_in[3, 5] = -10.0                              # negative: fp32 path
    layers = [(dev(t_mem), g3m[0], g3m[0], 0.0), (dev(t_smp), g3s[0], 0.0, a3s[0])]
    ref_s = torch.zeros(shape, device="cuda"); ref_r = torch.zeros(shape, device="cuda")
    abi.refract_layers(dev(i_in), 0.0, layers, ref_s, ref_r)
    src = dev(i_in)
    out_s = torch.zeros(shape, device="cuda"); out_r = torch.zeros(shape, device="cuda")
    junk = [torch.full(shape, 3.0, device="cuda") for _ in range(2)]
    total = torch.full((1,), 5.0, device="cuda", dtype=torch.float64)
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    abi.refract_layers(src, 0.0, layers, out_s, out_r, flag=flag, zero_fill=junk, clear_input=True, sum_ref=total,
                       intensity_scale=i0)
    assert int(flag.item()) == 0
    assert rel_l2(out_s.cpu().numpy(), ref_s.cpu().numpy()) < 3e-6
    assert rel_l2(out_r.cpu().numpy(), ref_r.cpu().numpy()) < 3e-6
    assert not src.any() and not junk[0].any() and not junk[1].any()
    assert abs((total.item() - 5.0) / ref_r.double().sum().item() - 1) < 1e-6
    # single beam, uniform input (the membrane hop), with the scalar zeroed by the kernel
    one_ref = torch.zeros(shape, device="cuda"); one = torch.zeros(shape, device="cuda")
    abi.refract_layers(None, i0, [(dev(t_mem), g3m[0], 0.0, a3s[0])], one_ref)
    abi.refract_layers(None, i0, [(dev(t_mem), g3m[0], 0.0, a3s[0])], one, zero_scalar=total, intensity_scale=i0)
    assert rel_l2(one.cpu().numpy(), one_ref.cpu().numpy()) < 3e-6 and total.item() == 0.0
    # a NaN intensity is reported, as the reference's guard does (refractionFileNumba2.py:81-82)
    bad = i_in.copy(); bad[shape[0] // 2, shape[1] // 2] = np.nan
    abi.refract_layers(dev(bad), 0.0, layers, out_s, out_r, flag=flag, intensity_scale=i0)
    assert int(flag.item()) & abi.FLAG_NONFINITE


@pytest.mark.parametrize("tile_config", [0], indirect=True)
@pytest.mark.parametrize("n_extra", [2, 3])
def test_tile_hop_kernels_with_three_and_four_layers(abi, tile_config, n_extra):
    """Membrane + 2 or 3 sample materials (the multi-material phantoms of createSampGeom.py:110-260): the tile
    kernels' 3- and 4-layer instantiations, one and two beams, against the direct kernel."""
    shape = (300, 600)
    rng = np.random.default_rng(3)
    x = np.linspace(0, 9, shape[0])[:, None]; y = np.linspace(0, 9, shape[1])[None, :]
    caps = np.sqrt(np.maximum(0.0, 0.2 - (np.mod(x, 1.0) - 0.5) ** 2 - (np.mod(y, 1.0) - 0.5) ** 2))
    E, pix, M, d3 = 52.0, 2.9256, 1.0254, 3.6
    g3m, _ = _layer_coeffs([5.97e-7], [5.37e-9], E, d3, M, pix)
    layers = [(dev((6e-4 * caps).astype(np.float32)), g3m[0], g3m[0], 0.0)]
    for k in range(n_extra):
        t = (7e-4 * np.exp(-((x - 3.0 - 1.5 * k) ** 2 + (y - 4.0 - 0.7 * k) ** 2) / (1.0 + 0.5 * k))).astype(np.float32)
        g, a = _layer_coeffs([9.52e-8 * (1 + 0.4 * k)], [4.4e-11 * (1 + k)], E, d3, M, pix)
        layers.append((dev(t), g[0], 0.0, a[0]))
    i0 = 7500.0
    i_in = dev((i0 * (0.6 + 0.8 * rng.random(shape))).astype(np.float32))
    ref_s = torch.zeros(shape, device="cuda"); ref_r = torch.zeros(shape, device="cuda")
    abi.refract_layers(i_in, 0.0, layers, ref_s, ref_r)
    out_s = torch.zeros(shape, device="cuda"); out_r = torch.zeros(shape, device="cuda")
    abi.refract_layers(i_in, 0.0, layers, out_s, out_r, intensity_scale=i0)
    assert rel_l2(out_s.cpu().numpy(), ref_s.cpu().numpy()) < 3e-6
    assert rel_l2(out_r.cpu().numpy(), ref_r.cpu().numpy()) < 3e-6
    one_ref = torch.zeros(shape, device="cuda"); one = torch.zeros(shape, device="cuda")
    single = [(t, g, 0.0, a) for (t, g, _, a) in layers]
    abi.refract_layers(None, i0, single, one_ref)
    abi.refract_layers(None, i0, single, one, intensity_scale=i0)
    assert rel_l2(one.cpu().numpy(), one_ref.cpu().numpy()) < 3e-6


def _misaligned(shape, fill=0.0):
    """A contiguous [nx, ny] float32 view that starts 4 bytes into its allocation (no 16-byte alignment)."""
    buf = torch.full((shape[0] * shape[1] + 1,), fill, device="cuda", dtype=torch.float32)
    return buf[1:].view(shape)


def test_tile_kernels_with_unaligned_images(abi):
    """Images that are not 16-byte aligned and a pitch that is not a multiple of 4: the vector flush / zero-fill
    paths must step aside (lean hop kernel and tile splat against the direct kernels)."""
    rng = np.random.default_rng(11)
    for shape in ((300, 520), (300, 523)):
        x = np.linspace(0, 9, shape[0])[:, None]; y = np.linspace(0, 9, shape[1])[None, :]
        caps = np.sqrt(np.maximum(0.0, 0.2 - (np.mod(x, 1.0) - 0.5) ** 2 - (np.mod(y, 1.0) - 0.5) ** 2))
        t_mem = dev((6e-4 * caps).astype(np.float32))
        E, pix, M, d3 = 52.0, 2.9256, 1.0254, 3.6
        g3m, a3m = _layer_coeffs([5.97e-7], [5.37e-9], E, d3, M, pix)
        i0 = 7500.0
        ref = torch.zeros(shape, device="cuda")
        abi.refract_layers(None, i0, [(t_mem, g3m[0], 0.0, a3m[0])], ref)
        out = _misaligned(shape)
        junk = [_misaligned(shape, 3.0) for _ in range(3)]
        abi.refract_layers(None, i0, [(t_mem, g3m[0], 0.0, a3m[0])], out, zero_fill=junk, intensity_scale=i0)
        assert rel_l2(out.cpu().numpy(), ref.cpu().numpy()) < 3e-6
        assert not any(j.any() for j in junk)
        # stand-alone splat, variant 3 into an unaligned image
        I = dev(rng.uniform(0.5, 2.0, shape).astype(np.float32))
        Dx = dev((2.5 * np.sin(1.3 * x + 0.4 * y)).astype(np.float32) * np.ones(shape, dtype=np.float32))
        Dy = dev((2.5 * np.cos(0.7 * x + 1.1 * y)).astype(np.float32) * np.ones(shape, dtype=np.float32))
        a = torch.zeros(shape, device="cuda"); b = _misaligned(shape)
        abi.splat(I, Dx, Dy, a, margin=15, variant=0)
        abi.splat(I, Dx, Dy, b, margin=15, variant=3)
        assert rel_l2(b.cpu().numpy(), a.cpu().numpy()) < 5e-6


@pytest.mark.parametrize("tile_config", [0], indirect=True)
def test_tile_hop_kernel_cannot_overflow(abi, tile_config):
    """Every ray of a tile focused into ONE cell (a lens): the fixed-point tile holds it (rays per tile x
    brightest fixed-point ray < 2^32) and the image gets the exact sum."""
    n = 256
    i = np.arange(n, dtype=np.float64)
    # parabolic thickness: gradient proportional to the distance from the centre -> all rays land at the centre
    E, pix, M, d3 = 52.0, 2.9256, 1.0254, 3.6
    g, _ = _layer_coeffs([1.0], [0.0], E, d3, M, pix)          # pixels per metre of 2-pixel difference
    c = (n - 1) / 2
    t = -(((i[:, None] - c) ** 2 + (i[None, :] - c) ** 2) / (4 * g[0]))    # D = g * (t[+1]-t[-1]) = (c - i)
    tm = dev(t.astype(np.float32))
    i0 = 1000.0
    out = torch.zeros((n, n), device="cuda"); ref = torch.zeros((n, n), device="cuda")
    abi.refract_layers(None, 1.9 * i0, [(tm, g[0], 0.0, 0.0)], out, intensity_scale=i0)
    abi.refract_layers(None, 1.9 * i0, [(tm, g[0], 0.0, 0.0)], ref)
    # (the fp32 reference adds ~16k rays into each of four cells in scheduling order: its own sum moves by ~1e-6)
    assert abs(out.double().sum().item() / ref.double().sum().item() - 1) < 1e-5
    assert abs(out.double().sum().item() / (n * n * 1.9 * i0) - 1) < 1e-5    # (four fp32 cells of ~3e7 each: ulp = 2)
    assert out.max().item() > 0.2 * n * n * 1.9 * i0          # really focused
    assert rel_l2(out.cpu().numpy(), ref.cpu().numpy()) < 3e-5     # (two fp32 scheduling orders of the same sums)


def test_membrane_from_field_matches_raster(abi):
    """Positions cut from the once-rasterised sphere field = positions rasterised one by one."""
    from ref_harness import synthetic_sphere_rows
    rows = synthetic_sphere_rows(0, 60000)
    mean_r, layers, dx, dy, pix = 50.0, 3, 300, 420, 2.853
    tab, ex, ey = po.membrane_sphere_table(rows, mean_r, dx, dy, pix)
    margin, margin2 = po.membrane_margin(mean_r, pix)
    table = dev(tab, torch.float64)
    reach = int(np.floor(tab[:, 2].max() / pix)) + 1
    assert reach <= margin2
    field = torch.empty((int(np.ceil(ex / pix)) + margin + 1, int(np.ceil(ey / pix)) + margin + 1), device="cuda")
    abi.raster_field(table, pix, field)
    np.random.seed(3)
    for _ in range(3):
        offs = po.draw_membrane_offsets(layers, mean_r, ex, ey, dx, dy, pix)
        a = torch.empty((dx, dy), device="cuda"); b = torch.full((dx, dy), -1.0, device="cuda")
        abi.raster_spheres(table, pix, offs, dx, dy, margin, a)
        abi.membrane_from_field(field, offs, margin, dx, dy, b)
        assert rel_l2(b.cpu().numpy(), a.cpu().numpy()) < 1e-6


def test_detector_images_in_one_launch(abi, golden):
    """paresis_detect_counts_multi = paresis_detect_counts image by image (same Poisson streams)."""
    from paresis_b200 import hostmath as hm
    g = golden("detector")
    for kcase in range(int(g["n_det"])):
        os_, d0, d1, fwhm, psf = g["det%d_cfg" % kcase]
        os_, d0, d1 = int(os_), int(d0), int(d1)
        imgs = [dev(g["det%d_in" % kcase]), dev(np.abs(g["det%d_in" % kcase][::-1].copy()) + 3.0)]
        src = dev(hm.gaussian_1d(fwhm / 2.355)) if fwhm != 0 else None
        pk = dev(hm.gaussian_1d(psf)) if psf != 0 else None
        for noise in (False, True):
            single = [torch.empty((d0, d1), device="cuda") for _ in imgs]
            multi = [torch.empty((d0, d1), device="cuda") for _ in imgs]
            for k, im in enumerate(imgs):
                abi.detect_counts(im, os_, d0, d1, src, pk, None, single[k], noise, 11, 100 + k)
            abi.detect_counts_multi(imgs, os_, d0, d1, src, pk, None, multi, noise, 11, [100, 101])
            for s, m in zip(single, multi):
                assert torch.equal(s, m), (kcase, noise)
            if noise:
                # ... and the fused draw is the stand-alone sampler applied to the expectation
                expect = torch.empty((d0, d1), device="cuda"); counts = torch.empty((d0, d1), device="cuda")
                abi.detect_counts(imgs[1], os_, d0, d1, src, pk, None, expect, False)
                abi.poisson(expect, counts, 11, 101)
                assert torch.equal(counts, multi[1]), kcase


def test_two_sphere_phantoms(abi, golden):
    """CreateSampleSpheresInCylinder / CreateSampleSpheresInParallelepiped (createSampGeom.py:110-260) against the reference."""
    g = golden("phantoms")
    for tag, kind in (("cyl_a", 0), ("cyl_b", 0), ("par_a", 1), ("par_b", 1)):
        dx, dy, pix = g[tag + "_cfg"]
        out = torch.empty((3, int(dx), int(dy)), device="cuda")
        abi.two_sphere_phantom(kind, int(dx), int(dy), float(pix), out)
        for m in range(3):
            assert rel_l2(out[m].cpu().numpy(), g[tag][m]) < 1e-6, (tag, m)
    with pytest.raises(abi.ParesisError, match="too big"):
        abi.two_sphere_phantom(0, 64, 64, 10.0, torch.empty((3, 64, 64), device="cuda"))


@pytest.mark.parametrize("shape,n_e", [((130, 290), 2), ((300, 257), 4), ((517, 1030), 3)])
def test_energy_group_hop_matches_energy_by_energy(abi, shape, n_e):
    """paresis_refract_group (several energies of a detector bin through the object hop at once) against the
    same energies one by one (paresis_refract_layers with out_ref), accumulating into the same images."""
    rng = np.random.default_rng(9)
    x = np.linspace(0, 9, shape[0])[:, None]
    y = np.linspace(0, 9, shape[1])[None, :]
    caps = np.sqrt(np.maximum(0.0, 0.2 - (np.mod(x, 1.0) - 0.5) ** 2 - (np.mod(y, 1.0) - 0.5) ** 2))
    tm, ts = dev((6e-4 * caps).astype(np.float32)), dev((9e-4 * np.exp(-((x - 4.5) ** 2 + (y - 4.2) ** 2))).astype(np.float32))
    pix, M, d3 = 2.9256, 1.0254, 3.6
    energies, scales = [30.0, 41.0, 52.0, 67.0][:n_e], [900.0, 4200.0, 7500.0, 150.0][:n_e]
    ref_s = torch.zeros(shape, device="cuda"); ref_r = torch.zeros(shape, device="cuda")
    got_s = torch.zeros(shape, device="cuda"); got_r = torch.zeros(shape, device="cuda")
    group, sums, want_sums = [], [], []
    for E, i0 in zip(energies, scales):
        dm, ds = 5.97e-7 * (52.0 / E) ** 2, 9.52e-8 * (52.0 / E) ** 2
        g3m, _ = _layer_coeffs([dm], [0.0], E, d3, M, pix)
        g3s, a3s = _layer_coeffs([ds], [4.4e-11 * (52.0 / E) ** 3], E, d3, M, pix)
        layers = [(tm, g3m[0], g3m[0], 0.0), (ts, g3s[0], 0.0, a3s[0])]
        i_in = (i0 * (0.6 + 0.8 * rng.random(shape))).astype(np.float32)
        i_in[rng.random(shape) > 0.997] *= 3.0          # brighter than 2 x its scale: fp32 path
        before = ref_r.double().sum().item()
        abi.refract_layers(dev(i_in), 0.0, layers, ref_s, ref_r)
        want_sums.append(ref_r.double().sum().item() - before)
        sums.append(torch.full((1,), 2.0, device="cuda", dtype=torch.float64))
        group.append((dev(i_in), i0, layers, sums[-1]))
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    abi.refract_group(group, got_s, got_r, flag)
    assert int(flag.item()) == 0
    assert rel_l2(got_s.cpu().numpy(), ref_s.cpu().numpy()) < 3e-6
    assert rel_l2(got_r.cpu().numpy(), ref_r.cpu().numpy()) < 3e-6
    for (inten, _, _, _), s_, w in zip(group, sums, want_sums):
        assert not inten.any()                                      # cleared behind the pass
        assert abs((s_.item() - 2.0) / w - 1) < 1e-5
