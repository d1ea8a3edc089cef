"""Oracle parity AT THE GRIDS bench.py TIMES (BASELINE.json configs 2, 3 and 5), through the PRODUCTION path.

Reference behaviour pinned here: Experiment.computeSampleAndReferenceImages_RT, Experiment.py:448-521 (the
per-energy loop with accumulate and bin close) around getMembraneSegmentedFromFile (getMembraneFromFile.py:60-171).

GPU side = exactly what bench.py's device job runs: ``ImageFormation.compute_rt_positions`` (one
``paresis_rt_run_positions`` call: membrane cut from the sphere field, the hop kernels, the grouped object hop of
a polychromatic bin, the multi-image detector, 3 positions in flight).  Oracle side = oracle/paresis_oracle.py in
fp64 from the sphere list onwards (its own membrane raster, its own hops and detector), same offsets, noise off.
Bound: 1e-5 relative L2 on every image (north star), 1e-6 on the membrane map.

  config 2: 2048^2, mono, positions 0 and 2 of a 3-position call
  config 3: 4096^2, 4 of the 64 energies (one detector bin -> the grouped object hop), position 0 and 1
  config 5: 8192^2, 2 of the 128 energies, position 1 (position 0 adds two more 8192^2 oracle hops; its extra
            images are covered at 2048^2 and 4096^2)
Distances are appended to $PARESIS_REPORT_L2 (profiles/r02_parity_distances.txt) by conftest.rel_l2.
"""
import importlib
import os

import numpy as np
import pytest

import paresis_oracle as po
from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from paresis_b200 import workspace
    ws = workspace.make_workspace(str(tmp_path_factory.mktemp("ws_grids")))
    old = os.getcwd()
    workspace.enter(ws)
    yield importlib.import_module("Experiment")
    os.chdir(old)


def _db(obj, energies):
    """{energy: (deltas, betas)} per material, from the lists the shim derived (Sample.py:83-152)."""
    out = {}
    for e in energies:
        out[e] = ([dict(obj.delta[m])[e] for m in range(len(obj.delta))], [dict(obj.beta[m])[e] for m in range(len(obj.beta))])
    return out


CASES = [
    # name, grid, spectrum indices kept (None = all), points of the call, points compared with the oracle, positions per launch
    pytest.param("B200_2048_mono", 2048, None, [0, 1, 2, 3, 4], [0, 2, 4], 4, id="config2-2048-mono-4-per-launch"),   # what bench.py times
    pytest.param("B200_2048_mono", 2048, None, [0, 1, 2], [1], 0, id="config2-2048-mono-3-streams"),
    pytest.param("B200_4096_poly64", 4096, [0, 21, 32, 63], [0, 1], [0, 1], 0, id="config3-4096-4of64"),
    pytest.param("B200_8192_poly128", 8192, [10, 100], [1, 2], [1], 0, id="config5-8192-2of128"),
]


@pytest.mark.parametrize("name,n,keep,points,check,per_launch", CASES)
def test_production_path_matches_oracle_at_benchmark_grid(shim, name, n, keep, points, check, per_launch):
    from paresis_b200 import geometry, workspace
    d = dict(experimentName=name, filepath="unused/", overSampling=2, nbExpPoints=len(points), simulation_type="RayT",
             expID="t", poissonNoise=False, returnDisplacement=False)
    e = shim.Experiment(d)
    assert tuple(int(v) for v in e.exp_dict['studyDimensions']) == (n, n)
    if keep is not None:
        e.mySource.mySpectrum = [e.mySource.mySpectrum[i] for i in keep]
    spectrum = list(e.mySource.mySpectrum)
    energies = [en for en, _ in spectrum]
    mem, smp = e.myMembrane, e.mySampleofInterest
    e.myDetector.det_param["myBinsThersholds"] = []
    thresholds = list(e._open_bins(0))
    assert thresholds == [energies[-1]]
    scene = e._scene(thresholds, per_position_membrane=True)
    plan = geometry.MembranePlan(mem, n, n, mem.membranePixelSize)
    assert plan.field() is not None            # the production membrane: windows of the sphere field
    np.random.seed(1000 + n)
    offsets = [plan.draw_offsets() for _ in points]
    eng = e._get_engine()
    res = eng.compute_rt_positions(scene, plan, offsets, points, n_slots=max(3, per_launch), per_launch=per_launch)
    torch.cuda.synchronize()
    eng.check_flag()

    # ---- oracle, fp64, from the sphere list
    det = e.myDetector.det_param
    src = e.mySource.source_dict
    setup = po.Setup(e.exp_dict['distSourceToMembrane'], e.exp_dict['distMembraneToObject'], e.exp_dict['distObjectToDetector'],
                     (n // 2, n // 2), det['myPixelSize'], 2, e.exp_dict['meanShotCount'], spectrum, src["mySize"], det['myPSF'],
                     energy_sampling=src["myEnergySampling"])
    assert abs(setup.study_pixel_um / e.exp_dict['studyPixelSize'] - 1) < 1e-14
    rows = workspace.synthetic_sphere_rows(0, 60000)      # what make_workspace wrote to Samples/Membranes/CuSn.txt
    mdb, sdb = _db(mem, energies), _db(smp, energies)
    # sample map: the GPU's (pinned to the reference at small grids by the geometry goldens), fp32 -> fp64
    sample_t = np.asarray(smp.myGeometry).astype(np.float64)
    assert sample_t.shape == (1, n, n)
    firsts = res["firsts"]
    for p in check:
        k = points.index(p)
        mem_t = po.membrane_segmented(rows, mem.myMeanSphereRadius, mem.myNbOfLayers, n, n, setup.membrane_pixel_um,
                                      mem.myPMMAThickness, offsets=offsets[k])
        got_t = res["thickness"][k].cpu().numpy()
        assert rel_l2(got_t, mem_t[0]) < 1e-6
        want = po.compute_rt(setup, mem_t, mdb, sample_t, sdb, p)
        del mem_t
        assert rel_l2(res["sample"][k].cpu().numpy(), want[0]) < TOL
        assert rel_l2(res["reference"][k].cpu().numpy(), want[1]) < TOL
        if p == 0:
            f = firsts.index(k)
            assert rel_l2(res["propag"][f].cpu().numpy(), want[2]) < TOL
            assert rel_l2(res["white"][f].cpu().numpy(), want[3]) < TOL
        # Experiment.py:485-486: mean energy from the per-energy means of the reference beam
        sums = res["sums"][k].cpu().numpy()
        assert sums.shape == (len(energies),) and (sums > 0).all()
        del want
