"""Energy sharding on the GPU: the sharded pipeline (partial sums -> linear detector -> reduction ->
Poisson) must reproduce the single-call pipeline.  World size 1 runs in-process on one GPU (bit
identical); world size 2 needs two GPUs and NCCL."""
import importlib
import os
import socket

import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _scene_and_engine(noise):
    from paresis_b200 import workspace
    import tempfile
    ws = workspace.make_workspace(tempfile.mkdtemp())
    workspace.enter(ws)
    shim = importlib.import_module("Experiment")
    d = dict(experimentName="B200_small_poly3", filepath="x/", overSampling=2, nbExpPoints=2, simulation_type="RayT",
             expID="t", poissonNoise=noise, seed=11)
    e = shim.Experiment(d)
    np.random.seed(5)
    e.myMembrane.getMyGeometry(e.exp_dict['studyDimensions'], e.myMembrane.membranePixelSize, 2, 0, 2)
    thresholds = list(e._open_bins(0))
    return e, e._scene(thresholds), e._get_engine()


@pytest.mark.parametrize("noise", [False, True])
def test_world1_matches_single_call(noise):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from paresis_b200 import shard
    e, scene, eng = _scene_and_engine(noise)
    for point in (0, 1):
        ref = eng.compute_rt(scene, point)
        got = shard.compute_rt_energy_sharded(eng, scene, point)
        for k in ("sample", "reference") + (("propag", "white") if point == 0 else ()):
            a, b = got[k].cpu().numpy(), ref[k].cpu().numpy()
            assert a.shape == b.shape == (2, 96, 128)
            if noise:
                # same Philox stream; the expectations agree to an ulp (fp32 REDs commit in any order), so
                # at most a handful of draws sitting exactly on a decision boundary may move
                assert np.mean(a != b) < 2e-3 and rel_l2(a, b) < 1e-3, k
            else:
                # (the single call sends the energies of a bin through paresis_refract_group, the sharded one
                # energy by energy through the lean hop kernel: two fixed-point roundings of the same deposits)
                assert rel_l2(a, b) < 3e-6, k
        assert np.allclose(got["mean_energy"], ref["mean_energy"], rtol=1e-7)     # fp32 partial sums of the reference beam: the miss list is drained in arrival order
        # the deferred form: same images, same means, once finish() has been called
        later = shard.compute_rt_energy_sharded(eng, scene, point, defer=True)
        assert "mean_energy" not in later._out
        fin = later.finish()
        assert fin is later.finish()
        for k in ("sample", "reference"):
            a, b = fin[k].cpu().numpy(), got[k].cpu().numpy()
            if noise:
                assert np.mean(a != b) < 2e-3 and rel_l2(a, b) < 1e-3, k
            else:
                assert rel_l2(a, b) < 3e-6, k          # fp32 REDs of the miss list commit in any order
        assert np.allclose(fin["mean_energy"], got["mean_energy"], rtol=1e-7)


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from paresis_b200 import shard
    e, scene, eng = _scene_and_engine(True)
    out = {}
    for point in (0, 1):
        res = shard.compute_rt_energy_sharded(eng, scene, point)
        later = shard.compute_rt_energy_sharded(eng, scene, point, defer=True).finish()      # every rank, same order
        if rank == 0:
            for k in ("sample", "reference"):
                assert float((later[k] != res[k]).float().mean()) < 2e-3, k
            assert np.allclose(later["mean_energy"], res["mean_energy"], rtol=1e-7)
        else:
            assert later is None
        if rank == 0:
            single = eng.compute_rt(scene, point)
            out[point] = ({k: res[k].cpu().numpy() for k in ("sample", "reference")},
                          {k: single[k].cpu().numpy() for k in ("sample", "reference")},
                          res["mean_energy"], single["mean_energy"])
        else:
            assert res is None
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_nccl_matches_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for point in (0, 1):
        sharded, single, me_s, me_1 = out[point]
        for k in sharded:
            # fp32 sums in a different order may move an expectation by an ulp and with it a few draws
            assert rel_l2(sharded[k], single[k]) < 2e-3, k
            assert abs(sharded[k].mean() / single[k].mean() - 1) < 1e-4
        assert np.allclose(me_s, me_1, rtol=1e-6)
