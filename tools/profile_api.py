"""cProfile of the e2e API path (one 20-position job) on a GPU box."""
import cProfile
import contextlib
import io
import os
import pstats
import sys
import tempfile


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paresis_b200 import workspace  # noqa: E402

ws = workspace.make_workspace(tempfile.mkdtemp())
workspace.enter(ws)
import Experiment as shim  # noqa: E402
import torch  # noqa: E402

d = dict(experimentName="B200_2048_mono", filepath="x/", overSampling=2, nbExpPoints=20, simulation_type="RayT", expID="p", seed=1)
with contextlib.redirect_stdout(io.StringIO()):
    exp = shim.Experiment(d)
mem = exp.myMembrane


def job():
    with contextlib.redirect_stdout(io.StringIO()):
        for point in range(20):
            mem.myGeometry = []
            mem.getMyGeometry(exp.exp_dict['studyDimensions'], mem.membranePixelSize, 2, point, 20)
            res = exp.computeSampleAndReferenceImages_RT(point)
            thick = mem.myGeometry[0]
    torch.cuda.synchronize()


job(); job(); job()
pr = cProfile.Profile()
pr.enable()
job()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
